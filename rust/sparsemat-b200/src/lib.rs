//! Safe wrappers that give the reference's hot path (sparsematrix.rs:146-158 `mvp`, vector.rs:50-63,
//! densevec.rs:51-73, linearsolver.rs:27-61) a device-resident implementation with the reference's names,
//! argument meaning and panics.  Source only: the build image has no Rust toolchain; every
//! call below is exercised through the same C ABI by tests/ (ctypes) and tests/cpp (C++ mirror).
use sparsemat::densevec::DenseVec;
use sparsemat::sparsemat_crs::SparseMatCRS;
use sparsemat::sparsemat_indexlist::SparseMatIndexList;
use sparsemat::sparsematrix::SparseMatrix;
use sparsemat::vector::Vector;
use sparsemat_b200_sys as sys;
use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::ptr;
use std::rc::Rc;

/// f32 / f64 (types.rs:70-77 FloatType) and u32 / u64 (types.rs:48-49) are the combinations the kernels carry.
pub trait GpuValue: Copy + Default + Into<f64> { const VT: i32; fn from_f64(v: f64) -> Self; }
impl GpuValue for f32 { const VT: i32 = sys::SMB200_F32; fn from_f64(v: f64) -> f32 { v as f32 } }
impl GpuValue for f64 { const VT: i32 = sys::SMB200_F64; fn from_f64(v: f64) -> f64 { v } }
pub trait GpuIndex: Copy { const IT: i32; }
impl GpuIndex for u32 { const IT: i32 = sys::SMB200_U32; }
impl GpuIndex for u64 { const IT: i32 = sys::SMB200_U64; }

fn check(status: sys::smb200_status) {
    match status {
        sys::SMB200_OK => (),
        sys::SMB200_ERR_DIM => panic!("Dimension mismatch"),                         // densevec.rs:52-54,61-63
        sys::SMB200_ERR_NOT_SQUARE => panic!("Matrix is not symmetric"),             // linearsolver.rs:30-32
        sys::SMB200_ERR_SIZE_MISMATCH => panic!("Matrix and vector size mismatch"),  // linearsolver.rs:33-36
        _ => panic!("smb200: {}", unsafe { CStr::from_ptr(sys::smb200_last_error()) }.to_string_lossy()),
    }
}

struct CtxInner(*mut sys::smb200_ctx);
impl Drop for CtxInner { fn drop(&mut self) { unsafe { sys::smb200_ctx_destroy(self.0); } } }
/// One CUDA device + stream; `!Send`/`!Sync` like the reference's `&mut` discipline (Rc inside).
#[derive(Clone)]
pub struct Context(Rc<CtxInner>);
impl Context {
    pub fn new(device: i32) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { sys::smb200_ctx_create(device, ptr::null_mut(), &mut h) });
        Context(Rc::new(CtxInner(h)))
    }
    pub fn sync(&self) { check(unsafe { sys::smb200_ctx_sync((self.0).0) }) }
}

/// densevec.rs:5-140 on the device.
pub struct DeviceVec<T: GpuValue> { h: *mut sys::smb200_vec, ctx: Context, _t: PhantomData<T> }
impl<T: GpuValue> Drop for DeviceVec<T> { fn drop(&mut self) { unsafe { sys::smb200_vec_free(self.h); } } }
impl<T: GpuValue> Clone for DeviceVec<T> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { sys::smb200_vec_clone(self.h, &mut h) });
        DeviceVec { h, ctx: self.ctx.clone(), _t: PhantomData }
    }
}
impl<T: GpuValue> DeviceVec<T> {
    pub fn with_dim(ctx: &Context, n: usize) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { sys::smb200_vec_create((ctx.0).0, T::VT, n as u64, &mut h) });
        DeviceVec { h, ctx: ctx.clone(), _t: PhantomData }
    }
    /// DenseVec::from_vec (densevec.rs:30-34); `values` is `dense.iter_ref().as_slice()`.
    pub fn from_slice(ctx: &Context, values: &[T]) -> Self {
        let v = Self::with_dim(ctx, values.len());
        check(unsafe { sys::smb200_vec_upload(v.h, values.as_ptr() as *const c_void, values.len() as u64) });
        ctx.sync();
        v
    }
    pub fn to_vec(&self) -> Vec<T> {
        let mut out = vec![T::default(); self.dim()];
        check(unsafe { sys::smb200_vec_download(self.h, out.as_mut_ptr() as *mut c_void, out.len() as u64) });
        out
    }
    pub fn dim(&self) -> usize { let mut n = 0u64; check(unsafe { sys::smb200_vec_dim(self.h, &mut n) }); n as usize }
    pub fn add(&mut self, rhs: &Self) { check(unsafe { sys::smb200_vec_add(self.h, rhs.h) }) }          // densevec.rs:51-58
    pub fn sub(&mut self, rhs: &Self) { check(unsafe { sys::smb200_vec_sub(self.h, rhs.h) }) }          // densevec.rs:60-67
    pub fn scale(&mut self, rhs: T) { check(unsafe { sys::smb200_vec_scale(self.h, rhs.into()) }) }     // densevec.rs:69-73
    pub fn inner_prod(&self, rhs: &Self) -> T {                                                          // vector.rs:50-53
        let mut out = 0.0;
        check(unsafe { sys::smb200_vec_dot(self.h, rhs.h, &mut out) });
        T::from_f64(out)
    }
    pub fn norm_squared(&self) -> T {                                                                    // vector.rs:56-58
        let mut out = 0.0;
        check(unsafe { sys::smb200_vec_norm2sq(self.h, &mut out) });
        T::from_f64(out)
    }
    pub fn norm(&self) -> f64 { let mut out = 0.0; check(unsafe { sys::smb200_vec_norm(self.h, &mut out) }); out }   // vector.rs:61-63
}

/// sparsemat_crs.rs:9-17 on the device; `mvp` is sparsematrix.rs:146-158.
pub struct GpuCrs<T: GpuValue, I: GpuIndex> { h: *mut sys::smb200_crs, ctx: Context, _t: PhantomData<(T, I)> }
impl<T: GpuValue, I: GpuIndex> Drop for GpuCrs<T, I> { fn drop(&mut self) { unsafe { sys::smb200_crs_free(self.h); } } }
impl<T: GpuValue, I: GpuIndex> GpuCrs<T, I> {
    /// Upload a finished CRS matrix.  `raw_parts()` is the accessor the overlay adds to sparsemat_crs.rs (its
    /// fields are private, sparsemat_crs.rs:9-17): `(values, columns, offset_rows)`.
    pub fn from_crs(ctx: &Context, m: &SparseMatCRS<T, I>) -> Self
    where SparseMatCRS<T, I>: for<'a> SparseMatrix<'a> {
        let (values, columns, offset_rows) = m.raw_parts();
        let mut h = ptr::null_mut();
        check(unsafe {
            sys::smb200_crs_upload((ctx.0).0, T::VT, I::IT, m.n_rows() as u64, m.n_cols() as u64, values.len() as u64,
                                   values.as_ptr() as *const c_void, columns.as_ptr() as *const c_void,
                                   offset_rows.as_ptr() as *const c_void, &mut h)
        });
        GpuCrs { h, ctx: ctx.clone(), _t: PhantomData }
    }
    /// SparseMatIndexList::to_crs (sparsemat_indexlist.rs:61-63) with the conversion done on the GPU;
    /// `raw_arrays()` is the overlay's reader for (columns, values, pos_start, index_list).
    pub fn from_indexlist(ctx: &Context, m: &SparseMatIndexList<T, I>) -> Self
    where SparseMatIndexList<T, I>: for<'a> SparseMatrix<'a> {
        let (columns, values, pos_start, index_list) = m.raw_arrays();
        let mut h = ptr::null_mut();
        check(unsafe {
            sys::smb200_crs_from_indexlist((ctx.0).0, T::VT, I::IT, m.n_rows() as u64, m.n_cols() as u64, values.len() as u64,
                                           columns.as_ptr() as *const c_void, values.as_ptr() as *const c_void,
                                           pos_start.as_ptr() as *const c_void, index_list.as_ptr() as *const c_void, &mut h)
        });
        GpuCrs { h, ctx: ctx.clone(), _t: PhantomData }
    }
    /// Binary CRS container (additive; the reference only writes text/PBM, sparsematrix.rs:304-338).
    pub fn save(&self, path: &std::path::Path) {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).expect("path contains NUL");
        check(unsafe { sys::smb200_crs_save(self.h, c.as_ptr()) });
    }
    /// Panics if the file holds another value/index type than `T`/`I`, or is damaged.
    pub fn load(ctx: &Context, path: &std::path::Path) -> Self {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes()).expect("path contains NUL");
        let (mut vt, mut it) = (0i32, 0i32);
        check(unsafe { sys::smb200_crsfile_info(c.as_ptr(), &mut vt, &mut it, ptr::null_mut()) });
        assert!(vt == T::VT && it == I::IT, "GpuCrs::load: the file holds another value/index type");
        let mut h = ptr::null_mut();
        check(unsafe { sys::smb200_crs_load((ctx.0).0, c.as_ptr(), &mut h) });
        GpuCrs { h, ctx: ctx.clone(), _t: PhantomData }
    }
    fn dims(&self) -> [u64; 3] { let mut d = [0u64; 3]; check(unsafe { sys::smb200_crs_dims(self.h, d.as_mut_ptr()) }); d }
    pub fn n_rows(&self) -> usize { self.dims()[0] as usize }
    pub fn n_cols(&self) -> usize { self.dims()[1] as usize }
    pub fn n_non_zero_entries(&self) -> usize { self.dims()[2] as usize }
    pub fn scale(&mut self, rhs: T) { check(unsafe { sys::smb200_crs_scale(self.h, rhs.into()) }) }     // sparsemat_crs.rs:153-157
    /// Device-resident product: a fresh vector of dim n_rows, like the reference.
    pub fn mvp(&self, rhs: &DeviceVec<T>) -> DeviceVec<T> {
        let y = DeviceVec::with_dim(&self.ctx, self.n_rows());
        check(unsafe { sys::smb200_spmv(self.h, rhs.h, y.h) });
        y
    }
    /// The reference's generic signature (any `Vector`): host values in, host values out — the call the
    /// `Mul<DenseVec<T>>` operator (sparsematrix.rs:435-443) maps to.
    pub fn mvp_host<'a, V: Vector<'a, Value = T>>(&self, rhs: &'a V) -> V {
        let x: Vec<T> = rhs.iter().collect();
        let mut y = vec![T::default(); self.n_rows()];
        check(unsafe { sys::smb200_spmv_host(self.h, x.as_ptr() as *const c_void, x.len() as u64, y.as_mut_ptr() as *mut c_void) });
        V::from_vec(y)
    }
    pub fn transpose(&self) -> Self {                                                                   // sparsematrix.rs:174-183
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::smb200_crs_transpose(self.h, &mut h) });
        GpuCrs { h, ctx: self.ctx.clone(), _t: PhantomData }
    }
    pub fn inner_prod(&self, lhs: &DeviceVec<T>, rhs: &DeviceVec<T>) -> T {                             // sparsematrix.rs:161-171
        let mut out = 0.0;
        check(unsafe { sys::smb200_bilinear(self.h, lhs.h, rhs.h, &mut out) });
        T::from_f64(out)
    }
}
impl<T: GpuValue, I: GpuIndex> std::ops::Mul<&DeviceVec<T>> for &GpuCrs<T, I> {
    type Output = DeviceVec<T>;
    fn mul(self, rhs: &DeviceVec<T>) -> DeviceVec<T> { self.mvp(rhs) }
}

/// linearsolver.rs:12-24.  `default()` keeps the reference's private constants; `new` is additive.
pub struct ConjugateGradient { tol: f64, iter_max: usize, relative: bool }
impl Default for ConjugateGradient { fn default() -> Self { ConjugateGradient { tol: 1e-12, iter_max: 10_000, relative: false } } }
impl ConjugateGradient {
    pub fn new(tol: f64, iter_max: usize, relative: bool) -> Self { ConjugateGradient { tol, iter_max, relative } }
    /// linearsolver.rs:27-61: x is updated in place; panics exactly where the reference does.
    pub fn solve<T: GpuValue, I: GpuIndex>(&self, mat: &GpuCrs<T, I>, b: &DeviceVec<T>, x: &mut DeviceVec<T>) {
        self.solve_with_stats(mat, b, x);
    }
    pub fn solve_with_stats<T: GpuValue, I: GpuIndex>(&self, mat: &GpuCrs<T, I>, b: &DeviceVec<T>, x: &mut DeviceVec<T>)
                                                      -> sys::smb200_cg_stats {
        let mut st = sys::smb200_cg_stats::default();
        check(unsafe { sys::smb200_cg_solve(mat.h, b.h, x.h, self.tol, self.relative as i32, self.iter_max as u64, &mut st) });
        st
    }
    /// Host-vector convenience with the reference's exact signature shape (b: &DenseVec, x: &mut DenseVec).
    pub fn solve_dense<T: GpuValue, I: GpuIndex>(&self, mat: &GpuCrs<T, I>, b: &DenseVec<T>, x: &mut DenseVec<T>)
    where DenseVec<T>: for<'a> Vector<'a, Value = T> {
        let db = DeviceVec::from_slice(&mat.ctx, b.iter_ref().as_slice());
        let mut dx = DeviceVec::from_slice(&mat.ctx, x.iter_ref().as_slice());
        self.solve(mat, &db, &mut dx);
        *x = DenseVec::from_vec(dx.to_vec());
    }
}
