//! Raw bindings of `include/smb200.h` (one declaration per entry point, same order as the header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub type smb200_status = i32;
pub const SMB200_OK: smb200_status = 0;
pub const SMB200_ERR_INVALID: smb200_status = 1;
pub const SMB200_ERR_DIM: smb200_status = 2;           // densevec.rs:52-54 "Dimension mismatch"
pub const SMB200_ERR_CUDA: smb200_status = 3;
pub const SMB200_ERR_NOT_SQUARE: smb200_status = 4;    // linearsolver.rs:30-32 "Matrix is not symmetric"
pub const SMB200_ERR_SIZE_MISMATCH: smb200_status = 5; // linearsolver.rs:33-36 "Matrix and vector size mismatch"
pub const SMB200_ERR_NCCL: smb200_status = 6;
pub const SMB200_ERR_UNSUPPORTED: smb200_status = 7;
pub const SMB200_ERR_OOM: smb200_status = 8;
pub const SMB200_ERR_IO: smb200_status = 9;            // crsfile_* / crs_save / crs_load

pub const SMB200_F32: i32 = 0;
pub const SMB200_F64: i32 = 1;
pub const SMB200_U32: i32 = 0;
pub const SMB200_U64: i32 = 1;

#[repr(C)] pub struct smb200_ctx { _p: [u8; 0] }
#[repr(C)] pub struct smb200_vec { _p: [u8; 0] }
#[repr(C)] pub struct smb200_crs { _p: [u8; 0] }
#[repr(C)] pub struct smb200_event { _p: [u8; 0] }
#[repr(C)] pub struct smb200_dist { _p: [u8; 0] }
#[repr(C)] pub struct smb200_par { _p: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct smb200_devinfo {
    pub device: i32,
    pub sm_count: i32,
    pub cc_major: i32,
    pub cc_minor: i32,
    pub l2_bytes: i64,
    pub l2_persist_max_bytes: i64,
    pub hbm_bytes: i64,
    pub name: [c_char; 128],
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct smb200_plan_info {
    pub variant: i32,
    pub lanes: i32,
    pub flags: u32,
    pub n_blocks: u64,
    pub n_rows: u64,
    pub n_cols: u64,
    pub nnz: u64,
    pub max_row_len: u64,
    pub mean_row_len: f64,
    pub algorithmic_bytes: u64,
    pub launches_per_spmv: u64,
    pub n_xwin_blocks: u64,
    pub nnz_c16: u64,
    pub stream_bytes: u64,
    pub rows_o16: u64,
    pub plan_bytes: u64,
    pub plan_ms: f64,
    pub nnz_v8: u64,
    pub sell_entries: u64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct smb200_cg_stats {
    pub iterations: u64,
    pub final_residual: f64,
    pub converged: i32,
    pub device_ms: f32,
    pub launches: u64,
}

extern "C" {
    pub fn smb200_version() -> i32;
    pub fn smb200_last_error() -> *const c_char;
    pub fn smb200_launch_count() -> u64;
    pub fn smb200_ctx_create(device: i32, stream: *mut c_void, out: *mut *mut smb200_ctx) -> smb200_status;
    pub fn smb200_ctx_destroy(ctx: *mut smb200_ctx) -> smb200_status;
    pub fn smb200_ctx_sync(ctx: *mut smb200_ctx) -> smb200_status;
    pub fn smb200_ctx_devinfo(ctx: *mut smb200_ctx, out: *mut smb200_devinfo) -> smb200_status;
    pub fn smb200_ctx_flush_l2(ctx: *mut smb200_ctx) -> smb200_status;
    pub fn smb200_event_create(ctx: *mut smb200_ctx, out: *mut *mut smb200_event) -> smb200_status;
    pub fn smb200_event_record(ev: *mut smb200_event) -> smb200_status;
    pub fn smb200_event_elapsed_ms(start: *mut smb200_event, stop: *mut smb200_event, ms: *mut f32) -> smb200_status;
    pub fn smb200_event_destroy(ev: *mut smb200_event) -> smb200_status;
    pub fn smb200_host_alloc(bytes: usize, out: *mut *mut c_void) -> smb200_status;
    pub fn smb200_host_free(p: *mut c_void) -> smb200_status;

    pub fn smb200_vec_create(ctx: *mut smb200_ctx, vt: i32, n: u64, out: *mut *mut smb200_vec) -> smb200_status;
    pub fn smb200_vec_wrap(ctx: *mut smb200_ctx, vt: i32, n: u64, device_ptr: *mut c_void, out: *mut *mut smb200_vec) -> smb200_status;
    pub fn smb200_vec_free(v: *mut smb200_vec) -> smb200_status;
    pub fn smb200_vec_device_ptr(v: *const smb200_vec, out: *mut *mut c_void) -> smb200_status;
    pub fn smb200_vec_copy(dst: *mut smb200_vec, src: *const smb200_vec) -> smb200_status;
    pub fn smb200_vec_fill(v: *mut smb200_vec, value: f64) -> smb200_status;
    pub fn smb200_vec_fill_uniform(v: *mut smb200_vec, seed: u64) -> smb200_status;
    pub fn smb200_vec_dim(v: *const smb200_vec, n: *mut u64) -> smb200_status;
    pub fn smb200_vec_upload(v: *mut smb200_vec, host: *const c_void, n: u64) -> smb200_status;
    pub fn smb200_vec_download(v: *const smb200_vec, host: *mut c_void, n: u64) -> smb200_status;
    pub fn smb200_vec_clone(v: *const smb200_vec, out: *mut *mut smb200_vec) -> smb200_status;
    pub fn smb200_vec_add(x: *mut smb200_vec, y: *const smb200_vec) -> smb200_status;
    pub fn smb200_vec_sub(x: *mut smb200_vec, y: *const smb200_vec) -> smb200_status;
    pub fn smb200_vec_scale(x: *mut smb200_vec, s: f64) -> smb200_status;
    pub fn smb200_vec_axpy(y: *mut smb200_vec, alpha: f64, x: *const smb200_vec) -> smb200_status;
    pub fn smb200_vec_scale_add(p: *mut smb200_vec, beta: f64, r: *const smb200_vec) -> smb200_status;
    pub fn smb200_vec_dot(x: *const smb200_vec, y: *const smb200_vec, out: *mut f64) -> smb200_status;
    pub fn smb200_vec_norm2sq(x: *const smb200_vec, out: *mut f64) -> smb200_status;
    pub fn smb200_vec_norm(x: *const smb200_vec, out: *mut f64) -> smb200_status;

    pub fn smb200_crs_upload(ctx: *mut smb200_ctx, vt: i32, it: i32, n_rows: u64, n_cols: u64, nnz: u64,
                             values: *const c_void, columns: *const c_void, offset_rows: *const c_void,
                             out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_crs_from_indexlist(ctx: *mut smb200_ctx, vt: i32, it: i32, n_rows: u64, n_cols: u64, nnz: u64,
                                     columns: *const c_void, values: *const c_void, pos_start: *const c_void,
                                     index_list: *const c_void, out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_crs_free(m: *mut smb200_crs) -> smb200_status;
    pub fn smb200_crsfile_write(path: *const c_char, vt: i32, it: i32, n_rows: u64, n_cols: u64, nnz: u64,
                                values: *const c_void, columns: *const c_void, offset_rows: *const c_void) -> smb200_status;
    pub fn smb200_crsfile_info(path: *const c_char, vt: *mut i32, it: *mut i32, out3: *mut u64) -> smb200_status;
    pub fn smb200_crsfile_read(path: *const c_char, values: *mut c_void, values_cap_bytes: u64, columns: *mut c_void,
                               columns_cap_bytes: u64, offset_rows: *mut c_void, offsets_cap_bytes: u64) -> smb200_status;
    pub fn smb200_crs_save(m: *const smb200_crs, path: *const c_char) -> smb200_status;
    pub fn smb200_crs_load(ctx: *mut smb200_ctx, path: *const c_char, out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_crs_dims(m: *const smb200_crs, out3: *mut u64) -> smb200_status;
    pub fn smb200_crs_download(m: *const smb200_crs, values: *mut c_void, columns: *mut c_void,
                               offset_rows: *mut c_void) -> smb200_status;
    pub fn smb200_crs_scale(m: *mut smb200_crs, s: f64) -> smb200_status;
    pub fn smb200_crs_types(m: *const smb200_crs, vt: *mut i32, it: *mut i32) -> smb200_status;
    pub fn smb200_crs_configure(m: *mut smb200_crs, variant: i32, lanes: i32, flags: u32) -> smb200_status;
    pub fn smb200_crs_plan_info(m: *const smb200_crs, out: *mut smb200_plan_info) -> smb200_status;
    pub fn smb200_gen_laplace(ctx: *mut smb200_ctx, vt: i32, it: i32, nx: u64, ny: u64, nz: u64, row_lo: u64, row_hi: u64,
                              out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_gen_powerlaw(ctx: *mut smb200_ctx, vt: i32, it: i32, n_rows: u64, n_cols: u64, seed_len: u64, seed_col: u64,
                               seed_val: u64, max_len: u64, out: *mut *mut smb200_crs) -> smb200_status;

    pub fn smb200_spmv(a: *mut smb200_crs, x: *const smb200_vec, y: *mut smb200_vec) -> smb200_status;
    pub fn smb200_spmv_host(a: *mut smb200_crs, x_host: *const c_void, nx: u64, y_host: *mut c_void) -> smb200_status;
    pub fn smb200_crs_transpose(a: *const smb200_crs, out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_bilinear(a: *mut smb200_crs, lhs: *const smb200_vec, rhs: *const smb200_vec, out: *mut f64) -> smb200_status;
    pub fn smb200_cg_solve(a: *mut smb200_crs, b: *const smb200_vec, x: *mut smb200_vec, tol: f64, relative: i32,
                           iter_max: u64, stats: *mut smb200_cg_stats) -> smb200_status;
    pub fn smb200_cg_history(a: *const smb200_crs, out: *mut f64, cap: u64, n: *mut u64) -> smb200_status;
    pub fn smb200_crs_diagonal(a: *const smb200_crs, d: *mut smb200_vec) -> smb200_status;
    pub fn smb200_pcg_jacobi_solve(a: *mut smb200_crs, b: *const smb200_vec, x: *mut smb200_vec, tol: f64, relative: i32,
                                   iter_max: u64, stats: *mut smb200_cg_stats) -> smb200_status;

    pub fn smb200_par_locate(n_blocks: u64, max_n_rows: u64, row: u64, block: *mut u64, local_row: *mut u64) -> smb200_status;
    pub fn smb200_par_create(ctx: *mut smb200_ctx, n_blocks: u64, max_n_rows: u64, vt: i32, it: i32, out: *mut *mut smb200_par) -> smb200_status;
    pub fn smb200_par_free(p: *mut smb200_par) -> smb200_status;
    pub fn smb200_par_owner(p: *const smb200_par, block: u64, rank: *mut i32) -> smb200_status;
    pub fn smb200_par_set_block_indexlist(p: *mut smb200_par, block: u64, n_rows: u64, n_cols: u64, nnz: u64, columns: *const c_void,
                                          values: *const c_void, pos_start: *const c_void, index_list: *const c_void) -> smb200_status;
    pub fn smb200_par_set_block_crs(p: *mut smb200_par, block: u64, n_rows: u64, n_cols: u64, nnz: u64, values: *const c_void,
                                    columns: *const c_void, offset_rows: *const c_void) -> smb200_status;
    pub fn smb200_par_dims(p: *const smb200_par, out3: *mut u64) -> smb200_status;
    pub fn smb200_par_block(p: *mut smb200_par, block: u64, out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_par_mvp(p: *mut smb200_par, x: *const smb200_vec, y: *mut smb200_vec) -> smb200_status;
    pub fn smb200_partition_rows(n_rows: u64, world: u32, align: u64, out_bounds: *mut u64) -> smb200_status;
    pub fn smb200_partition_rows_by_nnz(it: i32, n_rows: u64, offset_rows: *const c_void, world: u32, out_bounds: *mut u64) -> smb200_status;
    pub fn smb200_ghost_plan(it: i32, nnz: u64, columns_global: *const c_void, world: u32, rank: u32, bounds: *const u64,
                             columns_local_out: *mut c_void, ghosts: *mut u64, n_ghosts: *mut u64,
                             ghosts_per_owner: *mut u64) -> smb200_status;
    pub fn smb200_comm_unique_id(out128: *mut c_void) -> smb200_status;
    pub fn smb200_comm_init(ctx: *mut smb200_ctx, rank: i32, world: i32, uid128: *const c_void) -> smb200_status;
    pub fn smb200_dist_create(ctx: *mut smb200_ctx, vt: i32, it: i32, n_global: u64, bounds: *const u64, nnz_local: u64,
                              values: *const c_void, columns_global: *const c_void, offset_rows_local: *const c_void,
                              out: *mut *mut smb200_dist) -> smb200_status;
    pub fn smb200_comm_destroy(ctx: *mut smb200_ctx) -> smb200_status;
    pub fn smb200_dist_laplace(ctx: *mut smb200_ctx, vt: i32, it: i32, nx: u64, ny: u64, nz: u64, out: *mut *mut smb200_dist) -> smb200_status;
    pub fn smb200_dist_dims(d: *const smb200_dist, out4: *mut u64) -> smb200_status;
    pub fn smb200_dist_local(d: *mut smb200_dist, out: *mut *mut smb200_crs) -> smb200_status;
    pub fn smb200_dist_free(d: *mut smb200_dist) -> smb200_status;
    pub fn smb200_dist_vec_create(d: *mut smb200_dist, out: *mut *mut smb200_vec) -> smb200_status;
    pub fn smb200_dist_spmv(d: *mut smb200_dist, x: *mut smb200_vec, y: *mut smb200_vec) -> smb200_status;
    pub fn smb200_dist_dot(d: *mut smb200_dist, x: *const smb200_vec, y: *const smb200_vec, out: *mut f64) -> smb200_status;
    pub fn smb200_dist_barrier(d: *mut smb200_dist) -> smb200_status;
    pub fn smb200_dist_info(d: *mut smb200_dist, out6: *mut u64) -> smb200_status;
    pub fn smb200_dist_cg_solve(d: *mut smb200_dist, b: *const smb200_vec, x: *mut smb200_vec, tol: f64, relative: i32,
                                iter_max: u64, stats: *mut smb200_cg_stats) -> smb200_status;
    pub fn smb200_dist_cg_solve_sr(d: *mut smb200_dist, b: *const smb200_vec, x: *mut smb200_vec, tol: f64, relative: i32,
                                   iter_max: u64, stats: *mut smb200_cg_stats) -> smb200_status;
}
