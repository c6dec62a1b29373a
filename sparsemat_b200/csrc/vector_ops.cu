// DenseVec<T> on the device: storage management and the K5 (dot / norm) and K6 (elementwise)
// kernels.  Reference: densevec.rs:5-140, vector.rs:50-63, compositions at linearsolver.rs:47-59.
//
// Roofline: all of these are pure HBM streams.  Algorithmic bytes per element: add/sub/axpy/
// scale_add 3*sizeof(T) (2 reads + 1 write), scale 2*sizeof(T), dot 2*sizeof(T), norm 1*sizeof(T).
// Every thread moves 16 bytes per access (float4 / double2), grids are sized to a multiple of the SM
// count, and the elementwise arithmetic uses the never-contracted intrinsics so that results are
// bit-identical to the reference's separate multiply and add.
#include "common.cuh"
#include "reduce.cuh"
#include "rng.cuh"

namespace smb {

constexpr int kEwThreads = 256;
constexpr int kRedThreads = 256;

enum class Ew { Add, Sub, Scale, Axpy, ScaleAdd, Fill, Uniform };

template <class T, Ew OP>
__device__ __forceinline__ T ew_apply(T a, T b, T s) {
    if constexpr (OP == Ew::Add) return add_rn(a, b);
    else if constexpr (OP == Ew::Sub) return sub_rn(a, b);
    else if constexpr (OP == Ew::Scale) return mul_rn(a, s);
    else if constexpr (OP == Ew::Axpy) return add_rn(a, mul_rn(b, s));       // a += (b * s)
    else if constexpr (OP == Ew::ScaleAdd) return add_rn(mul_rn(a, s), b);   // a = (a * s) + b
    else return s;
}

// a[i] = op(a[i], b[i], s) for i < n.  VEC = true requires 16-byte aligned a and b.
template <class T, Ew OP, bool VEC>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(T* __restrict__ a, const T* __restrict__ b, uint64_t n, T s) {
    constexpr bool kReadsA = (OP != Ew::Fill);
    constexpr bool kReadsB = (OP == Ew::Add || OP == Ew::Sub || OP == Ew::Axpy || OP == Ew::ScaleAdd);
    const uint64_t tid = blockIdx.x * (uint64_t)kEwThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kEwThreads;
    if constexpr (VEC) {
        using V = typename Vec16<T>::type;
        constexpr int N = Vec16<T>::N;
        const uint64_t nvec = n / N;
        V* av = reinterpret_cast<V*>(a);
        const V* bv = reinterpret_cast<const V*>(b);
        for (uint64_t i = tid; i < nvec; i += stride) {
            Pack16<T> pa, pb;
            if constexpr (kReadsA) pa.v = av[i];
            if constexpr (kReadsB) pb.v = __ldg(bv + i);
#pragma unroll
            for (int k = 0; k < N; ++k) pa.e[k] = ew_apply<T, OP>(kReadsA ? pa.e[k] : T(0), kReadsB ? pb.e[k] : T(0), s);
            av[i] = pa.v;
        }
        for (uint64_t i = nvec * N + tid; i < n; i += stride)
            a[i] = ew_apply<T, OP>(kReadsA ? a[i] : T(0), kReadsB ? b[i] : T(0), s);
    } else {
        for (uint64_t i = tid; i < n; i += stride)
            a[i] = ew_apply<T, OP>(kReadsA ? a[i] : T(0), kReadsB ? b[i] : T(0), s);
    }
}

template <class T>
__global__ void __launch_bounds__(kEwThreads) uniform_kernel(T* __restrict__ a, uint64_t n, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * kEwThreads;
    for (uint64_t i = blockIdx.x * (uint64_t)kEwThreads + threadIdx.x; i < n; i += stride)
        a[i] = (T)rng::pm1(rng::rng1(seed, i));
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static unsigned ew_grid(const smb200_ctx* ctx, uint64_t n_items) {
    uint64_t need = (n_items + kEwThreads - 1) / kEwThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * 8;           // 8 CTAs of 256 threads fill an SM
    if (need < 1) need = 1;
    return (unsigned)(need < cap ? need : cap);
}

template <class T, Ew OP>
static smb200_status ew_launch_t(smb200_ctx* ctx, void* a, const void* b, uint64_t n, double s) {
    if (n == 0) return SMB200_OK;
    const bool vec = aligned16(a) && (b == nullptr || aligned16(b));
    if (vec) {
        unsigned g = ew_grid(ctx, n / Vec16<T>::N + 1);
        ew_kernel<T, OP, true><<<g, kEwThreads, 0, ctx->stream>>>((T*)a, (const T*)b, n, (T)s);
    } else {
        unsigned g = ew_grid(ctx, n);
        ew_kernel<T, OP, false><<<g, kEwThreads, 0, ctx->stream>>>((T*)a, (const T*)b, n, (T)s);
    }
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

template <Ew OP>
static smb200_status ew_launch(smb200_ctx* ctx, int vt, void* a, const void* b, uint64_t n, double s) {
    return vt == SMB200_F64 ? ew_launch_t<double, OP>(ctx, a, b, n, s) : ew_launch_t<float, OP>(ctx, a, b, n, s);
}

// ---- K5: dot / norm^2 -------------------------------------------------------------------------------
// Each thread folds the products it owns in T (a few dozen terms), the partials are combined in f64
// (warp shuffle -> CTA -> grid, fixed order) and the grand total is rounded to T once.
template <class T, bool VEC>
__global__ void __launch_bounds__(kRedThreads)
dot_kernel(const T* __restrict__ x, const T* __restrict__ y, uint64_t n, double* __restrict__ partials,
           unsigned int* __restrict__ ticket, double* __restrict__ result) {
    __shared__ double scratch[kRedThreads / 32 + 1];
    const uint64_t tid = blockIdx.x * (uint64_t)kRedThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kRedThreads;
    double acc = 0.0;
    if constexpr (VEC) {
        using V = typename Vec16<T>::type;
        constexpr int N = Vec16<T>::N;
        const uint64_t nvec = n / N;
        const V* xv = reinterpret_cast<const V*>(x);
        const V* yv = reinterpret_cast<const V*>(y);
        T lane_acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) lane_acc[k] = T(0);
        for (uint64_t i = tid; i < nvec; i += stride) {
            Pack16<T> px, py;
            px.v = __ldg(xv + i);
            py.v = (x == y) ? px.v : __ldg(yv + i);
#pragma unroll
            for (int k = 0; k < N; ++k) lane_acc[k] = add_rn(lane_acc[k], mul_rn(px.e[k], py.e[k]));
        }
#pragma unroll
        for (int k = 0; k < N; ++k) acc += (double)lane_acc[k];
        for (uint64_t i = nvec * N + tid; i < n; i += stride) acc += (double)mul_rn(x[i], y[i]);
    } else {
        T a = T(0);
        for (uint64_t i = tid; i < n; i += stride) a = add_rn(a, mul_rn(x[i], y[i]));
        acc = (double)a;
    }
    const double bsum = block_sum<kRedThreads>(acc, scratch);
    double total;
    if (grid_sum<kRedThreads>(bsum, partials, ticket, scratch, total)) {
        if (threadIdx.x == 0) *result = (double)(T)total;
    }
}

smb200_status dot_launch(smb200_ctx* ctx, int vt, const void* x, const void* y, uint64_t n, int slot) {
    const bool vec = aligned16(x) && aligned16(y);
    const size_t esz = vsize(vt);
    uint64_t items = vec ? n / (16 / esz) + 1 : n;
    uint64_t need = (items + kRedThreads - 1) / kRedThreads;
    uint64_t cap = (uint64_t)ctx->sm_count * 8;
    unsigned g = (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
    SMB_TRY(ensure_reduction_scratch(ctx, g));
    double* res = ctx->red_result + slot;
    if (vt == SMB200_F64) {
        if (vec) dot_kernel<double, true><<<g, kRedThreads, 0, ctx->stream>>>((const double*)x, (const double*)y, n, ctx->red_partials, ctx->red_ticket, res);
        else dot_kernel<double, false><<<g, kRedThreads, 0, ctx->stream>>>((const double*)x, (const double*)y, n, ctx->red_partials, ctx->red_ticket, res);
    } else {
        if (vec) dot_kernel<float, true><<<g, kRedThreads, 0, ctx->stream>>>((const float*)x, (const float*)y, n, ctx->red_partials, ctx->red_ticket, res);
        else dot_kernel<float, false><<<g, kRedThreads, 0, ctx->stream>>>((const float*)x, (const float*)y, n, ctx->red_partials, ctx->red_ticket, res);
    }
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

}  // namespace smb

using namespace smb;

static smb200_status vec_make(smb200_ctx* ctx, smb200_vtype vt, uint64_t n, uint64_t cap, smb200_vec** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "vec_create: NULL argument");
    SMB_REQUIRE(vt == SMB200_F32 || vt == SMB200_F64, SMB200_ERR_INVALID, "vec_create: bad value type %d", (int)vt);
    *out = nullptr;
    SMB_CUDA(cudaSetDevice(ctx->device));
    smb200_vec* v = new smb200_vec();
    v->ctx = ctx; v->vt = vt; v->n = n; v->cap = cap < n ? n : cap;
    smb200_status s = dev_alloc(&v->d, v->cap * vsize(vt));
    if (s != SMB200_OK) { delete v; return s; }
    cudaError_t e = cudaMemsetAsync(v->d, 0, v->cap * vsize(vt) + kPadBytes, ctx->stream);
    if (e != cudaSuccess) { cudaFree(v->d); delete v; SMB_CUDA(e); }
    ctx_retain(ctx);
    *out = v;
    return SMB200_OK;
}

namespace smb {
smb200_status vec_create_cap(smb200_ctx* ctx, int vt, uint64_t n, uint64_t cap, smb200_vec** out) {
    return vec_make(ctx, (smb200_vtype)vt, n, cap, out);
}
}  // namespace smb

extern "C" {

smb200_status smb200_vec_create(smb200_ctx* ctx, smb200_vtype vt, uint64_t n, smb200_vec** out) {
    return vec_make(ctx, vt, n, n, out);
}

smb200_status smb200_vec_wrap(smb200_ctx* ctx, smb200_vtype vt, uint64_t n, void* device_ptr, smb200_vec** out) {
    SMB_REQUIRE(ctx && out && (device_ptr || n == 0), SMB200_ERR_INVALID, "vec_wrap: NULL argument");
    SMB_REQUIRE(vt == SMB200_F32 || vt == SMB200_F64, SMB200_ERR_INVALID, "vec_wrap: bad value type %d", (int)vt);
    SMB_REQUIRE(((uintptr_t)device_ptr % vsize(vt)) == 0, SMB200_ERR_INVALID, "vec_wrap: pointer not element aligned");
    smb200_vec* v = new smb200_vec();
    v->ctx = ctx; v->vt = vt; v->n = n; v->cap = n; v->d = device_ptr; v->owned = false;
    ctx_retain(ctx);
    *out = v;
    return SMB200_OK;
}

smb200_status smb200_vec_free(smb200_vec* v) {
    if (!v) return SMB200_OK;
    if (v->owned && v->d) {
        cudaSetDevice(v->ctx->device);
        cudaStreamSynchronize(v->ctx->stream);
        cudaFree(v->d);
    }
    smb200_ctx* ctx = v->ctx;
    delete v;
    ctx_release(ctx);
    return SMB200_OK;
}

smb200_status smb200_vec_dim(const smb200_vec* v, uint64_t* n) {
    SMB_REQUIRE(v && n, SMB200_ERR_INVALID, "vec_dim: NULL argument");
    *n = v->n;
    return SMB200_OK;
}

smb200_status smb200_vec_device_ptr(const smb200_vec* v, void** out) {
    SMB_REQUIRE(v && out, SMB200_ERR_INVALID, "vec_device_ptr: NULL argument");
    *out = v->d;
    return SMB200_OK;
}

smb200_status smb200_vec_upload(smb200_vec* v, const void* host, uint64_t n) {
    SMB_REQUIRE(v && (host || n == 0), SMB200_ERR_INVALID, "vec_upload: NULL argument");
    SMB_REQUIRE(n <= v->n, SMB200_ERR_DIM, "vec_upload: %llu elements into a vector of dim %llu",
                (unsigned long long)n, (unsigned long long)v->n);
    if (n) SMB_CUDA(cudaMemcpyAsync(v->d, host, n * vsize(v->vt), cudaMemcpyHostToDevice, v->ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_vec_download(const smb200_vec* v, void* host, uint64_t n) {
    SMB_REQUIRE(v && (host || n == 0), SMB200_ERR_INVALID, "vec_download: NULL argument");
    SMB_REQUIRE(n <= v->n, SMB200_ERR_DIM, "vec_download: %llu elements from a vector of dim %llu",
                (unsigned long long)n, (unsigned long long)v->n);
    if (n) SMB_CUDA(cudaMemcpyAsync(host, v->d, n * vsize(v->vt), cudaMemcpyDeviceToHost, v->ctx->stream));
    SMB_CUDA(cudaStreamSynchronize(v->ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_vec_clone(const smb200_vec* v, smb200_vec** out) {
    SMB_REQUIRE(v && out, SMB200_ERR_INVALID, "vec_clone: NULL argument");
    SMB_TRY(vec_make(v->ctx, (smb200_vtype)v->vt, v->n, v->cap, out));
    if (v->n) SMB_CUDA(cudaMemcpyAsync((*out)->d, v->d, v->n * vsize(v->vt), cudaMemcpyDeviceToDevice, v->ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_vec_copy(smb200_vec* dst, const smb200_vec* src) {
    SMB_REQUIRE(dst && src, SMB200_ERR_INVALID, "vec_copy: NULL argument");
    SMB_REQUIRE(dst->vt == src->vt, SMB200_ERR_INVALID, "vec_copy: value types differ");
    SMB_REQUIRE(dst->n >= src->n, SMB200_ERR_DIM, "Dimension mismatch");
    if (src->n) SMB_CUDA(cudaMemcpyAsync(dst->d, src->d, src->n * vsize(src->vt), cudaMemcpyDeviceToDevice, dst->ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_vec_fill(smb200_vec* v, double value) {
    SMB_REQUIRE(v, SMB200_ERR_INVALID, "vec_fill: NULL argument");
    return ew_launch<Ew::Fill>(v->ctx, v->vt, v->d, nullptr, v->n, value);
}

smb200_status smb200_vec_fill_uniform(smb200_vec* v, uint64_t seed) {
    SMB_REQUIRE(v, SMB200_ERR_INVALID, "vec_fill_uniform: NULL argument");
    if (v->n == 0) return SMB200_OK;
    unsigned g = ew_grid(v->ctx, v->n);
    if (v->vt == SMB200_F64) uniform_kernel<double><<<g, kEwThreads, 0, v->ctx->stream>>>((double*)v->d, v->n, seed);
    else uniform_kernel<float><<<g, kEwThreads, 0, v->ctx->stream>>>((float*)v->d, v->n, seed);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

#define SMB_BINARY_CHECK(name)                                                                      \
    SMB_REQUIRE(x && y, SMB200_ERR_INVALID, name ": NULL argument");                                \
    SMB_REQUIRE(x->vt == y->vt, SMB200_ERR_INVALID, name ": value types differ");                   \
    SMB_REQUIRE(x->ctx == y->ctx, SMB200_ERR_INVALID, name ": vectors belong to different contexts")

smb200_status smb200_vec_add(smb200_vec* x, const smb200_vec* y) {
    SMB_BINARY_CHECK("vec_add");
    SMB_REQUIRE(x->n >= y->n, SMB200_ERR_DIM, "Dimension mismatch");      // densevec.rs:52-54
    return ew_launch<Ew::Add>(x->ctx, x->vt, x->d, y->d, y->n, 0.0);
}
smb200_status smb200_vec_sub(smb200_vec* x, const smb200_vec* y) {
    SMB_BINARY_CHECK("vec_sub");
    SMB_REQUIRE(x->n >= y->n, SMB200_ERR_DIM, "Dimension mismatch");      // densevec.rs:61-63
    return ew_launch<Ew::Sub>(x->ctx, x->vt, x->d, y->d, y->n, 0.0);
}
smb200_status smb200_vec_scale(smb200_vec* x, double s) {
    SMB_REQUIRE(x, SMB200_ERR_INVALID, "vec_scale: NULL argument");
    return ew_launch<Ew::Scale>(x->ctx, x->vt, x->d, nullptr, x->n, s);
}
smb200_status smb200_vec_axpy(smb200_vec* y, double alpha, const smb200_vec* x) {
    SMB_BINARY_CHECK("vec_axpy");
    SMB_REQUIRE(y->n >= x->n, SMB200_ERR_DIM, "Dimension mismatch");
    return ew_launch<Ew::Axpy>(y->ctx, y->vt, y->d, x->d, x->n, alpha);
}
smb200_status smb200_vec_scale_add(smb200_vec* x, double beta, const smb200_vec* y) {
    SMB_BINARY_CHECK("vec_scale_add");
    SMB_REQUIRE(x->n == y->n, SMB200_ERR_DIM, "Dimension mismatch");
    return ew_launch<Ew::ScaleAdd>(x->ctx, x->vt, x->d, y->d, x->n, beta);
}

smb200_status smb200_vec_dot(const smb200_vec* x, const smb200_vec* y, double* out) {
    SMB_REQUIRE(x && y && out, SMB200_ERR_INVALID, "vec_dot: NULL argument");
    SMB_REQUIRE(x->vt == y->vt, SMB200_ERR_INVALID, "vec_dot: value types differ");
    SMB_REQUIRE(x->ctx == y->ctx, SMB200_ERR_INVALID, "vec_dot: vectors belong to different contexts");
    const uint64_t n = x->n < y->n ? x->n : y->n;                          // zip stops at the shorter
    SMB_TRY(dot_launch(x->ctx, x->vt, x->d, y->d, n, 0));
    return fetch_result(x->ctx, 0, out);
}
smb200_status smb200_vec_norm2sq(const smb200_vec* x, double* out) {
    SMB_REQUIRE(x && out, SMB200_ERR_INVALID, "vec_norm2sq: NULL argument");
    SMB_TRY(dot_launch(x->ctx, x->vt, x->d, x->d, x->n, 0));
    return fetch_result(x->ctx, 0, out);
}
smb200_status smb200_vec_norm(const smb200_vec* x, double* out) {
    double sq = 0.0;
    SMB_TRY(smb200_vec_norm2sq(x, &sq));
    *out = sqrt(sq);                                                       // f64::sqrt(f64::from(T))
    return SMB200_OK;
}

}  // extern "C"
