// K7 — ConjugateGradient::solve (linearsolver.rs:27-61) as three fused kernels per iteration with
// every scalar (alpha, beta, r.r, the stop test) kept on the device:
//
//   A  ap = A p,  pAp = p.ap                    SpMV with the dot fused into its epilogue (spmv.cu)
//   B  r -= (ap*alpha); rr' = r.r                alpha = rr / pAp computed by every thread
//   C  x += (p*alpha); stop test sqrt(rr') < threshold; p = (p*beta) + r   beta = rr' / rr
//
// Elementwise arithmetic is the reference's: product rounded, then add (two roundings, never an FMA),
// alpha/beta are divisions in T.  Only the two reductions are re-ordered (fixed-order tree in f64).  The x update of
// linearsolver.rs:45-46 rides kernel C instead of B (same operation on the same operands, p is read once per iteration
// instead of twice).  Algorithmic bytes per iteration: SpMV bytes + 8*N*sizeof(T)  (B: r,ap in, r out; C: x,p,r in, x,p out)
// against ~24*N*sizeof(T) plus 3-4 allocations in the reference's clone-heavy loop.
//
// Iterations are replayed from a CUDA graph in batches; the host only polls a pinned copy of the
// (iteration, done) pair one batch behind the GPU, so the device never waits for the CPU.  After the
// stop test fires (in C, which has updated x first — the reference updates x before its `break`, linearsolver.rs:45-54)
// the remaining kernels of a batch exit immediately, leaving x, r, p untouched.  Two flags: S_DONE is raised by one
// thread of C while other CTAs of the same launch may still be starting, so C itself looks at S_DONE_SEEN, which the
// next B raises when it finds S_DONE set.
#include "common.cuh"
#include "cg_sr.cuh"
#include "halo.cuh"
#include "reduce.cuh"

#include <cmath>
#include <cstdlib>

namespace smb {

// device scalar block (doubles)
// S_PAP..S_PAP+2: up to three partial p.Ap sums (interior / lower / upper boundary launches of the
// distributed SpMV); contiguous so one all-reduce covers them.  Single GPU uses slot 0 only.
// S_RR_LOCAL: multi-GPU only — the rank-local r.r, all-reduced OUT OF PLACE into S_RR_NEW (after the stop test
// has fired the update kernels exit early and leave S_RR_LOCAL alone, so repeating the all-reduce is harmless).
// S_RES2: Jacobi-preconditioned solves only — r.r for the stop test (S_RR / S_RR_NEW then hold r.z, z = D^-1 r).
enum { S_RR = 0, S_PAP = 1, S_RR_NEW = 4, S_THRESH = 5, S_ITER = 6, S_DONE = 7, S_RR_LOCAL = 8, S_RES2 = 9, S_DONE_SEEN = 13, S_COUNT = 16 };   // 10..12: cg_sr.cuh

constexpr int kCgThreads = 256;

// Two grid-wide sums with one ticket (same pattern as reduce.cuh's grid_sum; partials holds 2 * nblocks doubles).
template <int THREADS>
__device__ __forceinline__ bool grid_sum2(double v0, double v1, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                          double* scratch, double& t0, double& t1) {
    __shared__ bool s_last2;
    const unsigned int nblocks = gridDim.x;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v0;
        partials[nblocks + blockIdx.x] = v1;
        __threadfence();
        s_last2 = atomicAdd(ticket, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (!s_last2) return false;
    __threadfence();
    double a0 = 0.0, a1 = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += THREADS) { a0 += __ldcg(partials + i); a1 += __ldcg(partials + nblocks + i); }
    t0 = block_sum<THREADS>(a0, scratch);
    t1 = block_sum<THREADS>(a1, scratch);
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

// r = b - ap; p = r; S[RR_NEW] = r.r   (linearsolver.rs:38-40).  PRE (Jacobi, additive): z = dinv * r; p = z; S[RR_NEW] = r.z
template <class T, bool PRE>
__global__ void __launch_bounds__(kCgThreads)
cg_init_kernel(const T* __restrict__ b, const T* __restrict__ ap, T* __restrict__ r, T* __restrict__ p, uint64_t n,
               double* __restrict__ S, int rr_slot, double* __restrict__ partials, unsigned int* __restrict__ ticket, int vec_ok,
               const T* __restrict__ dinv) {
    __shared__ double scratch[kCgThreads / 32 + 1];
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const uint64_t tid = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kCgThreads;
    const uint64_t nvec = vec_ok ? n / N : 0;      // a borrowed b / x (smb200_vec_wrap) may not be 16-byte aligned: element loop
    T lane_acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) lane_acc[k] = T(0);
    for (uint64_t i = tid; i < nvec; i += stride) {
        Pack16<T> pb, pa, pz;
        pb.v = __ldg(reinterpret_cast<const V*>(b) + i);
        pa.v = __ldg(reinterpret_cast<const V*>(ap) + i);
        if constexpr (PRE) pz.v = __ldg(reinterpret_cast<const V*>(dinv) + i);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            pb.e[k] = sub_rn(pb.e[k], pa.e[k]);
            if constexpr (PRE) pz.e[k] = mul_rn(pz.e[k], pb.e[k]); else pz.e[k] = pb.e[k];
            lane_acc[k] = add_rn(lane_acc[k], mul_rn(pb.e[k], pz.e[k]));
        }
        reinterpret_cast<V*>(r)[i] = pb.v;
        reinterpret_cast<V*>(p)[i] = pz.v;
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) acc += (double)lane_acc[k];
    for (uint64_t i = nvec * N + tid; i < n; i += stride) {
        const T v = sub_rn(b[i], ap[i]);
        T z = v;
        if constexpr (PRE) z = mul_rn(dinv[i], v);
        r[i] = v; p[i] = z;
        acc += (double)mul_rn(v, z);
    }
    const double bsum = block_sum<kCgThreads>(acc, scratch);
    double total;
    if (grid_sum<kCgThreads>(bsum, partials, ticket, scratch, total))
        if (threadIdx.x == 0) S[rr_slot] = (double)(T)total;
}

// B: r -= (ap * alpha); S[RR_NEW] = r.r      (linearsolver.rs:47-51)
//    PRE: S[RR_NEW] = r.(dinv * r) and S[RES2] = r.r
template <class T, bool PRE>
__global__ void __launch_bounds__(kCgThreads)
cg_update_r_kernel(T* __restrict__ r, const T* __restrict__ ap, uint64_t n,
                   double* __restrict__ S, int rr_slot, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                   const T* __restrict__ dinv, const ArDev* __restrict__ ar) {
    __shared__ double scratch[kCgThreads / 32 + 1];
    __shared__ double ar_sv[kMaxPeers][kArSlots], ar_in[kArSlots], ar_out[kArSlots];
    if (__ldcg(S + S_DONE) != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) S[S_DONE_SEEN] = 1.0;
        return;
    }
    const T alpha = div_rn((T)__ldcg(S + S_RR), (T)(__ldcg(S + S_PAP) + __ldcg(S + S_PAP + 1) + __ldcg(S + S_PAP + 2)));
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const uint64_t tid = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kCgThreads;
    const uint64_t nvec = n / N;
    T lane_acc[N], lane_acz[N];
#pragma unroll
    for (int k = 0; k < N; ++k) lane_acc[k] = lane_acz[k] = T(0);
    for (uint64_t i = tid; i < nvec; i += stride) {
        Pack16<T> pr, pa, pd;
        pa.v = __ldg(reinterpret_cast<const V*>(ap) + i);
        if constexpr (PRE) pd.v = __ldg(reinterpret_cast<const V*>(dinv) + i);
        pr.v = reinterpret_cast<V*>(r)[i];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            pr.e[k] = sub_rn(pr.e[k], mul_rn(pa.e[k], alpha));
            lane_acc[k] = add_rn(lane_acc[k], mul_rn(pr.e[k], pr.e[k]));
            if constexpr (PRE) lane_acz[k] = add_rn(lane_acz[k], mul_rn(pr.e[k], mul_rn(pd.e[k], pr.e[k])));
        }
        reinterpret_cast<V*>(r)[i] = pr.v;
    }
    double acc = 0.0, acz = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) { acc += (double)lane_acc[k]; acz += (double)lane_acz[k]; }
    for (uint64_t i = nvec * N + tid; i < n; i += stride) {
        const T rv = sub_rn(r[i], mul_rn(ap[i], alpha));
        r[i] = rv;
        acc += (double)mul_rn(rv, rv);
        if constexpr (PRE) acz += (double)mul_rn(rv, mul_rn(dinv[i], rv));
    }
    if constexpr (PRE) {
        const double b0 = block_sum<kCgThreads>(acc, scratch), b1 = block_sum<kCgThreads>(acz, scratch);
        double t0, t1;
        if (grid_sum2<kCgThreads>(b0, b1, partials, ticket, scratch, t0, t1))
            if (threadIdx.x == 0) { S[S_RES2] = (double)(T)t0; S[rr_slot] = (double)(T)t1; }
    } else {
        const double bsum = block_sum<kCgThreads>(acc, scratch);
        double total;
        if (grid_sum<kCgThreads>(bsum, partials, ticket, scratch, total)) {
            if (ar != nullptr) {
                // one rank per GPU: the CTA that finished last exchanges the ranks' r.r through peer memory (halo.cuh) and
                // leaves the global value, rounded to T like the reference's scalar, where kernel C expects it
                if (threadIdx.x == 0) ar_in[0] = total;
                __syncthreads();
                if (threadIdx.x < 32) ar_warp_allreduce(*ar, ar_in, ar_out, 1, ar_sv);
                __syncthreads();
                if (threadIdx.x == 0) S[S_RR_NEW] = (double)(T)ar_out[0];
            } else if (threadIdx.x == 0) {
                S[rr_slot] = (double)(T)total;
            }
        }
    }
}

// C: x += (p * alpha); stop test, bookkeeping, p = (p * beta) + r          (linearsolver.rs:45-46, 52-59)
//    PRE: the stop test looks at r.r (S[RES2]), beta = r.z' / r.z, p = (p * beta) + dinv * r
template <class T, bool PRE>
__global__ void __launch_bounds__(kCgThreads)
cg_update_xp_kernel(T* __restrict__ x, T* __restrict__ p, const T* __restrict__ r, uint64_t n, double* __restrict__ S,
                    double* __restrict__ history, uint64_t hist_cap, int vec_ok, const T* __restrict__ dinv) {
    if (__ldcg(S + S_DONE_SEEN) != 0.0) return;
    const double rr_old = __ldcg(S + S_RR);
    const T alpha = div_rn((T)rr_old, (T)(__ldcg(S + S_PAP) + __ldcg(S + S_PAP + 1) + __ldcg(S + S_PAP + 2)));    // as in B
    const double rr_new = __ldcg(S + S_RR_NEW);
    const double res = sqrt(PRE ? __ldcg(S + S_RES2) : rr_new);     // f64::sqrt(r_norm_squared.into())
    const bool done = res < __ldcg(S + S_THRESH);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t it = (uint64_t)S[S_ITER];
        if (history && it < hist_cap) history[it] = res;
        S[S_ITER] = (double)(it + 1);
        if (done) S[S_DONE] = 1.0;
    }
    const T beta = done ? T(0) : div_rn((T)rr_new, (T)rr_old);
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const uint64_t tid = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kCgThreads;
    const uint64_t nvec = vec_ok ? n / N : 0;      // a borrowed x (smb200_vec_wrap) may not be 16-byte aligned: element loop
    for (uint64_t i = tid; i < nvec; i += stride) {
        Pack16<T> px, pp, pr, pd;
        pp.v = reinterpret_cast<V*>(p)[i];
        px.v = reinterpret_cast<V*>(x)[i];
        if (!done) {
            pr.v = __ldg(reinterpret_cast<const V*>(r) + i);
            if constexpr (PRE) pd.v = __ldg(reinterpret_cast<const V*>(dinv) + i);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) px.e[k] = add_rn(px.e[k], mul_rn(pp.e[k], alpha));
        reinterpret_cast<V*>(x)[i] = px.v;
        if (!done) {
#pragma unroll
            for (int k = 0; k < N; ++k) pp.e[k] = add_rn(mul_rn(pp.e[k], beta), PRE ? mul_rn(pd.e[k], pr.e[k]) : pr.e[k]);
            reinterpret_cast<V*>(p)[i] = pp.v;
        }
    }
    for (uint64_t i = nvec * N + tid; i < n; i += stride) {
        const T pv = p[i];
        x[i] = add_rn(x[i], mul_rn(pv, alpha));
        if (!done) p[i] = add_rn(mul_rn(pv, beta), PRE ? mul_rn(dinv[i], r[i]) : r[i]);
    }
}

// ---- single-reduction variant (cg_sr.cuh) ------------------------------------------------------------
// U: p = r + (p * beta); s = w + (s * beta); x += (p * alpha); r -= (s * alpha); S[RR_NEW] = r.r of this rank (f64, unrounded)
template <class T>
__global__ void __launch_bounds__(kCgThreads)
cgsr_update_kernel(T* __restrict__ x, T* __restrict__ r, T* __restrict__ p, T* __restrict__ s, const T* __restrict__ w, uint64_t n,
                   double* __restrict__ S, double* __restrict__ partials, unsigned int* __restrict__ ticket, int vec_ok) {
    __shared__ double scratch[kCgThreads / 32 + 1];
    if (__ldcg(S + S_DONE) != 0.0) return;
    const T alpha = (T)__ldcg(S + SR_ALPHA), beta = (T)__ldcg(S + SR_BETA);
    const bool first = beta == T(0);           // p = r, s = w whatever an earlier solve left in p and s
    using V = typename Vec16<T>::type;
    constexpr int N = Vec16<T>::N;
    const uint64_t tid = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kCgThreads;
    const uint64_t nvec = vec_ok ? n / N : 0;
    T lane_acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) lane_acc[k] = T(0);
    for (uint64_t i = tid; i < nvec; i += stride) {
        Pack16<T> px, pr, pp, ps, pw;
        pw.v = __ldg(reinterpret_cast<const V*>(w) + i);
        pr.v = reinterpret_cast<V*>(r)[i];
        if (!first) { pp.v = reinterpret_cast<V*>(p)[i]; ps.v = reinterpret_cast<V*>(s)[i]; }
        px.v = reinterpret_cast<V*>(x)[i];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            pp.e[k] = first ? pr.e[k] : add_rn(pr.e[k], mul_rn(pp.e[k], beta));
            ps.e[k] = first ? pw.e[k] : add_rn(pw.e[k], mul_rn(ps.e[k], beta));
            px.e[k] = add_rn(px.e[k], mul_rn(pp.e[k], alpha));
            pr.e[k] = sub_rn(pr.e[k], mul_rn(ps.e[k], alpha));
            lane_acc[k] = add_rn(lane_acc[k], mul_rn(pr.e[k], pr.e[k]));
        }
        reinterpret_cast<V*>(p)[i] = pp.v;
        reinterpret_cast<V*>(s)[i] = ps.v;
        reinterpret_cast<V*>(x)[i] = px.v;
        reinterpret_cast<V*>(r)[i] = pr.v;
    }
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) acc += (double)lane_acc[k];
    for (uint64_t i = nvec * N + tid; i < n; i += stride) {
        const T rv0 = r[i];
        const T pv = first ? rv0 : add_rn(rv0, mul_rn(p[i], beta));
        const T sv = first ? w[i] : add_rn(w[i], mul_rn(s[i], beta));
        p[i] = pv; s[i] = sv;
        x[i] = add_rn(x[i], mul_rn(pv, alpha));
        const T rv = sub_rn(rv0, mul_rn(sv, alpha));
        r[i] = rv;
        acc += (double)mul_rn(rv, rv);
    }
    const double bsum = block_sum<kCgThreads>(acc, scratch);
    double total;
    if (grid_sum<kCgThreads>(bsum, partials, ticket, scratch, total))
        if (threadIdx.x == 0) S[S_RR_NEW] = total;
}

// S in a kernel of its own (one warp): local plans whose product is more than one launch, and the NCCL fallback (the host
// has all-reduced S[1..4] in front of it: ar == nullptr).
template <class T>
__global__ void __launch_bounds__(32)
cgsr_scalar_kernel(double* __restrict__ S, const ArDev* __restrict__ ar, double* __restrict__ history, unsigned long long hist_cap) {
    __shared__ double ar_sv[kMaxPeers][kArSlots], ar_in[kArSlots], ar_out[kArSlots];
    if (__ldcg(S + S_DONE) != 0.0) return;
    if (threadIdx.x == 0) {
        ar_in[0] = __ldcg(S + S_PAP) + __ldcg(S + S_PAP + 1) + __ldcg(S + S_PAP + 2);
        ar_in[1] = __ldcg(S + S_RR_NEW);
        ar_out[0] = ar_in[0]; ar_out[1] = ar_in[1];
    }
    __syncwarp();
    if (ar != nullptr) ar_warp_allreduce(*ar, ar_in, ar_out, 2, ar_sv);
    __syncwarp();
    if (threadIdx.x == 0) cgsr_scalars<T>(S, ar_out[0], ar_out[1], history, hist_cap);
}

void cg_free(CgWork& w) {
    if (w.r) cudaFree(w.r);
    if (w.p) cudaFree(w.p);
    if (w.ap) cudaFree(w.ap);
    if (w.s) cudaFree(w.s);
    if (w.scalars) cudaFree(w.scalars);
    if (w.scalars_host) cudaFreeHost(w.scalars_host);
    if (w.history) cudaFree(w.history);
    if (w.dinv) cudaFree(w.dinv);
    if (w.graph) cudaGraphExecDestroy(w.graph);
    w = CgWork();
}

static unsigned cg_grid(const smb200_ctx* ctx, uint64_t n, int vt) {
    const uint64_t items = n / (16 / vsize(vt)) + 1;
    uint64_t need = (items + kCgThreads - 1) / kCgThreads;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    return (unsigned)(need < cap ? need : cap);
}

// p_cap: allocated elements of p (dist solves keep ghost room behind the owned part)
smb200_status cg_prepare(smb200_ctx* ctx, CgWork& w, int vt, uint64_t n, uint64_t p_cap, uint64_t iter_max) {
    if (w.n != n || !w.r || w.cap < (p_cap > n ? p_cap : n) || w.hist_cap < (iter_max < (1u << 20) ? iter_max : (1u << 20))) {
        cudaStreamSynchronize(ctx->stream);
        cg_free(w);
        const size_t es = vsize(vt);
        SMB_TRY(dev_alloc(&w.r, (p_cap > n ? p_cap : n) * es));     // the single-reduction variant multiplies r: ghost room like p
        SMB_TRY(dev_alloc(&w.p, (p_cap > n ? p_cap : n) * es));
        SMB_CUDA(cudaMemsetAsync(w.r, 0, (p_cap > n ? p_cap : n) * es + kPadBytes, ctx->stream));
        w.cap = p_cap > n ? p_cap : n;
        SMB_TRY(dev_alloc(&w.ap, n * es));
        SMB_CUDA(cudaMemsetAsync(w.p, 0, (p_cap > n ? p_cap : n) * es + kPadBytes, ctx->stream));
        SMB_CUDA(cudaMalloc(&w.scalars, S_COUNT * sizeof(double)));
        SMB_CUDA(cudaHostAlloc(&w.scalars_host, 4 * S_COUNT * sizeof(double), cudaHostAllocDefault));
        w.hist_cap = iter_max < (1u << 20) ? iter_max : (1u << 20);
        if (w.hist_cap == 0) w.hist_cap = 1;
        SMB_CUDA(cudaMalloc(&w.history, w.hist_cap * sizeof(double)));
        w.n = n;
    }
    SMB_TRY(ensure_reduction_scratch(ctx, (size_t)ctx->sm_count * 16 + 16));     // two partials per CTA (preconditioned r update)
    return SMB200_OK;
}

smb200_status cg_init_launch(smb200_ctx* ctx, CgWork& w, int vt, const void* b, uint64_t n, const void* dinv, int rr_slot) {
    const unsigned g = cg_grid(ctx, n, vt);
    const int slot = rr_slot >= 0 ? rr_slot : ctx->world > 1 ? S_RR_LOCAL : S_RR_NEW;
    const int vec_ok = ((uintptr_t)b & 15u) == 0 ? 1 : 0;
    if (dinv) {
        if (vt == SMB200_F64) cg_init_kernel<double, true><<<g, kCgThreads, 0, ctx->stream>>>((const double*)b, (const double*)w.ap, (double*)w.r, (double*)w.p, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, vec_ok, (const double*)dinv);
        else cg_init_kernel<float, true><<<g, kCgThreads, 0, ctx->stream>>>((const float*)b, (const float*)w.ap, (float*)w.r, (float*)w.p, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, vec_ok, (const float*)dinv);
    } else if (vt == SMB200_F64) cg_init_kernel<double, false><<<g, kCgThreads, 0, ctx->stream>>>((const double*)b, (const double*)w.ap, (double*)w.r, (double*)w.p, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, vec_ok, nullptr);
    else cg_init_kernel<float, false><<<g, kCgThreads, 0, ctx->stream>>>((const float*)b, (const float*)w.ap, (float*)w.r, (float*)w.p, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, vec_ok, nullptr);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

smb200_status cg_r_launch(smb200_ctx* ctx, CgWork& w, int vt, uint64_t n, const void* dinv, const ArDev* ar) {
    const unsigned g = cg_grid(ctx, n, vt);
    const int slot = ctx->world > 1 ? S_RR_LOCAL : S_RR_NEW;
    if (dinv) {
        if (vt == SMB200_F64) cg_update_r_kernel<double, true><<<g, kCgThreads, 0, ctx->stream>>>((double*)w.r, (const double*)w.ap, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, (const double*)dinv, nullptr);
        else cg_update_r_kernel<float, true><<<g, kCgThreads, 0, ctx->stream>>>((float*)w.r, (const float*)w.ap, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, (const float*)dinv, nullptr);
    } else if (vt == SMB200_F64) cg_update_r_kernel<double, false><<<g, kCgThreads, 0, ctx->stream>>>((double*)w.r, (const double*)w.ap, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, nullptr, ar);
    else cg_update_r_kernel<float, false><<<g, kCgThreads, 0, ctx->stream>>>((float*)w.r, (const float*)w.ap, n, w.scalars, slot, ctx->red_partials, ctx->red_ticket, nullptr, ar);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

smb200_status cg_xp_launch(smb200_ctx* ctx, CgWork& w, int vt, void* x, uint64_t n, const void* dinv) {
    const unsigned g = cg_grid(ctx, n, vt);
    const int vec_ok = ((uintptr_t)x & 15u) == 0 ? 1 : 0;
    if (dinv) {
        if (vt == SMB200_F64) cg_update_xp_kernel<double, true><<<g, kCgThreads, 0, ctx->stream>>>((double*)x, (double*)w.p, (const double*)w.r, n, w.scalars, w.history, w.hist_cap, vec_ok, (const double*)dinv);
        else cg_update_xp_kernel<float, true><<<g, kCgThreads, 0, ctx->stream>>>((float*)x, (float*)w.p, (const float*)w.r, n, w.scalars, w.history, w.hist_cap, vec_ok, (const float*)dinv);
    } else if (vt == SMB200_F64) cg_update_xp_kernel<double, false><<<g, kCgThreads, 0, ctx->stream>>>((double*)x, (double*)w.p, (const double*)w.r, n, w.scalars, w.history, w.hist_cap, vec_ok, nullptr);
    else cg_update_xp_kernel<float, false><<<g, kCgThreads, 0, ctx->stream>>>((float*)x, (float*)w.p, (const float*)w.r, n, w.scalars, w.history, w.hist_cap, vec_ok, nullptr);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

// single-reduction variant: the workspace's fourth vector, allocated when first needed
smb200_status cgsr_prepare(smb200_ctx* ctx, CgWork& w, int vt) {
    if (!w.s) {
        SMB_TRY(dev_alloc(&w.s, w.cap * vsize(vt)));
        SMB_CUDA(cudaMemsetAsync(w.s, 0, w.cap * vsize(vt), ctx->stream));
    }
    return SMB200_OK;
}

smb200_status cgsr_update_launch(smb200_ctx* ctx, CgWork& w, int vt, void* x, uint64_t n) {
    const unsigned g = cg_grid(ctx, n, vt);
    const int vec_ok = ((uintptr_t)x & 15u) == 0 ? 1 : 0;
    if (vt == SMB200_F64) cgsr_update_kernel<double><<<g, kCgThreads, 0, ctx->stream>>>((double*)x, (double*)w.r, (double*)w.p, (double*)w.s, (const double*)w.ap, n, w.scalars, ctx->red_partials, ctx->red_ticket, vec_ok);
    else cgsr_update_kernel<float><<<g, kCgThreads, 0, ctx->stream>>>((float*)x, (float*)w.r, (float*)w.p, (float*)w.s, (const float*)w.ap, n, w.scalars, ctx->red_partials, ctx->red_ticket, vec_ok);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

smb200_status cgsr_scalar_launch(smb200_ctx* ctx, CgWork& w, int vt, const ArDev* ar) {
    if (vt == SMB200_F64) cgsr_scalar_kernel<double><<<1, 32, 0, ctx->stream>>>(w.scalars, ar, w.history, w.hist_cap);
    else cgsr_scalar_kernel<float><<<1, 32, 0, ctx->stream>>>(w.scalars, ar, w.history, w.hist_cap);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

// One CG iteration on the context stream (single GPU).
static smb200_status cg_iteration(smb200_crs* a, void* x, const void* dinv) {
    smb200_ctx* ctx = a->ctx;
    CgWork& w = a->cg;
    SMB_TRY(spmv_launch_cg(a, a->plan, 0, a->n_rows, w.p, w.ap, w.p, w.scalars, 0, true));
    SMB_TRY(cg_r_launch(ctx, w, a->vt, a->n_rows, dinv));
    SMB_TRY(cg_xp_launch(ctx, w, a->vt, x, a->n_rows, dinv));
    return SMB200_OK;
}

// The solver behind smb200_cg_solve (dinv == nullptr: the reference's ConjugateGradient) and smb200_pcg_jacobi_solve
// (dinv = 1 / diag(A): the additive Jacobi-preconditioned variant, same kernels with one more operand).
smb200_status cg_solve_impl(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                            uint64_t iter_max, smb200_cg_stats* stats, const void* dinv) {
    SMB_REQUIRE(a && b && x, SMB200_ERR_INVALID, "cg_solve: NULL argument");
    SMB_REQUIRE(b->vt == a->vt && x->vt == a->vt, SMB200_ERR_INVALID, "cg_solve: value types differ");
    SMB_REQUIRE(b->ctx == a->ctx && x->ctx == a->ctx, SMB200_ERR_INVALID, "cg_solve: operands belong to different contexts");
    SMB_REQUIRE(a->n_rows == a->n_cols, SMB200_ERR_NOT_SQUARE, "Matrix is not symmetric");            // linearsolver.rs:30-32
    SMB_REQUIRE(a->n_rows == b->n && a->n_rows == x->n, SMB200_ERR_SIZE_MISMATCH, "Matrix and vector size mismatch");  // :33-36
    smb200_ctx* ctx = a->ctx;
    CgWork& w = a->cg;
    const uint64_t n = a->n_rows;
    const uint64_t launches0 = g_launches;
    if (stats) memset(stats, 0, sizeof *stats);
    if (n == 0) return SMB200_OK;
    SMB_CUDA(cudaSetDevice(ctx->device));
    if (!a->plan.built) SMB_TRY(plan_build(a));
    SMB_TRY(cg_prepare(ctx, w, a->vt, n, n, iter_max));

    double threshold = tol;
    if (relative) {
        double bb = 0.0;
        SMB_TRY(dot_launch(ctx, a->vt, b->d, b->d, n, 0));
        SMB_TRY(fetch_result(ctx, 0, &bb));
        threshold = tol * sqrt(bb);
    }
    struct Events {                       // destroyed on every return path
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, poll[2] = {nullptr, nullptr};
        ~Events() { for (cudaEvent_t e : {ev0, ev1, poll[0], poll[1]}) if (e) cudaEventDestroy(e); }
    } evs;
    SMB_CUDA(cudaEventCreate(&evs.ev0));
    SMB_CUDA(cudaEventCreate(&evs.ev1));
    SMB_CUDA(cudaEventCreateWithFlags(&evs.poll[0], cudaEventDisableTiming));
    SMB_CUDA(cudaEventCreateWithFlags(&evs.poll[1], cudaEventDisableTiming));
    cudaEvent_t ev0 = evs.ev0, ev1 = evs.ev1;
    cudaEvent_t* poll_ev = evs.poll;
    SMB_CUDA(cudaEventRecord(ev0, ctx->stream));

    double init[S_COUNT] = {0};
    init[S_THRESH] = threshold;
    memcpy(w.scalars_host + 3 * S_COUNT, init, sizeof init);
    SMB_CUDA(cudaMemcpyAsync(w.scalars, w.scalars_host + 3 * S_COUNT, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    // r = b - A x ; p = r ; rr = r.r
    g_x_unpadded = !x->owned;                                   // the start vector may be borrowed memory
    const smb200_status st0 = spmv_launch_plan(a, a->plan, 0, n, x->d, w.ap, nullptr, 0);
    g_x_unpadded = false;
    SMB_TRY(st0);
    SMB_TRY(cg_init_launch(ctx, w, a->vt, b->d, n, dinv));

    const char* genv = getenv("SMB200_CG_GRAPH");
    const bool use_graph = !(genv && genv[0] == '0');
    const char* benv = getenv("SMB200_CG_BATCH");
    int batch = benv ? atoi(benv) : 8;
    if (batch < 1) batch = 1;

    uint64_t launched = 0;
    smb200_status st = SMB200_OK;
    uint64_t rounds = 0;
    bool finished = false;
    // first iteration eagerly: performs every lazy initialisation outside of stream capture
    if (iter_max > 0) { st = cg_iteration(a, x->d, dinv); launched = 1; }
    while (st == SMB200_OK && !finished) {
        // publish (iteration, done) of everything launched so far
        const int slot = (int)(rounds & 1);
        cudaError_t e = cudaMemcpyAsync(w.scalars_host + slot * S_COUNT, w.scalars, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(poll_ev[slot], ctx->stream);
        if (e != cudaSuccess) { set_error("cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
        // enqueue the next batch before looking at the previous snapshot
        uint64_t nb = iter_max - launched < (uint64_t)batch ? iter_max - launched : (uint64_t)batch;
        if (nb > 0) {
            if (use_graph && nb == (uint64_t)batch) {
                if (!w.graph || w.graph_batch != batch || w.graph_x != x->d || w.graph_partials != ctx->red_partials ||
                    w.graph_plan != a->plan.blk_rows || w.graph_dinv != dinv) {
                    if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; }
                    cudaGraph_t graph = nullptr;
                    e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
                    if (e == cudaSuccess) {
                        for (int k = 0; k < batch && st == SMB200_OK; ++k) st = cg_iteration(a, x->d, dinv);
                        cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
                        if (st == SMB200_OK && e2 != cudaSuccess) e = e2;
                    }
                    if (st == SMB200_OK && e == cudaSuccess) e = cudaGraphInstantiate(&w.graph, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                    if (e != cudaSuccess && st == SMB200_OK) { set_error("cg_solve: graph capture failed: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; }
                    if (st != SMB200_OK) break;
                    w.graph_batch = batch;
                    w.graph_x = x->d;
                    w.graph_partials = ctx->red_partials;
                    w.graph_plan = a->plan.blk_rows;
                    w.graph_dinv = dinv;
                    w.graph_kind = 0;
                }
                e = cudaGraphLaunch(w.graph, ctx->stream);
                if (e != cudaSuccess) { set_error("cg_solve: graph launch failed: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
                count_launch(4 * (uint64_t)batch);   // SpMV+dot, dot finalize, r update, x/p update
            } else {
                for (uint64_t k = 0; k < nb && st == SMB200_OK; ++k) st = cg_iteration(a, x->d, dinv);
                if (st != SMB200_OK) break;
            }
            launched += nb;
        }
        // look at the snapshot taken before this batch
        e = cudaEventSynchronize(poll_ev[slot]);
        if (e != cudaSuccess) { set_error("cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; break; }
        const double* snap = w.scalars_host + slot * S_COUNT;
        if (snap[S_DONE] != 0.0) finished = true;
        if (nb == 0) finished = true;
        ++rounds;
    }
    if (st == SMB200_OK) {
        cudaError_t e = cudaEventRecord(ev1, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(w.scalars_host, w.scalars, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("cg_solve: %s", cudaGetErrorString(e)); st = SMB200_ERR_CUDA; }
    }
    if (st == SMB200_OK) {
        const double* S = w.scalars_host;
        const uint64_t iters = (uint64_t)S[S_ITER];
        w.history_host.assign(iters < w.hist_cap ? iters : w.hist_cap, 0.0);
        if (!w.history_host.empty())
            cudaMemcpy(w.history_host.data(), w.history, w.history_host.size() * sizeof(double), cudaMemcpyDeviceToHost);
        if (stats) {
            stats->iterations = iters;
            stats->final_residual = sqrt(dinv ? S[S_RES2] : S[S_RR_NEW]);
            stats->converged = S[S_DONE] != 0.0;
            cudaEventElapsedTime(&stats->device_ms, ev0, ev1);
            stats->launches = g_launches - launches0;
        }
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    return st;
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_cg_solve(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                              uint64_t iter_max, smb200_cg_stats* stats) {
    return cg_solve_impl(a, b, x, tol, relative, iter_max, stats, nullptr);
}

smb200_status smb200_cg_history(const smb200_crs* a, double* out, uint64_t cap, uint64_t* n) {
    SMB_REQUIRE(a && n, SMB200_ERR_INVALID, "cg_history: NULL argument");
    const auto& h = a->cg.history_host;
    *n = h.size();
    if (out) for (uint64_t k = 0; k < cap && k < h.size(); ++k) out[k] = h[k];
    return SMB200_OK;
}

}  // extern "C"
