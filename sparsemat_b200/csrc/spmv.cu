// THE hot path: y = A x for a CRS matrix (SparseMatrix::mvp, sparsematrix.rs:146-158, over
// SparseMatCRS::iter_row, sparsemat_crs.rs:102-110).
//
// Roofline: HBM.  Algorithmic bytes per launch
//     B = nnz*(sizeof T + sizeof I) + (n_rows+1)*sizeof I + n_cols*sizeof T + n_rows*sizeof T
// (values and columns streamed once, offsets once, x once — its re-use is served by L1/L2 —, y once).
//
// Kernel families:
//   K1/K2  spmv_vector_kernel<LANES>   LANES threads per row (1 = scalar, 2..16 = sub-warp, 32 = warp),
//                                      shuffle reduction; LANES is chosen from the mean row length.
//   K3     spmv_stream_kernel          "merge-path, row-snapped": CTA k owns the rows whose merge
//                                      coordinate row + offset_rows[row] falls into [k*TARGET,(k+1)*TARGET),
//                                      i.e. an equal share of rows + non-zeros.  The CTA streams its
//                                      contiguous slice of values/columns with 128-bit loads, multiplies
//                                      by the gathered x on the fly, parks the products in shared memory
//                                      and then sums every row in STORAGE ORDER with one thread per row.
//                                      Rows of <= 64 non-zeros therefore reproduce the reference's
//                                      sequential `sum += x[col]*val` bit for bit; longer rows use a
//                                      warp (or the CTA) and fall under the documented tolerance.
//   K3-TMA spmv_stream_tma_kernel      same, but the slice is brought in by two cp.async.bulk (TMA)
//                                      copies completing on an mbarrier; products overwrite the values.
//   K4     spmv_banded_kernel          K3-TMA plus the block's x window [cmin,cmax] staged in shared
//                                      memory by a third bulk copy (banded / 2-D stencil matrices);
//          SMB200_FLAG_L2_PERSIST_X    L2 access-policy window over x for everything else.
// Every variant can fuse a dot product  sum_r w[r]*y[r]  into its epilogue (K7a, used by CG).
#include "common.cuh"
#include "cg_sr.cuh"
#include "halo.cuh"
#include "reduce.cuh"

#include <algorithm>
#include <chrono>
#include <type_traits>
#include <cstdlib>

namespace smb {

constexpr int kSpmvThreads = 256;
constexpr int kWarpRowMin = 64;      // rows longer than this are reduced by a warp inside the stream kernels
constexpr int kBigRow = 1024;        // fallback path: rows at least this long are reduced by the whole CTA
constexpr int kLongCap = 192;        // per-CTA list of rows deferred to the warp / CTA pass (cap / 65 at most)
constexpr unsigned kMaxCap = 12288;  // staging capacity limit implied by kLongCap
constexpr unsigned kBlkNoFit = 1u;   // PIPE block flags: slice larger than a stage -> direct path
constexpr unsigned kBlkLong = 2u;    //                   a row of more than kWarpRowMin entries -> two-pass row sums
constexpr unsigned kBlkShort = 4u;   //                   every row has <= kRowMajorMax entries -> row-major path
constexpr int kRowMajorMax = 32;
constexpr int kRingRowMax = 256;
constexpr unsigned long long kSliceAlign = 16;   // RING: slices start / end at multiples of 16 entries (16 bytes of 8-bit value codes)     // RING: longest row it takes (8 lanes per row x 32 entries per lane)
constexpr int kPipeMaxStages = 8;

struct DotArgs {
    const void* w;            // weights (nullptr = no fused dot)
    double* partials;
    unsigned int* ticket;
    double* result;
    // CG plumbing (cg.cu): the CTA that finishes last also rolls rr <- rr_new, and every CTA leaves
    // at once when the solver's stop flag is already set.
    double* roll_dst;
    const double* roll_src;
    const double* done;
    // distributed CG (dist.cu): the finalize kernel also sums the ranks' totals through peer memory (halo.cuh)
    const ArDev* ar;
};

template <class T> __device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// y[r] = sum, or y[r] += sum for the second and later column bands of a band-split product (BANDSPLIT); returns what was
// stored (the fused dot weighs the final value).
template <class T> __device__ __forceinline__ T store_y(T* __restrict__ y, uint64_t r, T sum, bool accumulate) {
    if (accumulate) sum = add_rn(y[r], sum);
    y[r] = sum;
    return sum;
}

// Fused dot epilogue of the SpMV kernels: every CTA leaves ONE f64 partial; spmv_dot_finalize_kernel (launched
// right behind on the same stream) folds them in index order.  No ticket atomics and no __threadfence in the hot
// kernel: with ~40k CTAs a single-address ticket costs more than a quarter of the product itself (measured).
template <class T, int THREADS = kSpmvThreads>
__device__ __forceinline__ void finish_dot(double acc, const DotArgs& d) {
    __shared__ double scratch[THREADS / 32 + 1];
    const double bsum = block_sum<THREADS>(acc, scratch);
    if (threadIdx.x == 0) d.partials[blockIdx.x] = bsum;
}

constexpr int kFinalizeThreads = 1024;
__global__ void __launch_bounds__(kFinalizeThreads)
spmv_dot_finalize_kernel(const double* __restrict__ partials, unsigned n, int is_f32, double* __restrict__ result,
                         double* __restrict__ roll_dst, const double* __restrict__ roll_src, const double* __restrict__ done,
                         const ArDev* __restrict__ ar, CgSrLaunch sr) {
    __shared__ double scratch[kFinalizeThreads / 32 + 1];
    __shared__ double ar_sv[kMaxPeers][kArSlots], ar_in[kArSlots], ar_out[kArSlots];
    if (done != nullptr && __ldcg(done) != 0.0) return;      // the SpMV CTAs left early: keep the previous values
    double acc = 0.0;
    for (unsigned i = threadIdx.x; i < n; i += kFinalizeThreads) acc += __ldcg(partials + i);
    double total = block_sum<kFinalizeThreads>(acc, scratch);
    if (sr.S != nullptr) {
        // single-reduction CG (cg_sr.cuh): this product's (A r).r and the r.r the update kernel left travel in ONE
        // all-reduce, and the iteration's scalar step runs right here
        if (threadIdx.x == 0) { ar_in[0] = total; ar_in[1] = __ldcg(sr.S + 4); ar_out[0] = ar_in[0]; ar_out[1] = ar_in[1]; }
        __syncthreads();
        if (ar != nullptr && threadIdx.x < 32) ar_warp_allreduce(*ar, ar_in, ar_out, 2, ar_sv);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (is_f32) cgsr_scalars<float>(sr.S, ar_out[0], ar_out[1], sr.history, sr.hist_cap);
            else cgsr_scalars<double>(sr.S, ar_out[0], ar_out[1], sr.history, sr.hist_cap);
        }
        return;
    }
    if (ar != nullptr) {
        // one rank per GPU: the ranks' totals are exchanged through peer memory right here (every rank runs this kernel at
        // the same point of its stream; the stop flag above is identical on all ranks), summed in rank order
        if (threadIdx.x == 0) ar_in[0] = total;
        __syncthreads();
        if (threadIdx.x < 32) ar_warp_allreduce(*ar, ar_in, ar_out, 1, ar_sv);
        __syncthreads();
        total = ar_out[0];
    }
    if (threadIdx.x == 0) {
        *result = is_f32 ? (double)(float)total : total;
        if (roll_dst) *roll_dst = *roll_src;
    }
}
__device__ __forceinline__ bool solver_done(const DotArgs& d) { return d.done != nullptr && __ldcg(d.done) != 0.0; }

// ---- K1 / K2: LANES threads per row -----------------------------------------------------------------
template <class T, class I, int LANES, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_vector_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                   uint64_t row_begin, uint64_t row_end, const T* __restrict__ x, T* __restrict__ y, DotArgs dot) {
    constexpr int ROWS = kSpmvThreads / LANES;
    if constexpr (DOT) { if (solver_done(dot)) return; }
    const uint64_t row = row_begin + (uint64_t)blockIdx.x * ROWS + threadIdx.x / LANES;
    const int sub = threadIdx.x % LANES;
    double acc = 0.0;
    T s = T(0);
    if (row < row_end) {
        const uint64_t a = (uint64_t)__ldg(offs + row), e = (uint64_t)__ldg(offs + row + 1);
        for (uint64_t k = a + sub; k < e; k += LANES)
            s = add_rn(s, mul_rn(__ldg(x + (size_t)__ldcs(cols + k)), __ldcs(vals + k)));
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) s = add_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (row < row_end && sub == 0) {
        y[row] = s;
        if constexpr (DOT) acc = (double)mul_rn(__ldg((const T*)dot.w + row), s);
    }
    if constexpr (DOT) finish_dot<T>(acc, dot);
}

// ---- shared pieces of the stream kernels ------------------------------------------------------------
// 4 consecutive elements of an array, 16-byte (T,I of 4 bytes) or 32-byte (8 bytes) aligned.
template <class E> struct Quad { E e[4]; };

// Streamed once: bypass L1 (it is kept for the x gathers) and mark the line evict-first in L2.  8-byte element
// quads are one 256-bit load (LDG.E.256, new on sm_100).
template <class E> __device__ __forceinline__ Quad<E> load_quad_stream(const E* p) {
    Quad<E> q;
    if constexpr (sizeof(E) == 4) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        memcpy(q.e, &v, 16);
    } else {
        uint4 v0, v1;
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w), "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "l"(p));
        memcpy(q.e, &v0, 16);
        memcpy(q.e + 2, &v1, 16);
    }
    return q;
}
// Band-split products (bandsplit.cu): the band of x being gathered must stay in L2 while everything else streams past
// it, so the streams are marked evict-first in L2 (4-byte quads too) and the gathers evict-last.
// (The .L2::evict_* qualifiers exist for 256-bit loads only; narrower accesses take a policy operand.)
struct L2Policy { uint64_t first, last; };
__device__ __forceinline__ L2Policy make_l2_policies() {
    L2Policy p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.last));
    return p;
}
template <class E> __device__ __forceinline__ Quad<E> load_quad_evict_first(const E* p, uint64_t pol) {
    if constexpr (sizeof(E) == 4) {
        Quad<E> q;
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
        memcpy(q.e, &v, 16);
        return q;
    } else {
        return load_quad_stream(p);
    }
}
template <class T> __device__ __forceinline__ T gather_hint(const T* p, uint64_t pol) {
    T v;
    if constexpr (sizeof(T) == 8) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
template <class T> __device__ __forceinline__ T load_hint(const T* p, uint64_t pol) {
    T v;
    if constexpr (sizeof(T) == 8) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    else asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
template <class T> __device__ __forceinline__ void store_hint(T* p, T v, uint64_t pol) {
    if constexpr (sizeof(T) == 8) asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
    else asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

template <class E> __device__ __forceinline__ Quad<E> load_quad_shared(const E* p) {
    Quad<E> q;
    if constexpr (sizeof(E) == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        memcpy(q.e, &v, 16);
    } else {
        const uint4 v0 = *reinterpret_cast<const uint4*>(p);
        const uint4 v1 = *(reinterpret_cast<const uint4*>(p) + 1);
        memcpy(q.e, &v0, 16);
        memcpy(q.e + 2, &v1, 16);
    }
    return q;
}
template <class E> __device__ __forceinline__ void store_quad_shared(E* p, const Quad<E>& q) {
    if constexpr (sizeof(E) == 4) {
        uint4 v; memcpy(&v, q.e, 16);
        *reinterpret_cast<uint4*>(p) = v;
    } else {
        uint4 v0, v1; memcpy(&v0, q.e, 16); memcpy(&v1, q.e + 2, 16);
        *reinterpret_cast<uint4*>(p) = v0;
        *(reinterpret_cast<uint4*>(p) + 1) = v1;
    }
}

// Row sums out of the staged products.  prod[k] holds the product of element a0 + k.
// Pass 1: one thread per row, storage order (bit-exact).  Pass 2: rows longer than kWarpRowMin, a warp each.
template <class T, class I, bool DOT, int THREADS = kSpmvThreads>
__device__ __forceinline__ double reduce_rows_from_smem(const T* prod, const I* __restrict__ offs, uint64_t r0, uint64_t r1,
                                                        uint64_t a0, T* __restrict__ y, const T* __restrict__ w,
                                                        unsigned int* s_long_count, unsigned int* s_long_rows, bool accumulate = false) {
    double acc = 0.0;
    for (uint64_t r = r0 + threadIdx.x; r < r1; r += THREADS) {
        const unsigned a = (unsigned)((uint64_t)__ldg(offs + r) - a0);
        const unsigned e = (unsigned)((uint64_t)__ldg(offs + r + 1) - a0);
        if (e - a <= (unsigned)kWarpRowMin) {
            T s = T(0);
            for (unsigned k = a; k < e; ++k) s = add_rn(s, prod[k]);
            s = store_y(y, r, s, accumulate);
            if constexpr (DOT) acc += (double)mul_rn(__ldg(w + r), s);
        } else {
            const unsigned slot = atomicAdd(s_long_count, 1u);
            s_long_rows[slot] = (unsigned)(r - r0);
        }
    }
    __syncthreads();
    const unsigned n_long = *s_long_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned j = warp; j < n_long; j += THREADS / 32) {
        const uint64_t r = r0 + s_long_rows[j];
        const unsigned a = (unsigned)((uint64_t)__ldg(offs + r) - a0);
        const unsigned e = (unsigned)((uint64_t)__ldg(offs + r + 1) - a0);
        T s = T(0);
        for (unsigned k = a + lane; k < e; k += 32) s = add_rn(s, prod[k]);
        s = warp_sum_t(s);
        if (lane == 0) {
            s = store_y(y, r, s, accumulate);
            if constexpr (DOT) acc += (double)mul_rn(__ldg(w + r), s);
        }
    }
    return acc;
}

// Blocks whose slice does not fit the staging buffer (they contain a very long row): classic
// CSR-vector inside the CTA — a warp per row, the whole CTA for rows >= kBigRow.
template <class T, class I, bool DOT, int THREADS = kSpmvThreads>
__device__ __forceinline__ double rows_direct(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                                              uint64_t r0, uint64_t r1, const T* __restrict__ x, T* __restrict__ y,
                                              const T* __restrict__ w, unsigned int* s_long_count, unsigned int* s_long_rows,
                                              double* scratch, bool accumulate = false) {
    double acc = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t r = r0 + warp; r < r1; r += THREADS / 32) {
        const uint64_t a = (uint64_t)__ldg(offs + r), e = (uint64_t)__ldg(offs + r + 1);
        if (e - a >= (uint64_t)kBigRow) {
            if (lane == 0) {
                const unsigned slot = atomicAdd(s_long_count, 1u);
                if (slot < (unsigned)kLongCap) s_long_rows[slot] = (unsigned)(r - r0);
            }
            continue;
        }
        T s = T(0);
        for (uint64_t k = a + lane; k < e; k += 32) s = add_rn(s, mul_rn(__ldg(x + (size_t)__ldcs(cols + k)), __ldcs(vals + k)));
        s = warp_sum_t(s);
        if (lane == 0) {
            s = store_y(y, r, s, accumulate);
            if constexpr (DOT) acc += (double)mul_rn(__ldg(w + r), s);
        }
    }
    __syncthreads();
    const unsigned n_big = min(*s_long_count, (unsigned)kLongCap);
    for (unsigned j = 0; j < n_big; ++j) {
        const uint64_t r = r0 + s_long_rows[j];
        const uint64_t a = (uint64_t)__ldg(offs + r), e = (uint64_t)__ldg(offs + r + 1);
        T s0 = T(0), s1 = T(0);
        uint64_t k = a + threadIdx.x;
        for (; k + THREADS < e; k += 2 * THREADS) {
            const T p0 = mul_rn(__ldg(x + (size_t)__ldcs(cols + k)), __ldcs(vals + k));
            const T p1 = mul_rn(__ldg(x + (size_t)__ldcs(cols + k + THREADS)), __ldcs(vals + k + THREADS));
            s0 = add_rn(s0, p0);
            s1 = add_rn(s1, p1);
        }
        if (k < e) s0 = add_rn(s0, mul_rn(__ldg(x + (size_t)__ldcs(cols + k)), __ldcs(vals + k)));
        const double tot = block_sum<THREADS>((double)s0 + (double)s1, scratch);
        if (threadIdx.x == 0) {
            const T s = store_y(y, r, (T)tot, accumulate);
            if constexpr (DOT) acc += (double)mul_rn(__ldg(w + r), s);
        }
    }
    return acc;
}

// ---- K3: stream kernel, register-staged 128-bit loads -----------------------------------------------
template <class T, class I, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, sizeof(T) == 4 ? 8 : 6)
spmv_stream_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                   const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz, unsigned cap, const T* __restrict__ x,
                   T* __restrict__ y, DotArgs dot, int accumulate) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* prod = reinterpret_cast<T*>(smem_raw);
    __shared__ unsigned int s_long_count;
    __shared__ unsigned int s_long_rows[kLongCap];
    __shared__ double scratch[kSpmvThreads / 32 + 1];
    if (threadIdx.x == 0) s_long_count = 0;
    // one round trip for the block's row range and element range (the plan stores offset_rows[blk_rows[k]]) and,
    // inside a CG solve, for the solver's stop flag
    const uint64_t r0 = (uint64_t)__ldg(blk_rows + blockIdx.x), r1 = (uint64_t)__ldg(blk_rows + blockIdx.x + 1);
    const uint64_t n0 = (uint64_t)__ldg(blk_nnz + blockIdx.x), n1 = (uint64_t)__ldg(blk_nnz + blockIdx.x + 1);
    double stop = 0.0;
    if constexpr (DOT) { if (dot.done != nullptr) stop = __ldcg(dot.done); }
    const uint64_t a0 = n0 & ~(uint64_t)3;
    if constexpr (DOT) { if (stop != 0.0) return; }
    double acc = 0.0;
    // Row offsets of the first two rows this thread will sum: requested before the stream loads, first touched
    // after the barrier, so their latency hides behind the whole product phase.
    constexpr int KP = 2;
    I pfa[KP], pfe[KP];
    T pfw[KP];                          // fused dot: the weights of those rows, requested just as early
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        const uint64_t r = r0 + threadIdx.x + (uint64_t)j * kSpmvThreads;
        pfa[j] = pfe[j] = 0;
        pfw[j] = T(0);
        if (r < r1) {
            pfa[j] = __ldg(offs + r); pfe[j] = __ldg(offs + r + 1);
            if constexpr (DOT) pfw[j] = __ldg((const T*)dot.w + r);
        }
    }
    __syncthreads();
    if (n1 - a0 <= (uint64_t)cap) {
        const unsigned groups = (unsigned)((n1 - a0 + 3) >> 2);
        const T* vbase = vals + a0;
        const I* cbase = cols + a0;
        unsigned g = threadIdx.x;
        // two quads in flight per thread: all streaming loads first, then the gathers
        for (; g + kSpmvThreads < groups; g += 2 * kSpmvThreads) {
            const Quad<I> c0 = load_quad_stream(cbase + 4 * (size_t)g);
            const Quad<I> c1 = load_quad_stream(cbase + 4 * (size_t)(g + kSpmvThreads));
            const Quad<T> v0 = load_quad_stream(vbase + 4 * (size_t)g);
            const Quad<T> v1 = load_quad_stream(vbase + 4 * (size_t)(g + kSpmvThreads));
            Quad<T> p0, p1;
#pragma unroll
            for (int k = 0; k < 4; ++k) p0.e[k] = __ldg(x + (size_t)c0.e[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) p1.e[k] = __ldg(x + (size_t)c1.e[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) { p0.e[k] = mul_rn(p0.e[k], v0.e[k]); p1.e[k] = mul_rn(p1.e[k], v1.e[k]); }
            store_quad_shared(prod + 4 * g, p0);
            store_quad_shared(prod + 4 * (g + kSpmvThreads), p1);
        }
        if (g < groups) {
            const Quad<I> c0 = load_quad_stream(cbase + 4 * (size_t)g);
            const Quad<T> v0 = load_quad_stream(vbase + 4 * (size_t)g);
            Quad<T> p0;
#pragma unroll
            for (int k = 0; k < 4; ++k) p0.e[k] = mul_rn(__ldg(x + (size_t)c0.e[k]), v0.e[k]);
            store_quad_shared(prod + 4 * g, p0);
        }
        __syncthreads();
        // rows whose offsets were prefetched and that are short enough for the storage-order sum
        uint64_t r_next = r0;
        bool all_short = true;
        double acc_fast = 0.0;
        // (accumulate mode — y += — skips this pass: a block that is redone generically below would add its rows twice)
        if (accumulate) all_short = false;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const uint64_t r = r0 + threadIdx.x + (uint64_t)j * kSpmvThreads;
            if (r < r1 && !accumulate) {
                const unsigned a = (unsigned)((uint64_t)pfa[j] - a0), e = (unsigned)((uint64_t)pfe[j] - a0);
                if (e - a <= (unsigned)kWarpRowMin) {
                    T sum = T(0);
                    for (unsigned k = a; k < e; ++k) sum = add_rn(sum, prod[k]);
                    y[r] = sum;
                    if constexpr (DOT) acc_fast += (double)mul_rn(pfw[j], sum);
                } else {
                    all_short = false;
                }
            }
        }
        r_next = r0 + (uint64_t)KP * kSpmvThreads;
        // everything else (rows beyond KP per thread, rows longer than kWarpRowMin) takes the generic two-pass path
        if (__syncthreads_or(!all_short)) { r_next = r0; acc_fast = 0.0; }    // redo the block generically
        acc = acc_fast;
        if (r_next < r1)
            acc += reduce_rows_from_smem<T, I, DOT>(prod, offs, r_next, r1, a0 + 0, y, (const T*)dot.w, &s_long_count, s_long_rows, accumulate != 0);
    } else {
        acc = rows_direct<T, I, DOT>(vals, cols, offs, r0, r1, x, y, (const T*)dot.w, &s_long_count, s_long_rows, scratch, accumulate != 0);
    }
    if constexpr (DOT) finish_dot<T>(acc, dot);
}

// ---- K3 for the column bands of a band-split product (bandsplit.cu) ------------------------------------
// Same shape as spmv_stream_kernel, tuned for parts with 1-3 entries per row: the matrix streams (and y, which every band
// but the first reads and all write) are evict-first in L2, the gathers evict-last, so the band of x survives; the row
// offsets, the old y and the dot weights of up to KP rows per thread are requested before the stream phase; whether the
// block takes the one-thread-per-row pass or the generic two-pass one is voted BEFORE anything is written (y += must
// happen once).
template <class T, class I, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, 4)
spmv_band_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                 const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz, unsigned cap, const T* __restrict__ x,
                 T* __restrict__ y, DotArgs dot, int accumulate) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* prod = reinterpret_cast<T*>(smem_raw);
    __shared__ unsigned int s_long_count;
    __shared__ unsigned int s_long_rows[kLongCap];
    __shared__ double scratch[kSpmvThreads / 32 + 1];
    if (threadIdx.x == 0) s_long_count = 0;
    const uint64_t r0 = (uint64_t)__ldg(blk_rows + blockIdx.x), r1 = (uint64_t)__ldg(blk_rows + blockIdx.x + 1);
    const uint64_t n0 = (uint64_t)__ldg(blk_nnz + blockIdx.x), n1 = (uint64_t)__ldg(blk_nnz + blockIdx.x + 1);
    double stop = 0.0;
    if constexpr (DOT) { if (dot.done != nullptr) stop = __ldcg(dot.done); }
    const uint64_t a0 = n0 & ~(uint64_t)3;
    if constexpr (DOT) { if (stop != 0.0) return; }
    const L2Policy pol = make_l2_policies();
    constexpr int KP = 4;
    I pfa[KP], pfe[KP];
    T pfy[KP], pfw[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        const uint64_t r = r0 + threadIdx.x + (uint64_t)j * kSpmvThreads;
        pfa[j] = pfe[j] = 0;
        pfy[j] = pfw[j] = T(0);
        if (r < r1) {
            pfa[j] = __ldg(offs + r); pfe[j] = __ldg(offs + r + 1);
            if (accumulate) pfy[j] = load_hint(y + r, pol.first);
            if constexpr (DOT) pfw[j] = __ldg((const T*)dot.w + r);
        }
    }
    double acc = 0.0;
    if (n1 - a0 <= (uint64_t)cap) {
        const unsigned groups = (unsigned)((n1 - a0 + 3) >> 2);
        const T* vbase = vals + a0;
        const I* cbase = cols + a0;
        unsigned g = threadIdx.x;
        for (; g + kSpmvThreads < groups; g += 2 * kSpmvThreads) {
            const Quad<I> c0 = load_quad_evict_first(cbase + 4 * (size_t)g, pol.first);
            const Quad<I> c1 = load_quad_evict_first(cbase + 4 * (size_t)(g + kSpmvThreads), pol.first);
            const Quad<T> v0 = load_quad_evict_first(vbase + 4 * (size_t)g, pol.first);
            const Quad<T> v1 = load_quad_evict_first(vbase + 4 * (size_t)(g + kSpmvThreads), pol.first);
            Quad<T> p0, p1;
#pragma unroll
            for (int k = 0; k < 4; ++k) p0.e[k] = gather_hint(x + (size_t)c0.e[k], pol.last);
#pragma unroll
            for (int k = 0; k < 4; ++k) p1.e[k] = gather_hint(x + (size_t)c1.e[k], pol.last);
#pragma unroll
            for (int k = 0; k < 4; ++k) { p0.e[k] = mul_rn(p0.e[k], v0.e[k]); p1.e[k] = mul_rn(p1.e[k], v1.e[k]); }
            store_quad_shared(prod + 4 * g, p0);
            store_quad_shared(prod + 4 * (g + kSpmvThreads), p1);
        }
        if (g < groups) {
            const Quad<I> c0 = load_quad_evict_first(cbase + 4 * (size_t)g, pol.first);
            const Quad<T> v0 = load_quad_evict_first(vbase + 4 * (size_t)g, pol.first);
            Quad<T> p0;
#pragma unroll
            for (int k = 0; k < 4; ++k) p0.e[k] = mul_rn(gather_hint(x + (size_t)c0.e[k], pol.last), v0.e[k]);
            store_quad_shared(prod + 4 * g, p0);
        }
        // vote: every row of the block is among the prefetched ones and short enough for the storage-order sum
        bool mine_ok = true;
#pragma unroll
        for (int j = 0; j < KP; ++j) mine_ok = mine_ok && ((uint64_t)pfe[j] - (uint64_t)pfa[j] <= (uint64_t)kWarpRowMin);
        const bool generic = __syncthreads_or(!mine_ok || (r1 - r0) > (uint64_t)KP * kSpmvThreads) != 0;   // (also: products are in place)
        if (!generic) {
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const uint64_t r = r0 + threadIdx.x + (uint64_t)j * kSpmvThreads;
                if (r < r1) {
                    const unsigned a = (unsigned)((uint64_t)pfa[j] - a0), e = (unsigned)((uint64_t)pfe[j] - a0);
                    T sum = T(0);
                    for (unsigned k = a; k < e; ++k) sum = add_rn(sum, prod[k]);
                    if (accumulate) sum = add_rn(pfy[j], sum);
                    store_hint(y + r, sum, pol.first);
                    if constexpr (DOT) acc += (double)mul_rn(pfw[j], sum);
                }
            }
        } else {
            acc = reduce_rows_from_smem<T, I, DOT>(prod, offs, r0, r1, a0, y, (const T*)dot.w, &s_long_count, s_long_rows, accumulate != 0);
        }
    } else {
        __syncthreads();
        acc = rows_direct<T, I, DOT>(vals, cols, offs, r0, r1, x, y, (const T*)dot.w, &s_long_count, s_long_rows, scratch, accumulate != 0);
    }
    if constexpr (DOT) finish_dot<T>(acc, dot);
}

// ---- TMA helpers (cp.async.bulk + mbarrier) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; size and both addresses are multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- K3-TMA / K4: slice (and optionally the x window) brought in by bulk copies ----------------------
template <class T, class I, bool DOT, bool XWIN>
__global__ void __launch_bounds__(kSpmvThreads)
spmv_stream_tma_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                       const I* __restrict__ blk_rows, const I* __restrict__ blk_win, unsigned cap, unsigned win_cap,
                       const T* __restrict__ x, T* __restrict__ y, DotArgs dot) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [values/products: cap*sizeof T][columns: cap*sizeof I][x window: win_cap*sizeof T]
    T* sv = reinterpret_cast<T*>(smem_raw);
    I* sc = reinterpret_cast<I*>(smem_raw + (size_t)cap * sizeof(T));
    T* sx = reinterpret_cast<T*>(smem_raw + (size_t)cap * (sizeof(T) + sizeof(I)));
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned int s_long_count;
    __shared__ unsigned int s_long_rows[kLongCap];
    __shared__ double scratch[kSpmvThreads / 32 + 1];
    if constexpr (DOT) { if (solver_done(dot)) return; }
    const uint64_t r0 = (uint64_t)__ldg(blk_rows + blockIdx.x), r1 = (uint64_t)__ldg(blk_rows + blockIdx.x + 1);
    const uint64_t n0 = (uint64_t)__ldg(offs + r0), n1 = (uint64_t)__ldg(offs + r1);
    const uint64_t a0 = n0 & ~(uint64_t)3;
    const unsigned count = (unsigned)(((n1 - a0) + 3) & ~(uint64_t)3);
    uint64_t w0 = 0;          // first x element of the window (aligned down to 16 bytes)
    unsigned wcount = 0;
    bool fits = (n1 - a0) <= (uint64_t)cap;
    if constexpr (XWIN) {
        const uint64_t cmin = (uint64_t)__ldg(blk_win + 2 * (size_t)blockIdx.x);
        const uint64_t cend = (uint64_t)__ldg(blk_win + 2 * (size_t)blockIdx.x + 1);
        constexpr uint64_t XA = 16 / sizeof(T);
        w0 = cmin & ~(XA - 1);
        const uint64_t wc = cend > w0 ? ((cend - w0 + XA - 1) & ~(XA - 1)) : 0;
        fits = fits && wc <= (uint64_t)win_cap;
        wcount = (unsigned)wc;
    }
    if (threadIdx.x == 0) {
        s_long_count = 0;
        mbar_init(&bar, 1);
    }
    __syncthreads();
    double acc = 0.0;
    if (fits) {
        if (threadIdx.x == 0) {
            unsigned bytes = count * (unsigned)(sizeof(T) + sizeof(I));
            if constexpr (XWIN) bytes += wcount * (unsigned)sizeof(T);
            mbar_expect_tx(&bar, bytes);
            bulk_g2s(sc, cols + a0, count * (unsigned)sizeof(I), &bar);
            bulk_g2s(sv, vals + a0, count * (unsigned)sizeof(T), &bar);
            if constexpr (XWIN) { if (wcount) bulk_g2s(sx, x + w0, wcount * (unsigned)sizeof(T), &bar); }
        }
        mbar_wait(&bar, 0);
        const unsigned groups = count >> 2;
        for (unsigned g = threadIdx.x; g < groups; g += kSpmvThreads) {
            const Quad<I> c = load_quad_shared(sc + 4 * g);
            Quad<T> v = load_quad_shared(sv + 4 * g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                T xv;
                if constexpr (XWIN) {
                    // elements outside [n0,n1) belong to neighbouring blocks: their columns may lie
                    // outside this block's window, so clamp (their products are never read)
                    const uint64_t rel = (uint64_t)c.e[k] - w0;
                    xv = sx[rel < (uint64_t)wcount ? rel : 0];
                } else {
                    xv = __ldg(x + (size_t)c.e[k]);
                }
                v.e[k] = mul_rn(xv, v.e[k]);
            }
            store_quad_shared(sv + 4 * g, v);
        }
        __syncthreads();
        acc = reduce_rows_from_smem<T, I, DOT>(sv, offs, r0, r1, a0, y, (const T*)dot.w, &s_long_count, s_long_rows);
    } else {
        acc = rows_direct<T, I, DOT>(vals, cols, offs, r0, r1, x, y, (const T*)dot.w, &s_long_count, s_long_rows, scratch);
    }
    if constexpr (DOT) finish_dot<T>(acc, dot);
}

// ---- K3 persistent: multi-stage TMA ring ---------------------------------------------------------------
// One CTA per SM slot walks the row blocks b = blockIdx.x, blockIdx.x + gridDim.x, ... .  Thread 0 keeps
// `stages - 1` blocks of values + columns in flight with cp.async.bulk into a shared-memory ring (one
// mbarrier per stage); all threads then multiply the landed slice by the gathered x in place and sum the
// rows in storage order.  One __syncthreads per block: the stage consumed in iteration i - 1 is refilled
// right after the barrier of iteration i, when every thread is known to have left it.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class E, int G> struct Grp { E e[G]; };
template <class E, int G> __device__ __forceinline__ Grp<E, G> lds_grp(const E* p) {
    static_assert(sizeof(E) * G == 16 || sizeof(E) * G == 8, "group must be 8 or 16 bytes");
    Grp<E, G> q;
    if constexpr (sizeof(E) * G == 16) { const uint4 v = *reinterpret_cast<const uint4*>(p); memcpy(q.e, &v, 16); }
    else { const uint2 v = *reinterpret_cast<const uint2*>(p); memcpy(q.e, &v, 8); }
    return q;
}
template <class E, int G> __device__ __forceinline__ void sts_grp(E* p, const Grp<E, G>& q) {
    if constexpr (sizeof(E) * G == 16) { uint4 v; memcpy(&v, q.e, 16); *reinterpret_cast<uint4*>(p) = v; }
    else { uint2 v; memcpy(&v, q.e, 8); *reinterpret_cast<uint2*>(p) = v; }
}

// ROWS = true: every row of the matrix has <= kRowMajorMax entries (stencils, FEM): only the row-major path is
// compiled in, which keeps the kernel lean enough for two 256-thread CTAs per SM.  ROWS = false: general rows.
template <class T, class I, int THREADS, bool DOT, bool ROWS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS)
spmv_pipe_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                 const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz, const unsigned char* __restrict__ blk_flags,
                 unsigned n_blocks, unsigned cap, unsigned stages, const T* __restrict__ x, T* __restrict__ y, DotArgs dot) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kPipeMaxStages];
    __shared__ unsigned long long s_desc[kPipeMaxStages][4];    // r0, r1, n0, n1 of the block staged there
    __shared__ unsigned int s_flags[kPipeMaxStages];
    __shared__ unsigned int s_long_count;
    __shared__ unsigned int s_long_rows[kLongCap];
    __shared__ double scratch[THREADS / 32 + 1];
    constexpr int G = 16 / (sizeof(T) > sizeof(I) ? sizeof(T) : sizeof(I));   // elements per 16-byte group
    constexpr int U = 4;                                                        // groups in flight per thread
    constexpr int K = sizeof(T) == 4 ? 3 : 2;                                   // rows per thread with prefetched offsets
    constexpr int CH = 8;                                                       // row-major path: entries per row in flight
    if constexpr (DOT) { if (solver_done(dot)) return; }
    const unsigned tid = threadIdx.x;
    const unsigned n_my = blockIdx.x < n_blocks ? (n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const size_t stage_bytes = (size_t)cap * (sizeof(T) + sizeof(I));
    if (tid == 0) {
        for (unsigned s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        s_long_count = 0;
    }
    __syncthreads();

    // thread 0: publish the descriptor of block j in its stage and start the bulk copies
    auto issue = [&](unsigned s, unsigned long long r0, unsigned long long r1, unsigned long long n0, unsigned long long n1,
                     unsigned fl) {
        s_desc[s][0] = r0; s_desc[s][1] = r1; s_desc[s][2] = n0; s_desc[s][3] = n1;
        s_flags[s] = fl;
        const unsigned long long a0 = n0 & ~3ull;
        const unsigned count = (fl & kBlkNoFit) ? 0u : (unsigned)(((n1 - a0) + 3ull) & ~3ull);
        if (count == 0) { mbar_arrive(&full[s]); return; }
        unsigned char* base = smem_raw + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], count * (unsigned)(sizeof(T) + sizeof(I)));
        bulk_g2s(base, vals + a0, count * (unsigned)sizeof(T), &full[s]);
        bulk_g2s(base + (size_t)cap * sizeof(T), cols + a0, count * (unsigned)sizeof(I), &full[s]);
    };
    auto blk = [&](unsigned j) { return (size_t)blockIdx.x + (size_t)j * gridDim.x; };

    // prologue: stages - 1 blocks in flight; the descriptors are fetched by the first lanes in parallel
    const unsigned pro = n_my < stages - 1 ? n_my : stages - 1;
    if (tid < 32) {
        unsigned long long r0 = 0, r1 = 0, n0 = 0, n1 = 0;
        unsigned fl = 0;
        if (tid < pro) {
            const size_t b = blk(tid);
            r0 = (unsigned long long)__ldg(blk_rows + b); r1 = (unsigned long long)__ldg(blk_rows + b + 1);
            n0 = (unsigned long long)__ldg(blk_nnz + b); n1 = (unsigned long long)__ldg(blk_nnz + b + 1);
            fl = ROWS ? 0u : __ldg(blk_flags + b);
        }
        for (unsigned j = 0; j < pro; ++j) {
            const unsigned long long a = __shfl_sync(0xffffffffu, r0, j), b2 = __shfl_sync(0xffffffffu, r1, j);
            const unsigned long long c = __shfl_sync(0xffffffffu, n0, j), d = __shfl_sync(0xffffffffu, n1, j);
            const unsigned f = __shfl_sync(0xffffffffu, fl, j);
            if (tid == 0) issue(j, a, b2, c, d, f);
        }
    }
    // thread 0 keeps the descriptor of the next block to issue in registers (its loads overlap one iteration)
    unsigned long long nd_r0 = 0, nd_r1 = 0, nd_n0 = 0, nd_n1 = 0;
    unsigned nd_fl = 0;
    unsigned issue_j = pro, issue_s = pro % stages;
    auto fetch_next = [&]() {
        if (issue_j < n_my) {
            const size_t b = blk(issue_j);
            nd_r0 = (unsigned long long)__ldg(blk_rows + b); nd_r1 = (unsigned long long)__ldg(blk_rows + b + 1);
            nd_n0 = (unsigned long long)__ldg(blk_nnz + b); nd_n1 = (unsigned long long)__ldg(blk_nnz + b + 1);
            nd_fl = ROWS ? 0u : __ldg(blk_flags + b);
        }
    };
    if (tid == 0) fetch_next();
    __syncthreads();          // the prologue's descriptors are visible to everybody

    // Row offsets of the K rows this thread sums per block, fetched ONE BLOCK AHEAD (raw values; they are only
    // touched after the next barrier, so their DRAM latency overlaps a whole iteration).
    I nxa[K], nxe[K];
    auto fetch_offsets = [&](unsigned st) {
        const uint64_t q0 = s_desc[st][0], q1 = s_desc[st][1];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const uint64_t r = q0 + tid + (uint64_t)j * THREADS;
            nxa[j] = nxe[j] = 0;
            if (r < q1) { nxa[j] = __ldg(offs + r); nxe[j] = __ldg(offs + r + 1); }
        }
    };
    if (n_my) fetch_offsets(0);

    double acc = 0.0;
    unsigned s = 0, parity = 0;
    for (unsigned i = 0; i < n_my; ++i) {
        mbar_wait(&full[s], parity);
        const uint64_t r0 = s_desc[s][0], r1 = s_desc[s][1], n0 = s_desc[s][2], n1 = s_desc[s][3];
        const unsigned fl = s_flags[s];
        const uint64_t a0 = n0 & ~(uint64_t)3;
        T* sv = reinterpret_cast<T*>(smem_raw + (size_t)s * stage_bytes);
        const I* sc = reinterpret_cast<const I*>(smem_raw + (size_t)s * stage_bytes + (size_t)cap * sizeof(T));
        I cua[K], cue[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { cua[j] = nxa[j]; cue[j] = nxe[j]; }
        if constexpr (ROWS) {
            // Short rows (stencils): one thread per row straight out of the staged slice.  Lanes of a warp hold
            // consecutive rows, so for a stencil the x gathers of one instruction fall into one or two cache lines,
            // and the slice is read with an odd word stride (no bank conflicts).  Storage-order sums: bit-exact.
            unsigned ka[K], ke[K];
            I c[K][CH];
            T v[K][CH], sum[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const bool live = r0 + tid + (uint64_t)j * THREADS < r1;
                ka[j] = live ? (unsigned)((uint64_t)cua[j] - a0) : 0u;
                ke[j] = live ? (unsigned)((uint64_t)cue[j] - a0) : 0u;
                sum[j] = T(0);
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (ka[j] + u < ke[j]) { c[j][u] = sc[ka[j] + u]; v[j][u] = sv[ka[j] + u]; }
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (ka[j] + u < ke[j]) v[j][u] = mul_rn(__ldg(x + (size_t)c[j][u]), v[j][u]);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) {
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (ka[j] + u < ke[j]) sum[j] = add_rn(sum[j], v[j][u]);
                for (unsigned k = ka[j] + CH; k < ke[j]; ++k) sum[j] = add_rn(sum[j], mul_rn(__ldg(x + (size_t)sc[k]), sv[k]));
                const uint64_t r = r0 + tid + (uint64_t)j * THREADS;
                if (r < r1) {
                    y[r] = sum[j];
                    if constexpr (DOT) acc += (double)mul_rn(__ldg((const T*)dot.w + r), sum[j]);
                }
            }
            for (uint64_t r = r0 + tid + (uint64_t)K * THREADS; r < r1; r += THREADS) {
                const unsigned a = (unsigned)((uint64_t)__ldg(offs + r) - a0), e = (unsigned)((uint64_t)__ldg(offs + r + 1) - a0);
                T s1 = T(0);
                for (unsigned k = a; k < e; ++k) s1 = add_rn(s1, mul_rn(__ldg(x + (size_t)sc[k]), sv[k]));
                y[r] = s1;
                if constexpr (DOT) acc += (double)mul_rn(__ldg((const T*)dot.w + r), s1);
            }
        } else if (!(fl & kBlkNoFit)) {
            const unsigned groups = (unsigned)(((n1 - a0) + 3) & ~(uint64_t)3) / G;
            for (unsigned g = tid; g < groups; g += THREADS * U) {
                Grp<I, G> c[U];
                Grp<T, G> v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned gi = g + u * THREADS;
                    if (gi < groups) { c[u] = lds_grp<I, G>(sc + (size_t)gi * G); v[u] = lds_grp<T, G>(sv + (size_t)gi * G); }
                }
                Grp<T, G> xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (g + u * THREADS < groups) {
#pragma unroll
                        for (int k = 0; k < G; ++k) xv[u].e[k] = __ldg(x + (size_t)c[u].e[k]);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned gi = g + u * THREADS;
                    if (gi < groups) {
#pragma unroll
                        for (int k = 0; k < G; ++k) v[u].e[k] = mul_rn(xv[u].e[k], v[u].e[k]);
                        sts_grp<T, G>(sv + (size_t)gi * G, v[u]);
                    }
                }
            }
            fence_proxy_async();      // the products were written through the generic proxy; the refill is async-proxy
        }
        __syncthreads();
        if (tid == 0 && issue_j < n_my) {
            // the stage consumed in iteration i - 1 is free now
            issue(issue_s, nd_r0, nd_r1, nd_n0, nd_n1, nd_fl);
            ++issue_j;
            if (++issue_s == stages) issue_s = 0;
            fetch_next();
        }
        // block i + 1 was published before the barrier above: request its row offsets now
        if (i + 1 < n_my) fetch_offsets(s + 1 == stages ? 0 : s + 1);
        if constexpr (ROWS) {
        } else if (fl & kBlkNoFit) {
            acc += rows_direct<T, I, DOT, THREADS>(vals, cols, offs, r0, r1, x, y, (const T*)dot.w, &s_long_count, s_long_rows, scratch);
            __syncthreads();
            if (tid == 0) s_long_count = 0;
        } else if (fl & kBlkLong) {
            acc += reduce_rows_from_smem<T, I, DOT, THREADS>(sv, offs, r0, r1, a0, y, (const T*)dot.w, &s_long_count, s_long_rows);
            __syncthreads();
            if (tid == 0) s_long_count = 0;
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const uint64_t r = r0 + tid + (uint64_t)j * THREADS;
                if (r < r1) {
                    T sum = T(0);
                    const unsigned ka = (unsigned)((uint64_t)cua[j] - a0), ke = (unsigned)((uint64_t)cue[j] - a0);
                    for (unsigned k = ka; k < ke; ++k) sum = add_rn(sum, sv[k]);
                    y[r] = sum;
                    if constexpr (DOT) acc += (double)mul_rn(__ldg((const T*)dot.w + r), sum);
                }
            }
            for (uint64_t r = r0 + tid + (uint64_t)K * THREADS; r < r1; r += THREADS) {
                const unsigned a = (unsigned)((uint64_t)__ldg(offs + r) - a0), e = (unsigned)((uint64_t)__ldg(offs + r + 1) - a0);
                T sum = T(0);
                for (unsigned k = a; k < e; ++k) sum = add_rn(sum, sv[k]);
                y[r] = sum;
                if constexpr (DOT) acc += (double)mul_rn(__ldg((const T*)dot.w + r), sum);
            }
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }
    if constexpr (DOT) finish_dot<T, THREADS>(acc, dot);
}

// ---- K4 persistent ring: everything staged by TMA -------------------------------------------------------------
// For matrices whose rows are short (<= kRowMajorMax entries: stencils, FEM).  One 512-thread CTA per SM walks the
// row blocks b = blockIdx.x, blockIdx.x + gridDim.x, ...  Thread 0 keeps `stages` blocks staged or in flight; one
// stage holds the block's slice of values and columns, its row offsets, up to kNSeg windows of x that cover every
// column the block touches (found once at plan time).  All of it arrives by cp.async.bulk on the stage's mbarrier, so
// the only synchronous memory operations left in the loop are shared-memory loads, the coalesced store of y and, for
// the fused dot, one coalesced load of the row's weight.  A producer warp (lane 0) fills the ring; the other 15 warps
// consume it without any block-wide barrier: a warp releases a stage (mbarrier `empty`) as soon as its own rows are done.  One thread per row sums in storage order (bit-exact vs the
// reference); lanes hold consecutive rows, so slice reads have an odd word stride and window reads are consecutive.
constexpr int kNSeg = 4;
constexpr int kSegWords = 4096;              // plan time: bitmap over 131072 column buckets
constexpr int kRingThreads = 512;

struct RingDesc {
    unsigned long long r0, r1, a0, r0a;
    unsigned long long lo[kNSeg];            // first column of each window (unused: all ones)
    long long delta[kNSeg];                  // window element index = column + delta
    unsigned xwin;                           // number of windows that cover the block; 0: gather x from global memory
    unsigned c16;                            // 1: the column area of the stage holds 16-bit indices into the staged windows
    unsigned o16;                            // 1: the offset area holds 16-bit row offsets relative to a0, entry 0 = row r0
};

template <class T, class I, int NSEG> struct RingWin {
    I lo1, lo2, lo3;
    unsigned d0, d1, d2, d3;          // window element index = (unsigned)column + d  (mod 2^32: windows are < 2^32 long)
    __device__ __forceinline__ unsigned at(I c) const {
        unsigned d = d0;
        if constexpr (NSEG > 1) { if (c >= lo1) d = d1; }
        if constexpr (NSEG > 2) { if (c >= lo2) d = d2; }
        if constexpr (NSEG > 3) { if (c >= lo3) d = d3; }
        return (unsigned)c + d;
    }
};

// Rows [r0, r1) of one staged block, one thread per row, starting at thread `lane_id` of `n_lanes` consumer threads.
// The stage's row offsets: full width (absolute, entry 0 = row r0a) or 16 bits (relative to a0, entry 0 = row r0).
template <class I, bool O16>
__device__ __forceinline__ void ring_row_span(const I* so, uint64_t r, uint64_t r0, uint64_t r0a, uint64_t a0, unsigned& ka, unsigned& ke) {
    if constexpr (O16) {
        const uint16_t* so16 = reinterpret_cast<const uint16_t*>(so);
        ka = so16[r - r0];
        ke = so16[r - r0 + 1];
    } else {
        ka = (unsigned)((uint64_t)so[r - r0a] - a0);
        ke = (unsigned)((uint64_t)so[r + 1 - r0a] - a0);
    }
}

template <class T, class I, bool DOT, int NSEG, bool O16 = false, bool DIST = false>
__device__ __forceinline__ double ring_rows(const T* sv, const I* sc, const I* so, const T* sx, const RingDesc& d,
                                            unsigned lane_id, unsigned n_lanes, const T* __restrict__ x, T* __restrict__ y,
                                            const T* __restrict__ w, const T* gx = nullptr, unsigned long long g0 = ~0ull) {
    const uint64_t r0 = d.r0, r1 = d.r1, a0 = d.a0, r0a = d.r0a;
    RingWin<T, I, NSEG> win;
    win.lo1 = (I)d.lo[1]; win.lo2 = (I)d.lo[2]; win.lo3 = (I)d.lo[3];
    win.d0 = (unsigned)d.delta[0]; win.d1 = (unsigned)d.delta[1]; win.d2 = (unsigned)d.delta[2]; win.d3 = (unsigned)d.delta[3];
    double acc = 0.0;
    for (uint64_t r = r0 + lane_id; r < r1; r += n_lanes) {
        unsigned ka, ke;
        ring_row_span<I, O16>(so, r, r0, r0a, a0, ka, ke);
        T wv = T(0);
        if constexpr (DOT) wv = __ldg(w + r);          // one coalesced load per row, in flight during the row's sum
        T sum = T(0);
#pragma unroll 4
        for (unsigned k = ka; k < ke; ++k) {
            const I c = sc[k];
            T xv;
            if constexpr (NSEG > 0) xv = sx[win.at(c)];
            else if constexpr (DIST) xv = (unsigned long long)c >= g0 ? __ldcg(gx + ((unsigned long long)c - g0)) : __ldg(x + (size_t)c);   // ghosts: written by peers during this launch
            else xv = __ldg(x + (size_t)c);
            sum = add_rn(sum, mul_rn(xv, sv[k]));
        }
        y[r] = sum;
        if constexpr (DOT) acc += (double)mul_rn(wv, sum);
    }
    return acc;
}

// Same with compressed columns: the plan stored, for every non-zero of a windowed block, the 16-bit position of its
// column inside the block's concatenated x windows, so a stage carries 2 bytes per column instead of sizeof(I) and the
// consumer needs no window search.  The arithmetic (operands, order, roundings) is unchanged.
// V8 (value indexing): the plan found at most 256 distinct values in every block (constant-coefficient stencils, unit-weight
// graphs) and stored one 8-bit code per non-zero plus the block's dictionary; the value area of the stage then holds the
// codes and `dict` the values themselves.  Same operands in the same order: results do not change by a bit.
template <class T, class I, bool DOT, bool O16, int LANES = 1, bool V8 = false>
__device__ __forceinline__ double ring_rows_c16(const T* sv, const uint16_t* sc, const I* so, const T* sx, const RingDesc& d,
                                                unsigned lane_id, unsigned n_lanes, T* __restrict__ y, const T* __restrict__ w,
                                                const T* dict = nullptr) {
    const uint64_t r0 = d.r0, r1 = d.r1, a0 = d.a0, r0a = d.r0a;
    double acc = 0.0;
    [[maybe_unused]] const uint8_t* vc = reinterpret_cast<const uint8_t*>(sv);
    if constexpr (LANES > 1) {
        // Rows of 33..256 entries (FEM with several unknowns per node): LANES consecutive threads share a row, lane l sums the
        // entries l, l + LANES, ... (consecutive shared-memory words across the group), a fixed xor-shuffle tree adds the
        // partial sums.  Deterministic, but not the reference's storage order: inside the north-star tolerance, not bit-exact.
        const unsigned sub = lane_id % LANES, grp = lane_id / LANES, n_grp = n_lanes / LANES;
        const uint64_t rows = r1 - r0;
        // every thread of a warp runs the same number of iterations (the shuffles below need the whole group converged)
        for (uint64_t i = grp; i < ((rows + n_grp - 1) / n_grp) * n_grp; i += n_grp) {
            const bool live = i < rows;
            const uint64_t r = r0 + (live ? i : 0);
            unsigned ka = 0, ke = 0;
            if (live) ring_row_span<I, O16>(so, r, r0, r0a, a0, ka, ke);
            T sum = T(0);
            for (unsigned k = ka + sub; k < ke; k += LANES) sum = add_rn(sum, mul_rn(sx[sc[k]], sv[k]));
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) sum = add_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
            if (live && sub == 0) {
                y[r] = sum;
                if constexpr (DOT) acc += (double)mul_rn(__ldg(w + r), sum);
            }
        }
    } else {
        for (uint64_t r = r0 + lane_id; r < r1; r += n_lanes) {
            unsigned ka, ke;
            ring_row_span<I, O16>(so, r, r0, r0a, a0, ka, ke);
            T wv = T(0);
            if constexpr (DOT) wv = __ldg(w + r);
            T sum = T(0);
#pragma unroll 4
            for (unsigned k = ka; k < ke; ++k) {
                if constexpr (V8) sum = add_rn(sum, mul_rn(sx[sc[k]], dict[vc[k]]));
                else sum = add_rn(sum, mul_rn(sx[sc[k]], sv[k]));
            }
            y[r] = sum;
            if constexpr (DOT) acc += (double)mul_rn(wv, sum);
        }
    }
    return acc;
}

// DIST = true: the launch is one distributed product (halo protocol of halo.cuh); false compiles all of it out.
struct NoHalo {};
template <bool DIST> struct HaloParam { using type = NoHalo; };
template <> struct HaloParam<true> { using type = HaloDev; };       // by value: its fields are read from the constant bank
// LANES = threads per row of the compressed-column path: 1 (rows <= 32 entries, storage-order sums, bit-exact) or 2 / 4 / 8.
// V8 = the stage's value area holds 8-bit codes into the block's dictionary (plan-time value indexing; LANES = 1 only).
template <class T, class I, bool DOT, bool DIST, int LANES, bool V8>
__global__ void __launch_bounds__(kRingThreads, 2)
spmv_ring_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                 const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz, const unsigned long long* __restrict__ seg_lo,
                 const unsigned* __restrict__ seg_len, unsigned n_blocks, unsigned cap, unsigned ocap, unsigned xcap,
                 unsigned colb, unsigned stages, int xwin_ok, const uint16_t* __restrict__ lcols, unsigned long long lcols_base,
                 const uint16_t* __restrict__ loffs, unsigned long long row_begin, const T* __restrict__ x, T* __restrict__ y,
                 DotArgs dot, const typename HaloParam<DIST>::type halo, unsigned rot, const uint8_t* __restrict__ vcodes,
                 const T* __restrict__ vdict) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kPipeMaxStages];    // producer -> consumers: the stage's bytes have landed
    __shared__ __align__(8) uint64_t empty[kPipeMaxStages];   // consumers -> producer: every consumer warp has left the stage
    __shared__ RingDesc s_desc[kPipeMaxStages];
    constexpr unsigned OA = 16 / sizeof(I);
    constexpr unsigned kConsumerWarps = kRingThreads / 32 - 1;
    const unsigned tid = threadIdx.x;
    const size_t o_cols = (size_t)cap * (V8 ? 1 : sizeof(T));  // V8: one code byte per entry (cap is a multiple of 16)
    const size_t o_offs = o_cols + (size_t)cap * colb;          // colb = 2: every block of the plan streams 16-bit columns
    const size_t o_x = o_offs + (size_t)(ocap + 8) * (loffs ? 2 : sizeof(I));   // loffs: ocap is a multiple of 8
    const size_t o_dict = o_x + (size_t)xcap * sizeof(T);
    const size_t stage_bytes = o_dict + (V8 ? 256 * sizeof(T) : 0);
    double stop = 0.0;
    if constexpr (DOT) { if (dot.done != nullptr) stop = __ldcg(dot.done); }
    if (stop != 0.0) return;
    const unsigned n_my = blockIdx.x < n_blocks ? (n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0)
        for (unsigned s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    // Distributed product (dist.cu, halo.cuh): this launch is epoch `epoch` of the halo protocol.  The ghost entries of x
    // live in half `epoch & 1` of the rank's ghost buffer, written by the neighbours' launches of the same epoch.
    unsigned long long epoch = 0, g0 = ~0ull;
    const T* gx = nullptr;
    if constexpr (DIST) {
        epoch = __ldcg(halo.epoch) + 1ull;            // one L2 round trip, first needed by the push / the first ghost block
        g0 = halo.g0;
        gx = (const T*)halo.ghost + (epoch & 1ull) * halo.ghost_stride;
    }
    [[maybe_unused]] bool waited = false;    // producer lane: the neighbours' flags of this epoch have been seen
    __syncthreads();
    double acc = 0.0;

    if (tid < 32) {
        // ---- producer warp: lane 0 runs ahead of the consumers, bounded only by the ring depth --------------------------
        if (tid == 0) {
            // The producer is ONE thread and every block costs it a round trip for the block's metadata plus a few hundred
            // instructions; with two stages it must stay well below the ~1.9 us the consumers need per block.  The distributed
            // instantiation (more work per block) requests the metadata of block j + 1 while block j is being issued.
            struct Meta { I r0, r1, n0, n1; unsigned long long lo[kNSeg]; unsigned len[kNSeg]; size_t b; };
            auto fetch = [&](unsigned j) {
                Meta m;
                // `rot` rotates the block order so that the rows that need ghost entries come last (their wait is then over
                // before it starts)
                size_t b = (size_t)blockIdx.x + (size_t)j * gridDim.x;
                if constexpr (DIST) { b += rot; if (b >= n_blocks) b -= n_blocks; }
                m.b = b;
                m.r0 = __ldg(blk_rows + b); m.r1 = __ldg(blk_rows + b + 1);
                m.n0 = __ldg(blk_nnz + b); m.n1 = __ldg(blk_nnz + b + 1);
#pragma unroll
                for (int i = 0; i < kNSeg; ++i) { m.lo[i] = __ldg(seg_lo + kNSeg * b + i); m.len[i] = __ldg(seg_len + kNSeg * b + i); }
                return m;
            };
            unsigned s = 0, parity = 0;
            [[maybe_unused]] Meta next;
            if constexpr (DIST) next = fetch(0);
            for (unsigned j = 0; j < n_my; ++j) {
                // (measured on the 256^3 f32 slab: the look-ahead takes the distributed launch from 136.5 to 133.3 us, but
                // costs the plain one 3 us — 131.2 instead of 128.2 —, so only the distributed instantiation uses it)
                Meta cur;
                if constexpr (DIST) { cur = next; if (j + 1 < n_my) next = fetch(j + 1); }
                else cur = fetch(j);
                const size_t b = cur.b;
                const unsigned long long r0 = (unsigned long long)cur.r0, r1 = (unsigned long long)cur.r1;
                const unsigned long long n0 = (unsigned long long)cur.n0, n1 = (unsigned long long)cur.n1;
                if (j >= stages) mbar_wait(&empty[s], parity ^ 1u);     // the consumers have drained this stage's previous block
                RingDesc& d = s_desc[s];
                unsigned char* base = smem_raw + (size_t)s * stage_bytes;
                // slices start at a multiple of 16 elements: 16-byte aligned for 1-byte value codes and 2-byte columns as well
                const unsigned long long a0 = n0 & ~(kSliceAlign - 1), r0a = r0 & ~(unsigned long long)(OA - 1);
                const unsigned count = (unsigned)(((n1 - a0) + (kSliceAlign - 1)) & ~(kSliceAlign - 1));
                // row offsets: full width from the CRS array, or the plan's 16-bit copy (block b's entries start at a multiple of 8)
                const bool o16 = loffs != nullptr;
                const unsigned ocount = o16 ? (unsigned)(((r1 - r0 + 1) + 7ull) & ~7ull)
                                            : (unsigned)(((r1 + 1 - r0a) + (OA - 1)) & ~(unsigned long long)(OA - 1));
                const unsigned obytes = ocount * (o16 ? 2u : (unsigned)sizeof(I));
                d.r0 = r0; d.r1 = r1; d.a0 = a0; d.r0a = r0a; d.o16 = o16 ? 1u : 0u;
                unsigned xtotal = 0, nseg = 0;
                [[maybe_unused]] bool need = false;     // DIST: the block reads ghost entries
#pragma unroll
                for (int i = 0; i < kNSeg; ++i) {
                    d.lo[i] = cur.len[i] ? cur.lo[i] : ~0ull;
                    d.delta[i] = (long long)xtotal - (long long)cur.lo[i];
                    xtotal += cur.len[i];
                    nseg += cur.len[i] ? 1u : 0u;
                    if constexpr (DIST) need = need || (cur.len[i] != 0u && cur.lo[i] + cur.len[i] > g0);
                }
                const bool xw = xwin_ok && xtotal > 0;
                const bool c16 = xw && lcols != nullptr;
                d.xwin = xw ? nseg : 0u;
                d.c16 = c16 ? 1u : 0u;
                const unsigned cbytes = count * (c16 ? 2u : (unsigned)sizeof(I));
                unsigned bytes = count * (unsigned)(V8 ? 1 : sizeof(T)) + cbytes + obytes;
                if (xw) bytes += xtotal * (unsigned)sizeof(T);
                if constexpr (V8) bytes += 256u * (unsigned)sizeof(T);
                if constexpr (DIST) {
                    // A block without windows gathers any column, a windowed one needs the ghosts if a window reaches past g0;
                    // and a product ends only after every neighbour's flag of its epoch was seen (halo.cuh, step 3): CTA 0
                    // waits at its last block at the latest.
                    if (!waited && (need || !xw || (blockIdx.x == 0 && j + 1 == n_my))) {
                        halo_wait(halo, epoch);
                        fence_proxy_async_global();
                        waited = true;
                    }
                }
                mbar_expect_tx(&full[s], bytes);
                if (count) {
                    if constexpr (V8) bulk_g2s(base, vcodes + (a0 - lcols_base), count, &full[s]);
                    else bulk_g2s(base, vals + a0, count * (unsigned)sizeof(T), &full[s]);
                    if (c16) bulk_g2s(base + o_cols, lcols + (a0 - lcols_base), cbytes, &full[s]);
                    else bulk_g2s(base + o_cols, cols + a0, cbytes, &full[s]);
                }
                if constexpr (V8) bulk_g2s(base + o_dict, vdict + 256 * b, 256u * (unsigned)sizeof(T), &full[s]);
                if (o16) bulk_g2s(base + o_offs, loffs + (((r0 - row_begin) + 8ull * b) & ~7ull), obytes, &full[s]);
                else bulk_g2s(base + o_offs, offs + r0a, obytes, &full[s]);
                if (xw) {
                    unsigned at = 0;
#pragma unroll
                    for (int i = 0; i < kNSeg; ++i)
                        if (cur.len[i]) {
                            unsigned char* dstw = base + o_x + (size_t)at * sizeof(T);
                            if constexpr (!DIST) {
                                bulk_g2s(dstw, x + cur.lo[i], cur.len[i] * (unsigned)sizeof(T), &full[s]);
                            } else {
                                // owned part of the window from x, ghost part from this epoch's half of the ghost buffer (a window
                                // across the end of the owned columns is two copies: g0 is a multiple of 64 elements)
                                const unsigned long long lo = cur.lo[i];
                                const unsigned n_own = lo >= g0 ? 0u : (unsigned)(g0 - lo < (unsigned long long)cur.len[i] ? g0 - lo : cur.len[i]);
                                if (n_own) bulk_g2s(dstw, x + lo, n_own * (unsigned)sizeof(T), &full[s]);
                                if (cur.len[i] > n_own)
                                    bulk_g2s(dstw + (size_t)n_own * sizeof(T), gx + (lo + n_own - g0), (cur.len[i] - n_own) * (unsigned)sizeof(T), &full[s]);
                            }
                            at += cur.len[i];
                        }
                }
                if (++s == stages) { s = 0; parity ^= 1u; }
            }
        }
    } else {
        // ---- consumer warps: no block-wide barrier; a warp releases the stage as soon as its own rows are done ------------
        const unsigned lane_id = tid - 32, n_lanes = kRingThreads - 32;
        if constexpr (DIST) {
            // x is final (stream order): while the first stages are still in flight, put the entries the neighbours need into
            // their ghost buffers; the CTA whose consumers finish last raises the neighbours' flags.  (Measured: letting the
            // 31 idle lanes of the producer warp do this instead slows lane 0 down — 0.145 vs 0.139 ms per product.)
            halo_push<T>(halo, x, epoch, (uint64_t)blockIdx.x * n_lanes + lane_id, (uint64_t)gridDim.x * n_lanes);
            asm volatile("bar.sync 1, %0;" ::"r"(n_lanes) : "memory");
            if (lane_id == 0) halo_arrive(halo, epoch, gridDim.x);
        }
        unsigned s = 0, parity = 0;
        for (unsigned i = 0; i < n_my; ++i) {
            mbar_wait(&full[s], parity);
            const RingDesc d = s_desc[s];
            const unsigned char* base = smem_raw + (size_t)s * stage_bytes;
            const T* sv = reinterpret_cast<const T*>(base);
            const I* sc = reinterpret_cast<const I*>(base + o_cols);
            const I* so = reinterpret_cast<const I*>(base + o_offs);
            const T* sx = reinterpret_cast<const T*>(base + o_x);
            const T* w = (const T*)dot.w;
            if (V8 || d.c16) {                    // (a V8 plan is packed: every block streams compressed columns)
                const uint16_t* sc16 = reinterpret_cast<const uint16_t*>(base + o_cols);
                const T* dict = reinterpret_cast<const T*>(base + o_dict);
                if (d.o16) acc += ring_rows_c16<T, I, DOT, true, LANES, V8>(sv, sc16, so, sx, d, lane_id, n_lanes, y, w, dict);
                else acc += ring_rows_c16<T, I, DOT, false, LANES, V8>(sv, sc16, so, sx, d, lane_id, n_lanes, y, w, dict);
            } else if constexpr (!V8) switch (d.xwin) {              // number of x windows of the block (block-uniform)
                case 0:
                    if (d.o16) acc += ring_rows<T, I, DOT, 0, true, DIST>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w, gx, g0);
                    else acc += ring_rows<T, I, DOT, 0, false, DIST>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w, gx, g0);
                    break;
                case 1: acc += ring_rows<T, I, DOT, 1>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w); break;
                case 2: acc += ring_rows<T, I, DOT, 2>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w); break;
                case 3: acc += ring_rows<T, I, DOT, 3>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w); break;
                default: acc += ring_rows<T, I, DOT, 4>(sv, sc, so, sx, d, lane_id, n_lanes, x, y, w); break;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            if (++s == stages) { s = 0; parity ^= 1u; }
        }
    }
    if constexpr (DIST) {
        // the CTA that finishes last closes the epoch and re-arms the arrival counters for the next launch
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(halo.ctr + 1, 1u) == gridDim.x - 1) {
                halo.ctr[1] = 0u;
                *halo.epoch = epoch;
            }
        }
    }
    if constexpr (DOT) finish_dot<T, kRingThreads>(acc, dot);
}

}  // namespace smb
#include "spmv_sell.cuh"
namespace smb {

// Plan time, one CTA per block: find <= kNSeg windows of x that cover every column of the block.  Columns are
// bucketed (2^shift columns per bucket) into a shared-memory bitmap, runs of set buckets (gaps of one bucket are
// bridged) become windows, a second pass takes the exact [min, max] of each; blocks whose windows do not fit
// `xcap` elements, or that need more than kNSeg, get no window and gather from global memory.
template <class I>
__global__ void __launch_bounds__(128)
block_segments_kernel(const I* __restrict__ cols, const I* __restrict__ blk_nnz, unsigned shift, unsigned xcap, unsigned xalign,
                      unsigned long long* __restrict__ seg_lo, unsigned* __restrict__ seg_len, unsigned long long* __restrict__ n_ok) {
    __shared__ unsigned bitmap[kSegWords];
    constexpr int kMaxRuns = 64;
    __shared__ unsigned long long s_lo[kNSeg], s_hi[kNSeg];
    __shared__ unsigned run_b0[kMaxRuns], run_b1[kMaxRuns];
    __shared__ int s_nruns;
    __shared__ unsigned s_bmin, s_bmax;
    const size_t b = blockIdx.x;
    const uint64_t n0 = (uint64_t)blk_nnz[b], n1 = (uint64_t)blk_nnz[b + 1];
    for (unsigned i = threadIdx.x; i < (unsigned)kSegWords; i += 128) bitmap[i] = 0u;
    if (threadIdx.x == 0) { s_bmin = 0xffffffffu; s_bmax = 0u; s_nruns = 0; }
    if (threadIdx.x < kNSeg) { s_lo[threadIdx.x] = ~0ull; s_hi[threadIdx.x] = 0ull; }
    __syncthreads();
    unsigned bmin = 0xffffffffu, bmax = 0u;
    for (uint64_t k = n0 + threadIdx.x; k < n1; k += 128) {
        const unsigned bk = (unsigned)((uint64_t)cols[k] >> shift);
        atomicOr(&bitmap[bk >> 5], 1u << (bk & 31));
        bmin = bk < bmin ? bk : bmin;
        bmax = bk > bmax ? bk : bmax;
    }
    if (n1 > n0) { atomicMin(&s_bmin, bmin); atomicMax(&s_bmax, bmax); }
    __syncthreads();
    if (threadIdx.x == 0 && n1 > n0) {
        int nr = 0;
        unsigned last = 0;
        for (unsigned wd = s_bmin >> 5; wd <= (s_bmax >> 5) && nr <= kMaxRuns; ++wd) {
            unsigned bits = bitmap[wd];
            while (bits) {
                const unsigned bk = wd * 32u + (unsigned)(__ffs((int)bits) - 1);
                bits &= bits - 1u;
                if (nr > 0 && bk <= last + 2u) run_b1[nr - 1] = bk;            // same run (a one-bucket gap is bridged)
                else if (nr == kMaxRuns) { nr = kMaxRuns + 1; break; }
                else { run_b0[nr] = bk; run_b1[nr] = bk; ++nr; }
                last = bk;
            }
        }
        // more runs than windows (several unknowns per node, wide stencils): close the narrowest gaps first; whether the
        // merged windows still fit the stage is decided on their exact bounds below
        while (nr > kNSeg && nr <= kMaxRuns) {
            int at = 0;
            unsigned best = 0xffffffffu;
            for (int i = 0; i + 1 < nr; ++i) { const unsigned gap = run_b0[i + 1] - run_b1[i]; if (gap < best) { best = gap; at = i; } }
            run_b1[at] = run_b1[at + 1];
            for (int i = at + 1; i + 1 < nr; ++i) { run_b0[i] = run_b0[i + 1]; run_b1[i] = run_b1[i + 1]; }
            --nr;
        }
        s_nruns = nr;
    }
    __syncthreads();
    const int nr = s_nruns;
    if (nr >= 1 && nr <= kNSeg) {
        for (uint64_t k = n0 + threadIdx.x; k < n1; k += 128) {
            const unsigned long long c = (unsigned long long)cols[k];
            const unsigned bk = (unsigned)(c >> shift);
            int i = 0;
            while (i + 1 < nr && bk > run_b1[i]) ++i;
            atomicMin(&s_lo[i], c);
            atomicMax(&s_hi[i], c + 1ull);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long lo[kNSeg];
        unsigned len[kNSeg], total = 0;
        bool ok = nr >= 1 && nr <= kNSeg;
        for (int i = 0; i < kNSeg; ++i) { lo[i] = ~0ull; len[i] = 0; }
        if (ok) {
            for (int i = 0; i < nr; ++i) {
                lo[i] = s_lo[i] & ~(unsigned long long)(xalign - 1);
                const unsigned long long l = (s_hi[i] - lo[i] + (xalign - 1)) & ~(unsigned long long)(xalign - 1);
                if (l > (unsigned long long)xcap) { ok = false; break; }
                len[i] = (unsigned)l;
                total += len[i];
            }
            if (total > xcap) ok = false;
        }
        for (int i = 0; i < kNSeg; ++i) { seg_lo[kNSeg * b + i] = ok ? lo[i] : ~0ull; seg_len[kNSeg * b + i] = ok ? len[i] : 0u; }
        if (ok) atomicAdd(n_ok, 1ull);
    }
}

// Plan time, one CTA per block: index compression.  For a block with windows, the position of every column inside the
// block's concatenated windows (the layout the producer stages them in) as a 16-bit number.
template <class I>
__global__ void __launch_bounds__(256)
ring_compress_kernel(const I* __restrict__ cols, const I* __restrict__ blk_nnz, const unsigned long long* __restrict__ seg_lo,
                     const unsigned* __restrict__ seg_len, unsigned long long lcols_base, uint16_t* __restrict__ lcols,
                     unsigned long long* __restrict__ n_c16) {
    const size_t b = blockIdx.x;
    unsigned long long lo[kNSeg];
    unsigned at[kNSeg], total = 0;
#pragma unroll
    for (int i = 0; i < kNSeg; ++i) {
        const unsigned len = seg_len[kNSeg * b + i];
        lo[i] = len ? seg_lo[kNSeg * b + i] : ~0ull;
        at[i] = total;
        total += len;
    }
    if (total == 0) return;                                   // no windows: the kernel reads the original columns
    const uint64_t n0 = (uint64_t)blk_nnz[b], n1 = (uint64_t)blk_nnz[b + 1];
    for (uint64_t k = n0 + threadIdx.x; k < n1; k += 256) {
        const unsigned long long c = (unsigned long long)cols[k];
        int i = 0;
#pragma unroll
        for (int j = 1; j < kNSeg; ++j) if (c >= lo[j]) i = j;
        lcols[k - lcols_base] = (uint16_t)((unsigned)(c - lo[i]) + at[i]);
    }
    if (threadIdx.x == 0) atomicAdd(n_c16, (unsigned long long)(n1 - n0));
}

// Plan time, one CTA per block: value indexing.  The block's distinct values (bit patterns) are collected in a shared-memory
// hash set; with at most 256 of them the sorted set becomes the block's dictionary and every non-zero gets its 8-bit code.
// A block with more distinct values raises *n_fail (the plan then keeps full-width values).  probe_only: count only.
template <class T, class I>
__global__ void __launch_bounds__(256)
ring_vdict_kernel(const T* __restrict__ vals, const I* __restrict__ blk_nnz, unsigned long long codes_base, uint8_t* __restrict__ codes,
                  T* __restrict__ dicts, unsigned long long* __restrict__ n_fail, int probe_only) {
    using U = typename std::conditional<sizeof(T) == 8, unsigned long long, unsigned>::type;
    constexpr unsigned kSlots = 1024;
    constexpr U kEmpty = ~(U)0;                       // (a NaN pattern; a value with exactly these bits is kept in s_has_empty)
    __shared__ U slots[kSlots];
    __shared__ U dict[256];
    __shared__ unsigned s_count, s_over, s_has_empty;
    const size_t b = blockIdx.x;
    const uint64_t n0 = (uint64_t)blk_nnz[b], n1 = (uint64_t)blk_nnz[b + 1];
    for (unsigned i = threadIdx.x; i < kSlots; i += 256) slots[i] = kEmpty;
    if (threadIdx.x == 0) { s_count = 0; s_over = 0; s_has_empty = 0; }
    __syncthreads();
    for (uint64_t k = n0 + threadIdx.x; k < n1 && !s_over; k += 256) {
        U bits;
        const T v = vals[k];
        memcpy(&bits, &v, sizeof bits);
        if (bits == kEmpty) { s_has_empty = 1; continue; }
        unsigned h = (unsigned)((bits * (U)0x9E3779B97F4A7C15ull) >> (sizeof(U) * 8 - 10));
        for (unsigned probe = 0; probe < kSlots; ++probe, h = (h + 1) & (kSlots - 1)) {
            const U old = atomicCAS(&slots[h], kEmpty, bits);
            if (old == bits) break;
            if (old == kEmpty) { if (atomicAdd(&s_count, 1u) >= 256u) s_over = 1; break; }
        }
    }
    __syncthreads();
    const unsigned distinct = s_count + s_has_empty;
    if (s_over || distinct > 256u) { if (threadIdx.x == 0) atomicAdd(n_fail, 1ull); return; }
    if (probe_only) return;
    // compact the set and sort it by bit pattern (rank by counting: <= 256 elements, one thread each)
    __shared__ U found[256];
    __shared__ unsigned s_nf;
    if (threadIdx.x == 0) s_nf = 0;
    __syncthreads();
    for (unsigned i = threadIdx.x; i < kSlots; i += 256)
        if (slots[i] != kEmpty) found[atomicAdd(&s_nf, 1u)] = slots[i];
    __syncthreads();
    if (threadIdx.x == 0 && s_has_empty) found[s_nf++] = kEmpty;
    __syncthreads();
    const unsigned nf = s_nf;
    dict[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < nf) {
        const U mine = found[threadIdx.x];
        unsigned rank = 0;
        for (unsigned i = 0; i < nf; ++i) rank += found[i] < mine ? 1u : 0u;      // distinct values: ranks are a permutation
        dict[rank] = mine;
    }
    __syncthreads();
    {
        T v;
        const U bits = dict[threadIdx.x];
        memcpy(&v, &bits, sizeof v);
        dicts[256 * b + threadIdx.x] = v;
    }
    for (uint64_t k = n0 + threadIdx.x; k < n1; k += 256) {
        U bits;
        const T v = vals[k];
        memcpy(&bits, &v, sizeof bits);
        unsigned lo = 0, hi = nf;                       // first index with dict[i] >= bits
        while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (dict[mid] < bits) lo = mid + 1; else hi = mid; }
        codes[k - codes_base] = (uint8_t)lo;
    }
}

// Plan time, one CTA per block: 16-bit row offsets relative to the start of the block's slice (a0 = blk_nnz[b] & ~7).
// Block b's rows + 1 entries start at ((r0 - row_begin) + 8 b) & ~7, a multiple of 8 entries (16 bytes) that never
// reaches into the previous block's entries.
template <class I>
__global__ void __launch_bounds__(128)
ring_offsets16_kernel(const I* __restrict__ offs, const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz,
                      unsigned long long row_begin, uint16_t* __restrict__ loffs) {
    const unsigned long long b = blockIdx.x;
    const unsigned long long r0 = (unsigned long long)blk_rows[b], r1 = (unsigned long long)blk_rows[b + 1];
    const unsigned long long a0 = (unsigned long long)blk_nnz[b] & ~(kSliceAlign - 1);
    const unsigned long long base = ((r0 - row_begin) + 8ull * b) & ~7ull;
    for (unsigned long long i = threadIdx.x; i <= r1 - r0; i += 128) loffs[base + i] = (uint16_t)((unsigned long long)offs[r0 + i] - a0);
}

// ---- plan construction --------------------------------------------------------------------------------
// Split points of the merge coordinate key(r) = (r - rb) + (offs[r] - offs[rb]) at multiples of `target`.
template <class I>
__global__ void split_rows_kernel(const I* __restrict__ offs, uint64_t rb, uint64_t re, uint64_t target, uint64_t row_weight,
                                  uint64_t n_blocks, I* __restrict__ blk_rows, I* __restrict__ blk_nnz) {
    const uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k > n_blocks) return;
    if (k == n_blocks) { blk_rows[k] = (I)re; blk_nnz[k] = offs[re]; return; }
    const uint64_t base = (uint64_t)offs[rb];
    const uint64_t want = k * target;
    uint64_t lo = rb, hi = re;          // smallest r in [rb,re] with key(r) >= want
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        const uint64_t key = (mid - rb) * row_weight + ((uint64_t)offs[mid] - base);
        if (key < want) lo = mid + 1; else hi = mid;
    }
    blk_rows[k] = (I)lo;
    blk_nnz[k] = offs[lo];
}

// PIPE: per block, does the slice fit a stage and does it hold a row longer than kWarpRowMin?
template <class I>
__global__ void block_flags_kernel(const I* __restrict__ offs, const I* __restrict__ blk_rows, uint64_t n_blocks, unsigned cap,
                                   unsigned char* __restrict__ flags) {
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const uint64_t r0 = (uint64_t)blk_rows[b], r1 = (uint64_t)blk_rows[b + 1];
    const uint64_t n0 = (uint64_t)offs[r0], n1 = (uint64_t)offs[r1];
    unsigned f = (n1 - (n0 & ~(uint64_t)3)) > (uint64_t)cap ? kBlkNoFit : 0u;
    uint64_t prev = n0, longest = 0;
    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t e = (uint64_t)offs[r + 1];
        longest = e - prev > longest ? e - prev : longest;
        if (longest > (uint64_t)kWarpRowMin) break;
        prev = e;
    }
    if (longest > (uint64_t)kWarpRowMin) f |= kBlkLong;
    else if (longest <= (uint64_t)kRowMajorMax && !(f & kBlkNoFit)) f |= kBlkShort;
    flags[b] = (unsigned char)f;
}

// Per block: [min column, max column + 1) over the block's own elements.
template <class I>
__global__ void __launch_bounds__(kSpmvThreads)
block_window_kernel(const I* __restrict__ cols, const I* __restrict__ offs, const I* __restrict__ blk_rows,
                    I* __restrict__ blk_win, unsigned long long* __restrict__ max_win) {
    __shared__ unsigned long long s_min[kSpmvThreads / 32], s_max[kSpmvThreads / 32];
    const uint64_t r0 = (uint64_t)blk_rows[blockIdx.x], r1 = (uint64_t)blk_rows[blockIdx.x + 1];
    const uint64_t n0 = (uint64_t)offs[r0], n1 = (uint64_t)offs[r1];
    unsigned long long mn = ~0ull, mx = 0ull;
    for (uint64_t k = n0 + threadIdx.x; k < n1; k += kSpmvThreads) {
        const unsigned long long c = (unsigned long long)cols[k];
        mn = c < mn ? c : mn;
        mx = c + 1 > mx ? c + 1 : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kSpmvThreads / 32; ++w) { mn = s_min[w] < mn ? s_min[w] : mn; mx = s_max[w] > mx ? s_max[w] : mx; }
        if (mx == 0) { mn = 0; }
        blk_win[2 * (size_t)blockIdx.x] = (I)mn;
        blk_win[2 * (size_t)blockIdx.x + 1] = (I)mx;
        atomicMax(max_win, mx - mn);
    }
}

void hostpipe_free(HostPipe& hp) {
    for (auto& p : hp.plans) plan_free(p);
    for (auto e : hp.ev_x) cudaEventDestroy(e);
    for (auto e : hp.ev_y) cudaEventDestroy(e);
    hp = HostPipe();
}

// [min column, max column + 1) over the elements of rows [r_lo, r_hi): which pieces of x a chunk needs
template <class I>
__global__ void col_range_kernel(const I* __restrict__ cols, const I* __restrict__ offs, uint64_t r_lo, uint64_t r_hi,
                                 unsigned long long* __restrict__ out2) {
    const uint64_t n0 = (uint64_t)offs[r_lo], n1 = (uint64_t)offs[r_hi];
    unsigned long long mn = ~0ull, mx = 0ull;
    for (uint64_t k = n0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n1; k += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long c = (unsigned long long)cols[k];
        mn = c < mn ? c : mn;
        mx = c + 1 > mx ? c + 1 : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(out2, mn); atomicMax(out2 + 1, mx); }
}

void plan_free(SpmvPlan& p) {
    if (p.blk_rows) cudaFree(p.blk_rows);
    if (p.blk_nnz) cudaFree(p.blk_nnz);
    if (p.blk_flags) cudaFree(p.blk_flags);
    if (p.seg_lo) cudaFree(p.seg_lo);
    if (p.seg_len) cudaFree(p.seg_len);
    if (p.lcols) cudaFree(p.lcols);
    if (p.loffs) cudaFree(p.loffs);
    if (p.vcodes) cudaFree(p.vcodes);
    if (p.vdict) cudaFree(p.vdict);
    if (p.sell_blocks) cudaFree(p.sell_blocks);
    if (p.sell_codes) cudaFree(p.sell_codes);
    if (p.sell_cols) cudaFree(p.sell_cols);
    if (p.sell_rowlen) cudaFree(p.sell_rowlen);
    if (p.sell_soff) cudaFree(p.sell_soff);
    if (p.blk_win) cudaFree(p.blk_win);
    for (smb200_crs* part : p.parts) smb200_crs_free(part);
    p = SpmvPlan();
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

static int pick_lanes(double mean_len) {
    if (mean_len <= 1.5) return 1;
    if (mean_len <= 3.0) return 2;
    if (mean_len <= 6.0) return 4;
    if (mean_len <= 12.0) return 8;
    if (mean_len <= 24.0) return 16;
    return 32;
}

// Staging capacity (elements) and merge target of the stream kernels.  Defaults keep 8 CTAs of 256
// threads resident per SM with room left for L1 (x gathers): f32 18 KB, f64 24 KB of products per CTA.
static void stream_shape(const smb200_crs* m, int variant, unsigned* cap, unsigned* target) {
    unsigned c, t;
    if (variant == SMB200_SPMV_STREAM) {
        c = m->vt == SMB200_F64 ? 3584u : 4608u;
    } else if (variant == SMB200_SPMV_RING) {
        // 36 KB of values + columns per stage (measured best on the 7-point Laplacian: f32/u32 4608, f64/u32 3072)
        const size_t per = vsize(m->vt) + isize(m->it);
        c = (unsigned)((36u * 1024u) / per) & ~3u;
        c = (unsigned)env_int("SMB200_RING_CAP", (int)c) & ~3u;
    } else if (variant == SMB200_SPMV_STREAM_PIPE) {
        // one ring stage: 32 KB of values + columns
        const size_t per = vsize(m->vt) + isize(m->it);
        c = (unsigned)((32u * 1024u) / per) & ~3u;
        c = (unsigned)env_int("SMB200_PIPE_CAP", (int)c) & ~3u;
    } else {
        // TMA variants stage values + columns: keep the footprint comparable
        const size_t per = vsize(m->vt) + isize(m->it);
        c = (unsigned)((36u * 1024u) / per) & ~3u;
    }
    c = (unsigned)env_int("SMB200_STREAM_CAP", (int)c) & ~3u;
    if (c < 256) c = 256;
    if (c > kMaxCap) c = kMaxCap;
    t = c - c / 9;                     // leave room for the row that straddles the target
    // rows are short: a block overshoots the target by < one row, and its slice is widened to multiples of 8 elements
    if (variant == SMB200_SPMV_RING) t = c - (unsigned)std::min<uint64_t>(std::max<uint64_t>(m->max_row_len, kRowMajorMax), c / 2) - 16;
    t = (unsigned)env_int("SMB200_STREAM_TARGET", (int)t);
    if (t + 8 > c) t = c - 8;
    *cap = c;
    *target = t;
}

struct PlanShape { unsigned cap = 0, target = 0, win_cap = 0; };
static PlanShape g_shape_of_plan(const smb200_crs* m, const SpmvPlan& p) {
    PlanShape s;
    s.cap = p.cap;
    s.target = p.target;
    if (p.variant == SMB200_SPMV_BANDED) {
        const unsigned xa = (unsigned)(16 / vsize(m->vt));
        s.win_cap = (unsigned)((p.max_win + 2 * xa) & ~(uint64_t)(xa - 1));
    }
    return s;
}

static smb200_status plan_build_range_impl(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                                           uint64_t rb, uint64_t re);

// AUTO: short-row matrices try the TMA ring first (it also wins on the L2-resident C1: 12.3 us warm / 24.0 us cold); it is kept when (almost) every block got its
// x windows (stencils, banded, FEM-like), otherwise the stream kernel — which gathers x through L1/L2 — is planned.
static smb200_status plan_build_range_auto(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                                           uint64_t rb, uint64_t re);

smb200_status plan_build_range(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                               uint64_t rb, uint64_t re) {
    const auto t0 = std::chrono::steady_clock::now();
    const smb200_status st = plan_build_range_auto(m, p, want_variant, want_lanes, flags, rb, re);
    if (st == SMB200_OK) {
        cudaStreamSynchronize(m->ctx->stream);
        p.build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return st;
}

// Device memory a plan holds beside the CRS arrays.
static uint64_t plan_device_bytes(const smb200_crs* m, const SpmvPlan& p) {
    const uint64_t is = isize(m->it);
    uint64_t b = 0;
    if (p.blk_rows) b += (p.n_blocks + 1) * is;
    if (p.blk_nnz) b += (p.n_blocks + 1) * is;
    if (p.blk_flags) b += p.n_blocks;
    if (p.seg_lo) b += (uint64_t)kNSeg * p.n_blocks * 8;
    if (p.seg_len) b += (uint64_t)kNSeg * p.n_blocks * 4;
    if (p.lcols) b += (p.n_c16 ? p.n_c16 : m->nnz) * 2 + kPadBytes;
    if (p.loffs) b += (p.n_o16 + 8 * p.n_blocks + 16) * 2 + kPadBytes;
    if (p.vcodes) b += p.n_v8 + kPadBytes;
    if (p.vdict) b += p.n_blocks * 256 * vsize(m->vt);
    if (p.sell_blocks) b += p.n_blocks * sizeof(SellBlock) + 3 * p.sell_entries + p.sell_rowbytes + 4 * p.sell_soffwords;
    if (p.blk_win) b += 2 * p.n_blocks * is;
    for (const smb200_crs* part : p.parts)
        b += part->nnz * (vsize(part->vt) + isize(part->it)) + (part->n_rows + 1) * isize(part->it) + plan_device_bytes(part, part->plan);
    return b;
}

static smb200_status plan_build_range_auto(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                                           uint64_t rb, uint64_t re) {
    int want = want_variant;
    if (want == SMB200_SPMV_AUTO) want = env_int("SMB200_SPMV_VARIANT", SMB200_SPMV_AUTO);
    if (want == SMB200_SPMV_AUTO && m->max_row_len <= (uint64_t)kRingRowMax && m->nnz > 0) {
        SMB_TRY(plan_build_range_impl(m, p, SMB200_SPMV_RING, want_lanes, flags, rb, re));
        // rows beyond 32 entries are only worth the ring when every block streams 16-bit columns (the multi-lane row sums
        // exist for that path only)
        const bool long_rows = m->max_row_len > (uint64_t)kRowMajorMax;
        if (p.variant == SMB200_SPMV_RING && (long_rows ? (p.colb == 2 && p.n_xwin == p.n_blocks) : p.n_xwin * 10 >= p.n_blocks * 8)) return SMB200_OK;
    }
    // x far larger than L2 and no column locality (the ring was not kept): column bands (bandsplit.cu).  The plan holds a
    // second copy of the matrix (12 instead of 16 bytes per f64/u64 entry), so it is only taken when that fits comfortably.
    if (want == SMB200_SPMV_AUTO && env_int("SMB200_BANDSPLIT_AUTO", 1) != 0 && rb == 0 && re == m->n_rows && m->x_extra == 0 &&
        m->n_cols * vsize(m->vt) > 2 * (uint64_t)m->ctx->l2_bytes && m->nnz >= 4 * m->n_rows) {
        size_t free_b = 0, total_b = 0;
        const uint64_t need = m->nnz * (vsize(m->vt) + 4) + 16 * (m->n_rows + 1) * 4;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (uint64_t)free_b > 2 * need)
            return plan_build_range_impl(m, p, SMB200_SPMV_BANDSPLIT, want_lanes, flags, rb, re);
    }
    return plan_build_range_impl(m, p, want_variant, want_lanes, flags, rb, re);
}

// Cut rows [rb, re) into blocks of `target` merge units (key = row_weight * rows + non-zeros): blk_rows / blk_nnz.
static smb200_status plan_cut(smb200_crs* m, SpmvPlan& p, uint64_t rb, uint64_t re, uint64_t ob, uint64_t oe, unsigned target,
                              uint64_t row_weight) {
    smb200_ctx* ctx = m->ctx;
    const size_t is = isize(m->it);
    if (p.blk_rows) { cudaFree(p.blk_rows); p.blk_rows = nullptr; }
    if (p.blk_nnz) { cudaFree(p.blk_nnz); p.blk_nnz = nullptr; }
    p.target = target;
    const uint64_t merge_len = (re - rb) * row_weight + (oe - ob);
    p.n_blocks = (merge_len + target - 1) / target;
    if (p.n_blocks == 0) p.n_blocks = 1;
    SMB_REQUIRE(p.n_blocks < 0x7FFFFFFFull, SMB200_ERR_UNSUPPORTED, "spmv: too many row blocks");
    SMB_CUDA(cudaMalloc(&p.blk_rows, (p.n_blocks + 1) * is));
    SMB_CUDA(cudaMalloc(&p.blk_nnz, (p.n_blocks + 1) * is));
    const unsigned g = (unsigned)((p.n_blocks + 1 + 255) / 256);
    if (m->it == SMB200_U64) split_rows_kernel<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)m->offsets, rb, re, target, row_weight, p.n_blocks, (uint64_t*)p.blk_rows, (uint64_t*)p.blk_nnz);
    else split_rows_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)m->offsets, rb, re, target, row_weight, p.n_blocks, (uint32_t*)p.blk_rows, (uint32_t*)p.blk_nnz);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

// RING: the x windows of every block (<= p.xcap elements in total, else the block gets none) -> seg_lo / seg_len / n_xwin.
static smb200_status ring_segments(smb200_crs* m, SpmvPlan& p) {
    smb200_ctx* ctx = m->ctx;
    if (p.seg_lo) { cudaFree(p.seg_lo); p.seg_lo = nullptr; }
    if (p.seg_len) { cudaFree(p.seg_len); p.seg_len = nullptr; }
    SMB_CUDA(cudaMalloc(&p.seg_lo, (size_t)kNSeg * p.n_blocks * sizeof(unsigned long long)));
    SMB_CUDA(cudaMalloc(&p.seg_len, (size_t)kNSeg * p.n_blocks * sizeof(unsigned)));
    unsigned long long* d_ok = nullptr;
    SMB_CUDA(cudaMalloc(&d_ok, sizeof(unsigned long long)));
    cudaMemsetAsync(d_ok, 0, sizeof(unsigned long long), ctx->stream);
    unsigned shift = 6;                                     // >= 64 columns per bucket, <= 131072 buckets
    while (((m->n_cols + ((1ull << shift) - 1)) >> shift) > (uint64_t)kSegWords * 32u) ++shift;
    const unsigned xalign = (unsigned)(16 / vsize(m->vt));
    if (m->nnz) {
        if (m->it == SMB200_U64) block_segments_kernel<uint64_t><<<(unsigned)p.n_blocks, 128, 0, ctx->stream>>>((const uint64_t*)m->columns, (const uint64_t*)p.blk_nnz, shift, p.xcap, xalign, p.seg_lo, p.seg_len, d_ok);
        else block_segments_kernel<uint32_t><<<(unsigned)p.n_blocks, 128, 0, ctx->stream>>>((const uint32_t*)m->columns, (const uint32_t*)p.blk_nnz, shift, p.xcap, xalign, p.seg_lo, p.seg_len, d_ok);
        count_launch();
    } else {
        cudaMemsetAsync(p.seg_len, 0, (size_t)kNSeg * p.n_blocks * sizeof(unsigned), ctx->stream);
        cudaMemsetAsync(p.seg_lo, 0xff, (size_t)kNSeg * p.n_blocks * sizeof(unsigned long long), ctx->stream);
    }
    unsigned long long h_ok = 0;
    cudaError_t e = cudaMemcpyAsync(&h_ok, d_ok, sizeof h_ok, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ok);
    SMB_CUDA(e);
    p.n_xwin = h_ok;
    return SMB200_OK;
}

// Largest slice (non-zeros, widened to multiples of 8 on both sides) and largest row count over the blocks of a cut.
template <class I>
__global__ void block_extent_kernel(const I* __restrict__ blk_rows, const I* __restrict__ blk_nnz, uint64_t n_blocks,
                                    unsigned long long* __restrict__ out2) {
    const uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long z = 0, r = 0;
    if (b < n_blocks) {
        const unsigned long long n0 = (unsigned long long)blk_nnz[b], n1 = (unsigned long long)blk_nnz[b + 1];
        z = ((n1 + (kSliceAlign - 1)) & ~(kSliceAlign - 1)) - (n0 & ~(kSliceAlign - 1));
        r = (unsigned long long)blk_rows[b + 1] - (unsigned long long)blk_rows[b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long z2 = __shfl_xor_sync(0xffffffffu, z, o), r2 = __shfl_xor_sync(0xffffffffu, r, o);
        z = z2 > z ? z2 : z;
        r = r2 > r ? r2 : r;
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(out2, z); atomicMax(out2 + 1, r); }
}

constexpr size_t kRingStageBudget = 55 * 1024;   // two CTAs/SM x two stages (+ static shared memory) inside 227 KB

// Value-indexed, packed ring plan -> sliced-ELLPACK stage order (spmv_sell.cuh).  Kept when the padding is below 20 % and the
// stage still fits; p.vcodes / p.lcols / p.loffs (CRS order) are then released.
template <class I>
static smb200_status ring_sell_build(smb200_crs* m, SpmvPlan& p) {
    smb200_ctx* ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const size_t ts = vsize(m->vt);
    const uint64_t n = p.n_blocks;
    std::vector<I> h_rows(n + 1);
    SMB_CUDA(cudaMemcpyAsync(h_rows.data(), p.blk_rows, (n + 1) * sizeof(I), cudaMemcpyDeviceToHost, st));
    SMB_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> slice_base(n + 1), rows_of(n);
    unsigned long long total_slices = 0;
    for (uint64_t b = 0; b < n; ++b) {
        rows_of[b] = (unsigned long long)h_rows[b + 1] - (unsigned long long)h_rows[b];
        if (rows_of[b] > 65536) return SMB200_OK;                      // (sell_fill_kernel's slice table)
        slice_base[b] = total_slices;
        total_slices += (rows_of[b] + 31) / 32;
    }
    slice_base[n] = total_slices;
    unsigned long long *d_sb = nullptr, *d_rows = nullptr, *d_sz = nullptr;
    unsigned* d_w = nullptr;
    auto cleanup = [&] { if (d_sb) cudaFree(d_sb); if (d_rows) cudaFree(d_rows); if (d_sz) cudaFree(d_sz); if (d_w) cudaFree(d_w); };
#define SELL_CUDA(expr) do { cudaError_t se__ = (expr); if (se__ != cudaSuccess) { cleanup(); SMB_CUDA(se__); } } while (0)
    SELL_CUDA(cudaMalloc(&d_sb, (n + 1) * 8));
    SELL_CUDA(cudaMalloc(&d_rows, n * 8));
    SELL_CUDA(cudaMalloc(&d_sz, 3 * n * 8));
    SELL_CUDA(cudaMalloc(&d_w, (total_slices + 1) * 4));
    SELL_CUDA(cudaMemcpyAsync(d_sb, slice_base.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
    SELL_CUDA(cudaMemcpyAsync(d_rows, rows_of.data(), n * 8, cudaMemcpyHostToDevice, st));
    sell_width_kernel<I><<<(unsigned)n, 256, 0, st>>>((const I*)m->offsets, (const I*)p.blk_rows, d_sb, n, d_w);
    sell_block_sizes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_w, d_sb, d_rows, n, d_sz, d_sz + n, d_sz + 2 * n);
    count_launch(2);
    std::vector<unsigned long long> sz(3 * n);
    SELL_CUDA(cudaMemcpyAsync(sz.data(), d_sz, 3 * n * 8, cudaMemcpyDeviceToHost, st));
    SELL_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> bases(3 * n);
    unsigned long long E = 0, R = 0, S = 0, ecap = 0, rcap = 0, scap = 0;
    for (uint64_t b = 0; b < n; ++b) {
        bases[b] = E; bases[n + b] = R; bases[2 * n + b] = S;
        E += sz[b]; R += sz[n + b]; S += sz[2 * n + b];
        ecap = std::max(ecap, sz[b]); rcap = std::max(rcap, sz[n + b]); scap = std::max(scap, sz[2 * n + b]);
    }
    const size_t stage = 3 * (size_t)ecap + (size_t)rcap + 4 * (size_t)scap + (size_t)p.xcap * ts + 256 * ts;
    const bool keep = (double)E <= 1.2 * (double)p.n_v8 + 4096.0 && (stage * 2 + 3072) * 2 <= 227u * 1024u && ecap < (1ull << 31);
    if (!keep) { cleanup(); return SMB200_OK; }
    SELL_CUDA(cudaMemcpyAsync(d_sz, bases.data(), 3 * n * 8, cudaMemcpyHostToDevice, st));
    SELL_CUDA(cudaMalloc(&p.sell_blocks, n * sizeof(SellBlock)));
    SELL_CUDA(cudaMalloc(&p.sell_codes, E + kPadBytes));
    SELL_CUDA(cudaMalloc(&p.sell_cols, 2 * E + kPadBytes));
    SELL_CUDA(cudaMalloc(&p.sell_rowlen, R + kPadBytes));
    SELL_CUDA(cudaMalloc(&p.sell_soff, 4 * S + kPadBytes));
    cudaMemsetAsync(p.sell_codes, 0, E + kPadBytes, st);
    cudaMemsetAsync(p.sell_cols, 0, 2 * E + kPadBytes, st);
    cudaMemsetAsync(p.sell_rowlen, 0, R + kPadBytes, st);
    cudaMemsetAsync(p.sell_soff, 0, 4 * S + kPadBytes, st);
    const unsigned long long g0 = m->x_extra ? ((m->n_rows + 63) & ~(uint64_t)63) : ~0ull;     // dist.cu: ghosts from n_rows rounded up to 64
    sell_fill_kernel<I><<<(unsigned)n, 256, 0, st>>>((const I*)m->offsets, (const I*)p.blk_rows, d_w, d_sb, d_sz, d_sz + n, d_sz + 2 * n,
                                                    p.vcodes, p.lcols, p.lcols_base, (unsigned)ts, p.seg_lo, p.seg_len, g0, p.sell_codes, p.sell_cols,
                                                    p.sell_rowlen, p.sell_soff, (SellBlock*)p.sell_blocks);
    count_launch();
    SELL_CUDA(cudaGetLastError());
    SELL_CUDA(cudaStreamSynchronize(st));
#undef SELL_CUDA
    cleanup();
    p.sell_entries = E; p.sell_rowbytes = R; p.sell_soffwords = S;
    p.sell_ecap = (unsigned)ecap; p.sell_rcap = (unsigned)rcap; p.sell_scap = (unsigned)scap;
    // the CRS-order compressed arrays are not read any more
    cudaFree(p.vcodes); p.vcodes = nullptr;
    cudaFree(p.lcols); p.lcols = nullptr;
    if (p.loffs) { cudaFree(p.loffs); p.loffs = nullptr; p.n_o16 = 0; }      // (a length byte per row instead)
    return SMB200_OK;
}

// RING plan.  A stage holds [values cap*T][columns cap*colb][row offsets (ocap+8)*I][x windows xcap*T].
//  (a) packed: every block windowed, columns staged as 16-bit window positions (colb = 2), capacities taken from the
//      measured extents of the cut, the merge target grown until a stage is full — the bytes in flight per SM stay what
//      they were with full-width columns although every non-zero now costs sizeof(I) - 2 bytes less;
//  (b) conservative: worst-case capacities, full-width column area; blocks that did get windows still stream 16-bit
//      columns (mixed), the others gather x from global memory.
static smb200_status ring_plan(smb200_crs* m, SpmvPlan& p, uint64_t rb, uint64_t re, uint64_t ob, uint64_t oe, unsigned cap_b,
                               unsigned target_b) {
    smb200_ctx* ctx = m->ctx;
    const size_t ts = vsize(m->vt), is = isize(m->it);
    const uint64_t rows = re - rb, nnz = oe - ob;
    const bool want_c16 = env_int("SMB200_RING_C16", 1) != 0;
    const bool want_o16 = env_int("SMB200_RING_O16", 1) != 0;
    bool packed = false;
    // (a) with `vb` bytes per value in a stage (sizeof T, or 1 with value indexing) and `extra` more bytes per stage (the dictionary)
    auto try_packed = [&](size_t vb, size_t extra) -> smb200_status {
        packed = false;
        const size_t budget = kRingStageBudget - extra;
        const double mean = (double)nnz / (double)rows;
        const double bytes_per_key = (mean * (double)(vb + 2) + (double)is) / (mean + 2.0);
        double t = (double)env_int("SMB200_RING_FILL", 70) * 0.01 * (double)budget / bytes_per_key;
        unsigned long long* d_ext = nullptr;
        SMB_CUDA(cudaMalloc(&d_ext, 2 * sizeof(unsigned long long)));
        for (int attempt = 0; attempt < 5 && !packed; ++attempt, t *= 0.9) {
            if (t > 60000.0) t = 60000.0;
            if (t < 512.0) break;
            smb200_status st = plan_cut(m, p, rb, re, ob, oe, (unsigned)t & ~7u, 2);
            if (st != SMB200_OK) { cudaFree(d_ext); return st; }
            cudaMemsetAsync(d_ext, 0, 2 * sizeof(unsigned long long), ctx->stream);
            const unsigned g = (unsigned)((p.n_blocks + 255) / 256);
            if (m->it == SMB200_U64) block_extent_kernel<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)p.blk_rows, (const uint64_t*)p.blk_nnz, p.n_blocks, d_ext);
            else block_extent_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)p.blk_rows, (const uint32_t*)p.blk_nnz, p.n_blocks, d_ext);
            count_launch();
            unsigned long long h_ext[2] = {0, 0};
            cudaError_t e = cudaMemcpyAsync(h_ext, d_ext, sizeof h_ext, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { cudaFree(d_ext); SMB_CUDA(e); }
            const size_t cap = (size_t)((h_ext[0] + (kSliceAlign - 1)) & ~(kSliceAlign - 1)), ocap = (size_t)((h_ext[1] + 7ull) & ~7ull);
            const size_t fixed = cap * (vb + 2) + (ocap + 8) * (want_o16 && cap + 16 < 65536 ? 2 : is);
            if (fixed + 64 * ts > budget) continue;
            size_t xc = (budget - fixed) / ts;
            if (xc > 65536) xc = 65536;
            p.cap = (unsigned)cap;
            p.ocap = (unsigned)ocap;
            p.xcap = (unsigned)xc & ~3u;
            st = ring_segments(m, p);
            if (st != SMB200_OK) { cudaFree(d_ext); return st; }
            packed = p.n_xwin == p.n_blocks;
        }
        cudaFree(d_ext);
        return SMB200_OK;
    };
    auto vdict_run = [&](int probe_only, unsigned n_blk, unsigned long long* h_fail) -> smb200_status {
        unsigned long long* d_fail = nullptr;
        SMB_CUDA(cudaMalloc(&d_fail, sizeof(unsigned long long)));
        cudaMemsetAsync(d_fail, 0, sizeof(unsigned long long), ctx->stream);
#define SMB_VDICT(T, I) ring_vdict_kernel<T, I><<<n_blk, 256, 0, ctx->stream>>>((const T*)m->values, (const I*)p.blk_nnz, p.lcols_base, p.vcodes, (T*)p.vdict, d_fail, probe_only)
        if (m->vt == SMB200_F64) { if (m->it == SMB200_U64) SMB_VDICT(double, uint64_t); else SMB_VDICT(double, uint32_t); }
        else { if (m->it == SMB200_U64) SMB_VDICT(float, uint64_t); else SMB_VDICT(float, uint32_t); }
#undef SMB_VDICT
        count_launch();
        cudaError_t e = cudaMemcpyAsync(h_fail, d_fail, sizeof *h_fail, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_fail);
        SMB_CUDA(e);
        return SMB200_OK;
    };
    const bool pack_ok = want_c16 && nnz > 0 && env_int("SMB200_RING_PACK", 1) != 0 && getenv("SMB200_RING_CAP") == nullptr;
    // Value indexing (whole-matrix plans of one-lane ring kernels): if the first block-sized stretch of values has few distinct
    // ones, cut the blocks for 1-byte values, then give every block its dictionary; any block with more than 256 distinct
    // values sends the plan back to full-width values.
    bool v8 = false;
    if (pack_ok && want_o16 && rb == 0 && re == m->n_rows && p.lanes <= 1 && env_int("SMB200_RING_V8", 1) != 0) {
        SMB_TRY(try_packed(1, 256 * ts));
        if (packed && (size_t)p.cap + 16 < 65536) {
            p.lcols_base = ob & ~(kSliceAlign - 1);
            unsigned long long fails = 0;
            SMB_TRY(vdict_run(1, (unsigned)std::min<uint64_t>(p.n_blocks, 64), &fails));      // probe: the first blocks only
            if (fails == 0) {
                const size_t n_c = (size_t)(oe - p.lcols_base) + kSliceAlign;
                SMB_CUDA(cudaMalloc(&p.vcodes, n_c + kPadBytes));
                SMB_CUDA(cudaMalloc(&p.vdict, p.n_blocks * 256 * ts));
                cudaMemsetAsync(p.vcodes, 0, n_c + kPadBytes, ctx->stream);
                SMB_TRY(vdict_run(0, (unsigned)p.n_blocks, &fails));
                v8 = fails == 0;
                if (!v8) { cudaFree(p.vcodes); cudaFree(p.vdict); p.vcodes = nullptr; p.vdict = nullptr; }
            }
        }
        if (!v8) packed = false;
        else p.n_v8 = nnz;
    }
    if (!packed && pack_ok) SMB_TRY(try_packed(ts, 0));
    p.colb = packed ? 2u : (unsigned)is;
    if (!packed) {
        p.cap = cap_b;
        SMB_TRY(plan_cut(m, p, rb, re, ob, oe, target_b, 2));
        // worst case of a cut with key = 2*rows + nnz: target/2 rows
        p.ocap = (target_b / 2 + 8) & ~3u;
        const size_t fixed = (size_t)cap_b * (ts + is) + (size_t)(p.ocap + 8) * is;
        unsigned xc = kRingStageBudget > fixed ? (unsigned)((kRingStageBudget - fixed) / ts) : 0u;
        if (xc > cap_b) xc = cap_b;
        p.xcap = (unsigned)env_int("SMB200_RING_XCAP", (int)xc) & ~3u;
        SMB_TRY(ring_segments(m, p));
    }
    // index compression: 16-bit window positions for the windowed blocks (the windows of one block hold <= xcap <= 65536
    // elements).  Costs 2 bytes per non-zero of plan memory and saves sizeof(I) - 2 of every column read.
    if (p.n_xwin > 0 && p.xcap <= 65536u && want_c16) {
        p.lcols_base = ob & ~(kSliceAlign - 1);
        const size_t n_l = (size_t)(oe - p.lcols_base) + kSliceAlign;
        SMB_CUDA(cudaMalloc(&p.lcols, n_l * sizeof(uint16_t) + kPadBytes));
        unsigned long long* d_n = nullptr;
        SMB_CUDA(cudaMalloc(&d_n, sizeof(unsigned long long)));
        cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), ctx->stream);
        cudaMemsetAsync(p.lcols, 0, n_l * sizeof(uint16_t) + kPadBytes, ctx->stream);
        if (m->it == SMB200_U64) ring_compress_kernel<uint64_t><<<(unsigned)p.n_blocks, 256, 0, ctx->stream>>>((const uint64_t*)m->columns, (const uint64_t*)p.blk_nnz, p.seg_lo, p.seg_len, p.lcols_base, p.lcols, d_n);
        else ring_compress_kernel<uint32_t><<<(unsigned)p.n_blocks, 256, 0, ctx->stream>>>((const uint32_t*)m->columns, (const uint32_t*)p.blk_nnz, p.seg_lo, p.seg_len, p.lcols_base, p.lcols, d_n);
        count_launch();
        unsigned long long h_n = 0;
        cudaError_t e2 = cudaMemcpyAsync(&h_n, d_n, sizeof h_n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(ctx->stream);
        cudaFree(d_n);
        SMB_CUDA(e2);
        p.n_c16 = h_n;
    }
    // packed plans also stream 16-bit row offsets (relative to the block's slice): another sizeof(I) - 2 bytes per row
    if (packed && want_o16 && p.lcols && (size_t)p.cap + 16 < 65536) {
        const size_t n_o = (size_t)rows + 8 * (size_t)p.n_blocks + 16;
        SMB_CUDA(cudaMalloc(&p.loffs, n_o * sizeof(uint16_t) + kPadBytes));
        cudaMemsetAsync(p.loffs, 0, n_o * sizeof(uint16_t) + kPadBytes, ctx->stream);
        if (m->it == SMB200_U64) ring_offsets16_kernel<uint64_t><<<(unsigned)p.n_blocks, 128, 0, ctx->stream>>>((const uint64_t*)m->offsets, (const uint64_t*)p.blk_rows, (const uint64_t*)p.blk_nnz, rb, p.loffs);
        else ring_offsets16_kernel<uint32_t><<<(unsigned)p.n_blocks, 128, 0, ctx->stream>>>((const uint32_t*)m->offsets, (const uint32_t*)p.blk_rows, (const uint32_t*)p.blk_nnz, rb, p.loffs);
        count_launch();
        SMB_CUDA(cudaGetLastError());
        p.loffs_row_begin = rb;
        p.n_o16 = rows;
    }
    if (v8 && p.lcols && p.n_c16 == nnz && env_int("SMB200_RING_SELL", 1) != 0) {
        if (m->it == SMB200_U64) SMB_TRY(ring_sell_build<uint64_t>(m, p));
        else SMB_TRY(ring_sell_build<uint32_t>(m, p));
    }
    return SMB200_OK;
}

static smb200_status plan_build_range_impl(smb200_crs* m, SpmvPlan& p, int want_variant, int want_lanes, uint32_t flags,
                                           uint64_t rb, uint64_t re) {
    smb200_ctx* ctx = m->ctx;
    plan_free(p);
    p.flags = flags;
    const uint64_t rows = re - rb;
    const double mean = m->n_rows ? (double)m->nnz / (double)m->n_rows : 0.0;
    int variant = want_variant;
    if (variant == SMB200_SPMV_AUTO) variant = env_int("SMB200_SPMV_VARIANT", SMB200_SPMV_AUTO);
    if (variant == SMB200_SPMV_AUTO) {
        // Row-length statistics decide: short/medium rows -> stream (balanced, bit-exact row sums);
        // long regular rows -> a warp per row reads them with full coalescing and no staging.
        variant = (mean >= 96.0 && m->max_row_len < 8 * (uint64_t)mean + 4096) ? SMB200_SPMV_VECTOR : SMB200_SPMV_STREAM;
        // small short-row matrices that live in L2 (config C1): nothing to stream from HBM, the two-phase stream
        // kernel only adds latency; one thread per row wins there (12.4 us vs 18.4 us warm on the 1024^2 Laplacian)
        const uint64_t bytes = m->nnz * (vsize(m->vt) + isize(m->it)) + (m->n_rows + 1) * isize(m->it) + (m->n_cols + m->n_rows) * vsize(m->vt);
        if (bytes * 3 < (uint64_t)ctx->l2_bytes * 2 && m->max_row_len <= 16) variant = SMB200_SPMV_SCALAR;
    }
    // RING needs short rows everywhere (its stages have no long-row path)
    if (variant == SMB200_SPMV_RING && m->max_row_len > (uint64_t)kRingRowMax) variant = SMB200_SPMV_STREAM;
    if (variant == SMB200_SPMV_BANDSPLIT) {
        // whole-matrix products only; a matrix of a single band (or a row range, or a distributed block) is STREAM's
        if (rb == 0 && re == m->n_rows && m->x_extra == 0 && rows > 0) {
            SMB_TRY(bandsplit_build(m, p, SMB200_SPMV_STREAM));
            if (p.variant == SMB200_SPMV_BANDSPLIT) { p.built = true; return SMB200_OK; }
        }
        variant = SMB200_SPMV_STREAM;
    }
    p.variant = variant;
    p.lanes = 0;
    if (variant == SMB200_SPMV_SCALAR) p.lanes = 1;
    if (variant == SMB200_SPMV_RING) {
        // threads per row: 1 while a row fits 32 entries (storage-order sums), else the fewest that bring a lane's share to <= 32
        p.lanes = 1;
        while (p.lanes < 8 && m->max_row_len > (uint64_t)kRowMajorMax * (uint64_t)p.lanes) p.lanes *= 2;
        const int forced = env_int("SMB200_RING_LANES", 0);
        if (forced == 1 || forced == 2 || forced == 4 || forced == 8) p.lanes = forced;
    }
    if (variant == SMB200_SPMV_VECTOR) {
        int l = want_lanes > 0 ? want_lanes : env_int("SMB200_SPMV_LANES", 0);
        if (l <= 0) l = pick_lanes(mean);
        SMB_REQUIRE(l == 1 || l == 2 || l == 4 || l == 8 || l == 16 || l == 32, SMB200_ERR_INVALID,
                    "spmv: lanes must be a power of two in [1,32], got %d", l);
        p.lanes = l;
    }
    if (variant >= SMB200_SPMV_STREAM && rows > 0) {
        unsigned cap, target;
        stream_shape(m, variant, &cap, &target);
        p.cap = cap;
        p.target = target;
        // total merge length of the range, from the two boundary offsets
        uint64_t ob = 0, oe = 0;
        const size_t is = isize(m->it);
        if (m->it == SMB200_U64) {
            SMB_CUDA(cudaMemcpyAsync(&ob, (const char*)m->offsets + rb * is, 8, cudaMemcpyDeviceToHost, ctx->stream));
            SMB_CUDA(cudaMemcpyAsync(&oe, (const char*)m->offsets + re * is, 8, cudaMemcpyDeviceToHost, ctx->stream));
            SMB_CUDA(cudaStreamSynchronize(ctx->stream));
        } else {
            uint32_t b32 = 0, e32 = 0;
            SMB_CUDA(cudaMemcpyAsync(&b32, (const char*)m->offsets + rb * is, 4, cudaMemcpyDeviceToHost, ctx->stream));
            SMB_CUDA(cudaMemcpyAsync(&e32, (const char*)m->offsets + re * is, 4, cudaMemcpyDeviceToHost, ctx->stream));
            SMB_CUDA(cudaStreamSynchronize(ctx->stream));
            ob = b32; oe = e32;
        }
        if (variant == SMB200_SPMV_RING) {
            SMB_TRY(ring_plan(m, p, rb, re, ob, oe, cap, target));
        } else {
            SMB_TRY(plan_cut(m, p, rb, re, ob, oe, target, 1));
        }
        if (variant == SMB200_SPMV_STREAM_PIPE) {
            SMB_CUDA(cudaMalloc(&p.blk_flags, p.n_blocks));
            const unsigned gf = (unsigned)((p.n_blocks + 127) / 128);
            if (m->it == SMB200_U64) block_flags_kernel<uint64_t><<<gf, 128, 0, ctx->stream>>>((const uint64_t*)m->offsets, (const uint64_t*)p.blk_rows, p.n_blocks, cap, p.blk_flags);
            else block_flags_kernel<uint32_t><<<gf, 128, 0, ctx->stream>>>((const uint32_t*)m->offsets, (const uint32_t*)p.blk_rows, p.n_blocks, cap, p.blk_flags);
            count_launch();
            SMB_CUDA(cudaGetLastError());
        }
        if (variant == SMB200_SPMV_BANDED) {
            SMB_CUDA(cudaMalloc(&p.blk_win, 2 * p.n_blocks * is));
            unsigned long long* d_max = nullptr;
            SMB_CUDA(cudaMalloc(&d_max, sizeof(unsigned long long)));
            cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), ctx->stream);
            if (m->it == SMB200_U64) block_window_kernel<uint64_t><<<(unsigned)p.n_blocks, kSpmvThreads, 0, ctx->stream>>>((const uint64_t*)m->columns, (const uint64_t*)m->offsets, (const uint64_t*)p.blk_rows, (uint64_t*)p.blk_win, d_max);
            else block_window_kernel<uint32_t><<<(unsigned)p.n_blocks, kSpmvThreads, 0, ctx->stream>>>((const uint32_t*)m->columns, (const uint32_t*)m->offsets, (const uint32_t*)p.blk_rows, (uint32_t*)p.blk_win, d_max);
            count_launch();
            unsigned long long h_max = 0;
            cudaError_t e = cudaMemcpyAsync(&h_max, d_max, sizeof h_max, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_max);
            SMB_CUDA(e);
            p.max_win = h_max;
            // the window must fit next to the slice in shared memory; otherwise fall back to K3-TMA
            const size_t per = vsize(m->vt) + isize(m->it);
            const size_t need = (size_t)cap * per + (size_t)(h_max + 8) * vsize(m->vt) + 2048;
            if (need > 200u * 1024u) {
                p.variant = SMB200_SPMV_STREAM_TMA;
                cudaFree(p.blk_win);
                p.blk_win = nullptr;
            }
        }
    }
    p.built = true;
    return SMB200_OK;
}

smb200_status plan_build(smb200_crs* m) {
    return plan_build_range(m, m->plan, m->want_variant, m->want_lanes, m->want_flags, 0, m->n_rows);
}

// ---- launch ---------------------------------------------------------------------------------------------
static thread_local unsigned g_last_pipe_grid = 0;   // CTAs of the most recent persistent launch (= its dot partials)

template <class T, class I, bool DOT>
static smb200_status launch_typed(smb200_crs* m, const SpmvPlan& p, uint64_t rb, uint64_t re, const void* x, void* y,
                                  const DotArgs& dot) {
    smb200_ctx* ctx = m->ctx;
    const T* vals = (const T*)m->values;
    const I* cols = (const I*)m->columns;
    const I* offs = (const I*)m->offsets;
    const T* xx = (const T*)x;
    T* yy = (T*)y;
    cudaStream_t st = g_redirect.stream ? g_redirect.stream : ctx->stream;
    if (p.variant == SMB200_SPMV_SCALAR || p.variant == SMB200_SPMV_VECTOR) {
        const uint64_t rows = re - rb;
        const int rows_per_cta = kSpmvThreads / p.lanes;
        const uint64_t grid = (rows + rows_per_cta - 1) / rows_per_cta;
        SMB_REQUIRE(grid < 0x7FFFFFFFull, SMB200_ERR_UNSUPPORTED, "spmv: grid too large");
        switch (p.lanes) {
#define SMB_VEC_CASE(L) case L: spmv_vector_kernel<T, I, L, DOT><<<(unsigned)grid, kSpmvThreads, 0, st>>>(vals, cols, offs, rb, re, xx, yy, dot); break;
            SMB_VEC_CASE(1) SMB_VEC_CASE(2) SMB_VEC_CASE(4) SMB_VEC_CASE(8) SMB_VEC_CASE(16) SMB_VEC_CASE(32)
#undef SMB_VEC_CASE
            default: SMB_FAIL(SMB200_ERR_INVALID, "spmv: bad lane count %d", p.lanes);
        }
    } else {
        const PlanShape sh = g_shape_of_plan(m, p);
        const int carve = env_int("SMB200_CARVEOUT", -1);
        if (p.variant == SMB200_SPMV_STREAM) {
            const size_t smem = (size_t)(sh.cap + 8) * sizeof(T);
            auto kern = g_spmv_band ? spmv_band_kernel<T, I, DOT> : spmv_stream_kernel<T, I, DOT>;
            if (smem > 32 * 1024) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (carve >= 0) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            kern<<<(unsigned)p.n_blocks, kSpmvThreads, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, (const I*)p.blk_nnz, sh.cap, xx, yy, dot,
                                                                   g_spmv_accumulate ? 1 : 0);
        } else if (p.variant == SMB200_SPMV_RING) {
            // bulk copies need 16-byte aligned sources whose rounded-up tails stay inside the allocation
            const int xwin_ok = (((uintptr_t)x & 15u) == 0 && !g_x_unpadded && env_int("SMB200_RING_XWIN", 1) != 0) ? 1 : 0;
            // without windows (borrowed, unpadded x) every block stages its full-width columns and no x
            const unsigned colb = (xwin_ok && p.colb == 2) ? 2u : (unsigned)sizeof(I);
            const unsigned xcap = xwin_ok ? p.xcap : 0u;
            if (p.sell_blocks && xwin_ok) {
                // value-indexed plan in sliced-ELLPACK stage order (spmv_sell.cuh)
                const size_t stage = 3 * (size_t)p.sell_ecap + (size_t)p.sell_rcap + 4 * (size_t)p.sell_scap + (size_t)xcap * sizeof(T) + 256 * sizeof(T);
                int stages = env_int("SMB200_RING_STAGES", 2);
                if (stages < 2) stages = 2;
                if (stages > kPipeMaxStages) stages = kPipeMaxStages;
                while (stages > 2 && (stage * stages + 3072) * 2 > 227u * 1024u) --stages;
                const size_t smem = stage * stages;
                auto kern = spmv_ring_sell_kernel<T, DOT, false>;
                auto kern_d = spmv_ring_sell_kernel<T, DOT, true>;
                if (g_halo.host) SMB_CUDA(cudaFuncSetAttribute(kern_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                else SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int resident = 0;
                if (g_halo.host) SMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern_d, kRingThreads, smem));
                else SMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kRingThreads, smem));
                if (resident < 1) resident = 1;
                if (resident > 2) resident = 2;
                uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)resident;
                if (grid > p.n_blocks) grid = p.n_blocks;
                if (g_ring_grid_cap > 0 && grid > (uint64_t)g_ring_grid_cap) grid = (uint64_t)g_ring_grid_cap;
                g_last_pipe_grid = (unsigned)grid;
                if (g_halo.host)
                    kern_d<<<(unsigned)grid, kRingThreads, smem, st>>>((const SellBlock*)p.sell_blocks, p.sell_codes, p.sell_cols, p.sell_rowlen, p.sell_soff,
                                                                     (const T*)p.vdict, (unsigned)p.n_blocks, p.sell_ecap, p.sell_rcap, p.sell_scap, xcap,
                                                                     (unsigned)stages, xx, yy, dot, *g_halo.host, (unsigned)(g_halo.rot % p.n_blocks));
                else
                    kern<<<(unsigned)grid, kRingThreads, smem, st>>>((const SellBlock*)p.sell_blocks, p.sell_codes, p.sell_cols, p.sell_rowlen, p.sell_soff,
                                                                   (const T*)p.vdict, (unsigned)p.n_blocks, p.sell_ecap, p.sell_rcap, p.sell_scap, xcap,
                                                                   (unsigned)stages, xx, yy, dot, NoHalo(), 0u);
            } else {
            // value indexing needs the packed layout (every block windowed) and one lane per row
            const bool v8 = p.vcodes != nullptr && xwin_ok && colb == 2 && p.lanes <= 1;
            const size_t stage = (size_t)sh.cap * ((v8 ? 1 : sizeof(T)) + colb) + (size_t)(p.ocap + 8) * (p.loffs ? 2 : sizeof(I)) +
                                 (size_t)xcap * sizeof(T) + (v8 ? 256 * sizeof(T) : 0);
            // two CTAs per SM, two stages each: one block is consumed while the next one lands, and the second CTA's
            // consumers fill the issue slots the first one leaves idle
            int ctas = env_int("SMB200_RING_CTAS", 2);
            if (ctas < 1) ctas = 1;
            if (ctas > 2) ctas = 2;
            int stages = env_int("SMB200_RING_STAGES", ctas == 2 ? 2 : 4);
            if (stages < 2) stages = 2;
            if (stages > kPipeMaxStages) stages = kPipeMaxStages;
            while (stages > 2 && (stage * stages + 3072) * ctas > 227u * 1024u) --stages;
            if ((stage * stages + 3072) * ctas > 227u * 1024u) ctas = 1;
            const size_t smem = stage * stages;
            if (smem > 224u * 1024u) {
                // A plan cut for compressed stages, launched without x windows (a borrowed, unpadded x): the full-width slices do
                // not fit the ring.  Rare and not worth a second plan: straight from the CRS arrays, one thread per row where the
                // plan sums rows in storage order (so the result stays bit-identical), a sub-warp per row otherwise.
                const int l = p.lanes <= 1 ? 1 : pick_lanes(m->n_rows ? (double)m->nnz / (double)m->n_rows : 1.0);
                const uint64_t rows = re - rb, rows_per_cta = (uint64_t)(kSpmvThreads / l);
                const uint64_t grid = (rows + rows_per_cta - 1) / rows_per_cta;
                SMB_REQUIRE(grid < 0x7FFFFFFFull, SMB200_ERR_UNSUPPORTED, "spmv: grid too large");
                if constexpr (DOT) SMB_TRY(ensure_reduction_scratch(ctx, grid));
                DotArgs d2 = dot;
                if constexpr (DOT) { if (!g_redirect.partials) d2.partials = ctx->red_partials; }
                switch (l) {
#define SMB_VEC_CASE(L) case L: spmv_vector_kernel<T, I, L, DOT><<<(unsigned)grid, kSpmvThreads, 0, st>>>(vals, cols, offs, rb, re, xx, yy, d2); break;
                    SMB_VEC_CASE(1) SMB_VEC_CASE(2) SMB_VEC_CASE(4) SMB_VEC_CASE(8) SMB_VEC_CASE(16) SMB_VEC_CASE(32)
#undef SMB_VEC_CASE
                }
                g_last_pipe_grid = (unsigned)grid;
                count_launch();
                SMB_CUDA(cudaGetLastError());
                if constexpr (DOT) {
                    spmv_dot_finalize_kernel<<<1, kFinalizeThreads, 0, st>>>(d2.partials, g_last_pipe_grid, sizeof(T) == 4 ? 1 : 0, d2.result,
                                                                           d2.roll_dst, d2.roll_src, d2.done, d2.ar, g_cgsr);
                    count_launch();
                    SMB_CUDA(cudaGetLastError());
                }
                return SMB200_OK;
            }

            const int rl = p.lanes > 1 ? p.lanes : 1;        // threads per row (plan: from the longest row)
            auto kern = v8 ? spmv_ring_kernel<T, I, DOT, false, 1, true>
                      : rl == 8 ? spmv_ring_kernel<T, I, DOT, false, 8, false> : rl == 4 ? spmv_ring_kernel<T, I, DOT, false, 4, false>
                      : rl == 2 ? spmv_ring_kernel<T, I, DOT, false, 2, false> : spmv_ring_kernel<T, I, DOT, false, 1, false>;
            auto kern_d = v8 ? spmv_ring_kernel<T, I, DOT, true, 1, true>
                        : rl == 8 ? spmv_ring_kernel<T, I, DOT, true, 8, false> : rl == 4 ? spmv_ring_kernel<T, I, DOT, true, 4, false>
                        : rl == 2 ? spmv_ring_kernel<T, I, DOT, true, 2, false> : spmv_ring_kernel<T, I, DOT, true, 1, false>;
            if (g_halo.host) SMB_CUDA(cudaFuncSetAttribute(kern_d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            else SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int resident = 0;
            if (g_halo.host) SMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern_d, kRingThreads, smem));
            else SMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kRingThreads, smem));
            if (resident < 1) resident = 1;
            if (resident > ctas) resident = ctas;
            int sms = ctx->sm_count - g_ring_reserve_sms;
            if (sms < 1) sms = 1;
            uint64_t grid = (uint64_t)sms * (uint64_t)resident;
            if (grid > p.n_blocks) grid = p.n_blocks;
            if (g_ring_grid_cap > 0 && grid > (uint64_t)g_ring_grid_cap) grid = (uint64_t)g_ring_grid_cap;
            g_last_pipe_grid = (unsigned)grid;
            if (g_halo.host) {
                kern_d<<<(unsigned)grid, kRingThreads, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, (const I*)p.blk_nnz, p.seg_lo, p.seg_len,
                                                                 (unsigned)p.n_blocks, sh.cap, p.ocap, xcap, colb, (unsigned)stages, xwin_ok,
                                                                 xwin_ok ? p.lcols : nullptr, (unsigned long long)p.lcols_base, p.loffs,
                                                                 (unsigned long long)p.loffs_row_begin, xx, yy, dot, *g_halo.host,
                                                                 (unsigned)(g_halo.rot % p.n_blocks), p.vcodes, (const T*)p.vdict);
            } else {
                kern<<<(unsigned)grid, kRingThreads, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, (const I*)p.blk_nnz, p.seg_lo, p.seg_len,
                                                               (unsigned)p.n_blocks, sh.cap, p.ocap, xcap, colb, (unsigned)stages, xwin_ok,
                                                               xwin_ok ? p.lcols : nullptr, (unsigned long long)p.lcols_base, p.loffs,
                                                               (unsigned long long)p.loffs_row_begin, xx, yy, dot, NoHalo(), 0u, p.vcodes,
                                                               (const T*)p.vdict);
            }
            }
        } else if (p.variant == SMB200_SPMV_STREAM_PIPE) {
            int stages = env_int("SMB200_PIPE_STAGES", 3);
            if (stages < 3) stages = 3;     // the kernel reads the descriptor of block i + 1 during iteration i
            if (stages > kPipeMaxStages) stages = kPipeMaxStages;
            const size_t per_stage = (size_t)sh.cap * (sizeof(T) + sizeof(I));
            while (stages > 3 && per_stage * stages > 200u * 1024u) --stages;
            const size_t smem = per_stage * stages;
            SMB_REQUIRE(smem <= 227u * 1024u, SMB200_ERR_INVALID, "spmv: pipeline stage of %zu bytes does not fit shared memory", per_stage);
            int ctas = env_int("SMB200_PIPE_CTAS", 0);
            if (ctas <= 0) { ctas = (int)((220u * 1024u) / (smem + 2048)); if (ctas < 1) ctas = 1; if (ctas > 4) ctas = 4; }
            const int threads = env_int("SMB200_PIPE_THREADS", 256);
            uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)ctas;
            if (grid > p.n_blocks) grid = p.n_blocks;
            const bool rows_mode = m->max_row_len <= (uint64_t)kRowMajorMax && env_int("SMB200_PIPE_ROWS", 1) != 0;
#define SMB_PIPE_LAUNCH(TH, RW)                                                                                          \
    do {                                                                                                                 \
        auto kern = spmv_pipe_kernel<T, I, TH, DOT, RW>;                                                                 \
        SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        int resident = 0;                                                                                                \
        SMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, TH, smem));                              \
        if (resident < 1) resident = 1;                                                                                  \
        if ((uint64_t)resident * ctx->sm_count < grid) grid = (uint64_t)resident * ctx->sm_count;                        \
        g_last_pipe_grid = (unsigned)grid;                                                                               \
        kern<<<(unsigned)grid, TH, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, (const I*)p.blk_nnz, p.blk_flags, \
                                               (unsigned)p.n_blocks, sh.cap, (unsigned)stages, xx, yy, dot);            \
    } while (0)
            if (threads == 512) { if (rows_mode) SMB_PIPE_LAUNCH(512, true); else SMB_PIPE_LAUNCH(512, false); }
            else { if (rows_mode) SMB_PIPE_LAUNCH(256, true); else SMB_PIPE_LAUNCH(256, false); }
#undef SMB_PIPE_LAUNCH
        } else if (p.variant == SMB200_SPMV_STREAM_TMA) {
            const size_t smem = (size_t)sh.cap * (sizeof(T) + sizeof(I));
            auto kern = spmv_stream_tma_kernel<T, I, DOT, false>;
            if (smem > 32 * 1024) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (carve >= 0) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            kern<<<(unsigned)p.n_blocks, kSpmvThreads, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, nullptr, sh.cap, 0u, xx, yy, dot);
        } else {
            const size_t smem = (size_t)sh.cap * (sizeof(T) + sizeof(I)) + (size_t)sh.win_cap * sizeof(T);
            auto kern = spmv_stream_tma_kernel<T, I, DOT, true>;
            if (smem > 32 * 1024) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (carve >= 0) SMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            kern<<<(unsigned)p.n_blocks, kSpmvThreads, smem, st>>>(vals, cols, offs, (const I*)p.blk_rows, (const I*)p.blk_win, sh.cap, sh.win_cap, xx, yy, dot);
        }
    }
    count_launch();
    SMB_CUDA(cudaGetLastError());
    if constexpr (DOT) {
        unsigned n_partials;
        if (p.variant == SMB200_SPMV_SCALAR || p.variant == SMB200_SPMV_VECTOR) n_partials = (unsigned)(((re - rb) + (kSpmvThreads / p.lanes) - 1) / (kSpmvThreads / p.lanes));
        else if (p.variant == SMB200_SPMV_STREAM_PIPE || p.variant == SMB200_SPMV_RING) n_partials = g_last_pipe_grid;
        else n_partials = (unsigned)p.n_blocks;
        spmv_dot_finalize_kernel<<<1, kFinalizeThreads, 0, st>>>(dot.partials, n_partials, sizeof(T) == 4 ? 1 : 0, dot.result,
                                                               dot.roll_dst, dot.roll_src, dot.done, dot.ar, g_cgsr);
        count_launch();
        SMB_CUDA(cudaGetLastError());
    }
    return SMB200_OK;
}

static smb200_status set_l2_window(smb200_ctx* ctx, const void* base, size_t bytes, bool on) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (on && ctx->l2_persist_max > 0) {
        static thread_local int limit_set_for = -1;
        if (limit_set_for != ctx->device) {
            SMB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max));
            limit_set_for = ctx->device;
        }
        int max_win = 0;
        SMB_CUDA(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
        size_t win = bytes < (size_t)max_win ? bytes : (size_t)max_win;
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = win;
        double ratio = (double)ctx->l2_persist_max / (double)(win ? win : 1);
        attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitRatio = 0.f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    SMB_CUDA(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    return SMB200_OK;
}

static smb200_status spmv_launch_impl(smb200_crs* m, const SpmvPlan& p, uint64_t rb, uint64_t re, const void* x, void* y,
                                      const void* w, double* result, double* roll_dst, const double* roll_src,
                                      const double* done) {
    if (re <= rb) return SMB200_OK;
    if (p.variant == SMB200_SPMV_BANDSPLIT) {
        // y = sum_b A_b x_b: one launch per column band; the first writes y, the others add; the fused dot (and the CG
        // roll) ride the last one, which stores the final y
        const size_t es = vsize(m->vt);
        for (size_t b = 0; b < p.parts.size(); ++b) {
            smb200_crs* part = p.parts[b];
            const bool last = b + 1 == p.parts.size();
            g_spmv_accumulate = b > 0;
            g_spmv_band = env_int("SMB200_BAND_KERNEL", 1) != 0;
            const smb200_status st = spmv_launch_impl(part, part->plan, 0, part->n_rows, (const char*)x + b * (size_t)p.band_width * es, y,
                                                      last ? w : nullptr, result, last ? roll_dst : nullptr, roll_src, done);
            g_spmv_accumulate = false;
            g_spmv_band = false;
            SMB_TRY(st);
        }
        return SMB200_OK;
    }
    smb200_ctx* ctx = m->ctx;
    DotArgs dot{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (w) {
        uint64_t blocks = p.variant >= SMB200_SPMV_STREAM ? p.n_blocks : ((re - rb) * (uint64_t)p.lanes + kSpmvThreads - 1) / kSpmvThreads;
        double* partials = ctx->red_partials;
        if (g_redirect.partials) {
            SMB_REQUIRE(blocks <= ctx->red_cap_aux, SMB200_ERR_INVALID, "spmv: side-stream partials buffer too small");
            partials = g_redirect.partials;
        } else {
            SMB_TRY(ensure_reduction_scratch(ctx, blocks));
            partials = ctx->red_partials;
        }
        dot = DotArgs{w, partials, ctx->red_ticket, result, roll_dst, roll_src, done, g_dot_ar};
    }
    static thread_local const void* window_on = nullptr;
    if (p.flags & SMB200_FLAG_L2_PERSIST_X) {
        if (window_on != x) { SMB_TRY(set_l2_window(ctx, x, (size_t)m->n_cols * vsize(m->vt), true)); window_on = x; }
    } else if (window_on) {
        SMB_TRY(set_l2_window(ctx, nullptr, 0, false));
        window_on = nullptr;
    }
#define SMB_LAUNCH(T, I) (w ? launch_typed<T, I, true>(m, p, rb, re, x, y, dot) : launch_typed<T, I, false>(m, p, rb, re, x, y, dot))
    if (m->vt == SMB200_F64) return m->it == SMB200_U64 ? SMB_LAUNCH(double, uint64_t) : SMB_LAUNCH(double, uint32_t);
    return m->it == SMB200_U64 ? SMB_LAUNCH(float, uint64_t) : SMB_LAUNCH(float, uint32_t);
#undef SMB_LAUNCH
}

smb200_status spmv_launch_plan(smb200_crs* m, const SpmvPlan& p, uint64_t rb, uint64_t re, const void* x, void* y,
                               const void* w, int dot_slot) {
    return spmv_launch_impl(m, p, rb, re, x, y, w, m->ctx->red_result + dot_slot, nullptr, nullptr, nullptr);
}

// CG kernel A: ap = A p with p.Ap fused into partial slot `slot`; S is the solver's scalar block
// (cg.cu: S_RR=0, S_PAP=1..3, S_RR_NEW=4, S_DONE=7).  `roll` makes the launch also do rr <- rr_new.
smb200_status spmv_launch_cg(smb200_crs* m, const SpmvPlan& p, uint64_t rb, uint64_t re, const void* x, void* y,
                             const void* w, double* S, int slot, bool roll) {
    // single-reduction iterations (cg_sr.cuh) keep gamma in S[0] themselves: no roll
    return spmv_launch_impl(m, p, rb, re, x, y, w, S + 1 + slot, roll && !g_cgsr.active ? S + 0 : nullptr, S + 4, S + 7);
}

smb200_status spmv_launch(smb200_crs* m, const void* x, void* y, const void* w, int dot_slot, uint64_t, uint64_t) {
    if (m->n_rows == 0) return SMB200_OK;
    if (!m->plan.built) SMB_TRY(plan_build(m));
    return spmv_launch_plan(m, m->plan, 0, m->n_rows, x, y, w, dot_slot);
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_crs_configure(smb200_crs* m, smb200_spmv_variant variant, int32_t lanes, uint32_t flags) {
    SMB_REQUIRE(m, SMB200_ERR_INVALID, "crs_configure: NULL argument");
    SMB_REQUIRE(variant >= SMB200_SPMV_AUTO && variant <= SMB200_SPMV_BANDSPLIT, SMB200_ERR_INVALID, "crs_configure: bad variant %d", (int)variant);
    m->want_variant = variant;
    m->want_lanes = lanes;
    m->want_flags = flags;
    cudaStreamSynchronize(m->ctx->stream);
    plan_free(m->plan);
    hostpipe_free(m->hp);
    // a captured batch of CG iterations holds the old plan's arrays
    if (m->cg.graph) { cudaGraphExecDestroy(m->cg.graph); m->cg.graph = nullptr; }
    if (m->n_rows == 0) return SMB200_OK;
    return plan_build(m);
}

smb200_status smb200_crs_plan_info(const smb200_crs* cm, smb200_plan_info* out) {
    SMB_REQUIRE(cm && out, SMB200_ERR_INVALID, "crs_plan_info: NULL argument");
    smb200_crs* m = const_cast<smb200_crs*>(cm);
    if (!m->plan.built && m->n_rows) SMB_TRY(plan_build(m));
    memset(out, 0, sizeof *out);
    out->variant = m->plan.variant;
    out->lanes = m->plan.lanes;
    out->flags = m->plan.flags;
    out->n_blocks = m->plan.n_blocks;
    out->n_rows = m->n_rows;
    out->n_cols = m->n_cols;
    out->nnz = m->nnz;
    out->max_row_len = m->max_row_len;
    out->mean_row_len = m->n_rows ? (double)m->nnz / (double)m->n_rows : 0.0;
    out->algorithmic_bytes = m->nnz * (vsize(m->vt) + isize(m->it)) + (m->n_rows + 1) * isize(m->it) +
                             m->n_cols * vsize(m->vt) + m->n_rows * vsize(m->vt);
    out->launches_per_spmv = m->plan.variant == SMB200_SPMV_BANDSPLIT ? m->plan.parts.size() : 1;
    out->n_xwin_blocks = m->plan.n_xwin;
    out->nnz_c16 = m->plan.n_c16;
    out->rows_o16 = m->plan.n_o16;
    out->stream_bytes = out->algorithmic_bytes - (m->plan.n_c16 + m->plan.n_o16) * (isize(m->it) - 2);
    out->nnz_v8 = m->plan.n_v8;
    if (m->plan.n_v8) out->stream_bytes = out->stream_bytes - m->plan.n_v8 * (vsize(m->vt) - 1) + m->plan.n_blocks * 256 * vsize(m->vt);
    if (m->plan.sell_blocks)        // padded 3-byte entries, a length byte per row, an offset word per slice, the dictionaries, x and y
        out->stream_bytes = 3 * m->plan.sell_entries + m->plan.sell_rowbytes + 4 * m->plan.sell_soffwords + m->plan.n_blocks * (256 * vsize(m->vt) + sizeof(SellBlock)) +
                            (m->n_cols + m->n_rows) * vsize(m->vt);
    out->sell_entries = m->plan.sell_entries;
    if (m->plan.variant == SMB200_SPMV_BANDSPLIT) out->stream_bytes = bandsplit_stream_bytes(m, m->plan);
    out->plan_bytes = plan_device_bytes(m, m->plan);
    out->plan_ms = m->plan.build_ms;
    return SMB200_OK;
}

smb200_status smb200_spmv(smb200_crs* a, const smb200_vec* x, smb200_vec* y) {
    SMB_REQUIRE(a && x && y, SMB200_ERR_INVALID, "spmv: NULL argument");
    SMB_REQUIRE(x->vt == a->vt && y->vt == a->vt, SMB200_ERR_INVALID, "spmv: value types differ");
    SMB_REQUIRE(x->ctx == a->ctx && y->ctx == a->ctx, SMB200_ERR_INVALID, "spmv: operands belong to different contexts");
    SMB_REQUIRE(x->d != y->d || a->n_rows == 0, SMB200_ERR_INVALID, "spmv: x and y alias");
    // rhs.get(col) would panic for col >= x.dim (densevec.rs:40-42); n_cols = max col + 1
    SMB_REQUIRE(x->cap >= a->n_cols && x->n + a->x_extra >= a->n_cols, SMB200_ERR_DIM,
                "Dimension mismatch: x has %llu entries, matrix has %llu columns", (unsigned long long)x->n,
                (unsigned long long)(a->n_cols - a->x_extra));
    SMB_REQUIRE(y->n >= a->n_rows, SMB200_ERR_DIM, "Dimension mismatch: y has %llu entries, matrix has %llu rows",
                (unsigned long long)y->n, (unsigned long long)a->n_rows);
    g_x_unpadded = !x->owned;             // borrowed memory (smb200_vec_wrap) has no padding behind its last element
    const smb200_status st = spmv_launch(a, x->d, y->d, nullptr, 0);
    g_x_unpadded = false;
    return st;
}

smb200_status smb200_spmv_host(smb200_crs* a, const void* x_host, uint64_t nx, void* y_host) {
    SMB_REQUIRE(a && (x_host || nx == 0) && (y_host || a->n_rows == 0), SMB200_ERR_INVALID, "spmv_host: NULL argument");
    SMB_REQUIRE(nx >= a->n_cols, SMB200_ERR_DIM, "Dimension mismatch: x has %llu entries, matrix has %llu columns",
                (unsigned long long)nx, (unsigned long long)a->n_cols);
    smb200_ctx* ctx = a->ctx;
    const size_t es = vsize(a->vt);
    const size_t xb = (size_t)a->n_cols * es, yb = (size_t)a->n_rows * es;
    if (ctx->stage_x_bytes < xb) {
        if (ctx->stage_x) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->stage_x); ctx->stage_x = nullptr; ctx->stage_x_bytes = 0; }
        SMB_TRY(dev_alloc(&ctx->stage_x, xb));
        ctx->stage_x_bytes = xb;
    }
    if (ctx->stage_y_bytes < yb) {
        if (ctx->stage_y) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->stage_y); ctx->stage_y = nullptr; ctx->stage_y_bytes = 0; }
        SMB_TRY(dev_alloc(&ctx->stage_y, yb));
        ctx->stage_y_bytes = yb;
    }
    if (a->n_rows == 0) return SMB200_OK;
    // ---- pipelined path: H2D pieces of x | row chunks | D2H slices of y on three streams ---------------------------
    HostPipe& hp = a->hp;
    if (!hp.built) {
        // ~8 MiB of y per chunk (measured best on C2: every chunk costs ~14 us of launches, events and copy set-up).  The
        // upload of the first piece is the head of the pipeline (nothing else can run yet) and the download of the last slice
        // its tail, so the chunks at both ends are smaller: taper 1 = half size, taper 2 = 1/8, 1/4, 1/2 of a full chunk.
        int chunks = env_int("SMB200_HOST_CHUNKS", 0);
        if (chunks <= 0) chunks = (int)((yb + (8u << 20) - 1) / (8u << 20));
        if (chunks > 56) chunks = 56;
        if (chunks < 1 || env_int("SMB200_HOST_PIPE", 1) == 0) chunks = 1;
        const int taper = chunks >= 4 ? env_int("SMB200_HOST_TAPER", 2) : 0;   // measured on C2: 0: 1.86, 1: 1.78, 2: 1.75 ms/step
        std::vector<double> weights;
        if (taper == 2) {
            weights = {0.125, 0.25, 0.5};
            weights.insert(weights.end(), (size_t)chunks - 2, 1.0);
            weights.insert(weights.end(), {0.5, 0.25, 0.125});
        } else if (taper == 1) {
            weights.assign((size_t)chunks + 1, 1.0);
            weights.front() = weights.back() = 0.5;
        } else {
            weights.assign((size_t)chunks, 1.0);
        }
        chunks = (int)weights.size();
        double total_w = 0.0;
        for (double w : weights) total_w += w;
        hp.n_chunks = chunks;
        hp.row_bounds.assign(chunks + 1, a->n_rows);
        hp.x_bounds.assign(chunks + 1, a->n_cols);
        double at = 0.0;
        for (int c = 0; c < chunks; ++c) {
            hp.row_bounds[c] = ((uint64_t)((double)a->n_rows * (at / total_w)) + 1023) / 1024 * 1024;
            hp.x_bounds[c] = ((uint64_t)((double)a->n_cols * (at / total_w)) + 1023) / 1024 * 1024;
            if (hp.row_bounds[c] > a->n_rows) hp.row_bounds[c] = a->n_rows;
            if (hp.x_bounds[c] > a->n_cols) hp.x_bounds[c] = a->n_cols;
            at += weights[c];
        }
        hp.row_bounds[0] = 0; hp.x_bounds[0] = 0;
        hp.last_piece.assign(chunks, chunks - 1);
        hp.plans.resize(chunks);
        if (chunks > 1) {
            if (!ctx->copy_in) SMB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
            if (!ctx->copy_out) SMB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
            unsigned long long* d_rng = nullptr;
            std::vector<unsigned long long> h_rng(2 * (size_t)chunks);
            SMB_CUDA(cudaMalloc(&d_rng, h_rng.size() * sizeof(unsigned long long)));
            for (int c = 0; c < chunks; ++c) { h_rng[2 * c] = ~0ull; h_rng[2 * c + 1] = 0ull; }
            SMB_CUDA(cudaMemcpyAsync(d_rng, h_rng.data(), h_rng.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
            for (int c = 0; c < chunks; ++c) {
                if (hp.row_bounds[c + 1] <= hp.row_bounds[c] || a->nnz == 0) continue;
                if (a->it == SMB200_U64) col_range_kernel<uint64_t><<<ctx->sm_count, 256, 0, ctx->stream>>>((const uint64_t*)a->columns, (const uint64_t*)a->offsets, hp.row_bounds[c], hp.row_bounds[c + 1], d_rng + 2 * c);
                else col_range_kernel<uint32_t><<<ctx->sm_count, 256, 0, ctx->stream>>>((const uint32_t*)a->columns, (const uint32_t*)a->offsets, hp.row_bounds[c], hp.row_bounds[c + 1], d_rng + 2 * c);
                count_launch();
            }
            cudaError_t e = cudaMemcpyAsync(h_rng.data(), d_rng, h_rng.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_rng);
            SMB_CUDA(e);
            // Piece k of x ends where chunk k's columns end (rounded up, monotone), so chunk k waits for piece k only: for a
            // banded matrix that is its own rows plus the band, and the product starts one piece — not two — into the upload.
            // (Scattered columns: chunk 0 reaches the end of x, piece 0 is all of x, the later pieces are empty.)
            for (int c = 0; c < chunks; ++c) {
                unsigned long long mx = (h_rng[2 * c + 1] + 1023ull) / 1024ull * 1024ull;     // max column + 1 (0: no entries)
                if (mx > a->n_cols) mx = a->n_cols;
                hp.x_bounds[c + 1] = mx > hp.x_bounds[c] ? mx : hp.x_bounds[c];
                hp.last_piece[c] = c;
            }
            hp.x_bounds[chunks] = a->n_cols;       // columns nobody reads are uploaded too: stage_x is a complete copy of x
            hp.ev_x.resize(chunks); hp.ev_y.resize(chunks);
            for (int c = 0; c < chunks; ++c) {
                SMB_CUDA(cudaEventCreateWithFlags(&hp.ev_x[c], cudaEventDisableTiming));
                SMB_CUDA(cudaEventCreateWithFlags(&hp.ev_y[c], cudaEventDisableTiming));
                SMB_TRY(plan_build_range(a, hp.plans[c], a->want_variant, a->want_lanes, a->want_flags, hp.row_bounds[c], hp.row_bounds[c + 1]));
            }
        }
        hp.built = true;
    }
    if (hp.n_chunks <= 1) {
        if (xb) SMB_CUDA(cudaMemcpyAsync(ctx->stage_x, x_host, xb, cudaMemcpyHostToDevice, ctx->stream));
        SMB_TRY(spmv_launch(a, ctx->stage_x, ctx->stage_y, nullptr, 0));
        if (yb) SMB_CUDA(cudaMemcpyAsync(y_host, ctx->stage_y, yb, cudaMemcpyDeviceToHost, ctx->stream));
        SMB_CUDA(cudaStreamSynchronize(ctx->stream));
        return SMB200_OK;
    }
    for (int k = 0; k < hp.n_chunks; ++k) {
        const size_t lo = (size_t)hp.x_bounds[k] * es, hi = (size_t)hp.x_bounds[k + 1] * es;
        if (hi > lo) SMB_CUDA(cudaMemcpyAsync((char*)ctx->stage_x + lo, (const char*)x_host + lo, hi - lo, cudaMemcpyHostToDevice, ctx->copy_in));
        SMB_CUDA(cudaEventRecord(hp.ev_x[k], ctx->copy_in));
    }
    int waited = -1;
    for (int c = 0; c < hp.n_chunks; ++c) {
        const uint64_t rb = hp.row_bounds[c], re = hp.row_bounds[c + 1];
        if (hp.last_piece[c] > waited) { SMB_CUDA(cudaStreamWaitEvent(ctx->stream, hp.ev_x[hp.last_piece[c]], 0)); waited = hp.last_piece[c]; }
        if (re > rb) SMB_TRY(spmv_launch_plan(a, hp.plans[c], rb, re, ctx->stage_x, ctx->stage_y, nullptr, 0));
        SMB_CUDA(cudaEventRecord(hp.ev_y[c], ctx->stream));
        SMB_CUDA(cudaStreamWaitEvent(ctx->copy_out, hp.ev_y[c], 0));
        if (re > rb) SMB_CUDA(cudaMemcpyAsync((char*)y_host + (size_t)rb * es, (const char*)ctx->stage_y + (size_t)rb * es, (size_t)(re - rb) * es, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    SMB_CUDA(cudaStreamSynchronize(ctx->copy_out));
    SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SMB200_OK;
}

smb200_status smb200_bilinear(smb200_crs* a, const smb200_vec* lhs, const smb200_vec* rhs, double* out) {
    SMB_REQUIRE(a && lhs && rhs && out, SMB200_ERR_INVALID, "bilinear: NULL argument");
    SMB_REQUIRE(lhs->vt == a->vt && rhs->vt == a->vt, SMB200_ERR_INVALID, "bilinear: value types differ");
    SMB_REQUIRE(rhs->cap >= a->n_cols && rhs->n + a->x_extra >= a->n_cols, SMB200_ERR_DIM, "Dimension mismatch");
    SMB_REQUIRE(lhs->n >= a->n_rows, SMB200_ERR_DIM, "Dimension mismatch");
    if (a->n_rows == 0) { *out = 0.0; return SMB200_OK; }
    smb200_ctx* ctx = a->ctx;
    const size_t yb = (size_t)a->n_rows * vsize(a->vt);
    if (ctx->stage_y_bytes < yb) {
        if (ctx->stage_y) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->stage_y); ctx->stage_y = nullptr; ctx->stage_y_bytes = 0; }
        SMB_TRY(dev_alloc(&ctx->stage_y, yb));
        ctx->stage_y_bytes = yb;
    }
    g_x_unpadded = !rhs->owned;
    const smb200_status st = spmv_launch(a, rhs->d, ctx->stage_y, lhs->d, 1);
    g_x_unpadded = false;
    SMB_TRY(st);
    return fetch_result(ctx, 1, out);
}

}  // extern "C"
