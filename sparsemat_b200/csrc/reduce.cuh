// Deterministic block / grid reductions shared by the dot, fused SpMV+dot and CG kernels.
//
// Cross-thread accumulation is always done in f64 and in a fixed order (given the launch shape), so
// results are reproducible run to run.  The grid stage is the "last block finishes" pattern: every
// CTA publishes its partial, takes a ticket, and the CTA that draws the last ticket folds all
// partials in index order.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace smb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; the result is valid in every thread.  `scratch` holds THREADS/32 + 1 doubles.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // scratch may still be in use by a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = (lane < WARPS) ? scratch[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

// Grid-wide sum of one value per CTA.  Returns true in every thread of the CTA that finished last;
// there `total` holds the sum of all CTAs' values, folded in CTA-index order.  The ticket counter is
// left at zero again for the next reduction on the stream.
template <int THREADS>
__device__ __forceinline__ bool grid_sum(double block_value, double* __restrict__ partials,
                                         unsigned int* __restrict__ ticket, double* scratch, double& total) {
    __shared__ bool s_last;
    const unsigned int nblocks = gridDim.x * gridDim.y;
    const unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        partials[bid] = block_value;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    double acc = 0.0;
    for (unsigned int i = threadIdx.x; i < nblocks; i += THREADS) acc += __ldcg(partials + i);
    total = block_sum<THREADS>(acc, scratch);
    if (threadIdx.x == 0) *ticket = 0u;
    return true;
}

// Exact (never contracted) arithmetic in the matrix/vector value type.
template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <class T> __device__ __forceinline__ T sub_rn(T a, T b);
template <> __device__ __forceinline__ float sub_rn<float>(float a, float b) { return __fsub_rn(a, b); }
template <> __device__ __forceinline__ double sub_rn<double>(double a, double b) { return __dsub_rn(a, b); }
template <class T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }

// 16-byte vector of T
template <class T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

template <class T> union Pack16 {
    typename Vec16<T>::type v;
    T e[Vec16<T>::N];
};

}  // namespace smb
