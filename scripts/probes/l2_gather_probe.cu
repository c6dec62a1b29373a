// Probe (B200): how large may a window of x be for random 8-byte gathers to hit L2 while a stream of evict-first data passes
// through the same L2?  Decides the band width of the band-split SpMV (bandsplit.cu).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2_gather_probe scripts/probes/l2_gather_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ double ld_hint(const double* p, uint64_t pol) {
    double v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p, uint64_t pol) {
    uint4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol)); return v;
}

// Every thread: per iteration one 16-byte streaming load (mode bit 1: evict_first hint) and G random gathers from the window
// (mode bit 0: evict_last hint).  stream_per_gather 16-byte loads per G gathers emulate the matrix stream (12 B / entry).
template <int G>
__global__ void probe(const double* __restrict__ x, uint64_t win, const uint4* __restrict__ stream, uint64_t n_stream16, int iters,
                      int mode, double* out) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t pl = make_policy_evict_last(), pf = make_policy_evict_first();
    double acc = 0.0;
    unsigned sacc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint64_t k = (uint64_t)it * nt + tid;
        if (n_stream16) {
#pragma unroll
            for (int q = 0; q < G; ++q) {      // 16 streamed bytes per gather, like the band-split product (12 B entry + y + offsets)
                const uint64_t at = ((uint64_t)it * G + q) * nt + tid;
                const uint4 s = (mode & 2) ? ld_stream(stream + (at % n_stream16), pf) : __ldg(stream + (at % n_stream16));
                sacc += s.x ^ s.y ^ s.z ^ s.w;
            }
        }
        double g[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint64_t c = mix(k * G + j) % win;
            g[j] = (mode & 1) ? ld_hint(x + c, pl) : __ldg(x + c);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) acc += g[j];
    }
    if (acc == 1.2345 || sacc == 77) out[0] = acc;
}

int main(int argc, char** argv) {
    const uint64_t stream_bytes = 4ull << 30;
    double* x; uint4* s; double* out;
    cudaMalloc(&x, 256ull << 20); cudaMalloc(&s, stream_bytes); cudaMalloc(&out, 8);
    cudaMemset(x, 0, 256ull << 20); cudaMemset(s, 0, stream_bytes);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, threads = 256, iters = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("window_MB mode(1=x evict_last,2=stream evict_first) stream  ms   gathers/s(G)  gather_GB/s(8B)  stream_GB/s\n");
    const int G = 4;
    for (int with_stream = 0; with_stream < 2; ++with_stream)
        for (int mode = 0; mode < 4; ++mode) {
            if (!with_stream && (mode & 2)) continue;
            for (uint64_t mb : {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 128, 192}) {
                const uint64_t win = (mb << 20) / 8;
                const uint64_t ns = with_stream ? stream_bytes / 16 : 0;
                for (int rep = 0; rep < 3; ++rep) {
                    if (rep == 1) cudaEventRecord(e0);
                    probe<G><<<grid, threads>>>(x, win, s, ns, iters, mode, out);
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 2;
                const double gathers = (double)grid * threads * iters * G;
                printf("%4llu %d %d %8.3f %10.2f %10.1f %10.1f\n", (unsigned long long)mb, mode, with_stream, ms, gathers / ms / 1e6,
                       gathers * 8 / ms / 1e6, with_stream ? (double)grid * threads * iters * 16 * G / ms / 1e6 : 0.0);
            }
        }
    return 0;
}
