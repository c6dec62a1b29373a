// Counter-based generator for the synthetic workloads.  Integer and correctly
// rounded IEEE arithmetic only, so the host oracle can reproduce every array bit for bit.
#pragma once
#include <cstdint>

namespace smb {
namespace rng {

__host__ __device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rng1(uint64_t seed, uint64_t i) { return mix(mix(seed) + i); }
__host__ __device__ __forceinline__ uint64_t rng2(uint64_t seed, uint64_t i, uint64_t k) { return mix(rng1(seed, i) + k); }
__host__ __device__ __forceinline__ double u01(uint64_t bits) { return (double)(bits >> 11) * 0x1.0p-53; }
__host__ __device__ __forceinline__ double pm1(uint64_t bits) { return 2.0 * u01(bits) - 1.0; }

}  // namespace rng
}  // namespace smb
