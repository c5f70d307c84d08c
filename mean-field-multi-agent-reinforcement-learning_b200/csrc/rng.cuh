// Counter-based and reference-parity random number generators shared by the battle and Ising kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfmarl {

// Philox4x32-10 (Salmon et al. 2011), counter-based: the draw for (seed, env, step, i) needs no state.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// minstd_rand0: x <- 16807 x mod (2^31 - 1)   (libstdc++ std::default_random_engine, GridWorld.h:106)
__device__ __forceinline__ uint32_t minstd_next(uint32_t s) {
    return (uint32_t)(((uint64_t)s * 16807ull) % 2147483647ull);
}

}  // namespace mfmarl
