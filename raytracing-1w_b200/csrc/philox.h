// philox.h — Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011), host + device.
// Counter-based: every (pixel, sample, bounce, purpose) tuple addresses its own 128-bit
// block, so no RNG state is stored in the ray queues.  Replaces the per-pixel sequential
// StdRng stream of main.rs:964 (parity with the reference is statistical by design).
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define RT1W_HD __host__ __device__ __forceinline__
#else
#define RT1W_HD inline
#endif

namespace rt1w {

struct Philox4 {
    uint32_t x, y, z, w;
};

RT1W_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = uint64_t(a) * uint64_t(b);
    lo = uint32_t(p), hi = uint32_t(p >> 32);
#endif
}

RT1W_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll 2 // five trips of two rounds: a fifth of the unrolled body for ~10 more instructions per call (the wave kernels are instruction-cache bound)
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// 24-bit uniform in [0,1): the f32 analogue of rand's `Standard` for floats.
RT1W_HD float u01(uint32_t x) { return float(x >> 8) * (1.0f / 16777216.0f); }

} // namespace rt1w
