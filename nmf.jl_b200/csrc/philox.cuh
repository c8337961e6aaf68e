// philox.cuh -- counter-based Philox4x32-10: one counter per element, so a draw depends neither on the launch geometry nor on
// how a matrix is sharded, and a host can regenerate it (tests/test_gpu_init.py does).  Used by nmfb200_randinit_* (api.cu) and by
// the Gaussian test matrix / random fill of the device NNDSVD (init_device.cuh).
#pragma once
#include <cstdint>

namespace nmfb200 {
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t& o0,
                                              uint32_t& o1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    o0 = c0;
    o1 = c1;
}
template <typename T> __device__ __forceinline__ T philox_uniform(uint64_t e, uint32_t stream, uint64_t seed);
template <> __device__ __forceinline__ float philox_uniform<float>(uint64_t e, uint32_t stream, uint64_t seed) {
    uint32_t a, b;
    philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), a, b);
    return (float)(a >> 8) * 5.9604644775390625e-08f;   // 24 bits -> [0, 1)
}
template <> __device__ __forceinline__ double philox_uniform<double>(uint64_t e, uint32_t stream, uint64_t seed) {
    uint32_t a, b;
    philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), a, b);
    return (double)((((uint64_t)a << 32) | b) >> 11) * 1.1102230246251565e-16;   // 53 bits -> [0, 1)
}
// Standard normal by Box-Muller from output words 0 and 1 of the same counter, evaluated in Float64 for both element types (so a
// NumPy host mirror reproduces it to the last bit or two): z = sqrt(-2 ln u1) cos(2 pi u2), u1 = (a + 0.5) 2^-32, u2 = b 2^-32.
__device__ __forceinline__ double philox_normal(uint64_t e, uint32_t stream, uint64_t seed) {
    uint32_t a, b;
    philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), a, b);
    const double u1 = ((double)a + 0.5) * 2.3283064365386963e-10, u2 = (double)b * 2.3283064365386963e-10;
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
}  // namespace nmfb200
