// gcd_kernels.cuh -- GreedyCD per-row coordinate kernels (greedycd.jl:125-165) and the un-fused arithmetic
// helpers, shared by the exact (SIMT) engine and the tensor-core engine.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace nmfb200 {
namespace {

// ---- un-fused arithmetic helpers (the reference's scalar loops are not FMA-contracted) ----------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
template <typename T> __device__ __forceinline__ T eps_of();
template <> __device__ __forceinline__ float eps_of<float>() { return 1.1920928955078125e-07f; }
template <> __device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }

template <typename T>
__device__ __forceinline__ void gcd_sd(T w, T g, T prr, T& s, T& d) {
    // S = max(0, W - G/(eps(T)+P[r,r])) - W ; D = -G*S - 0.5*P[r,r]*S^2   (greedycd.jl:127-128, :155-156)
    T t = sub_rn(w, div_rn(g, add_rn(eps_of<T>(), prr)));
    s = sub_rn(t > T(0) ? t : T(0), w);
    d = sub_rn(mul_rn(-g, s), mul_rn(mul_rn(T(0.5), prr), mul_rn(s, s)));
}

// warp-level (value, index) arg-max with first-max tie-break
template <typename T>
__device__ __forceinline__ void warp_argmax(T& v, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

constexpr int GCD_WARPS = 4;

// pass 1: per-row max_r D[i,r]  -> per-block max (greedycd.jl:132-137)
template <typename T>
__global__ void __launch_bounds__(GCD_WARPS * 32) gcd_rowmax_kernel(const T* __restrict__ F, int64_t sFr, int64_t sFc,
                                                                   const T* __restrict__ G, const T* __restrict__ P,
                                                                   int64_t rows, int k, T* __restrict__ blockmax) {
    __shared__ T wmax[GCD_WARPS];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    int64_t i = (int64_t)blockIdx.x * GCD_WARPS + warp;
    T best = T(-1.0);
    if (i < rows) {
        int bi = 0x7fffffff;
        T bv = (T)(-INFINITY);
        for (int r = lane; r < k; r += 32) {
            T s, d;
            gcd_sd(F[i * sFr + r * sFc], G[i * k + r], P[r + (int64_t)r * k], s, d);
            if (d > bv) { bv = d; bi = r; }
        }
        warp_argmax(bv, bi);
        best = bv;
    }
    if (lane == 0) wmax[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        T m = T(-1.0);
        for (int w = 0; w < GCD_WARPS; ++w) m = wmax[w] > m ? wmax[w] : m;
        blockmax[blockIdx.x] = m;
    }
}

template <typename T>
__global__ void max_partials_kernel(const T* __restrict__ part, int nparts, T* __restrict__ out) {
    __shared__ T red[256];
    T m = T(-1.0);
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) m = part[i] > m ? part[i] : m;
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = red[threadIdx.x + o] > red[threadIdx.x] ? red[threadIdx.x + o] : red[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// pass 2: the per-row greedy coordinate loop, one warp per row, row state in shared memory
// (greedycd.jl:139-165).  smem per warp: g[k], f[k], fnew[k]; per block: pdiag[k].
template <typename T>
__global__ void __launch_bounds__(GCD_WARPS * 32) gcd_rows_kernel(T* __restrict__ F, int64_t sFr, int64_t sFc,
                                                                 const T* __restrict__ G, const T* __restrict__ P,
                                                                 int64_t rows, int k, const T* __restrict__ p_init_ptr,
                                                                 unsigned long long* __restrict__ updates) {
    extern __shared__ unsigned char smem_raw[];
    T* pdiag = (T*)smem_raw;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    T* g = pdiag + k + (size_t)warp * 3 * k;
    T* f = g + k;
    T* fnew = f + k;
    for (int r = threadIdx.x; r < k; r += blockDim.x) pdiag[r] = P[r + (int64_t)r * k];
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * GCD_WARPS + warp;
    if (i >= rows) return;
    for (int r = lane; r < k; r += 32) {
        g[r] = G[i * k + r];
        f[r] = F[i * sFr + r * sFc];
        fnew[r] = T(0);
    }
    __syncwarp();
    const T thresh = mul_rn(T(0.001), p_init_ptr[0]);
    const int64_t maxsteps = (int64_t)k * k;
    unsigned long long nupd = 0;
    for (int64_t it = 0; it < maxsteps; ++it) {
        int bi = 0x7fffffff;
        T bv = (T)(-INFINITY);
        for (int r = lane; r < k; r += 32) {
            T s, d;
            gcd_sd(f[r], g[r], pdiag[r], s, d);
            if (d > bv) { bv = d; bi = r; }
        }
        warp_argmax(bv, bi);
        if (bv < thresh) break;
        T sq, dq;
        gcd_sd(f[bi], g[bi], pdiag[bi], sq, dq);
        __syncwarp();
        if (lane == 0) fnew[bi] = add_rn(fnew[bi], sq);
        // P[qi, r] (greedycd.jl:151).  P = O'O is symmetric (bit-for-bit: both triangles are the same sums of the same
        // products), so read the contiguous column P[:, qi] = P[r + qi*k] -- coalesced across the warp.
        const T* prow = P + (int64_t)bi * k;
        for (int r = lane; r < k; r += 32) g[r] = add_rn(g[r], mul_rn(sq, prow[r]));
        __syncwarp();
        ++nupd;
    }
    for (int r = lane; r < k; r += 32) {
        T v = add_rn(f[r], fnew[r]);
        if (v < T(0)) v = T(0);  // projectnn! (utils.jl:34-41)
        F[i * sFr + r * sFc] = v;
    }
    if (lane == 0 && nupd) atomicAdd(updates, nupd);
}


// Register-resident variant for the tensor-core engine: factor F and gradient G are row-major [R][KP], KP is a
// compile-time multiple of 32, every lane owns KP/32 components of its row in registers; the only memory traffic
// of a coordinate step is the contiguous row P[q, :] (KP/32 independent, coalesced loads per lane).
// 1/(eps + P[r,r]) is hoisted out of the loop (multiply instead of divide: the gradients are bf16-derived anyway).
template <int KP>
__global__ void __launch_bounds__(256) gcd_rows_tc_kernel(float* __restrict__ F, const float* __restrict__ G, const float* __restrict__ P,
                                                          int R, int k, const float* __restrict__ p_init_ptr,
                                                          unsigned long long* __restrict__ updates) {
    constexpr int NE = KP / 32;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    float g[NE], f[NE], fn[NE], prr[NE], rinv[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int r = lane + 32 * e;
        g[e] = G[row * KP + r];
        f[e] = F[row * KP + r];
        fn[e] = 0.f;
        prr[e] = P[(size_t)r * KP + r];
        rinv[e] = 1.0f / (1.1920928955078125e-07f + prr[e]);
    }
    const float thresh = 0.001f * p_init_ptr[0];       // nu * p_init (greedycd.jl:140,145)
    unsigned long long nupd = 0;
    const int maxsteps = k * k;                        // `for _ in 1:n_components^2` (:144): the real k, not the padded KP
    for (int it = 0; it < maxsteps; ++it) {
        float bv = -INFINITY, bs = 0.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < NE; ++e) {                 // S, D and the first arg-max (:153-158)
            const float t = f[e] - g[e] * rinv[e];
            const float s = fmaxf(t, 0.f) - f[e];
            const float d = -g[e] * s - 0.5f * prr[e] * s * s;
            if (d > bv) { bv = d; bi = lane + 32 * e; bs = s; }
        }
        // first arg-max over the warp with two REDUX instructions instead of 15 shuffles: the maximum of an order-preserving
        // integer image of D, then the smallest index among the lanes that hold it (findmax's tie rule, greedycd.jl:158)
        const unsigned ub = __float_as_uint(bv + 0.0f);   // -0 -> +0: the two zeros must tie, as they do for findmax
        const unsigned key = (ub & 0x80000000u) ? ~ub : (ub | 0x80000000u);
        const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
        bi = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)bi : 0x7fffffffu);
        bs = __shfl_sync(0xffffffffu, bs, bi & 31);    // component bi lives on lane bi % 32
        bv = __uint_as_float((kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax);
        if (bv < thresh) break;                        // :145-147
        const float* prow = P + (size_t)bi * KP;       // symmetric P: row q contiguous
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            if (bi == lane + 32 * e) fn[e] += bs;      // Wnew[i,q] += S[i,q] (:149)
            g[e] += bs * __ldg(prow + lane + 32 * e);  // G[i,:] += S[i,q] P[q,:] (:151)
        }
        ++nupd;
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) F[row * KP + lane + 32 * e] = fmaxf(f[e] + fn[e], 0.f);  // :164-165
    if (lane == 0 && nupd) atomicAdd(updates, nupd);
}

}  // namespace
}  // namespace nmfb200
