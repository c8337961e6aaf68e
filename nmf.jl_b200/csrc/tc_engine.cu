// tc_engine.cu -- the tensor-core engine for MultUpdate(:mse) (multupd.jl:83-116), Float32 API.
//
// One half-step "update factor F (R rows x k) against the other factor O" is ONE kernel
// (mu_update_kernel): a CTA owns 128 rows of F and computes, with tcgen05.mma into TMEM,
//     Num[r][a] = sum_c Xs[r][c] * O[c][a]        (X H' or (W'X)'; bf16 operands streamed by TMA)
//     Den[r][a] = sum_b F[r][b]  * P[b][a]        (P = O'O, k x k; bf16 hi/lo split => ~fp32 accuracy)
// and its epilogue applies   F <- F * max(0, Num - lambda) / (Den + delta)   straight out of TMEM
// (multupd.jl:101-103 / :112-114), writes the new F in the four forms the next kernels consume
// (fp32 master, bf16 hi/lo K-major tiles for Den, bf16 transposed copy as the next B operand) and the
// per-component stop_condition partial sums (common.jl:97-104).  No cuBLAS, no separate elementwise
// kernel.  H-step and W-step are the same kernel with the roles of the buffers swapped, because both
// factors are kept in "row-factor" layout ([rows][KP], KP = k padded to 64/128/256) and X is cached
// in bf16 in both orientations (Xr = [n][p] for the H-step, Xc = [p][n] for the W-step).
//
// The k x k Gram P = F'F of the freshly updated factor is a second small tcgen05 kernel
// (gram_kernel, split over row chunks, fp32 red.add into P, last CTA converts to bf16 hi/lo).
// stop_condition is finished by conv_kernel (one block).  The iteration loop is enqueued without
// host round trips: every kernel exits immediately once the device-side `converged` flag is set, and
// the host polls that flag every `check_every` iterations.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nmfb200 {
namespace {

using bf16 = __nv_bfloat16;
using namespace ptx;

struct TcState {
    int converged;
    int iters;
    float devmax;
    int pad;
};

// ---- kernel parameter block (tensor maps must live in __grid_constant__ param space) ---------------
struct UpdateParams {
    CUtensorMap tmA;    // Xs   bf16 [R][Kdim]     box 64 x 128
    CUtensorMap tmB;    // O^T  bf16 [KP][Kdim]    box 64 x KP
    CUtensorMap tmFhi;  // F hi bf16 [R][KP]       box 64 x 128
    CUtensorMap tmFlo;  // F lo
    CUtensorMap tmPhi;  // P hi bf16 [KP][KP]      box 64 x KP
    CUtensorMap tmPlo;  // P lo
    float* F;           // [R][KP] fp32 master, updated in place
    bf16* Fhi;          // [R][KP]
    bf16* Flo;          // [R][KP]
    bf16* FbT;          // [KP][ldT] transposed bf16 copy
    float* num_io;      // MODE 1: raw numerators out, MODE 2: reduced numerators in ([R][KP])
    float* conv_part;   // [tiles][2][KP]
    const TcState* state;
    int64_t ldT;
    int R, Kdim;
    float lambda, delta;
};

template <int KP>
struct UpdCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = KP * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = KP == 256 ? 4 : (KP == 128 ? 6 : 8);
    static constexpr int CONV_BYTES = 4 * 2 * KP * 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CONV_BYTES + 256 + 1024;
    static constexpr int TMEM_COLS = 2 * KP;
    static constexpr int NSLAB = KP / 64;
};

// sum v[j] over the 32 lanes of the warp; afterwards v[0] on lane l holds the total of column l
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            float send = up ? v[i] : v[i + o];
            float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}

__device__ __forceinline__ uint32_t pack_bf16x2(bf16 lo, bf16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// MODE 0: fused (single GPU).  MODE 1: numerators only -> num_io (row-sharded H-step, before the
// all-reduce).  MODE 2: no main loop, numerators read from num_io (after the all-reduce).
template <int KP, int MODE>
__global__ void __launch_bounds__(192, 1) mu_update_kernel(const __grid_constant__ UpdateParams prm) {
    using C = UpdCfg<KP>;
    if (prm.state->converged) return;  // uniform: the loop has already met stop_condition

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);
    float* conv_s = (float*)(smem + C::STAGES * C::STAGE_BYTES + 256);  // [4 warps][2][KP]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 128;
    const int nkb = (MODE == 2) ? 0 : (prm.Kdim + 63) / 64;
    constexpr int NPRE = (MODE == 1) ? 0 : 3 * C::NSLAB;
    const int total = NPRE + nkb;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&prm.tmA);
        prefetch_tmap(&prm.tmB);
        if (MODE != 1) {
            prefetch_tmap(&prm.tmFhi);
            prefetch_tmap(&prm.tmFlo);
            prefetch_tmap(&prm.tmPhi);
            prefetch_tmap(&prm.tmPlo);
        }
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int b = 0; b < total; ++b) {
                const int s = b % C::STAGES;
                const uint32_t ph = (uint32_t)(b / C::STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
                uint8_t* a_dst = smem + s * C::STAGE_BYTES;
                uint8_t* b_dst = a_dst + C::A_BYTES;
                if (b < NPRE) {  // Den = Fhi*Phi + Fhi*Plo + Flo*Phi
                    const int t = b / C::NSLAB, sl = b % C::NSLAB;
                    tma_load_2d(a_dst, t == 2 ? &prm.tmFlo : &prm.tmFhi, &full_bar[s], 64 * sl, r0);
                    tma_load_2d(b_dst, t == 1 ? &prm.tmPlo : &prm.tmPhi, &full_bar[s], 64 * sl, 0);
                } else {
                    const int kb = b - NPRE;
                    tma_load_2d(a_dst, &prm.tmA, &full_bar[s], 64 * kb, r0);
                    tma_load_2d(b_dst, &prm.tmB, &full_bar[s], 64 * kb, 0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, KP);
            for (int b = 0; b < total; ++b) {
                const int s = b % C::STAGES;
                const uint32_t ph = (uint32_t)(b / C::STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * C::STAGE_BYTES);
                const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + C::A_BYTES);
                const bool pre = b < NPRE;
                const uint32_t d = pre ? tmem_base + KP : tmem_base;
                const bool first = pre ? (b == 0) : (b == NPRE);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)  // 4 x (K = 16 bf16 = 32 B) per 128-B swizzle row
                    umma_bf16(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (!first || kk > 0) ? 1u : 0u);
                umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
            }
            umma_commit(tmem_full);
        }
        __syncwarp();
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int row = r0 + 32 * q + lane;
        const bool valid = row < prm.R;
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* convw = conv_s + q * 2 * KP;
        const float lambda = prm.lambda, delta = prm.delta;
#pragma unroll 1
        for (int c0 = 0; c0 < KP; c0 += 32) {
            uint32_t num_u[32], den_u[32];
            float f[32];
            if (MODE != 2) tmem_ld32(t_lane + c0, num_u);
            if (MODE != 1) tmem_ld32(t_lane + KP + c0, den_u);
            if (MODE == 1) {
                tmem_ld_wait();
                if (valid) {
                    float4* dst = (float4*)(prm.num_io + (size_t)row * KP + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(num_u[4 * j]), __uint_as_float(num_u[4 * j + 1]),
                                             __uint_as_float(num_u[4 * j + 2]), __uint_as_float(num_u[4 * j + 3]));
                }
                continue;
            }
            if (valid) {
                const float4* src = (const float4*)(prm.F + (size_t)row * KP + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = src[j];
                    f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                }
                if (MODE == 2) {
                    const float4* ns = (const float4*)(prm.num_io + (size_t)row * KP + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = ns[j];
                        num_u[4 * j] = __float_as_uint(v.x); num_u[4 * j + 1] = __float_as_uint(v.y);
                        num_u[4 * j + 2] = __float_as_uint(v.z); num_u[4 * j + 3] = __float_as_uint(v.w);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { f[j] = 0.f; if (MODE == 2) num_u[j] = 0u; }
            }
            tmem_ld_wait();
            float d2[32], s2[32];
            uint32_t hi_p[16], lo_p[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float fn[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float num = __uint_as_float(num_u[j + e]) - lambda;
                    num = (num > 0.f || num != num) ? num : 0.f;             // Julia max(0, x): NaN propagates
                    float den = __uint_as_float(den_u[j + e]) + delta;
                    float v = f[j + e] * __fdiv_rn(num, den);                // multupd.jl:102 / :113
                    fn[e] = valid ? v : 0.f;
                    float dd = fn[e] - f[j + e], ss = fn[e] + f[j + e];      // common.jl:98-99 / :103-104
                    d2[j + e] = dd * dd;
                    s2[j + e] = ss * ss;
                    f[j + e] = fn[e];
                }
                bf16 h0 = __float2bfloat16_rn(fn[0]), h1 = __float2bfloat16_rn(fn[1]);
                bf16 l0 = __float2bfloat16_rn(fn[0] - __bfloat162float(h0));
                bf16 l1 = __float2bfloat16_rn(fn[1] - __bfloat162float(h1));
                hi_p[j / 2] = pack_bf16x2(h0, h1);
                lo_p[j / 2] = pack_bf16x2(l0, l1);
            }
            if (valid) {
                float4* dst = (float4*)(prm.F + (size_t)row * KP + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                uint4* dh = (uint4*)(prm.Fhi + (size_t)row * KP + c0);
                uint4* dl = (uint4*)(prm.Flo + (size_t)row * KP + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dh[j] = make_uint4(hi_p[4 * j], hi_p[4 * j + 1], hi_p[4 * j + 2], hi_p[4 * j + 3]);
                    dl[j] = make_uint4(lo_p[4 * j], lo_p[4 * j + 1], lo_p[4 * j + 2], lo_p[4 * j + 3]);
                }
                // transposed bf16 copy: FbT[a][row]; a warp writes 32 consecutive rows (64 B) per component
                unsigned short* tb = (unsigned short*)prm.FbT + (size_t)c0 * prm.ldT + row;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    uint32_t pk = hi_p[j / 2];
                    tb[(size_t)j * prm.ldT] = (unsigned short)((j & 1) ? (pk >> 16) : (pk & 0xffffu));
                }
            }
            warp_transpose_reduce(d2, lane);
            warp_transpose_reduce(s2, lane);
            convw[c0 + lane] = d2[0];
            convw[KP + c0 + lane] = s2[0];
        }
        if (MODE != 1) {
            // combine the four lane quarters: named barrier over the 128 epilogue threads
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;  // 0..127
            for (int i = t; i < 2 * KP; i += 128) {
                float s = conv_s[i] + conv_s[2 * KP + i] + conv_s[4 * KP + i] + conv_s[6 * KP + i];
                prm.conv_part[(size_t)blockIdx.x * 2 * KP + i] = s;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- Gram: P += T T'  for T = FbT ([KP][R] bf16, rows of length R contiguous) ------------------------
struct GramParams {
    CUtensorMap tmT;  // bf16 [KP][R], box 64 x 128
    float* part;      // [gridDim.x][KP][KP] fp32 partial Grams (plain stores, reduced by gram_reduce_kernel)
    const TcState* state;
    int R, chunk;     // rows (K extent) per CTA, multiple of 64
};

template <int KP>
struct GramCfg {
    static constexpr int MT = (KP + 127) / 128;         // 128-row M tiles
    static constexpr int STAGE_BYTES = MT * 128 * 128;  // the tile is both A and B operand
    static constexpr int STAGES = 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
    static constexpr int TMEM_COLS = (MT * KP) < 32 ? 32 : (MT * KP);  // 64, 128, 512
};

template <int KP>
__global__ void __launch_bounds__(192, 1) gram_kernel(const __grid_constant__ GramParams prm) {
    using C = GramCfg<KP>;
    if (prm.state->converged) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_begin = blockIdx.x * prm.chunk;
    const int k_end = min(prm.R, k_begin + prm.chunk);
    const int nkb = (k_end - k_begin + 63) / 64;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&prm.tmT);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int b = 0; b < nkb; ++b) {
                const int s = b % C::STAGES;
                const uint32_t ph = (uint32_t)(b / C::STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
                // NOTE: columns >= k_end inside the last 64-block belong to the next CTA's chunk only if
                // chunk % 64 != 0; chunk is a multiple of 64, and columns >= R are zero-filled by TMA.
                for (int m = 0; m < C::MT; ++m)
                    tma_load_2d(smem + s * C::STAGE_BYTES + m * 128 * 128, &prm.tmT, &full_bar[s], k_begin + 64 * b, 128 * m);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, KP);
            for (int b = 0; b < nkb; ++b) {
                const int s = b % C::STAGES;
                const uint32_t ph = (uint32_t)(b / C::STAGES) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * C::STAGE_BYTES);
                const uint64_t bdesc = make_kmajor_sw128_desc(base);  // B = first KP rows of the tile
#pragma unroll
                for (int m = 0; m < C::MT; ++m) {
                    const uint64_t adesc = make_kmajor_sw128_desc(base + m * 128 * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + m * KP, adesc + 2 * kk, bdesc + 2 * kk, idesc, (b > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(tmem_full);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* part = prm.part + (size_t)blockIdx.x * KP * KP;
#pragma unroll 1
        for (int m = 0; m < C::MT; ++m) {
            const int a = 128 * m + 32 * q + lane;
#pragma unroll 1
            for (int c0 = 0; c0 < KP; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + m * KP + c0, v);
                tmem_ld_wait();
                if (a < KP) {
                    float4* dst = (float4*)(part + (size_t)a * KP + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                             __uint_as_float(v[4 * j + 3]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// P[e] = sum_g part[g][e]; writes the fp32 Gram and (do_split) its bf16 hi/lo split.  One element per thread.
__global__ void __launch_bounds__(256) gram_reduce_kernel(const float* __restrict__ part, int nparts, int nelem, float* __restrict__ P,
                                                          bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split,
                                                          const TcState* st) {
    if (st->converged) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nelem) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int g = 0;
    for (; g + 4 <= nparts; g += 4) {
        s0 += part[(size_t)g * nelem + i];
        s1 += part[(size_t)(g + 1) * nelem + i];
        s2 += part[(size_t)(g + 2) * nelem + i];
        s3 += part[(size_t)(g + 3) * nelem + i];
    }
    for (; g < nparts; ++g) s0 += part[(size_t)g * nelem + i];
    const float v = (s0 + s1) + (s2 + s3);
    P[i] = v;
    if (do_split) {
        bf16 hi = __float2bfloat16_rn(v);
        Phi[i] = hi;
        Plo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// multi-GPU: split an all-reduced fp32 Gram into bf16 hi/lo
__global__ void gram_split_kernel(const float* __restrict__ P, bf16* __restrict__ Phi, bf16* __restrict__ Plo, int n, const TcState* st) {
    if (st->converged) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = P[i];
        bf16 hi = __float2bfloat16_rn(v);
        Phi[i] = hi;
        Plo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// ---- stop_condition finish (common.jl:92-111): one block, thread a owns component a ---------------
// acc (double [4][KP]) = {dev_w, sum_w, dev_h, sum_h}; stage 0 = reduce tiles into acc, stage 1 = decide.
__global__ void conv_kernel(const float* __restrict__ partW, int tilesW, const float* __restrict__ partH, int tilesH, int KP, int k,
                            int update_H, double* __restrict__ acc, float tol, TcState* st, int do_reduce, int do_decide) {
    if (st->converged) return;
    __shared__ int fail;
    __shared__ float devs[256];
    __shared__ double red[4][1024];
    const int a = threadIdx.x;
    if (a == 0) fail = 0;
    __syncthreads();
    if (do_reduce) {
        // blockDim.x = 1024: KP components x (1024 / KP) tile groups, then a shared-memory pass over the groups
        const int groups = blockDim.x / KP;
        const int g = threadIdx.x / KP, c = threadIdx.x % KP;
        double dw = 0, sw = 0, dh = 0, sh = 0;
        if (g < groups) {
            for (int t = g; t < tilesW; t += groups) {
                dw += (double)partW[(size_t)t * 2 * KP + c];
                sw += (double)partW[(size_t)t * 2 * KP + KP + c];
            }
            if (update_H)
                for (int t = g; t < tilesH; t += groups) {
                    dh += (double)partH[(size_t)t * 2 * KP + c];
                    sh += (double)partH[(size_t)t * 2 * KP + KP + c];
                }
        }
        red[0][threadIdx.x] = dw; red[1][threadIdx.x] = sw; red[2][threadIdx.x] = dh; red[3][threadIdx.x] = sh;
        __syncthreads();
        if (threadIdx.x < KP) {
            for (int gg = 1; gg < groups; ++gg) {
                dw += red[0][gg * KP + c]; sw += red[1][gg * KP + c]; dh += red[2][gg * KP + c]; sh += red[3][gg * KP + c];
            }
            if (!update_H) { dh = 0; sh = 1; }
            acc[c] = dw; acc[KP + c] = sw; acc[2 * KP + c] = dh; acc[3 * KP + c] = sh;
        }
    }
    if (!do_decide) return;
    __syncthreads();
    float dev = 0.f;
    if (a < k) {
        float dw = (float)acc[a], sw = (float)acc[KP + a], dh = (float)acc[2 * KP + a], sh = (float)acc[3 * KP + a];
        float rw = dw / sw, rh = dh / sh;
        float m = (rw != rw) ? rw : ((rh != rh) ? rh : fmaxf(rw, rh));
        dev = sqrtf(m);
        if (sqrtf(dw) > tol * sqrtf(sw) || sqrtf(dh) > tol * sqrtf(sh)) atomicExch(&fail, 1);
    }
    if (a < 256) devs[a] = dev;
    __syncthreads();
    if (a == 0) {
        float dm = 0.f;
        for (int i = 0; i < k; ++i) dm = (dm != dm) ? dm : ((devs[i] != devs[i]) ? devs[i] : fmaxf(dm, devs[i]));
        st->devmax = dm;
        st->iters += 1;
        if (!fail) st->converged = 1;
    }
}

// ---- X caches ---------------------------------------------------------------------------------------
// Xr[j][i] = bf16(X[i + j*ldx])  (same orientation as the caller's column-major X)
__global__ void cvt_rows_kernel(const float* __restrict__ X, int64_t p, int64_t n, int64_t ldx, bf16* __restrict__ Xr, int64_t ldp) {
    const int64_t j = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ldp; i += (int64_t)gridDim.x * blockDim.x)
        Xr[j * ldp + i] = __float2bfloat16_rn(i < p ? X[i + j * ldx] : 0.f);
}
// Xc[i][j] = bf16(X[i + j*ldx])  (transposed), 32x32 tiles through shared memory
__global__ void cvt_transpose_kernel(const float* __restrict__ X, int64_t p, int64_t n, int64_t ldx, bf16* __restrict__ Xc, int64_t ldn) {
    __shared__ float tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + threadIdx.x, j = j0 + r;
        tile[r][threadIdx.x] = (i < p && j < n) ? X[i + j * ldx] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r, j = j0 + threadIdx.x;
        if (i < p && j < ldn) Xc[i * ldn + j] = __float2bfloat16_rn(tile[threadIdx.x][r]);
    }
}

// ---- factor packing / unpacking -----------------------------------------------------------------------
// src(r, a) = S[r*sr + a*sa] (r < R, a < k) -> Fm[r][a], Fhi, Flo ([R][KP]) and FbT[a][r] ([KP][ldT]); zero padded
__global__ void pack_factor_kernel(const float* __restrict__ S, int64_t sr, int64_t sa, int R, int k, int KP, float* __restrict__ Fm,
                                   bf16* __restrict__ Fhi, bf16* __restrict__ Flo, bf16* __restrict__ FbT, int64_t ldT) {
    const int64_t total = (int64_t)R * KP;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / KP;
        const int a = (int)(idx % KP);
        float v = a < k ? S[r * sr + a * sa] : 0.f;
        bf16 hi = __float2bfloat16_rn(v);
        Fm[idx] = v;
        Fhi[idx] = hi;
        Flo[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
        FbT[(int64_t)a * ldT + r] = hi;
    }
}
__global__ void unpack_factor_kernel(const float* __restrict__ Fm, int R, int k, int KP, float* __restrict__ D, int64_t sr, int64_t sa) {
    const int64_t total = (int64_t)R * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx / k;
        const int a = (int)(idx % k);
        D[r * sr + a * sa] = Fm[r * KP + a];
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        NMF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        NMF_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, NMFB200_ECUDA, "cuTensorMapEncodeTiled unavailable");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 matrix [rows][inner] with row pitch ld (elements); box = 64 (128 B) x box_rows, SWIZZLE_128B
CUtensorMap make_tmap_bf16(const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * sizeof(bf16)};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NMF_REQUIRE(r == CUDA_SUCCESS, NMFB200_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int pick_kp(int64_t k) { return k <= 64 ? 64 : (k <= 128 ? 128 : (k <= 256 ? 256 : 0)); }
inline int ew_grid(int64_t len) { return (int)std::min<int64_t>(ceil_div(len, 256), 148 * 16); }

struct Factor {  // one factor in row-factor layout
    int R = 0;
    int rowsT = 0;
    int64_t ldT = 0;
    float* m = nullptr;
    bf16 *hi = nullptr, *lo = nullptr, *bT = nullptr;
    float* P = nullptr;  // Gram of THIS factor (k x k), fp32 accumulator
    bf16 *Phi = nullptr, *Plo = nullptr;
    float* conv = nullptr;
    int tiles = 0;
};

template <int KP>
struct TcSolver {
    nmfb200_handle* h;
    cudaStream_t st;
    TcState* state;

    void launch_update(int mode, const Factor& F, const Factor& O, const bf16* Xs, int64_t ldX, int Kdim, float lambda, float delta,
                       float* num_io) {
        UpdateParams prm;
        prm.tmA = make_tmap_bf16(Xs, (uint64_t)Kdim, (uint64_t)F.R, (uint64_t)ldX, 128);
        prm.tmB = make_tmap_bf16(O.bT, (uint64_t)Kdim, KP, (uint64_t)O.ldT, KP);
        prm.tmFhi = make_tmap_bf16(F.hi, KP, (uint64_t)F.R, KP, 128);
        prm.tmFlo = make_tmap_bf16(F.lo, KP, (uint64_t)F.R, KP, 128);
        prm.tmPhi = make_tmap_bf16(O.Phi, KP, KP, KP, KP);
        prm.tmPlo = make_tmap_bf16(O.Plo, KP, KP, KP, KP);
        prm.F = F.m; prm.Fhi = F.hi; prm.Flo = F.lo; prm.FbT = F.bT; prm.ldT = F.ldT;
        prm.num_io = num_io;
        prm.conv_part = F.conv;
        prm.state = state;
        prm.R = F.R; prm.Kdim = Kdim; prm.lambda = lambda; prm.delta = delta;
        const int smem = UpdCfg<KP>::SMEM_BYTES;
        const bool timed = h->time_kernels && mode != 2;
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        if (mode == 0) mu_update_kernel<KP, 0><<<F.tiles, 192, smem, st>>>(prm);
        else if (mode == 1) mu_update_kernel<KP, 1><<<F.tiles, 192, smem, st>>>(prm);
        else mu_update_kernel<KP, 2><<<F.tiles, 192, smem, st>>>(prm);
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        h->launches += 1;
    }

    void launch_gram(const Factor& F, bool split, float* P_dst = nullptr) {
        GramParams g;
        g.tmT = make_tmap_bf16(F.bT, (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, 128);
        // ~128 CTAs at most, each a multiple of 64 rows and at least 256
        g.chunk = (int)std::max<int64_t>(256, round_up(ceil_div(F.R, 128), 64));
        const int grid = (int)ceil_div(F.R, g.chunk);
        g.part = h->buf_t<float>("tc.gram_part", (size_t)grid * KP * KP);
        g.state = state;
        g.R = F.R;
        gram_kernel<KP><<<grid, 192, GramCfg<KP>::SMEM_BYTES, st>>>(g);
        gram_reduce_kernel<<<(KP * KP + 255) / 256, 256, 0, st>>>(g.part, grid, KP * KP, P_dst ? P_dst : F.P, F.Phi, F.Plo, split ? 1 : 0, state);
        h->launches += 2;
    }

    static void set_attrs() {
        static bool done = false;
        if (done) return;
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(gram_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg<KP>::SMEM_BYTES));
        done = true;
    }
};

Factor alloc_factor(nmfb200_handle* h, const char* tag, int R, int KP) {
    Factor f;
    std::string t(tag);
    f.R = R;
    f.ldT = round_up(R, 64);
    f.tiles = (int)ceil_div(R, 128);
    f.m = h->buf_t<float>("tc." + t + ".m", (size_t)R * KP);
    f.hi = h->buf_t<bf16>("tc." + t + ".hi", (size_t)R * KP);
    f.lo = h->buf_t<bf16>("tc." + t + ".lo", (size_t)R * KP);
    f.rowsT = KP < 128 ? 128 : KP;  // gram_kernel loads 128-row M tiles: keep zero rows behind KP = 64
    f.bT = h->buf_t<bf16>("tc." + t + ".bT", (size_t)f.rowsT * f.ldT);
    f.P = h->buf_t<float>("tc." + t + ".P", (size_t)KP * KP);
    f.Phi = h->buf_t<bf16>("tc." + t + ".Phi", (size_t)KP * KP);
    f.Plo = h->buf_t<bf16>("tc." + t + ".Plo", (size_t)KP * KP);
    f.conv = h->buf_t<float>("tc." + t + ".conv", (size_t)f.tiles * 2 * KP);
    return f;
}

template <int KP>
void tc_solve_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs();
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const int64_t ldp = round_up(p, 64), ldn = round_up(n, 64);
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;

    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));

    // bf16 caches of X in both orientations (built once per set_X)
    bf16* Xr = h->buf_t<bf16>("tc.Xr", (size_t)n * ldp);
    bf16* Xc = h->buf_t<bf16>("tc.Xc", (size_t)p * ldn);
    if (h->tc_x_epoch != h->x_epoch) {
        const float* X = (const float*)h->dX;
        cvt_rows_kernel<<<dim3((unsigned)std::min<int64_t>(ceil_div(ldp, 256), 64), (unsigned)n), 256, 0, st>>>(X, p, n, h->ldx, Xr, ldp);
        cvt_transpose_kernel<<<dim3((unsigned)ceil_div(p, 32), (unsigned)ceil_div(ldn, 32)), dim3(32, 8), 0, st>>>(X, p, n, h->ldx, Xc, ldn);
        h->launches += 2;
        NMF_CUDA(cudaGetLastError());
        h->tc_x_epoch = h->x_epoch;
    }
    NMF_CUDA(cudaEventRecord(e0, st));

    Factor W = alloc_factor(h, "W", (int)p, KP), H = alloc_factor(h, "H", (int)n, KP);
    TcState* state = (TcState*)h->buf("tc.state", sizeof(TcState));
    double* acc = h->buf_t<double>("tc.acc", 4 * KP);
    NMF_CUDA(cudaMemsetAsync(state, 0, sizeof(TcState), st));
    NMF_CUDA(cudaMemsetAsync(W.bT, 0, (size_t)W.rowsT * W.ldT * sizeof(bf16), st));
    NMF_CUDA(cudaMemsetAsync(H.bT, 0, (size_t)H.rowsT * H.ldT * sizeof(bf16), st));

    // stage the caller's factors (column-major W p x k, H k x n)
    float *Wd = Wc, *Hd = Hc;
    int64_t ldwd = ldw, ldhd = ldh;
    if (!a.on_device) {
        Wd = h->buf_t<float>("tc.Wstage", (size_t)p * k);
        Hd = h->buf_t<float>("tc.Hstage", (size_t)k * n);
        ldwd = p;
        ldhd = k;
        NMF_CUDA(cudaMemcpy2DAsync(Wd, p * sizeof(float), Wc, ldw * sizeof(float), p * sizeof(float), k, cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hd, k * sizeof(float), Hc, ldh * sizeof(float), k * sizeof(float), n, cudaMemcpyHostToDevice, st));
    }
    pack_factor_kernel<<<ew_grid(p * KP), 256, 0, st>>>(Wd, 1, ldwd, (int)p, (int)k, KP, W.m, W.hi, W.lo, W.bT, W.ldT);
    pack_factor_kernel<<<ew_grid(n * KP), 256, 0, st>>>(Hd, ldhd, 1, (int)n, (int)k, KP, H.m, H.hi, H.lo, H.bT, H.ldT);
    h->launches += 2;
    NMF_CUDA(cudaGetLastError());

    TcSolver<KP> s{h, st, state};
    const bool multi = h->comm != nullptr;
    // multi-GPU (rows of X, W sharded; H replicated): packed all-reduce buffer [ (W_g' X_g)' : n x KP | W_g' W_g : KP x KP ]
    float* packed = multi ? h->buf_t<float>("tc.packed", (size_t)n * KP + (size_t)KP * KP) : nullptr;
    float* packed_P = multi ? packed + (size_t)n * KP : nullptr;
    s.launch_gram(W, !multi, packed_P);       // P_W = W'W for the first H-step (partial per rank when sharded)
    if (!a.update_H) s.launch_gram(H, true);  // H never changes: P_H once
    NMF_CUDA(cudaEventRecord(e1, st));

    h->ev_used = 0;
    int64_t enq = 0;
    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    while (enq < a.maxiter) {
        int64_t batch = std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            if (a.update_H) {
                if (!multi) {
                    s.launch_update(0, H, W, Xr, ldp, (int)p, lh, delta, nullptr);  // H-step: rows of H' against W
                } else {
                    s.launch_update(1, H, W, Xr, ldp, (int)p, lh, delta, packed);   // partial numerators of this shard
                    h->allreduce_sum(packed, (size_t)n * KP + (size_t)KP * KP);     // THE exchange step of the iteration
                    gram_split_kernel<<<(KP * KP + 255) / 256, 256, 0, st>>>(packed_P, W.Phi, W.Plo, KP * KP, state);
                    h->launches += 1;
                    s.launch_update(2, H, W, Xr, ldp, (int)p, lh, delta, packed);   // ratio with the reduced numerators
                }
                s.launch_gram(H, true);
            }
            s.launch_update(0, W, H, Xc, ldn, (int)n, lw, delta, nullptr);          // W-step (local rows)
            s.launch_gram(W, !multi, packed_P);
            if (!multi) {
                conv_kernel<<<1, 1024, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 1, 1);
                h->launches += 1;
            } else {
                conv_kernel<<<1, 1024, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 1, 0);
                h->allreduce_sum(acc, (size_t)2 * KP);  // dev_w, sum_w over all row shards; the H sums are replicated
                conv_kernel<<<1, 1024, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 0, 1);
                h->launches += 2;
            }
        }
        enq += batch;
        NMF_CUDA(cudaGetLastError());
        NMF_CUDA(cudaMemcpyAsync(&hs, state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        iters = hs.iters;
        devmax = hs.devmax;
        if (hs.converged) {
            converged = true;
            break;
        }
    }
    NMF_CUDA(cudaEventRecord(e2, st));

    // results back in the caller's layout
    unpack_factor_kernel<<<ew_grid(p * k), 256, 0, st>>>(W.m, (int)p, (int)k, KP, Wd, 1, ldwd);
    unpack_factor_kernel<<<ew_grid(n * k), 256, 0, st>>>(H.m, (int)n, (int)k, KP, Hd, ldhd, 1);
    h->launches += 2;
    NMF_CUDA(cudaGetLastError());
    // objective 0.5*||X - WH||^2 (multupd.jl:81): exact fp32 GEMM + fp64 reduction from the SIMT engine
    double objv = simt_objective_f32(h, 0, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);
    if (!a.on_device) {
        NMF_CUDA(cudaMemcpy2DAsync(Wc, ldw * sizeof(float), Wd, p * sizeof(float), p * sizeof(float), k, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hc, ldh * sizeof(float), Hd, k * sizeof(float), k * sizeof(float), n, cudaMemcpyDeviceToHost, st));
    }
    NMF_CUDA(cudaStreamSynchronize(st));
    float ms_up = 0, ms_loop = 0;
    cudaEventElapsedTime(&ms_up, e0, e1);
    cudaEventElapsedTime(&ms_loop, e1, e2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    out->niters = iters;
    out->converged = converged ? 1 : 0;
    out->engine = 1;
    out->objvalue = objv;
    out->last_dev = devmax;
    out->solve_ms = ms_loop;
    out->upload_ms = ms_up;
    out->coordinate_updates = 0;
    out->kernel_launches = h->launches;
    out->hot_kernel_ms = h->drain_event_pairs(&out->hot_kernel_launches);
}

}  // namespace

bool tc_supported(const nmfb200_handle* h, const SolveArgs& a) {
    if (a.alg != 0) return false;         // MultUpdate(:mse) only, so far
    if (a.verbose) return false;          // per-iteration objective: exact engine
    if (pick_kp(a.k) == 0) return false;
    if (h->p > (int64_t)INT32_MAX / 256 || h->n > (int64_t)INT32_MAX / 256) return false;
    return true;
}

void tc_solve(nmfb200_handle* h, const SolveArgs& a, float* W, int64_t ldw, float* H, int64_t ldh, nmfb200_result* out) {
    switch (pick_kp(a.k)) {
        case 64: tc_solve_kp<64>(h, a, W, ldw, H, ldh, out); break;
        case 128: tc_solve_kp<128>(h, a, W, ldw, H, ldh, out); break;
        case 256: tc_solve_kp<256>(h, a, W, ldw, H, ldh, out); break;
        default: throw Error{NMFB200_ENOTSUP, "k > 256 is not covered by the tensor-core engine"};
    }
}

void tc_release(nmfb200_handle*) {}

}  // namespace nmfb200
