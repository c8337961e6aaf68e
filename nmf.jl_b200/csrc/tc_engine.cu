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
// stop_condition is finished by conv_reduce_kernel (last block decides).  The loop is enqueued without
// host round trips: every kernel exits immediately once the device-side `converged` flag is set, and
// the host polls that flag every `check_every` iterations.
#include <cuda_bf16.h>

#include <algorithm>
#include <climits>
#include <utility>

#include "common.cuh"
#include "gcd_kernels.cuh"
#include "tc_ptx.cuh"

namespace nmfb200 {
namespace {

using bf16 = __nv_bfloat16;
using namespace ptx;

struct TcState {
    int converged;
    int iters;
    float devmax;
    unsigned int ticket;
};

// ---- peer-memory exchange (multi-GPU) ---------------------------------------------------------------
// The packed vector [numerators n x KP | W'W KP x KP | W-side stop sums 2 x KP] is treated as Rtot = n+KP+2
// rows of KP floats, cut into G contiguous segments of RS rows; rank j owns (reduces) segment j.
// Arena of every rank (IPC-mapped into all peers): flags | packed [G*RS][KP].  A rank's kernels write their
// partial sums into its own `packed`; the exchange kernel PULLS its segment from every peer over NVLink, sums
// in rank order and PUSHES the reduced segment into every peer's `packed` (reduce-scatter + all-gather fused).
struct XchgDev {
    float* packed[XCHG_MAX_RANKS];        // packed[j] = rank j's packed vector (mapped peer memory)
    unsigned int* flags[XCHG_MAX_RANKS];  // flags[j][phase * XCHG_MAX_RANKS + src]: "src reached epoch in phase"
    unsigned int* ticket;                 // local counter for the last-block pattern
    int G, rank, RS;
};

// ---- kernel parameter block (tensor maps must live in __grid_constant__ param space) ---------------
struct UpdateParams {
    CUtensorMap tmA;    // Xs   bf16 tile-contiguous [tiles*nkb*tile_rows][64], box 64 x tile_rows
    CUtensorMap tmB;    // O^T  bf16 [KP][Kdim]    box 64 x KP
    CUtensorMap tmFhi;  // F hi bf16 [R][KP]       box 64 x 128
    CUtensorMap tmFlo;  // F lo
    CUtensorMap tmPhi;  // P hi bf16 [KP][KP]      box 64 x KP
    CUtensorMap tmPlo;  // P lo
    CUtensorMap tmF32;  // F fp32 [R][KP]          box 32 x tile_rows (staged epilogue store)
    CUtensorMap tmT;    // F^T bf16 [KP][R]        box 64 x KP        (staged epilogue store of the transposed copy)
    float* gram_part;   // staged epilogue: [tiles][KP][KP] fp32 Gram contribution of each tile (nullptr = skip)
    float* F;           // [R][KP] fp32 master, updated in place
    bf16* Fhi;          // [R][KP]
    bf16* Flo;          // [R][KP]
    bf16* FbT;          // [KP][ldT] transposed bf16 copy
    float* num_io;      // MODE 1: raw numerators out, MODE 2: reduced numerators in ([R][KP]); MODE 5: num_splits k-split partials in
    int num_splits;     // MODE 5: numerators = sum over s < num_splits of num_io[s * num_split_stride + ...] (in order)
    int64_t num_split_stride;
    float* conv_part;   // [tiles][2][KP]   (MODE 3: [tiles] per-CTA max of D, greedycd.jl:132-137)
    const float* Pfull; // MODE 3: fp32 Gram of the other factor ([KP][KP]); its diagonal enters S and D
    const float* colsum; // MODE 4: column sums of the other factor (sW / sH of multupd.jl:176,188), [KP]
    const TcState* state;
    int64_t ldT;
    int R, Kdim;
    int tile_rows;      // rows of F owned by one CTA (<= 128, multiple of 8); the TMA boxes of A / Fhi / Flo have this many rows
    long long* timing;  // diagnostics (tc_debug bit 3): CTA 0 records clock64() at its phase boundaries, see TSTAMP
    float lambda, delta;
};

template <int KP>
struct UpdCfg {
    static constexpr int A_BYTES = 128 * 128;   // A part of a stage: up to 128 rows x 64 bf16
    static constexpr int B_BYTES = KP * 128;    // B part: KP rows x 64 bf16
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // One ring, A and B of a k-block travel together (one wait + one commit per block on the MMA thread).
    // Deeper / split rings were measured and bought nothing (profiles/r1b_pipeline_experiments.md).
    static constexpr int STAGES = KP == 256 ? 4 : (KP == 128 ? 6 : 8);
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int CONV_BYTES = 4 * 2 * KP * 4;
    static constexpr int SMEM_BYTES = RING_BYTES + CONV_BYTES + 1024 + 1024;  // ring | barriers (1 KB) | conv scratch | align slack
    static constexpr int TMEM_COLS = 2 * KP;
    static constexpr int NSLAB = KP / 64;
    static constexpr int THREADS = 320;  // w0 TMA producer, w1 MMA issuer, w2-9 epilogue (lane quarter = warp % 4, column half = (warp-2)/4)
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
// MODE 4: MultUpdate(:div): Xs is the quotient panel Q, no denominator MMAs; F <- F * Num / (colsum + lambda) (multupd.jl:177-179,189-191).
// MODE 3: GreedyCD gradient: G = F*P - Xs*O (+lambda) -> num_io, per-CTA max_r D[i,r] -> conv_part (greedycd.jl:117-137).
// MODE 5: MultUpdate(:div) after div_fused_kernel: no main loop, numerators = sum of the k-split partials in num_io, then as MODE 4.
template <int KP, int MODE>
__global__ void __launch_bounds__(UpdCfg<KP>::THREADS, 1) mu_update_kernel(const __grid_constant__ UpdateParams prm) {
    using C = UpdCfg<KP>;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + C::RING_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* gram_bar = tmem_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(gram_bar + 1);
    uint32_t* stop_slot = tmem_slot + 1;
    float* conv_s = (float*)(smem + C::RING_BYTES + 1024);  // [4 warps][2][KP]
    // Staged epilogue (KP <= 128, modes that write the factor): the ring is idle once the accumulators are complete
    constexpr bool STAGED = (KP <= 128) && (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5);
    uint8_t* const SF = smem;                              // fp32 tile:  KP/32 boxes of 128 rows x 128 B
    uint8_t* const SH = SF + (KP / 32) * 16384;            // bf16 hi:    KP/64 boxes
    uint8_t* const SL = SH + (KP / 64) * 16384;            // bf16 lo
    uint8_t* const ST = SL + (KP / 64) * 16384;            // transposed: 2 boxes of KP rows x 128 B (64 tile rows each)
    static_assert(!STAGED || (KP / 32 + 2 * (KP / 64)) * 16384 + 2 * KP * 128 <= C::RING_BYTES, "staging does not fit in the ring");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define TSTAMP(i) do { if (prm.timing != nullptr && blockIdx.x == 0) prm.timing[i] = clock64(); } while (0)
    if (threadIdx.x == 0) TSTAMP(0);
    if (prm.timing != nullptr && threadIdx.x == 0) {  // every CTA: global timer at entry (and exit, below)
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.timing[16 + 2 * blockIdx.x] = gt;
    }
    const int tile_rows = prm.tile_rows;
    const int r0 = blockIdx.x * tile_rows;
    const uint32_t a_bytes = (uint32_t)tile_rows * 128u;
    const int nkb = (MODE == 2 || MODE == 5) ? 0 : (prm.Kdim + 63) / 64;
    constexpr int NPRE = (MODE == 1 || MODE == 4 || MODE == 5) ? 0 : 3 * C::NSLAB;

    if (warp == 0 && lane == 0) {
        // Has the loop already met stop_condition?  ONE thread samples the flag for the whole CTA: under PDL (see
        // launch_update) the preceding kernel may be writing it right now, and the early exit below must be uniform.
        // A CTA that still sees 0 here streams its panel and skips the epilogue after pdl_wait().
        *stop_slot = (uint32_t)__ldcg(&prm.state->converged);
        prefetch_tmap(&prm.tmA);
        prefetch_tmap(&prm.tmB);
        if (MODE != 1 && MODE != 4 && MODE != 5) {
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
        mbar_init(gram_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (*stop_slot != 0u) {  // uniform early exit
        if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
        return;
    }
    // Block order: the nkb numerator blocks FIRST (they depend on nothing the preceding kernel writes), then the NPRE
    // denominator blocks (the Gram hi/lo they read is produced by the immediately preceding reduce kernel).
    // The two single-thread loops below are the latency-critical part of the kernel: no per-block branches, no
    // div/mod, everything loop-invariant is hoisted (an extra compare per block is measurable at 256 blocks).

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            uint8_t* dst = smem;
            int arow = blockIdx.x * nkb * tile_rows;   // tile-contiguous X: k-block kb of this tile starts at panel row arow0 + kb*tile_rows
            const uint32_t num_tx = a_bytes + (uint32_t)C::B_BYTES;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], num_tx);
                tma_load_2d(dst, &prm.tmA, &full_bar[s], 0, arow);
                tma_load_2d(dst + C::A_BYTES, &prm.tmB, &full_bar[s], 64 * kb, 0);
                arow += tile_rows;
                dst += C::STAGE_BYTES;
                if (++s == C::STAGES) { s = 0; ph ^= 1u; dst = smem; }
            }
            if (NPRE > 0) {
                pdl_wait();  // the Gram of the other factor comes from the preceding (reduce) kernel
#pragma unroll
                for (int bd = 0; bd < NPRE; ++bd) {  // Den = Fhi*Phi + Fhi*Plo + Flo*Phi
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], num_tx);
                    const int t = bd / C::NSLAB, sl = bd % C::NSLAB;   // compile-time after unrolling
                    tma_load_2d(dst, t == 2 ? &prm.tmFlo : &prm.tmFhi, &full_bar[s], 64 * sl, r0);
                    tma_load_2d(dst + C::A_BYTES, t == 1 ? &prm.tmPlo : &prm.tmPhi, &full_bar[s], 64 * sl, 0);
                    dst += C::STAGE_BYTES;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; dst = smem; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, KP);
            int s = 0;
            uint32_t ph = 0;
            const uint64_t adesc0 = make_kmajor_sw128_desc(smem_u32(smem));
            const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(smem + C::A_BYTES));
            uint64_t adesc = adesc0, bdesc = bdesc0;
            // one k-block: wait for its operands, 4 x (K = 16 bf16 = 32 B per 128-B swizzle row), free the stage
            auto block = [&](uint32_t d, uint32_t acc0) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                umma_bf16(d, adesc, bdesc, idesc, acc0);
#pragma unroll
                for (int kk = 1; kk < 4; ++kk) umma_bf16(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
                umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
                adesc += C::STAGE_BYTES >> 4;  // descriptor start address is in 16-byte units
                bdesc += C::STAGE_BYTES >> 4;
                if (++s == C::STAGES) { s = 0; ph ^= 1u; adesc = adesc0; bdesc = bdesc0; }
            };
            int kb = 0;
            if (nkb > 0) { block(tmem_base, 0u); kb = 1; TSTAMP(1); }   // first operands have landed
            for (; kb < nkb; ++kb) block(tmem_base, 1u);
            TSTAMP(2);                                                   // numerator blocks issued
            if (NPRE > 0) {
                block(tmem_base + KP, 0u);
#pragma unroll 1
                for (int bd = 1; bd < NPRE; ++bd) block(tmem_base + KP, 1u);
            }
            if (MODE != 5) umma_commit(tmem_full);
            TSTAMP(3);                                                   // all MMAs issued
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4, columns [chalf*KP/2, (chalf+1)*KP/2) =====
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int row = r0 + 32 * q + lane;
        const bool valid = (32 * q + lane) < tile_rows && row < prm.R;
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
        pdl_wait();  // from here on we read / overwrite what the preceding kernel wrote / read
        const bool stop = __ldcg(&prm.state->converged) != 0;  // uniform: the preceding kernel is complete
        if (threadIdx.x == 64) TSTAMP(4);    // preceding kernel complete
        if (MODE != 5) mbar_wait(tmem_full, 0);   // (parking the epilogue warps in a named barrier instead of this poll was measured: no difference)
        tc_fence_after();
        if (threadIdx.x == 64) TSTAMP(5);    // accumulators complete
        do {
        if (stop) break;  // converged while this kernel was streaming (PDL): leave F untouched
        float* convw = conv_s + q * 2 * KP;
        const float lambda = prm.lambda, delta = prm.delta;
        float gcd_rowmax = -1.0f;
        if (MODE == 3) {  // diagonal of P into shared memory (conv scratch is free in this mode)
            for (int i = threadIdx.x - 64; i < KP; i += 256) conv_s[i] = prm.Pfull[(size_t)i * KP + i];
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
#pragma unroll 1
        for (int c0 = chalf * (KP / 2); c0 < (chalf + 1) * (KP / 2); c0 += 32) {
            uint32_t num_u[32], den_u[32];
            float f[32];
            if (MODE != 2 && MODE != 5) tmem_ld32(t_lane + c0, num_u);
            if (MODE != 1 && MODE != 4 && MODE != 5) tmem_ld32(t_lane + KP + c0, den_u);
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
                if (MODE == 5) {  // k-split partial numerators of div_fused_kernel, summed in split order (deterministic)
                    const float* nbase = prm.num_io + (size_t)row * KP + c0;
                    float acc[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = __ldcg((const float4*)nbase + j);
                        acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
                    }
                    for (int sp = 1; sp < prm.num_splits; ++sp) {
                        const float4* ns = (const float4*)(nbase + (size_t)sp * prm.num_split_stride);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 v = __ldcg(ns + j);
                            acc[4 * j] += v.x; acc[4 * j + 1] += v.y; acc[4 * j + 2] += v.z; acc[4 * j + 3] += v.w;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) num_u[j] = __float_as_uint(acc[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { f[j] = 0.f; if (MODE == 2 || MODE == 5) num_u[j] = 0u; }
            }
            tmem_ld_wait();
            if (MODE == 3) {
                float g[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float gv = __uint_as_float(den_u[j]) - __uint_as_float(num_u[j]);       // G = F P - Z   (greedycd.jl:119-120)
                    if (lambda > 0.f) gv += lambda;                                           // :121-123
                    g[j] = gv;
                    const float prr = conv_s[c0 + j];
                    const float w = f[j];
                    const float t = w - gv / (1.1920928955078125e-07f + prr);                 // :127
                    const float sv = fmaxf(t, 0.f) - w;
                    const float dv = -gv * sv - 0.5f * prr * sv * sv;                         // :128
                    if (valid) gcd_rowmax = fmaxf(gcd_rowmax, dv);
                }
                if (valid) {
                    float4* dst = (float4*)(prm.num_io + (size_t)row * KP + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
                }
                continue;
            }
            float d2[32], s2[32];
            uint32_t hi_p[16], lo_p[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                float fn[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v;
                    if (MODE == 4 || MODE == 5) {
                        v = f[j + e] * __fdividef(__uint_as_float(num_u[j + e]), prm.colsum[c0 + j + e] + lambda);  // multupd.jl:178 / :190
                    } else {
                        float num = __uint_as_float(num_u[j + e]) - lambda;
                        num = (num > 0.f || num != num) ? num : 0.f;         // Julia max(0, x): NaN propagates
                        float den = __uint_as_float(den_u[j + e]) + delta;
                        v = f[j + e] * __fdividef(num, den);                 // multupd.jl:102 / :113 (2-ulp divide; operands are bf16-derived)
                    }
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
            if constexpr (STAGED) {
                // stage the four forms of the new tile in the (idle) ring, in the swizzled images the TMA stores expect
                const int rr = 32 * q + lane;
                uint8_t* sf = SF + (c0 >> 5) * 16384 + rr * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *(float4*)(sf + ((j ^ (rr & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                uint8_t* sh = SH + (c0 >> 6) * 16384 + rr * 128;
                uint8_t* sl = SL + (c0 >> 6) * 16384 + rr * 128;
                const int cb = (c0 & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    *(uint4*)(sh + (((cb + j) ^ (rr & 7)) << 4)) = make_uint4(hi_p[4 * j], hi_p[4 * j + 1], hi_p[4 * j + 2], hi_p[4 * j + 3]);
                    *(uint4*)(sl + (((cb + j) ^ (rr & 7)) << 4)) = make_uint4(lo_p[4 * j], lo_p[4 * j + 1], lo_p[4 * j + 2], lo_p[4 * j + 3]);
                }
                uint8_t* st = ST + (rr >> 6) * (KP * 128) + ((rr & 7) << 1);
                const int rch = (rr & 63) >> 3;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int a = c0 + j;
                    const uint32_t pk = hi_p[j / 2];
                    *(unsigned short*)(st + a * 128 + ((rch ^ (a & 7)) << 4)) = (unsigned short)((j & 1) ? (pk >> 16) : (pk & 0xffffu));
                }
            } else if (valid) {
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
        if (MODE == 3) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) gcd_rowmax = fmaxf(gcd_rowmax, __shfl_xor_sync(0xffffffffu, gcd_rowmax, o));
            asm volatile("bar.sync 1, 256;" ::: "memory");   // everybody is done reading the diagonal
            if (lane == 0) conv_s[warp - 2] = gcd_rowmax;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                float m = conv_s[0];
                for (int i = 1; i < 8; ++i) m = fmaxf(m, conv_s[i]);
                prm.conv_part[blockIdx.x] = m;
            }
        } else if (MODE != 1) {
            if constexpr (STAGED) {
                fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA / tensor-core (async) proxy
                tc_fence_before();     // our TMEM reads are complete (the Gram below reuses the Num columns)
            }
            if (threadIdx.x == 64) TSTAMP(6);  // this warp's ratio / staging done
            // combine the four lane quarters: named barrier over the 256 epilogue threads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if constexpr (STAGED) {
                if (threadIdx.x == 64) {
                    TSTAMP(7);                 // all epilogue warps done
#pragma unroll
                    for (int b = 0; b < KP / 32; ++b) tma_store_2d(&prm.tmF32, SF + b * 16384, 32 * b, r0);
#pragma unroll
                    for (int b = 0; b < KP / 64; ++b) {
                        tma_store_2d(&prm.tmFhi, SH + b * 16384, 64 * b, r0);
                        tma_store_2d(&prm.tmFlo, SL + b * 16384, 64 * b, r0);
                    }
                    tma_store_2d(&prm.tmT, ST, r0, 0);
                    if (tile_rows > 64) tma_store_2d(&prm.tmT, ST + KP * 128, r0 + 64, 0);
                    tma_store_commit();
                    if (prm.gram_part != nullptr) {  // Gram contribution of this tile: T T' (K = 128 rows), into the Num columns
                        tc_fence_after();
                        constexpr uint32_t gdesc_i = make_idesc(FMT_BF16, 128, KP);
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            const uint64_t td = make_kmajor_sw128_desc(smem_u32(ST + hb * KP * 128));
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, td + 2 * kk, td + 2 * kk, gdesc_i, (hb > 0 || kk > 0) ? 1u : 0u);
                        }
                        umma_commit(gram_bar);
                    }
                }
            }
            const int t = threadIdx.x - 64;  // 0..255
            for (int i = t; i < 2 * KP; i += 256) {
                float s = conv_s[i] + conv_s[2 * KP + i] + conv_s[4 * KP + i] + conv_s[6 * KP + i];
                prm.conv_part[(size_t)blockIdx.x * 2 * KP + i] = s;
            }
            if constexpr (STAGED) {
                if (prm.gram_part != nullptr) {
                    mbar_wait(gram_bar, 0);
                    tc_fence_after();
                    if (threadIdx.x == 64) TSTAMP(8);  // tile Gram MMAs complete
                    const int a = 32 * q + lane;
                    float* gp = prm.gram_part + ((size_t)blockIdx.x * KP + a) * KP;
#pragma unroll 1
                    for (int c0 = chalf * (KP / 2); c0 < (chalf + 1) * (KP / 2); c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + c0, v);
                        tmem_ld_wait();
                        if (a < KP) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                ((float4*)(gp + c0))[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        }
                    }
                }
                // the staging buffers must stay valid until the bulk stores have drained (waiting only for the smem reads,
                // .read, measured the same)
                if (threadIdx.x == 64) {
                    TSTAMP(9);                 // tile Gram written
                    tma_store_wait_all<0>();
                    TSTAMP(10);                // bulk stores have read their staging buffers
                }
            }
        }
        } while (0);
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) TSTAMP(11);
    if (prm.timing != nullptr && threadIdx.x == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.timing[16 + 2 * blockIdx.x + 1] = gt;
    }
#undef TSTAMP
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
        if (elect_one()) {
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
        if (elect_one()) {
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

// P[e] = sum_g part[g][e]; writes the fp32 Gram and (do_split) its bf16 hi/lo split.  Four lanes per element
// (each sums every 4th partial with 8 loads in flight), combined with two shuffles: fixed order => deterministic.
__global__ void __launch_bounds__(256) gram_reduce_kernel(const float* __restrict__ part, int nparts, int nelem, float* __restrict__ P,
                                                          bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split,
                                                          const TcState* st) {
    // The update kernel behind us may start streaming X as soon as every block has passed this point; it waits for our
    // completion before it reads P.  (Pre-launching THIS kernel behind the running update kernel was measured too:
    // its resident blocks polling in griddepcontrol.wait slow the single-thread TMA / MMA loops, 4770 -> 4400 it/s.)
    pdl_launch_dependents();
    if (st->converged) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = t & 3;
    const int i = t >> 2;
    float acc = 0.f;
    if (i < nelem) {
        int g = sub;
        for (; g + 28 < nparts; g += 32) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (size_t)(g + 4 * u) * nelem + i);
            acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
        for (; g < nparts; g += 4) acc += __ldcg(part + (size_t)g * nelem + i);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (i < nelem && sub == 0) {
        P[i] = acc;
        if (do_split) {
            bf16 hi = __float2bfloat16_rn(acc);
            Phi[i] = hi;
            Plo[i] = __float2bfloat16_rn(acc - __bfloat162float(hi));
        }
    }
}

// ---- exchange kernels -----------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Wait until every rank has published `epoch` in this rank's flag row of `phase` (local memory poll, bounded).
__device__ __forceinline__ void xchg_wait_all(const XchgDev& x, int phase, unsigned int epoch) {
    if ((int)threadIdx.x < x.G) {
        const unsigned int* f = x.flags[x.rank] + phase * XCHG_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if (clock64() - t0 > 20000000000LL) {  // ~10 s
                printf("nmfb200: peer barrier timed out (rank %d waiting for %d, phase %d, epoch %u)\n", x.rank, (int)threadIdx.x, phase, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void xchg_signal_all(const XchgDev& x, int phase, unsigned int epoch) {  // call from < G threads
    if ((int)threadIdx.x < x.G) {
        __threadfence_system();
        st_release_sys(x.flags[threadIdx.x] + phase * XCHG_MAX_RANKS + x.rank, epoch);
    }
}

// Fused reduce-scatter + all-gather over NVLink peer memory, one launch:
//   (1) block 0 publishes "my partial sums are complete" (true by stream order: the producing kernels ran before);
//   (2) every block waits for all ranks, then pulls its share of this rank's segment from all ranks' packed vectors
//       (coalesced 16-byte peer loads), sums in rank order (=> bit-identical on every rank) and pushes the result
//       into every rank's packed vector (peer stores);
//   (3) the last block to finish publishes "my reduced segment is in place" (phase 1); consumers wait on that.
__global__ void __launch_bounds__(256) xchg_reduce_gather_kernel(XchgDev x, int KP, unsigned int epoch) {
    __shared__ int is_last;
    if (blockIdx.x == 0) xchg_signal_all(x, 0, epoch);
    xchg_wait_all(x, 0, epoch);
    const size_t seg4 = (size_t)x.RS * KP / 4;  // float4 elements per segment
    const size_t off = (size_t)x.rank * seg4;
    // U independent elements per thread and trip: G*U 16-byte peer loads in flight hide the ~2-3 us NVLink latency
    constexpr int U = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < seg4; i0 += U * stride) {
        float4 v[U][XCHG_MAX_RANKS];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
#pragma unroll
            for (int src = 0; src < XCHG_MAX_RANKS; ++src)
                if (src < x.G && i < seg4) v[u][src] = __ldcg((const float4*)x.packed[src] + off + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * stride;
            if (i >= seg4) break;
            float4 s = v[u][0];
#pragma unroll
            for (int src = 1; src < XCHG_MAX_RANKS; ++src)
                if (src < x.G) { s.x += v[u][src].x; s.y += v[u][src].y; s.z += v[u][src].z; s.w += v[u][src].w; }
#pragma unroll
            for (int dst = 0; dst < XCHG_MAX_RANKS; ++dst)
                if (dst < x.G) ((float4*)x.packed[dst])[off + i] = s;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(x.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (is_last) {
        if (threadIdx.x == 0) *x.ticket = 0u;
        xchg_signal_all(x, 1, epoch);
    }
}

// ---- stop_condition finish (common.jl:92-111) ---------------------------------------------------------
// acc (double [4][KP]) = {dev_w, sum_w, dev_h, sum_h}.  conv_reduce_kernel: grid = 4 * KP/32 blocks of 8 warps;
// block (q, cb) sums quantity q of components [32cb, 32cb+32) over all tiles (warp w takes tiles w, w+8, ...;
// 8 loads in flight; fixed combination order => deterministic).  With do_decide the last block to finish
// (atomic ticket) applies the reference's test; multi-GPU runs the decision as a separate launch after the
// packed all-reduce (post_allreduce_kernel).
__device__ void conv_decide(const double* acc, int KP, int k, float tol, TcState* st, float* devs, int* fail) {
    const int a = threadIdx.x;
    if (a == 0) *fail = 0;
    __syncthreads();
    float dev = 0.f;
    if (a < k) {
        float dw = (float)__ldcg(acc + a), sw = (float)__ldcg(acc + KP + a), dh = (float)__ldcg(acc + 2 * KP + a),
              sh = (float)__ldcg(acc + 3 * KP + a);
        float rw = dw / sw, rh = dh / sh;
        float m = (rw != rw) ? rw : ((rh != rh) ? rh : fmaxf(rw, rh));  // Julia max(): NaN propagates (common.jl:105)
        dev = sqrtf(m);
        if (sqrtf(dw) > tol * sqrtf(sw) || sqrtf(dh) > tol * sqrtf(sh)) atomicExch(fail, 1);  // common.jl:106
    }
    if (a < 256) devs[a] = dev;
    __syncthreads();
    if (a == 0) {
        float dm = 0.f;
        for (int i = 0; i < k; ++i) dm = (dm != dm) ? dm : ((devs[i] != devs[i]) ? devs[i] : fmaxf(dm, devs[i]));
        st->devmax = dm;
        st->iters += 1;
        if (!*fail) st->converged = 1;
    }
}

__global__ void __launch_bounds__(256) conv_reduce_kernel(const float* __restrict__ partW, int tilesW, const float* __restrict__ partH,
                                                          int tilesH, int KP, int k, int update_H, double* __restrict__ acc, float tol,
                                                          TcState* st, int do_decide, float* __restrict__ wsums_f32) {
    if (st->converged) return;
    __shared__ double red[8][32];
    __shared__ float devs[256];
    __shared__ int fail, is_last;
    const int cbs = KP / 32;
    const int q = blockIdx.x / cbs, cb = blockIdx.x % cbs;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = cb * 32 + lane;
    const float* part = (q < 2 ? partW : partH) + (size_t)(q & 1) * KP + c;
    const int tiles = q < 2 ? tilesW : (update_H ? tilesH : 0);
    double s = 0.0;
    int t = w;
    for (; t + 56 < tiles; t += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (size_t)(t + 8 * u) * 2 * KP);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; t < tiles; t += 8) s += (double)__ldcg(part + (size_t)t * 2 * KP);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0) {
        double tot = red[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) tot += red[i][lane];
        if (q >= 2 && !update_H) tot = (q == 2) ? 0.0 : 1.0;  // H untouched: dev_h = 0 (sum_h only scales a ratio of 0)
        acc[(size_t)q * KP + c] = tot;
        if (wsums_f32 && q < 2) wsums_f32[(size_t)q * KP + c] = (float)tot;  // multi-GPU: rides in the packed all-reduce
    }
    if (!do_decide) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(&st->ticket, 1u);
        is_last = (prev == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) st->ticket = 0u;
    conv_decide(acc, KP, k, tol, st, devs, &fail);
}

// One launch for the two small reductions that follow the W-step: blocks [0, gram_blocks) reduce the per-tile Gram
// contributions (gram_reduce_kernel's work), the remaining 4*KP/32 blocks reduce the stop_condition partial sums and the
// last of them decides (conv_reduce_kernel's work).
__device__ __forceinline__ void gram_reduce_body(const float* __restrict__ part, int nparts, int nelem, float* __restrict__ P,
                                                 bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split, int block) {
    const int t = block * blockDim.x + threadIdx.x;
    const int sub = t & 3;
    const int i = t >> 2;
    float acc = 0.f;
    if (i < nelem) {
        int g = sub;
        for (; g + 28 < nparts; g += 32) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (size_t)(g + 4 * u) * nelem + i);
            acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
        for (; g < nparts; g += 4) acc += __ldcg(part + (size_t)g * nelem + i);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (i < nelem && sub == 0) {
        P[i] = acc;
        if (do_split) {
            bf16 hi = __float2bfloat16_rn(acc);
            Phi[i] = hi;
            Plo[i] = __float2bfloat16_rn(acc - __bfloat162float(hi));
        }
    }
}

__global__ void __launch_bounds__(256) gram_conv_reduce_kernel(const float* __restrict__ gpart, int nparts, int nelem, float* __restrict__ P,
                                                               bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split, int gram_blocks,
                                                               const float* __restrict__ partW, int tilesW, const float* __restrict__ partH,
                                                               int tilesH, int KP, int k, int update_H, double* __restrict__ acc, float tol,
                                                               TcState* st, int do_decide, float* __restrict__ wsums_f32) {
    pdl_launch_dependents();  // the next H-step may start streaming X now; it waits for us before it reads P / `converged`
    if (st->converged) return;
    if ((int)blockIdx.x < gram_blocks) {
        gram_reduce_body(gpart, nparts, nelem, P, Phi, Plo, do_split, blockIdx.x);
        return;
    }
    __shared__ double red[8][32];
    __shared__ float devs[256];
    __shared__ int fail, is_last;
    const int cblock = blockIdx.x - gram_blocks, nconv = gridDim.x - gram_blocks;
    const int cbs = KP / 32;
    const int q = cblock / cbs, cb = cblock % cbs;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = cb * 32 + lane;
    const float* part = (q < 2 ? partW : partH) + (size_t)(q & 1) * KP + c;
    const int tiles = q < 2 ? tilesW : (update_H ? tilesH : 0);
    double s = 0.0;
    int t = w;
    for (; t + 56 < tiles; t += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(part + (size_t)(t + 8 * u) * 2 * KP);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; t < tiles; t += 8) s += (double)__ldcg(part + (size_t)t * 2 * KP);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0) {
        double tot = red[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) tot += red[i][lane];
        if (q >= 2 && !update_H) tot = (q == 2) ? 0.0 : 1.0;
        acc[(size_t)q * KP + c] = tot;
        if (wsums_f32 && q < 2) wsums_f32[(size_t)q * KP + c] = (float)tot;
    }
    if (!do_decide) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int prev = atomicAdd(&st->ticket, 1u);
        is_last = (prev == (unsigned)nconv - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) st->ticket = 0u;
    conv_decide(acc, KP, k, tol, st, devs, &fail);
}

// multi-GPU, after the packed all-reduce: block 0 finishes stop_condition of the PREVIOUS iteration (its W-side
// sums travelled in the tail of the packed buffer; nothing of the current iteration has touched W or H yet),
// the other blocks split the reduced Gram W'W into bf16 hi/lo.
__global__ void __launch_bounds__(256) post_allreduce_kernel(double* __restrict__ acc, const float* __restrict__ wsums_f32, int has_prev,
                                                             int KP, int k, float tol, TcState* st, const float* __restrict__ P,
                                                             bf16* __restrict__ Phi, bf16* __restrict__ Plo, XchgDev x, unsigned int epoch) {
    pdl_launch_dependents();  // the MODE 2 ratio kernel may set itself up now; it waits for our completion before it reads anything
    if (x.G > 0) xchg_wait_all(x, 1, epoch);  // peer-memory exchange: every rank's reduced segment has landed here
    if (st->converged) return;
    __shared__ float devs[256];
    __shared__ int fail;
    if (blockIdx.x == 0) {
        if (!has_prev) return;
        for (int i = threadIdx.x; i < 2 * KP; i += blockDim.x) acc[i] = (double)__ldcg(wsums_f32 + i);
        __syncthreads();
        conv_decide(acc, KP, k, tol, st, devs, &fail);
        return;
    }
    if (P == nullptr) return;
    const int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x;
    if (i < KP * KP) {
        float v = P[i];
        bf16 hi = __float2bfloat16_rn(v);
        Phi[i] = hi;
        Plo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// ---- X caches: bf16, TILE-CONTIGUOUS ---------------------------------------------------------------------
// A panel with R rows and contraction length Kdim is stored as [tile][kb][TR rows][64 cols] (TR = rows per
// CTA, kb = 64-wide k-block): one TMA box = one contiguous TR*128-byte burst and a CTA streams one
// sequential region of HBM (instead of gathering 128-B segments from TR rows a full row pitch apart).
// Padding rows / columns are written as zeros.  element(r, c) = X[r*sr + c*sc].
// (a) contraction index contiguous in the source (sc == 1): direct
__global__ void cvt_tiled_direct_kernel(const float* __restrict__ X, int64_t sr, int R, int Kdim, int TR, int nkb,
                                        bf16* __restrict__ dst) {
    const int64_t row_slot = blockIdx.x;            // tile * TR + row-in-tile (x: up to 2^31-1 rows)
    const int tile = (int)(row_slot / TR), rr = (int)(row_slot % TR);
    const int64_t r = (int64_t)tile * TR + rr;
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < nkb * 64; c += gridDim.y * blockDim.x) {
        float v = (r < R && c < Kdim) ? X[r * sr + c] : 0.f;
        const int kb = c >> 6, cc = c & 63;
        dst[(((int64_t)tile * nkb + kb) * TR + rr) * 64 + cc] = __float2bfloat16_rn(v);
    }
}
// (b) row index contiguous in the source (sr == 1): 64 x 64 transpose through shared memory
__global__ void cvt_tiled_transpose_kernel(const float* __restrict__ X, int64_t sc, int R, int Kdim, int TR, int nkb, int tiles,
                                           bf16* __restrict__ dst) {
    __shared__ float tile_s[64][65];
    const int kb = blockIdx.x;
    const int64_t r_base = (int64_t)blockIdx.y * 64;   // 64 consecutive logical rows
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {  // y: column within the k-block, x: row (contiguous in X)
        int64_t r = r_base + threadIdx.x * 2;
        int c = kb * 64 + y;
        float v0 = (r < R && c < Kdim) ? X[r + (int64_t)c * sc] : 0.f;
        float v1 = (r + 1 < R && c < Kdim) ? X[r + 1 + (int64_t)c * sc] : 0.f;
        tile_s[threadIdx.x * 2][y] = v0;
        tile_s[threadIdx.x * 2 + 1][y] = v1;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 64; y += blockDim.y) {  // y: row within the 64-row group, x: column pair
        int64_t r = r_base + y;
        const int tile = (int)(r / TR), rr = (int)(r % TR);
        if (tile >= tiles) continue;  // the grid is rounded up to 64-row groups
        __nv_bfloat162 pk = __floats2bfloat162_rn(tile_s[y][threadIdx.x * 2], tile_s[y][threadIdx.x * 2 + 1]);
        *(__nv_bfloat162*)(dst + (((int64_t)tile * nkb + kb) * TR + rr) * 64 + threadIdx.x * 2) = pk;
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

// fp32 matrix [rows][inner] with row pitch ld (elements); box = 32 (128 B) x box_rows, SWIZZLE_128B
CUtensorMap make_tmap_f32(const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NMF_REQUIRE(r == CUDA_SUCCESS, NMFB200_ECUDA, "cuTensorMapEncodeTiled(f32) failed with code " + std::to_string((int)r));
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
    float* colsum = nullptr;  // [KP] column sums (MultUpdate :div)
    int tiles = 0;
    int tile_rows = 128;
};

// Launch with or without the programmatic-stream-serialization attribute (PDL).
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    NMF_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

template <int KP>
struct TcSolver {
    nmfb200_handle* h;
    cudaStream_t st;
    TcState* state;
    bool defer_gram_reduce = false;  // the caller will run gram_conv_reduce_kernel itself
    bool last_fused_gram = false;
    float* last_gram_part = nullptr;
    int num_splits = 1;              // MODE 5: k-split partial numerators behind num_io
    int64_t num_split_stride = 0;

    // gram: -1 = no Gram of the updated factor wanted; 0 / 1 = wanted, without / with the bf16 hi-lo split;
    // gram_dst = where the fp32 Gram goes (default F.P).  KP <= 128: the update kernel's staged epilogue produces the
    // per-tile contributions itself (only a reduce launch follows); KP = 256: separate gram_kernel pass.
    void launch_update(int mode, const Factor& F, const Factor& O, const bf16* Xs, int Kdim, float lambda, float delta,
                       float* num_io, float* conv_override = nullptr, int gram = -1, float* gram_dst = nullptr, bool pdl = false) {
        UpdateParams prm;
        const bool fused_gram = gram >= 0 && KP <= 128 && (mode == 0 || mode == 2);
        prm.gram_part = fused_gram ? h->buf_t<float>("tc.gram_part", (size_t)std::max(F.tiles, 1) * KP * KP) : nullptr;
        prm.tmF32 = make_tmap_f32(F.m, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmT = make_tmap_bf16(F.bT, (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, KP);
        prm.tile_rows = F.tile_rows;
        prm.timing = (h->tc_debug & 8) ? (long long*)h->buf("tc.timing", (16 + 2 * 4096) * sizeof(long long)) : nullptr;
        const uint64_t nkb = (uint64_t)ceil_div(Kdim, 64);
        prm.tmA = make_tmap_bf16(Xs, 64, (uint64_t)F.tiles * nkb * F.tile_rows, 64, (uint32_t)F.tile_rows);
        prm.tmB = make_tmap_bf16(O.bT, (uint64_t)Kdim, KP, (uint64_t)O.ldT, KP);
        prm.tmFhi = make_tmap_bf16(F.hi, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmFlo = make_tmap_bf16(F.lo, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmPhi = make_tmap_bf16(O.Phi, KP, KP, KP, KP);
        prm.tmPlo = make_tmap_bf16(O.Plo, KP, KP, KP, KP);
        prm.F = F.m; prm.Fhi = F.hi; prm.Flo = F.lo; prm.FbT = F.bT; prm.ldT = F.ldT;
        prm.num_io = num_io;
        prm.num_splits = num_splits;
        prm.num_split_stride = num_split_stride;
        prm.conv_part = conv_override ? conv_override : F.conv;
        prm.Pfull = O.P;
        prm.colsum = O.colsum;
        prm.state = state;
        prm.R = F.R; prm.Kdim = Kdim; prm.lambda = lambda; prm.delta = delta;
        const int smem = UpdCfg<KP>::SMEM_BYTES;
        const bool timed = h->time_kernels == 1 && mode != 2;
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        void (*kern)(const UpdateParams) = mode == 0   ? mu_update_kernel<KP, 0>
                                           : mode == 1 ? mu_update_kernel<KP, 1>
                                           : mode == 2 ? mu_update_kernel<KP, 2>
                                           : mode == 3 ? mu_update_kernel<KP, 3>
                                           : mode == 4 ? mu_update_kernel<KP, 4>
                                                       : mu_update_kernel<KP, 5>;
        // pdl: programmatic dependent launch -- start streaming X while the preceding reduce kernel is still running
        launch_k(kern, dim3((unsigned)F.tiles), dim3(UpdCfg<KP>::THREADS), (size_t)smem, st, pdl, prm);
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        h->launches += 1;
        last_fused_gram = fused_gram;
        last_gram_part = prm.gram_part;
        if (fused_gram && !defer_gram_reduce) {
            launch_k(gram_reduce_kernel, dim3((4 * KP * KP + 255) / 256), dim3(256), 0, st, false, (const float*)prm.gram_part, F.tiles, KP * KP,
                     gram_dst ? gram_dst : F.P, F.Phi, F.Plo, gram, (const TcState*)state);
            h->launches += 1;
        } else if (gram >= 0 && !fused_gram) {
            launch_gram(F, gram != 0, gram_dst);
        }
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
        gram_reduce_kernel<<<(4 * KP * KP + 255) / 256, 256, 0, st>>>(g.part, grid, KP * KP, P_dst ? P_dst : F.P, F.Phi, F.Plo, split ? 1 : 0, state);
        h->launches += 2;
    }

    static void set_attrs() {
        static bool done = false;
        if (done) return;
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(gram_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg<KP>::SMEM_BYTES));
        done = true;
    }
};

// ---- peer-memory exchange arena: allocate, export through CUDA IPC, import every peer's ------------------
void xchg_teardown(nmfb200_handle* h) {
    Xchg& x = h->xchg;
    for (int j = 0; j < XCHG_MAX_RANKS; ++j) {
        if (x.arena_peer[j] && j != x.rank) cudaIpcCloseMemHandle(x.arena_peer[j]);
        x.arena_peer[j] = nullptr;
    }
    if (x.arena_local) cudaFree(x.arena_local);
    x.arena_local = nullptr;
    x.ready = false;
}

constexpr size_t XCHG_FLAG_BYTES = 256;

// Collective over the communicator: every rank calls it with the same (n, KP).  Returns false (and leaves the
// NCCL path in charge) if peer mapping is not possible on this machine.
bool xchg_setup(nmfb200_handle* h, int64_t n, int KP, XchgDev* out) {
    Xchg& x = h->xchg;
    const int G = h->nranks;
    if (!h->tc_xchg || G > XCHG_MAX_RANKS || G < 2) return false;
    const size_t rtot = (size_t)n + KP + 2;
    const size_t RS = (rtot + G - 1) / G;
    if (!(x.ready && x.G == G && x.rank == h->rank && x.rows_per_seg == RS && x.row_floats == (size_t)KP)) {
        xchg_teardown(h);
        x.G = G;
        x.rank = h->rank;
        x.rows_per_seg = RS;
        x.row_floats = (size_t)KP;
        const size_t region = (size_t)G * RS * KP * sizeof(float);
        x.arena_bytes = XCHG_FLAG_BYTES + region;
        NMF_CUDA(cudaMalloc(&x.arena_local, x.arena_bytes));
        NMF_CUDA(cudaMemsetAsync(x.arena_local, 0, x.arena_bytes, h->stream));
        cudaIpcMemHandle_t mine;
        NMF_CUDA(cudaIpcGetMemHandle(&mine, x.arena_local));
        char* dsend = (char*)h->buf("tc.xchg_ipc_send", sizeof(mine));
        char* drecv = (char*)h->buf("tc.xchg_ipc_recv", sizeof(mine) * XCHG_MAX_RANKS);
        NMF_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
        NMF_NCCL(NcclApi::get().AllGather(dsend, drecv, sizeof(mine), ncclChar, h->comm, h->stream));
        std::vector<cudaIpcMemHandle_t> all(G);
        NMF_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(mine) * G, cudaMemcpyDeviceToHost, h->stream));
        NMF_CUDA(cudaStreamSynchronize(h->stream));
        int ok = 1;
        for (int j = 0; j < G; ++j) {
            if (j == x.rank) {
                x.arena_peer[j] = x.arena_local;
                continue;
            }
            void* p = nullptr;
            cudaError_t err = cudaIpcOpenMemHandle(&p, all[j], cudaIpcMemLazyEnablePeerAccess);
            if (err != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            x.arena_peer[j] = p;
        }
        // agree on the outcome: everybody falls back to NCCL if anybody could not map a peer
        int* dok = (int*)h->buf("tc.xchg_ok", sizeof(int));
        NMF_CUDA(cudaMemcpyAsync(dok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        NMF_NCCL(NcclApi::get().AllReduce(dok, dok, 1, ncclInt32, ncclMin, h->comm, h->stream));
        NMF_CUDA(cudaMemcpyAsync(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        NMF_CUDA(cudaStreamSynchronize(h->stream));
        if (!ok) {
            xchg_teardown(h);
            h->tc_xchg = 0;
            return false;
        }
        x.epoch = 0;
        x.ready = true;
    }
    out->G = G;
    out->rank = x.rank;
    out->RS = (int)RS;
    out->ticket = (unsigned int*)((char*)x.arena_local + 128);
    for (int j = 0; j < XCHG_MAX_RANKS; ++j) {
        char* base = (char*)x.arena_peer[j];
        out->flags[j] = base ? (unsigned int*)base : nullptr;
        out->packed[j] = base ? (float*)(base + XCHG_FLAG_BYTES) : nullptr;
    }
    return true;
}

// Rows of a factor per CTA of the update kernel (multiple of 8, <= 128).
int pick_tile_rows(int R, int forced) {
    if (forced >= 8 && forced <= 128 && forced % 8 == 0) return forced;
    // Measured (profiles/r1b_pipeline_experiments.md): the time per k-block does not shrink with the box height,
    // so full 128-row boxes always win, even when that leaves SMs idle (128 CTAs at 16384 rows).
    return R >= 128 ? 128 : (int)round_up(std::max(R, 8), 8);
}

Factor alloc_factor(nmfb200_handle* h, const char* tag, int R, int KP) {
    Factor f;
    std::string t(tag);
    f.R = R;
    f.ldT = round_up(R, 64);
    f.tile_rows = pick_tile_rows(R, h->tc_tile_rows);
    f.tiles = (int)ceil_div(R, f.tile_rows);
    f.m = h->buf_t<float>("tc." + t + ".m", (size_t)R * KP);
    f.hi = h->buf_t<bf16>("tc." + t + ".hi", (size_t)R * KP);
    f.lo = h->buf_t<bf16>("tc." + t + ".lo", (size_t)R * KP);
    f.rowsT = KP < 128 ? 128 : KP;  // gram_kernel loads 128-row M tiles: keep zero rows behind KP = 64
    f.bT = h->buf_t<bf16>("tc." + t + ".bT", (size_t)f.rowsT * f.ldT);
    f.P = h->buf_t<float>("tc." + t + ".P", (size_t)KP * KP);
    f.Phi = h->buf_t<bf16>("tc." + t + ".Phi", (size_t)KP * KP);
    f.Plo = h->buf_t<bf16>("tc." + t + ".Plo", (size_t)KP * KP);
    f.conv = h->buf_t<float>("tc." + t + ".conv", (size_t)f.tiles * 2 * KP);
    f.colsum = h->buf_t<float>("tc." + t + ".colsum", (size_t)KP);
    return f;
}

// ---- objective on tensor cores (multupd.jl:81,148; greedycd.jl:84) --------------------------------------------------
// 0.5*||X - WH||^2 or gkldiv(X, WH) without materialising WH: same pipeline as the quotient kernel, but the X tile
// is the caller's fp32 X (TMA, 2 boxes of 128 x 32 fp32 per 128 x 64 tile), WH = Rf*Cf' uses the bf16 hi/lo split
// of both factors (hi*hi + hi*lo + lo*hi, ~2^-17 relative; KP = 256: hi only, smem) and the epilogue reduces in fp64
// (StatsBase semantics: per-element terms in fp32, Float64 accumulator).  Rows = columns of X (j), k-blocks over i.
struct ObjParams {
    CUtensorMap tmX;    // X fp32 [n][p] (column-major p x n), row pitch ldx, box 32 x 128
    CUtensorMap tmRhi, tmRlo;   // H hi/lo bf16 [n][KP], box 64 x 128
    CUtensorMap tmChi, tmClo;   // W hi/lo bf16 [p][KP], box 64 x 64
    double* part;       // [gridDim.x * gridDim.y] partial sums
    int nkb, kchunk;
};

template <int KP>
struct ObjCfg {
    static constexpr bool SPLIT = KP <= 128;
    static constexpr int NT = SPLIT ? 2 : 1;             // hi (+ lo) copies
    static constexpr int NSLAB = KP / 64;
    static constexpr int RF_BYTES = NT * NSLAB * 128 * 128;
    static constexpr int C_BYTES = NT * NSLAB * 64 * 128;
    static constexpr int X_BYTES = 2 * 128 * 128;        // 128 rows x 64 fp32
    static constexpr int SC = 2, SX = 2;
    static constexpr int OFF_C = RF_BYTES;
    static constexpr int OFF_X = OFF_C + SC * C_BYTES;
    static constexpr int OFF_BAR = OFF_X + SX * X_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static constexpr int THREADS = 320;
    static constexpr int TMEM_COLS = 128;
};

template <int KP, int KL>
__global__ void __launch_bounds__(ObjCfg<KP>::THREADS, 1) objective_tc_kernel(const __grid_constant__ ObjParams prm) {
    using C = ObjCfg<KP>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* fullC = (uint64_t*)(smem + C::OFF_BAR);
    uint64_t* emptyC = fullC + C::SC;
    uint64_t* fullX = emptyC + C::SC;
    uint64_t* emptyX = fullX + C::SX;
    uint64_t* tfull = emptyX + C::SX;
    uint64_t* tempty = tfull + 2;
    uint64_t* rf_full = tempty + 2;
    uint32_t* tmem_slot = (uint32_t*)(rf_full + 1);
    double* red = (double*)(tmem_slot + 2);              // [8 warps]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = blockIdx.y * prm.kchunk;
    const int nkb = min(prm.nkb, kb0 + prm.kchunk) - kb0;
    const int row0 = blockIdx.x * 128;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&prm.tmX);
        prefetch_tmap(&prm.tmRhi);
        prefetch_tmap(&prm.tmChi);
        for (int s = 0; s < C::SC; ++s) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], 1); }
        for (int s = 0; s < C::SX; ++s) { mbar_init(&fullX[s], 1); mbar_init(&emptyX[s], 8); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
        mbar_init(rf_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(rf_full, C::RF_BYTES);
            for (int t = 0; t < C::NT; ++t)
                for (int sl = 0; sl < C::NSLAB; ++sl)
                    tma_load_2d(smem + (t * C::NSLAB + sl) * 128 * 128, t ? &prm.tmRlo : &prm.tmRhi, rf_full, 64 * sl, row0);
            int sc = 0, sx = 0;
            uint32_t phc = 0, phx = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&emptyX[sx], phx ^ 1u);
                mbar_arrive_expect_tx(&fullX[sx], C::X_BYTES);
                for (int b = 0; b < 2; ++b)
                    tma_load_2d(smem + C::OFF_X + sx * C::X_BYTES + b * 128 * 128, &prm.tmX, &fullX[sx], 64 * (kb0 + kb) + 32 * b, row0);
                mbar_wait(&emptyC[sc], phc ^ 1u);
                mbar_arrive_expect_tx(&fullC[sc], C::C_BYTES);
                for (int t = 0; t < C::NT; ++t)
                    for (int sl = 0; sl < C::NSLAB; ++sl)
                        tma_load_2d(smem + C::OFF_C + sc * C::C_BYTES + (t * C::NSLAB + sl) * 64 * 128, t ? &prm.tmClo : &prm.tmChi,
                                    &fullC[sc], 64 * sl, 64 * (kb0 + kb));
                if (++sx == C::SX) { sx = 0; phx ^= 1u; }
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, 64);
            mbar_wait(rf_full, 0);
            int sc = 0;
            uint32_t phc = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int b = kb & 1;
                mbar_wait(&tempty[b], (((uint32_t)kb >> 1) & 1u) ^ 1u);
                mbar_wait(&fullC[sc], phc);
                tc_fence_after();
                const uint32_t cbase = smem_u32(smem + C::OFF_C + sc * C::C_BYTES);
                bool first = true;
                // terms: (R hi, C hi), (R hi, C lo), (R lo, C hi)
#pragma unroll
                for (int term = 0; term < (C::SPLIT ? 3 : 1); ++term) {
                    const int tr = term == 2 ? 1 : 0, tc = term == 1 ? 1 : 0;
#pragma unroll
                    for (int sl = 0; sl < C::NSLAB; ++sl) {
                        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + (tr * C::NSLAB + sl) * 128 * 128));
                        const uint64_t bdesc = make_kmajor_sw128_desc(cbase + (tc * C::NSLAB + sl) * 64 * 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            umma_bf16(tmem_base + b * 64, adesc + 2 * kk, bdesc + 2 * kk, idesc, first ? 0u : 1u);
                            first = false;
                        }
                    }
                }
                umma_commit(&emptyC[sc]);
                umma_commit(&tfull[b]);
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int q = warp & 3, hf = e >> 2;
        const int r = 32 * q + lane;
        double acc = 0.0;
        int sx = 0;
        uint32_t phx = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1;
            mbar_wait(&tfull[b], ((uint32_t)kb >> 1) & 1u);
            tc_fence_after();
            uint32_t d[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + b * 64 + 32 * hf, d);
            mbar_wait(&fullX[sx], phx);
            const uint8_t* xt = smem + C::OFF_X + sx * C::X_BYTES + hf * 128 * 128 + r * 128;
            float4 xv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) xv[c] = *(const float4*)(xt + ((c ^ (r & 7)) << 4));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);   // TMEM buffer b may be overwritten
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float xs[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float x = xs[w], y = __uint_as_float(d[4 * c + w]);
                    if (KL) {
                        part += (x > 0.f) ? (x * logf(x / y) - x + y) : y;   // gkldiv term
                    } else {
                        const float df = x - y;
                        part += df * df;                                      // sqL2dist term
                    }
                }
            }
            acc += (double)part;   // 32 fp32 terms per step, then Float64 (StatsBase accumulates in Float64)
            // Release the X stage only now that its values have been CONSUMED: the shared-memory loads above are
            // asynchronous, and an arrive issued right behind them let the producer's TMA overwrite the stage while
            // they were still in flight (seen as a run-to-run wobble of ~1e-5 in the objective with KP = 64).
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyX[sx]);
            if (++sx == C::SX) { sx = 0; phx ^= 1u; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) red[e] = acc;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) {
            double tot = 0.0;
            for (int i = 0; i < 8; ++i) tot += red[i];
            prm.part[blockIdx.y * gridDim.x + blockIdx.x] = tot;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

__global__ void sum_double_kernel(const double* __restrict__ part, int n, double* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}
__global__ void abs_sum_kernel(const float* __restrict__ a, int64_t len, double* __restrict__ part) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += (int64_t)gridDim.x * 256) s += (double)fabsf(a[i]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// Objective of the current factors (row-factor layout with fresh hi/lo copies) on tensor cores.  Returns false if
// the shape is not covered (caller falls back to the exact fp32 GEMM + reduction of the SIMT engine).
template <int KP>
bool tc_objective(nmfb200_handle* h, int alg, const Factor& W, const Factor& H, double lambda_w, double lambda_h, double* out) {
    const int64_t p = h->p, n = h->n;
    if (n < 128 || p < 64 || (h->ldx % 4) != 0 || (((uintptr_t)h->dX) & 15) != 0) return false;
    cudaStream_t st = h->stream;
    static bool attr = false;
    if (!attr) {
        NMF_CUDA(cudaFuncSetAttribute(objective_tc_kernel<KP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ObjCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(objective_tc_kernel<KP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ObjCfg<KP>::SMEM_BYTES));
        attr = true;
    }
    ObjParams op;
    op.tmX = make_tmap_f32(h->dX, (uint64_t)p, (uint64_t)n, (uint64_t)h->ldx, 128);
    op.tmRhi = make_tmap_bf16(H.hi, KP, (uint64_t)n, KP, 128);
    op.tmRlo = make_tmap_bf16(H.lo, KP, (uint64_t)n, KP, 128);
    op.tmChi = make_tmap_bf16(W.hi, KP, (uint64_t)p, KP, 64);
    op.tmClo = make_tmap_bf16(W.lo, KP, (uint64_t)p, KP, 64);
    const int tiles = (int)ceil_div(n, 128);
    op.nkb = (int)ceil_div(p, 64);
    int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(296, tiles), op.nkb / 16));
    op.kchunk = (int)ceil_div(op.nkb, ksplit);
    ksplit = (int)ceil_div(op.nkb, op.kchunk);
    const int nparts = tiles * ksplit;
    double* part = h->buf_t<double>("tc.obj_part", (size_t)nparts + 2048 + 4);
    op.part = part;
    if (alg == 1) objective_tc_kernel<KP, 1><<<dim3(tiles, ksplit), ObjCfg<KP>::THREADS, ObjCfg<KP>::SMEM_BYTES, st>>>(op);
    else objective_tc_kernel<KP, 0><<<dim3(tiles, ksplit), ObjCfg<KP>::THREADS, ObjCfg<KP>::SMEM_BYTES, st>>>(op);
    double* res = part + nparts;  // [0] data term, [1] |W|_1, [2] |H|_1
    sum_double_kernel<<<1, 256, 0, st>>>(part, nparts, res);
    h->launches += 2;
    const bool l1w = alg == 2 && lambda_w > 0, l1h = alg == 2 && lambda_h > 0;
    double* scratch = res + 4;
    if (l1w) {
        abs_sum_kernel<<<1024, 256, 0, st>>>(W.m, (int64_t)W.R * KP, scratch);
        sum_double_kernel<<<1, 256, 0, st>>>(scratch, 1024, res + 1);
        h->launches += 2;
    }
    if (l1h) {
        abs_sum_kernel<<<1024, 256, 0, st>>>(H.m, (int64_t)H.R * KP, scratch + 1024);
        sum_double_kernel<<<1, 256, 0, st>>>(scratch + 1024, 1024, res + 2);
        h->launches += 2;
    }
    NMF_CUDA(cudaGetLastError());
    if (h->comm) {  // rows of X / W are sharded: the data term and |W|_1 are partial sums; H is replicated
        h->allreduce_sum(res, 2);
    }
    double hres[3] = {0, 0, 0};
    NMF_CUDA(cudaMemcpyAsync(hres, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    if (alg == 1) {
        *out = (double)(float)hres[0];                       // gkldiv returns Float64; Result{T} converts (common.jl:32)
    } else {
        float r = 0.5f * (float)hres[0];                     // convert(T, 0.5) * sqL2dist (multupd.jl:81)
        if (l1w) r = r + (float)lambda_w * (float)hres[1];   // greedycd.jl:85-90
        if (l1h) r = r + (float)lambda_h * (float)hres[2];
        *out = (double)r;
    }
    return true;
}

// bf16 tile-contiguous caches of X in both orientations (built once per set_X / tile shape)
void build_x_caches(nmfb200_handle* h, bf16** Xr_out, bf16** Xc_out) {
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n;
    const int trH = pick_tile_rows((int)n, h->tc_tile_rows), trW = pick_tile_rows((int)p, h->tc_tile_rows);
    const int64_t tilesH = ceil_div(n, trH), tilesW = ceil_div(p, trW);
    const int nkbH = (int)ceil_div(p, 64), nkbW = (int)ceil_div(n, 64);
    bf16* Xr_ = h->buf_t<bf16>("tc.Xr", (size_t)tilesH * nkbH * trH * 64);  // rows = columns of X, contraction over p
    bf16* Xc_ = h->buf_t<bf16>("tc.Xc", (size_t)tilesW * nkbW * trW * 64);  // rows = rows of X, contraction over n
    if (h->tc_x_epoch != h->x_epoch || h->tc_x_trH != trH || h->tc_x_trW != trW) {
        const float* X = (const float*)h->dX;
        cvt_tiled_direct_kernel<<<dim3((unsigned)(tilesH * trH), (unsigned)std::min<int64_t>(ceil_div((int64_t)nkbH * 64, 256), 64)), 256, 0, st>>>(
            X, h->ldx, (int)n, (int)p, trH, nkbH, Xr_);
        NMF_REQUIRE(trW % 8 == 0, NMFB200_EINVAL, "tile rows must be a multiple of 8");
        // the transpose kernel walks 64 logical rows per block; cover the padded row range of the last tile too
        cvt_tiled_transpose_kernel<<<dim3((unsigned)nkbW, (unsigned)ceil_div(tilesW * trW, 64)), dim3(32, 8), 0, st>>>(
            X, h->ldx, (int)p, (int)n, trW, nkbW, (int)tilesW, Xc_);
        h->launches += 2;
        NMF_CUDA(cudaGetLastError());
        h->tc_x_epoch = h->x_epoch;
        h->tc_x_trH = trH;
        h->tc_x_trW = trW;
    }
    *Xr_out = Xr_;
    *Xc_out = Xc_;
}

template <int KP>
void tc_solve_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs();
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;

    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));

    bf16 *Xr = nullptr, *Xc = nullptr;
    build_x_caches(h, &Xr, &Xc);
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
    // multi-GPU exchange of the packed vector [numerators n x KP | W'W | W-side stop sums]:
    //   p2p  : MODE 1 stores its rows into the owner rank's HBM, reduce + all-gather kernel, flag barriers (NVLink)
    //   nccl : ncclAllReduce on a local packed buffer
    XchgDev xd;
    std::memset(&xd, 0, sizeof(xd));
    const bool p2p = multi && xchg_setup(h, n, KP, &xd);
    if (!p2p) std::memset(&xd, 0, sizeof(xd));  // G = 0: consumers do not wait on peer flags
    const size_t packed_len = (size_t)n * KP + (size_t)KP * KP + 2 * KP;
    float* packed = !multi ? nullptr : (p2p ? xd.packed[xd.rank] : h->buf_t<float>("tc.packed", packed_len));
    float* packed_P = multi ? packed + (size_t)n * KP : nullptr;   // gram_reduce writes the local Gram partial here
    float* packed_ws = multi ? packed_P + (size_t)KP * KP : nullptr;  // conv_reduce writes the local W-side stop sums here
    float* ws_small = multi ? h->buf_t<float>("tc.ws_small", 2 * KP) : nullptr;  // stand-alone decision (end of batch)
    if (multi) NMF_CUDA(cudaMemsetAsync(packed_ws, 0, 2 * KP * sizeof(float), st));
    XchgDev xnone;
    std::memset(&xnone, 0, sizeof(xnone));
    s.launch_gram(W, !multi, packed_P);       // P_W = W'W for the first H-step (partial per rank when sharded)
    if (!a.update_H) s.launch_gram(H, true);  // H never changes: P_H once
    NMF_CUDA(cudaEventRecord(e1, st));

    h->ev_used = 0;
    int64_t enq = 0;
    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    const int post_blocks = 1 + (KP * KP + 255) / 256;
    // Single GPU: both update kernels are launched as programmatic dependents of the small reduce kernel in front of
    // them, so their X streaming overlaps that kernel and the launch gap (the reduce results are only needed by the
    // denominator blocks and the epilogue, which wait for it).
    const bool pdl = h->tc_pdl != 0;  // multi-GPU: the MODE 1 numerator kernel and the W-step follow gram_conv_reduce / gram_reduce too
    // verbose (common.jl:54-59, 76-82): objective before the loop and after every iteration, through the trace callback
    double v_objv = std::numeric_limits<double>::quiet_NaN(), v_t0 = 0;
    auto wall = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    auto objective_now = [&]() {
        double v = 0;
        NMF_REQUIRE(tc_objective<KP>(h, 0, W, H, 0.0, 0.0, &v), NMFB200_ENOTSUP, "verbose on the tensor-core engine needs the tensor-core objective");
        return v;
    };
    if (a.verbose) {
        v_t0 = wall();
        v_objv = objective_now();
        if (h->trace) h->trace(h->trace_user, 0, 0.0, v_objv, NAN, NAN);
    }
    while (enq < a.maxiter) {
        int64_t batch = a.verbose ? 1 : std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            bool pending = multi && i > 0;  // the previous iteration of this batch still awaits its decision
            h->mark("start");
            if (a.update_H) {
                if (!multi) {
                    s.launch_update(0, H, W, Xr, (int)p, lh, delta, nullptr, nullptr, 1, nullptr, pdl);  // H-step (+ tile Grams of the new H)
                    h->mark("updH");
                } else {
                    s.launch_update(1, H, W, Xr, (int)p, lh, delta, packed, nullptr, -1, nullptr, pdl);  // partial numerators of this shard
                    h->mark("mode1");
                    // THE exchange step: [numerators | W'W | W-side stop sums of the previous iteration]
                    unsigned int ep = 0;
                    if (p2p) {
                        ep = ++h->xchg.epoch;
                        xchg_reduce_gather_kernel<<<148, 256, 0, st>>>(xd, KP, ep);  // NVLink peer memory, one launch
                        h->launches += 1;
                        h->mark("xchg");
                    } else {
                        h->allreduce_sum(packed, packed_len);
                        h->mark("nccl");
                    }
                    post_allreduce_kernel<<<post_blocks, 256, 0, st>>>(acc, packed_ws, pending ? 1 : 0, KP, (int)k, tol, state, packed_P,
                                                                       W.Phi, W.Plo, xd, ep);
                    h->launches += 1;
                    pending = false;
                    h->mark("post");
                    s.launch_update(2, H, W, Xr, (int)p, lh, delta, packed, nullptr, 1, nullptr, pdl);   // ratio with the reduced numerators
                    h->mark("mode2");
                }
                h->mark("gramH");
            }
            if (pending) {  // update_H = false: no packed exchange to ride on
                NMF_CUDA(cudaMemcpyAsync(ws_small, packed_ws, 2 * KP * sizeof(float), cudaMemcpyDeviceToDevice, st));
                h->allreduce_sum(ws_small, (size_t)2 * KP);
                post_allreduce_kernel<<<1, 256, 0, st>>>(acc, ws_small, 1, KP, (int)k, tol, state, nullptr, nullptr, nullptr, xnone, 0u);
                h->launches += 1;
            }
            // W-step (local rows) + W'W for the next H-step (partial per rank when sharded; not needed if H is fixed)
            const int gramW = (a.update_H || !multi) ? (multi ? 0 : 1) : -1;
            s.defer_gram_reduce = true;
            s.launch_update(0, W, H, Xc, (int)n, lw, delta, nullptr, nullptr, gramW, packed_P, pdl);
            s.defer_gram_reduce = false;
            h->mark("updW");
            const int gram_blocks = (s.last_fused_gram && gramW >= 0) ? (4 * KP * KP + 255) / 256 : 0;
            // one launch: Gram reduce (if produced by the staged epilogue) + stop_condition reduce / decision
            launch_k(gram_conv_reduce_kernel, dim3(gram_blocks + 4 * (KP / 32)), dim3(256), 0, st, false, (const float*)s.last_gram_part, W.tiles,
                     KP * KP, packed_P ? packed_P : W.P, W.Phi, W.Plo, gramW > 0 ? 1 : 0, gram_blocks, (const float*)W.conv, W.tiles,
                     (const float*)H.conv, H.tiles, KP, (int)k, (int)a.update_H, acc, tol, state, multi ? 0 : 1, packed_ws);
            h->launches += 1;
            h->mark("conv");
        }
        if (multi) {  // decision of the last iteration of the batch
            NMF_CUDA(cudaMemcpyAsync(ws_small, packed_ws, 2 * KP * sizeof(float), cudaMemcpyDeviceToDevice, st));
            h->allreduce_sum(ws_small, (size_t)2 * KP);
            post_allreduce_kernel<<<1, 256, 0, st>>>(acc, ws_small, 1, KP, (int)k, tol, state, nullptr, nullptr, nullptr, xnone, 0u);
            h->launches += 1;
        }
        enq += batch;
        NMF_CUDA(cudaGetLastError());
        NMF_CUDA(cudaMemcpyAsync(&hs, state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        iters = hs.iters;
        devmax = hs.devmax;
        if (a.verbose) {
            const double pre = v_objv;
            v_objv = objective_now();
            if (h->trace) h->trace(h->trace_user, iters, wall() - v_t0, v_objv, v_objv - pre, (double)devmax);
        }
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
    double objv = v_objv;
    if (!a.verbose && !tc_objective<KP>(h, 0, W, H, 0.0, 0.0, &objv)) objv = simt_objective_f32(h, 0, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);
    if (h->tc_debug & 16) {  // diagnostics: is the objective kernel repeatable on fixed inputs?
        fprintf(stderr, "[nmfb200] objective x8:");
        for (int i = 0; i < 8; ++i) {
            double v = 0;
            tc_objective<KP>(h, 0, W, H, 0.0, 0.0, &v);
            fprintf(stderr, " %.9g", v);
        }
        fprintf(stderr, "\n");
    }
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
    h->report_marks(iters);
    if (h->tc_debug & 8) {  // phase clocks of CTA 0 in the last update launch (SM cycles since kernel entry)
        std::vector<long long> tv(16 + 2 * 4096);
        NMF_CUDA(cudaMemcpy(tv.data(), h->buf("tc.timing", tv.size() * sizeof(long long)), tv.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const long long* t = tv.data();
        const int nct = std::min(W.tiles, 4096);
        long long s_min = LLONG_MAX, s_max = 0, e_min = LLONG_MAX, e_max = 0;
        for (int c = 0; c < nct; ++c) {
            s_min = std::min(s_min, t[16 + 2 * c]); s_max = std::max(s_max, t[16 + 2 * c]);
            e_min = std::min(e_min, t[17 + 2 * c]); e_max = std::max(e_max, t[17 + 2 * c]);
        }
        fprintf(stderr, "[nmfb200] last update launch, %d CTAs (globaltimer ns): first entry 0, last entry %lld, first exit %lld, last exit %lld\n",
                nct, s_max - s_min, e_min - s_min, e_max - s_min);
        static const char* names[12] = {"entry", "first_operands", "first_den_block", "mma_issued", "pred_complete", "accum_complete",
                                        "ratio_done", "all_warps_done", "gram_mma_done", "gram_written", "stores_read", "exit"};
        fprintf(stderr, "[nmfb200] mu_update_kernel CTA 0 phase clocks (cycles since entry):");
        for (int i = 1; i < 12; ++i) fprintf(stderr, " %s=%lld", names[i], t[i] - t[0]);
        fprintf(stderr, "\n");
    }
}

// ---- MultUpdate(:div) on the tensor-core engine (multupd.jl:150-193) ----------------------------------------------
// Quotient kernel: Q = X ./ (W H + delta) (multupd.jl:172-174 / :184-186) produced tile by tile, never via a
// p x n fp32 intermediate: a CTA owns 128 rows of the "row factor" Rf (resident in smem), walks the k-blocks of
// its tile-contiguous X panel, and per 128 x 64 tile
//   MMA warp:      D[128 x 64] = Rf_tile * Cf_tile'   (tcgen05, K = KP, bf16 operands) into one of two TMEM buffers,
//   8 epilogue warps: Q = X_tile / (D + delta) from the X tile in smem (swizzled) -> bf16 Q tile in smem ->
//                  TMA store into the Q panel (same tile-contiguous layout as the X panel),
// so the update kernel (MODE 4) can stream Q exactly like it streams X.  HBM traffic: read X (2 B) + write Q (2 B).
struct QuotParams {
    CUtensorMap tmX;   // X panel  bf16 tile-contiguous [tiles*nkb*128][64], box 64 x 128 (load)
    CUtensorMap tmQ;   // Q panel, same geometry (store)
    CUtensorMap tmR;   // row factor hi  bf16 [R][KP],  box 64 x 128
    CUtensorMap tmC;   // col factor hi  bf16 [C][KP],  box 64 x 64
    const TcState* state;
    int nkb;           // k-blocks per tile = ceil(C / 64)
    int kchunk;        // k-blocks handled by one CTA: blockIdx.y walks [y*kchunk, min(nkb, (y+1)*kchunk))
    float delta;
};

template <int KP>
struct QuotCfg {
    static constexpr int NSLAB = KP / 64;
    static constexpr int RF_BYTES = NSLAB * 128 * 128;   // resident row-factor tile
    static constexpr int C_BYTES = NSLAB * 64 * 128;     // one stage of the column factor: 64 rows x KP
    static constexpr int X_BYTES = 128 * 128;
    static constexpr int SC = 4, SX = 4, SO = 2;
    static constexpr int OFF_C = RF_BYTES;
    static constexpr int OFF_X = OFF_C + SC * C_BYTES;
    static constexpr int OFF_O = OFF_X + SX * X_BYTES;
    static constexpr int OFF_BAR = OFF_O + SO * X_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static constexpr int THREADS = 320;                  // w0 producer, w1 MMA, w2..w9 epilogue
    static constexpr int TMEM_COLS = 128;                // 2 buffers x 64 fp32 columns
};

template <int KP>
__global__ void __launch_bounds__(QuotCfg<KP>::THREADS, 1) div_quot_kernel(const __grid_constant__ QuotParams prm) {
    using C = QuotCfg<KP>;
    if (prm.state->converged) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* fullC = (uint64_t*)(smem + C::OFF_BAR);
    uint64_t* emptyC = fullC + C::SC;
    uint64_t* fullX = emptyC + C::SC;
    uint64_t* emptyX = fullX + C::SX;
    uint64_t* tfull = emptyX + C::SX;    // [2]
    uint64_t* tempty = tfull + 2;        // [2]
    uint64_t* rf_full = tempty + 2;
    uint32_t* tmem_slot = (uint32_t*)(rf_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = blockIdx.y * prm.kchunk;                      // quotient tiles are independent: split k freely
    const int nkb = min(prm.nkb, kb0 + prm.kchunk) - kb0;         // k-blocks of this CTA
    const int row0 = blockIdx.x * 128;
    const int prow0 = (blockIdx.x * prm.nkb + kb0) * 128;         // first panel row of this CTA's first tile

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&prm.tmX);
        prefetch_tmap(&prm.tmQ);
        prefetch_tmap(&prm.tmR);
        prefetch_tmap(&prm.tmC);
        for (int s = 0; s < C::SC; ++s) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], 1); }
        for (int s = 0; s < C::SX; ++s) { mbar_init(&fullX[s], 1); mbar_init(&emptyX[s], 8); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
        mbar_init(rf_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: resident row-factor tile, then per k-block the X tile and the column-factor rows =====
        if (elect_one()) {
            mbar_arrive_expect_tx(rf_full, C::RF_BYTES);
            for (int sl = 0; sl < C::NSLAB; ++sl) tma_load_2d(smem + sl * 128 * 128, &prm.tmR, rf_full, 64 * sl, row0);
            int sc = 0, sx = 0;
            uint32_t phc = 0, phx = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&emptyX[sx], phx ^ 1u);
                mbar_arrive_expect_tx(&fullX[sx], C::X_BYTES);
                tma_load_2d(smem + C::OFF_X + sx * C::X_BYTES, &prm.tmX, &fullX[sx], 0, prow0 + kb * 128);
                mbar_wait(&emptyC[sc], phc ^ 1u);
                mbar_arrive_expect_tx(&fullC[sc], C::C_BYTES);
                for (int sl = 0; sl < C::NSLAB; ++sl)
                    tma_load_2d(smem + C::OFF_C + sc * C::C_BYTES + sl * 64 * 128, &prm.tmC, &fullC[sc], 64 * sl, 64 * (kb0 + kb));
                if (++sx == C::SX) { sx = 0; phx ^= 1u; }
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: D = Rf * Cf' for every k-block, alternating TMEM buffers =====
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, 64);
            mbar_wait(rf_full, 0);
            int sc = 0;
            uint32_t phc = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int b = kb & 1;
                mbar_wait(&tempty[b], (((uint32_t)kb >> 1) & 1u) ^ 1u);  // epilogue has drained this TMEM buffer
                mbar_wait(&fullC[sc], phc);
                tc_fence_after();
                const uint32_t cbase = smem_u32(smem + C::OFF_C + sc * C::C_BYTES);
#pragma unroll
                for (int sl = 0; sl < C::NSLAB; ++sl) {
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + sl * 128 * 128));
                    const uint64_t bdesc = make_kmajor_sw128_desc(cbase + sl * 64 * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + b * 64, adesc + 2 * kk, bdesc + 2 * kk, idesc, (sl > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&emptyC[sc]);
                umma_commit(&tfull[b]);
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: 8 warps; warp e handles TMEM lane quarter (warp % 4) and column half e / 4 =====
        const int e = warp - 2;
        const int q = warp & 3, hf = e >> 2;
        const int r = 32 * q + lane;                 // row inside the tile
        const float delta = prm.delta;
        int sx = 0;
        uint32_t phx = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1, ob = kb & 1;
            mbar_wait(&tfull[b], ((uint32_t)kb >> 1) & 1u);
            tc_fence_after();
            uint32_t d[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + b * 64 + 32 * hf, d);
            mbar_wait(&fullX[sx], phx);
            const uint8_t* xt = smem + C::OFF_X + sx * C::X_BYTES + r * 128;
            uint4 xv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) xv[c] = *(const uint4*)(xt + (((4 * hf + c) ^ (r & 7)) << 4));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);  // TMEM buffer b may be overwritten
            uint4 qv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t xin[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
                uint32_t qo[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float x0 = __uint_as_float(xin[w] << 16), x1 = __uint_as_float(xin[w] & 0xffff0000u);
                    const float d0 = __uint_as_float(d[8 * c + 2 * w]) + delta, d1 = __uint_as_float(d[8 * c + 2 * w + 1]) + delta;
                    qo[w] = pack_bf16x2(__float2bfloat16_rn(__fdividef(x0, d0)), __float2bfloat16_rn(__fdividef(x1, d1)));
                }
                qv[c] = make_uint4(qo[0], qo[1], qo[2], qo[3]);
            }
            // the X stage may be refilled only now that its values have been consumed (asynchronous shared-memory loads)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&emptyX[sx]);
            // output staging buffer ob: its previous TMA store (two k-blocks ago) must have finished reading smem
            if (threadIdx.x == 64) tma_store_wait_read<1>();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            uint8_t* ot = smem + C::OFF_O + ob * C::X_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) *(uint4*)(ot + (((4 * hf + c) ^ (r & 7)) << 4)) = qv[c];
            fence_proxy_async();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                tma_store_2d(&prm.tmQ, smem + C::OFF_O + ob * C::X_BYTES, 0, prow0 + kb * 128);
                tma_store_commit();
            }
            if (++sx == C::SX) { sx = 0; phx ^= 1u; }
        }
        if (threadIdx.x == 64) tma_store_wait_all<0>();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- fused quotient + numerator (MultUpdate :div) --------------------------------------------------------------
// One half-step without the Q panel round trip through HBM: per 128 x 64 tile of X
//   MMA 1:  D[128 x 64]  = Rf_tile * Cf_blk'        (K = KP)      -> TMEM buffer b           (as div_quot_kernel)
//   warps:  Q = X_tile / (D + delta) -> bf16 Q tile in shared memory, in the swizzled K-major image an A operand needs
//   MMA 2:  Num[128 x KP] += Q_tile * (Cf_blk)      (K = 64)      -> TMEM accumulator        (as mu_update_kernel MODE 4)
// so X is read once (2 B per cell and half-step instead of 6).  The numerators go to num_part[blockIdx.y][R][KP]
// (k-split partials, summed in order by mu_update_kernel<KP,5>, which also applies F .* Num ./ (colsum + lambda)).
struct DivFusedParams {
    CUtensorMap tmX;   // X panel  bf16 tile-contiguous [tiles*nkb*128][64], box 64 x 128
    CUtensorMap tmR;   // row factor hi  bf16 [R][KP],   box 64 x 128
    CUtensorMap tmC;   // col factor hi  bf16 [C][KP],   box 64 x 64     (B operand of MMA 1)
    CUtensorMap tmT;   // col factor hi transposed bf16 [KP][ldC], box 64 x KP (B operand of MMA 2)
    const TcState* state;
    float* num_part;   // [gridDim.y][R][KP]
    int R;
    int nkb, kchunk;
    float delta;
};

template <int KP>
struct DivFusedCfg {
    static constexpr int NSLAB = KP / 64;
    static constexpr int RF_BYTES = NSLAB * 128 * 128;   // resident row-factor tile
    static constexpr int C_BYTES = NSLAB * 64 * 128;     // 64 rows x KP   (MMA 1 B operand)
    static constexpr int T_BYTES = KP * 128;             // KP rows x 64   (MMA 2 B operand)
    static constexpr int X_BYTES = 128 * 128;
    static constexpr int SC = KP == 64 ? 4 : 2, ST = KP == 64 ? 4 : 3, SX = KP == 64 ? 4 : 3, SQ = 2;
    static constexpr int OFF_C = RF_BYTES;
    static constexpr int OFF_T = OFF_C + SC * C_BYTES;
    static constexpr int OFF_X = OFF_T + ST * T_BYTES;
    static constexpr int OFF_Q = OFF_X + SX * X_BYTES;
    static constexpr int OFF_BAR = OFF_Q + SQ * X_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static constexpr int NW = 16;                        // quotient warps: 4 per TMEM lane quarter, CPW columns of the 64-wide tile each
    static constexpr int CPW = 256 / NW;                 // 16 (NW = 16) or 32 (NW = 8)
    static constexpr int THREADS = 64 + 32 * NW;         // w0 producer, w1 MMA, w2.. quotient / epilogue
    static constexpr int TMEM_COLS = 256;                // D: 2 x 64 columns at [0,128); Num: KP columns at [128, 128+KP)
    static_assert(NW == 8 || NW == 16, "quotient warps");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int KP>
__global__ void __launch_bounds__(DivFusedCfg<KP>::THREADS, 1) div_fused_kernel(const __grid_constant__ DivFusedParams prm) {
    using C = DivFusedCfg<KP>;
    if (prm.state->converged) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* fullC = (uint64_t*)(smem + C::OFF_BAR);
    uint64_t* emptyC = fullC + C::SC;
    uint64_t* fullT = emptyC + C::SC;
    uint64_t* emptyT = fullT + C::ST;
    uint64_t* fullX = emptyT + C::ST;
    uint64_t* emptyX = fullX + C::SX;
    uint64_t* tfull = emptyX + C::SX;    // [2]  D buffer complete (MMA 1 -> warps)
    uint64_t* tempty = tfull + 2;        // [2]  D buffer drained  (warps -> MMA 1)
    uint64_t* qfull = tempty + 2;        // [SQ] Q tile written    (warps -> MMA 2)
    uint64_t* qempty = qfull + C::SQ;    // [SQ] Q tile consumed   (MMA 2 -> warps)
    uint64_t* rf_full = qempty + C::SQ;
    uint64_t* num_full = rf_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(num_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb0 = blockIdx.y * prm.kchunk;
    const int nkb = min(prm.nkb, kb0 + prm.kchunk) - kb0;         // k-blocks of this CTA (>= 1 by construction of the grid)
    const int row0 = blockIdx.x * 128;
    const int prow0 = (blockIdx.x * prm.nkb + kb0) * 128;         // first panel row of this CTA's first tile

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&prm.tmX);
        prefetch_tmap(&prm.tmR);
        prefetch_tmap(&prm.tmC);
        prefetch_tmap(&prm.tmT);
        for (int i = 0; i < C::SC; ++i) { mbar_init(&fullC[i], 1); mbar_init(&emptyC[i], 1); }
        for (int i = 0; i < C::ST; ++i) { mbar_init(&fullT[i], 1); mbar_init(&emptyT[i], 1); }
        for (int i = 0; i < C::SX; ++i) { mbar_init(&fullX[i], 1); mbar_init(&emptyX[i], C::NW); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], C::NW); }
        for (int i = 0; i < C::SQ; ++i) { mbar_init(&qfull[i], C::NW); mbar_init(&qempty[i], 1); }
        mbar_init(rf_full, 1);
        mbar_init(num_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_num = tmem_base + 128;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_arrive_expect_tx(rf_full, C::RF_BYTES);
            for (int sl = 0; sl < C::NSLAB; ++sl) tma_load_2d(smem + sl * 128 * 128, &prm.tmR, rf_full, 64 * sl, row0);
            int sc = 0, st = 0, sx = 0;
            uint32_t phc = 0, pht = 0, phx = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&emptyX[sx], phx ^ 1u);
                mbar_arrive_expect_tx(&fullX[sx], C::X_BYTES);
                tma_load_2d(smem + C::OFF_X + sx * C::X_BYTES, &prm.tmX, &fullX[sx], 0, prow0 + kb * 128);
                mbar_wait(&emptyC[sc], phc ^ 1u);
                mbar_arrive_expect_tx(&fullC[sc], C::C_BYTES);
                for (int sl = 0; sl < C::NSLAB; ++sl)
                    tma_load_2d(smem + C::OFF_C + sc * C::C_BYTES + sl * 64 * 128, &prm.tmC, &fullC[sc], 64 * sl, 64 * (kb0 + kb));
                mbar_wait(&emptyT[st], pht ^ 1u);
                mbar_arrive_expect_tx(&fullT[st], C::T_BYTES);
                tma_load_2d(smem + C::OFF_T + st * C::T_BYTES, &prm.tmT, &fullT[st], 64 * (kb0 + kb), 0);
                if (++sx == C::SX) { sx = 0; phx ^= 1u; }
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
                if (++st == C::ST) { st = 0; pht ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: MMA 1 of block kb, then MMA 2 of block kb - 1 (its Q tile is being produced meanwhile) =====
        if (elect_one()) {
            constexpr uint32_t idesc1 = make_idesc(FMT_BF16, 128, 64);
            constexpr uint32_t idesc2 = make_idesc(FMT_BF16, 128, KP);
            mbar_wait(rf_full, 0);
            int sc = 0, st = 0;
            uint32_t phc = 0, pht = 0;
            auto mma2 = [&](int j) {
                const int o = j % C::SQ;
                mbar_wait(&qfull[o], ((uint32_t)(j / C::SQ)) & 1u);
                mbar_wait(&fullT[st], pht);
                tc_fence_after();
                const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + C::OFF_Q + o * C::X_BYTES));
                const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + C::OFF_T + st * C::T_BYTES));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_num, adesc + 2 * kk, bdesc + 2 * kk, idesc2, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&qempty[o]);
                umma_commit(&emptyT[st]);
                if (++st == C::ST) { st = 0; pht ^= 1u; }
            };
            for (int kb = 0; kb < nkb; ++kb) {
                const int b = kb & 1;
                mbar_wait(&tempty[b], (((uint32_t)kb >> 1) & 1u) ^ 1u);  // the warps have drained this D buffer
                mbar_wait(&fullC[sc], phc);
                tc_fence_after();
                const uint32_t cbase = smem_u32(smem + C::OFF_C + sc * C::C_BYTES);
#pragma unroll
                for (int sl = 0; sl < C::NSLAB; ++sl) {
                    const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(smem + sl * 128 * 128));
                    const uint64_t bdesc = make_kmajor_sw128_desc(cbase + sl * 64 * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem_base + b * 64, adesc + 2 * kk, bdesc + 2 * kk, idesc1, (sl > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&emptyC[sc]);
                umma_commit(&tfull[b]);
                if (++sc == C::SC) { sc = 0; phc ^= 1u; }
                if (kb > 0) mma2(kb - 1);
            }
            mma2(nkb - 1);
            umma_commit(num_full);
        }
        __syncwarp();
    } else {
        // ===== quotient warps: warp e handles TMEM lane quarter (warp % 4) and columns [CPW*cp, CPW*cp + CPW) of the tile =====
        const int e = warp - 2;
        const int q = warp & 3, cp = e >> 2;
        const int r = 32 * q + lane;                 // row inside the tile
        const float delta = prm.delta;
        constexpr int CPW = C::CPW;                  // columns per warp
        constexpr int NCH = CPW / 8;                 // 16-byte chunks (8 bf16) per row and warp
        int sx = 0;
        uint32_t phx = 0;
        for (int kb = 0; kb < nkb; ++kb) {
            const int b = kb & 1, o = kb % C::SQ;
            mbar_wait(&tfull[b], ((uint32_t)kb >> 1) & 1u);
            tc_fence_after();
            uint32_t d[CPW];
            if constexpr (CPW == 32) tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + b * 64 + CPW * cp, *(uint32_t(*)[32])d);
            else tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + b * 64 + CPW * cp, *(uint32_t(*)[16])d);
            mbar_wait(&fullX[sx], phx);
            const uint8_t* xt = smem + C::OFF_X + sx * C::X_BYTES + r * 128;
            uint4 xv[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) xv[c] = *(const uint4*)(xt + (((NCH * cp + c) ^ (r & 7)) << 4));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);  // TMEM buffer b may be overwritten
            uint4 qv[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const uint32_t xin[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
                uint32_t qo[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float x0 = __uint_as_float(xin[w] << 16), x1 = __uint_as_float(xin[w] & 0xffff0000u);
                    const float d0 = __uint_as_float(d[8 * c + 2 * w]) + delta, d1 = __uint_as_float(d[8 * c + 2 * w + 1]) + delta;
                    // The kernel is bound by the special-function pipe (one MUFU.RCP and one F2F per element as written):
                    // ONE reciprocal for the pair, 1/d0 = d1/(d0*d1), and ONE packed conversion (cvt.rn.bf16x2.f32).
                    // d0, d1 >= delta = 3.45e-4 and the product stays finite while the entries of W*H stay below ~1e19.
                    const float rr = __fdividef(1.0f, d0 * d1);
                    const __nv_bfloat162 qq = __floats2bfloat162_rn(x0 * (rr * d1), x1 * (rr * d0));
                    qo[w] = *reinterpret_cast<const uint32_t*>(&qq);
                }
                qv[c] = make_uint4(qo[0], qo[1], qo[2], qo[3]);
            }
            // Q stage o: MMA 2 of block kb - SQ must have read it
            mbar_wait(&qempty[o], (((uint32_t)(kb / C::SQ)) & 1u) ^ 1u);
            uint8_t* ot = smem + C::OFF_Q + o * C::X_BYTES + r * 128;
#pragma unroll
            for (int c = 0; c < NCH; ++c) *(uint4*)(ot + (((NCH * cp + c) ^ (r & 7)) << 4)) = qv[c];
            fence_proxy_async();     // generic-proxy stores (Q) and consumed loads (X) before the async proxy touches either stage
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&emptyX[sx]);            // X stage may be refilled (its values have been consumed)
                mbar_arrive(&qfull[o]);              // this warp's part of the Q tile is in place
            }
            if (++sx == C::SX) { sx = 0; phx ^= 1u; }
        }
        // numerators of this (tile, k-chunk): TMEM -> num_part[blockIdx.y][row][.]; warp (q, cp) takes KP / (NW/4) columns
        mbar_wait(num_full, 0);
        tc_fence_after();
        const int row = row0 + r;
        float* dst_row = prm.num_part + ((size_t)blockIdx.y * prm.R + row) * KP;
        constexpr int NCOL = KP / (C::NW / 4);       // 16 or 32 (KP = 64), 32 or 64 (KP = 128)
        constexpr int STEP = NCOL >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = cp * NCOL; c0 < (cp + 1) * NCOL; c0 += STEP) {
            uint32_t v[STEP];
            if constexpr (STEP == 32) tmem_ld32(tmem_num + ((uint32_t)(32 * q) << 16) + c0, *(uint32_t(*)[32])v);
            else tmem_ld16(tmem_num + ((uint32_t)(32 * q) << 16) + c0, *(uint32_t(*)[16])v);
            tmem_ld_wait();
            if (row < prm.R) {
#pragma unroll
                for (int j = 0; j < STEP / 4; ++j)
                    ((float4*)(dst_row + c0))[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
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

// column sums of a row-factor [R][KP] (sW = sum(W,1), sH = sum(H,2): multupd.jl:176,188): per 128-row tile, then reduced
__global__ void __launch_bounds__(256) colsum_tiles_kernel(const float* __restrict__ Fm, int R, int KP, float* __restrict__ part,
                                                           const TcState* st) {
    if (st->converged) return;
    __shared__ float red[256];
    const int groups = 256 / KP > 0 ? 256 / KP : 1;
    const int g = threadIdx.x / KP, a = threadIdx.x % KP;
    const int r0 = blockIdx.x * 128;
    float s = 0.f;
    if (g < groups)
        for (int rr = g; rr < 128 && r0 + rr < R; rr += groups) s += Fm[(size_t)(r0 + rr) * KP + a];
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < KP) {
        for (int gg = 1; gg < groups; ++gg) s += red[gg * KP + a];
        part[(size_t)blockIdx.x * KP + a] = s;
    }
}
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, int tiles, int KP, float* __restrict__ out,
                                                            const TcState* st) {
    if (st->converged) return;
    __shared__ double red[8][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    double s = 0.0;
    for (int t = w; t < tiles; t += 8) s += (double)__ldcg(part + (size_t)t * KP + c);
    red[w][lane] = s;
    __syncthreads();
    if (w == 0) {
        double tot = red[0][lane];
        for (int i = 1; i < 8; ++i) tot += red[i][lane];
        out[c] = (float)tot;
    }
}

template <int KP>
void tc_solve_div_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs();
    static bool quot_attr = false;
    if (!quot_attr) {
        NMF_CUDA(cudaFuncSetAttribute(div_quot_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, QuotCfg<KP>::SMEM_BYTES));
        quot_attr = true;
    }
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = std::max((float)a.lambda_w, delta), lh = std::max((float)a.lambda_h, delta);  // multupd.jl:37-40
    const float tol = (float)a.tol;
    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    bf16 *Xr = nullptr, *Xc = nullptr;
    build_x_caches(h, &Xr, &Xc);
    NMF_CUDA(cudaEventRecord(e0, st));

    Factor W = alloc_factor(h, "W", (int)p, KP), H = alloc_factor(h, "H", (int)n, KP);
    NMF_REQUIRE(W.tile_rows == 128 && H.tile_rows == 128, NMFB200_ENOTSUP, "tensor-core :div path needs 128-row tiles");
    TcState* state = (TcState*)h->buf("tc.state", sizeof(TcState));
    double* acc = h->buf_t<double>("tc.acc", 4 * KP);
    const int nkbH = (int)ceil_div(p, 64), nkbW = (int)ceil_div(n, 64);
    const size_t q_elems = std::max((size_t)H.tiles * nkbH, (size_t)W.tiles * nkbW) * 128 * 64;
    bf16* Q = h->tc_div_fused != 0 ? nullptr : h->buf_t<bf16>("tc.Q", q_elems);
    float* cs_part = h->buf_t<float>("tc.colsum_part", (size_t)std::max(W.tiles, H.tiles) * KP);
    NMF_CUDA(cudaMemsetAsync(state, 0, sizeof(TcState), st));
    NMF_CUDA(cudaMemsetAsync(W.bT, 0, (size_t)W.rowsT * W.ldT * sizeof(bf16), st));
    NMF_CUDA(cudaMemsetAsync(H.bT, 0, (size_t)H.rowsT * H.ldT * sizeof(bf16), st));

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
    TcSolver<KP> s{h, st, state};
    NMF_CUDA(cudaEventRecord(e1, st));

    // one half-step: Q = X ./ (Rf Cf' + delta) over Rf's panel, column sums of Cf, then Rf <- Rf .* (Q Cf) ./ (colsum + lambda).
    // Fused form (default): div_fused_kernel keeps Q on chip (quotient tile -> shared memory -> second MMA) and writes k-split
    // partial numerators; mu_update_kernel<KP,5> sums them and applies the ratio.  Unfused form (option tc_div_fused=0):
    // div_quot_kernel writes a bf16 Q panel that mu_update_kernel<KP,4> streams like X.
    const bool fused = h->tc_div_fused != 0;
    if (fused) {
        static bool fattr = false;
        if (!fattr) {
            NMF_CUDA(cudaFuncSetAttribute(div_fused_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, DivFusedCfg<KP>::SMEM_BYTES));
            fattr = true;
        }
    }
    auto pick_ksplit = [](int tiles, int nkb) {  // fewest k-chunks (>= 16 k-blocks each) that fill whole waves of 148 CTAs to >= 93 %
        int best = 1;
        double best_eff = 0;
        for (int ks = 1; ks <= std::max(1, nkb / 16) && ks <= 16; ++ks) {
            const int ctas = tiles * ks;
            const double eff = (double)ctas / (double)(ceil_div(ctas, 148) * 148);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = ks; }
            if (eff >= 0.93) { best = ks; break; }
        }
        return best;
    };
    auto half_step = [&](Factor& Rf, Factor& Cf, const bf16* Xs, int nkb, int Kdim, float lambda) {
        const uint64_t prow = (uint64_t)Rf.tiles * nkb * 128;
        colsum_tiles_kernel<<<Cf.tiles, 256, 0, st>>>(Cf.m, Cf.R, KP, cs_part, state);
        colsum_reduce_kernel<<<KP / 32, 256, 0, st>>>(cs_part, Cf.tiles, KP, Cf.colsum, state);
        h->launches += 2;
        if (fused) {
            DivFusedParams fp;
            fp.tmX = make_tmap_bf16(Xs, 64, prow, 64, 128);
            fp.tmR = make_tmap_bf16(Rf.hi, KP, (uint64_t)Rf.R, KP, 128);
            fp.tmC = make_tmap_bf16(Cf.hi, KP, (uint64_t)Cf.R, KP, 64);
            fp.tmT = make_tmap_bf16(Cf.bT, (uint64_t)Kdim, KP, (uint64_t)Cf.ldT, KP);
            fp.state = state;
            fp.R = Rf.R;
            fp.nkb = nkb;
            int ksplit = pick_ksplit(Rf.tiles, nkb);
            fp.kchunk = (int)ceil_div(nkb, ksplit);
            ksplit = (int)ceil_div(nkb, fp.kchunk);
            fp.num_part = h->buf_t<float>("tc.div_num_part", (size_t)ksplit * Rf.R * KP);
            fp.delta = delta;
            div_fused_kernel<KP><<<dim3(Rf.tiles, ksplit), DivFusedCfg<KP>::THREADS, DivFusedCfg<KP>::SMEM_BYTES, st>>>(fp);
            h->launches += 1;
            s.num_splits = ksplit;
            s.num_split_stride = (int64_t)Rf.R * KP;
            s.launch_update(5, Rf, Cf, Xs, Kdim, lambda, delta, fp.num_part);
            s.num_splits = 1;
            s.num_split_stride = 0;
            return;
        }
        QuotParams qp;
        qp.tmX = make_tmap_bf16(Xs, 64, prow, 64, 128);
        qp.tmQ = make_tmap_bf16(Q, 64, prow, 64, 128);
        qp.tmR = make_tmap_bf16(Rf.hi, KP, (uint64_t)Rf.R, KP, 128);
        qp.tmC = make_tmap_bf16(Cf.hi, KP, (uint64_t)Cf.R, KP, 64);
        qp.state = state;
        qp.nkb = nkb;
        // enough CTAs for ~2 waves, at least 32 k-blocks each
        int ksplit = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(296, Rf.tiles), nkb / 32));
        qp.kchunk = (int)ceil_div(nkb, ksplit);
        ksplit = (int)ceil_div(nkb, qp.kchunk);
        qp.delta = delta;
        div_quot_kernel<KP><<<dim3(Rf.tiles, ksplit), QuotCfg<KP>::THREADS, QuotCfg<KP>::SMEM_BYTES, st>>>(qp);
        h->launches += 1;
        s.launch_update(4, Rf, Cf, Q, Kdim, lambda, delta, nullptr);
    };

    h->ev_used = 0;
    int64_t enq = 0;
    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    while (enq < a.maxiter) {
        int64_t batch = std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            if (a.update_H) half_step(H, W, Xr, nkbH, (int)p, lh);   // multupd.jl:171-181
            half_step(W, H, Xc, nkbW, (int)n, lw);                    // :183-192
            conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 1, nullptr);
            h->launches += 1;
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
    unpack_factor_kernel<<<ew_grid(p * k), 256, 0, st>>>(W.m, (int)p, (int)k, KP, Wd, 1, ldwd);
    unpack_factor_kernel<<<ew_grid(n * k), 256, 0, st>>>(H.m, (int)n, (int)k, KP, Hd, ldhd, 1);
    h->launches += 2;
    double objv = 0;  // gkldiv (multupd.jl:148)
    if (!tc_objective<KP>(h, 1, W, H, 0.0, 0.0, &objv)) objv = simt_objective_f32(h, 1, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);
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

// ---- GreedyCD on the tensor-core engine (greedycd.jl:94-178) ---------------------------------------------------
// After the per-row coordinate kernel has rewritten the fp32 master, rebuild the bf16 operand forms of the factor
// and the stop_condition partial sums against the copy taken before the half-step.  One block per 128-row tile.
__global__ void __launch_bounds__(256) gcd_repack_kernel(const float* __restrict__ Fm, const float* __restrict__ Fprev, int R, int KP,
                                                         bf16* __restrict__ Fhi, bf16* __restrict__ Flo, bf16* __restrict__ FbT, int64_t ldT,
                                                         float* __restrict__ conv_part) {
    __shared__ float red[2][256];
    const int groups = 256 / KP > 0 ? 256 / KP : 1;
    const int cols_per_thread = KP > 256 ? 0 : 1;
    (void)cols_per_thread;
    const int g = threadIdx.x / KP, a = threadIdx.x % KP;
    const int r0 = blockIdx.x * 128;
    float d2 = 0.f, s2 = 0.f;
    if (g < groups) {
        for (int rr = g; rr < 128; rr += groups) {
            const int r = r0 + rr;
            if (r >= R) break;
            const size_t idx = (size_t)r * KP + a;
            const float v = Fm[idx], o = Fprev[idx];
            const bf16 hi = __float2bfloat16_rn(v);
            Fhi[idx] = hi;
            Flo[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
            FbT[(size_t)a * ldT + r] = hi;
            const float dd = v - o, ss = v + o;
            d2 += dd * dd;
            s2 += ss * ss;
        }
    }
    red[0][threadIdx.x] = d2;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < KP) {
        for (int gg = 1; gg < groups; ++gg) { d2 += red[0][gg * KP + a]; s2 += red[1][gg * KP + a]; }
        conv_part[(size_t)blockIdx.x * 2 * KP + a] = d2;
        conv_part[(size_t)blockIdx.x * 2 * KP + KP + a] = s2;
    }
}

template <int KP>
void tc_solve_gcd_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs();
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;
    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    bf16 *Xr = nullptr, *Xc = nullptr;
    build_x_caches(h, &Xr, &Xc);
    NMF_CUDA(cudaEventRecord(e0, st));

    Factor W = alloc_factor(h, "W", (int)p, KP), H = alloc_factor(h, "H", (int)n, KP);
    // conv partials here are per 128-row tile of the repack kernel
    const int tilesW = (int)ceil_div(p, 128), tilesH = (int)ceil_div(n, 128);
    W.conv = h->buf_t<float>("tc.W.conv", (size_t)std::max(tilesW, W.tiles) * 2 * KP);
    H.conv = h->buf_t<float>("tc.H.conv", (size_t)std::max(tilesH, H.tiles) * 2 * KP);
    TcState* state = (TcState*)h->buf("tc.state", sizeof(TcState));
    double* acc = h->buf_t<double>("tc.acc", 4 * KP);
    const size_t rmax = (size_t)std::max(p, n);
    float* G = h->buf_t<float>("tc.gcd_G", rmax * KP);
    float* prev = h->buf_t<float>("tc.gcd_prev", rmax * KP);
    const int maxtiles = std::max(W.tiles, H.tiles);
    float* bmax = h->buf_t<float>("tc.gcd_bmax", (size_t)maxtiles + 1);
    unsigned long long* d_updates = (unsigned long long*)h->buf("tc.gcd_updates", 16);
    NMF_CUDA(cudaMemsetAsync(state, 0, sizeof(TcState), st));
    NMF_CUDA(cudaMemsetAsync(d_updates, 0, sizeof(unsigned long long), st));
    NMF_CUDA(cudaMemsetAsync(W.bT, 0, (size_t)W.rowsT * W.ldT * sizeof(bf16), st));
    NMF_CUDA(cudaMemsetAsync(H.bT, 0, (size_t)H.rowsT * H.ldT * sizeof(bf16), st));

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
    TcSolver<KP> s{h, st, state};
    s.launch_gram(H, true);  // P = HH' for the first W-step (greedycd.jl:117)
    NMF_CUDA(cudaEventRecord(e1, st));

    auto half_step = [&](Factor& F, Factor& O, const bf16* Xs, int Kdim, float lambda, int tiles128) {
        NMF_CUDA(cudaMemcpyAsync(prev, F.m, (size_t)F.R * KP * sizeof(float), cudaMemcpyDeviceToDevice, st));
        s.launch_update(3, F, O, Xs, Kdim, lambda, 0.f, G, bmax);                                  // G = F P - X O (+lambda), per-CTA max D
        max_partials_kernel<float><<<1, 256, 0, st>>>(bmax, F.tiles, bmax + maxtiles);            // p_init (:132-137)
        gcd_rows_tc_kernel<KP><<<(unsigned)ceil_div(F.R, 8), 256, 0, st>>>(F.m, G, O.P, F.R, bmax + maxtiles, d_updates);  // :139-165
        gcd_repack_kernel<<<tiles128, 256, 0, st>>>(F.m, prev, F.R, KP, F.hi, F.lo, F.bT, F.ldT, F.conv);
        h->launches += 3;
        s.launch_gram(F, true);                                                                   // Gram of the updated factor
    };

    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    while (iters < a.maxiter && !converged) {  // one host check per iteration: the row kernel has no early-exit flag
        half_step(W, H, Xc, (int)n, lw, tilesW);                                                   // W first (greedycd.jl:169-171)
        if (a.update_H) half_step(H, W, Xr, (int)p, lh, tilesH);                                   // then H (:173-177)
        conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, tilesW, H.conv, tilesH, KP, (int)k, a.update_H, acc, tol, state, 1, nullptr);
        h->launches += 1;
        NMF_CUDA(cudaGetLastError());
        NMF_CUDA(cudaMemcpyAsync(&hs, state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        iters = hs.iters;
        devmax = hs.devmax;
        converged = hs.converged != 0;
    }
    NMF_CUDA(cudaEventRecord(e2, st));
    unpack_factor_kernel<<<ew_grid(p * k), 256, 0, st>>>(W.m, (int)p, (int)k, KP, Wd, 1, ldwd);
    unpack_factor_kernel<<<ew_grid(n * k), 256, 0, st>>>(H.m, (int)n, (int)k, KP, Hd, ldhd, 1);
    h->launches += 2;
    double objv = 0;  // greedycd.jl:82-92
    if (!tc_objective<KP>(h, 2, W, H, a.lambda_w, a.lambda_h, &objv))
        objv = simt_objective_f32(h, 2, Wd, ldwd, Hd, ldhd, k, a.lambda_w, a.lambda_h);
    unsigned long long upd = 0;
    NMF_CUDA(cudaMemcpyAsync(&upd, d_updates, sizeof(upd), cudaMemcpyDeviceToHost, st));
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
    out->coordinate_updates = (int64_t)upd;
    out->kernel_launches = h->launches;
    out->hot_kernel_ms = h->drain_event_pairs(&out->hot_kernel_launches);
}

}  // namespace

bool tc_supported(const nmfb200_handle* h, const SolveArgs& a) {
    if (a.alg > 2) return false;          // ProjectedALS / CoordinateDescent / ALSPGrad: exact engine
    if (a.alg == 1) {                     // MultUpdate(:div): quotient kernel + update kernel; k <= 128, single GPU
        if (h->comm != nullptr || a.k > 128 || h->p < 128 || h->n < 128) return false;
        if (h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 20)) return false;
    }
    if (a.alg == 2) {                     // GreedyCD: bf16 gradients; single GPU; auto-selected for large problems only
        if (h->comm != nullptr) return false;
        if (h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 24)) return false;
    }
    if (a.verbose && (a.alg != 0 || h->n < 128 || h->p < 64 || (h->ldx % 4) != 0)) return false;  // per-iteration objective: MU-MSE only
    if (pick_kp(a.k) == 0) return false;
    if (h->p > (int64_t)INT32_MAX / 256 || h->n > (int64_t)INT32_MAX / 256) return false;
    return true;
}

void tc_solve(nmfb200_handle* h, const SolveArgs& a, float* W, int64_t ldw, float* H, int64_t ldh, nmfb200_result* out) {
    if (a.alg == 1) {
        switch (pick_kp(a.k)) {
            case 64: tc_solve_div_kp<64>(h, a, W, ldw, H, ldh, out); return;
            case 128: tc_solve_div_kp<128>(h, a, W, ldw, H, ldh, out); return;
            default: throw Error{NMFB200_ENOTSUP, "k > 128 is not covered by the tensor-core :div path"};
        }
    }
    if (a.alg == 2) {
        switch (pick_kp(a.k)) {
            case 64: tc_solve_gcd_kp<64>(h, a, W, ldw, H, ldh, out); return;
            case 128: tc_solve_gcd_kp<128>(h, a, W, ldw, H, ldh, out); return;
            case 256: tc_solve_gcd_kp<256>(h, a, W, ldw, H, ldh, out); return;
            default: throw Error{NMFB200_ENOTSUP, "k > 256 is not covered by the tensor-core engine"};
        }
    }
    switch (pick_kp(a.k)) {
        case 64: tc_solve_kp<64>(h, a, W, ldw, H, ldh, out); break;
        case 128: tc_solve_kp<128>(h, a, W, ldw, H, ldh, out); break;
        case 256: tc_solve_kp<256>(h, a, W, ldw, H, ldh, out); break;
        default: throw Error{NMFB200_ENOTSUP, "k > 256 is not covered by the tensor-core engine"};
    }
}

void tc_release(nmfb200_handle* h) { xchg_teardown(h); }

}  // namespace nmfb200
