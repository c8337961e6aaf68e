// tc_objective.cuh -- objective on tensor cores (multupd.jl:81,148; greedycd.jl:84)
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

// ---- objective on tensor cores (multupd.jl:81,148; greedycd.jl:84) --------------------------------------------------
// 0.5*||X - WH||^2 or gkldiv(X, WH) without materialising WH: same pipeline as the quotient kernel, but the X tile
// is the caller's fp32 X (TMA, 2 boxes of 128 x 32 fp32 per 128 x 64 tile), WH = Rf*Cf' uses the bf16 hi/lo split
// of both factors (hi*hi + hi*lo + lo*hi, ~2^-17 relative) and the epilogue reduces in fp64
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
    // hi/lo split of both factors for every KP.  KP = 256 pays for it with single-buffered operand stages (the resident H
    // tile alone is 128 KB): this kernel runs once per solve, and 1e-3 on objvalue (bf16 hi only) missed the 1e-4 bar.
    static constexpr bool SPLIT = true;
    static constexpr int NT = SPLIT ? 2 : 1;             // hi (+ lo) copies
    static constexpr int NSLAB = KP / 64;
    static constexpr int RF_BYTES = NT * NSLAB * 128 * 128;
    static constexpr int C_BYTES = NT * NSLAB * 64 * 128;
    static constexpr int X_BYTES = 2 * 128 * 128;        // 128 rows x 64 fp32
    static constexpr int SC = KP <= 128 ? 2 : 1, SX = KP <= 128 ? 2 : 1;
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

// Objective of the current factors (row-factor layout with fresh hi/lo copies) on tensor cores, in three steps so that a
// row-sharded solve can combine the partial sums of its ranks:
//   tc_objective_covers(...)  -- can the tensor-core kernel take this (shard of) X?  (alignment / minimum size)
//   tc_objective_enqueue(...) -- launches; returns a device pointer res[3] = {data term, |W|_1, |H|_1} (Float64 partial sums
//                                over the rows of X / W given; H is replicated)
//   tc_objective_value(...)   -- the reference's rounding of the combined sums (multupd.jl:81,148; greedycd.jl:82-92)
inline bool tc_objective_covers(const float* X, int64_t p, int64_t n, int64_t ldx) {
    return !(n < 128 || p < 64 || (ldx % 4) != 0 || (((uintptr_t)X) & 15) != 0);
}

template <int KP>
double* tc_objective_enqueue(nmfb200_handle* h, const std::string& pfx, int alg, const float* X, int64_t p, int64_t n, int64_t ldx,
                             const Factor& W, const Factor& H, double lambda_w, double lambda_h) {
    cudaStream_t st = h->stream;
    static bool attr_on[64] = {};   // per device ordinal (cudaFuncSetAttribute is per device)
    bool& attr = attr_on[h->device & 63];
    if (!attr) {
        NMF_CUDA(cudaFuncSetAttribute(objective_tc_kernel<KP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ObjCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(objective_tc_kernel<KP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ObjCfg<KP>::SMEM_BYTES));
        attr = true;
    }
    ObjParams op;
    op.tmX = make_tmap_f32(X, (uint64_t)p, (uint64_t)n, (uint64_t)ldx, 128);
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
    double* part = h->buf_t<double>(pfx + ".obj_part", (size_t)nparts + 2048 + 4);
    op.part = part;
    if (alg == 1) objective_tc_kernel<KP, 1><<<dim3(tiles, ksplit), ObjCfg<KP>::THREADS, ObjCfg<KP>::SMEM_BYTES, st>>>(op);
    else objective_tc_kernel<KP, 0><<<dim3(tiles, ksplit), ObjCfg<KP>::THREADS, ObjCfg<KP>::SMEM_BYTES, st>>>(op);
    double* res = part + nparts;  // [0] data term, [1] |W|_1, [2] |H|_1
    NMF_CUDA(cudaMemsetAsync(res, 0, 4 * sizeof(double), st));
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
    return res;
}

inline double tc_objective_value(int alg, const double hres[3], double lambda_w, double lambda_h) {
    if (alg == 1) return (double)(float)hres[0];              // gkldiv returns Float64; Result{T} converts (common.jl:32)
    float r = 0.5f * (float)hres[0];                          // convert(T, 0.5) * sqL2dist (multupd.jl:81)
    if (alg == 2 && lambda_w > 0) r = r + (float)lambda_w * (float)hres[1];   // greedycd.jl:85-90
    if (alg == 2 && lambda_h > 0) r = r + (float)lambda_h * (float)hres[2];
    return (double)r;
}

// Single-GPU form.  Returns false if the shape is not covered (caller falls back to the exact fp32 GEMM + reduction of
// the SIMT engine).
template <int KP>
bool tc_objective(nmfb200_handle* h, int alg, const Factor& W, const Factor& H, double lambda_w, double lambda_h, double* out) {
    if (!tc_objective_covers((const float*)h->dX, h->p, h->n, h->ldx)) return false;
    double* res = tc_objective_enqueue<KP>(h, "tc", alg, (const float*)h->dX, h->p, h->n, h->ldx, W, H, lambda_w, lambda_h);
    double hres[3] = {0, 0, 0};
    NMF_CUDA(cudaMemcpyAsync(hres, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    NMF_CUDA(cudaStreamSynchronize(h->stream));
    *out = tc_objective_value(alg, hres, lambda_w, lambda_h);
    return true;
}
