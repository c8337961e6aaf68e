// tc_div.cuh -- MultUpdate(:div) kernels: fused quotient + numerator, the older quotient-panel kernel, column sums
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

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
    // Launched as a programmatic dependent of colsum_reduce_kernel (whose output only the NEXT kernel reads): do not
    // complete before it has, so that plain stream order behind us still implies "column sums are final".
    pdl_wait();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// column sums of a row-factor [R][KP] (sW = sum(W,1), sH = sum(H,2): multupd.jl:176,188): per 128-row tile, then reduced
__global__ void __launch_bounds__(256) colsum_tiles_kernel(const float* __restrict__ Fm, int R, int KP, float* __restrict__ part,
                                                           const TcState* st) {
    pdl_launch_dependents();  // PDL chain colsum_tiles -> colsum_reduce -> div_fused: the big kernel does not wait for these two
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
    pdl_launch_dependents();  // div_fused_kernel (which never reads the column sums) may start now
    pdl_wait();               // ... while we wait for colsum_tiles_kernel's partial sums
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
