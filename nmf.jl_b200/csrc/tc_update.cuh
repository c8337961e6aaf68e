// tc_update.cuh -- device state, the update kernel (mu_update_kernel, all modes), the stand-alone Gram kernel and the Gram reduce
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

// Batched replicates (interf.jl:85-101 as ONE pass over X per half-step for all replicates, tc_engine.cu::tc_solve_batched_kp):
// replicate r owns components [r*blk, (r+1)*blk) of the stacked factors; stop_condition (common.jl:92-111) is applied per replicate.
constexpr int MAX_BATCH = 32;
struct BatchState {
    int done[MAX_BATCH];     // the replicate has met stop_condition (its factors were snapshotted at that iteration)
    int newly[MAX_BATCH];    // ... in the iteration just decided
    int niters[MAX_BATCH];   // iteration count at which it did
    float devmax[MAX_BATCH]; // `dev` of its last stop_condition call
};

struct TcState {
    int converged;
    int iters;
    float devmax;
    unsigned int ticket;
    unsigned int ticket2;  // all-blocks ticket of gram_conv_reduce_kernel (trace-identity objective)
    unsigned int pad_;
    double objv;           // verbose: objective of the iteration just finished, by the trace identity (tc_reduce.cuh)
    // batched replicates (all zero otherwise): the Gram reduce keeps only the diagonal blk x blk blocks of the KP x KP Gram (kp = KP),
    // and the stop decision is taken per replicate in *bs; `converged` is raised once every replicate is done
    int blk, kp, nrep;
    unsigned int chain;    // option tc_chain: tiles completed by all chained update launches of this solve
    unsigned int skew[3];  // option tc_skew: the same per group (A, B) of tiles, and A CTAs past the middle of their k-loop
    unsigned int pad3_;
    BatchState* bs;
};

// Gram element i = row * kp + col of a stacked factor: does it couple two different replicates?
__device__ __forceinline__ bool gram_masked(const TcState* st, int i) {
    const int blk = st->blk;
    if (blk <= 0) return false;
    const int kp = st->kp;
    return (i / kp) / blk != (i % kp) / blk;
}

// ---- row-sharded solves (multi-GPU, or G logical shards on one GPU): see tc_shard.cuh -----------------------
// Flag phases of the per-rank arena (flags[phase * XCHG_MAX_RANKS + src] = last epoch `src` has published):
constexpr int PH_NUM = 0;     // src's partial numerators for MY H rows are in my slot[src]
constexpr int PH_H = 1;       // src's updated H rows (bf16 transposed slab), its partial Gram H'H and H-side stop sums are here
constexpr int PH_PW = 2;      // src's partial Gram W'W and W-side stop sums are here
constexpr int PH_BAR = 3;     // plain barrier (rank alignment before a timed region)
constexpr int PH_GATHER = 4;  // src's fp32 H rows are here (end of solve / verbose)
constexpr int PH_HBT = 5;     // src's updated rows of H'^T (bf16 slab) are in my copy: the W-step may start streaming
constexpr int N_PHASES = 6;

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
    CUtensorMap tmAlo;  // precision mode bf16x3: the remainder panel Xs - bf16(Xs), same geometry as tmA
    CUtensorMap tmBlo;  //   ... and the transposed remainder of the other factor, same geometry as tmB
    CUtensorMap tmAnext; // MODE 0, pf_blocks > 0: the X panel the NEXT update launch streams (same geometry rules as tmA)
    int pf_blocks;      // > 0: once its own loads are issued, the producer asks L2 for the first pf_blocks k-blocks of the tile that CTA
                        // blockIdx.x of the next launch will stream -- HBM is otherwise idle while all CTAs sit in their epilogues
    int pf_tiles, pf_tile_rows, pf_panel_rows;   // next launch: tiles, rows per tile, rows of the tile-contiguous panel per tile (nkb * tile_rows)
    // option tc_chain (MODE 0, single GPU, staged epilogue): the whole iteration is one chain of programmatic dependents, and the hand-over
    // from one update launch to the next does not wait for a kernel boundary (measured: griddepcontrol.wait in the kernel behind returns
    // ~5 us after the last CTA has exited).  Every CTA counts itself in *chain_cnt once its bulk stores -- the new transposed tile among
    // them -- are complete; the next update launch, whose CTAs take their seats as soon as this launch's loads are issued (early_trigger
    // lets the reduce kernel in between become resident, and that kernel releases the next launch at its top), polls
    // *chain_flag >= chain_need (the tiles of all chained update launches before it) before it touches the other factor's transposed copy.
    const unsigned int* chain_flag;
    unsigned int* chain_cnt;
    unsigned int chain_need;
    int early_trigger;  // 1: launch_dependents once the last load is issued; 2: at the top of the kernel (tc_skew)
    // option tc_skew (on top of tc_chain): the tiles of a launch form two groups, A = [0, tile_split) and B, that run HALF A PERIOD APART, so
    // that one group streams while the other sits in its epilogues (the one phase in which HBM idled).  A launch reads the other factor's
    // group-A tiles in k-blocks [0, kb_split) and its group-B tiles behind them, so a CTA needs cnt A >= need_a to start and cnt B >= need_b
    // only at the middle of its k-loop -- exactly when the late group of the launch before it is done.  The offset is enforced, not hoped
    // for: a B CTA starts only when every A CTA of its own launch has passed its middle (skew_cnt[2] >= need_mid).  Every CTA triggers
    // launch_dependents at its top, so the CTAs of the next launch enter one by one as SMs become free.
    unsigned int* skew_cnt;   // [0] tiles of group A complete, [1] of group B, [2] A CTAs past their middle (all cumulative over launches)
    unsigned int need_a, need_b, need_mid;
    int tile_split, kb_split;
    int diag_nob;       // diagnostics (tc_debug bit 8, skewed launches only): skip the B operand loads behind the first k-block
    int flush_chunk;    // > 0 (KP <= 128): the numerator MMAs accumulate at most this many k-blocks in TMEM; the epilogue warps add
                        // each finished chunk to fp32 register sums (round to nearest) while the next chunk accumulates.  The
                        // tensor core's accumulator TRUNCATES (measured: ~0.5 ulp lost per MMA, a relative bias of ~3e-8 per
                        // step that adds up over the 1000+ steps of a long contraction); short chains keep the bias at 1e-6.
    float* cross_part;  // verbose W-step: [tiles] per-CTA sums of Num .* F_new = this tile's share of <X H', W> (trace identity; nullptr = skip)
    int x3;             // 1: numerators = A*B + A*Blo + Alo*B (three passes over the k-blocks), ~2^-16 relative instead of 2^-8
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
    long long* timing;  // diagnostics (tc_debug bit 3): CTA timing_cta records clock64() at its phase boundaries, see TSTAMP
    int timing_cta;
    long long* gtl;     // diagnostics (tc_debug bit 7): EVERY CTA records %globaltimer at 8 points of its life, [8 * blockIdx.x + i]: 0 entry,
                        // 1 producer released (chain counter seen), 2 last load issued, 3 last MMA issued, 4 accumulators complete,
                        // 5 epilogue warps done, 6 bulk stores complete (tile counted), 7 exit
    float lambda, delta;
    // ---- row-sharded solves (tc_shard.cuh); all zero for a single-GPU launch
    int tile0;          // first tile of this launch (MODE 2: the rank's own H rows only)
    int wait_first;     // producer: griddepcontrol.wait before the first operand load (W-step under PDL: H arrives from peers)
    int num_row0;       // MODE 2: numerator slots are indexed by (row - num_row0)
    int G;              // number of ranks (0 = not sharded)
    int tiles_per_owner;                       // MODE 1: tile t belongs to rank t / tiles_per_owner
    int tiles_total;
    unsigned int epoch;
    float* num_peer[XCHG_MAX_RANKS];           // MODE 1: [owner] -> slot[my rank] in the owner's arena (peer memory)
    unsigned int* num_flag[XCHG_MAX_RANKS];    // MODE 1: [owner] -> flags[PH_NUM][my rank] in the owner's arena
    unsigned int* own_cnt;                     // MODE 1: [G] local counters "tiles of owner o finished"
    int n_peer;                                // MODE 2: transposed tile is also stored into n_peer peer copies of F^T
    CUtensorMap tmT_peer[XCHG_MAX_RANKS - 1];
    CUtensorMap tmNum[XCHG_MAX_RANKS];         // MODE 1: [owner] -> this rank's slot in the owner's arena, fp32 [slot_rows][KP], box 32 x tile_rows
    const unsigned int* num_wait;              // MODE 2: this rank's PH_NUM flag row (G entries): wait for `epoch` before reading the slots
    int rank;                                  // this rank
    unsigned int* hbt_cnt;                     // MODE 2 / 6: local counter "own tiles finished"
    unsigned int* hbt_flag[XCHG_MAX_RANKS];    // MODE 2 / 6: [j] -> flags[PH_HBT][my rank] in rank j's arena
    const unsigned int* hbt_wait;              // MODE 0 (W-step): this rank's PH_HBT flag row: the producer waits for `epoch` from every
                                               // rank before its first operand load (the rows of H'^T arrive from the peers)
    int defer_signal;                          // MODE 1 / 2: do not wait for the peer stores and raise no flag here -- the kernel behind
                                               // this one (kernel boundary = all stores performed) publishes NUM / HBT
    int signal_hbt;                            // MODE 0 (W-step): CTA 0 raises PH_HBT for this rank at its start (deferred from MODE 2)
    const unsigned int* den_flag;              // MODE 0 (W-step): local flag "the Gram of the other factor for `epoch` is in place" -- set by
                                               // shard_post_kernel, which runs on a side stream concurrently with this kernel's main loop
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

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
// MODE 6: row-sharded fused H-step (tc_shard.cuh): every CTA computes the partial numerators of one H tile over this rank's rows of
//         X; a CTA whose tile belongs to ANOTHER rank sends it there (as MODE 1) and exits; a CTA whose tile this rank OWNS waits
//         for the other ranks' partials, adds them to its own (rank order) and finishes like MODE 0 -- ratio, the new rows in all
//         forms (the transposed bf16 tile also into every peer's copy), tile Gram, stop sums.  Tiles are visited starting behind
//         the own range, so the CTAs that only send are dispatched before the CTAs that wait.
template <int KP, int MODE>
__global__ void __launch_bounds__(UpdCfg<KP>::THREADS, 1) mu_update_kernel(const __grid_constant__ UpdateParams prm) {
    using C = UpdCfg<KP>;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + C::RING_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full = empty_bar + C::STAGES;
    uint64_t* gram_bar = tmem_full + 1;
    uint64_t* acc_full = gram_bar + 1;    // [2] chunk buffer b holds a finished chunk of the numerator (flush_chunk mode)
    uint64_t* acc_empty = acc_full + 2;   // [2] ... has been added to the register sums
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    uint32_t* stop_slot = tmem_slot + 1;
    float* conv_s = (float*)(smem + C::RING_BYTES + 1024);  // [4 warps][2][KP]
    // Staged epilogue (KP <= 128, modes that write the factor): the ring is idle once the accumulators are complete
    constexpr bool STAGED = (KP <= 128) && (MODE == 0 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6);
    constexpr bool FUSED = (MODE == 6);
    uint8_t* const SF = smem;                              // fp32 tile:  KP/32 boxes of 128 rows x 128 B
    uint8_t* const SH = SF + (KP / 32) * 16384;            // bf16 hi:    KP/64 boxes
    uint8_t* const SL = SH + (KP / 64) * 16384;            // bf16 lo
    uint8_t* const ST = SL + (KP / 64) * 16384;            // transposed: 2 boxes of KP rows x 128 B (64 tile rows each)
    static_assert(!STAGED || (KP / 32 + 2 * (KP / 64)) * 16384 + 2 * KP * 128 <= C::RING_BYTES, "staging does not fit in the ring");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define TSTAMP(i) do { if (prm.timing != nullptr && (int)blockIdx.x == prm.timing_cta) prm.timing[i] = clock64(); } while (0)
#define GSTAMP(i) do { if (prm.gtl != nullptr) { long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); prm.gtl[8 * blockIdx.x + (i)] = g_; } } while (0)
    if (threadIdx.x == 0) GSTAMP(0);
    if (threadIdx.x == 0) TSTAMP(0);
    if (prm.timing != nullptr && threadIdx.x == 0) {  // every CTA: global timer at entry (and exit, below)
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.timing[16 + 2 * blockIdx.x] = gt;
    }
    const int tile_rows = prm.tile_rows;
    int tile_ = (int)blockIdx.x + prm.tile0;
    if (FUSED && tile_ >= prm.tiles_total) tile_ -= prm.tiles_total;
    const int tile = tile_;
    const bool owner = !FUSED || (tile / prm.tiles_per_owner == prm.rank);     // CTA-uniform
    const bool push = (MODE == 1) || (FUSED && !owner);                         // this CTA only sends its numerators away
    const int out_idx = FUSED ? tile - prm.rank * prm.tiles_per_owner : (int)blockIdx.x;   // index into gram_part / conv_part
    const int r0 = tile * tile_rows;
    const uint32_t a_bytes = (uint32_t)tile_rows * 128u;
    const int nkb = (MODE == 2 || MODE == 5) ? 0 : (prm.Kdim + 63) / 64;
    const int npass = prm.x3 ? 3 : 1;   // precision mode bf16x3: hi*hi, hi*lo, lo*hi
    const int flush_ch = (KP <= 128 && nkb > 0) ? prm.flush_chunk : 0;
    const int nchunks = flush_ch > 0 ? (nkb * npass + flush_ch - 1) / flush_ch : 0;
    // TMEM columns of the finished numerators / denominators: [0, KP) / [KP, 2KP), or -- chunked -- the buffer of the last chunk
    // (the register sums are written back there) / the other one
    const uint32_t num_col = flush_ch > 0 ? (uint32_t)(((nchunks - 1) & 1) * KP) : 0u;
    const uint32_t den_col = (uint32_t)KP - num_col;
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
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (MODE == 0 && prm.early_trigger == 2 && threadIdx.x == 0) pdl_launch_dependents();   // tc_skew: successors enter as SMs become free
    if (*stop_slot != 0u) {  // uniform early exit
        if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
        return;
    }
    // Block order: the nkb numerator blocks FIRST (they depend on nothing the preceding kernel writes), then the NPRE
    // denominator blocks (the Gram hi/lo they read is produced by the immediately preceding reduce kernel).
    // The two single-thread loops below are the latency-critical part of the kernel: no per-block branches, no
    // div/mod, everything loop-invariant is hoisted (an extra compare per block is measurable at 256 blocks).

    if (MODE == 0 && prm.signal_hbt && blockIdx.x == 0 && threadIdx.x == 96) {
        // row-sharded W-step: the ratio kernel in front of us (same stream, complete) stored this rank's rows of H'^T into every
        // rank's copy without waiting for the stores; publish them now (one thread of an epilogue warp, off the producer's path)
        __threadfence_system();
        for (int j = 0; j < prm.G; ++j) st_release_sys(prm.hbt_flag[j], prm.epoch);
    }
    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            uint8_t* dst = smem;
            int arow = tile * nkb * tile_rows;   // tile-contiguous X: k-block kb of this tile starts at panel row arow0 + kb*tile_rows
            const uint32_t num_tx = a_bytes + (uint32_t)C::B_BYTES;
            if (prm.wait_first) pdl_wait();   // sharded W-step: the other factor's rows arrive from peers, confirmed by the kernel in front
            if (MODE == 0 && prm.hbt_wait != nullptr) {   // ... or by the peers' PH_HBT flags themselves
                for (int j = 0; j < prm.G; ++j) {
                    const long long t0 = clock64();
                    while ((int)(ld_acquire_sys(prm.hbt_wait + j) - prm.epoch) < 0) {
                        if (clock64() - t0 > 20000000000LL) { printf("nmfb200: H'^T flag wait timed out (rank %d, epoch %u)\n", j, prm.epoch); __trap(); }
                    }
                }
                asm volatile("fence.proxy.async;" ::: "memory");   // the peers' writes -> the TMA loads below
            }
            if (MODE == 0 && prm.chain_flag != nullptr) {
                const long long t0 = clock64();
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prm.chain_flag) : "memory");
                    if ((int)(seen - prm.chain_need) >= 0) break;
                    // a predecessor that met stop_condition skips its epilogue and never counts itself: nothing we compute will be kept
                    if (__ldcg(&prm.state->converged) != 0) break;
                    if (clock64() - t0 > 20000000000LL) { printf("nmfb200: chain counter wait timed out (need %u, seen %u)\n", prm.chain_need, seen); __trap(); }
                } while (true);
                asm volatile("fence.proxy.async;" ::: "memory");   // the predecessor's bulk stores -> the TMA loads below
            }
            const bool skewed = MODE == 0 && prm.skew_cnt != nullptr;
            // bounded poll of a cumulative counter; also leaves when stop_condition was met (launches that see it publish nothing)
            auto poll = [&](const unsigned int* p, unsigned int need) {
                const long long t0 = clock64();
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p) : "memory");
                    if ((int)(seen - need) >= 0) break;
                    if (__ldcg(&prm.state->converged) != 0) break;
                    if (clock64() - t0 > 20000000000LL) { printf("nmfb200: skew counter wait timed out (need %u, seen %u)\n", need, seen); __trap(); }
                } while (true);
                asm volatile("fence.proxy.async;" ::: "memory");
            };
            if (skewed) {
                const bool grp_b = tile >= prm.tile_split;
                if (grp_b) poll(prm.skew_cnt + 2, prm.need_mid);   // stay half a period behind group A
                poll(prm.skew_cnt + 0, prm.need_a);
                GSTAMP(1);
                const int kb1 = min(prm.kb_split, nkb);
                const bool nob = prm.diag_nob != 0;   // diagnostics (tc_debug bit 8): B operand loaded for the first STAGES k-blocks only (WRONG results)
                for (int kb = 0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const bool ldb = !nob || kb < C::STAGES;   // every stage's B half holds finite data
                    mbar_arrive_expect_tx(&full_bar[s], ldb ? num_tx : a_bytes);
                    tma_load_2d(dst, &prm.tmA, &full_bar[s], 0, arow);
                    if (ldb) tma_load_2d(dst + C::A_BYTES, &prm.tmB, &full_bar[s], 64 * kb, 0);
                    arow += tile_rows;
                    dst += C::STAGE_BYTES;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; dst = smem; }
                }
                if (!grp_b) atomicAdd(prm.skew_cnt + 2, 1u);        // group A: past the middle
                poll(prm.skew_cnt + 1, prm.need_b);
                for (int kb = kb1; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], nob ? a_bytes : num_tx);
                    tma_load_2d(dst, &prm.tmA, &full_bar[s], 0, arow);
                    if (!nob) tma_load_2d(dst + C::A_BYTES, &prm.tmB, &full_bar[s], 64 * kb, 0);
                    arow += tile_rows;
                    dst += C::STAGE_BYTES;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; dst = smem; }
                }
            }
            if (!skewed) GSTAMP(1);
            const int arow0 = arow;
            for (int pass = 0; pass < (skewed ? 0 : npass); ++pass) {   // one pass unless precision mode bf16x3
                const CUtensorMap* mA = pass == 2 ? &prm.tmAlo : &prm.tmA;
                const CUtensorMap* mB = pass == 1 ? &prm.tmBlo : &prm.tmB;
                arow = arow0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[s], num_tx);
                    tma_load_2d(dst, mA, &full_bar[s], 0, arow);
                    tma_load_2d(dst + C::A_BYTES, mB, &full_bar[s], 64 * kb, 0);
                    arow += tile_rows;
                    dst += C::STAGE_BYTES;
                    if (++s == C::STAGES) { s = 0; ph ^= 1u; dst = smem; }
                }
            }
            if (NPRE > 0 && owner) {
                pdl_wait();  // the Gram of the other factor comes from the preceding (reduce) kernel
                if (MODE == 0 && prm.den_flag != nullptr) {   // ... or, row-sharded, from a kernel on the side stream: wait for its flag
                    const long long t0 = clock64();
                    unsigned int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prm.den_flag) : "memory");
                        if (clock64() - t0 > 20000000000LL) { printf("nmfb200: Gram flag wait timed out (epoch %u)\n", prm.epoch); __trap(); }
                    } while ((int)(seen - prm.epoch) < 0);
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of P hi/lo -> the TMA loads below
                }
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
            GSTAMP(2);
            if (MODE == 0 && prm.early_trigger == 1) pdl_launch_dependents();   // every load is issued: the reduce kernel behind us may take its seats
            if (MODE == 0 && prm.pf_blocks > 0 && (int)blockIdx.x < prm.pf_tiles) {
                // every load of this launch is in flight: warm L2 with the head of the next launch's panel (no smem destination,
                // no completion tracking; a converged solve wastes them harmlessly)
                int prow = (int)blockIdx.x * prm.pf_panel_rows;
                for (int i = 0; i < prm.pf_blocks; ++i) {
                    tma_prefetch_2d(&prm.tmAnext, 0, prow);
                    prow += prm.pf_tile_rows;
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
            const int nkb_all = nkb * npass;
            if (flush_ch > 0) {
                // chunked accumulation: chunk c goes to buffer c & 1; a buffer is reused once the epilogue warps have drained it
                int buf = 0;
                for (int c = 0; c < nchunks; ++c) {
                    const int kend = min(kb + flush_ch, nkb_all);
                    if (c >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)((c >> 1) - 1) & 1u); tc_fence_after(); }
                    block(tmem_base + buf * KP, 0u);
                    if (c == 0) TSTAMP(1);
                    for (++kb; kb < kend; ++kb) block(tmem_base + buf * KP, 1u);
                    umma_commit(&acc_full[buf]);
                    buf ^= 1;
                }
                TSTAMP(2);
                if (NPRE > 0 && owner) {   // denominators: the buffer that does NOT hold the last chunk (= den_col)
                    if (nchunks >= 2) { mbar_wait(&acc_empty[buf], (uint32_t)(((nchunks - 2) >> 1)) & 1u); tc_fence_after(); }
                    block(tmem_base + den_col, 0u);
#pragma unroll 1
                    for (int bd = 1; bd < NPRE; ++bd) block(tmem_base + den_col, 1u);
                }
            } else {
            if (nkb > 0) { block(tmem_base, 0u); kb = 1; TSTAMP(1); }   // first operands have landed
            for (; kb < nkb_all; ++kb) block(tmem_base, 1u);
            TSTAMP(2);                                                   // numerator blocks issued
            if (NPRE > 0 && owner) {
                block(tmem_base + KP, 0u);
#pragma unroll 1
                for (int bd = 1; bd < NPRE; ++bd) block(tmem_base + KP, 1u);
            }
            }
            if (MODE != 5) umma_commit(tmem_full);
            TSTAMP(3);                                                   // all MMAs issued
            GSTAMP(3);
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4, columns [chalf*KP/2, (chalf+1)*KP/2) =====
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int row = r0 + 32 * q + lane;
        const bool valid = (32 * q + lane) < tile_rows && row < prm.R;
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
        if constexpr (KP <= 128) {
            if (flush_ch > 0) {
                // chunked accumulation: add every finished chunk to fp32 register sums (this thread: its row, its column half)
                float nsum[KP / 2];
#pragma unroll
                for (int j = 0; j < KP / 2; ++j) nsum[j] = 0.f;
                int buf = 0;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&acc_full[buf], (uint32_t)(c >> 1) & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < KP / 2; cc += 32) {
                        uint32_t v[32];
                        tmem_ld32(t_lane + buf * KP + chalf * (KP / 2) + cc, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) nsum[cc + j] += __uint_as_float(v[j]);
                    }
                    if (c + 1 < nchunks) {   // the last chunk's buffer stays with us (the sums go back into it)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);
                    }
                    buf ^= 1;
                }
#pragma unroll
                for (int cc = 0; cc < KP / 2; cc += 32) {
                    uint32_t v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(nsum[cc + j]);
                    tmem_st32(t_lane + num_col + chalf * (KP / 2) + cc, v);
                }
                tmem_st_wait();
            }
        }
        pdl_wait();  // from here on we read / overwrite what the preceding kernel wrote / read
        const bool stop = __ldcg(&prm.state->converged) != 0;  // uniform: the preceding kernel is complete
        if (threadIdx.x == 64) TSTAMP(4);    // preceding kernel complete
        if (MODE != 5) mbar_wait(tmem_full, 0);   // (parking the epilogue warps in a named barrier instead of this poll was measured: no difference)
        tc_fence_after();
        if (threadIdx.x == 64) { TSTAMP(5); GSTAMP(4); }   // accumulators complete
        do {
        if (stop) break;  // converged while this kernel was streaming (PDL): leave F untouched
        if ((MODE == 2 || (FUSED && owner)) && prm.G > 0 && prm.num_wait != nullptr) {
            // row-sharded H-step: every rank's partial numerators for these rows must have landed in this rank's slots
            // (MODE 6: every OTHER rank's -- this rank's own partial is in TMEM)
            const int t = (int)threadIdx.x - 64;
            if (t < prm.G && !(FUSED && t == prm.rank)) {
                const long long t0 = clock64();
                while ((int)(ld_acquire_sys(prm.num_wait + t) - prm.epoch) < 0) {
                    if (clock64() - t0 > 20000000000LL) {
                        printf("nmfb200: numerator flag wait timed out (waiting for rank %d, epoch %u)\n", t, prm.epoch);
                        __trap();
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) TSTAMP(12);   // the other ranks' partial numerators have arrived
        }
        float* convw = conv_s + q * 2 * KP;
        float cross_acc = 0.f;
        float* cross_s = (float*)(smem + C::RING_BYTES + 512);   // [8 epilogue warps], behind the barriers
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
            if (MODE != 2 && MODE != 5) tmem_ld32(t_lane + num_col + c0, num_u);
            if (MODE != 1 && MODE != 4 && MODE != 5 && !push) tmem_ld32(t_lane + den_col + c0, den_u);
            if (push) {
                tmem_ld_wait();
                if (prm.G > 0) {
                    // row-sharded: the tile goes to the slot its OWNER keeps for this rank.  Staged in the (idle) ring as the
                    // SWIZZLE_128B image and sent by bulk TMA stores below: NVLink carries 16 KB bursts, not 16-byte stores.
                    const int rr = 32 * q + lane;
                    uint8_t* sf = SF + (c0 >> 5) * 16384 + rr * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *(uint4*)(sf + ((j ^ (rr & 7)) << 4)) = make_uint4(num_u[4 * j], num_u[4 * j + 1], num_u[4 * j + 2], num_u[4 * j + 3]);
                } else if (valid) {
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
                if (MODE == 5 || MODE == 2) {
                    // MODE 5: k-split partial numerators of div_fused_kernel; MODE 2: the per-rank partial numerators of a
                    // row-sharded H-step (slot s = rank s).  Summed in slot order => deterministic, same on every rank.
                    const float* nbase = prm.num_io + (size_t)(row - prm.num_row0) * KP + c0;
                    float acc[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = __ldcg((const float4*)nbase + j);
                        acc[4 * j] = v.x; acc[4 * j + 1] = v.y; acc[4 * j + 2] = v.z; acc[4 * j + 3] = v.w;
                    }
                    int sp = 1;
                    for (; sp + 1 < prm.num_splits; sp += 2) {   // two slots per trip: 16 independent 16-byte loads in flight
                        const float4* n0 = (const float4*)(nbase + (size_t)sp * prm.num_split_stride);
                        const float4* n1 = (const float4*)(nbase + (size_t)(sp + 1) * prm.num_split_stride);
                        float4 v0[8], v1[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { v0[j] = __ldcg(n0 + j); v1[j] = __ldcg(n1 + j); }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {   // slot order is kept: (acc + slot sp) + slot sp+1
                            acc[4 * j] = (acc[4 * j] + v0[j].x) + v1[j].x; acc[4 * j + 1] = (acc[4 * j + 1] + v0[j].y) + v1[j].y;
                            acc[4 * j + 2] = (acc[4 * j + 2] + v0[j].z) + v1[j].z; acc[4 * j + 3] = (acc[4 * j + 3] + v0[j].w) + v1[j].w;
                        }
                    }
                    for (; sp < prm.num_splits; ++sp) {
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
            if (FUSED) {
                // numerator = sum over the ranks IN RANK ORDER of their partials: slot s for s != rank, TMEM for s == rank
                if (valid) {
                    const float* nbase = prm.num_io + (size_t)(row - prm.num_row0) * KP + c0;
                    float acc[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
                    for (int sp = 0; sp < prm.num_splits; ++sp) {
                        if (sp == prm.rank) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(num_u[j]);
                        } else {
                            const float4* ns = (const float4*)(nbase + (size_t)sp * prm.num_split_stride);
                            float4 v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = __ldcg(ns + j);
#pragma unroll
                            for (int j = 0; j < 8; ++j) { acc[4 * j] += v[j].x; acc[4 * j + 1] += v[j].y; acc[4 * j + 2] += v[j].z; acc[4 * j + 3] += v[j].w; }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) num_u[j] = __float_as_uint(acc[j]);
                }
            }
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
            const bool want_cross = MODE == 0 && prm.cross_part != nullptr;
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
                    if (want_cross) cross_acc += __uint_as_float(num_u[j + e]) * fn[e];   // <X H', W_new>, this row, this column
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
        if (push && prm.G > 0) {
            // send the staged tile, count it, and the last tile for an owner raises that owner's NUM flag
            fence_proxy_async();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                const int owner_rank = tile / prm.tiles_per_owner;
                const int srow = (tile % prm.tiles_per_owner) * tile_rows;
#pragma unroll
                for (int b = 0; b < KP / 32; ++b) tma_store_2d(&prm.tmNum[owner_rank], SF + b * 16384, 32 * b, srow);
                tma_store_commit();
                if (MODE == 1 && prm.defer_signal) {
                    tma_store_wait_read<0>();   // the staging buffers have been read; completion is the kernel boundary's business
                } else {
                    tma_store_wait_all<0>();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence_system();
                    const int owned = min(prm.tiles_per_owner, prm.tiles_total - owner_rank * prm.tiles_per_owner);
                    const unsigned prev = atomicAdd(prm.own_cnt + owner_rank, 1u);
                    if (prev == (unsigned)owned - 1u) {
                        prm.own_cnt[owner_rank] = 0u;
                        __threadfence_system();
                        st_release_sys(prm.num_flag[owner_rank], prm.epoch);
                    }
                }
            }
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
        } else if (MODE != 1 && !push) {
            if constexpr (STAGED) {
                fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA / tensor-core (async) proxy
                tc_fence_before();     // our TMEM reads are complete (the Gram below reuses the Num columns)
            }
            if (threadIdx.x == 64) TSTAMP(6);  // this warp's ratio / staging done
            if (MODE == 0 && prm.cross_part != nullptr) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cross_acc += __shfl_xor_sync(0xffffffffu, cross_acc, o);
                if (lane == 0) cross_s[warp - 2] = cross_acc;
            }
            // combine the four lane quarters: named barrier over the 256 epilogue threads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (MODE == 0 && prm.cross_part != nullptr && threadIdx.x == 64) {
                float c = 0.f;
                for (int i = 0; i < 8; ++i) c += cross_s[i];   // fixed order
                prm.cross_part[out_idx] = c;
            }
            if constexpr (STAGED) {
                if (threadIdx.x == 64) {
                    TSTAMP(7);                 // all epilogue warps done
                    GSTAMP(5);
#pragma unroll
                    for (int b = 0; b < KP / 32; ++b) tma_store_2d(&prm.tmF32, SF + b * 16384, 32 * b, r0);
#pragma unroll
                    for (int b = 0; b < KP / 64; ++b) {
                        tma_store_2d(&prm.tmFhi, SH + b * 16384, 64 * b, r0);
                        tma_store_2d(&prm.tmFlo, SL + b * 16384, 64 * b, r0);
                    }
                    tma_store_2d(&prm.tmT, ST, r0, 0);
                    if (tile_rows > 64) tma_store_2d(&prm.tmT, ST + KP * 128, r0 + 64, 0);
                    if (MODE == 2 || FUSED) {   // row-sharded H-step: the all-gather of the new rows is these stores into the peers' copies
                        for (int j = 0; j < prm.n_peer; ++j) {
                            tma_store_2d(&prm.tmT_peer[j], ST, r0, 0);
                            if (tile_rows > 64) tma_store_2d(&prm.tmT_peer[j], ST + KP * 128, r0 + 64, 0);
                        }
                    }
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
                prm.conv_part[(size_t)out_idx * 2 * KP + i] = s;
            }
            if constexpr (STAGED) {
                if (prm.gram_part != nullptr) {
                    mbar_wait(gram_bar, 0);
                    tc_fence_after();
                    if (threadIdx.x == 64) TSTAMP(8);  // tile Gram MMAs complete
                    const int a = 32 * q + lane;
                    float* gp = prm.gram_part + ((size_t)out_idx * KP + a) * KP;
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
                    if (MODE == 2 && prm.defer_signal) tma_store_wait_read<0>();
                    else tma_store_wait_all<0>();
                    if (MODE == 0 && prm.chain_cnt != nullptr) {   // option tc_chain: this tile is final in all four forms
                        asm volatile("fence.proxy.async;" ::: "memory");
                        __threadfence();
                        atomicAdd(prm.chain_cnt, 1u);
                    }
                    if (MODE == 0 && prm.skew_cnt != nullptr) {    // option tc_skew: the same, per group
                        asm volatile("fence.proxy.async;" ::: "memory");
                        __threadfence();
                        atomicAdd(prm.skew_cnt + (tile >= prm.tile_split ? 1 : 0), 1u);
                    }
                    if ((MODE == 2 || FUSED) && prm.G > 0 && !(MODE == 2 && prm.defer_signal)) {
                        // the peers' copies were written through the async proxy: order them, count this tile, and the last own
                        // tile tells every rank that this rank's rows of H'^T are in place (PH_HBT)
                        asm volatile("fence.proxy.async;" ::: "memory");
                        __threadfence_system();
                        const unsigned own_n = FUSED ? (unsigned)min(prm.tiles_per_owner, prm.tiles_total - prm.rank * prm.tiles_per_owner) : gridDim.x;
                        if (atomicAdd(prm.hbt_cnt, 1u) == own_n - 1u) {
                            *prm.hbt_cnt = 0u;
                            __threadfence_system();
                            for (int j = 0; j < prm.G; ++j) st_release_sys(prm.hbt_flag[j], prm.epoch);
                        }
                    }
                    TSTAMP(10);                // bulk stores have read their staging buffers
                    GSTAMP(6);
                }
            }
        }
        } while (0);
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) { TSTAMP(11); GSTAMP(7); }
    if (prm.timing != nullptr && threadIdx.x == 0) {
        long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        prm.timing[16 + 2 * blockIdx.x + 1] = gt;
    }
#undef TSTAMP
#undef GSTAMP
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ---- Gram: P += T T'  for T = FbT ([KP][R] bf16, rows of length R contiguous) ------------------------
struct GramParams {
    CUtensorMap tmTlo;  // precision mode bf16x3: the transposed remainder, same geometry as tmT
    int split;        // 1: T T' = hi hi' + hi lo' + lo hi'
    CUtensorMap tmT;  // bf16 [KP][R], box 64 x 128
    float* part;      // [gridDim.x][KP][KP] fp32 partial Grams (plain stores, reduced by gram_reduce_kernel)
    const TcState* state;
    int R, chunk;     // rows (K extent) per CTA, multiple of 64
    int k0, k1;       // the Gram covers rows [k0, k1) of the factor (k1 = R: all; a row-sharded solve sums its own H rows only)
};

template <int KP>
struct GramCfg {
    static constexpr int MT = (KP + 127) / 128;         // 128-row M tiles
    static constexpr int TILE_BYTES = MT * 128 * 128;   // the tile is both A and B operand
    static constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // hi tile | lo tile (the lo half is only filled in split mode)
    static constexpr int STAGES = KP == 256 ? 3 : 4;
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
    const int k_begin = prm.k0 + blockIdx.x * prm.chunk;
    const int k_end = min(prm.k1, k_begin + prm.chunk);
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
                mbar_arrive_expect_tx(&full_bar[s], prm.split ? C::STAGE_BYTES : C::TILE_BYTES);
                // NOTE: columns >= k_end inside the last 64-block belong to the next CTA's chunk only if
                // chunk % 64 != 0; chunk, k0 and k1 (unless k1 = R) are multiples of 64, and columns >= R are zero-filled by TMA.
                for (int m = 0; m < C::MT; ++m) {
                    tma_load_2d(smem + s * C::STAGE_BYTES + m * 128 * 128, &prm.tmT, &full_bar[s], k_begin + 64 * b, 128 * m);
                    if (prm.split)
                        tma_load_2d(smem + s * C::STAGE_BYTES + C::TILE_BYTES + m * 128 * 128, &prm.tmTlo, &full_bar[s], k_begin + 64 * b, 128 * m);
                }
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
                const int nterm = prm.split ? 3 : 1;   // (A, B) = (hi, hi), (hi, lo), (lo, hi)
                for (int term = 0; term < nterm; ++term) {
                    const uint32_t abase = base + (term == 2 ? C::TILE_BYTES : 0), bbase = base + (term == 1 ? C::TILE_BYTES : 0);
                    const uint64_t bdesc = make_kmajor_sw128_desc(bbase);  // B = first KP rows of the tile
#pragma unroll
                    for (int m = 0; m < C::MT; ++m) {
                        const uint64_t adesc = make_kmajor_sw128_desc(abase + m * 128 * 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(tmem_base + m * KP, adesc + 2 * kk, bdesc + 2 * kk, idesc, (b > 0 || kk > 0 || term > 0) ? 1u : 0u);
                    }
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
                                                          const TcState* st, int chained, int nvb) {
    // The update kernel behind us may start streaming X as soon as every block has passed this point; it waits for our
    // completion before it reads P.  (Pre-launching THIS kernel behind the running update kernel was measured too:
    // its resident blocks polling in griddepcontrol.wait slow the single-thread TMA / MMA loops, 4770 -> 4400 it/s.)
    pdl_launch_dependents();
    // option tc_chain: this kernel was itself launched as a programmatic dependent (resident since the update kernel in front of it
    // issued its last loads): wait for that kernel to complete before reading its tile Grams.  It then runs as a few CTAs that walk
    // the nvb virtual blocks of 256 threads (gridDim.x == nvb otherwise).
    if (chained) pdl_wait();
    if (st->converged) return;
    for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        const int t = vb * blockDim.x + threadIdx.x;
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
            if (gram_masked(st, i)) acc = 0.f;   // batched replicates: block-diagonal Gram
            P[i] = acc;
            if (do_split) {
                bf16 hi = __float2bfloat16_rn(acc);
                Phi[i] = hi;
                Plo[i] = __float2bfloat16_rn(acc - __bfloat162float(hi));
            }
        }
    }
}
