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
#include "tc_update.cuh"
#include "tc_reduce.cuh"
#include "tc_layout.cuh"

// Extra launch arguments of mu_update_kernel in a row-sharded solve (tc_shard.cuh); nullptr for a single-GPU launch.
struct ShardLaunch {
    int tile0 = 0, wait_first = 0, num_row0 = 0;
    int G = 0, tiles_per_owner = 1, tiles_total = 0;
    unsigned int epoch = 0;
    float* num_peer[XCHG_MAX_RANKS] = {};
    unsigned int* num_flag[XCHG_MAX_RANKS] = {};
    unsigned int* own_cnt = nullptr;
    int n_peer = 0;
    bf16* peer_bT[XCHG_MAX_RANKS - 1] = {};   // MODE 2: the peers' copies of F^T ([rowsT][ldT], same geometry as F.bT)
    int64_t slot_rows = 0;                    // MODE 1: rows of one numerator slot (num_peer[o] is fp32 [slot_rows][KP])
    const unsigned int* num_wait = nullptr;   // MODE 2: this rank's PH_NUM flag row
    const unsigned int* den_flag = nullptr;   // MODE 0: local "Gram of the other factor is in place" flag (side stream)
    int rank = 0;
    unsigned int* hbt_cnt = nullptr;          // MODE 2 / 6: counter of finished own tiles
    unsigned int* hbt_flag[XCHG_MAX_RANKS] = {};
    const unsigned int* hbt_wait = nullptr;   // MODE 0: this rank's PH_HBT flag row
    int defer_signal = 0, signal_hbt = 0;
};

template <int KP>
struct TcSolver {
    nmfb200_handle* h;
    cudaStream_t st;
    TcState* state;
    std::string pfx = "tc";          // prefix of the handle's named buffers (one set per logical rank)
    std::string gram_tag = "gram_part";  // name of the tile-Gram buffer the next launch fills
    const ShardLaunch* sl = nullptr; // row-sharded launch extras for the NEXT launch_update (reset by the caller)
    const bf16* Xs_lo = nullptr;     // precision mode bf16x3: the remainder panel matching the Xs of the NEXT launch_update
    bool force_x3 = false;           // split operands whatever the handle's precision option says (tc_xmul)
    float* cross_part = nullptr;     // verbose W-step: per-tile sums of Num .* F_new for the NEXT launch_update (trace identity)
    int gtl_slot = 0;                // diagnostics (tc_debug bit 7): which half of the per-CTA timeline buffer the NEXT launch_update fills
    bool chain = false;              // option tc_chain: update and reduce kernels of the loop form one chain of programmatic dependents
    bool skew = false;               // option tc_skew (needs chain): two groups of tiles half a period apart
    unsigned int skew_a = 0, skew_b = 0, skew_mid = 0;   //   cumulative tiles of group A / B and A CTAs of all skewed launches so far
    int chain_blocks = 148;          //   CTAs of a chained reduce kernel (they walk its virtual blocks; all resident early, none in the way)
    unsigned int chain_tiles = 0;    //   tiles of all chained update launches so far in this solve (what the next one waits for)
    const bf16* pf_X = nullptr;      // option tc_prefetch_next: the X panel of the launch AFTER the next launch_update ...
    int pf_tiles = 0, pf_tile_rows = 0, pf_nkb = 0;   // ... and its geometry (reset by the caller)
    bool defer_gram_reduce = false;  // the caller will run gram_conv_reduce_kernel itself
    bool last_fused_gram = false;
    bool last_chained = false;       // the last launch_update was part of the chain (its reduce kernel must publish)
    float* last_gram_part = nullptr;
    int last_gram_parts = 0;
    int num_splits = 1;              // MODE 5: k-split partial numerators behind num_io
    int64_t num_split_stride = 0;

    // gram: -1 = no Gram of the updated factor wanted; 0 / 1 = wanted, without / with the bf16 hi-lo split;
    // gram_dst = where the fp32 Gram goes (default F.P).  KP <= 128: the update kernel's staged epilogue produces the
    // per-tile contributions itself (only a reduce launch follows); KP = 256: separate gram_kernel pass.
    void launch_update(int mode, const Factor& F, const Factor& O, const bf16* Xs, int Kdim, float lambda, float delta,
                       float* num_io, float* conv_override = nullptr, int gram = -1, float* gram_dst = nullptr, bool pdl = false) {
        UpdateParams prm;
        // precision mode bf16x3 (MultUpdate :mse and the GreedyCD gradient): three passes over the k-blocks with the bf16
        // remainders of X and of the other factor; the Gram then comes from the split stand-alone kernel, not the tile epilogue
        const bool x3 = (h->tc_precision == 1 || force_x3) && Xs_lo != nullptr && O.bTlo != nullptr && (mode == 0 || mode == 3 || (mode == 1 && sl == nullptr));
        const bool fused_gram = gram >= 0 && KP <= 128 && (mode == 0 || mode == 2 || mode == 6) && !x3;
        std::memset(&prm, 0, sizeof(prm));
        prm.gram_part = fused_gram ? h->buf_t<float>(pfx + "." + gram_tag, (size_t)std::max(F.tiles, 1) * KP * KP) : nullptr;
        prm.tmF32 = make_tmap_f32(F.m, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmT = make_tmap_bf16(F.bT, (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, KP);
        prm.tile_rows = F.tile_rows;
        prm.timing = (h->tc_debug & 8) ? (long long*)h->buf("tc.timing", (16 + 2 * 4096) * sizeof(long long)) : nullptr;
        if ((h->tc_debug & 64) && mode == 0) prm.timing = nullptr;   // bit 6: keep the H-step's clocks (the W-step would overwrite them)
        const uint64_t nkb = (uint64_t)ceil_div(Kdim, 64);
        const int tile0 = sl ? sl->tile0 : 0;
        if ((h->tc_debug & 128) && mode == 0 && F.tiles <= 4096)
            prm.gtl = (long long*)h->buf("tc.gtl", 2 * 8 * 4096 * sizeof(long long)) + (size_t)gtl_slot * 8 * 4096;
        prm.timing_cta = mode == 6 ? F.tiles - 1 : 0;   // fused sharded H-step: the last CTA works on a tile this rank owns
        prm.tmA = make_tmap_bf16(Xs, 64, (uint64_t)(tile0 + F.tiles) * nkb * F.tile_rows, 64, (uint32_t)F.tile_rows);
        if (sl) {
            prm.tile0 = sl->tile0; prm.wait_first = sl->wait_first; prm.num_row0 = sl->num_row0;
            prm.G = sl->G; prm.tiles_per_owner = sl->tiles_per_owner; prm.tiles_total = sl->tiles_total; prm.epoch = sl->epoch;
            for (int j = 0; j < XCHG_MAX_RANKS; ++j) { prm.num_peer[j] = sl->num_peer[j]; prm.num_flag[j] = sl->num_flag[j]; }
            prm.own_cnt = sl->own_cnt;
            prm.n_peer = sl->n_peer;
            for (int j = 0; j < sl->n_peer; ++j)
                prm.tmT_peer[j] = make_tmap_bf16(sl->peer_bT[j], (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, KP);
            if (mode == 1 || mode == 6)
                for (int o = 0; o < sl->G; ++o)
                    prm.tmNum[o] = make_tmap_f32(sl->num_peer[o], KP, (uint64_t)sl->slot_rows, KP, (uint32_t)F.tile_rows);
            prm.num_wait = sl->num_wait;
            prm.den_flag = sl->den_flag;
            prm.rank = sl->rank;
            prm.hbt_cnt = sl->hbt_cnt;
            for (int j = 0; j < XCHG_MAX_RANKS; ++j) prm.hbt_flag[j] = sl->hbt_flag[j];
            prm.hbt_wait = sl->hbt_wait;
            prm.defer_signal = sl->defer_signal;
            prm.signal_hbt = sl->signal_hbt;
        }
        prm.tmB = make_tmap_bf16(O.bT, (uint64_t)Kdim, KP, (uint64_t)O.ldT, KP);
        if (KP <= 128 && (mode == 0 || mode == 1 || mode == 3 || mode == 6))
            prm.flush_chunk = h->tc_flush >= 0 ? h->tc_flush : 8;   // chunked accumulation (costs nothing measurable: 4784 vs 4781 it/s at config 2)
        if (x3) {
            prm.x3 = 1;
            prm.tmAlo = make_tmap_bf16(Xs_lo, 64, (uint64_t)(tile0 + F.tiles) * nkb * F.tile_rows, 64, (uint32_t)F.tile_rows);
            prm.tmBlo = make_tmap_bf16(O.bTlo, (uint64_t)Kdim, KP, (uint64_t)O.ldT, KP);
        }
        if (mode == 0 && !x3 && sl == nullptr && pf_X != nullptr && h->tc_prefetch_next > 0 && pf_tiles > 0) {
            prm.tmAnext = make_tmap_bf16(pf_X, 64, (uint64_t)pf_tiles * pf_nkb * pf_tile_rows, 64, (uint32_t)pf_tile_rows);
            prm.pf_blocks = std::min(h->tc_prefetch_next, pf_nkb);
            prm.pf_tiles = pf_tiles;
            prm.pf_tile_rows = pf_tile_rows;
            prm.pf_panel_rows = pf_nkb * pf_tile_rows;
        }
        const bool timed_ = h->time_kernels == 1 && mode != 2 && mode != 6;
        const bool chained = chain && pdl && mode == 0 && fused_gram && sl == nullptr && !timed_;
        const bool skewed = chained && skew && F.tile_rows == 128 && O.tile_rows == 128 && F.tiles >= 2 && O.tiles >= 2 &&
                            F.tiles <= h->sm_count && O.tiles <= h->sm_count;
        if (skewed) {
            const int split = F.tiles / 2;
            prm.skew_cnt = state->skew;
            prm.tile_split = split;
            prm.kb_split = 2 * (O.tiles / 2);   // a 128-row tile of the other factor = two 64-wide k-blocks
            prm.need_a = skew_a;
            prm.need_b = skew_b;
            prm.need_mid = skew_mid + (unsigned int)split;
            prm.early_trigger = 2;
            prm.diag_nob = (h->tc_debug & 256) ? 1 : 0;
            skew_a += (unsigned int)split;
            skew_b += (unsigned int)(F.tiles - split);
            skew_mid += (unsigned int)split;
        } else if (chained) {
            prm.chain_flag = &state->chain;
            prm.chain_cnt = &state->chain;
            prm.chain_need = chain_tiles;
            prm.early_trigger = 1;
            chain_tiles += (unsigned int)F.tiles;
        }
        prm.tmFhi = make_tmap_bf16(F.hi, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmFlo = make_tmap_bf16(F.lo, KP, (uint64_t)F.R, KP, (uint32_t)F.tile_rows);
        prm.tmPhi = make_tmap_bf16(O.Phi, KP, KP, KP, KP);
        prm.tmPlo = make_tmap_bf16(O.Plo, KP, KP, KP, KP);
        prm.F = F.m; prm.Fhi = F.hi; prm.Flo = F.lo; prm.FbT = F.bT; prm.ldT = F.ldT;
        prm.num_io = num_io;
        prm.cross_part = mode == 0 ? cross_part : nullptr;
        prm.num_splits = num_splits;
        prm.num_split_stride = num_split_stride;
        prm.conv_part = conv_override ? conv_override : F.conv;
        prm.Pfull = O.P;
        prm.colsum = O.colsum;
        prm.state = state;
        prm.R = F.R; prm.Kdim = Kdim; prm.lambda = lambda; prm.delta = delta;
        const int smem = UpdCfg<KP>::SMEM_BYTES;
        const bool timed = h->time_kernels == 1 && mode != 2 && mode != 6;
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        void (*kern)(const UpdateParams) = mode == 0   ? mu_update_kernel<KP, 0>
                                           : mode == 1 ? mu_update_kernel<KP, 1>
                                           : mode == 2 ? mu_update_kernel<KP, 2>
                                           : mode == 3 ? mu_update_kernel<KP, 3>
                                           : mode == 4 ? mu_update_kernel<KP, 4>
                                           : mode == 5 ? mu_update_kernel<KP, 5>
                                                       : mu_update_kernel<KP, 6>;
        // pdl: programmatic dependent launch -- start streaming X while the preceding reduce kernel is still running
        launch_k(kern, dim3((unsigned)F.tiles), dim3(UpdCfg<KP>::THREADS), (size_t)smem, st, pdl, prm);
        if (timed) NMF_CUDA(cudaEventRecord(h->next_event(), st));
        h->launches += 1;
        if (x3 && mode == 0) refresh_bTlo(F);   // the new remainder, transposed, for the other factor's next hi*lo pass
        last_fused_gram = fused_gram;
        last_gram_part = prm.gram_part;
        last_gram_parts = F.tiles;
        last_chained = chained;
        if (fused_gram && !defer_gram_reduce) {
            const int nvb = (4 * KP * KP + 255) / 256;
            launch_k(gram_reduce_kernel, dim3(chained ? std::min(nvb, chain_blocks) : nvb), dim3(256), 0, st, chained, (const float*)prm.gram_part,
                     F.tiles, KP * KP, gram_dst ? gram_dst : F.P, F.Phi, F.Plo, gram, (const TcState*)state, chained ? 1 : 0, nvb);
            h->launches += 1;
        } else if (gram >= 0 && !fused_gram) {
            launch_gram(F, gram != 0, gram_dst);
        }
    }

    void refresh_bTlo(const Factor& F) {
        transpose_lo_kernel<<<dim3((unsigned)ceil_div(F.R, 32), KP / 32), dim3(32, 8), 0, st>>>(F.lo, F.R, KP, F.bTlo, F.ldT);
        h->launches += 1;
    }
    // partial Grams of rows [k0, k1) of F (k1 < 0: all rows) -> last_gram_part / last_gram_parts; no reduce
    void launch_gram_parts(const Factor& F, int k0 = 0, int k1 = -1) {
        GramParams g;
        if (k1 < 0) k1 = F.R;
        std::memset(&g, 0, sizeof(g));
        g.tmT = make_tmap_bf16(F.bT, (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, 128);
        if (h->tc_precision == 1 && F.bTlo != nullptr) {   // T T' = hi hi' + hi lo' + lo hi'
            g.split = 1;
            g.tmTlo = make_tmap_bf16(F.bTlo, (uint64_t)F.R, (uint64_t)F.rowsT, (uint64_t)F.ldT, 128);
        }
        // ~128 CTAs at most, each a multiple of 64 rows and at least 256
        g.chunk = (int)std::max<int64_t>(256, round_up(ceil_div(std::max(k1 - k0, 1), 128), 64));
        const int grid = (int)std::max<int64_t>(1, ceil_div(k1 - k0, g.chunk));
        g.part = h->buf_t<float>(pfx + "." + gram_tag, (size_t)grid * KP * KP);
        g.state = state;
        g.R = F.R;
        g.k0 = k0;
        g.k1 = k1;
        gram_kernel<KP><<<grid, 192, GramCfg<KP>::SMEM_BYTES, st>>>(g);
        h->launches += 1;
        last_gram_part = g.part;
        last_gram_parts = grid;
    }
    void launch_gram(const Factor& F, bool split, float* P_dst = nullptr) {
        launch_gram_parts(F);
        gram_reduce_kernel<<<(4 * KP * KP + 255) / 256, 256, 0, st>>>(last_gram_part, last_gram_parts, KP * KP, P_dst ? P_dst : F.P, F.Phi, F.Plo,
                                                                       split ? 1 : 0, state, 0, (4 * KP * KP + 255) / 256);
        h->launches += 1;
    }

    // cudaFuncSetAttribute is per device: remember it per device ordinal (a process may hold handles on several GPUs)
    static void set_attrs(int device) {
        static bool done_on[64] = {};
        bool& done = done_on[device & 63];
        if (done) return;
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(mu_update_kernel<KP, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, UpdCfg<KP>::SMEM_BYTES));
        NMF_CUDA(cudaFuncSetAttribute(gram_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg<KP>::SMEM_BYTES));
        done = true;
    }
};

#include "tc_objective.cuh"
#include "tc_shard.cuh"

template <int KP>
void tc_solve_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs(h->device);
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;

    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));

    bf16 *Xr = nullptr, *Xc = nullptr, *Xr_lo = nullptr, *Xc_lo = nullptr;
    build_x_caches(h, &Xr, &Xc, &Xr_lo, &Xc_lo);
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
    s.chain = h->tc_chain != 0 && !a.verbose;
    s.chain_blocks = h->tc_chain > 1 ? h->tc_chain : h->sm_count;   // one CTA per SM unless told otherwise
    s.skew = s.chain && h->tc_skew != 0 && a.update_H;
    if (h->tc_precision == 1) { s.refresh_bTlo(W); s.refresh_bTlo(H); }
    s.launch_gram(W, true);                   // P_W = W'W for the first H-step
    if (!a.update_H) s.launch_gram(H, true);  // H never changes: P_H once
    NMF_CUDA(cudaEventRecord(e1, st));

    h->ev_used = 0;
    int64_t enq = 0;
    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    // Both update kernels are launched as programmatic dependents of the small reduce kernel in front of them, so their X
    // streaming overlaps that kernel and the launch gap (the reduce results are only needed by the denominator blocks and
    // the epilogue, which wait for it).
    const bool pdl = h->tc_pdl != 0;
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
    // verbose: the objective after every iteration comes from the trace identity 0.5*(||X||^2 - 2<XH',W> + <W'W,HH'>) -- the W-step's
    // numerators times the new W, and the two Grams, are on hand -- instead of a pass over X (311 us at config 2 = 1.5 iterations);
    // the line before the loop and the last one (= Result.objvalue) use the objective kernel
    const bool trace_id = a.verbose && h->tc_trace_identity != 0;
    TraceObj tr;
    std::memset(&tr, 0, sizeof(tr));
    if (trace_id) {
        tr.cross_part = h->buf_t<float>("tc.cross_part", (size_t)std::max(W.tiles, 1));
        tr.ntiles = W.tiles;
        tr.P_other = H.P;
        tr.xnorm2 = x_norm2(h);
    }
    if (a.verbose) {
        v_t0 = wall();
        v_objv = objective_now();
        if (h->trace) h->trace(h->trace_user, 0, 0.0, v_objv, NAN, NAN);
    }
    while (enq < a.maxiter) {
        int64_t batch = a.verbose ? 1 : std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            h->mark("start");
            if (a.update_H) {
                s.Xs_lo = Xr_lo;
                s.pf_X = Xc; s.pf_tiles = W.tiles; s.pf_tile_rows = W.tile_rows; s.pf_nkb = (int)ceil_div(n, 64);   // next: the W-step's panel
                s.gtl_slot = 0;
                s.launch_update(0, H, W, Xr, (int)p, lh, delta, nullptr, nullptr, 1, nullptr, pdl);  // H-step (+ tile Grams of the new H)
                h->mark("updH");
            }
            // W-step + W'W for the next H-step (not needed if H is fixed)
            const int gramW = a.update_H ? 1 : -1;
            s.defer_gram_reduce = true;
            s.Xs_lo = Xc_lo;
            s.cross_part = const_cast<float*>(tr.cross_part);
            s.gtl_slot = 1;
            if (a.update_H) { s.pf_X = Xr; s.pf_tiles = H.tiles; s.pf_tile_rows = H.tile_rows; s.pf_nkb = (int)ceil_div(p, 64); }   // next: an H-step
            else { s.pf_X = Xc; s.pf_tiles = W.tiles; s.pf_tile_rows = W.tile_rows; s.pf_nkb = (int)ceil_div(n, 64); }
            s.launch_update(0, W, H, Xc, (int)n, lw, delta, nullptr, nullptr, gramW, nullptr, pdl);
            s.pf_X = nullptr;
            s.cross_part = nullptr;
            s.defer_gram_reduce = false;
            h->mark("updW");
            const int gram_blocks = (s.last_fused_gram && gramW >= 0) ? (4 * KP * KP + 255) / 256 : 0;
            // one launch: Gram reduce (if the staged epilogue produced tile Grams; KP = 256 ran gram_kernel + reduce inside
            // launch_update) + stop_condition reduce / decision
            // option tc_chain: a W-step launched with early_trigger lets this kernel take its seats early; it then waits for the W-step
            const bool chW = s.last_chained;
            const int nvbW = gram_blocks + 4 * (KP / 32);
            launch_k(gram_conv_reduce_kernel, dim3(chW ? std::min(nvbW, s.chain_blocks) : nvbW), dim3(256), 0, st, chW, (const float*)s.last_gram_part,
                     W.tiles, KP * KP, W.P, W.Phi, W.Plo, 1, gram_blocks, (const float*)W.conv, W.tiles, (const float*)H.conv, H.tiles,
                     KP, (int)k, (int)a.update_H, acc, tol, state, 1, (float*)nullptr, tr, chW ? 1 : 0, nvbW);
            h->launches += 1;
            h->mark("conv");
        }
        enq += batch;
        NMF_CUDA(cudaGetLastError());
        NMF_CUDA(cudaMemcpyAsync(&hs, state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        iters = hs.iters;
        devmax = hs.devmax;
        if (a.verbose) {
            const double pre = v_objv;
            const bool last = hs.converged != 0 || enq >= a.maxiter;
            v_objv = (trace_id && !last && a.update_H) ? hs.objv : objective_now();
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
    if (h->tc_debug & 128) {  // per-CTA life lines of the last H-step and the last W-step launch (ns since the H-step's first entry)
        std::vector<long long> g(2 * 8 * 4096);
        NMF_CUDA(cudaMemcpy(g.data(), h->buf("tc.gtl", g.size() * sizeof(long long)), g.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        long long t0 = LLONG_MAX;
        for (int c = 0; c < std::min(H.tiles, 4096); ++c) t0 = std::min(t0, g[8 * c]);
        for (int slot = 0; slot < 2; ++slot) {
            const int nct = std::min(slot == 0 ? H.tiles : W.tiles, 4096);
            for (int c = 0; c < nct; ++c) {
                const long long* t = g.data() + (size_t)slot * 8 * 4096 + 8 * c;
                fprintf(stderr, "[gtl] %c %d %lld %lld %lld %lld %lld %lld %lld %lld\n", slot == 0 ? 'H' : 'W', c, t[0] - t0, t[1] - t0, t[2] - t0, t[3] - t0,
                        t[4] - t0, t[5] - t0, t[6] - t0, t[7] - t0);
            }
        }
    }
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

// ---- batched replicates (interf.jl:85-101) --------------------------------------------------------------
// `nrep` independent MultUpdate(:mse) solves of the same X advance TOGETHER: their factors are stacked along the component axis
// (W p x nrep*k, H nrep*k x n), so every pass over X feeds nrep numerators at once -- the contraction X'W_stack is the same kernel with
// a wider B operand (nrep x the arithmetic intensity of one solve; X is the HBM-bound operand).  The replicates do not interact because
// the k x k Grams of the denominators are kept BLOCK DIAGONAL (gram_masked, tc_update.cuh): Den_r = H_r (W_r'W_r).  stop_condition is
// evaluated per replicate (conv_decide_batched); a replicate that meets it is snapshotted at that iteration (batch_snapshot_kernel) while
// the stacked iteration carries on for the others, and the loop ends when all have passed or at maxiter -- each replicate returns the
// niters / converged / factors its own solve! would have produced.
template <int KPr>
double batched_objective(nmfb200_handle* h, const float* Wd, int64_t ldwd, const float* Hd, int64_t ldhd, int64_t k) {
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n;
    Factor Wr = alloc_factor(h, "Wrep", (int)p, KPr), Hr = alloc_factor(h, "Hrep", (int)n, KPr);
    pack_factor_kernel<<<ew_grid(p * KPr), 256, 0, st>>>(Wd, 1, ldwd, (int)p, (int)k, KPr, Wr.m, Wr.hi, Wr.lo, Wr.bT, Wr.ldT);
    pack_factor_kernel<<<ew_grid(n * KPr), 256, 0, st>>>(Hd, ldhd, 1, (int)n, (int)k, KPr, Hr.m, Hr.hi, Hr.lo, Hr.bT, Hr.ldT);
    h->launches += 2;
    double v = 0;
    if (!tc_objective<KPr>(h, 0, Wr, Hr, 0.0, 0.0, &v)) v = simt_objective_f32(h, 0, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);
    return v;
}

template <int KP>
void tc_solve_batched_kp(nmfb200_handle* h, const SolveArgs& a, int nrep, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs(h->device);
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k, kt = a.k * nrep;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;
    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    bf16 *Xr = nullptr, *Xc = nullptr, *Xr_lo = nullptr, *Xc_lo = nullptr;
    build_x_caches(h, &Xr, &Xc, &Xr_lo, &Xc_lo);
    NMF_CUDA(cudaEventRecord(e0, st));

    Factor W = alloc_factor(h, "W", (int)p, KP), H = alloc_factor(h, "H", (int)n, KP);
    TcState* state = (TcState*)h->buf("tc.state", sizeof(TcState));
    BatchState* bs = (BatchState*)h->buf("tc.batch", sizeof(BatchState));
    float* Wsnap = h->buf_t<float>("tc.W.snap", (size_t)p * KP);
    float* Hsnap = h->buf_t<float>("tc.H.snap", (size_t)n * KP);
    double* acc = h->buf_t<double>("tc.acc", 4 * KP);
    TcState hs0;
    std::memset(&hs0, 0, sizeof(hs0));
    hs0.blk = (int)k;
    hs0.kp = KP;
    hs0.nrep = nrep;
    hs0.bs = bs;
    NMF_CUDA(cudaMemcpyAsync(state, &hs0, sizeof(TcState), cudaMemcpyHostToDevice, st));
    NMF_CUDA(cudaMemsetAsync(bs, 0, sizeof(BatchState), st));
    NMF_CUDA(cudaMemsetAsync(W.bT, 0, (size_t)W.rowsT * W.ldT * sizeof(bf16), st));
    NMF_CUDA(cudaMemsetAsync(H.bT, 0, (size_t)H.rowsT * H.ldT * sizeof(bf16), st));

    float *Wd = Wc, *Hd = Hc;
    int64_t ldwd = ldw, ldhd = ldh;
    if (!a.on_device) {
        Wd = h->buf_t<float>("tc.Wstage", (size_t)p * kt);
        Hd = h->buf_t<float>("tc.Hstage", (size_t)kt * n);
        ldwd = p;
        ldhd = kt;
        NMF_CUDA(cudaMemcpy2DAsync(Wd, p * sizeof(float), Wc, ldw * sizeof(float), p * sizeof(float), kt, cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hd, kt * sizeof(float), Hc, ldh * sizeof(float), kt * sizeof(float), n, cudaMemcpyHostToDevice, st));
    }
    pack_factor_kernel<<<ew_grid(p * KP), 256, 0, st>>>(Wd, 1, ldwd, (int)p, (int)kt, KP, W.m, W.hi, W.lo, W.bT, W.ldT);
    pack_factor_kernel<<<ew_grid(n * KP), 256, 0, st>>>(Hd, ldhd, 1, (int)n, (int)kt, KP, H.m, H.hi, H.lo, H.bT, H.ldT);
    h->launches += 2;
    NMF_CUDA(cudaGetLastError());

    TcSolver<KP> s{h, st, state};
    if (h->tc_precision == 1) { s.refresh_bTlo(W); s.refresh_bTlo(H); }
    s.launch_gram(W, true);                   // block-diagonal W'W for the first H-step
    if (!a.update_H) s.launch_gram(H, true);
    NMF_CUDA(cudaEventRecord(e1, st));

    h->ev_used = 0;
    int64_t enq = 0;
    TcState hs;
    std::memset(&hs, 0, sizeof(hs));
    const bool pdl = h->tc_pdl != 0;
    const int snap_grid = ew_grid((p + n) * KP);
    while (enq < a.maxiter) {
        const int64_t batch = std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            if (a.update_H) {
                s.Xs_lo = Xr_lo;
                s.launch_update(0, H, W, Xr, (int)p, lh, delta, nullptr, nullptr, 1, nullptr, pdl);
            }
            const int gramW = a.update_H ? 1 : -1;
            s.defer_gram_reduce = true;
            s.Xs_lo = Xc_lo;
            s.launch_update(0, W, H, Xc, (int)n, lw, delta, nullptr, nullptr, gramW, nullptr, pdl);
            s.defer_gram_reduce = false;
            const int gram_blocks = (s.last_fused_gram && gramW >= 0) ? (4 * KP * KP + 255) / 256 : 0;
            TraceObj tr;
            std::memset(&tr, 0, sizeof(tr));
            launch_k(gram_conv_reduce_kernel, dim3(gram_blocks + 4 * (KP / 32)), dim3(256), 0, st, false, (const float*)s.last_gram_part,
                     W.tiles, KP * KP, W.P, W.Phi, W.Plo, 1, gram_blocks, (const float*)W.conv, W.tiles, (const float*)H.conv, H.tiles,
                     KP, (int)kt, (int)a.update_H, acc, tol, state, 1, (float*)nullptr, tr, 0, gram_blocks + 4 * (KP / 32));
            launch_k(batch_snapshot_kernel, dim3(snap_grid), dim3(256), 0, st, false, (const TcState*)state, KP, (const float*)W.m, Wsnap,
                     (int64_t)p * KP, (const float*)H.m, Hsnap, (int64_t)n * KP, 0);
            h->launches += 2;
        }
        enq += batch;
        NMF_CUDA(cudaGetLastError());
        NMF_CUDA(cudaMemcpyAsync(&hs, state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        if (hs.converged) break;
    }
    NMF_CUDA(cudaEventRecord(e2, st));
    // replicates that never met stop_condition: their factors are the current ones
    batch_snapshot_kernel<<<snap_grid, 256, 0, st>>>(state, KP, W.m, Wsnap, (int64_t)p * KP, H.m, Hsnap, (int64_t)n * KP, 1);
    unpack_factor_kernel<<<ew_grid(p * kt), 256, 0, st>>>(Wsnap, (int)p, (int)kt, KP, Wd, 1, ldwd);
    unpack_factor_kernel<<<ew_grid(n * kt), 256, 0, st>>>(Hsnap, (int)n, (int)kt, KP, Hd, ldhd, 1);
    h->launches += 3;
    NMF_CUDA(cudaGetLastError());
    BatchState hb;
    NMF_CUDA(cudaMemcpyAsync(&hb, bs, sizeof(BatchState), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    float ms_up = 0, ms_loop = 0;
    cudaEventElapsedTime(&ms_up, e0, e1);
    cudaEventElapsedTime(&ms_loop, e1, e2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    for (int r = 0; r < nrep; ++r) {   // objective of every replicate from ITS factors (multupd.jl:81)
        const float* Wr = Wd + (size_t)r * k * ldwd;
        const float* Hr = Hd + (size_t)r * k;
        double objv = 0;
        switch (pick_kp(k)) {
            case 64: objv = batched_objective<64>(h, Wr, ldwd, Hr, ldhd, k); break;
            case 128: objv = batched_objective<128>(h, Wr, ldwd, Hr, ldhd, k); break;
            default: objv = batched_objective<256>(h, Wr, ldwd, Hr, ldhd, k); break;
        }
        std::memset(&out[r], 0, sizeof(nmfb200_result));
        out[r].niters = hb.done[r] ? hb.niters[r] : hs.iters;
        out[r].converged = hb.done[r] ? 1 : 0;
        out[r].engine = 1;
        out[r].objvalue = objv;
        out[r].last_dev = hb.devmax[r];
        out[r].solve_ms = ms_loop;       // the loop is shared: device time of the whole batch
        out[r].upload_ms = ms_up;
    }
    if (!a.on_device) {
        NMF_CUDA(cudaMemcpy2DAsync(Wc, ldw * sizeof(float), Wd, p * sizeof(float), p * sizeof(float), kt, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hc, ldh * sizeof(float), Hd, kt * sizeof(float), kt * sizeof(float), n, cudaMemcpyDeviceToHost, st));
    }
    NMF_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < nrep; ++r) out[r].kernel_launches = h->launches;
    int64_t npairs = 0;
    const double hot = h->drain_event_pairs(&npairs);
    out[0].hot_kernel_ms = hot;
    out[0].hot_kernel_launches = npairs;
}

#include "tc_div.cuh"

// Row-sharded MultUpdate(:div) (multupd.jl:171-181): the k-split partial numerators W_g'Q_g of one rank summed in split order
// -> one [R][KP] buffer whose size does not depend on the rank's row count (it is all-reduced over the ranks next)
__global__ void sum_splits_kernel(const float4* __restrict__ part, int nsplits, int64_t stride4, int64_t n4, float4* __restrict__ out,
                                  const TcState* st) {
    if (st->converged) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldcg(part + i);
        for (int sp = 1; sp < nsplits; ++sp) {
            const float4 v = __ldcg(part + (size_t)sp * stride4 + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        out[i] = a;
    }
}
// stop_condition decision on sums that were all-reduced over the ranks (common.jl:105-106)
__global__ void __launch_bounds__(256) conv_decide_kernel(const double* __restrict__ acc, int KP, int k, float tol, TcState* st) {
    __shared__ float devs[256];
    __shared__ int fail;
    if (st->converged) return;
    conv_decide(acc, KP, k, tol, st, devs, &fail);
}

template <int KP>
void tc_solve_div_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs(h->device);
    static bool quot_attr_on[64] = {};
    bool& quot_attr = quot_attr_on[h->device & 63];
    if (!quot_attr) {
        NMF_CUDA(cudaFuncSetAttribute(div_quot_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, QuotCfg<KP>::SMEM_BYTES));
        quot_attr = true;
    }
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = std::max((float)a.lambda_w, delta), lh = std::max((float)a.lambda_h, delta);  // multupd.jl:37-40
    const float tol = (float)a.tol;
    const bool multi = h->comm != nullptr;   // rows of X / W sharded over ranks, H replicated: two NCCL all-reduces per H-step
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
        static bool fattr_on[64] = {};
        bool& fattr = fattr_on[h->device & 63];
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
    // reduce_ranks (row-sharded H-step): Cf = this rank's rows of W, so the column sums sW and the numerators W'Q are partial -- both are
    // all-reduced over the ranks (NCCL on the solver's stream) before the ratio, which every rank then applies to its replica of H
    auto half_step = [&](Factor& Rf, Factor& Cf, const bf16* Xs, int nkb, int Kdim, float lambda, bool reduce_ranks) {
        const uint64_t prow = (uint64_t)Rf.tiles * nkb * 128;
        const bool pdl = fused && h->tc_pdl != 0;
        colsum_tiles_kernel<<<Cf.tiles, 256, 0, st>>>(Cf.m, Cf.R, KP, cs_part, state);
        launch_k(colsum_reduce_kernel, dim3(KP / 32), dim3(256), 0, st, pdl, (const float*)cs_part, Cf.tiles, KP, Cf.colsum, (const TcState*)state);
        h->launches += 2;
        if (reduce_ranks) h->allreduce_sum(Cf.colsum, (size_t)KP);
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
            launch_k(div_fused_kernel<KP>, dim3(Rf.tiles, ksplit), dim3(DivFusedCfg<KP>::THREADS), (size_t)DivFusedCfg<KP>::SMEM_BYTES, st, pdl, fp);
            h->launches += 1;
            if (reduce_ranks) {
                // the number of k-splits follows the rank's own row count: sum them here, all-reduce a buffer every rank sizes alike
                float* red = h->buf_t<float>("tc.div_num_red", (size_t)Rf.R * KP);
                const int64_t n4 = (int64_t)Rf.R * KP / 4;
                sum_splits_kernel<<<ew_grid(n4), 256, 0, st>>>((const float4*)fp.num_part, ksplit, n4, n4, (float4*)red, state);
                h->launches += 1;
                h->allreduce_sum(red, (size_t)Rf.R * KP);
                s.num_splits = 1;
                s.num_split_stride = 0;
                s.launch_update(5, Rf, Cf, Xs, Kdim, lambda, delta, red);
                return;
            }
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
            if (a.update_H) half_step(H, W, Xr, nkbH, (int)p, lh, multi);   // multupd.jl:171-181
            half_step(W, H, Xc, nkbW, (int)n, lw, false);                    // :183-192 (local: H and its column sums are replicated)
            if (multi) {   // the W-side sums of stop_condition run over this rank's rows only
                conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 0, nullptr);
                h->allreduce_sum(acc, (size_t)2 * KP);
                conv_decide_kernel<<<1, 256, 0, st>>>(acc, KP, (int)k, tol, state);
                h->launches += 1;
            } else {
                conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, W.tiles, H.conv, H.tiles, KP, (int)k, a.update_H, acc, tol, state, 1, nullptr);
            }
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
    double objv = 0;  // gkldiv (multupd.jl:148); row-sharded: the data term is a sum over the ranks' rows
    if (h->all_ranks(tc_objective_covers((const float*)h->dX, h->p, h->n, h->ldx))) {
        double* res = tc_objective_enqueue<KP>(h, "tc", 1, (const float*)h->dX, h->p, h->n, h->ldx, W, H, 0.0, 0.0);
        h->allreduce_sum(res, 2);
        double hres[3] = {0, 0, 0};
        NMF_CUDA(cudaMemcpyAsync(hres, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        objv = tc_objective_value(1, hres, 0.0, 0.0);
    } else {
        objv = simt_objective_f32(h, 1, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);
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
    TcSolver<KP>::set_attrs(h->device);
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;
    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    bf16 *Xr = nullptr, *Xc = nullptr, *Xr_lo = nullptr, *Xc_lo = nullptr;
    build_x_caches(h, &Xr, &Xc, &Xr_lo, &Xc_lo);
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
    NMF_CUDA(cudaMemsetAsync(d_updates, 0, 2 * sizeof(unsigned long long), st));
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
    if (h->tc_precision == 1) { s.refresh_bTlo(W); s.refresh_bTlo(H); }
    s.launch_gram(H, true);  // P = HH' for the first W-step (greedycd.jl:117)
    NMF_CUDA(cudaEventRecord(e1, st));

    // Row-sharded (rows of X and W over the ranks, H replicated):
    //  * W-step: everything is local except p_init, a maximum over ALL rows of W (greedycd.jl:132-137): 1-float all-reduce (max).
    //  * H-step: the gradient G = H (W'W) - X'W is a sum over the ranks of  H (W_g'W_g) - X_g'W_g  -- exactly what the update kernel
    //    produces from this rank's rows and this rank's OWN Gram (whose bf16 split it already holds) -- so G is all-reduced once, lambda is
    //    added on rank 0 only, p_init is recomputed from the complete G (the fused per-CTA maximum saw a partial one), and the coordinate
    //    loop runs replicated on every rank with the all-reduced fp32 Gram.
    const bool multi = h->comm != nullptr;
    float* Pw_glob = multi ? h->buf_t<float>("tc.gcd_Pw_glob", (size_t)KP * KP) : nullptr;
    const int nb_h = (int)ceil_div(n, GCD_WARPS);
    float* bmax_h = multi ? h->buf_t<float>("tc.gcd_bmax_h", (size_t)nb_h + 1) : nullptr;
    auto half_step = [&](Factor& F, Factor& O, const bf16* Xs, const bf16* Xs_lo, int Kdim, float lambda, int tiles128, bool is_h) {
        NMF_CUDA(cudaMemcpyAsync(prev, F.m, (size_t)F.R * KP * sizeof(float), cudaMemcpyDeviceToDevice, st));
        s.Xs_lo = Xs_lo;
        const bool reduce_g = multi && is_h;
        s.launch_update(3, F, O, Xs, Kdim, (reduce_g && h->rank != 0) ? 0.f : lambda, 0.f, G, bmax);   // G = F P - X O (+lambda), per-CTA max D
        const float* P_rows = O.P;
        const float* p_init = bmax + maxtiles;
        if (reduce_g) {
            h->allreduce_sum(G, (size_t)F.R * KP);
            gcd_rowmax_kernel<float><<<nb_h, GCD_WARPS * 32, 0, st>>>(F.m, KP, 1, G, Pw_glob, F.R, KP, bmax_h);
            max_partials_kernel<float><<<1, 256, 0, st>>>(bmax_h, nb_h, bmax_h + nb_h);
            h->launches += 1;
            P_rows = Pw_glob;
            p_init = bmax_h + nb_h;
        } else {
            max_partials_kernel<float><<<1, 256, 0, st>>>(bmax, F.tiles, bmax + maxtiles);        // p_init (:132-137)
            if (multi) h->allreduce_max(bmax + maxtiles, 1);                                      // ... over the rows of every rank
        }
        gcd_rows_tc_kernel<KP><<<(unsigned)ceil_div(F.R, 8), 256, 0, st>>>(F.m, G, P_rows, F.R, (int)k, p_init, d_updates + (is_h ? 1 : 0));  // :139-165
        gcd_repack_kernel<<<tiles128, 256, 0, st>>>(F.m, prev, F.R, KP, F.hi, F.lo, F.bT, F.ldT, F.conv);
        h->launches += 3;
        if (h->tc_precision == 1) s.refresh_bTlo(F);
        s.launch_gram(F, true);                                                                   // Gram of the updated factor
        if (multi && !is_h) {   // W'W over all ranks for the H-step's coordinate loop; F.Phi / F.Plo keep this rank's own Gram
            NMF_CUDA(cudaMemcpyAsync(Pw_glob, F.P, (size_t)KP * KP * sizeof(float), cudaMemcpyDeviceToDevice, st));
            h->allreduce_sum(Pw_glob, (size_t)KP * KP);
        }
    };

    bool converged = false;
    int64_t iters = 0;
    float devmax = 0.f;
    TcState hs;
    while (iters < a.maxiter && !converged) {  // one host check per iteration: the row kernel has no early-exit flag
        half_step(W, H, Xc, Xc_lo, (int)n, lw, tilesW, false);                                     // W first (greedycd.jl:169-171)
        if (a.update_H) half_step(H, W, Xr, Xr_lo, (int)p, lh, tilesH, true);                      // then H (:173-177)
        if (multi) {   // the W-side sums of stop_condition run over this rank's rows only
            conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, tilesW, H.conv, tilesH, KP, (int)k, a.update_H, acc, tol, state, 0, nullptr);
            h->allreduce_sum(acc, (size_t)2 * KP);
            conv_decide_kernel<<<1, 256, 0, st>>>(acc, KP, (int)k, tol, state);
            h->launches += 1;
        } else {
            conv_reduce_kernel<<<4 * (KP / 32), 256, 0, st>>>(W.conv, tilesW, H.conv, tilesH, KP, (int)k, a.update_H, acc, tol, state, 1, nullptr);
        }
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
    double objv = 0;  // greedycd.jl:82-92; row-sharded: the data term and |W|_1 are sums over the ranks' rows, |H|_1 is replicated
    if (h->all_ranks(tc_objective_covers((const float*)h->dX, h->p, h->n, h->ldx))) {
        double* res = tc_objective_enqueue<KP>(h, "tc", 2, (const float*)h->dX, h->p, h->n, h->ldx, W, H, a.lambda_w, a.lambda_h);
        h->allreduce_sum(res, 2);
        double hres[3] = {0, 0, 0};
        NMF_CUDA(cudaMemcpyAsync(hres, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        objv = tc_objective_value(2, hres, a.lambda_w, a.lambda_h);
    } else {
        objv = simt_objective_f32(h, 2, Wd, ldwd, Hd, ldhd, k, a.lambda_w, a.lambda_h);
    }
    unsigned long long upd2[2] = {0, 0};   // coordinate steps on W rows (this rank's) and on H rows (replicated)
    NMF_CUDA(cudaMemcpyAsync(upd2, d_updates, sizeof(upd2), cudaMemcpyDeviceToHost, st));
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
    double w_upd = (double)upd2[0];
    if (multi) {   // steps on W summed over the ranks; the steps on the replicated H count once
        double* d = h->buf_t<double>("tc.gcd_upd_sum", 1);
        NMF_CUDA(cudaMemcpyAsync(d, &w_upd, sizeof(double), cudaMemcpyHostToDevice, st));
        h->allreduce_sum(d, 1);
        NMF_CUDA(cudaMemcpyAsync(&w_upd, d, sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
    }
    out->coordinate_updates = (int64_t)(w_upd + 0.5) + (int64_t)upd2[1];
    out->kernel_launches = h->launches;
    out->hot_kernel_ms = h->drain_event_pairs(&out->hot_kernel_launches);
}

}  // namespace

bool tc_supported(const nmfb200_handle* h, const SolveArgs& a) {
    if (a.alg > 2) return false;          // ProjectedALS / CoordinateDescent / ALSPGrad: exact engine
    // engine = auto: bf16 operands are worth it -- and an accepted trade -- only for problems large enough to be limited by
    // streaming X; below 2^20 cells every Float32 problem stays on the exact engine (fp32 parity with the reference, and
    // the reference's own tiny test problems -- laurberg6x3, 5x8 -- never see bf16 rounding).  engine = tc overrides.
    if (a.alg == 0 && h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 20)) return false;
    if (a.alg == 1) {                     // MultUpdate(:div): quotient kernel + update kernel; k <= 128, single GPU
        if (h->emulate_shards > 1 || a.k > 128 || h->p < 128 || h->n < 128) return false;
        if (h->comm != nullptr && h->tc_div_fused == 0) return false;   // row-sharded: the fused form only (its numerators are all-reduced)
        if (h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 20)) return false;
    }
    const bool sharded = h->comm != nullptr || h->emulate_shards > 1;
    if (sharded && h->tc_precision == 1 && h->engine_opt != 2) return false;   // parity mode on several ranks: the exact engine
    if (a.alg == 0 && h->comm != nullptr && !h->tc_xchg) return false;   // option tc_xchg=nccl: multi-GPU solves stay on the exact engine
    if (a.alg == 0 && sharded && (h->comm ? h->nranks : h->emulate_shards) > XCHG_MAX_RANKS) return false;
    if (a.alg == 2) {                     // GreedyCD: bf16 gradients; auto-selected for large problems only; several ranks: NCCL all-reduces
        if (h->emulate_shards > 1) return false;
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
    if (h->comm != nullptr || h->emulate_shards > 1) {   // rows of X / W sharded over ranks (real or logical): tc_shard.cuh
        switch (pick_kp(a.k)) {
            case 64: tc_solve_sharded_kp<64>(h, a, W, ldw, H, ldh, out); return;
            case 128: tc_solve_sharded_kp<128>(h, a, W, ldw, H, ldh, out); return;
            case 256: tc_solve_sharded_kp<256>(h, a, W, ldw, H, ldh, out); return;
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

bool tc_batched_supported(const nmfb200_handle* h, const SolveArgs& a, int nrep) {
    if (a.alg != 0 || a.verbose || nrep < 1 || nrep > MAX_BATCH || a.k * nrep > 256) return false;
    if (h->x_elt != 4 || h->engine_opt == 1 || h->comm != nullptr || h->emulate_shards > 1) return false;
    if (h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 20)) return false;   // small problems stay on the exact engine, one by one
    if (h->p > (int64_t)INT32_MAX / 256 || h->n > (int64_t)INT32_MAX / 256) return false;
    return true;
}

void tc_solve_batched(nmfb200_handle* h, const SolveArgs& a, int nrep, float* W, int64_t ldw, float* H, int64_t ldh, nmfb200_result* out) {
    switch (pick_kp(a.k * nrep)) {
        case 64: tc_solve_batched_kp<64>(h, a, nrep, W, ldw, H, ldh, out); break;
        case 128: tc_solve_batched_kp<128>(h, a, nrep, W, ldw, H, ldh, out); break;
        case 256: tc_solve_batched_kp<256>(h, a, nrep, W, ldw, H, ldh, out); break;
        default: throw Error{NMFB200_ENOTSUP, "replicates * k > 256 is not covered by the batched tensor-core solve"};
    }
}

namespace {
template <int KP>
void tc_xmul_kp(nmfb200_handle* h, int side, const float* O, int64_t sOr, int64_t sOc, int64_t k, float* out, int64_t sNr, int64_t sNc,
                bool prepare_only) {
    TcSolver<KP>::set_attrs(h->device);
    cudaStream_t st = h->stream;
    const int saved = h->tc_precision;
    h->tc_precision = 1;   // the remainder caches / transposed remainders below exist only in the split-operand mode
    try {
        bf16 *Xr = nullptr, *Xc = nullptr, *Xr_lo = nullptr, *Xc_lo = nullptr;
        build_x_caches(h, &Xr, &Xc, &Xr_lo, &Xc_lo);
        const int R = (int)(side == 0 ? h->n : h->p), Kdim = (int)(side == 0 ? h->p : h->n);
        Factor F = alloc_factor(h, side == 0 ? "xm.F0" : "xm.F1", R, KP);     // rows of the result (only their geometry is used)
        Factor Of = alloc_factor(h, side == 0 ? "xm.O0" : "xm.O1", Kdim, KP);  // the factor being contracted against
        TcState* state = (TcState*)h->buf("xm.state", sizeof(TcState));
        float* num = h->buf_t<float>("xm.num", (size_t)std::max(h->n, h->p) * KP);
        NMF_CUDA(cudaMemsetAsync(state, 0, sizeof(TcState), st));
        if (prepare_only) {   // bf16 caches of X and every buffer now exist: the first product inside the timed loop allocates nothing
            h->tc_precision = saved;
            return;
        }
        pack_factor_kernel<<<ew_grid((int64_t)Kdim * KP), 256, 0, st>>>(O, sOr, sOc, Kdim, (int)k, KP, Of.m, Of.hi, Of.lo, Of.bT, Of.ldT);
        h->launches += 1;
        TcSolver<KP> s{h, st, state};
        s.pfx = "xm";
        s.refresh_bTlo(Of);
        s.Xs_lo = side == 0 ? Xr_lo : Xc_lo;
        s.force_x3 = true;
        s.launch_update(1, F, Of, side == 0 ? Xr : Xc, Kdim, 0.f, 0.f, num, nullptr, -1, nullptr, false);
        unpack_factor_kernel<<<ew_grid((int64_t)R * k), 256, 0, st>>>(num, R, (int)k, KP, out, sNr, sNc);
        h->launches += 1;
        NMF_CUDA(cudaGetLastError());
    } catch (...) {
        h->tc_precision = saved;
        throw;
    }
    h->tc_precision = saved;
}
}  // namespace

bool tc_xmul(nmfb200_handle* h, int side, const float* O, int64_t sOr, int64_t sOc, int64_t k, float* out, int64_t sNr, int64_t sNc) {
    const bool prepare_only = O == nullptr;   // tc_xmul(h, side, nullptr, ...): allocate and build the caches, compute nothing
    if (h->engine_opt == 1 || h->x_elt != 4 || h->tc_xmul_opt == 0) return false;
    if (h->p < 128 || h->n < 128 || pick_kp(k) == 0) return false;
    if (h->engine_opt != 2 && h->p * h->n < ((int64_t)1 << 20)) return false;   // small problems: the exact engine, as for MultUpdate
    if (h->p > (int64_t)INT32_MAX / 256 || h->n > (int64_t)INT32_MAX / 256) return false;
    switch (pick_kp(k)) {
        case 64: tc_xmul_kp<64>(h, side, O, sOr, sOc, k, out, sNr, sNc, prepare_only); return true;
        case 128: tc_xmul_kp<128>(h, side, O, sOr, sOc, k, out, sNr, sNc, prepare_only); return true;
        case 256: tc_xmul_kp<256>(h, side, O, sOr, sOc, k, out, sNr, sNc, prepare_only); return true;
        default: return false;
    }
}

void tc_shard_geometry(int64_t n, int ranks, int rank, int64_t* own_row0, int64_t* own_row1, int64_t* tile_rows) {
    const ShardGeom g = shard_geom(ranks, 128, n, 0);   // KP does not enter the ownership
    *own_row0 = g.own_row0(rank);
    *own_row1 = g.own_row1(rank);
    *tile_rows = g.trH;
}

void tc_release(nmfb200_handle* h) {
    xchg_teardown(h);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    h->side_stream = nullptr;
    for (cudaStream_t vs : h->vstreams) cudaStreamDestroy(vs);
    h->vstreams.clear();
}

}  // namespace nmfb200
