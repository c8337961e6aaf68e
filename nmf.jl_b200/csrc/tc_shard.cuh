// tc_shard.cuh -- MultUpdate(:mse) on the tensor-core engine with the rows of X and W sharded over G ranks:
// G processes / GPUs over NVLink peer memory, or G LOGICAL ranks on one GPU (option "emulate_shards": same kernels, same
// flags, same arenas -- the 1-GPU test of the sharded mathematics).
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its anonymous namespace.
//
// Ownership.  W rows and the matching rows of X belong to one rank.  The rows of H' (128-row tiles) are divided too:
// rank o OWNS tiles [o*tpo, (o+1)*tpo) -- it alone applies the multiplicative ratio to them -- and every rank keeps a
// complete bf16 copy of H' transposed (the B operand of its W-step).  One iteration (multupd.jl:95-115), per rank:
//   K2  shard_post_kernel        waits PW; W'W = sum of the ranks' partial Grams (rank order), bf16 hi/lo; stop_condition of the
//                                 PREVIOUS iteration (its W-side sums travelled with PW).  Overlaps K7 in front and K1 behind (PDL)
//   K1  mu_update_kernel<KP,1>   partial numerators (W_g' X_g)' of ALL H tiles; each tile is staged in shared memory and sent by
//                                 bulk TMA stores into the slot its OWNER keeps for this rank (NVLink); the last tile for an
//                                 owner raises NUM
//   K3  mu_update_kernel<KP,2>   own tiles only: waits NUM (all ranks); numerators = sum of the G slots in rank order, Den by tcgen05, ratio,
//                                 new rows -> local fp32 / hi / lo, and the transposed bf16 tile by TMA store into EVERY
//                                 rank's copy of H'^T (the all-gather); per-tile Gram by tcgen05
//   K4  shard_push_kernel        partial Gram H'H over the own tiles + H-side stop sums -> every rank's slot; raises H
//   K5  shard_post_kernel        waits H; HH' = sum of partial Grams (rank order), hi/lo; H-side stop sums
//   K6  mu_update_kernel<KP,0>   W-step on the local rows (unchanged single-GPU kernel)
//   K7  shard_push_kernel        partial Gram W'W + W-side stop sums -> every rank's slot (double-buffered); raises PW
// There is no reduce-scatter / all-gather kernel and no NCCL call in the loop: the reduction is the slot sum in K3's
// prologue, the gather is K3's epilogue store.  Every H row is computed by exactly one rank from operands summed in rank
// order, so the replicated state is bit-identical everywhere; every wait is bounded (trap, not hang).
#pragma once

// ---- arena geometry ---------------------------------------------------------------------------------------------
struct ShardGeom {
    int G = 0, KP = 0;
    int64_t n = 0;
    int trH = 128, tilesH = 0, tpo = 1;   // H tile height, number of H tiles, tiles per owner
    int rowsT = 0;
    int64_t ldT = 0;
    size_t slot_rows = 0;                 // rows of one numerator slot = tpo * trH
    size_t gram_slot = 0;                 // bytes of one Gram slot: KP*KP floats + 2*KP doubles
    size_t off_cnt = 512, off_num = 1024, off_pw = 0, off_ph = 0, off_hbt = 0, off_hm = 0, bytes = 0;
    int own_tiles(int g) const { return std::max(0, std::min(tpo, tilesH - g * tpo)); }
    int64_t own_row0(int g) const { return std::min<int64_t>(n, (int64_t)g * tpo * trH); }
    int64_t own_row1(int g) const { return std::min<int64_t>(n, (int64_t)(g + 1) * tpo * trH); }
};

ShardGeom shard_geom(int G, int KP, int64_t n, int forced_tile_rows) {
    ShardGeom g;
    g.G = G; g.KP = KP; g.n = n;
    g.trH = pick_tile_rows((int)n, forced_tile_rows);
    g.tilesH = (int)ceil_div(n, g.trH);
    g.tpo = (int)ceil_div(g.tilesH, G);
    g.rowsT = KP < 128 ? 128 : KP;
    g.ldT = round_up(n, 64);
    g.slot_rows = (size_t)g.tpo * g.trH;
    g.gram_slot = (size_t)KP * KP * sizeof(float) + (size_t)2 * KP * sizeof(double);
    auto al = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
    g.off_pw = al(g.off_num + (size_t)G * g.slot_rows * KP * sizeof(float));
    g.off_ph = al(g.off_pw + (size_t)2 * G * g.gram_slot);
    g.off_hbt = al(g.off_ph + (size_t)G * g.gram_slot);
    g.off_hm = al(g.off_hbt + (size_t)g.rowsT * g.ldT * sizeof(bf16));
    g.bytes = al(g.off_hm + (size_t)n * KP * sizeof(float));
    return g;
}

struct ShardDev {  // device view of the arenas, as seen by one (logical) rank
    char* arena[XCHG_MAX_RANKS];
    int G, rank;
};

// ---- flags ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int* shard_flag(char* arena, int phase, int src) {
    return (unsigned int*)arena + phase * XCHG_MAX_RANKS + src;
}
// every thread of the block calls; returns when all ranks have published `epoch` for `phase` in THIS rank's arena
__device__ __forceinline__ void shard_wait(const ShardDev& x, int phase, unsigned int epoch) {
    if ((int)threadIdx.x < x.G) {
        const unsigned int* f = shard_flag(x.arena[x.rank], phase, threadIdx.x);
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if (clock64() - t0 > 20000000000LL) {  // ~10 s
                printf("nmfb200: peer flag wait timed out (rank %d waiting for rank %d, phase %d, epoch %u)\n", x.rank, (int)threadIdx.x, phase, epoch);
                __trap();
            }
        }
    }
    __syncthreads();
}
// threads < G publish `epoch` for `phase` in every rank's arena (everything this block wrote before is ordered in front)
__device__ __forceinline__ void shard_signal(const ShardDev& x, int phase, unsigned int epoch) {
    if ((int)threadIdx.x < x.G) {
        __threadfence_system();
        st_release_sys(shard_flag(x.arena[threadIdx.x], phase, x.rank), epoch);
    }
}

// ---- K4 / K7: partial Gram of this rank's tiles + its stop_condition partial sums -> every rank's slot ----------------
// blocks [0, gram_blocks): 4 lanes per Gram element (as gram_reduce_kernel); the next 2*KP/32 blocks: quantity q (0 = dev,
// 1 = sum) of 32 components over this rank's tiles, in Float64.  Everything is reduced into THIS rank's copy of its slot;
// the last block to finish then copies the finished slot (KP*KP floats + 2*KP doubles) to every peer with coalesced 16-byte
// stores and raises `phase` -- one burst per peer instead of 4-byte stores and a system-scope fence per block.
__global__ void __launch_bounds__(256) shard_push_kernel(ShardDev x, int phase, unsigned int epoch, const float* __restrict__ gpart,
                                                         int nparts, int KP, int gram_blocks, size_t slot_off, size_t slot_bytes,
                                                         const float* __restrict__ conv_part, int tiles, unsigned int* ticket,
                                                         const TcState* st) {
    pdl_launch_dependents();  // the kernel behind us may set itself up now; it synchronises with us through flags / griddepcontrol.wait
    if (st->converged) return;
    __shared__ double red[8][32];
    __shared__ int is_last;
    const int nelem = KP * KP;
    char* mine = x.arena[x.rank] + slot_off;
    if ((int)blockIdx.x < gram_blocks) {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        const int sub = t & 3, i = t >> 2;
        float acc = 0.f;
        if (i < nelem) {
            int g = sub;
            for (; g + 28 < nparts; g += 32) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldcg(gpart + (size_t)(g + 4 * u) * nelem + i);
                acc += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
            for (; g < nparts; g += 4) acc += __ldcg(gpart + (size_t)g * nelem + i);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (i < nelem && sub == 0) ((float*)mine)[i] = acc;
    } else {
        const int cblock = blockIdx.x - gram_blocks, cbs = KP / 32;
        const int q = cblock / cbs, cb = cblock % cbs;
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int c = cb * 32 + lane;
        const float* part = conv_part + (size_t)q * KP + c;
        double s = 0.0;
        for (int t = w; t < tiles; t += 8) s += (double)__ldcg(part + (size_t)t * 2 * KP);
        red[w][lane] = s;
        __syncthreads();
        if (w == 0) {
            double tot = red[0][lane];
#pragma unroll
            for (int i = 1; i < 8; ++i) tot += red[i][lane];
            ((double*)(mine + (size_t)nelem * sizeof(float)))[q * KP + c] = tot;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) *ticket = 0u;
    __threadfence();
    // the slot is complete in local memory: one coalesced copy per peer (gram_blocks == 0: only the sums part is fresh)
    const size_t first = gram_blocks > 0 ? 0 : (size_t)nelem * sizeof(float);
    const int n16 = (int)((slot_bytes - first) / 16);
    const uint4* src = (const uint4*)(mine + first);
    for (int i0 = threadIdx.x; i0 < n16; i0 += 4 * 256) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * 256 < n16) v[u] = __ldcg(src + i0 + u * 256);
        for (int j = 0; j < x.G; ++j) {
            if (j == x.rank) continue;
            uint4* dst = (uint4*)(x.arena[j] + slot_off + first);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * 256 < n16) dst[i0 + u * 256] = v[u];
        }
    }
    __syncthreads();
    shard_signal(x, phase, epoch);
}

// ---- K2 / K5: sum the ranks' slots in rank order -----------------------------------------------------------------
// block 0: the 2*KP stop sums -> acc_dst (Float64), optionally the decision of common.jl:105-106 (conv_decide);
// blocks 1..: Gram element i = sum_j slot_j[i] -> P (fp32) and its bf16 hi/lo split.
__global__ void __launch_bounds__(256) shard_post_kernel(ShardDev x, int phaseA, unsigned int epochA, int phaseB, unsigned int epochB,
                                                         size_t slots_off, size_t slot_stride, int KP, int k, int do_P,
                                                         float* __restrict__ P, bf16* __restrict__ Phi, bf16* __restrict__ Plo,
                                                         double* __restrict__ acc, int acc_off, int h_fixed, int decide, float tol,
                                                         TcState* st, unsigned int* done_flag, unsigned int* done_ticket) {
    pdl_launch_dependents();  // a dependent update kernel may set itself up; it waits for our completion before it reads
    if (st->converged) return;
    if (phaseA >= 0) shard_wait(x, phaseA, epochA);
    if (phaseB >= 0) shard_wait(x, phaseB, epochB);
    __shared__ float devs[256];
    __shared__ int fail, is_last;
    const int nelem = KP * KP;
    const char* base = x.arena[x.rank] + slots_off;
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < 2 * KP; i += blockDim.x) {
            double s = 0.0;
            for (int j = 0; j < x.G; ++j) s += __ldcg((const double*)(base + (size_t)j * slot_stride + (size_t)nelem * sizeof(float)) + i);
            acc[acc_off + i] = s;
            if (h_fixed) acc[2 * KP + i] = i < KP ? 0.0 : 1.0;  // H untouched: dev_h = 0 (sum_h only scales a ratio of 0)
        }
        __syncthreads();
        if (decide) conv_decide(acc, KP, k, tol, st, devs, &fail);
    } else if (do_P) {
        const int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x;
        if (i < nelem) {
            float v = 0.f;
            for (int j = 0; j < x.G; ++j) v += __ldcg((const float*)(base + (size_t)j * slot_stride) + i);
            P[i] = v;
            const bf16 hi = __float2bfloat16_rn(v);
            Phi[i] = hi;
            Plo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
    if (done_flag == nullptr) return;
    // side-stream use: the W-step running concurrently polls done_flag before its denominator blocks
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(done_ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        *done_ticket = 0u;
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(done_flag), "r"(epochA) : "memory");
    }
}

// ---- Ksum: the partial numerators of this rank's own H rows, summed over the ranks' slots in rank order ----------------
// One coalesced 16-byte load per slot and thread (the epilogue of the update kernel would read the same data row by row,
// 16 bytes per lane at a 512-byte stride: measured ~2.3 us per slot there, ~3 us for ALL slots here).
__global__ void __launch_bounds__(256) shard_slot_sum_kernel(ShardDev x, unsigned int epoch, size_t num_off, size_t slot_stride4, int64_t n4,
                                                             float4* __restrict__ out, const TcState* st, int signal_first) {
    pdl_launch_dependents();  // the ratio kernel behind us may set itself up; it waits for our completion before it reads `out`
    if (st->converged) return;
    // deferred NUM flag: the numerator kernel in front of us (same stream, complete) pushed this rank's partials into the owners'
    // slots without waiting for the stores; the kernel boundary has performed them, publish
    if (signal_first && blockIdx.x == 0) shard_signal(x, PH_NUM, epoch);
    shard_wait(x, PH_NUM, epoch);
    const float4* base = (const float4*)(x.arena[x.rank] + num_off);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v[XCHG_MAX_RANKS];
#pragma unroll
        for (int j = 0; j < XCHG_MAX_RANKS; ++j)
            if (j < x.G) v[j] = __ldcg(base + (size_t)j * slot_stride4 + i);
        float4 a = v[0];
#pragma unroll
        for (int j = 1; j < XCHG_MAX_RANKS; ++j)
            if (j < x.G) { a.x += v[j].x; a.y += v[j].y; a.z += v[j].z; a.w += v[j].w; }
        out[i] = a;
    }
}

// ---- small helpers ----------------------------------------------------------------------------------------------
// <<<1, 32>>> each; signal for every hosted rank first, then wait (logical ranks share one stream: a kernel that waited
// for a flag raised by a kernel queued behind it would never finish)
__global__ void shard_signal_kernel(ShardDev x, int phase, unsigned int epoch) { shard_signal(x, phase, epoch); }
__global__ void shard_wait_kernel(ShardDev x, int phase, unsigned int epoch, const TcState* st) {  // <<<1, 32>>>
    pdl_launch_dependents();  // a dependent W-step may set itself up; its producer waits for our completion before the first load
    if (st != nullptr && st->converged) return;  // the loop has stopped: nobody raises this flag any more
    shard_wait(x, phase, epoch);
}

// copy a 2D region of this rank's arena (rows x width bytes at byte offset off, row pitch `pitch`; 16-byte granules) into
// every other rank's arena at the same place; the last block raises `phase` (phase < 0: no flag)
__global__ void __launch_bounds__(256) shard_copy2d_kernel(ShardDev x, size_t off, size_t pitch, int rows, int width16, int phase,
                                                           unsigned int epoch, unsigned int* ticket, const TcState* st) {
    __shared__ int is_last;
    if (st != nullptr && st->converged) return;
    const int64_t total = (int64_t)rows * width16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const size_t o = off + (size_t)(i / width16) * pitch + (size_t)(i % width16) * 16;
        const uint4 v = *(const uint4*)(x.arena[x.rank] + o);
        for (int j = 0; j < x.G; ++j)
            if (j != x.rank) *(uint4*)(x.arena[j] + o) = v;
    }
    if (phase < 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) *ticket = 0u;
    shard_signal(x, phase, epoch);
}

__global__ void split_hi_lo_kernel(const float* __restrict__ F, int64_t len, bf16* __restrict__ hi, bf16* __restrict__ lo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = F[i];
        const bf16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ---- arenas: allocate; real ranks export theirs through CUDA IPC and import every peer's ---------------------------
void xchg_teardown(nmfb200_handle* h) {
    Xchg& x = h->xchg;
    for (int j = 0; j < XCHG_MAX_RANKS; ++j) {
        if (x.ipc && x.arena[j] && j != x.rank) cudaIpcCloseMemHandle(x.arena[j]);
        x.arena[j] = nullptr;
    }
    if (x.arena_local) cudaFree(x.arena_local);
    x.arena_local = nullptr;
    x.ready = false;
    x.ipc = false;
}

// Collective over the communicator (real ranks) or local (logical ranks).  Returns false if peer mapping is impossible.
bool xchg_setup(nmfb200_handle* h, const ShardGeom& geom, bool emulate) {
    Xchg& x = h->xchg;
    const int G = geom.G;
    const int rank = emulate ? 0 : h->rank;
    if (x.ready && x.G == G && x.rank == rank && x.arena_bytes == geom.bytes && x.ipc == !emulate) return true;
    xchg_teardown(h);
    x.G = G;
    x.rank = rank;
    x.arena_bytes = geom.bytes;
    x.epoch = 0;
    if (emulate) {
        for (int j = 0; j < G; ++j) {
            x.arena[j] = h->buf("shard.arena" + std::to_string(j), geom.bytes);
            NMF_CUDA(cudaMemsetAsync(x.arena[j], 0, geom.bytes, h->stream));
        }
        x.ipc = false;
        x.ready = true;
        return true;
    }
    NMF_CUDA(cudaMalloc(&x.arena_local, geom.bytes));
    NMF_CUDA(cudaMemsetAsync(x.arena_local, 0, geom.bytes, h->stream));
    cudaIpcMemHandle_t mine;
    NMF_CUDA(cudaIpcGetMemHandle(&mine, x.arena_local));
    char* dsend = (char*)h->buf("shard.ipc_send", sizeof(mine));
    char* drecv = (char*)h->buf("shard.ipc_recv", sizeof(mine) * XCHG_MAX_RANKS);
    NMF_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
    NMF_NCCL(NcclApi::get().AllGather(dsend, drecv, sizeof(mine), ncclChar, h->comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all(G);
    NMF_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(mine) * G, cudaMemcpyDeviceToHost, h->stream));
    NMF_CUDA(cudaStreamSynchronize(h->stream));
    x.ipc = true;
    bool ok = true;
    for (int j = 0; j < G && ok; ++j) {
        if (j == rank) {
            x.arena[j] = x.arena_local;
            continue;
        }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
        } else {
            x.arena[j] = p;
        }
    }
    // agree on the outcome (this is also the barrier that guarantees every arena was zeroed before anybody signals)
    if (!h->all_ranks(ok)) {
        xchg_teardown(h);
        return false;
    }
    x.ready = true;
    return true;
}

// ---- the solve ------------------------------------------------------------------------------------------------
template <int KP>
struct ShardRank {  // one (logical) rank hosted by this process
    int g = 0;
    std::string pfx;
    const float* X = nullptr;
    int64_t p = 0, ldx = 0, row0 = 0;    // this rank's rows [row0, row0 + p) of the (logical) whole
    bf16 *Xr = nullptr, *Xc = nullptr;
    Factor W, H, Hown;
    float* numsum = nullptr;             // [slot_rows][KP]: numerators of the own H rows, summed over the ranks (Ksum -> K3)
    float* h_gram_part = nullptr;        // tile Grams of the own H rows (K3 -> K4)
    int h_gram_parts = 0;
    int own_tiles = 0;                   // H tiles (of height geom.trH) this rank owns
    int64_t own_r0 = 0, own_r1 = 0;
    TcState* state = nullptr;
    double* acc = nullptr;
    unsigned int* ticket = nullptr;      // arena-resident counters: [0] K7, [1] K4, [2] K3 (PH_HBT), [3] K5 done, [4] copy kernels,
    unsigned int* own_cnt = nullptr;     //   [5] the "Gram of H is in place" flag itself; own_cnt: K1's per-owner tile counters
    ShardDev dev;
    ShardLaunch sl1, sl3, sl6, slf;
    TcSolver<KP> s;
};

template <int KP>
void tc_solve_sharded_kp(nmfb200_handle* h, const SolveArgs& a, float* Wc, int64_t ldw, float* Hc, int64_t ldh, nmfb200_result* out) {
    TcSolver<KP>::set_attrs(h->device);
    cudaStream_t st = h->stream;
    const bool emulate = h->comm == nullptr;
    const int G = emulate ? h->emulate_shards : h->nranks;
    NMF_REQUIRE(G >= 2 && G <= XCHG_MAX_RANKS, NMFB200_ENOTSUP, "row-sharded tensor-core solves cover 2..8 ranks");
    const int64_t n = h->n, k = a.k;
    const float delta = std::sqrt(std::numeric_limits<float>::epsilon());
    const float lw = (float)a.lambda_w, lh = (float)a.lambda_h, tol = (float)a.tol;
    const ShardGeom geom = shard_geom(G, KP, n, h->tc_tile_rows);
    NMF_REQUIRE(xchg_setup(h, geom, emulate), NMFB200_ENOTSUP, "peer memory between the GPUs is not available (tensor-core engine needs it)");
    Xchg& xc = h->xchg;

    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    NMF_CUDA(cudaEventRecord(e0, st));

    // stage the caller's factors on the device (column-major W rows x k, H k x n)
    const int64_t p_all = h->p;  // real rank: its shard; emulation: the whole matrix, cut below
    float *Wd = Wc, *Hd = Hc;
    int64_t ldwd = ldw, ldhd = ldh;
    if (!a.on_device) {
        Wd = h->buf_t<float>("tc.Wstage", (size_t)p_all * k);
        Hd = h->buf_t<float>("tc.Hstage", (size_t)k * n);
        ldwd = p_all;
        ldhd = k;
        NMF_CUDA(cudaMemcpy2DAsync(Wd, p_all * sizeof(float), Wc, ldw * sizeof(float), p_all * sizeof(float), k, cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hd, k * sizeof(float), Hc, ldh * sizeof(float), k * sizeof(float), n, cudaMemcpyHostToDevice, st));
    }

    const int V = emulate ? G : 1;  // logical ranks hosted here
    std::vector<ShardRank<KP>> R(V);
    // K3 works on its own H rows in tiles of 64 (twice the CTAs for the same rows: K3 is latency-, not bandwidth-bound)
    const int tr3 = (geom.trH == 128 && KP <= 128) ? 64 : geom.trH;
    for (int v = 0; v < V; ++v) {
        ShardRank<KP>& r = R[v];
        r.g = emulate ? v : h->rank;
        r.pfx = emulate ? "v" + std::to_string(v) : std::string("tc");
        if (emulate) {  // balanced contiguous row ranges, as nmf.jl_b200/dist.py::row_shard
            const int64_t base = p_all / G, extra = p_all % G;
            r.row0 = r.g * base + std::min<int64_t>(r.g, extra);
            r.p = base + (r.g < extra ? 1 : 0);
        } else {
            r.row0 = 0;
            r.p = p_all;
        }
        NMF_REQUIRE(r.p >= 1, NMFB200_EDIM, "a rank has no rows of X");
        r.X = (const float*)h->dX + r.row0;
        r.ldx = h->ldx;
        build_x_caches(h, r.pfx, r.X, r.p, n, r.ldx, &r.Xr, &r.Xc);
        char* arena = (char*)xc.arena[r.g];
        r.W = alloc_factor(h, r.pfx + ".W", (int)r.p, KP);
        // H: fp32 master and the transposed bf16 copy live in the arena (peers write into them)
        Factor& H = r.H;
        H.R = (int)n;
        H.ldT = geom.ldT;
        H.rowsT = geom.rowsT;
        H.tile_rows = geom.trH;
        H.tiles = geom.tilesH;
        H.m = (float*)(arena + geom.off_hm);
        H.bT = (bf16*)(arena + geom.off_hbt);
        H.hi = h->buf_t<bf16>(r.pfx + ".H.hi", (size_t)n * KP);
        H.lo = h->buf_t<bf16>(r.pfx + ".H.lo", (size_t)n * KP);
        H.P = h->buf_t<float>(r.pfx + ".H.P", (size_t)KP * KP);
        H.Phi = h->buf_t<bf16>(r.pfx + ".H.Phi", (size_t)KP * KP);
        H.Plo = h->buf_t<bf16>(r.pfx + ".H.Plo", (size_t)KP * KP);
        H.colsum = h->buf_t<float>(r.pfx + ".H.colsum", (size_t)KP);
        r.own_tiles = geom.own_tiles(r.g);
        r.own_r0 = geom.own_row0(r.g);
        r.own_r1 = geom.own_row1(r.g);
        r.Hown = H;
        r.Hown.tile_rows = tr3;
        r.Hown.tiles = (int)ceil_div(r.own_r1 - r.own_r0, tr3);
        r.Hown.conv = h->buf_t<float>(r.pfx + ".H.conv", (size_t)std::max(r.Hown.tiles, 1) * 2 * KP);
        H.conv = r.Hown.conv;
        r.numsum = h->buf_t<float>(r.pfx + ".numsum", std::max<size_t>(geom.slot_rows, 1) * KP);
        r.state = (TcState*)h->buf(r.pfx + ".state", sizeof(TcState));
        r.acc = h->buf_t<double>(r.pfx + ".acc", 4 * KP);
        r.ticket = (unsigned int*)(arena + geom.off_cnt + 64);
        r.own_cnt = (unsigned int*)(arena + geom.off_cnt);
        std::memset(&r.dev, 0, sizeof(r.dev));
        for (int j = 0; j < G; ++j) r.dev.arena[j] = (char*)xc.arena[j];
        r.dev.G = G;
        r.dev.rank = r.g;
        // K1: where this rank's partial numerators go
        ShardLaunch& s1 = r.sl1;
        s1.G = G; s1.tiles_per_owner = geom.tpo; s1.tiles_total = geom.tilesH; s1.own_cnt = r.own_cnt;
        s1.slot_rows = (int64_t)geom.slot_rows;
        for (int o = 0; o < G; ++o) {
            s1.num_peer[o] = (float*)((char*)xc.arena[o] + geom.off_num) + (size_t)r.g * geom.slot_rows * KP;
            s1.num_flag[o] = (unsigned int*)xc.arena[o] + PH_NUM * XCHG_MAX_RANKS + r.g;
        }
        // K3: own tiles; transposed tile into every other rank's H'^T
        ShardLaunch& s3 = r.sl3;
        s3.tile0 = (int)(r.own_r0 / tr3);
        s3.num_row0 = (int)r.own_r0;
        s3.G = G;
        s3.num_wait = nullptr;   // Ksum has waited for the NUM flags and summed the slots
        s3.own_cnt = r.ticket + 2;
        for (int j = 0; j < G; ++j) s3.num_flag[j] = (unsigned int*)xc.arena[j] + PH_HBT * XCHG_MAX_RANKS + r.g;
        s3.n_peer = 0;
        for (int j = 0; j < G; ++j)
            if (j != r.g) s3.peer_bT[s3.n_peer++] = (bf16*)((char*)xc.arena[j] + geom.off_hbt);
        if (KP > 128) s3.n_peer = 0;  // no staged epilogue at KP = 256: the slab is pushed by shard_copy2d_kernel
        r.sl6.wait_first = 1;
        r.sl6.den_flag = r.ticket + 5;
        r.sl6.G = G;
        s3.rank = r.g;
        s3.hbt_cnt = r.ticket + 2;
        for (int j = 0; j < G; ++j) s3.hbt_flag[j] = (unsigned int*)xc.arena[j] + PH_HBT * XCHG_MAX_RANKS + r.g;
        // fused H-step (MODE 6): K1's and K3's arguments in one launch; tiles are visited starting behind the own range
        ShardLaunch& sf = r.slf;
        sf = s1;
        sf.rank = r.g;
        sf.tile0 = (r.g + 1) * geom.tpo < geom.tilesH ? (r.g + 1) * geom.tpo : 0;
        sf.num_row0 = (int)r.own_r0;
        sf.num_wait = (const unsigned int*)arena + PH_NUM * XCHG_MAX_RANKS;
        sf.hbt_cnt = s3.hbt_cnt;
        for (int j = 0; j < G; ++j) sf.hbt_flag[j] = s3.hbt_flag[j];
        sf.n_peer = s3.n_peer;
        for (int j = 0; j < XCHG_MAX_RANKS - 1; ++j) sf.peer_bT[j] = s3.peer_bT[j];
        r.s = TcSolver<KP>{h, st, r.state};
        r.s.pfx = r.pfx;

        NMF_CUDA(cudaMemsetAsync(r.state, 0, sizeof(TcState), st));
        NMF_CUDA(cudaMemsetAsync(r.W.bT, 0, (size_t)r.W.rowsT * r.W.ldT * sizeof(bf16), st));
        NMF_CUDA(cudaMemsetAsync(H.bT, 0, (size_t)H.rowsT * H.ldT * sizeof(bf16), st));
        const float* Wsrc = emulate ? Wd + r.row0 : Wd;
        pack_factor_kernel<<<ew_grid(r.p * KP), 256, 0, st>>>(Wsrc, 1, ldwd, (int)r.p, (int)k, KP, r.W.m, r.W.hi, r.W.lo, r.W.bT, r.W.ldT);
        pack_factor_kernel<<<ew_grid(n * KP), 256, 0, st>>>(Hd, ldhd, 1, (int)n, (int)k, KP, H.m, H.hi, H.lo, H.bT, H.ldT);
        h->launches += 2;
    }
    NMF_CUDA(cudaGetLastError());

    const int gram_blocks = (4 * KP * KP + 255) / 256;
    const int post_blocks = 1 + (KP * KP + 255) / 256;
    const size_t gslot = geom.gram_slot;
    auto pw_off = [&](unsigned int e) { return geom.off_pw + (size_t)(e & 1u) * G * gslot; };
    // K7 (and the set-up): partial Gram W'W (if wanted) + W-side stop sums of rank r -> every rank
    auto push_W = [&](ShardRank<KP>& r, unsigned int e, bool with_gram, bool with_conv) {
        launch_k(shard_push_kernel, dim3((with_gram ? gram_blocks : 0) + 2 * (KP / 32)), dim3(256), 0, st, false, r.dev, (int)PH_PW, e,
                 (const float*)r.s.last_gram_part, with_gram ? r.s.last_gram_parts : 0, KP, with_gram ? gram_blocks : 0,
                 pw_off(e) + (size_t)r.g * gslot, gslot, (const float*)r.W.conv, with_conv ? r.W.tiles : 0, r.ticket, (const TcState*)r.state);
        h->launches += 1;
    };
    // K2: W'W of all ranks (epoch e_pw) -> P_W hi/lo, W-side stop sums, optionally the stop decision.  `chain`: launched as a
    // programmatic dependent of the push kernel in front of it (it synchronises with that kernel through the PW flag)
    auto post_W = [&](ShardRank<KP>& r, unsigned int e_pw, bool do_P, bool decide, bool chain) {
        launch_k(shard_post_kernel, dim3(do_P ? post_blocks : 1), dim3(256), 0, st, chain, r.dev, (int)PH_PW, e_pw, -1, 0u, pw_off(e_pw), gslot, KP,
                 (int)k, do_P ? 1 : 0, r.W.P, r.W.Phi, r.W.Plo, r.acc, 0, a.update_H ? 0 : 1, decide ? 1 : 0, tol, r.state, (unsigned int*)nullptr,
                 (unsigned int*)nullptr);
        h->launches += 1;
    };

    // set-up: partial Grams W'W (epoch e_init); a fixed H has its full Gram locally
    unsigned int e_prev = ++xc.epoch;
    for (auto& r : R) {
        r.s.launch_gram_parts(r.W);
        push_W(r, e_prev, true, false);
        if (!a.update_H) r.s.launch_gram(r.H, true);
    }
    // align the ranks on the DEVICE, then start the clock (a host barrier leaves the start skew inside the timed region)
    {
        const unsigned int eb = ++xc.epoch;
        for (auto& r : R) shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_BAR, eb);
        for (auto& r : R) shard_wait_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_BAR, eb, (const TcState*)nullptr);
        h->launches += 2 * V;
    }
    NMF_CUDA(cudaEventRecord(e1, st));

    h->ev_used = 0;
    const bool pdl = h->tc_pdl != 0;
    int64_t enq = 0, iters = 0;
    bool converged = false;
    float devmax = 0.f;
    std::vector<TcState> hs(V);
    // logical ranks: one stream each for the fused H-step launch (see `fused` below)
    cudaEvent_t ev_vfork = nullptr, ev_vjoin[XCHG_MAX_RANKS] = {};
    if (emulate) {
        while ((int)h->vstreams.size() < V) {
            cudaStream_t vs;
            NMF_CUDA(cudaStreamCreateWithFlags(&vs, cudaStreamNonBlocking));
            h->vstreams.push_back(vs);
        }
        NMF_CUDA(cudaEventCreateWithFlags(&ev_vfork, cudaEventDisableTiming));
        for (int v = 0; v < V; ++v) NMF_CUDA(cudaEventCreateWithFlags(&ev_vjoin[v], cudaEventDisableTiming));
    }
    // side stream for K4 / K5 (real ranks only; option tc_side_stream=0 keeps everything in stream order)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (!emulate && h->tc_side_stream && a.update_H && h->time_kernels != 2) {
        if (!h->side_stream) NMF_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        side = h->side_stream;
        NMF_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        NMF_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    const bool covers = [&] {
        bool ok = true;
        for (auto& r : R) ok = ok && tc_objective_covers(r.X, r.p, n, r.ldx);
        return h->all_ranks(ok);
    }();
    // end of solve / verbose: every rank needs all fp32 rows of H (each holds its own) and fresh hi/lo of all rows
    auto gather_H = [&]() {
        if (!a.update_H) return;
        const unsigned int eg = ++xc.epoch;
        for (auto& r : R) {
            const int64_t rows = r.own_r1 - r.own_r0;
            const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(148, ceil_div(rows * KP / 4, 256)));
            shard_copy2d_kernel<<<blocks, 256, 0, st>>>(r.dev, geom.off_hm + (size_t)r.own_r0 * KP * sizeof(float), (size_t)KP * sizeof(float),
                                                        (int)rows, KP / 4, (int)PH_GATHER, eg, r.ticket + 4, (const TcState*)nullptr);
        }
        for (auto& r : R) {
            shard_wait_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_GATHER, eg, (const TcState*)nullptr);
            split_hi_lo_kernel<<<ew_grid(n * KP), 256, 0, st>>>(r.H.m, n * KP, r.H.hi, r.H.lo);
        }
        h->launches += 3 * V;
    };
    auto objective_now = [&](int alg) -> double {
        double tot[3] = {0, 0, 0};
        if (covers) {
            std::vector<double*> res(V);
            for (int v = 0; v < V; ++v)
                res[v] = tc_objective_enqueue<KP>(h, R[v].pfx, alg, R[v].X, R[v].p, n, R[v].ldx, R[v].W, R[v].H, 0.0, 0.0);
            if (h->comm) h->allreduce_sum(res[0], 2);  // rows of X / W are sharded: data term and |W|_1 are partial; H is replicated
            for (int v = 0; v < V; ++v) {
                double hres[3];
                NMF_CUDA(cudaMemcpyAsync(hres, res[v], 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
                NMF_CUDA(cudaStreamSynchronize(st));
                tot[0] += hres[0];
                tot[1] += hres[1];
                tot[2] = hres[2];
            }
            return tc_objective_value(alg, tot, 0.0, 0.0);
        }
        return std::numeric_limits<double>::quiet_NaN();  // caller evaluates on the unpacked factors (exact engine)
    };
    double v_objv = std::numeric_limits<double>::quiet_NaN(), v_t0 = 0;
    auto wall = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    NMF_REQUIRE(!a.verbose || covers, NMFB200_ENOTSUP, "verbose on the tensor-core engine needs the tensor-core objective");
    if (a.verbose) {
        v_t0 = wall();
        v_objv = objective_now(0);
        if (h->trace) h->trace(h->trace_user, 0, 0.0, v_objv, NAN, NAN);
    }

    while (enq < a.maxiter) {
        const int64_t batch = a.verbose ? 1 : std::min<int64_t>(h->check_every, a.maxiter - enq);
        for (int64_t i = 0; i < batch; ++i) {
            const unsigned int e = ++xc.epoch;
            h->mark("start");
            // fused H-step (MODE 6): CTAs of own tiles WAIT inside the kernel for the other ranks' partials.  Real ranks run
            // concurrently by construction.  Logical ranks must be made concurrent: one stream each, and only when every CTA of
            // every logical rank fits on the GPU at once (one CTA per SM) -- otherwise the emulation keeps K1 and K3 separate.
            // Measured (profiles/r2_scaling.md): the fused form wins at G = 2 (0.239 vs 0.251 ms per iteration) and loses from
            // G = 4 on (0.300 vs 0.277 ms at G = 8): its 128/G owner CTAs sum the G slots row by row, 2.3 us per slot.  Default
            // (tc_fused_hstep = -1): fused for two ranks, K1 + Ksum + K3 for more.
            const bool want_fused = h->tc_fused_hstep < 0 ? G == 2 : h->tc_fused_hstep != 0;
            const bool fused = a.update_H && KP <= 128 && want_fused && (!emulate || (int64_t)G * geom.tilesH <= 120);
            if (a.update_H) {
                // K2 first: its inputs (the PW slots of the previous iteration) are ready long before K1 ends, and K1 only needs
                // its results in the epilogue of K3
                for (auto& r : R) post_W(r, e_prev, true, i > 0, pdl && V == 1);
                h->mark("K2 W'W + decision");
                if (fused && emulate) NMF_CUDA(cudaEventRecord(ev_vfork, st));
                for (auto& r : R) {  // K1 + K3 in one launch (MODE 6): CTAs of own tiles finish the update themselves
                    if (!fused) break;
                    if (emulate) {
                        NMF_CUDA(cudaStreamWaitEvent(h->vstreams[r.g], ev_vfork, 0));
                        r.s.st = h->vstreams[r.g];
                    }
                    r.slf.epoch = e;
                    r.s.sl = &r.slf;
                    r.s.defer_gram_reduce = true;
                    r.s.num_splits = G;
                    r.s.num_split_stride = (int64_t)geom.slot_rows * KP;
                    r.s.gram_tag = "gram_partH";
                    r.s.launch_update(6, r.H, r.W, r.Xr, (int)r.p, lh, delta, (float*)((char*)xc.arena[r.g] + geom.off_num), nullptr, 1, nullptr,
                                      pdl && !emulate);
                    if (emulate) {
                        NMF_CUDA(cudaEventRecord(ev_vjoin[r.g], r.s.st));
                        r.s.st = st;
                    }
                    r.s.num_splits = 1;
                    r.s.num_split_stride = 0;
                    r.s.defer_gram_reduce = false;
                    r.s.sl = nullptr;
                    r.h_gram_part = r.s.last_gram_part;
                    r.h_gram_parts = r.own_tiles;
                    r.s.gram_tag = "gram_part";
                    if (r.own_tiles == 0) {  // the other ranks wait for this rank's PH_HBT too
                        shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_HBT, e);
                        h->launches += 1;
                    }
                }
                if (fused && emulate)
                    for (auto& r : R) NMF_CUDA(cudaStreamWaitEvent(st, ev_vjoin[r.g], 0));
                if (fused) h->mark("K1+K3 fused H-step");
                for (auto& r : R) {  // K1
                    if (fused) break;
                    r.sl1.epoch = e;
                    r.sl1.defer_signal = (h->tc_defer_signal != 0) ? 1 : 0;
                    r.s.sl = &r.sl1;
                    r.s.launch_update(1, r.H, r.W, r.Xr, (int)r.p, lh, delta, nullptr, nullptr, -1, nullptr, pdl);
                    r.s.sl = nullptr;
                }
                if (!fused && h->tc_defer_signal != 0 && emulate)   // logical ranks share a stream: every NUM flag before any Ksum
                    for (auto& r : R) { shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_NUM, e); h->launches += 1; }
                if (!fused) h->mark("K1 numerators");
                const bool defer = !fused && h->tc_defer_signal != 0;
                for (auto& r : R) {  // Ksum
                    if (fused) break;
                    if (r.own_tiles == 0) {  // nothing to sum here, but the owners wait for this rank's NUM flag
                        if (defer) { shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_NUM, e); h->launches += 1; }
                        continue;
                    }
                    const int64_t n4 = (r.own_r1 - r.own_r0) * KP / 4;
                    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(148 * 4, ceil_div(n4, 256)));
                    launch_k(shard_slot_sum_kernel, dim3(blocks), dim3(256), 0, st, false, r.dev, e, geom.off_num, geom.slot_rows * KP / 4, n4,
                             (float4*)r.numsum, (const TcState*)r.state, defer && !emulate ? 1 : 0);
                    h->launches += 1;
                }
                if (!fused) h->mark("Ksum slots");
                for (auto& r : R) {  // K3
                    if (fused) break;
                    if (r.own_tiles == 0) {  // nothing to update here, but the other ranks wait for this rank's PH_HBT too
                        shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_HBT, e);
                        h->launches += 1;
                        continue;
                    }
                    r.sl3.epoch = e;
                    r.sl3.defer_signal = (defer && KP <= 128) ? 1 : 0;
                    r.s.sl = &r.sl3;
                    r.s.defer_gram_reduce = true;
                    r.s.num_splits = 1;
                    r.s.gram_tag = "gram_partH";   // K4 reads these on the side stream while K6 fills the W-step's tile Grams
                    // no tile Gram in the ratio kernel: H'H over the own rows comes from gram_kernel on the side stream (below), which
                    // takes the TMEM read-back and a 64 KB store per CTA out of the critical path
                    r.s.launch_update(2, r.Hown, r.W, r.Xr, (int)r.p, lh, delta, r.numsum, nullptr, -1, nullptr, pdl);
                    r.s.num_splits = 1;
                    r.s.num_split_stride = 0;
                    r.s.defer_gram_reduce = false;
                    r.s.sl = nullptr;
                    if (KP > 128) {  // no staged epilogue: slab to the peers by a copy kernel
                        const int w16 = (int)(round_up((r.own_r1 - r.own_r0) * sizeof(bf16), 16) / 16);
                        shard_copy2d_kernel<<<64, 256, 0, st>>>(r.dev, geom.off_hbt + (size_t)r.own_r0 * sizeof(bf16), (size_t)geom.ldT * sizeof(bf16),
                                                                KP, w16, (int)PH_HBT, e, r.ticket + 4, (const TcState*)r.state);
                        h->launches += 1;
                    }
                    r.s.gram_tag = "gram_part";
                    if (defer && KP <= 128 && emulate) {  // deferred HBT flag (a real rank's W-step raises it at its start)
                        shard_signal_kernel<<<1, 32, 0, st>>>(r.dev, (int)PH_HBT, e);
                        h->launches += 1;
                    }
                }
                if (!fused) h->mark("K3 own rows of H");
                // K4 + K5 only feed the denominator blocks at the END of the W-step: a real rank runs them on a side stream,
                // concurrently with the W-step's main loop (logical ranks share one stream; the flags are then already up)
                cudaStream_t sk = side ? side : st;
                if (side) {
                    NMF_CUDA(cudaEventRecord(ev_fork, st));
                    NMF_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
                }
                for (auto& r : R) {  // partial Grams H'H of the own rows (the fused H-step made them in its epilogue)
                    if (fused || r.own_tiles == 0) continue;
                    r.s.st = sk;
                    r.s.gram_tag = "gram_partH";
                    r.s.launch_gram_parts(r.H, (int)r.own_r0, (int)r.own_r1);
                    r.s.gram_tag = "gram_part";
                    r.s.st = st;
                    r.h_gram_part = r.s.last_gram_part;
                    r.h_gram_parts = r.s.last_gram_parts;
                }
                for (auto& r : R) {  // K4
                    const bool any = r.own_tiles > 0;
                    launch_k(shard_push_kernel, dim3(gram_blocks + 2 * (KP / 32)), dim3(256), 0, sk, false, r.dev, (int)PH_H, e,
                             (const float*)r.h_gram_part, any ? r.h_gram_parts : 0, KP, gram_blocks, geom.off_ph + (size_t)r.g * gslot,
                             gslot, (const float*)r.Hown.conv, any ? (fused ? r.own_tiles : r.Hown.tiles) : 0, r.ticket + 1, (const TcState*)r.state);
                    h->launches += 1;
                }
                if (!side) h->mark("K4 push H'H");
                for (auto& r : R) {  // K5
                    launch_k(shard_post_kernel, dim3(post_blocks), dim3(256), 0, sk, false, r.dev, (int)PH_H, e, -1, 0u, geom.off_ph, gslot, KP,
                             (int)k, 1, r.H.P, r.H.Phi, r.H.Plo, r.acc, 2 * KP, 0, 0, tol, r.state, r.ticket + 5, r.ticket + 3);
                    h->launches += 1;
                }
                if (!side) h->mark("K5 HH'");
                if (side) NMF_CUDA(cudaEventRecord(ev_join, side));
                // every rank's rows of H'^T must be in this rank's copy before the W-step streams it: its producer thread polls
                // the PH_HBT flags itself (real ranks); logical ranks share a stream, a wait kernel keeps the order simple there
                for (auto& r : R) {
                    if (!emulate) break;
                    launch_k(shard_wait_kernel, dim3(1), dim3(32), 0, st, false, r.dev, (int)PH_HBT, e, (const TcState*)r.state);
                    h->launches += 1;
                }
            }
            for (auto& r : R) {  // K6
                r.sl6.epoch = e;
                r.sl6.den_flag = a.update_H ? r.ticket + 5 : nullptr;
                r.sl6.hbt_wait = a.update_H ? (const unsigned int*)xc.arena[r.g] + PH_HBT * XCHG_MAX_RANKS : nullptr;
                r.sl6.signal_hbt = (a.update_H && !fused && h->tc_defer_signal != 0 && KP <= 128 && !emulate && r.own_tiles > 0) ? 1 : 0;
                for (int j = 0; j < G; ++j) r.sl6.hbt_flag[j] = r.sl3.hbt_flag[j];
                r.s.sl = &r.sl6;
                r.s.defer_gram_reduce = true;
                const int gramW = (a.update_H && KP <= 128) ? 1 : -1;
                r.s.launch_update(0, r.W, r.H, r.Xc, (int)n, lw, delta, nullptr, nullptr, gramW, nullptr, pdl && a.update_H && emulate);
                r.s.defer_gram_reduce = false;
                r.s.sl = nullptr;
                if (a.update_H && KP > 128) r.s.launch_gram_parts(r.W);
            }
            h->mark("K6 W-step");
            if (side && a.update_H) NMF_CUDA(cudaStreamWaitEvent(st, ev_join, 0));  // everything later on the main stream sees K5's results
            for (auto& r : R) push_W(r, e, a.update_H, true);  // K7
            h->mark("K7 push W'W");
            e_prev = e;
            if (!a.update_H)
                for (auto& r : R) post_W(r, e_prev, false, true, false);  // H fixed: nothing to ride on, decide every iteration
        }
        NMF_CUDA(cudaEventRecord(e2, st));  // end of the timed loop (re-recorded per batch; the last one counts)
        if (a.update_H)
            for (auto& r : R) post_W(r, e_prev, false, true, false);  // decision of the last iteration of the batch
        enq += batch;
        NMF_CUDA(cudaGetLastError());
        for (int v = 0; v < V; ++v) NMF_CUDA(cudaMemcpyAsync(&hs[v], R[v].state, sizeof(TcState), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        for (int v = 1; v < V; ++v)
            NMF_REQUIRE(hs[v].iters == hs[0].iters && hs[v].converged == hs[0].converged && hs[v].devmax == hs[0].devmax, NMFB200_ECUDA,
                        "logical ranks disagree on the stop decision (replicated state diverged)");
        iters = hs[0].iters;
        devmax = hs[0].devmax;
        if (a.verbose) {
            gather_H();
            const double pre = v_objv;
            v_objv = objective_now(0);
            if (h->trace) h->trace(h->trace_user, iters, wall() - v_t0, v_objv, v_objv - pre, (double)devmax);
        }
        if (hs[0].converged) {
            converged = true;
            break;
        }
    }

    // results back in the caller's layout (this rank's rows of W; the whole H)
    if (!a.verbose) gather_H();
    for (auto& r : R) {
        float* Wdst = emulate ? Wd + r.row0 : Wd;
        unpack_factor_kernel<<<ew_grid(r.p * k), 256, 0, st>>>(r.W.m, (int)r.p, (int)k, KP, Wdst, 1, ldwd);
        h->launches += 1;
    }
    unpack_factor_kernel<<<ew_grid(n * k), 256, 0, st>>>(R[0].H.m, (int)n, (int)k, KP, Hd, ldhd, 1);
    h->launches += 1;
    NMF_CUDA(cudaGetLastError());
    if (emulate && h->tc_debug & 32) {  // diagnostics: the replicated H must be bit-identical on every logical rank
        std::vector<float> h0((size_t)n * KP), hv((size_t)n * KP);
        NMF_CUDA(cudaMemcpyAsync(h0.data(), R[0].H.m, h0.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        for (int v = 1; v < V; ++v) {
            NMF_CUDA(cudaMemcpyAsync(hv.data(), R[v].H.m, hv.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
            NMF_CUDA(cudaStreamSynchronize(st));
            NMF_REQUIRE(std::memcmp(h0.data(), hv.data(), h0.size() * sizeof(float)) == 0, NMFB200_ECUDA, "replicated H differs between logical ranks");
        }
    }
    double objv = a.verbose ? v_objv : objective_now(0);
    if (objv != objv) objv = simt_objective_f32(h, 0, Wd, ldwd, Hd, ldhd, k, 0.0, 0.0);  // shard shape not covered: exact engine on the unpacked factors
    if (!a.on_device) {
        NMF_CUDA(cudaMemcpy2DAsync(Wc, ldw * sizeof(float), Wd, p_all * sizeof(float), p_all * sizeof(float), k, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hc, ldh * sizeof(float), Hd, k * sizeof(float), k * sizeof(float), n, cudaMemcpyDeviceToHost, st));
    }
    NMF_CUDA(cudaStreamSynchronize(st));
    if (side) NMF_CUDA(cudaStreamSynchronize(side));
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_vfork) cudaEventDestroy(ev_vfork);
    for (cudaEvent_t ev : ev_vjoin)
        if (ev) cudaEventDestroy(ev);
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
    if ((h->tc_debug & 8) && h->rank == 0) {  // diagnostics: the LAST update launch (a W-step); and see tc_debug & 64 for the H-step
        std::vector<long long> tv(16 + 2 * 4096);
        NMF_CUDA(cudaMemcpy(tv.data(), h->buf("tc.timing", tv.size() * sizeof(long long)), tv.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const long long* t = tv.data();
        const int nct = std::min(geom.tilesH, 4096);
        long long s_min = LLONG_MAX;
        for (int c = 0; c < nct; ++c) s_min = std::min(s_min, t[16 + 2 * c]);
        fprintf(stderr, "[nmfb200] last timed launch: per-CTA exit (us after the first entry):");
        for (int c = 0; c < nct; c += std::max(1, nct / 16)) fprintf(stderr, " %d:%.1f", c, (t[17 + 2 * c] - s_min) * 1e-3);
        fprintf(stderr, " last:%.1f\n", (t[17 + 2 * (nct - 1)] - s_min) * 1e-3);
        static const char* names[13] = {"entry", "first_operands", "numerators_issued", "mma_issued", "pred_complete", "accum_complete",
                                        "ratio_done", "all_warps_done", "gram_mma_done", "gram_written", "stores_done", "exit", "peers_arrived"};
        fprintf(stderr, "[nmfb200] timed CTA phase clocks (cycles since entry):");
        for (int i = 1; i < 13; ++i) fprintf(stderr, " %s=%lld", names[i], t[i] - t[0]);
        fprintf(stderr, "\n");
    }
}
