// tc_reduce.cuh -- stop_condition finish (common.jl:92-111), merged Gram + stop reduce (the sharded counterparts are in tc_shard.cuh)
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

// ---- stop_condition finish (common.jl:92-111) ---------------------------------------------------------
// acc (double [4][KP]) = {dev_w, sum_w, dev_h, sum_h}.  conv_reduce_kernel: grid = 4 * KP/32 blocks of 8 warps;
// block (q, cb) sums quantity q of components [32cb, 32cb+32) over all tiles (warp w takes tiles w, w+8, ...;
// 8 loads in flight; fixed combination order => deterministic).  With do_decide the last block to finish
// (atomic ticket) applies the reference's test; row-sharded solves decide in shard_post_kernel (tc_shard.cuh).
// Batched replicates: the reference's test per replicate (components [r*blk, (r+1)*blk)).  A replicate that passes for the first
// time is marked `newly` (batch_snapshot_kernel then keeps its factors as they are now -- the stacked iteration goes on for the
// others, and the replicates do not interact: block-diagonal Grams); the loop ends when every replicate has passed.
__device__ void conv_decide_batched(const double* acc, int KP, float tol, TcState* st, float* devs) {
    __shared__ int failc[256];
    const int a = threadIdx.x;
    const int blk = st->blk, nrep = st->nrep;
    BatchState* bs = st->bs;
    float dev = 0.f;
    int f = 0;
    if (a < blk * nrep) {
        float dw = (float)__ldcg(acc + a), sw = (float)__ldcg(acc + KP + a), dh = (float)__ldcg(acc + 2 * KP + a),
              sh = (float)__ldcg(acc + 3 * KP + a);
        float rw = dw / sw, rh = dh / sh;
        float m = (rw != rw) ? rw : ((rh != rh) ? rh : fmaxf(rw, rh));
        dev = sqrtf(m);
        f = (sqrtf(dw) > tol * sqrtf(sw) || sqrtf(dh) > tol * sqrtf(sh)) ? 1 : 0;
    }
    if (a < 256) { devs[a] = dev; failc[a] = f; }
    __syncthreads();
    if (a < nrep) {
        float dm = 0.f;
        int any = 0;
        for (int i = a * blk; i < (a + 1) * blk; ++i) {
            dm = (dm != dm) ? dm : ((devs[i] != devs[i]) ? devs[i] : fmaxf(dm, devs[i]));
            any |= failc[i];
        }
        int nw = 0;
        if (!bs->done[a]) {
            bs->devmax[a] = dm;
            if (!any) { bs->done[a] = 1; bs->niters[a] = st->iters + 1; nw = 1; }
        }
        bs->newly[a] = nw;
    }
    __syncthreads();
    if (a == 0) {
        int all = 1;
        float dm = 0.f;
        for (int r = 0; r < nrep; ++r) {
            all &= bs->done[r];
            const float d = bs->devmax[r];
            dm = (dm != dm) ? dm : ((d != d) ? d : fmaxf(dm, d));
        }
        st->devmax = dm;
        st->iters += 1;
        if (all) st->converged = 1;
    }
}

__device__ void conv_decide(const double* acc, int KP, int k, float tol, TcState* st, float* devs, int* fail) {
    if (st->bs != nullptr) { conv_decide_batched(acc, KP, tol, st, devs); return; }
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
                                                 bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split, int block, const TcState* st) {
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
        if (gram_masked(st, i)) acc = 0.f;   // batched replicates: block-diagonal Gram
        P[i] = acc;
        if (do_split) {
            bf16 hi = __float2bfloat16_rn(acc);
            Phi[i] = hi;
            Plo[i] = __float2bfloat16_rn(acc - __bfloat162float(hi));
        }
    }
}

// Trace-identity objective (verbose): 0.5*||X - WH||^2 = 0.5*(||X||^2 - 2<XH', W> + <W'W, HH'>) from quantities the
// iteration has on hand -- the W-step's numerators times the new W (per-tile sums in cross_part), the two k x k Grams -- instead
// of a pass over X.  Evaluated by the last block of gram_conv_reduce_kernel to finish (all-blocks ticket).
struct TraceObj {
    const float* cross_part;  // nullptr = off
    int ntiles;
    const float* P_other;     // Gram of the other factor (H H'), KP x KP fp32
    double xnorm2;            // ||X||^2 (Float64, once per set_X)
};

// One VIRTUAL block of the merged reduce (vb in [0, nvb)): a Gram block (vb < gram_blocks) or a stop_condition block.
__device__ void gram_conv_virtual_block(int vb, int nvb, const float* __restrict__ gpart, int nparts, int nelem, float* __restrict__ P,
                                        bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split, int gram_blocks,
                                        const float* __restrict__ partW, int tilesW, const float* __restrict__ partH, int tilesH, int KP, int k,
                                        int update_H, double* __restrict__ acc, float tol, TcState* st, int do_decide,
                                        float* __restrict__ wsums_f32, const TraceObj& tr) {
    __shared__ double red[8][32];
    __shared__ float devs[256];
    __shared__ int fail, is_last, is_last_all;
    if (vb < gram_blocks) {
        gram_reduce_body(gpart, nparts, nelem, P, Phi, Plo, do_split, vb, st);
    } else {
        const int cblock = vb - gram_blocks, nconv = nvb - gram_blocks;
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
        if (do_decide) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int prev = atomicAdd(&st->ticket, 1u);
                is_last = (prev == (unsigned)nconv - 1) ? 1 : 0;
            }
            __syncthreads();
            if (is_last) {
                __threadfence();
                if (threadIdx.x == 0) st->ticket = 0u;
                conv_decide(acc, KP, k, tol, st, devs, &fail);
            }
        }
    }
    if (tr.cross_part == nullptr) return;
    // ---- trace-identity objective: the last of ALL (virtual) blocks sees the complete Gram P of this factor
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last_all = (atomicAdd(&st->ticket2, 1u) == (unsigned)nvb - 1u) ? 1 : 0;
    __syncthreads();
    if (!is_last_all) return;
    __threadfence();
    double part = 0.0;
    for (int i = threadIdx.x; i < nelem; i += 256) part += (double)__ldcg(P + i) * (double)__ldcg(tr.P_other + i);   // <W'W, HH'>
    double cross = 0.0;
    for (int i = threadIdx.x; i < tr.ntiles; i += 256) cross += (double)__ldcg(tr.cross_part + i);                    // <XH', W>
    part -= 2.0 * cross;
    double* sred = &red[0][0];
    sred[threadIdx.x] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sred[threadIdx.x] += sred[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st->ticket2 = 0u;
        const double d = tr.xnorm2 + sred[0];
        st->objv = (double)(0.5f * (float)(d > 0.0 ? d : 0.0));   // convert(T, 0.5) * sqL2dist (multupd.jl:81)
    }
}

// The work is cut into nvb = gram_blocks + 4*KP/32 VIRTUAL blocks; a launch with gridDim.x == nvb gives each its own CTA (the default),
// a smaller grid walks them (option tc_chain: a handful of CTAs that are resident early, next to the update kernel's, and whose
// running time is hidden under the next update launch's 75 us of streaming).
__global__ void __launch_bounds__(256) gram_conv_reduce_kernel(const float* __restrict__ gpart, int nparts, int nelem, float* __restrict__ P,
                                                               bf16* __restrict__ Phi, bf16* __restrict__ Plo, int do_split, int gram_blocks,
                                                               const float* __restrict__ partW, int tilesW, const float* __restrict__ partH,
                                                               int tilesH, int KP, int k, int update_H, double* __restrict__ acc, float tol,
                                                               TcState* st, int do_decide, float* __restrict__ wsums_f32, TraceObj tr,
                                                               int chained, int nvb) {
    pdl_launch_dependents();  // the next H-step may start streaming X now; it waits for us before it reads P / `converged`
    if (chained) pdl_wait();   // option tc_chain: launched as a programmatic dependent of the W-step (see gram_reduce_kernel)
    if (st->converged) return;
    for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
        gram_conv_virtual_block(vb, nvb, gpart, nparts, nelem, P, Phi, Plo, do_split, gram_blocks, partW, tilesW, partH, tilesH, KP, k, update_H,
                                acc, tol, st, do_decide, wsums_f32, tr);
        __syncthreads();   // the shared scratch of one virtual block is reused by the next
    }
}

// Batched replicates: keep the factors of every replicate that met stop_condition in the iteration just decided (final != 0: of
// every replicate that never did -- the end of the loop).  Fm / snap: [rows][KP] fp32 masters, W rows then H rows as two calls' worth
// of work in one launch.  Launched behind gram_conv_reduce_kernel every iteration; copies nothing unless a flag is up.
__global__ void __launch_bounds__(256) batch_snapshot_kernel(const TcState* __restrict__ st, int KP, const float* __restrict__ Wm,
                                                             float* __restrict__ Wsnap, int64_t lenW, const float* __restrict__ Hm,
                                                             float* __restrict__ Hsnap, int64_t lenH, int final) {
    pdl_launch_dependents();   // the next H-step may start streaming; it waits for this kernel's completion before it touches H
    __shared__ int flag[MAX_BATCH];
    __shared__ int any;
    const BatchState* bs = st->bs;
    const int nrep = st->nrep, blk = st->blk;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    if ((int)threadIdx.x < nrep) {
        const int f = final ? (bs->done[threadIdx.x] ? 0 : 1) : bs->newly[threadIdx.x];
        flag[threadIdx.x] = f;
        if (f) any = 1;
    }
    __syncthreads();
    if (!any) return;
    const int64_t total = lenW + lenH;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const bool w = i < lenW;
        const int64_t j = w ? i : i - lenW;
        const int r = (int)(j % KP) / blk;
        if (r < nrep && flag[r]) {
            if (w) Wsnap[j] = Wm[j];
            else Hsnap[j] = Hm[j];
        }
    }
}
