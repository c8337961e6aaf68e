// tc_xchg.cuh -- multi-GPU exchange over NVLink peer memory: fused reduce-scatter / all-gather kernel, flag barriers, IPC arena set-up
// Part of the tensor-core engine; included only by tc_engine.cu, inside namespace nmfb200 and its
// anonymous namespace.
#pragma once

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
