// common.cuh -- handle, device-buffer cache, error plumbing and NCCL loader shared by the engines.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <nccl.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/nmfb200.h"

namespace nmfb200 {

struct Error {
    int status;
    std::string msg;
};

#define NMF_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            throw ::nmfb200::Error{e__ == cudaErrorMemoryAllocation ? NMFB200_ENOMEM : NMFB200_ECUDA, \
                                   std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" +    \
                                       __FILE__ + ":" + std::to_string(__LINE__) + ")"};           \
        }                                                                                          \
    } while (0)

#define NMF_REQUIRE(cond, status, text)                        \
    do {                                                       \
        if (!(cond)) throw ::nmfb200::Error{(status), (text)}; \
    } while (0)

// ---- NCCL through dlopen: no link-time dependency, binds to the libnccl.so.2 already in the process
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    static NcclApi& get() {
        static NcclApi api;
        if (!api.lib) {
            api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            NMF_REQUIRE(api.lib, NMFB200_ENCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
            api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            NMF_REQUIRE(api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString,
                        NMFB200_ENCCL, "libnccl.so.2 lacks a required symbol");
        }
        return api;
    }
};

#define NMF_NCCL(expr)                                                                                     \
    do {                                                                                                   \
        ncclResult_t r__ = (expr);                                                                         \
        if (r__ != ncclSuccess)                                                                            \
            throw ::nmfb200::Error{NMFB200_ENCCL, std::string(#expr) + ": " +                              \
                                                      ::nmfb200::NcclApi::get().GetErrorString(r__)};      \
    } while (0)

struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
};

template <typename T> struct NcclType;
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat32; };
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclFloat64; };

}  // namespace nmfb200

namespace nmfb200 {
constexpr int XCHG_MAX_RANKS = 8;
// Peer-memory arenas of a row-sharded solve (one per rank; a real rank maps every peer's arena through CUDA IPC, logical
// ranks of an emulated solve simply allocate theirs on the same GPU).  Layout and protocol: csrc/tc_shard.cuh.
struct Xchg {
    bool ready = false;
    bool ipc = false;                         // arenas of the other ranks are IPC mappings (real multi-GPU)
    int G = 0, rank = 0;
    size_t arena_bytes = 0;                   // geometry the arenas were built for
    void* arena_local = nullptr;              // real multi-GPU: this rank's arena (cudaMalloc)
    void* arena[XCHG_MAX_RANKS] = {};         // [rank] -> base (own / logical ranks: local allocation; peers: IPC mapping)
    unsigned int epoch = 0;
};
}  // namespace nmfb200

// The opaque handle of the C ABI.
struct nmfb200_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;

    // options
    int engine_opt = 0;  // 0 auto, 1 simt, 2 tc
    int check_every = 8;
    int time_kernels = 0;
    int tc_tile_rows = 0;  // 0 = auto
    int tc_debug = 0;      // diagnostics: bit 3 (8) = record and print the phase clocks of the update kernel
    int tc_xchg = 1;       // multi-GPU on the tensor-core engine needs peer memory; 0 = keep multi-GPU solves on the exact engine (NCCL)
    int tc_fused_hstep = -1; // row-sharded solves, k <= 128: 1 = one launch for the H-step (own tiles finish the update in the numerator
                             // kernel), 0 = numerators / slot sum / ratio as three launches, -1 = auto (fused for two ranks)
    int tc_defer_signal = 1; // row-sharded, > 2 ranks: the numerator / ratio kernels do not wait for their peer stores; the next kernel in
                             // the stream (kernel boundary) raises the NUM / HBT flags
    int tc_side_stream = 1;  // row-sharded solves: run the H-Gram exchange (K4/K5) on a side stream, concurrently with the W-step
    cudaStream_t side_stream = nullptr;
    std::vector<cudaStream_t> vstreams;  // logical ranks (emulate_shards): one stream each for launches that wait on each other
    int emulate_shards = 0;  // > 1: run the row-sharded tensor-core algorithm with this many LOGICAL ranks on this one GPU
    int tc_xmul_opt = 1;   // ProjectedALS / CoordinateDescent / ALSPGrad (Float32): X-sized products on the tensor cores (split operands)
    int tc_flush = -1;     // k-blocks per TMEM accumulation chunk of the update kernel (0 = one long chain; -1 = default: 8)
    int tc_chain = 1;      // MultUpdate(:mse), single GPU, k <= 128: 1 (default) = the iteration is one chain of programmatic dependents and the
                           // hand-over between update launches is a per-tile completion counter, not a kernel boundary (tc_update.cuh);
                           // the reduce kernels run as one CTA per SM walking their virtual blocks; > 1 = that many CTAs; 0 = off
    int tc_skew = 0;       // on top of tc_chain: two groups of tiles half a period apart (one streams while the other is in its epilogues)
    int sm_count = 148;
    int tc_prefetch_next = 0;  // update kernel (single GPU, bf16 mode): k-blocks of the NEXT launch's X panel each CTA prefetches into L2
                               // while it sits in its epilogue (0 = off)
    int tc_precision = 0;  // 0 = bf16 operands; 1 = bf16x3 (hi/lo split of X and of the streamed factor: fp32-class products)
    nmfb200::Xchg xchg;
    int tc_pdl = 1;        // 1 = launch the update kernels as programmatic dependents of the reduce kernel before them
    int tc_div_fused = 1;  // MultUpdate(:div): 1 = quotient tile stays on chip (div_fused_kernel), 0 = bf16 Q panel through HBM
    std::vector<cudaEvent_t> ev_pool;  // events for time_kernels
    size_t ev_used = 0;
    nmfb200_trace_fn trace = nullptr;
    void* trace_user = nullptr;

    // X (column-major p x n as given; device pointer, owned or adopted)
    int x_elt = 0;  // 0 none, 4 f32, 8 f64
    int64_t p = 0, n = 0, ldx = 0;
    const void* dX = nullptr;
    bool x_owned = false;
    uint64_t x_epoch = 0;     // bumped on every set_X; engines key their derived caches on it
    struct XCacheKey {        // what a pair of bf16 X caches was built for (per buffer-name prefix: whole X, or a row shard)
        uint64_t epoch = 0;
        const void* X = nullptr;
        int64_t p = 0;
        int trH = 0, trW = 0;
        int with_lo = 0;  // precision mode bf16x3: the remainder caches were built too
        bool covers(const XCacheKey& want) const {   // a cache built with the remainders also serves a request without them
            return epoch == want.epoch && X == want.X && p == want.p && trH == want.trH && trW == want.trW && with_lo >= want.with_lo;
        }
    };
    std::map<std::string, XCacheKey> tc_x_cache;
    double x_norm2 = 0;            // ||X||^2 of the resident X (verbose trace-identity objective)
    uint64_t x_norm2_epoch = ~0ull;
    int tc_trace_identity = 1;     // verbose on the tensor-core engine: per-iteration objective by the trace identity (0 = a pass over X)

    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;

    // named device buffers, kept across solves
    std::map<std::string, nmfb200::DevBuf> bufs;
    int64_t launches = 0;  // kernels launched by the current call

    void* buf(const std::string& name, size_t bytes) {
        nmfb200::DevBuf& b = bufs[name];
        if (b.bytes < bytes) {
            if (b.ptr) NMF_CUDA(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.bytes = 0;
            NMF_CUDA(cudaMalloc(&b.ptr, bytes ? bytes : 16));
            b.bytes = bytes ? bytes : 16;
        }
        return b.ptr;
    }
    template <typename T> T* buf_t(const std::string& name, size_t count) { return (T*)buf(name, count * sizeof(T)); }
    void drop(const std::string& name) {
        auto it = bufs.find(name);
        if (it != bufs.end()) {
            if (it->second.ptr) cudaFree(it->second.ptr);
            bufs.erase(it);
        }
    }
    // phase timeline (option time_kernels = 2): mark(label) records an event; the time since the previous mark is
    // charged to `label`.  Printed to stderr at the end of the solve (diagnostics for multi-GPU runs).
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    void mark(const char* label) {
        if (time_kernels != 2) return;
        cudaEvent_t e;
        NMF_CUDA(cudaEventCreate(&e));
        NMF_CUDA(cudaEventRecord(e, stream));
        marks.emplace_back(label, e);
    }
    void report_marks(int64_t iters) {
        if (marks.empty()) return;
        std::map<std::string, std::pair<double, int>> agg;
        std::vector<std::string> order;
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            auto& a = agg[marks[i].first];
            if (a.second == 0) order.push_back(marks[i].first);
            a.first += ms;
            a.second += 1;
        }
        fprintf(stderr, "[nmfb200 rank %d] phase timeline over %lld iterations (us per iteration):", rank, (long long)iters);
        for (auto& name : order) fprintf(stderr, "  %s=%.1f", name.c_str(), agg[name].first * 1e3 / (double)std::max<int64_t>(iters, 1));
        fprintf(stderr, "\n");
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            NMF_CUDA(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    }
    // sum of (ev[2i+1] - ev[2i]) over the events handed out since ev_used was reset; stream must be idle
    double drain_event_pairs(int64_t* npairs) {
        double ms = 0;
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, ev_pool[i], ev_pool[i + 1]);
            ms += t;
        }
        *npairs = (int64_t)(ev_used / 2);
        ev_used = 0;
        return ms;
    }
    void free_all() {
        for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
        ev_pool.clear();
        ev_used = 0;
        for (auto& kv : bufs)
            if (kv.second.ptr) cudaFree(kv.second.ptr);
        bufs.clear();
    }

    template <typename T> void allreduce_sum(T* ptr, size_t count) {
        if (!comm) return;
        NMF_NCCL(nmfb200::NcclApi::get().AllReduce(ptr, ptr, count, nmfb200::NcclType<T>::v, ncclSum, comm, stream));
    }
    // Collective AND of a per-rank predicate (host-synchronising; once per solve).  Decisions that pick a code path with
    // its own collectives (engine, tensor-core objective vs exact objective) must be the same on every rank, whatever the
    // local shard's row count / alignment says.
    bool all_ranks(bool mine) {
        if (!comm) return mine;
        int* d = (int*)buf("comm.agree", sizeof(int));
        int v = mine ? 1 : 0;
        NMF_CUDA(cudaMemcpyAsync(d, &v, sizeof(int), cudaMemcpyHostToDevice, stream));
        NMF_NCCL(nmfb200::NcclApi::get().AllReduce(d, d, 1, ncclInt32, ncclMin, comm, stream));
        NMF_CUDA(cudaMemcpyAsync(&v, d, sizeof(int), cudaMemcpyDeviceToHost, stream));
        NMF_CUDA(cudaStreamSynchronize(stream));
        return v != 0;
    }
    template <typename T> void allreduce_max(T* ptr, size_t count) {
        if (!comm) return;
        NMF_NCCL(nmfb200::NcclApi::get().AllReduce(ptr, ptr, count, nmfb200::NcclType<T>::v, ncclMax, comm, stream));
    }
};

namespace nmfb200 {

struct SolveArgs {
    int alg;  // 0 multmse, 1 multdiv, 2 greedycd, 3 projals, 4 cd, 5 alspgrad
    int64_t k;
    int64_t maxiter;
    double tol, lambda_w, lambda_h;
    int update_H, verbose, on_device;
    // CoordinateDescent (coorddesc.jl:24-46)
    double cd_alpha = 0, cd_l1ratio = 0;
    int cd_regularization = 0;  // 0 :both, 1 :components, 2 :transformation, 3 :none
    int cd_shuffle = 0;
    uint64_t cd_seed = 0;
    // ALSPGrad (alspgrad.jl:352-373)
    int64_t maxsubiter = 200;
    double tolg = 0;
};

// engines (one translation unit each)
template <typename T>
void simt_solve(nmfb200_handle* h, const SolveArgs& a, T* W, int64_t ldw, T* H, int64_t ldh, nmfb200_result* out);
double simt_objective_f32(nmfb200_handle* h, int alg, const float* W, int64_t ldw, const float* H, int64_t ldh, int64_t k,
                          double lambda_w, double lambda_h);
template <typename T>
void simt_mul_X(nmfb200_handle* h, int transpose_X, const T* B, int64_t ldb, int64_t c, T* C, int64_t ldc);
// NNDSVD on the device (csrc/init_device.cuh): rsvd(X, k) of initialization.jl:78 and _nndsvd! (:26-68) on the resident X
template <typename T>
void simt_rsvd(nmfb200_handle* h, int64_t k, uint64_t seed, T* U, int64_t ldu, T* S, T* V, int64_t ldv);
template <typename T>
void simt_nndsvd(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int variant, int zeroh, uint64_t seed, int on_device);
bool tc_supported(const nmfb200_handle* h, const SolveArgs& a);
// out(r, a) = sum_c Xs(r, c) * O(c, a) on the tensor cores with split (bf16 hi + lo) operands, for the algorithms whose remaining
// arithmetic stays on the exact engine (ProjectedALS, CoordinateDescent, ALSPGrad).  side 0: Xs = X' (rows = columns of X,
// O is p x k); side 1: Xs = X (O is n x k).  O(c, a) = O[c*sOr + a*sOc], out(r, a) = out[r*sNr + a*sNc]; device pointers.
// Returns false (nothing done) when the shape is not covered; the caller then uses its own GEMM.
bool tc_xmul(nmfb200_handle* h, int side, const float* O, int64_t sOr, int64_t sOc, int64_t k, float* out, int64_t sNr, int64_t sNc);
void tc_solve(nmfb200_handle* h, const SolveArgs& a, float* W, int64_t ldw, float* H, int64_t ldh, nmfb200_result* out);
// batched replicates: nrep MultUpdate(:mse) solves stacked along the component axis (W p x nrep*k, H nrep*k x n), one pass over X for all
bool tc_batched_supported(const nmfb200_handle* h, const SolveArgs& a, int nrep);
void tc_solve_batched(nmfb200_handle* h, const SolveArgs& a, int nrep, float* W, int64_t ldw, float* H, int64_t ldh, nmfb200_result* out);
void tc_release(nmfb200_handle* h);
void tc_shard_geometry(int64_t n, int ranks, int rank, int64_t* own_row0, int64_t* own_row1, int64_t* tile_rows);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace nmfb200
