// api.cu -- the extern "C" surface declared in include/nmfb200.h.  Validation mirrors the reference
// constructors (multupd.jl:27-31, greedycd.jl:25-28) and nmf_checksize (common.jl:5-16); every C++
// exception is converted to a status code + message at this boundary.
#include "common.cuh"
#include "philox.cuh"

using namespace nmfb200;

namespace {

template <typename T>
__global__ void count_not_nonneg_kernel(const T* __restrict__ X, int64_t p, int64_t n, int64_t ldx, unsigned long long* out) {
    unsigned long long c = 0;
    int64_t len = p * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i % p, col = i / p;
        if (!(X[r + col * ldx] >= T(0))) ++c;  // interf.jl:15 `all(t -> t >= zero(T), X)`: NaN fails too
    }
    c = __reduce_add_sync(0xffffffffu, (unsigned)c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// A(i, j) = U(e), e = e0 + i + j * e_ld  (rows x cols, column-major with leading dimension ld)
template <typename T>
__global__ void philox_fill_kernel(T* __restrict__ A, int64_t rows, int64_t cols, int64_t ld, uint64_t e0, uint64_t e_ld, uint32_t stream,
                                   uint64_t seed) {
    const int64_t total = rows * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t % rows, j = t / rows;
        A[i + j * ld] = philox_uniform<T>(e0 + (uint64_t)i + (uint64_t)j * e_ld, stream, seed);
    }
}
// column sums in T, sequentially per column like `sum(view(W, :, j))` would not be (Julia sums pairwise): one block per column,
// fixed tree => deterministic; the result is only a scale factor
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ A, int64_t rows, int64_t ld, T* __restrict__ out) {
    __shared__ T red[256];
    const T* col = A + (int64_t)blockIdx.x * ld;
    T s = T(0);
    for (int64_t i = threadIdx.x; i < rows; i += 256) s += col[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}
template <typename T>
__global__ void colscale_kernel(T* __restrict__ A, int64_t rows, int64_t cols, int64_t ld, const T* __restrict__ sums) {
    const int64_t total = rows * cols;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t % rows, j = t / rows;
        A[i + j * ld] *= T(1) / sums[j];   // `W[:, j] .*= 1 / sum(W[:, j])` (utils.jl:28-32)
    }
}

template <typename T>
void randinit_impl(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, uint64_t seed, int64_t row_offset, int64_t p_total,
                   int normalize, int zeroh, int on_device) {
    NMF_REQUIRE(W != nullptr && H != nullptr, NMFB200_EINVAL, "NULL argument");
    NMF_REQUIRE(h->x_elt != 0, NMFB200_ESTATE, "nmfb200_set_X must precede randinit (it defines p and n)");
    const int64_t p = h->p, n = h->n;
    NMF_REQUIRE(k >= 1 && ldw >= p && ldh >= k && row_offset >= 0 && p_total >= row_offset + p, NMFB200_EDIM, "inconsistent dimensions");
    cudaStream_t st = h->stream;
    T* dW = on_device ? W : h->buf_t<T>("init.W", (size_t)p * k);
    T* dH = on_device ? H : h->buf_t<T>("init.H", (size_t)k * n);
    const int64_t lw = on_device ? ldw : p, lh = on_device ? ldh : k;
    const int grid = 148 * 8;
    philox_fill_kernel<T><<<grid, 256, 0, st>>>(dW, p, k, lw, (uint64_t)row_offset, (uint64_t)p_total, 0u, seed);
    h->launches += 1;
    if (normalize) {
        T* sums = h->buf_t<T>("init.colsum", (size_t)k);
        colsum_kernel<T><<<(unsigned)k, 256, 0, st>>>(dW, p, lw, sums);
        h->allreduce_sum(sums, (size_t)k);   // row-sharded: the column sum runs over all ranks' rows
        colscale_kernel<T><<<grid, 256, 0, st>>>(dW, p, k, lw, sums);
        h->launches += 2;
    }
    if (zeroh) {
        NMF_CUDA(cudaMemset2DAsync(dH, lh * sizeof(T), 0, k * sizeof(T), n, st));
    } else {
        philox_fill_kernel<T><<<grid, 256, 0, st>>>(dH, k, n, lh, 0, (uint64_t)k, 1u, seed);
        h->launches += 1;
    }
    NMF_CUDA(cudaGetLastError());
    if (!on_device) {
        NMF_CUDA(cudaMemcpy2DAsync(W, ldw * sizeof(T), dW, p * sizeof(T), p * sizeof(T), k, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(H, ldh * sizeof(T), dH, k * sizeof(T), k * sizeof(T), n, cudaMemcpyDeviceToHost, st));
    }
    NMF_CUDA(cudaStreamSynchronize(st));
}

template <typename F>
int guarded(nmfb200_handle* h, F&& f) {
    if (!h) return NMFB200_EINVAL;
    try {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != h->device) NMF_CUDA(cudaSetDevice(h->device));
        h->launches = 0;
        f();
        h->err.clear();
        return NMFB200_OK;
    } catch (const Error& e) {
        h->err = e.msg;
        cudaGetLastError();
        return e.status;
    } catch (const std::exception& e) {
        h->err = e.what();
        return NMFB200_ECUDA;
    } catch (...) {
        h->err = "unknown exception";
        return NMFB200_ECUDA;
    }
}

template <typename T>
void set_X_impl(nmfb200_handle* h, const T* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg, bool on_device) {
    NMF_REQUIRE(X != nullptr, NMFB200_EINVAL, "X is NULL");
    NMF_REQUIRE(p > 0 && n > 0 && ldx >= p, NMFB200_EDIM, "invalid dimensions for X");
    h->x_owned = false;   // an owned buffer ("X") is kept and reused by buf() when the next matrix fits
    h->dX = nullptr;
    h->x_elt = 0;
    const T* dX = X;
    int64_t dld = ldx;
    if (!on_device) {
        T* buf = h->buf_t<T>("X", (size_t)p * n);
        NMF_CUDA(cudaMemcpy2DAsync(buf, p * sizeof(T), X, ldx * sizeof(T), p * sizeof(T), n, cudaMemcpyHostToDevice, h->stream));
        dX = buf;
        dld = p;
        h->x_owned = true;
    }
    if (check_nonneg) {
        unsigned long long* cnt = (unsigned long long*)h->buf("X.nonneg_count", 16);
        NMF_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), h->stream));
        count_not_nonneg_kernel<T><<<148 * 8, 256, 0, h->stream>>>(dX, p, n, dld, cnt);
        h->launches += 1;
        unsigned long long c = 0;
        NMF_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
        NMF_CUDA(cudaStreamSynchronize(h->stream));
        NMF_REQUIRE(c == 0, NMFB200_EINVAL, "The elements of X must be non-negative.");
    }
    NMF_CUDA(cudaStreamSynchronize(h->stream));
    h->dX = dX;
    h->ldx = dld;
    h->p = p;
    h->n = n;
    h->x_elt = (int)sizeof(T);
    h->x_epoch += 1;
}

// ---- sparse X (README.md:22 "Sparse NMF": the reference's solvers only touch X through mul!, so a SparseMatrixCSC works there) ----
// The compressed columns cross PCIe (12 or 16 bytes per stored entry instead of 4 or 8 per cell) and are expanded ON THE DEVICE into
// the dense column-major matrix every engine works on -- the tensor-core path streams dense bf16 tiles whatever the sparsity, and
// p * n * sizeof(T) has to fit in HBM.  Duplicate (row, column) entries are summed, as `sparse(I, J, V)` / scipy's toarray() do.
template <typename T>
__global__ void csc_scatter_kernel(const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval, const T* __restrict__ nzval,
                                   int64_t p, int64_t n, int64_t base, T* __restrict__ X, unsigned long long* __restrict__ bad) {
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        const int64_t e0 = colptr[col] - base, e1 = colptr[col + 1] - base;
        for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const int64_t r = rowval[e] - base;
            if (r < 0 || r >= p) { atomicAdd(bad, 1ull); continue; }
            atomicAdd(X + r + col * p, nzval[e]);
        }
    }
}

template <typename T>
void set_X_csc_impl(nmfb200_handle* h, const int64_t* colptr, const int64_t* rowval, const T* nzval, int64_t p, int64_t n, int index_base,
                    int check_nonneg) {
    NMF_REQUIRE(colptr != nullptr, NMFB200_EINVAL, "colptr is NULL");
    NMF_REQUIRE(p > 0 && n > 0, NMFB200_EDIM, "invalid dimensions for X");
    NMF_REQUIRE(index_base == 0 || index_base == 1, NMFB200_EINVAL, "index_base must be 0 or 1");
    const int64_t base = index_base;
    NMF_REQUIRE(colptr[0] == base, NMFB200_EINVAL, "colptr[0] must equal index_base");
    for (int64_t j = 0; j < n; ++j) NMF_REQUIRE(colptr[j + 1] >= colptr[j], NMFB200_EINVAL, "colptr must be non-decreasing");
    const int64_t nnz = colptr[n] - base;
    NMF_REQUIRE(nnz == 0 || (rowval != nullptr && nzval != nullptr), NMFB200_EINVAL, "rowval / nzval is NULL");
    cudaStream_t st = h->stream;
    h->x_owned = false;
    h->dX = nullptr;
    h->x_elt = 0;
    T* X = h->buf_t<T>("X", (size_t)p * n);
    int64_t* d_colptr = h->buf_t<int64_t>("X.csc_colptr", (size_t)n + 1);
    int64_t* d_rowval = h->buf_t<int64_t>("X.csc_rowval", (size_t)std::max<int64_t>(nnz, 1));
    T* d_nzval = h->buf_t<T>("X.csc_nzval", (size_t)std::max<int64_t>(nnz, 1));
    unsigned long long* cnt = (unsigned long long*)h->buf("X.nonneg_count", 16);
    NMF_CUDA(cudaMemsetAsync(X, 0, (size_t)p * n * sizeof(T), st));
    NMF_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), st));
    NMF_CUDA(cudaMemcpyAsync(d_colptr, colptr, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        NMF_CUDA(cudaMemcpyAsync(d_rowval, rowval, (size_t)nnz * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaMemcpyAsync(d_nzval, nzval, (size_t)nnz * sizeof(T), cudaMemcpyHostToDevice, st));
        csc_scatter_kernel<T><<<(unsigned)std::min<int64_t>(n, 148 * 16), 128, 0, st>>>(d_colptr, d_rowval, d_nzval, p, n, base, X, cnt + 1);
        h->launches += 1;
        if (check_nonneg) {   // interf.jl:15 on the stored entries (the implicit zeros pass)
            count_not_nonneg_kernel<T><<<148 * 8, 256, 0, st>>>(d_nzval, nnz, 1, nnz, cnt);
            h->launches += 1;
        }
    }
    unsigned long long c[2] = {0, 0};
    NMF_CUDA(cudaMemcpyAsync(c, cnt, sizeof(c), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    NMF_REQUIRE(c[1] == 0, NMFB200_EINVAL, "rowval holds a row index outside 1..p (index_base .. index_base + p - 1)");
    NMF_REQUIRE(c[0] == 0, NMFB200_EINVAL, "The elements of X must be non-negative.");
    h->dX = X;
    h->ldx = p;
    h->p = p;
    h->n = n;
    h->x_elt = (int)sizeof(T);
    h->x_owned = true;
    h->x_epoch += 1;
}

template <typename T>
void solve_args(nmfb200_handle* h, const SolveArgs& a, T* W, int64_t ldw, T* H, int64_t ldh, nmfb200_result* out) {
    NMF_REQUIRE(out != nullptr && W != nullptr && H != nullptr, NMFB200_EINVAL, "NULL argument");
    std::memset(out, 0, sizeof(*out));
    NMF_REQUIRE(h->x_elt != 0, NMFB200_ESTATE, "nmfb200_set_X must precede solve");
    NMF_REQUIRE(h->x_elt == (int)sizeof(T), NMFB200_ESTATE, "X was set with a different element type");
    if (a.alg <= 2) {  // the MultUpdate / GreedyCD constructors validate; the other three do not
        NMF_REQUIRE(a.maxiter > 1, NMFB200_EINVAL, "maxiter must be greater than 1.");   // multupd.jl:28, greedycd.jl:25
        NMF_REQUIRE(a.tol > 0, NMFB200_EINVAL, "tol must be positive.");                 // multupd.jl:29, greedycd.jl:26
        NMF_REQUIRE(a.lambda_w >= 0, NMFB200_EINVAL, "lambda_w must be non-negative.");  // multupd.jl:30, greedycd.jl:27
        NMF_REQUIRE(a.lambda_h >= 0, NMFB200_EINVAL, "lambda_h must be non-negative.");  // multupd.jl:31, greedycd.jl:28
    }
    NMF_REQUIRE(a.alg != 4 || (a.cd_regularization >= 0 && a.cd_regularization <= 3), NMFB200_EINVAL, "regularization must be 0..3");
    NMF_REQUIRE(a.k >= 1 && ldw >= h->p && ldh >= a.k, NMFB200_EDIM, "Dimensions of X, W, and H are inconsistent.");  // common.jl:12-14
    bool use_tc = false;
    if (sizeof(T) == 4 && h->engine_opt != 1) {
        use_tc = h->all_ranks(tc_supported(h, a));   // multi-GPU: one decision for all ranks (shards may differ in size / alignment)
        NMF_REQUIRE(use_tc || h->engine_opt != 2, NMFB200_ENOTSUP, "engine=tc requested but this problem is not covered by the tensor-core engine");
    } else {
        NMF_REQUIRE(h->engine_opt != 2, NMFB200_ENOTSUP, "engine=tc supports Float32 only");
    }
    if (use_tc) tc_solve(h, a, (float*)W, ldw, (float*)H, ldh, out);
    else simt_solve<T>(h, a, W, ldw, H, ldh, out);
}

template <typename T>
void solve_impl(nmfb200_handle* h, int alg, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int64_t maxiter, double tol,
                double lambda_w, double lambda_h, int update_H, int verbose, int on_device, nmfb200_result* out) {
    SolveArgs a{alg, k, maxiter, tol, lambda_w, lambda_h, update_H, verbose, on_device};
    solve_args<T>(h, a, W, ldw, H, ldh, out);
}

}  // namespace

extern "C" {

int nmfb200_version(void) { return NMFB200_VERSION; }

const char* nmfb200_status_string(int status) {
    switch (status) {
        case NMFB200_OK: return "ok";
        case NMFB200_EINVAL: return "invalid argument";
        case NMFB200_EDIM: return "dimension mismatch";
        case NMFB200_ECUDA: return "CUDA error";
        case NMFB200_ENCCL: return "NCCL error";
        case NMFB200_ENOMEM: return "out of device memory";
        case NMFB200_ESTATE: return "invalid call order";
        case NMFB200_ENOTSUP: return "not supported";
        case NMFB200_ENUMERIC: return "numerical breakdown";
        default: return "unknown status";
    }
}

int nmfb200_create(nmfb200_handle** out, int device, int flags) {
    (void)flags;
    if (!out) return NMFB200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return NMFB200_ECUDA; }
    if (device < 0 || device >= ndev) return NMFB200_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NMFB200_ECUDA;
    if (prop.major != 10) return NMFB200_ENOTSUP;  // sm_100a only: no other code path exists in this library
    if (cudaSetDevice(device) != cudaSuccess) return NMFB200_ECUDA;
    nmfb200_handle* h = new (std::nothrow) nmfb200_handle();
    if (!h) return NMFB200_ENOMEM;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return NMFB200_ECUDA; }
    h->stream = h->own_stream;
    *out = h;
    return NMFB200_OK;
}

int nmfb200_destroy(nmfb200_handle* h) {
    if (!h) return NMFB200_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    try { tc_release(h); } catch (...) {}
    if (h->comm) { try { NcclApi::get().CommDestroy(h->comm); } catch (...) {} }
    h->free_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return NMFB200_OK;
}

const char* nmfb200_last_error(const nmfb200_handle* h) { return h ? h->err.c_str() : "NULL handle"; }

int nmfb200_set_stream(nmfb200_handle* h, void* stream) {
    return guarded(h, [&] {
        NMF_CUDA(cudaStreamSynchronize(h->stream));
        h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    });
}

int nmfb200_set_option(nmfb200_handle* h, const char* key, const char* value) {
    return guarded(h, [&] {
        NMF_REQUIRE(key && value, NMFB200_EINVAL, "NULL option");
        std::string k(key), v(value);
        if (k == "engine") {
            if (v == "auto") h->engine_opt = 0;
            else if (v == "simt") h->engine_opt = 1;
            else if (v == "tc") h->engine_opt = 2;
            else throw Error{NMFB200_EINVAL, "engine must be auto|simt|tc"};
        } else if (k == "check_every") {
            int c = atoi(value);
            NMF_REQUIRE(c >= 1, NMFB200_EINVAL, "check_every must be >= 1");
            h->check_every = c;
        } else if (k == "tc_tile_rows") {
            h->tc_tile_rows = atoi(value);
        } else if (k == "tc_pdl") {
            h->tc_pdl = atoi(value);
        } else if (k == "tc_div_fused") {
            h->tc_div_fused = atoi(value);
        } else if (k == "tc_xchg") {
            if (v == "p2p") h->tc_xchg = 1;
            else if (v == "nccl") h->tc_xchg = 0;
            else throw Error{NMFB200_EINVAL, "tc_xchg must be p2p|nccl"};
        } else if (k == "tc_xmul") {
            h->tc_xmul_opt = atoi(value);
        } else if (k == "tc_skew") {
            h->tc_skew = atoi(value);
        } else if (k == "tc_chain") {
            h->tc_chain = atoi(value);
        } else if (k == "tc_prefetch_next") {
            int c = atoi(value);
            NMF_REQUIRE(c >= 0 && c <= 4096, NMFB200_EINVAL, "tc_prefetch_next must be 0..4096");
            h->tc_prefetch_next = c;
        } else if (k == "tc_flush") {
            h->tc_flush = atoi(value);
        } else if (k == "tc_fused_hstep") {
            h->tc_fused_hstep = atoi(value);
        } else if (k == "tc_trace_identity") {
            h->tc_trace_identity = atoi(value);
        } else if (k == "tc_defer_signal") {
            h->tc_defer_signal = atoi(value);
        } else if (k == "tc_side_stream") {
            h->tc_side_stream = atoi(value);
        } else if (k == "emulate_shards") {
            int g = atoi(value);
            NMF_REQUIRE(g >= 0 && g <= XCHG_MAX_RANKS, NMFB200_EINVAL, "emulate_shards must be 0..8");
            h->emulate_shards = g;
        } else if (k == "precision") {
            if (v == "bf16") h->tc_precision = 0;
            else if (v == "bf16x3") h->tc_precision = 1;
            else throw Error{NMFB200_EINVAL, "precision must be bf16|bf16x3"};
        } else if (k == "tc_debug") {
            h->tc_debug = atoi(value);
        } else if (k == "time_kernels") {
            h->time_kernels = atoi(value);
        } else {
            throw Error{NMFB200_EINVAL, "unknown option " + k};
        }
    });
}

int nmfb200_set_trace(nmfb200_handle* h, nmfb200_trace_fn fn, void* user) {
    return guarded(h, [&] { h->trace = fn; h->trace_user = user; });
}

int nmfb200_set_X_f32(nmfb200_handle* h, const float* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg) {
    return guarded(h, [&] { set_X_impl<float>(h, X, p, n, ldx, check_nonneg, false); });
}
int nmfb200_set_X_f64(nmfb200_handle* h, const double* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg) {
    return guarded(h, [&] { set_X_impl<double>(h, X, p, n, ldx, check_nonneg, false); });
}
int nmfb200_set_X_dev_f32(nmfb200_handle* h, const float* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg) {
    return guarded(h, [&] { set_X_impl<float>(h, X, p, n, ldx, check_nonneg, true); });
}
int nmfb200_set_X_dev_f64(nmfb200_handle* h, const double* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg) {
    return guarded(h, [&] { set_X_impl<double>(h, X, p, n, ldx, check_nonneg, true); });
}

int nmfb200_set_X_csc_f32(nmfb200_handle* h, const int64_t* colptr, const int64_t* rowval, const float* nzval, int64_t p, int64_t n,
                          int index_base, int check_nonneg) {
    return guarded(h, [&] { set_X_csc_impl<float>(h, colptr, rowval, nzval, p, n, index_base, check_nonneg); });
}
int nmfb200_set_X_csc_f64(nmfb200_handle* h, const int64_t* colptr, const int64_t* rowval, const double* nzval, int64_t p, int64_t n,
                          int index_base, int check_nonneg) {
    return guarded(h, [&] { set_X_csc_impl<double>(h, colptr, rowval, nzval, p, n, index_base, check_nonneg); });
}

#define NMFB200_DEFINE_SOLVE(NAME, ALG, T)                                                                                   \
    int NAME(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int64_t maxiter, T tol, T lambda_w,         \
             T lambda_h, int update_H, int verbose, int on_device, nmfb200_result* out) {                                    \
        return guarded(h, [&] {                                                                                              \
            solve_impl<T>(h, ALG, W, ldw, H, ldh, k, maxiter, (double)tol, (double)lambda_w, (double)lambda_h, update_H,     \
                          verbose, on_device, out);                                                                          \
        });                                                                                                                  \
    }

NMFB200_DEFINE_SOLVE(nmfb200_solve_multmse_f32, 0, float)
NMFB200_DEFINE_SOLVE(nmfb200_solve_multmse_f64, 0, double)
NMFB200_DEFINE_SOLVE(nmfb200_solve_multdiv_f32, 1, float)
NMFB200_DEFINE_SOLVE(nmfb200_solve_multdiv_f64, 1, double)
NMFB200_DEFINE_SOLVE(nmfb200_solve_greedycd_f32, 2, float)
NMFB200_DEFINE_SOLVE(nmfb200_solve_greedycd_f64, 2, double)
NMFB200_DEFINE_SOLVE(nmfb200_solve_projals_f32, 3, float)
NMFB200_DEFINE_SOLVE(nmfb200_solve_projals_f64, 3, double)

int nmfb200_solve_multmse_batched_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k, int32_t replicates,
                                      int64_t maxiter, float tol, float lambda_w, float lambda_h, int update_H, int on_device,
                                      nmfb200_result* out) {
    return guarded(h, [&] {
        NMF_REQUIRE(out != nullptr && W != nullptr && H != nullptr, NMFB200_EINVAL, "NULL argument");
        NMF_REQUIRE(h->x_elt != 0, NMFB200_ESTATE, "nmfb200_set_X must precede solve");
        NMF_REQUIRE(h->x_elt == 4, NMFB200_ESTATE, "X was set with a different element type");
        NMF_REQUIRE(replicates >= 1, NMFB200_EINVAL, "The value of replicates must be positive.");   // interf.jl:20
        NMF_REQUIRE(maxiter > 1, NMFB200_EINVAL, "maxiter must be greater than 1.");               // multupd.jl:28
        NMF_REQUIRE(tol > 0, NMFB200_EINVAL, "tol must be positive.");
        NMF_REQUIRE(lambda_w >= 0, NMFB200_EINVAL, "lambda_w must be non-negative.");
        NMF_REQUIRE(lambda_h >= 0, NMFB200_EINVAL, "lambda_h must be non-negative.");
        NMF_REQUIRE(k >= 1 && ldw >= h->p && ldh >= k * replicates, NMFB200_EDIM, "Dimensions of X, W, and H are inconsistent.");
        SolveArgs a{0, k, maxiter, (double)tol, (double)lambda_w, (double)lambda_h, update_H, 0, on_device};
        NMF_REQUIRE(tc_batched_supported(h, a, replicates), NMFB200_ENOTSUP,
                    "batched replicates need the tensor-core engine on one GPU: Float32, replicates * k <= 256, replicates <= 32, "
                    "at least 2^20 cells unless engine=tc");
        tc_solve_batched(h, a, replicates, W, ldw, H, ldh, out);
    });
}

#define NMFB200_DEFINE_SOLVE_CD(NAME, T)                                                                                     \
    int NAME(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int64_t maxiter, T tol, T alpha, T l1ratio, \
             int regularization, int shuffle, uint64_t seed, int update_H, int verbose, int on_device, nmfb200_result* out) { \
        return guarded(h, [&] {                                                                                              \
            SolveArgs a{4, k, maxiter, (double)tol, 0.0, 0.0, update_H, verbose, on_device};                                 \
            a.cd_alpha = (double)alpha;                                                                                      \
            a.cd_l1ratio = (double)l1ratio;                                                                                  \
            a.cd_regularization = regularization;                                                                            \
            a.cd_shuffle = shuffle;                                                                                          \
            a.cd_seed = seed;                                                                                                \
            solve_args<T>(h, a, W, ldw, H, ldh, out);                                                                        \
        });                                                                                                                  \
    }
NMFB200_DEFINE_SOLVE_CD(nmfb200_solve_cd_f32, float)
NMFB200_DEFINE_SOLVE_CD(nmfb200_solve_cd_f64, double)

#define NMFB200_DEFINE_SOLVE_ALSPGRAD(NAME, T)                                                                               \
    int NAME(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int64_t maxiter, int64_t maxsubiter, T tol, \
             T tolg, int update_H, int verbose, int on_device, nmfb200_result* out) {                                        \
        return guarded(h, [&] {                                                                                              \
            SolveArgs a{5, k, maxiter, (double)tol, 0.0, 0.0, update_H, verbose, on_device};                                 \
            a.maxsubiter = maxsubiter;                                                                                       \
            a.tolg = (double)tolg;                                                                                           \
            solve_args<T>(h, a, W, ldw, H, ldh, out);                                                                        \
        });                                                                                                                  \
    }
NMFB200_DEFINE_SOLVE_ALSPGRAD(nmfb200_solve_alspgrad_f32, float)
NMFB200_DEFINE_SOLVE_ALSPGRAD(nmfb200_solve_alspgrad_f64, double)

int nmfb200_mul_X_f32(nmfb200_handle* h, int transpose_X, const float* B, int64_t ldb, int64_t c, float* C, int64_t ldc) {
    return guarded(h, [&] {
        NMF_REQUIRE(h->x_elt == 4, NMFB200_ESTATE, h->x_elt ? "X was set with a different element type" : "nmfb200_set_X must precede mul_X");
        simt_mul_X<float>(h, transpose_X, B, ldb, c, C, ldc);
    });
}
int nmfb200_mul_X_f64(nmfb200_handle* h, int transpose_X, const double* B, int64_t ldb, int64_t c, double* C, int64_t ldc) {
    return guarded(h, [&] {
        NMF_REQUIRE(h->x_elt == 8, NMFB200_ESTATE, h->x_elt ? "X was set with a different element type" : "nmfb200_set_X must precede mul_X");
        simt_mul_X<double>(h, transpose_X, B, ldb, c, C, ldc);
    });
}

int nmfb200_randinit_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k, uint64_t seed, int64_t row_offset,
                         int64_t p_total, int normalize, int zeroh, int on_device) {
    return guarded(h, [&] { randinit_impl<float>(h, W, ldw, H, ldh, k, seed, row_offset, p_total, normalize, zeroh, on_device); });
}
int nmfb200_randinit_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k, uint64_t seed, int64_t row_offset,
                         int64_t p_total, int normalize, int zeroh, int on_device) {
    return guarded(h, [&] { randinit_impl<double>(h, W, ldw, H, ldh, k, seed, row_offset, p_total, normalize, zeroh, on_device); });
}

int nmfb200_rsvd_f32(nmfb200_handle* h, int64_t k, uint64_t seed, float* U, int64_t ldu, float* S, float* V, int64_t ldv) {
    return guarded(h, [&] { simt_rsvd<float>(h, k, seed, U, ldu, S, V, ldv); });
}
int nmfb200_rsvd_f64(nmfb200_handle* h, int64_t k, uint64_t seed, double* U, int64_t ldu, double* S, double* V, int64_t ldv) {
    return guarded(h, [&] { simt_rsvd<double>(h, k, seed, U, ldu, S, V, ldv); });
}
int nmfb200_nndsvd_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k, int variant, int zeroh, uint64_t seed,
                       int on_device) {
    return guarded(h, [&] { simt_nndsvd<float>(h, W, ldw, H, ldh, k, variant, zeroh, seed, on_device); });
}
int nmfb200_nndsvd_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k, int variant, int zeroh, uint64_t seed,
                       int on_device) {
    return guarded(h, [&] { simt_nndsvd<double>(h, W, ldw, H, ldh, k, variant, zeroh, seed, on_device); });
}

int nmfb200_comm_unique_id(void* out_id_128) {
    if (!out_id_128) return NMFB200_EINVAL;
    try {
        static_assert(sizeof(ncclUniqueId) == NMFB200_UNIQUE_ID_BYTES, "ncclUniqueId size");
        ncclUniqueId id;
        if (NcclApi::get().GetUniqueId(&id) != ncclSuccess) return NMFB200_ENCCL;
        std::memcpy(out_id_128, &id, sizeof(id));
        return NMFB200_OK;
    } catch (...) {
        return NMFB200_ENCCL;
    }
}

int nmfb200_shard_geometry(int64_t n, int ranks, int rank, int64_t* own_row0, int64_t* own_row1, int64_t* tile_rows) {
    if (!own_row0 || !own_row1 || !tile_rows || n < 1 || ranks < 1 || ranks > XCHG_MAX_RANKS || rank < 0 || rank >= ranks) return NMFB200_EINVAL;
    try {
        tc_shard_geometry(n, ranks, rank, own_row0, own_row1, tile_rows);
        return NMFB200_OK;
    } catch (...) {
        return NMFB200_EINVAL;
    }
}

int nmfb200_comm_init(nmfb200_handle* h, int rank, int nranks, const void* id_128) {
    return guarded(h, [&] {
        NMF_REQUIRE(id_128 && nranks >= 1 && rank >= 0 && rank < nranks, NMFB200_EINVAL, "invalid rank/nranks/id");
        if (h->comm) { NMF_NCCL(NcclApi::get().CommDestroy(h->comm)); h->comm = nullptr; }
        ncclUniqueId id;
        std::memcpy(&id, id_128, sizeof(id));
        NMF_NCCL(NcclApi::get().CommInitRank(&h->comm, nranks, id, rank));
        h->rank = rank;
        h->nranks = nranks;
    });
}

int nmfb200_comm_destroy(nmfb200_handle* h) {
    return guarded(h, [&] {
        if (h->comm) NMF_NCCL(NcclApi::get().CommDestroy(h->comm));
        h->comm = nullptr;
        h->rank = 0;
        h->nranks = 1;
    });
}

}  // extern "C"
