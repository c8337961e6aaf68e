// simt_engine.cu -- the exact engine: CUDA-core fp32 / fp64 kernels for MultUpdate(:mse), MultUpdate(:div)
// and GreedyCD.  It serves Float64 problems (BASELINE config 1), odd shapes, and is the on-GPU
// cross-check for the tensor-core engine.  Arithmetic is the reference's, in the reference's
// element type T, with these deliberate departures (all documented in DESIGN.md):
//   * MU-MSE uses (W'W)H and W(HH') instead of W'(WH) and (WH)H' (multupd.jl:99,110) -- the same
//     mathematics without the p x n intermediate.
//   * reductions run in parallel (pairwise order) instead of Julia's sequential order.
// Layouts on device are the caller's: column-major X (p x n), W (p x k), H (k x n).
#include <algorithm>

#include "common.cuh"
#include "gcd_kernels.cuh"
#include "philox.cuh"

namespace nmfb200 {
namespace {

// ---- generic strided GEMM: C(m,n) = sum_k A(m,k) B(k,n); 64x64x16 tiles, 256 threads, 4x4 per thread
constexpr int GBM = 64, GBN = 64, GBK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_kernel(int M, int N, int K, const T* __restrict__ A, int64_t sAm, int64_t sAk,
                                                   const T* __restrict__ B, int64_t sBk, int64_t sBn, T* __restrict__ C,
                                                   int64_t sCm, int64_t sCn, int64_t sCz, int kchunk) {
    __shared__ T As[GBK][GBM + 4];
    __shared__ T Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    const bool a_kfast = (sAk == 1), b_kfast = (sBk == 1);
    const int tx = tid % 16, ty = tid / 16;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

    for (int k0 = kbeg; k0 < kend; k0 += GBK) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int mm, kk;
            if (a_kfast) { kk = tid % GBK; mm = tid / GBK + 16 * r; } else { mm = tid % GBM; kk = tid / GBM + 4 * r; }
            int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < kend) ? A[(int64_t)gm * sAm + (int64_t)gk * sAk] : T(0);
            int nn, kb;
            if (b_kfast) { kb = tid % GBK; nn = tid / GBK + 16 * r; } else { nn = tid % GBN; kb = tid / GBN + 4 * r; }
            int gn = n0 + nn, gkb = k0 + kb;
            Bs[kb][nn] = (gn < N && gkb < kend) ? B[(int64_t)gkb * sBk + (int64_t)gn * sBn] : T(0);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GBK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
    T* Cz = C + (int64_t)blockIdx.z * sCz;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn < N) Cz[(int64_t)gm * sCm + (int64_t)gn * sCn] = acc[i][j];
        }
    }
}

template <typename T>
__global__ void reduce_splits_kernel(int M, int N, int splits, const T* __restrict__ part, T* __restrict__ C, int64_t sCm,
                                     int64_t sCn) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    T s = T(0);
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * M * N + idx];
    int m = idx / N, nn = idx % N;
    C[(int64_t)m * sCm + (int64_t)nn * sCn] = s;
}

// ---- MU-MSE ratio: F[i] *= max(0, Num[i]-lambda) / (Den[i]+delta)   (multupd.jl:101-103, :112-114)
template <typename T>
__global__ void mu_mse_ratio_kernel(T* __restrict__ F, const T* __restrict__ Num, const T* __restrict__ Den, int64_t len,
                                    T lambda, T delta) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        T num = sub_rn(Num[i], lambda);
        if (!(num > T(0))) num = (num != num) ? num : T(0);
        F[i] = mul_rn(F[i], div_rn(num, add_rn(Den[i], delta)));
    }
}

// ---- MU-Div pieces (multupd.jl:172-179, :184-191)
template <typename T>
__global__ void mu_div_quot_kernel(T* __restrict__ Q, const T* __restrict__ X, int64_t p, int64_t n, int64_t ldx,
                                   const T* __restrict__ WH, T delta) {
    int64_t len = p * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i % p, c = i / p;
        Q[i] = div_rn(X[r + c * ldx], add_rn(WH[i], delta));
    }
}

// out[j] = sum_i A[i*sI + j*sJ], one block per j (double accumulation, fixed order => deterministic)
template <typename T>
__global__ void strided_sum_kernel(const T* __restrict__ A, int64_t len, int64_t sI, int64_t sJ, T* __restrict__ out) {
    __shared__ double red[256];
    const T* a = A + (int64_t)blockIdx.x * sJ;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) s += (double)a[i * sI];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (T)red[0];
}

// H[i,j] *= WtQ[i,j] / (sW[i] + lambda)  (k x n col-major)   /   W[i,j] *= QHt[i,j] / (sH[j] + lambda) (p x k)
template <typename T>
__global__ void mu_div_scale_kernel(T* __restrict__ F, const T* __restrict__ Num, const T* __restrict__ s, int64_t rows,
                                    int64_t cols, int s_along_rows, T lambda) {
    int64_t len = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i % rows, c = i / rows;
        T d = add_rn(s_along_rows ? s[r] : s[c], lambda);
        F[i] = mul_rn(F[i], div_rn(Num[i], d));
    }
}

// ---- stop_condition partial sums (common.jl:92-111).  One block per component j:
// acc[4*j..] = {dev_w, sum_w, dev_h, sum_h} as doubles.  W p x k (col j contiguous), H k x n (row j, stride k).
template <typename T>
__global__ void stop_partials_kernel(const T* __restrict__ W, const T* __restrict__ preW, int64_t p, const T* __restrict__ H,
                                     const T* __restrict__ preH, int64_t n, int64_t k, double* __restrict__ acc) {
    __shared__ double red[4][256];
    const int j = blockIdx.x;
    double dw = 0, sw = 0, dh = 0, sh = 0;
    for (int64_t i = threadIdx.x; i < p; i += blockDim.x) {
        T a = W[i + j * p], b = preW[i + j * p];
        T d = sub_rn(a, b), s = add_rn(a, b);
        dw += (double)mul_rn(d, d);
        sw += (double)mul_rn(s, s);
    }
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        T a = H[j + i * k], b = preH[j + i * k];
        T d = sub_rn(a, b), s = add_rn(a, b);
        dh += (double)mul_rn(d, d);
        sh += (double)mul_rn(s, s);
    }
    red[0][threadIdx.x] = dw; red[1][threadIdx.x] = sw; red[2][threadIdx.x] = dh; red[3][threadIdx.x] = sh;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int q = 0; q < 4; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 4) acc[4 * j + threadIdx.x] = red[threadIdx.x][0];
}

// Final decision, one thread: flag[0] = converged, dev[0] = devmax over all components (comparisons in T)
template <typename T>
__global__ void stop_final_kernel(const double* __restrict__ acc, int64_t k, T tol, int* __restrict__ flag, double* __restrict__ dev) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool conv = true;
    T devmax = T(0);
    for (int64_t j = 0; j < k; ++j) {
        T dw = (T)acc[4 * j], sw = (T)acc[4 * j + 1], dh = (T)acc[4 * j + 2], sh = (T)acc[4 * j + 3];
        T rw = dw / sw, rh = dh / sh;
        T m = (rw != rw) ? rw : ((rh != rh) ? rh : (rw > rh ? rw : rh));
        T sq = sqrt(m);
        devmax = (devmax != devmax) ? devmax : ((sq != sq) ? sq : (sq > devmax ? sq : devmax));
        if (sqrt(dw) > tol * sqrt(sw) || sqrt(dh) > tol * sqrt(sh)) conv = false;
    }
    flag[0] = conv ? 1 : 0;
    dev[0] = (double)devmax;
}

// ---- objectives (StatsBase.sqL2dist / gkldiv semantics: differences in T, Float64 accumulator)
template <typename T, int MODE>  // MODE 0: sum (x-y)^2, 1: gkldiv, 2: sum |y| (x ignored), 3: sum y^2 (x ignored)
__global__ void objective_partials_kernel(const T* __restrict__ X, int64_t p, int64_t n, int64_t ldx, const T* __restrict__ Y,
                                          double* __restrict__ part) {
    __shared__ double red[256];
    int64_t len = p * n;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        if (MODE == 2) {
            s += (double)fabs(Y[i]);
        } else if (MODE == 3) {
            s += (double)mul_rn(Y[i], Y[i]);
        } else {
            int64_t r = i % p, c = i / p;
            T a = X[r + c * ldx], b = Y[i];
            if (MODE == 0) {
                T d = sub_rn(a, b);
                s += (double)mul_rn(d, d);
            } else {
                if (a > T(0)) {
                    T t = mul_rn(a, (T)log(div_rn(a, b)));
                    t = sub_rn(t, a);
                    t = add_rn(t, b);
                    s += (double)t;
                } else {
                    s += (double)b;
                }
            }
        }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

__global__ void sum_partials_kernel(const double* __restrict__ part, int nparts, double* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ---- GreedyCD (greedycd.jl:94-166) ---------------------------------------------------------------
// G[i][r] (row-major rows x k) <- (F P)[i][r] - Z[i][r] (+ lambda); FP and Z are row-major rows x k.
template <typename T>
__global__ void gcd_form_g_kernel(T* __restrict__ G, const T* __restrict__ Z, int64_t len, T lambda, int add_lambda) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        T g = sub_rn(G[i], Z[i]);
        if (add_lambda) g = add_rn(g, lambda);
        G[i] = g;
    }
}

// strided copy (2D) for packing caller matrices with ld != rows
template <typename T>
__global__ void copy2d_kernel(T* __restrict__ dst, int64_t ldd, const T* __restrict__ src, int64_t lds, int64_t rows, int64_t cols) {
    int64_t len = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i % rows, c = i / rows;
        dst[r + c * ldd] = src[r + c * lds];
    }
}


// ---- ProjectedALS pieces (projals.jl:77-107, utils.jl:15-24, :34-41, :63-84) ------------------------
template <typename T>
__global__ void add_diag_kernel(T* __restrict__ A, int k, T a) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) A[(int64_t)i * k + i] = add_rn(A[(int64_t)i * k + i], a);
}

// projectnn! (utils.jl:34-41): NaN < 0 is false => NaN preserved
template <typename T>
__global__ void clamp_nn_kernel(T* __restrict__ A, int64_t len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        if (A[i] < T(0)) A[i] = T(0);
}

// inv(A) for a symmetric positive definite k x k matrix (what potrf!/potrs! and potrf!/potri! deliver, utils.jl:63-84),
// by in-place Gauss-Jordan elimination without pivoting (stable for SPD) in Float64, one CTA.  work: k*k doubles.
// info[0] = j+1 if pivot j is not positive (LAPACK potrf! info > 0), else untouched.
template <typename T>
__global__ void __launch_bounds__(1024) spd_inverse_kernel(const T* __restrict__ A, int k, double* __restrict__ work, T* __restrict__ inv,
                                                           int* __restrict__ info) {
    extern __shared__ double sh[];  // [k] pivot column, [k] scaled pivot row
    double* colj = sh;
    double* rowj = sh + k;
    __shared__ int bad;
    const int n2 = k * k;
    if (threadIdx.x == 0) bad = 0;
    for (int e = threadIdx.x; e < n2; e += blockDim.x) work[e] = (double)A[e];
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        const double piv = work[(size_t)j * k + j];
        if (!(piv > 0.0)) {  // uniform: every thread reads the same element
            if (threadIdx.x == 0) info[0] = j + 1;
            bad = 1;
        }
        const double ipiv = 1.0 / piv;
        for (int e = threadIdx.x; e < k; e += blockDim.x) {
            colj[e] = work[(size_t)j * k + e];          // column j: element (e, j)   (col-major: e + j*k)
            rowj[e] = work[(size_t)e * k + j] * ipiv;   // row j / pivot: element (j, e)
        }
        __syncthreads();
        if (bad) return;
        for (int e = threadIdx.x; e < n2; e += blockDim.x) {
            const int r = e % k, c = e / k;
            double v;
            if (r == j) v = (c == j) ? ipiv : rowj[c];
            else if (c == j) v = -colj[r] * ipiv;
            else v = work[e] - colj[r] * rowj[c];
            work[e] = v;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n2; e += blockDim.x) inv[e] = (T)work[e];
}

// ---- CoordinateDescent sweep (coorddesc.jl:138-157) ------------------------------------------------------
// One thread per row i of the factor F(i, r) = Fp[i*sFr + r*sFc]; the row lives in shared memory (transposed:
// conflict-free), components are visited in `perm` order, the gradient is accumulated sequentially from
// -(XHt[i,t] - l1) exactly like the reference's scalar loop (un-fused multiply-add).  HHt (k x k col-major, l2
// already on the diagonal) is read through the read-only cache (warp-uniform addresses => broadcast).
// Z(i, t) = Zp[i*sZr + t*sZc] is X*O.  The violation sum (coorddesc.jl:150) is not formed: it is stored in the
// reference's state and never read.
template <typename T>
__global__ void cd_sweep_kernel(T* __restrict__ Fp, int64_t sFr, int64_t sFc, const T* __restrict__ HHt, const T* __restrict__ Zp,
                                int64_t sZr, int64_t sZc, int64_t rows, int k, const int* __restrict__ perm, T l1) {
    extern __shared__ unsigned char cd_smem_raw[];
    T* row = (T*)cd_smem_raw;  // [k][blockDim.x]
    const int tx = threadIdx.x, nt = blockDim.x;
    const int64_t i = (int64_t)blockIdx.x * nt + tx;
    if (i >= rows) return;
    for (int r = 0; r < k; ++r) row[r * nt + tx] = Fp[i * sFr + r * sFc];
    for (int tt = 0; tt < k; ++tt) {
        const int t = perm[tt];
        T x = Zp[i * sZr + t * sZc];
        if (l1 > T(0)) x = sub_rn(x, l1);                                   // coorddesc.jl:126-128
        T grad = -x;                                                        // :141
        const T* hrow = HHt + t;                                            // HHt[t, r] = HHt[t + r*k]
        for (int r = 0; r < k; ++r) grad = add_rn(grad, mul_rn(__ldg(hrow + (int64_t)r * k), row[r * nt + tx]));  // :143-145
        const T hess = __ldg(HHt + (int64_t)t * k + t);                     // :153
        if (hess != T(0)) {
            T v = sub_rn(row[t * nt + tx], div_rn(grad, hess));             // :155
            row[t * nt + tx] = (v > T(0) || v != v) ? v : T(0);             // max(v, 0): NaN propagates
        }
    }
    for (int r = 0; r < k; ++r) Fp[i * sFr + r * sFc] = row[r * nt + tx];
}

// ---- ALSPGrad pieces (alspgrad.jl:86-191, :242-347) ------------------------------------------------------
template <typename T>
__global__ void sub_inplace_kernel(T* __restrict__ G, const T* __restrict__ B, int64_t len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        G[i] = sub_rn(G[i], B[i]);                                          // :118-120
}
// Fn = max(F - alpha*G, 0), D = Fn - F  (:134-138)
template <typename T>
__global__ void pg_step_kernel(const T* __restrict__ F, const T* __restrict__ G, T alpha, T* __restrict__ Fn, T* __restrict__ D, int64_t len) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        T f = F[i];
        T v = sub_rn(f, mul_rn(alpha, G[i]));
        v = (v > T(0) || v != v) ? v : T(0);
        Fn[i] = v;
        D[i] = sub_rn(v, f);
    }
}
// MODE 0: sum g^2 over {g < 0 or x > 0} (projgradnorm, :9-19); 1: sum a*b (BLAS.dot); 2: sum (a-b)^2 (isapprox norm)
template <typename T, int MODE>
__global__ void pair_reduce_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t len, double* __restrict__ part) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        T x = a[i], y = b[i];
        if (MODE == 0) { if (x < T(0) || y > T(0)) s += (double)mul_rn(x, x); }
        else if (MODE == 1) s += (double)x * (double)y;
        else { T d = sub_rn(x, y); s += (double)mul_rn(d, d); }
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// `randperm` stand-in for CoordinateDescent(shuffle=true): splitmix64 + Fisher-Yates, identical to oracle ShufflePerm
struct ShufflePerm {
    uint64_t s;
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void perm(std::vector<int>& a, int k) {
        a.resize(k);
        for (int i = 0; i < k; ++i) a[i] = i;
        for (int i = k - 1; i >= 1; --i) {
            int j = (int)(next() % (uint64_t)(i + 1));
            std::swap(a[i], a[j]);
        }
    }
};

inline int ew_blocks(int64_t len) { return (int)std::min<int64_t>(ceil_div(len, 256), 148 * 16); }

template <typename T>
struct Simt {
    nmfb200_handle* h;
    cudaStream_t st;
    int64_t p, n, k;
    const T* X;
    int64_t ldx;

    void gemm(int M, int N, int K, const T* A, int64_t sAm, int64_t sAk, const T* B, int64_t sBk, int64_t sBn, T* C, int64_t sCm,
              int64_t sCn) {
        int tiles = (int)(ceil_div(M, GBM) * ceil_div(N, GBN));
        int splits = 1;
        if (tiles < 148 && K >= 4096) splits = (int)std::min<int64_t>(std::max<int64_t>(1, 296 / tiles), ceil_div(K, 1024));
        int kchunk = (int)(ceil_div(ceil_div(K, splits), GBK) * GBK);
        splits = (int)ceil_div(K, kchunk);
        dim3 grid((unsigned)ceil_div(N, GBN), (unsigned)ceil_div(M, GBM), (unsigned)splits);
        if (splits == 1) {
            gemm_kernel<T><<<grid, 256, 0, st>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, C, sCm, sCn, 0, kchunk);
            h->launches += 1;
        } else {
            T* part = h->buf_t<T>("simt.gemm_part", (size_t)splits * M * N);
            gemm_kernel<T><<<grid, 256, 0, st>>>(M, N, K, A, sAm, sAk, B, sBk, sBn, part, N, 1, (int64_t)M * N, kchunk);
            reduce_splits_kernel<T><<<(unsigned)ceil_div((int64_t)M * N, 256), 256, 0, st>>>(M, N, splits, part, C, sCm, sCn);
            h->launches += 2;
        }
        NMF_CUDA(cudaGetLastError());
    }

    // The two X-sized products of every iterative algorithm: side 0: out (n x k) = X' O with O p x k; side 1: out (p x k) = X O
    // with O n x k.  O(c, a) = O[c*sOr + a*sOc], out(r, a) = out[r*sNr + a*sNc].  Float32 problems from 2^20 cells on go to the
    // tcgen05 mainloop of the tensor-core engine with split (bf16 hi + lo) operands (tc_xmul, ~2^-16 relative per product,
    // fp32 accumulation); everything else, and every k x k product, stays on the fp32 / fp64 CUDA-core GEMM.
    bool used_tc = false;
    void xprod(int side, const T* O, int64_t sOr, int64_t sOc, T* out, int64_t sNr, int64_t sNc) {
        if constexpr (sizeof(T) == 4) {
            if (tc_xmul(h, side, (const float*)O, sOr, sOc, k, (float*)out, sNr, sNc)) {
                used_tc = true;
                return;
            }
        }
        if (side == 0) gemm((int)n, (int)k, (int)p, X, ldx, 1, O, sOr, sOc, out, sNr, sNc);
        else gemm((int)p, (int)k, (int)n, X, 1, ldx, O, sOr, sOc, out, sNr, sNc);
    }

    double reduce_objective(int mode, const T* Xp, int64_t rows, int64_t cols, int64_t ld, const T* Y) {
        int nb = ew_blocks(rows * cols);
        double* part = h->buf_t<double>("simt.obj_part", (size_t)nb + 1);
        if (mode == 0) objective_partials_kernel<T, 0><<<nb, 256, 0, st>>>(Xp, rows, cols, ld, Y, part);
        else if (mode == 1) objective_partials_kernel<T, 1><<<nb, 256, 0, st>>>(Xp, rows, cols, ld, Y, part);
        else if (mode == 3) objective_partials_kernel<T, 3><<<nb, 256, 0, st>>>(Xp, rows, cols, ld, Y, part);
        else objective_partials_kernel<T, 2><<<nb, 256, 0, st>>>(Xp, rows, cols, ld, Y, part);
        sum_partials_kernel<<<1, 256, 0, st>>>(part, nb, part + nb);
        h->launches += 2;
        h->allreduce_sum(part + nb, 1);  // row-sharded X: partial objective per rank
        double v = 0;
        NMF_CUDA(cudaMemcpyAsync(&v, part + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        return v;
    }

    // objective with W (p x k), H (k x n) column-major compact
    double objective(int alg, const T* W, const T* H, double lambda_w, double lambda_h, int64_t ldw = 0, int64_t ldh = 0) {
        if (ldw == 0) ldw = p;
        if (ldh == 0) ldh = k;
        T* WH = h->buf_t<T>("simt.WH", (size_t)p * n);
        gemm((int)p, (int)n, (int)k, W, 1, ldw, H, 1, ldh, WH, 1, p);
        if (alg == 1) return (double)(T)reduce_objective(1, X, p, n, ldx, WH);  // gkldiv, Result converts to T
        T r = mul_rn_host(T(0.5), (T)reduce_objective(0, X, p, n, ldx, WH));
        if (alg == 3) {  // projals.jl:66-75: + (0.5*lambda) * abs2(norm(F)), norm in T
            if (lambda_w > 0) {
                T nw = (T)std::sqrt(reduce_objective(3, nullptr, p, k, p, W));
                r = (T)(r + (T)(T(0.5) * (T)lambda_w) * (T)(nw * nw));
            }
            if (lambda_h > 0) {
                T nh = (T)std::sqrt(reduce_objective(3, nullptr, k, n, k, H) / (h->comm ? h->nranks : 1));
                r = (T)(r + (T)(T(0.5) * (T)lambda_h) * (T)(nh * nh));
            }
        }
        if (alg == 2) {  // greedycd.jl:85-90
            if (lambda_w > 0) {
                double l1 = reduce_objective(2, nullptr, p, k, p, W);
                r = (T)(r + (T)lambda_w * (T)l1);
            }
            if (lambda_h > 0) {
                // H is replicated across ranks: undo the allreduce multiplication
                double l1 = reduce_objective(2, nullptr, k, n, k, H) / (h->comm ? h->nranks : 1);
                r = (T)(r + (T)lambda_h * (T)l1);
            }
        }
        return (double)r;
    }
    static T mul_rn_host(T a, T b) { return a * b; }

    // one call of stop_condition; returns converged and dev
    bool stop(const T* W, const T* preW, const T* H, const T* preH, T tol, double* dev_out) {
        double* acc = h->buf_t<double>("simt.stop_acc", (size_t)4 * k + 2);
        int* flag = (int*)h->buf("simt.stop_flag", 16);
        stop_partials_kernel<T><<<(unsigned)k, 256, 0, st>>>(W, preW, p, H, preH, n, k, acc);
        h->launches += 1;
        if (h->comm) {
            // W rows are sharded: sum dev_w/sum_w over ranks; H sums are replicated -> divide back
            h->allreduce_sum(acc, (size_t)4 * k);
            // (the H entries were multiplied by nranks; stop_final works on ratios dev_h/sum_h and on
            //  sqrt(dev_h) > tol*sqrt(sum_h), both invariant under a common positive factor.)
        }
        stop_final_kernel<T><<<1, 32, 0, st>>>(acc, k, tol, flag, acc + 4 * k);
        h->launches += 1;
        int hflag = 0;
        double hdev = 0;
        NMF_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpyAsync(&hdev, acc + 4 * k, sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        *dev_out = hdev;
        return hflag != 0;
    }

    // ---- MU-MSE iteration (multupd.jl:83-116) in Gram form
    void iter_multmse(T* W, T* H, bool update_H, T lw, T lh, T delta) {
        T* num_h = h->buf_t<T>("simt.num_h", (size_t)k * n + (size_t)k * k);  // [A (k x n) | G (k x k)] packed for one allreduce
        T* gram = num_h + (size_t)k * n;
        T* den_h = h->buf_t<T>("simt.den_h", (size_t)k * n);
        T* num_w = h->buf_t<T>("simt.num_w", (size_t)p * k);
        T* den_w = h->buf_t<T>("simt.den_w", (size_t)p * k);
        if (update_H) {
            gemm((int)k, (int)n, (int)p, W, p, 1, X, 1, ldx, num_h, 1, k);  // W'X (k x n)
            gemm((int)k, (int)k, (int)p, W, p, 1, W, 1, p, gram, 1, k);    // W'W
            h->allreduce_sum(num_h, (size_t)k * n + (size_t)k * k);
            gemm((int)k, (int)n, (int)k, gram, 1, k, H, 1, k, den_h, 1, k);  // (W'W) H
            mu_mse_ratio_kernel<T><<<ew_blocks(k * n), 256, 0, st>>>(H, num_h, den_h, k * n, lh, delta);
            h->launches += 1;
        }
        T* gramh = h->buf_t<T>("simt.gramh", (size_t)k * k);
        gemm((int)p, (int)k, (int)n, X, 1, ldx, H, k, 1, num_w, 1, p);      // X H' (p x k)
        gemm((int)k, (int)k, (int)n, H, 1, k, H, k, 1, gramh, 1, k);        // H H'
        gemm((int)p, (int)k, (int)k, W, 1, p, gramh, 1, k, den_w, 1, p);    // W (H H')
        mu_mse_ratio_kernel<T><<<ew_blocks(p * k), 256, 0, st>>>(W, num_w, den_w, p * k, lw, delta);
        h->launches += 1;
        NMF_CUDA(cudaGetLastError());
    }

    // ---- MU-Div iteration (multupd.jl:150-193), as written with materialised WH and Q
    void iter_multdiv(T* W, T* H, bool update_H, T lw, T lh, T delta, bool first) {
        T* WH = h->buf_t<T>("simt.WH", (size_t)p * n);
        T* Q = h->buf_t<T>("simt.Q", (size_t)p * n);
        if (first) gemm((int)p, (int)n, (int)k, W, 1, p, H, 1, k, WH, 1, p);  // prepare_state (multupd.jl:138)
        if (update_H) {
            T* wtq = h->buf_t<T>("simt.num_h", (size_t)k * n + (size_t)k * k);  // [W'Q | sW] packed
            T* sW = wtq + (size_t)k * n;
            mu_div_quot_kernel<T><<<ew_blocks(p * n), 256, 0, st>>>(Q, X, p, n, ldx, WH, delta);
            h->launches += 1;
            gemm((int)k, (int)n, (int)p, W, p, 1, Q, 1, p, wtq, 1, k);
            strided_sum_kernel<T><<<(unsigned)k, 256, 0, st>>>(W, p, 1, p, sW);
            h->launches += 1;
            h->allreduce_sum(wtq, (size_t)k * n + (size_t)k);
            mu_div_scale_kernel<T><<<ew_blocks(k * n), 256, 0, st>>>(H, wtq, sW, k, n, 1, lh);
            h->launches += 1;
            gemm((int)p, (int)n, (int)k, W, 1, p, H, 1, k, WH, 1, p);
        }
        T* qht = h->buf_t<T>("simt.num_w", (size_t)p * k);
        T* sH = h->buf_t<T>("simt.sH", (size_t)k);
        mu_div_quot_kernel<T><<<ew_blocks(p * n), 256, 0, st>>>(Q, X, p, n, ldx, WH, delta);
        gemm((int)p, (int)k, (int)n, Q, 1, p, H, k, 1, qht, 1, p);
        strided_sum_kernel<T><<<(unsigned)k, 256, 0, st>>>(H, n, k, 1, sH);
        mu_div_scale_kernel<T><<<ew_blocks(p * k), 256, 0, st>>>(W, qht, sH, p, k, 0, lw);
        h->launches += 3;
        gemm((int)p, (int)n, (int)k, W, 1, p, H, 1, k, WH, 1, p);
        NMF_CUDA(cudaGetLastError());
    }

    // ---- GreedyCD half-step (greedycd.jl:94-166) for factor F (rows x k, element strides sFr/sFc)
    // against O (cols x k, strides sOr/sOc); Xv(r, c) = Xp[r*sXr + c*sXc].
    void gcd_half(T* F, int64_t sFr, int64_t sFc, const T* O, int64_t sOr, int64_t sOc, int64_t rows, int64_t cols, int64_t sXr,
                  int64_t sXc, T lambda, bool rows_sharded, unsigned long long* d_updates) {
        // packed [Z (rows x k row-major) | P (k x k)]: when the contraction dim is sharded both are allreduced
        T* Z = h->buf_t<T>("simt.gcd_Z", (size_t)std::max(p, n) * k + (size_t)k * k);
        T* P = Z + (size_t)rows * k;
        T* G = h->buf_t<T>("simt.gcd_G", (size_t)std::max(p, n) * k);
        gemm((int)k, (int)k, (int)cols, O, sOc, sOr, O, sOr, sOc, P, 1, k);          // P = O'O   (:117)
        gemm((int)rows, (int)k, (int)cols, X, sXr, sXc, O, sOr, sOc, Z, k, 1);      // Z = X O   (:118)
        if (!rows_sharded) h->allreduce_sum(Z, (size_t)rows * k + (size_t)k * k);   // H-step under row sharding
        gemm((int)rows, (int)k, (int)k, F, sFr, sFc, P, 1, k, G, k, 1);              // G = F P   (:119)
        gcd_form_g_kernel<T><<<ew_blocks(rows * k), 256, 0, st>>>(G, Z, rows * k, lambda, lambda > T(0) ? 1 : 0);  // :120-123
        int nb = (int)ceil_div(rows, GCD_WARPS);
        T* bmax = h->buf_t<T>("simt.gcd_bmax", (size_t)nb + 1);
        gcd_rowmax_kernel<T><<<nb, GCD_WARPS * 32, 0, st>>>(F, sFr, sFc, G, P, rows, (int)k, bmax);
        max_partials_kernel<T><<<1, 256, 0, st>>>(bmax, nb, bmax + nb);
        if (rows_sharded) h->allreduce_max(bmax + nb, 1);  // p_init is a max over ALL rows (:132-137)
        size_t smem = ((size_t)k + (size_t)GCD_WARPS * 3 * k) * sizeof(T);
        NMF_CUDA(cudaFuncSetAttribute(gcd_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gcd_rows_kernel<T><<<nb, GCD_WARPS * 32, smem, st>>>(F, sFr, sFc, G, P, rows, (int)k, bmax + nb, d_updates);
        h->launches += 4;
        NMF_CUDA(cudaGetLastError());
    }

    // ---- ProjectedALS iteration (projals.jl:77-107).  The k x k systems are solved through the explicit inverse
    // (Float64 Gauss-Jordan in one CTA) followed by a GEMM, for H (reference: potrs!) as well as W (reference: potri!).
    void spd_inverse(const T* A, T* inv, int* info) {
        double* work = h->buf_t<double>("simt.inv_work", (size_t)k * k);
        spd_inverse_kernel<T><<<1, 1024, 2 * (size_t)k * sizeof(double), st>>>(A, (int)k, work, inv, info);
        h->launches += 1;
    }
    void iter_projals(T* W, T* H, bool update_H, T lw, T lh, int* info) {
        T* num_h = h->buf_t<T>("simt.num_h", (size_t)k * n + (size_t)k * k);  // [W'X | W'W] packed for one allreduce
        T* gram = num_h + (size_t)k * n;
        T* inv = h->buf_t<T>("simt.inv", (size_t)k * k);
        if (update_H) {
            xprod(0, W, 1, p, num_h, k, 1);                                  // W'X        (projals.jl:92)
            gemm((int)k, (int)k, (int)p, W, p, 1, W, 1, p, gram, 1, k);     // W'W        (:91)
            h->allreduce_sum(num_h, (size_t)k * n + (size_t)k * k);
            if (lh != T(0)) { add_diag_kernel<T><<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(gram, (int)k, lh); h->launches += 1; }
            spd_inverse(gram, inv, info);                                    // pdsolve!   (:93)
            gemm((int)k, (int)n, (int)k, inv, 1, k, num_h, 1, k, H, 1, k);
            clamp_nn_kernel<T><<<ew_blocks(k * n), 256, 0, st>>>(H, k * n);  // projectnn! (:94)
            h->launches += 1;
        }
        T* gramh = h->buf_t<T>("simt.gramh", (size_t)k * k);
        T* num_w = h->buf_t<T>("simt.num_w", (size_t)p * k);
        gemm((int)k, (int)k, (int)n, H, 1, k, H, k, 1, gramh, 1, k);         // HH'        (:99)
        if (lw != T(0)) { add_diag_kernel<T><<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(gramh, (int)k, lw); h->launches += 1; }
        xprod(1, H, k, 1, num_w, 1, p);                                      // XH'        (:100)
        spd_inverse(gramh, inv, info);                                       // pdrsolve!  (:101)
        gemm((int)p, (int)k, (int)k, num_w, 1, p, inv, 1, k, W, 1, p);
        clamp_nn_kernel<T><<<ew_blocks(p * k), 256, 0, st>>>(W, p * k);      // projectnn! (:102)
        h->launches += 1;
        NMF_CUDA(cudaGetLastError());
    }

    // ---- CoordinateDescent half-step (coorddesc.jl:108-160): factor F (rows x k) against O (cols x k)
    void cd_half(T* F, int64_t sFr, int64_t sFc, const T* O, int64_t sOr, int64_t sOc, int64_t rows, int64_t cols, int64_t sXr,
                 int64_t sXc, T l1, T l2, bool contraction_sharded, ShufflePerm* rng) {
        // packed [Z = X*O (rows x k, row-major) | HHt = O'O (k x k)]
        T* Z = h->buf_t<T>("simt.gcd_Z", (size_t)std::max(p, n) * k + (size_t)k * k);
        T* HHt = Z + (size_t)rows * k;
        gemm((int)k, (int)k, (int)cols, O, sOc, sOr, O, sOr, sOc, HHt, 1, k);        // :112
        xprod(sXr == 1 ? 1 : 0, O, sOr, sOc, Z, k, 1);                              // :118
        if (contraction_sharded) h->allreduce_sum(Z, (size_t)rows * k + (size_t)k * k);
        if (l2 > T(0)) { add_diag_kernel<T><<<(unsigned)ceil_div(k, 256), 256, 0, st>>>(HHt, (int)k, l2); h->launches += 1; }  // :123-125
        std::vector<int> perm;
        if (rng) rng->perm(perm, (int)k);                                            // :129-133
        else { perm.resize(k); for (int i = 0; i < (int)k; ++i) perm[i] = i; }
        int* dperm = (int*)h->buf("simt.cd_perm", (size_t)k * sizeof(int));
        NMF_CUDA(cudaMemcpyAsync(dperm, perm.data(), (size_t)k * sizeof(int), cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaStreamSynchronize(st));  // perm is a stack object; the copy must have left it
        int nt = 128;
        while (nt > 32 && (size_t)nt * k * sizeof(T) > 200 * 1024) nt >>= 1;
        const size_t smem = (size_t)nt * k * sizeof(T);
        NMF_REQUIRE(smem <= 200 * 1024, NMFB200_ENOTSUP, "CoordinateDescent: k too large for the row kernel");
        NMF_CUDA(cudaFuncSetAttribute(cd_sweep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cd_sweep_kernel<T><<<(unsigned)ceil_div(rows, nt), nt, smem, st>>>(F, sFr, sFc, HHt, Z, k, 1, rows, (int)k, dperm, l1);  // :138-157
        h->launches += 1;
        NMF_CUDA(cudaGetLastError());
    }
    void iter_cd(T* W, T* H, bool update_H, T l1W, T l2W, T l1H, T l2H, ShufflePerm* rng) {
        cd_half(W, 1, p, H, k, 1, p, n, 1, ldx, l1W, l2W, false, rng);                        // coorddesc.jl:168
        if (update_H) cd_half(H, k, 1, W, 1, p, n, p, ldx, 1, l1H, l2H, h->comm != nullptr, rng);  // :171-176
    }

    // ---- ALSPGrad (alspgrad.jl:400-425) ----------------------------------------------------------------
    double pair_reduce(int mode, const T* a, const T* b, int64_t len) {
        int nb = ew_blocks(len);
        double* part = h->buf_t<double>("simt.pg_part", (size_t)nb + 1);
        if (mode == 0) pair_reduce_kernel<T, 0><<<nb, 256, 0, st>>>(a, b, len, part);
        else if (mode == 1) pair_reduce_kernel<T, 1><<<nb, 256, 0, st>>>(a, b, len, part);
        else pair_reduce_kernel<T, 2><<<nb, 256, 0, st>>>(a, b, len, part);
        sum_partials_kernel<<<1, 256, 0, st>>>(part, nb, part + nb);
        h->launches += 2;
        double v = 0;
        NMF_CUDA(cudaMemcpyAsync(&v, part + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        return v;
    }
    // _alspgrad_updateh! (left = true: G = gram*F - cross, F is k x n) / _alspgrad_updatew! (left = false:
    // G = F*gram - cross, F is p x k).  Control flow on the host, exactly the reference's; returns the sub-iteration count.
    int64_t pg_subsolve(T* F, const T* gram, const T* cross, bool left, int64_t maxsub, int traceiter, T tolg, T beta, T sigma) {
        const int64_t rows = left ? k : p, cols = left ? n : k, len = rows * cols;
        T* G = h->buf_t<T>("simt.pg_G", (size_t)std::max(k * n, p * k));
        T* Fn = h->buf_t<T>("simt.pg_Fn", (size_t)std::max(k * n, p * k));
        T* Fp = h->buf_t<T>("simt.pg_Fp", (size_t)std::max(k * n, p * k));
        T* D = h->buf_t<T>("simt.pg_D", (size_t)std::max(k * n, p * k));
        T* GD = h->buf_t<T>("simt.pg_GD", (size_t)std::max(k * n, p * k));
        auto mul = [&](T* dst, const T* src) {
            if (left) gemm((int)k, (int)n, (int)k, gram, 1, k, src, 1, k, dst, 1, k);
            else gemm((int)p, (int)k, (int)k, src, 1, p, gram, 1, k, dst, 1, p);
        };
        const size_t bytes = (size_t)len * sizeof(T);
        const T epsT = std::numeric_limits<T>::epsilon();
        int64_t t = 0;
        bool converged = false, decr_alpha = true;
        T alpha = T(1);
        while (!converged && t < maxsub) {
            ++t;
            mul(G, F);                                                               // :117
            sub_inplace_kernel<T><<<ew_blocks(len), 256, 0, st>>>(G, cross, len);    // :118-120
            h->launches += 1;
            const T pgnrm = (T)std::sqrt((T)pair_reduce(0, G, F, len));              // :123
            if (pgnrm < tolg) converged = true;
            int it = 0;
            if (!converged) {
                while (it < traceiter) {
                    ++it;
                    NMF_REQUIRE(std::isfinite((double)alpha), NMFB200_EINVAL, "alpha is not finite");   // :132
                    pg_step_kernel<T><<<ew_blocks(len), 256, 0, st>>>(F, G, alpha, Fn, D, len);  // :134-139
                    h->launches += 1;
                    const T dv1 = (T)pair_reduce(1, G, D, len);                      // :142
                    mul(GD, D);                                                      // :143
                    const T dv2 = (T)pair_reduce(1, GD, D, len);                     // :144
                    const bool suff_decr = (T)((T)((T)(T(1) - sigma) * dv1) + (T)(T(0.5) * dv2)) < T(0);  // :147
                    if (it == 1) {
                        decr_alpha = !suff_decr;                                     // :150
                        NMF_CUDA(cudaMemcpyAsync(Fp, F, bytes, cudaMemcpyDeviceToDevice, st));  // :151
                    }
                    if (decr_alpha) {
                        if (suff_decr) {
                            NMF_CUDA(cudaMemcpyAsync(F, Fn, bytes, cudaMemcpyDeviceToDevice, st));  // :156
                            break;
                        }
                        alpha = (T)(alpha * beta);                                   // :159
                    } else {
                        // isapprox(Hp, Hn, atol=eps(T)): rtol = 0 when atol > 0  =>  norm(Hp - Hn) <= eps(T)
                        const bool close = (T)std::sqrt((T)pair_reduce(2, Fp, Fn, len)) <= epsT;
                        if (!suff_decr || close) {                                   // :162
                            NMF_CUDA(cudaMemcpyAsync(F, Fp, bytes, cudaMemcpyDeviceToDevice, st));  // :163
                            break;
                        }
                        alpha = (T)(alpha / beta);                                   // :166
                        NMF_CUDA(cudaMemcpyAsync(Fp, Fn, bytes, cudaMemcpyDeviceToDevice, st));  // :167
                    }
                }
            }
        }
        return t;
    }
    int64_t iter_alspgrad(T* W, T* H, bool update_H, int64_t maxsub, T* tolg) {
        int64_t sub = 0;
        T* gram = h->buf_t<T>("simt.gramh", (size_t)k * k);
        if (update_H) {
            T* wtx = h->buf_t<T>("simt.num_h", (size_t)k * n + (size_t)k * k);
            gemm((int)k, (int)k, (int)p, W, p, 1, W, 1, p, gram, 1, k);              // set_w! (:55-59)
            xprod(0, W, 1, p, wtx, k, 1);
            const int64_t itH = pg_subsolve(H, gram, wtx, true, maxsub, 20, *tolg, T(0.2), T(0.01));   // :405-407
            sub += itH;
            if (itH == 1) *tolg = (T)((double)*tolg * 0.1);                          // :409-411
        }
        T* xht = h->buf_t<T>("simt.num_w", (size_t)p * k);
        gemm((int)k, (int)k, (int)n, H, 1, k, H, k, 1, gram, 1, k);                  // set_h! (:211-215)
        xprod(1, H, k, 1, xht, 1, p);
        const int64_t itW = pg_subsolve(W, gram, xht, false, maxsub, 20, *tolg, T(0.2), T(0.01));      // :415-417
        sub += itW;
        if (itW == 1) *tolg = (T)((double)*tolg * 0.1);                              // :419-421
        return sub;
    }

    void iter_greedycd(T* W, T* H, bool update_H, T lw, T lh, unsigned long long* d_updates) {
        // W-step: F = W (p x k), O = H' (n x k): O(c, a) = H[a + c*k]; X(r, c) = X[r + c*ldx]
        gcd_half(W, 1, p, H, k, 1, p, n, 1, ldx, lw, /*rows_sharded=*/h->comm != nullptr, d_updates);
        if (update_H) {
            // H-step: F = H' (n x k): F(j, a) = H[a + j*k]; O = W (p x k); X'(j, i) = X[i + j*ldx]
            gcd_half(H, k, 1, W, 1, p, n, p, ldx, 1, lh, /*rows_sharded=*/false, d_updates);
        }
    }
};

#include "init_device.cuh"

}  // namespace

template <typename T>
void simt_rsvd(nmfb200_handle* h, int64_t k, uint64_t seed, T* U, int64_t ldu, T* S, T* V, int64_t ldv) {
    rsvd_impl<T>(h, k, seed, U, ldu, S, V, ldv);
}
template <typename T>
void simt_nndsvd(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int variant, int zeroh, uint64_t seed, int on_device) {
    nndsvd_impl<T>(h, W, ldw, H, ldh, k, variant, zeroh, seed, on_device);
}
template void simt_rsvd<float>(nmfb200_handle*, int64_t, uint64_t, float*, int64_t, float*, float*, int64_t);
template void simt_rsvd<double>(nmfb200_handle*, int64_t, uint64_t, double*, int64_t, double*, double*, int64_t);
template void simt_nndsvd<float>(nmfb200_handle*, float*, int64_t, float*, int64_t, int64_t, int, int, uint64_t, int);
template void simt_nndsvd<double>(nmfb200_handle*, double*, int64_t, double*, int64_t, int64_t, int, int, uint64_t, int);

template <typename T>
void simt_solve(nmfb200_handle* h, const SolveArgs& a, T* Wc, int64_t ldw, T* Hc, int64_t ldh, nmfb200_result* out) {
    NMF_REQUIRE(h->x_elt == (int)sizeof(T), NMFB200_ESTATE, "set_X with the same element type must precede solve");
    Simt<T> s{h, h->stream, h->p, h->n, a.k, (const T*)h->dX, h->ldx};
    cudaStream_t st = h->stream;
    const int64_t p = h->p, n = h->n, k = a.k;
    NMF_REQUIRE(ldw >= p && ldh >= k, NMFB200_EDIM, "Dimensions of X, W, and H are inconsistent.");
    const T eps = std::numeric_limits<T>::epsilon();
    const T delta = std::sqrt(eps);
    T lw = (T)a.lambda_w, lh = (T)a.lambda_h;
    if (a.alg == 1) {  // multupd.jl:37-40
        lw = std::max(lw, delta);
        lh = std::max(lh, delta);
    }
    const T tol = (T)a.tol;
    NMF_REQUIRE(a.alg <= 2 || h->comm == nullptr || a.alg == 3 || a.alg == 4, NMFB200_ENOTSUP, "ALSPGrad is single-GPU in this build");
    // CoordinateDescentUpd (coorddesc.jl:61-80): l1/l2 weights per factor
    T l1W = T(0), l2W = T(0), l1H = T(0), l2H = T(0);
    if (a.alg == 4) {
        const T alpha = (T)a.cd_alpha, ratio = (T)a.cd_l1ratio;
        const T aH = (a.cd_regularization == 0 || a.cd_regularization == 1) ? alpha : T(0);
        const T aW = (a.cd_regularization == 0 || a.cd_regularization == 2) ? alpha : T(0);
        l1W = (T)(aW * ratio);
        l2W = (T)(aW * (T)(T(1) - ratio));
        l1H = (T)(aH * ratio);
        l2H = (T)(aH * (T)(T(1) - ratio));
    }
    ShufflePerm shuffle_rng{a.cd_seed};
    T tolg = (T)a.tolg;
    int64_t sub_iterations = 0;
    int* d_info = (int*)h->buf("simt.info", 16);
    NMF_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), h->stream));

    cudaEvent_t e0, e1, e2;
    NMF_CUDA(cudaEventCreate(&e0));
    NMF_CUDA(cudaEventCreate(&e1));
    NMF_CUDA(cudaEventCreate(&e2));
    NMF_CUDA(cudaEventRecord(e0, st));

    // stage W, H into compact device buffers
    T* W = h->buf_t<T>("simt.W", (size_t)p * k);
    T* H = h->buf_t<T>("simt.H", (size_t)k * n);
    T* preW = h->buf_t<T>("simt.preW", (size_t)p * k);
    T* preH = h->buf_t<T>("simt.preH", (size_t)k * n);
    if (a.on_device) {
        copy2d_kernel<T><<<ew_blocks(p * k), 256, 0, st>>>(W, p, Wc, ldw, p, k);
        copy2d_kernel<T><<<ew_blocks(k * n), 256, 0, st>>>(H, k, Hc, ldh, k, n);
    } else {
        NMF_CUDA(cudaMemcpy2DAsync(W, p * sizeof(T), Wc, ldw * sizeof(T), p * sizeof(T), k, cudaMemcpyHostToDevice, st));
        NMF_CUDA(cudaMemcpy2DAsync(H, k * sizeof(T), Hc, ldh * sizeof(T), k * sizeof(T), n, cudaMemcpyHostToDevice, st));
    }
    unsigned long long* d_updates = (unsigned long long*)h->buf("simt.gcd_updates", 16);
    NMF_CUDA(cudaMemsetAsync(d_updates, 0, sizeof(unsigned long long), st));
    if (sizeof(T) == 4 && a.alg >= 3) {   // tensor-core products (xprod): bf16 caches of X and buffers, built outside the timed loop
        tc_xmul(h, 0, nullptr, 0, 0, k, nullptr, 0, 0);
        tc_xmul(h, 1, nullptr, 0, 0, k, nullptr, 0, 0);
    }
    NMF_CUDA(cudaEventRecord(e1, st));

    double objv = std::numeric_limits<double>::quiet_NaN();
    double t_start = 0;
    auto now = []() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    if (a.verbose) {  // common.jl:54-59
        t_start = now();
        objv = s.objective(a.alg, W, H, lw, lh);
        if (h->trace) h->trace(h->trace_user, 0, 0.0, objv, NAN, NAN);
    }
    bool converged = false;
    int64_t t = 0;
    double dev = 0;
    while (!converged && t < a.maxiter) {  // common.jl:64-83
        ++t;
        NMF_CUDA(cudaMemcpyAsync(preW, W, (size_t)p * k * sizeof(T), cudaMemcpyDeviceToDevice, st));
        NMF_CUDA(cudaMemcpyAsync(preH, H, (size_t)k * n * sizeof(T), cudaMemcpyDeviceToDevice, st));
        if (a.alg == 0) s.iter_multmse(W, H, a.update_H != 0, lw, lh, delta);
        else if (a.alg == 1) s.iter_multdiv(W, H, a.update_H != 0, lw, lh, delta, t == 1);
        else if (a.alg == 2) s.iter_greedycd(W, H, a.update_H != 0, lw, lh, d_updates);
        else if (a.alg == 3) s.iter_projals(W, H, a.update_H != 0, lw, lh, d_info);
        else if (a.alg == 4) s.iter_cd(W, H, a.update_H != 0, l1W, l2W, l1H, l2H, a.cd_shuffle ? &shuffle_rng : nullptr);
        else sub_iterations += s.iter_alspgrad(W, H, a.update_H != 0, a.maxsubiter, &tolg);
        converged = s.stop(W, preW, H, preH, tol, &dev);
        if (a.alg == 3) {  // potrf! failure: the reference ignores info (utils.jl:68,78); we stop with a distinct status
            int info = 0;
            NMF_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
            NMF_CUDA(cudaStreamSynchronize(st));
            NMF_REQUIRE(info == 0, NMFB200_ENUMERIC, "matrix is not positive definite; Cholesky factorization failed (pivot " + std::to_string(info) + ").");
        }
        if (a.verbose) {
            double pre = objv;
            objv = s.objective(a.alg, W, H, lw, lh);
            if (h->trace) h->trace(h->trace_user, t, now() - t_start, objv, objv - pre, dev);
        }
    }
    NMF_CUDA(cudaEventRecord(e2, st));
    if (!a.verbose) objv = s.objective(a.alg, W, H, lw, lh);  // common.jl:85-87

    if (a.on_device) {
        copy2d_kernel<T><<<ew_blocks(p * k), 256, 0, st>>>(Wc, ldw, W, p, p, k);
        copy2d_kernel<T><<<ew_blocks(k * n), 256, 0, st>>>(Hc, ldh, H, k, k, n);
    } else {
        NMF_CUDA(cudaMemcpy2DAsync(Wc, ldw * sizeof(T), W, p * sizeof(T), p * sizeof(T), k, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(Hc, ldh * sizeof(T), H, k * sizeof(T), k * sizeof(T), n, cudaMemcpyDeviceToHost, st));
    }
    unsigned long long upd = 0;
    NMF_CUDA(cudaMemcpyAsync(&upd, d_updates, sizeof(upd), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    float ms_up = 0, ms_loop = 0;
    cudaEventElapsedTime(&ms_up, e0, e1);
    cudaEventElapsedTime(&ms_loop, e1, e2);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    out->niters = t;
    out->converged = converged ? 1 : 0;
    out->engine = s.used_tc ? 1 : 0;   // 1: the X-sized products ran on the tensor cores (split operands); the rest is exact fp32
    out->objvalue = objv;
    out->last_dev = dev;
    out->solve_ms = ms_loop;
    out->upload_ms = ms_up;
    out->coordinate_updates = (int64_t)upd;
    out->kernel_launches = h->launches;
    out->sub_iterations = sub_iterations;
    out->tolg_final = (double)tolg;
}

double simt_objective_f32(nmfb200_handle* h, int alg, const float* W, int64_t ldw, const float* H, int64_t ldh, int64_t k,
                          double lambda_w, double lambda_h) {
    NMF_REQUIRE(alg != 2 || (ldw == h->p && ldh == k), NMFB200_EINVAL, "compact factors required for the L1 terms");
    Simt<float> s{h, h->stream, h->p, h->n, k, (const float*)h->dX, h->ldx};
    return s.objective(alg, W, H, lambda_w, lambda_h, ldw, ldh);
}

// C = X * B or X' * B with host B, C (nmfb200_mul_X_*): the X-sized products of the randomised range finder behind
// NMF.nndsvd (initialization.jl:78), on the resident X.
template <typename T>
void simt_mul_X(nmfb200_handle* h, int transpose_X, const T* B, int64_t ldb, int64_t c, T* C, int64_t ldc) {
    const int64_t rowsB = transpose_X ? h->p : h->n, rowsC = transpose_X ? h->n : h->p;
    NMF_REQUIRE(B != nullptr && C != nullptr, NMFB200_EINVAL, "NULL argument");
    NMF_REQUIRE(c >= 1 && ldb >= rowsB && ldc >= rowsC, NMFB200_EDIM, "Dimensions of X, B and C are inconsistent.");
    NMF_REQUIRE(rowsB <= INT32_MAX && rowsC <= INT32_MAX && c <= INT32_MAX, NMFB200_EDIM, "dimension exceeds 2^31-1");
    cudaStream_t st = h->stream;
    T* dB = h->buf_t<T>("mulx.B", (size_t)rowsB * c);
    T* dC = h->buf_t<T>("mulx.C", (size_t)rowsC * c);
    NMF_CUDA(cudaMemcpy2DAsync(dB, rowsB * sizeof(T), B, ldb * sizeof(T), rowsB * sizeof(T), c, cudaMemcpyHostToDevice, st));
    Simt<T> s{h, st, h->p, h->n, c, (const T*)h->dX, h->ldx};
    if (transpose_X) s.gemm((int)h->n, (int)c, (int)h->p, s.X, h->ldx, 1, dB, 1, rowsB, dC, 1, rowsC);
    else s.gemm((int)h->p, (int)c, (int)h->n, s.X, 1, h->ldx, dB, 1, rowsB, dC, 1, rowsC);
    NMF_CUDA(cudaMemcpy2DAsync(C, ldc * sizeof(T), dC, rowsC * sizeof(T), rowsC * sizeof(T), c, cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
}
template void simt_mul_X<float>(nmfb200_handle*, int, const float*, int64_t, int64_t, float*, int64_t);
template void simt_mul_X<double>(nmfb200_handle*, int, const double*, int64_t, int64_t, double*, int64_t);

template void simt_solve<float>(nmfb200_handle*, const SolveArgs&, float*, int64_t, float*, int64_t, nmfb200_result*);
template void simt_solve<double>(nmfb200_handle*, const SolveArgs&, double*, int64_t, double*, int64_t, nmfb200_result*);

}  // namespace nmfb200
