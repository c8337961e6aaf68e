// init_device.cuh -- NMF.nndsvd on the device: RandomizedLinAlg.rsvd as called at initialization.jl:78
// (`Q = qr(X * randn(n, k)).Q; svd(Q' * X)`; no oversampling, no power iterations) and `_nndsvd!` (initialization.jl:26-68),
// so that nnmf(X, k) with the default init=:nndsvdar never leaves the GPU between set_X and the first iteration.
// Included by simt_engine.cu inside namespace nmfb200 { namespace { ... } }.
//
//   Omega = randn(n, k)      counter-based Philox + Box-Muller (philox.cuh): a host regenerates it (tests/test_gpu_init.py)
//   Y = X * Omega            the solver's X-sized product (tcgen05 mainloop with split operands for large Float32 problems)
//   Q = qr(Y).Q              CholeskyQR2 in Float64: G = Y'Y, G = R'R, Q = Y R^-1, twice.  Needs cond(Y) < ~1e6; a pivot that
//                            collapses (rank-deficient sample: k > rank(X)) returns NMFB200_ENUMERIC and the host layer falls
//                            back to LAPACK's Householder QR, as the reference uses
//   B' = X' * Q              the second X-sized product
//   svd(B)                   one-sided Jacobi (Hestenes) on the n x k matrix B' in Float64: column pairs are rotated until
//                            mutually orthogonal, B' J = V S with J = U_B accumulated; round-robin ordering, k/2 independent
//                            pairs per launch; high relative accuracy, no k x k Gram squaring of the small singular values
//   U = Q * U_B, sorted by decreasing singular value
//   W, H                     _nndsvd!: positive / negative parts per singular pair, one CTA per component
// Singular vectors are defined up to a joint sign of (u_j, v_j); _nndsvd! is invariant to it (the roles of the positive and negative
// parts swap).  Everything here is O((p + n) k^2) next to the two products with X.
#pragma once
// (philox.cuh is included by simt_engine.cu at file scope)

template <typename T>
__global__ void philox_normal_fill_kernel(T* __restrict__ A, int64_t len, uint32_t stream, uint64_t seed) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < len; t += (int64_t)gridDim.x * blockDim.x)
        A[t] = (T)philox_normal((uint64_t)t, stream, seed);
}

template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t len) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < len; t += (int64_t)gridDim.x * blockDim.x) dst[t] = (TD)src[t];
}

// G (k x k, column-major, Float64, symmetric positive definite) -> Rinv = R^-1 with G = R'R (R upper triangular).  One CTA.
// info[0] = j + 1 if the j-th pivot is not safely positive (below 1e-12 of the original diagonal entry: column j of the sample depends
// on the columns before it to ~1e-6 -- CholeskyQR cannot orthogonalise that).  G is overwritten with R.
__global__ void __launch_bounds__(1024) chol_rinv_kernel(double* __restrict__ G, int k, double* __restrict__ Rinv, int* __restrict__ info) {
    __shared__ int bad;
    __shared__ double rjj;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        if (threadIdx.x == 0) {
            const double d = G[(size_t)j * k + j];
            const double d0 = Rinv[(size_t)j * k + j];   // the caller stored the original diagonal here
            if (!(d > 1e-12 * d0) || !(d0 > 0.0)) { bad = 1; info[0] = j + 1; }
            rjj = sqrt(d);
        }
        __syncthreads();
        if (bad) return;
        const double r = rjj;
        for (int c = j + threadIdx.x; c < k; c += blockDim.x) G[(size_t)c * k + j] = G[(size_t)c * k + j] / r;   // row j of R: R(j, c) = G(j, c) / r
        __syncthreads();
        // trailing update: G(a, b) -= R(j, a) R(j, b) for j < a <= b
        const int m = k - j - 1;
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int a = j + 1 + e % m, b = j + 1 + e / m;
            if (a <= b) G[(size_t)b * k + a] -= G[(size_t)a * k + j] * G[(size_t)b * k + j];
        }
        __syncthreads();
    }
    // R^-1 by back substitution, one column per thread (R(i, j) is read at the same address by all threads: broadcast)
    for (int c = threadIdx.x; c < k; c += blockDim.x) {
        double* x = Rinv + (size_t)c * k;
        for (int i = c + 1; i < k; ++i) x[i] = 0.0;
        x[c] = 1.0 / G[(size_t)c * k + c];
        for (int i = c - 1; i >= 0; --i) {
            double s = 0.0;
            for (int jj = i + 1; jj <= c; ++jj) s += G[(size_t)jj * k + i] * x[jj];
            x[i] = -s / G[(size_t)i * k + i];
        }
    }
}
__global__ void save_diag_kernel(const double* __restrict__ G, int k, double* __restrict__ Rinv) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) Rinv[(size_t)j * k + j] = G[(size_t)j * k + j];
}
__global__ void set_identity_kernel(double* __restrict__ J, int k) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * k; e += gridDim.x * blockDim.x) J[e] = (e % k == e / k) ? 1.0 : 0.0;
}

// One round of a one-sided Jacobi sweep over the columns of A (n x k, column-major, ld n): block b owns the pair (i, j) that the
// round-robin ("circle") schedule assigns to it in this round -- the k/2 pairs of a round are disjoint, so the blocks are independent.
// kp = k rounded up to even (a pair with the padding index is a bye).  J (k x k) accumulates the rotations.
__global__ void __launch_bounds__(256) jacobi_round_kernel(double* __restrict__ A, int64_t n, double* __restrict__ J, int k, int kp, int round,
                                                           double tol, unsigned int* __restrict__ rotations) {
    __shared__ double red[3][256];
    __shared__ double cs[2];
    __shared__ int rotate;
    const int b = blockIdx.x, last = kp - 1;
    int i, j;
    if (b == 0) { i = round % last; j = last; }
    else { i = (round + b) % last; j = (round - b + last) % last; }
    if (i > j) { const int t = i; i = j; j = t; }
    if (j >= k) return;
    double* ai = A + (size_t)i * n;
    double* aj = A + (size_t)j * n;
    double alpha = 0.0, beta = 0.0, gamma = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += 256) {
        const double x = ai[r], y = aj[r];
        alpha += x * x;
        beta += y * y;
        gamma += x * y;
    }
    red[0][threadIdx.x] = alpha;
    red[1][threadIdx.x] = beta;
    red[2][threadIdx.x] = gamma;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        alpha = red[0][0]; beta = red[1][0]; gamma = red[2][0];
        const bool rot = fabs(gamma) > tol * sqrt(alpha * beta) && gamma != 0.0;
        rotate = rot ? 1 : 0;
        if (rot) {
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t);
            cs[0] = c;
            cs[1] = c * t;
            atomicAdd(rotations, 1u);
        }
    }
    __syncthreads();
    if (!rotate) return;
    const double c = cs[0], s = cs[1];
    for (int64_t r = threadIdx.x; r < n; r += 256) {
        const double x = ai[r], y = aj[r];
        ai[r] = c * x - s * y;
        aj[r] = s * x + c * y;
    }
    double* ji = J + (size_t)i * k;
    double* jj = J + (size_t)j * k;
    for (int r = threadIdx.x; r < k; r += 256) {
        const double x = ji[r], y = jj[r];
        ji[r] = c * x - s * y;
        jj[r] = s * x + c * y;
    }
}

__global__ void __launch_bounds__(256) colnorm_kernel(const double* __restrict__ A, int64_t n, double* __restrict__ out) {
    __shared__ double red[256];
    const double* a = A + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += 256) s += a[r] * a[r];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sqrt(red[0]);
}

// dst(:, jj) = src(:, perm[jj]) * (scale ? 1 / scale[perm[jj]] : 1)   (rows x k, column-major, ld rows); a zero scale gives a zero column
template <typename TD>
__global__ void gather_cols_kernel(const double* __restrict__ src, int64_t rows, int k, const int* __restrict__ perm, const double* __restrict__ scale,
                                   TD* __restrict__ dst) {
    const int64_t total = rows * k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t % rows;
        const int jj = (int)(t / rows), sj = perm[jj];
        double v = src[r + (size_t)sj * rows];
        if (scale != nullptr) v = scale[sj] > 0.0 ? v / scale[sj] : 0.0;
        dst[t] = (TD)v;
    }
}

// sum of the entries of X (p x n, ld ldx) in Float64: mean(X) of initialization.jl:37
template <typename T>
__global__ void __launch_bounds__(256) sum_entries_kernel(const T* __restrict__ X, int64_t p, int64_t n, int64_t ldx, double* __restrict__ part) {
    __shared__ double red[256];
    double s = 0.0;
    const int64_t total = p * n;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) s += (double)X[(t % p) + (t / p) * ldx];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// _nndsvd! (initialization.jl:26-68): one CTA per component j.  U p x k, V n x k (column-major), S [k]; W p x k (ldw), H k x n (ldh).
// fill[j] = v_j: 0 (:nndsvd), mean(X) (:nndsvda), mean(X) * 0.01 * rand (:nndsvdar).  posnegnorm sums in Float64 here (the reference
// sums sequentially in T) and rounds to T before the square root.
template <typename T>
__global__ void __launch_bounds__(256) nndsvd_split_kernel(const T* __restrict__ U, int64_t p, const T* __restrict__ V, int64_t n,
                                                           const T* __restrict__ S, T* __restrict__ W, int64_t ldw, T* __restrict__ H, int64_t ldh,
                                                           int inith, const T* __restrict__ fill) {
    __shared__ double red[4][256];
    const int j = blockIdx.x;
    const T* x = U + (size_t)j * p;
    const T* y = V + (size_t)j * n;
    double xp = 0.0, xn = 0.0, yp = 0.0, yn = 0.0;
    for (int64_t i = threadIdx.x; i < p; i += 256) { const double v = (double)x[i]; if (x[i] > T(0)) xp += v * v; else xn += v * v; }
    for (int64_t i = threadIdx.x; i < n; i += 256) { const double v = (double)y[i]; if (y[i] > T(0)) yp += v * v; else yn += v * v; }
    red[0][threadIdx.x] = xp; red[1][threadIdx.x] = xn; red[2][threadIdx.x] = yp; red[3][threadIdx.x] = yn;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int q = 0; q < 4; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + o];
        __syncthreads();
    }
    const T xpnrm = sqrt((T)red[0][0]), xnnrm = sqrt((T)red[1][0]), ypnrm = sqrt((T)red[2][0]), ynnrm = sqrt((T)red[3][0]);
    const T mp = xpnrm * ypnrm, mn = xnnrm * ynnrm;
    const T vj = fill[j];
    const bool pos = mp >= mn;
    const T ss = sqrt(S[j] * (pos ? mp : mn));
    const T cx = ss / (pos ? xpnrm : xnnrm), cy = ss / (pos ? ypnrm : ynnrm);
    for (int64_t i = threadIdx.x; i < p; i += 256) {
        const T xi = x[i];
        W[i + (size_t)j * ldw] = pos ? (xi > T(0) ? xi * cx : vj) : (xi < T(0) ? -(xi * cx) : vj);   // scalepos! / scaleneg!
    }
    if (inith) {
        for (int64_t i = threadIdx.x; i < n; i += 256) {
            const T yi = y[i];
            H[j + (size_t)i * ldh] = pos ? (yi > T(0) ? yi * cy : vj) : (yi < T(0) ? -(yi * cy) : vj);
        }
    }
}

template <typename T>
__global__ void nndsvd_fill_kernel(T* __restrict__ fill, int k, int variant, double mean, uint64_t seed) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
        T v0 = variant == 0 ? T(0) : (variant == 1 ? (T)mean : (T)(mean * 0.01));   // initialization.jl:36-37
        if (variant == 2) v0 *= philox_uniform<T>((uint64_t)j, 3u, seed);             // vj *= rand(T), :48-51
        fill[j] = v0;
    }
}

// rsvd(X, k) on the resident X.  Leaves U (p x k), S (k), V (n x k) in the handle's buffers init.U / init.S / init.V (element type T).
template <typename T>
void device_rsvd(nmfb200_handle* h, int64_t k, uint64_t seed, T** U_out, T** S_out, T** V_out) {
    NMF_REQUIRE(h->x_elt == (int)sizeof(T), NMFB200_ESTATE, h->x_elt ? "X was set with a different element type" : "nmfb200_set_X must precede rsvd");
    const int64_t p = h->p, n = h->n;
    NMF_REQUIRE(k >= 1 && k <= std::min(p, n), NMFB200_EINVAL, "The value of k should not exceed min(size(X)).");
    NMF_REQUIRE(k <= 1024 && p <= INT32_MAX && n <= INT32_MAX, NMFB200_ENOTSUP, "device rsvd: k <= 1024");
    NMF_REQUIRE(h->comm == nullptr, NMFB200_ENOTSUP, "device rsvd runs on one GPU (row-sharded handles initialise on the host)");
    cudaStream_t st = h->stream;
    const int ik = (int)k;
    Simt<T> s{h, st, p, n, k, (const T*)h->dX, h->ldx};
    Simt<double> sd{h, st, p, n, k, nullptr, 0};
    T* Omega = h->buf_t<T>("init.Omega", (size_t)n * k);
    T* Y = h->buf_t<T>("init.Y", (size_t)p * k);
    T* Bt = h->buf_t<T>("init.Bt", (size_t)n * k);
    double* Yd = h->buf_t<double>("init.Yd", (size_t)p * k);
    double* Qd = h->buf_t<double>("init.Qd", (size_t)p * k);
    double* Ad = h->buf_t<double>("init.Ad", (size_t)n * k);
    double* G = h->buf_t<double>("init.G", (size_t)k * k);
    double* Rinv = h->buf_t<double>("init.Rinv", (size_t)k * k);
    double* J = h->buf_t<double>("init.J", (size_t)k * k);
    double* Jp = h->buf_t<double>("init.Jp", (size_t)k * k);
    double* sig = h->buf_t<double>("init.sig", (size_t)k);
    int* perm = h->buf_t<int>("init.perm", (size_t)k);
    int* info = (int*)h->buf("init.info", 16);
    unsigned int* rot = (unsigned int*)(info + 1);
    T* U = h->buf_t<T>("init.U", (size_t)p * k);
    T* S = h->buf_t<T>("init.S", (size_t)k);
    T* V = h->buf_t<T>("init.V", (size_t)n * k);
    const int g = 148 * 8;

    philox_normal_fill_kernel<T><<<g, 256, 0, st>>>(Omega, n * k, 2u, seed);
    h->launches += 1;
    s.xprod(1, Omega, 1, n, Y, 1, p);                                      // Y = X * Omega
    convert_kernel<T, double><<<g, 256, 0, st>>>(Y, Yd, p * k);
    h->launches += 1;
    NMF_CUDA(cudaMemsetAsync(info, 0, 16, st));
    double* src = Yd;
    double* dst = Qd;
    for (int pass = 0; pass < 2; ++pass) {                                 // CholeskyQR2
        sd.gemm(ik, ik, (int)p, src, p, 1, src, 1, p, G, 1, k);            // G = Y'Y
        save_diag_kernel<<<1, 256, 0, st>>>(G, ik, Rinv);
        chol_rinv_kernel<<<1, 1024, 0, st>>>(G, ik, Rinv, info);
        sd.gemm((int)p, ik, ik, src, 1, p, Rinv, 1, k, dst, 1, p);         // Q = Y R^-1
        h->launches += 2;
        std::swap(src, dst);
    }
    double* Q = src;                                                       // the orthonormal basis (Float64)
    double* Ud = dst;                                                      // the other Float64 buffer: U = Q * U_B goes there
    int hinfo = 0;
    NMF_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    NMF_REQUIRE(hinfo == 0, NMFB200_ENUMERIC,
                "device rsvd: the sample X * randn(n, k) is numerically rank deficient (column " + std::to_string(hinfo) +
                    "): CholeskyQR cannot orthogonalise it -- use the host QR");
    T* Qt = Y;                                                             // Q in the element type of X, for the second product
    convert_kernel<double, T><<<g, 256, 0, st>>>(Q, Qt, p * k);
    h->launches += 1;
    s.xprod(0, Qt, 1, p, Bt, 1, n);                                        // B' = X' * Q  (n x k)
    convert_kernel<T, double><<<g, 256, 0, st>>>(Bt, Ad, n * k);
    set_identity_kernel<<<64, 256, 0, st>>>(J, ik);
    h->launches += 2;
    const int kp = ik + (ik & 1);
    if (kp >= 2) {
        bool done = false;
        for (int sweep = 0; sweep < 40 && !done; ++sweep) {
            NMF_CUDA(cudaMemsetAsync(rot, 0, sizeof(unsigned int), st));
            for (int round = 0; round < kp - 1; ++round) jacobi_round_kernel<<<kp / 2, 256, 0, st>>>(Ad, n, J, ik, kp, round, 1e-14, rot);
            h->launches += kp - 1;
            unsigned int hrot = 0;
            NMF_CUDA(cudaMemcpyAsync(&hrot, rot, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            NMF_CUDA(cudaStreamSynchronize(st));
            done = hrot == 0;
        }
        NMF_REQUIRE(done, NMFB200_ENUMERIC, "device rsvd: the Jacobi sweeps did not converge");
    }
    colnorm_kernel<<<ik, 256, 0, st>>>(Ad, n, sig);
    h->launches += 1;
    std::vector<double> hsig((size_t)k);
    NMF_CUDA(cudaMemcpyAsync(hsig.data(), sig, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
    std::vector<int> hperm((size_t)k);
    for (int j = 0; j < ik; ++j) hperm[j] = j;
    std::stable_sort(hperm.begin(), hperm.end(), [&](int a, int b) { return hsig[a] > hsig[b]; });
    std::vector<T> hS((size_t)k);
    for (int j = 0; j < ik; ++j) hS[j] = (T)hsig[hperm[j]];
    NMF_CUDA(cudaMemcpyAsync(perm, hperm.data(), (size_t)k * sizeof(int), cudaMemcpyHostToDevice, st));
    NMF_CUDA(cudaMemcpyAsync(S, hS.data(), (size_t)k * sizeof(T), cudaMemcpyHostToDevice, st));
    gather_cols_kernel<T><<<g, 256, 0, st>>>(Ad, n, ik, perm, sig, V);                    // V = B' J / sigma, sorted
    gather_cols_kernel<double><<<64, 256, 0, st>>>(J, k, ik, perm, nullptr, Jp);          // U_B, sorted
    sd.gemm((int)p, ik, ik, Q, 1, p, Jp, 1, k, Ud, 1, p);                                  // U = Q * U_B
    convert_kernel<double, T><<<g, 256, 0, st>>>(Ud, U, p * k);
    h->launches += 3;
    NMF_CUDA(cudaGetLastError());
    NMF_CUDA(cudaStreamSynchronize(st));   // hperm / hS are stack-owned host buffers
    *U_out = U;
    *S_out = S;
    *V_out = V;
}

template <typename T>
void rsvd_impl(nmfb200_handle* h, int64_t k, uint64_t seed, T* U, int64_t ldu, T* S, T* V, int64_t ldv) {
    NMF_REQUIRE(U != nullptr && S != nullptr && V != nullptr, NMFB200_EINVAL, "NULL argument");
    NMF_REQUIRE(ldu >= h->p && ldv >= h->n, NMFB200_EDIM, "inconsistent dimensions");
    T *dU, *dS, *dV;
    device_rsvd<T>(h, k, seed, &dU, &dS, &dV);
    cudaStream_t st = h->stream;
    NMF_CUDA(cudaMemcpy2DAsync(U, ldu * sizeof(T), dU, h->p * sizeof(T), h->p * sizeof(T), k, cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaMemcpy2DAsync(V, ldv * sizeof(T), dV, h->n * sizeof(T), h->n * sizeof(T), k, cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaMemcpyAsync(S, dS, k * sizeof(T), cudaMemcpyDeviceToHost, st));
    NMF_CUDA(cudaStreamSynchronize(st));
}

// NMF.nndsvd(X, k; zeroh, variant) (initialization.jl:70-101) without initdata.  variant: 0 :std, 1 :a, 2 :ar.
template <typename T>
void nndsvd_impl(nmfb200_handle* h, T* W, int64_t ldw, T* H, int64_t ldh, int64_t k, int variant, int zeroh, uint64_t seed, int on_device) {
    NMF_REQUIRE(W != nullptr && H != nullptr, NMFB200_EINVAL, "NULL argument");
    NMF_REQUIRE(variant >= 0 && variant <= 2, NMFB200_EINVAL, "Invalid value for variant");   // initialization.jl:77
    NMF_REQUIRE(h->x_elt != 0, NMFB200_ESTATE, "nmfb200_set_X must precede nndsvd");
    const int64_t p = h->p, n = h->n;
    NMF_REQUIRE(ldw >= p && ldh >= k, NMFB200_EDIM, "inconsistent dimensions");
    T *dU, *dS, *dV;
    device_rsvd<T>(h, k, seed, &dU, &dS, &dV);
    cudaStream_t st = h->stream;
    double mean = 0.0;
    if (variant != 0) {
        const int nb = 148 * 8;
        double* part = h->buf_t<double>("init.mean_part", nb);
        sum_entries_kernel<T><<<nb, 256, 0, st>>>((const T*)h->dX, p, n, h->ldx, part);
        h->launches += 1;
        std::vector<double> hp(nb);
        NMF_CUDA(cudaMemcpyAsync(hp.data(), part, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaStreamSynchronize(st));
        for (double v : hp) mean += v;
        mean /= (double)p * (double)n;
    }
    T* fill = h->buf_t<T>("init.fill", (size_t)k);
    nndsvd_fill_kernel<T><<<1, 256, 0, st>>>(fill, (int)k, variant, mean, seed);
    T* dW = on_device ? W : h->buf_t<T>("init.W", (size_t)p * k);
    T* dH = on_device ? H : h->buf_t<T>("init.H", (size_t)k * n);
    const int64_t lw = on_device ? ldw : p, lh = on_device ? ldh : k;
    if (zeroh) NMF_CUDA(cudaMemset2DAsync(dH, lh * sizeof(T), 0, k * sizeof(T), n, st));   // fill!(H, 0), initialization.jl:88
    nndsvd_split_kernel<T><<<(unsigned)k, 256, 0, st>>>(dU, p, dV, n, dS, dW, lw, dH, lh, zeroh ? 0 : 1, fill);
    h->launches += 2;
    NMF_CUDA(cudaGetLastError());
    if (!on_device) {
        NMF_CUDA(cudaMemcpy2DAsync(W, ldw * sizeof(T), dW, p * sizeof(T), p * sizeof(T), k, cudaMemcpyDeviceToHost, st));
        NMF_CUDA(cudaMemcpy2DAsync(H, ldh * sizeof(T), dH, k * sizeof(T), k * sizeof(T), n, cudaMemcpyDeviceToHost, st));
    }
    NMF_CUDA(cudaStreamSynchronize(st));
}
