/*
 * oracle_kernels.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C restatement of the scalar loops of JuliaStats/NMF.jl @ 2eed3ec (v1.0.3) that sit on the
 * per-iteration hot path.  Every function cites the reference lines it follows.  The GEMMs of the
 * reference (LinearAlgebra.mul! -> OpenBLAS) are done by NumPy/OpenBLAS in nmf_oracle.py; the loops
 * here are the hand-written Julia loops, kept *sequential and un-fused* (compile with
 * -ffp-contract=off) so the accumulation order is the one Julia executes.
 *
 * PARITY STATUS: Julia is not installable in the build image, so the oracle cannot be compared
 * bit-for-bit with a run of the reference ("parity unpinned" at bit level).  It is pinned against
 * every fixture the reference's own tests hold for this path (test/testproblems.jl laurberg6x3,
 * test/multupd.jl, test/greedycd.jl, test/coorddesc.jl, test/alspgrad.jl, test/interf.jl) -- see
 * tests/test_oracle.py.
 *
 * Conventions: all matrices column-major (Julia layout), 0-based indices here, 1-based in the
 * citations.  `ORACLE_T` is instantiated for float and double through the macro block at the end.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <float.h>

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)

#define DEFINE_ORACLE(T, SFX, EPS_T)                                                                   \
                                                                                                       \
/* src/multupd.jl:101-103 and :112-114 -- F[i] *= max(0, Num[i]-lambda) / (Den[i]+delta) */            \
void CAT(oracle_mu_mse_ratio, SFX)(T* F, const T* Num, const T* Den, int64_t len, T lambda, T delta) { \
    for (int64_t i = 0; i < len; ++i) {                                                                \
        T num = Num[i] - lambda;                                                                       \
        if (!(num > (T)0)) num = (num != num) ? num : (T)0; /* Julia max(0,NaN) = NaN */               \
        F[i] = F[i] * (num / (Den[i] + delta));                                                        \
    }                                                                                                  \
}                                                                                                      \
                                                                                                       \
/* src/multupd.jl:172-174 / :184-186 -- Q[i] = X[i] / (WH[i] + delta) */                               \
void CAT(oracle_mu_div_quot, SFX)(T* Q, const T* X, const T* WH, int64_t len, T delta) {               \
    for (int64_t i = 0; i < len; ++i) Q[i] = X[i] / (WH[i] + delta);                                   \
}                                                                                                      \
                                                                                                       \
/* src/multupd.jl:177-179 -- H[i,j] *= WtQ[i,j] / (sW[i] + lambda_h); H is k x n col-major */          \
void CAT(oracle_mu_div_scale_h, SFX)(T* H, const T* WtQ, const T* sW, int64_t k, int64_t n, T lam) {   \
    for (int64_t j = 0; j < n; ++j)                                                                    \
        for (int64_t i = 0; i < k; ++i) H[i + j * k] = H[i + j * k] * (WtQ[i + j * k] / (sW[i] + lam));\
}                                                                                                      \
                                                                                                       \
/* src/multupd.jl:189-191 -- W[i,j] *= QHt[i,j] / (sH[j] + lambda_w); W is p x k col-major */          \
void CAT(oracle_mu_div_scale_w, SFX)(T* W, const T* QHt, const T* sH, int64_t p, int64_t k, T lam) {   \
    for (int64_t j = 0; j < k; ++j)                                                                    \
        for (int64_t i = 0; i < p; ++i) W[i + j * p] = W[i + j * p] * (QHt[i + j * p] / (sH[j] + lam));\
}                                                                                                      \
                                                                                                       \
/* Julia's max(a,b): NaN if either is NaN */                                                           \
static T CAT(jl_max, SFX)(T a, T b) {                                                                  \
    if (a != a) return a;                                                                              \
    if (b != b) return b;                                                                              \
    return a > b ? a : b;                                                                              \
}                                                                                                      \
                                                                                                       \
/* src/common.jl:92-111 stop_condition.  W,preW p x k; H,preH k x n.  Returns converged (0/1) and      \
 * writes devmax (partial on early exit, exactly as the reference).  Accumulators are T. */            \
int CAT(oracle_stop_condition, SFX)(const T* W, const T* preW, const T* H, const T* preH,              \
                                    int64_t p, int64_t n, int64_t k, T eps, T* devmax_out) {           \
    T devmax = (T)0;                                                                                   \
    for (int64_t j = 0; j < k; ++j) {                                                                  \
        T dev_w = (T)0, sum_w = (T)0;                                                                  \
        for (int64_t i = 0; i < p; ++i) {                                                              \
            T d = W[i + j * p] - preW[i + j * p];                                                      \
            T s = W[i + j * p] + preW[i + j * p];                                                      \
            dev_w += d * d;                                                                            \
            sum_w += s * s;                                                                            \
        }                                                                                              \
        T dev_h = (T)0, sum_h = (T)0;                                                                  \
        for (int64_t i = 0; i < n; ++i) {                                                              \
            T d = H[j + i * k] - preH[j + i * k];                                                      \
            T s = H[j + i * k] + preH[j + i * k];                                                      \
            dev_h += d * d;                                                                            \
            sum_h += s * s;                                                                            \
        }                                                                                              \
        T m = CAT(jl_max, SFX)(dev_w / sum_w, dev_h / sum_h);                                          \
        devmax = CAT(jl_max, SFX)(devmax, (T)sqrt((double)m));                                         \
        /* tol is stored as a field of type T (multupd.jl:13, greedycd.jl:13) => comparison in T */    \
        if ((T)sqrt((double)dev_w) > (T)(eps * (T)sqrt((double)sum_w)) ||                              \
            (T)sqrt((double)dev_h) > (T)(eps * (T)sqrt((double)sum_h))) {                              \
            *devmax_out = devmax;                                                                      \
            return 0;                                                                                  \
        }                                                                                              \
    }                                                                                                  \
    *devmax_out = devmax;                                                                              \
    return 1;                                                                                          \
}                                                                                                      \
                                                                                                       \
/* StatsBase.sqL2dist (third-party, not vendored; Project.toml compat 0.25-0.34; deviation.jl):        \
 * r = 0.0 (Float64); r += abs2(a[i]-b[i]) with the difference formed in T.  Call sites:               \
 * src/multupd.jl:81, src/greedycd.jl:84. */                                                           \
double CAT(oracle_sql2dist, SFX)(const T* a, const T* b, int64_t len) {                                \
    double r = 0.0;                                                                                    \
    for (int64_t i = 0; i < len; ++i) {                                                                \
        T d = a[i] - b[i];                                                                             \
        r += (double)(T)(d * d);                                                                       \
    }                                                                                                  \
    return r;                                                                                          \
}                                                                                                      \
                                                                                                       \
/* StatsBase.gkldiv (same package/file): r += a>0 ? a*log(a/b) - a + b : b, Float64 accumulator.       \
 * Call site: src/multupd.jl:148. */                                                                   \
double CAT(oracle_gkldiv, SFX)(const T* a, const T* b, int64_t len) {                                  \
    double r = 0.0;                                                                                    \
    for (int64_t i = 0; i < len; ++i) {                                                                \
        T ai = a[i], bi = b[i];                                                                        \
        if (ai > (T)0) {                                                                               \
            T t = (T)(ai * (T)log((double)(T)(ai / bi)));                                              \
            t = (T)(t - ai);                                                                           \
            t = (T)(t + bi);                                                                           \
            r += (double)t;                                                                            \
        } else {                                                                                       \
            r += (double)bi;                                                                           \
        }                                                                                              \
    }                                                                                                  \
    return r;                                                                                          \
}                                                                                                      \
                                                                                                       \
/* first-max argmax over a strided row, Julia `argmax(view(D, i, :))` (NaN handling not needed:        \
 * D is finite on this path) */                                                                        \
static int64_t CAT(argmax_row, SFX)(const T* D, int64_t i, int64_t rows, int64_t k) {                  \
    int64_t q = 0;                                                                                     \
    T best = D[i];                                                                                     \
    for (int64_t r = 1; r < k; ++r) {                                                                  \
        T v = D[i + r * rows];                                                                         \
        if (v > best) { best = v; q = r; }                                                             \
    }                                                                                                  \
    return q;                                                                                          \
}                                                                                                      \
                                                                                                       \
/* src/greedycd.jl:125-165 -- everything of _update_GreedyCD! after G = F*P - Z (+lambda) is formed.   \
 * F (rows x k, col-major, the factor being updated, "W" in the reference), G rows x k (in/out),       \
 * P k x k, scratch S,D,Fnew rows x k, q rows.  Returns the number of coordinate updates performed     \
 * (not a reference output; used for throughput reporting). */                                         \
int64_t CAT(oracle_greedycd_rows, SFX)(T* F, T* G, const T* P, T* S, T* D, T* Fnew, int64_t* q,        \
                                       int64_t rows, int64_t k) {                                      \
    const T epsT = (T)EPS_T;                                                                           \
    for (int64_t r = 0; r < k; ++r) {                          /* :125-130 */                          \
        T prr = P[r + r * k];                                                                          \
        for (int64_t i = 0; i < rows; ++i) {                                                           \
            T w = F[i + r * rows], g = G[i + r * rows];                                                \
            T t = w - g / (epsT + prr);                                                                \
            T s = (t > (T)0 ? t : (T)0) - w;                                                           \
            S[i + r * rows] = s;                                                                       \
            D[i + r * rows] = -g * s - (T)0.5 * prr * (s * s);                                         \
        }                                                                                              \
    }                                                                                                  \
    T p_init = (T)-1.0;                                        /* :132-137 */                          \
    for (int64_t i = 0; i < rows; ++i) {                                                               \
        int64_t qi = CAT(argmax_row, SFX)(D, i, rows, k);                                              \
        q[i] = qi;                                                                                     \
        T v = D[i + qi * rows];                                                                        \
        if (v > p_init) p_init = v;                                                                    \
    }                                                                                                  \
    for (int64_t i = 0; i < rows * k; ++i) Fnew[i] = (T)0;     /* :139 */                              \
    const T nu = (T)0.001;                                     /* :140 */                              \
    int64_t updates = 0;                                                                               \
    for (int64_t i = 0; i < rows; ++i) {                       /* :142-163 */                          \
        int64_t qi = q[i];                                                                             \
        for (int64_t it = 0; it < k * k; ++it) {                                                       \
            if (D[i + qi * rows] < nu * p_init) break;                                                 \
            T sq = S[i + qi * rows];                                                                   \
            Fnew[i + qi * rows] += sq;                                                                 \
            for (int64_t r = 0; r < k; ++r) G[i + r * rows] += sq * P[qi + r * k];                     \
            for (int64_t r = 0; r < k; ++r) {                                                          \
                T prr = P[r + r * k];                                                                  \
                T w = F[i + r * rows], g = G[i + r * rows];                                            \
                T t = w - g / (epsT + prr);                                                            \
                T s = (t > (T)0 ? t : (T)0) - w;                                                       \
                S[i + r * rows] = s;                                                                   \
                D[i + r * rows] = -g * s - (T)0.5 * prr * (s * s);                                     \
            }                                                                                          \
            qi = CAT(argmax_row, SFX)(D, i, rows, k);                                                  \
            ++updates;                                                                                 \
        }                                                                                              \
        q[i] = qi;                                                                                     \
    }                                                                                                  \
    for (int64_t i = 0; i < rows * k; ++i) {                   /* :164-165 W + Wnew, projectnn! */     \
        T v = F[i] + Fnew[i];                                                                          \
        if (v < (T)0) v = (T)0;                                /* utils.jl:34-41: NaN preserved */     \
        F[i] = v;                                                                                      \
    }                                                                                                  \
    return updates;                                                                                    \
}                                                                                                      \
                                                                                                       \
/* src/utils.jl:34-41 projectnn!: negatives to zero (NaN < 0 is false => NaN preserved) */             \
void CAT(oracle_projectnn, SFX)(T* A, int64_t len) {                                                   \
    for (int64_t i = 0; i < len; ++i)                                                                  \
        if (A[i] < (T)0) A[i] = (T)0;                                                                  \
}                                                                                                      \
                                                                                                       \
/* src/coorddesc.jl:138-157 -- the sweep of _update_coord_descent! after HHt (+l2 on the diagonal)     \
 * and XHt (-l1) are formed.  F rows x k col-major ("W" in the reference, updated in place), HHt       \
 * k x k, XHt rows x k, perm = the component order (0-based; 1:k or randperm, :131-135).  Components   \
 * outer, rows inner, gradient accumulated sequentially starting from -XHt[i,t].  Returns the          \
 * violation sum (:150; stored in the state, never used for stopping). */                              \
T CAT(oracle_cd_sweep, SFX)(T* F, const T* HHt, const T* XHt, int64_t rows, int64_t k,                 \
                            const int64_t* perm) {                                                     \
    T violation = (T)0;                                                                                \
    for (int64_t tt = 0; tt < k; ++tt) {                                                               \
        int64_t t = perm[tt];                                                                          \
        for (int64_t i = 0; i < rows; ++i) {                                                           \
            T grad = -XHt[i + t * rows];                                                               \
            for (int64_t r = 0; r < k; ++r) grad += HHt[t + r * k] * F[i + r * rows];                  \
            T pg = (F[i + t * rows] == (T)0) ? (grad < (T)0 ? grad : (T)0) : grad;                     \
            violation += (T)fabs((double)pg);                                                          \
            T hess = HHt[t + t * k];                                                                   \
            if (hess != (T)0) {                                                                        \
                T v = F[i + t * rows] - grad / hess;                                                   \
                F[i + t * rows] = CAT(jl_max, SFX)(v, (T)0);                                           \
            }                                                                                          \
        }                                                                                              \
    }                                                                                                  \
    return violation;                                                                                  \
}                                                                                                      \
                                                                                                       \
/* src/alspgrad.jl:9-19 projgradnorm: sqrt of the sequential sum (in T) of g_i^2 over entries with     \
 * g_i < 0 or x_i > 0 */                                                                               \
T CAT(oracle_projgradnorm, SFX)(const T* g, const T* x, int64_t len) {                                 \
    T v = (T)0;                                                                                        \
    for (int64_t i = 0; i < len; ++i) {                                                                \
        T gi = g[i];                                                                                   \
        if (gi < (T)0 || x[i] > (T)0) v += gi * gi;                                                    \
    }                                                                                                  \
    return (T)sqrt((double)v);                                                                         \
}                                                                                                      \
                                                                                                       \
/* src/alspgrad.jl:134-138 (and :290-294): Fn = max(F - alpha*G, 0), D = Fn - F */                     \
void CAT(oracle_pg_step, SFX)(const T* F, const T* G, T alpha, T* Fn, T* D, int64_t len) {             \
    for (int64_t i = 0; i < len; ++i) {                                                                \
        T fi = F[i];                                                                                   \
        T v = CAT(jl_max, SFX)((T)(fi - (T)(alpha * G[i])), (T)0);                                     \
        Fn[i] = v;                                                                                     \
        D[i] = v - fi;                                                                                 \
    }                                                                                                  \
}

DEFINE_ORACLE(float, f32, FLT_EPSILON)
DEFINE_ORACLE(double, f64, DBL_EPSILON)
