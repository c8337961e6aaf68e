/*
 * nmfb200.h -- C ABI of libnmfb200.so, the B200 (sm_100a) accelerator for the per-iteration hot
 * path of JuliaStats/NMF.jl (reference @ 2eed3ec, v1.0.3).
 *
 * The boundary sits at `NMF.solve!(alg, X, W, H) -> NMF.Result{T}` (reference src/multupd.jl:45,
 * src/greedycd.jl:33), one level above the reference's internal updater protocol
 * (prepare_state / update_wh! / evaluate_objv, src/common.jl:43-89), because a per-`update_wh!`
 * boundary would force W and H across PCIe every iteration.  Each entry point below names the
 * reference interface it replaces.  A Julia `ccall` binding is in nmf.jl_b200/julia/NMFB200.jl,
 * the Python ctypes binding used by the tests is nmf.jl_b200/_lib.py; both bind exactly this file.
 *
 * Conventions
 *   - plain C, no CUDA / torch types in any signature; `void* stream` is a cudaStream_t.
 *   - every matrix is COLUMN-MAJOR with an explicit leading dimension (Julia `Matrix{T}` layout):
 *       X  p x n (ldx >= p), W  p x k (ldw >= p), H  k x n (ldh >= k).
 *     A C-order NumPy array of shape (n, p) is byte-identical to a Julia p x n matrix.
 *   - the caller owns all host buffers; W and H are updated IN PLACE (the reference mutates the
 *     caller's W, H: README.md:160-166) and X is read-only.
 *   - every function returns an nmfb200_status; nothing throws or exits across the ABI.
 *   - a handle is single-owner (one call at a time); distinct handles may be used from distinct
 *     host threads.  One handle drives one GPU; multi-GPU = one handle per process/GPU joined by
 *     nmfb200_comm_init (rows of X and W sharded across ranks, H replicated).
 */
#ifndef NMFB200_H
#define NMFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMFB200_VERSION 100 /* 0.1.0 */

typedef enum nmfb200_status {
    NMFB200_OK = 0,
    NMFB200_EINVAL = 1,  /* -> Julia ArgumentError       (multupd.jl:27-31, greedycd.jl:25-28, interf.jl:15-33) */
    NMFB200_EDIM = 2,    /* -> Julia DimensionMismatch   (common.jl:12-14) */
    NMFB200_ECUDA = 3,   /* CUDA runtime / driver error; text in nmfb200_last_error */
    NMFB200_ENCCL = 4,   /* NCCL error */
    NMFB200_ENOMEM = 5,  /* device allocation failed */
    NMFB200_ESTATE = 6,  /* call order violated (e.g. solve before set_X) */
    NMFB200_ENOTSUP = 7, /* valid request this build does not accelerate */
    NMFB200_ENUMERIC = 8 /* numerical breakdown: the k x k Gram of ProjectedALS is not positive definite.  The reference's
                            pdsolve!/pdrsolve! (utils.jl:63-84) ignore the `info` of LAPACK.potrf! and carry on with an
                            unfinished factor; this library stops instead (deliberate departure). */
} nmfb200_status;

/* Mirrors NMF.Result{T} (common.jl:21-34) minus the W/H aliases (the caller's own arrays). */
typedef struct nmfb200_result {
    int64_t niters;             /* Result.niters    (common.jl:24) */
    int32_t converged;          /* Result.converged (common.jl:25) */
    int32_t engine;             /* 0 = exact fp32/fp64 CUDA-core kernels, 1 = tcgen05 tensor-core kernels (ProjectedALS / CD / ALSPGrad:
                                   1 = their X-sized products ran on the tensor cores with split operands, the rest is exact) */
    double objvalue;            /* Result.objvalue  (common.jl:26), already rounded to T */
    double last_dev;            /* `dev` of the last stop_condition call (common.jl:73); full max, not partial */
    double solve_ms;            /* device time of the iteration loop (CUDA events), excl. transfers */
    double upload_ms;           /* host->device time of W/H (and X if passed through solve) */
    int64_t coordinate_updates; /* GreedyCD only: inner coordinate steps taken (greedycd.jl:144-162) */
    int64_t kernel_launches;    /* number of library kernels launched by this call */
    double hot_kernel_ms;       /* option "time_kernels"=1: summed device time of the dominant kernel's launches */
    int64_t hot_kernel_launches;/*   ... and how many launches that sum covers (0 when the option is off) */
    int64_t sub_iterations;     /* ALSPGrad only: projected-gradient sub-iterations summed over both factors (alspgrad.jl:114,270) */
    double tolg_final;          /* ALSPGrad only: the updater's tolg after its x0.1 decays (alspgrad.jl:409-411,419-421) */
} nmfb200_result;

typedef struct nmfb200_handle nmfb200_handle;

/* Called once per iteration when verbose != 0 -- replaces the table printed at common.jl:57-58,80-81.
 * iter = 0 is the pre-loop line (common.jl:56-58; objv_change and dev are NaN there). */
typedef void (*nmfb200_trace_fn)(void* user, int64_t iter, double elapsed_s, double objv,
                                 double objv_change, double dev);

/* ---- lifetime ---------------------------------------------------------------------------------- */
int nmfb200_version(void);
const char* nmfb200_status_string(int status);
/* device: CUDA ordinal.  flags: reserved, pass 0. */
int nmfb200_create(nmfb200_handle** out, int device, int flags);
int nmfb200_destroy(nmfb200_handle* h);
const char* nmfb200_last_error(const nmfb200_handle* h);
/* Run all work of this handle on the caller's stream (cudaStream_t); NULL = the handle's own. */
int nmfb200_set_stream(nmfb200_handle* h, void* stream);
/* Options (key, value):
 *   "engine"      = "auto" | "simt" | "tc"   -- simt: exact fp32/fp64 CUDA-core kernels;
 *                                               tc: tcgen05 bf16-operand / fp32-accumulate kernels
 *                                               (Float32 only); auto = tc for Float32 problems of at least
 *                                               2^20 cells (2^24 for GreedyCD), the exact engine below that.
 *   "check_every" = "<int>"                  -- host polls the device convergence flag every N
 *                                               iterations (results are independent of N).
 *   "tc_tile_rows" = "<int>"                 -- rows of a factor owned by one CTA of the tensor-core
 *                                               update kernel (multiple of 8 in [8,128]; 0 = auto).
 *   "precision"   = "bf16" | "bf16x3"        -- tensor-core engine, MultUpdate(:mse) and the GreedyCD gradient: bf16 = X and the
 *                                               streamed factor rounded to bf16 (default); bf16x3 = plus their bf16 remainders
 *                                               (hi*hi + hi*lo + lo*hi, three passes over X, split Gram): fp32-class products,
 *                                               W/H within ~2x the exact engine's distance to the reference, 2.65x the time.
 *   "emulate_shards" = "<G>"                 -- G in 2..8: run MultUpdate(:mse) as the ROW-SHARDED algorithm with G logical ranks on
 *                                               this one GPU (same kernels, arenas and flags as a G-GPU run; the caller passes the
 *                                               whole X, W, H); 0 (default) = off.  The 1-GPU test of the multi-GPU mathematics.
 *   "tc_xchg"      = "p2p" | "nccl"          -- multi-GPU MultUpdate(:mse): p2p (default) = tensor-core engine over NVLink peer
 *                                               memory (csrc/tc_shard.cuh); nccl = stay on the exact engine (ncclAllReduce).
 *   "tc_fused_hstep" = "-1" | "0" | "1"      -- row-sharded H-step as one launch (1), as numerators / slot sum / ratio (0), or
 *                                               auto (-1, default: one launch for two ranks).  Results are identical.
 *   "tc_side_stream" = "0" | "1"             -- row-sharded: H-Gram exchange on a side stream under the W-step (1, default).
 *   "tc_flush"     = "<int>"                 -- k-blocks per TMEM accumulation chunk of the update kernel (default 8; 0 = one chain:
 *                                               the tensor core's accumulator truncates, long chains bias the sums by ~3e-8 per MMA).
 *   "tc_xmul"      = "0" | "1"               -- ProjectedALS / CoordinateDescent / ALSPGrad (Float32, >= 2^20 cells): X-sized products
 *                                               on the tensor cores with split operands (1, default) or on the exact engine (0).
 *   "tc_pdl"       = "0" | "1"               -- 1 (default): the tensor-core update kernels are launched as programmatic
 *                                               dependents of the small reduce kernel in front of them (their X streaming
 *                                               overlaps it); 0: plain stream order.  Results are identical.
 *   "tc_chain"     = "0" | "1" | "<n>"        -- MultUpdate(:mse) on one GPU, k <= 128: 1 (default) = the kernels of the loop form one chain
 *                                               of programmatic dependents and an update launch is released by a counter of finished
 *                                               tiles of the launch before it instead of a kernel boundary (-6 % per iteration at
 *                                               config 2); the small reduce kernels then run as one CTA per SM (n > 1: n CTAs).
 *                                               0 = kernel-boundary hand-over.  Results are identical.
 *   "tc_skew"      = "0" | "1"               -- experiment on top of tc_chain, default 0: two groups of tiles half a period apart (one
 *                                               streams while the other is in its epilogues).  Identical results; measured: no gain,
 *                                               one SM streams ~52 GB/s whatever the others do (profiles/r2c_chain_handover.md).
 *   "tc_prefetch_next" = "<n>"               -- experiment, default 0: L2 prefetch of the first n k-blocks of the next launch's X
 *                                               panel from the epilogue (measured: no gain, profiles/r2b_prefetch_next.md).
 *   "tc_div_fused" = "0" | "1"               -- MultUpdate(:div) on the tensor-core engine: 1 (default) keeps the quotient
 *                                               tile X./(WH+delta) on chip between two tensor-core products; 0 writes it
 *                                               as a bf16 panel through HBM (older form, kept for comparison).
 *   "tc_debug"     = "<int>"                 -- diagnostics for profiling experiments (bit 3: phase clocks of the update
 *                                               kernel on stderr), 0 in production.
 *   "time_kernels" = "0" | "1"               -- bracket every launch of the dominant kernel with
 *                                               CUDA events and report the sum in nmfb200_result. */
int nmfb200_set_option(nmfb200_handle* h, const char* key, const char* value);
int nmfb200_set_trace(nmfb200_handle* h, nmfb200_trace_fn fn, void* user);

/* ---- data: the X argument of solve!(alg, X, W, H) ---------------------------------------------- */
/* Uploads (or adopts, for the _dev variants: X already resident on this handle's GPU) the data
 * matrix.  Stays resident across solves (replicates, interf.jl:85-101, reuse it).  Does the
 * non-negativity scan of interf.jl:15 only if check_nonneg != 0 (EINVAL on a negative entry). */
int nmfb200_set_X_f32(nmfb200_handle* h, const float* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg);
int nmfb200_set_X_f64(nmfb200_handle* h, const double* X, int64_t p, int64_t n, int64_t ldx, int check_nonneg);
int nmfb200_set_X_dev_f32(nmfb200_handle* h, const float* dX, int64_t p, int64_t n, int64_t ldx, int check_nonneg);
int nmfb200_set_X_dev_f64(nmfb200_handle* h, const double* dX, int64_t p, int64_t n, int64_t ldx, int check_nonneg);

/* X as a SparseMatrixCSC{T,Int64} (README.md:22 "Sparse NMF": the reference's solvers reach X only through mul!, so sparse X works
 * there).  colptr [n+1], rowval [nnz], nzval [nnz] are the three arrays of the Julia type (index_base = 1) or of a scipy csc_matrix
 * with int64 indices (index_base = 0).  Only the stored entries cross PCIe; they are expanded on the device into the dense
 * column-major matrix the kernels stream (p * n * sizeof(T) must fit in HBM -- the tensor-core path reads dense bf16 tiles whatever
 * the sparsity).  Duplicate entries are summed.  EINVAL: colptr not non-decreasing / not starting at index_base, a row index out of
 * range, or (check_nonneg != 0) a stored entry that is not >= 0. */
int nmfb200_set_X_csc_f32(nmfb200_handle* h, const int64_t* colptr, const int64_t* rowval, const float* nzval,
                          int64_t p, int64_t n, int index_base, int check_nonneg);
int nmfb200_set_X_csc_f64(nmfb200_handle* h, const int64_t* colptr, const int64_t* rowval, const double* nzval,
                          int64_t p, int64_t n, int index_base, int check_nonneg);

/* ---- solve!: one entry point per (algorithm, eltype) -------------------------------------------
 * Common arguments: W (p x k, ldw), H (k x n, ldh) initialised by the caller, updated in place.
 * `on_device` != 0: W and H are device pointers on this handle's GPU (no PCIe traffic).
 * maxiter/tol/lambda_w/lambda_h/update_H/verbose: the fields of the algorithm struct.  Validation
 * is the constructor's (EINVAL): maxiter > 1, tol > 0, lambda >= 0 for MultUpdate and GreedyCD; the
 * ProjectedALS / CoordinateDescent / ALSPGrad constructors of the reference validate nothing.  niters/converged/objvalue
 * follow nmf_skeleton! (common.jl:45-89) and stop_condition (common.jl:92-111). */

/* NMF.solve!(::MultUpdate{T} with obj=:mse, X, W, H)  -- multupd.jl:45-48, :83-116 */
int nmfb200_solve_multmse_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                              int64_t maxiter, float tol, float lambda_w, float lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_multmse_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                              int64_t maxiter, double tol, double lambda_w, double lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
/* solve_replicates! (interf.jl:85-101) for MultUpdate{Float32} with obj=:mse as ONE stacked iteration: `replicates` independent
 * solves of the resident X, replicate r initialised from columns [r*k, (r+1)*k) of W (p x replicates*k, ldw) and rows [r*k, (r+1)*k)
 * of H (replicates*k x n, ldh), all updated in place.  Every pass over X serves all replicates (the HBM-bound operand is read once
 * per half-step whatever `replicates` is); they do not interact (block-diagonal Grams), stop_condition is applied per replicate, and
 * out[r] (r < replicates) carries the niters / converged / objvalue that replicate's own solve! returns -- the caller picks the
 * smallest objvalue as interf.jl:94-98 does.  Tensor-core engine, one GPU, replicates * k <= 256, replicates <= 32; ENOTSUP otherwise
 * (the caller then loops over nmfb200_solve_multmse_f32).  solve_ms / upload_ms / kernel_launches of every out[r] describe the batch. */
int nmfb200_solve_multmse_batched_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                                      int32_t replicates, int64_t maxiter, float tol, float lambda_w, float lambda_h,
                                      int update_H, int on_device, nmfb200_result* out);
/* NMF.solve!(::MultUpdate{T} with obj=:div, X, W, H)  -- multupd.jl:45-51, :150-193.
 * The lambda floor max(lambda, sqrt(eps(T))) of the constructor (multupd.jl:37-40) is applied here. */
int nmfb200_solve_multdiv_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                              int64_t maxiter, float tol, float lambda_w, float lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_multdiv_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                              int64_t maxiter, double tol, double lambda_w, double lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
/* NMF.solve!(::GreedyCD{T}, X, W, H)  -- greedycd.jl:33-34, :94-178 */
int nmfb200_solve_greedycd_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                               int64_t maxiter, float tol, float lambda_w, float lambda_h, int update_H,
                               int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_greedycd_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                               int64_t maxiter, double tol, double lambda_w, double lambda_h, int update_H,
                               int verbose, int on_device, nmfb200_result* out);

/* NMF.solve!(::ProjectedALS{T}, X, W, H)  -- projals.jl:37-39, :77-107.  lambda_w / lambda_h are the L2 weights
 * (the constructor validates nothing: maxiter = 1 is accepted, projals.jl:26-34).  The k x k normal equations are
 * solved on the GPU (Gauss-Jordan inverse of the SPD Gram in Float64, one CTA); a Gram that is not positive definite
 * returns NMFB200_ENUMERIC (the reference does not check potrf!'s info, utils.jl:68,78, and would continue with garbage). */
int nmfb200_solve_projals_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                              int64_t maxiter, float tol, float lambda_w, float lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_projals_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                              int64_t maxiter, double tol, double lambda_w, double lambda_h, int update_H,
                              int verbose, int on_device, nmfb200_result* out);
/* NMF.solve!(::CoordinateDescent{T}, X, W, H)  -- coorddesc.jl:49-51, :108-181.
 * regularization: 0 :both, 1 :components, 2 :transformation, 3 :none (coorddesc.jl:65-71).
 * shuffle != 0 draws one permutation of the components per half-step (coorddesc.jl:131-132) from a splitmix64 /
 * Fisher-Yates generator seeded with `seed` (Julia's global RNG has no counterpart outside Julia). */
int nmfb200_solve_cd_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                         int64_t maxiter, float tol, float alpha, float l1ratio, int regularization, int shuffle,
                         uint64_t seed, int update_H, int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_cd_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                         int64_t maxiter, double tol, double alpha, double l1ratio, int regularization, int shuffle,
                         uint64_t seed, int update_H, int verbose, int on_device, nmfb200_result* out);
/* NMF.solve!(::ALSPGrad{T}, X, W, H)  -- alspgrad.jl:381-383, :400-425 with the sub-solvers :86-191 / :242-347
 * (traceiter = 20, beta = 0.2, sigma = 0.01 as hard-wired at alspgrad.jl:405-406,415-416). */
int nmfb200_solve_alspgrad_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k,
                               int64_t maxiter, int64_t maxsubiter, float tol, float tolg, int update_H,
                               int verbose, int on_device, nmfb200_result* out);
int nmfb200_solve_alspgrad_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k,
                               int64_t maxiter, int64_t maxsubiter, double tol, double tolg, int update_H,
                               int verbose, int on_device, nmfb200_result* out);

/* ---- initialisation support: the two X-sized products of the randomised range finder ------------
 * NMF.nndsvd (initialization.jl:70-101) obtains its singular triplets from RandomizedLinAlg.rsvd(X, k) =
 * `Q = qr(X * randn(n, k)).Q; svd(Q' * X)` (initialization.jl:78).  Everything in it is O((p+n) k^2) except the
 * two products with X; these run here on the X that is already resident for the solve:
 *   transpose_X == 0:  C (p x c, ldc) = X  * B (n x c, ldb)
 *   transpose_X != 0:  C (n x c, ldc) = X' * B (p x c, ldb)
 * B and C are column-major HOST arrays in the element type of X (EDIM / ESTATE as for solve). */
int nmfb200_mul_X_f32(nmfb200_handle* h, int transpose_X, const float* B, int64_t ldb, int64_t c, float* C, int64_t ldc);
int nmfb200_mul_X_f64(nmfb200_handle* h, int transpose_X, const double* B, int64_t ldb, int64_t c, double* C, int64_t ldc);

/* ---- initialisation on the device: NMF.randinit (initialization.jl:4-17) ------------------------
 * W ~ U[0,1)^(p x k), columns scaled to sum 1 when normalize != 0 (normalize1_cols!, utils.jl:26-32 -- what nnmf does for
 * init=:random, interf.jl:43); H ~ U[0,1)^(k x n), or zeros when zeroh != 0.  p, n are those of the X set on the handle.
 * The reference draws from Julia's global RNG, which nothing outside Julia can reproduce; this entry point is counter-based
 * instead: element e (column-major linear index; for W the index in the UNSHARDED matrix: row_offset + i + j * p_total) is
 * output word 0 (Float32: >> 8, * 2^-24) or words 0,1 (Float64: 53 bits, * 2^-53) of Philox4x32-10 with counter (e_lo, e_hi,
 * stream, 0) -- stream 0 for W, 1 for H -- and key (seed_lo, seed_hi).  So the draw does not depend on the launch geometry or
 * on how the rows are sharded (multi-GPU: pass the rank's row_offset and the total row count; column sums are all-reduced),
 * and a host can regenerate it (tests/test_gpu_init.py does).  W, H: column-major, host or (on_device != 0) device pointers. */
int nmfb200_randinit_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k, uint64_t seed,
                         int64_t row_offset, int64_t p_total, int normalize, int zeroh, int on_device);
int nmfb200_randinit_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k, uint64_t seed,
                         int64_t row_offset, int64_t p_total, int normalize, int zeroh, int on_device);

/* ---- initialisation on the device: NMF.nndsvd (initialization.jl:70-101) ------------------------
 * nmfb200_rsvd_*: RandomizedLinAlg.rsvd(X, k) as called at initialization.jl:78 on the resident X, entirely on the GPU:
 * Omega = randn(n, k) (Philox4x32-10 as above, stream 2, element e = i + j*n, Box-Muller in Float64 on words 0 and 1:
 * sqrt(-2 ln((w0 + 0.5) 2^-32)) cos(2 pi w1 2^-32)), Y = X*Omega, Q = qr(Y).Q by CholeskyQR2 in Float64, B' = X'*Q, svd(B) by
 * one-sided Jacobi in Float64, U = Q*U_B.  Outputs: U (p x k, ldu), S (k, decreasing), V (n x k, ldv), column-major HOST arrays.
 * ENUMERIC: the sample is numerically rank deficient (k > rank(X)): orthogonalise on the host instead (LAPACK Householder QR, which
 * is what the reference runs) -- the Python / Julia layers do that.  ENOTSUP on a row-sharded handle.
 * nmfb200_nndsvd_*: rsvd as above, then _nndsvd! (initialization.jl:26-68) -> W (p x k, ldw), H (k x n, ldh); host pointers, or
 * device pointers when on_device != 0.  variant: 0 :nndsvd, 1 :nndsvda (fill = mean(X)), 2 :nndsvdar (fill = mean(X)/100 * rand,
 * rand = Philox stream 3, element j); zeroh != 0: H = 0 (interf.jl:39,45-49 for :projals). */
int nmfb200_rsvd_f32(nmfb200_handle* h, int64_t k, uint64_t seed, float* U, int64_t ldu, float* S, float* V, int64_t ldv);
int nmfb200_rsvd_f64(nmfb200_handle* h, int64_t k, uint64_t seed, double* U, int64_t ldu, double* S, double* V, int64_t ldv);
int nmfb200_nndsvd_f32(nmfb200_handle* h, float* W, int64_t ldw, float* H, int64_t ldh, int64_t k, int variant, int zeroh,
                       uint64_t seed, int on_device);
int nmfb200_nndsvd_f64(nmfb200_handle* h, double* W, int64_t ldw, double* H, int64_t ldh, int64_t k, int variant, int zeroh,
                       uint64_t seed, int on_device);

/* ---- multi-GPU: rows of X / W sharded over ranks, H replicated -----------
 * No counterpart in the reference (single process).  One handle per rank/GPU.  The unique id is an
 * opaque 128-byte blob (an ncclUniqueId) created on rank 0 and distributed by the host program
 * (torch.distributed / MPI / sockets).  After comm_init every solve on the handle treats its X, W
 * as the rank's row shard; H, niters, converged and objvalue come back identical on all ranks.
 * MultUpdate(:mse) Float32 exchanges the k x n numerators, the k x k Grams and the convergence partial
 * sums once per iteration through NVLink peer memory (CUDA IPC, no NCCL call in the loop).  MultUpdate(:div)
 * and GreedyCD in Float32 stay on the tensor-core engine as well and all-reduce over NCCL on the solver's
 * stream: :div the column sums of W and the numerators W'Q (multupd.jl:175-176), GreedyCD the gradient of H,
 * W'W and -- a maximum -- p_init (greedycd.jl:132-137), both the W-side sums of stop_condition.  Everything
 * else (Float64, small problems, the other algorithms) all-reduces the same quantities on the exact engine. */
#define NMFB200_UNIQUE_ID_BYTES 128
int nmfb200_comm_unique_id(void* out_id_128);
/* Host-only helper (no GPU needed): which rows of H' (columns of H) rank `rank` of `ranks` OWNS in the row-sharded tensor-core
 * MultUpdate(:mse) solve -- it alone applies the multiplicative ratio to them; everybody else receives them (csrc/tc_shard.cuh).
 * H' is cut into tiles of `tile_rows` rows (128 for n >= 128, else n rounded up to 8), ceil(tiles / ranks) consecutive tiles per rank;
 * trailing ranks may own nothing (own_row0 == own_row1).  EINVAL for ranks outside 1..8 or rank outside 0..ranks-1. */
int nmfb200_shard_geometry(int64_t n, int ranks, int rank, int64_t* own_row0, int64_t* own_row1, int64_t* tile_rows);
int nmfb200_comm_init(nmfb200_handle* h, int rank, int nranks, const void* id_128);
int nmfb200_comm_destroy(nmfb200_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* NMFB200_H */
