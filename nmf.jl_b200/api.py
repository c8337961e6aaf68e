"""
Host-side mirror of NMF.jl's operator API for the accelerated path, above the C ABI
(include/nmfb200.h).  Julia is not available in the build image, so this Python layer is what the
tests drive; nmf.jl_b200/julia/NMFB200.jl is the same thing written as `ccall`s.

Same names, argument meaning, defaults and error behaviour as the reference:
  Result            common.jl:21-38        nnmf               interf.jl:3-83
  MultUpdate        multupd.jl:9-43        solve_replicates   interf.jl:85-101
  GreedyCD          greedycd.jl:10-31      randinit           initialization.jl:4-17
  solve             multupd.jl:45 / greedycd.jl:33  (`NMF.solve!(alg, X, W, H)`)
  ProjectedALS      projals.jl:18-39       CoordinateDescent  coorddesc.jl:24-51
  ALSPGrad          alspgrad.jl:352-383
SPA exists as an option type (spa.jl:8-15) but is not on the accelerated path; solve() raises
NotImplementedError for it rather than falling back to a CPU implementation.

Arrays: NumPy, shapes as in Julia (X p x n, W p x k, H k x n).  Column-major (Fortran-order) arrays
are passed to the library without a copy and updated in place; other layouts are staged through a
column-major copy and written back, so `W` and `H` are always mutated like `solve!` does.
"""
from __future__ import annotations

import ctypes
import math
import warnings
from typing import Optional

import numpy as np

from . import _lib


class ArgumentError(ValueError):
    """Julia ArgumentError."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch."""


class NmfB200Error(RuntimeError):
    """CUDA / NCCL / state errors reported by libnmfb200."""


class NumericalError(NmfB200Error):
    """NMFB200_ENUMERIC: the k x k Gram of ProjectedALS is not positive definite.  The reference ignores the `info` of
    LAPACK.potrf! (utils.jl:63-84) and carries on with an unfinished factor; this library stops."""


def _raise(status: int, msg: str):
    if status == _lib.EINVAL:
        raise ArgumentError(msg)
    if status == _lib.EDIM:
        raise DimensionMismatch(msg)
    if status == _lib.ENOTSUP:
        raise NotImplementedError(msg)
    if status == _lib.ENUMERIC:
        raise NumericalError(msg)
    raise NmfB200Error(f"[{_lib.load().nmfb200_status_string(status).decode()}] {msg}")


def _eps(T) -> float:
    return float(np.finfo(np.dtype(T)).eps)


# --------------------------------------------------------------------------------------------------
# Result (common.jl:21-38)
# --------------------------------------------------------------------------------------------------
class Result:
    __slots__ = ("W", "H", "niters", "converged", "objvalue", "info")

    def __init__(self, W: np.ndarray, H: np.ndarray, niters: int, converged: bool, objv, info=None):
        if W.shape[1] != H.shape[0]:
            raise DimensionMismatch("Inner dimensions of W and H mismatch.")
        self.W, self.H = W, H
        self.niters, self.converged = int(niters), bool(converged)
        self.objvalue = W.dtype.type(objv)
        self.info = info or {}

    def __eq__(self, other):  # common.jl:37
        return (isinstance(other, Result) and np.array_equal(self.W, other.W) and np.array_equal(self.H, other.H)
                and self.niters == other.niters and self.converged == other.converged and self.objvalue == other.objvalue)

    def __hash__(self):  # common.jl:38
        return hash((self.W.tobytes(), self.H.tobytes(), self.niters, self.converged, float(self.objvalue)))

    def __repr__(self):
        return (f"Result{{{self.W.dtype.name}}}(W {self.W.shape}, H {self.H.shape}, niters={self.niters}, "
                f"converged={self.converged}, objvalue={self.objvalue})")


# --------------------------------------------------------------------------------------------------
# algorithm option types
# --------------------------------------------------------------------------------------------------
class MultUpdate:
    """multupd.jl:9-43"""

    def __init__(self, T=np.float64, *, obj="mse", maxiter=100, verbose=False, tol=None, update_H=True,
                 lambda_w=0.0, lambda_h=0.0, lambda_=None):
        T = np.dtype(T)
        tol = np.cbrt(_eps(T)) if tol is None else tol
        if obj not in ("mse", "div"):
            raise ArgumentError("Invalid value for obj.")
        if not maxiter > 1:
            raise ArgumentError("maxiter must be greater than 1.")
        if not tol > 0:
            raise ArgumentError("tol must be positive.")
        if not lambda_w >= 0:
            raise ArgumentError("lambda_w must be non-negative.")
        if not lambda_h >= 0:
            raise ArgumentError("lambda_h must be non-negative.")
        if lambda_ is not None and lambda_ >= 0:
            warnings.warn("lambda is deprecated, use lambda_w and lambda_h instead.")
            lambda_w = lambda_ if lambda_w == 0 else lambda_w
            lambda_h = lambda_ if lambda_h == 0 else lambda_h
        if obj == "div":
            lambda_w = max(lambda_w, math.sqrt(_eps(T)))
            lambda_h = max(lambda_h, math.sqrt(_eps(T)))
        self.T, self.obj, self.maxiter, self.verbose = T, obj, int(maxiter), bool(verbose)
        self.tol, self.update_H = T.type(tol), bool(update_H)
        self.lambda_w, self.lambda_h = T.type(lambda_w), T.type(lambda_h)


class GreedyCD:
    """greedycd.jl:10-31"""

    def __init__(self, T=np.float64, *, maxiter=100, verbose=False, tol=None, update_H=True, lambda_w=0.0, lambda_h=0.0):
        T = np.dtype(T)
        tol = np.cbrt(_eps(T)) if tol is None else tol
        if not maxiter > 1:
            raise ArgumentError("maxiter must be greater than 1.")
        if not tol > 0:
            raise ArgumentError("tol must be positive.")
        if not lambda_w >= 0:
            raise ArgumentError("lambda_w must be non-negative.")
        if not lambda_h >= 0:
            raise ArgumentError("lambda_h must be non-negative.")
        self.T, self.maxiter, self.verbose = T, int(maxiter), bool(verbose)
        self.tol, self.update_H = T.type(tol), bool(update_H)
        self.lambda_w, self.lambda_h = T.type(lambda_w), T.type(lambda_h)


class ProjectedALS:
    """projals.jl:18-35 (no validation in the reference constructor)."""

    def __init__(self, T=np.float64, *, maxiter=100, verbose=False, tol=None, update_H=True, lambda_w=None, lambda_h=None):
        T = np.dtype(T)
        c = np.cbrt(_eps(T))
        self.T, self.maxiter, self.verbose = T, int(maxiter), bool(verbose)
        self.tol = T.type(c if tol is None else tol)
        self.update_H = bool(update_H)
        self.lambda_w = T.type(c if lambda_w is None else lambda_w)
        self.lambda_h = T.type(c if lambda_h is None else lambda_h)


class ALSPGrad:
    """alspgrad.jl:352-373."""

    def __init__(self, T=np.float64, *, maxiter=100, maxsubiter=200, tol=None, tolg=None, update_H=True, verbose=False):
        T = np.dtype(T)
        self.T, self.maxiter, self.maxsubiter = T, int(maxiter), int(maxsubiter)
        self.tol = T.type(np.cbrt(_eps(T)) if tol is None else tol)
        self.tolg = T.type(_eps(T) ** 0.25 if tolg is None else tolg)
        self.update_H, self.verbose = bool(update_H), bool(verbose)


_CD_REG = {"both": 0, "components": 1, "transformation": 2, "none": 3}


class CoordinateDescent:
    """coorddesc.jl:24-46.  `seed` feeds the library's permutation generator when shuffle=True (the reference draws
    `randperm` from Julia's global RNG, coorddesc.jl:131-132)."""

    def __init__(self, T=np.float64, *, maxiter=100, verbose=False, tol=None, update_H=True, alpha=0.0, l1ratio=0.0,
                 regularization="both", shuffle=False, seed=0):
        T = np.dtype(T)
        self.T, self.maxiter, self.verbose = T, int(maxiter), bool(verbose)
        self.tol = T.type(np.cbrt(_eps(T)) if tol is None else tol)
        self.update_H = bool(update_H)
        self.alpha, self.l1ratio = T.type(alpha), T.type(l1ratio)
        if regularization not in _CD_REG:
            raise ArgumentError("regularization must be one of :both, :components, :transformation, :none")
        self.regularization, self.shuffle, self.seed = regularization, bool(shuffle), int(seed)


class SPA:
    """spa.jl:8-15.  Not accelerated (non-iterative)."""

    def __init__(self, T=np.float64, *, obj="mse"):
        if obj not in ("mse", "div"):
            raise ArgumentError("Invalid value for obj.")
        self.T, self.obj = np.dtype(T), obj


# --------------------------------------------------------------------------------------------------
# Session: one GPU handle with X resident (what `solve!` would keep alive between replicates)
# --------------------------------------------------------------------------------------------------
_SFX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}


def _is_sparse(X) -> bool:
    """A scipy.sparse matrix / array (duck-typed: scipy is imported only if the caller already holds one)."""
    return hasattr(X, "tocsc") and hasattr(X, "nnz") and not isinstance(X, np.ndarray)


def _col_major(a: np.ndarray, dtype) -> np.ndarray:
    return np.asfortranarray(a, dtype=dtype)


class Session:
    def __init__(self, device: int = 0, engine: str = "auto", stream: Optional[int] = None):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        st = self._lib.nmfb200_create(ctypes.byref(h), int(device), 0)
        if st != _lib.OK:
            _raise(st, f"nmfb200_create(device={device}) failed: {self._lib.nmfb200_status_string(st).decode()}")
        self._h = h
        self._trace_cb = None
        self.dtype = None
        self.shape = None
        self._keep = None
        self._x_host = None
        self.set_option("engine", engine)
        if stream is not None:
            self._check(self._lib.nmfb200_set_stream(self._h, ctypes.c_void_p(stream)))

    # -- plumbing
    def _check(self, st):
        if st != _lib.OK:
            _raise(st, self._lib.nmfb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nmfb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, key: str, value) -> None:
        self._check(self._lib.nmfb200_set_option(self._h, key.encode(), str(value).encode()))

    def set_trace(self, fn) -> None:
        """fn(iter, elapsed_s, objv, objv_change, dev) -- the verbose table of common.jl:57-58,80-81."""
        if fn is None:
            self._trace_cb = _lib.TRACE_FN(0)
        else:
            self._trace_cb = _lib.TRACE_FN(lambda user, it, el, ob, ch, dv: fn(it, el, ob, ch, dv))
        self._check(self._lib.nmfb200_set_trace(self._h, self._trace_cb, None))

    # -- X
    def set_X(self, X, check_nonneg: bool = False) -> None:
        """X: a NumPy matrix, or a scipy.sparse matrix (README.md:22 "Sparse NMF") -- its CSC arrays are uploaded and expanded on
        the device (nmfb200_set_X_csc_*)."""
        if _is_sparse(X):
            return self._set_X_csc(X, check_nonneg)
        X = np.asarray(X)
        if X.ndim != 2:
            raise DimensionMismatch("X must be a matrix")
        T = X.dtype
        if T not in _SFX:
            raise ArgumentError(f"eltype {T} not supported (Float32 / Float64)")
        Xf = _col_major(X, T)
        p, n = Xf.shape
        fn = getattr(self._lib, f"nmfb200_set_X_{_SFX[T]}")
        self._check(fn(self._h, Xf.ctypes.data_as(ctypes.c_void_p), p, n, p, int(check_nonneg)))
        self.dtype, self.shape = T, (p, n)
        self._x_host = X  # identity of the matrix that is resident (solve(alg, X, ..., session=s) compares against it)

    def _set_X_csc(self, X, check_nonneg: bool) -> None:
        Xc = X.tocsc()
        T = np.dtype(Xc.dtype)
        if T not in _SFX:
            raise ArgumentError(f"eltype {T} not supported (Float32 / Float64)")
        p, n = Xc.shape
        colptr = np.ascontiguousarray(Xc.indptr, dtype=np.int64)
        rowval = np.ascontiguousarray(Xc.indices, dtype=np.int64)
        nzval = np.ascontiguousarray(Xc.data, dtype=T)
        fn = getattr(self._lib, f"nmfb200_set_X_csc_{_SFX[T]}")
        self._check(fn(self._h, colptr.ctypes.data_as(ctypes.c_void_p), rowval.ctypes.data_as(ctypes.c_void_p),
                       nzval.ctypes.data_as(ctypes.c_void_p), p, n, 0, int(check_nonneg)))
        self.dtype, self.shape = T, (p, n)
        self._x_host = X

    def set_X_device(self, ptr: int, p: int, n: int, ldx: int, dtype, check_nonneg: bool = False, keepalive=None) -> None:
        """X already resident on this GPU (column-major p x n at device address `ptr`)."""
        T = np.dtype(dtype)
        fn = getattr(self._lib, f"nmfb200_set_X_dev_{_SFX[T]}")
        self._check(fn(self._h, ctypes.c_void_p(ptr), p, n, ldx, int(check_nonneg)))
        self.dtype, self.shape, self._keep = T, (p, n), keepalive
        self._x_host = None

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        st = _lib.load().nmfb200_comm_unique_id(buf)
        if st != _lib.OK:
            _raise(st, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, rank: int, nranks: int, unique_id: bytes) -> None:
        buf = ctypes.create_string_buffer(unique_id, _lib.UNIQUE_ID_BYTES)
        self._check(self._lib.nmfb200_comm_init(self._h, rank, nranks, buf))

    # -- initialisation on the device
    def randinit(self, k: int, *, seed: int = 0, normalize: bool = False, zeroh: bool = False, row_offset: int = 0,
                 p_total: Optional[int] = None):
        """NMF.randinit (initialization.jl:4-17) generated on the GPU for the X that is resident: counter-based Philox4x32-10
        keyed by `seed` (nmfb200_randinit_*; the draw of element (i, j) depends only on its index in the unsharded matrix, so
        row-sharded ranks that pass their row_offset / p_total get the rows of the same W).  Returns (W, H), column-major."""
        if self.dtype is None:
            raise NmfB200Error("set_X must precede randinit")
        p, n = self.shape
        W = np.empty((p, k), dtype=self.dtype, order="F")
        H = np.empty((k, n), dtype=self.dtype, order="F")
        fn = getattr(self._lib, "nmfb200_randinit_" + _SFX[self.dtype])
        self._check(fn(self._h, W.ctypes.data, p, H.ctypes.data, k, k, int(seed) & ((1 << 64) - 1), int(row_offset),
                       int(p if p_total is None else p_total), int(normalize), int(zeroh), 0))
        return W, H

    def rsvd(self, k: int, *, seed: int = 0):
        """RandomizedLinAlg.rsvd(X, k) as called at initialization.jl:78, entirely on the GPU for the resident X (nmfb200_rsvd_*:
        Philox Gaussian test matrix keyed by `seed`, CholeskyQR2 and a one-sided Jacobi SVD in Float64, csrc/init_device.cuh).
        Returns (U p x k, S k, V n x k).  Raises NumericalError when the sample is numerically rank deficient (k > rank(X)): use the
        host range finder (`rsvd(session, k, rng)`, LAPACK Householder QR) then."""
        if self.dtype is None:
            raise NmfB200Error("set_X must precede rsvd")
        p, n = self.shape
        U = np.empty((p, k), dtype=self.dtype, order="F")
        S = np.empty(k, dtype=self.dtype)
        V = np.empty((n, k), dtype=self.dtype, order="F")
        fn = getattr(self._lib, "nmfb200_rsvd_" + _SFX[self.dtype])
        self._check(fn(self._h, k, int(seed) & ((1 << 64) - 1), U.ctypes.data, p, S.ctypes.data, V.ctypes.data, n))
        return U, S, V

    def nndsvd(self, k: int, *, variant: str = "std", zeroh: bool = False, seed: int = 0):
        """NMF.nndsvd(X, k; zeroh, variant) (initialization.jl:70-101) on the GPU for the resident X: rsvd as above, then `_nndsvd!`
        (one CTA per component).  The :nndsvdar fill draws come from Philox stream 3 of the same seed.  Returns (W, H)."""
        if self.dtype is None:
            raise NmfB200Error("set_X must precede nndsvd")
        if variant not in ("std", "a", "ar"):
            raise ArgumentError("Invalid value for variant")
        p, n = self.shape
        W = np.empty((p, k), dtype=self.dtype, order="F")
        H = np.empty((k, n), dtype=self.dtype, order="F")
        fn = getattr(self._lib, "nmfb200_nndsvd_" + _SFX[self.dtype])
        self._check(fn(self._h, W.ctypes.data, p, H.ctypes.data, k, k, {"std": 0, "a": 1, "ar": 2}[variant], int(zeroh),
                       int(seed) & ((1 << 64) - 1), 0))
        return W, H

    # -- solve
    def mul_X(self, B: np.ndarray, transpose: bool = False) -> np.ndarray:
        """X * B (transpose=False) or X' * B on the resident X (nmfb200_mul_X_*): the two X-sized products of the
        randomised range finder behind nndsvd (initialization.jl:78)."""
        if self.dtype is None:
            raise NmfB200Error("set_X must precede mul_X")
        p, n = self.shape
        B = _col_major(np.asarray(B), self.dtype)
        rows_b, rows_c = (p, n) if transpose else (n, p)
        if B.ndim != 2 or B.shape[0] != rows_b:
            raise DimensionMismatch("Dimensions of X and B are inconsistent.")
        c = B.shape[1]
        C = np.empty((rows_c, c), dtype=self.dtype, order="F")
        fn = getattr(self._lib, "nmfb200_mul_X_" + _SFX[self.dtype])
        self._check(fn(self._h, 1 if transpose else 0, B.ctypes.data, rows_b, c, C.ctypes.data, rows_c))
        return C

    def solve_raw(self, alg_name: str, T, W_ptr: int, ldw: int, H_ptr: int, ldh: int, k: int, maxiter: int, tol,
                  lambda_w, lambda_h, update_H: bool, verbose: bool, on_device: bool) -> _lib.NmfResult:
        T = np.dtype(T)
        fn = getattr(self._lib, f"nmfb200_solve_{alg_name}_{_SFX[T]}")
        res = _lib.NmfResult()
        self._check(fn(self._h, ctypes.c_void_p(W_ptr), ldw, ctypes.c_void_p(H_ptr), ldh, k, int(maxiter), float(tol),
                       float(lambda_w), float(lambda_h), int(update_H), int(verbose), int(on_device), ctypes.byref(res)))
        return res

    def solve(self, alg, W: np.ndarray, H: np.ndarray) -> Result:
        """NMF.solve!(alg, X, W, H) with X = the matrix given to set_X."""
        if isinstance(alg, MultUpdate):
            name = "multmse" if alg.obj == "mse" else "multdiv"
        elif isinstance(alg, GreedyCD):
            name = "greedycd"
        elif isinstance(alg, ProjectedALS):
            name = "projals"
        elif isinstance(alg, CoordinateDescent):
            name = "cd"
        elif isinstance(alg, ALSPGrad):
            name = "alspgrad"
        elif isinstance(alg, SPA):
            raise NotImplementedError(
                "SPA is not on the accelerated path (SURVEY.md section 8f); this package has no CPU fallback")
        else:
            raise TypeError(f"unknown algorithm type {type(alg).__name__}")
        if self.shape is None:
            raise NmfB200Error("set_X must precede solve")
        T = alg.T
        if self.dtype != T or W.dtype != T or H.dtype != T:
            raise TypeError(f"element types differ: alg {T}, X {self.dtype}, W {W.dtype}, H {H.dtype}")
        p, n = self.shape
        k = W.shape[1]
        if not (W.shape[0] == p and H.shape == (k, n)):  # nmf_checksize, common.jl:5-16
            raise DimensionMismatch("Dimensions of X, W, and H are inconsistent.")
        Wf = W if W.flags.f_contiguous else np.asfortranarray(W)
        Hf = H if H.flags.f_contiguous else np.asfortranarray(H)
        if alg.verbose and self._trace_cb is None:
            self.set_trace(_print_trace)
        if name == "cd":
            res = _lib.NmfResult()
            fn = getattr(self._lib, f"nmfb200_solve_cd_{_SFX[T]}")
            self._check(fn(self._h, ctypes.c_void_p(Wf.ctypes.data), p, ctypes.c_void_p(Hf.ctypes.data), k, k, alg.maxiter,
                           float(alg.tol), float(alg.alpha), float(alg.l1ratio), _CD_REG[alg.regularization], int(alg.shuffle),
                           alg.seed & ((1 << 64) - 1), int(alg.update_H), int(alg.verbose), 0, ctypes.byref(res)))
            r = res
        elif name == "alspgrad":
            res = _lib.NmfResult()
            fn = getattr(self._lib, f"nmfb200_solve_alspgrad_{_SFX[T]}")
            self._check(fn(self._h, ctypes.c_void_p(Wf.ctypes.data), p, ctypes.c_void_p(Hf.ctypes.data), k, k, alg.maxiter,
                           alg.maxsubiter, float(alg.tol), float(alg.tolg), int(alg.update_H), int(alg.verbose), 0,
                           ctypes.byref(res)))
            r = res
        else:
            r = self.solve_raw(name, T, Wf.ctypes.data, p, Hf.ctypes.data, k, k, alg.maxiter, alg.tol, alg.lambda_w,
                               alg.lambda_h, alg.update_H, alg.verbose, False)
        if Wf is not W:
            W[...] = Wf
        if Hf is not H:
            H[...] = Hf
        info = {"engine": "tc" if r.engine == 1 else "simt", "solve_ms": r.solve_ms, "upload_ms": r.upload_ms,
                "last_dev": r.last_dev, "coordinate_updates": r.coordinate_updates, "kernel_launches": r.kernel_launches,
                "hot_kernel_ms": r.hot_kernel_ms, "hot_kernel_launches": r.hot_kernel_launches,
                "sub_iterations": r.sub_iterations, "tolg_final": r.tolg_final}
        return Result(W, H, r.niters, bool(r.converged), r.objvalue, info)


    def solve_batched(self, alg, Ws, Hs):
        """`len(Ws)` independent MultUpdate(:mse) solves of the resident X as ONE stacked iteration
        (nmfb200_solve_multmse_batched_f32: every pass over X serves all of them; block-diagonal Grams keep them independent;
        stop_condition per replicate).  Ws[r] (p x k) and Hs[r] (k x n) are updated in place like `solve!`; returns one Result per
        replicate.  Raises NotImplementedError when the library does not cover the request (Float32, tensor-core engine, one GPU,
        len(Ws) * k <= 256) -- callers then loop over solve()."""
        if not (isinstance(alg, MultUpdate) and alg.obj == "mse"):
            raise NotImplementedError("batched replicates cover MultUpdate(:mse) only")
        if self.shape is None:
            raise NmfB200Error("set_X must precede solve")
        T = alg.T
        if T != np.dtype(np.float32) or alg.verbose:
            raise NotImplementedError("batched replicates: Float32, verbose=false")
        p, n = self.shape
        R = len(Ws)
        k = Ws[0].shape[1]
        for W, H in zip(Ws, Hs):
            if W.dtype != T or H.dtype != T or self.dtype != T:
                raise TypeError("element types differ")
            if not (W.shape == (p, k) and H.shape == (k, n)):
                raise DimensionMismatch("Dimensions of X, W, and H are inconsistent.")
        Wst = np.asfortranarray(np.concatenate(Ws, axis=1))   # p x R*k: replicate r = columns [r*k, (r+1)*k)
        Hst = np.asfortranarray(np.concatenate(Hs, axis=0))   # R*k x n: replicate r = rows [r*k, (r+1)*k)
        res = (_lib.NmfResult * R)()
        self._check(self._lib.nmfb200_solve_multmse_batched_f32(
            self._h, ctypes.c_void_p(Wst.ctypes.data), p, ctypes.c_void_p(Hst.ctypes.data), R * k, k, R, alg.maxiter, float(alg.tol),
            float(alg.lambda_w), float(alg.lambda_h), int(alg.update_H), 0, res))
        out = []
        for r in range(R):
            Ws[r][...] = Wst[:, r * k:(r + 1) * k]
            Hs[r][...] = Hst[r * k:(r + 1) * k, :]
            info = {"engine": "tc", "solve_ms": res[r].solve_ms, "upload_ms": res[r].upload_ms, "last_dev": res[r].last_dev,
                    "kernel_launches": res[r].kernel_launches, "batched": R}
            out.append(Result(Ws[r], Hs[r], res[r].niters, bool(res[r].converged), res[r].objvalue, info))
        return out


def _print_trace(it, elapsed, objv, change, dev):
    if it == 0:  # common.jl:57-58
        print("%-5s    %-13s    %-13s    %-13s    %-13s" % ("Iter", "Elapsed time", "objv", "objv.change", "(W & H).relchange"))
        print("%5d    %13.6e    %13.6e" % (0, 0.0, objv))
    else:  # common.jl:80-81
        print("%5d    %13.6e    %13.6e    %13.6e    %13.6e" % (it, elapsed, objv, change, dev))


# --------------------------------------------------------------------------------------------------
# solve! / randinit / nnmf
# --------------------------------------------------------------------------------------------------
def solve(alg, X: np.ndarray, W: np.ndarray, H: np.ndarray, *, device: int = 0, engine: str = "auto",
          session: Optional[Session] = None) -> Result:
    """NMF.solve!(alg, X, W, H) -> Result.  W and H are updated in place.  With `session`, X is uploaded unless it IS the
    matrix already resident there (same object); pass X=None to use whatever the session holds (set_X_device)."""
    own = session is None
    s = session or Session(device=device, engine=engine)
    try:
        if X is not None and (own or s.shape is None or s._x_host is not X):
            s.set_X(X)
        return s.solve(alg, W, H)
    finally:
        if own:
            s.close()


def randinit(p: int, n: int, k: int, T, *, normalize: bool = False, zeroh: bool = False, rng=None):
    """initialization.jl:4-12 (+ normalize1_cols!, utils.jl:26-32).  Host side, runs once per solve."""
    T = np.dtype(T)
    rng = rng if rng is not None else np.random.default_rng()
    W = np.asfortranarray(rng.random((p, k)), dtype=T)
    if normalize:
        for j in range(k):
            W[:, j] *= T.type(1) / W[:, j].sum(dtype=T)
    H = np.zeros((k, n), dtype=T, order="F") if zeroh else np.asfortranarray(rng.random((k, n)), dtype=T)
    return W, H


def rsvd(session: Session, k: int, rng=None):
    """RandomizedLinAlg.rsvd(X, k) as called at initialization.jl:78 (`Q = qr(X * randn(n, k)).Q`, `svd(Q' * X)`,
    `U = Q * U_B`; no oversampling, no power iterations).  The two products with X run on the GPU on the resident X
    (Session.mul_X); the thin QR of the p x k sample and the SVD of the k x n projection are LAPACK calls on the host,
    as in the reference.  The Gaussian test matrix comes from `rng` (Julia's global RNG cannot be reproduced).
    Returns (U, s, V) with V n x k."""
    rng = rng if rng is not None else np.random.default_rng()
    p, n = session.shape
    T = session.dtype
    Omega = rng.standard_normal((n, k)).astype(T)
    Q, _ = np.linalg.qr(session.mul_X(Omega))                     # Y = X * Omega on the GPU
    Bt = session.mul_X(np.asfortranarray(Q), transpose=True)      # B' = X' * Q on the GPU (n x k)
    Ub, s, Vt = np.linalg.svd(Bt.T, full_matrices=False)
    return (Q @ Ub)[:, :k], s[:k], Vt[:k, :].T


def _posnegnorm(x):
    """initialization.jl:103-115 (zeros are counted with the negative part, contributing 0)."""
    T = x.dtype.type
    pos = x > 0
    return np.sqrt(T(np.sum(x[pos] * x[pos], dtype=x.dtype))), np.sqrt(T(np.sum(x[~pos] * x[~pos], dtype=x.dtype)))


def nndsvd(X: np.ndarray, k: int, *, zeroh: bool = False, variant: str = "std", initdata=None, rng=None,
           session: Optional[Session] = None):
    """NMF.nndsvd (initialization.jl:70-101) with `_nndsvd!` (:26-68).  `initdata` = (U, S, V) of an SVD of X (the
    reference takes an `SVD` object); without it the triplets come from `rsvd`, whose X-sized products run on the
    GPU (`session` with X resident; one is opened on device 0 if none is given).  variant: "std" | "a" | "ar"."""
    X = X if _is_sparse(X) else np.asarray(X)
    T = np.dtype(X.dtype)
    p, n = X.shape
    if variant not in ("std", "a", "ar"):
        raise ArgumentError("Invalid value for variant")
    rng = rng if rng is not None else np.random.default_rng()
    if initdata is None:
        if session is None:
            with Session() as s:
                s.set_X(X)
                U, sv, V = rsvd(s, k, rng)
        else:
            U, sv, V = rsvd(session, k, rng)
    else:
        U, sv, V = initdata[0][:, :k], initdata[1][:k], initdata[2][:, :k]
    U, sv, V = np.asarray(U, dtype=T), np.asarray(sv, dtype=T), np.asarray(V, dtype=T)   # :31-33
    if variant == "std":
        v0 = T.type(0)
    elif variant == "a":
        v0 = T.type(X.mean(dtype=np.float64))            # convert(T, mean(X)), :37
    else:
        v0 = T.type(X.mean(dtype=np.float64) * 0.01)     # :37
    W = np.empty((p, k), dtype=T, order="F")
    Ht = np.empty((n, k), dtype=T, order="F")
    for j in range(k):
        x, y = U[:, j], V[:, j]
        xp, xn = _posnegnorm(x)
        yp, yn = _posnegnorm(y)
        mp, mn = T.type(xp * yp), T.type(xn * yn)
        vj = v0
        if variant == "ar":
            vj = T.type(vj * T.type(rng.random()))       # :49-51
        if mp >= mn:                                     # :54-58 / :65-67
            ss = np.sqrt(T.type(sv[j] * mp))
            W[:, j] = np.where(x > 0, x * T.type(ss / xp), vj)
            if not zeroh:
                Ht[:, j] = np.where(y > 0, y * T.type(ss / yp), vj)
        else:                                            # :59-63 / :68-70
            ss = np.sqrt(T.type(sv[j] * mn))
            W[:, j] = np.where(x < 0, -(x * T.type(ss / xn)), vj)
            if not zeroh:
                Ht[:, j] = np.where(y < 0, -(y * T.type(ss / yn)), vj)
    H = np.zeros((k, n), dtype=T, order="F") if zeroh else np.asfortranarray(Ht.T)       # :88-98
    return W, H


def solve_replicates(alg, session: Session, W, H, *, replicates: int, initH: bool, rng=None, batched: bool = True) -> Result:
    """interf.jl:85-101; X stays resident on the GPU across replicates.  `rng`: a NumPy Generator (restarts drawn on the host) or an
    int seed (restarts drawn on the GPU, Session.randinit with seed + replicate number).

    MultUpdate(:mse) in Float32 runs its replicates in groups of up to 256 // k as ONE stacked iteration each (Session.solve_batched:
    one pass over X per half-step serves the whole group -- SURVEY 8f-3).  The restarts are drawn in the reference's order (solve!
    of MultUpdate consumes no random numbers, so drawing a group's factors up front is the same stream), and the first replicate with
    the smallest objvalue wins as at interf.jl:94-98.  batched=False, or a request the library does not cover, loops one by one."""
    p, n = session.shape
    k = W.shape[1]

    def restart(rep):
        if isinstance(rng, (int, np.integer)):
            return session.randinit(k, seed=int(rng) + rep - 1, normalize=True, zeroh=not initH)
        return randinit(p, n, k, alg.T, normalize=True, zeroh=not initH, rng=rng)

    group = min(replicates, 256 // max(k, 1), 32)
    use_batch = (batched and replicates > 1 and group >= 2 and isinstance(alg, MultUpdate) and alg.obj == "mse"
                 and alg.T == np.dtype(np.float32) and not alg.verbose)
    ret, minobjv = None, None
    rep = 1
    while rep <= replicates:
        g = min(group, replicates - rep + 1) if use_batch else 1
        facs = [(W, H) if r == 1 else restart(r) for r in range(rep, rep + g)]
        results = None
        if g >= 2:
            try:
                results = session.solve_batched(alg, [f[0] for f in facs], [f[1] for f in facs])
            except NotImplementedError:   # small problem / exact engine / sharded handle: one by one below, same factors
                use_batch = False
        if results is None:
            results = [session.solve(alg, fw, fh) for fw, fh in facs]
        for tmp in results:
            if ret is None or minobjv > tmp.objvalue:
                ret, minobjv = tmp, tmp.objvalue
        rep += g
    return ret


_NNDSVD_VARIANT = {"nndsvd": "std", "nndsvda": "a", "nndsvdar": "ar"}
_NOT_ACCEL_INIT = ("spa",)
_NOT_ACCEL_ALG = ("spa",)


def nnmf(X: np.ndarray, k: int, *, init: str = "nndsvdar", initdata=None, alg: str = "greedycd", maxiter: int = 100,
         tol=None, replicates: int = 1, W0=None, H0=None, update_H: bool = True, verbose: bool = False, rng=None,
         device: int = 0, engine: str = "auto") -> Result:
    """interf.jl:3-83.  Same keyword names, defaults, validation order and messages.  `rng`, `device`
    and `engine` are additions (Julia's global RNG has no NumPy counterpart).  `rng` = a NumPy Generator: random factors are drawn
    on the host; `rng` = an int: init=:random and the random restarts of `replicates` are drawn on the GPU (Philox keyed by it)."""
    sparse = _is_sparse(X)
    X = X if sparse else np.asarray(X)
    T = np.dtype(X.dtype)
    if T not in _SFX:
        raise ArgumentError(f"eltype {T} not supported (Float32 / Float64)")
    tol = np.cbrt(_eps(T) / 100) if tol is None else tol
    if not bool(((X.data if sparse else X) >= 0).all()):  # interf.jl:15 (sparse: the stored entries; implicit zeros pass)
        raise ArgumentError("The elements of X must be non-negative.")
    p, n = X.shape
    if not k <= min(p, n):  # :18
        raise ArgumentError("The value of k should not exceed min(size(X)).")
    if not replicates >= 1:  # :20
        raise ArgumentError("The value of replicates must be positive.")
    if not update_H and init != "custom":  # :22-24
        warnings.warn("Only W will be updated.")
    if init == "custom":  # :26-33
        if W0 is None or H0 is None:
            raise ArgumentError("To use :custom initialization, set W0 and H0.")
        if not bool((np.asarray(W0) >= 0).all()):
            raise ArgumentError("The elements of W0 must be non-negative.")
        if tuple(W0.shape) != (p, k):
            raise ArgumentError("Invalid size for W0.")
        if not bool((np.asarray(H0) >= 0).all()):
            raise ArgumentError("The elements of H0 must be non-negative.")
        if tuple(H0.shape) != (k, n):
            raise ArgumentError("Invalid size for H0.")
    elif W0 is not None or H0 is not None:  # :35
        warnings.warn("Ignore W0 and H0 except for :custom initialization.")
    initH = alg != "projals"  # :39
    device_seed = int(rng) if isinstance(rng, (int, np.integer)) else None
    if init == "random" and device_seed is not None:  # :42-43, drawn on the GPU once the session holds X (below)
        W = H = None
    elif init == "random":  # :42-43
        W, H = randinit(p, n, k, T, normalize=True, zeroh=not initH, rng=rng)
    elif init == "custom":  # :52-53
        W, H = W0, H0
        if W.dtype != T or H.dtype != T:
            raise TypeError("W0 and H0 must have the element type of X")  # `W::Matrix{T}` assert, :57-58
    elif init in _NNDSVD_VARIANT:  # :44-49 -- needs X on the GPU for the range finder: done below, once the session holds X
        W = H = None
    elif init in _NOT_ACCEL_INIT:
        raise NotImplementedError(f"init=:{init} is not on the accelerated path yet (SURVEY.md section 8f); "
                                  "pass init='random', 'nndsvd', 'nndsvda', 'nndsvdar' or 'custom'")
    else:
        raise ArgumentError("Invalid value for init.")  # :55
    if alg == "multmse":  # :64-66
        inst = MultUpdate(T, obj="mse", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "multdiv":
        inst = MultUpdate(T, obj="div", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "greedycd":
        inst = GreedyCD(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "projals":  # :60-61
        inst = ProjectedALS(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "alspgrad":  # :62-63
        inst = ALSPGrad(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "cd":  # :68-69
        inst = CoordinateDescent(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg in _NOT_ACCEL_ALG:
        if alg == "spa" and init != "spa":
            raise ArgumentError("Invalid value for init, use :spa instead.")  # :74-76
        raise NotImplementedError(f"alg=:{alg} is not on the accelerated path yet (SURVEY.md section 8f)")
    else:
        raise ArgumentError("Invalid algorithm.")  # :79
    with Session(device=device, engine=engine) as s:
        s.set_X(X)
        if W is None and init == "random":
            W, H = s.randinit(k, seed=device_seed, normalize=True, zeroh=not initH)
        elif W is None and device_seed is not None and initdata is None:
            # rng = <int>: the whole initialiser runs on the GPU (range finder, QR, SVD, split: Session.nndsvd).  A sample that is
            # numerically rank deficient (k > rank(X)) needs the Householder QR the reference uses: host path with the same seed.
            try:
                W, H = s.nndsvd(k, variant=_NNDSVD_VARIANT[init], zeroh=not initH, seed=device_seed)
            except (NumericalError, NotImplementedError):
                W, H = nndsvd(X, k, zeroh=not initH, variant=_NNDSVD_VARIANT[init], rng=np.random.default_rng(device_seed), session=s)
        elif W is None:
            nrng = np.random.default_rng(device_seed) if device_seed is not None else (rng if rng is not None else np.random.default_rng())
            W, H = nndsvd(X, k, zeroh=not initH, variant=_NNDSVD_VARIANT[init], initdata=initdata, rng=nrng, session=s)
        return solve_replicates(inst, s, W, H, replicates=replicates, initH=initH, rng=rng)
