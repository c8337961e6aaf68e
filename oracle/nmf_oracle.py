"""
nmf_oracle.py -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package `nmf.jl_b200`.

CPU restatement of the per-iteration hot path of JuliaStats/NMF.jl @ 2eed3ec (v1.0.3), *as written*:
the GEMMs the reference hands to LinearAlgebra.mul! (OpenBLAS) are NumPy matmuls (OpenBLAS 0.3.30 in
this image), the hand-written Julia loops are the sequential C loops of oracle_kernels.c.  dtype T is
float32 or float64 end to end, exactly like `Matrix{T}` in the reference.

PARITY STATUS ("parity unpinned" at bit level): there is no Julia in the build image or on the GPU
box, and the reference holds no golden vectors -- its tests for this path are convergence/property
tests on the analytic fixture laurberg6x3 (test/testproblems.jl:6-13).  The oracle is pinned against
every one of those (tests/test_oracle.py) and otherwise *defines* parity for this project.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.

Reference map (file:line under /root/reference/src):
  nmf_checksize      common.jl:5-16          Result            common.jl:21-34
  nmf_skeleton       common.jl:45-89         stop_condition    common.jl:92-111
  MultUpdate (ctor)  multupd.jl:9-43         solve_multupdate  multupd.jl:45-52
  MultUpdMSE         multupd.jl:56-116       MultUpdDiv        multupd.jl:121-193
  GreedyCD (ctor)    greedycd.jl:10-31       GreedyCDUpd       greedycd.jl:36-178
  ProjectedALS       projals.jl:18-107       pdsolve/pdrsolve  utils.jl:63-84
  CoordinateDescent  coorddesc.jl:24-181     ALSPGrad          alspgrad.jl:9-425
  nndsvd             initialization.jl:26-137
  randinit           initialization.jl:4-17  normalize1_cols   utils.jl:26-32
  nnmf               interf.jl:3-83          solve_replicates  interf.jl:85-101
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "oracle_kernels.c")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle_kernels.c with gcc (un-fused arithmetic, like Julia's scalar loops)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


def _clib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64, vp, dbl = ctypes.c_int64, ctypes.c_void_p, ctypes.c_double
        for sfx, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            getattr(_lib, f"oracle_mu_mse_ratio_{sfx}").argtypes = [vp, vp, vp, i64, ct, ct]
            getattr(_lib, f"oracle_mu_mse_ratio_{sfx}").restype = None
            getattr(_lib, f"oracle_mu_div_quot_{sfx}").argtypes = [vp, vp, vp, i64, ct]
            getattr(_lib, f"oracle_mu_div_quot_{sfx}").restype = None
            getattr(_lib, f"oracle_mu_div_scale_h_{sfx}").argtypes = [vp, vp, vp, i64, i64, ct]
            getattr(_lib, f"oracle_mu_div_scale_h_{sfx}").restype = None
            getattr(_lib, f"oracle_mu_div_scale_w_{sfx}").argtypes = [vp, vp, vp, i64, i64, ct]
            getattr(_lib, f"oracle_mu_div_scale_w_{sfx}").restype = None
            getattr(_lib, f"oracle_stop_condition_{sfx}").argtypes = [vp, vp, vp, vp, i64, i64, i64, ct, vp]
            getattr(_lib, f"oracle_stop_condition_{sfx}").restype = ctypes.c_int
            getattr(_lib, f"oracle_sql2dist_{sfx}").argtypes = [vp, vp, i64]
            getattr(_lib, f"oracle_sql2dist_{sfx}").restype = dbl
            getattr(_lib, f"oracle_gkldiv_{sfx}").argtypes = [vp, vp, i64]
            getattr(_lib, f"oracle_gkldiv_{sfx}").restype = dbl
            getattr(_lib, f"oracle_greedycd_rows_{sfx}").argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64]
            getattr(_lib, f"oracle_greedycd_rows_{sfx}").restype = i64
            getattr(_lib, f"oracle_projectnn_{sfx}").argtypes = [vp, i64]
            getattr(_lib, f"oracle_projectnn_{sfx}").restype = None
            getattr(_lib, f"oracle_cd_sweep_{sfx}").argtypes = [vp, vp, vp, i64, i64, vp]
            getattr(_lib, f"oracle_cd_sweep_{sfx}").restype = ct
            getattr(_lib, f"oracle_projgradnorm_{sfx}").argtypes = [vp, vp, i64]
            getattr(_lib, f"oracle_projgradnorm_{sfx}").restype = ct
            getattr(_lib, f"oracle_pg_step_{sfx}").argtypes = [vp, vp, ct, vp, vp, i64]
            getattr(_lib, f"oracle_pg_step_{sfx}").restype = None
    return _lib


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _fn(name, dtype):
    return getattr(_clib(), f"{name}_{_sfx(dtype)}")


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _F(a, dtype=None) -> np.ndarray:
    """Column-major (Julia layout) array of the given dtype."""
    return np.asfortranarray(a, dtype=dtype)


def _isF(a: np.ndarray) -> bool:
    return a.flags.f_contiguous


def _mm(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """A @ B returned column-major without a layout copy: (B' A')' -- the same sgemm/dgemm call
    LinearAlgebra.mul! makes on column-major operands."""
    return (B.T @ A.T).T


# --------------------------------------------------------------------------------------------------
# common.jl
# --------------------------------------------------------------------------------------------------
class DimensionMismatch(ValueError):
    """Stands in for Julia's DimensionMismatch."""


class ArgumentError(ValueError):
    """Stands in for Julia's ArgumentError."""


def nmf_checksize(X, W, H):
    """common.jl:5-16"""
    p, n = X.shape
    k = W.shape[1]
    if not (W.shape[0] == p and H.shape == (k, n)):
        raise DimensionMismatch("Dimensions of X, W, and H are inconsistent.")
    return p, n, k


@dataclass
class Result:
    """common.jl:21-34 (objvalue is converted to T by the inner constructor, common.jl:32)"""

    W: np.ndarray
    H: np.ndarray
    niters: int
    converged: bool
    objvalue: float

    def __post_init__(self):
        if self.W.shape[1] != self.H.shape[0]:
            raise DimensionMismatch("Inner dimensions of W and H mismatch.")
        self.objvalue = self.W.dtype.type(self.objvalue)


def stop_condition(W, preW, H, preH, eps):
    """common.jl:92-111 -> (converged, devmax)"""
    T = W.dtype
    p, k = W.shape
    n = H.shape[1]
    assert _isF(W) and _isF(preW) and _isF(H) and _isF(preH)
    devmax = np.zeros(1, dtype=T)
    conv = _fn("oracle_stop_condition", T)(_p(W), _p(preW), _p(H), _p(preH), p, n, k, T.type(eps), _p(devmax))
    return bool(conv), devmax[0]


def nmf_skeleton(updater, X, W, H, maxiter: int, verbose: bool, tol, log=None):
    """common.jl:45-89.  W and H are mutated in place (must be column-major arrays of dtype T)."""
    T = W.dtype
    assert _isF(W) and _isF(H) and H.dtype == T
    objv = T.type(np.nan)
    state = updater.prepare_state(X, W, H)
    preW = np.empty_like(W, order="F")
    preH = np.empty_like(H, order="F")
    if verbose:
        objv = updater.evaluate_objv(state, X, W, H)
        if log is not None:
            log.append((0, float(objv), float("nan")))
    converged = False
    t = 0
    while not converged and t < maxiter:
        t += 1
        np.copyto(preW, W)
        np.copyto(preH, H)
        updater.update_wh(state, X, W, H)
        converged, dev = stop_condition(W, preW, H, preH, tol)
        if verbose:
            objv = updater.evaluate_objv(state, X, W, H)
            if log is not None:
                log.append((t, float(objv), float(dev)))
    if not verbose:
        objv = updater.evaluate_objv(state, X, W, H)
    return Result(W, H, t, converged, objv)


def sqL2dist(a, b) -> float:
    """StatsBase.sqL2dist (not vendored; restated, Float64 accumulator)."""
    a = np.ascontiguousarray(a.ravel(order="K"))
    b = np.ascontiguousarray(b.ravel(order="K"))
    return _fn("oracle_sql2dist", a.dtype)(_p(a), _p(b), a.size)


def gkldiv(a, b) -> float:
    """StatsBase.gkldiv (not vendored; restated, Float64 accumulator)."""
    a = np.ascontiguousarray(a.ravel(order="K"))
    b = np.ascontiguousarray(b.ravel(order="K"))
    return _fn("oracle_gkldiv", a.dtype)(_p(a), _p(b), a.size)


# --------------------------------------------------------------------------------------------------
# multupd.jl
# --------------------------------------------------------------------------------------------------
class MultUpdate:
    """multupd.jl:9-43 (constructor rules incl. maxiter > 1 and the :div lambda floor)."""

    def __init__(self, T=np.float64, obj="mse", maxiter=100, verbose=False, tol=None, update_H=True,
                 lambda_w=0.0, lambda_h=0.0, lambda_=None):
        T = np.dtype(T)
        eps = np.finfo(T).eps
        if tol is None:
            tol = np.cbrt(eps)
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
            lambda_w = max(lambda_w, np.sqrt(eps))
            lambda_h = max(lambda_h, np.sqrt(eps))
        self.T = T
        self.obj = obj
        self.maxiter = int(maxiter)
        self.verbose = bool(verbose)
        self.tol = T.type(tol)
        self.update_H = bool(update_H)
        self.lambda_w = T.type(lambda_w)
        self.lambda_h = T.type(lambda_h)


class MultUpdMSE:
    """multupd.jl:56-116"""

    def __init__(self, T, update_H, lambda_w, lambda_h, delta):
        self.T, self.update_H, self.lambda_w, self.lambda_h, self.delta = np.dtype(T), update_H, lambda_w, lambda_h, delta

    def prepare_state(self, X, W, H):  # :70-80
        p, n, k = nmf_checksize(X, W, H)
        return {"WH": _mm(W, H)}

    def evaluate_objv(self, s, X, W, H):  # :81
        return self.T.type(0.5) * self.T.type(sqL2dist(X, s["WH"]))

    def update_wh(self, s, X, W, H):  # :83-116
        T = self.T
        ratio = _fn("oracle_mu_mse_ratio", T)
        WH = s["WH"]
        if self.update_H:
            WtX = _mm(W.T, X)                       # :98
            WtWH = _mm(W.T, WH)                     # :99
            ratio(_p(H), _p(WtX), _p(WtWH), H.size, T.type(self.lambda_h), T.type(self.delta))  # :101-103
            WH = s["WH"] = _mm(W, H)                # :104
        XHt = _mm(X, H.T)                           # :109
        WHHt = _mm(WH, H.T)                         # :110
        ratio(_p(W), _p(XHt), _p(WHHt), W.size, T.type(self.lambda_w), T.type(self.delta))      # :112-114
        s["WH"] = _mm(W, H)                         # :115


class MultUpdDiv:
    """multupd.jl:121-193"""

    def __init__(self, T, update_H, lambda_w, lambda_h, delta):
        self.T, self.update_H, self.lambda_w, self.lambda_h, self.delta = np.dtype(T), update_H, lambda_w, lambda_h, delta

    def prepare_state(self, X, W, H):  # :136-147
        nmf_checksize(X, W, H)
        return {"WH": _mm(W, H), "Q": np.empty(X.shape, dtype=self.T, order="F")}

    def evaluate_objv(self, s, X, W, H):  # :148 (gkldiv returns Float64; Result converts to T)
        return gkldiv(X, s["WH"])

    def update_wh(self, s, X, W, H):  # :150-193
        T = self.T
        p, n = X.shape
        k = W.shape[1]
        Q = s["Q"]
        quot = _fn("oracle_mu_div_quot", T)
        if self.update_H:
            quot(_p(Q), _p(X), _p(s["WH"]), X.size, T.type(self.delta))            # :172-174
            WtQ = _mm(W.T, Q)                                                      # :175
            sW = np.ascontiguousarray(W.sum(axis=0, dtype=T))                      # :176
            _fn("oracle_mu_div_scale_h", T)(_p(H), _p(WtQ), _p(sW), k, n, T.type(self.lambda_h))  # :177-179
            s["WH"] = _mm(W, H)                                                    # :180
        quot(_p(Q), _p(X), _p(s["WH"]), X.size, T.type(self.delta))                # :184-186
        QHt = _mm(Q, H.T)                                                          # :187
        sH = np.ascontiguousarray(H.sum(axis=1, dtype=T))                          # :188
        _fn("oracle_mu_div_scale_w", T)(_p(W), _p(QHt), _p(sH), p, k, T.type(self.lambda_w))      # :189-191
        s["WH"] = _mm(W, H)                                                        # :192


def solve_multupdate(alg: MultUpdate, X, W, H, log=None) -> Result:
    """multupd.jl:45-52"""
    T = alg.T
    delta = T.type(np.sqrt(np.finfo(T).eps))
    cls = MultUpdMSE if alg.obj == "mse" else MultUpdDiv
    upd = cls(T, alg.update_H, alg.lambda_w, alg.lambda_h, delta)
    return nmf_skeleton(upd, X, W, H, alg.maxiter, alg.verbose, alg.tol, log=log)


# --------------------------------------------------------------------------------------------------
# greedycd.jl
# --------------------------------------------------------------------------------------------------
class GreedyCD:
    """greedycd.jl:10-31"""

    def __init__(self, T=np.float64, maxiter=100, verbose=False, tol=None, update_H=True, lambda_w=0.0, lambda_h=0.0):
        T = np.dtype(T)
        if tol is None:
            tol = np.cbrt(np.finfo(T).eps)
        if not maxiter > 1:
            raise ArgumentError("maxiter must be greater than 1.")
        if not tol > 0:
            raise ArgumentError("tol must be positive.")
        if not lambda_w >= 0:
            raise ArgumentError("lambda_w must be non-negative.")
        if not lambda_h >= 0:
            raise ArgumentError("lambda_h must be non-negative.")
        self.T = T
        self.maxiter = int(maxiter)
        self.verbose = bool(verbose)
        self.tol = T.type(tol)
        self.update_H = bool(update_H)
        self.lambda_w = T.type(lambda_w)
        self.lambda_h = T.type(lambda_h)


def greedycd_rows(F, G, P):
    """greedycd.jl:125-165 on column-major F (rows x k, updated in place), G (rows x k, clobbered),
    P (k x k).  Returns the number of coordinate updates performed."""
    T = F.dtype
    rows, k = F.shape
    assert _isF(F) and _isF(G) and _isF(P)
    S = np.empty((rows, k), dtype=T, order="F")
    D = np.empty((rows, k), dtype=T, order="F")
    Fnew = np.empty((rows, k), dtype=T, order="F")
    q = np.empty(rows, dtype=np.int64)
    return _fn("oracle_greedycd_rows", T)(_p(F), _p(G), _p(P), _p(S), _p(D), _p(Fnew), _p(q), rows, k)


class GreedyCDUpd:
    """greedycd.jl:36-178"""

    def __init__(self, T, update_H, lambda_w, lambda_h):
        self.T, self.update_H, self.lambda_w, self.lambda_h = np.dtype(T), update_H, lambda_w, lambda_h
        self.coordinate_updates = 0

    def prepare_state(self, X, W, H):  # :60-80
        nmf_checksize(X, W, H)
        return {}

    def evaluate_objv(self, s, X, W, H):  # :82-92
        T = self.T
        WH = _mm(W, H)
        r = T.type(0.5) * T.type(sqL2dist(X, WH))
        if self.lambda_w > 0:
            r = T.type(r + self.lambda_w * T.type(np.abs(W).sum(dtype=T)))
        if self.lambda_h > 0:
            r = T.type(r + self.lambda_h * T.type(np.abs(H).sum(dtype=T)))
        return r

    def _update(self, X, F, Ot, lam):
        """_update_GreedyCD! :94-166 for factor F (rows x k) against Ot (cols x k): X is rows x cols."""
        T = self.T
        P = _mm(Ot.T, Ot)                 # :117
        Z = _mm(X, Ot)                    # :118
        G = _mm(F, P)                     # :119
        G -= Z                            # :120
        if lam > 0:
            G += T.type(lam)              # :121-123
        self.coordinate_updates += greedycd_rows(F, G, P)

    def update_wh(self, s, X, W, H):  # :168-178
        Ht = _F(H.T)                                  # lazy transpose in the reference; values identical
        self._update(X, W, Ht, self.lambda_w)         # :171
        if self.update_H:
            self._update(X.T, Ht, W, self.lambda_h)   # :175-176 (writes through transpose(H))
            H[...] = Ht.T


def solve_greedycd(alg: GreedyCD, X, W, H, log=None) -> Result:
    """greedycd.jl:33-34"""
    upd = GreedyCDUpd(alg.T, alg.update_H, alg.lambda_w, alg.lambda_h)
    res = nmf_skeleton(upd, X, W, H, alg.maxiter, alg.verbose, alg.tol, log=log)
    res.coordinate_updates = upd.coordinate_updates
    return res


# --------------------------------------------------------------------------------------------------
# projals.jl
# --------------------------------------------------------------------------------------------------
def projectnn(A):
    """utils.jl:34-41"""
    assert A.flags.f_contiguous or A.flags.c_contiguous
    _fn("oracle_projectnn", A.dtype)(_p(A), A.size)


def adddiag(A, a):
    """utils.jl:15-24"""
    if a != 0.0:
        A[np.diag_indices(A.shape[0])] += A.dtype.type(a)
    return A


def _lapack(name, T):
    from scipy.linalg import lapack

    return getattr(lapack, ("s" if np.dtype(T) == np.float32 else "d") + name)


def pdsolve(A, x):
    """utils.jl:63-70: potrf!('U', A); potrs!('U', A, x)  (LAPACK in T)."""
    T = A.dtype
    c, info = _lapack("potrf", T)(A, lower=0, overwrite_a=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"potrf info={info}")   # Julia: PosDefException
    sol, info = _lapack("potrs", T)(c, x, lower=0)
    return _F(sol, dtype=T)


def pdrsolve(A, B):
    """utils.jl:72-84: x <- A * inv(B) with inv(B) by potrf!/potri!/copytri! (LAPACK in T)."""
    T = B.dtype
    c, info = _lapack("potrf", T)(B, lower=0, overwrite_a=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"potrf info={info}")
    inv, info = _lapack("potri", T)(c, lower=0)
    inv = np.triu(inv) + np.triu(inv, 1).T          # copytri!(B, 'U')
    return _mm(A, _F(inv, dtype=T))


class ProjectedALS:
    """projals.jl:18-35 (the reference constructor validates nothing)."""

    def __init__(self, T=np.float64, maxiter=100, verbose=False, tol=None, update_H=True, lambda_w=None, lambda_h=None):
        T = np.dtype(T)
        c = np.cbrt(np.finfo(T).eps)
        self.T = T
        self.maxiter = int(maxiter)
        self.verbose = bool(verbose)
        self.tol = T.type(c if tol is None else tol)
        self.update_H = bool(update_H)
        self.lambda_w = T.type(c if lambda_w is None else lambda_w)
        self.lambda_h = T.type(c if lambda_h is None else lambda_h)


class ProjectedALSUpd:
    """projals.jl:42-107"""

    def __init__(self, T, update_H, lambda_w, lambda_h):
        self.T, self.update_H, self.lambda_w, self.lambda_h = np.dtype(T), update_H, lambda_w, lambda_h

    def prepare_state(self, X, W, H):  # :53-64
        nmf_checksize(X, W, H)
        return {"WH": _mm(W, H)}

    def evaluate_objv(self, s, X, W, H):  # :66-75
        T = self.T
        r = T.type(0.5) * T.type(sqL2dist(X, s["WH"]))
        if self.lambda_w > 0:
            r = T.type(r + T.type(T.type(0.5) * self.lambda_w) * T.type(T.type(np.linalg.norm(W.ravel(order="K"))) ** 2))
        if self.lambda_h > 0:
            r = T.type(r + T.type(T.type(0.5) * self.lambda_h) * T.type(T.type(np.linalg.norm(H.ravel(order="K"))) ** 2))
        return r

    def update_wh(self, s, X, W, H):  # :77-107
        if self.update_H:
            WtW = adddiag(_mm(W.T, W), self.lambda_h)     # :91
            H[...] = _mm(W.T, X)                          # :92
            H[...] = pdsolve(WtW, H)                      # :93
            projectnn(H)                                  # :94
        HHt = adddiag(_mm(H, H.T), self.lambda_w)         # :99
        XHt = _mm(X, H.T)                                 # :100
        W[...] = pdrsolve(XHt, HHt)                       # :101
        projectnn(W)                                      # :102
        s["WH"] = _mm(W, H)                               # :105


def solve_projals(alg: ProjectedALS, X, W, H, log=None) -> Result:
    """projals.jl:37-39"""
    upd = ProjectedALSUpd(alg.T, alg.update_H, alg.lambda_w, alg.lambda_h)
    return nmf_skeleton(upd, X, W, H, alg.maxiter, alg.verbose, alg.tol, log=log)


# --------------------------------------------------------------------------------------------------
# coorddesc.jl
# --------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


class ShufflePerm:
    """Stand-in for `randperm(n_components)` (coorddesc.jl:131-132).  Julia's global RNG stream cannot be
    reproduced outside Julia, so the project defines its own: splitmix64 seeded by the caller, Fisher-Yates from the
    top.  The same generator is implemented in libnmfb200 (csrc/simt_engine.cu, `ShufflePerm`); one permutation is
    drawn per call of _update_coord_descent! (W-step, then H-step), the state carries over within a solve."""

    def __init__(self, seed: int):
        self.s = int(seed) & _M64

    def _next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def perm(self, k: int) -> np.ndarray:
        a = np.arange(k, dtype=np.int64)
        for i in range(k - 1, 0, -1):
            j = self._next() % (i + 1)
            a[i], a[j] = a[j], a[i]
        return a


class CoordinateDescent:
    """coorddesc.jl:24-46 (no validation in the reference constructor).  `seed` replaces Julia's global RNG for
    shuffle=true."""

    def __init__(self, T=np.float64, maxiter=100, verbose=False, tol=None, update_H=True, alpha=0.0, regularization="both",
                 l1ratio=0.0, shuffle=False, seed=0):
        T = np.dtype(T)
        self.T = T
        self.maxiter = int(maxiter)
        self.verbose = bool(verbose)
        self.tol = T.type(np.cbrt(np.finfo(T).eps) if tol is None else tol)
        self.update_H = bool(update_H)
        self.alpha = T.type(alpha)
        self.l1ratio = T.type(l1ratio)
        self.regularization = regularization
        self.shuffle = bool(shuffle)
        self.seed = int(seed)


class CoordinateDescentUpd:
    """coorddesc.jl:54-181"""

    def __init__(self, T, alpha, l1ratio, regularization, shuffle, update_H, seed=0):
        T = np.dtype(T)
        aW = aH = T.type(0)
        if regularization in ("both", "components"):        # :65-67
            aH = T.type(alpha)
        if regularization in ("both", "transformation"):    # :69-71
            aW = T.type(alpha)
        one = T.type(1)
        self.T = T
        self.l1W, self.l2W = T.type(aW * l1ratio), T.type(aW * T.type(one - l1ratio))   # :73-76
        self.l1H, self.l2H = T.type(aH * l1ratio), T.type(aH * T.type(one - l1ratio))
        self.shuffle, self.update_H = shuffle, update_H
        self.rng = ShufflePerm(seed)
        self.violation = T.type(0)

    def prepare_state(self, X, W, H):  # :100
        nmf_checksize(X, W, H)
        return {}

    def evaluate_objv(self, s, X, W, H):  # :102-105
        return self.T.type(0.5) * self.T.type(sqL2dist(X, _mm(W, H)))

    def _update(self, X, F, Ot, l1, l2):
        """_update_coord_descent! :108-160 for F (rows x k) against Ot (cols x k); X is rows x cols."""
        T = self.T
        k = F.shape[1]
        HHt = _mm(Ot.T, Ot)                      # :112
        XHt = _mm(X, Ot)                         # :118
        if l2 > 0.0:
            HHt[np.diag_indices(k)] += T.type(l2)    # :123-125
        if l1 > 0.0:
            XHt -= T.type(l1)                        # :126-128
        perm = self.rng.perm(k) if self.shuffle else np.arange(k, dtype=np.int64)   # :129-133
        return _fn("oracle_cd_sweep", T)(_p(F), _p(_F(HHt)), _p(_F(XHt)), F.shape[0], k, _p(perm))

    def update_wh(self, s, X, W, H):  # :163-181
        Ht = _F(H.T)
        v = self._update(X, W, Ht, self.l1W, self.l2W)                 # :168
        if self.update_H:
            v = self.T.type(v + self._update(X.T, Ht, W, self.l1H, self.l2H))   # :171-176
            H[...] = Ht.T
        self.violation = v


def solve_cd(alg: CoordinateDescent, X, W, H, log=None) -> Result:
    """coorddesc.jl:49-51"""
    upd = CoordinateDescentUpd(alg.T, alg.alpha, alg.l1ratio, alg.regularization, alg.shuffle, alg.update_H, alg.seed)
    return nmf_skeleton(upd, X, W, H, alg.maxiter, alg.verbose, alg.tol, log=log)


# --------------------------------------------------------------------------------------------------
# alspgrad.jl
# --------------------------------------------------------------------------------------------------
def projgradnorm(g, x):
    """alspgrad.jl:9-19"""
    g = np.ascontiguousarray(g.ravel(order="K"))
    x = np.ascontiguousarray(x.ravel(order="K"))
    return g.dtype.type(_fn("oracle_projgradnorm", g.dtype)(_p(g), _p(x), g.size))


def _dot(a, b):
    """BLAS.dot on the dense storage (alspgrad.jl:141,143)."""
    return a.dtype.type(np.dot(a.ravel(order="K"), b.ravel(order="K")))


def _alspgrad_sub(F, gram, cross, left: bool, maxiter, traceiter, tolg, beta, sigma):
    """_alspgrad_updateh! (alspgrad.jl:86-191, left=True: G = gram*F - cross) and _alspgrad_updatew!
    (alspgrad.jl:242-347, left=False: G = F*gram - cross).  F is updated in place; returns the number of
    sub-iterations t.  Both routines are the same algorithm up to the side the k x k Gram multiplies from."""
    T = F.dtype
    assert _isF(F)
    Fn = np.empty_like(F, order="F")
    Fp = np.empty_like(F, order="F")
    D = np.empty_like(F, order="F")
    step = _fn("oracle_pg_step", T)
    t = 0
    converged = False
    decr_alpha = True
    alpha = T.type(1)                                   # :112
    mul = (lambda M: _mm(gram, M)) if left else (lambda M: _mm(M, gram))
    while not converged and t < maxiter:
        t += 1
        G = mul(F)                                      # :117
        G -= cross                                      # :118-120
        pgnrm = projgradnorm(G, F)                      # :123
        if pgnrm < tolg:
            converged = True
        it = 0
        if not converged:
            while it < traceiter:
                it += 1
                if not np.isfinite(alpha):
                    raise FloatingPointError("alpha is not finite")     # :132 `error("α is not finite")`
                step(_p(F), _p(G), T.type(alpha), _p(Fn), _p(D), F.size)   # :134-139
                dv1 = _dot(G, D)                        # :142
                GD = mul(D)                             # :143
                dv2 = _dot(GD, D)                       # :144
                suff_decr = T.type(T.type(T.type(1 - sigma) * dv1) + T.type(T.type(0.5) * dv2)) < 0   # :147
                if it == 1:
                    decr_alpha = not suff_decr          # :150
                    np.copyto(Fp, F)                    # :151
                if decr_alpha:
                    if suff_decr:
                        np.copyto(F, Fn)                # :156
                        break
                    alpha = T.type(alpha * beta)        # :159
                else:
                    # isapprox(Hp, Hn, atol=eps(T)): rtol defaults to 0 when atol > 0  => norm(Hp - Hn) <= eps(T)
                    close = T.type(np.linalg.norm((Fp - Fn).ravel(order="K"))) <= np.finfo(T).eps
                    if (not suff_decr) or close:        # :162
                        np.copyto(F, Fp)                # :163
                        break
                    alpha = T.type(alpha / beta)        # :166
                    np.copyto(Fp, Fn)                   # :167
    return t


def alspgrad_updateh(X, W, H, maxiter=1000, traceiter=20, tolg=None, beta=0.2, sigma=0.01):
    """alspgrad.jl:64-84"""
    T = H.dtype
    tolg = T.type(np.cbrt(np.finfo(T).eps) if tolg is None else tolg)
    return _alspgrad_sub(H, _mm(W.T, W), _mm(W.T, X), True, maxiter, traceiter, tolg, T.type(beta), T.type(sigma))


def alspgrad_updatew(X, W, H, maxiter=1000, traceiter=20, tolg=None, beta=0.2, sigma=0.01):
    """alspgrad.jl:221-240"""
    T = W.dtype
    tolg = T.type(np.cbrt(np.finfo(T).eps) if tolg is None else tolg)
    return _alspgrad_sub(W, _mm(H, H.T), _mm(X, H.T), False, maxiter, traceiter, tolg, T.type(beta), T.type(sigma))


class ALSPGrad:
    """alspgrad.jl:352-373"""

    def __init__(self, T=np.float64, maxiter=100, maxsubiter=200, tol=None, tolg=None, update_H=True, verbose=False):
        T = np.dtype(T)
        eps = np.finfo(T).eps
        self.T = T
        self.maxiter, self.maxsubiter = int(maxiter), int(maxsubiter)
        self.tol = T.type(np.cbrt(eps) if tol is None else tol)
        self.tolg = T.type(eps ** 0.25 if tolg is None else tolg)
        self.update_H, self.verbose = bool(update_H), bool(verbose)


class ALSPGradUpd:
    """alspgrad.jl:375-425 (mutable: tolg shrinks by 10x whenever a sub-solve stops after one iteration)."""

    def __init__(self, T, update_H, maxsubiter, tolg):
        self.T, self.update_H, self.maxsubiter, self.tolg = np.dtype(T), update_H, maxsubiter, np.dtype(T).type(tolg)
        self.subiters = 0

    def prepare_state(self, X, W, H):  # :385-396
        nmf_checksize(X, W, H)
        return {"WH": _mm(W, H)}

    def evaluate_objv(self, s, X, W, H):  # :398
        return self.T.type(0.5) * self.T.type(sqL2dist(X, s["WH"]))

    def update_wh(self, s, X, W, H):  # :400-425
        T = self.T
        if self.update_H:
            itH = _alspgrad_sub(H, _mm(W.T, W), _mm(W.T, X), True, self.maxsubiter, 20, self.tolg, T.type(0.2), T.type(0.01))
            self.subiters += itH
            if itH == 1:
                self.tolg = T.type(self.tolg * 0.1)     # :409-411 (Float64 literal, stored back into ::T)
        itW = _alspgrad_sub(W, _mm(H, H.T), _mm(X, H.T), False, self.maxsubiter, 20, self.tolg, T.type(0.2), T.type(0.01))
        self.subiters += itW
        if itW == 1:
            self.tolg = T.type(self.tolg * 0.1)         # :419-421
        s["WH"] = _mm(W, H)                             # :424


def solve_alspgrad(alg: ALSPGrad, X, W, H, log=None) -> Result:
    """alspgrad.jl:381-383"""
    upd = ALSPGradUpd(alg.T, alg.update_H, alg.maxsubiter, alg.tolg)
    res = nmf_skeleton(upd, X, W, H, alg.maxiter, alg.verbose, alg.tol, log=log)
    res.subiters = upd.subiters
    res.tolg_final = upd.tolg
    return res


def solve(alg, X, W, H, log=None) -> Result:
    """NMF.solve!(alg, X, W, H) dispatch for the algorithm types on the accelerated path."""
    if isinstance(alg, MultUpdate):
        return solve_multupdate(alg, X, W, H, log=log)
    if isinstance(alg, GreedyCD):
        return solve_greedycd(alg, X, W, H, log=log)
    if isinstance(alg, ProjectedALS):
        return solve_projals(alg, X, W, H, log=log)
    if isinstance(alg, CoordinateDescent):
        return solve_cd(alg, X, W, H, log=log)
    if isinstance(alg, ALSPGrad):
        return solve_alspgrad(alg, X, W, H, log=log)
    raise TypeError(f"oracle has no restatement for {type(alg).__name__}")


# --------------------------------------------------------------------------------------------------
# initialization.jl / utils.jl / interf.jl
# --------------------------------------------------------------------------------------------------
def randinit(p, n, k, T, rng, normalize=False, zeroh=False):
    """initialization.jl:4-12 + utils.jl:26-32.  Julia's global RNG stream cannot be reproduced; the
    caller supplies a NumPy Generator (same distribution: U[0,1))."""
    T = np.dtype(T)
    W = _F(rng.random((p, k), dtype=np.float64), dtype=T)
    if normalize:
        for j in range(k):
            W[:, j] *= T.type(1) / W[:, j].sum(dtype=T)
    H = np.zeros((k, n), dtype=T, order="F") if zeroh else _F(rng.random((k, n), dtype=np.float64), dtype=T)
    return W, H


def rsvd(X, k, rng):
    """RandomizedLinAlg.rsvd(A, n) as NMF.jl calls it (initialization.jl:78; dependency un-vendored, compat "0.1",
    restated from its published source: `Q = rrange(A, n)` = thin Q of `A * randn(size(A, 2), n)` (no oversampling, no
    power iterations), then `svd_restricted`: `B = Q'A`, `svd(B)`, `U = Q * U_B`).  The Gaussian test matrix comes
    from the caller's NumPy Generator (Julia's global RNG stream cannot be reproduced).  Returns (U, s, V)."""
    X = np.asarray(X)
    Omega = rng.standard_normal((X.shape[1], k)).astype(X.dtype)
    Q, _ = np.linalg.qr(X @ Omega)
    Ub, s, Vt = np.linalg.svd(Q.T @ X, full_matrices=False)
    return (Q @ Ub)[:, :k], s[:k], Vt[:k, :].T


def _posnegnorm(x):
    """initialization.jl:103-115: zeros are counted with the negative part (contributing 0)."""
    T = x.dtype.type
    pn, nn = T(0), T(0)
    for xi in x:
        if xi > 0:
            pn = T(pn + xi * xi)
        else:
            nn = T(nn + xi * xi)
    return np.sqrt(pn), np.sqrt(nn)


def nndsvd(X, k, zeroh=False, variant="std", initdata=None, rng=None):
    """initialization.jl:26-101.  `initdata` = (U, S, V) of an SVD of X (the reference takes an `SVD` object and uses
    `.U[:,1:k], .S[1:k], .V[:,1:k]`); without it the factors come from `rsvd` above.  variant in {"std","a","ar"}.
    Element type follows X; sums are sequential in T like posnegnorm / mean."""
    X = _F(X)
    T = X.dtype
    p, n = X.shape
    if variant not in ("std", "a", "ar"):
        raise ArgumentError("Invalid value for variant")
    rng = rng if rng is not None else np.random.default_rng()
    if initdata is None:
        U, s, V = rsvd(X, k, rng)
    else:
        U, s, V = initdata[0][:, :k], initdata[1][:k], initdata[2][:, :k]
    U, s, V = U.astype(T), s.astype(T), V.astype(T)          # :31-33
    one = T.type(1)
    if variant == "std":
        v0 = T.type(0)
    elif variant == "a":
        v0 = T.type(X.mean(dtype=np.float64))                # convert(T, mean(X)), :37
    else:
        v0 = T.type(X.mean(dtype=np.float64) * 0.01)         # convert(T, mean(X) * 0.01), :37
    W = np.empty((p, k), dtype=T, order="F")
    Ht = np.empty((n, k), dtype=T, order="F")
    for j in range(k):
        x, y = U[:, j], V[:, j]
        xp, xn = _posnegnorm(x)
        yp, yn = _posnegnorm(y)
        mp, mn = T.type(xp * yp), T.type(xn * yn)
        vj = v0
        if variant == "ar":
            vj = T.type(vj * T.type(rng.random()))           # vj *= rand(T), :49-51
        if mp >= mn:                                         # :54-58
            ss = np.sqrt(T.type(s[j] * mp))
            W[:, j] = np.where(x > 0, x * T.type(ss / xp), vj)          # scalepos!, :117-126
            if not zeroh:
                Ht[:, j] = np.where(y > 0, y * T.type(ss / yp), vj)
        else:                                                # :59-63
            ss = np.sqrt(T.type(s[j] * mn))
            W[:, j] = np.where(x < 0, -(x * T.type(ss / xn)), vj)       # scaleneg!, :128-137
            if not zeroh:
                Ht[:, j] = np.where(y < 0, -(y * T.type(ss / yn)), vj)
    H = np.zeros((k, n), dtype=T, order="F") if zeroh else _F(Ht.T)      # :88-98
    return W, H


def solve_replicates(alg, X, W, H, replicates, initH, rng):
    """interf.jl:85-101"""
    ret = solve(alg, X, W, H)
    k = W.shape[1]
    minobjv = ret.objvalue
    for _ in range(2, replicates + 1):
        Wr, Hr = randinit(X.shape[0], X.shape[1], k, X.dtype, rng, normalize=True, zeroh=not initH)
        tmp = solve(alg, X, Wr, Hr)
        if minobjv > tmp.objvalue:
            ret, minobjv = tmp, tmp.objvalue
    return ret


def nnmf(X, k, init="nndsvdar", alg="greedycd", maxiter=100, tol=None, replicates=1, W0=None, H0=None,
         update_H=True, verbose=False, rng=None, initdata=None):
    """interf.jl:3-83 restricted to what the accelerated path covers: init in {:random, :custom, :nndsvd, :nndsvda, :nndsvdar},
    alg in {:multmse, :multdiv, :greedycd, :projals, :alspgrad, :cd}.  Validation order and messages follow the
    reference."""
    X = _F(X)
    T = X.dtype
    if tol is None:
        tol = np.cbrt(np.finfo(T).eps / 100)
    if not (X >= 0).all():
        raise ArgumentError("The elements of X must be non-negative.")
    p, n = X.shape
    if not k <= min(p, n):
        raise ArgumentError("The value of k should not exceed min(size(X)).")
    if not replicates >= 1:
        raise ArgumentError("The value of replicates must be positive.")
    if not update_H and init != "custom":
        warnings.warn("Only W will be updated.")
    if init == "custom":
        if W0 is None or H0 is None:
            raise ArgumentError("To use :custom initialization, set W0 and H0.")
        if not (np.asarray(W0) >= 0).all():
            raise ArgumentError("The elements of W0 must be non-negative.")
        if W0.shape != (p, k):
            raise ArgumentError("Invalid size for W0.")
        if not (np.asarray(H0) >= 0).all():
            raise ArgumentError("The elements of H0 must be non-negative.")
        if H0.shape != (k, n):
            raise ArgumentError("Invalid size for H0.")
    elif W0 is not None or H0 is not None:
        warnings.warn("Ignore W0 and H0 except for :custom initialization.")
    initH = alg != "projals"
    rng = rng if rng is not None else np.random.default_rng()
    if init == "random":
        W, H = randinit(p, n, k, T, rng, normalize=True, zeroh=not initH)
    elif init == "custom":
        W, H = _F(W0, dtype=T), _F(H0, dtype=T)
    elif init in ("nndsvd", "nndsvda", "nndsvdar"):          # interf.jl:44-49
        W, H = nndsvd(X, k, zeroh=not initH, variant={"nndsvd": "std", "nndsvda": "a", "nndsvdar": "ar"}[init], initdata=initdata, rng=rng)
    elif init == "spa":
        raise NotImplementedError("init=:spa is outside the restated hot path (SURVEY.md section 8f)")
    else:
        raise ArgumentError("Invalid value for init.")
    if alg == "multmse":
        inst = MultUpdate(T, obj="mse", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "multdiv":
        inst = MultUpdate(T, obj="div", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "greedycd":
        inst = GreedyCD(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "projals":
        inst = ProjectedALS(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "alspgrad":
        inst = ALSPGrad(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "cd":
        inst = CoordinateDescent(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "spa":
        raise NotImplementedError("alg=:spa is outside the restated hot path (SURVEY.md section 8f)")
    else:
        raise ArgumentError("Invalid algorithm.")
    return solve_replicates(inst, X, W, H, replicates, initH, rng)


# --------------------------------------------------------------------------------------------------
# fixtures of the reference's own tests
# --------------------------------------------------------------------------------------------------
def laurberg6x3(alpha, T=np.float64):
    """test/testproblems.jl:6-13 -> (X, W, H) with X = W*H, W = H'."""
    a = alpha
    H = np.array([[a, 1, 1, a, 0, 0], [1, a, 0, 0, a, 1], [0, 0, a, 1, 1, a]], dtype=T)
    W = H.T.copy()
    X = W @ H
    return _F(X), _F(W), _F(H)
