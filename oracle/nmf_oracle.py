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


def solve(alg, X, W, H, log=None) -> Result:
    """NMF.solve!(alg, X, W, H) dispatch for the algorithm types on the accelerated path."""
    if isinstance(alg, MultUpdate):
        return solve_multupdate(alg, X, W, H, log=log)
    if isinstance(alg, GreedyCD):
        return solve_greedycd(alg, X, W, H, log=log)
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
         update_H=True, verbose=False, rng=None):
    """interf.jl:3-83 restricted to what the accelerated path covers: init in {:random, :custom},
    alg in {:multmse, :multdiv, :greedycd}.  Validation order and messages follow the reference."""
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
    elif init in ("nndsvd", "nndsvda", "nndsvdar", "spa"):
        raise NotImplementedError(f"init=:{init} is outside the restated hot path (SURVEY.md section 8f)")
    else:
        raise ArgumentError("Invalid value for init.")
    if alg == "multmse":
        inst = MultUpdate(T, obj="mse", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "multdiv":
        inst = MultUpdate(T, obj="div", maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg == "greedycd":
        inst = GreedyCD(T, maxiter=maxiter, tol=tol, verbose=verbose, update_H=update_H)
    elif alg in ("projals", "alspgrad", "cd", "spa"):
        raise NotImplementedError(f"alg=:{alg} is outside the restated hot path (SURVEY.md section 8f)")
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
