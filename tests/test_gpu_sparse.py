"""Sparse X (README.md:22 "Sparse NMF"): a scipy.sparse matrix goes through nmfb200_set_X_csc_* -- its CSC arrays cross
PCIe and are expanded on the device -- and every solver then sees exactly the dense matrix.  Checked: bit-identical results to the
dense upload of X.toarray(), parity with the oracle on the dense matrix, the reference's argument checks, duplicates, empty columns."""
import numpy as np
import pytest

sp = pytest.importorskip("scipy.sparse")
pytestmark = pytest.mark.gpu


def _sparse_problem(p, n, density, dtype, seed):
    rng = np.random.default_rng(seed)
    X = sp.random(p, n, density=density, format="csc", dtype=dtype, random_state=rng, data_rvs=lambda m: rng.random(m).astype(dtype))
    return X, rng


@pytest.mark.parametrize("alg", ["multmse", "multdiv", "greedycd", "cd"])
@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_sparse_X_equals_dense_X(NMF, oracle, alg, T):
    p, n, k = 96, 130, 5
    X, rng = _sparse_problem(p, n, 0.15, T, seed=31)
    Xd = X.toarray()
    Xd[:, 7] = 0           # an empty column
    X = sp.csc_matrix(Xd)
    Xd = np.asfortranarray(Xd)
    W0, H0 = NMF.randinit(p, n, k, T, normalize=True, rng=rng)
    mk = {"multmse": lambda M: M.MultUpdate(T, obj="mse", maxiter=15, tol=1e-9),
          "multdiv": lambda M: M.MultUpdate(T, obj="div", maxiter=15, tol=1e-9),
          "greedycd": lambda M: M.GreedyCD(T, maxiter=8, tol=1e-9),
          "cd": lambda M: M.CoordinateDescent(T, maxiter=8, tol=1e-9)}[alg]
    Ws, Hs, Wd, Hd, Wo, Ho = (a.copy(order="F") for a in (W0, H0, W0, H0, W0, H0))
    rs = NMF.solve(mk(NMF), X, Ws, Hs, engine="simt")
    rd = NMF.solve(mk(NMF), Xd, Wd, Hd, engine="simt")
    ro = oracle.solve(mk(oracle), Xd, Wo, Ho)
    assert (Ws == Wd).all() and (Hs == Hd).all() and rs.objvalue == rd.objvalue and rs.niters == rd.niters   # the same dense matrix on the GPU
    tol = 1e-9 if T == np.float64 else 2e-4
    assert np.linalg.norm(Ws - Wo) <= tol * np.linalg.norm(Wo) and np.linalg.norm(Hs - Ho) <= tol * np.linalg.norm(Ho)
    assert abs(float(rs.objvalue) - float(ro.objvalue)) <= max(tol, 1e-4 if T == np.float32 else 0) * float(ro.objvalue)


def test_sparse_X_on_the_tensor_core_engine(NMF, oracle):
    p, n, k = 1024, 1280, 32
    X, rng = _sparse_problem(p, n, 0.05, np.float32, seed=33)
    Xd = np.asfortranarray(X.toarray())
    W0, H0 = NMF.randinit(p, n, k, np.float32, normalize=True, rng=rng)
    Ws, Hs, Wd, Hd, Wo, Ho = (a.copy(order="F") for a in (W0, H0, W0, H0, W0, H0))
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=10, tol=1e-9)
    rs = NMF.solve(alg, X, Ws, Hs)            # engine=auto: 2^20 cells and more -> tcgen05 kernels
    rd = NMF.solve(alg, Xd, Wd, Hd)
    assert rs.info["engine"] == "tc" and (Ws == Wd).all() and (Hs == Hd).all() and rs.objvalue == rd.objvalue
    ro = oracle.solve(oracle.MultUpdate(np.float32, obj="mse", maxiter=10, tol=1e-9), Xd, Wo, Ho)
    assert np.linalg.norm(Ws - Wo) <= 5e-3 * np.linalg.norm(Wo) and np.linalg.norm(Hs - Ho) <= 5e-3 * np.linalg.norm(Ho)
    assert abs(float(rs.objvalue) - float(ro.objvalue)) <= 1e-4 * float(ro.objvalue)


def test_nnmf_with_sparse_X_default_arguments(NMF):
    X, rng = _sparse_problem(60, 80, 0.3, np.float64, seed=35)
    r = NMF.nnmf(X, 4, maxiter=30, rng=rng)                     # init=:nndsvdar, alg=:greedycd (interf.jl:3-13)
    rd = NMF.nnmf(np.asfortranarray(X.toarray()), 4, maxiter=30, rng=np.random.default_rng(0))
    assert r.W.shape == (60, 4) and r.H.shape == (4, 80) and (r.W >= 0).all() and (r.H >= 0).all()
    assert float(r.objvalue) <= 1.2 * float(rd.objvalue)        # different random range finder draws, same quality
    Xneg = X.copy()
    Xneg.data[3] = -1.0
    with pytest.raises(NMF.ArgumentError, match="non-negative"):
        NMF.nnmf(Xneg, 4)
    with NMF.Session() as s:
        with pytest.raises(NMF.ArgumentError, match="non-negative"):
            s.set_X(Xneg, check_nonneg=True)
        s.set_X(X, check_nonneg=True)


def test_csc_duplicates_are_summed_and_bad_indices_rejected(NMF):
    import ctypes
    rows = np.array([0, 2, 2, 1], dtype=np.int64)
    cols = np.array([0, 0, 0, 2], dtype=np.int64)
    vals = np.array([1.0, 2.0, 3.0, 4.0])
    X = sp.csc_matrix((vals, (rows, cols)), shape=(3, 4))       # scipy sums the duplicate (2, 0) on construction ...
    raw = sp.csc_matrix((3, 4))
    raw.indptr = np.array([0, 3, 3, 4, 4], dtype=np.int64)      # ... so hand the library the un-merged arrays directly
    raw.indices, raw.data = rows.copy(), vals.copy()
    with NMF.Session(engine="simt") as s:
        s.set_X(raw)
        Y = s.mul_X(np.eye(4), transpose=False)                 # X * I
        assert np.array_equal(Y, X.toarray())
        lib = s._lib
        bad_rows = np.array([0, 5, 2, 1], dtype=np.int64)
        st = lib.nmfb200_set_X_csc_f64(s._h, raw.indptr.ctypes.data_as(ctypes.c_void_p), bad_rows.ctypes.data_as(ctypes.c_void_p),
                                       vals.ctypes.data_as(ctypes.c_void_p), 3, 4, 0, 0)
        assert st == NMF._lib.EINVAL
        bad_ptr = np.array([0, 3, 2, 4, 4], dtype=np.int64)
        st = lib.nmfb200_set_X_csc_f64(s._h, bad_ptr.ctypes.data_as(ctypes.c_void_p), rows.ctypes.data_as(ctypes.c_void_p),
                                       vals.ctypes.data_as(ctypes.c_void_p), 3, 4, 0, 0)
        assert st == NMF._lib.EINVAL
        one_based = lib.nmfb200_set_X_csc_f64(s._h, (raw.indptr + 1).ctypes.data_as(ctypes.c_void_p), (rows + 1).ctypes.data_as(ctypes.c_void_p),
                                              vals.ctypes.data_as(ctypes.c_void_p), 3, 4, 1, 0)   # Julia's 1-based arrays
        assert one_based == NMF._lib.OK
        s.dtype, s.shape = np.dtype(np.float64), (3, 4)
        assert np.array_equal(s.mul_X(np.eye(4)), X.toarray())
