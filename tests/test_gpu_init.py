"""NNDSVD initialisation (initialization.jl:26-101) with the X-sized products of the range finder on the GPU
(nmfb200_mul_X_*), and nnmf() with the reference's default arguments (init=:nndsvdar, alg=:greedycd, interf.jl:3-13).
Tolerances: mul_X is an fp32 / fp64 CUDA-core GEMM with a different summation order than OpenBLAS (fp32 1e-5, fp64 1e-12
relative to the row scale); the derived factors inherit it through a QR and an SVD (fp64 1e-8, fp32 2e-3)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _planted(rng, p, n, k, T):
    X = np.maximum(rng.random((p, k)) - 0.3, 0) @ np.maximum(rng.random((k, n)) - 0.3, 0)
    return np.asfortranarray(X, dtype=T)


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_mul_X_matches_numpy(NMF, T, tol):
    rng = np.random.default_rng(0)
    for (p, n, c) in [(200, 150, 5), (515, 1030, 37), (64, 4100, 128)]:
        X = np.asfortranarray(rng.random((p, n)), dtype=T)
        with NMF.Session() as s:
            s.set_X(X)
            B = rng.standard_normal((n, c)).astype(T)
            C = s.mul_X(B)
            ref = X.astype(np.float64) @ B.astype(np.float64)
            assert C.shape == (p, c) and C.dtype == T
            assert np.abs(C - ref).max() <= tol * np.abs(ref).max() * 4
            B2 = rng.standard_normal((p, c)).astype(T)
            C2 = s.mul_X(B2, transpose=True)
            ref2 = X.astype(np.float64).T @ B2.astype(np.float64)
            assert C2.shape == (n, c) and np.abs(C2 - ref2).max() <= tol * np.abs(ref2).max() * 4
            with pytest.raises(NMF.DimensionMismatch):
                s.mul_X(np.ones((n + 1, 2), dtype=T))


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-8), (np.float32, 2e-3)])
def test_nndsvd_on_gpu_matches_oracle(NMF, oracle, T, tol):
    rng = np.random.default_rng(1)
    X = np.asfortranarray(rng.random((300, 220)), dtype=T)
    for variant in ("std", "a", "ar"):
        for zeroh in (False, True):
            W, H = NMF.nndsvd(X, 6, variant=variant, zeroh=zeroh, rng=np.random.default_rng(7))
            Wo, Ho = oracle.nndsvd(X, 6, variant=variant, zeroh=zeroh, rng=np.random.default_rng(7))
            assert (W >= 0).all() and (H >= 0).all() and W.dtype == T and H.shape == (6, 220)
            assert np.linalg.norm(W - Wo) <= tol * np.linalg.norm(Wo)
            assert np.linalg.norm(H - Ho) <= tol * max(np.linalg.norm(Ho), 1e-300)
            if zeroh:
                assert (H == 0).all()


def test_nnmf_reference_defaults_run_on_the_gpu(NMF, oracle):
    """nnmf(X, k) exactly as a user of the reference calls it: init=:nndsvdar, alg=:greedycd, maxiter=100,
    tol=cbrt(eps(T)/100) (interf.jl:3-13).  Float64 keeps the greedy coordinate choices identical to the oracle's."""
    rng = np.random.default_rng(2)
    X = _planted(rng, 120, 90, 4, np.float64)
    r = NMF.nnmf(X, 4, rng=np.random.default_rng(11))
    ro = oracle.nnmf(X, 4, rng=np.random.default_rng(11))
    # the range finder's products differ from OpenBLAS's in the last bits, so the start differs by ~1e-12 and the greedy
    # trajectories may part on a near-tie: compare the outcome, not the path
    assert r.converged == ro.converged and abs(r.niters - ro.niters) <= 3
    assert np.linalg.norm(X - r.W @ r.H) <= 1e-4 * np.linalg.norm(X) and np.linalg.norm(X - ro.W @ ro.H) <= 1e-4 * np.linalg.norm(X)
    assert np.linalg.norm(r.W @ r.H - ro.W @ ro.H) <= 1e-4 * np.linalg.norm(X)
    # the same call in Float32 with every NNDSVD variant and every iterative algorithm of interf.jl:60-71
    Xf = X.astype(np.float32)
    for init in ("nndsvd", "nndsvda", "nndsvdar"):
        for alg in ("multmse", "multdiv", "greedycd", "projals", "alspgrad", "cd"):
            r = NMF.nnmf(Xf, 4, init=init, alg=alg, maxiter=200, rng=np.random.default_rng(5))
            assert np.isfinite(r.W).all() and np.isfinite(r.H).all() and (r.W >= 0).all() and (r.H >= 0).all()
            assert np.linalg.norm(Xf - r.W @ r.H) <= 0.2 * np.linalg.norm(Xf), (init, alg)   # NNDSVD zeros are absorbing under MU
    # initdata = an SVD computed by the caller (test/interf.jl:17-21)
    U, s, Vt = np.linalg.svd(X, full_matrices=False)
    r = NMF.nnmf(X, 4, alg="multmse", init="nndsvd", initdata=(U, s, Vt.T), maxiter=300)
    ro = oracle.nnmf(X, 4, alg="multmse", init="nndsvd", initdata=(U, s, Vt.T), maxiter=300)
    assert r.niters == ro.niters
    assert abs(float(r.objvalue) - float(ro.objvalue)) <= 1e-8 * float(ro.objvalue) + 1e-14
