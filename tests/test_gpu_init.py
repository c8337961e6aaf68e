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


# ---- randinit on the device (nmfb200_randinit_*): counter-based Philox4x32-10, regenerated here on the host --------------------
def _philox_words(e, stream, seed):
    """Words 0 and 1 of Philox4x32-10 with counter (e_lo, e_hi, stream, 0) and key (seed_lo, seed_hi); e: uint64 array."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c0, c1 = e & mask, e >> np.uint64(32)
    c2, c3 = np.full_like(e, stream), np.zeros_like(e)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                        # 32 x 32 -> 64 bit products
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0), p1 & mask, (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1), p0 & mask
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1


def _philox_uniform(shape, T, stream, seed, e0=0, e_ld=None):
    rows, cols = shape
    e_ld = rows if e_ld is None else e_ld
    i, j = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange(cols, dtype=np.uint64), indexing="ij")
    a, b = _philox_words(np.uint64(e0) + i + j * np.uint64(e_ld), stream, seed)
    if np.dtype(T) == np.float32:
        return ((a >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (((a << np.uint64(32)) | b) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_device_randinit_is_the_documented_philox_stream(NMF, T):
    p, n, k, seed = 301, 257, 7, 0x1234567890ABCDEF
    X = np.asfortranarray(np.random.default_rng(0).random((p, n)), dtype=T)
    with NMF.Session() as s:
        s.set_X(X)
        W, H = s.randinit(k, seed=seed)
        assert (W == _philox_uniform((p, k), T, 0, seed)).all() and (H == _philox_uniform((k, n), T, 1, seed)).all()
        assert W.min() >= 0 and W.max() < 1 and H.min() >= 0 and H.max() < 1 and abs(W.mean() - 0.5) < 0.05
        Wn, Hz = s.randinit(k, seed=seed, normalize=True, zeroh=True)                 # what nnmf asks for with alg=:projals
        np.testing.assert_allclose(Wn.sum(axis=0), 1.0, rtol=1e-5 if T == np.float32 else 1e-12)   # test/initialization.jl:22-27
        np.testing.assert_allclose(Wn * W.sum(axis=0, keepdims=True, dtype=np.float64), W, rtol=2e-5 if T == np.float32 else 1e-12)
        assert (Hz == 0).all()
    # a row shard draws the rows of the unsharded W (row_offset / p_total) and the same H
    lo, hi = 100, 230
    with NMF.Session() as s:
        s.set_X(np.asfortranarray(X[lo:hi]))
        Ws, Hs = s.randinit(k, seed=seed, row_offset=lo, p_total=p)
    assert (Ws == W[lo:hi]).all() and (Hs == H).all()


def test_nnmf_with_device_seed(NMF, oracle):
    """nnmf(..., init=:random, rng=<int>): the initial factors and the random restarts of `replicates` come from the device
    generator; the run is reproducible and equals a solve from the same (regenerated) factors."""
    rng = np.random.default_rng(3)
    X = np.asfortranarray(rng.random((96, 80)), dtype=np.float64)
    r1 = NMF.nnmf(X, 4, init="random", alg="multmse", maxiter=30, rng=77, replicates=3)
    r2 = NMF.nnmf(X, 4, init="random", alg="multmse", maxiter=30, rng=77, replicates=3)
    assert r1 == r2 and np.isfinite(float(r1.objvalue))
    objs = []
    for rep in range(3):
        W0 = _philox_uniform((96, 4), np.float64, 0, 77 + rep)
        W0 = np.asfortranarray(W0 * (1.0 / W0.sum(axis=0, keepdims=True)))
        H0 = np.asfortranarray(_philox_uniform((4, 80), np.float64, 1, 77 + rep))
        ro = oracle.solve(oracle.MultUpdate(np.float64, maxiter=30, tol=np.cbrt(np.finfo(np.float64).eps / 100)), X, W0, H0)
        objs.append(float(ro.objvalue))
    assert abs(float(r1.objvalue) - min(objs)) <= 1e-9 * min(objs)          # interf.jl:91-98 keeps the best replicate


# ---- NNDSVD entirely on the device (nmfb200_rsvd_* / nmfb200_nndsvd_*, csrc/init_device.cuh) -------------------------------------
def _philox_normal(rows, cols, T, seed):
    """The Gaussian test matrix of the device range finder: Philox stream 2, element e = i + j*rows, Box-Muller in Float64."""
    i, j = np.meshgrid(np.arange(rows, dtype=np.uint64), np.arange(cols, dtype=np.uint64), indexing="ij")
    a, b = _philox_words(i + j * np.uint64(rows), 2, seed)
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -32
    u2 = b.astype(np.float64) * 2.0 ** -32
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(T)


def _separated(rng, p, n, r, T):
    """Non-negative X with well separated leading singular values (so individual singular vectors are well conditioned)."""
    A = np.maximum(rng.random((p, r)) - 0.3, 0) * (2.0 ** -np.arange(r))
    B = np.maximum(rng.random((r, n)) - 0.3, 0)
    return np.asfortranarray(A @ B + 1e-3 * rng.random((p, n)), dtype=T)


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-9), (np.float32, 5e-4)])
def test_device_rsvd_matches_the_host_range_finder(NMF, T, tol):
    rng = np.random.default_rng(41)
    p, n, k, seed = 300, 220, 6, 2024
    X = _separated(rng, p, n, 8, T)
    with NMF.Session() as s:
        s.set_X(X)
        U, S, V = s.rsvd(k, seed=seed)
        U2, S2, V2 = s.rsvd(k, seed=seed)
    assert (U == U2).all() and (S == S2).all() and (V == V2).all()          # counter-based draw, fixed reduction orders
    # the same algorithm on the host from the regenerated test matrix (initialization.jl:78: qr(X * randn).Q, svd(Q' * X))
    Om = _philox_normal(n, k, T, seed).astype(np.float64)
    X64 = X.astype(np.float64)
    Q, _ = np.linalg.qr(X64 @ Om)
    Ub, s_ref, Vt = np.linalg.svd(Q.T @ X64, full_matrices=False)
    U_ref, V_ref = Q @ Ub, Vt.T
    assert (np.diff(S) <= 0).all()
    np.testing.assert_allclose(S, s_ref, rtol=tol)
    eye_tol = 1e-10 if T == np.float64 else 1e-5
    assert np.abs(U.astype(np.float64).T @ U - np.eye(k)).max() <= eye_tol and np.abs(V.astype(np.float64).T @ V - np.eye(k)).max() <= eye_tol
    for j in range(k):                                                     # singular pairs agree up to their joint sign
        su, sv = np.sign(U[:, j] @ U_ref[:, j]), np.sign(V[:, j] @ V_ref[:, j])
        assert su == sv
        assert np.linalg.norm(su * U[:, j] - U_ref[:, j]) <= 50 * tol and np.linalg.norm(sv * V[:, j] - V_ref[:, j]) <= 50 * tol
    rec, rec_ref = (U * S) @ V.T, (U_ref * s_ref) @ V_ref.T
    assert np.linalg.norm(rec - rec_ref) <= tol * np.linalg.norm(rec_ref)


@pytest.mark.parametrize("T,tol", [(np.float64, 1e-10), (np.float32, 2e-5)])
def test_device_nndsvd_matches_oracle_split_of_the_same_triplets(NMF, oracle, T, tol):
    rng = np.random.default_rng(43)
    p, n, k, seed = 260, 310, 5, 99
    X = _separated(rng, p, n, 7, T)
    with NMF.Session() as s:
        s.set_X(X)
        U, S, V = s.rsvd(k, seed=seed)
        out = {(v, z): s.nndsvd(k, variant=v, zeroh=z, seed=seed) for v in ("std", "a", "ar") for z in (False, True)}
    for v in ("std", "a"):
        for z in (False, True):
            W, H = out[(v, z)]
            Wo, Ho = oracle.nndsvd(X, k, zeroh=z, variant=v, initdata=(U, S, V))     # _nndsvd! on the device's own triplets
            assert (W >= 0).all() and (H >= 0).all() and W.dtype == T and H.shape == (k, n)   # test/initialization.jl:29-53
            assert np.linalg.norm(W - Wo) <= tol * np.linalg.norm(Wo)
            assert np.linalg.norm(H - Ho) <= tol * max(np.linalg.norm(Ho), 1e-300)
            if z:
                assert (H == 0).all()
    # :nndsvdar -- the fill of column j is convert(T, mean(X) * 0.01) * rand(T) with rand = Philox stream 3, element j
    Ws, Hs = out[("std", False)]
    War, Har = out[("ar", False)]
    v0 = T(X.mean(dtype=np.float64) * 0.01)
    fill = (v0 * _philox_uniform((k, 1), T, 3, seed)[:, 0]).astype(T)
    for j in range(k):
        zw, zh = Ws[:, j] == 0, Hs[j, :] == 0
        assert (War[~zw, j] == Ws[~zw, j]).all() and (Har[j, ~zh] == Hs[j, ~zh]).all()
        np.testing.assert_allclose(War[zw, j], fill[j], rtol=1e-6)
        np.testing.assert_allclose(Har[j, zh], fill[j], rtol=1e-6)


def test_nnmf_with_device_seed_runs_nndsvd_on_the_gpu(NMF):
    rng = np.random.default_rng(47)
    for T in (np.float64, np.float32):
        X = _separated(rng, 200, 160, 6, T)
        r = NMF.nnmf(X, 5, maxiter=60, rng=123)                       # defaults: init=:nndsvdar, alg=:greedycd (interf.jl:3-13)
        r_again = NMF.nnmf(X, 5, maxiter=60, rng=123)
        r_host = NMF.nnmf(X, 5, maxiter=60, rng=np.random.default_rng(123))
        assert r == r_again
        assert (r.W >= 0).all() and (r.H >= 0).all() and np.isfinite(float(r.objvalue))
        assert float(r.objvalue) <= 2.0 * float(r_host.objvalue) + 1e-12     # another test matrix, the same quality of start
    # k > rank(X): the sample X * randn(n, k) is rank deficient -- the device QR refuses, nnmf falls back to the host QR
    Xr = np.asfortranarray(np.maximum(rng.random((80, 3)) - 0.3, 0) @ np.maximum(rng.random((3, 70)) - 0.3, 0))
    with NMF.Session() as s:
        s.set_X(Xr)
        with pytest.raises(NMF.NumericalError):
            s.rsvd(5, seed=1)
    r = NMF.nnmf(Xr, 5, maxiter=100, rng=1, alg="multmse")
    assert np.isfinite(r.W).all() and np.linalg.norm(Xr - r.W @ r.H) <= 0.2 * np.linalg.norm(Xr)
