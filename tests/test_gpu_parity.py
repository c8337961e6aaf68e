"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Tolerances are written next to each assertion.  Run on the B200 box: pytest -m gpu."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _inst(mod, g):
    from test_oracle import golden_instance
    return golden_instance(mod, g)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_simt_engine_vs_golden(NMF, path):
    """Exact engine (fp32/fp64 CUDA cores) vs the committed oracle outputs."""
    g = np.load(path)
    T = g["X"].dtype
    W, H = np.asfortranarray(g["W0"]), np.asfortranarray(g["H0"])
    r = NMF.solve(_inst(NMF, g), np.asfortranarray(g["X"]), W, H, engine="simt")
    assert r.info["engine"] == "simt" and r.info["kernel_launches"] > 0
    assert r.niters == int(g["niters"]) and r.converged == bool(g["converged"])
    alg = str(g["alg"])
    # fp64: summation-order differences only; fp32: ~1e-6 per op compounded over <= 40 iterations
    tol = 1e-9 if T == np.float64 else 2e-4
    if alg == "greedycd":  # argmax decisions can flip on last-bit differences of the GEMMs
        tol = 1e-6 if T == np.float64 else 2e-3
    if alg == "projals":   # k x k solves amplify summation-order noise by cond(Gram + lambda I); GPU inverts in Float64
        tol = 1e-7 if T == np.float64 else 5e-3
    if alg == "alspgrad":  # Armijo decisions are discrete; dot products are accumulated in Float64 on the GPU
        tol = 1e-7 if T == np.float64 else 5e-3
    assert abs(float(r.objvalue) - float(g["objvalue"])) <= tol * abs(float(g["objvalue"]))
    if alg != "greedycd":
        assert _relerr(W, g["W"]) <= tol and _relerr(H, g["H"]) <= tol
    upd_H = bool(g["update_H"]) if "update_H" in g else bool(__import__("json").loads(str(g["opts"])).get("update_H", True))
    if not upd_H:
        assert (H == g["H0"]).all()  # test/interf.jl:35 -- bit-identical


def _start(oracle, T, rng):
    X, Wg, Hg = oracle.laurberg6x3(0.3, T)
    return X, np.asfortranarray(Wg + rng.random(Wg.shape).astype(T) * T(0.1)), Hg.copy(order="F")


# test/multupd.jl:4-21 on the GPU
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("obj", ["mse", "div"])
@pytest.mark.parametrize("lam", [0.0, 1e-4])
@pytest.mark.parametrize("engine", ["simt", "auto", "tc"])
def test_reference_kat_multupd_gpu(NMF, oracle, T, obj, lam, engine):
    """The reference's own known-answer test on every engine a user can get: the exact engine, the default (auto), and --
    Float32 :mse only, the one algorithm the tensor-core engine accepts at 6 x 6 -- the bf16 tensor-core engine forced."""
    if engine == "tc" and (T != np.float32 or obj != "mse"):
        pytest.skip("engine=tc covers Float32 only, and :div only from 128 x 128 on")
    X, W, H = _start(oracle, T, np.random.default_rng(21))
    with NMF.Session(engine=engine) as s:
        s.set_option("check_every", 64)
        s.set_X(X)
        s.solve(NMF.MultUpdate(T, obj=obj, maxiter=5000, tol=1e-9, lambda_w=lam, lambda_h=lam), W, H)
    assert (W >= 0).all() and (H >= 0).all() and not np.isnan(W).any() and not np.isnan(H).any()
    assert np.linalg.norm(X - W @ H) <= 1e-2


# test/greedycd.jl:5-20 on the GPU
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("lam", [0.0, 1e-5])
@pytest.mark.parametrize("engine", ["simt", "auto"])
def test_reference_kat_greedycd_gpu(NMF, oracle, T, lam, engine):
    X, W, H = _start(oracle, T, np.random.default_rng(22))
    NMF.solve(NMF.GreedyCD(T, maxiter=1000, tol=1e-9, lambda_w=lam, lambda_h=lam), X, W, H, engine=engine)
    assert (W >= 0).all() and (H >= 0).all() and not np.isnan(W).any() and not np.isnan(H).any()
    assert np.linalg.norm(X - W @ H) <= 1e-3


def test_greedycd_half_step_bit_exact(NMF, oracle):
    """With identical inputs one full GreedyCD iteration differs from the oracle only through GEMM
    summation order; on a tiny problem with exactly representable data the GEMMs are exact, so W and H
    must match bit for bit (integer-valued inputs scaled by powers of two)."""
    rng = np.random.default_rng(5)
    p, n, k = 24, 20, 4
    T = np.float32
    X = np.asfortranarray(rng.integers(0, 8, (p, n)) / 8.0, dtype=T)
    W0 = np.asfortranarray(rng.integers(0, 8, (p, k)) / 8.0, dtype=T)
    H0 = np.asfortranarray(rng.integers(0, 8, (k, n)) / 8.0, dtype=T)
    Wg, Hg = W0.copy(order="F"), H0.copy(order="F")
    Wo, Ho = W0.copy(order="F"), H0.copy(order="F")
    r = NMF.solve(NMF.GreedyCD(T, maxiter=2, tol=1e-30), X, Wg, Hg, engine="simt")
    ro = oracle.solve(oracle.GreedyCD(T, maxiter=2, tol=1e-30), X, Wo, Ho)
    assert r.niters == ro.niters == 2
    # first iteration is exact arithmetic in the GEMMs; allow a few ulp after the second
    np.testing.assert_allclose(Wg, Wo, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(Hg, Ho, rtol=2e-5, atol=1e-6)
    assert r.info["coordinate_updates"] == ro.coordinate_updates


@pytest.mark.parametrize("alg", ["multmse", "multdiv", "greedycd"])
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_nnmf_update_H_false(NMF, alg, T):  # test/interf.jl:31-37
    rng = np.random.default_rng(23)
    p, n, k = 5, 8, 3
    X = np.asfortranarray(np.maximum(rng.random((p, k)) - 0.3, 0) @ np.maximum(rng.random((k, n)) - 0.3, 0), dtype=T)
    W = np.asfortranarray(np.maximum(rng.random((p, k)) - 0.3, 0), dtype=T)
    H = np.asfortranarray(np.maximum(rng.random((k, n)) - 0.3, 0), dtype=T)
    ret = NMF.nnmf(X, k, alg=alg, init="custom", W0=W.copy(order="F"), H0=H.copy(order="F"), update_H=False, engine="simt")
    assert (ret.H == H).all()
    assert (ret.W != W).any()


def test_nnmf_replicates_and_warm_start(NMF):  # test/interf.jl:24-25
    rng = np.random.default_rng(24)
    X = np.asfortranarray(rng.random((30, 20)), dtype=np.float64)
    rep = NMF.nnmf(X, 3, replicates=5, maxiter=10, alg="multmse", init="random", rng=rng)
    one = NMF.nnmf(X, 3, replicates=1, maxiter=10, alg="multmse", init="random", rng=np.random.default_rng(24))
    assert rep.niters == 10
    ret = NMF.nnmf(X, 3, W0=rep.W, H0=rep.H, init="custom", alg="multmse", maxiter=10)
    assert float(ret.objvalue) <= float(rep.objvalue) * (1 + 1e-12)  # MU is monotone
    assert np.isfinite(float(one.objvalue))


def test_c_order_inputs_are_updated_in_place(NMF, oracle):
    rng = np.random.default_rng(25)
    X = rng.random((40, 30))                     # C order
    W = rng.random((40, 4))
    H = rng.random((4, 30))
    Wo, Ho = np.asfortranarray(W), np.asfortranarray(H)
    NMF.solve(NMF.MultUpdate(np.float64, maxiter=5, tol=1e-12), X, W, H, engine="simt")
    oracle.solve(oracle.MultUpdate(np.float64, maxiter=5, tol=1e-12), np.asfortranarray(X), Wo, Ho)
    assert _relerr(W, Wo) < 1e-10 and _relerr(H, Ho) < 1e-10


def test_convergence_flag_and_niters(NMF, oracle):
    """A loose tolerance stops both sides at the same iteration (common.jl:64-83)."""
    rng = np.random.default_rng(26)
    X = np.asfortranarray(rng.random((50, 40)))
    W0, H0 = NMF.randinit(50, 40, 4, np.float64, normalize=True, rng=rng)
    Wg, Hg, Wo, Ho = W0.copy(order="F"), H0.copy(order="F"), W0.copy(order="F"), H0.copy(order="F")
    r = NMF.solve(NMF.MultUpdate(np.float64, maxiter=500, tol=1e-3), X, Wg, Hg, engine="simt")
    ro = oracle.solve(oracle.MultUpdate(np.float64, maxiter=500, tol=1e-3), X, Wo, Ho)
    assert ro.converged and r.converged and r.niters == ro.niters
    assert abs(float(r.objvalue) - float(ro.objvalue)) <= 1e-9 * float(ro.objvalue)


def test_verbose_trace(NMF, capsys):  # common.jl:54-59, 76-82; test/interf.jl:40-42
    rng = np.random.default_rng(27)
    X = np.asfortranarray(rng.random((20, 15)))
    W, H = NMF.randinit(20, 15, 3, np.float64, normalize=True, rng=rng)
    rows = []
    with NMF.Session(engine="simt") as s:
        s.set_X(X)
        s.set_trace(lambda *a: rows.append(a))
        r = s.solve(NMF.MultUpdate(np.float64, maxiter=6, tol=1e-12, verbose=True), W, H)
    assert [a[0] for a in rows] == list(range(0, 7))
    objs = [a[2] for a in rows]
    assert all(objs[i + 1] <= objs[i] * (1 + 1e-12) for i in range(6))  # MU-MSE is monotone
    assert abs(objs[-1] - float(r.objvalue)) <= 1e-12 * objs[-1]


def test_abi_errors(NMF):
    rng = np.random.default_rng(28)
    X = np.asfortranarray(rng.random((10, 8)))
    with NMF.Session() as s:
        Wd, Hd = NMF.randinit(10, 8, 2, np.float64, rng=rng)
        with pytest.raises(NMF.NmfB200Error):  # solve before set_X
            s.solve_raw("multmse", np.float64, Wd.ctypes.data, 10, Hd.ctypes.data, 2, 2, 10, 1e-3, 0, 0, True, False, False)
        Xn = X.copy(order="F")
        Xn[3, 4] = -0.5
        with pytest.raises(NMF.ArgumentError, match="non-negative"):
            s.set_X(Xn, check_nonneg=True)
        s.set_X(X, check_nonneg=True)
        W, H = NMF.randinit(10, 8, 2, np.float64, rng=rng)
        with pytest.raises(NMF.DimensionMismatch):
            s.solve(NMF.MultUpdate(np.float64), W[:9], H)
        with pytest.raises(TypeError):
            s.solve(NMF.MultUpdate(np.float32), W, H)
        with pytest.raises(NotImplementedError):
            s.solve(NMF.SPA(np.float64), W, H)
        import ctypes
        res = NMF._lib.NmfResult()
        st = s._lib.nmfb200_solve_multmse_f64(s._h, W.ctypes.data_as(ctypes.c_void_p), 10, H.ctypes.data_as(ctypes.c_void_p), 2, 2,
                                              1, 1e-3, 0.0, 0.0, 1, 0, 0, ctypes.byref(res))
        assert st == NMF._lib.EINVAL and b"maxiter" in s._lib.nmfb200_last_error(s._h)
        st = s._lib.nmfb200_solve_multmse_f64(s._h, W.ctypes.data_as(ctypes.c_void_p), 9, H.ctypes.data_as(ctypes.c_void_p), 2, 2,
                                              5, 1e-3, 0.0, 0.0, 1, 0, 0, ctypes.byref(res))
        assert st == NMF._lib.EDIM


def test_medium_sizes_simt_vs_oracle(NMF, oracle):
    """Sizes the oracle finishes in seconds; ragged (non multiple of any tile) shapes."""
    rng = np.random.default_rng(29)
    for (p, n, k, alg) in [(333, 517, 17, "multmse"), (257, 129, 9, "multdiv"), (130, 203, 12, "greedycd")]:
        T = np.float32
        X = np.asfortranarray(rng.random((p, n)), dtype=T)
        W0, H0 = NMF.randinit(p, n, k, T, normalize=True, rng=rng)
        Wg, Hg, Wo, Ho = W0.copy(order="F"), H0.copy(order="F"), W0.copy(order="F"), H0.copy(order="F")
        if alg == "greedycd":
            a, b = NMF.GreedyCD(T, maxiter=5, tol=1e-9), oracle.GreedyCD(T, maxiter=5, tol=1e-9)
        else:
            a, b = NMF.MultUpdate(T, obj=alg[4:], maxiter=20, tol=1e-9), oracle.MultUpdate(T, obj=alg[4:], maxiter=20, tol=1e-9)
        r = NMF.solve(a, X, Wg, Hg, engine="simt")
        ro = oracle.solve(b, X, Wo, Ho)
        assert r.niters == ro.niters
        assert abs(float(r.objvalue) - float(ro.objvalue)) <= 1e-4 * float(ro.objvalue)  # north-star bar
        if alg != "greedycd":
            assert _relerr(Wg, Wo) <= 5e-4 and _relerr(Hg, Ho) <= 5e-4  # fp32 summation-order noise over 20 iterations
