"""ProjectedALS / CoordinateDescent / ALSPGrad through the C ABI against the oracle:
the reference's own known-answer tests (test/coorddesc.jl, test/alspgrad.jl, test/interf.jl) run on the GPU, plus
parity with the oracle on seeded problems.  Tolerances are stated at each assertion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _problem(NMF, p, n, k, T, seed, planted=False, zeroh=False):
    rng = np.random.default_rng(seed)
    if planted:
        X = np.maximum(rng.random((p, k)) - 0.3, 0) @ np.maximum(rng.random((k, n)) - 0.3, 0)
    else:
        X = rng.random((p, n))
    X = np.asfortranarray(X, dtype=T)
    W0, H0 = NMF.randinit(p, n, k, T, normalize=True, zeroh=zeroh, rng=rng)
    return X, W0, H0


def _start(oracle, T, rng):
    X, Wg, Hg = oracle.laurberg6x3(0.3, T)
    return X, np.asfortranarray(Wg + rng.random(Wg.shape).astype(T) * T(0.1)), Hg.copy(order="F")


# ---- test/coorddesc.jl:4-16 on the GPU
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_reference_kat_coorddesc_gpu(NMF, oracle, T):
    X, W, H = _start(oracle, T, np.random.default_rng(31))
    r = NMF.solve(NMF.CoordinateDescent(T, alpha=0.0, maxiter=1000, tol=1e-9), X, W, H)
    assert r.info["engine"] == "simt" and r.info["kernel_launches"] > 0
    assert np.linalg.norm(X - W @ H) <= 1e-4
    X, W, H = _start(oracle, T, np.random.default_rng(32))
    NMF.solve(NMF.CoordinateDescent(T, alpha=1e-4, l1ratio=0.5, shuffle=True, maxiter=1000, tol=1e-9, seed=5), X, W, H)
    assert np.linalg.norm(X - W @ H) <= 1e-2


# ---- test/alspgrad.jl:4-27: the reference only asserts on the sub-solvers and that solve! runs; on the GPU the
# sub-solvers are reached through solve! (update_H=false isolates the W sub-solve)
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_reference_kat_alspgrad_gpu(NMF, oracle, T):
    rng = np.random.default_rng(33)
    X, Wg, Hg = oracle.laurberg6x3(0.3, T)
    W = np.asfortranarray(rng.random(Wg.shape).astype(T))
    H = Hg.copy(order="F")
    eps = np.finfo(T).eps
    r = NMF.solve(NMF.ALSPGrad(T, maxiter=2, maxsubiter=1000, tol=1e-30, tolg=eps, update_H=False), X, W, H)
    assert (H == Hg).all() and (W >= 0).all()
    assert np.linalg.norm(W - Wg) <= eps ** 0.25           # `W ≈ Wg atol=eps(T)^(1/4)`
    assert r.info["sub_iterations"] >= 2
    W = np.asfortranarray(rng.random(Wg.shape).astype(T))
    H = np.asfortranarray(rng.random(Hg.shape).astype(T))
    r = NMF.solve(NMF.ALSPGrad(T), X, W, H)
    assert np.isfinite(float(r.objvalue)) and (W >= 0).all() and (H >= 0).all()


# ---- test/interf.jl:31-37 for the three algorithm types
@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("alg", ["projals", "alspgrad", "cd"])
def test_update_H_false_bit_identical(NMF, T, alg):
    rng = np.random.default_rng(34)
    p, n, k = 5, 8, 3
    X = np.asfortranarray(np.maximum(rng.random((p, k)) - 0.3, 0) @ np.maximum(rng.random((k, n)) - 0.3, 0), dtype=T)
    W = np.asfortranarray(np.maximum(rng.random((p, k)) - 0.3, 0), dtype=T)
    H = np.asfortranarray(np.maximum(rng.random((k, n)) - 0.3, 0) + 0.05, dtype=T)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ret = NMF.nnmf(X, k, alg=alg, init="custom", W0=W.copy(order="F"), H0=H.copy(order="F"), update_H=False)
    assert (ret.H == H).all() and (ret.W != W).any()


# ---- parity with the oracle on seeded problems
@pytest.mark.parametrize("T,p,n,k,iters,planted", [
    (np.float64, 96, 80, 6, 12, False),
    (np.float64, 130, 75, 10, 8, True),
    (np.float32, 200, 160, 8, 10, False),
    (np.float32, 257, 129, 16, 6, True),     # ragged sizes
])
def test_projals_vs_oracle(NMF, oracle, T, p, n, k, iters, planted):
    X, W0, H0 = _problem(NMF, p, n, k, T, seed=p + k, planted=planted, zeroh=True)
    Wg, Hg, Wo, Ho = W0.copy(order="F"), H0.copy(order="F"), W0.copy(order="F"), H0.copy(order="F")
    kw = dict(maxiter=iters, tol=1e-12)
    r = NMF.solve(NMF.ProjectedALS(T, **kw), X, Wg, Hg)
    ro = oracle.solve(oracle.ProjectedALS(T, **kw), X, Wo, Ho)
    assert r.niters == ro.niters and r.converged == ro.converged
    # the oracle solves the k x k systems with LAPACK in T, the GPU inverts in Float64: differences are bounded by
    # cond(Gram + lambda I) * eps(T)
    tol = 1e-8 if T == np.float64 else 5e-3
    ew, eh = _relerr(Wg, Wo), _relerr(Hg, Ho)
    eo = abs(float(r.objvalue) - float(ro.objvalue)) / float(ro.objvalue)
    print(f"projals {T.__name__} p={p} n={n} k={k}: errW={ew:.2e} errH={eh:.2e} errObj={eo:.2e}")
    assert (Wg >= 0).all() and (Hg >= 0).all()
    if T == np.float32 and planted:
        # rank-k data: cond(HH' + lambda I) ~ 1e6, so the reference's own Float32 path is only defined up to
        # cond * eps(Float32) ~ 1e-1 in W.  Bar: the GPU result is as close to the Float64 solution of the same
        # problem as the Float32 oracle is (factor 3), and the objective agrees to 5e-3.
        W64, H64 = W0.astype(np.float64, order="F"), H0.astype(np.float64, order="F")
        lam = float(np.float32(np.cbrt(np.finfo(np.float32).eps)))
        oracle.solve(oracle.ProjectedALS(np.float64, lambda_w=lam, lambda_h=lam, **kw), X.astype(np.float64, order="F"), W64, H64)
        print(f"   vs Float64 solution: gpu {_relerr(Wg, W64):.2e}/{_relerr(Hg, H64):.2e}  oracle32 {_relerr(Wo, W64):.2e}/{_relerr(Ho, H64):.2e}")
        assert _relerr(Wg, W64) <= 3 * _relerr(Wo, W64) + 1e-3 and _relerr(Hg, H64) <= 3 * _relerr(Ho, H64) + 1e-3
        assert eo <= 5e-3
    else:
        assert ew <= tol and eh <= tol and eo <= tol


def test_projals_not_positive_definite_is_a_numerical_error(NMF):
    """lambda = 0 and a zero column in W: W'W is singular.  The reference ignores potrf!'s info (utils.jl:68) and carries on
    with an unfinished factor; the library stops with NMFB200_ENUMERIC, distinct from ArgumentError."""
    rng = np.random.default_rng(35)
    X = np.asfortranarray(rng.random((20, 16)))
    W = np.asfortranarray(rng.random((20, 3)))
    W[:, 1] = 0.0
    H = np.zeros((3, 16), order="F")
    with pytest.raises(NMF.NumericalError, match="positive definite"):
        NMF.solve(NMF.ProjectedALS(np.float64, maxiter=3, lambda_w=0.0, lambda_h=0.0), X, W, H)


@pytest.mark.parametrize("T,p,n,k,iters,opts", [
    (np.float64, 90, 70, 6, 10, {}),
    (np.float64, 64, 100, 9, 8, dict(alpha=1e-3, l1ratio=0.5)),
    (np.float64, 75, 60, 5, 8, dict(alpha=1e-2, l1ratio=0.3, regularization="transformation", shuffle=True, seed=11)),
    (np.float32, 200, 150, 8, 10, {}),
    (np.float32, 130, 257, 12, 6, dict(alpha=1e-3, l1ratio=1.0, shuffle=True, seed=3)),
    (np.float32, 96, 64, 40, 5, dict(regularization="none", alpha=1.0)),
])
def test_cd_vs_oracle(NMF, oracle, T, p, n, k, iters, opts):
    X, W0, H0 = _problem(NMF, p, n, k, T, seed=2 * p + k)
    Wg, Hg, Wo, Ho = W0.copy(order="F"), H0.copy(order="F"), W0.copy(order="F"), H0.copy(order="F")
    kw = dict(maxiter=iters, tol=1e-12, **opts)
    r = NMF.solve(NMF.CoordinateDescent(T, **kw), X, Wg, Hg)
    ro = oracle.solve(oracle.CoordinateDescent(T, **kw), X, Wo, Ho)
    assert r.niters == ro.niters == iters
    # the sweep itself is the reference's sequential arithmetic; only the GEMM summation order differs
    tol = 1e-9 if T == np.float64 else 1e-3
    ew, eh = _relerr(Wg, Wo), _relerr(Hg, Ho)
    eo = abs(float(r.objvalue) - float(ro.objvalue)) / float(ro.objvalue)
    print(f"cd {T.__name__} p={p} n={n} k={k} {opts}: errW={ew:.2e} errH={eh:.2e} errObj={eo:.2e}")
    assert ew <= tol and eh <= tol and eo <= tol
    assert (Wg >= 0).all() and (Hg >= 0).all()


@pytest.mark.parametrize("T,p,n,k,iters,planted", [
    (np.float64, 80, 64, 5, 6, False),
    (np.float64, 60, 90, 8, 5, True),
    (np.float32, 150, 120, 6, 6, False),
])
def test_alspgrad_vs_oracle(NMF, oracle, T, p, n, k, iters, planted):
    X, W0, H0 = _problem(NMF, p, n, k, T, seed=3 * p + k, planted=planted)
    Wg, Hg, Wo, Ho = W0.copy(order="F"), H0.copy(order="F"), W0.copy(order="F"), H0.copy(order="F")
    kw = dict(maxiter=iters, tol=1e-12, maxsubiter=40)
    r = NMF.solve(NMF.ALSPGrad(T, **kw), X, Wg, Hg)
    ro = oracle.solve(oracle.ALSPGrad(T, **kw), X, Wo, Ho)
    assert r.niters == ro.niters
    eo = abs(float(r.objvalue) - float(ro.objvalue)) / float(ro.objvalue)
    ew, eh = _relerr(Wg, Wo), _relerr(Hg, Ho)
    print(f"alspgrad {T.__name__} p={p} n={n} k={k}: errW={ew:.2e} errH={eh:.2e} errObj={eo:.2e} "
          f"sub={r.info['sub_iterations']}/{ro.subiters} tolg={r.info['tolg_final']:.3g}/{float(ro.tolg_final):.3g}")
    # Float64: same Armijo decisions => same trajectory up to summation order.  Float32: a decision may flip near the
    # sufficient-decrease boundary; the objective stays within 5e-3.
    if T == np.float64:
        assert r.info["sub_iterations"] == ro.subiters
        assert ew <= 1e-7 and eh <= 1e-7 and eo <= 1e-9
    else:
        assert eo <= 5e-3
    assert (Wg >= 0).all() and (Hg >= 0).all()


def test_nnmf_dispatches_next_algs(NMF, oracle):
    """interf.jl:60-69: nnmf(alg=:projals|:alspgrad|:cd) with init=:random; projals starts from H = 0 (interf.jl:39)."""
    rng = np.random.default_rng(36)
    X = np.asfortranarray(np.maximum(rng.random((12, 3)) - 0.3, 0) @ np.maximum(rng.random((3, 10)) - 0.3, 0))
    for alg in ("projals", "alspgrad", "cd"):
        r = NMF.nnmf(X, 3, alg=alg, init="random", maxiter=30, rng=np.random.default_rng(1))
        assert r.W.shape == (12, 3) and r.H.shape == (3, 10) and np.isfinite(float(r.objvalue))
        assert float(r.objvalue) <= 0.5 * float(np.sum(X * X))


# ---- the X-sized products of ProjectedALS / CoordinateDescent / ALSPGrad on the tcgen05 mainloop --------------------
# Float32 problems from 2^20 cells on compute W'X and XH' with the update kernel of the tensor-core engine (split bf16 hi + lo
# operands, fp32 accumulation; tc_xmul in csrc/tc_engine.cu); the k x k algebra (Cholesky-equivalent inverse, sweeps, Armijo
# control flow) is unchanged.  Checked against the oracle AND against the same solve with the products on the exact engine
# (option tc_xmul=0): the tensor-core products must not cost more than a small multiple of the exact engine's own distance.
def _solve_opt(NMF, alg, X, W0, H0, opts=()):
    W, H = W0.copy(order="F"), H0.copy(order="F")
    with NMF.Session(engine="auto") as s:
        for key, val in opts:
            s.set_option(key, val)
        s.set_X(X)
        r = s.solve(alg, W, H)
    return r, W, H


@pytest.mark.parametrize("name,p,n,k,iters", [("projals", 1536, 1024, 32, 8), ("projals", 1100, 1300, 100, 5), ("cd", 1280, 1024, 24, 6),
                                              ("cd", 1024, 1536, 150, 3), ("alspgrad", 1200, 1024, 16, 3), ("projals", 6144, 4096, 64, 4),
                                              ("cd", 4096, 6144, 64, 3)])
def test_next_algorithms_with_tensor_core_products(NMF, oracle, name, p, n, k, iters):
    X, W0, H0 = _problem(NMF, p, n, k, np.float32, seed=p + k, zeroh=(name == "projals"))
    if name == "projals":
        alg, oalg = NMF.ProjectedALS(np.float32, maxiter=iters, tol=1e-30), oracle.ProjectedALS(np.float32, maxiter=iters, tol=1e-30)
    elif name == "cd":
        alg = NMF.CoordinateDescent(np.float32, maxiter=iters, tol=1e-30, alpha=1e-3, l1ratio=0.5)
        oalg = oracle.CoordinateDescent(np.float32, maxiter=iters, tol=1e-30, alpha=1e-3, l1ratio=0.5)
    else:
        alg, oalg = NMF.ALSPGrad(np.float32, maxiter=iters, tol=1e-30), oracle.ALSPGrad(np.float32, maxiter=iters, tol=1e-30)
    r, W, H = _solve_opt(NMF, alg, X, W0, H0)
    re, We, He = _solve_opt(NMF, alg, X, W0, H0, opts=(("tc_xmul", 0),))
    Wo, Ho = W0.copy(order="F"), H0.copy(order="F")
    ro = oracle.solve(oalg, X, Wo, Ho)
    assert r.info["engine"] == "tc" and re.info["engine"] == "simt"          # "tc" here = tensor-core products, exact rest
    assert r.niters == ro.niters == iters
    e_tc = max(_relerr(W, Wo), _relerr(H, Ho))
    e_ex = max(_relerr(We, Wo), _relerr(He, Ho))
    eo = abs(float(r.objvalue) - float(ro.objvalue)) / float(ro.objvalue)
    print(f"{name} p={p} n={n} k={k} it={iters}: W/H vs oracle: tensor-core products {e_tc:.1e}, exact products {e_ex:.1e}; objvalue {eo:.1e}; "
          f"loop {r.info['solve_ms']:.2f} ms vs {re.info['solve_ms']:.2f} ms")
    assert (W >= 0).all() and (H >= 0).all() and np.isfinite(W).all() and np.isfinite(H).all()
    assert e_tc <= 10 * e_ex + 1e-4
    assert eo <= 1e-4
