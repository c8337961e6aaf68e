"""Batched replicates (interf.jl:85-101): R MultUpdate(:mse) solves stacked along the component axis advance as ONE
iteration -- one pass over X per half-step for all of them (nmfb200_solve_multmse_batched_f32).  Every replicate must come out as
ITS OWN solve! would: compared with the oracle on the same initial factors (tensor-core tolerances, tests/test_gpu_tc.py), with the
single-replicate tensor-core solve (tolerance-bound stops), and through nnmf(replicates=...)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _problem(NMF, p, n, k, R, seed, planted=False):
    rng = np.random.default_rng(seed)
    if planted:
        X = np.maximum(rng.random((p, k)) - 0.3, 0) @ np.maximum(rng.random((k, n)) - 0.3, 0)
    else:
        X = rng.random((p, n))
    X = np.asfortranarray(X, dtype=np.float32)
    facs = [NMF.randinit(p, n, k, np.float32, normalize=True, rng=rng) for _ in range(R)]
    return X, facs


@pytest.mark.parametrize("p,n,k,R,iters", [
    (1024, 1152, 32, 4, 20),   # 4 x 32 = 128 stacked components: KP = 128, staged epilogue + fused tile Grams
    (1100, 1030, 16, 3, 12),   # 48 -> KP = 64, ragged tiles
    (1536, 1280, 64, 4, 10),   # 256 stacked components: KP = 256 (stand-alone Gram kernel)
    (1024, 1024, 100, 2, 10),  # 200 -> KP = 256, blocks not aligned to 32 / 64
])
def test_batched_replicates_vs_oracle(NMF, oracle, p, n, k, R, iters):
    X, facs = _problem(NMF, p, n, k, R, seed=p + n + k + R)
    Ws = [f[0].copy(order="F") for f in facs]
    Hs = [f[1].copy(order="F") for f in facs]
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=iters, tol=1e-9)
    with NMF.Session(engine="tc") as s:
        s.set_X(X)
        res = s.solve_batched(alg, Ws, Hs)
    assert len(res) == R
    for r in range(R):
        Wo, Ho = facs[r][0].copy(order="F"), facs[r][1].copy(order="F")
        ro = oracle.solve(oracle.MultUpdate(np.float32, obj="mse", maxiter=iters, tol=1e-9), X, Wo, Ho)
        assert res[r].niters == ro.niters == iters and not res[r].converged
        assert res[r].W is Ws[r] and np.isfinite(Ws[r]).all() and (Ws[r] >= 0).all() and (Hs[r] >= 0).all()
        ew, eh = _relerr(Ws[r], Wo), _relerr(Hs[r], Ho)
        eo = abs(float(res[r].objvalue) - float(ro.objvalue)) / float(ro.objvalue)
        print(f"batched p={p} n={n} k={k} R={R} rep={r}: errW={ew:.2e} errH={eh:.2e} errObj={eo:.2e}")
        assert ew <= 5e-3 and eh <= 5e-3   # the bf16 tolerance of the single solve (tests/test_gpu_tc.py)
        assert eo <= 1e-4


def test_batched_replicates_do_not_interact(NMF):
    """The same initial factors in another slot of the stack, next to other neighbours, give the same solve."""
    X, facs = _problem(NMF, 1024, 1024, 32, 4, seed=5)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=15, tol=1e-9)
    with NMF.Session(engine="tc") as s:
        s.set_X(X)
        a = s.solve_batched(alg, [f[0].copy(order="F") for f in facs], [f[1].copy(order="F") for f in facs])
        order = [2, 0, 3, 1]
        b = s.solve_batched(alg, [facs[i][0].copy(order="F") for i in order], [facs[i][1].copy(order="F") for i in order])
        one = s.solve(alg, facs[2][0].copy(order="F"), facs[2][1].copy(order="F"))
    for slot, i in enumerate(order):
        assert _relerr(b[slot].W, a[i].W) <= 2e-5 and _relerr(b[slot].H, a[i].H) <= 2e-5
        assert abs(float(b[slot].objvalue) - float(a[i].objvalue)) <= 1e-6 * float(a[i].objvalue)
    # and the stacked solve agrees with the single tensor-core solve of the same factors (different KP, same arithmetic)
    assert _relerr(a[2].W, one.W) <= 1e-4 and _relerr(a[2].H, one.H) <= 1e-4
    assert abs(float(a[2].objvalue) - float(one.objvalue)) <= 1e-5 * float(one.objvalue)


def test_batched_stop_condition_is_per_replicate(NMF):
    """Tolerance-bound: every replicate stops at the iteration its own solve! stops at and keeps the factors of that iteration,
    while the stacked iteration carries on for the others."""
    p, n, k, R = 1024, 1024, 8, 6
    X, facs = _problem(NMF, p, n, k, R, seed=11, planted=True)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=400, tol=2e-3)
    with NMF.Session(engine="tc") as s:
        s.set_X(X)
        singles = [s.solve(alg, f[0].copy(order="F"), f[1].copy(order="F")) for f in facs]
        batch = s.solve_batched(alg, [f[0].copy(order="F") for f in facs], [f[1].copy(order="F") for f in facs])
        s.set_option("check_every", 3)   # the host's polling interval must not matter
        batch3 = s.solve_batched(alg, [f[0].copy(order="F") for f in facs], [f[1].copy(order="F") for f in facs])
    its = [r.niters for r in singles]
    print("single niters", its, "batched niters", [r.niters for r in batch])
    assert len(set(its)) > 1, "the test problem should make the replicates stop at different iterations"
    for r in range(R):
        assert batch[r].converged == singles[r].converged
        assert abs(batch[r].niters - singles[r].niters) <= max(3, 0.05 * singles[r].niters)
        assert _relerr(batch[r].W @ batch[r].H, singles[r].W @ singles[r].H) <= 2e-3
        assert batch3[r].niters == batch[r].niters and batch3[r].converged == batch[r].converged
        assert (batch3[r].W == batch[r].W).all() and (batch3[r].H == batch[r].H).all()


def test_batched_update_H_false_and_maxiter_bound(NMF):
    X, facs = _problem(NMF, 1024, 1024, 16, 4, seed=13)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=6, tol=1e-9, update_H=False)
    Ws = [f[0].copy(order="F") for f in facs]
    Hs = [f[1].copy(order="F") for f in facs]
    with NMF.Session(engine="tc") as s:
        s.set_X(X)
        res = s.solve_batched(alg, Ws, Hs)
    for r in range(4):
        assert res[r].niters == 6 and not res[r].converged
        assert (Hs[r] == facs[r][1]).all()      # test/interf.jl:35
        assert (Ws[r] != facs[r][0]).any()


def test_nnmf_replicates_batched_equals_one_by_one(NMF):
    rng = np.random.default_rng(17)
    X = np.asfortranarray(rng.random((1024, 1100)), dtype=np.float32)
    kw = dict(alg="multmse", init="random", maxiter=12, tol=1e-9, engine="tc")
    a = NMF.nnmf(X, 24, replicates=5, rng=np.random.default_rng(3), **kw)
    assert a.info.get("batched") == 5           # 5 x 24 = 120 stacked components, one group
    with NMF.Session(engine="tc") as s:
        s.set_X(X)
        r2 = np.random.default_rng(3)
        W, H = NMF.randinit(1024, 1100, 24, np.float32, normalize=True, rng=r2)
        b = NMF.solve_replicates(NMF.MultUpdate(np.float32, obj="mse", maxiter=12, tol=1e-9), s, W, H, replicates=5, initH=True,
                                 rng=r2, batched=False)
    assert "batched" not in b.info
    assert abs(float(a.objvalue) - float(b.objvalue)) <= 1e-5 * float(b.objvalue)
    assert _relerr(a.W, b.W) <= 1e-3 and _relerr(a.H, b.H) <= 1e-3   # the same replicate won
    # more replicates than fit one stack: groups of 256 // k
    c = NMF.nnmf(X, 100, replicates=5, rng=np.random.default_rng(4), **kw)
    assert c.info.get("batched") in (1, 2) or "batched" not in c.info
    assert c.niters == 12


def test_batched_request_outside_coverage_falls_back(NMF):
    rng = np.random.default_rng(19)
    X64 = np.asfortranarray(rng.random((40, 30)))
    with NMF.Session() as s:
        s.set_X(X64)
        W, H = NMF.randinit(40, 30, 3, np.float64, normalize=True, rng=rng)
        with pytest.raises(NotImplementedError):
            s.solve_batched(NMF.MultUpdate(np.float64, maxiter=5), [W, W.copy()], [H, H.copy()])
    r = NMF.nnmf(X64, 3, replicates=3, alg="multmse", init="random", maxiter=10, rng=rng)   # Float64: one by one on the exact engine
    assert r.niters == 10 and r.info["engine"] == "simt"
    X32 = np.asfortranarray(rng.random((40, 30)), dtype=np.float32)                          # Float32 but tiny: ENOTSUP -> one by one
    r = NMF.nnmf(X32, 3, replicates=3, alg="multmse", init="random", maxiter=10, rng=rng)
    assert r.niters == 10 and r.info["engine"] == "simt"
