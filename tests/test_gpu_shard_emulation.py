"""The row-sharded tensor-core solver (csrc/tc_shard.cuh) with G LOGICAL ranks on ONE GPU (option "emulate_shards"):
the same kernels, arenas, flags and epochs a G-GPU run uses -- partial numerators into the owner's slots, slot sums in
rank order, owner-only ratio, all-gather by the epilogue's stores, partial-Gram exchange, stop decision riding one
iteration late -- so the driver's single-GPU `pytest -m gpu` exercises the sharded mathematics.
Checked against the oracle (same bars as the single-GPU tensor-core tests: objvalue 1e-4, W/H 5e-3 relative Frobenius),
against the unsharded tensor-core solve (summation order is the only difference: 1e-4 after ~10 iterations), and for the invariants of the
replicated state (tc_debug bit 5 makes the library compare H bit-for-bit across the logical ranks)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _problem(NMF, p, n, k, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.random((p, n)), dtype=np.float32)
    W0, H0 = NMF.randinit(p, n, k, np.float32, normalize=True, rng=rng)
    return X, W0, H0


def _solve(NMF, X, W0, H0, G, alg, check_every=7, opts=()):
    W, H = W0.copy(order="F"), H0.copy(order="F")
    with NMF.Session(engine="tc") as s:
        s.set_option("check_every", check_every)
        s.set_option("emulate_shards", G)
        s.set_option("tc_debug", 32)
        for k_, v_ in opts:
            s.set_option(k_, v_)
        s.set_X(X)
        r = s.solve(alg, W, H)
    return r, W, H


@pytest.mark.parametrize("p,n,k,iters,G", [
    (1024, 768, 96, 12, 2),     # KP = 128, 6 H tiles over 2 owners (two ranks: the fused H-step is the default)
    (1024, 768, 96, 12, -2),    # negative G: the other H-step form than the default (here: K1 + Ksum + K3 for two ranks)
    (1024, 768, 96, 12, 4),     # ... over 4 owners: 2 own two tiles, 2 own one; K3 runs 64-row tiles
    (1000, 1100, 64, 10, 3),    # ragged rows (1000 = 334 + 333 + 333: shards 2, 3 start at rows not divisible by 4), ragged H tail
    (2048, 1024, 128, 8, 8),    # 8 H tiles, 8 owners
    (515, 640, 32, 9, 4),       # the shape the round-1 multi-GPU check used (p % G != 0)
    (640, 300, 200, 6, 2),      # KP = 256: no staged epilogue, slab pushed by the copy kernel, Gram by gram_kernel
    (700, 200, 24, 8, 8),       # fewer H tiles (2) than ranks: six ranks own nothing
    (1536, 4224, 64, 4, 4),     # 33 H tiles, KP = 64
    (1024, 768, 96, 12, -4),    # ... and the fused H-step (MODE 6: own tiles finish inside the numerator kernel) for more than two
    (2048, 1024, 128, 8, -8),
])
def test_emulated_shards_vs_oracle_and_unsharded(NMF, oracle, p, n, k, iters, G):
    opts = ()
    if G < 0:
        G, opts = -G, (("tc_fused_hstep", 0 if G == -2 else 1),)
    X, W0, H0 = _problem(NMF, p, n, k, seed=p + n + k + G)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=iters, tol=1e-9)
    r, W, H = _solve(NMF, X, W0, H0, G, alg, opts=opts)
    r1, W1, H1 = _solve(NMF, X, W0, H0, 0, alg)
    Wo, Ho = W0.copy(order="F"), H0.copy(order="F")
    ro = oracle.solve(oracle.MultUpdate(np.float32, obj="mse", maxiter=iters, tol=1e-9), X, Wo, Ho)
    assert r.info["engine"] == "tc" and r.niters == ro.niters == iters and not r.converged
    assert np.isfinite(W).all() and np.isfinite(H).all() and (W >= 0).all() and (H >= 0).all()
    ew, eh = _relerr(W, Wo), _relerr(H, Ho)
    eo = abs(float(r.objvalue) - float(ro.objvalue)) / float(ro.objvalue)
    print(f"G={G} p={p} n={n} k={k}: vs oracle errW={ew:.2e} errH={eh:.2e} errObj={eo:.2e}; vs unsharded {_relerr(W, W1):.2e} {_relerr(H, H1):.2e}")
    assert ew <= 5e-3 and eh <= 5e-3 and eo <= 1e-4
    assert _relerr(W, W1) <= 1e-4 and _relerr(H, H1) <= 1e-4
    assert abs(float(r.objvalue) - float(r1.objvalue)) <= 2e-6 * float(r1.objvalue)


@pytest.mark.parametrize("G", [2, 4])
def test_emulated_shards_tolerance_bound_stop_matches_unsharded(NMF, oracle, G):
    """stop_condition rides one iteration late in the sharded loop (its W-side sums travel with the next exchange); the
    iteration it stops at, `converged` and the factors must not depend on that, nor on how often the host polls."""
    X, W0, H0 = _problem(NMF, 515, 640, 32, seed=77)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=400, tol=2e-3)
    r1, W1, H1 = _solve(NMF, X, W0, H0, 0, alg)
    outs = [_solve(NMF, X, W0, H0, G, alg, check_every=ce) for ce in (1, 7, 64)]
    ro = oracle.solve(oracle.MultUpdate(np.float32, obj="mse", maxiter=400, tol=2e-3), X, W0.copy(order="F"), H0.copy(order="F"))
    assert ro.converged and r1.converged
    for r, W, H in outs:
        assert r.converged and abs(r.niters - ro.niters) <= max(3, ro.niters // 20)
        assert r.niters == outs[0][0].niters and (W == outs[0][1]).all() and (H == outs[0][2]).all()   # independent of check_every
        assert abs(r.niters - r1.niters) <= max(3, r1.niters // 20)     # summation order moves a tolerance-bound stop by a step or two


def test_emulated_shards_update_H_false_and_session_reuse(NMF, oracle):
    X, W0, H0 = _problem(NMF, 900, 512, 48, seed=5)
    alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=6, tol=1e-9, update_H=False)
    r, W, H = _solve(NMF, X, W0, H0, 3, alg)
    Wo, Ho = W0.copy(order="F"), H0.copy(order="F")
    oracle.solve(oracle.MultUpdate(np.float32, obj="mse", maxiter=6, tol=1e-9, update_H=False), X, Wo, Ho)
    assert (H == H0).all() and _relerr(W, Wo) <= 5e-3                 # test/interf.jl:31-37
    # the arenas, flags and epochs survive across solves on one handle; results are repeatable bit for bit
    with NMF.Session(engine="tc") as s:
        s.set_option("emulate_shards", 4)
        s.set_X(X)
        outs = []
        for _ in range(3):
            Wg, Hg = W0.copy(order="F"), H0.copy(order="F")
            rr = s.solve(NMF.MultUpdate(np.float32, maxiter=7, tol=1e-9), Wg, Hg)
            outs.append((Wg, Hg, float(rr.objvalue)))
        assert all((o[0] == outs[0][0]).all() and (o[1] == outs[0][1]).all() and o[2] == outs[0][2] for o in outs)


def test_emulated_shards_verbose_trace(NMF, oracle):
    X, W0, H0 = _problem(NMF, 1024, 512, 32, seed=9)
    rows = []
    W, H = W0.copy(order="F"), H0.copy(order="F")
    with NMF.Session(engine="tc") as s:
        s.set_option("emulate_shards", 2)
        s.set_X(X)
        s.set_trace(lambda it, el, ob, ch, dv: rows.append((it, ob)))
        r = s.solve(NMF.MultUpdate(np.float32, maxiter=5, tol=1e-9, verbose=True), W, H)
    assert [t for t, _ in rows] == [0, 1, 2, 3, 4, 5]
    objs = [o for _, o in rows]
    assert all(objs[i + 1] <= objs[i] * (1 + 1e-6) for i in range(5)) and abs(objs[-1] - float(r.objvalue)) <= 1e-6 * objs[-1]
