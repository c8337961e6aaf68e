"""An INDEPENDENT implementation cross-checks the oracle where one exists in this image: NMF.jl's CoordinateDescent
(coorddesc.jl:1-3 says it is a translation of scikit-learn's coordinate-descent NMF solver) against scikit-learn itself.
 (1) one sweep: oracle_cd_sweep (oracle/oracle_kernels.c, restating coorddesc.jl:138-157) vs
     sklearn.decomposition._cdnmf_fast._update_cdnmf_fast on the same (W, HHt, XHt, permutation), incl. the violation sum;
 (2) whole solves: oracle.solve(CoordinateDescent) vs sklearn's _fit_coordinate_descent for a fixed number of
     iterations (W first, then H -- coorddesc.jl:166-175 and sklearn's loop agree), with and without elastic-net terms;
 (3) MultUpdate(:mse) vs sklearn's multiplicative-update solver is NOT comparable step by step (sklearn's MU divides by
     the denominator floored at EPSILON and updates W before H), so it is not used.
This narrows -- it does not remove -- the "parity unpinned" caveat of DESIGN.md section 6: Julia itself cannot run here."""
import numpy as np
import pytest

sk_fast = pytest.importorskip("sklearn.decomposition._cdnmf_fast")
sk_nmf = pytest.importorskip("sklearn.decomposition._nmf")


@pytest.mark.parametrize("seed,rows,k,shuffle", [(0, 40, 5, False), (1, 77, 9, True), (2, 128, 16, True), (3, 6, 3, False)])
def test_cd_sweep_matches_sklearn_update_cdnmf_fast(oracle, seed, rows, k, shuffle):
    rng = np.random.default_rng(seed)
    cols = 50
    Ht = rng.random((cols, k))
    X = rng.random((rows, cols))
    W = rng.random((rows, k))
    W[rng.random((rows, k)) < 0.2] = 0.0                      # zeros exercise the projected-gradient branch (:146-149)
    HHt = Ht.T @ Ht
    HHt[np.diag_indices(k)] += 0.01                           # l2
    XHt = X @ Ht - 0.02                                       # l1
    perm = rng.permutation(k) if shuffle else np.arange(k)
    W_sk = np.ascontiguousarray(W.copy())
    v_sk = sk_fast._update_cdnmf_fast(W_sk, np.ascontiguousarray(HHt), np.ascontiguousarray(XHt), perm.astype(np.intp))
    W_or, HHt_f, XHt_f, perm64 = np.asfortranarray(W.copy()), np.asfortranarray(HHt), np.asfortranarray(XHt), perm.astype(np.int64)
    v_or = oracle._fn("oracle_cd_sweep", np.dtype(np.float64))(oracle._p(W_or), oracle._p(HHt_f), oracle._p(XHt_f), rows, k, oracle._p(perm64))
    np.testing.assert_allclose(W_or, W_sk, rtol=1e-13, atol=1e-15)
    assert abs(v_or - v_sk) <= 1e-11 * max(1.0, abs(v_sk))


@pytest.mark.parametrize("p,n,k,iters,alpha,l1ratio", [(30, 24, 4, 1, 0.0, 0.0), (60, 45, 6, 7, 0.0, 0.0), (50, 70, 5, 5, 0.05, 0.5),
                                                       (41, 33, 8, 4, 0.1, 1.0)])
def test_cd_solve_matches_sklearn_fit_coordinate_descent(oracle, p, n, k, iters, alpha, l1ratio):
    rng = np.random.default_rng(p + n)
    X = rng.random((p, n))
    W0, H0 = rng.random((p, k)), rng.random((k, n))
    Wo, Ho = np.asfortranarray(W0.copy()), np.asfortranarray(H0.copy())
    oracle.solve(oracle.CoordinateDescent(np.float64, maxiter=iters, tol=1e-300, alpha=alpha, l1ratio=l1ratio), X, Wo, Ho)
    # sklearn: same elastic-net weights on both factors (regularization = :both), tol = 0 so that exactly `iters` sweeps run
    l1, l2 = alpha * l1ratio, alpha * (1.0 - l1ratio)
    Wsk, Hsk, nit = sk_nmf._fit_coordinate_descent(X, np.ascontiguousarray(W0.copy()), np.ascontiguousarray(H0.copy()), tol=0.0,
                                                  max_iter=iters, l1_reg_W=l1, l1_reg_H=l1, l2_reg_W=l2, l2_reg_H=l2,
                                                  update_H=True, shuffle=False)
    assert nit == iters
    np.testing.assert_allclose(Wo, Wsk, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(Ho, Hsk, rtol=1e-10, atol=1e-12)


def test_cd_update_H_false_matches_sklearn(oracle):
    rng = np.random.default_rng(5)
    X = rng.random((36, 28))
    W0, H0 = rng.random((36, 5)), rng.random((5, 28))
    Wo, Ho = np.asfortranarray(W0.copy()), np.asfortranarray(H0.copy())
    oracle.solve(oracle.CoordinateDescent(np.float64, maxiter=6, tol=1e-300, update_H=False), X, Wo, Ho)
    Wsk, Hsk, _ = sk_nmf._fit_coordinate_descent(X, np.ascontiguousarray(W0.copy()), np.ascontiguousarray(H0.copy()), tol=0.0, max_iter=6,
                                                 update_H=False, shuffle=False)
    np.testing.assert_allclose(Wo, Wsk, rtol=1e-10, atol=1e-12)
    assert (Ho == H0).all()


# ---- MultUpdate half-steps against scikit-learn's multiplicative-update helpers ----------------------------------------
# One oracle iteration with update_H=True is H-step then W-step (multupd.jl:95-115 / :171-192).  scikit-learn's helpers apply
# ONE half-step in place and return the updated factor.  The formulas agree except for the guard: NMF.jl
# adds delta = sqrt(eps(T)) to the denominator (:mse) / to WH (:div), scikit-learn floors at EPSILON.  In Float64 that is a
# relative difference of ~1.5e-8 per element, which bounds the tolerance below (1e-6): an independent check of the update
# algebra, not a bit-level pin.
@pytest.mark.parametrize("obj,beta", [("mse", 2), ("div", 1)])
def test_multupdate_half_steps_match_sklearn_helpers(oracle, obj, beta):
    rng = np.random.default_rng(17)
    p, n, k = 48, 37, 5
    X = np.asfortranarray(rng.random((p, n)) + 0.05)       # the oracle's element loops index memory linearly: Julia layout
    W0, H0 = rng.random((p, k)) + 0.05, rng.random((k, n)) + 0.05
    alg = oracle.MultUpdate(np.float64, obj=obj, maxiter=2, tol=1e-300)
    lam_h, lam_w = float(alg.lambda_h), float(alg.lambda_w)          # :div floors them at sqrt(eps) (multupd.jl:37-40)
    # oracle: a single update_wh! call
    Wo, Ho = np.asfortranarray(W0.copy()), np.asfortranarray(H0.copy())
    delta = np.float64(np.sqrt(np.finfo(np.float64).eps))
    upd = (oracle.MultUpdMSE if obj == "mse" else oracle.MultUpdDiv)(np.float64, True, alg.lambda_w, alg.lambda_h, delta)
    st = upd.prepare_state(X, Wo, Ho)
    upd.update_wh(st, X, Wo, Ho)
    # scikit-learn: H-step with the old W, then W-step with the new H.  Its l1 term is ADDED to the denominator, which is
    # where :div puts lambda (multupd.jl:178,190); for :mse NMF.jl subtracts lambda from the numerator instead, so use 0.
    l1h, l1w = (lam_h, lam_w) if obj == "div" else (0.0, 0.0)
    Hs = sk_nmf._multiplicative_update_h(X, W0.copy(), H0.copy(), beta, l1h, 0.0, 1.0)         # returns the updated H
    Ws = sk_nmf._multiplicative_update_w(X, W0.copy(), Hs.copy(), beta, l1w, 0.0, 1.0)[0]      # ... and the updated W
    np.testing.assert_allclose(Ho, Hs, rtol=1e-6)
    np.testing.assert_allclose(Wo, Ws, rtol=1e-6)
