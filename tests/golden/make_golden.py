"""Generates the committed golden fixtures tests/golden/*.npz.

The reference (Julia) cannot run in the build image and holds no golden vectors of its own, so these
are produced by the oracle restatement (oracle/nmf_oracle.py) on seeded inputs.  Each file stores the
inputs (X, W0, H0, options) and the oracle outputs (W, H, niters, converged, objvalue).
Run from the repo root:  python tests/golden/make_golden.py [name-substring ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nmf_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ONLY = sys.argv[1:]  # optional name filters: regenerate only the matching cases


def planted(rng, p, n, k, T):
    # test/interf.jl:7-9 recipe
    Wg = np.maximum(rng.random((p, k)) - 0.3, 0)
    Hg = np.maximum(rng.random((k, n)) - 0.3, 0)
    return np.asfortranarray(Wg @ Hg, dtype=T)


def case2(name, alg, T, p, n, k, data_seed, data="uniform", zeroh=False, **opts):
    """ProjectedALS / CoordinateDescent / ALSPGrad: options stored as JSON."""
    import json
    if ONLY and not any(o in name for o in ONLY):
        return
    rng = np.random.default_rng(data_seed)
    T = np.dtype(T)
    X = np.asfortranarray(rng.random((p, n)), dtype=T) if data == "uniform" else planted(rng, p, n, k, T)
    W0, H0 = O.randinit(p, n, k, T, rng, normalize=True, zeroh=zeroh)
    W, H = W0.copy(order="F"), H0.copy(order="F")
    cls = {"projals": O.ProjectedALS, "cd": O.CoordinateDescent, "alspgrad": O.ALSPGrad}[alg]
    r = O.solve(cls(T, **opts), X, W, H)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), X=X, W0=W0, H0=H0, W=W, H=H, niters=r.niters, converged=r.converged,
        objvalue=float(r.objvalue), alg=alg, k=k, opts=json.dumps(opts),
    )
    print(f"{name}: niters={r.niters} converged={r.converged} objvalue={float(r.objvalue):.9g}")


def case(name, alg, T, p, n, k, maxiter, tol, seed, data="uniform", **kw):
    if ONLY and not any(o in name for o in ONLY):
        return
    rng = np.random.default_rng(seed)
    T = np.dtype(T)
    X = np.asfortranarray(rng.random((p, n)), dtype=T) if data == "uniform" else planted(rng, p, n, k, T)
    W0, H0 = O.randinit(p, n, k, T, rng, normalize=True)
    W, H = W0.copy(order="F"), H0.copy(order="F")
    if alg in ("multmse", "multdiv"):
        inst = O.MultUpdate(T, obj=alg[4:], maxiter=maxiter, tol=tol, **kw)
    else:
        inst = O.GreedyCD(T, maxiter=maxiter, tol=tol, **kw)
    r = O.solve(inst, X, W, H)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), X=X, W0=W0, H0=H0, W=W, H=H, niters=r.niters, converged=r.converged,
        objvalue=float(r.objvalue), alg=alg, maxiter=maxiter, tol=float(tol), k=k,
        lambda_w=float(kw.get("lambda_w", 0.0)), lambda_h=float(kw.get("lambda_h", 0.0)),
        update_H=bool(kw.get("update_H", True)),
    )
    print(f"{name}: niters={r.niters} converged={r.converged} objvalue={float(r.objvalue):.9g}")


def case_nndsvd(name, T, p, n, k, seed, variant, zeroh=False, data="uniform"):
    """NNDSVD initialisation (initialization.jl:26-101) from a caller-supplied SVD (`initdata`), so that the fixture
    does not depend on a random range finder.  Stored under golden/init/ (the solver fixtures are globbed by name)."""
    if ONLY and not any(o in name for o in ONLY):
        return
    rng = np.random.default_rng(seed)
    T = np.dtype(T)
    X = np.asfortranarray(rng.random((p, n)), dtype=T) if data == "uniform" else planted(rng, p, n, k, T)
    U, S, Vt = np.linalg.svd(X.astype(np.float64), full_matrices=False)
    W, H = O.nndsvd(X, k, zeroh=zeroh, variant=variant, initdata=(U, S, Vt.T), rng=np.random.default_rng(seed + 1000))
    os.makedirs(os.path.join(OUT, "init"), exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "init", name + ".npz"), X=X, U=U[:, :k], S=S[:k], V=Vt[:k, :].T, W=W, H=H, k=k,
                        variant=variant, zeroh=zeroh, rng_seed=seed + 1000)
    print(f"{name}: variant={variant} zeroh={zeroh} |W|={np.linalg.norm(W):.9g} |H|={np.linalg.norm(H):.9g}")


if __name__ == "__main__":
    case_nndsvd("nndsvd_f64_std", np.float64, 40, 56, 5, 21, "std")
    case_nndsvd("nndsvd_f32_a_planted", np.float32, 48, 36, 4, 22, "a", data="planted")
    case_nndsvd("nndsvd_f64_ar_zeroh", np.float64, 32, 44, 6, 23, "ar", zeroh=True)
    case_nndsvd("nndsvd_f32_ar", np.float32, 36, 28, 3, 24, "ar")
    # BASELINE config 1: nnmf(rand(200,150), 5; alg=:multmse, init=:random, maxiter=50), Float64,
    # tol = cbrt(eps/100) as nnmf passes it (interf.jl:8)
    case("cfg1_multmse_f64", "multmse", np.float64, 200, 150, 5, 50, np.cbrt(np.finfo(np.float64).eps / 100), 0)
    case("multmse_f32_uniform", "multmse", np.float32, 96, 80, 8, 30, 1e-9, 1)
    case("multmse_f32_planted_reg", "multmse", np.float32, 64, 72, 6, 40, 1e-9, 2, data="planted", lambda_w=1e-3, lambda_h=2e-3)
    case("multmse_f64_noH", "multmse", np.float64, 40, 56, 4, 25, 1e-12, 3, update_H=False)
    case("multdiv_f64_uniform", "multdiv", np.float64, 72, 60, 5, 30, 1e-12, 4)
    case("multdiv_f32_planted", "multdiv", np.float32, 64, 48, 4, 30, 1e-9, 5, data="planted")
    case("greedycd_f64_uniform", "greedycd", np.float64, 60, 50, 4, 12, 1e-12, 6)
    case("greedycd_f32_planted_reg", "greedycd", np.float32, 48, 40, 4, 12, 1e-9, 7, data="planted", lambda_w=1e-4, lambda_h=1e-4)
    # SURVEY.md section 8f rows 1-2
    case2("projals_f64_uniform", "projals", np.float64, 60, 48, 4, 8, zeroh=True, maxiter=20, tol=1e-12)
    case2("projals_f32_planted", "projals", np.float32, 64, 56, 5, 9, data="planted", zeroh=True, maxiter=15, tol=1e-9, lambda_w=1e-2, lambda_h=2e-2)
    case2("projals_f64_noH", "projals", np.float64, 40, 36, 3, 10, maxiter=10, tol=1e-12, update_H=False)
    case2("cd_f64_uniform", "cd", np.float64, 56, 44, 4, 11, maxiter=15, tol=1e-12)
    case2("cd_f32_planted_reg_shuffle", "cd", np.float32, 48, 60, 5, 12, data="planted", maxiter=15, tol=1e-9, alpha=1e-3, l1ratio=0.5, shuffle=True, seed=42)
    case2("cd_f64_components_only", "cd", np.float64, 36, 40, 3, 13, maxiter=12, tol=1e-12, alpha=1e-2, l1ratio=0.25, regularization="components")
    case2("alspgrad_f64_uniform", "alspgrad", np.float64, 48, 40, 4, 14, maxiter=8, tol=1e-12, maxsubiter=50)
    case2("alspgrad_f32_planted", "alspgrad", np.float32, 40, 48, 3, 15, data="planted", maxiter=8, tol=1e-9, maxsubiter=50)
