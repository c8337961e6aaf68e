import sys, hashlib
sys.path.insert(0, '.')
import numpy as np
import nmf_jl_b200 as NMF

def problem(p, n, k, seed):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.random((p, n)), dtype=np.float32)
    W0, H0 = NMF.randinit(p, n, k, np.float32, normalize=True, rng=rng)
    return X, W0, H0

def h(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]

for (p, n, k, tol, every, maxiter) in [(384, 256, 8, 1e-9, 60, 60), (2048, 2048, 64, 1e-9, 30, 30), (1024, 4096, 128, 1e-9, 10, 10)]:
    X, W0, H0 = problem(p, n, k, 21 + k)
    for pdl in (0, 1):
        outs = {}
        with NMF.Session(engine="tc") as s:
            s.set_option("check_every", every)
            try:
                s.set_option("tc_pdl", pdl)
            except Exception:
                pass
            s.set_X(X)
            for rep in range(16):
                Wg, Hg = W0.copy(order="F"), H0.copy(order="F")
                r = s.solve(NMF.MultUpdate(np.float32, maxiter=maxiter, tol=tol), Wg, Hg)
                key = (r.niters, h(Wg), h(Hg), float(r.objvalue))
                outs[key] = outs.get(key, 0) + 1
        print(f"p={p} n={n} k={k} tol={tol} every={every} maxiter={maxiter} pdl={pdl}: {len(outs)} distinct outcomes", sorted(outs.items(), key=lambda kv: -kv[1])[:4], flush=True)
