import sys
sys.path.insert(0, '.')
import numpy as np
import nmf_jl_b200 as NMF
for (p, n, k) in [(2048, 2048, 64), (2048, 2048, 128), (4096, 1024, 64), (1024, 4096, 64), (512, 512, 32)]:
    rng = np.random.default_rng(5)
    X = np.asfortranarray(rng.random((p, n)), dtype=np.float32)
    W0, H0 = NMF.randinit(p, n, k, np.float32, normalize=True, rng=rng)
    with NMF.Session(engine="tc") as s:
        s.set_option("tc_debug", 16)
        s.set_X(X)
        for rep in range(3):
            Wg, Hg = W0.copy(order="F"), H0.copy(order="F")
            r = s.solve(NMF.MultUpdate(np.float32, maxiter=5, tol=1e-9), Wg, Hg)
            print(p, n, k, "objv", float(r.objvalue), flush=True)
