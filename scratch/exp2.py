import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nmf_jl_b200 as NMF
def run(p, n, k, tr, iters=40, debug=0):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    X = torch.rand((n, p), device="cuda", generator=g)      # column-major p x n
    W = torch.rand((k, p), device="cuda", generator=g); H = torch.rand((n, k), device="cuda", generator=g)
    s = NMF.Session(engine="tc")
    s.set_option("tc_tile_rows", tr); s.set_option("check_every", iters); s.set_option("tc_debug", debug)
    s.set_X_device(X.data_ptr(), p, n, p, np.float32, keepalive=X)
    for timed in (0, 1):
        s.set_option("time_kernels", timed)
        r = s.solve_raw("multmse", np.float32, W.data_ptr(), p, H.data_ptr(), k, k, iters, 1e-30, 0, 0, True, False, True)
    kms = r.hot_kernel_ms / r.hot_kernel_launches
    gb = (p * n * 2 + (p + n) * k * 4) / 1e9
    print(f"p={p} n={n} k={k} tile_rows={tr} ctas={-(-p//tr) if tr else '?'} debug={debug}: kernel {kms*1e3:.1f} us  {gb/kms*1e3/1e3:.2f} TB/s  iter {r.solve_ms/r.niters*1e3:.1f} us", flush=True)
    s.close()
for (p, tr) in [(16384, 128), (18944, 128), (14336, 112), (16384, 112), (16576, 112), (9472, 64), (16384, 64), (18944, 64), (37888, 128), (32768, 128)]:
    run(p, p, 128, tr)
