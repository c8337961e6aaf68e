# Is the update kernel bound by per-SM TMA ingest (A 16 KB + B 16 KB per k-block)?  tc_debug bit 0 drops the B loads.
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nmf_jl_b200 as NMF
def run(p, n, k, tr, iters=40, debug=0):
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    X = torch.rand((n, p), device="cuda", generator=g)
    W = torch.rand((k, p), device="cuda", generator=g); H = torch.rand((n, k), device="cuda", generator=g)
    s = NMF.Session(engine="tc")
    s.set_option("tc_tile_rows", tr); s.set_option("check_every", iters); s.set_option("tc_debug", debug)
    s.set_X_device(X.data_ptr(), p, n, p, np.float32, keepalive=X)
    for timed in (0, 1):
        s.set_option("time_kernels", timed)
        r = s.solve_raw("multmse", np.float32, W.data_ptr(), p, H.data_ptr(), k, k, iters, 1e-30, 0, 0, True, False, True)
    kms = r.hot_kernel_ms / r.hot_kernel_launches
    gb = (p * n * 2 + (p + n) * k * 4) / 1e9
    print(f"p={p} n={n} k={k} tile_rows={tr} debug={debug}: kernel {kms*1e3:.1f} us  {gb/kms*1e3/1e3:.2f} TB/s  iter {r.solve_ms/r.niters*1e3:.1f} us", flush=True)
    s.close()
for dbg in (0, 1):
    for (p, k) in [(16384, 128), (18944, 128), (16384, 64), (16384, 256)]:
        run(p, p, k, 128, debug=dbg)
