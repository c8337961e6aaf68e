#!/usr/bin/env python
"""bench.py -- NMF iteration throughput on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE NMF iteration (H half-step + W half-step + stop_condition) of MultUpdate(:mse).
Workload (synthetic, numpy default_rng, X ~ U[0,1), W0 column-sum-1, H0 ~ U[0,1)):
  N = 1 : BASELINE.json configs[1]  -- dense fp32 X 16384 x 16384, k = 128
  N > 1 : configs[4] family         -- X (16384*N) x 16384 row-sharded, 16384 rows per GPU (weak
          scaling; N = 8 is exactly configs[4]), one NCCL all-reduce of [W'W | W'X] per iteration.
value  = cells*iters/s = p_total * n * K / t  with t = device time of exactly K iterations, inputs
         resident in HBM (CUDA events on the solver's stream, max over ranks).
e2e    = the same metric through the public host API (NMF.solve with HOST buffers): includes the
         H2D copy of X, W0, H0 from pinned memory, the bf16 cache build, K iterations, the final
         objective and the D2H copy of W, H.  Median of 7 passes; every pass is listed with the library's
         own upload / loop times so that a slow pass can be attributed (host hiccup vs device).
parity_check: after the timed passes the returned factors are checked -- objvalue against an independent
         evaluation of 0.5*||X - WH||^2 from the returned W, H (torch, fp32 GEMM per column chunk, fp64 sum,
         summed over ranks), and, for N > 1, that the replicated H is bit-identical on every rank.
secondary: BASELINE configs[2] (MultUpdate :div) and configs[3] (GreedyCD) are timed in the same run (N = 1,
         device-resident, fewer iterations) and reported under the "secondary" key.
--impl reference: NMF.jl cannot run here (no Julia); the reference arm times the NumPy/OpenBLAS
         restatement of NMF.jl's update_wh! as written (oracle/, 6 sgemm + ratio loops) on ALL host cores
         (the BLAS thread count is set explicitly: torchrun exports OMP_NUM_THREADS=1), on the full
         p_total x n problem when host memory allows, else on one 16384-row slab, flagged "extrapolated".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

ROWS_PER_GPU = 16384
NCOLS = 16384
K = 128
METRIC = "NMF cells*iters/sec, MultUpdate(:mse) k=128 (iterations/sec in iters_per_sec)"
UNIT = "cells*iters/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.period = period_s
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for nm in dir(nv):
            if nm.startswith("nvmlClocksThrottleReason") or nm.startswith("nvmlClocksEventReason"):
                v = getattr(nv, nm)
                if isinstance(v, int) and v not in (0,) and "All" not in nm and "None" not in nm:
                    names[v] = nm.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and nm not in ("GpuIdle", "ApplicationsClocksSetting"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_problem(rows: int, n: int, k: int, seed: int, pinned: bool):
    """X (rows x n), W0 (rows x k), H0 (k x n): column-major float32 (returned as Julia-layout arrays)."""
    import torch
    rng = np.random.default_rng(seed)

    def alloc(shape_f):  # column-major array of shape (r, c) == C-order (c, r)
        r, c = shape_f
        t = torch.empty((c, r), dtype=torch.float32, pin_memory=pinned)
        return t, t.numpy().T

    tx, X = alloc((rows, n))
    for j0 in range(0, n, 1024):  # chunked fill keeps the float64 temporary small
        j1 = min(n, j0 + 1024)
        X[:, j0:j1] = rng.random((rows, j1 - j0), dtype=np.float32)
    tw, W = alloc((rows, k))
    W[...] = rng.random((rows, k), dtype=np.float32)
    W /= W.sum(axis=0, keepdims=True)  # normalize1_cols! (utils.jl:28-32), as nnmf's :random init does
    th, H = alloc((k, n))
    H[...] = np.random.default_rng(12345).random((k, n), dtype=np.float32)  # H0 identical on every rank
    return (tx, tw, th), X, W, H


def _blas_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core the box has."""
    from threadpoolctl import threadpool_info, threadpool_limits
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threadpool_limits(limits=ncpu, user_api="blas")
    threads = max([d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"] or [1])
    return threads, ncpu


def cpu_reference_rate(rows: int, n: int, k: int, steps: int, warmup: int, budget_s: float = 60.0):
    """NumPy/OpenBLAS restatement of update_wh!(::MultUpdMSE) (multupd.jl:83-116) timed on the host, all cores.
    At most `budget_s` seconds of timed work.  Returns (iterations/s, iterations timed, BLAS threads, seconds)."""
    import nmf_oracle as O
    threads, _ = _blas_all_cores()
    rng = np.random.default_rng(0)
    X = np.empty((rows, n), dtype=np.float32, order="F")
    for j0 in range(0, n, 1024):
        X[:, j0:j0 + 1024] = rng.random((rows, min(1024, n - j0)), dtype=np.float32)
    W, H = O.randinit(rows, n, k, np.float32, rng, normalize=True)
    upd = O.MultUpdMSE(np.float32, True, np.float32(0), np.float32(0), np.float32(np.sqrt(np.finfo(np.float32).eps)))
    st = upd.prepare_state(X, W, H)
    preW, preH = np.empty_like(W), np.empty_like(H)

    def one():  # one trip of the loop body of nmf_skeleton! (common.jl:64-74)
        np.copyto(preW, W)
        np.copyto(preH, H)
        upd.update_wh(st, X, W, H)
        O.stop_condition(W, preW, H, preH, 1e-9)

    for _ in range(max(1, min(warmup, 1))):
        one()
    t0 = time.perf_counter()
    done = 0
    while done < max(steps, 2) and (done < 2 or (time.perf_counter() - t0) < budget_s):
        one()
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, done, threads, dt


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    n_gpus = args.gpus
    p_total = ROWS_PER_GPU * n_gpus
    # the full p_total x n problem needs X and the reference's materialised WH (both p_total x n fp32) plus temporaries
    need = 2.3 * p_total * NCOLS * 4 + (4 << 30)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    full = n_gpus == 1 or avail > need
    rows = p_total if full else ROWS_PER_GPU
    its, done, threads, secs = cpu_reference_rate(rows, NCOLS, K, min(args.steps, 6), args.warmup, budget_s=60.0)
    scale = p_total // rows                      # 1 when the whole problem was timed
    it_full = its / scale                        # the as-written update costs O(p n k): linear in the number of rows
    value = p_total * NCOLS * it_full
    sample = (f"{done} iterations of the as-written MultUpdMSE update (6 sgemm + 2 ratio loops + stop_condition) on "
              f"{rows}x{NCOLS}, k={K}, {threads} BLAS threads, {secs:.1f} s" +
              ("" if full else f"; one {rows}-row slab of the {p_total}-row problem (host memory {avail / 2**30:.0f} GiB < "
                               f"{need / 2**30:.0f} GiB needed), rate divided by {scale}"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "iters_per_sec": it_full, "n_gpus": n_gpus,
        "steps": done, "warmup": 1, "ms_per_step": 1e3 * secs / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": (not full),
        "config": {"workload": f"MultUpdate(:mse) dense fp32 X {p_total}x{NCOLS}, k={K}", "rows_per_gpu": ROWS_PER_GPU,
                   "rows_timed": rows,
                   "note": "NMF.jl itself cannot run here (no Julia): NumPy/OpenBLAS restatement of its CPU path; "
                           "ms_per_step is the measured time of one timed step (of rows_timed rows)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import nmf_jl_b200 as NMF

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    rows, n, k = ROWS_PER_GPU, NCOLS, K
    p_total = rows * n_gpus
    hbm_peak, tf_peak, peak_src = peaks()

    keep, X, W0, H0 = make_problem(rows, n, k, seed=1000 + rank, pinned=True)
    tx, tw, th = keep
    dX = tx.cuda(non_blocking=True)
    dW = torch.empty_like(tw, device="cuda")
    dH = torch.empty_like(th, device="cuda")
    torch.cuda.synchronize()

    sess = NMF.Session(device=local_rank, engine="tc")
    if world > 1:
        uid = [NMF.Session.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sess.comm_init(rank, world, uid[0])
    sess.set_X_device(dX.data_ptr(), rows, n, rows, np.float32, keepalive=dX)
    sess.set_option("check_every", max(args.steps, args.warmup, 1))  # no host round trip inside the timed loop
    for kv in args.opt:
        key, val = kv.split("=")
        sess.set_option(key, val)

    def device_solve(iters, timed):
        dW.copy_(tw, non_blocking=True)
        dH.copy_(th, non_blocking=True)
        torch.cuda.synchronize()
        sess.set_option("time_kernels", (2 if args.timeline else 1) if timed else 0)
        return sess.solve_raw("multmse", np.float32, dW.data_ptr(), rows, dH.data_ptr(), k, k, max(iters, 2), 1e-30, 0.0, 0.0,
                              True, False, True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident number: W warm-up iterations, then windows of exactly K timed iterations.  The library aligns the
    # ranks on the DEVICE (flag barrier kernel) right before its start event and records the end event behind the last
    # iteration's kernels, so a window holds K iterations and nothing else; five windows, each max over ranks, median reported.
    device_solve(max(args.warmup, 3), timed=False)
    barrier()
    # roofline pass FIRST, right behind the warm-up: CUDA events around every mu_update_kernel launch (the dominant kernel).  The events
    # serialise the launches (no programmatic-dependent-launch overlap), so this pass is not the throughput number; it runs before the
    # long timed windows because those drive the board into its software power cap (clocks.reasons), and the peak it is compared with
    # (MEASURED_PEAKS.json: best of 10 copies) is a burst figure too.
    res_k = None if args.timeline else device_solve(args.steps, timed=True)
    barrier()
    windows = []
    with ClockSampler(local_rank) as clk:
        for _ in range(1 if args.timeline else 5):
            res = device_solve(args.steps, timed=args.timeline)   # the headline loop: no per-kernel events in the stream
            barrier()
            assert res.niters == max(args.steps, 2) and res.engine == 1, "tensor-core engine did not run the requested iterations"
            w = torch.tensor([res.solve_ms], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
            windows.append(float(w.item()))
    if res_k is None:
        res_k = res
    loop_ms = float(np.median(windows))
    iters = res.niters
    it_per_s = iters / (loop_ms * 1e-3)
    value = p_total * n * it_per_s

    # ---- end-to-end through the host API (host buffers in pinned memory)
    e2e_value = t_e2e = h2d = d2h = objv = float('nan')
    r2 = None
    passes = []
    parity = None
    if not args.no_e2e:
        e2e_iters = iters
        # the factors travel from / to PINNED host memory like X (a pageable W, H costs two staged copies of 8 MB each way)
        twh = torch.empty((k, rows), dtype=torch.float32, pin_memory=True)
        thh = torch.empty((n, k), dtype=torch.float32, pin_memory=True)
        Wh, Hh = twh.numpy().T, thh.numpy().T          # column-major rows x k and k x n views of the pinned buffers
        assert Wh.flags.f_contiguous and Hh.flags.f_contiguous and Wh.shape == W0.shape and Hh.shape == H0.shape
        sess2 = NMF.Session(device=local_rank, engine="tc")
        if world > 1:
            uid = [NMF.Session.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            sess2.comm_init(rank, world, uid[0])
        sess2.set_option("check_every", max(e2e_iters, 1))
        alg = NMF.MultUpdate(np.float32, obj="mse", maxiter=max(e2e_iters, 2), tol=1e-30)
        times = []
        for rep in range(8):  # first pass warms allocations; median of the next seven (the host side of a shared box is noisy)
            Wh[...] = W0
            Hh[...] = H0
            barrier()
            t0 = time.perf_counter()
            sess2.set_X(X)
            t1 = time.perf_counter()
            r2 = sess2.solve(alg, Wh, Hh)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if rep > 0:
                times.append(t2 - t0)
                passes.append({"seconds": round(t2 - t0, 5), "set_X_s": round(t1 - t0, 5), "solve_call_s": round(t2 - t1, 5),
                               "lib_upload_ms": round(r2.info["upload_ms"], 3), "lib_loop_ms": round(r2.info["solve_ms"], 3)})
        t_e2e = float(np.median(times))
        t_e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_e2e_t, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e_t.item())
        e2e_value = p_total * n * r2.niters / t_e2e
        h2d = (X.nbytes + W0.nbytes + H0.nbytes) / r2.niters
        d2h = (W0.nbytes + H0.nbytes + 8) / r2.niters
        objv = float(r2.objvalue)
        sess2.close()
        # ---- parity self-check on what the e2e solve returned (outside every timed region)
        tW = torch.from_numpy(np.ascontiguousarray(Wh.T)).cuda()       # (k, rows): column-major rows x k
        tH = torch.from_numpy(np.ascontiguousarray(Hh.T)).cuda()       # (n, k):   column-major k x n
        part = torch.zeros(1, dtype=torch.float64, device="cuda")
        for j0 in range(0, n, 2048):
            Rm = dX[j0:j0 + 2048] - tH[j0:j0 + 2048] @ tW              # rows j of X' minus (W H)'
            part += (Rm.double() ** 2).sum()
        hmax, hmin = tH.clone(), tH.clone()
        if dist is not None:
            dist.all_reduce(part)
            dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
        obj_indep = 0.5 * float(part.item())
        h_same = bool((hmax == hmin).all().item())
        rel = abs(objv - obj_indep) / obj_indep
        finite = bool(torch.isfinite(tW).all().item() and torch.isfinite(tH).all().item() and (tW >= 0).all().item() and (tH >= 0).all().item())
        parity = {"status": "ok" if (rel <= 1e-4 and h_same and finite) else "FAILED", "objvalue": objv, "objvalue_independent": obj_indep,
                  "objvalue_rel_err": rel, "H_identical_on_all_ranks": h_same, "factors_finite_nonnegative": finite}
        del tW, tH, hmax, hmin

    # ---- roofline of the dominant kernel (mu_update_kernel: one launch per half-step)
    launches = max(int(res_k.hot_kernel_launches), 1)
    kern_ms = res_k.hot_kernel_ms / launches if res_k.hot_kernel_launches else float("nan")
    # algorithmic bytes per launch: the bf16 X panel once + the factor read & written in fp32
    alg_bytes = rows * n * 2 + 2 * ((rows + n) / 2) * k * 4
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "update_kernel_traffic.json")) as f:
            traffic = float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        pass
    flops_iter = 4.0 * rows * n * k + 4.0 * k * k * (rows + n)
    tflops = flops_iter * it_per_s / 1e12

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "iters_per_sec": it_per_s, "n_gpus": n_gpus, "steps": iters,
            "warmup": max(args.warmup, 3), "ms_per_step": loop_ms / iters, "ms_per_step_windows": [w / iters for w in windows],
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 operands, f32 accumulate/state", "data": "synthetic",
            "config": {"workload": f"MultUpdate(:mse) dense fp32 X {p_total}x{n}, k={k}" +
                       (f" row-sharded over {n_gpus} GPUs" if n_gpus > 1 else " (BASELINE configs[1])"),
                       "rows_per_gpu": rows, "tol": 1e-30, "lambda": 0.0,
                       "l2": "inputs larger than L2: two 512 MiB bf16 X panels per iteration vs 126 MB L2, no flush needed",
                       "engine": "tc", "objvalue_e2e": objv},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "iters_per_sec": (r2.niters / t_e2e) if r2 else None, "seconds": t_e2e,
                    "passes": passes, "iters": r2.niters if r2 else 0, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "median of 7 passes (max over ranks); PCIe-bound: the 1 GiB upload of X per GPU dominates"},
            "parity_check": parity["status"] if parity else None, "parity": parity,
            "gpu_launches": int(res.kernel_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "traffic_source": "profiles/r2c_update_kernel_ncu_full.md (ncu --set full, same workload)",
                         "kernel": "mu_update_kernel<128,0>", "kernel_ms": kern_ms, "launches_timed": launches,
                         "loop_ms_per_step_with_kernel_events": res_k.solve_ms / max(res_k.niters, 1),
                         "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "tensor_tflops_whole_iteration": tflops, "tensor_frac_of_sustained_bf16": tflops / tf_peak},
        }
        # CPU baseline on rank 0 at N = 1 only (bounded sample, ~10-30 s)
        if n_gpus == 1 and not args.no_cpu:
            its, done, threads, secs = cpu_reference_rate(rows, n, k, steps=6, warmup=1, budget_s=25.0)
            line["cpu_baseline"] = {"value": p_total * n * its, "unit": UNIT, "iters_per_sec": its, "cores": threads, "kind": "port",
                                    "sample": f"{done} full-size iterations ({secs:.1f} s) of the NumPy/OpenBLAS restatement of NMF.jl's "
                                              f"MultUpdMSE update_wh! (as written, 6 sgemm) on {rows}x{n}, k={k}"}
        else:
            line["cpu_baseline"] = None
    sess.close()
    del sess, dX, dW, dH
    torch.cuda.empty_cache()
    if rank == 0:
        if n_gpus == 1 and not args.no_secondary:   # BASELINE configs[2], configs[3] in the same driver-run line
            sec = {}
            for wl in ("cfg3", "cfg4"):
                try:
                    sec[wl] = secondary_line(wl, "auto", steps=10, warmup=3, opts=[])
                except Exception as e:  # never lose the headline line to a secondary workload
                    sec[wl] = {"error": repr(e)}
                torch.cuda.empty_cache()
            sec["batched_replicates"] = []
            for R, kk in ((2, 128), (8, 32)):
                try:
                    sec["batched_replicates"].append(batched_line(R, kk, steps=20, warmup=3))
                except Exception as e:
                    sec["batched_replicates"].append({"error": repr(e), "replicates": R, "k": kk})
                torch.cuda.empty_cache()
            line["secondary"] = sec
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_secondary(args):
    print(json.dumps(secondary_line(args.workload, args.engine, args.steps, args.warmup, args.opt)), flush=True)


def secondary_line(workload, engine, steps, warmup, opts):
    """BASELINE configs[2] (MultUpdate :div, X 8192x65536, k=64) and configs[3] (GreedyCD, X 32768x32768, k=256):
    device-resident iteration rate through the C ABI (inputs generated on the device; CUDA-event time of the loop)."""
    import torch
    import nmf_jl_b200 as NMF

    class A:
        pass
    args = A()
    args.workload, args.engine, args.steps, args.warmup, args.opt = workload, engine, steps, warmup, opts
    cfg = {"cfg3": ("multdiv", 8192, 65536, 64), "cfg4": ("greedycd", 32768, 32768, 256)}[args.workload]
    alg, p, n, k = cfg
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    dX = torch.rand((n, p), device="cuda", generator=g)            # column-major p x n
    dW = torch.rand((k, p), device="cuda", generator=g)
    dW /= dW.sum(dim=1, keepdim=True)                               # columns of W sum to 1
    dH = torch.rand((n, k), device="cuda", generator=g)
    W0, H0 = dW.clone(), dH.clone()
    sess = NMF.Session(device=0, engine=args.engine)
    sess.set_X_device(dX.data_ptr(), p, n, p, np.float32, keepalive=dX)
    sess.set_option("check_every", max(args.steps, 1))
    for kv in args.opt:
        key, val = kv.split("=")
        sess.set_option(key, val)
    out = {}
    for iters, tag in ((max(args.warmup, 2), "warm"), (max(args.steps, 2), "timed")):
        dW.copy_(W0)
        dH.copy_(H0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = sess.solve_raw(alg, np.float32, dW.data_ptr(), p, dH.data_ptr(), k, k, iters, 1e-30, 0.0, 0.0, True, False, True)
        torch.cuda.synchronize()
        out[tag] = (r, time.perf_counter() - t0)
    r, wall = out["timed"]
    it_s = r.niters / (r.solve_ms * 1e-3)
    flops = (8.0 * p * n * k) if alg == "multdiv" else (4.0 * p * n * k + 4.0 * k * k * (p + n))
    # two definitions, side by side: SURVEY 8d's algorithmic bytes (two fp32 passes over X) and the bytes the kernels
    # actually stream (two passes over the bf16 cache of X) -- the second is the honest HBM fraction of this implementation
    bytes_fp32 = 2.0 * p * n * 4
    bytes_streamed = 2.0 * p * n * (2 if r.engine == 1 else 4)
    hbm_peak, tf_peak, src = peaks()
    line = {"metric": METRIC.replace("MultUpdate(:mse) k=128", f"{alg} k={k}"), "value": p * n * it_s, "unit": UNIT, "iters_per_sec": it_s,
            "n_gpus": 1, "steps": int(r.niters), "ms_per_step": r.solve_ms / r.niters, "dtype": "bf16 operands, f32 accumulate/state" if r.engine == 1 else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {alg} dense fp32 X {p}x{n}, k={k}", "engine": "tc" if r.engine == 1 else "simt"},
            "objvalue": r.objvalue, "coordinate_updates": int(r.coordinate_updates), "gpu_launches": int(r.kernel_launches),
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": src,
                         "achieved": bytes_streamed * it_s / 1e9, "frac": bytes_streamed * it_s / 1e9 / hbm_peak,
                         "note": "whole iteration, bytes the kernels stream (2 passes over the bf16 X cache)",
                         "achieved_fp32_definition": bytes_fp32 * it_s / 1e9, "frac_fp32_definition": bytes_fp32 * it_s / 1e9 / hbm_peak,
                         "tflops": flops * it_s / 1e12, "tensor_frac_of_sustained_bf16": flops * it_s / 1e12 / tf_peak}}
    sess.close()
    del dX, dW, dH, W0, H0
    return line


def batched_line(R, k, steps, warmup):
    """Batched replicates (interf.jl:85-101) on the configs[1] matrix: R MultUpdate(:mse) solves as one stacked
    iteration (nmfb200_solve_multmse_batched_f32, device-resident) next to ONE solve of the same k timed the same way."""
    import ctypes
    import torch
    import nmf_jl_b200 as NMF

    p = n = 16384
    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    dX = torch.rand((n, p), device="cuda", generator=g)
    dW = torch.rand((R * k, p), device="cuda", generator=g)
    dW /= dW.sum(dim=1, keepdim=True)
    dH = torch.rand((n, R * k), device="cuda", generator=g)
    W0, H0 = dW.clone(), dH.clone()
    sess = NMF.Session(device=0, engine="tc")
    sess.set_X_device(dX.data_ptr(), p, n, p, np.float32, keepalive=dX)
    sess.set_option("check_every", max(steps, 1))
    lib = sess._lib
    res = (NMF._lib.NmfResult * R)()
    ms_b = None
    for iters in (max(warmup, 2), max(steps, 2)):
        dW.copy_(W0)
        dH.copy_(H0)
        torch.cuda.synchronize()
        sess._check(lib.nmfb200_solve_multmse_batched_f32(sess._h, ctypes.c_void_p(dW.data_ptr()), p, ctypes.c_void_p(dH.data_ptr()), R * k, k, R,
                                                          iters, 1e-30, 0.0, 0.0, 1, 1, res))
        ms_b = res[0].solve_ms / res[0].niters
    objs = [res[r].objvalue for r in range(R)]
    ms_1 = None
    for iters in (max(warmup, 2), max(steps, 2)):   # one solve of the first replicate's factors (columns / rows [0, k))
        w1 = W0[:k].contiguous()
        h1 = H0[:, :k].contiguous()
        torch.cuda.synchronize()
        r1 = sess.solve_raw("multmse", np.float32, w1.data_ptr(), p, h1.data_ptr(), k, k, iters, 1e-30, 0.0, 0.0, True, False, True)
        ms_1 = r1.solve_ms / r1.niters
    hbm_peak, tf_peak, src = peaks()
    flops = R * (4.0 * p * n * k + 4.0 * k * k * (p + n))
    sess.close()
    del dX, dW, dH, W0, H0
    return {"workload": f"{R} replicates of MultUpdate(:mse) k={k} on dense fp32 X {p}x{n} as one stacked iteration",
            "replicates": R, "k": k, "steps": int(res[0].niters), "ms_per_stacked_iteration": ms_b,
            "replicate_iters_per_sec": R * 1e3 / ms_b, "one_solve_ms_per_iteration": ms_1, "one_by_one_replicate_iters_per_sec": 1e3 / ms_1,
            "speedup_vs_one_by_one": R * ms_1 / ms_b, "tflops": flops / (ms_b * 1e-3) / 1e12,
            "tensor_frac_of_sustained_bf16": flops / (ms_b * 1e-3) / 1e12 / tf_peak,
            "objvalue_first_replicate": objs[0], "objvalue_one_solve": r1.objvalue}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="experiments only: skip the end-to-end (host buffers) leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2] / configs[3] legs of the default run")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "batched"],
                    help="cfg2 = headline (default); cfg3/cfg4 = secondary; batched = batched replicates on the cfg2 matrix")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--timeline", action="store_true", help="print the per-phase event timeline of the iteration (stderr)")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value for experiments (device-resident leg)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "batched":
        print(json.dumps([batched_line(R, kk, args.steps, args.warmup) for R, kk in ((2, 128), (8, 32), (4, 64))]), flush=True)
        return
    if args.workload != "cfg2":
        run_secondary(args)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} processes (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
