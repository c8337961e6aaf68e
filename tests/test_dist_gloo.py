"""world_size-2 gloo tests (CPU): the host-side sharding logic and the algebra the data path relies on --
summing per-shard W_g'X_g and W_g'W_g over ranks reproduces the unsharded MU-MSE iteration, and the
stop_condition partial sums combine the same way.  No GPU, no library compute calls."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_shard_partition(NMF):
    for p in (1, 7, 16, 131072, 1001):
        for world in (1, 2, 3, 8):
            spans = [NMF.dist.row_shard(p, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == p
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        NMF.dist.row_shard(10, 2, 2)


class _FakeSession:
    """Duck-typed stand-in for Session: records what init_comm hands to the library."""
    made = 0

    def __init__(self):
        self.args = None

    @staticmethod
    def comm_unique_id():
        _FakeSession.made += 1
        return bytes(range(128))

    def comm_init(self, rank, world, uid):
        self.args = (rank, world, uid)


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import nmf_jl_b200 as NMF
    import nmf_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. unique-id plumbing: created on rank 0 only, identical bytes everywhere
        s = _FakeSession()
        NMF.dist.init_comm(s)
        assert s.args[0] == rank and s.args[1] == world and s.args[2] == bytes(range(128))
        assert _FakeSession.made == (1 if rank == 0 else 0)

        # 2. sharded MU-MSE iteration == unsharded (what the library does with NCCL, here with gloo + the oracle's loops)
        rng = np.random.default_rng(0)
        p, n, k = 37, 29, 4
        T = np.float64
        X = np.asfortranarray(rng.random((p, n)))
        W, H = O.randinit(p, n, k, T, rng, normalize=True)
        Wref, Href = W.copy(order="F"), H.copy(order="F")
        lo, hi = NMF.dist.row_shard(p, rank, world)
        Xg, Wg, Hg = X[lo:hi], np.asfortranarray(W[lo:hi]), H.copy(order="F")
        delta = np.sqrt(np.finfo(T).eps)
        for _ in range(5):
            preW, preH = Wg.copy(), Hg.copy()
            packed = torch.from_numpy(np.concatenate([(Wg.T @ Xg).ravel(), (Wg.T @ Wg).ravel()]))
            dist.all_reduce(packed)                         # the one exchange step per iteration
            A = packed[: k * n].numpy().reshape(k, n)
            G = packed[k * n:].numpy().reshape(k, k)
            Hg *= np.maximum(0, A) / (G @ Hg + delta)
            Wg *= np.maximum(0, Xg @ Hg.T) / (Wg @ (Hg @ Hg.T) + delta)
            part = torch.tensor([[((Wg[:, j] - preW[:, j]) ** 2).sum(), ((Wg[:, j] + preW[:, j]) ** 2).sum()] for j in range(k)])
            dist.all_reduce(part)                           # stop_condition partial sums of the W rows
        ref = O.solve(O.MultUpdate(T, maxiter=5, tol=1e-30), X, Wref, Href)
        assert ref.niters == 5
        np.testing.assert_allclose(Hg, Href, rtol=1e-10)
        np.testing.assert_allclose(Wg, Wref[lo:hi], rtol=1e-10)
        # W-side stop_condition sums, last iteration, against the full matrices
        preWfull = None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_sharded_iteration_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


# ---- round 2: the OWNERSHIP protocol of the row-sharded tensor-core solver (csrc/tc_shard.cuh), restated with gloo + NumPy ----------
def test_h_row_ownership_mirror_matches_library(NMF):
    """dist.h_row_ownership (Python) == nmfb200_shard_geometry (the C++ geometry the kernels use; host-only entry point)."""
    for n in (5, 96, 128, 129, 640, 1030, 16384, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = []
            for r in range(world):
                a, b, t = NMF.dist.shard_geometry(n, r, world)
                assert (a, b) == NMF.dist.h_row_ownership(n, r, world)
                assert t == (128 if n >= 128 else -(-max(n, 8) // 8) * 8) and (a % t == 0 or a == n)
                spans.append((a, b))
            assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        NMF.dist.shard_geometry(100, 3, 3)


def _owner_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import nmf_jl_b200 as NMF
    import nmf_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(x):  # every rank's array, in rank order
        out = [torch.empty_like(torch.from_numpy(x)) for _ in range(world)]
        dist.all_gather(out, torch.from_numpy(np.ascontiguousarray(x)))
        return [o.numpy() for o in out]

    try:
        rng = np.random.default_rng(1)
        p, n, k = 301, 389, 6                       # 389 rows of H': 4 tiles of 128 -> world 3 owns 2 / 2 / 0 tiles
        T = np.float64
        X = np.asfortranarray(rng.random((p, n)))
        W, H = O.randinit(p, n, k, T, rng, normalize=True)
        Wref, Href = W.copy(order="F"), H.copy(order="F")
        lo, hi = NMF.dist.row_shard(p, rank, world)
        Xg, Wg = X[lo:hi], np.asfortranarray(W[lo:hi])
        Ht = np.ascontiguousarray(H.T)              # H' (n x k): the rows this rank owns are the only ones it ever updates
        o0, o1 = NMF.dist.h_row_ownership(n, rank, world)
        owners = [NMF.dist.h_row_ownership(n, r, world) for r in range(world)]
        delta = np.sqrt(np.finfo(T).eps)
        tol, maxiter = 5e-3, 400
        converged, updates = False, 0
        Pw_parts = gather(Wg.T @ Wg)                # set-up: partial Grams W'W (PH_PW of the epoch before the loop)
        wsums_parts = None
        while True:
            # K2: W'W = sum of the ranks' partial Grams in rank order; stop_condition of the PREVIOUS iteration (also run
            # stand-alone after the last iteration of a host batch)
            Pw = sum(Pw_parts[1:], Pw_parts[0].copy())
            if wsums_parts is not None:
                ws = sum(wsums_parts[1:], wsums_parts[0].copy())
                hs = sum(hsums_parts[1:], hsums_parts[0].copy())
                if all(np.sqrt(ws[0, j]) <= tol * np.sqrt(ws[1, j]) and np.sqrt(hs[0, j]) <= tol * np.sqrt(hs[1, j]) for j in range(k)):
                    converged = True
                    break
            if updates == maxiter:
                break
            updates += 1
            preW, preHt = Wg.copy(), Ht.copy()
            # K1: partial numerators of ALL H rows over this rank's rows of X; every owner receives one slot per rank
            slots = gather(np.ascontiguousarray((Wg.T @ Xg).T))        # slots[s] = rank s's partial (n x k)
            # Ksum + K3: the OWNER sums its rows over the slots in rank order and applies the ratio to them -- nobody else does
            if o1 > o0:
                num = sum((s[o0:o1] for s in slots[1:]), slots[0][o0:o1].copy())
                Ht[o0:o1] *= np.maximum(0, num) / (Ht[o0:o1] @ Pw + delta)
            # the all-gather: every rank receives the owners' new rows (K3's epilogue stores)
            mine = np.ascontiguousarray(Ht[o0:o1]) if o1 > o0 else np.zeros((0, k))
            pad = np.zeros((max(b - a for a, b in owners), k))
            pad[: o1 - o0] = mine
            for r, rows in enumerate(gather(pad)):
                a, b = owners[r]
                Ht[a:b] = rows[: b - a]
            # K4 / K5: partial Gram H'H and H-side stop sums of the own rows, summed over ranks in rank order
            Ph = sum((g for g in gather(Ht[o0:o1].T @ Ht[o0:o1])[1:]), gather(Ht[o0:o1].T @ Ht[o0:o1])[0].copy())
            hsums_parts = gather(np.stack([((Ht[o0:o1] - preHt[o0:o1]) ** 2).sum(0), ((Ht[o0:o1] + preHt[o0:o1]) ** 2).sum(0)]))
            # K6: W-step on the local rows; K7: partial Gram W'W + W-side stop sums to everybody
            Wg *= np.maximum(0, Xg @ Ht) / (Wg @ Ph + delta)
            Pw_parts = gather(Wg.T @ Wg)
            wsums_parts = gather(np.stack([((Wg - preW) ** 2).sum(0), ((Wg + preW) ** 2).sum(0)]))
        ref = O.solve(O.MultUpdate(T, maxiter=maxiter, tol=tol), X, Wref, Href)
        assert ref.converged and converged and ref.niters == updates, (ref.niters, updates, ref.converged, converged)
        np.testing.assert_allclose(Ht.T, Href, rtol=1e-9)
        np.testing.assert_allclose(Wg, Wref[lo:hi], rtol=1e-9)
        same = gather(Ht)
        assert all((s == same[0]).all() for s in same)      # the replicated H is bit-identical on every rank
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, f"FAIL {type(e).__name__}: {e} {traceback.format_exc()[-400:]}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ownership_protocol_gloo(world):
    """The sharded algorithm of tc_shard.cuh, step for step (K1 ... K7, stop decision one iteration late), with gloo collectives
    in place of the peer-memory stores: same iteration count, `converged`, W and H as the unsharded oracle; H identical everywhere."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + world) % 90
    procs = [ctx.Process(target=_owner_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


# ---- round 2: the row-sharded tensor-core MultUpdate(:div) and GreedyCD (tc_engine.cu), restated with gloo + NumPy -----------------
def _greedy_rows_numpy(F, G, P, p_init, T):
    """greedycd.jl:139-165 for the rows given, with p_init supplied from outside (the sharded W-step gets it from an all-reduce)."""
    rows, k = F.shape
    eps = np.finfo(T).eps
    steps = 0
    for i in range(rows):
        g, f, fnew = G[i].copy(), F[i].copy(), np.zeros(k, dtype=T)

        def sd(w, gg):
            s = np.maximum(0, w - gg / (eps + np.diag(P))) - w
            return s, -gg * s - 0.5 * np.diag(P) * s * s
        s, d = sd(f, g)
        q = int(np.argmax(d))
        for _ in range(k * k):
            if not d[q] >= 0.001 * p_init:
                break
            sq = s[q]
            fnew[q] += sq
            g += sq * P[q, :]
            s, d = sd(f, g)        # the reference recomputes S against the row as it was at the start of the half-step (:155)
            q = int(np.argmax(d))
            steps += 1
        F[i] = np.maximum(0, f + fnew)
    return steps


def _row_max_d(F, G, P, T):
    eps = np.finfo(T).eps
    s = np.maximum(0, F - G / (eps + np.diag(P))) - F
    d = -G * s - 0.5 * np.diag(P) * s * s
    return max(-1.0, float(d.max(axis=1).max())) if F.shape[0] else -1.0


def _div_gcd_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import nmf_jl_b200 as NMF
    import nmf_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allsum(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    try:
        T = np.float64
        rng = np.random.default_rng(5)
        p, n, k = 41, 33, 3
        X = np.asfortranarray(rng.random((p, n)))
        W0, H0 = O.randinit(p, n, k, T, rng, normalize=True)
        lo, hi = NMF.dist.row_shard(p, rank, world)
        Xg = X[lo:hi]
        delta = np.sqrt(np.finfo(T).eps)

        # (1) MultUpdate(:div), multupd.jl:171-192: all-reduce [colsum(W_g) | W_g'Q_g], ratio on every rank's replica of H, W-step local
        Wg, H = np.asfortranarray(W0[lo:hi]), H0.copy(order="F")
        lam = delta                                        # the constructor's floor (multupd.jl:37-40)
        for _ in range(6):
            Q = Xg / (Wg @ H + delta)
            sW = allsum(Wg.sum(axis=0))                     # column sums over ALL rows of W
            WtQ = allsum(Wg.T @ Q)                          # numerators, summed over the ranks
            H *= WtQ / (sW[:, None] + lam)
            Q = Xg / (Wg @ H + delta)
            Wg *= (Q @ H.T) / (H.sum(axis=1)[None, :] + lam)
        Wr, Hr = W0.copy(order="F"), H0.copy(order="F")
        ref = O.solve(O.MultUpdate(T, obj="div", maxiter=6, tol=1e-30), X, Wr, Hr)
        assert ref.niters == 6
        np.testing.assert_allclose(H, Hr, rtol=1e-10)
        np.testing.assert_allclose(Wg, Wr[lo:hi], rtol=1e-10)
        obj = float(allsum(np.array([np.sum(np.where(Xg > 0, Xg * np.log(np.where(Xg > 0, Xg, 1) / (Wg @ H)) - Xg + Wg @ H, Wg @ H))]))[0])
        assert abs(obj - float(ref.objvalue)) <= 1e-9 * float(ref.objvalue)    # the objective's data term is a sum over the ranks' rows

        # (2) GreedyCD, greedycd.jl:94-178.  W-step local except p_init (max over all ranks).  H-step: every rank forms
        #     H (W_g'W_g) - X_g'W_g from ITS rows and ITS OWN Gram; the sum over ranks is the gradient (lambda added on rank 0 only);
        #     p_init from the complete gradient; the coordinate loop runs replicated with the all-reduced W'W.
        lam_w, lam_h = 1e-3, 2e-3
        Wg, Ht = np.asfortranarray(W0[lo:hi]), np.array(H0.T, order="C", copy=True)   # (ascontiguousarray would alias H0)
        for _ in range(3):
            Ph = Ht.T @ Ht                                                     # replicated
            Gw = Wg @ Ph - Xg @ Ht + lam_w
            pw = torch.tensor([_row_max_d(Wg, Gw, Ph, T)], dtype=torch.float64)
            dist.all_reduce(pw, op=dist.ReduceOp.MAX)                          # 1-float all-reduce (max)
            _greedy_rows_numpy(Wg, Gw, Ph, float(pw.item()), T)
            Pw_own = Wg.T @ Wg
            Gh = allsum(Ht @ Pw_own - Xg.T @ Wg + (lam_h if rank == 0 else 0.0))
            Pw = allsum(Pw_own)
            _greedy_rows_numpy(Ht, Gh, Pw, _row_max_d(Ht, Gh, Pw, T), T)       # replicated: identical on every rank
        Wr, Hr = W0.copy(order="F"), H0.copy(order="F")
        ref = O.solve(O.GreedyCD(T, maxiter=3, tol=1e-30, lambda_w=lam_w, lambda_h=lam_h), X, Wr, Hr)
        np.testing.assert_allclose(Ht.T, Hr, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(Wg, Wr[lo:hi], rtol=1e-8, atol=1e-12)
        same = [torch.empty_like(torch.from_numpy(Ht)) for _ in range(world)]
        dist.all_gather(same, torch.from_numpy(Ht))
        assert all((s_ == same[0]).all() for s_ in same)                        # the replicated H is bit-identical on every rank
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, f"FAIL {type(e).__name__}: {e} {traceback.format_exc()[-600:]}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_div_and_greedycd_protocol_gloo(world):
    """What tc_solve_div_kp / tc_solve_gcd_kp exchange between ranks (NCCL there, gloo here): the same iterates as the unsharded oracle."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() + 7 * world) % 90
    procs = [ctx.Process(target=_div_gcd_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
