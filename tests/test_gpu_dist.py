"""Multi-GPU parity (row-sharded X/W, replicated H, one NCCL all-reduce per iteration).  Needs >= 2 GPUs:
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`; skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_solve_matches_oracle():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-4000:], out.stderr[-2000:])
    assert out.returncode == 0 and "dist_gpu_check ok" in out.stdout
