"""Multi-GPU correctness on hardware (needs >= 2 visible GPUs: `gpurun --gpus 2`; skipped on a
1-GPU box): NCCL all-reduced gradients of 2 ranks == single-rank gradients on the concatenated
batch, equal and ragged shards. See tests/ddp_worker.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_allreduced_gradients_equal_single_rank_gradients(cuda, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    lines, r = [], None
    for port in ("29533", "29547"):  # a launcher / rendezvous failure (no worker output at all) is retried once
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "ddp_worker.py")]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
        lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        if lines or r.returncode == 0:
            break
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"ddp_grad_check_n{world}.json"), "w") as f:
        json.dump(lines, f, indent=1)
    assert r.returncode == 0, (lines, r.stderr[-1500:])
    assert len(lines) == world and all(l["ok"] for l in lines), lines
