"""The N > 1 path on hardware: 2 ranks over NCCL / NVLink with the product kernels (needs >= 2 GPUs: run under
`gpurun --gpus 2`; skipped on a single-GPU box).  Averaged gradients == gradients of the concatenated batch."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("mode,bound", [("fp32", 1e-4), ("bf16", 8e-2)])
def test_two_rank_nccl_gradients_equal_concatenated_batch(mode, bound):
    """fp32 mode: sharp (the kernels are deterministic in the batch composition to fp32 round-off).  bf16 mode: to the
    rounding-noise level of the bf16 tier (see tests/ddp_nccl_worker.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if mode == "fp32" else "29518", os.path.join(ROOT, "tests", "ddp_nccl_worker.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"ddp_nccl_2rank_{mode}.json"), "w") as f:
            f.write(line + "\n")
    assert out["world"] == 2 and out["n_params"] > 250
    assert out["worst"] < bound, out
    assert abs(out["loss_full"] - out["loss_mean_of_shards"]) < (1e-6 if mode == "fp32" else 1e-4) * abs(out["loss_full"]), out
