"""tools/hwcheck.cu is the Python-free harness whose B200 run (profiles/r01_hwcheck.txt) verified the head_dim 64 / 128
attention kernels, the multi-row LayerNorm and the Adam step.  Its verdicts are only as good as its C++ reference: this
test builds the harness, lets it dump inputs + double-precision references (`--dry`, no GPU) and checks them against the
PyTorch restatement of the attention core the pytest cases use (tests/test_gpu_stages._attn_reference + autograd)."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from tests.helpers import rel_l2
from tests.test_gpu_stages import _attn_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "monocular_depth_estimation_b200", "csrc")


def test_harness_reference_equals_pytorch_reference(tmp_path):
    if shutil.which("nvcc") is None or not os.path.exists(os.path.join(CSRC, "libcrf_sm100.so")):
        pytest.skip("needs nvcc and the built library")
    subprocess.run(["make", "-C", CSRC, "hwcheck"], check=True, capture_output=True)
    subprocess.run([os.path.join(ROOT, "tools", "hwcheck"), "--dry", str(tmp_path)], check=True, capture_output=True)
    tags = sorted(os.listdir(tmp_path))
    assert len(tags) == 5
    for tag in tags:
        d = os.path.join(tmp_path, tag)
        B, H, W, C, nH, shift, nWin, _ = (int(t) for t in np.fromfile(d + "/meta.i32", dtype=np.int32))
        T = B * H * W

        def bf(name, shape):
            u = np.fromfile(f"{d}/{name}", dtype=np.uint16).astype(np.uint32) << 16
            return torch.from_numpy(u.view(np.float32).reshape(shape).copy()).double()

        def f64(name, shape):
            return torch.from_numpy(np.fromfile(f"{d}/{name}", dtype=np.float64).reshape(shape))

        qk, vb, dout = bf("qk.bf16", (T, 2 * C)), bf("vb.bf16", (T, C)), bf("dout.bf16", (T, C))
        bias = torch.from_numpy(np.fromfile(d + "/bias.f32", dtype=np.float32)).double()
        table = torch.from_numpy(np.fromfile(d + "/table.f32", dtype=np.float32).reshape(169, nH)).double()
        for t in (qk, vb, bias, table):
            t.requires_grad_(True)
        o, _ = _attn_reference(qk, vb, bias, table, B, H, W, C, nH, shift)
        o.backward(dout)
        scale = (C // nH) ** -0.5
        assert rel_l2(f64("o.f64", (T, C)), o) < 1e-12, tag
        assert rel_l2(f64("dv.f64", (T, C)), vb.grad) < 1e-12, tag
        assert rel_l2(f64("dk.f64", (T, C)), qk.grad[:, C:]) < 1e-12, tag
        assert rel_l2(f64("dq.f64", (T, C)), qk.grad[:, :C] * scale) < 1e-6, tag      # the harness multiplies by a float scale
        assert rel_l2(f64("dtable.f64", (169, nH)), table.grad) < 1e-12, tag
        if bias.grad[C:].abs().max() > 0:
            assert rel_l2(f64("dbias.f64", (2 * C,))[C:], bias.grad[C:]) < 1e-12, tag
