"""head_dim 64 / 128 (csrc/crf_attn_wide.cu, BASELINE.json configs[2]) -- kernels written after the round-1 GPU budget
was spent: they compile for sm_100a but had not run on hardware when this file was committed.

Also here, for the same reason (first run belongs to round 2): the golden cases for head widths 64 / 16 and the
whole-model check against the reference model's own numbers (tests/golden/model_64x96.npz).

The wide-head parity cases (`*_wide*` in test_gpu_stages.py / test_gpu_block.py) therefore run here in a SUBPROCESS with
CRF_WIDE_HEADS=1, so that a device-side fault of an unverified kernel cannot poison the CUDA context of the verified
suite, and the result is reported as xfail / xpass (non-strict) instead of failing the run.  This file sorts last.
Once the cases are green on a B200, drop the opt-in switch and fold them into the plain suite.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.xfail(strict=False, reason="cases added after the round-1 GPU budget was spent: not yet run on hardware (CRF_WIDE_HEADS=1 selects them)")
@pytest.mark.parametrize("select", ["test_attn_fwd_wide", "test_attn_bwd_wide", "test_wide_head_block_vs_oracle",
                                    "test_head_width_golden", "test_full_model_matches_reference_model_golden"])
def test_wide_heads_isolated(select):
    env = dict(os.environ, CRF_WIDE_HEADS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_stages.py", "tests/test_gpu_block.py", "-q", "-x",
                        "-m", "gpu", "-k", select, "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout or "")[-3000:] + (r.stderr or "")[-1500:]
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"wide_heads_{select}.log"), "w") as f:
            f.write(tail)
    assert r.returncode == 0, tail
