"""GPU cases written after the round-1 GPU budget was spent: they had not run on hardware when this file was committed.

Group A -- new CHECKS of verified kernels: the reference's golden vectors for four 16-wide heads, and the whole product
model against the unmodified reference model's own numbers (tests/golden/model_64x96.npz).
Group B -- head_dim 64 / 128 (csrc/crf_attn_wide.cu, BASELINE.json configs[2]) driven through Python: stage-level forward /
backward, block level against the oracle, the reference's golden vectors for one 64-wide head.  (The kernels themselves
passed on a B200 through the Python-free harness tools/hwcheck -- profiles/r01_hwcheck.txt -- in the last seconds of the
round's GPU budget; these pytest cases have not run.)

Group C -- a candidate KERNEL for an existing stage: the multi-row LayerNorm forward (CRF_LN_ROWS=4, csrc/crf_misc.cu
ln_fwd_multirow_kernel; bit-identical arithmetic, more bytes in flight per warp), run through the existing LayerNorm cases.

Each group runs in its own SUBPROCESS with its opt-in switch set and a hard time limit, so that a device-side fault of an
unverified kernel cannot poison the CUDA context of the verified suite (this file also sorts last), and is reported as
xfail / xpass (non-strict) instead of failing the run.  The tails of the pytest output are kept under gpurun_out/.
Once a group is green on a B200, drop the switch and fold its cases into the plain suite.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GROUPS = {   # name: (pytest -k expression, extra environment)
    "new_checks": ("hd16 or test_full_model_matches_reference_model_golden", {}),
    "wide_heads": ("test_attn_fwd_wide or test_attn_bwd_wide or test_wide_head_block_vs_oracle or hd64",
                   {}),
    # Group C -- candidate kernel: LayerNorm forward with 4 rows in flight per warp (CRF_LN_ROWS=4), through the
    # existing LayerNorm / conversion parity cases and one whole block
    "ln_multirow": ("test_ln_fwd or test_layer_norm_standalone or test_colsum_cast_convert or test_config1_block",
                    {"CRF_LN_ROWS": "4"}),
    # Group D -- new kernel outside the block: the one-launch Adam step (crf_adam_step, training.LibAdam)
    "lib_adam": ("test_lib_adam_matches_torch_adam or test_prefetch_loader_cuda", {}),
}


@pytest.mark.xfail(strict=False, reason="cases added after the round-1 GPU budget was spent: not yet run on hardware")
@pytest.mark.parametrize("group", list(GROUPS))
def test_unverified_cases_isolated(group):
    kexpr, extra = GROUPS[group]
    env = dict(os.environ, CRF_TEST_UNVERIFIED="1", **extra)
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_stages.py", "tests/test_gpu_block.py", "-q", "-m", "gpu",
           "-k", kexpr, "-p", "no:cacheprovider"]
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
        rc, tail = r.returncode, (r.stdout or "")[-4000:] + (r.stderr or "")[-1500:]
    except subprocess.TimeoutExpired as e:
        rc, tail = -1, "TIMEOUT after 240 s\n" + ((e.stdout or b"").decode(errors="replace")[-3000:] if e.stdout else "")
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"unverified_{group}.log"), "w") as f:
            f.write(tail)
    assert rc == 0, tail
