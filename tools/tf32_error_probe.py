"""How much error does TF32 tensor-core arithmetic introduce on this path?  Runs the oracle block on the GPU twice
(fp32 vs torch's TF32 matmuls) and prints rel-L2 errors of the output and every gradient.  Feasibility probe for a
TF32 mode targeting BASELINE.json's rel 1e-3 tolerance."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crf_oracle as O  # noqa: E402


def run(tf32, B, H, W, C, nH, shift):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.set_float32_matmul_precision("high" if tf32 else "highest")
    gen = torch.Generator().manual_seed(0)
    p = {k: t.cuda().requires_grad_(True) for k, t in O.init_block_params(C, nH, gen).items()}
    x = torch.randn(B, H * W, C, generator=gen).cuda().requires_grad_(True)
    v = torch.randn(B, H, W, C, generator=gen).cuda().requires_grad_(True)
    dy = torch.randn(B, H * W, C, generator=gen).cuda()
    y = O.crf_block(x, v, H, W, p, nH, 7, shift)
    y.backward(dy)
    out = {"y": y.detach(), "dx": x.grad, "dv": v.grad}
    out.update({k: t.grad for k, t in p.items()})
    return out


def main():
    for (B, H, W, C, nH, shift) in [(2, 60, 80, 128, 4, 3), (8, 15, 20, 1024, 32, 0)]:
        a = run(False, B, H, W, C, nH, shift)
        b = run(True, B, H, W, C, nH, shift)
        errs = {k: float((a[k].double() - b[k].double()).norm() / a[k].double().norm()) for k in a}
        print(f"B{B} {H}x{W} C{C}: max {max(errs.values()):.2e}", {k: f"{e:.1e}" for k, e in errs.items()})


if __name__ == "__main__":
    main()
