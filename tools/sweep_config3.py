"""BASELINE.json configs[2]: CRF-block sweep over the four decoder scales of a 480x640 input, shifted / unshifted
windows and embed dims, mapping each point to its binding roofline (tensor vs HBM) and the fraction achieved.
head_dim 32 (all four native points: 128/4, 256/8, 512/16, 1024/32, plus 64/2), head_dim 16, and head_dim 64 / 128
(csrc/crf_attn_wide.cu: correct, not tuned) are swept.

    python tools/sweep_config3.py > profiles/r01_config3_sweep.md
Each point: one CRFBlock fwd+bwd at batch 8, timed with CUDA events (5 iterations after 2 warm-ups); algorithmic
FLOPs = 3 * (22 C^2 + 196 C) per token, minimum HBM bytes = 8 C * 2 per token (bf16 I/O)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_depth_estimation_b200 import CRFBlock  # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def main():
    dev = torch.device("cuda:0")
    B = 8
    print("| H x W | C | heads | shift | windows | ms fwd+bwd | windows/s | TFLOP/s (algorithmic) | frac of bf16 peak | "
          "GB/s (min bytes) | frac of HBM peak | binding roofline of a fully fused block |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for (H, W) in [(120, 160), (60, 80), (30, 40), (15, 20)]:
        for C in [64, 128, 256, 512, 1024]:
            if B * H * W * C * 4 * 40 > 60e9:
                continue
            points = [(C // 32, 0), (C // 32, 3)] + ([(C // 16, 3)] if C <= 512 else [])
            points += [(C // 64, 3)] + ([(C // 128, 3)] if C >= 128 else [])
            for nH, shift in points:
                torch.manual_seed(0)
                blk = CRFBlock(C, nH, C, shift_size=shift).to(dev)
                blk.H, blk.W = H, W
                x = torch.randn(B, C, H, W, device=dev).flatten(2).transpose(1, 2).requires_grad_(True)
                v = torch.randn(B, C, H, W, device=dev).permute(0, 2, 3, 1).requires_grad_(True)
                dy = torch.randn(B, H * W, C, device=dev)

                def step():
                    y = blk(x, v, None)
                    y.backward(dy)
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                T = B * H * W
                hp, wp = -(-H // 7) * 7, -(-W // 7) * 7
                nwin = B * (hp // 7) * (wp // 7)
                flops = 3.0 * T * (22 * C * C + 196 * C)
                byts = 8.0 * C * 2 * T
                tf, gb = flops / ms / 1e9, byts / ms / 1e6
                bound = "tensor" if flops / (PEAKS["bf16_tflops"] * 1e12) > byts / (PEAKS["hbm_gbs"] * 1e9) else "hbm"
                print(f"| {H}x{W} | {C} | {nH} | {shift} | {nwin} | {ms:.3f} | {nwin / ms * 1e3:.0f} | {tf:.1f} | "
                      f"{tf / PEAKS['bf16_tflops']:.3f} | {gb:.0f} | {gb / PEAKS['hbm_gbs']:.3f} | {bound} |")


if __name__ == "__main__":
    main()
