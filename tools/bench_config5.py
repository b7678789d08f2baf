"""BASELINE.json configs[4]: high-resolution inference, 960x1280, batch 16, bf16, through the fused CRF decoder
(25 760 / 6 624 / 1 728 / 480 windows per block at the four scales: the large-window-count, HBM-bound stress case).

    python tools/bench_config5.py [--batch 16] [--height 960] [--width 1280] [--iters 5] > profiles/rNN_config5.json

Prints one JSON object: whole-model images/s (MobileNetV3-large encoder + NeWCRFs decoder, torch.no_grad, bf16
autocast, channels-last), and per library kernel label the CUDA-event time with its algorithmic TFLOP/s and GB/s
(KernelTimer, same labels as bench.py's kernel_breakdown) -- windows/s per decoder stage follow from the attention
labels.  Inputs are device-resident; this is a measurement tool, not the driver's bench (that is bench.py, configs[1]).
Written after the round-1 GPU budget was spent: first run belongs to round 2.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocular_depth_estimation_b200 import _lib  # noqa: E402
from monocular_depth_estimation_b200.model import CRF_DIMS, NUM_HEADS, PTModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=960)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_config5: needs a CUDA device (sm_100a); there is no CPU path")
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = PTModel().to(dev).eval().to(memory_format=torch.channels_last)
    img = torch.rand(args.batch, 3, args.height, args.width, device=dev).contiguous(memory_format=torch.channels_last)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(img)

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    assert out.shape == (args.batch, 1, args.height, args.width) and bool(torch.isfinite(out.float()).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.crf_kernel_launches()
    e0.record()
    for _ in range(args.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    launches = (lib.crf_kernel_launches() - n0) // args.iters

    lib.crf_timing_enable(1)
    for _ in range(args.iters):
        step()
    torch.cuda.synchronize()
    lib.crf_timing_enable(0)
    need = lib.crf_timing_report(None, 0)
    buf = ctypes.create_string_buffer(need + 16)
    lib.crf_timing_report(buf, need + 16)
    kernels = json.loads(buf.value.decode())
    rows = []
    for k in sorted(kernels, key=lambda k: -k["total_ms"]):
        avg_ms = k["total_ms"] / max(k["launches"], 1)
        rows.append({"kernel": k["kernel"], "launches_per_iter": k["launches"] / args.iters, "avg_us": avg_ms * 1e3,
                     "tflops": k["flops"] / avg_ms / 1e9 if avg_ms > 0 else 0.0,
                     "gbs": k["bytes"] / avg_ms / 1e6 if avg_ms > 0 else 0.0})
    lib_ms = sum(k["total_ms"] for k in kernels) / args.iters
    stages = []
    for s, (C, nH) in enumerate(zip(CRF_DIMS, NUM_HEADS)):
        H, W = args.height // (4 << s), args.width // (4 << s)
        nwin = args.batch * (-(-H // 7)) * (-(-W // 7))
        att = [r for r in rows if r["kernel"].startswith(f"attn_fwd_B{args.batch}_{H}x{W}_C{C}_")]
        stages.append({"scale": f"1/{4 << s}", "H": H, "W": W, "C": C, "heads": nH, "windows_per_block": nwin,
                       "attn_fwd_avg_us": [round(r["avg_us"], 2) for r in att],
                       "attn_fwd_windows_per_s": [nwin / (r["avg_us"] * 1e-6) for r in att]})
    print(json.dumps({
        "config": f"inference {args.height}x{args.width}, batch {args.batch}, bf16 autocast, channels-last, no_grad "
                  "(BASELINE.json configs[4])",
        "images_per_s": args.batch / (ms * 1e-3), "ms_per_batch": ms, "library_kernel_ms_per_batch": lib_ms,
        "library_launches_per_batch": int(launches), "stages": stages, "kernels": rows[:24]}))


if __name__ == "__main__":
    main()
