"""Whole-step GPU timeline summary (torch.profiler/kineto): which kernels -- ours and the stock PyTorch ones around the
hot path -- take the step's GPU time, and how much of the wall time the GPU is busy.  Development tool, not a bench."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_depth_estimation_b200.model import PTModel  # noqa: E402
from monocular_depth_estimation_b200.training import train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/step_profile.txt")
    ap.add_argument("--cudnn-benchmark", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = a.cudnn_benchmark
    model = PTModel().to(dev).train().to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(model.parameters(), 1e-4, fused=True)
    img = torch.rand(a.batch, 3, 480, 640, device=dev).contiguous(memory_format=torch.channels_last)
    dep = torch.rand(a.batch, 1, 480, 640, device=dev)
    for _ in range(3):
        train_step(model, opt, img, dep)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        for _ in range(a.steps):
            train_step(model, opt, img, dep)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    rows = [(e.key, e.device_time_total / 1e3 / a.steps, e.count / a.steps) for e in ka
            if e.device_type == torch.autograd.DeviceType.CUDA]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    with open(a.out, "w") as f:
        f.write(f"GPU kernel time per step: {tot:.3f} ms over {sum(r[2] for r in rows):.0f} launches\n")
        for k, ms, n in rows[:70]:
            f.write(f"{ms:9.3f} ms {n:7.1f}x  {k[:150]}\n")
        # GPU time attributed to framework-level ops (children included): where the stock-PyTorch glue spends it
        ops = [(e.key, e.device_time_total / 1e3 / a.steps, e.count / a.steps) for e in ka
               if e.device_type != torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
        ops.sort(key=lambda r: -r[1])
        f.write("\nGPU ms per step by operator (inclusive)\n")
        for k, ms, n in ops[:60]:
            f.write(f"{ms:9.3f} ms {n:7.1f}x  {k[:110]}\n")
    with open(a.out, "a") as f:
        f.write("\nLargest layout / dtype copies by input shape (GPU ms per step)\n")
        kas = prof.key_averages(group_by_input_shape=True)
        rows = [(e.key, str(e.input_shapes)[:90], e.device_time_total / 1e3 / a.steps, e.count / a.steps) for e in kas
                if e.key in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::_to_copy", "aten::add_", "aten::sum",
                             "aten::pixel_shuffle", "aten::pixel_unshuffle", "aten::mul", "aten::add")]
        rows.sort(key=lambda r: -r[2])
        for k, shp, ms, n in rows[:45]:
            f.write(f"{ms:9.3f} ms {n:6.1f}x  {k:22s} {shp}\n")
    print(open(a.out).read()[-6500:])


if __name__ == "__main__":
    main()
