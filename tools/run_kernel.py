"""Launch individual kernels of the library at BASELINE config-2 shapes (for ncu captures and quick timing).

    python tools/run_kernel.py attn --stage 0 --iters 3
    python tools/run_kernel.py block --stage 0          # one CRFBlock fwd+bwd
Development / profiling tool; prints CUDA-event timings per call."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monocular_depth_estimation_b200 import _lib as L  # noqa: E402
from monocular_depth_estimation_b200 import ops  # noqa: E402

STAGES = {0: (120, 160, 128, 4), 1: (60, 80, 256, 8), 2: (30, 40, 512, 16), 3: (15, 20, 1024, 32)}


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["attn", "block", "gemm", "gemms"])
    ap.add_argument("--stage", type=int, default=0)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shift", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="", help="gemms: run only the shape with this label (e.g. \"d_fc2*gelu'\")")
    a = ap.parse_args()
    H, W, C, nH = STAGES[a.stage]
    B, dev = a.batch, torch.device("cuda:0")
    T = B * H * W
    torch.manual_seed(0)
    if a.what == "attn":
        qk = (torch.randn(T, 2 * C, device=dev) * 0.7).to(torch.bfloat16)
        vb = torch.randn(T, C, device=dev).to(torch.bfloat16)
        dout = torch.randn(T, C, device=dev).to(torch.bfloat16)
        bias = torch.randn(2 * C, device=dev) * 0.5
        table = torch.randn(169, nH, device=dev) * 0.5
        d = ops.make_desc(B, H, W, C, nH, a.shift, device=0)
        o, lse = ops.attn_fwd(d, qk, vb, bias, 32 ** -0.5, table)
        print("attn_fwd ms", timed(lambda: ops.attn_fwd(d, qk, vb, bias, 32 ** -0.5, table), a.iters))
        print("attn_bwd ms", timed(lambda: ops.attn_bwd(d, qk, vb, bias, 32 ** -0.5, table, lse, dout), a.iters))
    elif a.what == "block":
        from monocular_depth_estimation_b200 import CRFBlock
        blk = CRFBlock(C, nH, C, shift_size=a.shift).to(dev)
        blk.H, blk.W = H, W
        x = torch.randn(B, C, H, W, device=dev).flatten(2).transpose(1, 2).requires_grad_(True)
        v = torch.randn(B, C, H, W, device=dev).permute(0, 2, 3, 1).requires_grad_(True)

        def step():
            y = blk(x, v, None)
            y.backward(torch.ones_like(y))
        print("block fwd+bwd ms", timed(step, a.iters))
    elif a.what == "gemms":
        # every fprop / dgrad projection shape of this stage: (label, M, N, K, b_major, epilogue)
        shapes = [("qk", T, 2 * C, C, 0, L.EPI_STORE_BF16), ("proj+res", T, C, C, 0, L.EPI_BIAS_RES_F32),
                  ("fc1+gelu", T, 4 * C, C, 0, L.EPI_BIAS_GELU), ("fc2+res", T, C, 4 * C, 0, L.EPI_BIAS_RES_F32),
                  ("d_fc2*gelu'", T, 4 * C, C, 1, L.EPI_MUL_DGELU), ("d_fc1", T, C, 4 * C, 1, L.EPI_STORE_F32),
                  ("d_proj", T, C, C, 1, L.EPI_STORE_BF16), ("d_qk", T, C, 2 * C, 1, L.EPI_STORE_F32)]
        for name, M, N, K, bmaj, epi in shapes:
            if a.only and name != a.only:
                continue
            A = torch.randn(M, K, device=dev).to(torch.bfloat16)
            Wt = (torch.randn(K, N, device=dev) if bmaj else torch.randn(N, K, device=dev)).to(torch.bfloat16)
            f32 = epi in (L.EPI_STORE_F32, L.EPI_BIAS_RES_F32)
            out0 = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
            out1 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == L.EPI_BIAS_GELU else None
            aux = (torch.randn(M, N, device=dev) if epi == L.EPI_BIAS_RES_F32 else
                   torch.randn(M, N, device=dev).to(torch.bfloat16) if epi == L.EPI_MUL_DGELU else None)
            bias = torch.randn(N, device=dev) if not bmaj else None
            ms = timed(lambda: ops.gemm(A, Wt, M, N, K, b_major=bmaj, epilogue=epi, out0=out0, out1=out1, bias=bias,
                                        aux1=aux), a.iters)
            print(f"{name:12s} M{M} N{N} K{K}: {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
    else:
        M, N, K = T, 4 * C, C
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        Wt = torch.randn(N, K, device=dev).to(torch.bfloat16)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        act = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev)
        print("fc1+gelu ms", timed(lambda: ops.gemm(A, Wt, M, N, K, epilogue=L.EPI_BIAS_GELU, out0=pre, out1=act,
                                                     bias=bias), a.iters))


if __name__ == "__main__":
    main()
