"""Shared test helpers: golden-fixture loading and error metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYER_CASES = ["layer_9x10_c64", "layer_15x20_c64", "block_14x21_c128"]
HEAD_CASES = ["layer_8x13_c64_hd64", "layer_8x13_c64_hd16"]   # head widths other than 32 (BASELINE configs[2])


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 in float64 -- the 'rel' of BASELINE.json's tolerances."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def golden_layer_inputs(g, device="cpu"):
    """Rebuild the (possibly strided) inputs and the per-block parameter dicts of a layer fixture."""
    B, H, W, C, nH, depth, strided = [int(t) for t in g["meta"]]
    x = torch.from_numpy(g["x"]).to(device)
    v = torch.from_numpy(g["v"]).to(device)
    if strided:  # re-create the NCHW-view strides NewCRF.forward produces (newcrf_layers.py:426-427)
        x = x.transpose(1, 2).contiguous().transpose(1, 2)
        v = v.permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1)
    blocks = []
    for i in range(depth):
        pre = f"sd.blocks.{i}."
        blocks.append({k[len(pre):]: torch.from_numpy(t).to(device) for k, t in g.items()
                       if k.startswith(pre) and not k.endswith("relative_position_index")})
    return (B, H, W, C, nH, depth), x, v, blocks
