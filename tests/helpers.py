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


def fill_by_name(module):
    """Deterministic, construction-order-independent values for every parameter and BatchNorm statistic of a model:
    each tensor is drawn from a generator seeded by the CRC32 of its state_dict name.  The reference model (in the build
    container, tests/golden/make_golden.py), the oracle model and the product model therefore get identical weights
    from their NAMES alone -- no 180 MB checkpoint has to travel to the GPU box."""
    import zlib
    with torch.no_grad():
        for name, t in list(module.named_parameters()) + list(module.named_buffers()):
            if name.endswith("relative_position_index") or name.endswith("num_batches_tracked"):
                continue
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            r = torch.randn(t.shape, generator=g)
            if name.endswith("running_var"):
                t.copy_(1.0 + 0.2 * r.abs())
            elif name.endswith("running_mean"):
                t.copy_(0.1 * r)
            elif name.endswith("relative_position_bias_table"):
                t.copy_(0.2 * r)
            elif t.dim() >= 2:                      # conv / linear weights: He scaling keeps activations O(1)
                t.copy_(r * (2.0 / t[0].numel()) ** 0.5)
            elif name.endswith("weight"):           # LayerNorm / BatchNorm scales
                t.copy_(1.0 + 0.1 * r)
            else:                                   # biases
                t.copy_(0.1 * r)
    return module


# parameters whose gradients the full-model fixture stores (small tensors spread over encoder, bridge, every stage, head)
MODEL_GRAD_KEYS = ["Unet.0.original_model.features.0.0.weight", "Unet.1.conv0.bias", "Unet.1.conv1.weight",
                   "Unet.1.crf3.norm_crf.weight", "Unet.1.crf3.crf_layer.blocks.1.attn.relative_position_bias_table",
                   "Unet.1.crf2.crf_layer.blocks.0.attn.qk.bias", "Unet.1.crf2.proj_v.bias",
                   "Unet.1.crf1.crf_layer.blocks.1.mlp.fc2.bias", "Unet.1.crf1.norm_crf.bias",
                   "Unet.1.crf0.crf_layer.blocks.0.norm1.weight", "Unet.1.crf0.crf_layer.blocks.1.attn.proj.bias",
                   "Unet.1.crf0.proj_x.bias"]
