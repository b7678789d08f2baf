"""ORACLE -- test infrastructure, not product code (see oracle/crf_oracle.py for the rules on who may import it).

CPU restatement of the full model the hot path sits in: torchvision MobileNetV3-large feature stack + the NeWCRFs
decoder (/root/reference/src/model_mobileV3_large_newCRFs.py:60-193, NewCRF wrapper newcrf_layers.py:367-434), with the
CRF layers evaluated by the fp32 oracle in crf_oracle.py.  Used for (a) the drop-in parity test of the whole model
and (b) the host-CPU baseline / `bench.py --impl reference` arm (the reference itself is Python and cannot travel to
the GPU box; this is its CPU path restated with the same torch ops).

Module / parameter names match the reference so one state_dict loads into the reference, this oracle and the product.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import crf_oracle as O


class _Holder(nn.Module):
    """Parameter container with the reference's attribute names; no forward."""


def _block_holder(C, nH, ws=7):
    b = _Holder()
    b.norm1 = nn.LayerNorm(C)
    b.attn = _Holder()
    b.attn.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, nH))
    nn.init.trunc_normal_(b.attn.relative_position_bias_table, std=.02)
    b.attn.register_buffer("relative_position_index", torch.from_numpy(O.relative_position_index(ws)))
    b.attn.qk = nn.Linear(C, 2 * C)
    b.attn.proj = nn.Linear(C, C)
    b.norm2 = nn.LayerNorm(C)
    b.mlp = _Holder()
    b.mlp.fc1 = nn.Linear(C, 4 * C)
    b.mlp.fc2 = nn.Linear(4 * C, C)
    return b


class OracleNewCRF(nn.Module):
    def __init__(self, input_dim, embed_dim, v_dim, num_heads, window_size=7, depth=2):
        super().__init__()
        self.embed_dim, self.num_heads, self.window_size = embed_dim, num_heads, window_size
        self.proj_x = nn.Conv2d(input_dim, embed_dim, 3, padding=1) if input_dim != embed_dim else None
        self.proj_v = nn.Conv2d(v_dim, embed_dim, 3, padding=1) if v_dim != embed_dim else None
        self.crf_layer = _Holder()
        self.crf_layer.blocks = nn.ModuleList([_block_holder(embed_dim, num_heads, window_size) for _ in range(depth)])
        self.norm_crf = nn.LayerNorm(embed_dim)

    def forward(self, x, v):
        if self.proj_x is not None:
            x = self.proj_x(x)
        if self.proj_v is not None:
            v = self.proj_v(v)
        B, C, H, W = x.shape
        tokens = x.flatten(2).transpose(1, 2)                                   # newcrf_layers.py:426
        v_hwc = v.transpose(1, 2).transpose(2, 3)                               # :427
        blocks = [dict(b.named_parameters()) for b in self.crf_layer.blocks]
        y = O.basic_crf_layer(tokens, v_hwc, H, W, blocks, self.num_heads, self.window_size)   # :429
        y = self.norm_crf(y)                                                    # :430-431
        return y.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()              # :432


class OracleDecoder(nn.Module):
    def __init__(self):
        super().__init__()
        heads, dims, vdims, enc = (4, 8, 16, 32), (128, 256, 512, 1024), (64, 128, 256, 512), (24, 40, 112, 160, 960)
        self.conv0 = nn.Conv2d(enc[4], vdims[3], 1)
        self.crf3 = OracleNewCRF(enc[3], dims[3], vdims[3], heads[3])
        self.crf2 = OracleNewCRF(enc[2], dims[2], vdims[2], heads[2])
        self.crf1 = OracleNewCRF(enc[1], dims[1], vdims[1], heads[1])
        self.crf0 = OracleNewCRF(enc[0], dims[0], vdims[0], heads[0])
        self.conv1 = nn.Conv2d(dims[0], 1, 3, padding=1)

    def forward(self, feats):
        e3 = self.crf3(feats[16], self.conv0(feats[17]))                        # model_...newCRFs.py:113-115
        e2 = self.crf2(feats[13], F.pixel_shuffle(e3, 2))                       # :116-117
        e1 = self.crf1(feats[7], F.pixel_shuffle(e2, 2))                        # :118-119
        e0 = self.crf0(feats[4], F.pixel_shuffle(e1, 2))                        # :120-121
        return F.interpolate(torch.sigmoid(self.conv1(e0)), scale_factor=4, mode="bilinear", align_corners=False)


class OracleEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        import torchvision.models as tvm
        self.original_model = tvm.mobilenet_v3_large(weights=None)              # no network: random init

    def forward(self, x):
        feats = [x]
        for layer in self.original_model.features:                              # :178-182
            feats.append(layer(feats[-1]))
        return feats


class OraclePTModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.Unet = nn.Sequential(OracleEncoder(), OracleDecoder())

    def forward(self, x):
        return self.Unet(x)


def depth_norm(d):
    return (d - d.min()) / (d.max() - d.min())                                  # src/utils.py:7-8


def ssim_l1_loss(pred, target):
    """src/train.py:94-100 with src/loss.py:57-88."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    x, y = F.pad(pred, (1, 1, 1, 1), mode="reflect"), F.pad(target, (1, 1, 1, 1), mode="reflect")
    mx, my = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sx, sy = F.avg_pool2d(x ** 2, 3, 1) - mx ** 2, F.avg_pool2d(y ** 2, 3, 1) - my ** 2
    sxy = F.avg_pool2d(x * y, 3, 1) - mx * my
    ssim = torch.clamp((1 - (2 * mx * my + c1) * (2 * sxy + c2) / ((mx ** 2 + my ** 2 + c1) * (sx + sy + c2))) / 2, 0, 1)
    return ssim.mean() + 0.1 * F.l1_loss(pred, target)


def train_step(model, optimizer, image, depth):
    """One iteration of the reference loop (src/train.py:86-114) in fp32 on the host CPU."""
    loss = ssim_l1_loss(model(image), depth_norm(depth))
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss
