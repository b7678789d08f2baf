"""Training-step plumbing around the hot path: the reference loop's loss (src/train.py:89-100: SSIM + 0.1 * L1 on
min-max-normalised depth; SSIM from src/loss.py:57-88) and the data-parallel wrapper the reference lacks
(one process per GPU, NCCL gradient all-reduce over NVLink).  Windows never cross images, so sharding the batch
needs no data-path collective; the only exchange is the parameter-gradient average, which DDP overlaps with backward.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F


def depth_norm(depth):
    """src/utils.py:7-8 -- per-batch min-max normalisation."""
    return (depth - depth.min()) / (depth.max() - depth.min())


def ssim_loss(x, y):
    """monodepth2-style SSIM on 3x3 average pools with reflection padding (src/loss.py:57-88)."""
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    x, y = F.pad(x, (1, 1, 1, 1), mode="reflect"), F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sx = F.avg_pool2d(x * x, 3, 1) - mu_x * mu_x
    sy = F.avg_pool2d(y * y, 3, 1) - mu_y * mu_y
    sxy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + c1) * (2 * sxy + c2)
    d = (mu_x * mu_x + mu_y * mu_y + c1) * (sx + sy + c2)
    return torch.clamp((1 - n / d) / 2, 0, 1).mean()


def depth_loss(pred, depth_n):
    """loss = 1.0 * SSIM + 0.1 * L1 (src/train.py:94-100).  CUDA tensors go through the library's two-kernel
    forward / backward (functional.depth_loss); the composition of torch ops below is what it replaces and remains
    the CPU path of the gloo tests."""
    if pred.is_cuda:
        from .functional import depth_loss as _fused
        return _fused(pred, depth_n)
    return ssim_loss(pred, depth_n) + 0.1 * F.l1_loss(pred, depth_n)


def init_distributed():
    """One process per GPU (torchrun): returns (rank, local_rank, world_size, device)."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        if device.type == "cuda":
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=device)
        else:
            dist.init_process_group("gloo")
    return rank, local_rank, world, device


def wrap_ddp(model, device, world, grad_dtype=None):
    """The DDP wrapper `train.py` lacks: gradients are averaged with an NCCL all-reduce, bucketed and launched while
    backward is still running (crf3's 29 M parameters become ready first and are reduced under the rest).

    grad_dtype=torch.bfloat16 (or CRF_DDP_GRAD_BF16=1) exchanges the buckets in bf16 -- 90 MB instead of 180 MB per
    step -- through DDP's stock bf16_compress_hook: each rank's gradient is divided by the world size,
    rounded to bf16, summed by the all-reduce in bf16 and widened back to fp32.  Off by default: the averaged gradient
    then carries a bf16 rounding (rel ~ 2^-9) the fp32 exchange does not have."""
    if world <= 1:
        return model
    if grad_dtype is None and os.environ.get("CRF_DDP_GRAD_BF16", "0") == "1":
        grad_dtype = torch.bfloat16
    if grad_dtype not in (None, torch.float32, torch.bfloat16):
        raise ValueError(f"wrap_ddp: grad_dtype must be None, torch.float32 or torch.bfloat16 (got {grad_dtype})")
    from torch.nn.parallel import DistributedDataParallel as DDP
    # The reference keeps torchvision's ImageNet classifier head inside the encoder module although forward() never
    # uses it (model_mobileV3_large_newCRFs.py:176-182).  DDP requires every trainable parameter to receive a
    # gradient, so those parameters are frozen (they never get a gradient in the reference either).
    for name, p in model.named_parameters():
        if ".original_model.classifier." in name:
            p.requires_grad_(False)
    if device.type == "cuda":
        # broadcast_buffers=False: BatchNorm running statistics stay rank-local (46 BN layers would otherwise add a
        # broadcast of ~140 small buffers to every forward); gradients are what is averaged.
        net = DDP(model, device_ids=[device.index], bucket_cap_mb=64, broadcast_buffers=False,
                  gradient_as_bucket_view=os.environ.get("CRF_DDP_BUCKET_VIEW", "1") != "0")
    else:
        net = DDP(model)
    if grad_dtype == torch.bfloat16:
        import torch.distributed as dist
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net.register_comm_hook(dist.group.WORLD, default_hooks.bf16_compress_hook)
    return net


def train_step(model, optimizer, image, depth, autocast_dtype=torch.bfloat16):
    """One iteration of the reference loop (src/train.py:86-114): forward, loss, zero_grad, backward, Adam step."""
    depth_n = depth_norm(depth)
    use_amp = autocast_dtype is not None and image.is_cuda
    with torch.autocast("cuda", dtype=autocast_dtype, enabled=use_amp):
        pred = model(image)
    loss = depth_loss(pred if pred.is_cuda else pred.float(), depth_n)
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return loss


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, nn.Module) else model


def save_checkpoint(path, model, optimizer, epoch, loss):
    """The reference loop's checkpoint (src/train.py:143-155): a dict with `epoch`, `model_state_dict`,
    `optimizer_state_dict`, `loss`.  A DDP-wrapped model is unwrapped first, so the keys are the reference's
    (`Unet.0...`, `Unet.1...`, never `module.`-prefixed) and the file loads into the reference, a single-GPU run or
    another DDP run alike.  Under torchrun call it on rank 0 only (parameters are replicated)."""
    loss_t = loss.detach().cpu() if torch.is_tensor(loss) else torch.tensor(float(loss))
    torch.save({"epoch": int(epoch), "model_state_dict": _unwrap(model).state_dict(),
                "optimizer_state_dict": optimizer.state_dict(), "loss": loss_t}, path)


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """Resume as src/train.py:56-67 does: returns (epoch, loss).  Accepts checkpoints written by the reference loop
    (same format and keys) and ones whose keys carry a DDP `module.` prefix."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ck["model_state_dict"].items()}
    _unwrap(model).load_state_dict(sd, strict=True)
    if optimizer is not None and "optimizer_state_dict" in ck:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return int(ck.get("epoch", 0)), ck.get("loss")
