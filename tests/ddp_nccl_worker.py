"""Worker of tests/test_gpu_ddp_nccl.py (one process per GPU under torchrun, NCCL over NVLink): the DDP-averaged
parameter gradients of a batch sharded over the ranks must equal the single-process gradients of the concatenated
batch -- with the PRODUCT kernels (CRF blocks on the sm_100a library), SURVEY.md 4 item 5 / 8e.  The depth target is
pre-normalised (DepthNorm is a per-batch min-max: per-rank normalisation would differ from the global one) and the
model runs in eval mode (BatchNorm batch statistics would differ between a shard and the whole batch; gradients flow
all the same).  Rank 0 prints one JSON line: worst relative error per parameter group, the all-reduce checked by value.

argv[1] = "fp32": CRF blocks in the fp32 precision mode and TF32 off for the stock convolutions -- every kernel is then
deterministic to fp32 round-off in the batch composition, so the comparison is sharp (1e-4).  "bf16": the default
arithmetic; cuDNN picks different TF32 algorithms for a batch of 2 and a batch of 4, the 1e-4 differences flip bf16
roundings downstream, and the two runs carry two different realisations of the bf16 rounding noise: they agree to the
noise level of the bf16 tier (~1e-2 in the decoder, amplified in the encoder), which is what that mode asserts."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monocular_depth_estimation_b200 import training as T  # noqa: E402
from monocular_depth_estimation_b200.model import PTModel  # noqa: E402
from tests.helpers import fill_by_name, rel_l2  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, local_rank, world, device = T.init_distributed()
    assert world >= 2 and device.type == "cuda"
    torch.cuda.set_device(device)
    if mode == "fp32":
        import monocular_depth_estimation_b200 as pkg
        pkg.set_precision("fp32")
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    per_rank, H, W = 2, 64, 96
    gen = torch.Generator().manual_seed(21)
    image = torch.rand(per_rank * world, 3, H, W, generator=gen).to(device)
    depth = torch.rand(per_rank * world, 1, H, W, generator=gen).to(device)          # already in [0, 1]
    model = fill_by_name(PTModel()).eval().to(device)
    # single-process gradients of the WHOLE batch (every rank computes them: identical weights everywhere)
    loss_full = T.depth_loss(model(image), depth)
    loss_full.backward()
    full = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    # data-parallel: this rank's shard through DDP (bucketed NCCL all-reduce, mean over ranks)
    net = T.wrap_ddp(model, device, world)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    loss = T.depth_loss(net(image[sl]), depth[sl])
    loss.backward()
    torch.cuda.synchronize()
    errs = {}
    for k, p in model.named_parameters():
        if p.grad is None or k not in full:
            continue
        grp = ".".join(k.split(".")[:3])
        errs[grp] = max(errs.get(grp, 0.0), rel_l2(p.grad, full[k]))
    # the mean of the shard losses is the whole-batch loss (equal shard sizes)
    lt = loss.detach().clone()
    dist.all_reduce(lt)
    out = {"mode": mode, "world": world, "worst": max(errs.values()), "n_params": len(full), "loss_full": float(loss_full),
           "loss_mean_of_shards": float(lt) / world, "by_group": errs}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
