"""Host-side logic that needs no GPU: module tree / state_dict compatibility with the reference, constructor
validation, the product path's refusal to run without CUDA."""
import pytest
import torch

import monocular_depth_estimation_b200 as pkg
from tests.helpers import load_golden


def test_state_dict_keys_and_shapes_match_reference_checkpoint():
    g = load_golden("layer_9x10_c64")
    ref_sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    layer = pkg.BasicCRFLayer(dim=64, depth=2, num_heads=2, v_dim=64, window_size=7)
    sd = layer.state_dict()
    assert sorted(sd.keys()) == sorted(ref_sd.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref_sd[k].shape), k
        assert str(v.dtype).split(".")[-1] == str(ref_sd[k].dtype), k
    idx = sd["blocks.0.attn.relative_position_index"]
    assert idx.dtype == torch.int64 and (idx.numpy() == ref_sd["blocks.0.attn.relative_position_index"]).all()
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in ref_sd.items()}, strict=True)


def test_blocks_alternate_shift_and_scale():
    layer = pkg.BasicCRFLayer(dim=128, depth=4, num_heads=4, v_dim=128)
    assert [b.shift_size for b in layer.blocks] == [0, 3, 0, 3]
    assert abs(layer.blocks[0].attn.scale - 32 ** -0.5) < 1e-12
    assert len(layer.blocks[0].fused_params()) == 13


def test_constructor_validation():
    with pytest.raises(AssertionError, match="shift_size must in 0-window_size"):
        pkg.CRFBlock(64, 2, 64, window_size=7, shift_size=9)
    with pytest.raises(NotImplementedError):
        pkg.CRFBlock(64, 2, 64, drop_path=0.1)
    with pytest.raises(NotImplementedError):
        pkg.WindowAttention(64, (7, 7), 2, 64, attn_drop=0.5)


def test_no_cpu_fallback():
    blk = pkg.CRFBlock(64, 2, 64)
    blk.H, blk.W = 7, 7
    with pytest.raises(RuntimeError, match="no CPU path"):
        blk(torch.randn(1, 49, 64), torch.randn(1, 7, 7, 64), None)


def test_model_tree_matches_reference_prefixes():
    from monocular_depth_estimation_b200.model import PTModel
    m = PTModel()
    keys = m.state_dict().keys()
    assert "Unet.1.crf0.crf_layer.blocks.1.attn.qk.weight" in keys
    assert "Unet.1.crf3.norm_crf.weight" in keys and "Unet.1.conv0.weight" in keys
    assert "Unet.0.original_model.features.0.0.weight" in keys
    n = sum(p.numel() for p in m.parameters())
    assert abs(n / 1e6 - 45.07) < 3.0   # 45.07 M (encoder incl. the unused classifier here)


def test_glue_host_paths_without_cuda():
    """The glue around the hot path keeps stock-PyTorch behaviour for CPU tensors (the gloo tests run the model's
    host logic on CPU), while everything that IS the hot path refuses to run without CUDA."""
    import torch.nn.functional as F
    from monocular_depth_estimation_b200 import functional as CF
    from monocular_depth_estimation_b200 import newcrf_layers as NL
    from monocular_depth_estimation_b200 import training as TR
    from oracle import model_oracle as MO
    torch.manual_seed(0)
    # PixelShuffle(2): falls through to torch on CPU
    x = torch.randn(2, 8, 3, 5)
    assert torch.equal(CF.pixel_shuffle2(x), F.pixel_shuffle(x, 2))
    # projection conv + bias: plain conv on CPU; the backward of the bias Function sums over (B, H, W)
    conv = torch.nn.Conv2d(4, 6, 3, padding=1)
    inp = torch.randn(2, 4, 5, 7)
    assert torch.equal(NL._project(conv, inp), conv(inp))
    y = torch.randn(2, 6, 5, 7, requires_grad=True)
    b = torch.randn(6, requires_grad=True)
    g = torch.randn(2, 6, 5, 7)
    NL._ConvBiasFn.apply(y, b).backward(g)
    assert torch.allclose(b.grad, g.sum((0, 2, 3)), atol=1e-5) and torch.equal(y.grad, g)
    # loss: CPU composition of torch ops == the oracle's
    p, t = torch.rand(2, 1, 9, 11), torch.rand(2, 1, 9, 11)
    assert torch.allclose(TR.depth_loss(p, t), MO.ssim_l1_loss(p, t), atol=1e-6)
    # the hot path itself has no CPU path
    for call in (lambda: CF.layer_norm(torch.randn(4, 64), torch.ones(64), torch.zeros(64)),
                 lambda: CF.depth_loss(p, t),
                 lambda: CF.crf_layer(torch.randn(1, 49, 64), torch.randn(1, 7, 7, 64), 7, 7, [[None] * 13], 2)):
        with pytest.raises(RuntimeError, match="CUDA"):
            call()


def test_checkpoint_format_round_trip(tmp_path):
    """src/train.py:143-155 / :56-67: {'epoch', 'model_state_dict', 'optimizer_state_dict', 'loss'}; a checkpoint of a
    wrapped model carries the reference's keys and resumes model and Adam state exactly."""
    from monocular_depth_estimation_b200 import training as TR

    class Wrapped(torch.nn.Module):          # stands in for DistributedDataParallel (same `.module` convention)
        def __init__(self, m):
            super().__init__()
            self.module = m

    torch.manual_seed(0)
    a = pkg.NewCRF(input_dim=8, embed_dim=64, v_dim=8, num_heads=2)
    opt = torch.optim.Adam(a.parameters(), 1e-3)
    for p in a.parameters():
        p.grad = torch.randn_like(p)
    opt.step()
    path = str(tmp_path / "global_checkpoint.pth")
    TR.save_checkpoint(path, Wrapped(a), opt, epoch=7, loss=torch.tensor(0.25))
    raw = torch.load(path, weights_only=False)
    assert sorted(raw.keys()) == ["epoch", "loss", "model_state_dict", "optimizer_state_dict"]
    assert sorted(raw["model_state_dict"].keys()) == sorted(a.state_dict().keys())     # no `module.` prefix
    b = pkg.NewCRF(input_dim=8, embed_dim=64, v_dim=8, num_heads=2)
    opt_b = torch.optim.Adam(b.parameters(), 1e-3)
    epoch, loss = TR.load_checkpoint(path, b, opt_b)
    assert epoch == 7 and float(loss) == 0.25
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    sa, sb = opt.state_dict()["state"], opt_b.state_dict()["state"]
    assert all(torch.equal(sa[i]["exp_avg"], sb[i]["exp_avg"]) for i in sa)
    # a checkpoint whose keys carry the DDP prefix loads too
    raw["model_state_dict"] = {"module." + k: v for k, v in raw["model_state_dict"].items()}
    torch.save(raw, path)
    TR.load_checkpoint(path, pkg.NewCRF(input_dim=8, embed_dim=64, v_dim=8, num_heads=2))


def test_synthetic_loader_and_prefetch_on_cpu():
    """The loader stand-in yields the reference loop's sample dicts (src/train.py:86-89); on a CPU device the prefetcher
    is a pass-through that preserves order and count."""
    from monocular_depth_estimation_b200 import training as TR
    a = list(TR.synthetic_batches(3, 2, 32, 48, seed=1, pin=False))
    b = list(TR.prefetch_to_device(TR.synthetic_batches(3, 2, 32, 48, seed=1, pin=False), "cpu"))
    assert len(a) == len(b) == 3
    for x, y in zip(a, b):
        assert sorted(x) == ["depth", "image"]
        assert x["image"].shape == (2, 3, 32, 48) and x["depth"].shape == (2, 1, 32, 48)
        assert torch.equal(x["image"], y["image"]) and torch.equal(x["depth"], y["depth"])
        assert 0 <= float(x["image"].min()) and float(x["image"].max()) <= 1 and float(x["depth"].min()) >= 1
    assert list(TR.prefetch_to_device(iter(()), "cpu")) == []
