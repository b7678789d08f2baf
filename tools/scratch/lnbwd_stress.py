import sys, torch
sys.path.insert(0, ".")
from monocular_depth_estimation_b200 import ops
dev = "cuda"
def rb(*s, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).to(torch.bfloat16).to(dev)
shapes = [(38400, 256, 512), (38400, 256, 1024), (153600, 128, 512), (1000, 256, 1024)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split(','))]
for (T, C, K) in shapes:
    g = torch.Generator(device="cpu").manual_seed(T + C + K)
    x = (torch.randn(T, C, generator=g) * 1.5 + 0.3).to(dev)
    gam = (1.0 + 0.1 * torch.randn(C, generator=g)).to(dev)
    dy, W = rb(T, K, seed=31), rb(K, C, seed=32, scale=K ** -0.5)
    dres = torch.randn(T, C, generator=g).to(dev)
    stats = torch.stack([x.mean(1), (x.var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
    dx0, dxb0, dg0, db0 = ops.dgrad_ln_bwd(dy, W, x, stats, gam, dres)
    torch.cuda.synchronize()
    for mode in ("both", "f32", "bf16"):
        bad = 0
        for it in range(12):
            dx, dxb, _, _ = ops.dgrad_ln_bwd(dy, W, x, stats, gam, dres, want_f32=mode != "bf16", want_bf16=mode != "f32")
            torch.cuda.synchronize()
            for name, a_, b_ in (("dx", dx, dx0), ("dxb", dxb, dxb0)):
                if a_ is None:
                    continue
                ne = (a_ != b_)
                if ne.any():
                    bad += 1
                    idx = ne.nonzero()
                    rows = idx[:, 0].unique()
                    print(f"T{T} C{C} K{K} mode {mode} it {it} {name}: {int(ne.sum())} elems differ, rows {rows[:8].tolist()} "
                          f"(n rows {len(rows)}), cols {idx[:, 1].min().item()}..{idx[:, 1].max().item()}, "
                          f"maxdiff {(a_.float() - b_.float()).abs().max().item():.3e}")
        print(f"T{T} C{C} K{K} mode {mode}: {bad} mismatching outputs of 12 runs")
