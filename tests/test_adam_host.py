"""The library Adam step (csrc/crf_adam.cu, training.LibAdam), checked without a GPU:
the per-element update (csrc/crf_adam_math.h, shared verbatim by the device kernel) is compiled with g++ and compared
with torch.optim.Adam over several steps; the per-tensor CTA counts cover every element exactly once."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "monocular_depth_estimation_b200", "csrc")

SHIM = r"""
#include "crf_adam_math.h"
extern "C" void adam_steps(float* p, const float* g, float* m, float* v, int n, int steps, double lr, double b1, double b2,
                           double eps, double wd) {
  for (int t = 1; t <= steps; ++t) {
    const crf::AdamCoef c = crf::adam_coef(lr, b1, b2, eps, wd, (double)t);
    for (int i = 0; i < n; ++i) crf::adam_update(c, p[i], g[(t - 1) * n + i], m[i], v[i]);
  }
}
"""


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("adam")
    src, so = d / "shim.cpp", d / "libadam.so"
    src.write_text(SHIM)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    fp = ctypes.POINTER(ctypes.c_float)
    lib.adam_steps.argtypes = [fp, fp, fp, fp, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 5
    lib.adam_steps.restype = None
    return lib


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_element_update_matches_torch_adam(shim, wd):
    n, steps, lr, b1, b2, eps = 4099, 25, 1e-4, 0.9, 0.999, 1e-8     # the reference loop: Adam(lr=1e-4), defaults
    gen = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=gen)
    grads = torch.randn(steps, n, generator=gen) * torch.logspace(-6, 1, n)   # gradient scales from 1e-6 to 10
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd)
    for t in range(steps):
        ref.grad = grads[t].clone()
        opt.step()
    p = p0.numpy().copy()
    m, v = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    g = np.ascontiguousarray(grads.numpy())
    fp = ctypes.POINTER(ctypes.c_float)
    shim.adam_steps(p.ctypes.data_as(fp), g.ctypes.data_as(fp), m.ctypes.data_as(fp), v.ctypes.data_as(fp), n, steps,
                    lr, b1, b2, eps, wd)
    st = opt.state[ref]
    upd_ref = (ref.detach() - p0).numpy()
    upd = p - p0.numpy()
    # parameters are O(1) and move by ~lr per step: agreement to two ulps of the parameter after 25 steps, i.e. the
    # accumulated displacement (~2.5e-3) agrees to ~1e-4 of itself
    assert np.abs(upd - upd_ref).max() <= 2.5e-7 * max(1.0, float(np.abs(p0.numpy()).max()))
    assert np.abs(upd - upd_ref).sum() <= 2e-4 * np.abs(upd_ref).sum()
    # per-element scale of the effective gradient g + wd * p (the moments are sums with cancellation)
    gs = np.abs(g).max(axis=0) + wd * (np.abs(p0.numpy()) + 1e-2)
    assert (np.abs(m - st["exp_avg"].numpy()) <= 2e-6 * np.abs(m) + 2e-6 * gs).all()
    assert (np.abs(v - st["exp_avg_sq"].numpy()) <= 2e-6 * np.abs(v) + 2e-6 * gs * gs).all()


def test_chunk_counts_cover_every_element_once():
    from monocular_depth_estimation_b200.training import LibAdam
    numels = [1, 3, 16384, 16385, 49 * 13, 1024 * 4096, 7, 32768]
    chunk = LibAdam.CHUNK
    assert chunk % 4 == 0 and chunk >= 1024
    for n, c in zip(numels, LibAdam.chunk_counts(numels, chunk)):
        assert sum(min(n, (k + 1) * chunk) - k * chunk for k in range(c)) == n and (c - 1) * chunk < n


def test_lib_adam_refuses_cpu_parameters():
    from monocular_depth_estimation_b200.training import LibAdam
    w = torch.nn.Parameter(torch.randn(8))
    w.grad = torch.randn(8)
    with pytest.raises(RuntimeError, match="CUDA"):
        LibAdam([w], lr=1e-4).step()


def test_dense_and_same_layout_helpers():
    """LibAdam updates memory element-wise: parameters only need to be dense (any permutation of a contiguous layout,
    e.g. channels-last convolution weights), and a gradient must have the parameter's element order -- strides of
    size-1 dimensions do not matter (1x1 convolution weights in channels-last report different ones)."""
    from monocular_depth_estimation_b200.training import _dense, _same_layout
    a = torch.randn(6, 4, 3, 3)
    cl = a.contiguous(memory_format=torch.channels_last)
    assert _dense(a) and _dense(cl) and _dense(a.permute(2, 0, 3, 1)) and _dense(torch.randn(5)[:3])
    assert not _dense(a[:, ::2]) and not _dense(torch.randn(10)[::2]) and not _dense(a[:, :, :2])
    assert _same_layout(a, a.clone()) and not _same_layout(a, cl) and _same_layout(cl, cl.clone())
    w11 = torch.randn(24, 40, 1, 1)
    w11_cl = w11.contiguous(memory_format=torch.channels_last)      # same element order, different size-1 strides
    assert _same_layout(w11, w11_cl) and _dense(w11_cl)
    assert not _same_layout(a, torch.randn(6, 4, 3, 2))
