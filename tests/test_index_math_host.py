"""Integer index math of the device code, checked on the host (no GPU).

1. `WindowGeom` (csrc/crf_window.cuh) -- the closed-form pad + roll + window_partition map and the shift-mask region ids
   that the stand-alone index kernels and the attention kernels use -- is compiled with g++ (the struct is
   __host__ __device__) and compared, bit for bit, with the oracle's numpy maps, which are pinned to the reference
   (tests/test_oracle_golden.py), over every H, W in 1..30 and the geometries of BASELINE.json's configs.
2. The float-reciprocal division the attention kernels use for (window -> image, row, column)
   (csrc/crf_attn_async.cu: fast_div) is emulated in IEEE fp32 with numpy and shown to be exact on its whole domain
   (0 <= a < 2^22, every divisor the geometries above produce).
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import crf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "monocular_depth_estimation_b200", "csrc")

SHIM = r"""
#define __host__
#define __device__
#include "crf_window.cuh"
extern "C" void window_maps(int H, int W, int ws, int shift, int* dims, int* src, int* region) {
  crf::WindowGeom g(H, W, ws, shift);
  dims[0] = g.Hp; dims[1] = g.Wp; dims[2] = g.nWw; dims[3] = g.nW;
  if (!src) return;
  for (int w = 0; w < g.nW; ++w)
    for (int p = 0; p < ws * ws; ++p) {
      src[w * ws * ws + p] = g.source(w, p);
      region[w * ws * ws + p] = g.region(w, p);
    }
}
"""


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("geom")
    src, so = d / "shim.cpp", d / "libgeom.so"
    src.write_text(SHIM)
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-I", CSRC, str(src), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    ip = ctypes.POINTER(ctypes.c_int)
    lib.window_maps.argtypes = [ctypes.c_int] * 4 + [ip, ip, ip]
    lib.window_maps.restype = None
    return lib


def _maps(lib, H, W, shift):
    ip = ctypes.POINTER(ctypes.c_int)
    dims = np.zeros(4, dtype=np.int32)
    lib.window_maps(H, W, 7, shift, dims.ctypes.data_as(ip), None, None)
    n = int(dims[3]) * 49
    src, reg = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    lib.window_maps(H, W, 7, shift, dims.ctypes.data_as(ip), src.ctypes.data_as(ip), reg.ctypes.data_as(ip))
    return dims, src.reshape(-1, 49), reg.reshape(-1, 49)


GEOMS = [(H, W) for H in range(1, 31) for W in range(1, 31)] + \
        [(120, 160), (60, 80), (30, 40), (15, 20), (240, 320), (63, 84), (126, 161), (245, 322)]


def test_window_geom_matches_oracle_maps(shim):
    for (H, W) in GEOMS:
        for shift in (0, 3):
            dims, src, reg = _maps(shim, H, W, shift)
            Hp, Wp = O.padded_size(H, 7), O.padded_size(W, 7)
            assert dims.tolist() == [Hp, Wp, Wp // 7, (Hp // 7) * (Wp // 7)], (H, W)
            assert np.array_equal(src, O.window_source_index(H, W, 7, shift)), (H, W, shift)
            if shift:
                assert np.array_equal(reg, O.window_region_ids(H, W, 7, shift)), (H, W, shift)
                # the mask the attention kernels derive from the ids == the reference's additive mask
                mask = np.where(reg[:, :, None] == reg[:, None, :], 0.0, -100.0).astype(np.float32)
                assert np.array_equal(mask, O.shift_mask(H, W, 7, shift)), (H, W, shift)


def test_every_token_has_exactly_one_window_slot(shim):
    """window_reverse o window_partition is the identity on real tokens: each token is the source of exactly one slot
    (this is what lets the backward kernels write dq / dk / dv rows without atomics)."""
    for (H, W) in [(1, 1), (7, 7), (8, 13), (9, 10), (15, 20), (23, 17), (30, 40), (60, 80)]:
        for shift in (0, 3):
            _, src, _ = _maps(shim, H, W, shift)
            real = src[src >= 0]
            assert real.size == H * W and np.array_equal(np.sort(real), np.arange(H * W)), (H, W, shift)


def _fast_div(a, d):
    """numpy restatement of fast_div(a, d, 1.0f / d) in fp32, vectorised over a."""
    rd = np.float32(1.0) / np.float32(d)
    q = np.trunc((a.astype(np.float32) + np.float32(0.5)) * rd).astype(np.int64)
    r = a - q * d
    q = np.where(r < 0, q - 1, np.where(r >= d, q + 1, q))
    return q


def test_float_reciprocal_division_is_exact():
    a = np.arange(1 << 22, dtype=np.int64)
    divisors = set()
    for (H, W) in GEOMS + [(480, 640), (960, 1280)]:
        Hp, Wp = O.padded_size(H, 7), O.padded_size(W, 7)
        divisors.update({Wp // 7, (Hp // 7) * (Wp // 7)})
    for d in sorted(divisors):
        assert np.array_equal(_fast_div(a, d), a // d), d
