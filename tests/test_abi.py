"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol
include/crf_sm100.h declares; host-side argument validation returns errors instead of crashing."""
import ctypes
import os
import re

import pytest

from monocular_depth_estimation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "crf_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("extension not built (run __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/crf_sm100.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "python binding list and header disagree"


def test_version_and_error_reporting_without_gpu():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("extension not built")
    lib = _lib.lib()
    assert lib.crf_abi_version() == 1
    d = _lib.BlockDesc()
    d.B, d.H, d.W, d.C, d.num_heads, d.window, d.shift = 1, 7, 7, 96, 3, 7, 0   # C not a multiple of 64
    s = ctypes.c_size_t()
    assert lib.crf_block_sizes(ctypes.byref(d), ctypes.byref(s), None, None) != 0
    assert b"multiple of 64" in lib.crf_last_error()
    d.C, d.num_heads, d.shift = 128, 4, 7
    assert lib.crf_block_sizes(ctypes.byref(d), ctypes.byref(s), None, None) != 0
    assert b"shift_size must in 0-window_size" in lib.crf_last_error()
    d.shift = 3
    assert lib.crf_block_sizes(ctypes.byref(d), ctypes.byref(s), None, None) == 0 and s.value > 0


def test_struct_sizes_match_header_layout():
    # 12 int32 + 7 int64 ; 13 pointers + 2 floats ; 13 pointers
    assert ctypes.sizeof(_lib.BlockDesc) == 12 * 4 + 7 * 8
    assert ctypes.sizeof(_lib.BlockParams) == 13 * 8 + 8
    assert ctypes.sizeof(_lib.BlockGrads) == 13 * 8


def test_ctypes_structs_match_the_c_header_field_by_field(tmp_path):
    """Compile a tiny C program against include/crf_sm100.h (gcc) that prints sizeof and every offsetof, and compare
    with the ctypes mirrors in _lib.py -- a silent mismatch here would corrupt arguments at the drop-in boundary."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"crf_block_desc": _lib.BlockDesc, "crf_block_params": _lib.BlockParams,
               "crf_block_grads": _lib.BlockGrads, "crf_layer_args": _lib.LayerArgs, "crf_gemm_args": _lib.GemmArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "crf_sm100.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"

