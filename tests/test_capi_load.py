"""The C-ABI library loads and exports every symbol include/cucudecide.h declares; without a GPU it
refuses to create a handle (there is no CPU compute path)."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(cucd):
    lib = cucd.load_library()
    names = cucd.declared_symbols()
    assert len(names) >= 32
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cucd_abi_version() == 3


def test_header_cites_reference_for_every_entry_point(cucd):
    text = open(os.path.join(os.path.dirname(cucd.LIB_PATH), "..", "include", "cucudecide.h")).read()
    for anchor in ("TEncSearch.cpp:2327-2361", "TEncSlice.cpp:878-1173", "TEncCu.cpp:589-600", "TEncSearch.cpp:421", "TComRdCost.h:109",
                   "TEncSearch.cpp:1092-1387", "TEncSearch.cpp:4340-4376", "tools_YS.cpp:1682-1839", "TEncPreanalyzer.cpp:64-139", "TEncSearch.cpp:2660"):
        assert anchor in text


def test_no_cpu_fallback_without_gpu(cucd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cucd.CucdError) as e:
        cucd.Engine(416, 240)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_link_the_oracle(cucd):
    """the shipped library must not reference oracle/ symbols"""
    import subprocess
    out = subprocess.run(["nm", "-D", cucd.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "hmref_" not in out


def test_host_tcm_fit_matches_golden(cucd):
    from _util import golden
    g = golden("obf_ai8.npz")
    # rebuild the histogram of |coeff/8| with numpy from the DCT definition (TEncSlice.cpp:55-77, 925-926)
    _, W, H, bd = [int(v) for v in g["f0_meta"]]
    org = g["f0_org"].astype(np.int64)
    blk = org.reshape(H // 4, 4, W // 4, 4).transpose(0, 2, 1, 3)          # [by][bx][y][x]
    T = np.array([[64, 64, 64, 64], [83, 36, -36, -83], [64, -64, -64, 64], [36, -83, 83, -36]], np.int64)
    s1, s2 = bd - 7, 8
    tmp = (np.einsum("kx,abyx->abky", T, blk) + (1 << (s1 - 1))) >> s1      # [by][bx][k][y]: first stage output transposed
    coef = (np.einsum("ly,abky->ablk", T, tmp) + (1 << (s2 - 1))) >> s2     # [by][bx][l (vertical freq)][k (horizontal freq)]
    c = coef.reshape(-1, 16)
    hist = np.zeros((16, 4096), np.uint32)
    for f in range(1, 16):
        bins = np.abs(c[:, f]) >> 3
        hist[f] = np.bincount(bins, minlength=4096)[:4096]
    yc, thr = cucd.tcm_fit(hist, c.shape[0])
    assert np.array_equal(yc[1:], g["f0_yc"][1:])
    assert np.array_equal(thr[1:], (g["f0_yc"][1:] * 8).astype(np.int32))
