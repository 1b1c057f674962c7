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
    assert lib.cucd_abi_version() == 5


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


def test_packed_cost_table_format_kernel_store_vs_host_helpers(cucd, emul):
    """the packed CTU table (include/cucudecide.h): the kernels' store code (replayed on the CPU), the C helpers
    cucd_unpack_costs / cucd_packed_cost and the numpy restatement agree; maxima and the 'not inside' codes survive"""
    lib = cucd.load_library()
    lib.cucd_unpack_costs.argtypes = [C.c_void_p, C.c_void_p]; lib.cucd_unpack_costs.restype = None
    lib.cucd_packed_cost.argtypes = [C.c_void_p, C.c_int, C.c_int]; lib.cucd_packed_cost.restype = C.c_uint32
    rng = np.random.default_rng(4)
    cost = np.zeros((341, 35), np.uint32)
    cost[:21] = rng.integers(0, 1 << 21, (21, 35))
    cost[21:85] = rng.integers(0, 32737, (64, 35))
    cost[85:] = rng.integers(0, 8161, (256, 35))
    cost[85, 0] = 8160; cost[340, 34] = 8160; cost[21, 0] = 32736; cost[0, 0] = 0xFFFFFFFF
    cost[100:104] = 0xFFFFFFFF; cost[30] = 0xFFFFFFFF; cost[7] = 0xFFFFFFFF        # PUs outside the picture
    for nthreads in (256, 97):
        packed = np.full(cucd.PACKED_CTU_BYTES + 16, 0xAB, np.uint8)
        assert emul.emul_store_packed_ctu(cost.ctypes.data_as(C.c_void_p), nthreads, packed.ctypes.data_as(C.c_void_p)) == cucd.PACKED_CTU_BYTES == 21980
        assert (packed[cucd.PACKED_CTU_BYTES:] == 0xAB).all()                       # nothing written past the table
        wide = np.zeros((341, 35), np.uint32)
        lib.cucd_unpack_costs(packed.ctypes.data, wide.ctypes.data)
        assert np.array_equal(wide, cost)
        assert np.array_equal(cucd.unpack_costs(packed[:cucd.PACKED_CTU_BYTES])[0], cost)
        for pu in (0, 7, 20, 21, 30, 84, 85, 86, 100, 103, 104, 339, 340):
            for mode in (0, 1, 17, 34):
                assert lib.cucd_packed_cost(packed.ctypes.data, pu, mode) == int(cost[pu, mode])


def _extreme_outlier_plane(bd, seed):
    """a textured picture whose first block rows carry the 4x4 patterns that maximise single AC coefficients (a whole picture of
    them would make every coefficient equal, and the reference's TCM fit degenerates into ~1e9 iterations on such input)"""
    from _util import textured_plane
    W = H = 128
    org = textured_plane(W, H, bd, seed=seed).copy()
    m = (1 << bd) - 1
    yy, xx = np.mgrid[0:4, 0:4]
    pats = [np.where((xx == 0) | (xx == 3), m, 0), np.where((yy == 0) | (yy == 3), m, 0), np.where(xx < 2, m, 0), np.where(yy < 2, m, 0),
            np.where((xx + yy) & 1, m, 0), np.where((xx == 1) | (xx == 2), m, 0), np.where(xx & 1, m, 0), np.where(yy & 1, m, 0)]
    for i, pat in enumerate(pats):
        org[0:4, 4 * i:4 * i + 4] = pat
        org[4:8, 4 * i:4 * i + 4] = m - pat
    return org


def test_outlier_plane_values_fit_a_byte(oracle):
    """cucd_frame_out.outlier_u8: |AC coefficient of the 4x4 transform| / 100 <= 163 for 8..10-bit content (DC is dropped)"""
    from _util import oracle_outlier_frame
    worst = 0
    for bd in (8, 10):
        obf, outl, _ = oracle_outlier_frame(oracle, _extreme_outlier_plane(bd, 3), bd)
        worst = max(worst, int(outl.max()))
        assert obf.max() <= 15
    assert 160 <= worst <= 163
