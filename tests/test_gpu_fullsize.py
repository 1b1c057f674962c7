"""Full-size checks at BASELINE.json's resolutions.  The oracle is too slow for whole 1080p/2160p
pictures inside a test, so these combine (1) the oracle on a seeded subset of CTUs, and (2)
size-independent properties of the cost tables."""
import numpy as np
import pytest

from _util import P, i16p, u32p, oracle_outlier_frame, oracle_rmd_frame, pseudo_recon, textured_plane

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bd,W,H", [(8, 1920, 1080), (10, 3840, 2160)])
def test_fullsize_replay_subset_and_properties(cucd, oracle, bd, W, H):
    org = textured_plane(W, H, bd, seed=20261018)
    rec = pseudo_recon(org, bd)
    with cucd.Engine(W, H, bit_depth=bd, max_pictures=2) as eng:
        out = eng.frame(org, rec)
        cost = out["rmd_cost"]
        nctu = eng.ctus_per_pic
        wc = eng.ctus_per_row
        # (1) oracle on a seeded subset of CTUs: corners, edges (partial bottom row at 1080p), interior
        rng = np.random.default_rng(1)
        subset = sorted(set([0, wc - 1, nctu - wc, nctu - 1] + [int(c) for c in rng.integers(0, nctu, 8)]))
        for c in subset:
            want = oracle_rmd_frame(oracle, org, rec, bd, ctu_begin=c, ctu_end=c + 1)[0]
            assert np.array_equal(cost[c], want), c
        # (2a) PUs outside the picture carry the sentinel, every other entry is a real cost
        hc = (H + 63) // 64
        inside = np.ones((nctu, 341), bool)
        for c in range(nctu - wc, nctu):
            rows = H - (hc - 1) * 64
            pu = 0
            for d in range(5):
                n = 64 >> d
                for z in range(1 << (2 * d)):
                    py = sum(((z >> (2 * b + 1)) & 1) << b for b in range(d))
                    inside[c, pu] = (py + 1) * n <= rows
                    pu += 1
        assert np.array_equal((cost != 0xFFFFFFFF).all(axis=2), inside)
        assert np.array_equal((cost == 0xFFFFFFFF).any(axis=2), ~inside)
        # (2b) idempotence: same inputs, same table; and a second picture in the batch does not disturb the first
        again = eng.frames([org, rec], [rec, org])
        assert np.array_equal(again[0]["rmd_cost"], cost)
        # (2c) predicting a picture from itself: the DC/planar/angular cost of a CONSTANT picture is exactly 0
    flat = np.full((H, W), 1 << (bd - 1), np.int16)
    with cucd.Engine(W, H, bit_depth=bd) as eng:
        z = eng.frame(flat, flat)
    valid = z["rmd_cost"] != 0xFFFFFFFF
    assert (z["rmd_cost"][valid] == 0).all()
    assert (z["obf"] == 0).all() and (z["outlier"] == 0).all() and (z["ctu_src_had"] == 0).all()


def test_fullsize_features_vs_oracle(cucd, oracle):
    W, H, bd = 1920, 1080, 8
    org = textured_plane(W, H, bd, seed=3)
    with cucd.Engine(W, H, bit_depth=bd) as eng:
        out = eng.frame(org)
    obf, outl, yc = oracle_outlier_frame(oracle, org, bd)
    assert np.array_equal(out["yc"][1:], yc[1:])
    assert np.array_equal(out["obf"], obf)
    assert np.array_equal(out["outlier"], outl)
    # checksum of checksums: N_Outlier summed over any depth equals the sum of the OBF plane over whole CUs
    for d in range(4):
        s = 64 >> d
        assert int(out[f"n_outlier{d}"].sum()) == int(obf[: (H // s) * (s // 4), : (W // s) * (s // 4)].sum())
