"""tcgen05 (kind::i8) Hadamard SATD building block against the oracle's xCalcHADs8x8 restatement."""
import numpy as np
import pytest

from _util import P, i16p

pytestmark = pytest.mark.gpu


def test_tensor_core_satd_matches_oracle(cucd, oracle):
    rng = np.random.default_rng(5)
    n = 128 * 40
    org = rng.integers(0, 256, (n, 64)).astype(np.uint8)
    pred = rng.integers(0, 256, (n, 64)).astype(np.uint8)
    org[:128] = 255; pred[:128] = 0                      # extreme: DC only
    yy, xx = np.mgrid[0:8, 0:8]
    org[128:256] = np.where((xx + yy) & 1, 255, 0).ravel()  # extreme: highest frequency
    pred[128:256] = np.where((xx + yy) & 1, 0, 255).ravel()
    got, ms = cucd.exp_satd_tc(org, pred, iters=3)
    for k in list(range(0, 300)) + list(rng.integers(0, n, 300)):
        o = org[k].astype(np.int16).reshape(8, 8).copy()
        p = pred[k].astype(np.int16).reshape(8, 8).copy()
        want = oracle.oracle_satd(8, P(o, i16p), 8, P(p, i16p), 8, 8, 8)
        assert int(got[k]) == want, k
    print(f"tensor-core SATD: {n} tiles, {ms * 1e3:.1f} us per launch")
