"""Known-answer vectors for the CU texture features (a12, getTMVFeature) and the AQ activity (a13, TEncPreanalyzer),
made by calling the REFERENCE's own compiled functions through oracle/_ref/libhmref.so (oracle/build_ref.sh) on small
seeded planes.  Run in the build container (needs /root/reference to have been built):  python tests/golden/gen_golden_texture.py
Writes tests/golden/texture.npz (a few tens of KB)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _util  # noqa: E402
from _util import P, i16p, f64p  # noqa: E402


def main():
    ref = _util.load_hmref()
    assert ref is not None, "oracle/_ref/libhmref.so missing: run oracle/build_ref.sh"
    out = {}
    for tag, bd, W, H in (("a8", 8, 136, 72), ("b10", 10, 72, 136)):
        org = _util.textured_plane(W, H, bd, seed=41)
        org[:16, :16] = np.random.default_rng(9).choice([0, (1 << bd) - 1], size=(16, 16))     # extreme corner
        cus = np.array(_util.all_cus(W, H), np.int32)
        feat = np.zeros((len(cus), 5, 26))
        for i, (x, y, l) in enumerate(cus):
            ref.hmref_tmv_features(P(org, i16p), W, W, H, int(x), int(y), 1 << int(l), C.c_void_p(feat[i].ctypes.data))
        acts = [np.zeros(((H + (64 >> d) - 1) // (64 >> d), (W + (64 >> d) - 1) // (64 >> d))) for d in range(4)]
        ptrs = (C.c_void_p * 4)(*[a.ctypes.data for a in acts])
        avg = np.zeros(4)
        ref.hmref_aq_activity(P(org, i16p), W, W, H, 4, ptrs, P(avg, f64p))
        out[f"{tag}_meta"] = np.array([W, H, bd], np.int32)
        out[f"{tag}_org"] = org
        out[f"{tag}_cus"] = cus
        out[f"{tag}_tmv"] = feat
        for d in range(4):
            out[f"{tag}_act{d}"] = acts[d]
        out[f"{tag}_avg"] = avg
    np.savez_compressed(os.path.join(HERE, "texture.npz"), **out)
    print("wrote texture.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
