#!/usr/bin/env python3
"""Known-answer vectors for the fractional-pel ME refinement: sampled candidates of xPatternRefinement
(TEncSearch.cpp:808-865) dumped from the REAL reference encoder - source block, the reference window around the integer MV
(rows / columns -4 .. size+4), the candidate's quarter-pel offset and the distortion DistFunc returned (Hadamard with
--HadamardME=1, SAD with --HadamardME=0).  Low-delay P runs with AMP, so every PU shape occurs.
Needs /root/reference (through oracle/_ref); not run on the GPU box.  Usage: python tests/golden/gen_golden_frac.py"""
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import ENC, LDP, run_encoder, synth_clip  # noqa: E402


def read_frac(path):
    data = open(path, "rb").read()
    off, recs = 0, []
    while off < len(data):
        hdr = struct.unpack_from("<8i", data, off); off += 32
        assert hdr[0] == 0x52465243
        w, h = hdr[1], hdr[2]
        org = np.frombuffer(data, np.int16, w * h, off).copy(); off += 2 * w * h
        win = np.frombuffer(data, np.int16, (w + 9) * (h + 9), off).copy(); off += 2 * (w + 9) * (h + 9)
        recs.append((hdr[1:], org, win))
    return recs


def pack(recs, limit, seed):
    rng = np.random.default_rng(seed)
    by = {}
    for i, r in enumerate(recs):
        w, h, bd, had, qx, qy, _ = r[0]
        by.setdefault((w, h, had, qx & 3, qy & 3), []).append(i)
    chosen = []
    for k in sorted(by):                       # one of every (shape, fraction) class first
        chosen.append(int(rng.choice(by[k])))
    rest = [i for i in rng.permutation(len(recs)) if i not in set(chosen)]
    chosen = (chosen + [int(i) for i in rest])[:max(limit, 0)] if len(chosen) < limit else [int(i) for i in rng.permutation(chosen)[:limit]]
    out = {"hdr": np.array([recs[i][0] for i in chosen], np.int32)}     # w, h, bitDepth, hadamard, qx, qy, dist
    out["org"] = np.concatenate([recs[i][1] for i in chosen])
    out["win"] = np.concatenate([recs[i][2] for i in chosen])
    return out, len(by)


def main():
    if not os.path.exists(ENC):
        sys.exit("oracle/_ref/TAppEncoder missing - run oracle/build_ref.sh (needs /root/reference)")
    W, H = 416, 240
    for name, bd, extra, seed in (("had8", 8, [], 20261024), ("sad10", 10, ["--HadamardME=0"], 20261025)):
        cfg = [c for c in LDP if not (extra and c.startswith("--HadamardME"))] + extra
        with tempfile.TemporaryDirectory(prefix="cucd_gold_") as wd:
            open(os.path.join(wd, "clip.yuv"), "wb").write(synth_clip(W, H, 3, bd, seed))
            run_encoder(wd, W, H, 3, bd, 32, cfg, {"CUCD_DUMP_FRAC": "frac.bin", "CUCD_DUMP_FRAC_EVERY": "157"})
            recs = read_frac(os.path.join(wd, "frac.bin"))
            out, classes = pack(recs, 420, seed)
            np.savez_compressed(os.path.join(HERE, f"frac_{name}.npz"), **out)
            shapes = sorted({(int(h[0]), int(h[1])) for h in out["hdr"]})
            print(f"{name}: {len(recs)} candidates dumped ({classes} shape x fraction classes), kept {len(out['hdr'])}; shapes {shapes}")


if __name__ == "__main__":
    main()
