#!/usr/bin/env python3
"""Known-answer vectors for the intra luma TU coding path: xIntraCodingTUBlock
(TEncSearch.cpp:1092-1387; luma and, for SURVEY 8f.4, the Cb/Cr blocks of estIntraPredChromaQT) = predIntraAng -> residual -> TComTrQuant::transformNxN -> invTransformNxN -> reconstruction ->
SSE, dumped from the REAL reference encoder's own call sites (hooks cucd_hook_tu_* of oracle/ref_shims/cucd_dump.h).

Three encoder runs on small seeded clips:
  tu_rdoq8.npz   default configuration (RDOQ=1): pins prediction, forward transform / transform skip (m_plTempCoeff),
                 de-quantisation + inverse transform + reconstruction + SSE from the levels RDOQ chose
  tu_hdq8.npz    --RDOQ=0 --RDOQTS=0 (SignHideFlag=1): additionally pins the plain quantiser xQuant + signBitHidingHDQ
  tu_hdq10.npz   the same at 10 bit with --SignHideFlag=0
Needs /root/reference (through oracle/_ref); not run on the GPU box.  Usage: python tests/golden/gen_golden_tu.py
"""
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import AI, ENC, run_encoder, synth_clip  # noqa: E402


def read_tu(path):
    data = open(path, "rb").read()
    off, recs = 0, []
    while off < len(data):
        hdr = struct.unpack_from("<16i", data, off); off += 64
        assert hdr[0] == 0x55545243
        n = hdr[4]
        r = dict(poc=hdr[1], x=hdr[2], y=hdr[3], n=n, mode=hdr[5], bd=hdr[6], ts=hdr[7], load=hdr[8], qp=hdr[9], intra=hdr[10],
                 sbh=hdr[11], rdoq=hdr[12], abs_sum=hdr[13], dist=hdr[14] & 0xFFFFFFFF, comp=hdr[15])
        r["border"] = np.frombuffer(data, np.int16, 4 * n + 1, off).copy(); off += 2 * (4 * n + 1)
        r["org"] = np.frombuffer(data, np.int16, n * n, off).copy(); off += 2 * n * n
        r["pred"] = np.frombuffer(data, np.int16, n * n, off).copy(); off += 2 * n * n
        r["coef"] = np.frombuffer(data, np.int32, n * n, off).copy(); off += 4 * n * n
        r["level"] = np.frombuffer(data, np.int32, n * n, off).copy(); off += 4 * n * n
        r["reco"] = np.frombuffer(data, np.int16, n * n, off).copy(); off += 2 * n * n
        recs.append(r)
    return recs


def pack(recs, quota, seed):
    rng = np.random.default_rng(seed)
    out = {}
    for (n, chroma), q in [((n, c), q // (2 if c else 1)) for n, q in quota.items() for c in (0, 1)]:
        for ts in (0, 1):
            cand = [r for r in recs if r["n"] == n and r["ts"] == ts and (r["comp"] > 0) == bool(chroma)]
            if not cand:
                continue
            # prefer records with non-zero levels (absSum > 0) 3:1, keep every mode represented
            order = rng.permutation(len(cand))
            nz = [i for i in order if cand[i]["abs_sum"] > 0]
            z = [i for i in order if cand[i]["abs_sum"] == 0]
            chosen = nz[: (3 * q) // 4] + z[: q // 4]
            seen_modes = {cand[i]["mode"] for i in chosen}
            for i in order:
                if cand[i]["mode"] not in seen_modes:
                    chosen.append(i); seen_modes.add(cand[i]["mode"])
            sel = [cand[i] for i in chosen]
            tag = ("c" if chroma else "n") + f"{n}" + ("ts" if ts else "")
            out[tag + "_hdr"] = np.array([[r[k] for k in ("poc", "x", "y", "mode", "bd", "ts", "load", "qp", "intra", "sbh", "rdoq", "abs_sum", "dist", "comp")]
                                          for r in sel], np.int64)
            for k in ("border", "org", "pred", "coef", "level", "reco"):
                out[tag + "_" + k] = np.stack([r[k] for r in sel])
    return out


def main():
    if not os.path.exists(ENC):
        sys.exit("oracle/_ref/TAppEncoder missing - run oracle/build_ref.sh (needs /root/reference)")
    W, H = 416, 240
    jobs = [("rdoq8", 8, 2, 27, [], 20261021, "37"),
            ("hdq8", 8, 2, 32, ["--RDOQ=0", "--RDOQTS=0"], 20261022, "37"),
            ("hdq10", 10, 1, 22, ["--RDOQ=0", "--RDOQTS=0", "--SignHideFlag=0"], 20261023, "29")]
    quota = {32: 30, 16: 60, 8: 120, 4: 160}
    for name, bd, frames, qp, extra, seed, every in jobs:
        with tempfile.TemporaryDirectory(prefix="cucd_gold_") as wd:
            open(os.path.join(wd, "clip.yuv"), "wb").write(synth_clip(W, H, frames, bd, seed))
            run_encoder(wd, W, H, frames, bd, qp, AI + extra, {"CUCD_DUMP_TU": "tu.bin", "CUCD_DUMP_TU_EVERY": every})
            recs = read_tu(os.path.join(wd, "tu.bin"))
            out = pack(recs, quota, seed)
            np.savez_compressed(os.path.join(HERE, f"tu_{name}.npz"), **out)
            print(f"{name}: {len(recs)} TUs dumped; kept", {k: v.shape[0] for k, v in out.items() if k.endswith("_hdr")})


if __name__ == "__main__":
    main()
