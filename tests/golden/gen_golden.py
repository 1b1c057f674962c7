#!/usr/bin/env python3
"""Generate the committed known-answer fixtures in tests/golden/ from the REAL reference encoder.

Runs oracle/_ref/TAppEncoder (the patched, hook-instrumented build of /root/reference made by
oracle/build_ref.sh) on small seeded synthetic clips with the KAT dump hooks of
oracle/ref_shims/cucd_dump.h switched on, then sub-samples the dumps into compact .npz files:

  rmd_<clip>.npz   per-PU records from the reference's rough-mode-decision loop
                   (TEncSearch.cpp:2252-2361): unfiltered + filtered border, source block, uiSad[35],
                   neighbour flags; plus the (x, y, N, flags) of EVERY PU the encoder visited
  obf_<clip>.npz   per-picture records of TEncSlice::getOutlierWithDCT (TEncSlice.cpp:878-1173):
                   source luma, Yc[1..15], OBF plane, Outlier plane; per-CU Num_OBF / N_Outlier
                   (TEncCu.cpp:589-600)
  me_<clip>.npz    sampled integer-ME SAD probes of xTZSearchHelp (TEncSearch.cpp:336-437)

This script needs /root/reference (through oracle/_ref) and is NOT run on the GPU box; the .npz
files are what travels.  Usage:  python tests/golden/gen_golden.py
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
ENC = os.path.join(ROOT, "oracle", "_ref", "TAppEncoder")

COMMON = ["--MaxCUWidth=64", "--MaxCUHeight=64", "--MaxPartitionDepth=4", "--QuadtreeTULog2MaxSize=5",
          "--QuadtreeTULog2MinSize=2", "--QuadtreeTUMaxDepthIntra=3", "--QuadtreeTUMaxDepthInter=3",
          "--SEIDecodedPictureHash=1", "--FEN=1", "--FDM=1", "--TransformSkip=1", "--TransformSkipFast=1",
          "--SAO=1", "-fr", "30"]
AI = ["--IntraPeriod=1", "--GOPSize=1"]
LDP = ["--IntraPeriod=-1", "--GOPSize=4", "--FastSearch=1", "--SearchRange=64", "--BipredSearchRange=4",
       "--HadamardME=1", "--AMP=1",
       "--Frame1=P 1 3 0.4624 0 0 0 4 4 -1 -5 -9 -13 0",
       "--Frame2=P 2 2 0.4624 0 0 0 4 4 -1 -2 -6 -10 1 -1 5 1 1 1 0 1",
       "--Frame3=P 3 3 0.4624 0 0 0 4 4 -1 -3 -7 -11 1 -1 5 0 1 1 1 1",
       "--Frame4=P 4 1 0.578 0 0 0 4 4 -1 -4 -8 -12 1 -1 5 0 1 1 1 1"]


def synth_clip(W, H, frames, bit_depth, seed):
    """Seeded 4:2:0 clip: blocky random texture that translates, flat and gradient regions (so the
    strong-smoothing test of TComPattern.cpp:201-214 fires), sinusoids and mild noise."""
    rng = np.random.default_rng(seed)
    field = rng.integers(0, 256, (H // 8 + 8, W // 8 + 8)).astype(np.float64)
    tex = np.kron(field, np.ones((8, 8)))
    out = []
    x = np.arange(W)[None, :]
    y = np.arange(H)[:, None]
    for t in range(frames):
        base = tex[8 + t: 8 + t + H, 8 + 2 * t: 8 + 2 * t + W]
        Y = 0.55 * base + 40 + 25 * np.sin((x + 3 * t) / 23.0) + 20 * np.cos((y - 2 * t) / 17.0)
        Y = Y + rng.uniform(-6, 6, (H, W))
        # flat / smooth-gradient panels
        Y[H // 2:, W // 2:] = 60 + 0.20 * x[:, W // 2:] + 0.15 * y[H // 2:, :] + rng.uniform(-0.6, 0.6, (H - H // 2, W - W // 2))
        Y[: H // 4, : W // 4] = 200
        Y = np.clip(np.rint(Y), 0, 255)
        xc = np.arange(W // 2)[None, :]
        yc = np.arange(H // 2)[:, None]
        Cb = np.clip(np.rint(128 + 30 * np.sin((xc + t) / 29.0) + 0 * yc), 0, 255)
        Cr = np.clip(np.rint(128 + 30 * np.cos((yc + t) / 23.0) + 0 * xc), 0, 255)
        if bit_depth > 8:
            sh = bit_depth - 8
            planes = [(p.astype(np.int64) << sh) + rng.integers(0, 1 << sh, p.shape) for p in (Y, Cb, Cr)]
            out.append(b"".join(p.astype("<u2").tobytes() for p in planes))
        else:
            out.append(b"".join(p.astype(np.uint8).tobytes() for p in (Y, Cb, Cr)))
    return b"".join(out)


def run_encoder(workdir, W, H, frames, bit_depth, qp, extra, env):
    args = [ENC, "-i", "clip.yuv", "-wdt", str(W), "-hgt", str(H), "-f", str(frames), "-q", str(qp),
            "-b", "out.bin", "-o", "rec.yuv", f"--InputBitDepth={bit_depth}", f"--InternalBitDepth={bit_depth}",
            "--Profile=" + ("main10" if bit_depth > 8 else "main")] + COMMON + extra
    e = dict(os.environ)
    e.update(env)
    with open(os.path.join(workdir, "enc.log"), "w") as log:
        subprocess.run(args, cwd=workdir, env=e, stdout=log, stderr=subprocess.STDOUT, check=True)


def read_rmd(path):
    data = open(path, "rb").read()
    off, recs = 0, []
    while off < len(data):
        hdr = struct.unpack_from("<8i", data, off)
        off += 32
        assert hdr[0] == 0x444D5243
        poc, x, y, n, bd, nf, _ = hdr[1:]
        flags = np.frombuffer(data, np.uint8, nf, off).copy(); off += nf
        unf = np.frombuffer(data, np.int16, 4 * n + 1, off).copy(); off += 2 * (4 * n + 1)
        fil = np.frombuffer(data, np.int16, 4 * n + 1, off).copy(); off += 2 * (4 * n + 1)
        org = np.frombuffer(data, np.int16, n * n, off).copy(); off += 2 * n * n
        sad = np.frombuffer(data, np.uint32, 35, off).copy(); off += 140
        recs.append(dict(poc=poc, x=x, y=y, n=n, bd=bd, flags=flags, unf=unf, fil=fil, org=org, sad=sad))
    return recs


def plain_121(unf):
    f = unf.astype(np.int32).copy()
    f[1:-1] = (unf[:-2].astype(np.int32) + 2 * unf[1:-1] + unf[2:] + 2) >> 2
    return f.astype(np.int16)


def pack_rmd(recs, W, H, quota, seed):
    rng = np.random.default_rng(seed)
    out = {}
    # every visited PU's neighbour flags (deduplicated) - pins the availability rule
    seen = {}
    for r in recs:
        seen[(r["x"], r["y"], r["n"])] = r["flags"]
    keys = sorted(seen)
    out["vis_xyn"] = np.array(keys, np.int16)
    out["vis_flags"] = np.concatenate([seen[k] for k in keys]).astype(np.uint8)
    for n, q in quota.items():
        cand = [r for r in recs if r["n"] == n]
        if not cand:
            continue
        chosen, pats = [], set()
        order = rng.permutation(len(cand))
        # strong-smoothing hits first, then one per distinct flag pattern, then random fill
        strong = [i for i in order if n == 32 and not np.array_equal(cand[i]["fil"], plain_121(cand[i]["unf"]))]
        for i in strong[: q // 4]:
            chosen.append(i)
        for i in order:
            p = cand[i]["flags"].tobytes()
            if p not in pats and len(chosen) < q:
                pats.add(p)
                if i not in chosen:
                    chosen.append(i)
        for i in order:
            if len(chosen) >= q:
                break
            if i not in chosen:
                chosen.append(i)
        sel = [cand[i] for i in chosen]
        out[f"n{n}_xy"] = np.array([[r["poc"], r["x"], r["y"]] for r in sel], np.int16)
        out[f"n{n}_flags"] = np.stack([r["flags"] for r in sel])
        out[f"n{n}_unf"] = np.stack([r["unf"] for r in sel])
        out[f"n{n}_fil"] = np.stack([r["fil"] for r in sel])
        out[f"n{n}_org"] = np.stack([r["org"] for r in sel])
        out[f"n{n}_sad"] = np.stack([r["sad"] for r in sel])
    out["meta"] = np.array([W, H, recs[0]["bd"]], np.int32)
    return out


def read_obf(workdir, max_frames):
    data = open(os.path.join(workdir, "obf.bin"), "rb").read()
    yd = open(os.path.join(workdir, "yc.bin"), "rb").read()
    cu = np.frombuffer(open(os.path.join(workdir, "cu.bin"), "rb").read(), np.int32).reshape(-1, 7)
    off = yoff = 0
    out = {}
    k = 0
    while off < len(data) and k < max_frames:
        _, poc, w, h, bd = struct.unpack_from("<5i", data, off); off += 20
        org = np.frombuffer(data, np.int16, w * h, off).reshape(h, w).copy(); off += 2 * w * h
        obf = np.frombuffer(data, np.int16, (w // 4) * (h // 4), off).reshape(h // 4, w // 4).copy(); off += 2 * (w // 4) * (h // 4)
        outl = np.frombuffer(data, np.int16, w * h, off).reshape(h, w).copy(); off += 2 * w * h
        (n,) = struct.unpack_from("<i", yd, yoff); yoff += 4
        yc = np.frombuffer(yd, np.float64, n, yoff).copy(); yoff += 8 * n
        out[f"f{k}_org"] = org
        out[f"f{k}_obf"] = obf
        out[f"f{k}_outlier"] = outl
        out[f"f{k}_yc"] = yc
        out[f"f{k}_cu"] = cu[cu[:, 0] == poc][:, 1:].copy()  # depth, x, y, size, Num_OBF, N_Outlier
        out[f"f{k}_meta"] = np.array([poc, w, h, bd], np.int32)
        k += 1
    out["nframes"] = np.array([k], np.int32)
    # the un-instrumented byproduct files the fork always writes (8-bit truncated planes)
    obf_yuv = np.fromfile(os.path.join(workdir, "OBF.yuv"), np.uint8)
    out["obf_yuv_frame0"] = obf_yuv[: (w // 4) * (h // 4)].reshape(h // 4, w // 4).copy()
    return out


def read_me(path, limit, seed):
    data = open(path, "rb").read()
    off, recs = 0, []
    while off < len(data):
        hdr = struct.unpack_from("<8i", data, off); off += 32
        assert hdr[0] == 0x454D5243
        cols, rows = hdr[1], hdr[2]
        org = np.frombuffer(data, np.int16, cols * rows, off).copy(); off += 2 * cols * rows
        ref = np.frombuffer(data, np.int16, cols * rows, off).copy(); off += 2 * cols * rows
        recs.append((hdr[1:], org, ref))
    rng = np.random.default_rng(seed)
    # keep every (cols, rows, subShift) shape class represented
    by = {}
    for i, r in enumerate(recs):
        by.setdefault(tuple(r[0][:3]), []).append(i)
    chosen = []
    per = max(4, limit // max(1, len(by)))
    for k in sorted(by):
        idx = rng.permutation(by[k])[:per]
        chosen.extend(int(i) for i in idx)
    chosen = chosen[:limit]
    out = {"hdr": np.array([recs[i][0] for i in chosen], np.int32)}  # cols, rows, subShift, bitDepth, mvx, mvy, sad
    out["org"] = np.concatenate([recs[i][1] for i in chosen])
    out["ref"] = np.concatenate([recs[i][2] for i in chosen])
    return out, len(recs), sorted(by)


def main():
    if not os.path.exists(ENC):
        sys.exit("oracle/_ref/TAppEncoder missing - run oracle/build_ref.sh (needs /root/reference)")
    W, H = 416, 240
    jobs = [
        ("ai8", 8, 4, 32, AI, 20261018),
        ("ai10", 10, 2, 27, AI, 20261019),
    ]
    for name, bd, frames, qp, cfg, seed in jobs:
        with tempfile.TemporaryDirectory(prefix="cucd_gold_") as wd:
            open(os.path.join(wd, "clip.yuv"), "wb").write(synth_clip(W, H, frames, bd, seed))
            run_encoder(wd, W, H, frames, bd, qp, cfg,
                        {"CUCD_DUMP_RMD": "rmd.bin", "CUCD_DUMP_OBF": "obf.bin", "CUCD_DUMP_OBF_YC": "yc.bin", "CUCD_DUMP_CU": "cu.bin"})
            recs = read_rmd(os.path.join(wd, "rmd.bin"))
            quota = {64: 24, 32: 64, 16: 120, 8: 200, 4: 320}
            np.savez_compressed(os.path.join(HERE, f"rmd_{name}.npz"), **pack_rmd(recs, W, H, quota, seed))
            np.savez_compressed(os.path.join(HERE, f"obf_{name}.npz"), **read_obf(wd, 2))
            md5 = subprocess.run(["md5sum", os.path.join(wd, "out.bin")], capture_output=True, text=True).stdout.split()[0]
            print(f"{name}: {len(recs)} RMD PUs dumped, bitstream md5 {md5}")
    with tempfile.TemporaryDirectory(prefix="cucd_gold_") as wd:
        open(os.path.join(wd, "clip.yuv"), "wb").write(synth_clip(W, H, 3, 8, 20261020))
        run_encoder(wd, W, H, 3, 8, 32, LDP, {"CUCD_DUMP_ME": "me.bin", "CUCD_DUMP_ME_EVERY": "53"})
        me, total, shapes = read_me(os.path.join(wd, "me.bin"), 260, 7)
        np.savez_compressed(os.path.join(HERE, "me_ldp8.npz"), **me)
        print(f"ldp8: {total} sampled ME probes, kept {len(me['hdr'])}, shapes {shapes}")


if __name__ == "__main__":
    main()
