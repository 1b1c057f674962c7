#!/usr/bin/env python3
"""Reference-encoder answers for the BASELINE configurations AT THEIR STATED SIZES (BASELINE.json configs[1..4]).

Runs the CPU-only reference build oracle/_ref/TAppEncoder on seeded synthetic clips of 1920x1080 / 3840x2160 and records, per
case, the MD5 of the bitstream and of the reconstruction, the bitstream size and the encoder's wall-clock seconds in
tests/golden/encoder_md5.json.  tests/test_gpu_encoder_md5.py then runs ONLY the integrated build (TAppEncoderCucd, every cost
from the GPU) on the GPU box and compares with these answers: a full-size CPU encode takes minutes and needs no GPU, so it is
done once here, in the build container.

Needs /root/reference (through oracle/_ref); NOT run on the GPU box.  Usage:
    python tests/golden/gen_golden_md5.py [case ...]        (no argument = every case; cases run in parallel)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402

OUT = os.path.join(HERE, "encoder_md5.json")

# HM random-access GOP8 (B pictures, two reference lists), closed intra periods (DecodingRefreshType=2: BASELINE configs[4])
RA = ["--IntraPeriod=16", "--GOPSize=8", "--DecodingRefreshType=2", "--FastSearch=1", "--SearchRange=64", "--BipredSearchRange=4", "--HadamardME=1", "--AMP=1",
      "--Frame1=B 8 1 0.442 0 0 0 4 4 -8 -10 -12 -16 0",
      "--Frame2=B 4 2 0.3536 0 0 0 2 3 -4 -6 4 1 4 5 1 1 0 0 1",
      "--Frame3=B 2 3 0.3536 0 0 0 2 4 -2 -4 2 6 1 2 4 1 1 1 1",
      "--Frame4=B 1 4 0.68 0 0 0 2 4 -1 1 3 7 1 1 5 1 0 1 1 1",
      "--Frame5=B 3 4 0.68 0 0 0 2 4 -1 -3 1 5 1 -2 5 1 1 1 1 0",
      "--Frame6=B 6 3 0.3536 0 0 0 2 4 -2 -4 -6 2 1 -3 5 1 1 1 1 0",
      "--Frame7=B 5 4 0.68 0 0 0 2 4 -1 -5 1 3 1 1 5 1 0 1 1 1",
      "--Frame8=B 7 4 0.68 0 0 0 2 4 -1 -3 -7 1 1 -2 5 1 1 1 1 0"]

# name -> (W, H, bit depth, pictures, QP, coding structure, clip seed)
# 1080p: 17 CTU rows, the last one 56 samples high (forced splits down to 8x8, TEncCu.cpp:488-489, 648);
# LDP / RA: TZ search range 64 at real picture size (TEncSearch.cpp:3865-3881).
CASES = {
    "ai1080p8_q22": (1920, 1080, 8, 1, 22, "AI", 20261101),
    "ai1080p8_q27": (1920, 1080, 8, 1, 27, "AI", 20261105),
    "ai1080p8_q37": (1920, 1080, 8, 2, 37, "AI", 20261102),
    "ldp1080p8_q32": (1920, 1080, 8, 2, 32, "LDP", 20261103),
    "ra1080p10_q32": (1920, 1080, 10, 3, 32, "RA", 20261104),
    "ra1080p10_gop8_q32": (1920, 1080, 10, 9, 32, "RA", 20261104),
    "ai2160p10_q32": (3840, 2160, 10, 1, 32, "AI", 20261106),
}
STRUCT = {"AI": gg.AI, "LDP": gg.LDP, "RA": RA}


def encoder_args(binary, W, H, frames, bd, qp, structure):
    return [binary, "-i", "clip.yuv", "-wdt", str(W), "-hgt", str(H), "-f", str(frames), "-q", str(qp), "-b", "out.bin", "-o", "rec.yuv",
            f"--InputBitDepth={bd}", f"--InternalBitDepth={bd}", "--Profile=" + ("main10" if bd > 8 else "main")] + gg.COMMON + STRUCT[structure]


def file_md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def run_case(name):
    W, H, bd, frames, qp, structure, seed = CASES[name]
    with tempfile.TemporaryDirectory(prefix="cucd_md5gold_") as wd:
        open(os.path.join(wd, "clip.yuv"), "wb").write(gg.synth_clip(W, H, frames, bd, seed))
        t0 = time.perf_counter()
        r = subprocess.run(encoder_args(gg.ENC, W, H, frames, bd, qp, structure), cwd=wd, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res = {"width": W, "height": H, "bit_depth": bd, "frames": frames, "qp": qp, "structure": structure, "seed": seed,
               "bitstream_md5": file_md5(os.path.join(wd, "out.bin")), "rec_md5": file_md5(os.path.join(wd, "rec.yuv")),
               "bitstream_bytes": os.path.getsize(os.path.join(wd, "out.bin")), "cpu_encoder_seconds_build_container": round(dt, 1)}
    print(name, json.dumps(res), flush=True)
    return name, res


def main():
    if not os.path.exists(gg.ENC):
        sys.exit("oracle/_ref/TAppEncoder missing - run oracle/build_ref.sh (needs /root/reference)")
    names = sys.argv[1:] or list(CASES)
    done = json.load(open(OUT)) if os.path.exists(OUT) else {}
    with ThreadPoolExecutor(max_workers=min(len(names), max(1, (os.cpu_count() or 2) - 2))) as ex:
        for name, res in ex.map(run_case, names):
            done[name] = res
            json.dump(done, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
