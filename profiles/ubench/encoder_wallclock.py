#!/usr/bin/env python3
"""Wall-clock of the integrated encoder (run on the GPU box): the CPU-only reference build, the integrated build calling the
library in-process (one synchronous call per PU), and K concurrent instances sharing the GPU through cucd_server.

  python profiles/ubench/encoder_wallclock.py [--width 1920 --height 1080 --qp 32 --instances 1,4,8,12] > gpurun_out/encoder_wallclock.json

Every integrated run is checked byte for byte against the CPU-only run of the same clip.  Prints one JSON document."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gen_golden as gg  # noqa: E402
import gen_golden_md5 as gm  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
SERVER = os.path.join(ROOT, "fast-cu-decision-hevc_b200", "cucd_server")


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def launch(binary, wd, W, H, frames, bd, qp, structure, env=None):
    e = dict(os.environ); e["CUCD_SHIM_TMV"] = "0"; e.update(env or {})      # no TMV verification calls: they are test instrumentation, not integration
    return subprocess.Popen(gm.encoder_args(os.path.join(REF, binary), W, H, frames, bd, qp, structure), cwd=wd, stdout=subprocess.DEVNULL,
                            stderr=subprocess.PIPE, text=True, env=e)


def run_set(binary, clips, W, H, frames, bd, qp, structure, env=None):
    """all clips concurrently, one process each; returns (wall seconds, [bitstream md5], [stderr tail])"""
    dirs = [tempfile.mkdtemp(prefix="cucd_wc_") for _ in clips]
    for d, c in zip(dirs, clips):
        open(os.path.join(d, "clip.yuv"), "wb").write(c)
    t0 = time.perf_counter()
    procs = [launch(binary, d, W, H, frames, bd, qp, structure, env) for d in dirs]
    errs = [p.communicate()[1] for p in procs]
    dt = time.perf_counter() - t0
    for p, e in zip(procs, errs):
        if p.returncode != 0:
            raise RuntimeError(f"{binary} failed: {e[-1000:]}")
    sums = [md5(os.path.join(d, "out.bin")) for d in dirs]
    for d in dirs:
        subprocess.run(["rm", "-rf", d])
    return dt, sums, [e.strip().splitlines()[-1] if e.strip() else "" for e in errs]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--bit-depth", type=int, default=8); ap.add_argument("--qp", type=int, default=32)
    ap.add_argument("--frames", type=int, default=1); ap.add_argument("--structure", default="AI")
    ap.add_argument("--instances", default="1,4,8,12"); ap.add_argument("--window-us", default="20")
    a = ap.parse_args()
    W, H, bd = a.width, a.height, a.bit_depth
    ks = [int(v) for v in a.instances.split(",")]
    clips = [gg.synth_clip(W, H, a.frames, bd, 20261200 + i) for i in range(max(ks))]
    ctus = a.frames * ((W + 63) // 64) * ((H + 63) // 64)
    out = {"clip": f"{W}x{H} {bd}-bit {a.structure} {a.frames} picture(s) QP {a.qp}", "host_threads": os.cpu_count(), "rows": []}
    ref_md5 = {}
    for k in ks:
        dt, sums, _ = run_set("TAppEncoder", clips[:k], W, H, a.frames, bd, a.qp, a.structure)
        for i, s in enumerate(sums):
            ref_md5[i] = s
        out["rows"].append({"build": "TAppEncoder (CPU only)", "instances": k, "wall_s": round(dt, 2), "ctus_per_s": round(k * ctus / dt, 1)})
    dt, sums, tails = run_set("TAppEncoderCucd", clips[:1], W, H, a.frames, bd, a.qp, a.structure)
    out["rows"].append({"build": "TAppEncoderCucd in-process (one synchronous library call per PU)", "instances": 1, "wall_s": round(dt, 2), "ctus_per_s": round(ctus / dt, 1),
                        "byte_identical": sums[0] == ref_md5[0], "shim": tails[0]})
    for min_n in (16, 32):
        dt, sums, tails = run_set("TAppEncoderCucd", clips[:1], W, H, a.frames, bd, a.qp, a.structure, env={"CUCD_SHIM_MIN_N": str(min_n)})
        out["rows"].append({"build": f"TAppEncoderCucd in-process, PUs smaller than {min_n} kept on the CPU (CUCD_SHIM_MIN_N={min_n})", "instances": 1, "wall_s": round(dt, 2),
                            "ctus_per_s": round(ctus / dt, 1), "byte_identical": sums[0] == ref_md5[0], "shim": tails[0]})
    for k in ks:
        name = f"/cucd_wc_{os.getpid()}_{k}"
        srv = subprocess.Popen([SERVER, "--name", name, "--clients", str(k), "--window-us", a.window_us], stdout=subprocess.PIPE, text=True)
        time.sleep(1.0)
        try:
            dt, sums, tails = run_set("TAppEncoderCucd", clips[:k], W, H, a.frames, bd, a.qp, a.structure, env={"CUCD_SERVER": name})
            stats = json.loads(srv.communicate(timeout=60)[0].strip().splitlines()[-1])
        finally:
            if srv.poll() is None:
                srv.terminate()
        out["rows"].append({"build": "TAppEncoderCucd through cucd_server", "instances": k, "wall_s": round(dt, 2), "ctus_per_s": round(k * ctus / dt, 1),
                            "byte_identical": all(sums[i] == ref_md5[i] for i in range(k)), "server": stats, "shim": tails[0]})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
