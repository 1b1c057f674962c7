"""End-to-end drop-in proof (north_star: "the produced bitstream and the TAppDecoder MD5 of the reconstruction match the
reference encoder's byte for byte"): oracle/_ref/TAppEncoderCucd is the reference encoder with exactly the lines of
INTEGRATION.md S0/S1/S2 added (oracle/ref_shims/cucd_dump.h, -DCUCD_INTEGRATION) and linked against libcucudecide.so, so its
per-picture OBF/Outlier features and every rough-mode-decision SATD come from the GPU.  Its bitstream and reconstruction must
equal those of the CPU-only reference build, and the reference decoder must verify every picture hash.
Both binaries are built in the build container (oracle/build_ref.sh) and travel in oracle/_ref/."""
import hashlib
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


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


def _encode(binary, wd, W, H, frames, bd, qp, cfg, env=None):
    import gen_golden as gg
    args = [os.path.join(REF, binary), "-i", "clip.yuv", "-wdt", str(W), "-hgt", str(H), "-f", str(frames), "-q", str(qp), "-b", "out.bin",
            "-o", "rec.yuv", f"--InputBitDepth={bd}", f"--InternalBitDepth={bd}", "--Profile=" + ("main10" if bd > 8 else "main")] + gg.COMMON + cfg
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(args, cwd=wd, capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


@pytest.mark.parametrize("cfg,bd,frames,qp", [("AI", 8, 5, 32), ("AI", 10, 2, 27), ("LDP", 8, 3, 32), ("RA", 10, 9, 32), ("AITU", 8, 1, 32), ("AITU", 10, 1, 27)])
def test_bitstream_md5_matches_reference_encoder(cfg, bd, frames, qp):
    """AI: S1 + S2 on the GPU (5 frames reach the fork's Testing state, so the OBF-driven early decisions are live);
    LDP (I + 2 P pictures, TZ search range 64, AMP): additionally every integer-ME SAD of the uni-directional searches (S3) and
    every candidate of their half-/quarter-pel refinement (8f.3).
    RA (I + one GOP8 of B pictures, 10 bit): the same with two reference lists, including the bi-predictive refinement
    (search key 2 * org - other prediction through cucd_me_sad_surface_src / cucd_me_subpel_cost_src).
    AITU (CUCD_SHIM_TU=1): additionally every luma and chroma TU of xIntraCodingTUBlock goes through cucd_intra_tu_forward (its
    prediction replaces the encoder's, its transform output must equal m_plTempCoeff), the host's RDOQ, and cucd_intra_tu_recon
    (its reconstruction replaces the encoder's samples, its SSE must equal getDistPart)."""
    import gen_golden as gg
    for b in ("TAppEncoder", "TAppEncoderCucd", "TAppDecoder"):
        if not os.path.exists(os.path.join(REF, b)):
            pytest.skip(f"oracle/_ref/{b} not built (needs /root/reference in the build container)")
    W, H = 416, 240
    clip = gg.synth_clip(W, H, frames, bd, 20261030 + bd)
    out = {}
    for binary in ("TAppEncoder", "TAppEncoderCucd"):
        with tempfile.TemporaryDirectory(prefix="cucd_md5_") as wd:
            open(os.path.join(wd, "clip.yuv"), "wb").write(clip)
            r = _encode(binary, wd, W, H, frames, bd, qp, {"AI": gg.AI, "AITU": gg.AI, "LDP": gg.LDP, "RA": RA}[cfg],
                        env={"CUCD_SHIM_TU": "1"} if cfg == "AITU" else None)
            bits = open(os.path.join(wd, "out.bin"), "rb").read()
            rec = open(os.path.join(wd, "rec.yuv"), "rb").read()
            out[binary] = (hashlib.md5(bits).hexdigest(), hashlib.md5(rec).hexdigest(), len(bits))
            if binary == "TAppEncoderCucd":
                assert "RMD PUs on the GPU" in r.stderr, r.stderr[-500:]          # the GPU path really ran
                n_gpu = int(r.stderr.split("pictures,")[1].split("RMD PUs")[0])
                assert n_gpu > 1000
                if cfg in ("LDP", "RA"):
                    n_me = int(r.stderr.split("GPU,")[1].split("ME searches")[0])
                    assert n_me > 1000
                    import re
                    assert int(re.search(r"(\d+) sub-pel refinements on the GPU", r.stderr).group(1)) > 1000
                    n_bi = int(re.search(r"(\d+) of them bi-predictive", r.stderr).group(1))
                    assert (n_bi > 100) if cfg == "RA" else (n_bi == 0)       # B pictures: the bi-predictive refinement runs on the GPU too
                if cfg == "AITU":
                    import re
                    assert int(re.search(r"(\d+) TUs coded on the GPU", r.stderr).group(1)) > 50000
                if cfg == "AI" and bd == 8:      # Training pictures dump TMV features: every set was recomputed on the GPU and compared
                    import re
                    assert int(re.search(r"(\d+) TMV feature sets verified", r.stderr).group(1)) > 100
                print(r.stderr.strip().splitlines()[-1])
                d = subprocess.run([os.path.join(REF, "TAppDecoder"), "-b", "out.bin", "-o", "dec.yuv", "-d", "0"],
                                   cwd=wd, capture_output=True, text=True, timeout=300)
                assert d.returncode == 0
                assert d.stdout.count("(OK)") == frames and "ERROR" not in d.stdout.upper().replace("(OK)", "")
                assert hashlib.md5(open(os.path.join(wd, "dec.yuv"), "rb").read()).hexdigest() == out[binary][1]
    assert out["TAppEncoder"] == out["TAppEncoderCucd"], out


# ---- the BASELINE configurations at their stated sizes (1920x1080 / 3840x2160) --------------------------------------------------------
# The CPU-only reference encoder was run once in the build container (tests/golden/gen_golden_md5.py -> encoder_md5.json); here only
# the integrated build runs, on the GPU box, and must reproduce the recorded bitstream and reconstruction byte for byte.
def _fullsize_cases():
    import json
    path = os.path.join(ROOT, "tests", "golden", "encoder_md5.json")
    return sorted(json.load(open(path)).items()) if os.path.exists(path) else []


@pytest.mark.parametrize("name,ref", _fullsize_cases(), ids=[n for n, _ in _fullsize_cases()])
def test_fullsize_bitstream_md5_matches_recorded_reference(name, ref):
    """1080p: 17 CTU rows, the last 56 samples high -> forced splits down to 8x8 (TEncCu.cpp:488-489, 648) go through the live
    encoder; LDP / RA: the full +-64 TZ window at real picture size (TEncSearch.cpp:3865-3881); QP 22 / 27 / 37; Main10 at 2160p."""
    import time
    import gen_golden as gg
    import gen_golden_md5 as gm
    if "gop8" in name and not os.environ.get("CUCD_LONG_TESTS"):
        # nine 1080p Main10 pictures with B-picture GOP8: ~20 M single-PU requests, 18 minutes on the GPU box.  Last run (round 2, every
        # hook on, profiles/r02_summary.md): byte-identical.  CUCD_LONG_TESTS=1 runs it.
        pytest.skip("long case: set CUCD_LONG_TESTS=1")
    for b in ("TAppEncoderCucd", "TAppDecoder"):
        if not os.path.exists(os.path.join(REF, b)):
            pytest.skip(f"oracle/_ref/{b} not built (needs /root/reference in the build container)")
    W, H, bd, frames, qp = ref["width"], ref["height"], ref["bit_depth"], ref["frames"], ref["qp"]
    with tempfile.TemporaryDirectory(prefix="cucd_md5_") as wd:
        open(os.path.join(wd, "clip.yuv"), "wb").write(gg.synth_clip(W, H, frames, bd, ref["seed"]))
        t0 = time.perf_counter()
        # full-size random access: the AMVP / merge candidate distortions (one 35 us request each, ~1 M per picture) stay on the CPU to
        # bound the suite's run time; the 416x240 RA case above and the long gop8 case run them on the GPU
        env = dict(os.environ, CUCD_SHIM_MC="0") if (ref["structure"] == "RA" and "gop8" not in name) else None
        r = subprocess.run(gm.encoder_args(os.path.join(REF, "TAppEncoderCucd"), W, H, frames, bd, qp, ref["structure"]), cwd=wd, capture_output=True,
                           text=True, timeout=2400, env=env)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "RMD PUs on the GPU" in r.stderr, r.stderr[-500:]
        got = (gm.file_md5(os.path.join(wd, "out.bin")), gm.file_md5(os.path.join(wd, "rec.yuv")), os.path.getsize(os.path.join(wd, "out.bin")))
        print(f"{name}: TAppEncoderCucd {dt:.1f} s (CPU-only reference in the build container: {ref['cpu_encoder_seconds_build_container']} s); " + r.stderr.strip().splitlines()[-1])
        assert got == (ref["bitstream_md5"], ref["rec_md5"], ref["bitstream_bytes"]), (name, got)
        if ref["structure"] != "AI":
            assert int(r.stderr.split("GPU,")[1].split("ME searches")[0]) > 10000
        d = subprocess.run([os.path.join(REF, "TAppDecoder"), "-b", "out.bin", "-o", "dec.yuv", "-d", "0"], cwd=wd, capture_output=True, text=True, timeout=600)
        assert d.returncode == 0 and d.stdout.count("(OK)") == frames
        assert gm.file_md5(os.path.join(wd, "dec.yuv")) == ref["rec_md5"]


def test_three_encoder_instances_through_cucd_server_stay_byte_identical():
    """three encoder PROCESSES (All-Intra, low-delay P and 10-bit All-Intra clips) share the GPU through cucd_server
    (include/cucd_ipc.h): their S2 requests are coalesced into common batches, S1 / S3 / sub-pel requests are served per instance.
    Every bitstream must equal the CPU-only reference encoder's."""
    import json
    import time
    import gen_golden as gg
    server = os.path.join(ROOT, "fast-cu-decision-hevc_b200", "cucd_server")
    for b in ("TAppEncoder", "TAppEncoderCucd"):
        if not os.path.exists(os.path.join(REF, b)):
            pytest.skip(f"oracle/_ref/{b} not built (needs /root/reference in the build container)")
    assert os.path.exists(server), "cucd_server not built"
    W, H = 416, 240
    jobs = [("AI", 8, 3, 32, gg.AI), ("LDP", 8, 3, 32, gg.LDP), ("AI", 10, 2, 27, gg.AI)]
    name = f"/cucd_test_{os.getpid()}"
    srv = subprocess.Popen([server, "--name", name, "--clients", str(len(jobs))], stdout=subprocess.PIPE, text=True)
    time.sleep(1.0)
    dirs, procs = [], []
    try:
        for i, (cfg, bd, frames, qp, struct) in enumerate(jobs):
            wd = tempfile.mkdtemp(prefix="cucd_srv_")
            dirs.append(wd)
            open(os.path.join(wd, "clip.yuv"), "wb").write(gg.synth_clip(W, H, frames, bd, 20261300 + i))
            args = [os.path.join(REF, "TAppEncoderCucd"), "-i", "clip.yuv", "-wdt", str(W), "-hgt", str(H), "-f", str(frames), "-q", str(qp), "-b", "out.bin", "-o", "rec.yuv",
                    f"--InputBitDepth={bd}", f"--InternalBitDepth={bd}", "--Profile=" + ("main10" if bd > 8 else "main")] + gg.COMMON + struct
            procs.append(subprocess.Popen(args, cwd=wd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=dict(os.environ, CUCD_SERVER=name)))
        errs = [p.communicate(timeout=900)[1] for p in procs]
        for p, e in zip(procs, errs):
            assert p.returncode == 0, e[-1500:]
            assert "through cucd_server" in e
        stats = json.loads(srv.communicate(timeout=60)[0].strip().splitlines()[-1])
        assert stats["closed"] == len(jobs) and stats["rmd_requests"] > 3000 and stats["me_surfaces"] > 100 and stats["subpel"] > 100 and stats["frames"] == 8
        assert stats["rmd_batches"] < stats["rmd_requests"]            # requests of different instances really shared batches
        print(stats)
        for i, (cfg, bd, frames, qp, struct) in enumerate(jobs):
            with tempfile.TemporaryDirectory(prefix="cucd_md5_") as wd:
                open(os.path.join(wd, "clip.yuv"), "wb").write(gg.synth_clip(W, H, frames, bd, 20261300 + i))
                _encode("TAppEncoder", wd, W, H, frames, bd, qp, struct)
                assert open(os.path.join(wd, "out.bin"), "rb").read() == open(os.path.join(dirs[i], "out.bin"), "rb").read(), (cfg, bd)
                assert hashlib.md5(open(os.path.join(wd, "rec.yuv"), "rb").read()).hexdigest() == hashlib.md5(open(os.path.join(dirs[i], "rec.yuv"), "rb").read()).hexdigest()
    finally:
        if srv.poll() is None:
            srv.terminate()
        for d in dirs:
            subprocess.run(["rm", "-rf", d])
