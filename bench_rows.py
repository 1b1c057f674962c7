#!/usr/bin/env python3
"""bench_rows.py - secondary measurements: the SURVEY.md section-8 rows that bench.py's headline line does not cover.

bench.py times the frame path (a1-a5, a8-a11).  This script times the other operators of include/cucudecide.h on one B200,
each on a 1080p-sized workload, and prints ONE JSON line per row:
  kernel   throughput from the device time between the first and last kernel of the call (cucd_last_kernel_time: CUDA events
           on the library's stream, copies outside) and the HBM roofline fraction of that time for the row's ALGORITHMIC bytes
           (every input read once, every output written once; DESIGN.md section 4 states the per-unit figures)
  e2e      the same C-ABI call with host buffers, wall clock around the bare call (descriptor arrays prebuilt, copies inside): the caller's
           buffers page-locked once through cucd_pin_host_buffer (what an encoder does with its long-lived buffers); e2e_pageable: the same
           buffers before they were pinned.  h2d_bytes / d2h_bytes let the reader subtract the PCIe time
  cpu      the reference's own function (oracle/_ref/libhmref.so, kind "reference") or the oracle port (kind "port") on a
           bounded sample, on `cores` host threads
Rows: s2 (cucd_intra_rmd_batch), f1 (cucd_queue_*: K concurrent instances), s3 (cucd_me_sad_surface), f3 (cucd_me_subpel_cost),
a12 (cucd_tmv_features), a13 (cucd_aq_activity), f2 (cucd_intra_tu_code / _forward / _recon).  Usage: python bench_rows.py [--iters 5] > gpurun_out/rows.jsonl
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util  # noqa: E402
from _util import P, i16p, i32p, u32p, f64p  # noqa: E402

W, H = 1920, 1080


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 7700.0, "fallback (B200_PROFILING.md nominal)"


def timed(fn, eng, iters, pin=()):
    """returns (kernel s, call s with the buffers in `pin` page-locked, call s with pageable buffers)"""
    fn()                                     # warm-up (allocations inside the handle)
    fn()
    ps = []
    for _ in range(max(2, iters // 2)):
        t0 = time.perf_counter()
        fn()
        ps.append(time.perf_counter() - t0)
    for a in pin:
        eng.pin_host_buffer(a)
    fn()
    ks, ws = [], []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        ws.append(time.perf_counter() - t0)
        ks.append(eng.last_kernel_time_ms() * 1e-3)
    for a in pin:
        eng.unpin_host_buffer(a)
    return float(np.median(ks)), float(np.median(ws)), float(np.median(ps))


def cpu_parallel(work, n_items, threads):
    """run work(i) for i in range(n_items) on `threads` host threads (ctypes releases the GIL); returns seconds"""
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(n_items)))
    return time.perf_counter() - t0


def emit(row, unit, units, k_s, w_s, algo_bytes, cpu, extra=None, p_s=None, h2d=None, d2h=None):
    peak, src = hbm_peak()
    gbs = algo_bytes / k_s / 1e9
    d = {"row": row, "unit": unit, "units_per_call": units,
         "kernel": {"value": units / k_s, "ms": k_s * 1e3, "algorithmic_bytes": algo_bytes, "achieved_gbs": gbs, "hbm_peak_gbs": peak,
                    "frac": gbs / peak, "peak_source": src},
         "e2e": {"value": units / w_s, "ms": w_s * 1e3, "call_over_kernel": w_s / k_s, "h2d_bytes": h2d, "d2h_bytes": d2h},
         "e2e_pageable": None if p_s is None else {"value": units / p_s, "ms": p_s * 1e3},
         "cpu": cpu}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu-threads", type=int, default=min(16, os.cpu_count() or 1))
    a = ap.parse_args()
    cucd = importlib.import_module("fast-cu-decision-hevc_b200")
    oracle = _util.load_oracle()
    ref = _util.load_hmref()
    T = a.cpu_threads
    org = _util.textured_plane(W, H, 8, seed=5)
    rec = _util.pseudo_recon(org, 8)
    rng = np.random.default_rng(1)
    eng = cucd.Engine(W, H, bit_depth=8)

    # ---- S2: batched RMD with caller-supplied borders: the 341 PUs of 510 CTUs' worth of blocks (mixed sizes) -------------
    sizes, orgs, brds = [], [], []
    for n, cnt in ((64, 480), (32, 1980), (16, 8040), (8, 32400), (4, 129600)):
        cnt = cnt // 4                      # a quarter picture per call keeps the host-side packing small
        sizes += [int(np.log2(n))] * cnt
        orgs.append(rng.integers(0, 256, cnt * n * n).astype(np.int16))
        brds.append(rng.integers(0, 256, cnt * (4 * n + 1)).astype(np.int16))
    o, b = np.concatenate(orgs), np.concatenate(brds)
    n_pu = len(sizes)
    pu_arr, _ = cucd.Engine.pu_descs(sizes)
    sad_out = np.zeros((n_pu, 35), np.uint32)
    k_s, w_s, p_s = timed(lambda: eng.intra_rmd_batch_raw(pu_arr, n_pu, o, b, sad_out), eng, a.iters, pin=(o, b, sad_out))
    algo = o.nbytes + b.nbytes + n_pu * 35 * 4
    # CPU: the reference's own prediction + Hadamard functions per PU
    samp = list(range(0, n_pu, 97))
    off_o = np.cumsum([0] + [1 << (2 * s) for s in sizes]); off_b = np.cumsum([0] + [(4 << s) + 1 for s in sizes])
    lib, kind = (ref, "reference") if ref is not None else (oracle, "port")
    fn = lib.hmref_rmd_pu if ref is not None else lib.oracle_rmd_pu

    def work_s2(i):
        j = samp[i]; n = 1 << sizes[j]
        out = np.zeros(35, np.uint32)
        fn(8, n, 1, C.c_void_p(o.ctypes.data + 2 * int(off_o[j])), n, C.c_void_p(b.ctypes.data + 2 * int(off_b[j])), P(out, u32p))
    sec = cpu_parallel(work_s2, len(samp), T)
    emit("s2_intra_rmd_batch", "PU/s", n_pu, k_s, w_s, algo,
         {"value": len(samp) / sec, "unit": "PU/s", "cores": T, "kind": kind, "sample": f"every 97th PU of the batch ({len(samp)} PUs, same size mix) in {sec:.2f} s"},
         p_s=p_s, h2d=int(o.nbytes + b.nbytes + n_pu * 16), d2h=int(sad_out.nbytes))

    # ---- f1: the coalescing queue: K encoder instances (host threads), each submitting one CU's worth of PUs per request and waiting
    #      for it (the live encoder's serial dependency), for K = 1, 4, 16: what coalescing buys over one-request-per-launch ----------------
    #      (a C++ driver: Python threads would measure the GIL) ------------------------------------------------------------------------------
    import subprocess
    exe = os.path.join(ROOT, "profiles", "ubench", "queue_bench")
    pkg = os.path.join(ROOT, "fast-cu-decision-hevc_b200")
    try:
        subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "profiles", "ubench", "queue_bench.cpp"),
                        "-L" + pkg, "-lcucudecide", "-Wl,-rpath," + pkg, "-lpthread", "-o", exe], check=True, capture_output=True)
        sys.stdout.write(subprocess.run([exe, "2000"], check=True, capture_output=True, text=True).stdout)
        sys.stdout.flush()
    except (subprocess.CalledProcessError, OSError) as e:
        print(json.dumps({"row": "f1_rmd_queue", "unavailable": str(e)[:200]}), flush=True)

    # ---- S3: integer-ME SAD surfaces: every whole 32x32 PU of the picture, +-32 window, FEN row sub-sampling ---------------
    pad = 80
    refp = np.pad(rec, pad, mode="edge")
    eng.set_cur_picture(org); eng.set_ref_picture(0, refp, pad, pad)
    descs = [dict(x=x, y=y, w=32, h=32, ref_idx=0, left=-32, right=32, top=-32, bottom=32, sub_shift=1)
             for y in range(0, H - 31, 32) for x in range(0, W - 31, 32)]
    n_pu = len(descs)
    me_arr, _, me_total = cucd.Engine.me_descs(descs)
    surf = np.zeros(me_total, np.uint32)
    k_s, w_s, p_s = timed(lambda: eng.me_sad_surface_raw(me_arr, n_pu, surf), eng, a.iters, pin=(surf,))
    algo = n_pu * (32 * 32 * 2 + 96 * 96 * 2 + 65 * 65 * 4)
    samp = descs[::40]
    Wp = W + 2 * pad
    fn = ref.hmref_sad_surface if ref is not None else oracle.oracle_sad_surface

    def work_s3(i):
        d = samp[i]
        blk = np.ascontiguousarray(org[d["y"]:d["y"] + 32, d["x"]:d["x"] + 32])
        out = np.zeros(65 * 65, np.uint32)
        base = refp.ctypes.data + 2 * ((d["y"] + pad) * Wp + d["x"] + pad)
        fn(8, P(blk, i16p), 32, 32, 32, C.c_void_p(base), Wp, -32, 32, -32, 32, 1, P(out, u32p))
    sec = cpu_parallel(work_s3, len(samp), T)
    emit("s3_me_sad_surface", "candidate SAD/s", n_pu * 65 * 65, k_s, w_s, algo,
         {"value": len(samp) * 65 * 65 / sec, "unit": "candidate SAD/s", "cores": T, "kind": kind,
          "sample": f"{len(samp)} of the {n_pu} PUs (32x32, 65x65 window, iSubShift 1) in {sec:.2f} s"},
         {"pus_per_call": n_pu}, p_s=p_s, h2d=int(n_pu * 40), d2h=int(surf.nbytes))

    # ---- f3: fractional-pel refinement of the same PUs around a pseudo-random integer MV, Hadamard -------------------------------------------
    sdescs = [dict(x=d["x"], y=d["y"], w=32, h=32, ref_idx=0, mvx=int(rng.integers(-16, 17)), mvy=int(rng.integers(-16, 17)), use_hadamard=1) for d in descs]
    sp_arr, _ = cucd.Engine.subpel_descs(sdescs)
    sp_out = np.zeros((len(sdescs), 49), np.uint32)
    k_s, w_s, p_s = timed(lambda: eng.me_subpel_cost_raw(sp_arr, len(sdescs), sp_out), eng, a.iters, pin=(sp_out,))
    algo = len(sdescs) * (32 * 32 * 2 + 41 * 41 * 2 + 49 * 4)
    samp = sdescs[::40]

    def work_f3(i):
        d = samp[i]
        blk = np.ascontiguousarray(org[d["y"]:d["y"] + 32, d["x"]:d["x"] + 32])
        out = np.zeros(49, np.uint32)
        base = refp.ctypes.data + 2 * ((d["y"] + pad) * Wp + d["x"] + pad)
        oracle.oracle_subpel_surface(8, P(blk, i16p), 32, 32, 32, C.c_void_p(base), Wp, d["mvx"], d["mvy"], 1, P(out, u32p))
    sec = cpu_parallel(work_f3, len(samp), T)
    emit("f3_me_subpel_cost", "PU refinement/s", len(sdescs), k_s, w_s, algo,
         {"value": len(samp) / sec, "unit": "PU refinement/s", "cores": T, "kind": "port",
          "sample": f"{len(samp)} of the {len(sdescs)} PUs (32x32, 49 quarter-pel positions, Hadamard) through the oracle in {sec:.2f} s"},
         p_s=p_s, h2d=int(len(sdescs) * 32), d2h=int(sp_out.nbytes))

    # ---- a12: TMV features of every whole CU of the picture (depths 0..3) -------------------------------------------------------
    cus = _util.all_cus(W, H)
    cu_arr, n_cu = cucd.Engine.cu_descs(cus)
    feat = np.zeros((n_cu, 5, 26))
    k_s, w_s, p_s = timed(lambda: eng.tmv_features_raw(cu_arr, n_cu, feat), eng, a.iters, pin=(feat,))
    algo = sum((1 << (2 * l)) * 2 + 130 * 8 for _, _, l in cus)
    samp = cus[::61]

    def work_a12(i):
        x, y, l = samp[i]
        out = np.zeros(130)
        if ref is not None:
            ref.hmref_tmv_features(P(org, i16p), W, W, H, x, y, 1 << l, P(out, f64p))
        else:
            oracle.oracle_tmv_features(C.c_void_p(org.ctypes.data + 2 * (y * W + x)), W, 1 << l, P(out, f64p))
    sec = cpu_parallel(work_a12, len(samp), 1)          # the reference shim copies the whole plane per call: 1 thread, plane copy included
    emit("a12_tmv_features", "CU/s", len(cus), k_s, w_s, algo,
         {"value": len(samp) / sec, "unit": "CU/s", "cores": 1, "kind": kind, "sample": f"every 61st CU ({len(samp)}) in {sec:.2f} s (driver copies the plane per call)"},
         p_s=p_s, h2d=int(n_cu * 12), d2h=int(feat.nbytes))

    # ---- a13: AQ activity, 4 layers ------------------------------------------------------------------------------------------------
    k_s, w_s, _ = timed(lambda: eng.aq_activity(4), eng, a.iters)
    units = sum(((W + (64 >> d) - 1) // (64 >> d)) * ((H + (64 >> d) - 1) // (64 >> d)) for d in range(4))
    algo = 4 * W * H * 2 + units * 8
    t0 = time.perf_counter()
    _util.oracle_aq_activity(oracle, org, 4)
    sec = time.perf_counter() - t0
    emit("a13_aq_activity", "AQ unit/s", units, k_s, w_s, algo,
         {"value": units / sec, "unit": "AQ unit/s", "cores": 1, "kind": "port", "sample": f"one 1080p picture, 4 layers, in {sec:.3f} s"})

    # ---- f2: intra TU coding: every whole TU position of the picture at 32/16/8/4 with a random mode ------------------------------------
    tus, orgs, brds = [], [], []
    for n in (32, 16, 8, 4):
        for y in range(0, H - n + 1, n):
            rows = org[y:y + n]
            for x in range(0, W - n + 1, n):
                tus.append((int(np.log2(n)), int(rng.integers(0, 35)), 32, 0))
                orgs.append(rows[:, x:x + n].ravel())
    o = np.concatenate(orgs)
    nb = sum((4 << t[0]) + 1 for t in tus)
    b = np.resize(o, nb).astype(np.int16)            # borders drawn from the picture's own samples
    n_tu = len(tus)
    samples = o.size
    _, tu_arr, _ = cucd.Engine._tu_descs(tus)
    coef = np.zeros(samples, np.int32); pix = np.zeros(samples, np.int16); dist = np.zeros(n_tu, np.uint32); asum = np.zeros(n_tu, np.int32)
    pins = (o, b, coef, pix, dist, asum)
    for name, stage, out_bytes in (("f2_intra_tu_forward", 0, samples * 6), ("f2_intra_tu_code", 1, samples * 6 + n_tu * 8), ("f2_intra_tu_recon", 2, samples * 2 + n_tu * 4)):
        k_s, w_s, p_s = timed(lambda: eng.intra_tu_raw(stage, tu_arr, n_tu, o, b, coef, pix, dist, asum), eng, a.iters, pin=pins)
        algo = o.nbytes + b.nbytes + out_bytes + (samples * 4 if stage == 2 else 0)
        cpu = None
        if name == "f2_intra_tu_code":
            samp = list(range(0, n_tu, 53))
            off_o = np.cumsum([0] + [1 << (2 * t[0]) for t in tus]); off_b = np.cumsum([0] + [(4 << t[0]) + 1 for t in tus])

            def work_tu(i):
                j = samp[i]; n = 1 << tus[j][0]
                _util.oracle_intra_tu(oracle, 8, n, tus[j][1], 32, 0, o[off_o[j]:off_o[j + 1]], b[off_b[j]:off_b[j + 1]], 1)
            sec = cpu_parallel(work_tu, len(samp), T)
            cpu = {"value": len(samp) / sec, "unit": "TU/s", "cores": T, "kind": "port",
                   "sample": f"every 53rd TU ({len(samp)}, same size mix) through the oracle's matrix-form chain in {sec:.2f} s"}
        emit(name, "TU/s", n_tu, k_s, w_s, algo, cpu, {"luma_samples_per_call": int(samples)}, p_s=p_s,
             h2d=int(o.nbytes + b.nbytes + n_tu * 16 + (samples * 4 if stage == 2 else 0)), d2h=int(out_bytes))
    eng.close()


if __name__ == "__main__":
    main()
