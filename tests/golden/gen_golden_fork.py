#!/usr/bin/env python3
"""Golden for the fork-aware frame mode (cucd_set_decision_switches): what the REAL encoder visits on Testing pictures.

Runs oracle/_ref/TAppEncoder (All-Intra, 416x240, 6 pictures: POC 0-1 Training, 2 Verifying, 3-5 Testing) with the RMD, OBF and
decision-switch dump hooks on and stores, for every Testing picture: the OBF plane, the per-depth Skip2Nx2N / TerminateCU switches
(g_bDecisionSwitch as SetDecisionSwitch left them after the verify picture) and the exact set of luma PUs (x, y, N) the rough-mode-
decision loop TEncSearch.cpp:2327-2361 was entered for.  tests/ check that the library's pruned enumeration is that set.

Needs /root/reference (through oracle/_ref); not run on the GPU box.  Usage: python tests/golden/gen_golden_fork.py
"""
import os
import struct
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as gg  # noqa: E402


def main():
    W, H, frames = 416, 240, 6
    out = {}
    for name, bd, qp, seed in (("a8", 8, 32, 20261040), ("b8", 8, 22, 20261041)):
        with tempfile.TemporaryDirectory(prefix="cucd_fork_") as wd:
            open(os.path.join(wd, "clip.yuv"), "wb").write(gg.synth_clip(W, H, frames, bd, seed))
            gg.run_encoder(wd, W, H, frames, bd, qp, gg.AI, {"CUCD_DUMP_RMD": "rmd.bin", "CUCD_DUMP_OBF": "obf.bin", "CUCD_DUMP_OBF_YC": "yc.bin",
                                                            "CUCD_DUMP_CU": "cu.bin", "CUCD_DUMP_SWITCHES": "sw.bin"})
            recs = gg.read_rmd(os.path.join(wd, "rmd.bin"))
            sw = np.frombuffer(open(os.path.join(wd, "sw.bin"), "rb").read(), np.int32).reshape(-1, 9)
            obf = gg.read_obf(wd, frames)
            pocs = []
            for k in range(int(obf["nframes"][0])):
                poc = int(obf[f"f{k}_meta"][0])
                row = sw[sw[:, 0] == poc]
                if poc % 60 < 3 or not len(row):
                    continue
                vis = sorted({(r["x"], r["y"], r["n"]) for r in recs if r["poc"] == poc})
                out[f"{name}_p{poc}_obf"] = obf[f"f{k}_obf"]
                out[f"{name}_p{poc}_org"] = obf[f"f{k}_org"]
                out[f"{name}_p{poc}_skip"] = row[0, 1:5].astype(np.uint8)
                out[f"{name}_p{poc}_term"] = row[0, 5:9].astype(np.uint8)
                out[f"{name}_p{poc}_visited"] = np.array(vis, np.int16)
                pocs.append(poc)
                print(f"{name} POC {poc}: switches skip {row[0, 1:5]} term {row[0, 5:9]}, {len(vis)} RMD PUs visited")
            out[f"{name}_pocs"] = np.array(pocs, np.int32)
            out[f"{name}_meta"] = np.array([W, H, bd], np.int32)
    np.savez_compressed(os.path.join(HERE, "fork_ai8.npz"), **out)


if __name__ == "__main__":
    main()
