"""bench.py's reference arm runs on the host cores alone, so its JSON line (the contract the driver parses) can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--pics", "1",
                        "--width", "416", "--height", "240"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cu_decision_ctus_per_s_1080p_all_intra" and d["unit"] == "CTU/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_smoke_wide_me_block_runs_against_a_stand_in_engine(oracle):
    """__graft_entry__._smoke_me_wide (the part of smoke() that drives the dy-lane SAD kernel and the one-CTA sub-pel kernel) with a
    stand-in engine that answers from the oracle: the window / pointer arithmetic of the check itself is exercised without a GPU"""
    import ctypes as C
    import importlib
    import numpy as np
    from _util import P, i16p, u32p, pseudo_recon, textured_plane
    ge = importlib.import_module("__graft_entry__")
    W, H, bd = 200, 136, 8
    org = textured_plane(W, H, bd, seed=1)
    rec = pseudo_recon(org, bd)

    class StandIn:
        def set_ref_picture(self, idx, plane, mx, my):
            self.plane, self.mx, self.my = plane, mx, my

        def _base(self, d, dx=0, dy=0):
            S = self.plane.shape[1]
            return C.c_void_p(self.plane.ctypes.data + 2 * ((d["y"] + self.my + dy) * S + d["x"] + self.mx + dx)), S

        def me_sad_surface(self, descs):
            res = []
            for d in descs:
                blk = np.ascontiguousarray(org[d["y"]:d["y"] + d["h"], d["x"]:d["x"] + d["w"]])
                out = np.zeros((d["bottom"] - d["top"] + 1, d["right"] - d["left"] + 1), np.uint32)
                base, S = self._base(d)
                oracle.oracle_sad_surface(bd, P(blk, i16p), d["w"], d["w"], d["h"], base, S, d["left"], d["right"], d["top"], d["bottom"], d["sub_shift"], P(out, u32p))
                res.append(out)
            return res

        def me_subpel_cost(self, descs):
            res = []
            for d in descs:
                blk = np.ascontiguousarray(org[d["y"]:d["y"] + d["h"], d["x"]:d["x"] + d["w"]])
                out = np.zeros(49, np.uint32)
                base, S = self._base(d)
                oracle.oracle_subpel_surface(bd, P(blk, i16p), d["w"], d["w"], d["h"], base, S, d["mvx"], d["mvy"], d["use_hadamard"], P(out, u32p))
                res.append(out.reshape(7, 7))
            return res

    ge._smoke_me_wide(StandIn(), oracle, org, rec, bd)
