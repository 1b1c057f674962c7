#!/bin/bash
# GPU run of me_subpel_small_kernel: parity of the ME / sub-pel entry points, the live LDP encoder (byte-identical bitstream), then the
# contract lines of the two inter configurations
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_subpel.py tests/test_gpu_encoder_md5.py -x -q -m gpu -k "me_ or subpel or bipred or LDP-8-3-32" > $O/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3d_pytest.log
for c in ldp1080p ra1080p10; do python bench.py --config $c > $O/r3d_bench_$c.json 2> $O/r3d_bench_$c.err; echo "$c rc=$?"; done
python - <<'P'
import json
for n in ("ldp1080p","ra1080p10"):
    try:
        d=json.loads(open(f"gpurun_out/r3d_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["ms_per_step"],2), round(d["roofline"]["launch_ms"],2), round(d["e2e"]["value"]), d["paths_agree"], d["gpu_launches"])
    except Exception as e: print(n, "failed", e)
P
