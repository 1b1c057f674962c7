#!/bin/bash
# first GPU run of the dy-lane ME SAD kernel: parity of every ME entry point, then the two inter configurations
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_subpel.py -x -q -m gpu -k "me_ or subpel or bipred" > $O/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r3a_pytest.log
python bench.py --config ldp1080p --no-cpu-baseline --steps 5 --warmup 3 > $O/r3a_ldp.json 2> $O/r3a_ldp.err; echo "ldp rc=$?"
python bench.py --config ra1080p10 --no-cpu-baseline --steps 3 --warmup 3 > $O/r3a_ra.json 2> $O/r3a_ra.err; echo "ra rc=$?"
python - <<'P'
import json
for n in ("ldp","ra"):
    try:
        d=json.loads(open(f"gpurun_out/r3a_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["ms_per_step"],2), round(d["roofline"]["launch_ms"],2), round(d["e2e"]["value"]), d["paths_agree"], d["gpu_launches"])
    except Exception as e: print(n, "failed", e)
P
