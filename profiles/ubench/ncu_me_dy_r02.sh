#!/bin/bash
# ncu --set full of the dy-lane ME SAD kernel (8-bit: ldp1080p, 10-bit: ra1080p10), one launch each, after the plain run exited 0
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for c in ldp1080p ra1080p10; do
  A="python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline"
  $A > $O/r3b_plain_$c.log 2>&1 && $NCU -k regex:"me_sad" -c 2 -o $O/r3b_prof_$c -f $A > $O/r3b_ncu_$c.log 2>&1; echo "$c ncu rc=$?"
done
