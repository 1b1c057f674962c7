#!/bin/bash
# re-capture of the frame kernels after the last kernel change of round 2 (dedicated issuing warps): same recipe as ncu_r02.sh
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
A="python bench.py --pics 4 --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes --e2e-instances 1"
[ "$1" = "tc3" ] || $A > $O/r2y_plainA.log 2>&1 && $NCU -k regex:"rmd_frame_tc2" -s 3 -c 2 -o $O/r2y_prof_frame -f $A > $O/r2y_ncuA.log 2>&1; echo "ncuA rc=$?"
[ "$1" = "tc3" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes > $O/r2y_ncuD.log 2>&1; echo "ncuD rc=$?"
# ... and of rmd_frame_tc3_kernel after the residual-before-MMA-2 change
if [ "$1" = "tc3" ]; then
C="python bench.py --config ai2160p10 --pics 1 --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes --e2e-instances 1"
$C > $O/r2y_plainC.log 2>&1 && $NCU -k regex:"rmd_frame_tc3" -s 3 -c 1 -o $O/r2y_prof_tc3 -f $C > $O/r2y_ncuC.log 2>&1; echo "ncuC rc=$?"
fi
