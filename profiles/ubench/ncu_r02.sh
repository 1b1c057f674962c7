#!/bin/bash
# ncu --set full captures of every kernel of the library + the launch list of the default bench (run on the GPU box from the repo root).
# Each command is first run without ncu (B200_PROFILING.md: capture only what exits 0 unprofiled); numbers printed under ncu are not bench values.
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
A="python bench.py --pics 4 --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes --e2e-instances 1"
B="python bench_rows.py --iters 1"
C="python bench.py --config ai2160p10 --pics 1 --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes --e2e-instances 1"
$A > $O/r2k_plainA.log 2>&1 && $NCU -k regex:"rmd_frame_tc2|feature_|ctu_src" -s 12 -c 8 -o $O/r2k_prof_frame -f $A > $O/r2k_ncuA.log 2>&1; echo "ncuA rc=$?"
$C > $O/r2k_plainC.log 2>&1 && $NCU -k regex:"rmd_frame_tc3" -s 3 -c 1 -o $O/r2k_prof_tc3 -f $C > $O/r2k_ncuC.log 2>&1; echo "ncuC rc=$?"
$B > $O/r2k_plainB.log 2>&1 && {
  $NCU -k regex:"me_sad" -s 2 -c 2 -o $O/r2k_prof_mesad -f $B > $O/r2k_ncuB1.log 2>&1; echo "ncuB1 rc=$?"
  $NCU -k regex:"me_subpel" -s 2 -c 1 -o $O/r2k_prof_subpel -f $B > $O/r2k_ncuB2.log 2>&1; echo "ncuB2 rc=$?"
  $NCU -k regex:"tmv_feature" -s 2 -c 1 -o $O/r2k_prof_tmv -f $B > $O/r2k_ncuB3.log 2>&1; echo "ncuB3 rc=$?"
}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hm-planes > $O/r2k_ncuD.log 2>&1; echo "ncuD rc=$?"
