#!/bin/bash
# last measurement batch of round 2 (after the dy-lane ME SAD kernel and the grouped sub-pel passes): contract line of every BASELINE
# configuration, the reference arm, the per-operator rows, smoke(), then ncu --set full of the ME kernels (after the plain runs exited 0)
O=gpurun_out
python bench.py > $O/r3z_bench_ai1080p8.json 2> $O/r3z_bench_ai1080p8.err; echo "ai1080p8 rc=$?"
python bench.py --fork-aware > $O/r3z_bench_ai1080p8_fork_aware.json 2> $O/r3z_bench_fork.err; echo "fork-aware rc=$?"
for c in ai2160p10 ldp1080p ra1080p10; do python bench.py --config $c > $O/r3z_bench_$c.json 2> $O/r3z_bench_$c.err; echo "$c rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > $O/r3z_bench_reference.json 2> $O/r3z_bench_reference.err; echo "reference rc=$?"
python bench_rows.py > $O/r3z_rows.jsonl 2> $O/r3z_rows.err; echo "rows rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
NCU="ncu --set full --clock-control none --import-source on"
for c in ldp1080p ra1080p10; do
  A="python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline"
  $NCU -k regex:"me_sad_dy|me_subpel" -c 2 -o $O/r3z_prof_$c -f $A > $O/r3z_ncu_$c.log 2>&1; echo "$c ncu rc=$?"
done
