#!/bin/bash
# final measurement batch of round 2 (run from the repo root on the GPU box): the contract line of every BASELINE configuration,
# the reference arm, the per-operator rows and the integrated-encoder wall clock
O=gpurun_out
python bench.py > $O/r2z_bench_ai1080p8.json 2> $O/r2z_bench_ai1080p8.err; echo "ai1080p8 rc=$?"
python bench.py --fork-aware > $O/r2z_bench_ai1080p8_fork_aware.json 2> $O/r2z_bench_fork.err; echo "fork-aware rc=$?"
for c in ai2160p10 ldp1080p ra1080p10; do python bench.py --config $c > $O/r2z_bench_$c.json 2> $O/r2z_bench_$c.err; echo "$c rc=$?"; done
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2z_bench_reference.json 2> $O/r2z_bench_reference.err; echo "reference rc=$?"
python bench_rows.py > $O/r2z_rows.jsonl 2> $O/r2z_rows.err; echo "rows rc=$?"
python profiles/ubench/encoder_wallclock.py --instances 1,8 > $O/r2z_wallclock.json 2> $O/r2z_wallclock.err; echo "wallclock rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
