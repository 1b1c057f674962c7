#!/bin/bash
# 8-GPU box: the copy-pattern ceiling (no kernels) and bench.py at N = 1, 2, 4, 8 ranks.  Run from the repo root.
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python profiles/ubench/pcie_pattern.py > $O/r2m_pcie_n$n.json 2> $O/r2m_pcie_n$n.err
  else $TR --nproc-per-node $n --master-port 2951$n profiles/ubench/pcie_pattern.py > $O/r2m_pcie_n$n.json 2> $O/r2m_pcie_n$n.err; fi
  echo "pcie n=$n rc=$?"
done
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hm-planes > $O/r2m_bench_n$n.json 2> $O/r2m_bench_n$n.err
  else $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-hm-planes > $O/r2m_bench_n$n.json 2> $O/r2m_bench_n$n.err; fi
  echo "bench n=$n rc=$?"
done
nproc; nvidia-smi topo -m 2>/dev/null | head -12 > $O/r2m_topo.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" > $O/r2m_cpu.txt
