#!/bin/bash
# Fresh ncu evidence for both kernels: plain run first, then the launch list, then one --set full
# capture per kernel (after warm-up launches).  Usage: bash tools/prof.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
B="python bench.py --no-cpu-baseline --genome-mb 10 --steps 2 --warmup 1"
$B > gpurun_out/bench_small_$TAG.log 2>&1 || { tail -20 gpurun_out/bench_small_$TAG.log; exit 1; }
tail -1 gpurun_out/bench_small_$TAG.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 4 -c 2 -f -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
