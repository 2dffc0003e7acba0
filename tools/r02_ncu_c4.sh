#!/bin/bash
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
S="python bench.py --workload c4 --genome-mb 10 --no-cpu-baseline --no-cli --steps 1 --warmup 1 --parity-reads 16 --parity-kmers 2e5"
ncu --set full --clock-control none --import-source on -k regex:'k_wall_a' -s 1 -c 1 -f -o $O/prof_c4_$TAG $S > $O/ncu_c4_$TAG.log 2>&1
ls -la $O/prof_c4_$TAG.ncu-rep; tail -2 $O/ncu_c4_$TAG.log | cut -c1-200
