#!/bin/bash
# ncu --set full with sources for every kernel of the classification path (40 Mb slice of the bench
# workload: 60 k reads), one capture of each kernel from the second pass over the data
#   bash tools/r02_ncu_full.sh <tag>
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
S="python bench.py --no-cpu-baseline --no-cli --genome-mb 40 --steps 1 --warmup 1 --parity-reads 16 --parity-kmers 2e5"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$TAG.csv $S > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(decode|wall_a|wall_b|wall_c|rel|unrel_a|unrel_b|emit)' -s 9 -c 9 -f -o $O/prof_$TAG $S > $O/ncu_full_$TAG.log 2>&1
ls -la $O/prof_$TAG.ncu-rep
tail -3 $O/ncu_full_$TAG.log
