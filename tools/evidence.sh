#!/bin/bash
# Round evidence on one B200: GPU tests, the default bench line (both arms), the CLI on config c1
# against the reference binary, the ncu launch list and --set full captures.
#   bash tools/evidence.sh <tag>       (outputs under gpurun_out/)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_full_$TAG.log 2> $O/bench_full_$TAG.err || tail -5 $O/bench_full_$TAG.err
tail -1 $O/bench_full_$TAG.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2>&1; tail -1 $O/bench_ref_$TAG.log | cut -c1-300
python tools/run_config.py c1 > $O/config_c1_$TAG.json 2>&1; tail -1 $O/config_c1_$TAG.json | cut -c1-200
python tools/run_config.py c4 --scale 0.004 > $O/config_c4_$TAG.json 2>&1; tail -1 $O/config_c4_$TAG.json | cut -c1-200
# ncu: launch list and source-level capture on a 40 Mb slice (60 k reads: enough to fill the 38 k lane groups of k_wall), DRAM traffic on the default workload
S="python bench.py --no-cpu-baseline --genome-mb 40 --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv $S > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 15 -c 5 -f -o $O/prof_$TAG $S > $O/ncu_full_$TAG.log 2>&1
F="python bench.py --no-cpu-baseline --steps 1 --warmup 1"
ncu --set full --clock-control none -k regex:k_ -s 15 -c 5 -f -o $O/prof_full_$TAG $F > $O/ncu_fullwl_$TAG.log 2>&1
ls -la $O/prof_$TAG.ncu-rep $O/prof_full_$TAG.ncu-rep
# profile producer: parity with the harness counter + stage times, then the launch list and one
# full capture of its kernels; the forced 3-pass run is the open bug of round 1
CPG_COUNT_TIMING=1 python tools/producer_check.py 2000000 30 40,32,21 > $O/producer_$TAG.log 2>&1; tail -4 $O/producer_$TAG.log
CPG_COUNT_PASSES=3 python tools/producer_check.py 300000 20 40 > $O/producer_passes_$TAG.log 2>&1; tail -2 $O/producer_passes_$TAG.log
P="python tools/producer_check.py 1000000 30 40"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/producer_launches_$TAG.csv $P > $O/ncu_producer_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_kmer|k_run|k_scatter|k_enc' -c 8 -f -o $O/prof_producer_$TAG $P > $O/ncu_producer_$TAG.log 2>&1
