#!/bin/bash
# Round evidence on one B200 (outputs under gpurun_out/, copied by hand into profiles/):
#   bash tools/evidence.sh <tag>
# GPU tests + both workloads of the bench, the full default bench line (CLI figure, CPU baseline) and the
# reference arm, DRAM traffic per kernel + launch list + ncu --set full with sources, file-level parity at full
# size (exact counts from the GPU profiler) with every differing read traced.
# Multi-GPU: `gpurun --gpus N -- bash tools/r02_multi8.sh <tag> N`  (mind the budget: N x the wall time).
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/r02_both.sh $TAG
python bench.py > $O/bench_full_$TAG.log 2> $O/bench_full_$TAG.err || tail -5 $O/bench_full_$TAG.err
python tools/benchsum.py full=$O/bench_full_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2>&1; tail -1 $O/bench_ref_$TAG.log | cut -c1-300
bash tools/r02_traffic.sh $TAG c2
python tools/run_big.py c2 --genome-mb 100 --keep /tmp/big_c2 > $O/big_c2_$TAG.json 2> $O/big_c2_$TAG.err; tail -1 $O/big_c2_$TAG.json | cut -c1-300
python tools/trace_file_flips.py /tmp/big_c2 40 $O/flipreads_c2_$TAG.npz > $O/flips_c2_$TAG.json 2>&1; rm -rf /tmp/big_c2
bash tools/r02_c4b.sh $TAG
