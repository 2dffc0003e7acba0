#!/bin/bash
# Multi-GPU evidence: run under `gpurun --gpus N`.  The whole -m gpu suite (the CLI -G<n>
# test runs for every n <= devices, the producer's key-range passes), then the sharded bench at N.
#   bash tools/r02_multi.sh <tag> <N>
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; N=${2:-2}
O=gpurun_out
nvidia-smi -L > $O/gpus_$TAG.log
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee $O/pytest_gpu_$TAG.log
python tools/producer_debug.py > $O/producer_debug_$TAG.log 2>&1; tail -6 $O/producer_debug_$TAG.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
   bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_n${N}_$TAG.log 2> $O/bench_n${N}_$TAG.err || tail -5 $O/bench_n${N}_$TAG.err
tail -1 $O/bench_n${N}_$TAG.log | cut -c1-1500
