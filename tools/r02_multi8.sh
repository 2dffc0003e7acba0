#!/bin/bash
# 8-GPU evidence: ClassPro -G<n> for every n against the reference binary, then the sharded bench at N
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; N=${2:-8}
O=gpurun_out
nvidia-smi -L > $O/gpus_$TAG.log; nproc >> $O/gpus_$TAG.log
python -m pytest tests/test_gpu.py -m gpu -q -s -k "every_gpu_count" 2>&1 | tail -14 | tee $O/pytest_gpu_$TAG.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --no-cli --parity-reads 256 --parity-kmers 4e6 > $O/bench_n${N}_$TAG.log 2> $O/bench_n${N}_$TAG.err || tail -5 $O/bench_n${N}_$TAG.err
python tools/benchsum.py N$N=$O/bench_n${N}_$TAG.log
python - <<PY
import json
d=json.loads(open("$O/bench_n${N}_$TAG.log").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "imbalance", d["run"]["rank_time_imbalance"], "gen_s", d["run"]["gen_seconds"])
for s in d["run"]["shards"]: print(s)
PY
