#!/bin/bash
# GPU parity, then the bench line of the default workload and of the repeat-rich one
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
python -m pytest tests/test_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu_$TAG.log
for wl in ${WLS:-c2 c4}; do
python bench.py --workload $wl --no-cpu-baseline --no-cli --steps 3 --warmup 3 > $O/bench_${wl}_$TAG.log 2> $O/bench_${wl}_$TAG.err || tail -5 $O/bench_${wl}_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench_${wl}_$TAG.log").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("$wl", d["config"]["kmers"], "resident", round(d["ms_per_step"],1), {x:(round(k[x]["ms"],2)) for x in k}, {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "e2e", round(d["e2e"]["ms_per_step"],1), "flips", d["parity_sample"]["flips"], "bad", d["reads_with_errors"])
PY
done
for v in "$@"; do
  [ "$v" = "$TAG" ] && continue
  CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so python bench.py --workload c4 --no-cpu-baseline --no-cli --steps 3 --warmup 3 --parity-reads 64 --parity-kmers 1e6 > $O/bench_c4_${TAG}_$v.log 2> $O/bench_c4_${TAG}_$v.err || tail -3 $O/bench_c4_${TAG}_$v.err
  python - <<PY
import json
d=json.loads(open("$O/bench_c4_${TAG}_$v.log").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("c4 variant $v", "resident", round(d["ms_per_step"],1), {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "flips", d["parity_sample"]["flips"])
PY
done
