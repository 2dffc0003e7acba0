#!/bin/bash
# k_wall_a/b/c: GPU parity, then the bench line for the default build and for lane-group variants
#   bash tools/r02_wall.sh <tag> [variant ...]     (variants: names of classpro_b200/build/lib_<name>.so)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; shift
O=gpurun_out
python -m pytest tests/test_gpu.py -m gpu -x -q 2>&1 | tail -30 | tee $O/pytest_gpu_$TAG.log
B="python bench.py --no-cpu-baseline --no-cli --steps 3 --warmup 3"
$B > $O/bench_$TAG.log 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench_$TAG.log").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("default", round(d["ms_per_step"],1), {x:(round(k[x]["ms"],2)) for x in k}, {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "e2e", round(d["e2e"]["ms_per_step"],1), "flips", d["parity_sample"]["flips"], "bad", d["reads_with_errors"])
PY
for vv in "$@"; do
  v=${vv%%:*}; E=""; [ "$vv" != "$v" ] && E=${vv#*:}
  env $E CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so $B --parity-reads 64 --parity-kmers 1e6 > $O/bench_${TAG}_$v.log 2> $O/bench_${TAG}_$v.err || tail -3 $O/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${TAG}_$v.log").read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print("$vv", round(d["ms_per_step"],1), {x:(round(k[x]["ms"],2)) for x in k}, {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "flips", d["parity_sample"]["flips"], "bad", d["reads_with_errors"])
except Exception as e: print("$v failed", e)
PY
done
