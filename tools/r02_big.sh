#!/bin/bash
# file-level parity at full / large size (exact counts from the GPU profiler) + the c4 bench workload
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
python tools/run_big.py c2 --genome-mb 100 > $O/big_c2_$TAG.json 2> $O/big_c2_$TAG.err; tail -1 $O/big_c2_$TAG.json | cut -c1-1500
python tools/run_big.py c4 --genome-mb 50 > $O/big_c4_$TAG.json 2> $O/big_c4_$TAG.err; tail -1 $O/big_c4_$TAG.json | cut -c1-1500
python bench.py --workload c4 --no-cpu-baseline --no-cli --steps 3 --warmup 3 > $O/bench_c4_$TAG.log 2> $O/bench_c4_$TAG.err || tail -5 $O/bench_c4_$TAG.err
python - <<PY
import json
d=json.loads(open("$O/bench_c4_$TAG.log").read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print("c4", d["config"]["kmers"], "resident", round(d["ms_per_step"],1), {x:(round(k[x]["ms"],2)) for x in k}, {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "e2e", round(d["e2e"]["ms_per_step"],1), "flips", d["parity_sample"]["flips"], "bad", d["reads_with_errors"])
PY
