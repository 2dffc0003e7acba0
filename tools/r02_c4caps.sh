#!/bin/bash
# which capacity sends the reads of the repeat-rich workload to the retry launch?
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
B="python bench.py --workload c4 --genome-mb 10 --no-cpu-baseline --no-cli --steps 2 --warmup 3 --parity-reads 32 --parity-kmers 5e5"
run() { name=$1; shift; env "$@" $B > $O/c4caps_${TAG}_$name.log 2> $O/c4caps_${TAG}_$name.err || tail -3 $O/c4caps_${TAG}_$name.err
python - <<PY
import json
try:
    d=json.loads(open("$O/c4caps_${TAG}_$name.log").read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print("$name", "resident", round(d["ms_per_step"],1), "retry", round(k["retry_launch"]["ms"],1), {a:round(b,2) for a,b in list(k["k_wall"]["launches"].items())+list(k["k_unrel"]["launches"].items())}, "rel", round(k["k_rel"]["ms"],1), "flips", d["parity_sample"]["flips"])
except Exception as e: print("$name failed", e)
PY
}
run default X=1
run scratch4 CPG_SCRATCH_DIV=4
run hdr6 CPG_HDR_DIV=6
run big20 CPG_BIG_DIV=20
run all CPG_SCRATCH_DIV=4 CPG_HDR_DIV=6 CPG_BIG_DIV=20 CPG_UPRE_DIV=12 CPG_POOL_DIV=4
