#!/bin/bash
# whole -m gpu suite, the default bench line (with the CLI figure), then the e2e pass with other batch counts
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu_$TAG.log
python bench.py --steps 3 --warmup 3 > $O/bench_$TAG.log 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
for nb in 2 3 6; do
  python bench.py --no-cpu-baseline --no-cli --steps 3 --warmup 3 --batches $nb --parity-reads 32 --parity-kmers 5e5 > $O/bench_${TAG}_b$nb.log 2> $O/bench_${TAG}_b$nb.err || tail -3 $O/bench_${TAG}_b$nb.err
done
python - <<PY
import json
for t in ["", "_b2", "_b3", "_b6"]:
    try:
        d=json.loads(open("$O/bench_$TAG%s.log" % t).read().strip().splitlines()[-1])
        print(t or "b4", "resident", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "e2e class strings", round(d["e2e"]["class_strings"]["ms_per_step"],1), "same", d["e2e"]["matches_resident_result"], "cli", (d.get("cli") or {}).get("wall_s"), (d.get("cli") or {}).get("stages"))
    except Exception as e: print(t, "failed", e)
PY
