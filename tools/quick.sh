#!/bin/bash
# quick check on one B200: GPU tests + the default bench without the CPU baseline, per-kernel times
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']
print('value %.3e e2e %.3e (%.1f ms) : dec %.2f wall %.1f rel %.1f unrel %.1f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], k['k_decode']['ms'], k['k_wall']['ms'], k['k_rel']['ms'], k['k_unrel']['ms']))"
