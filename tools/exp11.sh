#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/gpu_check.py 2>&1 | grep '^{' | cut -c1-150
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']
        print('$1', 'value %.3e e2e %.3e dec %.2f ms cls %.2f ms' % (d['value'], d['e2e']['value'], k['k_decode']['ms'], k['k_classify']['ms']), k['k_classify']['phase_share'])
"; }
$B --genome-mb 100 2>&1 | pick 100mb_split
CPG_FUSED=1 $B --genome-mb 100 2>&1 | pick 100mb_fused
$B --genome-mb 10 2>&1 | pick 10mb_split
python tools/run_config.py c1 2>&1 | tail -1 | cut -c1-1400
