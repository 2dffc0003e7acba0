#!/bin/bash
# experiment: why is the whole-dataset batch slower per k-mer than a 10 Mb batch?
cd $GRAFT_REPO_ROOT
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']
        print('$1', 'value %.3e e2e %.3e dec %.2f ms cls %.2f ms' % (d['value'], d['e2e']['value'], k['k_decode']['ms'], k['k_classify']['ms']), k['k_classify']['phase_share'])
"; }
$B --genome-mb 10 2>&1 | pick 10mb_default
CPG_FORCE_P=60000 $B --genome-mb 10 2>&1 | pick 10mb_forceP60000
$B --genome-mb 100 2>&1 | pick 100mb_default
CPG_ORDER_CHUNK=16384 $B --genome-mb 100 2>&1 | pick 100mb_chunk16k
CPG_ORDER_CHUNK=4736 $B --genome-mb 100 2>&1 | pick 100mb_chunk4736
