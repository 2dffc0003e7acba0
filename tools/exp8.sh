#!/bin/bash
cd $GRAFT_REPO_ROOT
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']
        print('$1', 'value %.3e e2e %.3e dec %.2f ms cls %.2f ms' % (d['value'], d['e2e']['value'], k['k_decode']['ms'], k['k_classify']['ms']), k['k_classify']['phase_share'])
"; }
$B --genome-mb 100 2>&1 | pick 100mb_g8_default
CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_nosync.so $B --genome-mb 100 2>&1 | pick 100mb_g8_nosync
CPG_ORDER_CHUNK=9472 $B --genome-mb 100 2>&1 | pick 100mb_g8_chunk9472
CPG_ORDER_CHUNK=37888 $B --genome-mb 100 2>&1 | pick 100mb_g8_chunk37888
