#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "decode" 2>&1 | tail -2
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']
        print('$1', 'value %.3e e2e %.3e dec %.2f ms (%.0f GB/s) cls %.2f ms' % (d['value'], d['e2e']['value'], k['k_decode']['ms'], k['k_decode']['GBps'], k['k_classify']['ms']))
"; }
$B --genome-mb 100 2>&1 | pick 100mb_fill
CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_dcb6.so $B --genome-mb 100 2>&1 | pick 100mb_fill_minblocks6
CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_dcb8.so $B --genome-mb 100 2>&1 | pick 100mb_fill_minblocks8
