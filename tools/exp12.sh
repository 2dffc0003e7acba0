#!/bin/bash
cd $GRAFT_REPO_ROOT
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; c=k['k_classify']; ps=c['phase_share']
        print('$1', 'value %.3e e2e %.3e cls %.2f ms : wall %.1f rel %.1f unrel %.1f retry %.2f' % (d['value'], d['e2e']['value'], c['ms'], c['ms']*ps['wall'], c['ms']*ps['reliable_dp'], c['ms']*ps['unreliable_emit'], c['ms']*ps['barrier_wait']))
"; }
$B 2>&1 | pick base_8_8_8
for v in W4 W16 W32 R16 U4 U16; do
CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so $B 2>&1 | pick $v
done
