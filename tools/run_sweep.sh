#!/bin/bash
# File-level parity of the CLI against the reference binary on scaled versions of BASELINE.json
# configs 2 and 5 (coverage / read-length sweep, -c auto vs fixed); configs 1 and 4 are in
# tools/evidence.sh.  One JSON line per run ("identical": true = byte-identical .class).
cd $GRAFT_REPO_ROOT
O=gpurun_out/sweep_${1:-x}.jsonl
: > $O
python tools/run_config.py c2 --scale 0.05 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.003 --cov 10 --len 10000 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.003 --cov 10 --len 10000 --fixed-c 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.003 --cov 50 --len 15000 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.002 --cov 100 --len 25000 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.002 --cov 100 --len 25000 --fixed-c 2>&1 | tail -1 >> $O
python tools/run_config.py c5 --scale 0.003 --cov 20 --len 20000 --fixed-c 2>&1 | tail -1 >> $O
python -c "
import json,sys
for l in open('$O'):
    try: d=json.loads(l)
    except Exception: print('BAD', l[:200]); continue
    print(d['config'], d['genome_len'], d['cov'], d.get('args'), 'reads', d['reads'], 'kmers', d['kmers'], 'identical', d.get('identical'), 'flips', d.get('flipped_chars'), 'ref_s', d.get('ref_s'), 'gpu_s', d.get('gpu_s'), d.get('gpu_error','')[:200], d.get('ref_error','')[:200])
"
