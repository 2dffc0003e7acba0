#!/bin/bash
# quick check on one B200: GPU tests + the default bench without the CPU baseline, per-kernel times
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-cli --steps 3 --warmup 3 > gpurun_out/quick.log 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python tools/benchsum.py quick=gpurun_out/quick.log
