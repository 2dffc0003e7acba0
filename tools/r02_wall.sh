#!/bin/bash
# k_wall_a/b/c: GPU parity, then the bench line for the default build and for lane-group variants
#   bash tools/r02_wall.sh <tag> [variant ...]     (variants: names of classpro_b200/build/lib_<name>.so)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; shift
O=gpurun_out
python -m pytest tests/test_gpu.py -m gpu -x -q 2>&1 | tail -30 | tee $O/pytest_gpu_$TAG.log
B="python bench.py --no-cpu-baseline --no-cli --steps 3 --warmup 3"
$B > $O/bench_$TAG.log 2> $O/bench_$TAG.err || tail -5 $O/bench_$TAG.err
for vv in "$@"; do
  v=${vv%%:*}; E=""; [ "$vv" != "$v" ] && E=${vv#*:}
  env $E CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so $B --parity-reads 64 --parity-kmers 1e6 > $O/bench_${TAG}_$v.log 2> $O/bench_${TAG}_$v.err || tail -3 $O/bench_${TAG}_$v.err
  done
