#!/bin/bash
# GPU parity, then the bench line of the default workload and of the repeat-rich one (+ variants of the library)
#   bash tools/r02_both.sh <tag> [variant ...]      WLS="c2 c4" chooses the workloads
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; shift
O=gpurun_out
python -m pytest tests/test_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest_gpu_$TAG.log
for wl in ${WLS:-c2 c4}; do
  python bench.py --workload $wl --no-cpu-baseline --no-cli --steps 3 --warmup 3 > $O/bench_${wl}_$TAG.log 2> $O/bench_${wl}_$TAG.err || tail -5 $O/bench_${wl}_$TAG.err
  python tools/benchsum.py $wl=$O/bench_${wl}_$TAG.log
  for v in "$@"; do
    CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so python bench.py --workload $wl --no-cpu-baseline --no-cli --steps 3 --warmup 3 --parity-reads 64 --parity-kmers 1e6 > $O/bench_${wl}_${TAG}_$v.log 2> $O/bench_${wl}_${TAG}_$v.err || tail -3 $O/bench_${wl}_${TAG}_$v.err
    python tools/benchsum.py $wl/$v=$O/bench_${wl}_${TAG}_$v.log
  done
done
