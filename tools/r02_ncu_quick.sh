#!/bin/bash
# quick ncu counters of chosen kernels for the default build and variants (40 Mb slice of the bench workload)
#   bash tools/r02_ncu_quick.sh <tag> <kernel regex> [variant ...]
cd $GRAFT_REPO_ROOT
TAG=$1; KR=$2; shift; shift
O=gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
S="python bench.py --no-cpu-baseline --no-cli --genome-mb 40 --steps 1 --warmup 1 --parity-reads 16 --parity-kmers 2e5"
for v in default "$@"; do
  L=""; [ "$v" != default ] && L="CLASSPRO_B200_LIB=$PWD/classpro_b200/build/lib_$v.so"
  env $L ncu --metrics $M --clock-control none -k regex:$KR -s 8 -c 8 --csv --log-file $O/ncuq_${TAG}_$v.csv $S > $O/ncuq_${TAG}_$v.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/ncuq_${TAG}_$v.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iid=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[iid],r[ik]),{})[r[im]]=r[iv]
for (i,k),m in d.items():
    print("$v",k.split("(")[0], {a.split(".")[0].replace("smsp__","").replace("sm__",""):b for a,b in m.items()})
PY
done
