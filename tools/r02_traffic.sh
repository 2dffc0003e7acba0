#!/bin/bash
# DRAM traffic per launch of every kernel on the DEFAULT size of a workload (one ncu pass with the two DRAM
# counters: `--set full` cannot finish its ~40 replays of the 80 ms kernels on this size), the launch list of the
# same command, and -- default workload only -- the --set full capture with sources on a 40 Mb slice.
#   bash tools/r02_traffic.sh <tag> <workload c2|c4>
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; WL=${2:-c2}
O=gpurun_out
F="python bench.py --workload $WL --no-cpu-baseline --no-cli --steps 1 --warmup 1 --parity-reads 16 --parity-kmers 2e5"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'^k_' -s 27 -c 9 --csv --log-file $O/traffic_${WL}_$TAG.csv $F > $O/ncu_traffic_${WL}_$TAG.log 2>&1
python - <<PY
import csv, json
rows=[r for r in csv.reader(open("$O/traffic_${WL}_$TAG.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iu=h.index("Metric Unit"); iid=h.index("ID"); ig=h.index("Grid Size") if "Grid Size" in h else None
U={"byte":1.,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9,"Tbyte":1e12}
d={}
for r in rows[1:]:
    name=r[ik].split("(")[0]
    e=d.setdefault((r[iid],name),{})
    v=float(r[iv].replace(",",""))
    if r[im].startswith("dram__bytes"): e["bytes"]=e.get("bytes",0.)+v*U.get(r[iu],1.)
    else: e["dur"]="%s %s"%(r[iv],r[iu])
out={"workload":"bench.py --workload $WL, default size: one resident step (launches 28..36 of `$F`), ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"}
for (i,name),e in d.items():
    key = "retry_launch" if name=="k_classify" else name
    out[key]=e.get("bytes"); out[key+"_duration_under_ncu"]=e.get("dur")
json.dump(out,open("$O/traffic_r02_${WL}_$TAG.json","w"),indent=1)
print(json.dumps(out)[:1500])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_${WL}_$TAG.csv $F > $O/ncu_launches_${WL}_$TAG.log 2>&1
if [ "$WL" = c2 ]; then
  S="python bench.py --no-cpu-baseline --no-cli --genome-mb 40 --steps 1 --warmup 1 --parity-reads 16 --parity-kmers 2e5"
  ncu --set full --clock-control none --import-source on -k regex:'^k_(decode|wall_a|wall_b|wall_c|rel|unrel_a|unrel_b|emit)' -s 27 -c 8 -f -o $O/prof_$TAG $S > $O/ncu_full_$TAG.log 2>&1
  ls -la $O/prof_$TAG.ncu-rep
fi
