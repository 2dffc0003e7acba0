cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
O=gpurun_out; TAG=r1j
S="python bench.py --no-cpu-baseline --genome-mb 40 --steps 2 --warmup 1"
$S 2>&1 | tail -1 | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv $S > $O/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 15 -c 5 -f -o $O/prof_$TAG $S > $O/ncu_full_$TAG.log 2>&1
ls -la $O/prof_$TAG.ncu-rep
