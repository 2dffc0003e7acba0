"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel in an
`ncu --set full` report -> profiles/traffic.json, which bench.py copies into `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/prof_full.ncu-rep "<workload description>" > profiles/traffic.json

Launches of one kernel are averaged; k_classify here is the retry launch (a 4-CTA grid that normally
finds nothing to do).
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0]
        grid = int(float(r[col["launch__grid_size"]]))
        if name == "k_classify" and grid < 16:
            name = "retry_launch"
        tot = 0.
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[key]]) * UNIT[units[col[key]]]
        dur = float(r[col["gpu__time_duration.sum"]])
        acc.setdefault(name, []).append((tot, dur, units[col["gpu__time_duration.sum"]]))
    res = {"source": rep.split("/")[-1], "workload": what}
    for name, v in acc.items():
        tot = sum(x[0] for x in v) / len(v)
        res[name] = tot if tot == tot else None        # NaN: ncu could not collect the DRAM counters for this launch
        res[name + "_launches"] = len(v)
        res[name + "_duration_under_ncu"] = "%.3f %s" % (sum(x[1] for x in v) / len(v), v[0][2])
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
