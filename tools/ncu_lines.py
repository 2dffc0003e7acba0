"""Summarise `ncu --page source --csv --print-source cuda,sass` by CUDA source line / function.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_classify > src.csv
    python tools/ncu_lines.py src.csv [top]
"""
import csv
import collections
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = csv.reader(open(path, newline=""))
    cur_file = None
    hdr = None
    lines = []        # (file, line, text, samples, inst, thread_inst)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or cur_file is None or r[0] == "" or r[0] == "Function Name":
            continue
        try:
            ln = int(r[0])
            smp = int(r[hdr["# Samples"]] or 0)
            ins = int(r[hdr["Instructions Executed"]] or 0)
            thr = int(r[hdr["Thread Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        lines.append((cur_file, ln, r[1].strip(), smp, ins, thr))
    tot_s = sum(x[3] for x in lines) or 1
    tot_i = sum(x[4] for x in lines) or 1
    print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
    byfile = collections.Counter()
    byfile_i = collections.Counter()
    for f, ln, t, s, i, th in lines:
        byfile[f] += s
        byfile_i[f] += i
    print("\nper file: samples%  inst%")
    for f, s in byfile.most_common():
        print("  %-22s %5.1f%%  %5.1f%%" % (f, 100. * s / tot_s, 100. * byfile_i[f] / tot_i))
    print("\ntop lines by stall samples:")
    for f, ln, t, s, i, th in sorted(lines, key=lambda x: -x[3])[:top]:
        print("  %5.1f%% s %5.1f%% i  thr %4.1f  %s:%d  %s" % (100. * s / tot_s, 100. * i / tot_i, th / max(i, 1), f, ln, t[:90]))


if __name__ == "__main__":
    main()
