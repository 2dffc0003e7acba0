"""Turn an ncu report into the markdown summary committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r01_xxx.md
"""
import collections
import csv
import io
import subprocess
import sys

RAW = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps/SM)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg active threads per warp instruction"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", "avg active threads per warp instruction (%)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe utilisation"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps per scheduler cycle"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
    ("smsp__sass_branch_targets_threads_divergent.sum", "divergent branch targets"),
    ("sm__sass_branch_targets_threads_uniform.pct", "uniform branch targets"),
    ("smsp__sass_average_branch_targets_threads_uniform.pct", "branch efficiency"),
]


def ncu(*args):
    return subprocess.run(["ncu"] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` (cold caches, serialised "
          "launches: compare shares, not absolutes).\n")
    kernels = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?").split("(")[0]
        kernels.append(name)
        print("## %s\n" % name)
        print("| metric | value |")
        print("|---|---|")
        for key, label in RAW:
            if key in d and d[key] != "":
                print("| %s (`%s`) | %s %s |" % (label, key, d[key], units[hdr.index(key)]))
        print()
    for name in sorted(set(kernels)):
        src = ncu("-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + name)
        cur, h, lines, stalls = None, None, [], collections.Counter()
        for r in csv.reader(io.StringIO(src)):
            if not r:
                continue
            if r[0] == "File Path":
                cur = r[1].split("/")[-1]
                continue
            if r[0] == "Line No":
                h = {x: i for i, x in enumerate(r)}
                hl = r
                continue
            if h is None or cur is None or r[0] in ("", "Function Name"):
                continue
            try:
                ln = int(r[0])
                smp = int(r[h["# Samples"]] or 0)
                ins = int(r[h["Instructions Executed"]] or 0)
                thr = int(r[h["Thread Instructions Executed"]] or 0)
            except (ValueError, IndexError):
                continue
            lines.append((cur, ln, r[1].strip(), smp, ins, thr))
            for i, x in enumerate(hl):
                if x.startswith("stall_") and "Not Issued" not in x:
                    try:
                        stalls[x] += int(r[i] or 0)
                    except ValueError:
                        pass
        if not lines:
            continue
        ts, ti = sum(x[3] for x in lines) or 1, sum(x[4] for x in lines) or 1
        print("## %s: warp stall reasons (sampled)\n" % name)
        tot = sum(stalls.values()) or 1
        print(", ".join("%s %.1f%%" % (k[6:], 100. * v / tot) for k, v in stalls.most_common(8)))
        print("\n## %s: where the samples and instructions are, by source file\n" % name)
        bf, bi = collections.Counter(), collections.Counter()
        for f, ln, t, s, i, th in lines:
            bf[f] += s
            bi[f] += i
        print("| file | stall samples | warp instructions |")
        print("|---|---|---|")
        for f, s in bf.most_common(8):
            print("| %s | %.1f%% | %.1f%% |" % (f, 100. * s / ts, 100. * bi[f] / ti))
        print("\n## %s: hottest source lines\n" % name)
        print("| samples | instr | active thr | line | source |")
        print("|---|---|---|---|---|")
        for f, ln, t, s, i, th in sorted(lines, key=lambda x: -x[3])[:14]:
            print("| %.1f%% | %.1f%% | %.1f | %s:%d | `%s` |" % (100. * s / ts, 100. * i / ti, th / max(i, 1), f, ln,
                                                                 t[:80].replace("|", "\\|")))
        print()


if __name__ == "__main__":
    main()
