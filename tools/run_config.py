"""Run one BASELINE.json-style configuration end to end on files: generate the dataset (exact
canonical 40-mer counts, tools/cpsim.c), classify it with the reference binary (oracle/_ref/ClassPro,
if present) and with this repository's ClassPro CLI on the GPU(s), and compare the .class files byte
for byte.  Prints one JSON line with sizes, timings and the number of differing characters.

    python tools/run_config.py c1 [--scale 1.0] [--gpus N] [--keep DIR]

configs (BASELINE.json):
  c1  MHC-like 5 Mb diploid genome, 40x HiFi reads (~20 kb)          (the reference's own test/1-run.sh case)
  c2  100 Mb, 1 % het, 30x        -> scaled by --scale (exact counting of 3e9 k-mers needs ~100 GB)
  c4  repeat-rich 500 Mb, 40x     -> scaled by --scale
  c5  coverage / read-length sweep point: --cov, --len, with -c/-r given to both programs
"""
import argparse
import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit  # noqa: E402

CLI = os.path.join(ROOT, "classpro_b200", "ClassPro")

CONFIGS = {
    "c1": dict(genome_len=5_000_000, cov=40., het=0.005, seg_dups=4, len_mean=20000, len_sd=2000, nparts=4),
    "c2": dict(genome_len=100_000_000, cov=30., het=0.01, len_mean=20000, len_sd=2000, nparts=8),
    "c4": dict(genome_len=500_000_000, cov=40., het=0.005, repeat_frac=0.5, seg_dups=8, len_mean=20000, len_sd=2000,
               nparts=8),
    "c5": dict(genome_len=1_000_000_000, cov=30., het=0.005, len_mean=20000, len_sd=2000, nparts=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="genome length multiplier")
    ap.add_argument("--cov", type=float, default=None)
    ap.add_argument("--len", type=int, default=None)
    ap.add_argument("--fixed-c", action="store_true", help="pass -c<true D> -r<len> to both programs")
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0, help="reference -T (default: all cores, max 32)")
    ap.add_argument("--seed", type=int, default=101)
    ap.add_argument("--keep", default=None)
    args = ap.parse_args()

    p = dict(CONFIGS[args.config])
    p["genome_len"] = max(100000, int(p["genome_len"] * args.scale))
    if args.cov:
        p["cov"] = args.cov
    if args.len:
        p["len_mean"] = args.len
        p["len_sd"] = args.len // 10
    extra = []
    if args.fixed_c:
        extra = ["-c%d" % int(round(p["cov"])), "-r%d" % p["len_mean"]]
    tmp = args.keep or tempfile.mkdtemp(prefix="cprun_")
    os.makedirs(tmp, exist_ok=True)
    t0 = time.time()
    sim = cpkit.simulate(write_to=tmp, root="reads", seed=args.seed, **p)
    t_gen = time.time() - t0
    fasta = os.path.join(tmp, "reads.fasta")
    out = {"config": args.config, "genome_len": p["genome_len"], "cov": p["cov"], "reads": sim.nreads,
           "kmers": sim.total_kmers, "prof_bytes_per_kmer": round(len(sim.prof) / max(1, sim.total_kmers), 4),
           "gen_s": round(t_gen, 1), "args": extra}
    del sim

    ref_class = None
    if cpkit.have_reference():
        cores = len(os.sched_getaffinity(0))
        T = args.threads or min(cores, 32)
        t0 = time.time()
        pr = subprocess.run([cpkit.REF_BIN, "-v", "-T%d" % T] + extra + [fasta], cwd=tmp, stdout=subprocess.PIPE,
                            stderr=subprocess.PIPE, text=True)
        out["ref_s"] = round(time.time() - t0, 2)
        out["ref_threads"] = T
        if pr.returncode != 0:
            out["ref_error"] = pr.stderr[-300:]
        else:
            ref_class = os.path.join(tmp, "reads.ref.class")
            os.replace(os.path.join(tmp, "reads.class"), ref_class)
            out["ref_phase"] = [l for l in pr.stderr.splitlines() if l.startswith("Resources for phase")][-1:]
            out["ref_kmers_per_s"] = round(out["kmers"] / out["ref_s"])

    cmd = [CLI, "-v"] + extra + (["-G%d" % args.gpus] if args.gpus else []) + [fasta]
    t0 = time.time()
    pr = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out["gpu_s"] = round(time.time() - t0, 2)
    if pr.returncode != 0:
        out["gpu_error"] = pr.stderr[-500:]
        print(json.dumps(out))
        return 1
    out["gpu_kmers_per_s"] = round(out["kmers"] / out["gpu_s"])
    out["gpu_stderr_tail"] = pr.stderr.splitlines()[-5:]
    mine = os.path.join(tmp, "reads.class")
    if ref_class:
        same = filecmp.cmp(mine, ref_class, shallow=False)
        out["identical"] = same
        if not same:
            a, b = cpkit.class_lines(mine), cpkit.class_lines(ref_class)
            out["records"] = [len(a), len(b)]
            out["flipped_chars"] = sum(sum(1 for x, y in zip(u, v) if x != y) + abs(len(u) - len(v)) for u, v in zip(a, b))
            out["flip_fraction"] = out["flipped_chars"] / max(1, out["kmers"])
    print(json.dumps(out))
    if not args.keep:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
