"""File-level parity at sizes the host counter of tools/cpsim.c cannot reach: the reads of a BASELINE.json-style
configuration are generated chromosome by chromosome in parallel threads (no counting), the FastK files
(<root>.hist / .prof / .pidx) with the EXACT canonical 40-mer counts of the whole read set come from this
repository's `profiler` (cpg_count_kmers / cpg_encode_profiles on the GPU), and then the unmodified reference
binary and this repository's ClassPro classify the same files; the two .class files are compared byte for byte.

    python tools/run_big.py c2 --genome-mb 100          (BASELINE.json configs[1], full size)
    python tools/run_big.py c4 --genome-mb 50           (configs[3], repeat-rich, scaled)

One JSON line: sizes, the profiler's time, both wall times, identical / flipped characters."""
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit          # noqa: E402
import bench          # noqa: E402

CLI = os.environ.get("CPG_CLI") or os.path.join(ROOT, "classpro_b200", "ClassPro")
PROFILER = os.environ.get("CPG_PROFILER") or os.path.join(ROOT, "classpro_b200", "profiler")

SIM = {
    "c2": dict(cov=30., sim=dict(het=0.01, exact=2, len_mean=20000, len_sd=2000, len_min=5000, len_max=50000)),
    "c4": dict(cov=40., sim=dict(het=0.005, repeat_frac=0.5, seg_dups=8, exact=2, len_mean=20000, len_sd=2000,
                                 len_min=5000, len_max=50000)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(SIM))
    ap.add_argument("--genome-mb", type=float, default=50.)
    ap.add_argument("--chunk-mb", type=float, default=5.)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--keep", default=None)
    args = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    tmp = args.keep or tempfile.mkdtemp(prefix="cpbig_")
    os.makedirs(tmp, exist_ok=True)
    wl = SIM[args.config]
    nch = max(1, int(round(args.genome_mb / args.chunk_mb)))
    t0 = time.time()
    per = bench.gen_chunks(list(range(nch)), wl, args.chunk_mb, max(1, min(cores, 16)), write_to=tmp)
    fasta = os.path.join(tmp, "reads.fasta")
    with open(fasta, "wb") as fo:
        for c in range(nch):
            src = os.path.join(tmp, "c%d.fasta" % c)
            with open(src, "rb") as fi:
                shutil.copyfileobj(fi, fo, 1 << 24)
            os.remove(src)
    out = {"config": args.config, "genome_mb": args.genome_mb, "cov": wl["cov"], "reads": sum(p[0] for p in per.values()),
           "kmers": sum(p[1] for p in per.values()), "gen_s": round(time.time() - t0, 1),
           "profiles": "exact canonical 40-mer counts of the whole read set (classpro_b200/profiler, GPU)"}
    t0 = time.time()
    pr = subprocess.run([PROFILER, "-v", "-k40", "-p8", fasta], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out["profiler_s"] = round(time.time() - t0, 2)
    if pr.returncode != 0:
        out["profiler_error"] = pr.stderr[-500:]
        print(json.dumps(out))
        return 1
    out["profiler_stderr_tail"] = pr.stderr.splitlines()[-3:]
    ref_class = None
    if cpkit.have_reference():
        T = min(cores, 16)
        t0 = time.time()
        pr = subprocess.run([cpkit.REF_BIN, "-v", "-T%d" % T, fasta], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        out["ref_s"] = round(time.time() - t0, 2)
        out["ref_threads"] = T
        if pr.returncode != 0:
            out["ref_error"] = pr.stderr[-300:]
        else:
            ref_class = os.path.join(tmp, "reads.ref.class")
            os.replace(os.path.join(tmp, "reads.class"), ref_class)
            out["ref_phase"] = [l for l in pr.stderr.splitlines() if l.startswith("Resources for phase")][-1:]
    cmd = [CLI, "-v", "-T%d" % min(cores, 16)] + (["-G%d" % args.gpus] if args.gpus else []) + [fasta]
    t0 = time.time()
    pr = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    out["gpu_s"] = round(time.time() - t0, 2)
    if pr.returncode != 0:
        out["gpu_error"] = pr.stderr[-500:]
        print(json.dumps(out))
        return 1
    out["gpu_stderr_tail"] = pr.stderr.splitlines()[-4:]
    if ref_class:
        mine = os.path.join(tmp, "reads.class")
        same = filecmp.cmp(mine, ref_class, shallow=False)
        out["identical"] = same
        out["wall_ratio_ref_over_gpu"] = round(out["ref_s"] / out["gpu_s"], 2)
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
