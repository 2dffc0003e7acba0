#!/usr/bin/env python
"""bench.py -- classified k-mers/second of the ClassPro classification path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step is one pass of the hot path (profile decode -> walls -> reliable DP -> unreliable intervals
-> class strings) over the rank's shard of ONE global synthetic read set.  At N=1 that set is
BASELINE.json configs[1]: 100 Mb synthetic diploid genome, 1 % heterozygosity, 30x HiFi-like reads
(~20 kb), k=40.  At N GPUs the global set is N times that (weak scaling: 8 x 100 Mb = 27 % of the
human-scale configs[2]); it is cut into contiguous read ranges balanced by cumulative compressed-
profile bytes (classpro_b200/shard.py), every rank generates only the chromosomes its range
touches, the model comes from the histogram of the whole set, and no rank ever needs another rank's
reads: the only collectives are the all-gather of the per-read profile sizes the ranges are cut from,
the sum of the histograms and the max-reduction of the timings.  --scaling strong keeps the global
set at --genome-mb whatever N is.

  value  k-mers/s with the shard already resident in HBM (kernels only, CUDA events)
  e2e    k-mers/s through cpg_submit/cpg_collect with pinned HOST buffers: H2D copies of the packed
         reads + compressed profiles and the D2H copy of the result are inside the timed region
         (double buffered over two streams).  It starts from parsed, 2-bit packed reads and fetched
         profile bytes; the file-to-file program is measured separately:
  cli    (N=1) the ClassPro program of this repository, file to file, on the FASTA + FastK files of
         the same read set the reference arm reads: wall clock of the whole process
  parity_sample   a seeded sample of the bench reads classified by the oracle (oracle/) and compared
         with the GPU result; the run FAILS above 1e-6 flipped k-mers

--impl reference: the unmodified reference program (oracle/_ref/ClassPro -T<cores, max 16>) on the
files of the same read set, every step one whole run of the program (start-up, per-thread setup and
output included).  The first run is the warm-up and the calibration: if K runs of the full set do not
fit the time budget (--ref-budget-s) the steps use the largest prefix of the chromosomes that does, and
the line says so.
"""
import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = 40
FLIP_BUDGET = 1e-6

WORKLOADS = {
    # BASELINE.json configs[1]; profiles from the simulator's ground-truth coverage (tools/cpsim.c --fast)
    "c2": dict(name="100 Mb synthetic diploid genome, 1% heterozygosity, 30x HiFi-like reads (~20 kb), k=40 "
                    "(BASELINE.json configs[1])",
               genome_mb=100., cov=30.,
               sim=dict(het=0.01, snp_only=1, exact=0, len_mean=20000, len_sd=2000, len_min=5000, len_max=50000),
               profiles="ground-truth coverage (tools/cpsim.c --fast)"),
    # BASELINE.json configs[3] scaled to what one GPU counts in one pass: repeat-rich genome, reads with
    # substitution / homopolymer-indel errors, EXACT canonical 40-mer counts of the read set from the
    # profile producer of this repository (cpg_count_kmers / cpg_encode_profiles on the GPU)
    "c4": dict(name="repeat-rich synthetic diploid genome (50% tandem + interspersed repeats, low-complexity runs), 40x "
                    "HiFi-like reads (~20 kb), k=40 (BASELINE.json configs[3], scaled)",
               genome_mb=50., cov=40.,
               sim=dict(het=0.005, repeat_frac=0.5, seg_dups=8, exact=2, len_mean=20000, len_sd=2000, len_min=5000,
                        len_max=50000),
               profiles="exact canonical 40-mer counts of the read set (GPU profile producer, cpg_count_kmers)"),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- workload
def chunk_seed(c):
    return 1001 + c            # chromosome c of the global read set, whatever rank generates it


def gen_chunks(chunks, wl, chunk_mb, threads, write_to=None):
    """Chromosomes `chunks` (global indices) of the read set, generated in parallel threads (the C
    generator releases the GIL).  With write_to: FASTA + FastK files <write_to>/c<index>.* are written
    and only (reads, k-mers) of each chromosome is kept."""
    import cpkit
    out = {}

    def one(c):
        kw = dict(seed=chunk_seed(c), genome_len=int(chunk_mb * 1e6), cov=wl["cov"], nparts=1)
        kw.update(wl["sim"])
        if write_to is not None:
            s = cpkit.simulate(write_to=write_to, root="c%d" % c, **kw)
            out[c] = (s.nreads, int(np.maximum(s.rlen.astype(np.int64) - K + 1, 0).sum()))
        else:
            out[c] = cpkit.simulate(**kw)

    pending = list(chunks)
    while pending:
        now, pending = pending[:threads], pending[threads:]
        ths = [threading.Thread(target=one, args=(c,)) for c in now]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    return out


def join_files(tmp, chunks, root="reads"):
    """One FASTA and one FastK file set out of the per-chromosome files gen_chunks wrote: chromosome c
    becomes profile part number position+1 (FastK parts are contiguous read ranges)."""
    fasta = os.path.join(tmp, root + ".fasta")
    hist = None
    first = 0
    with open(fasta, "wb") as fo:
        for p, c in enumerate(chunks):
            src = os.path.join(tmp, "c%d.fasta" % c)
            with open(src, "rb") as fi:
                shutil.copyfileobj(fi, fo, 1 << 24)
            os.remove(src)
            with open(os.path.join(tmp, "c%d.hist" % c), "rb") as f:
                head = f.read(12)
                il, ih = struct.unpack("<qq", f.read(16))
                h = np.frombuffer(f.read(), dtype=np.int64)
            if hist is None:
                hist = [head, il, ih, h.copy()]
            else:
                hist[1] += il
                hist[2] += ih
                hist[3] = hist[3] + h
            os.remove(os.path.join(tmp, "c%d.hist" % c))
            os.remove(os.path.join(tmp, "c%d.prof" % c))
            os.replace(os.path.join(tmp, ".c%d.prof.1" % c), os.path.join(tmp, ".%s.prof.%d" % (root, p + 1)))
            with open(os.path.join(tmp, ".c%d.pidx.1" % c), "rb") as f:
                k, _, n = struct.unpack("<iqq", f.read(20))
                idx = f.read()
            with open(os.path.join(tmp, ".%s.pidx.%d" % (root, p + 1)), "wb") as f:
                f.write(struct.pack("<iqq", k, first, n))
                f.write(idx)
            os.remove(os.path.join(tmp, ".c%d.pidx.1" % c))
            first += n
    with open(os.path.join(tmp, root + ".hist"), "wb") as f:
        f.write(hist[0])
        f.write(struct.pack("<qq", hist[1], hist[2]))
        f.write(hist[3].tobytes())
    with open(os.path.join(tmp, root + ".prof"), "wb") as f:
        f.write(struct.pack("<ii", K, len(chunks)))
    return fasta, first


class HostData:
    """The rank's shard in pinned host memory, plus per-batch views."""

    def __init__(self, parts, n_batches, hist):
        """parts: list of (sim, first_read, end_read) in global read order."""
        from classpro_b200.abi import PinnedArray, pack_codes
        import classpro_b200 as cp
        rlen = np.concatenate([s.rlen[a:b] for s, a, b in parts]).astype(np.int32)
        n = len(rlen)
        self.n_reads = n
        self.kmers = int(np.maximum(rlen.astype(np.int64) - K + 1, 0).sum())
        self.bases = int(rlen.astype(np.int64).sum())
        self.hist = hist
        poff = np.zeros(n + 1, np.int64)
        np.cumsum((rlen.astype(np.int64) + 3) // 4, out=poff[1:])
        self.pin_seq = PinnedArray(int(poff[-1]) + 64)
        seq = self.pin_seq.u8
        at = 0
        for s, a, b in parts:
            so = s.seq_off[a:b + 1] - s.seq_off[a]
            pk, po = pack_codes(s.seq[s.seq_off[a]:s.seq_off[b]], so, s.rlen[a:b])
            seq[at:at + int(po[-1])] = pk[:int(po[-1])]
            at += int(po[-1])
        self.rlen, self.seq_off = rlen, poff
        self.seq_bytes = int(poff[-1])
        self.pin_prof = None
        self.pin_cls = PinnedArray(self.bases + 64)
        self._n_batches = n_batches
        self._cp = cp
        if all(len(s.prof) or s.nreads == 0 for s, a, b in parts) and not any(s.params.exact == 2 for s, a, b in parts):
            prof_len = np.concatenate([np.diff(s.prof_off[a:b + 1]) for s, a, b in parts])
            pro = np.zeros(n + 1, np.int64)
            np.cumsum(prof_len, out=pro[1:])
            self.set_profiles(np.concatenate([s.prof[s.prof_off[a]:s.prof_off[b]] for s, a, b in parts]), pro)

    def set_profiles(self, prof_bytes, pro):
        from classpro_b200.abi import PinnedArray
        cp = self._cp
        n, rlen, poff, seq = self.n_reads, self.rlen, self.seq_off, self.pin_seq.u8
        self.pin_prof = PinnedArray(int(pro[-1]) + 64)
        prof = self.pin_prof.u8
        prof[:int(pro[-1])] = prof_bytes[:int(pro[-1])]
        self.prof_bytes = int(pro[-1])
        self.prof_off = pro
        self.whole = cp.Batch(seq, poff, rlen, prof, pro, 2)
        cum = np.cumsum(rlen.astype(np.int64))
        cuts = [0]
        for b in range(1, self._n_batches):
            cuts.append(int(np.searchsorted(cum, self.bases * b // self._n_batches)))
        cuts.append(n)
        self.batches = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            so = poff[a:b + 1] - poff[a]
            po = pro[a:b + 1] - pro[a]
            bt = cp.Batch(seq[poff[a]:poff[b] + 16], so, rlen[a:b], prof[pro[a]:pro[b] + 16], po, 2)
            c0 = int(cum[a - 1]) if a > 0 else 0
            c1 = int(cum[b - 1])
            self.batches.append((bt, self.pin_cls.u8[c0:c1 + 1]))
        self.h2d_bytes = self.seq_bytes + self.prof_bytes + n * (8 + 8 + 4 + 8 + 8 + 4)
        self.d2h_bytes = self.bases + 4 * n

    def read_ascii(self, i):
        a, r = int(self.seq_off[i]), int(self.rlen[i])
        pk = self.pin_seq.u8[a:a + (r + 3) // 4]
        codes = np.empty(4 * len(pk), np.uint8)
        codes[0::4], codes[1::4], codes[2::4], codes[3::4] = pk & 3, (pk >> 2) & 3, (pk >> 4) & 3, pk >> 6
        return np.frombuffer(b"ACGT", dtype=np.uint8)[codes[:r]].tobytes()

    def free(self):
        self.whole = None
        self.batches = []
        for p in (self.pin_seq, self.pin_prof, self.pin_cls):
            if p is not None:
                p.free()


def parity_sample(data, cls, kmer_target, read_target, seed, threads):
    """Classify a seeded sample of the shard's reads with the oracle (counts from the oracle's own
    decoder of the compressed profile) and count the characters that differ from the GPU's."""
    import cpkit

    class _H:
        pass
    h = _H()
    h.hist, h.kmer = data.hist, K
    om = cpkit.oracle_model(h, 0, 20000)
    rng = np.random.default_rng(seed)
    order = rng.permutation(data.n_reads)
    picked, km = [], 0
    for i in order:
        if data.rlen[i] < K:
            continue
        picked.append(int(i))
        km += int(data.rlen[i]) - K + 1
        if len(picked) >= read_target and km >= kmer_target:
            break
    cls_off = np.zeros(data.n_reads + 1, np.int64)
    np.cumsum(data.rlen.astype(np.int64), out=cls_off[1:])
    res = [0, 0, []]
    lock = threading.Lock()

    def work(ids):
        ow = cpkit.OracleWork(clean=True)
        for i in ids:
            cap = int(data.rlen[i]) - K + 1
            n, counts = cpkit.oracle_decode(data.pin_prof.u8[data.prof_off[i]:data.prof_off[i + 1]], cap)
            assert n == cap
            a = ow.classify(om, data.read_ascii(i), counts)
            b = cls[cls_off[i]:cls_off[i + 1]].tobytes()
            f = 0
            if a != b:
                f = int((np.frombuffer(a, np.uint8) != np.frombuffer(b, np.uint8)).sum())
            with lock:
                res[0] += cap
                res[1] += f
                if f:
                    res[2].append(i)

    ths = [threading.Thread(target=work, args=(picked[t::threads],)) for t in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return {"reads": len(picked), "kmers": res[0], "flips": res[1], "flipped_reads": sorted(res[2])[:16]}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms
def run_reference_once(fasta, kmers, threads, extra=()):
    """One whole run of the reference's CPU implementation on the files.  (k-mers/s, kind, cores, detail)."""
    import cpkit
    if cpkit.have_reference():
        t0 = time.time()
        p = subprocess.run([cpkit.REF_BIN, "-v", "-T%d" % threads] + list(extra) + [fasta], cwd=os.path.dirname(fasta),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("reference ClassPro failed: " + p.stderr[-500:])
        phase = total = None
        for line in p.stderr.splitlines():
            if line.startswith("Resources for phase:"):
                phase = line
            if line.startswith("Total Resources:"):
                total = line
        return kmers / wall, "reference", threads, {"wall_s": round(wall, 3), "phase_line": phase, "total_line": total}
    # reference not built here: the oracle port, single thread
    out = fasta + ".oracle.class"
    t0 = time.time()
    rc = cpkit.oracle_lib().cpo_run_file(fasta.encode(), fasta[:-len(".fasta")].encode(), 0, 20000, out.encode(), 0, None)
    wall = time.time() - t0
    if rc != 0:
        raise RuntimeError("oracle port failed rc=%d" % rc)
    return kmers / wall, "port", 1, {"wall_s": round(wall, 3)}


def user_seconds(line):
    """'Resources for phase:  12.3 (s.ms) user ...' -> 12.3 (src/benchmark.c formats; minutes / hours forms too)."""
    if not line:
        return None
    try:
        tok = line.split(":", 1)[1].split()
        v = tok[0]
        parts = [float(x) for x in v.split(":")]
        s = 0.
        for x in parts:
            s = s * 60. + x
        return s
    except Exception:
        return None


def dataset_files(tmp, wl, nchunks, chunk_mb, threads, root="reads"):
    sims = gen_chunks(range(nchunks), wl, chunk_mb, threads, write_to=tmp)
    per = [sims[c] for c in range(nchunks)]
    fasta, nreads = join_files(tmp, list(range(nchunks)), root)
    return fasta, per


def workload_config(wl, args, world, nchunks_total, reads, kmers):
    """The same dict in both arms: what is classified."""
    return {"workload": wl["name"], "profiles": wl["profiles"], "genome_mb_total": round(nchunks_total * args.chunk_mb, 1),
            "chromosome_mb": args.chunk_mb, "coverage": wl["cov"], "reads": reads, "kmers": kmers,
            "sharding": "contiguous read ranges balanced by cumulative compressed-profile bytes" if world > 1 else "one GPU"}


def reference_arm(args, rank, world, wl):
    if rank != 0:
        return
    if wl["sim"].get("exact") == 2:
        print(json.dumps({"impl": "reference", "unavailable": "workload %s takes its profiles from the GPU producer; "
                          "the reference arm runs the default workload" % args.workload}))
        return
    cores = host_cores()
    threads = max(1, min(cores, args.cpu_threads if args.cpu_threads > 0 else 16))
    t_start = time.time()
    nchunks = max(1, int(round(args.genome_mb / args.chunk_mb)))          # the N=1 read set: what one GPU classifies per step
    with tempfile.TemporaryDirectory(prefix="cpbench_") as tmp:
        full = os.path.join(tmp, "full")
        os.makedirs(full)
        fasta, per = dataset_files(full, wl, nchunks, args.chunk_mb, min(cores, 16))
        reads_full, kmers_full = sum(p[0] for p in per), sum(p[1] for p in per)
        t_gen = time.time() - t_start
        # warm-up + calibration: one run of the full read set
        v_full, kind, used, d_full = run_reference_once(fasta, kmers_full, threads)
        t_full = d_full["wall_s"]
        left = args.ref_budget_s - (time.time() - t_start)
        use_chunks = nchunks
        if kind == "reference" and args.steps * t_full > left:
            # the largest prefix of the chromosomes whose K runs fit (run time ~ fixed setup + rate * k-mers):
            # second calibration point on one chromosome
            one = os.path.join(tmp, "one")
            os.makedirs(one)
            f1, p1 = dataset_files(one, wl, 1, args.chunk_mb, 1)
            _, _, _, d1 = run_reference_once(f1, p1[0][1], threads)
            rate = max(1e-9, (t_full - d1["wall_s"]) / max(1, kmers_full - p1[0][1]))
            setup = max(0., d1["wall_s"] - rate * p1[0][1])
            left = args.ref_budget_s - (time.time() - t_start) - 8.
            per_step = max(1e-3, left / args.steps)
            km = max(0., (per_step - setup) / rate)
            use_chunks = int(max(1, min(nchunks, km // max(1, kmers_full // nchunks))))
            shutil.rmtree(one, ignore_errors=True)
        t1_detail = None
        if use_chunks < nchunks:
            shutil.rmtree(full, ignore_errors=True)
            os.makedirs(full)
            fasta, per = dataset_files(full, wl, use_chunks, args.chunk_mb, min(cores, 16))
        reads, kmers = sum(p[0] for p in per), sum(p[1] for p in per)
        t0 = time.time()
        detail = d_full
        for _ in range(args.steps):
            v, kind, used, detail = run_reference_once(fasta, kmers, threads)
        dt = time.time() - t0
        # per-core figure: -T1 on one chromosome, user time of its classification phase
        try:
            if kind == "reference" and args.ref_budget_s - (time.time() - t_start) > 60:
                one = os.path.join(tmp, "t1")
                os.makedirs(one)
                f1, p1 = dataset_files(one, wl, 1, min(args.chunk_mb, 2.), 1)
                v1, _, _, d1 = run_reference_once(f1, p1[0][1], 1)
                us = user_seconds(d1.get("phase_line"))
                t1_detail = {"kmers": p1[0][1], "wall_s": d1["wall_s"], "phase_user_s": us,
                             "kmers_per_user_s": (p1[0][1] / us) if us else None, "phase_line": d1.get("phase_line")}
        except Exception as e:
            t1_detail = {"error": str(e)[:200]}
    value = kmers * args.steps / dt
    same = use_chunks == nchunks
    cfg = workload_config(wl, args, 1, nchunks, reads_full, kmers_full)
    sample = ("the whole read set per step (%d reads, %d k-mers)" % (reads, kmers)) if same else \
             ("the first %d of %d chromosomes per step (%d reads, %d k-mers): %d runs of the whole set (%.1f s each) do not fit "
              "the %d s budget" % (use_chunks, nchunks, reads, kmers, args.steps, t_full, args.ref_budget_s))
    line = {"impl": "reference", "metric": "classified k-mers/sec", "value": value, "unit": "k-mers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": used, "kind": kind,
                             "sample": "ClassPro -T%d, whole program runs (start-up, per-thread setup, output included) on %s; "
                                       "1 warm-up run (the full set)" % (used, sample),
                             "same_read_set_as_gpu_arm": same,
                             "full_set_run": {"reads": reads_full, "kmers": kmers_full, "wall_s": t_full, "value": v_full,
                                              "phase_line": d_full.get("phase_line"), "total_line": d_full.get("total_line")},
                             "last_step": detail, "t1": t1_detail, "gen_seconds": round(t_gen, 1)},
            "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": cores}
    print(json.dumps(line))


def cli_file_to_file(wl, args, nchunks, cores):
    """The ClassPro program of this repository on the files of the N=1 read set: wall clock of the whole process."""
    cli = os.path.join(ROOT, "classpro_b200", "ClassPro")
    threads = max(1, min(cores, 16))
    with tempfile.TemporaryDirectory(prefix="cpcli_") as tmp:
        fasta, per = dataset_files(tmp, wl, nchunks, args.chunk_mb, min(cores, 16))
        reads, kmers = sum(p[0] for p in per), sum(p[1] for p in per)
        best = None
        for _ in range(2):                      # the second run has the CUDA driver's caches and the page cache warm
            t0 = time.time()
            p = subprocess.run([cli, "-v", "-T%d" % threads, "-G1", fasta], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                               text=True)
            wall = time.time() - t0
            if p.returncode != 0:
                raise RuntimeError("ClassPro failed: " + p.stderr[-400:])
            if best is None or wall < best[0]:
                best = (wall, [l for l in p.stderr.splitlines() if "timeline" in l or "stage seconds" in l])
        size = os.path.getsize(fasta[:-len(".fasta")] + ".class")
    return {"value": kmers / best[0], "unit": "k-mers/s", "wall_s": round(best[0], 3), "reads": reads, "kmers": kmers,
            "threads": threads, "class_bytes": size, "what": "classpro_b200/ClassPro -T%d -G1 <fasta>, whole process, best of 2, on the "
            "same FASTA + FastK files the reference arm reads" % threads, "stages": best[1]}


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--genome-mb", type=float, default=0., help="per GPU (weak) or in all (strong); default: the workload's")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cov", type=float, default=0.)
    ap.add_argument("--chunk-mb", type=float, default=5.)
    ap.add_argument("--gen-threads", type=int, default=0)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--cpu-sample-mbases", type=float, default=360.)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true")
    ap.add_argument("--parity-reads", type=int, default=512)
    ap.add_argument("--parity-kmers", type=float, default=1.2e7)
    ap.add_argument("--ref-budget-s", type=float, default=600.)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.genome_mb <= 0:
        args.genome_mb = wl["genome_mb"]
    if args.cov > 0:
        wl["cov"] = args.cov

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world, wl)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import classpro_b200 as cp
    from classpro_b200 import abi
    from classpro_b200.shard import plan_chunk_shards
    cores = host_cores()
    gen_threads = args.gen_threads or max(1, min(16, cores // max(1, world)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def min_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MIN if world > 1 else None)

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM if world > 1 else None)

    # ---- the global read set and this rank's shard of it
    t_gen = time.time()
    per_gpu_chunks = max(1, int(round(args.genome_mb / args.chunk_mb)))
    nchunks_total = per_gpu_chunks * world if args.scaling == "weak" else max(world, per_gpu_chunks)
    home = list(range(nchunks_total * rank // world, nchunks_total * (rank + 1) // world))
    sims = gen_chunks(home, wl, args.chunk_mb, gen_threads)
    hist = np.sum([sims[c].hist for c in home], axis=0)
    # per-chromosome read counts and per-read profile sizes of the whole set (the weights the ranges are cut from);
    # reads-only workloads (profiles counted later) are cut by read length instead
    def weights(s):
        return np.diff(s.prof_off) if s.params.exact != 2 else s.rlen.astype(np.int64)
    mine = [(c, weights(sims[c])) for c in home]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        allw = [cw for g in gathered for cw in g]
        ht = torch.from_numpy(hist.astype(np.int64)).cuda()
        dist.all_reduce(ht)
        hist = ht.cpu().numpy()
    else:
        allw = mine
    allw.sort(key=lambda cw: cw[0])
    plans, reads_total = plan_chunk_shards(allw, world)
    beg, end, need = plans[rank]
    extra = [c for c, a, b in need if c not in sims]
    if extra:                                    # a range border inside a neighbour's chromosome: generate it here too
        sims.update(gen_chunks(extra, wl, args.chunk_mb, gen_threads))
    parts = [(sims[c], a, b) for c, a, b in need]
    n_generated = len(sims)
    data = HostData(parts, args.batches, hist)
    producer = None
    if data.pin_prof is None:
        # reads-only workload: exact counts + FastK profiles from the GPU producer (one call over the whole shard)
        t_p = time.time()
        counts, cnt_off, phist = abi.count_kmers(K, data.pin_seq.u8[:data.seq_bytes + 32], data.seq_off, data.rlen, local)
        t_c = time.time() - t_p
        prof, pro = abi.encode_profiles(counts, cnt_off, local)
        producer = {"count_s": round(t_c, 2), "encode_s": round(time.time() - t_p - t_c, 2), "kmers": int(cnt_off[-1])}
        del counts
        if world > 1:
            raise SystemExit("workload %s is single-GPU (its counts are those of the rank's own reads)" % args.workload)
        data.hist = hist = phist
        data.set_profiles(prof, pro)
        del prof
    del sims, parts
    t_gen = time.time() - t_gen
    model = cp.Model.from_hist(K, hist[1:32768], hist[32768], hist[32769], read_len=20000)
    ctx = cp.Context(model, local)

    # ---- resident: whole shard in HBM, kernels only
    ctx.upload(data.whole)
    ctx.run_resident(max(args.warmup, 3))
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    t0 = time.time()
    ms_dec, ms_cls, launches = ctx.run_resident(args.steps)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    phase = ctx.phase_cycles()
    wall_ns = ctx.wall_ns()
    n_cand, n_ivl, n_rel, n_vis = ctx.batch_stats()
    cls_res, status = ctx.download(data.whole)
    n_bad = int(sum_over_ranks(float((status & cp.ST_FATAL != 0).sum())))
    my_ms = ms_dec + ms_cls
    step_ms = max_over_ranks(my_ms)
    step_ms_min = min_over_ranks(my_ms)
    total_kmers = sum_over_ranks(float(data.kmers))
    value = total_kmers / (step_ms * 1e-3)

    # ---- end to end: pinned host buffers in, results out, double buffered.  The result of a batch is its
    # interval tables (CPG_RESULT_INTERVALS: 4 bytes per interval instead of 1 byte per base; the class
    # strings are a host-side expansion, cpg_expand_intervals, done where the characters are written --
    # in ClassPro's output formatter).  The class-string form of the same pass is timed beside it.
    def e2e_pass(intervals, ivres):
        inflight = []

        def fin(s, i, b, o):
            if intervals:
                if ivres[i] is None:
                    ivres[i] = abi.IntervalResult(b.n, ctx.intervals_bound(s), pinned=True)
                ctx.collect_intervals(s, ivres[i])
            else:
                ctx.collect(s, b, o)

        for i, (bt, out) in enumerate(data.batches):
            slot = i & 1
            if len(inflight) == 2:
                fin(*inflight.pop(0))
            ctx.submit(slot, bt)
            inflight.append((slot, i, bt, out))
        for x in inflight:
            fin(*x)

    def e2e_time(intervals, ivres):
        ctx.set_result_mode(intervals)
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_pass(intervals, ivres)
        barrier()
        t0 = time.time()
        for _ in range(args.steps):
            e2e_pass(intervals, ivres)
        barrier()
        return max_over_ranks((time.time() - t0) / args.steps)

    e2e_cls_s = e2e_time(False, None)
    same_cls = bool(np.array_equal(data.pin_cls.u8[:data.bases], cls_res[:data.bases]))
    ivres = [None] * len(data.batches)
    e2e_s = e2e_time(True, ivres)
    ctx.set_result_mode(False)
    e2e_value = total_kmers / e2e_s
    d2h_ivl = 0
    same = same_cls
    at = 0
    for (bt, out), rv in zip(data.batches, ivres):       # outside the timed region: expand and compare with the resident result
        d2h_ivl += 4 * rv.used + 16 * bt.n
        ex = rv.expand(K, bt.rlen)
        nb = int(bt.cls_off[-1])
        same = same and bool(np.array_equal(ex[:nb], cls_res[at:at + nb]))
        at += nb
        rv.free()
    # every collective is called by EVERY rank, outside the rank-0 block that prints the line (an all-reduce
    # inside it once cost a 10-minute NCCL timeout at N = 8)
    d2h_ivl_all = int(sum_over_ranks(float(d2h_ivl)))
    h2d_all = int(sum_over_ranks(float(data.h2d_bytes)))
    d2h_cls_all = int(sum_over_ranks(float(data.d2h_bytes)))
    same = bool(min_over_ranks(1.0 if same else 0.0) > 0.5)

    # ---- parity: a seeded sample of this shard's reads through the oracle
    par = parity_sample(data, cls_res, args.parity_kmers / world, max(16, args.parity_reads // world), 4242 + rank,
                        max(1, min(8, cores // max(1, world))))
    par_tot = {"reads": int(sum_over_ranks(float(par["reads"]))), "kmers": int(sum_over_ranks(float(par["kmers"]))),
               "flips": int(sum_over_ranks(float(par["flips"])))}
    par_tot["flip_fraction"] = par_tot["flips"] / max(1, par_tot["kmers"])
    par_tot["budget"] = FLIP_BUDGET
    par_tot["checker"] = "oracle/classpro_oracle.c (pinned to the unmodified reference, tests/test_oracle.py) on a seeded sample of the bench reads"
    if par["flipped_reads"]:
        par_tot["flipped_reads_rank%d" % rank] = par["flipped_reads"]

    # ---- roofline of the dominant kernel (+ the streaming decode kernel)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n, r, c = data.kmers, data.bases, data.prof_bytes
    # algorithmic bytes.  k_decode is decode + candidate scan fused: compressed
    # bytes in, counts (2 B/k-mer) and the candidate bit map (1 bit/k-mer) out.  The classification
    # kernels read the counts, the bit map and the 2-bit bases and write one class byte per base.
    bytes_dec = c + 2 * n + n // 8
    bytes_cls = 2 * n + n // 8 + (r + 3) // 4 + r
    tot_ns = max(1, sum(phase))
    ms_ph = [ms_cls * x / tot_ns for x in phase]
    traffic_file = "traffic_r02.json" if args.workload == "c2" else "traffic_r02_c4.json"
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
    except Exception:
        traffic = {}
    if not (world == 1 and args.genome_mb == 0.):
        traffic = {}                                    # measured on the default size of the workload only

    def sub(total_ms, parts, i):
        return total_ms * parts[i] / max(1, sum(parts))

    def K_(ms, nbytes, traf, note):
        d = {"ms": ms, "bytes": int(nbytes), "GBps": nbytes / (max(ms, 1e-6) * 1e-3) / 1e9,
             "frac": nbytes / (max(ms, 1e-6) * 1e-3) / 1e9 / peak, "traffic": traf, "note": note}
        return d

    # algorithmic bytes per launch: what a kernel must read and write, from the units
    # it works on -- counts n, bases r, compressed bytes c, candidates, intervals, reliable / visited intervals
    w3, u3 = wall_ns[:3], wall_ns[3:]
    kern = {
        "k_decode": K_(ms_dec, bytes_dec, traffic.get("k_decode"),
                       "profile decode + wall-candidate scan fused: c + 2n + n/8 bytes; issue bound"),
        "k_wall_a": K_(sub(ms_ph[0], w3, 0), n // 8 + n_cand * (4 + 8 + 16) + (n_cand // 5) * 216, traffic.get("k_wall_a"),
                       "pure, one wall candidate per lane: candidate bit map n/8, per candidate 2 counts + a 64-bit window of "
                       "bases in, a 16-byte header out, a 216-byte record for ~1 in 5 (estimate)"),
        "k_wall_b": K_(sub(ms_ph[0], w3, 1), n_cand * 16 + (n_cand // 5) * 216 + n_ivl * 48, traffic.get("k_wall_b"),
                       "order-dependent replay per read: candidate records in, 48-byte intervals out; DRAM-latency bound"),
        "k_wall_c": K_(sub(ms_ph[0], w3, 2), n_ivl * (96 + 8 * (K - 1)) + n_rel * 48, traffic.get("k_wall_c"),
                       "pure, one interval per lane: interval in and out, 4(K-1) counts, reliable ones copied"),
        "k_rel": K_(ms_ph[1], n_rel * (48 + 48 + 1) + n_ivl, traffic.get("k_rel"),
                    "reliable-interval DP: reliable intervals in, working copies, class codes out; FP64-issue bound "
                    "(44 % of its instructions are Bessel recurrences at 10 active lanes, FP64 pipe 46 %, profiles/r02_ncu_full.md)"),
        "k_unrel_a": K_(sub(ms_ph[2], u3, 0), n_ivl * 48 + n_vis * 104, traffic.get("k_unrel_a"),
                        "pure, one visited interval per lane: intervals in, 104-byte task records out"),
        "k_unrel_b": K_(sub(ms_ph[2], u3, 1), n_ivl * 48 + n_vis * 104, traffic.get("k_unrel_b"),
                        "the two sweeps per read on the recorded values; DRAM-latency bound"),
        "k_emit": K_(sub(ms_ph[2], u3, 2), n_ivl * 17 + r, traffic.get("k_emit"),
                     "class strings, streaming: 17 bytes per interval in, r bytes out (skipped when results are interval tables)"),
        "retry_launch": {"ms": ms_ph[3]},
    }
    phases = {"wall": ms_ph[0], "rel": ms_ph[1], "unrel_emit": ms_ph[2]}
    dom = max((k for k in kern if k != "retry_launch"), key=lambda k: kern[k]["ms"])
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": kern[dom]["frac"], "peak_source": peak_src,
            "traffic": kern[dom]["traffic"],
            "algorithmic_bytes_per_launch": kern[dom]["bytes"],
            "note": "the dominant kernel is not a streaming kernel: its bytes are the interval tables it works on; what bounds it "
                    "is in its note.  `pipeline` is the whole step against the bytes each read touches (SURVEY 8d, unfused: "
                    "c + 4n + n/4 + r/4 + r), `k_decode` the streaming kernel the north star's 50 % is about",
            "pipeline": {"ms": ms_dec + ms_cls, "bytes": int(bytes_dec + bytes_cls),
                         "GBps": (bytes_dec + bytes_cls) / ((ms_dec + ms_cls) * 1e-3) / 1e9,
                         "frac": (bytes_dec + bytes_cls) / ((ms_dec + ms_cls) * 1e-3) / 1e9 / peak},
            "units": {"kmers": n, "bases": r, "compressed_profile_bytes": c, "wall_candidates": n_cand, "intervals": n_ivl,
                      "reliable_intervals": n_rel, "intervals_visited_by_unreliable_sweeps": n_vis},
            "classification_ms": ms_cls, "classification_bytes": bytes_cls, "phases_ms": phases,
            "kernels": kern}

    ctx.close()
    shard_info = None
    if world > 1:
        g = [None] * world
        dist.all_gather_object(g, {"rank": rank, "reads": [int(beg), int(end)], "kmers": data.kmers, "ms": round(my_ms, 3),
                                   "chromosomes_generated": n_generated})
        shard_info = g
    rc_exit = 0
    if rank == 0:
        cpu = cli = None
        if world == 1 and not args.no_cli:
            try:
                cli = cli_file_to_file(wl, args, per_gpu_chunks, cores) if wl["sim"].get("exact") != 2 else None
            except Exception as e:
                cli = {"value": None, "unavailable": str(e)[:300]}
        if not args.no_cpu_baseline and world == 1 and wl["sim"].get("exact") != 2:
            try:
                threads = max(1, min(cores, args.cpu_threads if args.cpu_threads > 0 else 16))
                with tempfile.TemporaryDirectory(prefix="cpbench_") as tmp:
                    nck = max(1, int(round(args.cpu_sample_mbases / wl["cov"] / args.chunk_mb)))
                    fasta, per = dataset_files(tmp, wl, nck, args.chunk_mb, min(cores, 16))
                    km = sum(p[1] for p in per)
                    v, kind, used, detail = run_reference_once(fasta, km, threads)
                    cpu = {"value": v, "unit": "k-mers/s", "cores": used, "kind": kind,
                           "sample": "the first %d chromosomes of the read set (%d reads, %d k-mers), ClassPro -T%d, one whole run of "
                                     "the program incl. start-up and per-thread setup (the --impl reference arm runs the whole set)"
                                     % (nck, sum(p[0] for p in per), km, used),
                           "detail": detail, "host_cores": cores}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        cfg = workload_config(wl, args, world, nchunks_total, reads_total, int(total_kmers))
        line = {"metric": "classified k-mers/sec", "value": value, "unit": "k-mers/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "run": {"reads_rank0": data.n_reads, "kmers_rank0": data.kmers, "bases_rank0": data.bases,
                        "compressed_profile_bytes_per_kmer": c / max(1, n),
                        "l2": "inputs larger than L2 (%.1f GB of counts+classes per step vs 126 MB)" % ((2 * n + r) / 1e9),
                        "e2e_batches": len(data.batches), "gen_seconds": round(t_gen, 1), "producer": producer,
                        "ms_per_step_min_over_ranks": step_ms_min,
                        "rank_time_imbalance": (step_ms / step_ms_min) if step_ms_min > 0 else None, "shards": shard_info},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": h2d_all,
                        "d2h_bytes_per_step": d2h_ivl_all,
                        "ms_per_step": e2e_s * 1e3,
                        "matches_resident_result": same,
                        "result": "interval tables, 4 B per interval (cpg_collect_intervals); expanded to the class strings on "
                                  "the host (cpg_expand_intervals) outside the timed region and compared with the resident result",
                        "class_strings": {"value": total_kmers / e2e_cls_s, "ms_per_step": e2e_cls_s * 1e3,
                                          "d2h_bytes_per_step": d2h_cls_all,
                                          "what": "the same pass returning 1 byte per base (cpg_collect)"},
                        "starts_from": "parsed, 2-bit packed reads and fetched profile bytes in pinned host memory "
                                       "(file parsing and output formatting are in `cli`)"},
                "cli": cli,
                "parity_sample": par_tot,
                "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cpu,
                "reads_with_errors": n_bad}
        print(json.dumps(line))
        if par_tot["flip_fraction"] > FLIP_BUDGET:
            sys.stderr.write("bench.py: parity sample over budget: %d of %d k-mers differ from the oracle\n"
                             % (par_tot["flips"], par_tot["kmers"]))
            rc_exit = 3
        if not same:
            sys.stderr.write("bench.py: the end-to-end result differs from the resident result\n")
            rc_exit = 3
    data.free()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(rc_exit)


if __name__ == "__main__":
    main()
