#!/usr/bin/env python
"""bench.py -- classified k-mers/second of the ClassPro classification path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank/GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step is one pass of the hot path (profile decode -> walls -> reliable DP -> unreliable intervals
-> class strings) over the rank's whole synthetic dataset.  At N=1 the workload is BASELINE.json
configs[1]: 100 Mb synthetic diploid genome, 1 % heterozygosity, 30x HiFi-like reads (~20 kb),
k=40.  Reads are independent, so ranks never exchange data (weak scaling: every rank classifies
its own dataset of that size; the only collective is the max-reduction of the timings).

  value  k-mers/s with the batch already resident in HBM (kernels only, CUDA events)
  e2e    k-mers/s through cpg_submit/cpg_collect with pinned HOST buffers: H2D copies of the packed
         reads + compressed profiles and the D2H copy of the class strings are inside the timed
         region (double buffered over two streams)
"""
import argparse
import json
import os
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
WORKLOAD = ("100 Mb synthetic diploid genome, 1% heterozygosity, 30x HiFi-like reads (~20 kb), k=40 "
            "(BASELINE.json configs[1])")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------- workload
def gen_chunk(seed, genome_len, cov, out, idx):
    import cpkit
    out[idx] = cpkit.simulate(seed=seed, genome_len=genome_len, cov=cov, het=0.01, snp_only=1, exact=0,
                              len_mean=20000, len_sd=2000, len_min=5000, len_max=50000, nparts=1)


def make_workload(rank, genome_mb, cov, chunk_mb, threads):
    """Ground-truth-coverage profiles (tools/cpsim.c, mode 'fast') for genome_mb megabases generated
    as independent chunk_mb chromosomes in parallel threads (the C generator releases the GIL)."""
    nchunks = max(1, int(round(genome_mb / chunk_mb)))
    sims = [None] * nchunks
    seeds = [1000 * (rank + 1) + c for c in range(nchunks)]
    pending = list(range(nchunks))
    while pending:
        now, pending = pending[:threads], pending[threads:]
        ths = [threading.Thread(target=gen_chunk, args=(seeds[c], int(chunk_mb * 1e6), cov, sims, c)) for c in now]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    return sims


class HostData:
    """The rank's dataset in pinned host memory, plus per-batch views."""

    def __init__(self, sims, n_batches):
        from classpro_b200.abi import PinnedArray, pack_codes
        import classpro_b200 as cp
        rlen = np.concatenate([s.rlen for s in sims]).astype(np.int32)
        n = len(rlen)
        self.n_reads = n
        self.kmers = int((rlen.astype(np.int64) - K + 1).sum())
        self.bases = int(rlen.astype(np.int64).sum())
        self.hist = np.sum([s.hist for s in sims], axis=0)
        # packed sequence
        poff = np.zeros(n + 1, np.int64)
        np.cumsum((rlen.astype(np.int64) + 3) // 4, out=poff[1:])
        self.pin_seq = PinnedArray(int(poff[-1]) + 64)
        seq = self.pin_seq.u8
        at = 0
        r0 = 0
        for s in sims:
            pk, po = pack_codes(s.seq, s.seq_off, s.rlen)
            seq[at:at + int(po[-1])] = pk[:int(po[-1])]
            at += int(po[-1])
            r0 += s.nreads
        prof_len = np.concatenate([np.diff(s.prof_off) for s in sims])
        pro = np.zeros(n + 1, np.int64)
        np.cumsum(prof_len, out=pro[1:])
        self.pin_prof = PinnedArray(int(pro[-1]) + 64)
        prof = self.pin_prof.u8
        at = 0
        for s in sims:
            prof[at:at + len(s.prof)] = s.prof
            at += len(s.prof)
        self.prof_bytes = int(pro[-1])
        self.seq_bytes = int(poff[-1])
        self.rlen, self.seq_off, self.prof_off = rlen, poff, pro
        self.pin_cls = PinnedArray(self.bases + 64)
        self.whole = cp.Batch(seq, poff, rlen, prof, pro, 2)
        # contiguous batches balanced by bases
        cum = np.cumsum(rlen.astype(np.int64))
        cuts = [0]
        for b in range(1, n_batches):
            cuts.append(int(np.searchsorted(cum, self.bases * b // n_batches)))
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

    def free(self):
        self.whole = None
        self.batches = []
        for p in (self.pin_seq, self.pin_prof, self.pin_cls):
            p.free()


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
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_files(tmp, seed, mbases):
    """A bounded sample of the same workload written as FASTA + FastK files for the reference."""
    import cpkit
    glen = max(200000, int(mbases * 1e6 / 30))
    sim = cpkit.simulate(write_to=tmp, root="sample", seed=seed, genome_len=glen, cov=30., het=0.01, snp_only=1,
                         exact=0, len_mean=20000, len_sd=2000, len_min=5000, len_max=50000, nparts=1)
    return os.path.join(tmp, "sample.fasta"), sim


def run_cpu_once(fasta, sim, threads):
    """One timed run of the reference's own CPU implementation on the sample.  Returns
    (k-mers/s, kind, cores, detail)."""
    import cpkit
    kmers = int((np.maximum(sim.rlen.astype(np.int64) - K + 1, 0)).sum())
    if cpkit.have_reference():
        t0 = time.time()
        p = subprocess.run([cpkit.REF_BIN, "-v", "-T%d" % threads, fasta], cwd=os.path.dirname(fasta),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("reference ClassPro failed: " + p.stderr[-500:])
        phase = None
        for line in p.stderr.splitlines():
            if line.startswith("Resources for phase:"):
                phase = line
        return kmers / wall, "reference", threads, {"wall_s": round(wall, 3), "phase_line": phase}
    # reference not built here: the oracle port, single thread
    out = fasta + ".oracle.class"
    t0 = time.time()
    rc = cpkit.oracle_lib().cpo_run_file(fasta.encode(), fasta[:-len(".fasta")].encode(), 0, 20000, out.encode(), 0, None)
    wall = time.time() - t0
    if rc != 0:
        raise RuntimeError("oracle port failed rc=%d" % rc)
    return kmers / wall, "port", 1, {"wall_s": round(wall, 3)}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    threads = max(1, min(cores, args.cpu_threads if args.cpu_threads > 0 else 16))
    with tempfile.TemporaryDirectory(prefix="cpbench_") as tmp:
        fasta, sim = cpu_sample_files(tmp, 777, args.cpu_sample_mbases)
        kmers = int((np.maximum(sim.rlen.astype(np.int64) - K + 1, 0)).sum())
        for _ in range(args.warmup if args.warmup < 2 else 1):
            run_cpu_once(fasta, sim, threads)
        t0 = time.time()
        kind = "reference"
        detail = None
        for _ in range(args.steps):
            v, kind, used, detail = run_cpu_once(fasta, sim, threads)
        dt = time.time() - t0
    value = kmers * args.steps / dt
    line = {"impl": "reference", "metric": "classified k-mers/sec", "value": value, "unit": "k-mers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "%d reads / %d k-mers of the same generator per step" % (sim.nreads, kmers)},
            "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": used, "kind": kind,
                             "sample": "%d reads, %d k-mers, ClassPro -T%d incl. its per-thread setup" % (sim.nreads, kmers, used),
                             "detail": detail},
            "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": cores}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-mb", type=float, default=100.)
    ap.add_argument("--cov", type=float, default=30.)
    ap.add_argument("--chunk-mb", type=float, default=5.)
    ap.add_argument("--gen-threads", type=int, default=0)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--cpu-sample-mbases", type=float, default=360.)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import classpro_b200 as cp
    cores = host_cores()
    gen_threads = args.gen_threads or max(1, min(16, cores // max(1, world)))
    t_gen = time.time()
    sims = make_workload(rank, args.genome_mb, args.cov, args.chunk_mb, gen_threads)
    data = HostData(sims, args.batches)
    del sims
    t_gen = time.time() - t_gen
    model = cp.Model.from_hist(K, data.hist[1:32768], data.hist[32768], data.hist[32769], read_len=20000)
    ctx = cp.Context(model, local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident: whole dataset in HBM, kernels only
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
    cls_res, status = ctx.download(data.whole)
    n_bad = int((status & cp.ST_FATAL != 0).sum())
    step_ms = max_over_ranks(ms_dec + ms_cls)
    total_kmers = sum_over_ranks(float(data.kmers))
    value = total_kmers / (step_ms * 1e-3)

    # ---- end to end: pinned host buffers in, class strings out, double buffered
    def e2e_pass():
        inflight = []
        for i, (bt, out) in enumerate(data.batches):
            slot = i & 1
            if len(inflight) == 2:
                s, b, o = inflight.pop(0)
                ctx.collect(s, b, o)
            ctx.submit(slot, bt)
            inflight.append((slot, bt, out))
        for s, b, o in inflight:
            ctx.collect(s, b, o)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_pass()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        e2e_pass()
    barrier()
    e2e_s = max_over_ranks((time.time() - t0) / args.steps)
    e2e_value = total_kmers / e2e_s
    same = bool(np.array_equal(data.pin_cls.u8[:data.bases], cls_res[:data.bases]))

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
    # bytes in, counts (2 B/k-mer) and the candidate bit map (1 bit/k-mer) out.  k_classify reads the
    # counts, the bit map and the 2-bit bases and writes one class byte per base.
    bytes_dec = c + 2 * n + n // 8
    bytes_cls = 2 * n + n // 8 + (r + 3) // 4 + r
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    # classification = k_wall + k_rel + k_unrel (+ the retry launch): device time of each from the
    # CUDA events between them (cpg_phase_cycles returns nanoseconds of the last timed run)
    tot_ns = max(1, sum(phase))
    ms_ph = [ms_cls * x / tot_ns for x in phase]
    bytes_wall = 2 * n + n // 8 + (r + 3) // 4          # counts + candidate bits + 2-bit bases
    kern = {
        "k_decode": {"ms": ms_dec, "bytes": bytes_dec, "GBps": bytes_dec / (ms_dec * 1e-3) / 1e9,
                     "frac": bytes_dec / (ms_dec * 1e-3) / 1e9 / peak, "traffic": (traffic or {}).get("k_decode"),
                     "note": "profile decode + wall-candidate scan fused: c + 2n + n/8 bytes; issue bound"},
        "k_wall": {"ms": ms_ph[0], "bytes": bytes_wall, "GBps": bytes_wall / (max(ms_ph[0], 1e-6) * 1e-3) / 1e9,
                   "frac": bytes_wall / (max(ms_ph[0], 1e-6) * 1e-3) / 1e9 / peak, "traffic": (traffic or {}).get("k_wall"),
                   "note": "wall detection + reliable intervals: 2n + n/8 + r/4 bytes in, interval tables out; "
                           "DRAM-latency / FP64-latency bound (see stall mix in profiles/)"},
        "k_rel": {"ms": ms_ph[1], "bytes": None, "traffic": (traffic or {}).get("k_rel"),
                  "note": "reliable-interval DP on the interval tables (48 B per interval): FP64 dependency chains"},
        "k_unrel": {"ms": ms_ph[2], "bytes": r, "traffic": (traffic or {}).get("k_unrel"),
                    "note": "unreliable intervals + class string (r bytes out)"},
        "retry_launch": {"ms": ms_ph[3]},
    }
    dom = max(("k_decode", "k_wall", "k_rel", "k_unrel"), key=lambda k: kern[k]["ms"])
    dom_bytes = kern[dom]["bytes"] if kern[dom]["bytes"] is not None else bytes_cls
    dom_ms = kern[dom]["ms"]
    roof = {"bound": "hbm", "kernel": dom, "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
            "traffic": (traffic or {}).get(dom),
            "algorithmic_bytes_per_launch": dom_bytes,
            "classification_ms": ms_cls, "classification_bytes": bytes_cls,
            "kernels": kern}

    ctx.close()
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                threads = max(1, min(cores, args.cpu_threads if args.cpu_threads > 0 else 16))
                with tempfile.TemporaryDirectory(prefix="cpbench_") as tmp:
                    fasta, sim = cpu_sample_files(tmp, 777, args.cpu_sample_mbases)
                    v, kind, used, detail = run_cpu_once(fasta, sim, threads)
                    cpu = {"value": v, "unit": "k-mers/s", "cores": used, "kind": kind,
                           "sample": "%d reads / %d k-mers of the same generator, ClassPro -T%d, whole run incl. setup"
                                     % (sim.nreads, int((sim.rlen.astype(np.int64) - K + 1).sum()), used),
                           "detail": detail, "host_cores": cores}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "k-mers/s", "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}
        line = {"metric": "classified k-mers/sec", "value": value, "unit": "k-mers/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "reads_per_gpu": data.n_reads, "kmers_per_gpu": data.kmers,
                           "bases_per_gpu": data.bases, "compressed_profile_bytes_per_kmer": c / n,
                           "l2": "inputs larger than L2 (%.1f GB of counts+classes per step vs 126 MB)" % ((2 * n + r) / 1e9),
                           "e2e_batches": len(data.batches), "profiles": "ground-truth coverage (tools/cpsim.c --fast)",
                           "gen_seconds": round(t_gen, 1)},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "k-mers/s", "h2d_bytes_per_step": data.h2d_bytes,
                        "d2h_bytes_per_step": data.d2h_bytes, "ms_per_step": e2e_s * 1e3,
                        "matches_resident_result": same},
                "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cpu,
                "reads_with_errors": n_bad}
        print(json.dumps(line))
    data.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
