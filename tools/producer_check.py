"""Quick GPU check (CPG_COUNT_TIMING=1: device times of the stages on stderr)
Quick GPU check of the profile producer (cpg_count_kmers / cpg_encode_profiles) against the harness
counter of tools/cpsim.c; numpy + ctypes only.  usage: python tools/producer_check.py [genome_len] [cov]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit as kit                                     # noqa: E402
from classpro_b200 import abi                           # noqa: E402

glen = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 20.
L = abi.lib()
L.cpg_count_error.restype = C.c_char_p
KS = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [40, 32, 21]
for K in KS:
    sim = kit.simulate(kmer=K, seed=5, genome_len=glen, cov=cov, het=0.01, len_mean=10000, short_reads=1, repeat_frac=0.2)
    pseq, seq_off = abi.pack_codes(sim.seq, sim.seq_off, sim.rlen)
    n = sim.nreads
    cnt_off = np.zeros(n + 1, np.int64); counts = np.zeros(sim.total_kmers + 1, np.uint16); hist = np.zeros(32770, np.int64)
    rlen = np.ascontiguousarray(sim.rlen, np.int32)
    if os.environ.get("CPG_COUNT_TIMING"):                 # once untimed: CUDA context, first-use costs
        L.cpg_count_kmers(0, K, n, C.c_void_p(pseq.ctypes.data), C.c_void_p(seq_off.ctypes.data), C.c_void_p(rlen.ctypes.data),
                          C.c_void_p(cnt_off.ctypes.data), C.c_void_p(counts.ctypes.data), C.c_void_p(hist.ctypes.data))
    t0 = time.time()
    rc = L.cpg_count_kmers(0, K, n, C.c_void_p(pseq.ctypes.data), C.c_void_p(seq_off.ctypes.data), C.c_void_p(rlen.ctypes.data),
                           C.c_void_p(cnt_off.ctypes.data), C.c_void_p(counts.ctypes.data), C.c_void_p(hist.ctypes.data))
    t1 = time.time()
    assert rc == 0, L.cpg_count_error()
    prof = np.zeros(2 * sim.total_kmers + 16, np.uint8); prof_off = np.zeros(n + 1, np.int64)
    rc = L.cpg_encode_profiles(0, n, C.c_void_p(counts.ctypes.data), C.c_void_p(cnt_off.ctypes.data), C.c_void_p(prof.ctypes.data),
                               C.c_int64(len(prof)), C.c_void_p(prof_off.ctypes.data))
    t2 = time.time()
    assert rc == 0, L.cpg_count_error()
    ok = (np.array_equal(counts[:sim.total_kmers], sim.counts), np.array_equal(hist, sim.hist),
          np.array_equal(prof_off, sim.prof_off), np.array_equal(prof[:prof_off[n]], sim.prof))
    print("K=%d: %d k-mers, counts/hist/offsets/bytes equal to the harness: %s; count %.3f s, encode %.3f s (host wall, incl. copies and allocation)"
          % (K, sim.total_kmers, ok, t1 - t0, t2 - t1), flush=True)
    if not all(ok):                                        # what a debugging session wants to see first
        got, want = counts[:sim.total_kmers], sim.counts
        bad = np.flatnonzero(got != want)
        print("  %d of %d counts differ; first: %s" % (len(bad), len(want), [(int(i), int(got[i]), int(want[i])) for i in bad[:8]]))
        print("  sum over distinct k-mers of count: got %d, harness %d, k-mers %d; distinct: got %d, harness %d"
              % (int((hist[1:32767] * np.arange(1, 32767)).sum()) + int(hist[32769]), int((sim.hist[1:32767] * np.arange(1, 32767)).sum()) + int(sim.hist[32769]),
                 sim.total_kmers, int(hist[1:32768].sum()), int(sim.hist[1:32768].sum())))
        print("  got < want at %d positions, got > want at %d, got == 0 at %d" % (int((got < want).sum()), int((got > want).sum()), int((got == 0).sum())))
    assert all(ok)
print("producer_check: OK")
