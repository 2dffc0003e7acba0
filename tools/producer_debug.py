"""Stage-by-stage comparison of the profile producer's key-range passes with its single pass, on a GPU box.
Runs cpg_count_kmers twice on the same reads with CPG_COUNT_DEBUG_DIR set (every stage of every pass is
written to disk by the library): once with one pass, once with CPG_COUNT_PASSES=<p>, then checks, pass by
pass: the appended (key, index) multiset against the single-pass keys filtered by cpg_key_pass, the sort,
the run ids, the run starts and the scattered counts.  usage: python tools/producer_debug.py [passes] [genome_len]"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit as kit                                     # noqa: E402
from classpro_b200 import abi                           # noqa: E402

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 3
GLEN = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
K = 40
M64 = np.uint64(0xffffffffffffffff)


def key_pass(hi, lo, npass):
    with np.errstate(over="ignore"):
        x = lo ^ (hi * np.uint64(0x9e3779b97f4a7c15))
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
        return (((x >> np.uint64(32)) * np.uint64(npass)) >> np.uint64(32)).astype(np.int64)


def run(sim, pseq, seq_off, env, d):
    L = abi.lib()
    L.cpg_count_error.restype = C.c_char_p
    for k in ("CPG_COUNT_PASSES", "CPG_COUNT_DEBUG_DIR"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["CPG_COUNT_DEBUG_DIR"] = d
    n = sim.nreads
    cnt_off = np.zeros(n + 1, np.int64)
    counts = np.zeros(sim.total_kmers + 1, np.uint16)
    hist = np.zeros(32770, np.int64)
    rlen = np.ascontiguousarray(sim.rlen, np.int32)
    rc = L.cpg_count_kmers(0, K, n, C.c_void_p(pseq.ctypes.data), C.c_void_p(seq_off.ctypes.data), C.c_void_p(rlen.ctypes.data),
                           C.c_void_p(cnt_off.ctypes.data), C.c_void_p(counts.ctypes.data), C.c_void_p(hist.ctypes.data))
    print("rc", rc, L.cpg_count_error())
    return counts[:sim.total_kmers], hist


def ld(d, name, p, dt):
    return np.fromfile(os.path.join(d, "%s.%d.bin" % (name, p)), dtype=dt)


sim = kit.simulate(kmer=K, seed=3, genome_len=GLEN, cov=12., het=0.01, len_mean=6000, short_reads=1, repeat_frac=0.3)
pseq, seq_off = abi.pack_codes(sim.seq, sim.seq_off, sim.rlen)
d1 = tempfile.mkdtemp(prefix="pd1_")
dp = tempfile.mkdtemp(prefix="pdp_")
c1, h1 = run(sim, pseq, seq_off, {}, d1)
print("single pass counts == harness:", np.array_equal(c1, sim.counts))
cp_, hp = run(sim, pseq, seq_off, {"CPG_COUNT_PASSES": str(NP)}, dp)
print("%d passes counts == harness:" % NP, np.array_equal(cp_, sim.counts), " hist:", np.array_equal(hp, sim.hist))
klo, khx = ld(d1, "keys_lo", 0, np.uint64), ld(d1, "keys_hx", 0, np.uint64)
n = len(klo)
hi = khx >> np.uint64(48)
assert np.array_equal(khx & np.uint64((1 << 48) - 1), np.arange(n, dtype=np.uint64)), "single-pass index column"
ps = key_pass(hi, klo, NP)
prev_counts = None
for p in range(NP):
    want = np.flatnonzero(ps == p)
    try:
        alo, ahx = ld(dp, "keys_lo", p, np.uint64), ld(dp, "keys_hx", p, np.uint64)
    except FileNotFoundError:
        print("pass %d: no dump (empty pass?) expected %d keys" % (p, len(want)))
        continue
    idx = (ahx & np.uint64((1 << 48) - 1)).astype(np.int64)
    print("pass %d: appended %d keys, expected %d" % (p, len(alo), len(want)))
    o = np.argsort(idx, kind="stable")
    ok_idx = np.array_equal(idx[o], want)
    ok_keys = ok_idx and np.array_equal(alo[o], klo[want]) and np.array_equal(ahx[o] >> np.uint64(48), hi[want])
    print("   index set ok: %s, keys ok: %s" % (ok_idx, ok_keys))
    if not ok_idx:
        print("   dup indices: %d, missing: %d, extra: %d" % (len(idx) - len(np.unique(idx)), len(np.setdiff1d(want, idx)), len(np.setdiff1d(idx, want))))
    slo, shx = ld(dp, "sort_lo", p, np.uint64), ld(dp, "sort_hx", p, np.uint64)
    shi = shx >> np.uint64(48)
    sorted_ok = bool(np.all((shi[1:] > shi[:-1]) | ((shi[1:] == shi[:-1]) & (slo[1:] >= slo[:-1]))))
    same_set = np.array_equal(np.sort(shx & np.uint64((1 << 48) - 1)), np.sort(ahx & np.uint64((1 << 48) - 1)))
    sidx = (shx & np.uint64((1 << 48) - 1)).astype(np.int64)
    pair_ok = bool(np.array_equal(slo, klo[np.minimum(sidx, n - 1)]) and np.array_equal(shi, hi[np.minimum(sidx, n - 1)]))
    print("   sorted: %s, same index multiset after sort: %s, (key,index) pairs intact: %s" % (sorted_ok, same_set, pair_ok))
    rid, start = ld(dp, "rid", p, np.uint32), ld(dp, "start", p, np.uint32)
    head = np.ones(len(slo), bool)
    head[1:] = (slo[1:] != slo[:-1]) | (shi[1:] != shi[:-1])
    rid_want = np.cumsum(head).astype(np.uint32)
    st_want = np.concatenate([np.flatnonzero(head), [len(slo)]]).astype(np.uint32)
    print("   rid ok: %s, start ok: %s (runs %d)" % (np.array_equal(rid, rid_want), np.array_equal(start[:len(st_want)], st_want), len(st_want) - 1))
    cnt = ld(dp, "counts", p, np.uint16)
    runlen = np.diff(st_want.astype(np.int64))
    cw = np.minimum(runlen[rid_want.astype(np.int64) - 1], 32767).astype(np.uint16)
    got = cnt[sidx]
    print("   counts of this pass's positions ok: %s (%d wrong)" % (np.array_equal(got, cw), int((got != cw).sum())))
    if prev_counts is not None:
        touched = np.flatnonzero(prev_counts != cnt)
        outside = np.setdiff1d(touched, sidx)
        print("   positions changed by this pass that are not its own: %d" % len(outside))
    prev_counts = cnt
