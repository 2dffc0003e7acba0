"""C-ABI surface (loads, exports everything include/classpro_gpu.h declares, fails loudly without
a GPU), host model parity, packing, and the multi-rank sharding logic under gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "classpro_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import classpro_b200 as cp
    L = cp.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libclasspro_b200.so does not export %s" % n


def test_library_does_not_link_the_oracle():
    """The product must not route through oracle/: no cpo_* symbol, no liboracle dependency."""
    import classpro_b200 as cp
    out = subprocess.run(["nm", "-D", cp.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "cpo_" not in out
    ldd = subprocess.run(["ldd", cp.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in ldd and "hostsim" not in ldd


def test_no_gpu_means_loud_failure():
    import classpro_b200 as cp
    if cp.lib().cpg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = cp.Model.from_cov(40, 0, 30)
    with pytest.raises(cp.CpgError) as e:
        cp.Context(m)
    assert "no CPU fallback" in str(e.value)


def test_host_model_matches_oracle(kit):
    import classpro_b200 as cp
    sim = kit.simulate(seed=51, genome_len=40000, cov=30., het=0.01)
    for cov_opt, rl in ((0, 20000), (41, 15000)):
        om = kit.oracle_model(sim, cov_opt, rl)
        gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=rl)
        assert gm.cov == list(om.cov)
        assert gm.c.dr_ratio == om.dr_ratio and gm.c.cmax == om.cmax
        assert np.array_equal(np.frombuffer(gm.c.logfact, dtype=np.float64), np.frombuffer(om.logfact, dtype=np.float64))
    with pytest.raises(cp.CpgError):      # D-coverage too high for the 8-bit threshold table (wall.c:174-177)
        cp.Model.from_cov(40, 0, 250)


def test_model_errors(kit):
    import classpro_b200 as cp
    flat = np.ones(32767, dtype=np.int64)      # no peak >= 10 (hist.c:65-68)
    with pytest.raises(cp.CpgError):
        cp.Model.from_hist(40, flat, 1, 1)


def test_pack_seq():
    import classpro_b200 as cp
    L = cp.lib()
    s = b"ACGTTGCAAC"
    out = np.zeros(4, dtype=np.uint8)
    assert L.cpg_pack_seq(s, len(s), out.ctypes.data) == 0
    codes = [(out[i >> 2] >> ((i & 3) * 2)) & 3 for i in range(len(s))]
    assert codes == ["ACGT".index(chr(c)) for c in s]
    assert L.cpg_pack_seq(b"ACGNT", 5, out.ctypes.data) == 1
    assert L.cpg_pack_seq(b"acgt", 4, out.ctypes.data) == 1      # the reference compares raw bytes
    from classpro_b200.abi import pack_codes
    rng = np.random.default_rng(3)
    rl = np.array([5, 8, 13, 40], dtype=np.int32)
    so = np.concatenate([[0], np.cumsum(rl)]).astype(np.int64)
    codes = rng.integers(0, 4, size=int(so[-1])).astype(np.uint8)
    pk, po = pack_codes(codes, so, rl)
    for r in range(len(rl)):
        asc = bytes(b"ACGT"[c] for c in codes[so[r]:so[r + 1]])
        ref = np.zeros((rl[r] + 3) // 4, dtype=np.uint8)
        L.cpg_pack_seq(asc, int(rl[r]), ref.ctypes.data)
        assert np.array_equal(ref, pk[po[r]:po[r + 1]])


def test_expand_intervals():
    """cpg_expand_intervals (host code of the library): 'N' x (K-1), then every interval's class over its stretch
    -- the class string of CPG_RESULT_CLASSES mode (src/ClassPro.c:114-117,265-271).  Also: a table that stops short
    of the profile leaves the rest untouched, entries past the profile are clipped, reads shorter than K are all N."""
    import classpro_b200 as cp
    L = cp.lib()
    rng = np.random.default_rng(9)
    K = 40
    for _ in range(200):
        rlen = int(rng.integers(K, 3000))
        plen = rlen - K + 1
        ncut = int(rng.integers(0, min(plen, 60)))
        ends = np.unique(np.concatenate([rng.integers(1, plen + 1, size=ncut), [plen]])).astype(np.int64)
        codes = rng.integers(0, 4, size=len(ends))
        ivl = ((ends << 3) | codes).astype(np.uint32)
        out = np.full(rlen + 8, ord("#"), dtype=np.uint8)
        L.cpg_expand_intervals(K, rlen, ivl.ctypes.data, len(ivl), out.ctypes.data)
        want = bytearray(b"N" * (K - 1))
        b = 0
        for e, c in zip(ends, codes):
            want += bytes([b"ERHD"[c]]) * int(e - b)
            b = int(e)
        assert out[:rlen].tobytes() == bytes(want) and (out[rlen:] == ord("#")).all()
    out = np.full(64, ord("#"), dtype=np.uint8)
    L.cpg_expand_intervals(K, 30, None, 0, out.ctypes.data)                  # shorter than K: rlen 'N's
    assert out[:30].tobytes() == b"N" * 30 and out[30] == ord("#")
    ivl = np.array([(5 << 3) | 2, (500 << 3) | 3], dtype=np.uint32)          # second entry past the profile: clipped
    L.cpg_expand_intervals(K, 50, ivl.ctypes.data, 2, out.ctypes.data)
    assert out[:50].tobytes() == b"N" * 39 + b"HHHHH" + b"DDDDDD" and out[50] == ord("#")


def test_shard_ranges_properties():
    from classpro_b200.shard import shard_ranges, reference_thread_ranges
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 1000):
        w = rng.integers(1, 5000, size=n)
        for ranks in (1, 2, 3, 8):
            rs = shard_ranges(w, ranks)
            assert len(rs) == ranks and rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            if n >= 100:
                tot = [int(w[a:b].sum()) for a, b in rs]
                assert max(tot) - min(tot) <= 2 * int(w.max())
    assert reference_thread_ranges(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]


def test_plan_chunk_shards_properties():
    """The bench's shard plan over a read set that exists as chunks: ranges cover everything once, the chunk
    pieces of a rank add up to its range, and a rank only touches chunks that overlap its range."""
    from classpro_b200.shard import plan_chunk_shards, shard_ranges
    rng = np.random.default_rng(9)
    for nchunks, ranks in ((1, 1), (3, 2), (8, 8), (5, 8), (20, 3)):
        cw = [(100 + c, rng.integers(50, 4000, size=int(rng.integers(1, 300)))) for c in range(nchunks)]
        plans, total = plan_chunk_shards(cw, ranks)
        assert total == sum(len(w) for c, w in cw) and len(plans) == ranks
        allw = np.concatenate([w for c, w in cw])
        assert [(b, e) for b, e, _ in plans] == shard_ranges(allw, ranks)
        seen = []
        first = dict(zip([c for c, w in cw], np.cumsum([0] + [len(w) for c, w in cw[:-1]])))
        for beg, end, need in plans:
            assert sum(b - a for c, a, b in need) == end - beg
            for c, a, b in need:
                seen.extend(range(first[c] + a, first[c] + b))
        assert seen == list(range(total))


BENCH_SHARD_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import bench
from classpro_b200.shard import plan_chunk_shards
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
wl = dict(bench.WORKLOADS["c2"]); wl["cov"] = 8.
C = 3                                            # chromosomes of the global set: rank 0 is home to 1, rank 1 to 2
home = list(range(C * rank // world, C * (rank + 1) // world))
sims = bench.gen_chunks(home, wl, 0.05, 2)
mine = [(c, np.diff(sims[c].prof_off)) for c in home]
g = [None] * world
dist.all_gather_object(g, mine)
allw = sorted([cw for x in g for cw in x], key=lambda cw: cw[0])
plans, total = plan_chunk_shards(allw, world)
beg, end, need = plans[rank]
extra = [c for c, a, b in need if c not in sims]
sims.update(bench.gen_chunks(extra, wl, 0.05, 2))
# what this rank holds: per global read id, a checksum of its sequence and profile bytes
out = {}
first = {}
at = 0
for c, w in allw:
    first[c] = at; at += len(w)
for c, a, b in need:
    s = sims[c]
    for i in range(a, b):
        out[first[c] + i] = (int(s.rlen[i]), int(s.read_prof(i).astype(np.int64).sum()), int(s.seq[s.seq_off[i]:s.seq_off[i + 1]].astype(np.int64).sum()))
assert sorted(out) == list(range(beg, end))
g2 = [None] * world
dist.all_gather_object(g2, (out, extra))
if rank == 0:
    whole = bench.gen_chunks(range(C), wl, 0.05, 2)       # the same set generated in one place
    want = {}
    at = 0
    for c in range(C):
        s = whole[c]
        for i in range(s.nreads):
            want[at + i] = (int(s.rlen[i]), int(s.read_prof(i).astype(np.int64).sum()), int(s.seq[s.seq_off[i]:s.seq_off[i + 1]].astype(np.int64).sum()))
        at += s.nreads
    got = {}
    for o, e in g2:
        assert not (set(o) & set(got))
        got.update(o)
    assert got == want and at == total
    print("BENCH_SHARD_OK", world, total, [e for o, e in g2])
dist.destroy_process_group()
"""


def test_bench_shard_plan_two_ranks_gloo(kit, tmp_path):
    """bench.py's N>1 flow on CPU: every rank generates its home chromosomes, the per-read profile sizes are
    all-gathered, the ranges are cut, a rank whose range reaches into a neighbour's chromosome generates that one
    too -- together the ranks hold exactly the reads of the set generated in one place, each once."""
    script = tmp_path / "worker.py"
    script.write_text(BENCH_SHARD_WORKER % {"root": ROOT})
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "BENCH_SHARD_OK 2" in p.stdout


WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import cpkit
from classpro_b200.shard import shard_ranges
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sim = cpkit.simulate(seed=61, genome_len=30000, cov=20., het=0.01)
w = np.diff(sim.prof_off)
beg, end = shard_ranges(w, world)[rank]
# each rank classifies ONLY its own contiguous range (host-sim of the device logic stands in for
# the GPU here); no data-path collective is needed
gm = cpkit.gpu_model_from_sim(cpkit.hostsim_lib(), sim)
mine = [cpkit.hostsim_classify(gm, sim.read_ascii(i).tobytes(), sim.read_counts(i), 2)[1] for i in range(beg, end)]
kmers = sum(len(x) - 39 for x in mine)
# the only cross-rank steps: ordered gather of the outputs on rank 0, max of the timings
gathered = [None] * world
dist.all_gather_object(gathered, (beg, end, mine))
t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
tot = torch.tensor([float(kmers)]); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
if rank == 0:
    om = cpkit.oracle_model(sim); ow = cpkit.OracleWork(clean=True)
    full = [ow.classify(om, sim.read_ascii(i).tobytes(), sim.read_counts(i)) for i in range(sim.nreads)]
    got = []
    for b, e, part in sorted(gathered):
        got.extend(part)
    assert got == full, "sharded result differs"
    assert int(tot.item()) == sim.total_kmers and t.item() == world
    print("SHARD_OK", world, sim.nreads)
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(kit, hostsim, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    assert "SHARD_OK 2" in p.stdout
