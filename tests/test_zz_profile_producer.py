"""Profile producer: exact canonical k-mer counts of a read set at every read
position, the count histogram, and the encoder side of the FastK profile codec -- what FastK does
before ClassPro runs (FastK itself is not in the reference tree; the reference only reads its files).

Checker: the harness counter and encoder of tools/cpsim.c (exact mode), whose files the unmodified
reference binary reads in every file-level test of this suite, and the oracle's decoder
(src/libfastk.c:1467-1535 restated) for the round trip.

CPU tests run the kernels' element functions (classpro_b200/csrc/cpg_count.cuh) through the test-only
host build tests/hostsim/countsim.cpp; GPU tests call cpg_count_kmers / cpg_encode_profiles of
libclasspro_b200.so.  (The file sorts last on purpose: with the round's last GPU seconds the product calls
were compared with the harness on a B200 through tools/producer_check.py -- equal counts, histogram,
offsets and bytes for K = 40, 32, 21, profiles/r01_producer_check.log -- but not in this pytest form.)"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_SO = os.path.join(ROOT, "tests", "hostsim", "_build", "libcountsim.so")


def _bind(L, count, encode):
    f = getattr(L, count)
    f.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    g = getattr(L, encode)
    g.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    return f, g


class Producer:
    def __init__(self, L, count, encode, errfn=None):
        self.count_fn, self.encode_fn = _bind(L, count, encode)
        self.errfn = errfn

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("rc %d: %s" % (rc, self.errfn().decode() if self.errfn else ""))

    def count(self, K, pseq, seq_off, rlen):
        n = len(rlen)
        rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        pseq = np.ascontiguousarray(pseq, dtype=np.uint8)
        cnt_off = np.zeros(n + 1, dtype=np.int64)
        total = int(np.maximum(rlen.astype(np.int64) - K + 1, 0).sum())
        counts = np.zeros(max(total, 1), dtype=np.uint16)
        hist = np.zeros(32770, dtype=np.int64)
        self._check(self.count_fn(0, K, n, pseq.ctypes.data, seq_off.ctypes.data, rlen.ctypes.data,
                                  cnt_off.ctypes.data, counts.ctypes.data, hist.ctypes.data))
        assert cnt_off[n] == total
        return counts[:total], cnt_off, hist

    def encode(self, counts, cnt_off, cap=None):
        n = len(cnt_off) - 1
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        cnt_off = np.ascontiguousarray(cnt_off, dtype=np.int64)
        cap = 2 * len(counts) + 16 if cap is None else cap
        prof = np.zeros(max(cap, 1), dtype=np.uint8)
        prof_off = np.zeros(n + 1, dtype=np.int64)
        self._check(self.encode_fn(0, n, counts.ctypes.data, cnt_off.ctypes.data, prof.ctypes.data, cap, prof_off.ctypes.data))
        return prof[:prof_off[n]], prof_off


@pytest.fixture(scope="module")
def hostsim(kit):
    kit.build_hostsim()
    return Producer(C.CDLL(SIM_SO), "sim_count_kmers", "sim_encode_profiles")


@pytest.fixture(scope="module")
def device():
    from classpro_b200 import abi
    L = abi.lib()
    L.cpg_count_error.restype = C.c_char_p
    return Producer(L, "cpg_count_kmers", "cpg_encode_profiles", L.cpg_count_error)


def greedy_encode(c):
    """tools/cpsim.c:241-268 restated: the token choice the harness files are written with."""
    out = bytearray()
    if len(c) == 0:
        return bytes(out)
    d = int(c[0])
    out += bytes([0x80 | (d >> 8), d & 0xff]) if d >= 128 else bytes([d])
    i = 1
    while i < len(c):
        if int(c[i]) == d:
            run = 1
            while i + run < len(c) and int(c[i + run]) == d and run < 63:
                run += 1
            out.append(run)
            i += run
            continue
        diff = int(c[i]) - d
        if -32 <= diff <= 31:
            out.append(0x40 | (diff & 0x3f))
        else:
            x = diff & 0x7fff
            out += bytes([0x80 | (x >> 8), x & 0xff])
        d = int(c[i])
        i += 1
    return bytes(out)


def adversarial_counts(rng, n_reads):
    reads = []
    for r in range(n_reads):
        kind = r % 8
        if kind == 0:
            c = np.zeros(0, dtype=np.uint16)                      # read shorter than K
        elif kind == 1:
            c = np.array([rng.choice([1, 127, 128, 32767])], dtype=np.uint16)
        elif kind == 2:                                           # runs of exactly 62..64, 125..128 equal counts
            parts = [np.full(m + 1, v, dtype=np.uint16) for m, v in
                     zip([62, 63, 64, 125, 126, 127, 128, 1, 2, 189], rng.integers(1, 300, 10))]
            c = np.concatenate(parts)
        elif kind == 3:                                           # deltas around the one- / two-byte boundary
            steps = rng.choice([-33, -32, -31, -1, 1, 31, 32, 33, 0, 0], 400)
            c = np.clip(1000 + np.cumsum(steps), 1, 32767).astype(np.uint16)
        elif kind == 4:                                           # extremes
            c = rng.choice([1, 2, 32767, 16384, 127, 128, 129], 300).astype(np.uint16)
        elif kind == 5:                                           # one long plateau
            c = np.full(int(rng.integers(1, 700)), int(rng.integers(1, 32768)), dtype=np.uint16)
        else:                                                     # HiFi-like: plateaus with dips
            c = np.repeat(rng.integers(1, 60, 40), rng.integers(1, 200, 40)).astype(np.uint16)
        reads.append(c)
    off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum([len(c) for c in reads], out=off[1:])
    return reads, (np.concatenate(reads) if reads else np.zeros(0, np.uint16)), off


def check_counts(P, kit, K, **kw):
    from classpro_b200 import abi
    sim = kit.simulate(kmer=K, **kw)
    pseq, seq_off = abi.pack_codes(sim.seq, sim.seq_off, sim.rlen)
    counts, cnt_off, hist = P.count(K, pseq, seq_off, sim.rlen)
    assert np.array_equal(cnt_off, sim.cnt_off)
    assert np.array_equal(counts, sim.counts), np.flatnonzero(counts != sim.counts)[:10]
    assert np.array_equal(hist, sim.hist), np.flatnonzero(hist != sim.hist)[:10]
    assert int((hist[1:32768] * np.arange(1, 32768)).sum()) >= sim.total_kmers - int(hist[32767]) * 32767
    return sim, counts, cnt_off


def check_encoder(P, kit, rng):
    L = kit.oracle_lib()
    reads, flat, off = adversarial_counts(rng, 64)
    prof, prof_off = P.encode(flat, off)
    for r, c in enumerate(reads):
        got = prof[prof_off[r]:prof_off[r + 1]].tobytes()
        assert got == greedy_encode(c), r
        back = np.zeros(len(c) + 8, dtype=np.uint16)
        buf = np.frombuffer(got, dtype=np.uint8).copy() if got else np.zeros(1, np.uint8)
        n = L.cpo_decode_profile(buf.ctypes.data, len(got), back.ctypes.data, len(back))
        assert n == len(c) and np.array_equal(back[:n], c), r
    with pytest.raises(RuntimeError):
        P.encode(flat, off, cap=int(prof_off[-1]) - 1)
    e, eo = P.encode(np.zeros(0, np.uint16), np.zeros(4, np.int64))
    assert len(e) == 0 and not eo.any()


def check_strands(P):
    """A read and its reverse complement share every k-mer; a k-mer that is its own reverse
    complement counts once per occurrence."""
    from classpro_b200 import abi
    rng = np.random.default_rng(3)
    for K in (5, 16, 31, 32, 33, 40):
        x = rng.integers(0, 4, 300).astype(np.uint8)
        rc = (3 - x)[::-1].copy()
        pal = np.concatenate([x[:K // 2], (3 - x[:K // 2])[::-1]]) if K % 2 == 0 else x[:K]
        codes = np.concatenate([x, rc, pal, x[:K - 1]])
        rlen = np.array([len(x), len(rc), len(pal), K - 1], dtype=np.int32)
        seq_off = np.concatenate([[0], np.cumsum(rlen)]).astype(np.int64)
        pseq, poff = abi.pack_codes(codes, seq_off, rlen)
        counts, cnt_off, hist = P.count(K, pseq, poff, rlen)
        # brute force on (k-mer, reverse complement) tuples
        table = {}
        keys = []
        for r in range(4):
            s = codes[seq_off[r]:seq_off[r + 1]]
            for p in range(len(s) - K + 1):
                f = tuple(s[p:p + K]); b = tuple((3 - s[p:p + K])[::-1])
                k = min(f, b)
                keys.append(k)
                table[k] = table.get(k, 0) + 1
        want = np.array([table[k] for k in keys], dtype=np.uint16)
        assert np.array_equal(counts, want), K
        assert cnt_off[4] == cnt_off[3] and hist[1:].sum() - hist[32768] - hist[32769] == len(table)


# ------------------------------------------------------------------------------ CPU: the element functions
@pytest.mark.parametrize("K", [40, 21, 32, 33])
def test_counts_and_histogram_hostsim(kit, hostsim, K):
    sim, counts, cnt_off = check_counts(hostsim, kit, K, seed=11 + K, genome_len=30000, cov=12., het=0.01, len_mean=3000,
                                        short_reads=1, repeat_frac=0.2)
    prof, prof_off = hostsim.encode(counts, cnt_off)
    assert np.array_equal(prof_off, sim.prof_off) and np.array_equal(prof, sim.prof)


@pytest.mark.parametrize("passes", [2, 7, 64])
def test_key_range_passes_hostsim(kit, hostsim, monkeypatch, passes):
    """Read sets whose keys do not fit in device memory at once are counted in passes over hashed key
    ranges (CPG_COUNT_PASSES forces them): same counts and histogram."""
    monkeypatch.setenv("CPG_COUNT_PASSES", str(passes))
    check_counts(hostsim, kit, 40, seed=3, genome_len=20000, cov=10., het=0.01, len_mean=3000, short_reads=1, repeat_frac=0.3)
    check_strands(hostsim)


@pytest.mark.parametrize("npass,grid", [(3, 5), (7, 2)])
def test_pass_kernels_on_host_threads(kit, hostsim, npass, grid):
    """The source text of k_pass_sizes / k_kmer_keys_pass (the only producer kernels that are not loops
    around an element function) on host threads: 256 threads per CTA, warp votes and shuffles as
    rendezvous (tests/hostsim/passemu.cpp).  Every k-mer is appended in exactly one pass, equal keys
    in the same pass, as many as the sizing kernel says, and counting the appended keys gives the
    harness counts."""
    from classpro_b200 import abi
    L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "_build", "libpassemu.so"))
    L.pe_pass_sizes.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.pe_keys_pass.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_uint64, C.c_void_p, C.c_void_p]
    L.pe_keys_pass.restype = C.c_uint64
    K = 40
    sim = kit.simulate(kmer=K, seed=4, genome_len=6000, cov=8., het=0.01, len_mean=1500, short_reads=1, repeat_frac=0.3)
    pseq, seq_off = abi.pack_codes(sim.seq, sim.seq_off, sim.rlen)
    n, nr = sim.total_kmers, sim.nreads
    cnt_off = np.ascontiguousarray(sim.cnt_off, np.int64)
    sizes = np.zeros(64, np.uint64)
    L.pe_pass_sizes(grid, nr, pseq.ctypes.data, seq_off.ctypes.data, cnt_off.ctypes.data, K, npass, sizes.ctypes.data)
    assert int(sizes.sum()) == n and not sizes[npass:].any()
    seen = np.zeros(n, np.int32)
    where = {}
    counts = np.zeros(n, np.int64)
    for p in range(npass):
        cap = int(sizes[p])
        klo = np.zeros(cap + 2, np.uint64); khx = np.zeros(cap + 2, np.uint64)
        got = L.pe_keys_pass(grid, nr, pseq.ctypes.data, seq_off.ctypes.data, cnt_off.ctypes.data, K, p, npass, cap,
                             klo.ctypes.data, khx.ctypes.data)
        assert got == cap, (p, got, cap)
        klo, khx = klo[:cap], khx[:cap]
        idx = (khx & np.uint64((1 << 48) - 1)).astype(np.int64)
        np.add.at(seen, idx, 1)
        keys = list(zip(klo.tolist(), (khx >> np.uint64(48)).tolist()))
        tally = {}
        for k in keys:
            assert where.setdefault(k, p) == p
            tally[k] = tally.get(k, 0) + 1
        counts[idx] = [tally[k] for k in keys]
    assert (seen == 1).all()
    assert np.array_equal(np.minimum(counts, 32767).astype(np.uint16), sim.counts)


def test_encoder_hostsim(kit, hostsim):
    check_encoder(hostsim, kit, np.random.default_rng(5))


def test_strands_and_palindromes_hostsim(hostsim):
    check_strands(hostsim)


def test_saturation_hostsim(hostsim):
    """One 12-mer 40 000 times: count 32767 everywhere, one distinct k-mer in the top bin with its instances."""
    from classpro_b200 import abi
    K = 12
    codes = np.zeros(40000 + K - 1, dtype=np.uint8)              # poly-A
    rlen = np.array([len(codes)], dtype=np.int32)
    pseq, poff = abi.pack_codes(codes, np.array([0, len(codes)], dtype=np.int64), rlen)
    counts, cnt_off, hist = hostsim.count(K, pseq, poff, rlen)
    assert (counts == 32767).all() and hist[32767] == 1 and hist[32769] == 40000 and hist[1:32767].sum() == 0


def test_producer_errors_without_gpu():
    """No CPU fallback: on a machine without a CUDA device the product calls fail with a message
    (argument errors are reported before the device is touched)."""
    import torch
    from classpro_b200 import abi
    L = abi.lib()
    L.cpg_count_error.restype = C.c_char_p
    P = Producer(L, "cpg_count_kmers", "cpg_encode_profiles", L.cpg_count_error)
    rlen = np.array([100], dtype=np.int32)
    with pytest.raises(RuntimeError, match="K <= 40"):
        P.count(41, np.zeros(64, np.uint8), np.array([0, 25], np.int64), rlen)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            P.count(40, np.zeros(64, np.uint8), np.array([0, 25], np.int64), rlen)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            P.encode(np.ones(10, np.uint16), np.array([0, 10], np.int64))


FILESET = ["{r}.hist", "{r}.prof", ".{r}.pidx.1", ".{r}.prof.1", ".{r}.pidx.2", ".{r}.prof.2", ".{r}.pidx.3", ".{r}.prof.3"]


def check_program(kit, exe, tmp_path):
    """The `profiler` program on a FASTA file: the same .hist / .prof / .pidx.N / .prof.N bytes as the
    harness writer (whose files the unmodified reference binary reads in the file-level tests), from
    FASTA, from wrapped FASTQ.gz, and with -N; the reference classifies from the produced files."""
    import filecmp, gzip, shutil, subprocess
    d = str(tmp_path / "sim"); o = str(tmp_path / "out")
    os.makedirs(o)
    sim = kit.simulate(write_to=d, root="q", seed=21, genome_len=30000, cov=14., het=0.01, len_mean=4000, short_reads=1, nparts=3)
    shutil.copy(os.path.join(d, "q.fasta"), os.path.join(o, "q.fasta"))
    p = subprocess.run([exe, "-v", "-p3", os.path.join(o, "q.fasta")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    assert "%d 40-mers" % sim.total_kmers in p.stderr
    for f in FILESET:
        assert filecmp.cmp(os.path.join(d, f.format(r="q")), os.path.join(o, f.format(r="q")), shallow=False), f
    with gzip.open(os.path.join(o, "z.fastq.gz"), "wb") as g:
        for i in range(sim.nreads):
            s = sim.read_ascii(i).tobytes()
            g.write(b"@r%d\n" % i + s[:50] + b"\n" + s[50:] + b"\n+\n" + b"@" * len(s) + b"\n")
    p = subprocess.run([exe, "-p3", "-N" + os.path.join(o, "w"), os.path.join(o, "z.fastq.gz")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-800:]
    for f in FILESET:
        assert filecmp.cmp(os.path.join(d, f.format(r="q")), os.path.join(o, f.format(r="w")), shallow=False), f
    if kit.have_reference():
        a = kit.run_reference(os.path.join(d, "q.fasta"), threads=1)
        b = kit.run_reference(os.path.join(o, "q.fasta"), threads=1)
        assert filecmp.cmp(a, b, shallow=False)
    open(os.path.join(o, "n.fasta"), "wb").write(b">r\nACGTNACGT\n")
    p = subprocess.run([exe, os.path.join(o, "n.fasta")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "outside ACGT" in p.stderr
    p = subprocess.run([exe, "-k41", os.path.join(o, "q.fasta")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "[1,40]" in p.stderr


def test_profiler_program_hostsim(kit, hostsim, tmp_path):
    check_program(kit, os.path.join(ROOT, "tests", "hostsim", "_build", "profiler"), tmp_path)


# ------------------------------------------------------------------------------ GPU: the product path
@pytest.mark.gpu
@pytest.mark.parametrize("K", [40, 21, 32, 33])
def test_counts_and_histogram_gpu(kit, device, K):
    sim, counts, cnt_off = check_counts(device, kit, K, seed=11 + K, genome_len=200000, cov=15., het=0.01, len_mean=8000,
                                        short_reads=1, repeat_frac=0.2)
    prof, prof_off = device.encode(counts, cnt_off)
    assert np.array_equal(prof_off, sim.prof_off) and np.array_equal(prof, sim.prof)


@pytest.mark.gpu
@pytest.mark.parametrize("passes", [3, 64])
def test_key_range_passes_gpu(kit, device, monkeypatch, passes):
    monkeypatch.setenv("CPG_COUNT_PASSES", str(passes))
    check_counts(device, kit, 40, seed=3, genome_len=100000, cov=12., het=0.01, len_mean=6000, short_reads=1, repeat_frac=0.3)
    check_strands(device)


@pytest.mark.gpu
def test_encoder_and_strands_gpu(kit, device):
    check_encoder(device, kit, np.random.default_rng(5))
    check_strands(device)


@pytest.mark.gpu
def test_produced_profiles_feed_the_classifier_gpu(kit, device, tmp_path):
    """End to end without the harness counter: reads -> cpg_count_kmers -> cpg_encode_profiles ->
    model from the produced histogram -> cpg_classify == the oracle on the harness files."""
    import classpro_b200 as cp
    from test_gpu import make_batch, compare_with_oracle
    sim = kit.simulate(seed=9, genome_len=150000, cov=25., het=0.01, len_mean=8000)
    pseq, seq_off = cp.abi.pack_codes(sim.seq, sim.seq_off, sim.rlen)
    counts, cnt_off, hist = device.count(40, pseq, seq_off, sim.rlen)
    prof, prof_off = device.encode(counts, cnt_off)
    assert np.array_equal(hist, sim.hist) and np.array_equal(prof, sim.prof)
    sim.counts, sim.cnt_off, sim.hist, sim.prof, sim.prof_off = counts, cnt_off, hist, prof, prof_off     # produced, not simulated
    om = kit.oracle_model(sim)
    gm = cp.Model.from_hist(sim.kmer, hist[1:32768], hist[32768], hist[32769])
    ctx = cp.Context(gm)
    batch, keep = make_batch(cp, sim)
    cls, status = ctx.classify(batch)
    ctx.close()
    assert not (status & cp.ST_FATAL).any()
    kmers, flips = compare_with_oracle(kit, sim, batch, keep, cls, om)
    assert flips <= int(kmers * 1e-6)


@pytest.mark.gpu
def test_profiler_program_gpu(kit, device, tmp_path):
    check_program(kit, os.path.join(ROOT, "classpro_b200", "profiler"), tmp_path)
