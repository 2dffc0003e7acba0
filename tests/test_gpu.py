"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle, the golden
fixtures of the reference and -- where present -- the reference binary.  Need a B200 (-m gpu).

Bar: bit-exact counts for the decoder; for the class strings, byte identity with the oracle, with
the north-star's floating-point allowance (characters that flip because the CUDA and glibc exp/log
differ in the last place) bounded by FLIP_BUDGET = 1e-6 of the classified k-mers and reported."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FLIP_BUDGET = 1e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "classpro_b200", "ClassPro")


def make_batch(cp, sim, keep=None, seq_bits=2):
    from classpro_b200.abi import pack_codes
    if keep is None:
        keep = np.nonzero(sim.rlen >= sim.kmer)[0]
    rl = sim.rlen[keep].astype(np.int32)
    so = np.zeros(len(keep) + 1, np.int64)
    np.cumsum(rl, out=so[1:])
    codes = np.concatenate([sim.seq[sim.seq_off[i]:sim.seq_off[i + 1]] for i in keep]) if len(keep) else np.zeros(0, np.uint8)
    if seq_bits == 2:
        seq, soff = pack_codes(codes, so, rl)
    else:
        seq, soff = np.frombuffer(b"ACGT", dtype=np.uint8)[codes], so
    parts = [sim.read_prof(i) for i in keep]
    prof = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    po = np.zeros(len(keep) + 1, np.int64)
    np.cumsum([len(p) for p in parts], out=po[1:])
    return cp.Batch(seq, soff, rl, prof, po, seq_bits), keep


def compare_with_oracle(kit, sim, batch, keep, cls, om):
    ow = kit.OracleWork(clean=True)
    kmers = flips = 0
    for k, i in enumerate(keep):
        a = ow.classify(om, sim.read_ascii(i).tobytes(), sim.read_counts(i))
        b = cls[batch.cls_off[k]:batch.cls_off[k + 1]].tobytes()
        kmers += len(a) - sim.kmer + 1
        if a != b:
            flips += sum(1 for x, y in zip(a, b) if x != y)
    return kmers, flips


@pytest.fixture(scope="module")
def cp():
    import classpro_b200
    assert classpro_b200.lib().cpg_device_count() > 0, "no CUDA device: the -m gpu tests need a B200"
    return classpro_b200


DATASETS = [
    ("plain", dict(seed=71, genome_len=150000, cov=30., het=0.006), 0, 20000),
    ("repeats", dict(seed=72, genome_len=200000, cov=40., het=0.01, repeat_frac=0.5, seg_dups=3), 0, 20000),
    ("hicov_long", dict(seed=73, genome_len=120000, cov=100., het=0.01, repeat_frac=0.3, len_mean=25000, len_sd=3000), 0, 25000),
    ("noisy_lc", dict(seed=74, genome_len=100000, cov=35., het=0.001, err_indel_hp=0.003, err_sub=0.002, repeat_frac=0.8), 0, 20000),
    ("lowcov_opts", dict(seed=75, genome_len=100000, cov=12., het=0.01, len_mean=10000, short_reads=1), 11, 10000),
]


@pytest.mark.parametrize("knob", ["CPG_SCRATCH_DIV", "CPG_POOL_DIV", "CPG_HDR_DIV", "CPG_BIG_DIV", "CPG_FUSED"])
def test_retry_launch_matches_oracle(kit, cp, monkeypatch, knob):
    """The other routes through the kernels give the same class strings.
    CPG_SCRATCH_DIV: interval tables of 64 entries in the scratch blocks of the phase kernels, so
    nearly every read is flagged and classified by the retry launch with full-size tables.
    CPG_POOL_DIV: an interval pool of 4096 entries for the batch: the first reads fit, the rest are
    flagged by k_wall_b (a mix of both routes in one batch).
    CPG_HDR_DIV / CPG_BIG_DIV: 4096 candidate headers / big candidate records for the batch: the reads that
    find the arrays full are flagged by k_wall_a.
    CPG_FUSED: every read through the single-kernel path."""
    monkeypatch.setenv(knob, "1" if knob == "CPG_FUSED" else "1000000")
    name, params, cov_opt, read_len = DATASETS[1]
    sim = kit.simulate(**params)
    om = kit.oracle_model(sim, cov_opt, read_len)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=read_len)
    ctx = cp.Context(gm)
    batch, keep = make_batch(cp, sim)
    cls, status = ctx.classify(batch)
    ctx.close()
    assert not (status & cp.ST_FATAL).any() and not (status & (1 << 20)).any()
    kmers, flips = compare_with_oracle(kit, sim, batch, keep, cls, om)
    assert flips <= max(0, int(kmers * FLIP_BUDGET)), "%d of %d k-mers differ from the oracle" % (flips, kmers)


@pytest.mark.parametrize("knob", [None, "CPG_SCRATCH_DIV", "CPG_FUSED"])
def test_interval_results_expand_to_the_class_strings(kit, cp, monkeypatch, knob):
    """CPG_RESULT_INTERVALS: the packed interval tables of a batch, expanded on the host, are the class strings
    of CPG_RESULT_CLASSES mode byte for byte -- on the main path, with most reads through the retry launch, and
    on the single-kernel path."""
    from classpro_b200 import abi
    if knob:
        monkeypatch.setenv(knob, "1" if knob == "CPG_FUSED" else "1000000")
    name, params, cov_opt, read_len = DATASETS[1]
    sim = kit.simulate(**params)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=read_len)
    ctx = cp.Context(gm)
    batch, keep = make_batch(cp, sim)
    cls, status = ctx.classify(batch)
    ctx.set_result_mode(True)
    with pytest.raises(abi.CpgError):
        ctx.submit(0, batch)
        ctx.collect(0, batch)                      # wrong call for this mode; the batch stays in flight
    res = abi.IntervalResult(batch.n, ctx.intervals_bound(0))
    ctx.collect_intervals(0, res)
    assert (res.status[:batch.n] == status).all()
    assert 0 < res.used <= res.cap and int(res.cnt[:batch.n].sum()) <= res.used
    ex = res.expand(sim.kmer, batch.rlen)
    nb = int(batch.cls_off[-1])
    assert np.array_equal(ex[:nb], cls[:nb])
    small = abi.IntervalResult(batch.n, 8)
    ctx.submit(1, batch)
    with pytest.raises(abi.CpgError):
        ctx.collect_intervals(1, small)            # too small: rejected, still in flight
    ctx.collect_intervals(1, res)
    assert np.array_equal(res.expand(sim.kmer, batch.rlen)[:nb], cls[:nb])
    ctx.set_result_mode(False)
    cls2, _ = ctx.classify(batch)
    assert np.array_equal(cls2[:nb], cls[:nb])
    ctx.close()


@pytest.mark.parametrize("name,params,cov_opt,read_len", DATASETS, ids=[d[0] for d in DATASETS])
def test_classify_matches_oracle(kit, cp, name, params, cov_opt, read_len):
    sim = kit.simulate(**params)
    om = kit.oracle_model(sim, cov_opt, read_len)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=read_len)
    ctx = cp.Context(gm)
    batch, keep = make_batch(cp, sim)
    cls, status = ctx.classify(batch)
    ctx.close()
    assert not (status & cp.ST_FATAL).any()
    kmers, flips = compare_with_oracle(kit, sim, batch, keep, cls, om)
    print("%s: %d k-mers, %d flipped" % (name, kmers, flips))
    assert flips <= max(0, int(kmers * FLIP_BUDGET)), "%d of %d k-mers differ from the oracle" % (flips, kmers)


def test_decode_matches_counts(kit, cp):
    sim = kit.simulate(seed=76, genome_len=200000, cov=40., het=0.01, repeat_frac=0.5)
    ctx = cp.Context(cp.Model.from_cov(40, 0, 30))
    caps = np.maximum(sim.rlen.astype(np.int64) - sim.kmer + 1, 0)
    counts, cnt_off, plen = ctx.decode_profiles(sim.prof, sim.prof_off, caps)
    assert np.array_equal(plen, caps)
    assert np.array_equal(counts[:sim.total_kmers], sim.counts)
    ctx.close()


def test_decode_adversarial_streams(kit, cp):
    """Arbitrary byte streams: wrap-around, 15-bit mask, zero-length runs, truncation by the cap."""
    from test_device_logic_hostsim import random_stream
    rng = np.random.default_rng(77)
    streams = [random_stream(rng, int(rng.integers(0, 600)), bool(i & 1)) for i in range(400)]
    streams.append(b"")
    exp = [kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 200000) for s in streams]
    prof = np.frombuffer(b"".join(streams), dtype=np.uint8)
    po = np.zeros(len(streams) + 1, np.int64)
    np.cumsum([len(s) for s in streams], out=po[1:])
    caps = np.array([max(1, n // (1 + (i % 3 == 0))) for i, (n, _) in enumerate(exp)], dtype=np.int64)
    ctx = cp.Context(cp.Model.from_cov(40, 0, 30))
    counts, cnt_off, plen = ctx.decode_profiles(prof, po, caps)
    ctx.close()
    for i, (n, o) in enumerate(exp):
        assert plen[i] == n, i
        m = min(n, int(caps[i]))
        assert np.array_equal(counts[cnt_off[i]:cnt_off[i] + m], o[:m]), i


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_golden_fixture_through_cli(kit, cp, tmp_path, name):
    """Reference-generated golden .class reproduced by the ClassPro CLI of this repo."""
    from test_oracle import unpack_golden
    fasta, golden, args = unpack_golden(name, str(tmp_path))
    p = subprocess.run([CLI, "-v"] + args + [fasta], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-1500:]
    out = fasta[:-6] + ".class"
    if not filecmp.cmp(out, golden, shallow=False):
        a, b = kit.class_lines(out), kit.class_lines(golden)
        assert len(a) == len(b)
        hdr_a = open(out, "rb").read().split(b"\n")[0::4]
        hdr_b = open(golden, "rb").read().split(b"\n")[0::4]
        assert hdr_a == hdr_b
        flips = sum(sum(1 for x, y in zip(u, v) if x != y) + abs(len(u) - len(v)) for u, v in zip(a, b))
        kmers = sum(max(len(v) - 39, 0) for v in b)
        assert flips <= int(kmers * FLIP_BUDGET), "%d of %d k-mers differ from the reference" % (flips, kmers)


def test_cli_matches_live_reference(kit, cp, tmp_path):
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    kit.simulate(write_to=str(tmp_path), root="x", seed=78, genome_len=120000, cov=30., het=0.01, repeat_frac=0.3,
                 short_reads=1, nparts=3)
    fasta = str(tmp_path / "x.fasta")
    ref = kit.run_reference(fasta, threads=2)
    os.rename(ref, ref + ".ref")
    # small batches (many hand-overs between reader, GPU worker and writer) and a single big one,
    # with one and with several packing threads
    for args in (["-B2"], ["-B1", "-T1"], ["-B1", "-T7"], ["-B500", "-T3"]):
        p = subprocess.run([CLI] + args + [fasta], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-1500:]
        assert filecmp.cmp(ref, ref + ".ref", shallow=False), "CLI %s output differs from the reference binary's" % args
        os.remove(ref)


def test_cli_every_gpu_count_matches_reference(kit, cp, tmp_path):
    """The real multi-GPU shape: ONE read set, contiguous batches of reads handed to one worker per
    GPU, one writer that restores read order.  For every n <= devices present: ClassPro -G<n> gives the bytes
    of the reference binary, and every one of the n GPUs classified some of the reads."""
    import re
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    ndev = cp.lib().cpg_device_count()
    kit.simulate(write_to=str(tmp_path), root="m", seed=91, genome_len=1500000, cov=30., het=0.01, repeat_frac=0.2,
                 len_mean=15000, len_sd=3000, short_reads=1, nparts=5)
    fasta = str(tmp_path / "m.fasta")
    ref = kit.run_reference(fasta, threads=4)
    os.rename(ref, ref + ".ref")
    for n in range(1, ndev + 1):
        p = subprocess.run([CLI, "-v", "-G%d" % n, "-B1", "-T4", fasta], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr[-1500:]
        assert filecmp.cmp(ref, ref + ".ref", shallow=False), "ClassPro -G%d output differs from the reference binary's" % n
        os.remove(ref)
        m = re.search(r"per GPU \(batches/k-mers\):((?: \d+/\d+)+)", p.stderr)
        assert m, p.stderr[-800:]
        per = [tuple(int(x) for x in t.split("/")) for t in m.group(1).split()]
        assert len(per) == n and all(b > 0 and k > 0 for b, k in per), "a GPU stayed idle: %s" % per
        print("-G%d: per GPU (batches, k-mers) %s" % (n, per))


def test_edge_batches(kit, cp):
    sim = kit.simulate(seed=79, genome_len=60000, cov=25., het=0.01, len_mean=9000)
    om = kit.oracle_model(sim)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769])
    ctx = cp.Context(gm)
    # empty batch
    empty, _ = make_batch(cp, sim, keep=np.zeros(0, dtype=np.int64))
    cls, st = ctx.classify(empty)
    assert len(st) == 0
    # one read
    one, keep = make_batch(cp, sim, keep=np.array([3]))
    cls, st = ctx.classify(one)
    assert compare_with_oracle(kit, sim, one, keep, cls, om)[1] == 0
    # raw-byte sequences (seq_bits = 8) give the same classes as packed ones
    b2, keep = make_batch(cp, sim)
    b8, _ = make_batch(cp, sim, seq_bits=8)
    c2, _ = ctx.classify(b2)
    c8, _ = ctx.classify(b8)
    assert np.array_equal(c2[:b2.cls_off[-1]], c8[:b8.cls_off[-1]])
    # double-buffered submit/collect == synchronous classify; order of reads does not matter
    half = len(keep) // 2
    ba, ka = make_batch(cp, sim, keep=keep[:half])
    bb, kb = make_batch(cp, sim, keep=keep[half:][::-1].copy())
    ctx.submit(0, ba)
    ctx.submit(1, bb)
    ca, _ = ctx.collect(0, ba)
    cb, _ = ctx.collect(1, bb)
    assert compare_with_oracle(kit, sim, ba, ka, ca, om)[1] == 0
    assert compare_with_oracle(kit, sim, bb, kb, cb, om)[1] == 0
    # a profile that does not match its read length is reported, not classified (ClassPro.c:234-237)
    i0, i1 = 0, 1
    while sim.rlen[i1] == sim.rlen[i0]:
        i1 += 1
    ok, _ = make_batch(cp, sim, keep=np.array([i0, i1]))
    swapped = np.concatenate([sim.read_prof(i1), sim.read_prof(i0)])
    po = np.array([0, len(sim.read_prof(i1)), len(swapped)], dtype=np.int64)
    bad = cp.Batch(ok.seq, ok.seq_off, ok.rlen, swapped, po, 2)
    cls, st = ctx.classify(bad)
    assert (st[0] & 1) and (st[1] & 1)
    with pytest.raises(cp.CpgError):
        ctx.classify(bad, allow_read_errors=False)
    ctx.close()


def test_minimal_and_maximal_reads(kit, cp):
    """plen = 1 and the longest read the reference accepts (60000 bases)."""
    sim = kit.simulate(seed=80, genome_len=120000, cov=2., het=0.01, len_mean=60000, len_sd=0, len_min=60000,
                       len_max=60000 - 64)
    sim2 = kit.simulate(seed=81, genome_len=30000, cov=20., het=0.01)
    om = kit.oracle_model(sim2)
    gm = cp.Model.from_hist(sim2.kmer, sim2.hist[1:32768], sim2.hist[32768], sim2.hist[32769])
    ctx = cp.Context(gm)
    assert sim.rlen.max() <= 60000
    b, keep = make_batch(cp, sim)
    cls, st = ctx.classify(b)
    assert not (st & cp.ST_FATAL).any()
    assert compare_with_oracle(kit, sim, b, keep, cls, om)[1] == 0
    # reads cut down to K and K+1 bases
    ow = kit.OracleWork(clean=True)
    from classpro_b200.abi import pack_reads
    s, c = sim2.read_ascii(0).tobytes(), sim2.read_counts(0)
    reads, profs = [], []
    for n in (1, 2, 5, 41):
        reads.append(s[:n + 39])
        enc = [int(c[0])] if c[0] < 128 else [0x80 | (int(c[0]) >> 8), int(c[0]) & 0xff]
        for j in range(1, n):
            d = int(c[j]) - int(c[j - 1])
            enc += [0x40 | (d & 0x3f)] if -32 <= d <= 31 and d != 0 else ([1] if d == 0 else [0x80 | ((d & 0x7fff) >> 8), d & 0xff])
        profs.append(bytes(enc))
    pk, po = pack_reads(reads)
    pr = np.frombuffer(b"".join(profs), dtype=np.uint8)
    pro = np.concatenate([[0], np.cumsum([len(p) for p in profs])]).astype(np.int64)
    b = cp.Batch(pk, po, np.array([len(r) for r in reads], dtype=np.int32), pr, pro, 2)
    cls, st = ctx.classify(b)
    assert not st.any()
    for k, n in enumerate((1, 2, 5, 41)):
        a = ow.classify(om, reads[k], c[:n])
        assert cls[b.cls_off[k]:b.cls_off[k + 1]].tobytes() == a
    ctx.close()


def test_full_size_properties(kit, cp):
    """At bench scale (ground-truth-coverage profiles) the oracle is too slow for every read:
    check size-independent properties -- idempotence, batch-split invariance, class alphabet, the
    'N' prefix -- and spot-check a sample of reads against the oracle."""
    sim = kit.simulate(seed=82, genome_len=3000000, cov=30., het=0.01, snp_only=1, exact=0, len_mean=20000, len_sd=2000,
                       len_min=5000)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769])
    om = kit.oracle_model(sim)
    ctx = cp.Context(gm)
    b, keep = make_batch(cp, sim)
    c1, st = ctx.classify(b)
    assert not (st & cp.ST_FATAL).any()
    c2, _ = ctx.classify(b)
    assert np.array_equal(c1, c2)                                  # idempotent
    total = int(b.cls_off[-1])
    assert set(np.unique(c1[:total]).tolist()) <= set(b"NEHDR")
    for k in range(0, len(keep), 97):
        assert c1[b.cls_off[k]:b.cls_off[k] + 39].tobytes() == b"N" * 39
        assert b"N" not in c1[b.cls_off[k] + 39:b.cls_off[k + 1]].tobytes()
    h = len(keep) // 3                                             # split invariance
    bx, kx = make_batch(cp, sim, keep=keep[h:2 * h])
    cx, _ = ctx.classify(bx)
    assert np.array_equal(cx[:bx.cls_off[-1]], c1[b.cls_off[h]:b.cls_off[2 * h]])
    sample = keep[::max(1, len(keep) // 60)]
    bs, ks = make_batch(cp, sim, keep=sample)
    cs, _ = ctx.classify(bs)
    kmers, flips = compare_with_oracle(kit, sim, bs, ks, cs, om)
    assert flips <= int(kmers * FLIP_BUDGET)
    ctx.close()


def test_cli_fastq_gz_and_header_quirks(kit, cp, tmp_path):
    """FASTQ.gz input, and the header quirks that are part of the byte contract:
    no comment on the first record -> "(null)", a tab-separated comment, a comment-less record after
    one with a comment -> the stale comment is printed again; wrapped sequence lines."""
    import gzip
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    sim = kit.simulate(write_to=str(tmp_path), root="q", seed=83, genome_len=40000, cov=20., het=0.01, len_mean=7000,
                       short_reads=1, nparts=2)
    os.remove(str(tmp_path / "q.fasta"))
    with gzip.open(str(tmp_path / "q.fastq.gz"), "wb") as f:
        for i in range(sim.nreads):
            s = sim.read_ascii(i).tobytes()
            if i == 0:
                hdr = b"@r0"
            elif i % 5 == 1:
                hdr = b"@r%d\tcomm ent %d" % (i, i)
            elif i % 5 == 2:
                hdr = b"@r%d" % i
            else:
                hdr = b"@" + sim.headers[i].replace(b"Sim ", b"Sim_", 1)
            f.write(hdr + b"\n")
            if i % 3 == 0 and len(s) > 100:          # wrapped sequence (FASTA style wrapping is legal in kseq)
                f.write(s[:61] + b"\n" + s[61:] + b"\n")
            else:
                f.write(s + b"\n")
            f.write(b"+\n" + b"I" * len(s) + b"\n")
    fq = str(tmp_path / "q.fastq.gz")
    ref = kit.run_reference(fq, threads=1)
    os.rename(ref, ref + ".ref")
    p = subprocess.run([CLI, fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-1500:]
    assert filecmp.cmp(ref, ref + ".ref", shallow=False)
    head = open(ref, "rb").read(200)
    assert head.startswith(b"@r0 (null)\n")


def test_cli_errors(kit, cp, tmp_path):
    """Same failure behaviour as the reference for a missing profile, -M and .db inputs."""
    kit.simulate(write_to=str(tmp_path), root="e", seed=84, genome_len=20000, cov=12., het=0.01, len_mean=5000)
    fa = str(tmp_path / "e.fasta")
    os.remove(str(tmp_path / "e.prof"))
    p = subprocess.run([CLI, fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "Cannot open" in p.stderr
    p = subprocess.run([CLI, "-Mmodel", fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and "-M" in p.stderr
    p = subprocess.run([CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 1 and p.stderr.startswith("Usage: ClassPro")


def test_cli_raw_byte_sequences(kit, cp, tmp_path):
    """Lower-case bases: the reference compares raw characters (src/context.c), so 'a' != 'A' changes the
    run lengths; the CLI must ship such batches as bytes (seq_bits = 8) and still match."""
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    kit.simulate(write_to=str(tmp_path), root="lc", seed=85, genome_len=60000, cov=25., het=0.01, repeat_frac=0.5,
                 len_mean=8000)
    fa = str(tmp_path / "lc.fasta")
    rng = np.random.default_rng(3)
    lines = open(fa, "rb").read().split(b"\n")
    for i in range(1, len(lines), 2):
        s = bytearray(lines[i])
        for p in rng.integers(0, max(1, len(s)), size=max(1, len(s) // 200)):
            s[p:p + 3] = bytes(s[p:p + 3]).lower()
        lines[i] = bytes(s)
    open(fa, "wb").write(b"\n".join(lines))
    ref = kit.run_reference(fa, threads=1)
    os.rename(ref, ref + ".ref")
    p = subprocess.run([CLI, fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr[-1500:]
    assert filecmp.cmp(ref, ref + ".ref", shallow=False)
