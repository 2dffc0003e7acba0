"""The device-side per-read logic (classpro_b200/csrc/*.cuh) compiled for the host with a warp
width of 1 (tests/hostsim), against the oracle.  This is how the CUDA sources are unit-tested on a
machine without a GPU; it is a test build only, the product has no CPU path."""
import ctypes as C

import numpy as np
import pytest


def compare_dataset(kit, sim, cov_opt=0, read_len=20000, seq_bits=2, intervals=True):
    om = kit.oracle_model(sim, cov_opt, read_len)
    gm = kit.gpu_model_from_sim(kit.hostsim_lib(), sim, cov_opt, read_len)
    ow = kit.OracleWork(clean=True)
    status = 0
    for i in range(sim.nreads):
        if sim.rlen[i] < sim.kmer:
            continue
        s, c = sim.read_ascii(i).tobytes(), sim.read_counts(i)
        if intervals:
            a, ia, ma = ow.classify(om, s, c, True)
            st, b, ib, mb = kit.hostsim_classify(gm, s, c, seq_bits, True)
            assert ma == mb and ia == ib, "read %d: interval tables differ" % i
        else:
            a = ow.classify(om, s, c)
            st, b = kit.hostsim_classify(gm, s, c, seq_bits)
        assert a == b, "read %d: class strings differ" % i
        status |= st
    return status


def test_model_matches_oracle(kit, hostsim):
    sim = kit.simulate(seed=41, genome_len=40000, cov=30., het=0.01)
    for cov_opt, rl in ((0, 20000), (33, 12000)):
        om = kit.oracle_model(sim, cov_opt, rl)
        gm = kit.gpu_model_from_sim(hostsim, sim, cov_opt, rl)
        assert list(om.cov) == list(gm.cov) and om.cmax == gm.cmax
        assert om.dr_ratio == gm.dr_ratio and om.hc_erate == gm.hc_erate
        assert np.array_equal(np.frombuffer(om.logfact, dtype=np.float64), np.frombuffer(gm.logfact, dtype=np.float64))
        oc = np.frombuffer(om.cthres, dtype=np.uint8).reshape(3, 21, 256, 2, 2)
        gc = np.frombuffer(gm.cthres, dtype=np.uint8).reshape(36, 256, 2, 2)
        row = 0
        for t, lmax in enumerate((20, 10, 6)):
            for l in range(1, lmax + 1):
                assert np.array_equal(oc[t, l], gc[row]), (t, l)
                row += 1
            assert np.array_equal(np.array(om.pe[t][:lmax + 1]), np.array(gm.pe[t][:lmax + 1]))


def test_plain_dataset(kit, hostsim):
    sim = kit.simulate(seed=42, genome_len=50000, cov=30., het=0.006)
    assert compare_dataset(kit, sim) == 0


def test_repeat_rich_dataset(kit, hostsim):
    sim = kit.simulate(seed=43, genome_len=60000, cov=40., het=0.01, repeat_frac=0.6, seg_dups=2)
    compare_dataset(kit, sim)


def test_options_and_raw_bytes(kit, hostsim):
    sim = kit.simulate(seed=44, genome_len=40000, cov=50., het=0.02, len_mean=9000, short_reads=1)
    compare_dataset(kit, sim, cov_opt=47, read_len=9000, seq_bits=8, intervals=False)


def test_noisy_low_complexity(kit, hostsim):
    sim = kit.simulate(seed=45, genome_len=40000, cov=35., het=0.001, err_indel_hp=0.003, err_sub=0.002,
                       repeat_frac=0.8)
    compare_dataset(kit, sim)


def test_minimal_reads(kit, hostsim):
    """Reads of length K, K+1 (profiles of 1 and 2 counts) and constant profiles."""
    sim = kit.simulate(seed=46, genome_len=30000, cov=20., het=0.01)
    om = kit.oracle_model(sim)
    gm = kit.gpu_model_from_sim(hostsim, sim)
    ow = kit.OracleWork(clean=True)
    s = sim.read_ascii(0).tobytes()
    c = sim.read_counts(0)
    for n in (1, 2, 3, 39, 40, 41, 80, 81):
        a = ow.classify(om, s[:n + 39], c[:n])
        st, b = kit.hostsim_classify(gm, s[:n + 39], c[:n], 2)
        assert a == b, n
    flat = np.full(500, 30, dtype=np.uint16)
    a = ow.classify(om, s[:539], flat)
    st, b = kit.hostsim_classify(gm, s[:539], flat, 2)
    assert a == b
    high = np.full(500, 3000, dtype=np.uint16)
    a = ow.classify(om, s[:539], high)
    st, b = kit.hostsim_classify(gm, s[:539], high, 2)
    assert a == b and set(a[39:]) == {ord("R")}


def test_compact_scratch_retry(kit, hostsim):
    """A read that outgrows the compact interval tables of the main launch is flagged and classified
    again with full-size tables: same result.  Tables of 6 entries force the second attempt for
    nearly every read; the 32-thread emulation runs the abort paths with lane groups of 16."""
    sim = kit.simulate(seed=21, genome_len=20000, cov=16., het=0.01, repeat_frac=0.4, len_mean=2500, len_sd=500,
                       len_min=600)
    om = kit.oracle_model(sim)
    ow = kit.OracleWork(clean=True)
    for lib, nreads, group in ((hostsim, min(sim.nreads, 60), 1), (kit.hostsim32_lib(), 5, 16)):
        gm = kit.gpu_model_from_sim(lib, sim)
        if group > 1:
            lib.hs_set_group(group)
        lib.hs_set_small_caps(6)
        try:
            for i in range(nreads):
                s, c = sim.read_ascii(i).tobytes(), sim.read_counts(i)
                a, ia, ma = ow.classify(om, s, c, True)
                st, b, ib, mb = kit.hostsim_classify(gm, s, c, 2, True, lib=lib)
                assert st == 0 and a == b and ia == ib and ma == mb, i
            assert lib.hs_retries() >= nreads // 2
        finally:
            lib.hs_set_small_caps(0)
            if group > 1:
                lib.hs_set_group(32)


def test_context_closed_form_exhaustive(kit, hostsim):
    """cpg_context.cuh's on-demand run lengths equal the reference sweep (src/context.c) at every
    base, for every sequence of length <= 7 over ACGT and <= 12 over a two-letter alphabet."""
    L = kit.oracle_lib()

    def check(seq):
        n = len(seq)
        lc = np.zeros((n + 4, 3), dtype=np.uint8)
        rc = np.full((n + 4, 3), 0xEE, dtype=np.uint8)
        lc[0] = (1, 0, 0)
        L.cpo_seq_context(lc.ctypes.data, rc.ctypes.data, seq, n)
        for p in range(n):
            for t in range(3):
                assert hostsim.hs_ctx(seq, n, p, 0, t) == lc[p, t], (seq, p, t, "left")
                assert hostsim.hs_ctx(seq, n, p, 1, t) == rc[p, t], (seq, p, t, "right")
                assert hostsim.hs_ctx2(seq, n, p, 0, t, p & 3) == lc[p, t], (seq, p, t, "left, packed")
                assert hostsim.hs_ctx2(seq, n, p, 1, t, n & 3) == rc[p, t], (seq, p, t, "right, packed")

    import itertools
    for n in range(2, 7):
        for tup in itertools.product(b"ACGT", repeat=n):
            check(bytes(tup))
    for n in range(7, 12):
        for tup in itertools.product(b"AC", repeat=n):
            check(bytes(tup))


def test_context_random_low_complexity(kit, hostsim):
    L = kit.oracle_lib()
    rng = np.random.default_rng(7)
    for _ in range(60):
        parts = []
        while sum(map(len, parts)) < 600:
            kind = rng.integers(0, 4)
            if kind == 0:
                parts.append(bytes(rng.choice(list(b"ACGT"), size=rng.integers(1, 30)).tolist()))
            else:
                u = bytes(rng.choice(list(b"ACGT"), size=kind).tolist())
                parts.append((u * 40)[:int(rng.integers(kind, kind * 30))])
        seq = b"".join(parts)
        n = len(seq)
        lc = np.zeros((n + 4, 3), dtype=np.uint8)
        rc = np.full((n + 4, 3), 0xEE, dtype=np.uint8)
        lc[0] = (1, 0, 0)
        L.cpo_seq_context(lc.ctypes.data, rc.ctypes.data, seq, n)
        if lc[:n].max() >= 127:
            continue   # beyond the cap the reference reads cells it never wrote
        for p in range(0, n, 3):
            for t in range(3):
                assert hostsim.hs_ctx(seq, n, p, 0, t) == lc[p, t]
                assert hostsim.hs_ctx(seq, n, p, 1, t) == rc[p, t]
                assert hostsim.hs_ctx2(seq, n, p, 0, t, p & 3) == lc[p, t]
                assert hostsim.hs_ctx2(seq, n, p, 1, t, (p >> 2) & 3) == rc[p, t]


def test_context_packed_long_runs(kit, hostsim):
    """Packed (32 bases per window) and raw (base by base) evaluation agree on runs longer than a
    window and longer than the 127 cap, at every position and buffer alignment."""
    rng = np.random.default_rng(11)
    for unit in (b"A", b"AC", b"ACG", b"T", b"GT", b"TTC"):
        for copies in (20, 45, 130, 200):
            seq = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 9))).tolist()) + unit * copies \
                + bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 9))).tolist())
            n = len(seq)
            step = 1 if n < 150 else 5
            for p in list(range(0, n, step)) + [n - 1]:
                for t in range(3):
                    for right in (0, 1):
                        assert hostsim.hs_ctx2(seq, n, p, right, t, p & 3) == hostsim.hs_ctx(seq, n, p, right, t), \
                            (unit, copies, p, t, right)


def test_context_long_runs_match_the_reference_sweep(kit, hostsim):
    """Runs LONGER than the 127 cap against the reference sweep itself (the oracle's restatement of
    src/context.c on dense arrays), not only packed against raw.  Left contexts and the di-/trinucleotide right
    contexts are the capped closed form at any length.  The homopolymer right context is back-filled by the
    reference over the last min(L,127) bases of a run with mirrored CAPPED left lengths (src/context.c:24-26): in a
    run of 253+ bases the last base gets 127, not 1 -- reproduced; the bases before that stretch are never written
    there (cells keep the 0xEE the test put in): the device defines them as 127 (and flags the read when it uses
    one).  This is what made 49 of 100 k reads of the repeat-rich 50 Mb dataset differ from the reference binary
    (profiles/r02_file_parity_c4_50mb_first.json)."""
    L = kit.oracle_lib()
    rng = np.random.default_rng(5)
    for unit in (b"A", b"AC", b"ACG", b"T", b"GT", b"TTC"):
        for copies in (100, 127, 128, 129, 200, 253, 254, 300, 400):
            pre = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 9))).tolist())
            post = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 9))).tolist())
            seq = pre + unit * copies + post
            n = len(seq)
            lc = np.zeros((n + 4, 3), dtype=np.uint8)
            rc = np.full((n + 4, 3), 0xEE, dtype=np.uint8)
            lc[0] = (1, 0, 0)
            L.cpo_seq_context(lc.ctypes.data, rc.ctypes.data, seq, n)
            step = 1 if copies in (128, 254) else 3
            for p in list(range(0, n, step)) + [n - 1]:
                for t in range(3):
                    want_r = 127 if rc[p, t] == 0xEE else rc[p, t]
                    assert hostsim.hs_ctx(seq, n, p, 0, t) == lc[p, t], (unit, copies, p, t, "left")
                    assert hostsim.hs_ctx2(seq, n, p, 0, t, p & 3) == lc[p, t], (unit, copies, p, t, "left, packed")
                    assert hostsim.hs_ctx(seq, n, p, 1, t) == want_r, (unit, copies, p, t, "right")
                    assert hostsim.hs_ctx2(seq, n, p, 1, t, n & 3) == want_r, (unit, copies, p, t, "right, packed")


def test_math_shortcuts_are_bit_identical(kit, hostsim):
    """cpg_math.cuh answers two loops without running them; both must leave every bit / every decision as the
    reference's loops do (oracle: src/prob.c:76-112 restated, src/bessel.c:482-521 restated).
    (1) a binomial tail whose first term underflows to 0: the reference adds up to 32 767 further terms, all 0;
    (2) cpg_lp_trans_thr: a Skellam log-probability that is only compared with a threshold is -inf when the bound
        n(log(lambda/n)+1) < threshold-2 decides the comparison -- then the exact value must be below the threshold;
        otherwise it is the exact value, bit for bit."""
    import ctypes as C
    import math
    L = kit.oracle_lib()
    sim = kit.simulate(seed=5, genome_len=60000, cov=25., het=0.01, len_mean=8000)
    om = kit.oracle_model(sim)
    gm = kit.gpu_model_from_sim(hostsim, sim)
    L.cpo_binom_test_g.restype = C.c_double
    L.cpo_binom_test_g.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
    L.cpo_bessi.restype = C.c_double
    L.cpo_bessi.argtypes = [C.c_int, C.c_double]
    hostsim.hs_binom_tail.restype = C.c_double
    hostsim.hs_binom_tail.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int)]
    hostsim.hs_lp_trans_thr.restype = C.c_double
    hostsim.hs_lp_trans_thr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    rng = np.random.default_rng(3)
    bad = C.c_int()
    n_under = 0
    cases = [(0, 32767, 0.004), (32767, 32767, 0.004), (20000, 32767, 0.004), (5, 32767, 0.802), (100, 30000, 0.5),
             (1, 2, 0.004), (0, 0, 0.1), (3000, 3000, 0.05), (2, 20000, 0.3)]
    for _ in range(4000):
        n = int(rng.choice([rng.integers(1, 80), rng.integers(80, 3000), rng.integers(3000, 32768)]))
        k = int(rng.integers(0, n + 1))
        p = float(rng.choice([0.004, 0.01, 0.1, 0.2, 0.802, rng.uniform(0.002, 0.9)]))
        cases.append((k, n, p))
    for k, n, p in cases:
        a = L.cpo_binom_test_g(C.byref(om), k, n, p, 0)
        b = hostsim.hs_binom_tail(C.byref(gm), k, n, p, C.byref(bad))
        assert np.float64(a).tobytes() == np.float64(b).tobytes(), (k, n, p, a, b)
        n_under += (a == 0.0 or a == 1.0)
    assert n_under > 100                     # the underflow branch is really exercised

    def exact(read_len, b, e, cb, ce, cov):
        lam = float(cov) * abs(e - b) / read_len
        with np.errstate(all="ignore"):
            return float(np.float64(-2. * lam) + np.log(np.float64(L.cpo_bessi(abs(ce - cb), 2. * lam))))

    def same_bits(x, y):
        return np.float64(x).tobytes() == np.float64(y).tobytes() or (math.isnan(x) and math.isnan(y))

    n_skip = n_odd = 0
    for thres in (-23.025851, -9.210340):
        for _ in range(3000):
            cov = int(rng.choice([rng.integers(1, 60), rng.integers(60, 5000), rng.integers(5000, 32768)]))
            d = int(rng.integers(1, 450))
            cb = int(rng.integers(0, cov + 1))
            ce = int(rng.choice([cb + rng.integers(-3, 4), rng.integers(0, 32768)]))
            ce = max(0, min(32767, ce))
            b0 = int(rng.integers(0, 20000))
            got = hostsim.hs_lp_trans_thr(20000, b0, b0 + d, cb, ce, cov, thres, 1)
            plain = hostsim.hs_lp_trans_thr(20000, b0, b0 + d, cb, ce, cov, thres, 0)
            ex = exact(20000, b0, b0 + d, cb, ce, cov)
            assert same_bits(plain, ex), (cov, d, cb, ce, plain, ex)
            if got == -math.inf and not same_bits(plain, got):
                n_skip += 1
                assert plain < thres - 1.0, (cov, d, cb, ce, plain)          # decided, with room to spare
            else:
                assert same_bits(got, plain)
            # both forms the reference uses (src/wall.c:366 `>=`, :390 and :1028 `<`): NaN and +inf included
            assert (got >= thres) == (plain >= thres) and (got < thres) == (plain < thres)
            n_odd += (math.isnan(plain) or plain == math.inf)
    assert n_skip > 20 and n_odd > 20            # exp overflow inside bessi0 (2 lambda > 709): kept as the reference has it


def random_stream(rng, n_tokens, adversarial):
    """A FastK token stream; adversarial = arbitrary bytes (wrap-around and mask corner cases)."""
    first = int(rng.integers(0, 32768))
    out = bytearray()
    if first >= 128 or rng.random() < 0.2:
        out += bytes([0x80 | (first >> 8), first & 0xff])
    else:
        out.append(first)
    for _ in range(n_tokens):
        k = rng.random()
        if adversarial:
            b = int(rng.integers(0, 256))
            out.append(b)
            if b & 0x80:
                out.append(int(rng.integers(0, 256)))
        elif k < 0.5:
            out.append(int(rng.integers(0, 64)))
        elif k < 0.9:
            out.append(0x40 | int(rng.integers(0, 64)))
        else:
            out += bytes([0x80 | int(rng.integers(0, 128)), int(rng.integers(0, 256))])
    return bytes(out)


@pytest.mark.parametrize("adversarial", [False, True])
def test_decode_streams(kit, hostsim, adversarial):
    """The scan-based decoder equals the byte-serial one on well-formed and on arbitrary streams
    (16-bit wrap of short deltas, 15-bit mask of long ones, zero-length runs, cap truncation)."""
    rng = np.random.default_rng(11 + adversarial)
    for it in range(300):
        s = random_stream(rng, int(rng.integers(0, 400)), adversarial)
        n1, o1 = kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 100000)
        for cap in (100000, max(1, n1 // 2)):
            n2, o2 = kit.hostsim_decode(np.frombuffer(s, dtype=np.uint8), cap)
            assert n2 == n1
            assert np.array_equal(o2, o1[:min(cap, n1)])


def test_decode_empty_profile(kit, hostsim):
    n, o = kit.hostsim_decode(np.zeros(0, dtype=np.uint8), 10)
    assert n == 0


@pytest.mark.parametrize("group", [32, 16, 8, 4])
def test_warp32_emulation_classify(kit, group):
    """The same device sources with 32 host threads playing the lanes of one warp: every ballot,
    shuffle, reduction and group barrier is a rendezvous, lanes run asynchronously in between.
    Catches collectives reached by only some lanes (hang) and missing synchronisation (mismatch).
    `group` = lanes per read: the warp classifies 32/group copies of the read at the same time, each
    lane group with its own scratch (the kernels run with groups of 8 and 4), and the groups must agree."""
    L32 = kit.hostsim32_lib()
    assert L32.hs_set_group(group) == 0
    sim = kit.simulate(seed=47, genome_len=20000, cov=14., het=0.01, repeat_frac=0.4, len_mean=2200, len_sd=400,
                       len_min=500)
    om = kit.oracle_model(sim)
    gm = kit.gpu_model_from_sim(L32, sim)
    ow = kit.OracleWork(clean=True)
    try:
        for i in range(min(sim.nreads, 10 if group == 16 else 4)):
            s, c = sim.read_ascii(i).tobytes(), sim.read_counts(i)
            a, ia, ma = ow.classify(om, s, c, True)
            st, b, ib, mb = kit.hostsim_classify(gm, s, c, 2, True, lib=L32)
            assert st == 0 and a == b and ia == ib and ma == mb, i
    finally:
        L32.hs_set_group(32)


def test_decoder_candidate_bitmap(kit, hostsim):
    """The bit map the decoder writes on the side equals the candidate definition applied to the
    decoded counts, on simulated profiles and on arbitrary byte streams (wrap-around included)."""
    sim = kit.simulate(seed=5, genome_len=30000, cov=25., het=0.01, repeat_frac=0.3, len_mean=3000, len_sd=600,
                       len_min=500)
    for i in range(min(sim.nreads, 40)):
        c = sim.read_counts(i)
        for rcov in (12, 60, 40000):
            n, o, bits = kit.hostsim_decode_cand(sim.read_prof(i), len(c), rcov)
            assert n == len(c) and np.array_equal(o, c)
            assert np.array_equal(bits, kit.candidate_bits(c, rcov)), (i, rcov)
    rng = np.random.default_rng(77)
    L32 = kit.hostsim32_lib()
    for it in range(60):
        s = random_stream(rng, int(rng.integers(1, 400)), bool(it & 1))
        n1, o1 = kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 100000)
        for lib in (None, L32) if it < 12 else (None,):
            n2, o2, bits = kit.hostsim_decode_cand(np.frombuffer(s, dtype=np.uint8), 100000, 200, lib=lib)
            assert n2 == n1 and np.array_equal(o2, o1)
            assert np.array_equal(bits, kit.candidate_bits(o1, 200)), it
    # a capacity smaller than the stream: bits past the capacity are never written
    s = random_stream(rng, 300, False)
    n1, o1 = kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 100000)
    cap = max(1, n1 // 2)
    n2, o2, bits = kit.hostsim_decode_cand(np.frombuffer(s, dtype=np.uint8), cap, 200)
    assert n2 == n1 and np.array_equal(o2, o1[:cap]) and np.array_equal(bits, kit.candidate_bits(o1[:cap], 200))


def test_warp32_emulation_decode(kit):
    L32 = kit.hostsim32_lib()
    rng = np.random.default_rng(13)
    for it in range(40):
        s = random_stream(rng, int(rng.integers(0, 300)), bool(it & 1))
        n1, o1 = kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 100000)
        n2, o2 = kit.hostsim_decode(np.frombuffer(s, dtype=np.uint8), 100000, lib=L32)
        assert n2 == n1 and np.array_equal(o2, o1)


def test_reads_that_differed_from_the_reference_binary(kit, hostsim):
    """Six of the 49 reads of the repeat-rich 50 Mb file-level run of round 2 whose class line differed from the
    unmodified reference binary's (tests/golden/long_hp_runs.npz: sequence, exact counts, the REFERENCE's class
    line, the histogram of the run; written by tools/trace_file_flips.py on the GPU box).  All of them have a
    homopolymer run longer than 127 bases next to a wall.  The oracle and the host build of the device code must
    both give the reference's line."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long_hp_runs.npz"))

    class S:
        pass
    sim = S()
    sim.kmer = 40
    sim.hist = d["hist"]
    om = kit.oracle_model(sim, 0, 20000)
    gm = kit.gpu_model_from_sim(hostsim, sim, 0, 20000)
    ow = kit.OracleWork(clean=True)
    reads = sorted(int(k[4:]) for k in d.files if k.startswith("seq_"))
    assert len(reads) == 6
    for r in reads:
        seq, cnt, ref = d["seq_%d" % r].tobytes(), d["cnt_%d" % r], d["ref_%d" % r].tobytes()
        assert ow.classify(om, seq, cnt) == ref, r
        st, cls = kit.hostsim_classify(gm, seq, cnt, lib=hostsim)
        assert cls == ref, r
        assert not (st & ~32), (r, st)
