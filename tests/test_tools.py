"""The evaluation tools next to the path: class2acc (host only) and prof2class
(profile decode + class map on the GPU), each against the unmodified reference tool built into
oracle/_ref/ by oracle/Makefile."""
import gzip
import os
import random
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "classpro_b200")
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")


def run(cmd, cwd=None):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return p.returncode, p.stdout, p.stderr


@pytest.fixture(scope="module")
def class_pair(tmp_path_factory):
    """Truth = a golden .class of the reference; estimate = the same with 5 % of the class characters
    redrawn (and a few reads with many E's, to exercise -f)."""
    d = tmp_path_factory.mktemp("c2a")
    truth = d / "truth.class"
    with gzip.open(os.path.join(GOLD, "g1.class.gz"), "rt") as f:
        lines = f.read().split("\n")
    truth.write_text("\n".join(lines))
    rnd = random.Random(5)
    out = []
    for i, l in enumerate(lines):
        if i % 4 == 3:
            rate = 0.6 if (i // 4) % 7 == 3 else 0.05
            l = "".join((rnd.choice("EHDR") if (c != "N" and rnd.random() < rate) else c) for c in l)
        out.append(l)
    est = d / "est.class"
    est.write_text("\n".join(out))
    return str(est), str(truth)


@pytest.mark.parametrize("opts", [[], ["-e3", "-r5", "-f50"], ["-s", "-e4"], ["-e0", "-m10", "-n60"], ["-r30"]])
def test_class2acc_matches_reference(kit, class_pair, opts):
    ref = os.path.join(REF, "class2acc")
    if not os.path.exists(ref):
        pytest.skip("reference class2acc not built (no /root/reference here)")
    kit.build_product()
    est, truth = class_pair
    a = run([ref] + opts + [est, truth])
    b = run([os.path.join(BIN, "class2acc")] + opts + [est, truth])
    assert a[0] == 0 and b[0] == 0, (a[2], b[2])
    assert a[1] == b[1]
    assert b"Confusion Matrix" in b[1]


def test_class2acc_errors(kit, class_pair, tmp_path):
    kit.build_product()
    est, truth = class_pair
    exe = os.path.join(BIN, "class2acc")
    assert run([exe, est])[0] == 1                                    # usage
    short = tmp_path / "short.class"
    short.write_text("\n".join(open(truth).read().split("\n")[:8]) + "\n")
    rc, _, err = run([exe, est, str(short)])
    assert rc == 1 and b"# seqs in" in err
    rc, _, err = run([exe, "-pno_such_root", est, truth])
    assert rc == 1 and b"Cannot open" in err


@pytest.mark.parametrize("opts", [["-e0"], ["-w500"], ["-w1000", "-e2", "-r5"], ["-w77", "-f40"]])
def test_class2acc_with_profiles_matches_reference(kit, tmp_path, opts):
    """-p<FastK root> (coverages of a read from the counts of its true H / D-mers) and -w<int> (one line per
    window), src/class2acc.c:98-104,174-185,228-247,272-279, on a dataset with several profile parts: the estimate
    is the reference's own .class, the truth the same with some class characters redrawn."""
    ref = os.path.join(REF, "class2acc")
    if not os.path.exists(ref) or not kit.have_reference():
        pytest.skip("reference tools not built (no /root/reference here)")
    kit.build_product()
    kit.simulate(write_to=str(tmp_path), root="w", seed=31, genome_len=150000, cov=25., het=0.01, repeat_frac=0.2,
                 len_mean=9000, len_sd=1500, nparts=3)
    fasta = str(tmp_path / "w.fasta")
    est = kit.run_reference(fasta, threads=2)
    lines = open(est).read().split("\n")
    rnd = random.Random(7)
    out = []
    for i, l in enumerate(lines):
        if i % 4 == 3:
            l = "".join((rnd.choice("EHDR") if (c != "N" and rnd.random() < 0.04) else c) for c in l)
        out.append(l)
    truth = tmp_path / "truth.class"
    truth.write_text("\n".join(out))
    root = str(tmp_path / "w")
    a = run([ref, "-p" + root] + opts + [est, str(truth)])
    b = run([os.path.join(BIN, "class2acc"), "-p" + root] + opts + [est, str(truth)])
    assert a[0] == 0 and b[0] == 0, (a[2][-500:], b[2][-500:])
    assert a[1] == b[1]
    assert b"H1-cov=" in b[1] or opts == ["-f40"]


@pytest.mark.gpu
def test_prof2class_matches_reference(kit, tmp_path):
    """Any FastK profile is a valid relative profile; the simulator's read profiles (clipped into the
    0..5 range so that all four classes occur) go through both tools."""
    ref = os.path.join(REF, "prof2class")
    assert os.path.exists(ref), "oracle/_ref/prof2class missing (built by oracle/Makefile where /root/reference exists)"
    import numpy as np
    d1, d2 = tmp_path / "a", tmp_path / "b"
    for d in (d1, d2):
        d.mkdir()
    sim = kit.simulate(write_to=str(d1), root="reads", seed=91, genome_len=60000, cov=3., het=0.01, repeat_frac=0.3,
                       len_mean=6000, len_sd=1500, len_min=30, short_reads=1, nparts=3)
    assert sim.nreads > 10
    for f in os.listdir(d1):
        shutil.copy(os.path.join(d1, f), os.path.join(d2, f))
    a = run([ref, "reads", "reads.fasta"], cwd=str(d1))
    b = run([os.path.join(BIN, "prof2class"), "reads", "reads.fasta"], cwd=str(d2))
    assert a[0] == 0, a[2]
    assert b[0] == 0, b[2]
    ca, cb = open(d1 / "reads.class", "rb").read(), open(d2 / "reads.class", "rb").read()
    assert ca == cb
    cls = b"".join(ca.split(b"\n")[3::4])
    assert all(ch in cls for ch in b"HDRN")         # read profiles have no zero counts: no E


@pytest.mark.gpu
def test_prof2class_abi_on_arbitrary_streams(kit):
    """cpg_prof2class on token streams with zero counts, wrap-around and 15-bit tokens: the class
    string is the count -> class map of src/prof2class.c:236-258 applied to the oracle's decode."""
    import numpy as np
    import classpro_b200 as cp
    from test_device_logic_hostsim import random_stream
    rng = np.random.default_rng(3)
    K = 40
    streams, lens = [], []
    for it in range(200):
        s = random_stream(rng, int(rng.integers(1, 500)), bool(it & 1))
        n, o = kit.oracle_decode(np.frombuffer(s, dtype=np.uint8), 200000)
        if n > 50000:
            continue
        streams.append((s, o[:n]))
        lens.append(n + K - 1)
    streams.append((b"", np.zeros(0, np.uint16)))                 # a read shorter than K: no profile, all 'N'
    lens.append(17)
    prof = np.frombuffer(b"".join(s for s, _ in streams), dtype=np.uint8)
    poff = np.zeros(len(streams) + 1, np.int64)
    np.cumsum([len(s) for s, _ in streams], out=poff[1:])
    model = cp.Model.from_cov(K, 10, 20, 20000)
    ctx = cp.Context(model)
    cls, coff, status = ctx.prof2class(prof, poff, np.array(lens, np.int32))
    ctx.close()
    assert not status.any()
    lut = np.full(65536, ord("R"), np.uint8)
    lut[0], lut[1], lut[2] = ord("E"), ord("H"), ord("D")
    for r, (s, o) in enumerate(streams):
        got = cls[coff[r]:coff[r + 1]]
        want = np.concatenate([np.full(min(K - 1, lens[r]), ord("N"), np.uint8), lut[o]])
        assert np.array_equal(got, want), r
