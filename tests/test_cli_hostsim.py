"""The host side of the product -- the ClassPro and prof2class programs: option parsing, FASTX
reader, batching, the packing pool, per-GPU workers, the ordered writer, the byte contract of the
records -- on machines without a GPU.  tests/hostsim/build.sh links the programs' own sources
against a TEST-ONLY stand-in for the device (tests/hostsim/fakedev.cpp: the device sources compiled
for the host, one read at a time); nothing of this is in the product library."""
import filecmp
import re
import gzip
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "hostsim", "_build")
CLI = os.path.join(BUILD, "ClassPro")
P2C = os.path.join(BUILD, "prof2class")


@pytest.fixture(scope="module")
def progs(kit):
    kit.build_hostsim()
    assert os.path.exists(CLI) and os.path.exists(P2C)
    return CLI


def run(cmd, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, env=e, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.mark.parametrize("name", ["g1", "g2"])
@pytest.mark.parametrize("opts,devices", [([], 1), (["-B1", "-T1"], 1), (["-B1", "-T5"], 3), (["-B2", "-G2", "-T2"], 4)])
def test_golden_fixture_through_the_programs_host_side(kit, progs, tmp_path, name, opts, devices):
    """Golden .class of the reference, reproduced through reader -> pool -> workers -> writer with
    batches of 1-2 megabases, 1-5 packing threads and 1-3 worker threads ("GPUs")."""
    from test_oracle import unpack_golden
    fasta, golden, args = unpack_golden(name, str(tmp_path))
    p = run([CLI, "-v"] + args + opts + [fasta], env={"CPG_FAKE_DEVICES": str(devices)})
    assert p.returncode == 0, p.stderr[-1500:]
    assert filecmp.cmp(fasta[:-6] + ".class", golden, shallow=False)
    assert "Classified" in p.stderr


def test_fastq_gz_header_quirks_and_short_reads(kit, progs, tmp_path):
    """FASTQ.gz input, wrapped sequence lines, reads shorter than K, and the header quirks that are
    part of the byte contract: no comment on the first record -> "(null)", a
    tab-separated comment, a comment-less record after one with a comment -> the stale comment."""
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
            if i % 3 == 0 and len(s) > 100:
                f.write(s[:61] + b"\n" + s[61:] + b"\n")
            else:
                f.write(s + b"\n")
            f.write(b"+\n" + b"I" * len(s) + b"\n")
    fq = str(tmp_path / "q.fastq.gz")
    ref = kit.run_reference(fq, threads=1)
    os.rename(ref, ref + ".ref")
    for opts in ([], ["-B1", "-T3"]):
        p = run([CLI] + opts + [fq], env={"CPG_FAKE_DEVICES": "2"})
        assert p.returncode == 0, p.stderr[-1500:]
        assert filecmp.cmp(ref, ref + ".ref", shallow=False), opts
        os.remove(ref)


def test_program_errors(kit, progs, tmp_path):
    from test_oracle import unpack_golden
    fasta, golden, args = unpack_golden("g1", str(tmp_path))
    p = run([CLI])
    assert p.returncode == 1 and "Usage" in p.stderr
    p = run([CLI, str(tmp_path / "nothing.fasta")])
    assert p.returncode == 1 and "Cannot open" in p.stderr
    p = run([CLI, "-Mmodel", fasta])
    assert p.returncode == 1 and "-M" in p.stderr
    p = run([CLI, "-Tx", fasta])
    assert p.returncode == 1 and "not an integer" in p.stderr
    p = run([CLI, "-q", fasta])
    assert p.returncode == 1 and "illegal option" in p.stderr
    p = run([CLI, fasta], env={"CPG_FAKE_DEVICES": "0"})
    assert p.returncode == 1 and "no CUDA device" in p.stderr
    os.remove(str(tmp_path / ".g1.prof.1"))
    p = run([CLI, fasta])
    assert p.returncode == 1 and "misssing" in p.stderr         # the reference's own spelling (src/libfastk.c:1302)


def test_prof2class_host_side(kit, progs, tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "prof2class")
    if not os.path.exists(ref):
        pytest.skip("reference prof2class not built (no /root/reference here)")
    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir()
    d2.mkdir()
    kit.simulate(write_to=str(d1), root="reads", seed=91, genome_len=60000, cov=3., het=0.01, repeat_frac=0.3,
                 len_mean=6000, len_sd=1500, len_min=30, short_reads=1, nparts=3)
    for f in os.listdir(d1):
        shutil.copy(os.path.join(d1, f), os.path.join(d2, f))
    a = run([ref, "reads", "reads.fasta"], cwd=str(d1))
    b = run([P2C, "reads.prof", "reads.fasta"], cwd=str(d2))
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    assert filecmp.cmp(str(d1 / "reads.class"), str(d2 / "reads.class"), shallow=False)


def test_raw_byte_sequences_and_mixed_batches(kit, progs, tmp_path):
    """Lower-case bases: the reference compares raw characters (src/context.c), so 'a' != 'A' changes the
    run lengths; a batch with such a read is shipped as bytes (seq_bits = 8), the others 2-bit packed --
    with 1 Mbase batches both kinds occur in one run."""
    import numpy as np
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    kit.simulate(write_to=str(tmp_path), root="lc", seed=85, genome_len=60000, cov=25., het=0.01, repeat_frac=0.5,
                 len_mean=8000)
    fa = str(tmp_path / "lc.fasta")
    rng = np.random.default_rng(3)
    lines = open(fa, "rb").read().split(b"\n")
    for i in range(1, len(lines) // 2, 2):                  # first half of the reads only
        s = bytearray(lines[i])
        for p in rng.integers(0, max(1, len(s)), size=max(1, len(s) // 200)):
            s[p:p + 3] = bytes(s[p:p + 3]).lower()
        lines[i] = bytes(s)
    open(fa, "wb").write(b"\n".join(lines))
    ref = kit.run_reference(fa, threads=1)
    os.rename(ref, ref + ".ref")
    for opts in (["-B1", "-T2"], []):
        p = run([CLI] + opts + [fa])
        assert p.returncode == 0, p.stderr[-1500:]
        assert filecmp.cmp(ref, ref + ".ref", shallow=False), opts
        os.remove(ref)


def test_empty_and_tiny_inputs(kit, progs, tmp_path):
    """No reads at all, and a single read shorter than K (src/ClassPro.c:209-226)."""
    from test_oracle import unpack_golden
    fasta, golden, args = unpack_golden("g1", str(tmp_path))
    recs = open(fasta, "rb").read().split(b">")[1:]
    # the profile index still lists every read; only the FASTA is cut short, as when a run is interrupted
    open(fasta, "wb").write(b">" + recs[0])
    p = run([CLI] + args + [fasta])
    assert p.returncode == 0, p.stderr[-800:]
    out = open(fasta[:-6] + ".class", "rb").read().split(b"\n")
    want = open(golden, "rb").read().split(b"\n")
    assert out[:4] == want[:4] and len(out) == 5
    open(fasta, "wb").write(b"")
    p = run([CLI] + args + [fasta])
    assert p.returncode == 0 and os.path.getsize(fasta[:-6] + ".class") == 0


@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_host_side_under_sanitizers(kit, progs, tmp_path, san):
    """The programs' host side (reader, packing pool, three worker threads, writer) built with
    -fsanitize=address / thread: no report, same bytes.  (Found the lazily initialised table of
    cpg_pack_seq, which several packing threads raced on.)"""
    from test_oracle import unpack_golden
    host, hs = os.path.join(ROOT, "classpro_b200", "host"), os.path.join(ROOT, "tests", "hostsim")
    cf = ["-O1", "-g", "-ffp-contract=off", "-fsanitize=" + san, "-I" + os.path.join(ROOT, "include"), "-I" + host]
    objs = []
    for cc, src in (("gcc", os.path.join(host, "cpg_model.c")), ("gcc", os.path.join(host, "cpg_pack.c")),
                    ("gcc", os.path.join(host, "classpro_main.c")), ("g++", os.path.join(hs, "fakedev.cpp"))):
        o = str(tmp_path / (os.path.basename(src) + ".o"))
        p = run([cc] + cf + ["-w", "-c", src, "-o", o])
        if p.returncode != 0 and "sanitize" in p.stderr:
            pytest.skip("no %s sanitizer in this toolchain" % san)
        assert p.returncode == 0, p.stderr[-800:]
        objs.append(o)
    exe = str(tmp_path / "ClassPro_san")
    p = run(["g++", "-fsanitize=" + san, "-o", exe] + objs + ["-lz", "-lpthread", "-lm"])
    if p.returncode != 0:
        pytest.skip("cannot link with -fsanitize=%s: %s" % (san, p.stderr[-200:]))
    fasta, golden, args = unpack_golden("g1", str(tmp_path / "d"))
    p = run([exe, "-B1", "-T4"] + args + [fasta], env={"CPG_FAKE_DEVICES": "3", "ASAN_OPTIONS": "detect_leaks=0"})
    assert p.returncode == 0, p.stderr[-3000:]
    assert "Sanitizer" not in p.stderr, p.stderr[-3000:]
    assert filecmp.cmp(fasta[:-6] + ".class", golden, shallow=False)


@pytest.mark.parametrize("group", [4, 8])
def test_warp_emulation_under_thread_sanitizer(kit, progs, tmp_path, group):
    """The device sources with 32 host threads playing the lanes of a warp (lane groups of 4 / 8, as
    in k_wall_b / k_rel / k_unrel_b), built with -fsanitize=thread: a scratch or shared word written by one
    lane and read by another without a group barrier in between is a data race it reports.  30 reads
    of the golden fixture through decode and all three phases: no report, same bytes."""
    from test_oracle import unpack_golden
    host, hs = os.path.join(ROOT, "classpro_b200", "host"), os.path.join(ROOT, "tests", "hostsim")
    cf = ["-O1", "-g", "-ffp-contract=off", "-fsanitize=thread", "-w", "-I" + os.path.join(ROOT, "include"), "-I" + host]
    objs = []
    for cc, src, extra in (("gcc", os.path.join(host, "cpg_model.c"), []), ("gcc", os.path.join(host, "cpg_pack.c"), []),
                           ("gcc", os.path.join(host, "classpro_main.c"), []),
                           ("g++", os.path.join(hs, "fakedev.cpp"), ["-DCPG_HOSTSIM=32"])):
        o = str(tmp_path / (os.path.basename(src) + ".o"))
        p = run([cc] + cf + extra + ["-c", src, "-o", o])
        if p.returncode != 0 and "sanitize" in p.stderr:
            pytest.skip("no thread sanitizer in this toolchain")
        assert p.returncode == 0, p.stderr[-800:]
        objs.append(o)
    exe = str(tmp_path / "ClassPro_t32")
    p = run(["g++", "-fsanitize=thread", "-o", exe] + objs + ["-lz", "-lpthread", "-lm"])
    if p.returncode != 0:
        pytest.skip("cannot link with -fsanitize=thread: %s" % p.stderr[-200:])
    fasta, golden, args = unpack_golden("g1", str(tmp_path / "d"))
    recs = open(fasta, "rb").read().split(b">")[1:31]
    open(fasta, "wb").write(b"".join(b">" + r for r in recs))
    p = run([exe, "-T2"] + args + [fasta], env={"CPG_FAKE_GROUP": str(group)})
    assert p.returncode == 0, p.stderr[-3000:]
    assert "Sanitizer" not in p.stderr, p.stderr[-3000:]
    want = b"\n".join(open(golden, "rb").read().split(b"\n")[:4 * len(recs)]) + b"\n"
    assert open(fasta[:-6] + ".class", "rb").read() == want


def test_fasta_format_variants(kit, progs, tmp_path):
    """Record formats kseq accepts (src/kseq.h:177-218) and the program must echo byte for byte like
    the reference: wrapped sequence lines of several widths, CRLF line ends, blank lines between
    records, no comment / tab separator / extra blanks in the header."""
    import numpy as np
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    rng = np.random.default_rng(1)
    d = str(tmp_path)
    sim = kit.simulate(write_to=d, root="q", seed=1, genome_len=30000, cov=12., het=0.01, len_mean=4000,
                       short_reads=1, nparts=2)
    fa = os.path.join(d, "q.fasta")
    for variant in range(6):
        out = bytearray()
        for i in range(sim.nreads):
            s = sim.read_ascii(i).tobytes()
            hdr = sim.headers[i]
            k = rng.integers(0, 8)
            if k == 0:
                hdr = hdr.split(b" ")[0]
            elif k == 1:
                hdr = hdr.replace(b" ", b"\t", 1)
            elif k == 2:
                hdr = hdr + b"  trailing  "
            elif k == 3:
                hdr = hdr.replace(b" ", b"  ", 1)
            nl = b"\r\n" if rng.random() < 0.15 else b"\n"
            out += b">" + hdr + nl
            w = int(rng.choice([0, 0, 60, 80, 7, 1000]))
            if w and len(s) > 0:
                for a in range(0, len(s), w):
                    out += s[a:a + w] + nl
            else:
                out += s + nl
            if rng.random() < 0.1:
                out += nl
        open(fa, "wb").write(bytes(out))
        ref = kit.run_reference(fa, threads=1)
        os.replace(ref, ref + ".ref")
        p = run([CLI, "-B1", "-T2", fa])
        assert p.returncode == 0, p.stderr[-1500:]
        assert filecmp.cmp(ref, ref + ".ref", shallow=False), variant


def test_parallel_parse_and_mapped_writer(kit, progs, tmp_path):
    """The chunk-parallel FASTX parser and the mapped, parallel writer against the
    reference binary: FASTQ whose quality lines begin with / contain '@', '>' and '+' (a guessed
    piece start that is not a record start must be caught and parsed again), wrapped records,
    comments that a later comment-less record repeats across piece boundaries, reads shorter than
    K, CRLF; windows of 30-200 kb cut into 1-9 pieces.  Same bytes as with the serial reader /
    write() writer (CPG_SERIAL_IO=1) and as the reference."""
    import numpy as np
    if not kit.have_reference():
        pytest.skip("oracle/_ref/ClassPro not present")
    rng = np.random.default_rng(7)
    d = str(tmp_path)
    sim = kit.simulate(write_to=d, root="q", seed=5, genome_len=40000, cov=15., het=0.01, len_mean=3000,
                       short_reads=1, nparts=3)
    os.remove(os.path.join(d, "q.fasta"))
    qchars = np.frombuffer(b"@>+I!~5@@>", dtype=np.uint8)
    stats = {}
    for variant, ext in enumerate(["fastq", "fastq", "fasta", "fastq"]):
        out = bytearray()
        for i in range(sim.nreads):
            s = sim.read_ascii(i).tobytes()
            name = b"r%d" % i
            k = rng.integers(0, 4) if variant != 3 else (1 if i == sim.nreads // 2 else 0)
            hdr = name if k == 0 else name + b" c%d x" % i if k == 1 else name + b"\tt%d" % i if k == 2 else name + b" "
            nl = b"\r\n" if (variant == 2 and rng.random() < 0.2) else b"\n"
            mark = b"@" if ext == "fastq" else b">"
            out += mark + hdr + nl
            w = int(rng.choice([0, 0, 0, 50, 211])) if variant != 0 else 0
            lines = [s[a:a + w] for a in range(0, len(s), w)] if (w and s) else [s]
            for ln in lines:
                out += ln + nl
            if ext == "fastq":
                q = qchars[rng.integers(0, len(qchars), len(s))].tobytes()
                if len(s) and rng.random() < 0.5:
                    q = b"@" + q[1:]
                out += b"+" + (name if rng.random() < 0.3 else b"") + b"\n"
                qlines = [q[a:a + w] for a in range(0, len(q), w)] if (w and q) else [q]
                for ln in qlines:
                    out += ln + b"\n"
        src = os.path.join(d, "q." + ext)
        open(src, "wb").write(bytes(out))
        ref = kit.run_reference(src, threads=1)
        os.replace(ref, ref + ".ref")
        for env in ({"CPG_SERIAL_IO": "1"},
                    {"CPG_BATCH_BASES": "30000", "CPG_PARSE_PIECES": "9"},
                    {"CPG_BATCH_BASES": "200000", "CPG_PARSE_PIECES": "4"},
                    {"CPG_BATCH_BASES": "70001", "CPG_PARSE_PIECES": "1"},
                    {}):
            e = dict(env)
            e["CPG_FAKE_DEVICES"] = "2"
            p = run([CLI, "-v", "-T3", src], env=e)
            assert p.returncode == 0, p.stderr[-1500:]
            assert ("Parsing" in p.stderr) == ("CPG_SERIAL_IO" not in env)
            m = re.search(r"parser: (\d+) pieces parsed again from their true start, (\d+) headers completed", p.stderr)
            stats[(variant, env.get("CPG_PARSE_PIECES", "serial" if "CPG_SERIAL_IO" in env else "default"))] = (int(m.group(1)), int(m.group(2)))
            assert filecmp.cmp(ref, ref + ".ref", shallow=False), (variant, env)
            os.remove(ref)
        os.remove(src)
    # the cases the test is about did occur: wrong guesses in wrapped FASTQ (none in FASTA), carried comments
    assert stats[(1, "9")][0] > 0 and stats[(2, "9")][0] == 0 and stats[(1, "serial")] == (0, 0)
    assert stats[(3, "9")][1] > 0 and stats[(2, "9")][1] > 0, stats


def test_parallel_parse_fuzz(kit, progs, tmp_path):
    """Differential fuzz of the chunk-parallel parser + mapped writer against the serial parser +
    write() of the same program (CPG_SERIAL_IO=1; that pair is checked against the reference binary
    above) on randomly formatted FASTA/FASTQ: windows smaller than a read, more pieces than records,
    no newline at the end of the file, blank lines, CRLF, '>' and '@' inside headers and quality
    strings, reads shorter than K.  Null device (class strings of 'X'): the test is about records,
    offsets and order."""
    import numpy as np
    rng = np.random.default_rng(11)
    d = str(tmp_path)
    sim = kit.simulate(write_to=d, root="f", seed=2, genome_len=20000, cov=8., het=0.01, len_mean=2500,
                       short_reads=1, nparts=2)
    os.remove(os.path.join(d, "f.fasta"))
    qchars = np.frombuffer(b"@>+#I5", dtype=np.uint8)
    for trial in range(24):
        fastq = trial % 2 == 1
        out = bytearray()
        for i in range(sim.nreads):
            s = sim.read_ascii(i).tobytes()
            nl = b"\r\n" if rng.random() < 0.1 else b"\n"
            hdr = b"r%d" % i + [b"", b" a>b @c", b"\t@x", b" "][int(rng.integers(0, 4))]
            out += (b"@" if fastq else b">") + hdr + nl
            w = int(rng.choice([0, 0, 37, 120]))
            for ln in ([s[a:a + w] for a in range(0, len(s), w)] if (w and s) else [s]):
                out += ln + nl
            if fastq:
                q = qchars[rng.integers(0, len(qchars), len(s))].tobytes()
                out += b"+\n"
                for ln in ([q[a:a + w] for a in range(0, len(q), w)] if (w and q) else [q]):
                    out += ln + b"\n"
            if rng.random() < 0.1:
                out += nl
        if trial % 3 == 0:
            out = out.rstrip(b"\r\n")                              # no newline at the end of the file
        src = os.path.join(d, "f.fastq" if fastq else "f.fasta")
        open(src, "wb").write(bytes(out))
        cls = os.path.join(d, "f.class")
        base = {"CPG_FAKE_NULL": "1", "CPG_FAKE_DEVICES": "2"}
        p = run([CLI, "-T2", "-c8", src], env=dict(base, CPG_SERIAL_IO="1"))
        assert p.returncode == 0, p.stderr[-800:]
        want = open(cls, "rb").read()
        assert want.count(b"\n") == 4 * sim.nreads
        env = dict(base, CPG_BATCH_BASES=str(int(rng.choice([700, 5000, 40000, 10**7]))),
                   CPG_PARSE_PIECES=str(int(rng.choice([1, 2, 5, 16, 40]))))
        p = run([CLI, "-T%d" % rng.integers(1, 5), "-c8", src], env=env)
        assert p.returncode == 0, (env, p.stderr[-800:])
        assert open(cls, "rb").read() == want, (trial, env)
        os.remove(src)
