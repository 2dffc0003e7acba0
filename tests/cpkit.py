"""Test/bench harness helpers (ctypes bindings of the checker libraries).

Everything here is test infrastructure:
  * tools/libcpsim.so        seeded synthetic data generator (tools/cpsim.c)
  * oracle/_ref/liboracle.so CPU restatement of the reference (oracle/classpro_oracle.c)
  * oracle/_ref/ClassPro     the unmodified reference binary (built where /root/reference exists)
  * tests/hostsim/_build/libhostsim.so   host-compiled device logic (warp width 1), CPU tests only
The product library is bound separately in classpro_b200/abi.py.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_BIN = os.path.join(REF_DIR, "ClassPro")
TOOLS_DIR = os.path.join(ROOT, "tools")
HOSTSIM_DIR = os.path.join(ROOT, "tests", "hostsim")

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _run(cmd, **kw):
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)


def build_oracle():
    """Compile oracle/'s C restatement (and, where the reference tree exists, the reference)."""
    _run(["make", "-C", ORACLE_DIR, "--no-print-directory", "all"])
    return os.path.join(REF_DIR, "liboracle.so")


def build_product():
    """libclasspro_b200.so + the command-line programs (nvcc cross-compiles without a GPU)."""
    _run(["make", "-C", os.path.join(ROOT, "classpro_b200"), "--no-print-directory", "all"])
    build_oracle()


def build_cpsim():
    so = os.path.join(TOOLS_DIR, "libcpsim.so")
    src = os.path.join(TOOLS_DIR, "cpsim.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"])
    exe = os.path.join(TOOLS_DIR, "cpsim")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        _run(["gcc", "-O2", "-DCPSIM_MAIN", "-o", exe, src, "-lm"])
    return so


def build_hostsim():
    """tests/hostsim/build.sh, once: skipped when every output is newer than every source, and under a
    file lock, because the two ranks of the gloo test (and parallel test workers) call this at the same
    time and must not load a library another process is still writing."""
    import fcntl
    import glob
    bdir = os.path.join(HOSTSIM_DIR, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libhostsim.so")
    outs = [so] + [os.path.join(bdir, f) for f in ("libhostsim32.so", "libcountsim.so", "libpassemu.so", "ClassPro", "prof2class", "profiler")]
    srcs = (glob.glob(os.path.join(ROOT, "classpro_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "classpro_b200", "host", "*"))
            + glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(HOSTSIM_DIR, "*.cpp"))
            + [os.path.join(HOSTSIM_DIR, "build.sh")])
    with open(os.path.join(bdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        newest = max(os.path.getmtime(f) for f in srcs)
        if not all(os.path.exists(o) and os.path.getmtime(o) >= newest for o in outs):
            _run(["bash", os.path.join(HOSTSIM_DIR, "build.sh")])
    return so


# ----------------------------------------------------------------------------- cpsim
class SimParams(C.Structure):
    _fields_ = [("seed", C.c_int64), ("genome_len", C.c_int64), ("het", C.c_double),
                ("snp_only", C.c_int), ("repeat_frac", C.c_double), ("seg_dups", C.c_int64),
                ("cov", C.c_double), ("len_mean", C.c_int64), ("len_sd", C.c_int64),
                ("len_min", C.c_int64), ("len_max", C.c_int64),
                ("err_sub", C.c_double), ("err_indel_base", C.c_double), ("err_indel_hp", C.c_double),
                ("kmer", C.c_int64), ("nparts", C.c_int64), ("exact", C.c_int), ("short_reads", C.c_int)]


class _SimDataC(C.Structure):
    _fields_ = [("kmer", C.c_int), ("nreads", C.c_int64), ("total_bases", C.c_int64),
                ("total_kmers", C.c_int64), ("prof_bytes", C.c_int64),
                ("seq", C.POINTER(C.c_uint8)), ("seq_off", C.POINTER(C.c_int64)),
                ("rlen", C.POINTER(C.c_int32)), ("counts", C.POINTER(C.c_uint16)),
                ("cnt_off", C.POINTER(C.c_int64)), ("prof", C.POINTER(C.c_uint8)),
                ("prof_off", C.POINTER(C.c_int64)), ("hist", C.POINTER(C.c_int64)),
                ("hdr", C.POINTER(C.c_char)), ("hdr_off", C.POINTER(C.c_int64))]


_cpsim = None


def cpsim_lib():
    global _cpsim
    if _cpsim is None:
        _cpsim = C.CDLL(build_cpsim())
        _cpsim.cpsim_generate.argtypes = [C.POINTER(SimParams), C.POINTER(_SimDataC)]
        _cpsim.cpsim_write_files.argtypes = [C.POINTER(SimParams), C.POINTER(_SimDataC), C.c_char_p, C.c_char_p]
        _cpsim.cpsim_free.argtypes = [C.POINTER(_SimDataC)]
        _cpsim.cpsim_default_params.argtypes = [C.POINTER(SimParams)]
    return _cpsim


class SimData:
    """Synthetic dataset held in numpy arrays."""

    def __init__(self, params, d):
        n = d.nreads
        self.params = params
        self.kmer = d.kmer
        self.nreads = int(n)
        self.total_bases = int(d.total_bases)
        self.total_kmers = int(d.total_kmers)

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt, copy=True)
        self.seq = arr(d.seq, d.total_bases, np.uint8)            # codes 0..3
        self.seq_off = arr(d.seq_off, n + 1, np.int64)
        self.rlen = arr(d.rlen, n, np.int32)
        self.counts = arr(d.counts, d.total_kmers, np.uint16)
        self.cnt_off = arr(d.cnt_off, n + 1, np.int64)
        self.prof = arr(d.prof, d.prof_bytes, np.uint8)
        self.prof_off = arr(d.prof_off, n + 1, np.int64)
        self.hist = arr(d.hist, 32770, np.int64)
        hdr = C.string_at(d.hdr, int(d.hdr_off[n])) if n else b""
        ho = arr(d.hdr_off, n + 1, np.int64)
        self.headers = [hdr[ho[i]:ho[i + 1]] for i in range(n)]
        self.ascii = BASES[self.seq]                               # 'A','C','G','T' bytes

    def read_ascii(self, i):
        return self.ascii[self.seq_off[i]:self.seq_off[i + 1]]

    def read_counts(self, i):
        return self.counts[self.cnt_off[i]:self.cnt_off[i + 1]]

    def read_prof(self, i):
        return self.prof[self.prof_off[i]:self.prof_off[i + 1]]


def sim_params(**kw):
    p = SimParams()
    cpsim_lib().cpsim_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def simulate(write_to=None, root="x", **kw):
    """Generate a dataset; optionally also write FASTA + FastK files into directory `write_to`."""
    lib = cpsim_lib()
    p = sim_params(**kw)
    d = _SimDataC()
    if lib.cpsim_generate(C.byref(p), C.byref(d)) != 0:
        raise RuntimeError("cpsim_generate failed")
    try:
        if write_to is not None:
            os.makedirs(write_to, exist_ok=True)
            lib.cpsim_write_files(C.byref(p), C.byref(d), write_to.encode(), root.encode())
        return SimData(p, d)
    finally:
        lib.cpsim_free(C.byref(d))


# ----------------------------------------------------------------------------- oracle
class OracleModel(C.Structure):
    _fields_ = [("K", C.c_int), ("read_len", C.c_int), ("cov", C.c_uint16 * 4),
                ("dr_ratio", C.c_double), ("cmax", C.c_int), ("hc_erate", C.c_double),
                ("lmax", C.c_int * 3), ("pe", (C.c_double * 21) * 3),
                ("cthres", C.c_uint8 * (3 * 21 * 256 * 2 * 2)),
                ("logfact", C.c_double * 32768)]


class OracleIntvl(C.Structure):
    _fields_ = [("b", C.c_int32), ("e", C.c_int32), ("cb", C.c_uint16), ("ce", C.c_uint16),
                ("ccb", C.c_uint16), ("cce", C.c_uint16), ("is_rel", C.c_uint8), ("asgn", C.c_int8),
                ("pe", C.c_double), ("pe_o_b", C.c_double), ("pe_o_e", C.c_double)]


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        so = os.path.join(REF_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "classpro_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build_oracle()
        L = C.CDLL(so)
        L.cpo_model_from_hist.argtypes = [C.POINTER(OracleModel), C.c_int, C.c_int, C.c_int, C.c_int64,
                                          C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int]
        L.cpo_model_from_cov.argtypes = [C.POINTER(OracleModel), C.c_int, C.c_int, C.c_int, C.c_int]
        L.cpo_decode_profile.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.cpo_work_new.restype = C.c_void_p
        L.cpo_work_free.argtypes = [C.c_void_p]
        L.cpo_work_set_clean.argtypes = [C.c_void_p, C.c_int]
        L.cpo_classify_read.argtypes = [C.POINTER(OracleModel), C.c_void_p, C.c_char_p, C.c_int,
                                        C.c_void_p, C.c_int, C.c_char_p]
        L.cpo_last_intervals.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cpo_last_intervals.restype = C.POINTER(OracleIntvl)
        L.cpo_seq_context.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.cpo_run_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                   C.POINTER(C.c_int64)]
        L.cpo_bessi.argtypes = [C.c_int, C.c_double]
        L.cpo_bessi.restype = C.c_double
        _oracle = L
    return _oracle


def _hist_args(sim):
    h = np.ascontiguousarray(sim.hist[1:32768])
    return (1, 32767, int(sim.hist[32768]), int(sim.hist[32769]), h)


def oracle_model(sim, cov_opt=0, read_len=20000):
    m = OracleModel()
    low, high, il, ih, h = _hist_args(sim)
    rc = oracle_lib().cpo_model_from_hist(C.byref(m), sim.kmer, low, high, il, ih,
                                          h.ctypes.data_as(C.POINTER(C.c_int64)), cov_opt, read_len, 0)
    if rc != 0:
        raise RuntimeError("cpo_model_from_hist rc=%d" % rc)
    return m


class OracleWork:
    def __init__(self, clean=True):
        self.L = oracle_lib()
        self.w = self.L.cpo_work_new()
        self.L.cpo_work_set_clean(self.w, 1 if clean else 0)

    def classify(self, model, seq_ascii, counts, want_intervals=False):
        rlen = int(len(seq_ascii))
        plen = int(len(counts))
        cls = C.create_string_buffer(rlen + 1)
        counts = np.ascontiguousarray(counts, dtype=np.uint16)
        n = self.L.cpo_classify_read(C.byref(model), self.w, bytes(seq_ascii), rlen,
                                     counts.ctypes.data, plen, cls)
        if n < 0:
            raise RuntimeError("cpo_classify_read rc=%d" % n)
        if not want_intervals:
            return cls.raw[:rlen]
        N, M = C.c_int(), C.c_int()
        iv = self.L.cpo_last_intervals(self.w, C.byref(N), C.byref(M))
        out = [(iv[i].b, iv[i].e, iv[i].cb, iv[i].ce, iv[i].is_rel,
                iv[i].ccb if iv[i].is_rel else 0, iv[i].cce if iv[i].is_rel else 0,
                iv[i].asgn, iv[i].pe, iv[i].pe_o_b, iv[i].pe_o_e if i + 1 < N.value else 0.0)
               for i in range(N.value)]
        return cls.raw[:rlen], out, M.value

    def __del__(self):
        try:
            self.L.cpo_work_free(self.w)
        except Exception:
            pass


def oracle_decode(prof_bytes, cap):
    prof_bytes = np.ascontiguousarray(prof_bytes, dtype=np.uint8)
    out = np.zeros(max(cap, 1), dtype=np.uint16)
    n = oracle_lib().cpo_decode_profile(prof_bytes.ctypes.data, len(prof_bytes), out.ctypes.data, cap)
    return n, out[:min(n, cap)]


# ----------------------------------------------------------------------------- host-sim of the device logic
class GpuModel(C.Structure):
    """Mirror of cpg_model (include/classpro_gpu.h)."""
    _fields_ = [("kmer", C.c_int32), ("read_len", C.c_int32), ("cov", C.c_uint16 * 4),
                ("dr_ratio", C.c_double), ("cmax", C.c_int32), ("hc_erate", C.c_double),
                ("lmax", C.c_int32 * 3), ("pe", (C.c_double * 21) * 3),
                ("cthres", C.c_uint8 * (36 * 256 * 4)), ("logfact", C.c_double * 32768)]


class GpuIntvl(C.Structure):
    _fields_ = [("b", C.c_int32), ("e", C.c_int32), ("cb", C.c_uint16), ("ce", C.c_uint16),
                ("ccb", C.c_uint16), ("cce", C.c_uint16), ("is_rel", C.c_uint8), ("asgn", C.c_int8),
                ("pad", C.c_uint8 * 6), ("pe", C.c_double), ("peob", C.c_double), ("peoe", C.c_double)]


_hostsim = None
_hostsim32 = None


def _bind_hostsim(L):
    L.cpg_model_from_hist.argtypes = [C.POINTER(GpuModel), C.c_int, C.c_int, C.c_int, C.c_int64,
                                      C.c_int64, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int]
    L.cpg_model_from_cov.argtypes = [C.POINTER(GpuModel), C.c_int, C.c_int, C.c_int, C.c_int]
    L.hs_classify_read.argtypes = [C.POINTER(GpuModel), C.c_char_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_char_p, C.POINTER(GpuIntvl), C.POINTER(C.c_int),
                                   C.POINTER(C.c_int)]
    L.hs_decode_profile.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    L.hs_decode_profile_cand.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.hs_ctx.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.hs_ctx2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    assert L.hs_sizeof_intvl() == C.sizeof(GpuIntvl)
    return L


def hostsim32_lib():
    """The 32-host-threads-per-warp emulation (slow; small inputs only)."""
    global _hostsim32
    if _hostsim32 is None:
        build_hostsim()
        _hostsim32 = _bind_hostsim(C.CDLL(os.path.join(HOSTSIM_DIR, "_build", "libhostsim32.so")))
        assert _hostsim32.hs_warp_width() == 32
    return _hostsim32


def hostsim_lib():
    global _hostsim
    if _hostsim is None:
        L = _bind_hostsim(C.CDLL(build_hostsim()))
        _hostsim = L
    return _hostsim


def gpu_model_from_sim(lib, sim, cov_opt=0, read_len=20000):
    """Host one-shot model through the product's own C code (classpro_b200/host/cpg_model.c)."""
    m = GpuModel()
    low, high, il, ih, h = _hist_args(sim)
    rc = lib.cpg_model_from_hist(C.byref(m), sim.kmer, low, high, il, ih,
                                 h.ctypes.data_as(C.POINTER(C.c_int64)), cov_opt, read_len, 0)
    if rc != 0:
        raise RuntimeError("cpg_model_from_hist rc=%d" % rc)
    return m


def hostsim_classify(model, seq_ascii, counts, seq_bits=2, want_intervals=False, lib=None):
    L = lib or hostsim_lib()
    rlen, plen = int(len(seq_ascii)), int(len(counts))
    cls = C.create_string_buffer(rlen + 1)
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    iv = (GpuIntvl * (plen + 2))() if want_intervals else None
    N, M = C.c_int(), C.c_int()
    st = L.hs_classify_read(C.byref(model), bytes(seq_ascii), rlen, seq_bits, counts.ctypes.data, plen,
                            cls, iv, C.byref(N), C.byref(M))
    if not want_intervals:
        return st, cls.raw[:rlen]
    out = [(iv[i].b, iv[i].e, iv[i].cb, iv[i].ce, iv[i].is_rel,
            iv[i].ccb if iv[i].is_rel else 0, iv[i].cce if iv[i].is_rel else 0,
            iv[i].asgn, iv[i].pe, iv[i].peob, iv[i].peoe if i + 1 < N.value else 0.0)
           for i in range(N.value)]
    return st, cls.raw[:rlen], out, M.value


def hostsim_decode(prof_bytes, cap, lib=None):
    prof_bytes = np.ascontiguousarray(prof_bytes, dtype=np.uint8)
    out = np.zeros(max(cap, 1), dtype=np.uint16)
    n = (lib or hostsim_lib()).hs_decode_profile(prof_bytes.ctypes.data, len(prof_bytes), out.ctypes.data, cap)
    return n, out[:min(n, cap)]


def hostsim_decode_cand(prof_bytes, cap, rcov, lib=None):
    """Counts plus the wall-candidate bit map the decoder writes on the side (as a bool array)."""
    prof_bytes = np.ascontiguousarray(prof_bytes, dtype=np.uint8)
    out = np.zeros(max(cap, 1), dtype=np.uint16)
    cand = np.full((max(cap, 1) + 31) // 32, 0xdeadbeef, dtype=np.uint32)       # the decoder zeroes it
    n = (lib or hostsim_lib()).hs_decode_profile_cand(prof_bytes.ctypes.data, len(prof_bytes), out.ctypes.data, cap,
                                                      cand.ctypes.data, rcov)
    bits = np.unpackbits(cand.view(np.uint8), bitorder="little")[:max(cap, 1)].astype(bool)
    return n, out[:min(n, cap)], bits[:min(n, cap)]


def candidate_bits(counts, rcov):
    """Wall candidates by definition (src/wall.c:594-608): i >= 1, |c[i]-c[i-1]| >= 3, min < rcov."""
    c = counts.astype(np.int64)
    bits = np.zeros(len(c), dtype=bool)
    if len(c) > 1:
        bits[1:] = (np.abs(np.diff(c)) >= 3) & (np.minimum(c[1:], c[:-1]) < rcov)
    return bits


# ----------------------------------------------------------------------------- reference binary
def have_reference():
    return os.path.exists(REF_BIN)


def run_reference(fasta, args=(), threads=1, cwd=None):
    """Run the unmodified reference binary; returns the path of the .class it wrote."""
    cmd = [REF_BIN, "-T%d" % threads] + list(args) + [fasta]
    subprocess.run(cmd, check=True, cwd=cwd or os.path.dirname(fasta),
                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    root = fasta
    for ext in (".fasta.gz", ".fastq.gz", ".fa.gz", ".fq.gz", ".fasta", ".fastq", ".fa", ".fq"):
        if root.endswith(ext):
            root = root[:-len(ext)]
            break
    return root + ".class"


def class_lines(path):
    """The 4th line of every record of a .class file."""
    out = []
    with open(path, "rb") as f:
        for i, line in enumerate(f):
            if i % 4 == 3:
                out.append(line.rstrip(b"\n"))
    return out
