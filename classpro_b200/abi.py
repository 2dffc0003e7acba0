"""ctypes binding of libclasspro_b200.so (include/classpro_gpu.h).  Plumbing only."""
import ctypes as C
import os
import numpy as np

LIB_PATH = os.environ.get("CLASSPRO_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libclasspro_b200.so")

ST_FATAL = 1 | 2 | 4 | 8 | 64   # conditions on which the reference exits (cpg_common.h)


class CpgError(RuntimeError):
    pass


class CModel(C.Structure):
    """cpg_model"""
    _fields_ = [("kmer", C.c_int32), ("read_len", C.c_int32), ("cov", C.c_uint16 * 4),
                ("dr_ratio", C.c_double), ("cmax", C.c_int32), ("hc_erate", C.c_double),
                ("lmax", C.c_int32 * 3), ("pe", (C.c_double * 21) * 3),
                ("cthres", C.c_uint8 * (36 * 256 * 4)), ("logfact", C.c_double * 32768)]


class CBatch(C.Structure):
    """cpg_batch"""
    _fields_ = [("n_reads", C.c_int32), ("seq_bits", C.c_int32), ("seq", C.c_void_p),
                ("seq_off", C.c_void_p), ("rlen", C.c_void_p), ("prof", C.c_void_p),
                ("prof_off", C.c_void_p)]


class CResult(C.Structure):
    """cpg_result"""
    _fields_ = [("cls", C.c_void_p), ("cls_off", C.c_void_p), ("status", C.c_void_p)]


_lib = None


def lib():
    """Load the shared library; there is no fallback if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CpgError("%s not found: build it with `make -C classpro_b200` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.cpg_model_from_hist.argtypes = [C.POINTER(CModel), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.cpg_model_load.argtypes = [C.POINTER(CModel), C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.cpg_model_from_cov.argtypes = [C.POINTER(CModel), C.c_int, C.c_int, C.c_int, C.c_int]
        L.cpg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(CModel), C.c_int64, C.c_int32]
        L.cpg_destroy.argtypes = [C.c_void_p]
        L.cpg_last_error.argtypes = [C.c_void_p]
        L.cpg_last_error.restype = C.c_char_p
        L.cpg_pack_seq.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.cpg_classify.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(CResult)]
        L.cpg_submit.argtypes = [C.c_void_p, C.c_int, C.POINTER(CBatch)]
        L.cpg_collect.argtypes = [C.c_void_p, C.c_int, C.POINTER(CResult)]
        L.cpg_prof2class.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cpg_decode_profiles.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.cpg_upload.argtypes = [C.c_void_p, C.POINTER(CBatch)]
        L.cpg_run_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       C.POINTER(C.c_int)]
        L.cpg_download.argtypes = [C.c_void_p, C.POINTER(CResult)]
        L.cpg_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.cpg_wall_ns.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.cpg_batch_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.cpg_set_result_mode.argtypes = [C.c_void_p, C.c_int]
        L.cpg_intervals_bound.argtypes = [C.c_void_p, C.c_int]
        L.cpg_intervals_bound.restype = C.c_int64
        L.cpg_collect_intervals.argtypes = [C.c_void_p, C.c_int, C.POINTER(CResultIvl)]
        L.cpg_expand_intervals.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.cpg_expand_intervals.restype = None
        L.cpg_host_alloc.argtypes = [C.c_size_t]
        L.cpg_host_alloc.restype = C.c_void_p
        L.cpg_host_free.argtypes = [C.c_void_p]
        L.cpg_status_string.argtypes = [C.c_int32]
        L.cpg_status_string.restype = C.c_char_p
        L.cpg_version.restype = C.c_char_p
        L.cpg_device_count.restype = C.c_int
        L.cpg_count_kmers.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.cpg_encode_profiles.argtypes = [C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.cpg_count_error.restype = C.c_char_p
        _lib = L
    return _lib


class Model:
    """Host one-shot model (cpg_model)."""

    def __init__(self, c):
        self.c = c

    @classmethod
    def from_hist(cls, kmer, hist, ilow, ihigh, low=1, cov_opt=0, read_len=20000, verbose=0):
        hist = np.ascontiguousarray(hist, dtype=np.int64)
        m = CModel()
        rc = lib().cpg_model_from_hist(C.byref(m), kmer, low, low + len(hist) - 1, int(ilow), int(ihigh),
                                       hist.ctypes.data, cov_opt, read_len, verbose)
        if rc:
            raise CpgError("cpg_model_from_hist rc=%d" % rc)
        return cls(m)

    @classmethod
    def load(cls, fk_root, cov_opt=0, read_len=20000, verbose=0):
        m = CModel()
        rc = lib().cpg_model_load(C.byref(m), fk_root.encode(), cov_opt, read_len, verbose)
        if rc:
            raise CpgError("cpg_model_load rc=%d" % rc)
        return cls(m)

    @classmethod
    def from_cov(cls, kmer, h, d, read_len=20000):
        m = CModel()
        rc = lib().cpg_model_from_cov(C.byref(m), kmer, h, d, read_len)
        if rc:
            raise CpgError("cpg_model_from_cov rc=%d" % rc)
        return cls(m)

    @property
    def kmer(self):
        return self.c.kmer

    @property
    def cov(self):
        return list(self.c.cov)


class CResultIvl(C.Structure):
    _fields_ = [("ivl", C.c_void_p), ("ivl_cap", C.c_int64), ("ivl_at", C.c_void_p), ("ivl_n", C.c_void_p),
                ("status", C.c_void_p), ("ivl_used", C.c_int64)]


class IntervalResult:
    """Host buffers of a compact result (cpg_result_ivl); pinned=True takes them from cpg_host_alloc."""

    def __init__(self, n_reads, cap, pinned=False):
        self.n = int(n_reads)
        self.cap = int(cap)
        if pinned:
            self._pin = [PinnedArray(4 * max(self.cap, 1)), PinnedArray(8 * (self.n + 1)), PinnedArray(4 * (self.n + 1)),
                         PinnedArray(4 * (self.n + 1))]
            self.ivl = self._pin[0].view(np.uint32)
            self.at = self._pin[1].view(np.int64)
            self.cnt = self._pin[2].view(np.int32)
            self.status = self._pin[3].view(np.int32)
        else:
            self.ivl = np.zeros(max(self.cap, 1), np.uint32)
            self.at = np.zeros(self.n + 1, np.int64)
            self.cnt = np.zeros(self.n + 1, np.int32)
            self.status = np.zeros(self.n + 1, np.int32)
        self.used = 0
        self.c = CResultIvl(self.ivl.ctypes.data, self.cap, self.at.ctypes.data, self.cnt.ctypes.data,
                            self.status.ctypes.data, 0)

    def expand(self, K, rlen, out=None):
        """Class strings of all reads, read after read (cpg_expand_intervals: host code)."""
        L = lib()
        rlen = np.asarray(rlen, np.int64)
        off = np.zeros(self.n + 1, np.int64)
        np.cumsum(rlen[:self.n], out=off[1:])
        if out is None:
            out = np.zeros(int(off[-1]) + 1, np.uint8)
        for r in range(self.n):
            L.cpg_expand_intervals(K, int(rlen[r]), self.ivl.ctypes.data + 4 * int(self.at[r]), int(self.cnt[r]),
                                   out.ctypes.data + int(off[r]))
        return out

    def free(self):
        for p in getattr(self, "_pin", []):
            p.free()


class PinnedArray:
    """numpy view of page-locked host memory from cpg_host_alloc."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().cpg_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise CpgError("cpg_host_alloc(%d) failed" % nbytes)
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.u8 = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype):
        return self.u8.view(dtype)

    def free(self):
        if self.ptr:
            self.u8 = None
            lib().cpg_host_free(self.ptr)
            self.ptr = None


def pack_reads(ascii_reads):
    """2-bit pack a list of ASCII reads (bytes/np.uint8 arrays) -> (packed u8 array, seq_off)."""
    L = lib()
    offs = np.zeros(len(ascii_reads) + 1, dtype=np.int64)
    for i, r in enumerate(ascii_reads):
        offs[i + 1] = offs[i] + (len(r) + 3) // 4
    out = np.zeros(int(offs[-1]) + 16, dtype=np.uint8)
    for i, r in enumerate(ascii_reads):
        b = np.ascontiguousarray(np.frombuffer(bytes(r), dtype=np.uint8))
        if L.cpg_pack_seq(b.ctypes.data, len(b), out.ctypes.data + int(offs[i])):
            raise CpgError("read %d has a character outside ACGT: ship it with seq_bits=8" % i)
    return out, offs


def pack_codes(codes, seq_off, rlen):
    """Vectorised 2-bit packing of base codes (0..3) stored read after read."""
    n = len(rlen)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum((rlen.astype(np.int64) + 3) // 4, out=offs[1:])
    out = np.zeros(int(offs[-1]) + 16, dtype=np.uint8)
    for i in range(n):
        c = codes[seq_off[i]:seq_off[i + 1]]
        pad = (-len(c)) % 4
        if pad:
            c = np.concatenate([c, np.zeros(pad, dtype=np.uint8)])
        c = c.reshape(-1, 4)
        out[offs[i]:offs[i + 1]] = c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)
    return out, offs


class Batch:
    """A cpg_batch over numpy arrays (kept alive by this object)."""

    def __init__(self, seq, seq_off, rlen, prof, prof_off, seq_bits=2):
        self.seq = np.ascontiguousarray(seq, dtype=np.uint8)
        self.seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
        self.rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        self.prof = np.ascontiguousarray(prof, dtype=np.uint8)
        self.prof_off = np.ascontiguousarray(prof_off, dtype=np.int64)
        self.n = len(self.rlen)
        self.c = CBatch(self.n, seq_bits, self.seq.ctypes.data, self.seq_off.ctypes.data,
                        self.rlen.ctypes.data, self.prof.ctypes.data, self.prof_off.ctypes.data)
        self.cls_off = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(self.rlen.astype(np.int64), out=self.cls_off[1:])

    def kmers(self, K):
        return int(np.maximum(self.rlen.astype(np.int64) - K + 1, 0).sum())


class Context:
    """cpg_ctx on one GPU."""

    def __init__(self, model, device=0):
        self.L = lib()
        self.model = model
        h = C.c_void_p()
        rc = self.L.cpg_create(C.byref(h), device, C.byref(model.c), 0, 0)
        if rc:
            raise CpgError("cpg_create: %s" % self.L.cpg_last_error(None).decode())
        self.h = h

    def _err(self, what, rc):
        return CpgError("%s rc=%d: %s" % (what, rc, self.L.cpg_last_error(self.h).decode()))

    def _result(self, batch, cls=None):
        if cls is None:
            cls = np.zeros(int(batch.cls_off[-1]) + 1, dtype=np.uint8)
        status = np.zeros(batch.n + 1, dtype=np.int32)
        res = CResult(cls.ctypes.data, batch.cls_off.ctypes.data, status.ctypes.data)
        return res, cls, status

    def classify(self, batch, cls=None, allow_read_errors=True):
        """Host buffers in, host buffers out (cpg_classify). Returns (cls bytes array, status)."""
        res, cls, status = self._result(batch, cls)
        rc = self.L.cpg_classify(self.h, C.byref(batch.c), C.byref(res))
        if rc and not (rc == 5 and allow_read_errors):
            raise self._err("cpg_classify", rc)
        return cls, status[:batch.n]

    def submit(self, slot, batch):
        rc = self.L.cpg_submit(self.h, slot, C.byref(batch.c))
        if rc:
            raise self._err("cpg_submit", rc)

    def collect(self, slot, batch, cls=None, allow_read_errors=True):
        res, cls, status = self._result(batch, cls)
        rc = self.L.cpg_collect(self.h, slot, C.byref(res))
        if rc and not (rc == 5 and allow_read_errors):
            raise self._err("cpg_collect", rc)
        return cls, status[:batch.n]

    def set_result_mode(self, intervals):
        """False: class strings (cpg_collect); True: packed interval tables (collect_intervals)."""
        rc = self.L.cpg_set_result_mode(self.h, 1 if intervals else 0)
        if rc:
            raise self._err("cpg_set_result_mode", rc)

    def intervals_bound(self, slot):
        return int(self.L.cpg_intervals_bound(self.h, slot))

    def collect_intervals(self, slot, res, allow_read_errors=True):
        """res: IntervalResult with room for intervals_bound(slot) entries."""
        rc = self.L.cpg_collect_intervals(self.h, slot, C.byref(res.c))
        if rc and not (rc == 5 and allow_read_errors):
            raise self._err("cpg_collect_intervals", rc)
        res.used = int(res.c.ivl_used)
        return res

    def upload(self, batch):
        rc = self.L.cpg_upload(self.h, C.byref(batch.c))
        if rc:
            raise self._err("cpg_upload", rc)

    def run_resident(self, iters=1):
        a, b, n = C.c_float(), C.c_float(), C.c_int()
        rc = self.L.cpg_run_resident(self.h, iters, C.byref(a), C.byref(b), C.byref(n))
        if rc:
            raise self._err("cpg_run_resident", rc)
        return a.value, b.value, n.value

    def phase_cycles(self):
        """Device time of k_wall, k_rel, k_unrel and the retry launch in the last timed resident run,
        nanoseconds (per-phase clock cycles with CPG_FUSED=1)."""
        out = (C.c_uint64 * 4)()
        rc = self.L.cpg_phase_cycles(self.h, out)
        if rc:
            raise self._err("cpg_phase_cycles", rc)
        return list(out)

    def wall_ns(self):
        """Device time of k_wall_a, k_wall_b, k_wall_c, k_unrel_a, k_unrel_b, k_emit in the last timed resident run, nanoseconds."""
        out = (C.c_uint64 * 6)()
        rc = self.L.cpg_wall_ns(self.h, out)
        if rc:
            raise self._err("cpg_wall_ns", rc)
        return list(out)

    def batch_stats(self):
        """(wall candidates, intervals, reliable intervals, intervals visited by the unreliable sweeps) of the resident batch."""
        out = (C.c_int64 * 4)()
        rc = self.L.cpg_batch_stats(self.h, out)
        if rc:
            raise self._err("cpg_batch_stats", rc)
        return list(out)

    def download(self, batch, allow_read_errors=True):
        res, cls, status = self._result(batch)
        rc = self.L.cpg_download(self.h, C.byref(res))
        if rc and not (rc == 5 and allow_read_errors):
            raise self._err("cpg_download", rc)
        return cls, status[:batch.n]

    def decode_profiles(self, prof, prof_off, caps):
        """Fetch_Profile replacement alone: returns (counts, cnt_off, plen)."""
        prof = np.ascontiguousarray(prof, dtype=np.uint8)
        prof_off = np.ascontiguousarray(prof_off, dtype=np.int64)
        n = len(prof_off) - 1
        cnt_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.asarray(caps, dtype=np.int64), out=cnt_off[1:])
        counts = np.zeros(int(cnt_off[-1]) + 8, dtype=np.uint16)
        plen = np.zeros(n + 1, dtype=np.int32)
        rc = self.L.cpg_decode_profiles(self.h, n, prof.ctypes.data, prof_off.ctypes.data,
                                        cnt_off.ctypes.data, counts.ctypes.data, plen.ctypes.data)
        if rc:
            raise self._err("cpg_decode_profiles", rc)
        return counts, cnt_off, plen[:n]

    def prof2class(self, prof, prof_off, rlen):
        """prof2class on the device: relative profiles -> ground-truth class strings (reads concatenated).
        Returns (cls, cls_off, status); raises if a profile length is not rlen-K+1."""
        prof = np.ascontiguousarray(prof, dtype=np.uint8)
        prof_off = np.ascontiguousarray(prof_off, dtype=np.int64)
        rlen = np.ascontiguousarray(rlen, dtype=np.int32)
        n = len(rlen)
        cls_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(rlen.astype(np.int64), out=cls_off[1:])
        cls = np.zeros(int(cls_off[-1]) + 16, dtype=np.uint8)
        status = np.zeros(n + 1, dtype=np.int32)
        rc = self.L.cpg_prof2class(self.h, n, prof.ctypes.data, prof_off.ctypes.data, rlen.ctypes.data,
                                   cls.ctypes.data, status.ctypes.data)
        if rc:
            raise self._err("cpg_prof2class", rc)
        return cls, cls_off, status[:n]

    def close(self):
        if self.h:
            self.L.cpg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- profile producer (include/classpro_gpu.h: cpg_count_kmers / cpg_encode_profiles) ----
def count_kmers(kmer, pseq, seq_off, rlen, device=0):
    """Exact canonical k-mer counts of a 2-bit packed read set on the GPU: (counts, cnt_off, hist)."""
    L = lib()
    n = len(rlen)
    rlen = np.ascontiguousarray(rlen, dtype=np.int32)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.int64)
    pseq = np.ascontiguousarray(pseq, dtype=np.uint8)
    cnt_off = np.zeros(n + 1, dtype=np.int64)
    total = int(np.maximum(rlen.astype(np.int64) - kmer + 1, 0).sum())
    counts = np.zeros(max(total, 1), dtype=np.uint16)
    hist = np.zeros(32770, dtype=np.int64)
    rc = L.cpg_count_kmers(device, kmer, n, pseq.ctypes.data, seq_off.ctypes.data, rlen.ctypes.data,
                           cnt_off.ctypes.data, counts.ctypes.data, hist.ctypes.data)
    if rc:
        raise CpgError("cpg_count_kmers: rc %d: %s" % (rc, L.cpg_count_error().decode()))
    return counts[:total], cnt_off, hist


def encode_profiles(counts, cnt_off, device=0):
    """FastK token streams of the counts of every read on the GPU: (prof, prof_off)."""
    L = lib()
    n = len(cnt_off) - 1
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    cnt_off = np.ascontiguousarray(cnt_off, dtype=np.int64)
    cap = 2 * len(counts) + 16
    prof = np.zeros(cap, dtype=np.uint8)
    prof_off = np.zeros(n + 1, dtype=np.int64)
    rc = L.cpg_encode_profiles(device, n, counts.ctypes.data, cnt_off.ctypes.data, prof.ctypes.data, cap, prof_off.ctypes.data)
    if rc:
        raise CpgError("cpg_encode_profiles: rc %d: %s" % (rc, L.cpg_count_error().decode()))
    return prof[:prof_off[n]], prof_off
