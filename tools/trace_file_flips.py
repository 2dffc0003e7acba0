"""Attribute the class characters that differ between this repository's ClassPro and the reference binary on a
file-level run (tools/run_big.py --keep DIR, or tools/run_config.py --keep DIR) to comparisons that sit on a tie.

For every read whose 4th line differs: the differing ranges, then the read is classified by the oracle
(oracle/classpro_oracle.c, which must reproduce the REFERENCE's line) with its tie log on (cpo_set_trace: every
arg-max comparison and every coverage truncation whose relative gap is below 1e-6 is printed); the log lines with
a relative gap <= 1e-9 are the candidates the north star allows a flip to sit on.

    python tools/trace_file_flips.py DIR [max_reads]      -> one JSON document on stdout
"""
import json
import os
import re
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit  # noqa: E402

K = 40


def records(path):
    with open(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                return
            s = f.readline().rstrip(b"\n")
            f.readline()
            c = f.readline().rstrip(b"\n")
            yield h.rstrip(b"\n"), s, c


class Profiles:
    """Reader of the FastK profile files: part p holds reads first[p] .. first[p]+n[p]-1."""

    def __init__(self, d, root):
        with open(os.path.join(d, root + ".prof"), "rb") as f:
            self.kmer, self.nparts = struct.unpack("<ii", f.read(8))
        self.parts = []
        first = 0
        for p in range(1, self.nparts + 1):
            with open(os.path.join(d, ".%s.pidx.%d" % (root, p)), "rb") as f:
                _, _, n = struct.unpack("<iqq", f.read(20))
                off = np.frombuffer(f.read(8 * n), dtype=np.int64)
            self.parts.append((first, n, off, os.path.join(d, ".%s.prof.%d" % (root, p))))
            first += n

    def read(self, i):
        for first, n, off, path in self.parts:
            if first <= i < first + n:
                j = i - first
                b = 0 if j == 0 else int(off[j - 1])
                e = int(off[j])
                with open(path, "rb") as f:
                    f.seek(b)
                    return np.frombuffer(f.read(e - b), dtype=np.uint8)
        raise IndexError(i)


def main():
    d = sys.argv[1]
    max_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    root = "reads"
    mine, ref = os.path.join(d, root + ".class"), os.path.join(d, root + ".ref.class")
    diffs = []
    kmers = 0
    for i, (a, b) in enumerate(zip(records(mine), records(ref))):
        kmers += max(0, len(a[1]) - K + 1)
        if a[2] != b[2]:
            x, y = np.frombuffer(a[2], np.uint8), np.frombuffer(b[2], np.uint8)
            n = min(len(x), len(y))
            w = np.nonzero(x[:n] != y[:n])[0]
            runs = []
            if len(w):
                cut = np.nonzero(np.diff(w) > 1)[0]
                st = np.concatenate([[0], cut + 1])
                en = np.concatenate([cut, [len(w) - 1]])
                runs = [[int(w[s]), int(w[e]) + 1, chr(y[w[s]]), chr(x[w[s]])] for s, e in zip(st, en)]
            diffs.append({"read": i, "rlen": len(a[1]), "n_diff": int(len(w)) + abs(len(x) - len(y)),
                          "ranges_ref_gpu": runs, "seq": a[1], "ref": b[2]})
    out = {"dir": d, "kmers": kmers, "reads_differing": len(diffs), "flipped_characters": sum(x["n_diff"] for x in diffs),
           "flip_fraction": sum(x["n_diff"] for x in diffs) / max(1, kmers), "reads": []}
    if diffs:
        P = Profiles(d, root)
        hist = np.fromfile(os.path.join(d, root + ".hist"), dtype=np.uint8)
        k, low, high = struct.unpack("<iii", hist[:12].tobytes())
        il, ih = struct.unpack("<qq", hist[12:28].tobytes())
        h = np.frombuffer(hist[28:].tobytes(), dtype=np.int64)

        class S:            # what cpkit.oracle_model looks at
            pass
        sim = S()
        sim.kmer = k
        sim.hist = np.zeros(32770, np.int64)
        sim.hist[low:high + 1] = h
        sim.hist[32768], sim.hist[32769] = il, ih
        om = cpkit.oracle_model(sim, 0, 20000)
        L = cpkit.oracle_lib()
        ow = cpkit.OracleWork(clean=True)
        hs = cpkit.hostsim_lib()
        gm = cpkit.gpu_model_from_sim(hs, sim, 0, 20000)
        dump = {"hist": sim.hist}
        for x in diffs[:max_reads]:
            n, counts = cpkit.oracle_decode(P.read(x["read"]), x["rlen"] - K + 1)
            # the device logic compiled for the host (glibc exp/log): equal to the reference = the GPU's
            # difference comes from the CUDA math library's last place, not from the logic
            st, hcls = cpkit.hostsim_classify(gm, x["seq"], counts, lib=hs)
            dump["seq_%d" % x["read"]] = np.frombuffer(x["seq"], np.uint8)
            dump["cnt_%d" % x["read"]] = counts
            dump["ref_%d" % x["read"]] = np.frombuffer(x["ref"], np.uint8)
            log = os.path.join(d, "trace_%d.log" % x["read"])
            fd = os.open(log, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
            saved = os.dup(2)
            try:
                os.dup2(fd, 2)
                L.cpo_set_trace(1)
                o = ow.classify(om, x["seq"], counts)
            finally:
                L.cpo_set_trace(0)
                os.dup2(saved, 2)
                os.close(fd)
                os.close(saved)
            ties = []
            for l in open(log):
                m = re.search(r"(rel\.gap|within) ([0-9.e+-]+)", l)
                if m and float(m.group(2)) <= 1e-9:
                    ties.append(l.strip())
            out["reads"].append({"read": x["read"], "rlen": x["rlen"], "n_diff": x["n_diff"],
                                 "ranges_[from,to,ref,gpu]": x["ranges_ref_gpu"],
                                 "oracle_equals_reference": bool(o == x["ref"]),
                                 "host_build_of_device_code_equals_reference": bool(hcls == x["ref"]),
                                 "comparisons_within_1e-9": ties[:12], "n_comparisons_within_1e-9": len(ties)})
            os.remove(log)
        if len(sys.argv) > 3:
            np.savez_compressed(sys.argv[3], **dump)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
