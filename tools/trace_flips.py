"""Find the reads whose class string differs between the GPU and the oracle on one simulated dataset
and write them (index, differing ranges, both strings) to gpurun_out/flips_<tag>.json, so that the
comparison behind each flip can be traced off-line with the oracle.

    python tools/trace_flips.py <tag> <cov_opt> <read_len> key=value ...      (cpkit.simulate parameters)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpkit  # noqa: E402
import classpro_b200 as cp  # noqa: E402
from classpro_b200.abi import pack_codes  # noqa: E402


def main():
    tag, cov_opt, read_len = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    kw = {}
    for a in sys.argv[4:]:
        k, v = a.split("=")
        kw[k] = float(v) if "." in v else int(v)
    sim = cpkit.simulate(**kw)
    keep = np.nonzero(sim.rlen >= sim.kmer)[0]
    om = cpkit.oracle_model(sim, cov_opt, read_len)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=read_len)
    rl = sim.rlen[keep]
    codes = np.concatenate([sim.seq[sim.seq_off[i]:sim.seq_off[i + 1]] for i in keep])
    so = np.zeros(len(keep) + 1, np.int64)
    np.cumsum(rl, out=so[1:])
    packed, poff = pack_codes(codes, so, rl)
    parts = [sim.read_prof(i) for i in keep]
    prof = np.concatenate(parts)
    pro = np.zeros(len(keep) + 1, np.int64)
    np.cumsum([len(p) for p in parts], out=pro[1:])
    batch = cp.Batch(packed, poff, rl, prof, pro, 2)
    ctx = cp.Context(gm)
    cls, status = ctx.classify(batch)
    ctx.close()
    ow = cpkit.OracleWork(clean=True)
    out, kmers, flips = [], 0, 0
    for k, i in enumerate(keep):
        a = ow.classify(om, sim.read_ascii(i).tobytes(), sim.read_counts(i))
        b = cls[batch.cls_off[k]:batch.cls_off[k + 1]].tobytes()
        kmers += len(a) - sim.kmer + 1
        if a != b:
            diff = [j for j in range(len(a)) if a[j] != b[j]]
            flips += len(diff)
            out.append({"read": int(i), "rlen": len(a), "status": int(status[k]), "n_diff": len(diff),
                        "first": diff[0], "last": diff[-1], "oracle": a.decode(), "gpu": b.decode()})
    res = {"tag": tag, "params": kw, "cov_opt": cov_opt, "read_len": read_len, "kmers": kmers, "flips": flips, "reads": out}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "flips_%s.json" % tag), "w"))
    print("kmers %d flips %d reads %d" % (kmers, flips, len(out)))
    for r in out[:10]:
        print(r["read"], r["rlen"], r["n_diff"], r["first"], r["last"], r["status"])


if __name__ == "__main__":
    main()
