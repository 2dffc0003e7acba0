"""Regenerates the golden fixtures of tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/ClassPro, built from /root/reference by oracle/Makefile) on seeded synthetic inputs.

    python tests/golden/make_golden.py

For every dataset <name> it stores
    <name>.fasta.gz            the reads (input)
    <name>.hist.gz             FastK histogram (input; gzip only to keep the repo small)
    <name>.prof, <name>.pidx.N, <name>.prof.N   FastK profile stub, index and data parts (inputs;
                               stored without the leading dot of the hidden part files)
    <name>.class.gz            the reference's output (golden)
    <name>.args                the ClassPro options used
The reference cannot travel to the GPU box; these files can.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpkit  # noqa: E402

DATASETS = {
    # name: (cpsim parameters, ClassPro options)
    "g1": (dict(seed=21, genome_len=25000, cov=24., het=0.01, repeat_frac=0.3, seg_dups=0, len_mean=6000,
                len_sd=1500, len_min=1000, len_max=20000, short_reads=1, nparts=2), []),
    "g2": (dict(seed=22, genome_len=20000, cov=40., het=0.004, repeat_frac=0.6, len_mean=9000, len_sd=2000,
                len_min=2000, len_max=20000, err_indel_hp=0.002, nparts=1), ["-c38", "-r9000"]),
}


def main():
    if not cpkit.have_reference():
        raise SystemExit("oracle/_ref/ClassPro is missing: run `make -C oracle` where /root/reference exists")
    for name, (params, args) in DATASETS.items():
        with tempfile.TemporaryDirectory() as tmp:
            sim = cpkit.simulate(write_to=tmp, root=name, **params)
            fasta = os.path.join(tmp, name + ".fasta")
            out = cpkit.run_reference(fasta, args=args, threads=1)
            with open(fasta, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".fasta.gz"), "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
            with open(os.path.join(tmp, name + ".hist"), "rb") as f, \
                    gzip.GzipFile(os.path.join(HERE, name + ".hist.gz"), "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
            with open(out, "rb") as f, gzip.GzipFile(os.path.join(HERE, name + ".class.gz"), "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
            shutil.copy(os.path.join(tmp, name + ".prof"), os.path.join(HERE, name + ".prof"))
            p = 1
            while os.path.exists(os.path.join(tmp, ".%s.pidx.%d" % (name, p))):
                shutil.copy(os.path.join(tmp, ".%s.pidx.%d" % (name, p)), os.path.join(HERE, "%s.pidx.%d" % (name, p)))
                shutil.copy(os.path.join(tmp, ".%s.prof.%d" % (name, p)), os.path.join(HERE, "%s.prof.%d" % (name, p)))
                p += 1
            with open(os.path.join(HERE, name + ".args"), "w") as f:
                f.write(" ".join(args) + "\n")
            print(name, "reads", sim.nreads, "k-mers", sim.total_kmers)


if __name__ == "__main__":
    main()
