"""Quick on-GPU parity + timing check (development aid; the real tests live in tests/)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cpkit
import classpro_b200 as cp
from classpro_b200.abi import pack_codes

def run(name, cov_opt=0, read_len=20000, **kw):
    sim = cpkit.simulate(**kw)
    keep = np.nonzero(sim.rlen >= sim.kmer)[0]
    om = cpkit.oracle_model(sim, cov_opt, read_len)
    gm = cp.Model.from_hist(sim.kmer, sim.hist[1:32768], sim.hist[32768], sim.hist[32769], cov_opt=cov_opt, read_len=read_len)
    assert list(om.cov) == gm.cov, (list(om.cov), gm.cov)
    print("[%s] ctx..." % name, flush=True)
    ctx = cp.Context(gm)
    rl = sim.rlen[keep]
    seq_parts = [sim.seq[sim.seq_off[i]:sim.seq_off[i+1]] for i in keep]
    codes = np.concatenate(seq_parts); so = np.zeros(len(keep)+1, np.int64); np.cumsum(rl, out=so[1:])
    packed, poff = pack_codes(codes, so, rl)
    prof_parts = [sim.read_prof(i) for i in keep]
    prof = np.concatenate(prof_parts); pro = np.zeros(len(keep)+1, np.int64); np.cumsum([len(p) for p in prof_parts], out=pro[1:])
    # decode parity
    caps = rl.astype(np.int64) - sim.kmer + 1
    print("[%s] decode..." % name, flush=True)
    counts, cnt_off, plen = ctx.decode_profiles(prof, pro, caps)
    print("[%s] decode done" % name, flush=True)
    dec_bad = 0
    for k, i in enumerate(keep):
        if plen[k] != caps[k] or not np.array_equal(counts[cnt_off[k]:cnt_off[k+1]], sim.read_counts(i)): dec_bad += 1
    batch = cp.Batch(packed, poff, rl, prof, pro, 2)
    print("[%s] classify %d reads..." % (name, len(keep)), flush=True)
    t0 = time.time(); cls, status = ctx.classify(batch); t1 = time.time()
    print("[%s] classify done %.3fs" % (name, t1-t0), flush=True)
    ow = cpkit.OracleWork(clean=True)
    flips = 0; bad_reads = 0; nk = 0
    t2 = time.time()
    for k, i in enumerate(keep):
        a = ow.classify(om, sim.read_ascii(i).tobytes(), sim.read_counts(i))
        b = cls[batch.cls_off[k]:batch.cls_off[k+1]].tobytes()
        nk += len(a) - sim.kmer + 1
        if a != b:
            bad_reads += 1
            flips += sum(1 for x, y in zip(a, b) if x != y)
            if bad_reads <= 3: print("   diff read", i, "status", status[k], "len", len(a))
    t3 = time.time()
    ctx.upload(batch)
    ctx.run_resident(2)
    md, mc, nl = ctx.run_resident(5)
    res = dict(name=name, reads=int(len(keep)), kmers=int(nk), decode_mismatch=dec_bad, reads_differing=bad_reads, flips=flips,
               status_or=int(np.bitwise_or.reduce(status)) if len(status) else 0,
               e2e_s=round(t1-t0, 4), oracle_s=round(t3-t2, 3), ms_decode=round(md, 4), ms_classify=round(mc, 3),
               gkmers_per_s_resident=round(nk/((md+mc)*1e-3)/1e9, 3),
               decode_GBps=round((len(prof)+2*nk)/(md*1e-3)/1e9, 1))
    print(json.dumps(res)); sys.stdout.flush()
    ctx.close()
    return res

if __name__ == "__main__":
    out = []
    out.append(run("tiny", seed=1, genome_len=20000, cov=10., het=0.005, len_mean=5000, len_sd=500))
    out.append(run("basic", seed=1, genome_len=100000, cov=30., het=0.005))
    out.append(run("repeat", seed=2, genome_len=200000, cov=40., het=0.01, repeat_frac=0.5, seg_dups=3))
    out.append(run("hicov", seed=6, genome_len=120000, cov=100., het=0.01, repeat_frac=0.3, len_mean=25000, len_sd=3000))
    out.append(run("noisy", seed=8, genome_len=100000, cov=35., het=0.001, err_indel_hp=0.003, err_sub=0.002, repeat_frac=0.7))
    out.append(run("big", seed=11, genome_len=2000000, cov=30., het=0.01, len_mean=20000, len_sd=2000))
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
