/*******************************************************************************************
 *  fakedev.cpp -- TEST-ONLY stand-in for the device side of libclasspro_b200.so, so that the
 *  command-line programs (classpro_b200/host/classpro_main.c, cpg_prof2class.c) -- option
 *  parsing, FASTX reader, batching, packing pool, per-GPU workers, ordered writer, the byte
 *  contract of the records -- can be exercised by the CPU test-suite on machines without a GPU.
 *
 *  It implements the subset of include/classpro_gpu.h those programs call.  Every read goes
 *  through the DEVICE sources compiled for the host with a lane-group width of 1 (hostsim.cpp):
 *  decode_profile (counts + wall-candidate bit map), then classify_read -- the same per-read calls
 *  the kernels make.  Built by tests/hostsim/build.sh into tests/hostsim/_build/ only; it is not a
 *  fallback: the product library has none and fails without a CUDA device.
 *******************************************************************************************/
#ifndef CPG_HOSTSIM
#define CPG_HOSTSIM 1            /* -DCPG_HOSTSIM=32: every read on the 32-thread warp emulation (slow; used
                                    with -fsanitize=thread to look for missing group barriers) */
#endif
#include "hostsim.cpp"
#include <string.h>

struct cpg_ctx
  { cpg_model model; cpg_dmodel dm;
    std::vector<uint8_t> cls[2]; std::vector<int32_t> status[2], rlen[2]; int n[2]; int64_t bytes[2]; int busy[2];
    char err[256];
  };

static char g_fake_err[256] = "";

extern "C" {

int cpg_device_count(void)
{ const char *e = getenv("CPG_FAKE_DEVICES");         /* the tests choose how many "GPUs" there are */
  return e ? atoi(e) : 1;
}
const char *cpg_last_error(const cpg_ctx *ctx) { return ctx ? ctx->err : g_fake_err; }
const char *cpg_version(void) { return "classpro_b200 test stand-in (host build of the device sources)"; }
const char *cpg_status_string(int32_t st)
{ if (st == 0) return "ok";
  if (st & CPG_ST_BAD_PROFILE) return "profile length != read length - K + 1";
  if (st & CPG_ST_EINTVL_OVF)  return "# E-intvls >= plen";
  return "error";
}
void *cpg_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void  cpg_host_free(void *p) { free(p); }

int cpg_create(cpg_ctx **out, int device, const cpg_model *model, int64_t, int32_t)
{ if (device < 0 || device >= cpg_device_count()) { snprintf(g_fake_err,sizeof(g_fake_err),"device %d out of range",device); return CPG_EINVAL; }
  cpg_ctx *c = new cpg_ctx();
  c->model = *model;
  cpg_dmodel &dm = c->dm;
  const cpg_model *m = &c->model;
  dm.K = m->kmer; dm.read_len = m->read_len; dm.cmax = m->cmax;
  for (int t = 0; t < 3; t++) dm.lmax[t] = m->lmax[t];
  for (int s = 0; s < 4; s++) dm.cov[s] = m->cov[s];
  dm.dr_ratio = m->dr_ratio; dm.hc_erate = m->hc_erate;
  for (int t = 0; t < 3; t++) for (int l = 0; l < 21; l++) dm.pe[t][l] = m->pe[t][l];
  dm.cthres = m->cthres; dm.logfact = m->logfact;
  cpg_model_fill_logs(&dm,0,1);
  c->busy[0] = c->busy[1] = 0; c->err[0] = 0;
  *out = c;
  return CPG_OK;
}
void cpg_destroy(cpg_ctx *c) { delete c; }

/* one read: the per-read calls of k_decode and of the classification kernels.
   CPG_FAKE_NULL=1 skips them (class string of 'X's): then a run of the program times its host side
   alone -- parser, packing pool, queues, writer -- on inputs of any size. */
static int fake_read(cpg_ctx *c, const cpg_batch *b, int i, uint8_t *cls)
{ const int K = c->dm.K, rlen = b->rlen[i], cap = rlen-K+1;
  static const int null_dev = (getenv("CPG_FAKE_NULL") != 0);
  if (null_dev) { memset(cls,'N',(size_t)K-1); memset(cls+K-1,'X',(size_t)cap); return 0; }
#if CPG_HOSTSIM == 32
  { /* the warp emulation: 32 host threads per read, lane groups of CPG_FAKE_GROUP lanes (default 4) */
    static const int grp = getenv("CPG_FAKE_GROUP") ? atoi(getenv("CPG_FAKE_GROUP")) : 4;
    hs_set_group(grp);
    std::vector<uint16_t> cnt((size_t)cap+64,0);
    const int64_t po = b->prof_off[i];
    int n = hs_decode_profile(b->prof+po,b->prof_off[i+1]-po,cnt.data(),cap);
    if (n != cap) return CPG_ST_BAD_PROFILE;
    std::vector<char> asc((size_t)rlen+1,0);
    const uint8_t *sp = b->seq+b->seq_off[i];
    for (int j = 0; j < rlen; j++)
      asc[j] = (b->seq_bits == 2) ? "ACGT"[(sp[j >> 2] >> ((j & 3)*2)) & 3] : (char)sp[j];
    return hs_classify_read(&c->model,asc.data(),rlen,b->seq_bits,cnt.data(),cap,(char *)cls,NULL,NULL,NULL);
  }
#else
  static thread_local std::vector<uint16_t> cnt; static thread_local std::vector<uint32_t> cand;
  static thread_local std::vector<uint8_t> sq; static thread_local HsWork Wk; static thread_local int sized = 0;
  static thread_local unsigned tab[DC_SLOTS];
  cnt.assign((size_t)cap+64,0); cand.assign((size_t)cap/32+4,0xffffffffu);
  const int64_t po = b->prof_off[i];
  int n = decode_profile(b->prof+po,b->prof_off[i+1]-po,cnt.data(),cap,0,tab,cand.data(),c->dm.cov[ST_R]);
  if (n != cap) return CPG_ST_BAD_PROFILE;
  const size_t sb = (b->seq_bits == 2) ? (size_t)(rlen+3)/4 : (size_t)rlen;
  sq.assign(sb+128,0x5A);                                   /* padded: cpg_win reads aligned words around it */
  memcpy(sq.data()+64,b->seq+b->seq_off[i],sb);
  if (sized < cap) { Wk.size(cap+64); sized = cap+64; }
  Wk.set_caps(sized,0);
  static thread_local cpg_wshared ws; static thread_local RelShared sh[2];
  WCtx W; W.lane = 0; W.M = &c->dm; W.cthres = c->dm.cthres; W.ws = &ws; W.status = 0;
  W.glane = 0; W.gsize = 1; W.gbase = 0; W.gmask = 1u;
  ReadCtx R;
  R.prof = cnt.data(); R.plen = cap; R.rlen = rlen; R.seq.p = sq.data()+64; R.seq.bits = b->seq_bits;
  R.cand = cand.data(); R.nslots = 0; R.N = R.M = 0;
  HsWork &Kk = Wk;
  R.S.mark = Kk.mark.data(); R.S.slot = Kk.slot.data(); R.S.perr = Kk.perr.data(); R.S.eint = Kk.eint.data();
  R.S.intvl = Kk.intvl.data(); R.S.rint = Kk.rint.data(); R.S.wint = Kk.wint.data();
  R.S.bp = Kk.bp.data(); R.S.asg_f = Kk.af.data(); R.S.asg_b = Kk.ab.data();
  R.S.rpos = Kk.rpos.data(); R.S.ord = Kk.ord.data(); R.S.srt = Kk.srt.data(); R.S.MC = Kk.mc; R.S.upre = Kk.upre.data();
  R.S.capS = Kk.capS; R.S.capE = Kk.capE; R.S.capI = Kk.capI;
  R.S.tlog = Kk.tlog.data(); R.S.capT = Kk.capT; R.S.capC = Kk.capC; R.S.hdr = Kk.hdr.data(); R.S.big = Kk.big.data();
  R.S.key = Kk.key.data();
  R.hdr = 0; R.big = 0; R.ncand = 0; R.ntlog = 0; R.prune = 0;
  return classify_read(R,W,sh,cls);
#endif
}

int cpg_submit(cpg_ctx *c, int slot, const cpg_batch *b)
{ if (!c || !b || slot < 0 || slot > 1 || c->busy[slot]) { if (c) snprintf(c->err,sizeof(c->err),"cpg_submit: bad argument"); return CPG_EINVAL; }
  const int n = b->n_reads;
  int64_t tot = 0;
  for (int i = 0; i < n; i++)
    { if (b->rlen[i] < c->dm.K) { snprintf(c->err,sizeof(c->err),"read %d of the batch: rlen %d outside [K=%d,..]",i,b->rlen[i],c->dm.K); return CPG_EINVAL; }
      tot += b->rlen[i];
    }
  c->cls[slot].assign((size_t)tot+1,0); c->status[slot].assign((size_t)n+1,0); c->rlen[slot].assign(b->rlen,b->rlen+n);
  int64_t at = 0;
  for (int i = 0; i < n; i++) { c->status[slot][i] = fake_read(c,b,i,c->cls[slot].data()+at); at += b->rlen[i]; }
  c->n[slot] = n; c->bytes[slot] = tot; c->busy[slot] = 1;
  return CPG_OK;
}

int cpg_collect(cpg_ctx *c, int slot, cpg_result *res)
{ if (!c || !res || slot < 0 || slot > 1 || !c->busy[slot]) { if (c) snprintf(c->err,sizeof(c->err),"cpg_collect: bad argument"); return CPG_EINVAL; }
  c->busy[slot] = 0;
  memcpy(res->cls,c->cls[slot].data(),(size_t)c->bytes[slot]);
  int bad = 0;
  for (int i = 0; i < c->n[slot]; i++)
    { if (res->status) res->status[i] = c->status[slot][i];
      if (c->status[slot][i] & (1|2|4|8|64))
        { if (!bad) snprintf(c->err,sizeof(c->err),"read %d of the batch: %s",i,cpg_status_string(c->status[slot][i]));
          bad = 1;
        }
    }
  return bad ? CPG_EREAD : CPG_OK;
}

/* compact results (include/classpro_gpu.h): here simply the run-length form of the class strings the host
   build of the device sources produced (equal neighbours merge: the expansion is the same) */
int cpg_set_result_mode(cpg_ctx *c, int mode) { (void)mode; return c ? CPG_OK : CPG_EINVAL; }
int64_t cpg_intervals_bound(cpg_ctx *c, int slot) { return (c && slot >= 0 && slot <= 1) ? c->bytes[slot]+16 : 0; }
int cpg_collect_intervals(cpg_ctx *c, int slot, cpg_result_ivl *res)
{ if (!c || !res || slot < 0 || slot > 1 || !c->busy[slot]) { if (c) snprintf(c->err,sizeof(c->err),"cpg_collect_intervals: bad argument"); return CPG_EINVAL; }
  if (res->ivl_cap < c->bytes[slot]) { snprintf(c->err,sizeof(c->err),"cpg_result_ivl.ivl too small"); return CPG_EINVAL; }
  c->busy[slot] = 0;
  const int K = c->dm.K;
  int64_t at = 0, used = 0; int bad = 0;
  for (int i = 0; i < c->n[slot]; i++)
    { const int rlen = c->rlen[slot][i], plen = rlen-K+1;
      const uint8_t *s = c->cls[slot].data()+at+K-1;
      res->ivl_at[i] = used;
      int n = 0;
      for (int j = 0; j < plen; )
        { int e = j+1;
          while (e < plen && s[e] == s[j]) e++;
          const unsigned code = (s[j] == 'E') ? 0u : (s[j] == 'R') ? 1u : (s[j] == 'H') ? 2u : (s[j] == 'D') ? 3u : 4u;
          res->ivl[used++] = ((unsigned)e << 3) | code; n++;
          j = e;
        }
      res->ivl_n[i] = n;
      at += rlen;
      if (res->status) res->status[i] = c->status[slot][i];
      if (c->status[slot][i] & (1|2|4|8|64))
        { if (!bad) snprintf(c->err,sizeof(c->err),"read %d of the batch: %s",i,cpg_status_string(c->status[slot][i]));
          bad = 1;
        }
    }
  res->ivl_used = used;
  return bad ? CPG_EREAD : CPG_OK;
}

int cpg_classify(cpg_ctx *c, const cpg_batch *b, cpg_result *r)
{ int rc = cpg_submit(c,0,b);
  return rc ? rc : cpg_collect(c,0,r);
}

int cpg_prof2class(cpg_ctx *c, int32_t n, const uint8_t *prof, const int64_t *prof_off, const int32_t *rlen,
                   uint8_t *cls, int32_t *status)
{ const int K = c->dm.K;
  unsigned tab[DC_SLOTS];
  std::vector<uint16_t> cnt;
  int64_t at = 0; int bad = 0;
  for (int i = 0; i < n; i++)
    { const int cap = rlen[i] >= K ? rlen[i]-K+1 : 0;
      cnt.assign((size_t)cap+64,0);
      int m = decode_profile(prof+prof_off[i],prof_off[i+1]-prof_off[i],cnt.data(),cap,0,tab,0,0);
      const int st = (m == cap) ? 0 : CPG_ST_BAD_PROFILE;
      if (status) status[i] = st;
      if (st && !bad) { snprintf(c->err,sizeof(c->err),"read %d of the batch: %s",i,cpg_status_string(st)); bad = 1; }
      for (int j = 0; j < rlen[i]; j++)
        { char ch = 'N';
          if (j >= K-1) { unsigned v = cnt[j-(K-1)]; ch = v == 0 ? 'E' : v == 1 ? 'H' : v == 2 ? 'D' : 'R'; }
          cls[at+j] = (uint8_t)ch;
        }
      at += rlen[i];
    }
  return bad ? CPG_EREAD : CPG_OK;
}

}
