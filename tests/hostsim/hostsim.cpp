/*******************************************************************************************
 *  hostsim.cpp -- TEST-ONLY builds of the device-side per-read logic for machines without a GPU.
 *
 *  The CUDA sources classpro_b200/csrc/cpg_*.cuh compile as plain C++ when CPG_HOSTSIM is defined:
 *    CPG_HOSTSIM=1   warp width 1: one host thread runs the per-read logic serially
 *                    (libhostsim.so, fast, used for dataset-level parity with the oracle);
 *    CPG_HOSTSIM=32  warp width 32: 32 host threads play the lanes of one warp and meet at a
 *                    barrier for every __syncwarp/ballot/shuffle/reduction (libhostsim32.so).
 *                    Lanes run truly asynchronously between rendezvous, so a collective executed
 *                    by only some lanes deadlocks (test timeout) and a missing __syncwarp shows up
 *                    as wrong or unstable output -- stricter than the hardware.
 *  Built by tests/ into tests/hostsim/_build/, never linked into or loaded by
 *  libclasspro_b200.so; not a fallback: the product path has none.
 *******************************************************************************************/
#ifndef CPG_HOSTSIM
#define CPG_HOSTSIM 1
#endif
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#if CPG_HOSTSIM == 32
#include <pthread.h>
#endif
#include "../../classpro_b200/csrc/cpg_unrel.cuh"
#include "../../classpro_b200/csrc/cpg_decode.cuh"
#include "../../include/classpro_gpu.h"

#if CPG_HOSTSIM == 32
static pthread_barrier_t g_bar;
static unsigned g_slot[32];
static thread_local int t_lane = 0;
/* one barrier per aligned power-of-two lane group: index = log2(size), base/size */
static pthread_barrier_t g_gbar[6][32];
static pthread_barrier_t *bar_of(unsigned mask)
{ if (mask == 0xffffffffu) return &g_bar;
  int size = __builtin_popcount(mask), base = __builtin_ctz(mask);
  return &g_gbar[__builtin_ctz((unsigned)size)][base/size];
}
void cpg_sim_barrier(void) { pthread_barrier_wait(&g_bar); }
unsigned cpg_sim_ballot(int pred)
{ g_slot[t_lane] = pred ? 1u : 0u;
  pthread_barrier_wait(&g_bar);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) m |= g_slot[l] << l;
  pthread_barrier_wait(&g_bar);
  return m;
}
unsigned cpg_sim_shfl(unsigned v, int src)
{ g_slot[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  unsigned r = g_slot[src & 31];
  pthread_barrier_wait(&g_bar);
  return r;
}
unsigned cpg_sim_shfl_up(unsigned v, int d)
{ g_slot[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  unsigned r = (t_lane >= d) ? g_slot[t_lane-d] : v;
  pthread_barrier_wait(&g_bar);
  return r;
}
void cpg_sim_group_barrier(unsigned mask) { pthread_barrier_wait(bar_of(mask)); }
unsigned cpg_sim_gballot(unsigned mask, int pred)
{ pthread_barrier_t *b = bar_of(mask);
  g_slot[t_lane] = pred ? 1u : 0u;
  pthread_barrier_wait(b);
  unsigned m = 0;
  for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) m |= g_slot[l] << l;
  pthread_barrier_wait(b);
  return m;
}
int cpg_sim_gsum(unsigned mask, int v)
{ pthread_barrier_t *b = bar_of(mask);
  g_slot[t_lane] = (unsigned)v;
  pthread_barrier_wait(b);
  int s = 0;
  for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) s += (int)g_slot[l];
  pthread_barrier_wait(b);
  return s;
}
int cpg_sim_sum(int v) { return cpg_sim_gsum(0xffffffffu,v); }
unsigned cpg_sim_gshfl(unsigned mask, unsigned v, int src)
{ pthread_barrier_t *b = bar_of(mask);
  g_slot[t_lane] = v;
  pthread_barrier_wait(b);
  unsigned r = g_slot[src & 31];
  pthread_barrier_wait(b);
  return r;
}
#endif

static int g_group = CPG_WARP;        /* lanes per read (hs_set_group) */
static int g_small_caps = 0;          /* > 0: first attempt with interval tables of this many entries (hs_set_small_caps) */
static int g_retries = 0;

struct HsWork
  { std::vector<uint8_t> mark; std::vector<uint16_t> slot; std::vector<uint32_t> cand, key; std::vector<double> perr; std::vector<cpg_eintvl> eint;
    std::vector<int32_t> tlog; std::vector<cpg_chdr> hdr; std::vector<cpg_cbig> big; int capT = 0, capC = 0;
    std::vector<cpg_intvl> intvl, rint, wint; std::vector<uint16_t> bp;
    std::vector<uint8_t> af, ab, rpos; std::vector<int32_t> ord, srt; std::vector<cpg_upre> upre; int mc = 0;
    int capS = 0, capE = 0, capI = 0;
    /* the interval tables are allocated with exactly `cap` entries (heap, so that an access past a
       compact table is visible to a memory checker) */
    void set_caps(int P, int small)
      { capS = capE = capI = P+2;
        if (small > 0) { capI = capE = small; capS = 2*small; }
        capT = 3*capS; capC = capI;
        perr.assign((size_t)capS*4,0.); eint.assign(capE,cpg_eintvl()); intvl.assign(capI,cpg_intvl());
        srt.assign(capI,0); ord.assign(capI,0); key.assign(capI,0); upre.assign(capI,cpg_upre());
        tlog.assign(capT,0); hdr.assign(capC,cpg_chdr()); big.assign(capC,cpg_cbig());
      }
    void size(int P)
      { int MC = P/2+8;
        mark.assign(P+2+32,0); slot.assign(P+2,0xffff);        /* the flag bytes are zero between reads (checked below) */ cand.assign(P/32+2,0); perr.assign((size_t)(P+2)*4,0.); eint.resize(P+2); intvl.resize(P+2);
        rint.resize(MC); wint.resize(2*MC); bp.assign(2*MC,0); af.assign(MC,0); ab.assign(MC,0);
        rpos.assign(2*MC,0); mc = MC; srt.assign(P+2,0); ord.assign(P+2,0);
      }
  };

struct LaneJob
  { int lane, glane, gsize, gbase; unsigned gmask; const cpg_dmodel *dm; cpg_wshared *ws; RelShared *sh; ReadCtx R; uint8_t *cls; int status;
    int N, M;
    /* decode job */
    const uint8_t *src; int64_t len; uint16_t *out; int cap; unsigned *offs; int n; uint32_t *cand; int rcov;
  };

static void *lane_classify(void *arg)
{ LaneJob *J = (LaneJob *)arg;
#if CPG_HOSTSIM == 32
  t_lane = J->lane;
#endif
  WCtx W; W.lane = J->lane; W.M = J->dm; W.cthres = J->dm->cthres; W.ws = J->ws; W.status = 0;
  W.glane = J->glane; W.gsize = J->gsize; W.gbase = J->gbase; W.gmask = J->gmask;
  ReadCtx R = J->R;
  J->status = classify_read(R,W,J->sh,J->cls);
  J->N = R.N; J->M = R.M;
  return NULL;
}

static void *lane_decode(void *arg)
{ LaneJob *J = (LaneJob *)arg;
#if CPG_HOSTSIM == 32
  t_lane = J->lane;
#endif
  J->n = decode_profile(J->src,J->len,J->out,J->cap,J->lane,J->offs,J->cand,J->rcov);
  return NULL;
}

static void run_lanes(void *(*fn)(void *), LaneJob *jobs)
{
#if CPG_HOSTSIM == 32
  pthread_t th[32];
  pthread_barrier_init(&g_bar,NULL,32);
  for (int k = 0; k < 6; k++) for (int g = 0; g < (32 >> k); g++) pthread_barrier_init(&g_gbar[k][g],NULL,1u << k);
  for (int l = 0; l < 32; l++) pthread_create(&th[l],NULL,fn,&jobs[l]);
  for (int l = 0; l < 32; l++) pthread_join(th[l],NULL);
  pthread_barrier_destroy(&g_bar);
  for (int k = 0; k < 6; k++) for (int g = 0; g < (32 >> k); g++) pthread_barrier_destroy(&g_gbar[k][g]);
#else
  fn(&jobs[0]);
#endif
}

extern "C" {

int hs_warp_width(void) { return CPG_WARP; }

/* classify one read; seq = rlen raw characters; returns the OR of the lanes' status bits, fills
 * cls[rlen] and (optionally) the interval table */
int hs_classify_read(const cpg_model *m, const char *seq, int rlen, int seq_bits,
                     const uint16_t *prof, int plen, char *cls,
                     cpg_intvl *ivl_out, int *N_out, int *M_out)
{ cpg_dmodel dm;
  dm.K = m->kmer; dm.read_len = m->read_len; dm.cmax = m->cmax;
  for (int t = 0; t < 3; t++) dm.lmax[t] = m->lmax[t];
  for (int s = 0; s < 4; s++) dm.cov[s] = m->cov[s];
  dm.dr_ratio = m->dr_ratio; dm.hc_erate = m->hc_erate;
  for (int t = 0; t < 3; t++) for (int l = 0; l < 21; l++) dm.pe[t][l] = m->pe[t][l];
  dm.cthres = m->cthres; dm.logfact = m->logfact;
  cpg_model_fill_logs(&dm,0,1);

  std::vector<uint8_t> packed;
  cpg_seq S;
  if (seq_bits == 2)
    { /* 16 bytes of padding on both sides: cpg_win reads aligned words around the sequence */
      packed.assign((size_t)(rlen+3)/4+64,0xA5);
      if (cpg_pack_seq(seq,rlen,packed.data()+16)) return -1;
      S.p = packed.data()+16; S.bits = 2;
    }
  else { S.p = (const uint8_t *)seq; S.bits = 8; }

  /* every lane group of the simulated warp classifies the same read with its own scratch; the
     outputs of the groups must agree */
  const int G = g_group, NG = CPG_WARP/G;
  static HsWork Wk[32]; static int sized[32];
  static cpg_wshared ws[32]; static RelShared sh[32][2];
  std::vector<std::vector<char> > cls_g(NG);
  LaneJob jobs[CPG_WARP];
  int st = 0;
  for (int attempt = 0; attempt < 2; attempt++)      /* 0: as the phase kernels (compact tables if asked for, pruned records); 1: as the retry launch */
  {
  for (int l = 0; l < CPG_WARP; l++)
    { LaneJob &J = jobs[l];
      const int g = l/G;
      if (sized[g] < plen) { Wk[g].size(plen+64); sized[g] = plen+64; }
      if (l%G == 0) Wk[g].set_caps(sized[g],attempt == 0 ? g_small_caps : 0);
      if (g > 0 && (int)cls_g[g].size() < rlen+1) cls_g[g].assign((size_t)rlen+1,0);
      J.lane = l; J.glane = l%G; J.gsize = G; J.gbase = g*G;
      J.gmask = ((G >= 32) ? 0xffffffffu : ((1u << G)-1u)) << J.gbase;
      J.dm = &dm; J.ws = &ws[g]; J.sh = sh[g]; J.cls = (g == 0) ? (uint8_t *)cls : (uint8_t *)cls_g[g].data(); J.status = 0;
      ReadCtx &R = J.R;
      HsWork &K = Wk[g];
      R.prof = prof; R.plen = plen; R.rlen = rlen; R.seq = S; R.nslots = 0; R.N = R.M = 0;
      /* candidate bit map straight from the definition (src/wall.c:594-608); the decoder's own bit
         map is checked against the same definition by the decode tests */
      if (l%G == 0)
        { std::fill(K.cand.begin(),K.cand.end(),0u);
          for (int i = 1; i < plen; i++)
            if (dc_is_cand(prof[i-1],prof[i],dm.cov[ST_R])) K.cand[i >> 5] |= 1u << (i & 31);
        }
      R.cand = K.cand.data();
      R.S.mark = K.mark.data(); R.S.slot = K.slot.data(); R.S.perr = K.perr.data(); R.S.eint = K.eint.data();
      R.S.intvl = K.intvl.data(); R.S.rint = K.rint.data(); R.S.wint = K.wint.data();
      R.S.bp = K.bp.data(); R.S.asg_f = K.af.data(); R.S.asg_b = K.ab.data();
      R.S.rpos = K.rpos.data(); R.S.ord = K.ord.data(); R.S.srt = K.srt.data(); R.S.MC = K.mc; R.S.upre = K.upre.data();
      R.S.capS = K.capS; R.S.capE = K.capE; R.S.capI = K.capI;
      R.S.tlog = K.tlog.data(); R.S.capT = K.capT; R.S.capC = K.capC; R.S.hdr = K.hdr.data(); R.S.big = K.big.data();
      R.S.key = K.key.data();
      R.hdr = 0; R.big = 0; R.ncand = 0; R.ntlog = 0; R.prune = (attempt == 0);   /* first attempt as k_wall_a records, second as the retry launch */
    }
  run_lanes(lane_classify,jobs);
  st = 0;
  for (int l = 0; l < CPG_WARP; l++) st |= jobs[l].status;
  if (attempt == 0 && (st & CPG_ST_RETRY)) { g_retries++; continue; }      /* what the retry launch does on the device */
  break;
  }
  st = 0;
  for (int l = 0; l < CPG_WARP; l++)
    { st |= jobs[l].status;
      if (jobs[l].N != jobs[0].N || jobs[l].M != jobs[0].M) st |= 1<<30;      /* lanes disagree */
    }
  for (int g = 1; g < NG; g++) if (memcmp(cls,cls_g[g].data(),(size_t)rlen) != 0) st |= 1<<29;   /* groups disagree */
  for (int g = 0; g < NG; g++)                                   /* the wall stage must leave its flag bytes clean */
    for (size_t q = 0; q < Wk[g].mark.size(); q++) if (Wk[g].mark[q]) { st |= 1<<28; break; }
  if (ivl_out) for (int i = 0; i < jobs[0].N; i++) ivl_out[i] = Wk[0].intvl[i];
  if (N_out) *N_out = jobs[0].N;
  if (M_out) *M_out = jobs[0].M;
  return st;
}

/* first attempt of the next hs_classify_read calls with interval tables of n entries (0: full
   size only); hs_retries() counts the reads that needed the second, full-size attempt */
int hs_set_small_caps(int n) { g_small_caps = n; g_retries = 0; return 0; }
int hs_retries(void) { return g_retries; }

/* lanes per read for the next hs_classify_read calls (a power of two <= the warp width) */
int hs_set_group(int g)
{ if (g < 1 || g > CPG_WARP || (g & (g-1))) return -1;
  g_group = g;
  return 0;
}

/* cand: NULL or ceil(cap/32) words for the wall-candidate bit map (rcov = repeat threshold) */
int hs_decode_profile_cand(const uint8_t *src, int64_t len, uint16_t *out, int cap, uint32_t *cand, int rcov)
{ static unsigned offs[DC_SLOTS];
  LaneJob jobs[CPG_WARP];
  for (int l = 0; l < CPG_WARP; l++)
    { jobs[l].lane = l; jobs[l].src = src; jobs[l].len = len; jobs[l].out = out; jobs[l].cap = cap;
      jobs[l].offs = offs; jobs[l].n = 0; jobs[l].cand = cand; jobs[l].rcov = rcov;
    }
  run_lanes(lane_decode,jobs);
  for (int l = 1; l < CPG_WARP; l++) if (jobs[l].n != jobs[0].n) return -1000000;
  return jobs[0].n;
}

int hs_decode_profile(const uint8_t *src, int64_t len, uint16_t *out, int cap)
{ return hs_decode_profile_cand(src,len,out,cap,NULL,0); }

int hs_ctx(const char *seq, int rlen, int p, int right, int t)
{ cpg_seq S; S.p = (const uint8_t *)seq; S.bits = 8;
  return right ? cpg_rctx(S,rlen,p,t) : cpg_lctx(S,rlen,p,t);
}

/* the same query on the 2-bit packed form of the sequence (pad = byte offset of the packed
   sequence inside its buffer, to exercise every alignment of the window loads) */
int hs_ctx2(const char *seq, int rlen, int p, int right, int t, int pad)
{ std::vector<uint8_t> packed((size_t)(rlen+3)/4+64,0x5A);
  if (cpg_pack_seq(seq,rlen,packed.data()+16+pad)) return -1;
  cpg_seq S; S.p = packed.data()+16+pad; S.bits = 2;
  /* both forms must agree: the single-type function and the three-at-once span form */
  int one = right ? cpg_rctx(S,rlen,p,t) : cpg_lctx(S,rlen,p,t);
  int all[3];
  if (right) cpg_rctx3(S,rlen,p,all); else cpg_lctx3(S,rlen,p,all);
  if (all[t] != one) return -1000-all[t];
  return one;
}

int hs_sizeof_intvl(void) { return (int)sizeof(cpg_intvl); }

/* the two shortcuts of cpg_math.cuh against the oracle's plain loops (tests/test_device_logic_hostsim.py) */
double hs_binom_tail(const cpg_model *m, int k, int n, double p, int *bad)
{ cpg_rate r; r.p = p; r.lp = log(p); r.l1mp = log(1-p);
  *bad = 0;
  return cpg_binom_tail_lane(m->logfact,k,n,r,bad);
}
double hs_lp_trans_thr(int read_len, int b, int e, int cb, int ce, int cov, double thres, int use_bound)
{ cpg_dmodel dm; memset(&dm,0,sizeof(dm)); dm.read_len = read_len;
  WCtx W; memset(&W,0,sizeof(W)); W.M = &dm; W.gsize = 1;
  return use_bound ? cpg_lp_trans_thr(W,b,e,cb,ce,(uint16_t)cov,thres) : cpg_lp_trans(W,b,e,cb,ce,(uint16_t)cov);
}

}
