/*******************************************************************************************
 *  hostsim.cpp -- TEST-ONLY build of the device-side per-read logic for machines without a GPU.
 *
 *  The CUDA sources classpro_b200/csrc/cpg_*.cuh are written so that, with CPG_HOSTSIM defined,
 *  they compile as plain C++ with a warp width of 1.  This file wraps them behind a tiny C
 *  interface so that the CPU test-suite (`pytest -m "not gpu"`) can compare the device logic with
 *  the oracle read by read.  It is built by tests/ into tests/hostsim/_build/, is never linked
 *  into or loaded by libclasspro_b200.so, and is not a fallback: the product path has none.
 *******************************************************************************************/
#define CPG_HOSTSIM 1
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../classpro_b200/csrc/cpg_unrel.cuh"
#include "../../classpro_b200/csrc/cpg_decode.cuh"
#include "../../include/classpro_gpu.h"

struct HsWork
  { std::vector<uint32_t> mark; std::vector<double> perr; std::vector<cpg_eintvl> eint;
    std::vector<cpg_intvl> intvl, rint, wint; std::vector<uint16_t> bp;
    std::vector<uint8_t> af, ab, rpos, fixed; std::vector<int32_t> ord;
    void size(int P)
      { int MC = P/2+8;
        mark.assign(P+2,0); perr.assign((size_t)(P+2)*4,0.); eint.resize(P+2); intvl.resize(P+2);
        rint.resize(MC); wint.resize(MC); bp.assign(MC,0); af.assign(MC,0); ab.assign(MC,0);
        rpos.assign(MC,0); fixed.assign(P+2,0); ord.assign(P+2,0);
      }
  };

extern "C" {

/* classify one read; seq = rlen raw characters (seq_bits 8) ; returns status bits, fills cls[rlen],
 * and (optionally) the interval table */
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

  std::vector<uint8_t> packed;
  cpg_seq S;
  if (seq_bits == 2)
    { packed.assign((size_t)(rlen+3)/4,0);
      if (cpg_pack_seq(seq,rlen,packed.data())) return -1;
      S.p = packed.data(); S.bits = 2;
    }
  else { S.p = (const uint8_t *)seq; S.bits = 8; }

  static HsWork Wk; static int sized = 0;
  if (sized < plen) { Wk.size(plen+64); sized = plen+64; }
  cpg_wshared ws; RelShared sh;
  WCtx W; W.lane = 0; W.M = &dm; W.cthres = dm.cthres; W.ws = &ws; W.status = 0;
  ReadCtx R;
  R.prof = prof; R.plen = plen; R.rlen = rlen; R.seq = S; R.nslots = 0; R.N = R.M = 0;
  R.S.mark = Wk.mark.data(); R.S.perr = Wk.perr.data(); R.S.eint = Wk.eint.data();
  R.S.intvl = Wk.intvl.data(); R.S.rint = Wk.rint.data(); R.S.wint = Wk.wint.data();
  R.S.bp = Wk.bp.data(); R.S.asg_f = Wk.af.data(); R.S.asg_b = Wk.ab.data();
  R.S.rpos = Wk.rpos.data(); R.S.ord = Wk.ord.data(); R.S.fixed = Wk.fixed.data();
  int st = classify_read(R,W,&sh,(uint8_t *)cls);
  if (ivl_out) for (int i = 0; i < R.N; i++) ivl_out[i] = R.S.intvl[i];
  if (N_out) *N_out = R.N;
  if (M_out) *M_out = R.M;
  return st;
}

int hs_decode_profile(const uint8_t *src, int64_t len, uint16_t *out, int cap)
{ int offs[2*CPG_WARP];
  return decode_profile(src,len,out,cap,0,offs);
}

int hs_ctx(const char *seq, int rlen, int p, int right, int t)
{ cpg_seq S; S.p = (const uint8_t *)seq; S.bits = 8;
  return right ? cpg_rctx(S,rlen,p,t) : cpg_lctx(S,rlen,p,t);
}

int hs_sizeof_intvl(void) { return (int)sizeof(cpg_intvl); }

}
