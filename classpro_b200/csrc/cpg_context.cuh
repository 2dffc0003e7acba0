/*******************************************************************************************
 *  cpg_context.cuh -- sequence context on demand.
 *
 *  The reference materialises six run-length bytes per base with a serial sweep and back-fill
 *  (src/context.c:8-108) but reads them only at wall candidates and interval ends (~1-2 % of the
 *  positions).  Here each value is evaluated where it is needed, from the packed sequence, with a
 *  closed form that equals the reference's arrays at every base as long as no run reaches the
 *  127 cap (tests/test_context.py proves this exhaustively for all short sequences and on random
 *  low-complexity sequences; beyond the cap the reference reads cells it never wrote):
 *
 *    L_HP(p) = length of the homopolymer run ending at p                       (lctx[p][HP])
 *    L_DS(p) = 0 if p == 0 or s[p] == s[p-1], else number of consecutive copies of the
 *              dinucleotide (s[p-1],s[p]) ending at p                          (lctx[p][DS])
 *    L_TS(p) = 0 if p < 2 or s[p-2] == s[p-1] == s[p], else number of consecutive copies of the
 *              trinucleotide ending at p                                       (lctx[p][TS])
 *    R_*(p)  = the mirror images, counted from p to the right                  (rctx[p][*])
 *  all capped at 127.  ctx[DROP][i] = lctx[i+K-2], ctx[GAIN][i] = rctx[i] (src/ClassPro.c:138-142).
 *******************************************************************************************/
#ifndef CPG_CONTEXT_CUH
#define CPG_CONTEXT_CUH
#include "cpg_common.h"

/* the view is taken by value: it lives in registers for the duration of a query */
CPG_DEV int cpg_base(const cpg_seq S, int i)
{ return (S.bits == 8) ? (int)S.p[i] : (int)((S.p[i >> 2] >> ((i & 3)*2)) & 3); }

CPG_DEV int cpg_cap127(int x) { return x > 127 ? 127 : x; }

CPG_DEV_HELPER int cpg_lctx(const cpg_seq S, int rlen, int p, int t)
{ (void)rlen;
  if (t == CT_HP)
    { int c = cpg_base(S,p), n = 1;
      CPG_LOOP while (n < 127 && p-n >= 0 && cpg_base(S,p-n) == c) n++;
      return n;
    }
  if (t == CT_DS)
    { if (p == 0) return 0;
      int a = cpg_base(S,p-1), b = cpg_base(S,p);
      if (a == b) return 0;
      int u = 1, q = p;
      CPG_LOOP while (u < 127 && q >= 3 && cpg_base(S,q-3) == a && cpg_base(S,q-2) == b) { u++; q -= 2; }
      return u;
    }
  if (p < 2) return 0;
  int a = cpg_base(S,p-2), b = cpg_base(S,p-1), c = cpg_base(S,p);
  if (a == b && b == c) return 0;
  int u = 1, q = p;
  CPG_LOOP while (u < 127 && q >= 5 && cpg_base(S,q-5) == a && cpg_base(S,q-4) == b && cpg_base(S,q-3) == c)
    { u++; q -= 3; }
  return u;
}

CPG_DEV_HELPER int cpg_rctx(const cpg_seq S, int rlen, int p, int t)
{ if (t == CT_HP)
    { int c = cpg_base(S,p), n = 1;
      CPG_LOOP while (n < 127 && p+n < rlen && cpg_base(S,p+n) == c) n++;
      return n;
    }
  if (t == CT_DS)
    { if (p >= rlen-1) return 0;
      int a = cpg_base(S,p), b = cpg_base(S,p+1);
      if (a == b) return 0;
      int u = 1, q = p;
      CPG_LOOP while (u < 127 && q+3 <= rlen-1 && cpg_base(S,q+2) == a && cpg_base(S,q+3) == b) { u++; q += 2; }
      return u;
    }
  if (p > rlen-3) return 0;
  int a = cpg_base(S,p), b = cpg_base(S,p+1), c = cpg_base(S,p+2);
  if (a == b && b == c) return 0;
  int u = 1, q = p;
  CPG_LOOP while (u < 127 && q+5 <= rlen-1 && cpg_base(S,q+3) == a && cpg_base(S,q+4) == b && cpg_base(S,q+5) == c)
    { u++; q += 3; }
  return u;
}

/* ctx[wtype][i][t] of the reference, i a profile position */
CPG_DEV int cpg_ctx_at(const cpg_seq S, int rlen, int K, int wtype, int i, int t)
{ return (wtype == WT_DROP) ? cpg_lctx(S,rlen,i+K-2,t) : cpg_rctx(S,rlen,i,t); }

#endif
