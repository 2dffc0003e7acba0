/*******************************************************************************************
 *  cpg_context.cuh -- sequence context on demand.
 *
 *  The reference materialises six run-length bytes per base with a serial sweep and back-fill
 *  (src/context.c:8-108) but reads them only at wall candidates and interval ends (~1-2 % of the
 *  positions).  Here each value is evaluated where it is needed, from the packed sequence, with a
 *  closed form that equals the reference's arrays at every base (tests/test_device_logic_hostsim.py:
 *  exhaustively for all short sequences, on random low-complexity sequences, and against the
 *  reference sweep on runs longer than the 127 cap, where the homopolymer right context has a
 *  back-fill quirk of its own, cpg_rctx_hp_ref below):
 *
 *    L_HP(p) = length of the homopolymer run ending at p                       (lctx[p][HP])
 *    L_DS(p) = 0 if p == 0 or s[p] == s[p-1], else number of consecutive copies of the
 *              dinucleotide (s[p-1],s[p]) ending at p                          (lctx[p][DS])
 *    L_TS(p) = 0 if p < 2 or s[p-2] == s[p-1] == s[p], else number of consecutive copies of the
 *              trinucleotide ending at p                                       (lctx[p][TS])
 *    R_*(p)  = the mirror images, counted from p to the right                  (rctx[p][*])
 *  all capped at 127.  ctx[DROP][i] = lctx[i+K-2], ctx[GAIN][i] = rctx[i] (src/ClassPro.c:138-142).
 *  Raw-byte sequences (reads with characters outside ACGT) walk the bases one by one; 2-bit packed
 *  sequences compare 32 bases at a time (see below).
 *******************************************************************************************/
#ifndef CPG_CONTEXT_CUH
#define CPG_CONTEXT_CUH
#include "cpg_common.h"

/* the view is taken by value: it lives in registers for the duration of a query */
CPG_DEV int cpg_base(const cpg_seq S, int i)
{ return (S.bits == 8) ? (int)S.p[i] : (int)((S.p[i >> 2] >> ((i & 3)*2)) & 3); }

CPG_DEV int cpg_cap127(int x) { return x > 127 ? 127 : x; }

/* rctx[p][HP] as the reference's back-fill leaves it in a homopolymer run LONGER than the cap
 * (src/context.c:24-26,60-61).  When a run [s,t] of L bases ends, the fill starts lctx[t] = min(L,127)
 * positions before its end and copies the mirrored -- and capped -- left run lengths:
 *     rctx[j] = lctx[2t+1-min(L,127)-j]        for j = t+1-min(L,127) .. t
 * For L <= 127 that is the right run length t-j+1.  For L > 127 it is min(L-126+(t-j),127) over the last
 * 127 bases of the run (e.g. 127, not 1, at the last base of a run of 253 or more), and the bases before
 * them are never written: the reference then reads whatever an earlier read left there.  The first is part
 * of the contract and is reproduced; the second is undefined there and is DEFINED here as the cap.
 * (Dinucleotide and trinucleotide runs are filled over their whole length by a walk, src/context.c:36-41,
 * 55-60: their capped closed form is what the reference computes, whatever their length.)
 *   r128: min(right run length at p, 128), p included;  l254: min(left run length at p, 254), p included */
CPG_DEV int cpg_rctx_hp_ref(int r128, int l254)
{ if (r128 >= 128) return 127;
  if (l254+r128-1 <= 127) return r128;
  return cpg_cap127(l254+2*r128-128);
}

CPG_DEV_HELPER int cpg_lctx_raw(const cpg_seq S, int rlen, int p, int t)
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

CPG_DEV_HELPER int cpg_rctx_raw(const cpg_seq S, int rlen, int p, int t)
{ if (t == CT_HP)
    { int c = cpg_base(S,p), n = 1, l = 1;
      CPG_LOOP while (n < 128 && p+n < rlen && cpg_base(S,p+n) == c) n++;
      CPG_LOOP while (l < 254 && p-l >= 0 && cpg_base(S,p-l) == c) l++;
      return cpg_rctx_hp_ref(n,l);
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

/* ---- 2-bit packed sequences: 32 bases per 64-bit window, runs by XOR + count-zeros ----
 * With Zb(p,d) = number of consecutive j = p, p-1, ... (j-d >= 0) with s[j] == s[j-d] and Zf(p,d)
 * its mirror image (j = p, p+1, ..., j+d <= rlen-1, s[j] == s[j+d]) the closed forms above are
 *   L_HP = 1+Zb(p,1)   L_DS = 1+Zb(p,2)/2   L_TS = 1+Zb(p,3)/3   (R_* with Zf), capped at 127:
 * a further copy of the unit ends at p exactly when d more positions agree with the base d to
 * their left. */
#ifdef CPG_HOSTSIM
CPG_DEV uint32_t cpg_funnel_r(uint32_t lo, uint32_t hi, unsigned sh)
{ return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31)); }
CPG_DEV int cpg_ctz64(uint64_t x) { return __builtin_ctzll(x); }
CPG_DEV int cpg_clz64(uint64_t x) { return __builtin_clzll(x); }
#else
CPG_DEV uint32_t cpg_funnel_r(uint32_t lo, uint32_t hi, unsigned sh) { return __funnelshift_r(lo,hi,sh); }
CPG_DEV int cpg_ctz64(uint64_t x) { return __ffsll((long long)x)-1; }
CPG_DEV int cpg_clz64(uint64_t x) { return __clzll((long long)x); }
#endif

/* bases s .. s+31 (s >= 0), base s in bits 0..1.  Reads the three aligned words at and after the
   byte of base s: up to 3 bytes before it and 11 after it; s can be up to 32 bases past the end of
   the read (cpg_rctx3), so the staging buffers are padded by 64 bytes. */
CPG_DEV uint64_t cpg_win(const uint8_t *p, int s)
{ const size_t ad = (size_t)(p+(s >> 2));
  const uint32_t *w = reinterpret_cast<const uint32_t *>(ad & ~(size_t)3);
  const unsigned bo = (unsigned)(ad & 3)*8u+(unsigned)(s & 3)*2u;          /* 0..30 */
  const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  return (((uint64_t)cpg_funnel_r(w1,w2,bo)) << 32) | cpg_funnel_r(w0,w1,bo);
}
/* bases q-31 .. q, base q in the top two bits; bases below 0 read as zeros */
CPG_DEV uint64_t cpg_win_end(const uint8_t *p, int q)
{ const int s = q-31;
  if (s >= 0) return cpg_win(p,s);
  if (s <= -32) return 0;
  return cpg_win(p,0) << (2*(-s));
}

CPG_DEV_HELPER int cpg_zb(const uint8_t *sq, int p, int d, int zmax)
{ int lim = p-d+1;
  if (lim > zmax) lim = zmax;
  int z = 0;
  CPG_LOOP while (z < lim)
    { const uint64_t e = cpg_win_end(sq,p-z) ^ cpg_win_end(sq,p-z-d);
      const int c = e ? (cpg_clz64(e) >> 1) : 32;
      z += c;
      if (c < 32) break;
    }
  return z < lim ? z : (lim > 0 ? lim : 0);
}

CPG_DEV_HELPER int cpg_zf(const uint8_t *sq, int rlen, int p, int d, int zmax)
{ int lim = rlen-d-p;
  if (lim > zmax) lim = zmax;
  int z = 0;
  CPG_LOOP while (z < lim)
    { const uint64_t e = cpg_win(sq,p+z) ^ cpg_win(sq,p+z+d);
      const int c = e ? (cpg_ctz64(e) >> 1) : 32;
      z += c;
      if (c < 32) break;
    }
  return z < lim ? z : (lim > 0 ? lim : 0);
}

CPG_DEV_HELPER int cpg_lctx(const cpg_seq S, int rlen, int p, int t)
{ if (S.bits == 8) return cpg_lctx_raw(S,rlen,p,t);
  if (t == CT_HP) return 1+cpg_zb(S.p,p,1,126);
  if (t == CT_DS)
    { if (p == 0 || cpg_base(S,p-1) == cpg_base(S,p)) return 0;
      return 1+(cpg_zb(S.p,p,2,252) >> 1);
    }
  if (p < 2) return 0;
  { const int a = cpg_base(S,p-2), b = cpg_base(S,p-1), c = cpg_base(S,p);
    if (a == b && b == c) return 0;
  }
  return 1+cpg_zb(S.p,p,3,378)/3;
}

CPG_DEV_HELPER int cpg_rctx(const cpg_seq S, int rlen, int p, int t)
{ if (S.bits == 8) return cpg_rctx_raw(S,rlen,p,t);
  if (t == CT_HP)
    { const int r = 1+cpg_zf(S.p,rlen,p,1,127);
      if (p == 0 || cpg_base(S,p-1) != cpg_base(S,p)) return cpg_rctx_hp_ref(r,1);
      return cpg_rctx_hp_ref(r,1+cpg_zb(S.p,p,1,253));
    }
  if (t == CT_DS)
    { if (p >= rlen-1 || cpg_base(S,p) == cpg_base(S,p+1)) return 0;
      return 1+(cpg_zf(S.p,rlen,p,2,252) >> 1);
    }
  if (p > rlen-3) return 0;
  { const int a = cpg_base(S,p), b = cpg_base(S,p+1), c = cpg_base(S,p+2);
    if (a == b && b == c) return 0;
  }
  return 1+cpg_zf(S.p,rlen,p,3,378)/3;
}

/* All three run lengths at one base from ONE 128-bit span of the packed sequence (the six
 * windows of the single-type functions are shifts of it).  A run that fills its 32-base window,
 * or a raw-byte sequence, falls back on the single-type function. */
CPG_DEV_HELPER void cpg_lctx3(const cpg_seq S, int rlen, int p, int out[3])
{ if (S.bits != 2) { for (int t = 0; t < 3; t++) out[t] = cpg_lctx(S,rlen,p,t); return; }
  const uint64_t x = cpg_win_end(S.p,p), z = cpg_win_end(S.p,p-32);
  const int b0 = (int)(x >> 62) & 3, b1 = (int)(x >> 60) & 3, b2 = (int)(x >> 58) & 3;     /* s[p], s[p-1], s[p-2] */
  int zb[3];
  for (int d = 1; d <= 3; d++)
    { const uint64_t y = (x << (2*d)) | (z >> (64-2*d));        /* bases p-31-d .. p-d, base p-d on top */
      const uint64_t e = x ^ y;
      int c = e ? (cpg_clz64(e) >> 1) : 32;
      const int lim = p-d+1;
      if (c >= 32 && lim > 32) c = -1;                          /* the run leaves the window */
      else if (c > lim) c = lim > 0 ? lim : 0;
      zb[d-1] = c;
    }
  if (zb[0] < 0) out[0] = cpg_lctx(S,rlen,p,CT_HP); else out[0] = cpg_cap127(1+zb[0]);
  if (p == 0 || b1 == b0) out[1] = 0;
  else if (zb[1] < 0) out[1] = cpg_lctx(S,rlen,p,CT_DS);
  else out[1] = 1+(zb[1] >> 1);
  if (p < 2 || (b2 == b1 && b1 == b0)) out[2] = 0;
  else if (zb[2] < 0) out[2] = cpg_lctx(S,rlen,p,CT_TS);
  else out[2] = 1+zb[2]/3;
}

CPG_DEV_HELPER void cpg_rctx3(const cpg_seq S, int rlen, int p, int out[3])
{ if (S.bits != 2) { for (int t = 0; t < 3; t++) out[t] = cpg_rctx(S,rlen,p,t); return; }
  const uint64_t x = cpg_win(S.p,p), z = cpg_win(S.p,p+32);
  const int b0 = (int)x & 3, b1 = (int)(x >> 2) & 3, b2 = (int)(x >> 4) & 3;                /* s[p], s[p+1], s[p+2] */
  int zf[3];
  for (int d = 1; d <= 3; d++)
    { const uint64_t y = (x >> (2*d)) | (z << (64-2*d));        /* bases p+d .. p+d+31 */
      const uint64_t e = x ^ y;
      int c = e ? (cpg_ctz64(e) >> 1) : 32;
      const int lim = rlen-d-p;
      if (c >= 32 && lim > 32) c = -1;
      else if (c > lim) c = lim > 0 ? lim : 0;
      zf[d-1] = c;
    }
  if (zf[0] < 0 || (p > 0 && cpg_base(S,p-1) == b0)) out[0] = cpg_rctx(S,rlen,p,CT_HP);      /* inside a run: it may be a long one */
  else out[0] = 1+zf[0];
  if (p >= rlen-1 || b0 == b1) out[1] = 0;
  else if (zf[1] < 0) out[1] = cpg_rctx(S,rlen,p,CT_DS);
  else out[1] = 1+(zf[1] >> 1);
  if (p > rlen-3 || (b0 == b1 && b1 == b2)) out[2] = 0;
  else if (zf[2] < 0) out[2] = cpg_rctx(S,rlen,p,CT_TS);
  else out[2] = 1+zf[2]/3;
}

CPG_DEV void cpg_ctx3_at(const cpg_seq S, int rlen, int K, int wtype, int i, int out[3])
{ if (wtype == WT_DROP) cpg_lctx3(S,rlen,i+K-2,out); else cpg_rctx3(S,rlen,i,out); }

/* ctx[wtype][i][t] of the reference, i a profile position */
CPG_DEV int cpg_ctx_at(const cpg_seq S, int rlen, int K, int wtype, int i, int t)
{ return (wtype == WT_DROP) ? cpg_lctx(S,rlen,i+K-2,t) : cpg_rctx(S,rlen,i,t); }

#endif
