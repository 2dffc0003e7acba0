/*******************************************************************************************
 *  cpg_decode.cuh -- FastK profile decoder for one read, one warp.
 *
 *  Replaces the decode loop of Fetch_Profile, src/libfastk.c:1467-1535 (the lseek/read side,
 *  src/libfastk.c:1424-1462, becomes one contiguous host read per batch + one H2D copy).
 *
 *  The stream is  <first count: 1 or 2 bytes>  then tokens
 *      00rrrrrr            repeat the current count r more times
 *      01sddddd            6-bit signed delta, 16-bit wrap-around add
 *      1sdddddd dddddddd   15-bit two's-complement delta, sum masked to 15 bits
 *  The reference walks it byte by byte.  Here a warp takes 128 bytes per step, 4 per lane:
 *   1. token boundaries: a byte is the 2nd byte of a long token iff an odd number of
 *      high-bit-set bytes immediately precede it back to the last high-bit-clear byte.  Runs that
 *      span whole lanes have even length, so a lane only needs the trailing run of the nearest
 *      lower lane that is not all-high (one ballot, one count-leading-ones, one shuffle) or, if
 *      there is none, the carry of the previous step;
 *   2. values: inclusive warp scan of the per-lane delta sums mod 2^16.  The 15-bit mask of long
 *      tokens is honoured exactly: the low 15 bits of the running count always equal those of the
 *      plain sum, so the count right after the last long token (a max-scan of lane indices) is the
 *      plain sum up to there masked to 15 bits, and 16-bit wrap-around adds continue from it;
 *   3. expansion: scans of the output counts and of the number of emitting tokens give a compact
 *      per-warp table (value, end offset) in stream order; the lanes then write the step's
 *      outputs in aligned groups of eight counts (one 16-byte store per group), each finding its
 *      first token by binary search and moving on at most one token per count.  The ragged head
 *      and tail of a step are written count by count.
 *  Bit-exact on arbitrary byte streams (wrap-around, mask, zero-length runs), not only on
 *  well-formed ones.
 *
 *  Fused candidate scan.  find_wall starts from the positions i >= 1 whose count differs from the
 *  previous one by at least MIN_CNT_CHANGE while the smaller of the two is below the repeat
 *  threshold (src/wall.c:590-608).  Such a position is always the first output of a token (counts
 *  do not change inside a run), so the decoder, which has both values in registers, sets bit i of
 *  the read's candidate bit map right there; k_classify then walks the bit map (1 bit per
 *  position) instead of sweeping the counts again (16 bits per position).
 *******************************************************************************************/
#ifndef CPG_DECODE_CUH
#define CPG_DECODE_CUH
#include "cpg_common.h"

#if defined(CPG_HOSTSIM) && CPG_HOSTSIM == 32
CPG_DEV unsigned dc_ballot(int p) { return cpg_sim_ballot(p); }
CPG_DEV unsigned dc_shfl(unsigned v, int src) { return cpg_sim_shfl(v,src); }
CPG_DEV unsigned dc_shfl_up(unsigned v, int d, int lane) { unsigned r = cpg_sim_shfl_up(v,d); return lane >= d ? r : 0u; }
CPG_DEV unsigned dc_scan_add(unsigned v, int lane)
{ for (int d = 1; d < 32; d <<= 1) { unsigned t = cpg_sim_shfl_up(v,d); if (lane >= d) v += t; }
  return v;
}
CPG_DEV int dc_scan_max(int v, int lane)
{ for (int d = 1; d < 32; d <<= 1) { int t = (int)cpg_sim_shfl_up((unsigned)v,d); if (lane >= d && t > v) v = t; }
  return v;
}
CPG_DEV int dc_clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
#elif defined(CPG_HOSTSIM)
CPG_DEV unsigned dc_ballot(int p) { return p ? 1u : 0u; }
CPG_DEV unsigned dc_shfl(unsigned v, int src) { (void)src; return v; }
CPG_DEV unsigned dc_shfl_up(unsigned v, int d, int lane) { (void)v; (void)d; (void)lane; return 0u; }
CPG_DEV unsigned dc_scan_add(unsigned v, int lane) { (void)lane; return v; }
CPG_DEV int      dc_scan_max(int v, int lane) { (void)lane; return v; }
CPG_DEV int      dc_clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
#else
CPG_DEV unsigned dc_ballot(int p) { return __ballot_sync(0xffffffffu,p); }
CPG_DEV unsigned dc_shfl(unsigned v, int src) { return __shfl_sync(0xffffffffu,v,src); }
CPG_DEV unsigned dc_shfl_up(unsigned v, int d, int lane)
{ unsigned r = __shfl_up_sync(0xffffffffu,v,d); return lane >= d ? r : 0u; }
CPG_DEV unsigned dc_scan_add(unsigned v, int lane)
{ for (int d = 1; d < 32; d <<= 1)
    { unsigned t = __shfl_up_sync(0xffffffffu,v,d); if (lane >= d) v += t; }
  return v;
}
CPG_DEV int dc_scan_max(int v, int lane)
{ for (int d = 1; d < 32; d <<= 1)
    { int t = __shfl_up_sync(0xffffffffu,v,d); if (lane >= d && t > v) v = t; }
  return v;
}
CPG_DEV int dc_clz(unsigned v) { return __clz((int)v); }
#endif

#ifndef DC_BPL
#define DC_BPL    4                      /* bytes per lane per step (4 or 8; 8 measured slower: 4.25 vs 3.84 ms) */
#endif
#define DC_SLOTS  (DC_BPL*CPG_WARP)      /* token slots per step */

/* The step's table holds one word per EMITTING token, in stream order: count value << 16 | end
 * offset (exclusive) of its outputs inside the step.  Owner of step-relative output t = first
 * token whose end offset is > t. */
CPG_DEV int dc_owner(const unsigned *tab, int ntok, int t)
{ int lo = 0, hi = ntok-1;
  while (lo < hi)
    { int mid = (lo+hi) >> 1;
      if ((int)(tab[mid] & 0xffffu) > t) hi = mid; else lo = mid+1;
    }
  return lo;
}

/* Decodes `len` bytes at `src` into at most `cap` counts at `out`; returns the
 * decoded length (which may exceed cap, as Fetch_Profile's return value does).  `tab` is a
 * per-warp shared-memory array of DC_SLOTS words (a step emits at most DC_SLOTS*63 <= 16128 <
 * 65536 counts, so offsets fit 16 bits). */
#if defined(CPG_HOSTSIM)
CPG_DEV void dc_or(uint32_t *w, uint32_t bit) { __atomic_fetch_or(w,bit,__ATOMIC_RELAXED); }
#else
CPG_DEV void dc_or(uint32_t *w, uint32_t bit) { atomicOr(w,bit); }
#endif

/* the candidate test of src/wall.c:594-608 on two neighbouring counts */
CPG_DEV int dc_is_cand(unsigned a, unsigned b, int rcov)
{ const int d = (a > b) ? (int)(a-b) : (int)(b-a);
  return d >= CPG_MIN_CNT_CHANGE && (int)((a < b) ? a : b) < rcov;
}

/* `cand` (may be NULL) is the read's candidate bit map, ceil(cap/32) words, zeroed here. */
CPG_DEV_NOINL int decode_profile(const uint8_t *src, int64_t len, uint16_t *out, int cap,
                                 int lane, unsigned *tab, uint32_t *cand, int rcov)
{ if (cand != 0)
    { for (int w = lane; w < (cap+31)/32; w += CPG_WARP) cand[w] = 0u;
      CPG_SYNCWARP();
    }
  if (len <= 0) return 0;
  unsigned x0 = src[0];
  unsigned v_in; int64_t off;
  if (x0 & 0x80) { v_in = ((x0 & 0x7f) << 8) | (len > 1 ? src[1] : 0); off = 2; }
  else           { v_in = x0; off = 1; }
  if (lane == 0 && cap > 0) out[0] = (uint16_t)v_in;
  int n = 1;
  unsigned carry = 0, carry_hi = 0;      /* the first byte of the step is the 2nd byte of a long token */
  const int sh = (int)((((size_t)out) >> 1) & 7);      /* misalignment of the row, in counts */

  for (; off < len; off += DC_SLOTS)
    { /* ---- this lane's bytes ---- */
      const int64_t p0 = off+(int64_t)lane*DC_BPL;
      unsigned b[DC_BPL]; int nv = 0;
#ifdef CPG_HOSTSIM
      for (int k = 0; k < DC_BPL; k++)
        { int ok = (p0+k < len);
          b[k] = ok ? src[p0+k] : 0u;
          nv += ok;
        }
#else
      { /* aligned 32-bit loads + funnel shifts instead of byte loads; the profile buffer is padded,
           so reading up to 7 bytes past the read's stream stays inside it */
        unsigned word[DC_BPL/4];
        for (int k = 0; k < DC_BPL/4; k++) word[k] = 0;
        nv = (p0 >= len) ? 0 : ((len-p0 >= DC_BPL) ? DC_BPL : (int)(len-p0));
        if (nv > 0)
          { const size_t a = (size_t)(src+p0);
            const unsigned *wp = reinterpret_cast<const unsigned *>(a & ~(size_t)3);
            const unsigned shb = (unsigned)(a & 3)*8;
            unsigned lo32 = wp[0];
            for (int k = 0; k < DC_BPL/4; k++)
              { const unsigned hi32 = (shb || k+1 < DC_BPL/4) ? wp[k+1] : 0u;
                word[k] = __funnelshift_r(lo32,hi32,shb);
                lo32 = hi32;
              }
          }
        for (int k = 0; k < DC_BPL; k++) b[k] = (k < nv) ? ((word[k >> 2] >> (8*(k & 3))) & 0xffu) : 0u;
      }
#endif
      /* trailing run of high-bit-set bytes among the valid ones; full = all present and high */
      int trail = 0;
      for (int k = nv-1; k >= 0 && (b[k] & 0x80); k--) trail++;
      const int full = (nv == DC_BPL && trail == DC_BPL);
      const unsigned F = dc_ballot(full);
      int z = 0;
      if (lane > 0) { z = dc_clz(~(F << (32-lane))); if (z > lane) z = lane; }
      const int srcl = lane-1-z;
      const unsigned tsrc = dc_shfl((unsigned)trail,srcl < 0 ? 0 : srcl);
      int sec = (srcl < 0) ? (int)carry : (int)(tsrc & 1);
      const unsigned up = dc_shfl_up(b[DC_BPL-1],1,lane);
      unsigned prev = (lane == 0) ? carry_hi : up;

      /* ---- tokens of this lane: output count, delta, mask flag per byte slot ---- */
      unsigned cnt[DC_BPL], add[DC_BPL]; int msk[DC_BPL];
      unsigned C = 0, A = 0, A2 = 0; int has_mask = 0;
      for (int k = 0; k < DC_BPL; k++)
        { cnt[k] = 0; add[k] = 0; msk[k] = 0;
          if (k < nv)
            { const unsigned x = b[k];
              if (sec)
                { unsigned w = (prev & 0x40) ? ((prev << 8) & 0xffffu) : ((prev << 8) & 0x7fffu);
                  add[k] = (w | x) & 0xffffu; cnt[k] = 1; msk[k] = 1; sec = 0;
                }
              else if (x & 0x80) sec = 1;
              else if ((x & 0xc0) == 0) cnt[k] = x;
              else { add[k] = (x & 0x20) ? ((x & 0x1fu) | 0xffe0u) : (x & 0x1fu); cnt[k] = 1; }
              prev = x;
            }
          C += cnt[k]; A += add[k];
          if (msk[k]) { has_mask = 1; A2 = 0; } else A2 += add[k];
        }
      /* ---- count at the start of this lane ---- */
      const unsigned S = dc_scan_add(A,lane) & 0xffffu;
      const unsigned Mk = (S-(A2 & 0xffffu)) & 0xffffu;              /* plain sum up to the lane's last long token */
      const int q = dc_scan_max(has_mask ? lane : -1,lane);
      const unsigned Sp = dc_shfl_up(S,1,lane);                      /* 0 for lane 0 */
      const int qp = (int)dc_shfl_up((unsigned)(q+1),1,lane)-1;      /* -1 for lane 0 */
      const unsigned Mq = dc_shfl(Mk,qp < 0 ? 0 : qp);
      unsigned v;
      if (qp < 0) v = (v_in+Sp) & 0xffffu;
      else        v = ((((v_in+Mq) & 0x7fffu)+((Sp-Mq) & 0xffffu)) & 0xffffu);
      /* ---- token table: one entry per emitting token (value, end offset), compacted ---- */
      const unsigned incl = dc_scan_add(C,lane);
      unsigned o = incl-C;
      const int total = (int)dc_shfl(incl,CPG_WARP-1);
      unsigned m = 0;
      for (int k = 0; k < DC_BPL; k++) m += (cnt[k] != 0);
      const unsigned mincl = dc_scan_add(m,lane);
      unsigned slot = mincl-m;
      const int ntok = (int)dc_shfl(mincl,CPG_WARP-1);
      for (int k = 0; k < DC_BPL; k++)
        { const unsigned vb = v;
          v = msk[k] ? ((v+add[k]) & 0x7fffu) : ((v+add[k]) & 0xffffu);
          if (cnt[k] != 0)
            { if (cand != 0 && dc_is_cand(vb,v,rcov))
                { const int p = n+(int)o;                          /* first output of this token */
                  if (p < cap) dc_or(cand+(p >> 5),1u << (p & 31));
                }
              o += cnt[k]; tab[slot++] = (v << 16) | o;
            }
        }
      const unsigned v_end = v;
      CPG_SYNCWARP();

      /* ---- expansion: outputs n .. n+total-1 ---- */
      if (total > 0)
        { const int end = n+total;
          /* groups of 8 counts whose address is 16-byte aligned: positions p with (p+sh) % 8 == 0 */
          const int g0 = ((n+sh+7) & ~7)-sh, g1 = ((end+sh) & ~7)-sh;
          if (g0 < g1)
            { /* each lane takes a contiguous range of groups: one table search per lane and step,
                 then a plain walk (stores of a warp instruction are 16*gc bytes apart; L2 merges them) */
              const int G = (g1-g0) >> 3, gc = (G+CPG_WARP-1)/CPG_WARP;
              const int gb = lane*gc, ge = (gb+gc < G) ? gb+gc : G;
              if (gb < ge)
                { int p = g0+8*gb, t = p-n, s = dc_owner(tab,ntok,t);
                  unsigned cur = tab[s], cur_end = cur & 0xffffu;
                  for (int g = gb; g < ge; g++, p += 8)
                    { unsigned w[4] = {0,0,0,0};
                      if ((unsigned)t >= cur_end) { cur = tab[++s]; cur_end = cur & 0xffffu; }       /* every token emits >= 1 */
                      if ((unsigned)(t+8) <= cur_end)
                        { /* the whole group lies inside one token (long runs): splat its value */
                          const unsigned vv = (cur >> 16) | (cur & 0xffff0000u);
                          w[0] = w[1] = w[2] = w[3] = vv;
                          t += 8;
                        }
                      else
                        /* count by count, fully unrolled and predicated.  (Forming the group token by
                           token -- splat + one XOR per token boundary -- was measured slower, 4.93 vs
                           3.84 ms: its trip count differs from lane to lane.) */
                        for (int e = 0; e < 8; e++, t++)
                          { if ((unsigned)t >= cur_end) { cur = tab[++s]; cur_end = cur & 0xffffu; }
                            w[e >> 1] |= (cur >> 16) << ((e & 1)*16);
                          }
                      if (p+8 <= cap)
                        {
#ifdef CPG_HOSTSIM
                          for (int e = 0; e < 8; e++) out[p+e] = (uint16_t)((w[e >> 1] >> ((e & 1)*16)) & 0xffffu);
#else
                          *reinterpret_cast<uint4 *>(out+p) = make_uint4(w[0],w[1],w[2],w[3]);
#endif
                        }
                      else
                        for (int e = 0; e < 8; e++)
                          if (p+e < cap) out[p+e] = (uint16_t)((w[e >> 1] >> ((e & 1)*16)) & 0xffffu);
                    }
                }
            }
          /* ragged head [n,min(g0,end)) and tail [max(g1,g0),end): at most 7 counts each */
          const int hend = (g0 < end) ? g0 : end;
          const int tbeg = (g1 > g0) ? g1 : hend;
          const int nh = hend-n, nt = end-tbeg;
          for (int i = lane; i < nh+nt; i += CPG_WARP)
            { const int p = (i < nh) ? n+i : tbeg+(i-nh);
              if (p < cap) out[p] = (uint16_t)(tab[dc_owner(tab,ntok,p-n)] >> 16);
            }
        }
      CPG_SYNCWARP();

      n += total;
      v_in = dc_shfl(v_end,CPG_WARP-1);
      carry = dc_shfl((unsigned)sec,CPG_WARP-1);
      carry_hi = dc_shfl(prev,CPG_WARP-1);
    }
  return n;
}

#endif
