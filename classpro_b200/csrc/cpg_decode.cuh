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
 *  The reference walks it byte by byte.  Here a warp takes 32 bytes per step:
 *   1. token boundaries: a byte is the 2nd byte of a long token iff an odd number of
 *      high-bit-set bytes immediately precede it back to the last high-bit-clear byte (or to the
 *      carry of the previous step) -- one ballot + a count-leading-ones per lane;
 *   2. values: inclusive warp scan of the deltas mod 2^16; the 15-bit mask of long tokens is
 *      honoured exactly by restarting from the low 15 bits of the plain sum at the last long
 *      token (a max-scan of lane indices), because the low 15 bits of the running count always
 *      equal those of the plain sum;
 *   3. expansion: exclusive scan of the per-token output counts, then the lanes write the
 *      outputs of the step coalesced, each finding its token by binary search over the 32 offsets.
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
CPG_DEV unsigned dc_shfl_up(unsigned v, int d, int lane) { (void)d; (void)lane; return v; }
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

/* Decodes `len` bytes at `src` into at most `cap` counts at `out`; returns the decoded length
 * (which may exceed cap, as Fetch_Profile's return value does).  `offs` is a per-warp shared-memory
 * array of 2*CPG_WARP ints. */
CPG_DEV_NOINL int decode_profile(const uint8_t *src, int64_t len, uint16_t *out, int cap,
                                 int lane, int *offs)
{ if (len <= 0) return 0;
  int *vals = offs+CPG_WARP;
  unsigned x0 = src[0];
  unsigned v_in; int64_t off;
  if (x0 & 0x80) { v_in = ((x0 & 0x7f) << 8) | (len > 1 ? src[1] : 0); off = 2; }
  else           { v_in = x0; off = 1; }
  if (lane == 0 && cap > 0) out[0] = (uint16_t)v_in;
  int n = 1;
  unsigned carry = 0, carry_hi = 0;      /* lane 0 of the step is the 2nd byte of a long token */

  for (; off < len; off += CPG_WARP)
    { const int valid = (off+lane < len);
      const unsigned x = valid ? src[off+lane] : 0u;
      const unsigned hi = valid && (x & 0x80);
      const unsigned H = dc_ballot(hi);
      /* consecutive high-bit-set bytes right before this lane */
      int c = 0;
      if (lane > 0)
        { unsigned below = H << (32-lane);            /* bit 31 = lane-1 */
          c = dc_clz(~below);
          if (c > lane) c = lane;
        }
      if (c == lane) c += (int)carry;
      const int second = valid && (c & 1);
      const unsigned up = dc_shfl_up(x,1,lane);          /* every lane takes part in the shuffle */
      const unsigned prev = (lane == 0) ? carry_hi : up;

      unsigned a = 0, cnt = 0; int masked = 0;
      if (valid)
        { if (second)
            { unsigned w = (prev & 0x40) ? ((prev << 8) & 0xffffu) : ((prev << 8) & 0x7fffu);
              a = (w | x) & 0xffffu; cnt = 1; masked = 1;
            }
          else if ((x & 0xc0) == 0) cnt = x;
          else if (!(x & 0x80))
            { a = (x & 0x20) ? ((x & 0x1fu) | 0xffe0u) : (x & 0x1fu); cnt = 1; }
        }
      const unsigned S = dc_scan_add(a,lane) & 0xffffu;
      const int q = dc_scan_max(masked ? lane : -1,lane);
      const unsigned Sq = dc_shfl(S,q < 0 ? 0 : q);
      unsigned v;
      if (q < 0) v = (v_in+S) & 0xffffu;
      else       v = ((((v_in+Sq) & 0x7fffu)+((S-Sq) & 0xffffu)) & 0xffffu);
      const unsigned incl = dc_scan_add(cnt,lane);
      const unsigned excl = incl-cnt;
      const int total = (int)dc_shfl(incl,CPG_WARP-1);

      offs[lane] = (int)excl;
      vals[lane] = (int)v;          /* run tokens leave the count unchanged, so v is what they repeat */
      CPG_SYNCWARP();
      for (int t = lane; t < total; t += CPG_WARP)
        { /* owner of output t = LAST lane whose exclusive offset is <= t (tokens that emit
             nothing share the offset of their successor) */
          int lo = 0, hi2 = CPG_WARP-1;
          while (lo < hi2)
            { int mid = (lo+hi2+1) >> 1;
              if (offs[mid] <= t) lo = mid; else hi2 = mid-1;
            }
          if (n+t < cap) out[n+t] = (uint16_t)vals[lo];
        }
      CPG_SYNCWARP();

      n += total;
      v_in = dc_shfl(v,CPG_WARP-1);
      /* carry: the last byte of a full step starts a long token */
      const unsigned last_first_hi = valid && !second && (x & 0x80);
      carry = dc_shfl(last_first_hi,CPG_WARP-1);
      carry_hi = dc_shfl(x,CPG_WARP-1);
    }
  return n;
}

#endif
