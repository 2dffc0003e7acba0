/*******************************************************************************************
 *  cpg_count.cuh -- element functions of the profile PRODUCER: what FastK does
 *  before ClassPro runs -- exact canonical k-mer counts of the read set at every read position,
 *  the count histogram, and the encoder side of the profile codec (src/libfastk.c:1467-1535 is the
 *  decoder; FastK itself is not in the reference tree).
 *
 *  Every function here is a pure function of one element index over flat arrays, so that the
 *  kernels of cpg_count.cu are grid-stride loops around them and the same source compiles as
 *  plain C++ (CPG_HOSTSIM) for the CPU test-suite (tests/hostsim/countsim.cpp).
 *
 *  Pipeline (cpg_count.cu):
 *    keys     k-mer at position p of read r -> canonical 2K-bit key, index m = cnt_off[r]+p
 *    sort     radix sort of (key, m) pairs: 64 low key bits, then the high bits (stable)
 *    heads    run id of every sorted element = inclusive sum of "key differs from the one before"
 *    starts   first sorted position of every run
 *    scatter  counts[m] = min(run length, 32767); one histogram entry per run
 *    encode   counts -> FastK tokens: per position the number of bytes it emits (0..2), an
 *             exclusive sum for the byte offsets, then the bytes
 *******************************************************************************************/
#ifndef CPG_COUNT_CUH
#define CPG_COUNT_CUH
#include <stdint.h>

#ifdef CPG_HOSTSIM
  #define CPG_HD static inline
#else
  #define CPG_HD __host__ __device__ __forceinline__
#endif

#define CPG_CNT_MAX 32767          /* counts saturate here (FastK's high bin, src/libfastk.c:72-83) */

/* reverse the order of the 32 two-bit groups of a word: base j <-> base 31-j */
CPG_HD uint64_t cpg_rev2_64(uint64_t v)
{ v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
  v = ((v >> 4) & 0x0f0f0f0f0f0f0f0full) | ((v & 0x0f0f0f0f0f0f0f0full) << 4);
  v = ((v >> 8) & 0x00ff00ff00ff00ffull) | ((v & 0x00ff00ff00ff00ffull) << 8);
  v = ((v >> 16) & 0x0000ffff0000ffffull) | ((v & 0x0000ffff0000ffffull) << 16);
  return (v >> 32) | (v << 32);
}

/* Canonical key of the k-mer whose first base sits at bit `bit` of the 2-bit packed stream W
   (base i of the stream in bits 2i, 2i+1; A,C,G,T = 0..3; 32 <= 2K <= 80 here, any K <= 64 in
   principle).  g = the k-mer with its base j in bits 2j; the reverse complement in the same
   form is the group reversal of ~g; the key is the smaller of the two as a 2K-bit number, so a
   k-mer and its reverse complement share it and different pairs never do.  (Any such choice
   gives the same counts; tools/cpsim.c, the harness counter, takes the minimum of the
   first-base-highest forms.)  W must be readable up to the word holding bit+2K+63. */
CPG_HD void cpg_kmer_key(const uint64_t *W, int64_t bit, int K, uint64_t *khi, uint64_t *klo)
{ const int64_t w = bit >> 6; const int s = (int)(bit & 63);
  const uint64_t a0 = W[w], a1 = W[w+1], a2 = W[w+2];
  uint64_t lo = s ? (a0 >> s) | (a1 << (64-s)) : a0;
  uint64_t hi = s ? (a1 >> s) | (a2 << (64-s)) : a1;
  const int nb = 2*K;
  uint64_t mlo, mhi;
  if (nb >= 64) { mlo = ~0ull; mhi = nb == 64 ? 0ull : (nb == 128 ? ~0ull : ((1ull << (nb-64))-1)); }
  else          { mlo = (1ull << nb)-1; mhi = 0ull; }
  lo &= mlo; hi &= mhi;
  /* reverse complement: reverse the 64 groups of the 128-bit value ~g, then drop the 128-2K low bits */
  const uint64_t clo = ~lo & mlo, chi = ~hi & mhi;
  uint64_t rhi = cpg_rev2_64(clo), rlo = cpg_rev2_64(chi);         /* 128-bit reversal: halves swap */
  const int sh = 128-nb;
  if (sh >= 64) { rlo = sh == 64 ? rhi : rhi >> (sh-64); rhi = 0ull; }
  else if (sh > 0) { rlo = (rlo >> sh) | (rhi << (64-sh)); rhi >>= sh; }
  const int fwd = (hi < rhi) || (hi == rhi && lo <= rlo);
  *khi = fwd ? hi : rhi; *klo = fwd ? lo : rlo;
}

#define CPG_HIDX_SHIFT 48                      /* sort value = high key bits << 48 | element index */
#define CPG_HIDX_MASK  ((1ull << CPG_HIDX_SHIFT)-1)

/* element m = cnt_off[r]+p of the key arrays */
CPG_HD void cpg_key_element(const uint64_t *W, int64_t bit0, int p, int64_t m, int K, uint64_t *klo, uint64_t *khidx)
{ uint64_t hi, lo;
  cpg_kmer_key(W,bit0+2*(int64_t)p,K,&hi,&lo);
  klo[m] = lo;
  khidx[m] = (hi << CPG_HIDX_SHIFT) | (uint64_t)m;
}

/* Key-range passes for read sets whose keys do not fit in device memory at once: pass of a key =
   a hash of it scaled to [0,npass), so equal keys always meet in the same pass and a pass holds about
   1/npass of the k-mers whatever the composition of the reads */
CPG_HD uint32_t cpg_key_pass(uint64_t hi, uint64_t lo, uint32_t npass)
{ uint64_t x = lo ^ (hi * 0x9e3779b97f4a7c15ull);
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)(((x >> 32) * (uint64_t)npass) >> 32);
}

/* sorted element i starts a run of equal keys */
CPG_HD uint32_t cpg_run_head(int64_t i, const uint64_t *klo, const uint64_t *khidx)
{ return (i == 0 || klo[i] != klo[i-1] || (khidx[i] >> CPG_HIDX_SHIFT) != (khidx[i-1] >> CPG_HIDX_SHIFT)) ? 1u : 0u; }

/* rid = inclusive sum of the head flags: run ids 1..nruns.  start[q-1] = first sorted position of
   run q, start[nruns] = n */
CPG_HD void cpg_run_start(int64_t i, int64_t n, const uint32_t *rid, uint32_t *start)
{ const uint32_t q = rid[i];
  if (i == 0 || rid[i-1] != q) start[q-1] = (uint32_t)i;
  if (i == n-1) start[q] = (uint32_t)n;
}

/* writes the saturated count of sorted element i to its read position; returns the unsaturated run
   length if i is the first element of its run (one histogram entry per distinct k-mer), else 0 */
CPG_HD uint32_t cpg_scatter_count(int64_t i, const uint64_t *khidx, const uint32_t *rid, const uint32_t *start, uint16_t *counts)
{ const uint32_t q = rid[i];
  const uint32_t b = start[q-1], c = start[q]-b;
  counts[khidx[i] & CPG_HIDX_MASK] = (uint16_t)(c > CPG_CNT_MAX ? CPG_CNT_MAX : c);
  return (uint32_t)i == b ? c : 0u;
}

/* ---- encoder (tools/cpsim.c:241-268 is the harness statement of the same token choice) ----------
 * Token stream of one read's counts c[0..n):  the first count as one byte (< 128) or two
 * (0x80|c>>8, c&0xff); then, left to right, a position whose count differs from the one before
 * emits a delta token -- one byte 0x40|6-bit two's complement for -32..31, else two bytes
 * 0x80|15-bit two's complement, high byte first -- and a stretch of m positions equal to the one
 * before emits ceil(m/63) run tokens (value = positions covered, 63 except the last).  A run token
 * is charged to the LAST position it covers, which knows its value from what lies behind it and one
 * count ahead: position j of its stretch (0-based) emits iff j%63 == 62 or the stretch ends there.
 *   since = number of positions between the last change (or the first count) and p, exclusive:
 *           p - last_change_position - 1 = j */
CPG_HD int cpg_enc_token(const uint16_t *c, int32_t p, int32_t n, int32_t since, uint8_t out[2])
{ const uint16_t v = c[p];
  if (p == 0)
    { if (v >= 128) { out[0] = (uint8_t)(0x80 | (v >> 8)); out[1] = (uint8_t)(v & 0xff); return 2; }
      out[0] = (uint8_t)v; return 1;
    }
  const uint16_t d = c[p-1];
  if (v != d)
    { const int diff = (int)v-(int)d;
      if (diff >= -32 && diff <= 31) { out[0] = (uint8_t)(0x40 | (diff & 0x3f)); return 1; }
      const unsigned x = (unsigned)diff & 0x7fffu;
      out[0] = (uint8_t)(0x80 | (x >> 8)); out[1] = (uint8_t)(x & 0xff); return 2;
    }
  const int j = since % 63;
  if (j == 62 || p == n-1 || c[p+1] != v) { out[0] = (uint8_t)(j+1); return 1; }
  return 0;
}

/* chg[m] = m+1 where the count differs from the one before it in the read (or is the first), else 0;
   an inclusive max-scan over all reads turns it into 1 + index of the last change at or before m
   (the first count of every read is a change, so nothing leaks from read to read) */
CPG_HD uint32_t cpg_enc_change(const uint16_t *c, int p, int64_t m)
{ return (p == 0 || c[p] != c[p-1]) ? (uint32_t)(m+1) : 0u; }

/* bytes position p emits; with prof != NULL also written, at boff[m] */
CPG_HD int cpg_enc_position(const uint16_t *c, int p, int n, int64_t m, const uint32_t *last, const int64_t *boff, uint8_t *prof)
{ uint8_t t[2];
  const int since = (int)(m-(int64_t)last[m]);             /* m - (last[m]-1) - 1 */
  const int k = cpg_enc_token(c,p,n,since,t);
  if (prof)
    { const int64_t o = boff[m];
      if (k > 0) prof[o] = t[0];
      if (k > 1) prof[o+1] = t[1];
    }
  return k;
}

#endif
