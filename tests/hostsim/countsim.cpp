/*******************************************************************************************
 *  countsim.cpp -- TEST-ONLY host build of the profile producer (classpro_b200/csrc/cpg_count.cu):
 *  the same two entry points over the SAME element functions (cpg_count.cuh compiled with
 *  CPG_HOSTSIM), every kernel a plain loop over its index space, the device library calls replaced
 *  by their definitions (stable sort by the low key word then by the high bits; inclusive sum;
 *  inclusive max-scan; exclusive sum).  Checks the logic of the kernels on machines without a GPU
 *  against the harness counter of tools/cpsim.c; never linked into libclasspro_b200.so.
 *******************************************************************************************/
#define CPG_HOSTSIM 1
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <numeric>
#include "../../classpro_b200/csrc/cpg_count.cuh"

extern "C" int sim_count_kmers(int device, int32_t kmer, int32_t n_reads, const uint8_t *seq, const int64_t *seq_off,
                               const int32_t *rlen, int64_t *cnt_off, uint16_t *counts, int64_t *hist)
{ (void)device;
  if (2*kmer > 64+(64-CPG_HIDX_SHIFT)) return 1;
  cnt_off[0] = 0;
  for (int i = 0; i < n_reads; i++) cnt_off[i+1] = cnt_off[i]+(rlen[i] >= kmer ? rlen[i]-kmer+1 : 0);
  const int64_t n = cnt_off[n_reads];
  memset(hist,0,sizeof(int64_t)*32770);
  if (n == 0) return 0;
  /* the device copy of seq: 8-byte aligned and readable 32 bytes past its end */
  std::vector<uint64_t> W((size_t)(seq_off[n_reads]+32+7)/8,0);
  memcpy(W.data(),seq,(size_t)seq_off[n_reads]);
  /* passes (CPG_COUNT_PASSES, as in cpg_count.cu): the keys of one pass appended in read order */
  const char *e = getenv("CPG_COUNT_PASSES");
  const int npass = (e && atoi(e) > 0) ? (atoi(e) > 64 ? 64 : atoi(e)) : 1;
  for (int pass = 0; pass < npass; pass++)
    { std::vector<uint64_t> klo, khx;
      if (npass == 1)
        { klo.resize((size_t)n); khx.resize((size_t)n);
          for (int r = 0; r < n_reads; r++)                                              /* k_kmer_keys */
            for (int p = 0; p < (int)(cnt_off[r+1]-cnt_off[r]); p++)
              cpg_key_element(W.data(),8*seq_off[r],p,cnt_off[r]+p,kmer,klo.data(),khx.data());
        }
      else
        for (int r = 0; r < n_reads; r++)                                                /* k_kmer_keys_pass */
          for (int p = 0; p < (int)(cnt_off[r+1]-cnt_off[r]); p++)
            { uint64_t hi, lo;
              cpg_kmer_key(W.data(),8*seq_off[r]+2*(int64_t)p,kmer,&hi,&lo);
              if (cpg_key_pass(hi,lo,(uint32_t)npass) != (uint32_t)pass) continue;
              klo.push_back(lo); khx.push_back((hi << CPG_HIDX_SHIFT) | (uint64_t)(cnt_off[r]+p));
            }
      const int64_t m = (int64_t)klo.size();
      if (m == 0) continue;
      /* SortPairs(keys = low word, bits [0,min(64,2K))), then SortPairs(keys = value word, bits [48,48+2K-64)) */
      std::vector<int64_t> perm((size_t)m);
      std::iota(perm.begin(),perm.end(),(int64_t)0);
      const int lo_bits = 2*kmer < 64 ? 2*kmer : 64;
      const uint64_t lomask = lo_bits == 64 ? ~0ull : ((1ull << lo_bits)-1);
      std::stable_sort(perm.begin(),perm.end(),[&](int64_t a, int64_t b) { return (klo[a] & lomask) < (klo[b] & lomask); });
      if (2*kmer > 64)
        std::stable_sort(perm.begin(),perm.end(),[&](int64_t a, int64_t b) { return (khx[a] >> CPG_HIDX_SHIFT) < (khx[b] >> CPG_HIDX_SHIFT); });
      std::vector<uint64_t> slo((size_t)m), shx((size_t)m);
      for (int64_t i = 0; i < m; i++) { slo[i] = klo[perm[i]]; shx[i] = khx[perm[i]]; }
      std::vector<uint32_t> rid((size_t)m), start((size_t)m+1);
      for (int64_t i = 0; i < m; i++) rid[i] = cpg_run_head(i,slo.data(),shx.data());    /* k_run_heads */
      for (int64_t i = 1; i < m; i++) rid[i] += rid[i-1];                                 /* InclusiveSum */
      for (int64_t i = m-1; i >= 0; i--) cpg_run_start(i,m,rid.data(),start.data());      /* k_run_starts (any order) */
      for (int64_t i = 0; i < m; i++)                                                     /* k_scatter_counts */
        { const uint32_t c = cpg_scatter_count(i,shx.data(),rid.data(),start.data(),counts);
          if (c)
            { if (c < CPG_CNT_MAX) hist[c] += 1;
              else { hist[CPG_CNT_MAX] += 1; hist[32769] += c; }
            }
        }
    }
  hist[32768] = hist[1];
  return 0;
}

extern "C" int sim_encode_profiles(int device, int32_t n_reads, const uint16_t *counts, const int64_t *cnt_off,
                                   uint8_t *prof, int64_t prof_cap, int64_t *prof_off)
{ (void)device;
  const int64_t n = n_reads > 0 ? cnt_off[n_reads] : 0;
  if (n == 0) { for (int i = 0; i <= n_reads; i++) prof_off[i] = 0; return 0; }
  std::vector<uint32_t> last((size_t)n);
  std::vector<uint8_t> nb((size_t)n+1,0);
  std::vector<int64_t> boff((size_t)n+1);
  for (int r = 0; r < n_reads; r++)                                                  /* k_enc_change */
    for (int p = 0; p < (int)(cnt_off[r+1]-cnt_off[r]); p++)
      last[cnt_off[r]+p] = cpg_enc_change(counts+cnt_off[r],p,cnt_off[r]+p);
  for (int64_t i = 1; i < n; i++) last[i] = std::max(last[i],last[i-1]);             /* InclusiveScan(max) */
  for (int r = n_reads-1; r >= 0; r--)                                               /* k_enc_tokens<false> */
    for (int p = (int)(cnt_off[r+1]-cnt_off[r])-1; p >= 0; p--)
      nb[cnt_off[r]+p] = (uint8_t)cpg_enc_position(counts+cnt_off[r],p,(int)(cnt_off[r+1]-cnt_off[r]),cnt_off[r]+p,last.data(),NULL,NULL);
  boff[0] = 0;
  for (int64_t i = 0; i < n; i++) boff[i+1] = boff[i]+nb[i];                          /* ExclusiveScan(sum) over n+1 */
  if (boff[n] > prof_cap) return 1;
  for (int r = 0; r < n_reads; r++)                                                  /* k_enc_tokens<true> */
    { prof_off[r] = boff[cnt_off[r]];
      for (int p = 0; p < (int)(cnt_off[r+1]-cnt_off[r]); p++)
        cpg_enc_position(counts+cnt_off[r],p,(int)(cnt_off[r+1]-cnt_off[r]),cnt_off[r]+p,last.data(),boff.data(),prof);
    }
  prof_off[n_reads] = boff[n];
  return 0;
}

#ifdef COUNTSIM_AS_ABI
/* the product's entry points over the loops above, for the CPU tests of the `profiler` program's own
   source (tests/hostsim/build.sh links host/cpg_profiler.c against this file instead of the library) */
extern "C" int cpg_count_kmers(int device, int32_t kmer, int32_t n_reads, const uint8_t *seq, const int64_t *seq_off,
                               const int32_t *rlen, int64_t *cnt_off, uint16_t *counts, int64_t *hist)
{ return sim_count_kmers(device,kmer,n_reads,seq,seq_off,rlen,cnt_off,counts,hist); }
extern "C" int cpg_encode_profiles(int device, int32_t n_reads, const uint16_t *counts, const int64_t *cnt_off,
                                   uint8_t *prof, int64_t prof_cap, int64_t *prof_off)
{ return sim_encode_profiles(device,n_reads,counts,cnt_off,prof,prof_cap,prof_off); }
extern "C" const char *cpg_count_error(void) { return "countsim"; }
#endif
