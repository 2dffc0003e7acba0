/*******************************************************************************************
 *  cpg_count.cu -- profile producer on the GPU: cpg_count_kmers and
 *  cpg_encode_profiles of include/classpro_gpu.h.  sm_100a; grid-stride kernels around the element
 *  functions of cpg_count.cuh, CUB (the toolkit's) for the radix sort and the two prefix sums.
 *
 *  All of this is HBM-bound integer work.  Algorithmic bytes per k-mer, n k-mers, r reads
 *  (K <= 32: the second sort pass falls away):
 *    k_kmer_keys      r/4 in (2-bit bases, every 64-bit word read by <= 33 neighbouring threads: L1),
 *                     16 out (64 low key bits; high key bits << 48 | index)
 *    sort             16 in + 16 out per radix pass; 8 passes of 8 bits for the low word, 2 for
 *                     the 16 high bits of a 40-mer: 320 B per k-mer, the bulk of the whole job
 *    k_run_heads      16 in, 4 out;  inclusive sum 4 in, 4 out;  k_run_starts 8 in, <= 4 out
 *    k_scatter_counts 16 in, 2 out (scattered: the index order of a sorted run is random)
 *    encoder          k_enc_change 2 in, 4 out; max-scan 4+4; k_enc_size 6 in, 1 out; sum 1 in,
 *                     8 out; k_enc_write 15 in, c out  (a fused one-warp-per-read encoder, the mirror
 *                     image of k_decode, would read 2 and write c: next step)
 *  Device memory: 2 x 16 B per k-mer for the sort's double buffers plus its histogram scratch,
 *  4 + 4 B for run ids and run starts (the latter in the idle half of a double buffer), 2 B for the
 *  counts: ~40 B per k-mer, i.e. ~4 * 10^9 k-mers in 180 GB (BASELINE config 2 has 3 * 10^9).
 *  Larger read sets need key-range passes (equal keys always land in the same pass) and, over
 *  several GPUs, an all-to-all of keys by range: the one real exchange step of this row.
 *******************************************************************************************/
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "../../include/classpro_gpu.h"
#include "cpg_count.cuh"

static thread_local char t_err[512] = "";

static int cnt_err(int code, const char *fmt, ...)
{ va_list ap; va_start(ap,fmt); vsnprintf(t_err,sizeof(t_err),fmt,ap); va_end(ap);
  return code;
}

extern "C" const char *cpg_count_error(void) { return t_err; }

#define CT_THREADS 256
#define HIDX_SHIFT CPG_HIDX_SHIFT

/* one CTA per read at a time (reads are handed out round robin), threads stride the positions */
__global__ void __launch_bounds__(CT_THREADS)
k_kmer_keys(int n_reads, const uint64_t *__restrict__ W, const int64_t *__restrict__ seq_off,
            const int64_t *__restrict__ cnt_off, int K, uint64_t *__restrict__ klo, uint64_t *__restrict__ khidx)
{ for (int r = blockIdx.x; r < n_reads; r += gridDim.x)
    { const int64_t m0 = cnt_off[r]; const int n = (int)(cnt_off[r+1]-m0);
      const int64_t bit0 = 8*seq_off[r];
      for (int p = threadIdx.x; p < n; p += CT_THREADS) cpg_key_element(W,bit0,p,m0+p,K,klo,khidx);
    }
}

#include "cpg_count_pass.cuh"

__global__ void __launch_bounds__(CT_THREADS)
k_run_heads(int64_t n, const uint64_t *__restrict__ klo, const uint64_t *__restrict__ khidx, uint32_t *__restrict__ head)
{ for (int64_t i = blockIdx.x*(int64_t)CT_THREADS+threadIdx.x; i < n; i += (int64_t)gridDim.x*CT_THREADS)
    head[i] = cpg_run_head(i,klo,khidx);
}

__global__ void __launch_bounds__(CT_THREADS)
k_run_starts(int64_t n, const uint32_t *__restrict__ rid, uint32_t *__restrict__ start)
{ for (int64_t i = blockIdx.x*(int64_t)CT_THREADS+threadIdx.x; i < n; i += (int64_t)gridDim.x*CT_THREADS)
    cpg_run_start(i,n,rid,start);
}

/* counts back to read order + the histogram of distinct k-mers: the low bins (where nearly all
   distinct k-mers are) in shared memory, flushed once per CTA */
#define HIST_SMEM 1024
__global__ void __launch_bounds__(CT_THREADS)
k_scatter_counts(int64_t n, const uint64_t *__restrict__ khidx, const uint32_t *__restrict__ rid,
                 const uint32_t *__restrict__ start, uint16_t *__restrict__ counts, unsigned long long *__restrict__ hist)
{ __shared__ unsigned int sh[HIST_SMEM];
  for (int j = threadIdx.x; j < HIST_SMEM; j += CT_THREADS) sh[j] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x*(int64_t)CT_THREADS+threadIdx.x; i < n; i += (int64_t)gridDim.x*CT_THREADS)
    { const uint32_t c = cpg_scatter_count(i,khidx,rid,start,counts);
      if (c)                                               /* once per distinct k-mer */
        { if (c < HIST_SMEM) atomicAdd(&sh[c],1u);
          else if (c < CPG_CNT_MAX) atomicAdd(&hist[c],1ull);
          else { atomicAdd(&hist[CPG_CNT_MAX],1ull); atomicAdd(&hist[32769],(unsigned long long)c); }   /* + instances of the top bin */
        }
    }
  __syncthreads();
  for (int j = threadIdx.x; j < HIST_SMEM; j += CT_THREADS)
    if (sh[j]) atomicAdd(&hist[j],(unsigned long long)sh[j]);
}

/* invariant of a finished count: every position was written by exactly one pass (the array starts as 0xffff,
   which no count can be: they saturate at 32767) */
__global__ void __launch_bounds__(CT_THREADS)
k_count_unwritten(int64_t n, const uint16_t *__restrict__ counts, unsigned long long *__restrict__ bad)
{ unsigned int c = 0;
  for (int64_t i = blockIdx.x*(int64_t)CT_THREADS+threadIdx.x; i < n; i += (int64_t)gridDim.x*CT_THREADS)
    c += (counts[i] == 0xffffu);
  c = __reduce_add_sync(0xffffffffu,c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(bad,(unsigned long long)c);
}

/* ---- encoder ---- */
__global__ void __launch_bounds__(CT_THREADS)
k_enc_change(int n_reads, const uint16_t *__restrict__ counts, const int64_t *__restrict__ cnt_off, uint32_t *__restrict__ chg)
{ for (int r = blockIdx.x; r < n_reads; r += gridDim.x)
    { const int64_t m0 = cnt_off[r]; const int n = (int)(cnt_off[r+1]-m0);
      for (int p = threadIdx.x; p < n; p += CT_THREADS) chg[m0+p] = cpg_enc_change(counts+m0,p,m0+p);
    }
}

template<bool WRITE>
__global__ void __launch_bounds__(CT_THREADS)
k_enc_tokens(int n_reads, const uint16_t *__restrict__ counts, const int64_t *__restrict__ cnt_off,
             const uint32_t *__restrict__ last, uint8_t *__restrict__ nbytes, const int64_t *__restrict__ boff,
             uint8_t *__restrict__ prof, int64_t *__restrict__ prof_off)
{ for (int r = blockIdx.x; r < n_reads; r += gridDim.x)
    { const int64_t m0 = cnt_off[r]; const int n = (int)(cnt_off[r+1]-m0);
      if (WRITE && threadIdx.x == 0) prof_off[r] = boff[m0];       /* boff has one entry past the last count */
      for (int p = threadIdx.x; p < n; p += CT_THREADS)
        { const int k = cpg_enc_position(counts+m0,p,n,m0+p,last,boff,WRITE ? prof : NULL);
          if (!WRITE) nbytes[m0+p] = (uint8_t)k;
        }
    }
}

struct Sum64 { __host__ __device__ __forceinline__ int64_t operator()(int64_t a, int64_t b) const { return a+b; } };
struct MaxOp { __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = cnt_err(CPG_ECUDA,"%s: %s",#x,cudaGetErrorString(e_)); goto done; } } while (0)
#define DMALLOC(p,bytes) do { cudaError_t e_ = cudaMalloc((void **)&(p),(bytes)); \
    if (e_ != cudaSuccess) { rc = cnt_err(CPG_ENOMEM,"cannot allocate %zu bytes of device memory (%s)",(size_t)(bytes),cudaGetErrorString(e_)); goto done; } } while (0)

/* CPG_COUNT_DEBUG_DIR=<dir>: every stage of every pass is copied back and written to <dir>/<name>.<pass>.bin
   (tools/producer_debug.py compares the passes with the single-pass arrays) */
static void dbg_dump(cudaStream_t st, const char *name, int pass, const void *dptr, size_t bytes)
{ const char *dir = getenv("CPG_COUNT_DEBUG_DIR");
  if (dir == NULL) return;
  void *h = malloc(bytes ? bytes : 1);
  if (h == NULL) return;
  cudaStreamSynchronize(st);
  cudaError_t e = cudaMemcpy(h,dptr,bytes,cudaMemcpyDeviceToHost);
  char path[4096]; snprintf(path,sizeof(path),"%s/%s.%d.bin",dir,name,pass);
  FILE *f = fopen(path,"wb");
  if (f) { fwrite(h,1,bytes,f); fclose(f); }
  if (e != cudaSuccess) fprintf(stderr,"dbg_dump %s pass %d: %s\n",name,pass,cudaGetErrorString(e));
  free(h);
}

static int grid_for(int device, int64_t work_items)
{ int sms = 148;
  cudaDeviceGetAttribute(&sms,cudaDevAttrMultiProcessorCount,device);
  int64_t want = (work_items+CT_THREADS-1)/CT_THREADS, cap = (int64_t)sms*8;        /* 8 CTAs of 256 threads per SM */
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

extern "C" int cpg_count_kmers(int device, int32_t kmer, int32_t n_reads, const uint8_t *seq, const int64_t *seq_off,
                               const int32_t *rlen, int64_t *cnt_off, uint16_t *counts, int64_t *hist)
{ if (n_reads < 0 || kmer < 1 || (n_reads > 0 && (!seq || !seq_off || !rlen)) || !cnt_off || !hist || 2*kmer > 64+(64-HIDX_SHIFT))
    return cnt_err(CPG_EINVAL,"cpg_count_kmers: bad argument (1 <= K <= %d)",(64+64-HIDX_SHIFT)/2);
  int rc = CPG_OK;
  cnt_off[0] = 0;
  for (int i = 0; i < n_reads; i++) cnt_off[i+1] = cnt_off[i]+(rlen[i] >= kmer ? rlen[i]-kmer+1 : 0);
  const int64_t n = cnt_off[n_reads];
  memset(hist,0,sizeof(int64_t)*32770);
  if (n == 0) return CPG_OK;
  if (counts == NULL) return cnt_err(CPG_EINVAL,"cpg_count_kmers: counts is NULL");
  if (n >= ((int64_t)1 << HIDX_SHIFT)) return cnt_err(CPG_EINVAL,"cpg_count_kmers: %lld k-mers in one call (limit 2^%d)",(long long)n,HIDX_SHIFT);
  if (cudaSetDevice(device) != cudaSuccess) return cnt_err(CPG_ECUDA,"no CUDA device %d: the profile producer has no CPU fallback",device);

  const size_t seq_bytes = (size_t)seq_off[n_reads];
  uint8_t *d_seq = NULL; int64_t *d_seq_off = NULL, *d_cnt_off = NULL;
  uint64_t *d_lo[2] = { NULL, NULL }, *d_hx[2] = { NULL, NULL };
  uint32_t *d_rid = NULL; uint16_t *d_counts = NULL; unsigned long long *d_hist = NULL, *d_sizes = NULL; void *d_tmp = NULL;
  cudaStream_t st = 0;
  cudaEvent_t ev[5] = { NULL, NULL, NULL, NULL, NULL };
  { CU(cudaStreamCreate(&st));
    DMALLOC(d_seq,seq_bytes+32); DMALLOC(d_seq_off,sizeof(int64_t)*(size_t)(n_reads+1)); DMALLOC(d_cnt_off,sizeof(int64_t)*(size_t)(n_reads+1));
    DMALLOC(d_counts,sizeof(uint16_t)*(size_t)n); DMALLOC(d_hist,sizeof(unsigned long long)*32770);
    DMALLOC(d_sizes,sizeof(unsigned long long)*(MAX_PASSES+1));
    CU(cudaMemsetAsync(d_seq+seq_bytes,0,32,st));
    CU(cudaMemcpyAsync(d_seq,seq,seq_bytes,cudaMemcpyHostToDevice,st));
    CU(cudaMemcpyAsync(d_seq_off,seq_off,sizeof(int64_t)*(size_t)(n_reads+1),cudaMemcpyHostToDevice,st));
    CU(cudaMemcpyAsync(d_cnt_off,cnt_off,sizeof(int64_t)*(size_t)(n_reads+1),cudaMemcpyHostToDevice,st));
    CU(cudaMemsetAsync(d_hist,0,sizeof(unsigned long long)*32770,st));
    CU(cudaMemsetAsync(d_counts,0xff,sizeof(uint16_t)*(size_t)n,st));
    const int gr = grid_for(device,(int64_t)n_reads*CT_THREADS);

    /* passes: all keys at once if 36 bytes per k-mer (two 16-byte sort buffers in double, run ids) plus the
       sort's scratch fit in what is left of the device memory; otherwise key-range passes (equal keys always land
       in the same pass): as many as it takes for the largest pass to fit.  CPG_COUNT_PASSES=<p> forces p passes. */
    unsigned long long sizes[MAX_PASSES]; int npass = 1; int64_t cap = n;
    { size_t mfree = 0, mtotal = 0;
      CU(cudaMemGetInfo(&mfree,&mtotal));
      const double budget = 0.92*(double)mfree;
      const char *e = getenv("CPG_COUNT_PASSES");
      int forced = e ? atoi(e) : 0;
      if (forced > MAX_PASSES) forced = MAX_PASSES;
      if (forced > 0) npass = forced;
      else if (37.0*(double)n*1.02 > budget || n >= (int64_t)0xfffffff0u)
        { npass = (int)(37.0*(double)n*1.05/budget)+1;
          if (npass < 2) npass = 2;
          if (npass > MAX_PASSES)
            { rc = cnt_err(CPG_ENOMEM,"cpg_count_kmers: %lld k-mers need ~%.1f GB of device memory per pass even with %d passes, %.1f GB free",
                           (long long)n,37e-9*(double)n/MAX_PASSES,MAX_PASSES,1e-9*(double)mfree);
              goto done;
            }
        }
      for (;;)
        { sizes[0] = (unsigned long long)n;
          if (npass > 1)
            { /* sizes of the passes: the append kernel itself with a capacity of 0 (it counts, stores nothing), one
                 launch per pass -- the very code that fills the buffers afterwards.  (A separate sizing kernel with
                 a shared-memory histogram, k_pass_sizes, put every key in pass 0 on a B200 although its source is
                 right on host threads: profiles/r02_producer_debug.log; it is kept for the CPU suite only.) */
              CU(cudaMemsetAsync(d_sizes,0,sizeof(unsigned long long)*MAX_PASSES,st));
              for (int p = 0; p < npass; p++)
                k_kmer_keys_pass<<<gr,CT_THREADS,0,st>>>(n_reads,(const uint64_t *)d_seq,d_seq_off,d_cnt_off,kmer,p,npass,d_sizes+p,0ull,NULL,NULL);
              CU(cudaGetLastError());
              CU(cudaMemcpyAsync(sizes,d_sizes,sizeof(unsigned long long)*(size_t)npass,cudaMemcpyDeviceToHost,st));
              CU(cudaStreamSynchronize(st));
              unsigned long long tot = 0;
              for (int p = 0; p < npass; p++) tot += sizes[p];
              if ((int64_t)tot != n)
                { rc = cnt_err(CPG_ECUDA,"cpg_count_kmers: the %d passes hold %llu of %lld k-mers",npass,tot,(long long)n); goto done; }
            }
          cap = 0;
          for (int p = 0; p < npass; p++) if ((int64_t)sizes[p] > cap) cap = (int64_t)sizes[p];
          if (forced > 0 || npass >= MAX_PASSES || (37.0*(double)cap <= budget && cap < (int64_t)0xfffffff0u)) break;
          npass++;                                               /* a heavy repeat made one pass larger than its share */
        }
      if (cap >= (int64_t)0xfffffff0u)
        { rc = cnt_err(CPG_EINVAL,"cpg_count_kmers: %lld k-mers in one pass (limit 2^32); use more passes",(long long)cap); goto done; }
    }
    for (int b = 0; b < 2; b++) { DMALLOC(d_lo[b],sizeof(uint64_t)*(size_t)(cap+2)); DMALLOC(d_hx[b],sizeof(uint64_t)*(size_t)(cap+2)); }
    DMALLOC(d_rid,sizeof(uint32_t)*(size_t)(cap+2));
    const int lo_bits = 2*kmer < 64 ? 2*kmer : 64, hi_bits = 2*kmer > 64 ? 2*kmer-64 : 0;
    size_t tmp_bytes = 0;
    { cub::DoubleBuffer<uint64_t> Q_lo(d_lo[0],d_lo[1]), Q_hx(d_hx[0],d_hx[1]);
      size_t t1 = 0, t2 = 0, t3 = 0;
      CU(cub::DeviceRadixSort::SortPairs(NULL,t1,Q_lo,Q_hx,cap,0,lo_bits,st));
      if (hi_bits) CU(cub::DeviceRadixSort::SortPairs(NULL,t2,Q_hx,Q_lo,cap,HIDX_SHIFT,HIDX_SHIFT+hi_bits,st));
      CU(cub::DeviceScan::InclusiveSum(NULL,t3,d_rid,d_rid,cap,st));
      tmp_bytes = t1 > t2 ? t1 : t2; if (t3 > tmp_bytes) tmp_bytes = t3;
    }
    DMALLOC(d_tmp,tmp_bytes+16);
    const int timing = getenv("CPG_COUNT_TIMING") != NULL;
    float ms[4] = { 0.f, 0.f, 0.f, 0.f };
    if (timing) for (int i = 0; i < 5; i++) CU(cudaEventCreate(&ev[i]));
#define MARK(i) if (timing) CU(cudaEventRecord(ev[i],st))

    for (int pass = 0; pass < npass; pass++)
      { int64_t m = (int64_t)sizes[pass];
        if (m == 0) continue;
        MARK(0);
        if (npass == 1)
          k_kmer_keys<<<gr,CT_THREADS,0,st>>>(n_reads,(const uint64_t *)d_seq,d_seq_off,d_cnt_off,kmer,d_lo[0],d_hx[0]);
        else
          { CU(cudaMemsetAsync(d_sizes+MAX_PASSES,0,sizeof(unsigned long long),st));
            k_kmer_keys_pass<<<gr,CT_THREADS,0,st>>>(n_reads,(const uint64_t *)d_seq,d_seq_off,d_cnt_off,kmer,pass,npass,d_sizes+MAX_PASSES,(unsigned long long)cap,d_lo[0],d_hx[0]);
            /* the number of keys of the pass is what was appended; k_pass_sizes only sized the buffers */
            unsigned long long filled = 0;
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(&filled,d_sizes+MAX_PASSES,sizeof(filled),cudaMemcpyDeviceToHost,st));
            CU(cudaStreamSynchronize(st));
            if ((int64_t)filled != m)
              { rc = cnt_err(CPG_ECUDA,"cpg_count_kmers: pass %d of %d: %llu keys appended, %lld expected",pass,npass,filled,(long long)m);
                goto done;
              }
          }
        CU(cudaGetLastError());
        MARK(1);
        dbg_dump(st,"keys_lo",pass,d_lo[0],sizeof(uint64_t)*(size_t)m); dbg_dump(st,"keys_hx",pass,d_hx[0],sizeof(uint64_t)*(size_t)m);
        cub::DoubleBuffer<uint64_t> B_lo(d_lo[0],d_lo[1]), B_hx(d_hx[0],d_hx[1]);
        size_t tb = tmp_bytes;
        CU(cub::DeviceRadixSort::SortPairs(d_tmp,tb,B_lo,B_hx,m,0,lo_bits,st));
        tb = tmp_bytes;
        if (hi_bits) CU(cub::DeviceRadixSort::SortPairs(d_tmp,tb,B_hx,B_lo,m,HIDX_SHIFT,HIDX_SHIFT+hi_bits,st));
        const uint64_t *s_lo = B_lo.Current(), *s_hx = B_hx.Current();
        uint32_t *d_start = (uint32_t *)B_lo.Alternate();          /* idle half of a double buffer: m+1 entries fit in 8(m+2) bytes */
        const int g = grid_for(device,m);
        MARK(2);
        dbg_dump(st,"sort_lo",pass,s_lo,sizeof(uint64_t)*(size_t)m); dbg_dump(st,"sort_hx",pass,s_hx,sizeof(uint64_t)*(size_t)m);
        k_run_heads<<<g,CT_THREADS,0,st>>>(m,s_lo,s_hx,d_rid);
        CU(cudaGetLastError());
        tb = tmp_bytes;
        CU(cub::DeviceScan::InclusiveSum(d_tmp,tb,d_rid,d_rid,m,st));
        k_run_starts<<<g,CT_THREADS,0,st>>>(m,d_rid,d_start);
        CU(cudaGetLastError());
        MARK(3);
        dbg_dump(st,"rid",pass,d_rid,sizeof(uint32_t)*(size_t)m); dbg_dump(st,"start",pass,d_start,sizeof(uint32_t)*(size_t)(m+1));
        k_scatter_counts<<<g,CT_THREADS,0,st>>>(m,s_hx,d_rid,d_start,d_counts,d_hist);
        CU(cudaGetLastError());
        MARK(4);
        dbg_dump(st,"counts",pass,d_counts,sizeof(uint16_t)*(size_t)n);
        if (timing)
          { CU(cudaStreamSynchronize(st));
            for (int i = 0; i < 4; i++) { float t; CU(cudaEventElapsedTime(&t,ev[i],ev[i+1])); ms[i] += t; }
          }
      }
    /* fail closed: every position counted exactly once, and the histogram accounts for every k-mer */
    { unsigned long long unwritten = 0;
      CU(cudaMemsetAsync(d_sizes,0,sizeof(unsigned long long),st));
      k_count_unwritten<<<grid_for(device,n),CT_THREADS,0,st>>>(n,d_counts,d_sizes);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(&unwritten,d_sizes,sizeof(unwritten),cudaMemcpyDeviceToHost,st));
      CU(cudaMemcpyAsync(hist,d_hist,sizeof(int64_t)*32770,cudaMemcpyDeviceToHost,st));
      CU(cudaStreamSynchronize(st));
      int64_t inst = hist[32769];
      for (int c = 1; c < CPG_CNT_MAX; c++) inst += (int64_t)c*hist[c];
      if (unwritten != 0 || inst != n)
        { rc = cnt_err(CPG_ECUDA,"cpg_count_kmers: inconsistent result (%llu positions never counted; histogram accounts for %lld of %lld k-mers, %d pass%s)",
                       unwritten,(long long)inst,(long long)n,npass,npass > 1 ? "es" : "");
          goto done;
        }
    }
    CU(cudaMemcpyAsync(counts,d_counts,sizeof(uint16_t)*(size_t)n,cudaMemcpyDeviceToHost,st));
    CU(cudaStreamSynchronize(st));
    hist[32768] = hist[1];                                          /* instances of the low bin (count 1) */
    if (timing)
      fprintf(stderr,"cpg_count_kmers: n = %lld, K = %d, %d pass%s: keys %.3f ms (%.0f GB/s of 16 B/k-mer out), sort %.3f ms (%.0f GB/s of %d B/k-mer), "
                     "runs %.3f ms, scatter+hist %.3f ms; %.2f G k-mers/s on the device\n",
              (long long)n,kmer,npass,npass > 1 ? "es" : "",ms[0],16e-6*n/ms[0],ms[1],1e-6*n*32*((lo_bits+7)/8+(hi_bits+7)/8)/ms[1],
              32*((lo_bits+7)/8+(hi_bits+7)/8),ms[2],ms[3],1e-6*n/(ms[0]+ms[1]+ms[2]+ms[3]));
#undef MARK
  }
done:
  for (int i = 0; i < 5; i++) if (ev[i]) cudaEventDestroy(ev[i]);
  cudaFree(d_seq); cudaFree(d_seq_off); cudaFree(d_cnt_off);
  for (int b = 0; b < 2; b++) { cudaFree(d_lo[b]); cudaFree(d_hx[b]); }
  cudaFree(d_rid); cudaFree(d_counts); cudaFree(d_hist); cudaFree(d_sizes); cudaFree(d_tmp);
  if (st) cudaStreamDestroy(st);
  return rc;
}

extern "C" int cpg_encode_profiles(int device, int32_t n_reads, const uint16_t *counts, const int64_t *cnt_off,
                                   uint8_t *prof, int64_t prof_cap, int64_t *prof_off)
{ if (n_reads < 0 || !cnt_off || !prof_off || (n_reads > 0 && cnt_off[n_reads] > 0 && (!counts || !prof)))
    return cnt_err(CPG_EINVAL,"cpg_encode_profiles: bad argument");
  int rc = CPG_OK;
  const int64_t n = n_reads > 0 ? cnt_off[n_reads] : 0;
  if (n == 0) { for (int i = 0; i <= n_reads; i++) prof_off[i] = 0; return CPG_OK; }
  if (n >= (int64_t)0xfffffff0u) return cnt_err(CPG_EINVAL,"cpg_encode_profiles: %lld counts in one call (limit 2^32)",(long long)n);
  if (cudaSetDevice(device) != cudaSuccess) return cnt_err(CPG_ECUDA,"no CUDA device %d: the profile producer has no CPU fallback",device);

  uint16_t *d_counts = NULL; int64_t *d_cnt_off = NULL, *d_boff = NULL, *d_prof_off = NULL;
  uint32_t *d_last = NULL; uint8_t *d_nb = NULL, *d_prof = NULL; void *d_tmp = NULL;
  cudaStream_t st = 0;
  int64_t total = 0;
  { CU(cudaStreamCreate(&st));
    DMALLOC(d_counts,sizeof(uint16_t)*(size_t)(n+1)); DMALLOC(d_cnt_off,sizeof(int64_t)*(size_t)(n_reads+1));
    DMALLOC(d_last,sizeof(uint32_t)*(size_t)n); DMALLOC(d_nb,(size_t)n+1); DMALLOC(d_boff,sizeof(int64_t)*(size_t)(n+1));
    DMALLOC(d_prof_off,sizeof(int64_t)*(size_t)(n_reads+1));
    CU(cudaMemcpyAsync(d_counts,counts,sizeof(uint16_t)*(size_t)n,cudaMemcpyHostToDevice,st));
    CU(cudaMemcpyAsync(d_cnt_off,cnt_off,sizeof(int64_t)*(size_t)(n_reads+1),cudaMemcpyHostToDevice,st));
    CU(cudaMemsetAsync(d_nb+n,0,1,st));
    size_t t1 = 0, t2 = 0;
    CU(cub::DeviceScan::InclusiveScan(NULL,t1,d_last,d_last,MaxOp(),n,st));
    CU(cub::DeviceScan::ExclusiveScan(NULL,t2,d_nb,d_boff,Sum64(),(int64_t)0,n+1,st));
    size_t tmp_bytes = t1 > t2 ? t1 : t2;
    DMALLOC(d_tmp,tmp_bytes+16);
    const int g = grid_for(device,(int64_t)n_reads*CT_THREADS);
    k_enc_change<<<g,CT_THREADS,0,st>>>(n_reads,d_counts,d_cnt_off,d_last);
    CU(cudaGetLastError());
    CU(cub::DeviceScan::InclusiveScan(d_tmp,tmp_bytes,d_last,d_last,MaxOp(),n,st));
    k_enc_tokens<false><<<g,CT_THREADS,0,st>>>(n_reads,d_counts,d_cnt_off,d_last,d_nb,NULL,NULL,NULL);
    CU(cudaGetLastError());
    CU(cub::DeviceScan::ExclusiveScan(d_tmp,tmp_bytes,d_nb,d_boff,Sum64(),(int64_t)0,n+1,st));   /* uint8 in, int64 sums */
    CU(cudaMemcpyAsync(&total,d_boff+n,sizeof(int64_t),cudaMemcpyDeviceToHost,st));
    CU(cudaStreamSynchronize(st));
    if (total > prof_cap)
      { rc = cnt_err(CPG_EINVAL,"cpg_encode_profiles: %lld bytes needed, capacity %lld",(long long)total,(long long)prof_cap); goto done; }
    DMALLOC(d_prof,(size_t)total+16);
    k_enc_tokens<true><<<g,CT_THREADS,0,st>>>(n_reads,d_counts,d_cnt_off,d_last,NULL,d_boff,d_prof,d_prof_off);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(prof,d_prof,(size_t)total,cudaMemcpyDeviceToHost,st));
    CU(cudaMemcpyAsync(prof_off,d_prof_off,sizeof(int64_t)*(size_t)n_reads,cudaMemcpyDeviceToHost,st));
    CU(cudaStreamSynchronize(st));
    prof_off[n_reads] = total;
  }
done:
  cudaFree(d_counts); cudaFree(d_cnt_off); cudaFree(d_last); cudaFree(d_nb); cudaFree(d_boff); cudaFree(d_prof_off);
  cudaFree(d_prof); cudaFree(d_tmp);
  if (st) cudaStreamDestroy(st);
  return rc;
}
