/*******************************************************************************************
 *  cpg_count_pass.cuh -- the two kernels of the profile producer's key-range passes (EXPERIMENTAL, see
 *  cpg_count.cu): k_pass_sizes and k_kmer_keys_pass.  In a header of their own so that the CPU suite
 *  can run their source text on host threads (tests/hostsim/passemu.cpp: 256 threads per CTA, warp
 *  votes and shuffles as rendezvous) -- they are the only code of the producer that is not a loop
 *  around an element function.
 *******************************************************************************************/
#ifndef CPG_COUNT_PASS_CUH
#define CPG_COUNT_PASS_CUH
#include "cpg_count.cuh"

#ifndef CT_THREADS
#define CT_THREADS 256
#endif
#ifndef HIDX_SHIFT
#define HIDX_SHIFT CPG_HIDX_SHIFT
#endif

/* several passes (cpg_key_pass): how many k-mers each pass will hold */
#define MAX_PASSES 64
__global__ void __launch_bounds__(CT_THREADS)
k_pass_sizes(int n_reads, const uint64_t *__restrict__ W, const int64_t *__restrict__ seq_off,
             const int64_t *__restrict__ cnt_off, int K, int npass, unsigned long long *__restrict__ sizes)
{ __shared__ unsigned int sh[MAX_PASSES];
  if (threadIdx.x < MAX_PASSES) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int r = blockIdx.x; r < n_reads; r += gridDim.x)
    { const int n = (int)(cnt_off[r+1]-cnt_off[r]);
      const int64_t bit0 = 8*seq_off[r];
      for (int p = threadIdx.x; p < n; p += CT_THREADS)
        { uint64_t hi, lo;
          cpg_kmer_key(W,bit0+2*(int64_t)p,K,&hi,&lo);
          atomicAdd(&sh[cpg_key_pass(hi,lo,(uint32_t)npass)],1u);
        }
      __syncthreads();                                   /* flush per read: a 32-bit counter cannot overflow */
      if ((int)threadIdx.x < npass && sh[threadIdx.x]) { atomicAdd(&sizes[threadIdx.x],(unsigned long long)sh[threadIdx.x]); sh[threadIdx.x] = 0; }
      __syncthreads();
    }
}

/* the keys of one pass, appended in any order (the sort does not care): one atomic per warp and step */
__global__ void __launch_bounds__(CT_THREADS)
k_kmer_keys_pass(int n_reads, const uint64_t *__restrict__ W, const int64_t *__restrict__ seq_off,
                 const int64_t *__restrict__ cnt_off, int K, int pass, int npass, unsigned long long *__restrict__ fill,
                 unsigned long long cap, uint64_t *__restrict__ klo, uint64_t *__restrict__ khidx)
{ const unsigned lane = threadIdx.x & 31u;
  for (int r = blockIdx.x; r < n_reads; r += gridDim.x)
    { const int64_t m0 = cnt_off[r]; const int n = (int)(cnt_off[r+1]-m0);
      const int64_t bit0 = 8*seq_off[r];
      for (int p0 = 0; p0 < n; p0 += CT_THREADS)          /* every thread of the CTA runs every step: full-warp ballots */
        { const int p = p0+(int)threadIdx.x;
          uint64_t hi = 0, lo = 0; bool keep = false;
          if (p < n)
            { cpg_kmer_key(W,bit0+2*(int64_t)p,K,&hi,&lo);
              keep = cpg_key_pass(hi,lo,(uint32_t)npass) == (uint32_t)pass;
            }
          const unsigned bal = __ballot_sync(0xffffffffu,keep);
          if (bal == 0) continue;
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(fill,(unsigned long long)__popc(bal));
          base = __shfl_sync(0xffffffffu,base,0);
          if (keep)
            { const unsigned long long o = base+(unsigned long long)__popc(bal & ((1u << lane)-1u));
              if (o < cap)                                 /* the host compares the fill count with the capacity */
                { klo[o] = lo;
                  khidx[o] = (hi << HIDX_SHIFT) | (uint64_t)(m0+p);
                }
            }
        }
    }
}

#endif
