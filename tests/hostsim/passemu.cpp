/*******************************************************************************************
 *  passemu.cpp -- TEST-ONLY: the source text of the two key-range-pass kernels
 *  (classpro_b200/csrc/cpg_count_pass.cuh) on host threads.  A CTA is 256 pthreads; __syncthreads is
 *  a barrier of the CTA, __ballot_sync / __shfl_sync are rendezvous of the 32 threads of a warp,
 *  atomicAdd is an atomic add; CTAs run one after the other (so a __shared__ array is a static).
 *  Exports pe_pass_sizes / pe_keys_pass for tests/test_zz_profile_producer.py, which compares them
 *  with a plain restatement (numbers of keys per pass; the multiset of keys appended in a pass).
 *******************************************************************************************/
#define CPG_HOSTSIM 1
#include <stdint.h>
#include <pthread.h>
#include <vector>
#include <string.h>

struct Dim3 { unsigned x; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 gridDim;
static pthread_barrier_t g_cta, g_warp[8];
static unsigned long long g_slot[8][32];

#define __global__ static
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __syncthreads() pthread_barrier_wait(&g_cta)
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p,v,__ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p,v,__ATOMIC_RELAXED); }
static unsigned __ballot_sync(unsigned, bool pred)
{ const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_slot[w][l] = pred ? 1 : 0;
  pthread_barrier_wait(&g_warp[w]);
  unsigned m = 0;
  for (int i = 0; i < 32; i++) m |= (unsigned)g_slot[w][i] << i;
  pthread_barrier_wait(&g_warp[w]);
  return m;
}
static unsigned long long __shfl_sync(unsigned, unsigned long long v, int src)
{ const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_slot[w][l] = v;
  pthread_barrier_wait(&g_warp[w]);
  const unsigned long long r = g_slot[w][src & 31];
  pthread_barrier_wait(&g_warp[w]);
  return r;
}

#include "../../classpro_b200/csrc/cpg_count_pass.cuh"

struct Args
  { int which, tid, bid, n_reads, K, pass, npass; const uint64_t *W; const int64_t *seq_off, *cnt_off;
    unsigned long long *sizes, *fill, cap; uint64_t *klo, *khidx; };

static void *thread_main(void *p)
{ Args *a = (Args *)p;
  threadIdx.x = (unsigned)a->tid; blockIdx.x = (unsigned)a->bid;
  if (a->which == 0) k_pass_sizes(a->n_reads,a->W,a->seq_off,a->cnt_off,a->K,a->npass,a->sizes);
  else k_kmer_keys_pass(a->n_reads,a->W,a->seq_off,a->cnt_off,a->K,a->pass,a->npass,a->fill,a->cap,a->klo,a->khidx);
  return NULL;
}

static void launch(Args proto, int grid)
{ gridDim.x = (unsigned)grid;
  pthread_barrier_init(&g_cta,NULL,CT_THREADS);
  for (int w = 0; w < 8; w++) pthread_barrier_init(&g_warp[w],NULL,32);
  for (int b = 0; b < grid; b++)
    { std::vector<Args> a(CT_THREADS,proto); std::vector<pthread_t> th(CT_THREADS);
      for (int t = 0; t < CT_THREADS; t++) { a[t].tid = t; a[t].bid = b; pthread_create(&th[t],NULL,thread_main,&a[t]); }
      for (int t = 0; t < CT_THREADS; t++) pthread_join(th[t],NULL);
    }
  pthread_barrier_destroy(&g_cta);
  for (int w = 0; w < 8; w++) pthread_barrier_destroy(&g_warp[w]);
}

static std::vector<uint64_t> padded(const uint8_t *seq, int64_t bytes)
{ std::vector<uint64_t> W((size_t)(bytes+32+7)/8,0);
  memcpy(W.data(),seq,(size_t)bytes);
  return W;
}

extern "C" void pe_pass_sizes(int grid, int n_reads, const uint8_t *seq, const int64_t *seq_off, const int64_t *cnt_off,
                              int K, int npass, unsigned long long *sizes)
{ std::vector<uint64_t> W = padded(seq,seq_off[n_reads]);
  Args a; memset(&a,0,sizeof(a));
  a.which = 0; a.n_reads = n_reads; a.K = K; a.npass = npass; a.W = W.data(); a.seq_off = seq_off; a.cnt_off = cnt_off; a.sizes = sizes;
  launch(a,grid);
}

extern "C" unsigned long long pe_keys_pass(int grid, int n_reads, const uint8_t *seq, const int64_t *seq_off, const int64_t *cnt_off,
                                           int K, int pass, int npass, unsigned long long cap, uint64_t *klo, uint64_t *khidx)
{ std::vector<uint64_t> W = padded(seq,seq_off[n_reads]);
  unsigned long long fill = 0;
  Args a; memset(&a,0,sizeof(a));
  a.which = 1; a.n_reads = n_reads; a.K = K; a.pass = pass; a.npass = npass; a.W = W.data(); a.seq_off = seq_off; a.cnt_off = cnt_off;
  a.fill = &fill; a.cap = cap; a.klo = klo; a.khidx = khidx;
  launch(a,grid);
  return fill;
}
