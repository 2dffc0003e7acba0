/*******************************************************************************************
 *  cpg_kernels.cu -- sm_100a kernels and the C ABI of libclasspro_b200.so.
 *
 *  Nine launches per batch, all persistent (grid = a multiple of the SM count, warps / lane groups
 *  pull reads from an atomic queue in processing order):
 *
 *   k_decode    one warp per read: FastK profile bytes -> uint16 counts + the wall-candidate bit
 *               map (cpg_decode.cuh).  Streaming: c + 2n + n/8 bytes per read.
 *   k_wall_a    one warp per read, one wall CANDIDATE per lane: the pure part of wall detection
 *               (context, count thresholds, every probability pass A can ask for) -> a 16-byte
 *               header per candidate, a 216-byte record for the one in ten that needs it.
 *   k_wall_b    one lane group per read: the order-dependent replay of find_wall on those records
 *               (probability cache, paired flags, E-intervals, cuts) -> the read's interval table
 *               in the batch's interval pool.  Touches no count and no base.
 *   k_wall_c    one interval per lane: end counts, corrected counts, Skellam plausibility; the
 *               reliable intervals are appended to the read's table (cpg_wall.cuh).
 *   k_rel       reliable-interval DP, forward and backward (cpg_rel.cuh), on the pooled tables.
 *   k_unrel_a   one interval per lane: the pure part of the unreliable pass (neighbours, ten task
 *               values per interval the sweeps will visit).
 *   k_unrel_b   the two order-dependent sweeps on those values (cpg_unrel.cuh).
 *   k_emit      class strings, streaming: r bytes per read out.
 *   k_classify  the three phases in one kernel, on 2 CTAs with worst-case scratch: the retry
 *               launch for reads that outgrew the compact scratch blocks or the pool (normally
 *               none; it returns at once).  CPG_FUSED=1 runs every read through it instead.
 *  The classification kernels are bound by FP64 dependency chains (Bessel recurrences), DRAM
 *  latency and serial control flow, not by bandwidth.
 *
 *  No tensor cores: nothing on this path is a dense contraction (integer/byte scans and scalar
 *  FP64 recurrences).  Compiled with -fmad=false: see cpg_math.cuh.
 *******************************************************************************************/
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "../../include/classpro_gpu.h"
#include "cpg_unrel.cuh"
#include "cpg_decode.cuh"

#define DECODE_THREADS   256
#ifndef DECODE_MIN_BLOCKS
#define DECODE_MIN_BLOCKS 4
#endif
#ifndef CLASSIFY_THREADS
#define CLASSIFY_THREADS 128
#endif
#ifndef CLASSIFY_MIN_BLOCKS
#define CLASSIFY_MIN_BLOCKS 4
#endif

struct BatchDev
  { int32_t        n_reads;
    int32_t        seq_bits;
    const uint8_t *seq;
    const int64_t *seq_off;
    const int32_t *rlen;
    const uint8_t *prof;
    const int64_t *prof_off;
    uint16_t      *cnt;
    const int64_t *cnt_off;      /* multiples of 32 counts (rows are 64-byte aligned) */
    uint32_t      *cand;         /* wall-candidate bit map, bit cnt_off[r]+i = position i of read r (NULL: not wanted) */
    int32_t       *plen;
    uint8_t       *cls;
    const int64_t *cls_off;
    int32_t       *status;
    const int32_t *order;
    int32_t       *queue;        /* work counters: [0] decode, [1] classify/wall_a, [2] retry launch; [3] reads flagged
                                    for retry; [4] reliable DP, [5] unrel_a, [6] wall_b, [7] wall_c, [24] unrel_b, [25] emit */
    struct ReadRec *rec;         /* per read: where the wall kernels left its candidate records and interval tables */
    cpg_intvl     *pool;         /* interval pool of the batch: intvl[N] then rint[M] of each read */
    unsigned long long *pool_cursor;
    int64_t        pool_cap;     /* entries */
    cpg_chdr      *hdr;          /* candidate headers of the batch, a read's in position order */
    cpg_cbig      *big;          /* big candidate records */
    unsigned long long *hdr_cursor, *big_cursor;
    int64_t        hdr_cap, big_cap;
    uint32_t      *ivl;          /* compact result (CPG_RESULT_INTERVALS): packed interval of pool entry i at ivl[i]; NULL = class strings */
    int64_t       *ivl_at;       /* per read: its first entry */
    int32_t       *ivl_n;        /* per read: its number of intervals */
    cpg_upre      *upre;         /* recorded task values of the unreliable pass, one per interval its sweeps visit */
    unsigned long long *upre_cursor;
    int64_t        upre_cap;
    unsigned long long *phase_cycles;   /* [4] summed per-warp cycles of the three phases (+ idle at the CTA barriers) */
  };

struct ReadRec { int64_t off, hoff, uoff; int32_t N, M, ncand, mcap, nf, pad; };

/* Each kind of kernel has its own region of the scratch arena, laid out for what it uses:
   SM_WALL  k_wall_b: flag bytes, slot indices, probability slots, E-intervals, touch log
   SM_REL   k_rel:    DP working copies, back pointers, path strings
   SM_UNREL k_unrel_b: list of the intervals the sweeps visit, their keys and sweep order
   SM_FULL  k_classify (retry launch): all of it, worst-case capacities */
enum { SM_WALL = 0, SM_REL = 1, SM_UNREL = 2, SM_FULL = 3 };
#define N_OFF 18
struct ScratchDev
  { uint8_t *base;
    size_t   stride;             /* bytes per lane group */
    int32_t  mode;
    int32_t  P;                  /* longest profile the layout is sized for */
    int32_t  MC;                 /* reliable-interval capacity */
    int32_t  capS, capE, capI;   /* probability slots, E-intervals, intervals (cpg_common.h: cpg_scratch) */
    int32_t  capT, capC;         /* touch log, candidate records (SM_FULL only) */
  };

static inline __host__ __device__ size_t align_up(size_t x, size_t a) { return (x+a-1)/a*a; }

/* layout of one lane group's scratch */
__host__ __device__ static inline size_t scratch_layout(const ScratchDev &SC, size_t off[N_OFF])
{ const int P = SC.P, MC = SC.MC, md = SC.mode;
  const bool w = (md == SM_WALL || md == SM_FULL), r = (md == SM_REL || md == SM_FULL), u = (md == SM_UNREL || md == SM_FULL);
  const bool f = (md == SM_FULL);
  size_t o = 0;
  off[0]  = o; o = align_up(o+(w ? (size_t)(P+2+32) : 0),16);                      /* mark  */
  off[13] = o; o = align_up(o+(w ? sizeof(uint16_t)*(size_t)(P+2) : 0),16);        /* slot  */
  off[1]  = o; o = align_up(o+(w ? sizeof(double)*4*(size_t)SC.capS : 0),16);      /* perr  */
  off[2]  = o; o = align_up(o+(w ? sizeof(cpg_eintvl)*(size_t)SC.capE : 0),16);    /* eint  */
  off[14] = o; o = align_up(o+(w ? sizeof(int32_t)*(size_t)SC.capT : 0),16);       /* tlog  */
  off[3]  = o; o = align_up(o+(f ? sizeof(cpg_intvl)*(size_t)SC.capI : 0),16);     /* intvl */
  off[4]  = o; o = align_up(o+(f ? sizeof(cpg_intvl)*(size_t)MC : 0),16);          /* rint  */
  off[15] = o; o = align_up(o+(f ? sizeof(cpg_chdr)*(size_t)SC.capC : 0),16);      /* hdr   */
  off[16] = o; o = align_up(o+(f ? sizeof(cpg_cbig)*(size_t)SC.capC : 0),16);      /* big   */
  off[5]  = o; o = align_up(o+(r ? sizeof(cpg_intvl)*2*(size_t)MC : 0),16);        /* wint (fw, bw) */
  off[6]  = o; o = align_up(o+(r ? sizeof(uint16_t)*2*(size_t)MC : 0),16);         /* bp   (fw, bw) */
  off[7]  = o; o = align_up(o+(r ? (size_t)MC : 0),16);                            /* asg_f */
  off[8]  = o; o = align_up(o+(r ? (size_t)MC : 0),16);                            /* asg_b */
  off[9]  = o; o = align_up(o+(r ? 2*(size_t)MC : 0),16);                          /* rpos (fw, bw) */
  off[10] = o; o = align_up(o+(u ? sizeof(int32_t)*(size_t)SC.capI : 0),16);       /* ord   */
  off[11] = o; o = align_up(o+(u ? sizeof(int32_t)*(size_t)SC.capI : 0),16);       /* srt   */
  off[17] = o; o = align_up(o+(u ? sizeof(uint32_t)*(size_t)SC.capI : 0),16);      /* key   */
  off[12] = o; o = align_up(o+(f ? sizeof(cpg_upre)*(size_t)SC.capI : 0),16);      /* upre  */
  return align_up(o,256);
}

/* Capacities of the tables.  Worst case (SM_FULL): every position its own interval.  Main
   launches: one interval per 16 positions (HiFi profiles need about one per 70), one probability
   slot per 32 (they need one per 500); a read that needs more is flagged and classified again by
   the retry launch. */
static void scratch_caps(ScratchDev *SC, int P, int K, int mode)
{ SC->P = P; SC->MC = P/K+8; SC->mode = mode;
  if (mode == SM_FULL) { SC->capS = SC->capE = SC->capI = SC->capC = P+2; SC->capT = 3*(P+2); }
  else
    { int div = 16;
      { const char *f = getenv("CPG_SCRATCH_DIV"); if (f && atoi(f) > 0) div = atoi(f); }   /* test knob: force retries */
      SC->capI = P/div+64; SC->capE = P/div+64; SC->capS = P/(2*div)+64; SC->capT = P/div+192; SC->capC = 0;
      if (SC->capI > P+2) SC->capI = P+2;
      if (SC->capE > P+2) SC->capE = P+2;
      if (SC->capS > P+2) SC->capS = P+2;
    }
}

__device__ __forceinline__ void bind_scratch(ReadCtx &R, uint8_t *sb, const size_t off[N_OFF], const ScratchDev &SC)
{ R.S.mark  = sb+off[0];
  R.S.slot  = reinterpret_cast<uint16_t *>(sb+off[13]);
  R.S.perr  = reinterpret_cast<double *>(sb+off[1]);
  R.S.eint  = reinterpret_cast<cpg_eintvl *>(sb+off[2]);
  R.S.tlog  = reinterpret_cast<int32_t *>(sb+off[14]);
  R.S.intvl = reinterpret_cast<cpg_intvl *>(sb+off[3]);
  R.S.rint  = reinterpret_cast<cpg_intvl *>(sb+off[4]);
  R.S.hdr   = reinterpret_cast<cpg_chdr *>(sb+off[15]);
  R.S.big   = reinterpret_cast<cpg_cbig *>(sb+off[16]);
  R.S.wint  = reinterpret_cast<cpg_intvl *>(sb+off[5]);
  R.S.bp    = reinterpret_cast<uint16_t *>(sb+off[6]);
  R.S.asg_f = sb+off[7];
  R.S.asg_b = sb+off[8];
  R.S.rpos  = sb+off[9];
  R.S.ord   = reinterpret_cast<int32_t *>(sb+off[10]);
  R.S.srt   = reinterpret_cast<int32_t *>(sb+off[11]);
  R.S.key   = reinterpret_cast<uint32_t *>(sb+off[17]);
  R.S.MC = SC.MC; R.S.capS = SC.capS; R.S.capE = SC.capE; R.S.capI = SC.capI; R.S.capT = SC.capT; R.S.capC = SC.capC;
  R.S.upre  = reinterpret_cast<cpg_upre *>(sb+off[12]);
  R.hdr = 0; R.big = 0; R.ncand = 0; R.ntlog = 0; R.prune = 0;
}

__device__ __forceinline__ int next_read(int32_t *counter, int lane)
{ int r = 0;
  if (lane == 0) r = atomicAdd(counter,1);
  return __shfl_sync(0xffffffffu,r,0);
}

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(DECODE_THREADS,DECODE_MIN_BLOCKS)
k_decode(BatchDev B, int K, int rcov)
{ __shared__ unsigned s_tab[DECODE_THREADS/32][DC_SLOTS];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (;;)
    { int q = next_read(B.queue+0,lane);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      const int64_t po = B.prof_off[r];
      const int64_t len = B.prof_off[r+1]-po;
      const int cap = (B.rlen[r] >= K) ? B.rlen[r]-K+1 : 0;       /* reads shorter than K have empty profiles */
      const int64_t co = B.cnt_off[r];
      int n = decode_profile(B.prof+po,len,B.cnt+co,cap,lane,s_tab[wib],B.cand ? B.cand+(co >> 5) : 0,rcov);
      if (lane == 0)
        { B.plen[r] = n;
          B.status[r] = (n == cap) ? CPG_ST_OK : CPG_ST_BAD_PROFILE;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Lanes per read in k_classify: a read is owned by an aligned group of CPG_GROUP lanes, so a warp
   works on 32/CPG_GROUP reads whose (mostly serial, latency-bound) instruction streams interleave;
   the lane-parallel task lists of the per-read code are at most 23 long and mostly shorter.
   Registers cap the kernel at 32 warps per SM, so narrower groups are the way to more independent
   streams: measured on the 100 Mb workload, k_classify takes 497 / 343 / 309 ms with groups of
   32 / 16 / 8 lanes (profiles/r01_history.md). */
#ifndef CPG_GROUP
#define CPG_GROUP 8
#endif
#define CLASSIFY_GROUPS (CLASSIFY_THREADS/CPG_GROUP)
/* CTA-synchronous phases; -DCPG_NO_PHASE_SYNC lets every group run ahead */
#ifdef CPG_NO_PHASE_SYNC
#define CPG_PHASE_SYNC() do { } while (0)
#else
#define CPG_PHASE_SYNC() __syncthreads()
#endif

/* the count-threshold table sits in shared memory unless the per-group blocks need the room
   (groups of 8 lanes: 128 reads per CTA); then it is read through the read-only global path */
#if CPG_GROUP >= 16
#define CPG_CTHRES_SMEM 1
#else
#define CPG_CTHRES_SMEM 0
#endif

struct ClassifyShared
  {
#if CPG_CTHRES_SMEM
    uint8_t     cthres[CPG_LROWS*256*4];
#endif
    cpg_dmodel  model;
    cpg_wshared ws[CLASSIFY_GROUPS];
    RelShared   rel[CLASSIFY_GROUPS][2];
  };

/* retry = 0: every read of the batch, compact scratch blocks.  retry = 1: the reads the first launch
   flagged CPG_ST_RETRY, full-size scratch blocks (a handful of CTAs; normally finds nothing). */
__global__ void __launch_bounds__(CLASSIFY_THREADS,CLASSIFY_MIN_BLOCKS)
k_classify(BatchDev B, cpg_dmodel M, ScratchDev SC, int retry)
{ extern __shared__ __align__(16) unsigned char smem_raw[];
  ClassifyShared &sh = *reinterpret_cast<ClassifyShared *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int gib = threadIdx.x/CPG_GROUP;              /* group in the CTA */
  const int glane = threadIdx.x & (CPG_GROUP-1);
  const int gbase = lane-glane;
  const unsigned gmask = ((CPG_GROUP >= 32) ? 0xffffffffu : ((1u << CPG_GROUP)-1u)) << gbase;

#if CPG_CTHRES_SMEM
  for (int i = threadIdx.x; i < CPG_LROWS*256; i += blockDim.x)
    reinterpret_cast<uint32_t *>(sh.cthres)[i] = reinterpret_cast<const uint32_t *>(M.cthres)[i];
  const uint8_t *cthres = sh.cthres;
#else
  const uint8_t *cthres = M.cthres;
#endif
  if (threadIdx.x == 0) sh.model = M;
  __syncthreads();
  cpg_model_fill_logs(&sh.model,threadIdx.x,blockDim.x);
  __syncthreads();

  const size_t gg = (size_t)blockIdx.x*CLASSIFY_GROUPS+gib;
  uint8_t *sb = SC.base+gg*SC.stride;
  size_t off[N_OFF];
  scratch_layout(SC,off);

  /* One read per lane group, CLASSIFY_GROUPS reads per CTA at a time, taken from the queue in
     processing order (neighbouring reads have similar lengths, so the phases of a CTA finish
     close together). */
  __shared__ int s_base;
  if (retry && B.queue[3] == 0) return;                /* nothing was flagged (the usual case) */
  for (;;)
    { __syncthreads();
      if (threadIdx.x == 0) s_base = atomicAdd(B.queue+1+retry,CLASSIFY_GROUPS);
      __syncthreads();
      const int base = s_base;
      if (base >= B.n_reads) break;
      const int q = base+gib;
      int active = (q < B.n_reads);
      int r = 0, rlen = 0, plen = 0;
      if (active)
        { r = B.order[q];
          rlen = B.rlen[r]; plen = rlen-M.K+1;
          const int st0 = B.status[r];
          if (retry) { if (!(st0 & CPG_ST_RETRY)) active = 0; }
          else if (st0 != CPG_ST_OK) active = 0;             /* undecodable profile: left to the host */
          if (!active) { }
          else if (plen > SC.P) { if (glane == 0) B.status[r] = CPG_ST_BAD_PROFILE; active = 0; }
        }
      WCtx W;
      W.lane = lane; W.M = &sh.model; W.cthres = cthres; W.ws = &sh.ws[gib]; W.status = 0;
      W.glane = glane; W.gsize = CPG_GROUP; W.gbase = gbase; W.gmask = gmask;
      ReadCtx R;
      R.prof = B.cnt+(active ? B.cnt_off[r] : 0); R.plen = plen; R.rlen = rlen;
      R.seq.p = B.seq+(active ? B.seq_off[r] : 0); R.seq.bits = B.seq_bits;
      R.nslots = 0; R.N = 0; R.M = 0;
      R.cand    = B.cand+(active ? (B.cnt_off[r] >> 5) : 0);
      bind_scratch(R,sb,off,SC);

      long long t0 = clock64();
      if (active) classify_phase1(R,W);
      long long t1 = clock64();
      CPG_PHASE_SYNC();
      long long t2 = clock64();
      if (active) classify_phase2(R,W,sh.rel[gib]);
      long long t3 = clock64();
      CPG_PHASE_SYNC();
      long long t4 = clock64();
      if (active)
        { int st = classify_phase3(R,W,B.cls+B.cls_off[r]);
          st = __reduce_or_sync(gmask,st);
          if (B.ivl != 0 && !(st & CPG_ST_RETRY))
            { /* compact results: this read's table goes into the pool, where k_pack looks for it */
              long long at = 0;
              if (glane == 0) at = (long long)atomicAdd(B.pool_cursor,(unsigned long long)R.N);
              at = __shfl_sync(gmask,at,gbase);
              if (at+R.N > B.pool_cap) st |= CPG_ST_RETRY;            /* reported as an internal error */
              else
                { uint4 *dst = reinterpret_cast<uint4 *>(B.pool+at);
                  const uint4 *src = reinterpret_cast<const uint4 *>(R.S.intvl);
                  for (int i = glane; i < 3*R.N; i += CPG_GROUP) dst[i] = src[i];
                  if (glane == 0) { ReadRec rc = B.rec[r]; rc.off = at; rc.N = R.N; B.rec[r] = rc; }
                }
            }
          if (glane == 0)
            { B.status[r] = st;
              if (st & CPG_ST_RETRY) atomicAdd(B.queue+3,1);
            }
        }
      long long t5 = clock64();
      if (glane == 0 && B.phase_cycles)
        { atomicAdd(B.phase_cycles+0,(unsigned long long)(t1-t0));
          atomicAdd(B.phase_cycles+1,(unsigned long long)(t3-t2));
          atomicAdd(B.phase_cycles+2,(unsigned long long)(t5-t4));
          atomicAdd(B.phase_cycles+3,(unsigned long long)((t2-t1)+(t4-t3)));
        }
    }
}

/* ------------------------------------------------------------------------------------------
 *  The main path: one kernel per phase.  The per-read code is large and branchy and the phases
 *  share almost none of it, so each phase as its own persistent kernel keeps the instruction
 *  working set of an SM small without the CTA-wide barriers k_classify needs for the same effect,
 *  lets every lane group pull its next read on its own, and gives each phase its own register
 *  allocation and its own lane mapping (candidate / read / interval per lane).  Between the
 *  kernels a read lives in HBM as its candidate records, then as its interval table in the
 *  batch's pool (one atomicAdd per read).  k_classify above stays as the retry path (reads that
 *  outgrow the compact scratch, the record arrays or the pool).
 * ------------------------------------------------------------------------------------------ */
/* Lanes per read of each phase kernel (powers of two <= 32); profiles/README.md has the sweeps. */
#ifndef WALLB_GROUP
#define WALLB_GROUP 2
#endif
#ifndef WALLC_GROUP
#define WALLC_GROUP 32
#endif
#ifndef REL_GROUP
#define REL_GROUP   CPG_GROUP
#endif
#ifndef UNREL_GROUP
#define UNREL_GROUP CPG_GROUP
#endif
/* CTA shape of the phase kernels: they have no CTA-wide barrier, so the shape only decides how many
   warps fit an SM through the register cap (PHASE_THREADS x PHASE_MIN_BLOCKS threads per SM) */
#ifndef PHASE_THREADS
#define PHASE_THREADS    CLASSIFY_THREADS
#endif
#ifndef PHASE_MIN_BLOCKS
#define PHASE_MIN_BLOCKS CLASSIFY_MIN_BLOCKS
#endif
#ifndef WALLA_THREADS
#define WALLA_THREADS    256
#endif
#ifndef WALLA_MIN_BLOCKS
#define WALLA_MIN_BLOCKS 2
#endif
#define WALLA_QCAP (1024+32)     /* candidate positions queued per warp: one 1024-position chunk plus a remainder */

template<int G, bool CTHRES> struct PhaseShared
  { uint8_t     cthres[CTHRES ? CPG_LROWS*256*4 : 16];        /* the count-threshold table */
    cpg_dmodel  model;
    cpg_wshared ws[PHASE_THREADS/G];
  };
template<int G> struct RelPhaseShared
  { cpg_dmodel  model;
    cpg_wshared ws[PHASE_THREADS/G];
    RelShared   rel[PHASE_THREADS/G][2];
  };
struct WaTask;
struct WallAShared;

struct GroupId { int lane, gib, glane, gbase, gsize; unsigned gmask; };
template<int G> __device__ __forceinline__ GroupId group_id()
{ GroupId g;
  g.lane = threadIdx.x & 31; g.gib = threadIdx.x/G; g.glane = threadIdx.x & (G-1);
  g.gbase = g.lane-g.glane; g.gsize = G;
  g.gmask = ((G >= 32) ? 0xffffffffu : ((1u << G)-1u)) << g.gbase;
  return g;
}
/* next position of the processing order for this lane group */
__device__ __forceinline__ int group_next(int32_t *counter, const GroupId &g)
{ int q = 0;
  if (g.glane == 0) q = atomicAdd(counter,1);
  return __shfl_sync(g.gmask,q,g.gbase);
}
__device__ __forceinline__ void init_wctx(WCtx &W, const GroupId &g, const cpg_dmodel *M, const uint8_t *cthres, cpg_wshared *ws, int status)
{ W.lane = g.lane; W.M = M; W.cthres = cthres; W.ws = ws; W.status = status;
  W.glane = g.glane; W.gsize = g.gsize; W.gbase = g.gbase; W.gmask = g.gmask;
}

/* phase 1a: the pure step of wall detection, one wall candidate per lane (cpg_wall.cuh, wa_).
   A warp takes a read, streams its candidate bit map (1024 positions per step), queues the candidate
   positions in shared memory and works them off 32 at a time, so that the lanes stay full although
   only one position in ~80 is a candidate.  That is stage 0 (context + count thresholds, cheap).  The
   candidates that get past it (one in ten) need the probabilities (wa_tasks, expensive): the lanes of a
   warp wait for each other (SIMT), so these are queued a second time -- across reads -- and evaluated
   32 at a time as well.  Reads: 2 counts, the bases around the k-mer's end, the threshold table (shared
   memory) per candidate; writes: a header per candidate, in position order, a record per queued one. */
struct WaTask { int32_t r, pos; uint32_t idx, packed, cnts; };     /* packed: info | t << 8 | l << 12; cnts: cout | cin << 16 */
#define WALLA_TCAP 64
#ifndef WALLA_PRUNE
#define WALLA_PRUNE 1            /* wa_tasks: no partner values for an error type whose own probability is below the threshold */
#endif
struct WallAShared
  { uint8_t     cthres[CPG_LROWS*256*4];
    cpg_dmodel  model;
    int32_t     q[WALLA_THREADS/32][WALLA_QCAP];
    WaTask      tq[WALLA_THREADS/32][WALLA_TCAP];
  };

__device__ __forceinline__ void wa_run_tasks(const BatchDev &B, const WCtx &W, const WaTask *tq, int n, int lane)
{ if (lane < n)
    { const WaTask T = tq[lane];
      const int rlen = B.rlen[T.r], plen = rlen-W.M->K+1;
      cpg_seq seq; seq.p = B.seq+B.seq_off[T.r]; seq.bits = B.seq_bits;
      WaCand C;
      const unsigned info = T.packed & 0xffu;
      C.wtype = (info & CH_GAIN) ? WT_GAIN : WT_DROP;
      C.t = (T.packed >> 8) & 0xf; C.l = (T.packed >> 12) & 0xff;
      C.cout = (uint16_t)(T.cnts & 0xffffu); C.cin = (uint16_t)(T.cnts >> 16);
      C.cng = (int)C.cout-(int)C.cin; C.erate = W.M->pe[C.t][C.l];
      wa_tasks(B.cnt+B.cnt_off[T.r],plen,seq,rlen,W,T.pos,C,info,B.big+T.idx,WALLA_PRUNE);
    }
}

__global__ void __launch_bounds__(WALLA_THREADS,WALLA_MIN_BLOCKS)
k_wall_a(BatchDev B, cpg_dmodel M)
{ extern __shared__ __align__(16) unsigned char smem_raw[];
  WallAShared &sh = *reinterpret_cast<WallAShared *>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < CPG_LROWS*256; i += blockDim.x)
    reinterpret_cast<uint32_t *>(sh.cthres)[i] = reinterpret_cast<const uint32_t *>(M.cthres)[i];
  if (threadIdx.x == 0) sh.model = M;
  __syncthreads();
  cpg_model_fill_logs(&sh.model,threadIdx.x,blockDim.x);
  __syncthreads();
  int32_t *pq = sh.q[wib];
  WaTask *tq = sh.tq[wib];
  int ntq = 0;
  WCtx W;
  W.lane = lane; W.M = &sh.model; W.cthres = sh.cthres; W.ws = 0; W.status = 0;
  W.glane = 0; W.gsize = 1; W.gbase = lane; W.gmask = 1u << lane;
  const unsigned lt = (1u << lane)-1u;
  for (;;)
    { const int q = next_read(B.queue+1,lane);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      ReadRec rc; rc.off = 0; rc.hoff = 0; rc.uoff = 0; rc.N = 0; rc.M = 0; rc.ncand = 0; rc.mcap = 0; rc.nf = 0; rc.pad = 0;
      const int rlen = B.rlen[r], plen = rlen-M.K+1;
      if (B.status[r] != CPG_ST_OK) { if (lane == 0) B.rec[r] = rc; continue; }     /* undecodable profile: left to the host */
      const uint16_t *prof = B.cnt+B.cnt_off[r];
      const uint32_t *cand = B.cand+(B.cnt_off[r] >> 5);
      cpg_seq seq; seq.p = B.seq+B.seq_off[r]; seq.bits = B.seq_bits;
      const int nwords = (plen+31) >> 5;
      const unsigned tail = (plen & 31) ? ((1u << (plen & 31))-1u) : 0xffffffffu;
      int ncand = 0;
      for (int w = lane; w < nwords; w += 32) ncand += __popc(cand[w] & (w == nwords-1 ? tail : 0xffffffffu));
      ncand = __reduce_add_sync(0xffffffffu,ncand);
      long long hoff = 0;
      if (lane == 0) hoff = (long long)atomicAdd(B.hdr_cursor,(unsigned long long)ncand);
      hoff = __shfl_sync(0xffffffffu,hoff,0);
      if (hoff+ncand > B.hdr_cap)
        { if (lane == 0) { B.rec[r] = rc; B.status[r] = CPG_ST_RETRY; atomicAdd(B.queue+3,1); }
          continue;
        }
      cpg_chdr *hdr = B.hdr+hoff;
      int npend = 0, done = 0, overflow = 0;
      for (int wb = 0; wb < nwords || npend > 0; wb += 32)
        { if (wb < nwords)
            { const int w = wb+lane;
              unsigned cw = (w < nwords) ? (cand[w] & (w == nwords-1 ? tail : 0xffffffffu)) : 0u;
              int incl = __popc(cw);
              const int mine = incl;
              #pragma unroll
              for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu,incl,d); if (lane >= d) incl += t; }
              const int total = __shfl_sync(0xffffffffu,incl,31);
              int k = npend+incl-mine;
              while (cw) { const int bit = __ffs((int)cw)-1; cw &= cw-1; pq[k++] = (w << 5)+bit; }
              __syncwarp();
              npend += total;
            }
          const int flush = (wb+32 >= nwords);          /* last chunk: work off the remainder too */
          int head = 0;
          while (npend-head >= 32 || (flush && npend-head > 0))
            { const int nact = min(32,npend-head);
              const int active = lane < nact;
              const int pos = active ? pq[head+lane] : 0;
              unsigned info = 0; WaCand C; C.t = 0; C.l = 0; C.cout = 0; C.cin = 0;
              if (active) info = wa_stage0(prof,seq,rlen,W,pos,C);
              int need = active && (info & (CH_REACH_S|CH_REACH_O));
              const unsigned m = __ballot_sync(0xffffffffu,need);
              unsigned long long bb = 0;
              if (lane == 0 && m) bb = atomicAdd(B.big_cursor,(unsigned long long)__popc(m));
              bb = __shfl_sync(0xffffffffu,bb,0);
              const unsigned long long idx = bb+(unsigned)__popc(m & lt);
              if (need && (long long)idx >= B.big_cap) { overflow = 1; need = 0; }
              if (active)
                { cpg_chdr H; H.pos = pos; H.info = info; H.big = (uint32_t)idx; H.pad = 0;
                  *reinterpret_cast<uint4 *>(hdr+done+lane) = *reinterpret_cast<const uint4 *>(&H);
                }
              /* second queue: the expensive part, for full warps of candidates that need it */
              const unsigned m2 = __ballot_sync(0xffffffffu,need);
              if (need)
                { WaTask T; T.r = r; T.pos = pos; T.idx = (uint32_t)idx;
                  T.packed = info | ((unsigned)C.t << 8) | ((unsigned)C.l << 12);
                  T.cnts = (unsigned)C.cout | ((unsigned)C.cin << 16);
                  tq[ntq+__popc(m2 & lt)] = T;
                }
              __syncwarp();
              ntq += __popc(m2);
              if (ntq >= 32)
                { wa_run_tasks(B,W,tq,32,lane);
                  __syncwarp();
                  const int left = ntq-32;
                  WaTask T; if (lane < left) T = tq[32+lane];
                  __syncwarp();
                  if (lane < left) tq[lane] = T;
                  __syncwarp();
                  ntq = left;
                }
              done += nact; head += nact;
            }
          /* bring the remainder to the front of the queue */
          const int left = npend-head;
          if (head > 0 && left > 0)
            { const int v = (lane < left) ? pq[head+lane] : 0;
              __syncwarp();
              if (lane < left) pq[lane] = v;
            }
          __syncwarp();
          npend = left;
          if (flush) break;
        }
      overflow = __any_sync(0xffffffffu,overflow);
      if (lane == 0)
        { rc.hoff = hoff; rc.ncand = ncand;
          B.rec[r] = rc;
          if (overflow) { B.status[r] = CPG_ST_RETRY; atomicAdd(B.queue+3,1); }
        }
    }
  if (ntq > 0) wa_run_tasks(B,W,tq,ntq,lane);
}

/* phase 1b: the order-dependent replay of find_wall for one read per lane group (cpg_wall.cuh, wb_):
   candidate records in, interval table (without counts) out, straight into the pool. */
__global__ void __launch_bounds__(PHASE_THREADS,PHASE_MIN_BLOCKS)
k_wall_b(BatchDev B, cpg_dmodel M, ScratchDev SC)
{ constexpr int G = WALLB_GROUP;
  __shared__ cpg_dmodel s_model;
  const GroupId g = group_id<G>();
  if (threadIdx.x == 0) s_model = M;
  __syncthreads();
  uint8_t *sb = SC.base+((size_t)blockIdx.x*(PHASE_THREADS/G)+g.gib)*SC.stride;
  size_t off[N_OFF];
  scratch_layout(SC,off);
  for (;;)
    { const int q = group_next(B.queue+6,g);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      if (B.status[r] != CPG_ST_OK) continue;                 /* undecodable profile, or flagged for the retry launch */
      const int rlen = B.rlen[r], plen = rlen-M.K+1;
      if (plen > SC.P) { if (g.glane == 0) B.status[r] = CPG_ST_BAD_PROFILE; continue; }
      ReadRec rc = B.rec[r];
      WCtx W; init_wctx(W,g,&s_model,M.cthres,0,0);
      ReadCtx R;
      R.prof = 0; R.plen = plen; R.rlen = rlen; R.seq.p = 0; R.seq.bits = B.seq_bits; R.cand = 0;
      R.nslots = 0; R.N = 0; R.M = 0;
      bind_scratch(R,sb,off,SC);
      R.hdr = B.hdr+rc.hoff; R.big = B.big; R.ncand = rc.ncand;
      const int NS = wb_walls(R,W);
      int N = 0, mcap = 0;
      long long at = 0;
      if (!(W.status & CPG_ST_ABORT))
        { N = wb_cuts(R,W,NS,0,0,&mcap);
          /* the per-interval arrays of k_unrel_b's scratch blocks hold capI entries */
          if (N > SC.capI) W.status |= CPG_ST_RETRY;
          else
            { /* a place in the pool for intvl[N] and the (at most mcap) reliable ones behind them */
              if (g.glane == 0) at = (long long)atomicAdd(B.pool_cursor,(unsigned long long)(N+mcap));
              at = __shfl_sync(g.gmask,at,g.gbase);
              if (at+N+mcap > B.pool_cap) W.status |= CPG_ST_RETRY;
              else wb_cuts(R,W,NS,B.pool+at,N,0);
            }
        }
      wb_clean(R,W);
      const int st = __reduce_or_sync(g.gmask,W.status);
      if (g.glane == 0)
        { rc.off = at; rc.N = (st & CPG_ST_ABORT) ? 0 : N; rc.M = 0; rc.mcap = mcap;
          B.rec[r] = rc;
          if (st) B.status[r] = st;
          if (st & CPG_ST_RETRY) atomicAdd(B.queue+3,1);
        }
      __syncwarp(g.gmask);
    }
}

/* phase 1c: one interval per lane: end counts, corrected counts, plausibility (cpg_wall.cuh, wc_);
   the reliable intervals of a read go behind its table, in order. */
__global__ void __launch_bounds__(PHASE_THREADS,PHASE_MIN_BLOCKS)
k_wall_c(BatchDev B, cpg_dmodel M)
{ constexpr int G = WALLC_GROUP;
  __shared__ cpg_dmodel s_model;
  const GroupId g = group_id<G>();
  if (threadIdx.x == 0) s_model = M;
  __syncthreads();
  for (;;)
    { const int q = group_next(B.queue+7,g);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      const int st0 = B.status[r];
      if (st0 & (CPG_ST_BAD_PROFILE|CPG_ST_ABORT)) continue;
      ReadRec rc = B.rec[r];
      if (rc.N == 0) continue;
      const int rlen = B.rlen[r], plen = rlen-M.K+1;
      WCtx W; init_wctx(W,g,&s_model,M.cthres,0,0);
      cpg_seq seq; seq.p = B.seq+B.seq_off[r]; seq.bits = B.seq_bits;
      cpg_intvl *v = B.pool+rc.off;
      const int Mrel = wc_read(B.cnt+B.cnt_off[r],plen,seq,rlen,W,v,rc.N,v+rc.N);
      const int st = __reduce_or_sync(g.gmask,W.status);
      if (g.glane == 0)
        { B.rec[r].M = Mrel;
          if (st) B.status[r] = st0 | st;
        }
      __syncwarp(g.gmask);
    }
}

/* phase 2: forward/backward DP over the reliable intervals (cpg_rel.cuh).  The lane groups of a warp run
   the same code on different reads, so the warp executes them together wherever their control flow agrees
   (SIMT): an explicit warp-wide barrier per DP step changed neither the instruction count nor the active
   lanes per instruction (profiles/r02_rel_lockstep.md). */
__global__ void __launch_bounds__(PHASE_THREADS,PHASE_MIN_BLOCKS)
k_rel(BatchDev B, cpg_dmodel M, ScratchDev SC)
{ constexpr int G = REL_GROUP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RelPhaseShared<G> &sh = *reinterpret_cast<RelPhaseShared<G> *>(smem_raw);
  const GroupId g = group_id<G>();
  if (threadIdx.x == 0) sh.model = M;
  __syncthreads();
  cpg_model_fill_logs(&sh.model,threadIdx.x,blockDim.x);
  __syncthreads();
  uint8_t *sb = SC.base+((size_t)blockIdx.x*(PHASE_THREADS/G)+g.gib)*SC.stride;
  size_t off[N_OFF];
  scratch_layout(SC,off);
  for (;;)
    { const int q = group_next(B.queue+4,g);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      const int st0 = B.status[r];
      if (st0 & (CPG_ST_BAD_PROFILE|CPG_ST_ABORT)) continue;
      const ReadRec rc = B.rec[r];
      if (rc.M == 0) continue;
      WCtx W; init_wctx(W,g,&sh.model,M.cthres,&sh.ws[g.gib],0);
      ReadCtx R;
      const int rlen = B.rlen[r];
      R.prof = B.cnt+B.cnt_off[r]; R.plen = rlen-M.K+1; R.rlen = rlen;
      R.seq.p = B.seq+B.seq_off[r]; R.seq.bits = B.seq_bits; R.cand = 0;
      R.nslots = 0; R.N = rc.N; R.M = rc.M;
      bind_scratch(R,sb,off,SC);
      R.S.intvl = B.pool+rc.off; R.S.rint = B.pool+rc.off+rc.N;
      classify_reliable(R,W,sh.rel[g.gib]);
      const int st = __reduce_or_sync(g.gmask,W.status);
      if (g.glane == 0 && st != 0) B.status[r] = st0 | st;
      __syncwarp(g.gmask);
    }
}

/* phase 3a: the pure step of the unreliable pass, one interval per lane (cpg_unrel.cuh, un_pre_interval):
   a warp takes a read, queues the intervals its sweeps will visit (the ones not fixed by the DP, about a
   third) and works them off 32 at a time: nearest reliable H / D neighbours and the ten task values. */
#define UNRELA_QCAP 64
__global__ void __launch_bounds__(WALLA_THREADS,WALLA_MIN_BLOCKS)
k_unrel_a(BatchDev B, cpg_dmodel M)
{ __shared__ cpg_dmodel s_model;
  __shared__ int32_t s_q[WALLA_THREADS/32][UNRELA_QCAP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_model = M;
  __syncthreads();
  cpg_model_fill_logs(&s_model,threadIdx.x,blockDim.x);
  __syncthreads();
  int32_t *pq = s_q[wib];
  WCtx W;
  W.lane = lane; W.M = &s_model; W.cthres = M.cthres; W.ws = 0; W.status = 0;
  W.glane = 0; W.gsize = 1; W.gbase = lane; W.gmask = 1u << lane;
  const unsigned lt = (1u << lane)-1u;
  for (;;)
    { const int q = next_read(B.queue+5,lane);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      const int st0 = B.status[r];
      if (st0 & (CPG_ST_BAD_PROFILE|CPG_ST_RETRY|CPG_ST_EINTVL_OVF)) continue;
      ReadRec rc = B.rec[r];
      const int N = rc.N;
      const cpg_intvl *v = B.pool+rc.off;
      int nf = 0;
      for (int i = lane; i < N; i += 32) nf += !un_is_fixed(v[i]);
      nf = __reduce_add_sync(0xffffffffu,nf);
      long long uoff = 0;
      if (lane == 0) uoff = (long long)atomicAdd(B.upre_cursor,(unsigned long long)nf);
      uoff = __shfl_sync(0xffffffffu,uoff,0);
      if (uoff+nf > B.upre_cap)
        { if (lane == 0) { B.status[r] = st0 | CPG_ST_RETRY; atomicAdd(B.queue+3,1); }
          continue;
        }
      if (lane == 0) { B.rec[r].uoff = uoff; B.rec[r].nf = nf; }
      cpg_upre *U = B.upre+uoff;
      int npend = 0, done = 0;
      for (int base = 0; base < N || npend > 0; base += 32)
        { if (base < N)
            { const int i = base+lane;
              const int open = (i < N) && !un_is_fixed(v[i]);
              const unsigned m = __ballot_sync(0xffffffffu,open);
              if (open) pq[npend+__popc(m & lt)] = i;
              __syncwarp();
              npend += __popc(m);
            }
          const int flush = (base+32 >= N);
          int head = 0;
          while (npend-head >= 32 || (flush && npend-head > 0))
            { const int nact = min(32,npend-head);
              if (lane < nact) un_pre_interval(W,v,N,pq[head+lane],U+done+lane);
              done += nact; head += nact;
            }
          const int left = npend-head;
          if (head > 0 && left > 0)
            { const int x = (lane < left) ? pq[head+lane] : 0;
              __syncwarp();
              if (lane < left) pq[lane] = x;
            }
          __syncwarp();
          npend = left;
          if (flush) break;
        }
    }
}

/* phase 3b: the sweeps of the unreliable pass on the recorded values (cpg_unrel.cuh: un_sweeps) */
__global__ void __launch_bounds__(PHASE_THREADS,PHASE_MIN_BLOCKS)
k_unrel_b(BatchDev B, cpg_dmodel M, ScratchDev SC)
{ constexpr int G = UNREL_GROUP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PhaseShared<G,false> &sh = *reinterpret_cast<PhaseShared<G,false> *>(smem_raw);
  const GroupId g = group_id<G>();
  if (threadIdx.x == 0) sh.model = M;
  __syncthreads();
  cpg_model_fill_logs(&sh.model,threadIdx.x,blockDim.x);
  __syncthreads();
  uint8_t *sb = SC.base+((size_t)blockIdx.x*(PHASE_THREADS/G)+g.gib)*SC.stride;
  size_t off[N_OFF];
  scratch_layout(SC,off);
  for (;;)
    { const int q = group_next(B.queue+24,g);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      const int st0 = B.status[r];
      if (st0 & (CPG_ST_BAD_PROFILE|CPG_ST_RETRY)) continue;
      const ReadRec rc = B.rec[r];
      WCtx W; init_wctx(W,g,&sh.model,M.cthres,&sh.ws[g.gib],st0);
      ReadCtx R;
      const int rlen = B.rlen[r];
      R.prof = B.cnt+B.cnt_off[r]; R.plen = rlen-M.K+1; R.rlen = rlen;
      R.seq.p = B.seq+B.seq_off[r]; R.seq.bits = B.seq_bits; R.cand = 0;
      R.nslots = 0; R.N = rc.N; R.M = rc.M;
      bind_scratch(R,sb,off,SC);
      R.S.intvl = B.pool+rc.off; R.S.rint = B.pool+rc.off+rc.N;
      if (!(W.status & CPG_ST_ABORT))
        { const int nf = un_list(R,W);                       /* == rc.nf: same rule, same order as k_unrel_a */
          un_sweeps(R,W,nf,B.upre+rc.uoff);
        }
      const int st = __reduce_or_sync(g.gmask,W.status);
      if (g.glane == 0 && st != st0) B.status[r] = st;
      __syncwarp(g.gmask);
    }
}

/* phase 4: class strings (src/ClassPro.c:114-117,265-271), streaming: one warp per read, 32 intervals
   loaded at a time (one per lane), their stretches written 32 consecutive bytes per instruction.
   (12+5) N bytes in, r bytes out per read. */
__global__ void __launch_bounds__(256)
k_emit(BatchDev B, int K)
{ const int lane = threadIdx.x & 31;
  for (;;)
    { const int q = next_read(B.queue+25,lane);
      if (q >= B.n_reads) break;
      const int r = B.order[q];
      if (B.status[r] & (CPG_ST_BAD_PROFILE|CPG_ST_RETRY)) continue;       /* the host's / the retry launch's */
      const ReadRec rc = B.rec[r];
      uint8_t *out = B.cls+B.cls_off[r];
      for (int j = lane; j < K-1; j += 32) out[j] = 'N';
      out += K-1;
      const cpg_intvl *v = B.pool+rc.off;
      for (int base = 0; base < rc.N; base += 32)
        { const int i = base+lane;
          int b = 0, e = 0, c = '?';
          if (i < rc.N)
            { b = v[i].b; e = v[i].e;
              const int a = v[i].asgn;
              c = (a == ST_E) ? 'E' : (a == ST_R) ? 'R' : (a == ST_H) ? 'H' : (a == ST_D) ? 'D' : '?';
            }
          const int cnt = min(32,rc.N-base);
          for (int k = 0; k < cnt; k++)
            { const int kb = __shfl_sync(0xffffffffu,b,k), ke = __shfl_sync(0xffffffffu,e,k);
              const uint8_t kc = (uint8_t)__shfl_sync(0xffffffffu,c,k);
              for (int j = kb+lane; j < ke; j += 32) out[j] = kc;
            }
        }
    }
}

/* compact results: (end << 3 | class) of every interval, 4 bytes instead of the 48 of the table; 12 N bytes
   in, 4 N out per read */
__global__ void __launch_bounds__(256)
k_pack(BatchDev B)
{ const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x*blockDim.x) >> 5;
  for (int r = (blockIdx.x*blockDim.x+threadIdx.x) >> 5; r < B.n_reads; r += warps)
    { const int st = B.status[r];
      ReadRec rc = B.rec[r];
      if (st & (CPG_ST_BAD_PROFILE|CPG_ST_ABORT)) { rc.N = 0; rc.off = 0; }
      const cpg_intvl *v = B.pool+rc.off;
      for (int i = lane; i < rc.N; i += 32) B.ivl[rc.off+i] = ((uint32_t)v[i].e << 3) | ((uint32_t)v[i].asgn & 7u);
      if (lane == 0) { B.ivl_at[r] = rc.off; B.ivl_n[r] = rc.N; }
    }
}

/* ------------------------------------------------------------------------------------------
 *  prof2class (src/prof2class.c:236-258): a RELATIVE profile -- counts of the read's k-mers in a
 *  genome or haplotype table -- mapped to ground-truth classes, 0 -> E, 1 -> H, 2 -> D, more -> R,
 *  behind K-1 'N's.  One CTA per read at a time; streaming: 2n bytes in, r bytes out.
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
k_count2class(BatchDev B, int K)
{ for (int r = blockIdx.x; r < B.n_reads; r += gridDim.x)
    { const uint16_t *c = B.cnt+B.cnt_off[r];
      uint8_t *o = B.cls+B.cls_off[r];
      const int rlen = B.rlen[r];
      for (int i = threadIdx.x; i < rlen; i += blockDim.x)
        { char ch = 'N';
          if (i >= K-1)
            { const unsigned v = c[i-(K-1)];
              ch = (v == 0) ? 'E' : (v == 1) ? 'H' : (v == 2) ? 'D' : 'R';
            }
          o[i] = (uint8_t)ch;
        }
    }
}

/* ==========================================================================================
 *  C ABI
 * ========================================================================================== */
struct DevBuf { void *p; size_t cap; };

struct Slot
  { cudaStream_t stream;
    DevBuf seq, seq_off, rlen, prof, prof_off, cnt, cnt_off, cand, plen, cls, cls_off, status, order, queue, rec, pool, hdr, big, upre, ivl, ivl_at, ivl_n;
    /* small host-side (pinned) staging for arrays the library computes itself */
    int64_t *h_cnt_off; int32_t *h_order; size_t h_cap;
    int32_t *h_status; size_t h_status_cap;
    unsigned long long *h_cursor;        /* pinned: the pool cursor after the kernels (compact results) */
    int64_t  pool_cap;
    int32_t  n_reads; int64_t cls_bytes; int32_t maxP;
    int      busy;
    cudaEvent_t kdone;           /* recorded after this slot's kernels */
    int      kdone_valid;
    BatchDev B;
  };

struct cpg_ctx
  { int        device;
    cpg_model  model;
    cpg_dmodel dmodel;
    void      *d_cthres, *d_logfact;
    Slot       slot[2];
    /* scratch arenas, one set per slot: the kernels of the two slots run at the same time (the CTAs of a
       slot's next kernel move in as those of the other slot's persistent kernels run out of reads) */
    DevBuf     scratch[2], scratch_big[2];
    ScratchDev SC[2], SCbig[2];                                 /* SC: k_classify as the main launch (CPG_FUSED) */
    ScratchDev SCw[2], SCr[2], SCu[2];                          /* regions of `scratch` for k_wall_b, k_rel, k_unrel */
    int        scratch_P;
    int        retry_blocks;
    int        fused;                 /* CPG_FUSED=1: the single-kernel path (k_classify) for every read */
    int        result_mode;           /* CPG_RESULT_CLASSES / CPG_RESULT_INTERVALS */
    int        walla_blocks, wallb_blocks, wallc_blocks, rel_blocks, unrela_blocks, unrel_blocks;
    size_t     walla_smem, rel_smem, unrel_smem;
    cudaEvent_t evp[7];               /* between the phase kernels */
    uint64_t   phase_ns[4];           /* wall, reliable DP, unreliable + emit, retry launch: last timed run */
    uint64_t   wall_ns[3];            /* k_wall_a, k_wall_b, k_wall_c of that run */
    uint64_t   unrel_ns[3];           /* k_unrel_a, k_unrel_b, k_emit */
    int        n_sm, decode_blocks, classify_blocks;
    size_t     classify_smem;
    cudaEvent_t ev[3];
    char       err[512];
  };

static char g_err[512] = "";

static int set_err(cpg_ctx *c, int code, const char *fmt, ...)
{ va_list ap; va_start(ap,fmt);
  vsnprintf(c ? c->err : g_err,512,fmt,ap);
  va_end(ap);
  return code;
}

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
  return set_err(ctx,CPG_ECUDA,"%s failed: %s (%s:%d)",#call,cudaGetErrorString(e__),__FILE__,__LINE__); } while (0)

static int reserve(cpg_ctx *ctx, DevBuf *b, size_t bytes)
{ if (bytes <= b->cap) return CPG_OK;
  if (b->p) { cudaFree(b->p); b->p = NULL; b->cap = 0; }
  size_t want = bytes+bytes/4+256;
  cudaError_t e = cudaMalloc(&b->p,want);
  if (e != cudaSuccess)
    return set_err(ctx,CPG_ENOMEM,"cudaMalloc(%zu) failed: %s",want,cudaGetErrorString(e));
  b->cap = want;
  return CPG_OK;
}

extern "C" const char *cpg_version(void) { return "classpro_b200 0.1 (sm_100a)"; }

extern "C" const char *cpg_last_error(const cpg_ctx *ctx) { return ctx ? ctx->err : g_err; }

extern "C" int cpg_device_count(void)
{ int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" const char *cpg_status_string(int32_t st)
{ if (st == 0) return "ok";
  if (st & CPG_ST_BAD_PROFILE) return "profile length != read length - K + 1";
  if (st & CPG_ST_EINTVL_OVF)  return "# E-intvls >= plen";
  if (st & CPG_ST_NO_PROB)     return "no valid probability for an interval";
  if (st & CPG_ST_INTERP)      return "invalid points for interpolation";
  if (st & CPG_ST_BINOM)       return "k > n in a binomial";
  if (st & CPG_ST_UNDEF_TRACE) return "all DP states impossible (reference behaviour undefined)";
  if (st & 128)                return "profile[plen] read (reference reads stale memory)";
  if (st & CPG_ST_RETRY)       return "internal error: read left unclassified by the retry launch";
  return "unknown";
}

/* fatal = conditions on which the reference exits */
#define CPG_ST_FATAL (CPG_ST_BAD_PROFILE|CPG_ST_EINTVL_OVF|CPG_ST_NO_PROB|CPG_ST_INTERP|CPG_ST_BINOM|CPG_ST_RETRY)

extern "C" void cpg_destroy(cpg_ctx *ctx)
{ if (ctx == NULL) return;
  cudaSetDevice(ctx->device);
  for (int s = 0; s < 2; s++)
    { Slot *S = &ctx->slot[s];
      DevBuf *bufs[] = { &S->seq,&S->seq_off,&S->rlen,&S->prof,&S->prof_off,&S->cnt,&S->cnt_off,&S->cand,&S->plen,
                         &S->cls,&S->cls_off,&S->status,&S->order,&S->queue,&S->rec,&S->pool,&S->hdr,&S->big,&S->upre,&S->ivl,&S->ivl_at,&S->ivl_n };
      for (unsigned i = 0; i < sizeof(bufs)/sizeof(bufs[0]); i++) if (bufs[i]->p) cudaFree(bufs[i]->p);
      if (S->h_cnt_off) cudaFreeHost(S->h_cnt_off);
      if (S->h_order) cudaFreeHost(S->h_order);
      if (S->h_status) cudaFreeHost(S->h_status);
      if (S->h_cursor) cudaFreeHost(S->h_cursor);
      if (S->kdone) cudaEventDestroy(S->kdone);
      if (S->stream) cudaStreamDestroy(S->stream);
    }
  for (int s = 0; s < 2; s++)
    { if (ctx->scratch[s].p) cudaFree(ctx->scratch[s].p);
      if (ctx->scratch_big[s].p) cudaFree(ctx->scratch_big[s].p);
    }
  if (ctx->d_cthres) cudaFree(ctx->d_cthres);
  if (ctx->d_logfact) cudaFree(ctx->d_logfact);
  for (int i = 0; i < 3; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 7; i++) if (ctx->evp[i]) cudaEventDestroy(ctx->evp[i]);
  free(ctx);
}

extern "C" int cpg_create(cpg_ctx **out, int device, const cpg_model *model,
                          int64_t max_batch_bases, int32_t max_batch_reads)
{ (void)max_batch_bases; (void)max_batch_reads;     /* buffers grow on demand */
  cpg_ctx *ctx = NULL;
  if (out == NULL || model == NULL) return set_err(NULL,CPG_EINVAL,"cpg_create: NULL argument");
  *out = NULL;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return set_err(NULL,CPG_ECUDA,"no CUDA device available (%s): this library has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return set_err(NULL,CPG_EINVAL,"device %d out of range [0,%d)",device,ndev);
  if (model->kmer < 4 || model->kmer > 4096 || model->read_len <= 0 || model->cmax < 1 || model->cmax > 255)
    return set_err(NULL,CPG_EINVAL,"cpg_create: implausible model (kmer=%d read_len=%d cmax=%d)",
                   model->kmer,model->read_len,model->cmax);
  ctx = (cpg_ctx *)calloc(1,sizeof(cpg_ctx));
  if (ctx == NULL) return set_err(NULL,CPG_ENOMEM,"out of host memory");
  ctx->device = device;
  ctx->model = *model;
#define CU_C(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    set_err(NULL,CPG_ECUDA,"%s failed: %s",#call,cudaGetErrorString(e__)); cpg_destroy(ctx); return CPG_ECUDA; } } while (0)
  CU_C(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_C(cudaGetDeviceProperties(&prop,device));
  ctx->n_sm = prop.multiProcessorCount;
  CU_C(cudaMalloc(&ctx->d_cthres,sizeof(model->cthres)));
  CU_C(cudaMalloc(&ctx->d_logfact,sizeof(model->logfact)));
  CU_C(cudaMemcpy(ctx->d_cthres,model->cthres,sizeof(model->cthres),cudaMemcpyHostToDevice));
  CU_C(cudaMemcpy(ctx->d_logfact,model->logfact,sizeof(model->logfact),cudaMemcpyHostToDevice));
  cpg_dmodel &d = ctx->dmodel;
  d.K = model->kmer; d.read_len = model->read_len; d.cmax = model->cmax;
  for (int t = 0; t < 3; t++) d.lmax[t] = model->lmax[t];
  for (int s = 0; s < 4; s++) d.cov[s] = model->cov[s];
  d.dr_ratio = model->dr_ratio; d.hc_erate = model->hc_erate;
  memcpy(d.pe,model->pe,sizeof(d.pe));
  d.cthres = (const uint8_t *)ctx->d_cthres;
  d.logfact = (const double *)ctx->d_logfact;
  for (int s = 0; s < 2; s++)
    { CU_C(cudaHostAlloc((void **)&ctx->slot[s].h_cursor,64,cudaHostAllocDefault));
      CU_C(cudaStreamCreateWithFlags(&ctx->slot[s].stream,cudaStreamNonBlocking));
      CU_C(cudaEventCreateWithFlags(&ctx->slot[s].kdone,cudaEventDisableTiming));
    }
  for (int i = 0; i < 3; i++) CU_C(cudaEventCreate(&ctx->ev[i]));
  for (int i = 0; i < 7; i++) CU_C(cudaEventCreate(&ctx->evp[i]));
  { const char *f = getenv("CPG_FUSED"); ctx->fused = (f && atoi(f) > 0); }
  ctx->walla_smem = sizeof(WallAShared); ctx->rel_smem = sizeof(RelPhaseShared<REL_GROUP>);
  ctx->unrel_smem = sizeof(PhaseShared<UNREL_GROUP,false>);
  CU_C(cudaFuncSetAttribute(k_wall_a,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)ctx->walla_smem));
  CU_C(cudaFuncSetAttribute(k_unrel_b,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)ctx->unrel_smem));
  CU_C(cudaFuncSetAttribute(k_rel,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)ctx->rel_smem));
  { int o = 0;
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_wall_a,WALLA_THREADS,ctx->walla_smem));
    ctx->walla_blocks = ctx->n_sm*(o < 1 ? 1 : o);
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_wall_b,PHASE_THREADS,0));
    ctx->wallb_blocks = ctx->n_sm*(o < 1 ? 1 : o);
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_wall_c,PHASE_THREADS,0));
    ctx->wallc_blocks = ctx->n_sm*(o < 1 ? 1 : o);
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_rel,PHASE_THREADS,ctx->rel_smem));
    ctx->rel_blocks = ctx->n_sm*(o < 1 ? 1 : o);
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_unrel_b,PHASE_THREADS,ctx->unrel_smem));
    ctx->unrel_blocks = ctx->n_sm*(o < 1 ? 1 : o);
    CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o,k_unrel_a,WALLA_THREADS,0));
    ctx->unrela_blocks = ctx->n_sm*(o < 1 ? 1 : o);
  }

  ctx->classify_smem = sizeof(ClassifyShared);
  CU_C(cudaFuncSetAttribute(k_classify,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)ctx->classify_smem));
  int occ = 0;
  CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_classify,CLASSIFY_THREADS,ctx->classify_smem));
  if (occ < 1) occ = 1;
  ctx->classify_blocks = ctx->n_sm*occ;
  ctx->retry_blocks = 2;
  CU_C(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_decode,DECODE_THREADS,0));
  if (occ < 1) occ = 1;
  ctx->decode_blocks = ctx->n_sm*occ;
#undef CU_C
  *out = ctx;
  return CPG_OK;
}

/* scratch arenas: compact blocks for every resident lane group of the main launch, full-size
   blocks for the few groups of the retry launch */
static int ensure_scratch(cpg_ctx *ctx, int P)
{ if (ctx->scratch[0].p && P <= ctx->scratch_P) return CPG_OK;
  size_t off[N_OFF];
  const int K = ctx->model.kmer;
  ScratchDev SC = ctx->SC[0], SB = ctx->SCbig[0], Sw = ctx->SCw[0], Sr = ctx->SCr[0], Su = ctx->SCu[0];
  scratch_caps(&SB,P,K,SM_FULL); SB.stride = scratch_layout(SB,off);
  size_t total = 0, o_r = 0, o_u = 0;
  if (ctx->fused)
    { /* the single-kernel path as the main launch: compact tables, every array */
      scratch_caps(&SC,P,K,SM_WALL); SC.mode = SM_FULL; SC.capC = SC.capI; SC.stride = scratch_layout(SC,off);
      total = SC.stride*(size_t)ctx->classify_blocks*CLASSIFY_GROUPS;
    }
  else
    { scratch_caps(&Sw,P,K,SM_WALL);  Sw.stride = scratch_layout(Sw,off);
      scratch_caps(&Sr,P,K,SM_REL);   Sr.stride = scratch_layout(Sr,off);
      scratch_caps(&Su,P,K,SM_UNREL); Su.stride = scratch_layout(Su,off);
      o_r = Sw.stride*(size_t)ctx->wallb_blocks*(PHASE_THREADS/WALLB_GROUP);
      o_u = o_r+Sr.stride*(size_t)ctx->rel_blocks*(PHASE_THREADS/REL_GROUP);
      total = o_u+Su.stride*(size_t)ctx->unrel_blocks*(PHASE_THREADS/UNREL_GROUP);
    }
  for (int s = 0; s < 2; s++) cudaStreamSynchronize(ctx->slot[s].stream);
  for (int s = 0; s < 2; s++)
    { int rc = reserve(ctx,&ctx->scratch[s],total);
      if (rc) return rc;
      rc = reserve(ctx,&ctx->scratch_big[s],SB.stride*(size_t)ctx->retry_blocks*CLASSIFY_GROUPS);
      if (rc) return rc;
      /* the flag bytes of the wall replay are zero between reads: the arenas start as zeros */
      if (cudaMemset(ctx->scratch[s].p,0,ctx->scratch[s].cap) != cudaSuccess || cudaMemset(ctx->scratch_big[s].p,0,ctx->scratch_big[s].cap) != cudaSuccess)
        return set_err(ctx,CPG_ECUDA,"cudaMemset of the scratch arenas failed: %s",cudaGetErrorString(cudaGetLastError()));
    }
  /* the memsets run on the legacy default stream, which the (non-blocking) slot streams do not wait for:
     nothing may be launched on them before the arenas are really zero */
  if (cudaDeviceSynchronize() != cudaSuccess)
    return set_err(ctx,CPG_ECUDA,"zeroing the scratch arenas failed: %s",cudaGetErrorString(cudaGetLastError()));
  for (int s = 0; s < 2; s++)
    { uint8_t *base = (uint8_t *)ctx->scratch[s].p;
      SC.base = base; Sw.base = base; Sr.base = base+o_r; Su.base = base+o_u; SB.base = (uint8_t *)ctx->scratch_big[s].p;
      ctx->SC[s] = SC; ctx->SCbig[s] = SB; ctx->SCw[s] = Sw; ctx->SCr[s] = Sr; ctx->SCu[s] = Su;
    }
  ctx->scratch_P = P;
  return CPG_OK;
}

static int host_reserve(cpg_ctx *ctx, Slot *S, int n)
{ if ((size_t)n+1 <= S->h_cap) return CPG_OK;
  if (S->h_cnt_off) cudaFreeHost(S->h_cnt_off);
  if (S->h_order) cudaFreeHost(S->h_order);
  if (S->h_status) cudaFreeHost(S->h_status);
  S->h_cnt_off = NULL; S->h_order = NULL; S->h_status = NULL; S->h_cap = 0;
  size_t cap = (size_t)n+n/4+64;
  CU(cudaHostAlloc((void **)&S->h_cnt_off,sizeof(int64_t)*cap,cudaHostAllocDefault));
  CU(cudaHostAlloc((void **)&S->h_order,sizeof(int32_t)*cap,cudaHostAllocDefault));
  CU(cudaHostAlloc((void **)&S->h_status,sizeof(int32_t)*cap,cudaHostAllocDefault));
  S->h_cap = cap;
  return CPG_OK;
}

/* Validate a batch, derive count offsets and the longest-first order, upload everything. */
static int stage_batch(cpg_ctx *ctx, Slot *S, const cpg_batch *b, const int64_t *cls_off)
{ const int n = b->n_reads, K = ctx->model.kmer;
  if (n < 0 || (n > 0 && (!b->seq || !b->seq_off || !b->rlen || !b->prof || !b->prof_off)))
    return set_err(ctx,CPG_EINVAL,"cpg_batch: NULL array");
  if (b->seq_bits != 2 && b->seq_bits != 8) return set_err(ctx,CPG_EINVAL,"cpg_batch: seq_bits must be 2 or 8");
  int rc = host_reserve(ctx,S,n);
  if (rc) return rc;
  int64_t co = 0; int maxP = 1; int maxR = 0;
  for (int i = 0; i < n; i++)
    { const int rl = b->rlen[i];
      if (rl < K || rl > CPG_MAX_RLEN)
        return set_err(ctx,CPG_EINVAL,"read %d of the batch: rlen %d outside [K=%d,%d]",i,rl,K,CPG_MAX_RLEN);
      const int64_t need = (b->seq_bits == 2) ? (rl+3)/4 : rl;
      if (b->seq_off[i+1]-b->seq_off[i] < need || b->prof_off[i+1] < b->prof_off[i])
        return set_err(ctx,CPG_EINVAL,"read %d of the batch: inconsistent offsets",i);
      S->h_cnt_off[i] = co;
      co += (rl-K+1+31) & ~31;
      if (rl-K+1 > maxP) maxP = rl-K+1;
      if (rl > maxR) maxR = rl;
    }
  S->h_cnt_off[n] = co;
  /* Processing order: counting sort by length, longest first, inside chunks of as many consecutive
     reads as k_classify has resident lane groups (reads in flight).  Reads of one CTA have similar
     lengths (its phases end together), the reads in flight stay within a few hundred MB of HBM, and
     the CTAs of a wave differ in length, so they do not all sit in the same phase at the same time.
     Measured on the 100 Mb workload: chunk = resident warps 465 ms; half 569 ms; twice 525 ms; whole
     batch (plain longest-first) 608 ms (profiles/r01_history.md).  CPG_ORDER_CHUNK overrides. */
  { int *bucket = (int *)calloc((size_t)maxR+2,sizeof(int));
    if (bucket == NULL) return set_err(ctx,CPG_ENOMEM,"out of host memory");
    int chunk = ctx->fused ? ctx->classify_blocks*CLASSIFY_GROUPS : ctx->n_sm*256;
    { const char *f = getenv("CPG_ORDER_CHUNK"); if (f && atoi(f) > 0) chunk = atoi(f); }
    if (chunk < 1) chunk = 1;
    for (int c0 = 0; c0 < n; c0 += chunk)
      { const int c1 = (c0+chunk < n) ? c0+chunk : n;
        memset(bucket,0,sizeof(int)*((size_t)maxR+2));
        for (int i = c0; i < c1; i++) bucket[maxR-b->rlen[i]+1]++;
        for (int l = 1; l <= maxR+1; l++) bucket[l] += bucket[l-1];
        for (int i = c0; i < c1; i++) S->h_order[c0+bucket[maxR-b->rlen[i]]++] = i;
      }
    free(bucket);
  }
  rc = ensure_scratch(ctx,maxP);
  if (rc) return rc;

  const size_t seq_bytes = n ? (size_t)b->seq_off[n] : 0, prof_bytes = n ? (size_t)b->prof_off[n] : 0;
  const size_t cls_bytes = n ? (size_t)cls_off[n] : 0;
  if ((rc = reserve(ctx,&S->seq,seq_bytes+64)) || (rc = reserve(ctx,&S->seq_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->rlen,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->prof,prof_bytes+16))
      || (rc = reserve(ctx,&S->prof_off,sizeof(int64_t)*(n+1))) || (rc = reserve(ctx,&S->cnt,sizeof(uint16_t)*(size_t)co+16))
      || (rc = reserve(ctx,&S->cnt_off,sizeof(int64_t)*(n+1))) || (rc = reserve(ctx,&S->plen,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->cand,(size_t)co/8+64))
      || (rc = reserve(ctx,&S->cls,cls_bytes+16)) || (rc = reserve(ctx,&S->cls_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->status,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->order,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->queue,128)) || (rc = reserve(ctx,&S->rec,sizeof(ReadRec)*(size_t)(n+1))))
    return rc;
  /* interval pool: one entry per POOL_DIV profile positions (HiFi profiles need about one per 55);
     a read that does not fit is flagged and goes through the retry launch */
  int pool_div = 12;
  { const char *f = getenv("CPG_POOL_DIV"); if (f && atoi(f) > 0) pool_div = atoi(f); }     /* test knob */
  const int64_t pool_cap = co/pool_div+4096;
  if ((rc = reserve(ctx,&S->pool,sizeof(cpg_intvl)*(size_t)pool_cap))) return rc;
  S->pool_cap = pool_cap;
  if (ctx->result_mode == CPG_RESULT_INTERVALS
      && ((rc = reserve(ctx,&S->ivl,sizeof(uint32_t)*(size_t)pool_cap)) || (rc = reserve(ctx,&S->ivl_at,sizeof(int64_t)*(size_t)(n+1)))
          || (rc = reserve(ctx,&S->ivl_n,sizeof(int32_t)*(size_t)(n+1)))))
    return rc;
  /* candidate records: a header for one position in HDR_DIV and a big record for one in BIG_DIV.  Measured
     needs per position (host build of the device code): plain HiFi profiles 1.5 % / 0.26 %, repeat-rich
     0.43 compressed bytes per k-mer 1.6 % / 1.0 %, noisy low-complexity 2.4 % / 2.1 %; same fallback.  (With
     one big record per 160 positions nearly every read of the repeat-rich workload went to the retry
     launch: 17 s instead of 0.2 s, profiles/r02_c4_caps.log.) */
  int hdr_div = 24, big_div = 32;
  { const char *f = getenv("CPG_HDR_DIV"); if (f && atoi(f) > 0) hdr_div = atoi(f); }       /* test knobs */
  { const char *f = getenv("CPG_BIG_DIV"); if (f && atoi(f) > 0) big_div = atoi(f); }
  const int64_t hdr_cap = co/hdr_div+4096, big_cap = co/big_div+4096;
  /* recorded task values of the unreliable pass: one per interval its sweeps visit (about a third of the
     intervals, one per ~200 positions) */
  int upre_div = 32;
  { const char *f = getenv("CPG_UPRE_DIV"); if (f && atoi(f) > 0) upre_div = atoi(f); }
  const int64_t upre_cap = co/upre_div+4096;
  if (!ctx->fused && ((rc = reserve(ctx,&S->hdr,sizeof(cpg_chdr)*(size_t)hdr_cap)) || (rc = reserve(ctx,&S->big,sizeof(cpg_cbig)*(size_t)big_cap))
                      || (rc = reserve(ctx,&S->upre,sizeof(cpg_upre)*(size_t)upre_cap))))
    return rc;
  cudaStream_t st = S->stream;
  if (n > 0)
    { CU(cudaMemcpyAsync(S->seq.p,b->seq,seq_bytes,cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->seq_off.p,b->seq_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->rlen.p,b->rlen,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->prof.p,b->prof,prof_bytes,cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->prof_off.p,b->prof_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->cnt_off.p,S->h_cnt_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->cls_off.p,cls_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
      CU(cudaMemcpyAsync(S->order.p,S->h_order,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
    }
  S->n_reads = n; S->cls_bytes = (int64_t)cls_bytes; S->maxP = maxP;
  BatchDev &B = S->B;
  B.n_reads = n; B.seq_bits = b->seq_bits;
  B.seq = (const uint8_t *)S->seq.p; B.seq_off = (const int64_t *)S->seq_off.p;
  B.rlen = (const int32_t *)S->rlen.p; B.prof = (const uint8_t *)S->prof.p;
  B.prof_off = (const int64_t *)S->prof_off.p; B.cnt = (uint16_t *)S->cnt.p;
  B.cnt_off = (const int64_t *)S->cnt_off.p; B.plen = (int32_t *)S->plen.p;
  B.cand = (uint32_t *)S->cand.p;
  B.cls = (uint8_t *)S->cls.p; B.cls_off = (const int64_t *)S->cls_off.p;
  B.status = (int32_t *)S->status.p; B.order = (const int32_t *)S->order.p;
  B.queue = (int32_t *)S->queue.p;
  B.phase_cycles = (unsigned long long *)((char *)S->queue.p+32);
  B.pool_cursor = (unsigned long long *)((char *)S->queue.p+64);
  B.rec = (ReadRec *)S->rec.p; B.pool = (cpg_intvl *)S->pool.p; B.pool_cap = pool_cap;
  B.hdr_cursor = (unsigned long long *)((char *)S->queue.p+72);
  B.big_cursor = (unsigned long long *)((char *)S->queue.p+80);
  B.hdr = (cpg_chdr *)S->hdr.p; B.big = (cpg_cbig *)S->big.p; B.hdr_cap = hdr_cap; B.big_cap = big_cap;
  B.ivl = (ctx->result_mode == CPG_RESULT_INTERVALS) ? (uint32_t *)S->ivl.p : (uint32_t *)0;
  B.ivl_at = (int64_t *)S->ivl_at.p; B.ivl_n = (int32_t *)S->ivl_n.p;
  B.upre_cursor = (unsigned long long *)((char *)S->queue.p+88);
  B.upre = (cpg_upre *)S->upre.p; B.upre_cap = upre_cap;
  return CPG_OK;
}

static int launch_kernels(cpg_ctx *ctx, Slot *S, int timed)
{ cudaStream_t st = S->stream;
  if (S->n_reads == 0) return CPG_OK;
  /* The two slots overlap their copies with each other's kernels AND their kernels with each other: every
     slot has its own scratch arenas, and as the CTAs of one slot's persistent kernel run out of reads the
     CTAs of the other slot's kernel take their place (the tail of a batch would leave SMs idle otherwise). */
  const int si = (S == &ctx->slot[0]) ? 0 : 1;
  CU(cudaMemsetAsync(S->queue.p,0,128,st));
  if (timed) CU(cudaEventRecord(ctx->ev[0],st));
  k_decode<<<ctx->decode_blocks,DECODE_THREADS,0,st>>>(S->B,ctx->model.kmer,(int)ctx->model.cov[1]);
  if (timed) CU(cudaEventRecord(ctx->ev[1],st));
  if (ctx->fused)
    k_classify<<<ctx->classify_blocks,CLASSIFY_THREADS,ctx->classify_smem,st>>>(S->B,ctx->dmodel,ctx->SC[si],0);
  else
    { k_wall_a<<<ctx->walla_blocks,WALLA_THREADS,ctx->walla_smem,st>>>(S->B,ctx->dmodel);
      if (timed) CU(cudaEventRecord(ctx->evp[3],st));
      k_wall_b<<<ctx->wallb_blocks,PHASE_THREADS,0,st>>>(S->B,ctx->dmodel,ctx->SCw[si]);
      if (timed) CU(cudaEventRecord(ctx->evp[4],st));
      k_wall_c<<<ctx->wallc_blocks,PHASE_THREADS,0,st>>>(S->B,ctx->dmodel);
      if (timed) CU(cudaEventRecord(ctx->evp[0],st));
      k_rel<<<ctx->rel_blocks,PHASE_THREADS,ctx->rel_smem,st>>>(S->B,ctx->dmodel,ctx->SCr[si]);
      if (timed) CU(cudaEventRecord(ctx->evp[1],st));
      k_unrel_a<<<ctx->unrela_blocks,WALLA_THREADS,0,st>>>(S->B,ctx->dmodel);
      if (timed) CU(cudaEventRecord(ctx->evp[5],st));
      k_unrel_b<<<ctx->unrel_blocks,PHASE_THREADS,ctx->unrel_smem,st>>>(S->B,ctx->dmodel,ctx->SCu[si]);
      if (timed) CU(cudaEventRecord(ctx->evp[6],st));
      if (S->B.ivl == 0) k_emit<<<ctx->n_sm*8,256,0,st>>>(S->B,ctx->model.kmer);
      if (timed) CU(cudaEventRecord(ctx->evp[2],st));
    }
  k_classify<<<ctx->retry_blocks,CLASSIFY_THREADS,ctx->classify_smem,st>>>(S->B,ctx->dmodel,ctx->SCbig[si],1);
  if (S->B.ivl != 0)
    { k_pack<<<ctx->n_sm*8,256,0,st>>>(S->B);
      CU(cudaMemcpyAsync(S->h_cursor,S->B.pool_cursor,sizeof(unsigned long long),cudaMemcpyDeviceToHost,st));
    }
  if (timed) CU(cudaEventRecord(ctx->ev[2],st));
  CU(cudaEventRecord(S->kdone,st));
  S->kdone_valid = 1;
  CU(cudaGetLastError());
  return CPG_OK;
}

static int fetch_result(cpg_ctx *ctx, Slot *S, cpg_result *res)
{ cudaStream_t st = S->stream;
  const int n = S->n_reads;
  if (n > 0)
    { CU(cudaMemcpyAsync(res->cls,S->cls.p,(size_t)S->cls_bytes,cudaMemcpyDeviceToHost,st));
      CU(cudaMemcpyAsync(S->h_status,S->status.p,sizeof(int32_t)*n,cudaMemcpyDeviceToHost,st));
    }
  CU(cudaStreamSynchronize(st));
  int bad = 0;
  for (int i = 0; i < n; i++)
    { if (res->status) res->status[i] = S->h_status[i];
      if (S->h_status[i] & CPG_ST_FATAL)
        { if (!bad) set_err(ctx,CPG_EREAD,"read %d of the batch: %s",i,cpg_status_string(S->h_status[i]));
          bad = 1;
        }
    }
  return bad ? CPG_EREAD : CPG_OK;
}

extern "C" int cpg_submit(cpg_ctx *ctx, int slot, const cpg_batch *batch)
{ if (ctx == NULL || batch == NULL || slot < 0 || slot > 1) return set_err(ctx,CPG_EINVAL,"cpg_submit: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[slot];
  if (S->busy) return set_err(ctx,CPG_EINVAL,"cpg_submit: slot %d not collected",slot);
  /* class offsets default to the prefix sums of rlen; the caller's cpg_result must match */
  int rc = host_reserve(ctx,S,batch->n_reads);
  if (rc) return rc;
  static thread_local int64_t *tmp = NULL; static thread_local size_t tmp_cap = 0;
  if ((size_t)batch->n_reads+1 > tmp_cap)
    { free(tmp); tmp_cap = (size_t)batch->n_reads+1024; tmp = (int64_t *)malloc(sizeof(int64_t)*tmp_cap);
      if (tmp == NULL) { tmp_cap = 0; return set_err(ctx,CPG_ENOMEM,"out of host memory"); }
    }
  tmp[0] = 0;
  for (int i = 0; i < batch->n_reads; i++) tmp[i+1] = tmp[i]+batch->rlen[i];
  rc = stage_batch(ctx,S,batch,tmp);
  if (rc) return rc;
  /* the offsets buffer is consumed by an async copy from pageable memory, which CUDA stages
     before returning, so reusing tmp on the next call is safe */
  rc = launch_kernels(ctx,S,0);
  if (rc) return rc;
  S->busy = 1;
  return CPG_OK;
}

extern "C" int cpg_collect(cpg_ctx *ctx, int slot, cpg_result *res)
{ if (ctx == NULL || res == NULL || slot < 0 || slot > 1) return set_err(ctx,CPG_EINVAL,"cpg_collect: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[slot];
  if (!S->busy) return set_err(ctx,CPG_EINVAL,"cpg_collect: slot %d has no batch in flight",slot);
  if (ctx->result_mode != CPG_RESULT_CLASSES) return set_err(ctx,CPG_EINVAL,"cpg_collect: the context returns intervals (cpg_collect_intervals)");
  /* a rejected result leaves the batch in flight (the slot stays busy): call again with a valid one */
  if (res->cls == NULL && S->cls_bytes > 0) return set_err(ctx,CPG_EINVAL,"cpg_result.cls is NULL");
  if (res->cls_off && (res->cls_off[0] != 0 || res->cls_off[S->n_reads] != S->cls_bytes))
    return set_err(ctx,CPG_EINVAL,"cpg_result.cls_off must be the prefix sums of rlen");
  S->busy = 0;
  return fetch_result(ctx,S,res);
}

extern "C" int cpg_set_result_mode(cpg_ctx *ctx, int mode)
{ if (ctx == NULL || (mode != CPG_RESULT_CLASSES && mode != CPG_RESULT_INTERVALS)) return set_err(ctx,CPG_EINVAL,"cpg_set_result_mode: bad argument");
  if (ctx->slot[0].busy || ctx->slot[1].busy) return set_err(ctx,CPG_EINVAL,"cpg_set_result_mode: a batch is in flight");
  ctx->result_mode = mode;
  return CPG_OK;
}

extern "C" int64_t cpg_intervals_bound(cpg_ctx *ctx, int slot)
{ if (ctx == NULL || slot < 0 || slot > 1) return 0;
  return ctx->slot[slot].pool_cap;
}

extern "C" int cpg_collect_intervals(cpg_ctx *ctx, int slot, cpg_result_ivl *res)
{ if (ctx == NULL || res == NULL || slot < 0 || slot > 1) return set_err(ctx,CPG_EINVAL,"cpg_collect_intervals: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[slot];
  if (!S->busy) return set_err(ctx,CPG_EINVAL,"cpg_collect_intervals: slot %d has no batch in flight",slot);
  if (ctx->result_mode != CPG_RESULT_INTERVALS) return set_err(ctx,CPG_EINVAL,"cpg_collect_intervals: the context returns class strings (cpg_set_result_mode)");
  const int n = S->n_reads;
  if (n > 0 && (res->ivl == NULL || res->ivl_at == NULL || res->ivl_n == NULL)) return set_err(ctx,CPG_EINVAL,"cpg_result_ivl: NULL array");
  cudaStream_t st = S->stream;
  CU(cudaStreamSynchronize(st));                          /* kernels done: the number of pool entries is known */
  const int64_t used = (n > 0) ? (int64_t)*S->h_cursor : 0;
  if (used > res->ivl_cap) return set_err(ctx,CPG_EINVAL,"cpg_result_ivl.ivl holds %lld entries, the batch needs %lld",(long long)res->ivl_cap,(long long)used);
  S->busy = 0;
  res->ivl_used = used;
  if (n > 0)
    { CU(cudaMemcpyAsync(res->ivl,S->ivl.p,sizeof(uint32_t)*(size_t)used,cudaMemcpyDeviceToHost,st));
      CU(cudaMemcpyAsync(res->ivl_at,S->ivl_at.p,sizeof(int64_t)*(size_t)n,cudaMemcpyDeviceToHost,st));
      CU(cudaMemcpyAsync(res->ivl_n,S->ivl_n.p,sizeof(int32_t)*(size_t)n,cudaMemcpyDeviceToHost,st));
      CU(cudaMemcpyAsync(S->h_status,S->status.p,sizeof(int32_t)*n,cudaMemcpyDeviceToHost,st));
    }
  CU(cudaStreamSynchronize(st));
  int bad = 0;
  for (int i = 0; i < n; i++)
    { if (res->status) res->status[i] = S->h_status[i];
      if (S->h_status[i] & CPG_ST_FATAL)
        { if (!bad) set_err(ctx,CPG_EREAD,"read %d of the batch: %s",i,cpg_status_string(S->h_status[i]));
          bad = 1;
        }
    }
  return bad ? CPG_EREAD : CPG_OK;
}

extern "C" int cpg_classify(cpg_ctx *ctx, const cpg_batch *batch, cpg_result *res)
{ int rc = cpg_submit(ctx,0,batch);
  if (rc) return rc;
  return cpg_collect(ctx,0,res);
}

extern "C" int cpg_upload(cpg_ctx *ctx, const cpg_batch *batch)
{ if (ctx == NULL || batch == NULL) return set_err(ctx,CPG_EINVAL,"cpg_upload: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[0];
  int64_t *tmp = (int64_t *)malloc(sizeof(int64_t)*((size_t)batch->n_reads+1));
  if (tmp == NULL) return set_err(ctx,CPG_ENOMEM,"out of host memory");
  tmp[0] = 0;
  for (int i = 0; i < batch->n_reads; i++) tmp[i+1] = tmp[i]+batch->rlen[i];
  int rc = stage_batch(ctx,S,batch,tmp);
  if (rc == CPG_OK && cudaStreamSynchronize(S->stream) != cudaSuccess)
    rc = set_err(ctx,CPG_ECUDA,"upload failed: %s",cudaGetErrorString(cudaGetLastError()));
  free(tmp);
  return rc;
}

extern "C" int cpg_run_resident(cpg_ctx *ctx, int iters, float *ms_decode, float *ms_classify, int *launches)
{ if (ctx == NULL || iters < 1) return set_err(ctx,CPG_EINVAL,"cpg_run_resident: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[0];
  double td = 0., tc = 0.;
  for (int it = 0; it < iters; it++)
    { int rc = launch_kernels(ctx,S,1);
      if (rc) return rc;
      CU(cudaStreamSynchronize(S->stream));
      float a = 0.f, b = 0.f;
      if (S->n_reads > 0)
        { CU(cudaEventElapsedTime(&a,ctx->ev[0],ctx->ev[1]));
          CU(cudaEventElapsedTime(&b,ctx->ev[1],ctx->ev[2]));
          if (!ctx->fused)
            { float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
              CU(cudaEventElapsedTime(&p0,ctx->ev[1],ctx->evp[0]));
              CU(cudaEventElapsedTime(&p1,ctx->evp[0],ctx->evp[1]));
              CU(cudaEventElapsedTime(&p2,ctx->evp[1],ctx->evp[2]));
              CU(cudaEventElapsedTime(&p3,ctx->evp[2],ctx->ev[2]));
              ctx->phase_ns[0] = (uint64_t)(p0*1e6); ctx->phase_ns[1] = (uint64_t)(p1*1e6);
              ctx->phase_ns[2] = (uint64_t)(p2*1e6); ctx->phase_ns[3] = (uint64_t)(p3*1e6);
              CU(cudaEventElapsedTime(&p0,ctx->ev[1],ctx->evp[3]));
              CU(cudaEventElapsedTime(&p1,ctx->evp[3],ctx->evp[4]));
              CU(cudaEventElapsedTime(&p2,ctx->evp[4],ctx->evp[0]));
              ctx->wall_ns[0] = (uint64_t)(p0*1e6); ctx->wall_ns[1] = (uint64_t)(p1*1e6); ctx->wall_ns[2] = (uint64_t)(p2*1e6);
              CU(cudaEventElapsedTime(&p0,ctx->evp[1],ctx->evp[5]));
              CU(cudaEventElapsedTime(&p1,ctx->evp[5],ctx->evp[6]));
              CU(cudaEventElapsedTime(&p2,ctx->evp[6],ctx->evp[2]));
              ctx->unrel_ns[0] = (uint64_t)(p0*1e6); ctx->unrel_ns[1] = (uint64_t)(p1*1e6); ctx->unrel_ns[2] = (uint64_t)(p2*1e6);
            }
        }
      td += a; tc += b;
    }
  if (ms_decode) *ms_decode = (float)(td/iters);
  if (ms_classify) *ms_classify = (float)(tc/iters);
  if (launches) *launches = (ctx->fused ? 3 : 9)*iters;
  return CPG_OK;
}

extern "C" int cpg_phase_cycles(cpg_ctx *ctx, uint64_t out[4])
{ if (ctx == NULL || out == NULL) return set_err(ctx,CPG_EINVAL,"cpg_phase_cycles: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[0];
  if (S->queue.p == NULL) { out[0] = out[1] = out[2] = out[3] = 0; return CPG_OK; }
  CU(cudaStreamSynchronize(S->stream));
  if (!ctx->fused)
    { for (int i = 0; i < 4; i++) out[i] = ctx->phase_ns[i];
      return CPG_OK;
    }
  CU(cudaMemcpy(out,(char *)S->queue.p+32,32,cudaMemcpyDeviceToHost));
  return CPG_OK;
}

extern "C" int cpg_wall_ns(cpg_ctx *ctx, uint64_t out[6])
{ if (ctx == NULL || out == NULL) return set_err(ctx,CPG_EINVAL,"cpg_wall_ns: bad argument");
  for (int i = 0; i < 3; i++) out[i] = ctx->fused ? 0 : ctx->wall_ns[i];
  for (int i = 0; i < 3; i++) out[3+i] = ctx->fused ? 0 : ctx->unrel_ns[i];
  return CPG_OK;
}

/* sums over the reads of the batch on slot 0 (after cpg_run_resident): wall candidates, intervals, reliable
   intervals, intervals the unreliable sweeps visit -- the units the per-kernel algorithmic bytes are made of */
extern "C" int cpg_batch_stats(cpg_ctx *ctx, int64_t out[4])
{ if (ctx == NULL || out == NULL) return set_err(ctx,CPG_EINVAL,"cpg_batch_stats: bad argument");
  CU(cudaSetDevice(ctx->device));
  Slot *S = &ctx->slot[0];
  out[0] = out[1] = out[2] = out[3] = 0;
  if (S->n_reads == 0 || S->rec.p == NULL || ctx->fused) return CPG_OK;
  CU(cudaStreamSynchronize(S->stream));
  ReadRec *h = (ReadRec *)malloc(sizeof(ReadRec)*(size_t)S->n_reads);
  if (h == NULL) return set_err(ctx,CPG_ENOMEM,"out of host memory");
  cudaError_t e = cudaMemcpy(h,S->rec.p,sizeof(ReadRec)*(size_t)S->n_reads,cudaMemcpyDeviceToHost);
  if (e == cudaSuccess)
    for (int i = 0; i < S->n_reads; i++) { out[0] += h[i].ncand; out[1] += h[i].N; out[2] += h[i].M; out[3] += h[i].nf; }
  free(h);
  if (e != cudaSuccess) return set_err(ctx,CPG_ECUDA,"cpg_batch_stats: %s",cudaGetErrorString(e));
  return CPG_OK;
}

extern "C" int cpg_download(cpg_ctx *ctx, cpg_result *res)
{ if (ctx == NULL || res == NULL) return set_err(ctx,CPG_EINVAL,"cpg_download: bad argument");
  CU(cudaSetDevice(ctx->device));
  return fetch_result(ctx,&ctx->slot[0],res);
}

extern "C" int cpg_decode_profiles(cpg_ctx *ctx, int32_t n, const uint8_t *prof, const int64_t *prof_off,
                                   const int64_t *cnt_off, uint16_t *counts, int32_t *plen)
{ if (ctx == NULL || n < 0 || (n > 0 && (!prof || !prof_off || !cnt_off || !counts || !plen)))
    return set_err(ctx,CPG_EINVAL,"cpg_decode_profiles: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return CPG_OK;
  Slot *S = &ctx->slot[0];
  if (S->busy) return set_err(ctx,CPG_EINVAL,"cpg_decode_profiles: slot 0 busy");
  const int K = ctx->model.kmer;
  int rc = host_reserve(ctx,S,n);
  if (rc) return rc;
  const size_t prof_bytes = (size_t)prof_off[n], cnt_n = (size_t)cnt_off[n];
  int32_t *rl = (int32_t *)malloc(sizeof(int32_t)*(size_t)n);
  if (rl == NULL) return set_err(ctx,CPG_ENOMEM,"out of host memory");
  for (int i = 0; i < n; i++) { rl[i] = (int32_t)(cnt_off[i+1]-cnt_off[i])+K-1; S->h_order[i] = i; }
  if ((rc = reserve(ctx,&S->prof,prof_bytes+16)) || (rc = reserve(ctx,&S->prof_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->cnt,sizeof(uint16_t)*cnt_n+16)) || (rc = reserve(ctx,&S->cnt_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->rlen,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->plen,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->status,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->order,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->queue,sizeof(int32_t)*4)))
    { free(rl); return rc; }
  cudaStream_t st = S->stream;
  BatchDev B; memset(&B,0,sizeof(B));
  B.n_reads = n; B.prof = (const uint8_t *)S->prof.p; B.prof_off = (const int64_t *)S->prof_off.p;
  B.cnt = (uint16_t *)S->cnt.p; B.cnt_off = (const int64_t *)S->cnt_off.p; B.rlen = (const int32_t *)S->rlen.p;
  B.plen = (int32_t *)S->plen.p; B.status = (int32_t *)S->status.p; B.order = (const int32_t *)S->order.p;
  B.queue = (int32_t *)S->queue.p;
  cudaError_t e = cudaSuccess;
#define TRY(x) if (e == cudaSuccess) e = (x)
  TRY(cudaMemcpyAsync(S->prof.p,prof,prof_bytes,cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->prof_off.p,prof_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->cnt_off.p,cnt_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->rlen.p,rl,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->order.p,S->h_order,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
  TRY(cudaMemsetAsync(S->queue.p,0,sizeof(int32_t)*4,st));
  TRY(cudaMemsetAsync(S->cnt.p,0,sizeof(uint16_t)*cnt_n,st));
  if (e == cudaSuccess) k_decode<<<ctx->decode_blocks,DECODE_THREADS,0,st>>>(B,K,0);
  TRY(cudaGetLastError());
  TRY(cudaMemcpyAsync(counts,S->cnt.p,sizeof(uint16_t)*cnt_n,cudaMemcpyDeviceToHost,st));
  TRY(cudaMemcpyAsync(plen,S->plen.p,sizeof(int32_t)*n,cudaMemcpyDeviceToHost,st));
  TRY(cudaStreamSynchronize(st));
#undef TRY
  free(rl);
  if (e != cudaSuccess) return set_err(ctx,CPG_ECUDA,"cpg_decode_profiles: %s",cudaGetErrorString(e));
  return CPG_OK;
}

extern "C" int cpg_prof2class(cpg_ctx *ctx, int32_t n, const uint8_t *prof, const int64_t *prof_off,
                              const int32_t *rlen, uint8_t *cls, int32_t *status)
{ if (ctx == NULL || n < 0 || (n > 0 && (!prof || !prof_off || !rlen || !cls)))
    return set_err(ctx,CPG_EINVAL,"cpg_prof2class: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return CPG_OK;
  Slot *S = &ctx->slot[0];
  if (S->busy) return set_err(ctx,CPG_EINVAL,"cpg_prof2class: slot 0 busy");
  const int K = ctx->model.kmer;
  int rc = host_reserve(ctx,S,n);
  if (rc) return rc;
  int64_t *cls_off = (int64_t *)malloc(sizeof(int64_t)*((size_t)n+1));
  if (cls_off == NULL) return set_err(ctx,CPG_ENOMEM,"out of host memory");
  int64_t co = 0, lo = 0;
  for (int i = 0; i < n; i++)
    { if (rlen[i] < 0 || rlen[i] > CPG_MAX_RLEN || prof_off[i+1] < prof_off[i])
        { free(cls_off); return set_err(ctx,CPG_EINVAL,"cpg_prof2class: read %d: bad length or offsets",i); }
      S->h_cnt_off[i] = co; cls_off[i] = lo; S->h_order[i] = i;
      const int plen = rlen[i] >= K ? rlen[i]-K+1 : 0;
      co += (plen+31) & ~31; lo += rlen[i];
    }
  S->h_cnt_off[n] = co; cls_off[n] = lo;
  const size_t prof_bytes = (size_t)prof_off[n];
  if ((rc = reserve(ctx,&S->prof,prof_bytes+16)) || (rc = reserve(ctx,&S->prof_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->cnt,sizeof(uint16_t)*(size_t)co+64)) || (rc = reserve(ctx,&S->cnt_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->rlen,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->plen,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->status,sizeof(int32_t)*(n+1))) || (rc = reserve(ctx,&S->order,sizeof(int32_t)*(n+1)))
      || (rc = reserve(ctx,&S->cls,(size_t)lo+16)) || (rc = reserve(ctx,&S->cls_off,sizeof(int64_t)*(n+1)))
      || (rc = reserve(ctx,&S->queue,128)))
    { free(cls_off); return rc; }
  cudaStream_t st = S->stream;
  BatchDev B; memset(&B,0,sizeof(B));
  B.n_reads = n; B.prof = (const uint8_t *)S->prof.p; B.prof_off = (const int64_t *)S->prof_off.p;
  B.cnt = (uint16_t *)S->cnt.p; B.cnt_off = (const int64_t *)S->cnt_off.p; B.rlen = (const int32_t *)S->rlen.p;
  B.plen = (int32_t *)S->plen.p; B.status = (int32_t *)S->status.p; B.order = (const int32_t *)S->order.p;
  B.cls = (uint8_t *)S->cls.p; B.cls_off = (const int64_t *)S->cls_off.p;
  B.queue = (int32_t *)S->queue.p;
  cudaError_t e = cudaSuccess;
#define TRY(x) if (e == cudaSuccess) e = (x)
  TRY(cudaMemcpyAsync(S->prof.p,prof,prof_bytes,cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->prof_off.p,prof_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->cnt_off.p,S->h_cnt_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->cls_off.p,cls_off,sizeof(int64_t)*(n+1),cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->rlen.p,rlen,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
  TRY(cudaMemcpyAsync(S->order.p,S->h_order,sizeof(int32_t)*n,cudaMemcpyHostToDevice,st));
  TRY(cudaMemsetAsync(S->queue.p,0,128,st));
  if (e == cudaSuccess)
    { k_decode<<<ctx->decode_blocks,DECODE_THREADS,0,st>>>(B,K,0);
      k_count2class<<<ctx->n_sm*8,256,0,st>>>(B,K);
    }
  TRY(cudaGetLastError());
  TRY(cudaMemcpyAsync(cls,S->cls.p,(size_t)lo,cudaMemcpyDeviceToHost,st));
  TRY(cudaMemcpyAsync(S->h_status,S->status.p,sizeof(int32_t)*n,cudaMemcpyDeviceToHost,st));
  TRY(cudaStreamSynchronize(st));
#undef TRY
  free(cls_off);
  if (e != cudaSuccess) return set_err(ctx,CPG_ECUDA,"cpg_prof2class: %s",cudaGetErrorString(e));
  int bad = 0;
  for (int i = 0; i < n; i++)
    { if (status) status[i] = S->h_status[i];
      if (S->h_status[i] & CPG_ST_BAD_PROFILE)
        { if (!bad) set_err(ctx,CPG_EREAD,"read %d of the batch: %s",i,cpg_status_string(S->h_status[i]));
          bad = 1;
        }
    }
  return bad ? CPG_EREAD : CPG_OK;
}

/* pinned host memory for callers that want true asynchronous copies */
extern "C" void *cpg_host_alloc(size_t bytes)
{ void *p = NULL;
  if (cudaHostAlloc(&p,bytes ? bytes : 1,cudaHostAllocDefault) != cudaSuccess) return NULL;
  return p;
}
extern "C" void cpg_host_free(void *p) { if (p) cudaFreeHost(p); }
