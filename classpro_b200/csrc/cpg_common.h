/*******************************************************************************************
 *  cpg_common.h -- data layouts shared by the CUDA kernels and the C-ABI implementation.
 *
 *  The per-read logic in cpg_*.cuh is written in "warp-uniform" style: the 32 lanes of the warp
 *  that owns a read execute the same scalar control flow on the same values (no divergence),
 *  split the data-parallel inner loops (profile sweeps, binomial tail terms, the 4x4 transition
 *  table, interval sorting) by lane, and exchange results through a small per-warp block of
 *  shared memory.  With CPG_HOSTSIM defined the same sources compile as plain C++ with a warp
 *  width of 1; that build exists ONLY under tests/ (tests/hostsim) to unit-test the device logic
 *  on machines without a GPU.  It is never linked into libclasspro_b200.so.
 *
 *  file:line citations are relative to /root/reference/.
 *******************************************************************************************/
#ifndef CPG_COMMON_H
#define CPG_COMMON_H

#include <stdint.h>
#include <math.h>

#ifdef CPG_HOSTSIM
  #include <string.h>
  #define CPG_DEV        static inline
  #define CPG_DEV_NOINL  static
  #define CPG_DEV_HELPER static inline
  #define CPG_DEV_MATHFN static inline
  #define CPG_LDG(p)     (*(p))
  #define CPG_INF        ((double)INFINITY)
  #define CPG_LOOP
  #define CPG_UNROLL4
  #if CPG_HOSTSIM == 32
    /* 32 host threads play the lanes of one warp; every warp primitive is a rendezvous, so a
       collective reached by only some lanes, or a missing __syncwarp, shows up as a hang or as
       nondeterministic output in the CPU test-suite (tests/hostsim/hostsim.cpp) */
    #define CPG_WARP       32
    void     cpg_sim_barrier(void);
    unsigned cpg_sim_ballot(int pred);
    unsigned cpg_sim_shfl(unsigned v, int src);
    unsigned cpg_sim_shfl_up(unsigned v, int d);
    int      cpg_sim_sum(int v);
    unsigned cpg_sim_gballot(unsigned mask, int pred);
    int      cpg_sim_gsum(unsigned mask, int v);
    unsigned cpg_sim_gshfl(unsigned mask, unsigned v, int src);
    #define CPG_SYNCWARP() cpg_sim_barrier()
    void     cpg_sim_group_barrier(unsigned mask);
    #define CPG_SYNCGROUP(W) cpg_sim_group_barrier((W).gmask)
  #else
    #define CPG_WARP       1
    #define CPG_SYNCWARP() do { } while (0)
    #define CPG_SYNCGROUP(W) do { } while (0)
  #endif
#else
  #include <cuda_runtime.h>
  #define CPG_DEV        __device__ __forceinline__
  #define CPG_DEV_NOINL  __device__ __noinline__
  /* small helpers: inlined at every use (measured faster than one out-of-line copy each:
     107 ms vs 118 ms for k_classify on the 10 Mb bench); -DCPG_NOINLINE_HELPERS flips it */
  #ifndef CPG_NOINLINE_HELPERS
    #define CPG_DEV_HELPER __device__ __forceinline__
  #else
    #define CPG_DEV_HELPER __device__ __noinline__
  #endif
  /* exp/log: inlined (default) or one shared copy (-DCPG_NOINLINE_MATH) */
  #ifdef CPG_NOINLINE_MATH
    #define CPG_DEV_MATHFN __device__ __noinline__
  #else
    #define CPG_DEV_MATHFN __device__ __forceinline__
  #endif
  #define CPG_WARP       32
  #define CPG_SYNCWARP() __syncwarp()
  #define CPG_SYNCGROUP(W) __syncwarp((W).gmask)
  #define CPG_LDG(p)     __ldg(p)
  /* loops of the per-read logic are not unrolled: with ~30 warps per SM in different phases of a
     large kernel, the instruction-cache footprint matters more than loop overhead */
  #ifdef CPG_UNROLL_LOOPS
    #define CPG_LOOP
  #else
    #define CPG_LOOP     _Pragma("unroll 1")
  #endif
  /* short loops of independent loads: unrolled so that the loads are in flight together */
  #define CPG_UNROLL4    _Pragma("unroll 4")
  #define CPG_INF        (__longlong_as_double(0x7ff0000000000000LL))
#endif

/* enums of src/ClassPro.h:57-60,122 */
enum { ST_E = 0, ST_R = 1, ST_H = 2, ST_D = 3, ST_N = 4 };
enum { CT_HP = 0, CT_DS = 1, CT_TS = 2, CT_N = 3 };
enum { ET_SELF = 0, ET_OTHERS = 1 };
enum { WT_DROP = 0, WT_GAIN = 1 };
enum { TH_INIT = 0, TH_FINAL = 1 };

#define CPG_MAX_CNT    32767      /* src/const.c:38 */
#define CPG_MAX_RLEN   60000      /* src/const.c:57 */
#define CPG_LROWS      36         /* 20 + 10 + 6 context rows (src/wall.c:123) */

/* constants of src/const.c:56-73 */
#define CPG_MAX_N_HC        5
#define CPG_MIN_CNT_CHANGE  3
#define CPG_MAX_CNT_CHANGE  5
#define CPG_PE_INIT_SELF    0.001
#define CPG_PE_INIT_OTHERS  0.05
#define CPG_PE_FINAL        1e-5          /* PE_THRES[FINAL][SELF] == PE_THRES[FINAL][OTHERS] */
#define CPG_THRES_DIFF_EO   (-23.025851)
#define CPG_THRES_DIFF_REL  (-9.210340)
#define CPG_OFFSET          1000
#define CPG_R_LOGP          (-10.)
#define CPG_E_PO_BASE       (-10.)
#define CPG_PE_MEAN         0.01

/* per-read status bits returned to the host */
#define CPG_ST_OK             0
#define CPG_ST_BAD_PROFILE    1    /* decoded length != rlen-K+1 (src/ClassPro.c:234-237: exit(1)) */
#define CPG_ST_EINTVL_OVF     2    /* "# E-intvls >= plen" (src/wall.c:783-788: exit(1)) */
#define CPG_ST_NO_PROB        4    /* "No valid probability for interval" (class_unrel.c:221-226) */
#define CPG_ST_INTERP         8    /* invalid interpolation points (src/util.c:26-31: exit(1)) */
#define CPG_ST_UNDEF_TRACE   16    /* all DP states impossible at the last interval: the reference
                                      reads a stale path row (class_rel.c:62-73,606-613) */
#define CPG_ST_LONG_RUN      32    /* a low-complexity run reached the 127 cap: the reference reads
                                      never-written right-context cells (context.c:26-27) */
#define CPG_ST_BINOM         64    /* k > n in a binomial (src/prob.c:51-55: exit(1)) */
/* internal, never reaches the host: the read outgrew the compact scratch block of its lane group
   and is classified again by the second k_classify launch, which has full-size blocks */
#define CPG_ST_RETRY    (1<<20)
#define CPG_ST_ABORT    (CPG_ST_EINTVL_OVF|CPG_ST_RETRY)

/* Device-resident model: host one-shot results (src/ClassPro.c:536-554, src/wall.c:167-244) */
typedef struct
  { int32_t  K;
    int32_t  read_len;
    int32_t  cmax;
    int32_t  lmax[3];
    uint16_t cov[4];
    double   dr_ratio;
    double   hc_erate;
    double   pe[3][21];
    /* logarithms of the model's constants, filled by cpg_model_fill_logs() with the same log() the
       per-read code uses, so that replacing a log of a constant by its table entry changes no bit */
    double   lpe[3][21], l1mpe[3][21];   /* log pe[t][l], log(1-pe[t][l]) */
    double   l_hc, l1m_hc;               /* log hc_erate, log(1-hc_erate) */
    double   l_p1, l1m_p1;               /* log 0.1, log(1-0.1)   (src/class_unrel.c p_errorin rate) */
    double   l_p99, l1m_p99;             /* log(1-PE_MEAN), log(1-(1-PE_MEAN)) */
    double   lcov[4];                    /* log cov[s] */
    /* device pointers */
    const uint8_t *cthres;     /* [CPG_LROWS][256][2(thresT)][2(etype)], row = lrow(t,l) */
    const double  *logfact;    /* [32768] */
  } cpg_dmodel;

/* row of context (type t, length l>=1) in the flattened threshold table */
#define CPG_LROW(t,l) (((t) == 0 ? 0 : ((t) == 1 ? 20 : 30)) + (l) - 1)

typedef struct
  { int32_t  b, e;
    uint16_t cb, ce, ccb, cce;
    uint8_t  is_rel;
    int8_t   asgn;
    uint8_t  pad[6];
    double   pe, peob, peoe;
  } cpg_intvl;                  /* src/ClassPro.h:159-170 */

typedef struct { int32_t b, e; double pe; } cpg_eintvl;   /* src/ClassPro.h:153-157 */

/* Sequence view: 2-bit packed (A,C,G,T = 0..3, base i in bits 2*(i&3) of byte i>>2) or raw bytes */
typedef struct { const uint8_t *p; int32_t bits; } cpg_seq;

/* What the pure step of the unreliable pass (k_unrel_a, cpg_unrel.cuh) records for an interval the sweeps
   will visit: its four nearest reliable H / D neighbours and the ten task values under them */
typedef struct
  { int32_t nb[4];      /* H left, H right, D left, D right (interval index, -1 = none; -2 = forced R, no values) */
    double  val[10];
    int32_t st;         /* status bits the evaluation raised: they count only if the values are used */
    int32_t pad;
  } cpg_upre;           /* 104 bytes */

/* ---- wall stage, step 1 -> step 2: what the pure, candidate-parallel step (k_wall_a, cpg_wall.cuh "wa_")
 *      leaves for the order-dependent replay of a read (k_wall_b, "wb_").  One header per wall candidate of
 *      the read, in position order; one big record for every candidate that gets past the count thresholds
 *      of src/wall.c:643-675 for at least one error type (about one in ten on HiFi profiles). ---- */
typedef struct
  { int32_t  pos;       /* profile position i of the candidate */
    uint32_t info;      /* CH_* bits */
    uint32_t big;       /* index of its cpg_cbig record (valid iff info & (CH_REACH_S|CH_REACH_O)) */
    uint32_t pad;
  } cpg_chdr;
#define CH_REACH_S  0x01u   /* SELF gets past the thresholds (before the paired flags are looked at) */
#define CH_REACH_O  0x02u   /* OTHERS likewise */
#define CH_ONOW     0x04u   /* OTHERS: a wall by the thresholds alone (src/wall.c:672-675) */
#define CH_GAIN     0x08u   /* wall type: count gain (else drop) */
#define CH_LONG     0x10u   /* a context run at the candidate reached the 127 cap */

typedef struct
  { double   own[2];    /* p_errorin of the candidate itself under its context's error rate, per error type */
    double   term[23];  /* partner probabilities, layout in cpg_wall.cuh */
    int32_t  lc_j;      /* low-complexity partner position */
    uint8_t  lc_kind;   /* 0 = no partner, 1 = read boundary, 2 = regular */
    uint8_t  nhc;       /* high-complexity partners inside the profile (the reference's loop stops at the first outside) */
    uint8_t  bad;       /* k > n in a binomial */
    uint8_t  lr_walk;   /* the low-complexity walk met a run at the 127 cap */
    uint16_t ok;        /* bit e*7: the low-complexity partner passes the count tests of error type e;
                           bit e*7+1+n: high-complexity partner n does */
    uint16_t pad[3];
  } cpg_cbig;           /* 216 bytes */

/* Scratch block of a lane group in global memory.  P = longest profile of the batch.  The
 * per-position arrays have P entries; the tables have capS/capE/capI/capC entries: a few per
 * cent of P in the blocks of the main launch (a read that outgrows them is flagged
 * CPG_ST_RETRY), P+2 -- the worst case -- in the blocks of the retry launch.
 * mark[] is CLEAN (all zero) between reads: the replay logs every position it touches and zeroes
 * exactly those when the read is done, so no per-position sweep is ever needed. */
typedef struct
  { uint8_t    *mark;     /* [P+2+32] flag byte per profile position 0..plen */
    uint16_t   *slot;     /* [P+2]   probability slot of a position, valid where the flag byte says so */
    double     *perr;     /* [capS*4] slot-major: [slot][etype][wtype] */
    cpg_eintvl *eint;     /* [capE] */
    cpg_intvl  *intvl;    /* [capI] */
    int32_t     capS, capE, capI;
    int32_t    *tlog;     /* [capT] positions whose flag byte is not zero (may hold duplicates) */
    int32_t     capT, capC;
    cpg_chdr   *hdr;      /* [capC] candidate headers (single-kernel path; the phase kernels read the batch's arrays) */
    cpg_cbig   *big;      /* [capC] */
    cpg_intvl  *rint;     /* [MC]  reliable intervals (copy) */
    cpg_intvl  *wint;     /* [2*MC] DP working copies (forward, backward) */
    uint16_t   *bp;       /* [2*MC] back pointers: 4 x 3 bits */
    uint8_t    *asg_f;    /* [MC] */
    uint8_t    *asg_b;    /* [MC] */
    uint8_t    *rpos;     /* [2*MC] */
    int32_t     MC;
    int32_t    *ord;      /* [capI] unreliable pass: the intervals the sweeps visit, in index order */
    int32_t    *srt;      /* [capI] ... their positions in ord[], in sweep order */
    uint32_t   *key;      /* [capI] ... their sort keys */
    cpg_upre   *upre;     /* [capI] ... their recorded task values (single-kernel path; the phase kernels read the batch's array) */
  } cpg_scratch;

/* Exchange block of a lane group (shared memory on the device): task results */
typedef struct
  { double term[24];
  } cpg_wshared;

#endif
