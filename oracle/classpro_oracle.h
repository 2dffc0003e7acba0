/*******************************************************************************************
 *  classpro_oracle.h -- CPU restatement of ClassPro's per-read classification path.
 *
 *  TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by the
 *  product (classpro_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 *  --impl reference legs may use it, and only as the checker.
 *
 *  Parity status: PINNED.  The restatement is checked byte-for-byte against the unmodified
 *  reference binary (oracle/_ref/ClassPro, compiled from /root/reference/src by oracle/Makefile)
 *  on seeded synthetic datasets (tests/test_oracle_vs_reference.py, tests/golden/).
 *
 *  All file:line citations are relative to /root/reference/.
 *******************************************************************************************/
#ifndef CLASSPRO_ORACLE_H
#define CLASSPRO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { CPO_E = 0, CPO_R = 1, CPO_H = 2, CPO_D = 3, CPO_NSTATE = 4 };   /* src/ClassPro.h:57 */
enum { CPO_HP = 0, CPO_DS = 1, CPO_TS = 2, CPO_NCTYPE = 3 };           /* src/ClassPro.h:58 */
enum { CPO_SELF = 0, CPO_OTHERS = 1 };                                 /* src/ClassPro.h:59 */
enum { CPO_DROP = 0, CPO_GAIN = 1 };                                   /* src/ClassPro.h:60 */
enum { CPO_INIT = 0, CPO_FINAL = 1 };                                  /* src/ClassPro.h:122 */

#define CPO_MAX_CNT   32767            /* src/const.c:38  */
#define CPO_MAX_RLEN  60000            /* src/const.c:57  */
#define CPO_LMAX0     20               /* src/const.c:60  (MAX_N_LC) */

/* Host one-shot model: src/ClassPro.c:536-554, src/hist.c:28-143, src/wall.c:120-244 */
typedef struct
  { int      K;
    int      read_len;                      /* -r (src/ClassPro.c:516) */
    uint16_t cov[CPO_NSTATE];               /* GLOBAL_COV (src/ClassPro.c:544-547) */
    double   dr_ratio;                      /* src/ClassPro.c:548 */
    int      cmax;                          /* src/wall.c:178 */
    double   hc_erate;                      /* src/wall.c:180 */
    int      lmax[CPO_NCTYPE];              /* src/wall.c:123 */
    double   pe[CPO_NCTYPE][CPO_LMAX0+1];   /* src/wall.c:140-142 */
    /* cthres[t][l][cout][thresT][etype], cout < cmax <= 255 (src/wall.c:190-224) */
    uint8_t  cthres[CPO_NCTYPE][CPO_LMAX0+1][256][2][2];
    double   logfact[CPO_MAX_CNT+1];        /* src/prob.c:12-19 */
  } cpo_model;

typedef struct
  { int32_t  b, e;
    uint16_t cb, ce, ccb, cce;
    uint8_t  is_rel;
    int8_t   asgn;
    double   pe, pe_o_b, pe_o_e;
  } cpo_intvl;                              /* src/ClassPro.h:159-170 */

typedef struct cpo_work cpo_work;

/* Model construction.  hist points at the int64 bins hist[low..high] as stored in <root>.hist. */
int  cpo_model_from_hist(cpo_model *M, int kmer, int low, int high, int64_t ilowcnt, int64_t ihighcnt,
                         const int64_t *hist, int cov_opt, int read_len, int verbose);
int  cpo_model_load(cpo_model *M, const char *fk_root, int cov_opt, int read_len, int verbose);
/* Model from explicit coverages (as -c would, with D given and H = D>>1 unless h > 0). */
int  cpo_model_from_cov(cpo_model *M, int kmer, int h, int d, int read_len);

/* Stage functions */
int  cpo_decode_profile(const uint8_t *bytes, int64_t len, uint16_t *out, int cap);   /* libfastk.c:1414-1562 */
void cpo_seq_context(uint8_t (*lctx)[3], uint8_t (*rctx)[3], const char *seq, int rlen); /* context.c:8-108 */
double cpo_bessi(int n, double x);                                                    /* bessel.c:482-521 */
double cpo_binom_test_g(const cpo_model *M, int k, int n, double pe, int exact);     /* prob.c:76-112 */

cpo_work *cpo_work_new(void);
void      cpo_work_free(cpo_work *W);
/* clean = 1: index plen of the wall/perror scratch and the right-context buffer are reset per read
 * (the device definition); clean = 0: scratch persists across reads as in one
 * reference thread. */
void      cpo_work_set_clean(cpo_work *W, int clean);

/* Classify one read.  seq: rlen chars; prof: plen = rlen-K+1 counts.  cls receives rlen chars + NUL
 * ('N' x (K-1) then E/H/D/R per k-mer).  Returns the number of intervals N (>= 1), or <0 on error. */
int  cpo_classify_read(const cpo_model *M, cpo_work *W, const char *seq, int rlen,
                       const uint16_t *prof, int plen, char *cls);
/* Access to the intervals of the last classified read (after classify_unrel). */
const cpo_intvl *cpo_last_intervals(const cpo_work *W, int *N, int *Mrel);

/* Whole-file driver: reads <fastx> (+ .hist/.prof of fk_root), writes out_path in the format of
 * src/ClassPro.c:289.  Returns 0 on success.  nkmers (optional) receives the classified k-mers. */
int  cpo_run_file(const char *fastx, const char *fk_root, int cov_opt, int read_len,
                  const char *out_path, int verbose, int64_t *nkmers);

#ifdef __cplusplus
}
#endif
/* log near-ties of the order decisions to stderr (see classpro_oracle.c: trace_cmp) */
void cpo_set_trace(int on);

#endif
