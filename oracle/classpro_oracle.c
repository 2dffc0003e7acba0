/*******************************************************************************************
 *  classpro_oracle.c -- CPU restatement of ClassPro's per-read classification path.
 *
 *  TEST INFRASTRUCTURE ONLY (see classpro_oracle.h).  Parity status: PINNED against the
 *  unmodified reference binary oracle/_ref/ClassPro (byte-identical .class on the seeded
 *  datasets of tests/).  Plain serial C; data structures are dense per-position arrays as in the
 *  reference so that every decision can be compared one to one.  The floating-point expression
 *  order of the reference is kept (compile WITHOUT -ffast-math / -ffp-contract=fast).
 *
 *  All file:line citations are relative to /root/reference/.
 *******************************************************************************************/
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>
#include <zlib.h>
#include "classpro_oracle.h"

#define MINI(a,b) ((a) < (b) ? (a) : (b))
#define MAXI(a,b) ((a) > (b) ? (a) : (b))

/* constants of src/const.c:56-73 */
static const int    N_SIGMA_RCOV   = 5;
static const int    MAX_N_HC       = 5;
static const int    MIN_CNT_CHANGE = 3;
static const int    MAX_CNT_CHANGE = 5;
static const double PE_THRES[2][2] = { {0.001, 0.05}, {1e-5, 1e-5} };
static const double THRES_DIFF_EO  = -23.025851;
static const double THRES_DIFF_REL = -9.210340;
static const int    OFFSET         = 1000;
static const int    N_SIGMA_R      = 2;
static const double R_LOGP         = -10.;
static const double E_PO_BASE      = -10.;
static const double PE_MEAN        = 0.01;
static const char   STOC[4]        = { 'E','R','H','D' };

/*********************************************************************************************
 *  Numeric primitives: src/bessel.c:390-521, src/prob.c:12-112, src/util.c:9-55
 *********************************************************************************************/

/* bessel.c:390-411 */
static double bessel_i0(double x)
{ double ax = fabs(x), y, ans;
  if (ax < 3.75)
    { y = x/3.75; y = y*y;
      ans = 1.0+y*(3.5156229+y*(3.0899424+y*(1.2067492+y*(0.2659732+y*(0.360768e-1+y*0.45813e-2)))));
    }
  else
    { y = 3.75/ax;
      ans = (exp(ax)/sqrt(ax))*(0.39894228+y*(0.1328592e-1+y*(0.225319e-2+y*(-0.157565e-2+y*(0.916281e-2
            +y*(-0.2057706e-1+y*(0.2635537e-1+y*(-0.1647633e-1+y*0.392377e-2))))))));
    }
  return ans;
}

/* bessel.c:416-439 */
static double bessel_i1(double x)
{ double ax = fabs(x), y, ans;
  if (ax < 3.75)
    { y = x/3.75; y = y*y;
      ans = ax*(0.5+y*(0.87890594+y*(0.51498869+y*(0.15084934+y*(0.2658733e-1+y*(0.301532e-2+y*0.32411e-3))))));
    }
  else
    { y = 3.75/ax;
      ans = 0.2282967e-1+y*(-0.2895312e-1+y*(0.1787654e-1-y*0.420059e-2));
      ans = 0.39894228+y*(-0.3988024e-1+y*(-0.362018e-2+y*(0.163801e-2+y*(-0.1031555e-1+y*ans))));
      ans *= (exp(ax)/sqrt(ax));
    }
  return x < 0.0 ? -ans : ans;
}

/* bessel.c:482-521: downward recurrence from 2*(n+floor(sqrt(40 n))) with 1e10 rescaling */
double cpo_bessi(int n, double x)
{ if (n < 0) { fprintf(stderr,"n<0 @ bessi\n"); exit(1); }
  if (n == 0) return bessel_i0(x);
  if (n == 1) return bessel_i1(x);
  if (x == 0.0) return 0.0;
  double tox = 2.0/fabs(x), bip = 0.0, ans = 0.0, bi = 1.0, bim;
  for (int j = 2*(n+(int)sqrt(40.0*n)); j > 0; j--)
    { bim = bip+j*tox*bi;
      bip = bi;
      bi = bim;
      if (fabs(bi) > 1.0e10)
        { ans *= 1.0e-10; bi *= 1.0e-10; bip *= 1.0e-10; }
      if (j == n) ans = bip;
    }
  ans *= bessel_i0(x)/bi;
  return (x < 0.0 && (n%2) == 1) ? -ans : ans;
}

/* prob.c:22-31 (count clamp with a note on stderr) */
static int clamp_cnt(int n)
{ if (n > CPO_MAX_CNT)
    { fprintf(stderr,"K-mer count (%d) > MAX_KMER_CNT (%d) (due to D/R ratio?)\n",n,CPO_MAX_CNT);
      return CPO_MAX_CNT;
    }
  return n;
}

/* prob.c:33-39; k is a 16-bit count in the reference */
static double lp_poisson(const cpo_model *M, uint16_t k16, int lambda)
{ int k = clamp_cnt(k16);
  return k*log((double)lambda)-lambda-M->logfact[k];
}

/* prob.c:41-44 */
static double lp_skellam(int k, double lambda)
{ return -2.*lambda+log(cpo_bessi(abs(k),2.*lambda)); }

static void check_binom(int *k, int *n)   /* prob.c:47-57 */
{ *k = clamp_cnt(*k); *n = clamp_cnt(*n);
  if (*k > *n) { fprintf(stderr,"k (%d) > n (%d) in Binom\n",*k,*n); exit(1); }
}

/* prob.c:59-65 */
static double lp_binom(const cpo_model *M, uint16_t k16, uint16_t n16, double p)
{ int k = k16, n = n16; check_binom(&k,&n);
  return M->logfact[n]-M->logfact[k]-M->logfact[n-k]+k*log(p)+(n-k)*log(1-p);
}

/* prob.c:67-73 */
static double lp_binom_pre(const cpo_model *M, int k, int n, double lpe, double l1mpe)
{ return M->logfact[n]-M->logfact[k]-M->logfact[n-k]+k*lpe+(n-k)*l1mpe; }

/* prob.c:76-112: one-sided binomial tail, truncated when a term drops below a tenth of the first */
double cpo_binom_test_g(const cpo_model *M, int k, int n, double pe, int exact)
{ k &= 0xffff; n &= 0xffff;
  check_binom(&k,&n);
  const double lpe = log(pe), l1mpe = log(1-pe), mean = n*pe;
  double p, p_first, p_curr;
  if ((double)k >= mean)
    { p = p_first = exp(lp_binom_pre(M,k,n,lpe,l1mpe));
      for (int x = k+1; x <= n; x++)
        { p += p_curr = exp(lp_binom_pre(M,x,n,lpe,l1mpe));
          if (!exact && 10*p_curr < p_first) break;
        }
    }
  else
    { p = p_first = (k == 0) ? 0. : exp(lp_binom_pre(M,k-1,n,lpe,l1mpe));
      for (int x = k-2; x >= 0; x--)
        { p += p_curr = exp(lp_binom_pre(M,x,n,lpe,l1mpe));
          if (!exact && 10*p_curr < p_first) break;
        }
      p = 1-p;
    }
  return p;
}

/* util.c:35-44: cov is a 16-bit count in the reference signature */
static double lp_trans(const cpo_model *M, int b, int e, int cb, int ce, uint16_t cov)
{ return lp_skellam(ce-cb,(double)cov*abs(e-b)/M->read_len); }

/* util.c:46-55 */
static double p_errorin(const cpo_model *M, int etype, double erate, uint16_t cout, uint16_t cin)
{ if (!(cin <= cout)) { fprintf(stderr,"Violate cin (%d) <= cout (%d)\n",cin,cout); exit(1); }
  return cpo_binom_test_g(M,(etype == CPO_SELF) ? cin : cout-cin,cout,erate,0);
}

/* util.c:24-33 */
static double lin_interp(int x, int p1, uint16_t c1, int p2, uint16_t c2)
{ if (!(p1 < x && x < p2))
    { fprintf(stderr,"Invalid points for interpolation: x1=%d, x=%d, x2=%d\n",p1,x,p2); exit(1); }
  return (double)c1+((double)c2-c1)*(x-p1)/(p2-p1);
}

/*********************************************************************************************
 *  Host one-shot model
 *********************************************************************************************/

static void model_tables(cpo_model *M)
{ /* prob.c:14-19 */
  M->logfact[0] = 0.;
  for (int n = 1; n <= CPO_MAX_CNT; n++)
    M->logfact[n] = M->logfact[n-1]+log((double)n);
  /* ClassPro.c:544-548, util.c:9-11 */
  M->cov[CPO_E] = 1;
  M->cov[CPO_R] = (uint16_t)(M->cov[CPO_D]+(uint16_t)(sqrt((double)M->cov[CPO_D])*N_SIGMA_RCOV));
  M->dr_ratio = 1.+(double)N_SIGMA_R*(1./sqrt((double)M->cov[CPO_D]));
}

/* wall.c:120-244 (default closed-form error model; -M is out of scope: parity unpinned, no GSL) */
static int model_thresholds(cpo_model *M)
{ if (M->cov[CPO_R] > 255)
    { fprintf(stderr,"Too high REPEAT coverage (%d) > 255\n",M->cov[CPO_R]); return 1; }
  M->cmax = (uint8_t)M->cov[CPO_R];
  memset(M->cthres,0,sizeof(M->cthres));
  for (int t = 0; t < CPO_NCTYPE; t++)
    { M->lmax[t] = (uint8_t)(CPO_LMAX0/(t+1));
      M->pe[t][0] = 0.;
      for (int l = 1; l <= M->lmax[t]; l++)
        M->pe[t][l] = 0.002*l*l+0.002;
    }
  M->hc_erate = M->pe[CPO_HP][1];
  for (int t = 0; t < CPO_NCTYPE; t++)
    for (int l = 1; l <= M->lmax[t]; l++)
      { double pe = M->pe[t][l], lpe = log(pe), l1mpe = log(1-pe);
        for (int cout = 1; cout < M->cmax; cout++)
          { int found[2][2] = {{0,0},{0,0}};
            uint8_t ct[2];
            ct[CPO_SELF] = (uint8_t)cout; ct[CPO_OTHERS] = 0;
            for (int s = 0; s < 2; s++)
              for (int e = 0; e < 2; e++)
                M->cthres[t][l][cout][s][e] = ct[e];
            double psum = 1.;
            for (int cin = 0; cin <= cout; cin++)
              { if (found[0][0] && found[1][0] && found[0][1] && found[1][1]) break;
                ct[CPO_SELF] = (uint8_t)cin; ct[CPO_OTHERS] = (uint8_t)(cout-cin);
                psum -= exp(lp_binom_pre(M,cin,cout,lpe,l1mpe));
                for (int s = 0; s < 2; s++)
                  for (int e = 0; e < 2; e++)
                    if (!found[s][e] && psum < PE_THRES[s][e])
                      { M->cthres[t][l][cout][s][e] = ct[e]; found[s][e] = 1; }
              }
          }
      }
  return 0;
}

int cpo_model_from_cov(cpo_model *M, int kmer, int h, int d, int read_len)
{ memset(M,0,sizeof(*M));
  M->K = kmer; M->read_len = read_len;
  M->cov[CPO_D] = (uint16_t)d;
  M->cov[CPO_H] = (uint16_t)(h > 0 ? h : (d >> 1));
  model_tables(M);
  return model_thresholds(M);
}

/* hist.c:28-143 on top of libfastk.c:51-147 (Load_Histogram + Modify_Histogram(low,high,0)) */
int cpo_model_from_hist(cpo_model *M, int kmer, int low, int high, int64_t ilowcnt, int64_t ihighcnt,
                        const int64_t *raw, int cov_opt, int read_len, int verbose)
{ memset(M,0,sizeof(*M));
  M->K = kmer; M->read_len = read_len;
  int Hc, Dc;
  if (verbose) fprintf(stderr,"Global histogram inspection:\n");
  if (cov_opt > 0)
    { Dc = cov_opt; Hc = cov_opt >> 1;
      if (verbose) fprintf(stderr,"    Specified (H,D) cov   = (%d,%d)\n",Hc,Dc);
    }
  else
    { /* instance-count histogram: interior bins times their count, boundary bins swapped with
         the hidden instance totals (libfastk.c:22-47) */
      int64_t *hist = malloc(sizeof(int64_t)*(size_t)(high-low+3));
      if (hist == NULL) return 1;
      int64_t *h = hist-low;
      for (int i = low; i <= high; i++) h[i] = raw[i-low];
      for (int i = low+1; i < high; i++) h[i] *= i;
      h[high+1] = h[low];  h[low]  = ilowcnt;
      h[high+2] = h[high]; h[high] = ihighcnt;

      int maxcnt = 0; int64_t maxpk = 0;
      for (int i = MAXI(2,low); i < MINI(1000,high); i++)
        if (h[i-1] < h[i] && h[i] > h[i+1] && maxpk < h[i])
          { maxcnt = i; maxpk = h[i]; }
      if (maxcnt < 10)
        { fprintf(stderr,"[ERROR] Could not find any peak count >= 10 in the histogram. Revise data and use the `-c` option.");
          free(hist); return 2;
        }
      if (verbose)
        fprintf(stderr,"    Tallest peak count    = %d (# of k-mers = %lld)\n",maxcnt,(long long)maxpk);
      double m = (double)maxcnt/2, s = sqrt(m);
      int lmaxcnt = 0, is_lpeak = 0; int64_t lmaxpk = 0;
      for (int i = (int)round(m-s); i <= (int)round(m+s); i++)
        if (lmaxpk < h[i])
          { lmaxcnt = i; lmaxpk = h[i]; is_lpeak = (h[i-1] < h[i] && h[i] > h[i+1]) ? 1 : 0; }
      m = (double)maxcnt*2; s = sqrt(m);
      int rmaxcnt = 0, is_rpeak = 0; int64_t rmaxpk = 0;
      for (int i = (int)round(m-s); i <= (int)round(m+s); i++)
        if (rmaxpk < h[i])
          { rmaxcnt = i; rmaxpk = h[i]; is_rpeak = (h[i-1] < h[i] && h[i] > h[i+1]) ? 1 : 0; }
      if (lmaxpk > rmaxpk) { Dc = maxcnt; Hc = is_lpeak ? lmaxcnt : (maxcnt >> 1); }
      else                 { Hc = maxcnt; Dc = is_rpeak ? rmaxcnt : (maxcnt << 1); }
      if (verbose) fprintf(stderr,"    Estimated (H,D) cov   = (%d,%d)\n",Hc,Dc);
      free(hist);
    }
  M->cov[CPO_H] = (uint16_t)Hc;
  M->cov[CPO_D] = (uint16_t)Dc;
  model_tables(M);
  if (verbose) fprintf(stderr,"    Estimated R-threshold = %d\n",M->cov[CPO_R]);
  return model_thresholds(M);
}

/*********************************************************************************************
 *  Profile codec: libfastk.c:1467-1535
 *********************************************************************************************/
int cpo_decode_profile(const uint8_t *p, int64_t len, uint16_t *out, int cap)
{ if (len == 0) return 0;
  const uint8_t *q = p+len;
  uint16_t x = *p++, d;
  if (x & 0x80) d = (uint16_t)(((x & 0x7f) << 8) | *p++);
  else d = x;
  int n = 1;
  if (cap > 0) out[0] = d;
  while (p < q)
    { x = *p++;
      if ((x & 0xc0) == 0)                      /* run of x more copies */
        { for (int i = 0; i < x; i++, n++)
            if (n < cap) out[n] = d;
        }
      else
        { if (x & 0x80)                         /* 15-bit two's-complement delta, masked sum */
            { if (x & 0x40) x = (uint16_t)(x << 8);
              else          x = (uint16_t)((x << 8) & 0x7fff);
              x |= *p++;
              d = (uint16_t)((d+x) & 0x7fff);
            }
          else if (x & 0x20)                    /* 6-bit negative delta, 16-bit wrap */
            d = (uint16_t)(d+((x & 0x1fu) | 0xffe0u));
          else
            d = (uint16_t)(d+(x & 0x1fu));
          if (n < cap) out[n] = d;
          n++;
        }
    }
  return n;
}

/*********************************************************************************************
 *  Sequence context: context.c:8-108.  lctx[0] = {1,0,0} and lctx[1][TS] = 0 are set once by
 *  the caller (ClassPro.c:139-140) and never rewritten; rctx is never cleared.
 *********************************************************************************************/
void cpo_seq_context(uint8_t (*lctx)[3], uint8_t (*rctx)[3], const char *seq, int rlen)
{ int in_hp, in_ds = 0, in_ts = 0;
  const int last = rlen-1;
  for (int i = 1; i < rlen; i++)
    { in_hp = (seq[i-1] == seq[i]);
      in_ds = in_ts = 0;
      if (in_hp)
        { lctx[i][CPO_HP] = (uint8_t)MINI(lctx[i-1][CPO_HP]+1,127);
          lctx[i][CPO_DS] = rctx[i-1][CPO_DS] = 0;
        }
      else
        { lctx[i][CPO_HP] = 1;
          lctx[i][CPO_DS] = rctx[i-1][CPO_DS] = 1;
          /* the homopolymer that just ended: mirror its left lengths into right lengths */
          for (int j = i-lctx[i-1][CPO_HP], n = 0; j < i; j++, n++)
            rctx[j][CPO_HP] = lctx[i-1-n][CPO_HP];
          if (i >= 3 && seq[i-3] == seq[i-1] && seq[i-2] == seq[i])
            { lctx[i][CPO_DS] = (uint8_t)MINI(lctx[i-2][CPO_DS]+1,127);
              in_ds = 1;
            }
        }
      if (!in_ds)
        { int l = i-1;
          while (lctx[l][CPO_DS] > 1) l--;
          if (l < i-1)
            for (int j = l-1, n = 0; j < i; j++, n++)
              rctx[j-1][CPO_DS] = lctx[i-1-n][CPO_DS];
        }
      if (i >= 2)
        { if (in_hp && seq[i-2] == seq[i-1])
            lctx[i][CPO_TS] = rctx[i-2][CPO_TS] = 0;
          else if (i >= 5 && seq[i-5] == seq[i-2] && seq[i-4] == seq[i-1] && seq[i-3] == seq[i])
            { lctx[i][CPO_TS] = (uint8_t)MINI(lctx[i-3][CPO_TS]+1,127);
              in_ts = 1;
            }
          else
            lctx[i][CPO_TS] = rctx[i-1][CPO_TS] = rctx[i-2][CPO_TS] = 1;
          if (!in_ts)
            { int l = i-1;
              while (lctx[l][CPO_TS] > 1) l--;
              if (l < i-1)
                for (int j = l-2, n = 0; j < i; j++, n++)
                  rctx[j-2][CPO_TS] = lctx[i-1-n][CPO_TS];
            }
        }
    }
  for (int j = rlen-lctx[last][CPO_HP], n = 0; j < rlen; j++, n++)
    rctx[j][CPO_HP] = lctx[last-n][CPO_HP];
  if (in_ds)
    { int l = last;
      while (lctx[l][CPO_DS] > 1) l--;
      if (l < last)
        for (int j = l-1, n = 0; j < rlen; j++, n++)
          rctx[j-1][CPO_DS] = lctx[last-n][CPO_DS];
    }
  if (in_ts)
    { int l = last;
      while (lctx[l][CPO_TS] > 1) l--;
      if (l < last)
        for (int j = l-2, n = 0; j < rlen; j++, n++)
          rctx[j-2][CPO_TS] = lctx[last-n][CPO_TS];
    }
  rctx[last][CPO_DS] = rctx[last][CPO_TS] = rctx[rlen-2][CPO_TS] = 0;
}

/*********************************************************************************************
 *  Per-read scratch
 *********************************************************************************************/
typedef struct { int b, e; double pe; } errintvl;          /* ClassPro.h:153-157 */
typedef struct { int pos; uint16_t cnt; } poscnt;           /* ClassPro.h:201-204 */

#define MCAP 2048      /* reliable intervals per read: each spans >= K positions, 60000/40 = 1500 */

struct cpo_work
  { int       clean;
    /* context (ClassPro.c:136-142) */
    uint8_t (*lctx_base)[3];        /* _lctx */
    uint8_t (*rctx)[3];
    /* wall detection (ClassPro.h:172-177) */
    uint8_t  *wall;                 /* [MAX_RLEN+1] flag bytes */
    double  (*perror)[2][2];        /* [MAX_RLEN+1][etype][wtype] */
    errintvl *eintvl, *ointvl;
    /* intervals (ClassPro.c:132-133) */
    cpo_intvl *intvl, *rintvl;
    int       N, M;
    /* reliable DP (ClassPro.h:210-219) */
    uint16_t  COV[4];
    int       forward;
    double   *dp;                   /* [MCAP*4] */
    poscnt  (*st)[4];               /* [MCAP*4][4] */
    char    **bt;                   /* [MCAP*4+1] rows of MCAP */
    double   *dh_ratio;             /* [MCAP*4] */
    uint8_t  *rpos;                 /* [MCAP] */
    cpo_intvl *wintvl;              /* working copy */
    /* decoded profile + class string for the file driver */
    uint16_t *profile;
    char     *rasgn;
  };

enum { W_BY_S = 0x01, W_BY_O = 0x10, W_PAIR_S = 0x02, W_PAIR_O = 0x20,
       W_PAIR_MULT = 0x40, W_ERROR = 0x80 };                 /* wall.c:264-269 */
static const uint8_t W_BY[2]   = { W_BY_S, W_BY_O };
static const uint8_t W_PAIR[2] = { W_PAIR_S, W_PAIR_O };

cpo_work *cpo_work_new(void)
{ cpo_work *W = calloc(1,sizeof(cpo_work));
  const size_t R = CPO_MAX_RLEN;
  W->lctx_base = calloc(R+64,3);
  W->rctx      = calloc(R+64,3);
  W->wall      = calloc(R+1,1);
  W->perror    = calloc(R+1,sizeof(double[2][2]));
  W->eintvl    = calloc(R+1,sizeof(errintvl));
  W->ointvl    = calloc(R+1,sizeof(errintvl));
  W->intvl     = calloc(R+1,sizeof(cpo_intvl));
  W->rintvl    = calloc(R+1,sizeof(cpo_intvl));
  W->dp        = calloc(MCAP*4,sizeof(double));
  W->st        = calloc(MCAP*4,sizeof(poscnt[4]));
  W->bt        = calloc(MCAP*4+4,sizeof(char*));
  for (int i = 0; i < MCAP*4+4; i++) W->bt[i] = calloc(MCAP,1);
  W->dh_ratio  = calloc(MCAP*4,sizeof(double));
  W->rpos      = calloc(MCAP,1);
  W->wintvl    = calloc(MCAP,sizeof(cpo_intvl));
  W->profile   = calloc(R+1,sizeof(uint16_t));
  W->rasgn     = calloc(R+2,1);
  W->lctx_base[0][CPO_HP] = 1;                                 /* ClassPro.c:139-140 */
  W->lctx_base[0][CPO_DS] = W->lctx_base[0][CPO_TS] = W->lctx_base[1][CPO_TS] = 0;
  return W;
}

void cpo_work_free(cpo_work *W)
{ if (W == NULL) return;
  free(W->lctx_base); free(W->rctx); free(W->wall); free(W->perror); free(W->eintvl); free(W->ointvl);
  free(W->intvl); free(W->rintvl); free(W->dp); free(W->st);
  for (int i = 0; i < MCAP*4+4; i++) free(W->bt[i]);
  free(W->bt); free(W->dh_ratio); free(W->rpos); free(W->wintvl); free(W->profile); free(W->rasgn);
  free(W);
}

void cpo_work_set_clean(cpo_work *W, int clean) { W->clean = clean; }

const cpo_intvl *cpo_last_intervals(const cpo_work *W, int *N, int *Mrel)
{ if (N) *N = W->N;
  if (Mrel) *Mrel = W->M;
  return W->intvl;
}

/*********************************************************************************************
 *  Wall detection: wall.c:264-958
 *********************************************************************************************/
typedef struct
  { const cpo_model *M;
    cpo_work        *W;
    const uint16_t  *prof;
    int              plen;
    const uint8_t  (*ctx[2])[3];     /* ctx[DROP] = _lctx+K-2, ctx[GAIN] = rctx (ClassPro.c:138-142) */
  } wallctx;

/* wall.c:310-315: first writer wins, with the caller's error rate */
static void perror_once(wallctx *C, int i, int e, int w, uint16_t cout, uint16_t cin, double erate)
{ if (C->W->perror[i][e][w] == -INFINITY)
    C->W->perror[i][e][w] = p_errorin(C->M,e,erate,cout,cin);
}

/* wall.c:317-322 */
static double lp_diff_pair(wallctx *C, int i, int j)
{ const uint16_t *p = C->prof;
  int n_drop = (int)p[i-1]-p[i], n_gain = (int)p[j]-p[j-1];
  uint16_t cov = MAXI(p[i-1],p[j]);
  return lp_trans(C->M,i,j,n_drop,n_gain,cov);
}

/* wall.c:324-329; the reference passes cin through an 8-bit parameter */
static int thres_ng(int e, uint8_t cin, uint8_t ct)
{ return (e == CPO_SELF) ? (cin >= ct) : (cin < ct); }

/* wall.c:331-416: given a DROP at i, look for the matching GAIN about K-1 positions ahead */
static int pair_after_drop(wallctx *C, int i, uint16_t cout, uint16_t cin, int e, int t, int l,
                           double erate, errintvl *out)
{ const cpo_model *M = C->M;
  const uint16_t *prof = C->prof;
  const int plen = C->plen, K = M->K, ipk = i+K-1, ulen = t+1;
  double (*perr)[2][2] = C->W->perror;
  int max_j = -1; double max_pe = -INFINITY, pe;

  int m = ulen*l, n = 0, j;
  for (;;)
    { int idx = i+ulen*(n+1);
      if (idx >= plen || C->ctx[CPO_DROP][idx][t] != m+n+1) break;
      n++;
    }
  j = ipk+n-m;
  if (j <= i) return 0;
  if (j >= plen)
    { j = plen;
      pe = perr[i][e][CPO_DROP]*perr[i][e][CPO_DROP];
    }
  else
    { uint16_t cin_j = prof[j-1], cout_j = prof[j];
      pe = -INFINITY;
      if (cin_j <= cout_j
          && !(cout_j < M->cmax && thres_ng(e,(uint8_t)cin_j,M->cthres[t][l][cout_j][CPO_FINAL][e]))
          && (e == CPO_SELF || lp_diff_pair(C,i,j) >= THRES_DIFF_EO))
        { perror_once(C,j,e,CPO_GAIN,cout_j,cin_j,erate);
          pe = perr[i][e][CPO_DROP]*perr[j][e][CPO_GAIN];
        }
    }
  if (max_pe < pe) { max_j = j; max_pe = pe; }

  for (n = 0; n <= MAX_N_HC; n++)
    { j = ipk+n;
      if (j >= plen) break;
      uint16_t cin_j = prof[j-1], cout_j = prof[j];
      if (!(cin_j <= cout_j)) continue;
      if ((cout < M->cmax && thres_ng(e,(uint8_t)cin,M->cthres[CPO_HP][1][cout][CPO_FINAL][e]))
          || (cout_j < M->cmax && thres_ng(e,(uint8_t)cin_j,M->cthres[CPO_HP][1][cout_j][CPO_FINAL][e])))
        continue;
      if (e == CPO_OTHERS && lp_diff_pair(C,i,j) < THRES_DIFF_EO) continue;
      double pe_i = p_errorin(M,e,M->hc_erate,cout,cin);
      double pe_j = p_errorin(M,e,M->hc_erate,cout_j,cin_j);
      pe = pe_i*pe_j;
      if (max_pe < pe) { max_j = j; max_pe = pe; }
    }
  if (max_j == -1) return 0;
  out->b = i; out->e = max_j; out->pe = max_pe;
  return 1;
}

/* wall.c:418-507: given a GAIN at i, look for the matching DROP about K-1 positions behind */
static int pair_before_gain(wallctx *C, int i, uint16_t cout, uint16_t cin, int e, int t, int l,
                            double erate, errintvl *out)
{ const cpo_model *M = C->M;
  const uint16_t *prof = C->prof;
  const int K = M->K, imk = i-K+1, ulen = t+1;
  double (*perr)[2][2] = C->W->perror;
  int max_j = -1; double max_pe = -INFINITY, pe;

  int m = ulen*l, n = 0, j;
  for (;;)
    { int idx = i-ulen*(n+1);
      if (idx <= 0) break;
      if (C->ctx[CPO_GAIN][idx][t] != m+n+1) break;
      n++;
    }
  j = imk-n+m;
  if (j >= i) return 0;
  if (j <= 0)
    { j = 0;
      pe = perr[i][e][CPO_GAIN]*perr[i][e][CPO_GAIN];
    }
  else
    { uint16_t cout_j = prof[j-1], cin_j = prof[j];
      pe = -INFINITY;
      if (cin_j <= cout_j
          && !(cout_j < M->cmax && thres_ng(e,(uint8_t)cin_j,M->cthres[t][l][cout_j][CPO_FINAL][e]))
          && (e == CPO_SELF || lp_diff_pair(C,j,i) >= THRES_DIFF_EO))
        { perror_once(C,j,e,CPO_DROP,cout_j,cin_j,erate);
          pe = perr[j][e][CPO_DROP]*perr[i][e][CPO_GAIN];
        }
    }
  if (max_pe < pe) { max_j = j; max_pe = pe; }

  for (n = 0; n <= MAX_N_HC; n++)
    { j = imk-n;
      if (j <= 0) break;
      uint16_t cout_j = prof[j-1], cin_j = prof[j];
      if (!(cin_j <= cout_j)) continue;
      if ((cout < M->cmax && thres_ng(e,(uint8_t)cin,M->cthres[CPO_HP][1][cout][CPO_FINAL][e]))
          || (cout_j < M->cmax && thres_ng(e,(uint8_t)cin_j,M->cthres[CPO_HP][1][cout_j][CPO_FINAL][e])))
        continue;
      if (e == CPO_OTHERS && lp_diff_pair(C,j,i) < THRES_DIFF_EO) continue;
      double pe_i = p_errorin(M,e,M->hc_erate,cout,cin);
      double pe_j = p_errorin(M,e,M->hc_erate,cout_j,cin_j);
      pe = pe_i*pe_j;
      if (max_pe < pe) { max_j = j; max_pe = pe; }
    }
  if (max_j == -1) return 0;
  out->b = max_j; out->e = i; out->pe = max_pe;
  return 1;
}

/* wall.c:519-528 with glibc's stable merge sort: the (int) cast of a probability difference in
 * (-1,1) is 0, so the order is (b,e) then input order.  Stable insertion sort here. */
static int ei_less(const errintvl *x, const errintvl *y)   /* strictly less in the qsort order */
{ if (x->b == y->b)
    { if (x->e == y->e) return ((int)(y->pe-x->pe)) < 0;
      return x->e-y->e < 0;
    }
  return x->b-y->b < 0;
}

static void ei_sort(errintvl *a, int n)
{ for (int i = 1; i < n; i++)
    { errintvl v = a[i];
      int j = i-1;
      while (j >= 0 && ei_less(&v,&a[j])) { a[j+1] = a[j]; j--; }
      a[j+1] = v;
    }
}

/* wall.c:548-568 */
static int ei_unique(errintvl *a, int n)
{ ei_sort(a,n);
  if (n >= 2)
    { int i = 1;
      while (i < n && !(a[i-1].b == a[i].b && a[i-1].e == a[i].e)) i++;
      for (int j = i+1; j < n; j++)
        if (!(a[i-1].b == a[j].b && a[i-1].e == a[j].e))
          a[i++] = a[j];
      n = i;
    }
  return n;
}

/* wall.c:530-546 */
static int ei_find(const errintvl *a, int l, int r, int b, int e)
{ while (l <= r)
    { int m = (l+r)/2;
      if (a[m].b == b)
        { if (a[m].e == e) return m;
          if (e > a[m].e) l = m+1; else r = m-1;
        }
      else if (b > a[m].b) l = m+1;
      else r = m-1;
    }
  return -1;
}

#define EI_CHECK(n) do { if ((n) >= plen) { fprintf(stderr,"# E-intvls >= plen\n"); exit(1); } } while (0)

/* wall.c:570-958 */
static int find_walls(const cpo_model *M, cpo_work *W, const uint16_t *prof, int plen)
{ const int K = M->K;
  uint8_t *wall = W->wall;
  double (*perr)[2][2] = W->perror;
  errintvl *eint = W->eintvl, *oint = W->ointvl;
  cpo_intvl *intvl = W->intvl;
  wallctx C;
  C.M = M; C.W = W; C.prof = prof; C.plen = plen;
  C.ctx[CPO_DROP] = (const uint8_t (*)[3])(W->lctx_base+K-2);
  C.ctx[CPO_GAIN] = (const uint8_t (*)[3])W->rctx;

  const int reset_to = W->clean ? plen+1 : plen;
  for (int i = 0; i < reset_to; i++)
    { wall[i] = 0;
      for (int e = 0; e < 2; e++)
        for (int w = 0; w < 2; w++)
          perr[i][e][w] = -INFINITY;
    }

  /* pass A (wall.c:588-707) */
  uint8_t ct[2] = {0,0};
  int eidx = 0, oidx = 0;
  for (int i = 1; i < plen; i++)
    { uint16_t cim1 = prof[i-1], ci = prof[i];
      if (MINI(cim1,ci) >= M->cov[CPO_R]) continue;
      uint16_t cng = (uint16_t)abs((int)cim1-ci);
      if (cng < MIN_CNT_CHANGE) continue;
      int wtype; uint16_t cin, cout;
      if (cim1 > ci) { wtype = CPO_DROP; cin = ci;   cout = cim1; }
      else           { wtype = CPO_GAIN; cin = cim1; cout = ci;   }

      int maxt = -1, maxl = -1; double maxpe = -INFINITY;
      for (int t = 0; t < CPO_NCTYPE; t++)
        { int l = MINI(C.ctx[wtype][i][t],M->lmax[t]);
          double pe = M->pe[t][l];
          if (maxpe < pe) { maxpe = pe; maxt = t; maxl = l; }
        }

      for (int e = CPO_SELF; e <= CPO_OTHERS; e++)
        { if (wall[i] & W_PAIR[e]) continue;
          if (cout < M->cmax)
            { for (int s = 0; s < 2; s++) ct[s] = M->cthres[maxt][maxl][cout][s][e];
              if (!(cng > MAX_CNT_CHANGE || cin < MAXI(ct[CPO_INIT],3))) continue;
            }
          errintvl I;
          int found;
          if (e == CPO_SELF)
            { if (cout < M->cmax && cin >= ct[CPO_FINAL]) continue;
              perror_once(&C,i,e,wtype,cout,cin,maxpe);
              if (perr[i][e][wtype] < PE_THRES[CPO_FINAL][e]) continue;
              found = (wtype == CPO_DROP) ? pair_after_drop(&C,i,cout,cin,e,maxt,maxl,maxpe,&I)
                                          : pair_before_gain(&C,i,cout,cin,e,maxt,maxl,maxpe,&I);
              if (found && I.pe >= PE_THRES[CPO_FINAL][e])
                { wall[I.b] |= W_BY[e];   wall[I.e] |= W_BY[e];
                  wall[I.b] |= W_PAIR[e]; wall[I.e] |= W_PAIR[e];
                  eint[eidx++] = I;
                }
            }
          else
            { if (cng >= M->cov[CPO_H] || (cout < M->cmax && cin < ct[CPO_FINAL]))
                { wall[i] |= W_BY_O; continue; }
              perror_once(&C,i,e,wtype,cout,cin,maxpe);
              if (perr[i][e][wtype] < PE_THRES[CPO_FINAL][e])
                { wall[i] |= W_BY_O; continue; }
              found = (wtype == CPO_DROP) ? pair_after_drop(&C,i,cout,cin,e,maxt,maxl,maxpe,&I)
                                          : pair_before_gain(&C,i,cout,cin,e,maxt,maxl,maxpe,&I);
              if (found && I.pe >= PE_THRES[CPO_FINAL][e])
                { wall[I.b] |= W_PAIR[e]; wall[I.e] |= W_PAIR[e];
                  oint[oidx++] = I;
                  continue;
                }
              wall[i] |= W_BY_O;
            }
        }
    }
  int NS = eidx, NO = oidx;

  /* pass B (wall.c:721-735) */
  for (int i = 0; i < NO; i++)
    { wall[oint[i].b] &= (uint8_t)~W_BY_O; wall[oint[i].e] &= (uint8_t)~W_BY_O; }
  for (int i = 0; i < NS; i++)
    for (int j = eint[i].b+1; j < eint[i].e; j++)
      wall[j] &= (uint8_t)~W_BY_O;
  NS = ei_unique(eint,eidx);
  NO = ei_unique(oint,oidx);
  (void)NO;

  /* pass C (wall.c:759-860): E-intervals made of several errors, and boundary E-intervals */
  int midx = NS;
  const double TH = PE_THRES[CPO_FINAL][CPO_SELF];
  for (int i = 1; i < plen; i++)
    { if (!((wall[i] & W_BY_O) && !(wall[i] & W_BY_S))) continue;
      if (wall[i] & W_PAIR_MULT) continue;
      for (int w = CPO_DROP; w <= CPO_GAIN; w++)
        { double pe_i = perr[i][CPO_SELF][w], pe_j, pe;
          if (pe_i < TH) continue;
          if (w == CPO_DROP)
            { for (int j = i+1; j < MINI(i+200,plen+1); j++)
                { if (j == plen)
                    { if ((pe = pe_i*pe_i) < TH) continue;
                      eint[midx].b = i; eint[midx].e = plen; eint[midx].pe = pe;
                      wall[i] |= W_PAIR_MULT;
                      midx++; EI_CHECK(midx);
                    }
                  if (!(wall[j] & W_BY_S) && !(wall[j] & W_BY_O)) continue;
                  if (ei_find(eint,0,NS-1,i,j) == -1)
                    { pe_j = perr[j][CPO_SELF][CPO_GAIN];
                      if ((pe = pe_i*pe_j) >= TH)
                        { eint[midx].b = i; eint[midx].e = j; eint[midx].pe = pe;
                          wall[i] |= W_PAIR_MULT; wall[j] |= W_PAIR_MULT;
                          midx++; EI_CHECK(midx);
                        }
                    }
                  if (wall[j] & W_BY_O) break;
                }
            }
          else
            { for (int j = i-1; j >= MAXI(i-200,0); j--)
                { if (j == 0)
                    { if ((pe = pe_i*pe_i) < TH) continue;
                      eint[midx].b = 0; eint[midx].e = i; eint[midx].pe = pe;
                      wall[i] |= W_PAIR_MULT;
                      midx++; EI_CHECK(midx);
                    }
                  if (!(wall[j] & W_BY_S) && !(wall[j] & W_BY_O)) continue;
                  if (ei_find(eint,0,NS-1,j,i) == -1)
                    { pe_j = perr[j][CPO_SELF][CPO_DROP];
                      if ((pe = pe_i*pe_j) >= TH)
                        { eint[midx].b = j; eint[midx].e = i; eint[midx].pe = pe;
                          wall[i] |= W_PAIR_MULT; wall[j] |= W_PAIR_MULT;
                          midx++; EI_CHECK(midx);
                        }
                    }
                  if (wall[j] & W_BY_O) break;
                }
            }
        }
    }
  for (int i = NS; i < midx; i++)
    for (int j = eint[i].b+1; j < eint[i].e; j++)
      wall[j] &= (uint8_t)~W_BY_O;
  if (NS < midx) { NS = midx; ei_sort(eint,NS); }

  /* pass D (wall.c:877-919): append the hull of every chain of overlapping E-intervals; the
     loop bound is re-read, so appended hulls are visited too */
  { int i = 0;
    while (i < NS-1)
      { int max_e = eint[i].e; double max_pe = eint[i].pe;
        int j = i;
        while (j < NS-1 && eint[j+1].b <= eint[j].e)
          { max_e = MAXI(max_e,eint[j+1].e);
            max_pe = (max_pe > eint[j+1].pe) ? max_pe : eint[j+1].pe;
            j++;
          }
        if (i < j)
          { eint[NS].b = eint[i].b; eint[NS].e = max_e; eint[NS].pe = max_pe;
            NS++; EI_CHECK(NS);
          }
        i = j+1;
      }
  }
  ei_sort(eint,NS);
  for (int i = 0; i < NS; i++)
    for (int j = eint[i].b; j < eint[i].e; j++)
      wall[j] |= W_ERROR;

  /* pass E (wall.c:921-948) */
  int N = 0, b = 0;
  for (int i = 1; i <= plen; i++)
    if (i == plen || ((wall[i-1] & W_ERROR) != 0) != ((wall[i] & W_ERROR) != 0)
        || (!(wall[i] & W_ERROR) && (wall[i] & W_BY_O)))
      { int e = i;
        int k = ei_find(eint,0,NS-1,b,e);
        cpo_intvl *I = &intvl[N];
        I->b = b; I->e = e; I->cb = prof[b]; I->ce = prof[e-1];
        I->is_rel = 0;
        I->pe = (k != -1) ? log(eint[k].pe) : -INFINITY;
        double pob = (perr[b][CPO_OTHERS][CPO_DROP] > perr[b][CPO_OTHERS][CPO_GAIN])
                       ? perr[b][CPO_OTHERS][CPO_DROP] : perr[b][CPO_OTHERS][CPO_GAIN];
        double poe = (perr[e][CPO_OTHERS][CPO_DROP] > perr[e][CPO_OTHERS][CPO_GAIN])
                       ? perr[e][CPO_OTHERS][CPO_DROP] : perr[e][CPO_OTHERS][CPO_GAIN];
        I->pe_o_b = (pob != -INFINITY) ? log(pob) : -INFINITY;
        I->pe_o_e = (poe != -INFINITY) ? log(poe) : -INFINITY;
        I->asgn = CPO_NSTATE;
        N++;
        b = e;
      }
  return N;
}

/*********************************************************************************************
 *  Reliable intervals: wall.c:960-1051
 *********************************************************************************************/
/* wall.c:960-1014.  The inner loops of lines 999-1006 declare a position variable that hides the
 * interval index, so the max() adjustments are written to intvl[<position>]; emulated literally
 * on an array of MAX_RLEN interval slots. */
static void correct_wall_cnt(const cpo_model *M, cpo_work *W, int idx, const uint16_t *prof)
{ const int K = M->K;
  cpo_intvl *intvl = W->intvl;
  const cpo_intvl I = intvl[idx];
  const uint8_t (*lc)[3] = (const uint8_t (*)[3])(W->lctx_base+K-2);   /* ctx[DROP] */
  const uint8_t (*rc)[3] = (const uint8_t (*)[3])W->rctx;              /* ctx[GAIN] */
  int n_gain = 0, n_drop = 0, lmax, first, last;

  last = MINI(I.b+K-1,I.e-1);
  for (int p = I.b; p < last; p++) n_gain += MAXI((int)prof[p+1]-prof[p],0);
  if (I.b+K-1 < I.e)
    { lmax = 0;
      for (int t = 0; t < 3; t++) { int l = rc[I.b+K-1][t]*(t+1); if (lmax < l) lmax = l; }
      last = I.b+lmax;
      for (int p = I.b; p < last; p++) n_gain -= MAXI((int)prof[p]-prof[p+1],0);
    }
  first = MAXI(I.e-K+1,I.b);
  for (int p = first; p < I.e-1; p++) n_drop += MAXI((int)prof[p]-prof[p+1],0);
  if (I.b < I.e-K+1)
    { lmax = 0;
      for (int t = 0; t < 3; t++) { int l = lc[I.e-K+1][t]*(t+1); if (lmax < l) lmax = l; }
      first = I.e-lmax;
      for (int p = first; p < I.e-1; p++) n_drop -= MAXI((int)prof[p+1]-prof[p],0);
    }
  intvl[idx].ccb = (uint16_t)MINI(I.cb+MAXI(n_gain,0),CPO_MAX_CNT);
  intvl[idx].cce = (uint16_t)MINI(I.ce+MAXI(n_drop,0),CPO_MAX_CNT);

  last = MINI(I.b+2*K,I.e);
  for (int p = I.b; p < last; p++)
    if (intvl[p].ccb < prof[p]) intvl[p].ccb = prof[p];
  first = MAXI(I.e-2*K,I.b);
  for (int p = first; p < I.e; p++)
    if (intvl[p].cce < prof[p]) intvl[p].cce = prof[p];
}

/* wall.c:1016-1051 */
static int find_reliable(const cpo_model *M, cpo_work *W, int N, const uint16_t *prof)
{ cpo_intvl *intvl = W->intvl, *rintvl = W->rintvl;
  int Mrel = 0;
  const double logpthres = log(PE_THRES[CPO_FINAL][CPO_SELF]);
  for (int i = 0; i < N; i++)
    { if (intvl[i].e-intvl[i].b < M->K) continue;
      if (MAXI(intvl[i].cb,intvl[i].ce) >= M->cov[CPO_R]) continue;
      if (intvl[i].pe >= logpthres) continue;
      correct_wall_cnt(M,W,i,prof);
      if (lp_trans(M,intvl[i].b,intvl[i].e,intvl[i].ccb,intvl[i].cce,
                   (uint16_t)((intvl[i].ccb+intvl[i].cce)/2)) < THRES_DIFF_REL) continue;
      if (MAXI(intvl[i].ccb,intvl[i].cce) == CPO_MAX_CNT) continue;
      intvl[i].is_rel = 1;
      rintvl[Mrel++] = intvl[i];
    }
  return Mrel;
}

/*********************************************************************************************
 *  Reliable-interval DP: class_rel.c:41-963
 *********************************************************************************************/
#define RIDX(i,s) ((i)*4+(s))

static inline int pred_of(int x, int F)  { return F ? x-1 : x+1; }
static inline int succ_of(int x, int F)  { return F ? x+1 : x-1; }
static inline int off_pos(int x, int F)  { return F ? x-OFFSET : x+OFFSET; }
static inline int beg_pos(const cpo_intvl *I, int F) { return F ? I->b : I->e-1; }
static inline uint16_t beg_cnt(const cpo_intvl *I, int F) { return F ? I->ccb : I->cce; }
static inline int end_pos(const cpo_intvl *I, int F) { return F ? I->e-1 : I->b; }
static inline uint16_t end_cnt(const cpo_intvl *I, int F) { return F ? I->cce : I->ccb; }

/* Tie tracing: with cpo_set_trace(1) every strict comparison between two
   log-probabilities that decides an arg-max, and every double -> int truncation of a coverage
   ratio, is logged to stderr with its relative gap, so that a class character that differs on the
   GPU (CUDA exp/log vs glibc, <= 1 ulp) can be attributed to the comparison it sits on. */
static int cpo_trace = 0;
void cpo_set_trace(int on) { cpo_trace = on; }
static void trace_cmp(const char *where, int i, int a, int b, double x, double y)
{ if (!cpo_trace || x == -INFINITY || y == -INFINITY) return;
  double m = fabs(x) > fabs(y) ? fabs(x) : fabs(y);
  double g = (m > 0.) ? fabs(x-y)/m : 0.;
  if (g < 1e-6) fprintf(stderr,"TIE %s @%d: %d vs %d: %.17g vs %.17g rel.gap %.3g\n",where,i,a,b,x,y,g);
}
static void trace_trunc(const char *where, int i, double v)
{ if (!cpo_trace || !(v == v) || fabs(v) > 1e9) return;
  double f = v-floor(v), d = f < 0.5 ? f : 1.-f;
  if (d < 1e-6*fabs(v)+1e-12) fprintf(stderr,"TRUNC %s @%d: %.17g is within %.3g of an integer\n",where,i,v,d);
}

/* class_rel.c:62-73 */
static int best_state(const double *dp, int i)
{ double mx = -INFINITY; int ms = CPO_NSTATE;
  for (int s = 0; s < 4; s++)
    { if (ms != CPO_NSTATE) trace_cmp("best_state",i,ms,s,mx,dp[RIDX(i,s)]);
      if (mx < dp[RIDX(i,s)]) { mx = dp[RIDX(i,s)]; ms = s; }
    }
  return ms;
}

/* class_rel.c:80-96: s or t may be the wildcard CPO_NSTATE */
static int best_tr(const double *dp, double tr[4][4], int i, int s, int t, int F, double *out_logp)
{ int ip = pred_of(i,F);
  double mx = -INFINITY; int mxx = CPO_NSTATE;
  for (int x = 0; x < 4; x++)
    { int _s = (s < 4) ? s : x, _t = (t < 4) ? t : x;
      double lp = dp[RIDX(ip,_s)]+tr[_s][_t];
      if (mxx != CPO_NSTATE) trace_cmp(s < 4 ? "best_tr(to)" : "best_tr(from)",i*10+((s < 4) ? s : t),mxx,x,mx,lp);
      if (mx < lp) { mx = lp; mxx = x; }
    }
  if (out_logp) *out_logp = mx;
  return mxx;
}

/* class_rel.c:98-107 */
static int nearest_with(int forward, int i, int s, const char *asgn, int L)
{ int idx = i;
  if (forward) while (idx < L && asgn[idx] != (char)s) idx++;
  else         while (idx >= 0 && asgn[idx] != (char)s) idx--;
  return idx;
}

/* class_rel.c:113-156 */
static double dh_ratio_of(int init_s, const char *asgn, const cpo_intvl *intvl, int L, int F)
{ int idx[4];
  idx[0] = F ? L : -1;
  int s = init_s;
  for (int i = 0; i < 3; i++)
    { idx[i+1] = nearest_with(!F,pred_of(idx[i],F),s,asgn,L);
      if ((F && idx[i+1] < 0) || (!F && idx[i+1] >= L)) return -INFINITY;
      s = (s == CPO_H) ? CPO_D : CPO_H;
    }
  int s1p = beg_pos(&intvl[idx[1]],F); uint16_t s1c = beg_cnt(&intvl[idx[1]],F);
  int tp  = end_pos(&intvl[idx[2]],F); uint16_t tc  = end_cnt(&intvl[idx[2]],F);
  int s2p = end_pos(&intvl[idx[3]],F); uint16_t s2c = end_cnt(&intvl[idx[3]],F);
  if (!F) { int p = s1p; uint16_t c = s1c; s1p = s2p; s1c = s2c; s2p = p; s2c = c; }
  double est = lin_interp(tp,s2p,s2c,s1p,s1c);
  return (init_s == CPO_D) ? est/tc : tc/est;
}

/* class_rel.c:158-170 */
static double rel_lp_e(const cpo_model *M, const cpo_intvl *I, const uint16_t *COV)
{ double po = lp_poisson(M,I->ccb,COV[CPO_E])+lp_poisson(M,I->cce,COV[CPO_E])+E_PO_BASE;
  return (po > I->pe) ? po : I->pe;
}

/* class_rel.c:172-211 */
static double rel_lp_r(const cpo_model *M, const cpo_intvl *I, poscnt pr, int F, const uint16_t *COV)
{ uint16_t bc = beg_cnt(I,F);
  double sf = -INFINITY;
  double er = (bc < pr.cnt) ? lp_binom(M,bc,pr.cnt,1-PE_MEAN) : -INFINITY;
  double lp = (sf > er) ? sf : er;
  if (lp > R_LOGP) return lp;
  uint16_t mx = MAXI(I->ccb,I->cce);
  if (mx >= COV[CPO_R]) return R_LOGP;
  if (mx >= pr.cnt) return R_LOGP;
  return lp;
}

/* class_rel.c:213-240: the H-track transition is overwritten by the D-track one scaled by the
 * running D/H ratio whenever that ratio exists */
static double rel_lp_h(const cpo_model *M, cpo_work *W, int idx, int s, poscnt *sp, int F)
{ const cpo_intvl *I = &W->wintvl[idx];
  int bp = beg_pos(I,F); uint16_t bc = beg_cnt(I,F);
  poscnt st = sp[CPO_H];
  double sf = lp_trans(M,pred_of(st.pos,F),bp,st.cnt,bc,st.cnt);
  double r = W->dh_ratio[RIDX(pred_of(idx,F),s)];
  if (r != -INFINITY)
    { st = sp[CPO_D];
      trace_trunc("lp_h r*bc",idx,r*bc);
      sf = lp_trans(M,pred_of(st.pos,F),bp,st.cnt,(int)(r*bc),st.cnt);
    }
  return sf+0.;
}

/* class_rel.c:242-270: the ratio-scaled H-track value is computed and then discarded */
static double rel_lp_d(const cpo_model *M, cpo_work *W, int idx, int s, poscnt *sp, int F)
{ const cpo_intvl *I = &W->wintvl[idx];
  int bp = beg_pos(I,F); uint16_t bc = beg_cnt(I,F);
  double r = W->dh_ratio[RIDX(pred_of(idx,F),s)];
  if (r != -INFINITY)
    { poscnt st = sp[CPO_H];
      (void)lp_trans(M,pred_of(st.pos,F),bp,st.cnt,(int)((double)bc/r),st.cnt);
    }
  poscnt st = sp[CPO_D];
  double sf = lp_trans(M,pred_of(st.pos,F),bp,st.cnt,bc,st.cnt);
  return sf+0.;
}

/* class_rel.c:272-277 */
static double rel_lp(const cpo_model *M, cpo_work *W, int s, int t, int idx, poscnt *sp)
{ const int F = W->forward;
  if (t == CPO_E) return rel_lp_e(M,&W->wintvl[idx],W->COV);
  if (t == CPO_H) return rel_lp_h(M,W,idx,s,sp,F);
  if (t == CPO_D) return rel_lp_d(M,W,idx,s,sp,F);
  return rel_lp_r(M,&W->wintvl[idx],sp[CPO_R],F,W->COV);
}

/* class_rel.c:279-513 */
static void rel_update(const cpo_model *M, cpo_work *W, int i, int Mrel)
{ const int F = W->forward;
  const uint16_t *COV = W->COV;
  double *dp = W->dp; poscnt (*st)[4] = W->st; char **bt = W->bt;
  double *dhr = W->dh_ratio; cpo_intvl *intvl = W->wintvl;
  const cpo_intvl I = intvl[i];
  const int ep = end_pos(&I,F); const uint16_t ec = end_cnt(&I,F);
  const int ip = pred_of(i,F);

  double tr[4][4];
  for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) tr[s][t] = -INFINITY;
  for (int s = 0; s < 4; s++)
    { int idx = RIDX(ip,s);
      if (dp[idx] == -INFINITY)
        { for (int t = 0; t < 4; t++) tr[s][t] = 0.; continue; }
      for (int t = 0; t < 4; t++) tr[s][t] = exp(rel_lp(M,W,s,t,i,st[idx]));
    }
  double psum = 0.;
  for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) psum += tr[s][t];
  if (psum == 0.)
    { fprintf(stderr,"No possible state @ %d\n",i);
      for (int s = 0; s < 4; s++) tr[s][CPO_E] = 1.;
      psum = 4.;
    }
  for (int s = 0; s < 4; s++) for (int t = 0; t < 4; t++) tr[s][t] = log(tr[s][t]/psum);

  /* every live predecessor prefers R: freeze (class_rel.c:348-380) */
  int only_r = 1;
  for (int s = 0; s < 4; s++)
    { int mt = best_tr(dp,tr,i,s,CPO_NSTATE,F,NULL);
      if (mt != CPO_NSTATE && mt != CPO_R) { only_r = 0; break; }
    }
  if (only_r)
    { W->rpos[i] = 1;
      intvl[i] = intvl[ip];
      for (int s = 0; s < 4; s++)
        { int idx = RIDX(i,s), idp = RIDX(ip,s);
          dp[idx] = dp[idp];
          if (dp[idx] == -INFINITY) continue;
          if (F) for (int k = 0; k < i; k++) bt[idx][k] = bt[idp][k];
          else   for (int k = i+1; k < Mrel; k++) bt[idx][k] = bt[idp][k];
          bt[idx][i] = (char)s;
          for (int t = 0; t < 4; t++) st[idx][t] = st[idp][t];
        }
      return;
    }

  int mh = best_tr(dp,tr,i,CPO_NSTATE,CPO_H,F,NULL);
  int md = best_tr(dp,tr,i,CPO_NSTATE,CPO_D,F,NULL);
  if (mh == CPO_H && md == CPO_D)
    tr[CPO_H][CPO_H] = tr[CPO_D][CPO_D] = (tr[CPO_H][CPO_H] < tr[CPO_D][CPO_D]) ? tr[CPO_H][CPO_H] : tr[CPO_D][CPO_D];

  for (int t = 0; t < 4; t++)
    { double mlp;
      int ms = best_tr(dp,tr,i,CPO_NSTATE,t,F,&mlp);
      int idx = RIDX(i,t), idp = RIDX(ip,ms);
      dp[idx] = mlp;
      if (ms == CPO_NSTATE) continue;
      if (F) for (int k = 0; k < i; k++) bt[idx][k] = bt[idp][k];
      else   for (int k = i+1; k < Mrel; k++) bt[idx][k] = bt[idp][k];
      bt[idx][i] = (char)t;

      if (t == CPO_E)
        { for (int s = CPO_R; s <= CPO_D; s++) st[idx][s] = st[idp][s]; }
      else if (t == CPO_R)
        { for (int s = CPO_H; s <= CPO_D; s++)
            { st[idx][s].pos = off_pos(ep,F); st[idx][s].cnt = st[idp][s].cnt; }
          uint16_t rc = MINI(ec,COV[CPO_R]);
          if (st[idp][CPO_R].cnt < rc) st[idx][CPO_R] = st[idp][CPO_R];
          else { st[idx][CPO_R].pos = off_pos(ep,F); st[idx][CPO_R].cnt = rc; }
        }
      else
        { int ch, cd, cr;
          double r = dh_ratio_of(t,F ? bt[idx] : bt[idx]+i,F ? intvl : intvl+i,F ? i+1 : Mrel-i,F);
          int other = (t == CPO_H) ? CPO_D : CPO_H, has_other = 0;
          if (r == -INFINITY)
            { if (F) { for (int k = 0; k < i; k++) if (bt[idx][k] == other) has_other = 1; }
              else   { for (int k = i+1; k < Mrel; k++) if (bt[idx][k] == other) has_other = 1; }
            }
          if (t == CPO_H)
            { ch = ec;
              if (r == -INFINITY) cd = has_other ? st[idp][CPO_D].cnt : ch+COV[CPO_H];
              else { trace_trunc("r*ch",i,r*ch); cd = (int)(r*ch); dhr[idx] = r; }
            }
          else
            { cd = ec;
              if (r == -INFINITY) ch = has_other ? st[idp][CPO_H].cnt : MAXI(cd/2,cd-COV[CPO_H]);
              else { trace_trunc("cd/r",i,(double)cd/r); ch = (int)((double)cd/r); dhr[idx] = r; }
            }
          trace_trunc("dr_ratio*cd",i,M->dr_ratio*cd);
          cr = (int)(M->dr_ratio*cd);
          st[idx][CPO_H].pos = off_pos(ep,F); st[idx][CPO_H].cnt = (uint16_t)ch;
          st[idx][CPO_D].pos = off_pos(ep,F); st[idx][CPO_D].cnt = (uint16_t)cd;
          st[idx][CPO_R].pos = off_pos(ep,F); st[idx][CPO_R].cnt = (uint16_t)cr;
        }
      if (!(st[idx][CPO_H].cnt < st[idx][CPO_D].cnt && st[idx][CPO_D].cnt < st[idx][CPO_R].cnt))
        dp[idx] = -INFINITY;
    }
}

/* class_rel.c:515-614 */
static char *rel_pass(const cpo_model *M, cpo_work *W, int Mrel, int plen)
{ const int F = W->forward;
  const uint16_t *COV = W->COV;
  double *dp = W->dp; poscnt (*st)[4] = W->st; char **bt = W->bt;
  cpo_intvl *intvl = W->wintvl;
  for (int i = 0; i < Mrel; i++)
    { for (int s = 0; s < 4; s++) { dp[RIDX(i,s)] = -INFINITY; W->dh_ratio[RIDX(i,s)] = -INFINITY; }
      W->rpos[i] = 0;
      intvl[i] = W->rintvl[i];
    }
  const int POS_INIT = off_pos(F ? 0 : plen,F);
  int i = F ? 0 : Mrel-1;
  const cpo_intvl I = intvl[i];
  int idx;
  for (int s = 0; s < 4; s++)
    { idx = RIDX(i,s);
      for (int t = CPO_R; t <= CPO_D; t++) { st[idx][t].pos = POS_INIT; st[idx][t].cnt = COV[t]; }
      bt[idx][i] = (char)s;
    }
  idx = RIDX(i,CPO_E);
  dp[idx] = rel_lp_e(M,&intvl[i],COV);
  idx = RIDX(i,CPO_R);
  dp[idx] = rel_lp_r(M,&intvl[i],st[idx][CPO_R],F,COV);
  st[idx][CPO_R].pos = end_pos(&I,F);
  st[idx][CPO_R].cnt = MINI(end_cnt(&I,F),COV[CPO_R]);
  idx = RIDX(i,CPO_H);
  dp[idx] = lp_poisson(M,beg_cnt(&I,F),COV[CPO_H]);
  st[idx][CPO_H].pos = end_pos(&I,F);
  st[idx][CPO_H].cnt = end_cnt(&I,F);
  st[idx][CPO_D].pos = off_pos(end_pos(&I,F),F);
  st[idx][CPO_D].cnt = (uint16_t)(end_cnt(&I,F)+COV[CPO_H]);
  idx = RIDX(i,CPO_D);
  dp[idx] = lp_poisson(M,beg_cnt(&I,F),COV[CPO_D]);
  st[idx][CPO_H].pos = off_pos(end_pos(&I,F),F);
  st[idx][CPO_H].cnt = (uint16_t)MAXI(end_cnt(&I,F)/2,(int)end_cnt(&I,F)-COV[CPO_H]);
  st[idx][CPO_D].pos = end_pos(&I,F);
  st[idx][CPO_D].cnt = end_cnt(&I,F);

  double psum = 0.;
  for (int s = 0; s < 4; s++) psum += exp(dp[RIDX(i,s)]);
  for (int s = 0; s < 4; s++) dp[RIDX(i,s)] = log(exp(dp[RIDX(i,s)])/psum);

  for (;;)
    { i = succ_of(i,F);
      if ((F && i >= Mrel) || (!F && i < 0)) break;
      rel_update(M,W,i,Mrel);
    }
  i = F ? Mrel-1 : 0;
  int ms = best_state(dp,i);
  idx = RIDX(i,ms);       /* ms == NSTATE addresses the next row, as in the reference */
  for (int j = 0; j < Mrel; j++)
    if (W->rpos[j]) bt[idx][j] = CPO_R;
  return bt[idx];
}

typedef struct { char *asgn; double hdrr; } relres;

/* mean (ccb+cce)/2 coverage, length weighted, over intervals whose state passes `want`
   (want < 0: all) -- the integer accumulation of class_rel.c:634-664 */
static double mean_cov(const cpo_intvl *r, const char *asgn, int Mrel, int want)
{ int lsum = 0, csum = 0;
  for (int i = 0; i < Mrel; i++)
    if (want < 0 || asgn[i] == want)
      { int l = r[i].e-r[i].b;
        lsum += l;
        csum += (r[i].ccb+r[i].cce)*l/2;
      }
  return (double)csum/lsum;
}

/* class_rel.c:623-845 (forward and backward drivers share everything but two indices) */
static relres rel_direction(const cpo_model *M, cpo_work *W, int Mrel, int plen, int F)
{ const cpo_intvl *r = W->rintvl;
  const uint16_t *G = M->cov;
  W->forward = F;
  for (int s = 0; s < 4; s++) W->COV[s] = G[s];
  char *asgn = rel_pass(M,W,Mrel,plen);
  int no_h = 1;
  for (int i = 0; i < Mrel; i++) if (asgn[i] == CPO_H) no_h = 0;
  if (no_h)
    { int anchor = -1;          /* first D (forward) / last D (backward) */
      for (int i = 0; i < Mrel; i++)
        if (asgn[i] == CPO_D) { if (F) { if (anchor == -1) anchor = i; } else anchor = i; }
      if (anchor >= 0)
        { double mean_d = mean_cov(r,asgn,Mrel,CPO_D);
          if (mean_d < G[CPO_D])
            { W->COV[CPO_H] = F ? r[anchor].ccb : r[anchor].cce;
              W->COV[CPO_D] = (uint16_t)(W->COV[CPO_H]+G[CPO_H]);
              asgn = rel_pass(M,W,Mrel,plen);
              no_h = 1;
              for (int i = 0; i < Mrel; i++) if (asgn[i] == CPO_H) no_h = 0;
              if (no_h)
                { mean_d = mean_cov(r,asgn,Mrel,CPO_D);
                  if (fabs(mean_d-G[CPO_H]) <= fabs(mean_d-G[CPO_D]))
                    for (int i = 0; i < Mrel; i++) if (asgn[i] == CPO_D) asgn[i] = CPO_H;
                }
            }
        }
    }
  int all_h = 1;
  for (int i = 0; i < Mrel; i++) if (asgn[i] != CPO_H) all_h = 0;
  if (all_h)
    { double mean_h = mean_cov(r,asgn,Mrel,-1);
      if (fabs(mean_h-G[CPO_H]) >= fabs(mean_h-G[CPO_D]))
        for (int i = 0; i < Mrel; i++) asgn[i] = CPO_D;
    }
  int n = 0;
  for (int i = 0; i < Mrel; i++) if (asgn[i] == CPO_H) n++;
  if (n >= Mrel*0.7)
    { double mean_h = mean_cov(r,asgn,Mrel,CPO_H);
      if (fabs(mean_h-G[CPO_H]) >= fabs(mean_h-G[CPO_D]))
        for (int i = 0; i < Mrel; i++)
          { if (asgn[i] == CPO_H) asgn[i] = CPO_D;
            else if (asgn[i] == CPO_D) asgn[i] = CPO_R;
          }
    }
  int fd = -1, ld = -1, fh = -1, lh = -1;
  for (int i = 0; i < Mrel; i++)
    { if (asgn[i] == CPO_D) { if (fd == -1) fd = i; ld = i; }
      else if (asgn[i] == CPO_H) { if (fh == -1) fh = i; lh = i; }
    }
  relres res;
  res.asgn = asgn;
  res.hdrr = (fd >= 0 && fh >= 0)
               ? ((double)r[fd].ccb/r[fh].ccb)/((double)r[ld].cce/r[lh].cce) : 1.;
  return res;
}

/* class_rel.c:847-869: state codes are tested as booleans (E = 0 is "false", and the first test
 * compares against `true` = 1 = REPEAT) */
static int eq_prefix(const cpo_intvl *r, int Mrel)
{ if (r[0].asgn != 1) return 0;
  int i = 0;
  while (i < Mrel && r[i].asgn) i++;
  for (; i < Mrel; i++) if (r[i].asgn) return 0;
  return 1;
}
static int eq_suffix(const cpo_intvl *r, int Mrel)
{ if (r[Mrel-1].asgn != 1) return 0;
  int i = Mrel-2;
  while (i >= 0 && r[i].asgn) i--;
  for (; i >= 0; i--) if (r[i].asgn) return 0;
  return 1;
}

/* class_rel.c:871-963 */
static void classify_reliable(const cpo_model *M, cpo_work *W, int Mrel, int N, int plen)
{ if (Mrel == 0) return;
  cpo_intvl *r = W->rintvl;
  relres f = rel_direction(M,W,Mrel,plen,1);
  for (int i = 0; i < Mrel; i++) r[i].asgn = f.asgn[i];
  relres b = rel_direction(M,W,Mrel,plen,0);
  int eq = 1;
  for (int i = 0; i < Mrel; i++) if (r[i].asgn != b.asgn[i]) { eq = 0; break; }
  if (!eq)
    { if (eq_prefix(r,Mrel)) { }
      else if (eq_suffix(r,Mrel))
        { for (int i = 0; i < Mrel; i++) r[i].asgn = b.asgn[i]; }
      else
        { trace_cmp("fw/bw |hdrr-1|",0,0,1,fabs(f.hdrr-1.),fabs(b.hdrr-1.));
          if (!(fabs(f.hdrr-1.) <= fabs(b.hdrr-1.)))
            for (int i = 0; i < Mrel; i++) r[i].asgn = b.asgn[i];
        }
    }
  for (int ri = 0, ii = 0; ri < Mrel; ri++, ii++)
    { while (ii < N && !W->intvl[ii].is_rel) ii++;
      if (ii >= N || r[ri].b != W->intvl[ii].b || r[ri].e != W->intvl[ii].e)
        { fprintf(stderr,"Inconsistent reliable interval\n"); exit(1); }
      W->intvl[ii].asgn = r[ri].asgn;
    }
}

/*********************************************************************************************
 *  Unreliable intervals: class_unrel.c:11-300
 *********************************************************************************************/
/* class_unrel.c:11-25 */
static void nn_fixed(int idx, int s, const cpo_intvl *v, int N, int ret[2])
{ int l = idx-1;
  while (l >= 0 && !(v[l].asgn == s && v[l].is_rel)) l--;
  ret[0] = (l < 0) ? -1 : l;
  int r = idx+1;
  while (r < N && !(v[r].asgn == s && v[r].is_rel)) r++;
  ret[1] = (r >= N) ? -1 : r;
}

/* class_unrel.c:27-51 */
static uint16_t est_cov(const cpo_model *M, int x, int idx, const cpo_intvl *v, int N, int s, int from_est)
{ int nn[2];
  nn_fixed(idx,s,v,N,nn);
  int l = nn[0], r = nn[1];
  if (l != -1 && r != -1) return (uint16_t)lin_interp(x,v[l].e-1,v[l].cce,v[r].b,v[r].ccb);
  if (l != -1) return v[l].cce;
  if (r != -1) return v[r].ccb;
  if (from_est) return 0;
  uint16_t c = est_cov(M,x,idx,v,N,(s == CPO_H) ? CPO_D : CPO_H,1);
  if (c > 0) return (uint16_t)((s == CPO_H) ? c/2 : c*2);
  return M->cov[s];
}

/* class_unrel.c:53-65 */
static double un_lp_e(const cpo_model *M, const cpo_intvl *I)
{ double po = lp_poisson(M,I->cb,M->cov[CPO_E])+lp_poisson(M,I->ce,M->cov[CPO_E])+E_PO_BASE;
  return (I->pe > po) ? I->pe : po;
}

/* class_unrel.c:67-113 */
static double un_lp_r(const cpo_model *M, int idx, const cpo_intvl *v, int N)
{ const cpo_intvl *I = &v[idx];
  if (MAXI(I->cb,I->ce) >= M->cov[CPO_R]) return 0.;
  int nn[2];
  nn_fixed(idx,CPO_D,v,N,nn);
  int l = nn[0], r = nn[1];
  uint16_t dl, dr;
  if (l == -1 && r == -1) dl = dr = M->cov[CPO_D];
  else if (l == -1) dl = dr = v[r].cb;
  else if (r == -1) dl = dr = v[l].ce;
  else { dl = v[l].ce; dr = v[r].cb; }
  uint16_t rl = (uint16_t)(M->dr_ratio*dl), rr = (uint16_t)(M->dr_ratio*dr);
  if (I->cb >= rl || I->ce >= rr) return R_LOGP;
  double a = lp_binom(M,I->cb,rl,1-PE_MEAN);
  double b = lp_binom(M,I->ce,rr,1-PE_MEAN);
  return a+b;
}

static inline double max3(double a, double b, double c)
{ double m = (a > b) ? a : b; return (m > c) ? m : c; }

/* class_unrel.c:115-175 */
static double un_lp_hd(const cpo_model *M, int s, int idx, const cpo_intvl *v, int N)
{ const cpo_intvl *I = &v[idx];
  int nn[2];
  nn_fixed(idx,s,v,N,nn);
  int lrel = nn[0], rrel = nn[1];
  double lpl, lpr;
  { double er = -INFINITY, sf = -INFINITY, sfer = -INFINITY;
    int l = idx-1;
    if (l >= 0 && v[l].asgn == s) er = I->pe_o_b;
    if (lrel != -1) sf = lp_trans(M,v[lrel].e-1,I->b,v[lrel].cce,I->cb,v[lrel].cce);
    uint16_t est = est_cov(M,I->b,idx,v,N,s,0);
    if (est >= I->cb) sfer = log(p_errorin(M,CPO_OTHERS,0.1,est,I->cb));
    lpl = max3(er,sf,sfer);
  }
  { double er = -INFINITY, sf = -INFINITY, sfer = -INFINITY;
    int r = idx+1;
    if (r < N && v[r].asgn == s) er = I->pe_o_e;
    if (rrel != -1) sf = lp_trans(M,I->e-1,v[rrel].b,I->ce,v[rrel].ccb,v[rrel].ccb);
    uint16_t est = est_cov(M,I->e-1,idx,v,N,s,0);
    if (est >= I->ce) sfer = log(p_errorin(M,CPO_OTHERS,0.1,est,I->ce));
    lpr = max3(er,sf,sfer);
  }
  if (lpl == -INFINITY && lpr == -INFINITY)
    { lpl = lp_poisson(M,I->cb,M->cov[s]); lpr = lp_poisson(M,I->ce,M->cov[s]); }
  else if (lpl == -INFINITY) lpl = lpr;
  else if (lpr == -INFINITY) lpr = lpl;
  return lpl+lpr;
}

/* class_unrel.c:192-237 */
static void un_update(const cpo_model *M, int idx, cpo_intvl *v, int N)
{ const cpo_intvl I = v[idx];
  if (MAXI(I.cb,I.ce) >= M->cov[CPO_R]) { v[idx].asgn = CPO_R; return; }
  double mx = -INFINITY; int ms = -1;
  for (int s = CPO_E; s <= CPO_D; s++)
    { double lp = (s == CPO_E) ? un_lp_e(M,&v[idx])
                : (s == CPO_R) ? un_lp_r(M,idx,v,N) : un_lp_hd(M,s,idx,v,N);
      if (mx < lp) { mx = lp; ms = s; }
    }
  if (ms == -1) { fprintf(stderr,"No valid probability for interval %d\n",idx); exit(1); }
  if (I.asgn != ms) v[idx].asgn = (int8_t)ms;
}

/* class_unrel.c:248-275 */
static void classify_unreliable(const cpo_model *M, cpo_intvl *v, int N)
{ uint8_t *fixed = malloc((size_t)N+1);
  int *ord = malloc(sizeof(int)*((size_t)N+1));
  for (int i = 0; i < N; i++)
    { fixed[i] = (v[i].is_rel && (v[i].asgn == CPO_H || v[i].asgn == CPO_D));
      ord[i] = i;
    }
  /* stable sort by min(cb,ce) ascending (glibc qsort = merge sort) */
  for (int i = 1; i < N; i++)
    { int x = ord[i], kx = MINI(v[x].cb,v[x].ce), j = i-1;
      while (j >= 0 && MINI(v[ord[j]].cb,v[ord[j]].ce) > kx) { ord[j+1] = ord[j]; j--; }
      ord[j+1] = x;
    }
  for (int i = N-1; i >= 0; i--) if (!fixed[ord[i]]) un_update(M,ord[i],v,N);
  for (int i = 0; i < N; i++)    if (!fixed[ord[i]]) un_update(M,ord[i],v,N);
  free(fixed); free(ord);
}

/*********************************************************************************************
 *  One read: ClassPro.c:229-271
 *********************************************************************************************/
int cpo_classify_read(const cpo_model *M, cpo_work *W, const char *seq, int rlen,
                      const uint16_t *prof, int plen, char *cls)
{ const int K = M->K;
  if (rlen > CPO_MAX_RLEN || rlen != plen+K-1 || plen < 1) return -1;
  if (W->clean)
    { /* device definition of the reference's stale reads: right-context cells that
         the sweep never writes are 0, and profile[plen] (read by wall.c:977-978 when a
         low-complexity run reaches the end of the read) equals profile[plen-1] */
      memset(W->rctx,0,(size_t)(rlen+8)*3);
      if (prof != W->profile) memcpy(W->profile,prof,sizeof(uint16_t)*(size_t)plen);
      W->profile[plen] = W->profile[plen-1];
      prof = W->profile;
    }
  cpo_seq_context(W->lctx_base,W->rctx,seq,rlen);
  int N = find_walls(M,W,prof,plen);
  int Mrel = find_reliable(M,W,N,prof);
  if (Mrel > MCAP-2) { fprintf(stderr,"oracle: too many reliable intervals\n"); return -2; }
  W->N = N; W->M = Mrel;
  classify_reliable(M,W,Mrel,N,plen);
  classify_unreliable(M,W->intvl,N);
  for (int i = 0; i < K-1; i++) cls[i] = 'N';
  for (int i = 0; i < N; i++)
    { char c = STOC[(int)W->intvl[i].asgn];
      for (int j = W->intvl[i].b; j < W->intvl[i].e; j++) cls[K-1+j] = c;
    }
  cls[rlen] = '\0';
  return N;
}

/*********************************************************************************************
 *  File driver (FASTX + FastK files -> .class), semantics of kseq.h:177-218, ClassPro.c:181-289
 *********************************************************************************************/
typedef struct { char *s; size_t l, m; } str_t;
static void str_put(str_t *b, const char *p, size_t n)
{ if (b->l+n+1 > b->m) { b->m = (b->l+n+1)*2; b->s = realloc(b->s,b->m); }
  memcpy(b->s+b->l,p,n); b->l += n; b->s[b->l] = 0;
}

static int read_all(const char *path, char **buf, size_t *len)
{ gzFile f = gzopen(path,"r");
  if (f == NULL) return 1;
  size_t cap = 1<<24, n = 0;
  char *b = malloc(cap);
  for (;;)
    { if (cap-n < (1<<20)) { cap *= 2; b = realloc(b,cap); }
      int r = gzread(f,b+n,(unsigned)MINI((size_t)(1<<30),cap-n));
      if (r <= 0) break;
      n += (size_t)r;
    }
  gzclose(f);
  *buf = b; *len = n;
  return 0;
}

typedef struct
  { int      kmer, nparts;
    int64_t  nreads;
    int64_t *index;      /* [nreads+1] end offsets inside their part */
    int64_t *nbase;      /* [nparts] cumulative reads */
    char     prefix[4096];
  } profidx;

static int profidx_open(profidx *P, const char *fk_root)
{ char path[4200];
  const char *sl = strrchr(fk_root,'/');
  char dir[4096], root[1024];
  if (sl) { snprintf(dir,sizeof(dir),"%.*s",(int)(sl-fk_root),fk_root); snprintf(root,sizeof(root),"%s",sl+1); }
  else    { snprintf(dir,sizeof(dir),"."); snprintf(root,sizeof(root),"%s",fk_root); }
  size_t rl = strlen(root);
  if (rl > 5 && strcasecmp(root+rl-5,".prof") == 0) root[rl-5] = 0;
  snprintf(path,sizeof(path),"%s/%s.prof",dir,root);
  FILE *f = fopen(path,"rb");
  if (f == NULL) return 1;
  int32_t smer, nthreads;
  if (fread(&smer,4,1,f) != 1 || fread(&nthreads,4,1,f) != 1) { fclose(f); return 1; }
  fclose(f);
  snprintf(P->prefix,sizeof(P->prefix),"%s/.%s.",dir,root);
  P->kmer = smer; P->nparts = nthreads;
  P->nbase = calloc((size_t)nthreads,sizeof(int64_t));
  int64_t total = 0;
  for (int p = 0; p < nthreads; p++)
    { snprintf(path,sizeof(path),"%spidx.%d",P->prefix,p+1);
      f = fopen(path,"rb");
      if (f == NULL) { fprintf(stderr,"Profile part %s is misssing ?\n",path); return 2; }
      int32_t k; int64_t n;
      if (fread(&k,4,1,f) != 1 || fread(&n,8,1,f) != 1 || fread(&n,8,1,f) != 1) { fclose(f); return 2; }
      fclose(f);
      total += n;
    }
  P->index = calloc((size_t)total+1,sizeof(int64_t));
  int64_t nr = 0;
  for (int p = 0; p < nthreads; p++)
    { snprintf(path,sizeof(path),"%spidx.%d",P->prefix,p+1);
      f = fopen(path,"rb");
      int32_t k; int64_t n;
      if (fread(&k,4,1,f) != 1 || fread(&n,8,1,f) != 1 || fread(&n,8,1,f) != 1) { fclose(f); return 2; }
      if (fread(P->index+nr+1,8,(size_t)n,f) != (size_t)n) { fclose(f); return 2; }
      fclose(f);
      nr += n;
      P->nbase[p] = nr;
    }
  P->nreads = nr;
  return 0;
}

int cpo_model_load(cpo_model *M, const char *fk_root, int cov_opt, int read_len, int verbose)
{ char path[4200];
  snprintf(path,sizeof(path),"%s.hist",fk_root);
  FILE *f = fopen(path,"rb");
  if (f == NULL) { fprintf(stderr,"Cannot open %s\n",path); return 1; }
  int32_t kmer, low, high; int64_t il, ih;
  if (fread(&kmer,4,1,f) != 1 || fread(&low,4,1,f) != 1 || fread(&high,4,1,f) != 1
      || fread(&il,8,1,f) != 1 || fread(&ih,8,1,f) != 1) { fclose(f); return 1; }
  int64_t *h = malloc(sizeof(int64_t)*(size_t)(high-low+1));
  if (fread(h,8,(size_t)(high-low+1),f) != (size_t)(high-low+1)) { fclose(f); free(h); return 1; }
  fclose(f);
  int rc = cpo_model_from_hist(M,kmer,low,high,il,ih,h,cov_opt,read_len,verbose);
  free(h);
  return rc;
}

int cpo_run_file(const char *fastx, const char *fk_root, int cov_opt, int read_len,
                 const char *out_path, int verbose, int64_t *nkmers)
{ profidx P; memset(&P,0,sizeof(P));
  if (profidx_open(&P,fk_root)) { fprintf(stderr,"oracle: cannot open %s.prof\n",fk_root); return 1; }
  cpo_model *M = malloc(sizeof(cpo_model));
  if (cpo_model_load(M,fk_root,cov_opt,read_len,verbose)) return 2;
  M->K = P.kmer;
  const int K = P.kmer;
  char *buf; size_t len;
  if (read_all(fastx,&buf,&len)) { fprintf(stderr,"oracle: cannot open %s\n",fastx); return 3; }
  FILE *out = fopen(out_path,"wb");
  if (out == NULL) return 4;
  cpo_work *W = cpo_work_new();
  for (int i = 0; i < K-1; i++) W->rasgn[i] = 'N';

  str_t name = {0}, comment = {0}, seq = {0};
  int have_comment = 0;            /* comment.s is NULL until the first comment (prints "(null)") */
  size_t p = 0;
  int last_char = 0;
  int64_t id = 0, tot = 0;
  int part = -1; FILE *pf = NULL;
  uint8_t *cbuf = malloc(4*CPO_MAX_RLEN+16);
  while (1)
    { /* kseq_read */
      if (last_char == 0)
        { while (p < len && buf[p] != '>' && buf[p] != '@') p++;
          if (p >= len) break;
          last_char = buf[p++];
        }
      comment.l = 0; seq.l = 0; name.l = 0;
      size_t q = p;
      while (q < len && !isspace((unsigned char)buf[q])) q++;
      if (q >= len && q == p) break;
      str_put(&name,buf+p,q-p);
      int c = (q < len) ? buf[q] : 0;
      p = (q < len) ? q+1 : q;
      if (c != '\n' && q < len)
        { q = p;
          while (q < len && buf[q] != '\n') q++;
          comment.l = 0;
          str_put(&comment,buf+p,q-p);
          if (comment.l > 1 && comment.s[comment.l-1] == '\r') comment.s[--comment.l] = 0;
          have_comment = 1;
          p = (q < len) ? q+1 : q;
        }
      int cc = -1;
      if (seq.s == NULL) str_put(&seq,"",0);
      while (p < len)
        { cc = buf[p++];
          if (cc == '>' || cc == '+' || cc == '@') break;
          if (cc == '\n') { cc = -1; continue; }
          char ch = (char)cc;
          str_put(&seq,&ch,1);
          q = p;
          while (q < len && buf[q] != '\n') q++;
          str_put(&seq,buf+p,q-p);
          if (seq.l > 1 && seq.s[seq.l-1] == '\r') seq.s[--seq.l] = 0;
          p = (q < len) ? q+1 : q;
          cc = -1;
        }
      if (cc == '>' || cc == '@') last_char = cc;
      else if (cc == '+')
        { while (p < len && buf[p] != '\n') p++;
          if (p < len) p++;
          size_t ql = 0;
          while (p < len && ql < seq.l)
            { q = p;
              while (q < len && buf[q] != '\n') q++;
              size_t ll = q-p;
              if (ll > 0 && buf[q-1] == '\r' && ql+ll > 1) ll--;
              ql += ll;
              p = (q < len) ? q+1 : q;
            }
          last_char = 0;
        }
      else last_char = 0, p = len;

      if (id >= P.nreads) break;
      int rlen = (int)seq.l;
      if (rlen > CPO_MAX_RLEN)
        { fprintf(stderr,"rlen (%d) > MAX_READ_LEN for FASTX inputs (%d)\n",rlen,CPO_MAX_RLEN); return 5; }
      /* header: "@%s %s" (ClassPro.c:188) */
      fprintf(out,"@%s %s\n",name.s ? name.s : "",have_comment ? comment.s : "(null)");
      if (rlen <= K-1)
        { /* ClassPro.c:209-226: "%*s" prints the whole stale class string, padded to rlen */
          fprintf(out,"%s\n+\n%*s\n",seq.s,rlen,W->rasgn);
          id++;
          continue;
        }
      /* profile fetch: libfastk.c:1414-1462 */
      int w = 0;
      while (w < P.nparts && id >= P.nbase[w]) w++;
      if (w != part)
        { if (pf) fclose(pf);
          char path[4200];
          snprintf(path,sizeof(path),"%sprof.%d",P.prefix,w+1);
          pf = fopen(path,"rb");
          if (pf == NULL) { fprintf(stderr,"Profile part %s is misssing ?\n",path); return 6; }
          part = w;
        }
      int64_t off = (id == 0 || (w > 0 && id == P.nbase[w-1])) ? 0 : P.index[id];
      int64_t clen = P.index[id+1]-off;
      if (clen > 4*CPO_MAX_RLEN) { fprintf(stderr,"oracle: profile too long\n"); return 7; }
      fseeko(pf,off,SEEK_SET);
      if (fread(cbuf,1,(size_t)clen,pf) != (size_t)clen) return 7;
      int plen = cpo_decode_profile(cbuf,clen,W->profile,CPO_MAX_RLEN);
      if (rlen != plen+K-1)
        { fprintf(stderr,"Read %lld: rlen (%d) != plen+Km1 (%d)\n",(long long)id+1,rlen,plen+K-1); return 8; }
      if (cpo_classify_read(M,W,seq.s,rlen,W->profile,plen,W->rasgn) < 0) return 9;
      fprintf(out,"%s\n+\n%s\n",seq.s,W->rasgn);
      tot += plen;
      id++;
    }
  fclose(out);
  if (pf) fclose(pf);
  if (nkmers) *nkmers = tot;
  cpo_work_free(W);
  free(buf); free(cbuf); free(name.s); free(comment.s); free(seq.s);
  free(P.index); free(P.nbase); free(M);
  return 0;
}

#ifdef CPO_MAIN
/* cpo_classify [-v] [-c<int>] [-r<int>] [-N<fk_root>] <reads.fasta> <out.class> */
int main(int argc, char **argv)
{ int verbose = 0, cov = 0, rl = 20000;
  const char *fk = NULL, *pos[2]; int np = 0;
  for (int i = 1; i < argc; i++)
    { if (argv[i][0] == '-')
        { switch (argv[i][1])
            { case 'v': verbose = 1; break;
              case 'c': cov = atoi(argv[i]+2); break;
              case 'r': rl = atoi(argv[i]+2); break;
              case 'N': fk = argv[i]+2; break;
              default: fprintf(stderr,"unknown option %s\n",argv[i]); return 1;
            }
        }
      else if (np < 2) pos[np++] = argv[i];
    }
  if (np != 2) { fprintf(stderr,"usage: cpo_classify [-v] [-c<int>] [-r<int>] [-N<fk_root>] <reads> <out.class>\n"); return 1; }
  char root[4096];
  if (fk == NULL)
    { snprintf(root,sizeof(root),"%s",pos[0]);
      static const char *ext[] = { ".fastq.gz",".fasta.gz",".fq.gz",".fa.gz",".fastq",".fasta",".fq",".fa" };
      size_t l = strlen(root);
      for (int e = 0; e < 8; e++)
        { size_t el = strlen(ext[e]);
          if (l > el && strcmp(root+l-el,ext[e]) == 0) { root[l-el] = 0; break; }
        }
      fk = root;
    }
  int64_t nk = 0;
  int rc = cpo_run_file(pos[0],fk,cov,rl,pos[1],verbose,&nk);
  if (verbose) fprintf(stderr,"oracle: classified %lld k-mers (rc=%d)\n",(long long)nk,rc);
  return rc;
}
#endif
