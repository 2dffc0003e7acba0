/*******************************************************************************************
 *  cpg_math.cuh -- FP64 primitives of the classification path, device side.
 *
 *  Replaces src/bessel.c:390-521 (bessi0/bessi1/bessi), src/prob.c:22-112 and src/util.c:9-55.
 *  Every expression keeps the reference's operation order; the translation unit is compiled with
 *  -fmad=false so that no multiply-add is contracted (gcc emits none for baseline x86-64), and
 *  +,-,*,/,sqrt are IEEE-754 correctly rounded on both sides.  exp/log come from the CUDA math
 *  library (<= 1 ulp) where the reference uses glibc; DESIGN.md "Floating-point parity" covers the
 *  consequences.
 *******************************************************************************************/
#ifndef CPG_MATH_CUH
#define CPG_MATH_CUH
#include "cpg_common.h"

/* Context of the lane group that owns a read: the same values in every lane of the group except
 * `lane` / `glane`.  The group is a whole warp or an aligned power-of-two part of one; the
 * reliable-interval DP splits it once more (forward and backward pass on the two halves). */
struct WCtx
  { int               lane;      /* lane in the warp */
    const cpg_dmodel *M;
    const uint8_t    *cthres;    /* shared-memory copy of M->cthres (or M->cthres itself) */
    cpg_wshared      *ws;
    int               status;
    int               glane, gsize, gbase;   /* lane in the group, lanes in the group, warp lane of group lane 0 */
    unsigned          gmask;                 /* warp-level mask of the group's lanes */
  };

/* one copy of each in the kernel: the per-read code is executed by up to 32 warps per SM that sit in
   different phases, so instruction-cache footprint matters more than call overhead */
CPG_DEV_MATHFN double cpg_exp(double x) { return exp(x); }
CPG_DEV_MATHFN double cpg_log(double x) { return log(x); }

CPG_DEV int imin(int a, int b) { return a < b ? a : b; }
CPG_DEV int imax(int a, int b) { return a > b ? a : b; }
CPG_DEV int iabs(int a) { return a < 0 ? -a : a; }
/* the reference's MAX macro on doubles: (x) > (y) ? (x) : (y) */
CPG_DEV double dmax_ref(double x, double y) { return x > y ? x : y; }

/* src/bessel.c:390-411 */
CPG_DEV_HELPER double cpg_bessi0(double x)
{ double ax = fabs(x), y, ans;
  if (ax < 3.75)
    { y = x/3.75; y = y*y;
      ans = 1.0+y*(3.5156229+y*(3.0899424+y*(1.2067492+y*(0.2659732+y*(0.360768e-1+y*0.45813e-2)))));
    }
  else
    { y = 3.75/ax;
      ans = (cpg_exp(ax)/sqrt(ax))*(0.39894228+y*(0.1328592e-1+y*(0.225319e-2+y*(-0.157565e-2
            +y*(0.916281e-2+y*(-0.2057706e-1+y*(0.2635537e-1+y*(-0.1647633e-1+y*0.392377e-2))))))));
    }
  return ans;
}

/* src/bessel.c:416-439 */
CPG_DEV_HELPER double cpg_bessi1(double x)
{ double ax = fabs(x), y, ans;
  if (ax < 3.75)
    { y = x/3.75; y = y*y;
      ans = ax*(0.5+y*(0.87890594+y*(0.51498869+y*(0.15084934+y*(0.2658733e-1+y*(0.301532e-2
            +y*0.32411e-3))))));
    }
  else
    { y = 3.75/ax;
      ans = 0.2282967e-1+y*(-0.2895312e-1+y*(0.1787654e-1-y*0.420059e-2));
      ans = 0.39894228+y*(-0.3988024e-1+y*(-0.362018e-2+y*(0.163801e-2+y*(-0.1031555e-1+y*ans))));
      ans *= (cpg_exp(ax)/sqrt(ax));
    }
  return x < 0.0 ? -ans : ans;
}

/* keeps the compiler from turning the (rare, three-multiply) rescale into always-executed
 * predicated code inside the recurrence */
#ifdef CPG_HOSTSIM
#define CPG_NO_IFCVT() do { } while (0)
#else
#define CPG_NO_IFCVT() asm volatile("")
#endif

/* src/bessel.c:482-521: Miller downward recurrence started at j = 2*(n+floor(sqrt(40 n))),
 * rescaled by 1e-10 whenever |bi| passes 1e10, normalised with I0.  The loop is split at j == n:
 * before that point `ans` is still 0 (its rescale multiplies are no-ops and are skipped), at
 * j == n it takes the value of bip, afterwards it is rescaled with the others.  j is kept as a
 * double (exact) so no integer->double conversion sits in the dependent chain, and the steps are
 * written in pairs so that bi/bip swap roles instead of being copied.
 * Same operations on the same values, in the same order, as the reference. */
CPG_DEV_NOINL double cpg_bessi(int n, double x)
{ if (n == 0) return cpg_bessi0(x);
  if (n == 1) return cpg_bessi1(x);
  if (x == 0.0) return 0.0;
  const double tox = 2.0/fabs(x), big = 1.0e10, small = 1.0e-10;
  const int start = 2*(n+(int)sqrt(40.0*n));
  double jd = (double)start;
  /* One step of the recurrence without register shuffling: X holds bip and becomes the new bi,
     Y holds bi and becomes the new bip.  Two steps bring the roles back. */
#define CPG_BSTEP(X,Y,RESCALE_ANS) \
    { X = X+jd*tox*Y; \
      if (fabs(X) > big) { CPG_NO_IFCVT(); RESCALE_ANS X *= small; Y *= small; } \
      jd -= 1.0; }
  double P = 0.0, Q = 1.0, ans;                     /* bip = P, bi = Q */
  int c = start-n+1;                                /* steps j = start .. n: ans is still 0 */
#ifndef CPG_HOSTSIM
#pragma unroll 1
#endif
  for (; c >= 2; c -= 2) { CPG_BSTEP(P,Q,) CPG_BSTEP(Q,P,) }
  if (c) { CPG_BSTEP(P,Q,) double t = P; P = Q; Q = t; }
  ans = P;                                          /* if (j == n) ans = bip */
  c = n-1;                                          /* steps j = n-1 .. 1 */
#ifndef CPG_HOSTSIM
#pragma unroll 1
#endif
  for (; c >= 2; c -= 2) { CPG_BSTEP(P,Q,ans *= small;) CPG_BSTEP(Q,P,ans *= small;) }
  if (c) { CPG_BSTEP(P,Q,ans *= small;) Q = P; }
#undef CPG_BSTEP
  ans *= cpg_bessi0(x)/Q;
  return (x < 0.0 && (n%2) == 1) ? -ans : ans;
}

/* src/prob.c:22-31: counts above 32767 are clamped (the reference also prints a note) */
CPG_DEV int cpg_clamp_cnt(int n) { return n > CPG_MAX_CNT ? CPG_MAX_CNT : n; }

/* An error rate with its logarithms (taken from the model's tables) */
struct cpg_rate { double p, lp, l1mp; };
CPG_DEV cpg_rate cpg_rate_pe(const cpg_dmodel *M, int t, int l) { cpg_rate r = { M->pe[t][l], M->lpe[t][l], M->l1mpe[t][l] }; return r; }
CPG_DEV cpg_rate cpg_rate_hc(const cpg_dmodel *M) { cpg_rate r = { M->hc_erate, M->l_hc, M->l1m_hc }; return r; }
CPG_DEV cpg_rate cpg_rate_p1(const cpg_dmodel *M) { cpg_rate r = { 0.1, M->l_p1, M->l1m_p1 }; return r; }

/* Logarithms of the model constants; lane/thread `tid` of `nth` fills its share (the caller
   synchronises afterwards). */
CPG_DEV void cpg_model_fill_logs(cpg_dmodel *m, int tid, int nth)
{ for (int i = tid; i < 63; i += nth)
    { const int t = i/21, l = i%21;
      const double p = m->pe[t][l];
      m->lpe[t][l] = cpg_log(p); m->l1mpe[t][l] = cpg_log(1-p);
    }
  if (tid == 0)
    { const double p = 1-CPG_PE_MEAN;
      m->l_hc = cpg_log(m->hc_erate); m->l1m_hc = cpg_log(1-m->hc_erate);
      m->l_p1 = cpg_log(0.1);         m->l1m_p1 = cpg_log(1-0.1);
      m->l_p99 = cpg_log(p);          m->l1m_p99 = cpg_log(1-p);
      for (int s = 0; s < 4; s++) m->lcov[s] = cpg_log((double)m->cov[s]);
    }
}

/* src/prob.c:33-39; lambda is almost always one of the four global coverages */
CPG_DEV_HELPER double cpg_lp_poisson(const WCtx &W, uint16_t k16, int lambda)
{ int k = cpg_clamp_cnt(k16);
  const cpg_dmodel *M = W.M;
  double ll;
  if      (lambda == M->cov[ST_E]) ll = M->lcov[ST_E];
  else if (lambda == M->cov[ST_H]) ll = M->lcov[ST_H];
  else if (lambda == M->cov[ST_D]) ll = M->lcov[ST_D];
  else if (lambda == M->cov[ST_R]) ll = M->lcov[ST_R];
  else ll = cpg_log((double)lambda);
  return k*ll-lambda-CPG_LDG(M->logfact+k);
}

/* src/prob.c:41-44 */
CPG_DEV_HELPER double cpg_lp_skellam(int k, double lambda)
{ return -2.*lambda+cpg_log(cpg_bessi(k < 0 ? -k : k,2.*lambda)); }

/* src/util.c:35-44; `cov` is a 16-bit count in the reference's signature */
CPG_DEV double cpg_lp_trans(const WCtx &W, int b, int e, int cb, int ce, uint16_t cov)
{ int d = e-b; if (d < 0) d = -d;
  return cpg_lp_skellam(ce-cb,(double)cov*d/W.M->read_len);
}

/* cpg_lp_trans where only "is it >= thres" is asked (src/wall.c:366,390 THRES_DIFF_EO; :1028 THRES_DIFF_REL).
 * The Bessel recurrence takes 2(n+sqrt(40n)) steps, n = |ce-cb|: tens of thousands where the counts are in the
 * thousands (repeat-rich profiles: one lane of a warp in such a loop, the other 31 waiting).  There the answer is
 * known without it: I_n(x) < cosh(x) (x/2)^n / n!  (x > 0) and n! >= (n/e)^n give
 *     logp_skellam(k,lambda) = -2 lambda + log I_n(2 lambda)  <  n (log(lambda/n) + 1),
 * and when that bound is below the threshold by more than 2 -- far more than any rounding of the recurrence --
 * the exact value is below it too.  Returned then: -inf (every use is the comparison).  Otherwise the exact value.
 * Only for 2 lambda < 700: beyond that exp(2 lambda) overflows inside the reference's bessi0 (src/bessel.c:401) and
 * its result is +inf, or NaN where the recurrence underflowed to 0 -- and NaN passes the reference's `< threshold`
 * tests (src/wall.c:390,1028).  That behaviour is kept by evaluating those (rare: coverage x distance > 7e6) as it does. */
CPG_DEV_HELPER double cpg_lp_trans_thr(const WCtx &W, int b, int e, int cb, int ce, uint16_t cov, double thres)
{ int d = e-b; if (d < 0) d = -d;
  const int k = ce-cb, n = k < 0 ? -k : k;
  const double lambda = (double)cov*d/W.M->read_len;
  if (n >= 32 && 2.*lambda < 700.)
    { if (!(lambda > 0.)) return -CPG_INF;
      if ((double)n*(cpg_log(lambda/(double)n)+1.) < thres-2.) return -CPG_INF;
    }
  return cpg_lp_skellam(k,lambda);
}

/* src/prob.c:59-65 with p = 1-PE_MEAN, the only value the path uses (src/class_rel.c:186,
   src/class_unrel.c:98-99) */
CPG_DEV_HELPER double cpg_lp_binom99(WCtx &W, uint16_t k16, uint16_t n16)
{ int k = cpg_clamp_cnt(k16), n = cpg_clamp_cnt(n16);
  if (k > n) { W.status |= CPG_ST_BINOM; return -CPG_INF; }
  const double *lf = W.M->logfact;
  return CPG_LDG(lf+n)-CPG_LDG(lf+k)-CPG_LDG(lf+(n-k))+k*W.M->l_p99+(n-k)*W.M->l1m_p99;
}

/* src/prob.c:76-112 with exact = false: one-sided binomial tail summed in the reference's order and
 * cut after the first term below a tenth of the first one.  Evaluated by ONE lane, term after term
 * exactly as the reference loops: the callers want several independent tails at once, so each
 * lane takes one. */
CPG_DEV_NOINL double cpg_binom_tail_lane(const double *lf, int k, int n, const cpg_rate r, int *bad)
{ k = cpg_clamp_cnt(k & 0xffff); n = cpg_clamp_cnt(n & 0xffff);
  if (k > n) { *bad = 1; return 0.; }
  const double lpe = r.lp, l1mpe = r.l1mp, mean = n*r.p;
  const double lfn = CPG_LDG(lf+n);
  double p, p_first, t;
#define CPG_LBP(x) (lfn-CPG_LDG(lf+(x))-CPG_LDG(lf+(n-(x)))+(x)*lpe+(n-(x))*l1mpe)
  /* A first term that underflows to 0 can never stop the loop (10*t < 0 is false): the reference then adds
     n-k (or k-1) further terms, each farther out in the tail than the first and so 0 as well -- up to 32 767 exps
     for a sum that stays 0 (counts in the thousands, repeat-rich profiles: one lane of a warp in that loop, 31
     waiting).  Same result without them. */
  if ((double)k >= mean)
    { p = p_first = cpg_exp(CPG_LBP(k));
      if (p_first != 0.)
        CPG_LOOP for (int x = k+1; x <= n; x++)
          { p += t = cpg_exp(CPG_LBP(x));
            if (10*t < p_first) break;
          }
    }
  else
    { p = p_first = (k == 0) ? 0. : cpg_exp(CPG_LBP(k-1));
      if (p_first != 0.)
        CPG_LOOP for (int x = k-2; x >= 0; x--)
          { p += t = cpg_exp(CPG_LBP(x));
            if (10*t < p_first) break;
          }
      p = 1-p;
    }
#undef CPG_LBP
  return p;
}

/* p_errorin (src/util.c:46-55) for one lane; the caller guarantees cin <= cout */
CPG_DEV double cpg_p_errorin_lane(const double *lf, int etype, const cpg_rate erate, int cout, int cin, int *bad)
{ return cpg_binom_tail_lane(lf,(etype == ET_SELF) ? cin : cout-cin,cout,erate,bad); }

/* src/util.c:24-33 */
CPG_DEV double cpg_lin_interp(WCtx &W, int x, int p1, uint16_t c1, int p2, uint16_t c2)
{ if (!(p1 < x && x < p2)) W.status |= CPG_ST_INTERP;
  return (double)c1+((double)c2-c1)*(x-p1)/(p2-p1);
}

#endif
