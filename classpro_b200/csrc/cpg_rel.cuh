/*******************************************************************************************
 *  cpg_rel.cuh -- classification of the reliable intervals of one read (forward + backward
 *  Viterbi-like passes and their reconciliation), one warp.
 *
 *  Replaces classify_rel and everything under it, src/class_rel.c:41-963.
 *
 *  Restructured for the GPU (results identical):
 *   - the reference copies the whole state path of the chosen predecessor into each of the four
 *     states at every interval (O(M^2) bytes per pass, src/class_rel.c:401-407) and re-scans it in
 *     calc_dh_ratio / has_h / has_d.  Those scans only ever need, per path, the last H interval,
 *     the last D interval, the last H before the last D and the last D before the last H; this
 *     summary is carried per state, and the path itself is recovered once at the end from 3-bit
 *     back pointers.  Only two DP columns are live; they sit in the warp's shared-memory block.
 *   - the 16 transition log-probabilities of a step (each at most one Bessel recurrence) are
 *     evaluated by 16 lanes at once, and the 16 logs of the normalisation likewise; sums keep the
 *     reference's order.
 *******************************************************************************************/
#ifndef CPG_REL_CUH
#define CPG_REL_CUH
#include "cpg_wall.cuh"

struct RelState
  { double   dp, dhr;
    int      pos[4];          /* st[.][R,H,D].pos  (index 0 unused) */
    uint16_t cnt[4];          /* st[.][R,H,D].cnt */
    int      lastH, lastD;    /* path summary, interval indices or -1 */
    int      hbd, dbh;        /* last H before lastD / last D before lastH */
  };

struct RelShared
  { RelState col[2][4];
    double   tr[16];
    int      bpb[4];          /* back pointer of each target state, gathered by lane 0 */
  };

struct RelRun
  { int       F;              /* forward? */
    uint16_t  COV[4];
    int       M, plen;
    RelShared *sh;
    cpg_intvl *wint;          /* this direction's working copy, back pointers and freeze flags */
    uint16_t  *bp;
    uint8_t   *rpos;
  };

CPG_DEV int rl_pred(int x, int F) { return F ? x-1 : x+1; }
CPG_DEV int rl_off(int x, int F)  { return F ? x-CPG_OFFSET : x+CPG_OFFSET; }
CPG_DEV int rl_begpos(const cpg_intvl &I, int F) { return F ? I.b : I.e-1; }
CPG_DEV int rl_endpos(const cpg_intvl &I, int F) { return F ? I.e-1 : I.b; }
CPG_DEV uint16_t rl_begcnt(const cpg_intvl &I, int F) { return F ? I.ccb : I.cce; }
CPG_DEV uint16_t rl_endcnt(const cpg_intvl &I, int F) { return F ? I.cce : I.ccb; }

/* src/class_rel.c:158-170 */
CPG_DEV_HELPER double rl_lp_e(const WCtx &W, const cpg_intvl &I, const uint16_t *COV)
{ double po = cpg_lp_poisson(W,I.ccb,COV[ST_E])+cpg_lp_poisson(W,I.cce,COV[ST_E])+CPG_E_PO_BASE;
  return dmax_ref(po,I.pe);
}

/* src/class_rel.c:172-211 */
CPG_DEV_HELPER double rl_lp_r(WCtx &W, const cpg_intvl &I, uint16_t pr_cnt, int F, const uint16_t *COV)
{ uint16_t bc = rl_begcnt(I,F);
  double sf = -CPG_INF;
  double er = (bc < pr_cnt) ? cpg_lp_binom99(W,bc,pr_cnt) : -CPG_INF;
  double lp = dmax_ref(sf,er);
  if (lp > CPG_R_LOGP) return lp;
  uint16_t mx = (uint16_t)imax(I.ccb,I.cce);
  if (mx >= COV[ST_R]) return CPG_R_LOGP;
  if (mx >= pr_cnt) return CPG_R_LOGP;
  return lp;
}

/* src/class_rel.c:213-270, argument side.  H: the H-track transition is replaced by the D-track
 * one scaled by the predecessor's D/H ratio whenever that ratio exists.  D: always the plain
 * D-track transition (the ratio-scaled H-track value is computed and dropped by the reference).
 * Only the Skellam arguments are formed here: the evaluation itself (a Bessel recurrence) happens
 * at ONE call site for all lanes, so lanes on different branches do not serialise it. */
CPG_DEV void rl_hd_args(const WCtx &W, int t, const cpg_intvl &I, const RelState &P, int F,
                        int &k, double &lambda)
{ int bp = rl_begpos(I,F); uint16_t bc = rl_begcnt(I,F);
  int b, cb, ce; uint16_t cov;
  if (t == ST_H && P.dhr != -CPG_INF)
    { b = rl_pred(P.pos[ST_D],F); cb = P.cnt[ST_D]; ce = (int)(P.dhr*bc); cov = P.cnt[ST_D]; }
  else if (t == ST_H)
    { b = rl_pred(P.pos[ST_H],F); cb = P.cnt[ST_H]; ce = bc; cov = P.cnt[ST_H]; }
  else
    { b = rl_pred(P.pos[ST_D],F); cb = P.cnt[ST_D]; ce = bc; cov = P.cnt[ST_D]; }
  int d = bp-b; if (d < 0) d = -d;
  k = ce-cb;
  lambda = (double)cov*d/W.M->read_len;          /* src/util.c:43 */
}

/* src/class_rel.c:80-96 with s (or t) as the wildcard */
CPG_DEV int rl_best_from(const RelState *prv, const double *tr, int t, double *out)   /* wildcard s */
{ double mx = -CPG_INF; int ms = ST_N;
  CPG_LOOP for (int x = 0; x < 4; x++)
    { double lp = prv[x].dp+tr[x*4+t];
      if (mx < lp) { mx = lp; ms = x; }
    }
  if (out) *out = mx;
  return ms;
}
CPG_DEV int rl_best_to(const RelState *prv, const double *tr, int s)                  /* wildcard t */
{ double mx = -CPG_INF; int mt = ST_N;
  CPG_LOOP for (int x = 0; x < 4; x++)
    { double lp = prv[s].dp+tr[s*4+x];
      if (mx < lp) { mx = lp; mt = x; }
    }
  return mt;
}

/* src/class_rel.c:113-156 on the path summary of predecessor P extended by state t at interval i */
CPG_DEV_HELPER double rl_dh_ratio(WCtx &W, const cpg_intvl *v, int t, int i, const RelState &P, int F)
{ int i2 = (t == ST_H) ? P.lastD : P.lastH;
  if (i2 < 0) return -CPG_INF;
  int i3 = (t == ST_H) ? P.hbd : P.dbh;
  if (i3 < 0) return -CPG_INF;
  int s1p = rl_begpos(v[i],F);  uint16_t s1c = rl_begcnt(v[i],F);
  int tp  = rl_endpos(v[i2],F); uint16_t tc  = rl_endcnt(v[i2],F);
  int s2p = rl_endpos(v[i3],F); uint16_t s2c = rl_endcnt(v[i3],F);
  if (!F) { int p = s1p; uint16_t c = s1c; s1p = s2p; s1c = s2c; s2p = p; s2c = c; }
  double est = cpg_lin_interp(W,tp,s2p,s2c,s1p,s1c);
  return (t == ST_D) ? est/tc : tc/est;
}

CPG_DEV void rl_extend_path(RelState &dst, const RelState &P, int t, int i)
{ dst.lastH = P.lastH; dst.lastD = P.lastD; dst.hbd = P.hbd; dst.dbh = P.dbh;
  if (t == ST_H)      { dst.dbh = P.lastD; dst.lastH = i; }
  else if (t == ST_D) { dst.hbd = P.lastH; dst.lastD = i; }
}

/* src/class_rel.c:279-513 */
CPG_DEV_NOINL void rl_update(ReadCtx &R, WCtx &W, const RelRun &U, int i, RelState *prv, RelState *cur)
{ (void)R;
  const int F = U.F;
  const cpg_dmodel *M = W.M;
  cpg_intvl *wint = U.wint;
  const cpg_intvl I = wint[i];
  const int ep = rl_endpos(I,F); const uint16_t ec = rl_endcnt(I,F);
  const int ip = rl_pred(i,F);
  double *tr = U.sh->tr;

  /* The 16 transitions as tasks.  The 8 with target H or D cost one Bessel recurrence each, whose length
     grows with the count difference |k| (2(|k|+sqrt(40|k|)) steps): about 20 for the state that fits the
     interval, about 80 for the other one.  The lanes of a warp wait for the longest recurrence of a round
     (SIMT), so the rounds are made homogeneous: a lane holds the H and the D transition of ONE predecessor
     state and takes the longer of the two first.  With 4 lanes per chain: round 0 = the long recurrences of
     all chains of the warp, round 1 = the short ones (before: every round as long as the longest).
     Then the E / R targets (no recurrence). */
  CPG_LOOP for (int s0 = W.glane; s0 < 4; s0 += W.gsize)
    { int kk[2] = {0,0}; double ll[2] = {0.,0.};
      const int live = (prv[s0].dp != -CPG_INF);
      if (live)
        { rl_hd_args(W,ST_H,I,prv[s0],F,kk[0],ll[0]);
          rl_hd_args(W,ST_D,I,prv[s0],F,kk[1],ll[1]);
        }
      const int first = (iabs(kk[1]) > iabs(kk[0]));            /* the longer recurrence first */
      CPG_LOOP for (int rr = 0; rr < 2; rr++)
        { const int x = rr ? !first : first;
          double v = 0.;
          if (live) v = cpg_exp(cpg_lp_skellam(kk[x],ll[x])+0.);
          tr[s0*4+ST_H+x] = v;
        }
    }
  CPG_LOOP for (int q0 = W.glane; q0 < 8; q0 += W.gsize)
    { const int s = q0 >> 1, t = ST_E+(q0 & 1);
      double v = 0.;
      if (prv[s].dp != -CPG_INF)
        { if (t == ST_E) v = cpg_exp(rl_lp_e(W,I,U.COV));
          else           v = cpg_exp(rl_lp_r(W,I,prv[s].cnt[ST_R],F,U.COV));
        }
      tr[s*4+t] = v;
    }
  CPG_SYNCGROUP(W);
  double psum = 0.;
  CPG_LOOP for (int q = 0; q < 16; q++) psum += tr[q];
  int fix = (psum == 0.);
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int q = W.glane; q < 16; q += W.gsize)
    { double v = tr[q];
      if (fix) v = ((q & 3) == ST_E) ? 1. : v;
      tr[q] = cpg_log(v/(fix ? 4. : psum));
    }
  CPG_SYNCGROUP(W);

  /* every live predecessor prefers R: freeze this interval (src/class_rel.c:348-380) */
  int only_r = 1;
  CPG_LOOP for (int s = 0; s < 4; s++)
    { int mt = rl_best_to(prv,tr,s);
      if (mt != ST_N && mt != ST_R) { only_r = 0; break; }
    }
  if (only_r)
    { if (W.glane == 0)
        { U.rpos[i] = 1;
          wint[i] = wint[ip];
          uint16_t bp = 0;
          CPG_LOOP for (int s = 0; s < 4; s++)
            { cur[s].dp = prv[s].dp;
              cur[s].dhr = -CPG_INF;
              bp |= (uint16_t)(((prv[s].dp == -CPG_INF) ? ST_N : s) << (3*s));
              if (prv[s].dp == -CPG_INF) continue;
              CPG_LOOP for (int t = 0; t < 4; t++) { cur[s].pos[t] = prv[s].pos[t]; cur[s].cnt[t] = prv[s].cnt[t]; }
              rl_extend_path(cur[s],prv[s],s,i);
            }
          U.bp[i] = bp;
        }
      CPG_SYNCGROUP(W);
      return;
    }

  int mh = rl_best_from(prv,tr,ST_H,0), md = rl_best_from(prv,tr,ST_D,0);
  if (mh == ST_H && md == ST_D)
    { double a = tr[ST_H*4+ST_H], b = tr[ST_D*4+ST_D];
      double m = (a < b) ? a : b;
      CPG_SYNCGROUP(W);
      if (W.glane == 0) { tr[ST_H*4+ST_H] = m; tr[ST_D*4+ST_D] = m; }
      CPG_SYNCGROUP(W);
    }

  /* the four target states are independent: one lane each */
  CPG_LOOP for (int t = W.glane; t < 4; t += W.gsize)
    { double mlp;
      int ms = rl_best_from(prv,tr,t,&mlp);
      U.sh->bpb[t] = ms;
      RelState ns;
      ns.dp = mlp; ns.dhr = -CPG_INF;
      CPG_LOOP for (int k = 0; k < 4; k++) { ns.pos[k] = 0; ns.cnt[k] = 0; }
      ns.lastH = ns.lastD = ns.hbd = ns.dbh = -1;
      if (ms != ST_N)
        { const RelState P = prv[ms];
          rl_extend_path(ns,P,t,i);
          if (t == ST_E)
            { CPG_LOOP for (int s = ST_R; s <= ST_D; s++) { ns.pos[s] = P.pos[s]; ns.cnt[s] = P.cnt[s]; } }
          else if (t == ST_R)
            { CPG_LOOP for (int s = ST_H; s <= ST_D; s++) { ns.pos[s] = rl_off(ep,F); ns.cnt[s] = P.cnt[s]; }
              uint16_t rc = (uint16_t)imin(ec,U.COV[ST_R]);
              if (P.cnt[ST_R] < rc) { ns.pos[ST_R] = P.pos[ST_R]; ns.cnt[ST_R] = P.cnt[ST_R]; }
              else                  { ns.pos[ST_R] = rl_off(ep,F); ns.cnt[ST_R] = rc; }
            }
          else
            { int ch, cd, cr;
              double r = rl_dh_ratio(W,U.wint,t,i,P,F);
              if (t == ST_H)
                { ch = ec;
                  if (r == -CPG_INF) cd = (P.lastD >= 0) ? P.cnt[ST_D] : ch+U.COV[ST_H];
                  else { cd = (int)(r*ch); ns.dhr = r; }
                }
              else
                { cd = ec;
                  if (r == -CPG_INF) ch = (P.lastH >= 0) ? P.cnt[ST_H] : imax(cd/2,cd-U.COV[ST_H]);
                  else { ch = (int)((double)cd/r); ns.dhr = r; }
                }
              cr = (int)(M->dr_ratio*cd);
              ns.pos[ST_H] = ns.pos[ST_D] = ns.pos[ST_R] = rl_off(ep,F);
              ns.cnt[ST_H] = (uint16_t)ch; ns.cnt[ST_D] = (uint16_t)cd; ns.cnt[ST_R] = (uint16_t)cr;
            }
          if (!(ns.cnt[ST_H] < ns.cnt[ST_D] && ns.cnt[ST_D] < ns.cnt[ST_R])) ns.dp = -CPG_INF;
        }
      cur[t] = ns;
    }
  CPG_SYNCGROUP(W);
  if (W.glane == 0)
    U.bp[i] = (uint16_t)(U.sh->bpb[0] | (U.sh->bpb[1] << 3) | (U.sh->bpb[2] << 6) | (U.sh->bpb[3] << 9));
  CPG_SYNCGROUP(W);
}

/* src/class_rel.c:515-614; the state path is written to `asgn` */
CPG_DEV_NOINL void rl_pass(ReadCtx &R, WCtx &W, const RelRun &U, uint8_t *asgn)
{ const int F = U.F, Mrel = U.M;
  const uint16_t *COV = U.COV;
  cpg_intvl *wint = U.wint;
  CPG_LOOP for (int i = W.glane; i < Mrel; i += W.gsize) { wint[i] = R.S.rint[i]; U.rpos[i] = 0; U.bp[i] = 0; }
  CPG_SYNCGROUP(W);

  const int POS_INIT = rl_off(F ? 0 : U.plen,F);
  int i = F ? 0 : Mrel-1;
  const cpg_intvl I = wint[i];
  RelState *c0 = U.sh->col[0], *c1 = U.sh->col[1];
  { const int ep = rl_endpos(I,F); const uint16_t ec = rl_endcnt(I,F), bc = rl_begcnt(I,F);
    double d[4];
    d[ST_E] = rl_lp_e(W,I,COV);
    d[ST_R] = rl_lp_r(W,I,COV[ST_R],F,COV);
    d[ST_H] = cpg_lp_poisson(W,bc,COV[ST_H]);
    d[ST_D] = cpg_lp_poisson(W,bc,COV[ST_D]);
    double psum = 0.;
    CPG_LOOP for (int s = 0; s < 4; s++) psum += cpg_exp(d[s]);
    CPG_LOOP for (int s = 0; s < 4; s++) d[s] = cpg_log(cpg_exp(d[s])/psum);
    if (W.glane == 0)
      { CPG_LOOP for (int s = 0; s < 4; s++)
          { RelState &X = c0[s];
            X.dp = d[s]; X.dhr = -CPG_INF;
            CPG_LOOP for (int t = ST_R; t <= ST_D; t++) { X.pos[t] = POS_INIT; X.cnt[t] = COV[t]; }
            X.pos[0] = 0; X.cnt[0] = 0;
            X.lastH = X.lastD = X.hbd = X.dbh = -1;
          }
        c0[ST_R].pos[ST_R] = ep; c0[ST_R].cnt[ST_R] = (uint16_t)imin(ec,COV[ST_R]);
        c0[ST_H].pos[ST_H] = ep; c0[ST_H].cnt[ST_H] = ec;
        c0[ST_H].pos[ST_D] = rl_off(ep,F); c0[ST_H].cnt[ST_D] = (uint16_t)(ec+COV[ST_H]);
        c0[ST_H].lastH = i;
        c0[ST_D].pos[ST_H] = rl_off(ep,F); c0[ST_D].cnt[ST_H] = (uint16_t)imax(ec/2,(int)ec-COV[ST_H]);
        c0[ST_D].pos[ST_D] = ep; c0[ST_D].cnt[ST_D] = ec;
        c0[ST_D].lastD = i;
      }
    CPG_SYNCGROUP(W);
  }

  RelState *prv = c0, *cur = c1;
  CPG_LOOP for (;;)
    { i = F ? i+1 : i-1;
      if ((F && i >= Mrel) || (!F && i < 0)) break;
      rl_update(R,W,U,i,prv,cur);
      RelState *t = prv; prv = cur; cur = t;
    }

  /* traceback (src/class_rel.c:605-613) from back pointers */
  i = F ? Mrel-1 : 0;
  int s = ST_N; { double mx = -CPG_INF; for (int x = 0; x < 4; x++) if (mx < prv[x].dp) { mx = prv[x].dp; s = x; } }
  if (s == ST_N) { W.status |= CPG_ST_UNDEF_TRACE; s = ST_E; }
  CPG_SYNCGROUP(W);
  if (W.glane == 0)
    { const int first = F ? 0 : Mrel-1;
      CPG_LOOP for (;;)
        { asgn[i] = (uint8_t)(U.rpos[i] ? ST_R : s);
          if (i == first) break;
          int p = (U.bp[i] >> (3*s)) & 7;
          s = (p == ST_N) ? ST_E : p;
          i = F ? i-1 : i+1;
        }
    }
  CPG_SYNCGROUP(W);
}

/* integer accumulation of src/class_rel.c:634-664 */
CPG_DEV_HELPER double rl_mean_cov(const cpg_intvl *r, const uint8_t *asgn, int Mrel, int want)
{ int lsum = 0, csum = 0;
  CPG_LOOP for (int i = 0; i < Mrel; i++)
    if (want < 0 || asgn[i] == want)
      { int l = r[i].e-r[i].b;
        lsum += l;
        csum += (r[i].ccb+r[i].cce)*l/2;
      }
  return (double)csum/lsum;
}

CPG_DEV int rl_has(const uint8_t *asgn, int Mrel, int s)
{ CPG_LOOP for (int i = 0; i < Mrel; i++) if (asgn[i] == s) return 1;
  return 0;
}

CPG_DEV_HELPER void rl_relabel(uint8_t *asgn, int Mrel, int from1, int to1, int from2, int to2, const WCtx &W)
{ CPG_SYNCGROUP(W);
  CPG_LOOP for (int i = W.glane; i < Mrel; i += W.gsize)
    { uint8_t a = asgn[i];
      if (from1 < 0 || a == from1) asgn[i] = (uint8_t)to1;
      else if (a == from2) asgn[i] = (uint8_t)to2;
    }
  CPG_SYNCGROUP(W);
}

/* src/class_rel.c:623-845: one direction with its optional re-run and relabel heuristics */
CPG_DEV_NOINL double rl_direction(ReadCtx &R, WCtx &W, RelShared *sh, int F, int Mrel, int plen, uint8_t *asgn)
{ const cpg_intvl *r = R.S.rint;
  const uint16_t *G = W.M->cov;
  RelRun U;
  U.F = F; U.M = Mrel; U.plen = plen; U.sh = sh;
  U.wint = R.S.wint+(F ? 0 : R.S.MC); U.bp = R.S.bp+(F ? 0 : R.S.MC); U.rpos = R.S.rpos+(F ? 0 : R.S.MC);
  CPG_LOOP for (int s = 0; s < 4; s++) U.COV[s] = G[s];
  rl_pass(R,W,U,asgn);
  if (!rl_has(asgn,Mrel,ST_H))
    { int anchor = -1;
      CPG_LOOP for (int i = 0; i < Mrel; i++)
        if (asgn[i] == ST_D) { if (F) { if (anchor == -1) anchor = i; } else anchor = i; }
      if (anchor >= 0)
        { double mean_d = rl_mean_cov(r,asgn,Mrel,ST_D);
          if (mean_d < G[ST_D])
            { U.COV[ST_H] = F ? r[anchor].ccb : r[anchor].cce;
              U.COV[ST_D] = (uint16_t)(U.COV[ST_H]+G[ST_H]);
              rl_pass(R,W,U,asgn);
              if (!rl_has(asgn,Mrel,ST_H))
                { mean_d = rl_mean_cov(r,asgn,Mrel,ST_D);
                  if (fabs(mean_d-G[ST_H]) <= fabs(mean_d-G[ST_D]))
                    rl_relabel(asgn,Mrel,ST_D,ST_H,-2,0,W);
                }
            }
        }
    }
  { int all_h = 1;
    CPG_LOOP for (int i = 0; i < Mrel; i++) if (asgn[i] != ST_H) all_h = 0;
    if (all_h)
      { double mean_h = rl_mean_cov(r,asgn,Mrel,-1);
        if (fabs(mean_h-G[ST_H]) >= fabs(mean_h-G[ST_D]))
          rl_relabel(asgn,Mrel,-1,ST_D,-2,0,W);
      }
  }
  { int n = 0;
    CPG_LOOP for (int i = 0; i < Mrel; i++) if (asgn[i] == ST_H) n++;
    if (n >= Mrel*0.7)
      { double mean_h = rl_mean_cov(r,asgn,Mrel,ST_H);
        if (fabs(mean_h-G[ST_H]) >= fabs(mean_h-G[ST_D]))
          rl_relabel(asgn,Mrel,ST_H,ST_D,ST_D,ST_R,W);
      }
  }
  int fd = -1, ld = -1, fh = -1, lh = -1;
  CPG_LOOP for (int i = 0; i < Mrel; i++)
    { if (asgn[i] == ST_D) { if (fd == -1) fd = i; ld = i; }
      else if (asgn[i] == ST_H) { if (fh == -1) fh = i; lh = i; }
    }
  return (fd >= 0 && fh >= 0) ? ((double)r[fd].ccb/r[fh].ccb)/((double)r[ld].cce/r[lh].cce) : 1.;
}

/* src/class_rel.c:847-869: state codes tested as booleans, first/last compared with `true` */
CPG_DEV int rl_eq_prefix(const uint8_t *a, int Mrel)
{ if (a[0] != 1) return 0;
  int i = 0;
  CPG_LOOP while (i < Mrel && a[i]) i++;
  CPG_LOOP for (; i < Mrel; i++) if (a[i]) return 0;
  return 1;
}
CPG_DEV int rl_eq_suffix(const uint8_t *a, int Mrel)
{ if (a[Mrel-1] != 1) return 0;
  int i = Mrel-2;
  CPG_LOOP while (i >= 0 && a[i]) i--;
  CPG_LOOP for (; i >= 0; i--) if (a[i]) return 0;
  return 1;
}

/* src/class_rel.c:871-963.  The forward and the backward pass are independent (the reference runs
 * them one after the other): with a full warp the two half-warps run them at the same time, each
 * with its own lane group, DP columns, working copy and back pointers. */
CPG_DEV_NOINL void classify_reliable(ReadCtx &R, WCtx &W, RelShared *sh)
{ const int Mrel = R.M, N = R.N;
  if (Mrel == 0) return;
  uint8_t *af = R.S.asg_f, *ab = R.S.asg_b;
  double hf, hb;
  CPG_SYNCGROUP(W);
  if (W.gsize >= 2)
    { const int hs = W.gsize >> 1, h = (W.glane >= hs);
      WCtx G = W;
      G.glane = W.glane-(h ? hs : 0); G.gsize = hs; G.gbase = W.gbase+(h ? hs : 0);
      G.gmask = ((hs >= 32) ? 0xffffffffu : ((1u << hs)-1u)) << G.gbase;
      G.status = 0;
      double hd = rl_direction(R,G,sh+h,h == 0,Mrel,R.plen,h ? ab : af);
      W.status |= G.status;
      if (G.glane == 0) W.ws->term[h] = hd;
      CPG_SYNCGROUP(W);
      hf = W.ws->term[0]; hb = W.ws->term[1];
    }
  else
    { hf = rl_direction(R,W,sh,1,Mrel,R.plen,af);
      hb = rl_direction(R,W,sh+1,0,Mrel,R.plen,ab);
    }
  int eq = 1;
  CPG_LOOP for (int i = 0; i < Mrel; i++) if (af[i] != ab[i]) { eq = 0; break; }
  int use_b = 0;
  if (!eq)
    { if (rl_eq_prefix(af,Mrel)) use_b = 0;
      else if (rl_eq_suffix(af,Mrel)) use_b = 1;
      else use_b = !(fabs(hf-1.) <= fabs(hb-1.));
    }
  const uint8_t *fin = use_b ? ab : af;
  CPG_SYNCGROUP(W);
  if (W.glane == 0)
    { cpg_intvl *v = R.S.intvl;
      CPG_LOOP for (int ri = 0, ii = 0; ri < Mrel; ri++, ii++)
        { CPG_LOOP while (ii < N && !v[ii].is_rel) ii++;
          if (ii >= N) break;
          v[ii].asgn = (int8_t)fin[ri];
        }
    }
  CPG_SYNCGROUP(W);
}

#endif
