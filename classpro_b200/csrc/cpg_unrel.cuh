/*******************************************************************************************
 *  cpg_unrel.cuh -- classification of the remaining (unreliable) intervals of one read and the
 *  whole per-read pipeline, one warp.
 *
 *  Replaces classify_unrel and everything under it (src/class_unrel.c:11-300) and the loop body
 *  src/ClassPro.c:229-271 (context is evaluated on demand, see cpg_context.cuh).
 *
 *  The two sweeps are order dependent (an interval's new state is visible to the ones handled
 *  after it, src/class_unrel.c:260-274) and stay serial; the stable sort by min(cb,ce) becomes a
 *  lane-parallel rank computation with the interval index as tie break, which is the order
 *  glibc's merge-sort qsort produces for the single-key comparator (src/class_unrel.c:244-246).
 *******************************************************************************************/
#ifndef CPG_UNREL_CUH
#define CPG_UNREL_CUH
#include "cpg_rel.cuh"

/* src/class_unrel.c:27-51 */
CPG_DEV_HELPER uint16_t un_est_cov(WCtx &W, int x, const cpg_intvl *v, int s, const int nb[2][2])
{ int l = nb[s == ST_D][0], r = nb[s == ST_D][1];
  if (l != -1 && r != -1) return (uint16_t)cpg_lin_interp(W,x,v[l].e-1,v[l].cce,v[r].b,v[r].ccb);
  if (l != -1) return v[l].cce;
  if (r != -1) return v[r].ccb;
  /* nothing of state s around: fall back on the other state, once */
  uint16_t c;
  l = nb[s != ST_D][0]; r = nb[s != ST_D][1];
  if (l != -1 && r != -1) c = (uint16_t)cpg_lin_interp(W,x,v[l].e-1,v[l].cce,v[r].b,v[r].ccb);
  else if (l != -1) c = v[l].cce;
  else if (r != -1) c = v[r].ccb;
  else c = 0;
  if (c > 0) return (uint16_t)((s == ST_H) ? c/2 : c*2);
  return W.M->cov[s];
}

/* src/class_unrel.c:53-65 */
CPG_DEV_HELPER double un_lp_e(const WCtx &W, const cpg_intvl &I)
{ const int ce = W.M->cov[ST_E];
  double po = cpg_lp_poisson(W,I.cb,ce)+cpg_lp_poisson(W,I.ce,ce)+CPG_E_PO_BASE;
  return dmax_ref(I.pe,po);
}

/* src/class_unrel.c:67-113 */
CPG_DEV_HELPER double un_lp_r(WCtx &W, int idx, const cpg_intvl *v, const int nb[2][2])
{ const cpg_intvl &I = v[idx];
  const cpg_dmodel *M = W.M;
  if (imax(I.cb,I.ce) >= M->cov[ST_R]) return 0.;
  const int l = nb[1][0], r = nb[1][1];
  uint16_t dl, dr;
  if (l == -1 && r == -1) dl = dr = M->cov[ST_D];
  else if (l == -1) dl = dr = v[r].cb;
  else if (r == -1) dl = dr = v[l].ce;
  else { dl = v[l].ce; dr = v[r].cb; }
  uint16_t rl = (uint16_t)(M->dr_ratio*dl), rr = (uint16_t)(M->dr_ratio*dr);
  if (I.cb >= rl || I.ce >= rr) return CPG_R_LOGP;
  double a = cpg_lp_binom99(W,I.cb,rl);
  double b = cpg_lp_binom99(W,I.ce,rr);
  return a+b;
}

/* src/class_unrel.c:115-237.  The ten independent pieces of an update -- the E and R
 * log-probabilities and, for H and D, the Skellam transition and the error-in-others tail on
 * either side -- are tasks; the arg-max (first maximum wins, order E,R,H,D) is formed afterwards.
 *   task 0: E     task 1: R     task 2+t, t = 4*h+2*side+kind (h: 0=H,1=D; side: 0=left,1=right;
 *   kind 0 = Skellam transition from/to the nearest fixed interval of that state,
 *   kind 1 = log binomial tail of the count against the interpolated coverage)
 * All ten are pure functions of the interval and of its four nearest reliable H / D neighbours
 * nb[][]: they change only when a neighbour does.  The sweeps change few states, so the tasks are
 * evaluated BEFORE the sweeps, for every interval a sweep will visit, on the states as they are
 * then (un_pre_interval: one interval per lane, k_unrel_a), and an update inside a sweep takes
 * the recorded values whenever the interval's neighbours are still the recorded ones (un_update:
 * a few loads and the arg-max); only otherwise are they evaluated again, there and then. */
CPG_DEV_HELPER void un_task_args(WCtx &W, const cpg_intvl &I, const cpg_intvl *v, const int nb[2][2], int t,
                                 int &mkind, int &mk, double &ma)
{ const int s = (t & 4) ? ST_D : ST_H, right = (t >> 1) & 1, kind = t & 1;
  mkind = 0; mk = 0; ma = 0.;
  if (kind == 0)
    { const int l = nb[s == ST_D][0], r = nb[s == ST_D][1];
      if (!right && l != -1)
        { int d = I.b-(v[l].e-1); if (d < 0) d = -d;
          mk = (int)I.cb-(int)v[l].cce; ma = (double)v[l].cce*d/W.M->read_len; mkind = 1;
        }
      if (right && r != -1)
        { int d = v[r].b-(I.e-1); if (d < 0) d = -d;
          mk = (int)v[r].ccb-(int)I.ce; ma = (double)v[r].ccb*d/W.M->read_len; mkind = 1;
        }
    }
  else
    { uint16_t est = un_est_cov(W,right ? I.e-1 : I.b,v,s,nb);
      uint16_t c = right ? I.ce : I.cb;
      if (est >= c) { mkind = 2; mk = est; ma = (double)c; }
    }
}

CPG_DEV_HELPER double un_task_eval(const WCtx &W, int mkind, int mk, double ma, int *bad)
{ double val = -CPG_INF;
  if (mkind == 1) val = cpg_lp_skellam(mk,ma);
  if (mkind == 2) val = cpg_log(cpg_p_errorin_lane(W.M->logfact,ET_OTHERS,cpg_rate_p1(W.M),mk,(int)ma,bad));
  return val;
}

/* nearest reliable interval assigned H (h = 0) / D (h = 1) on either side of idx, by one lane
   (src/class_unrel.c:11-25; most intervals are such, so the walks are short) */
CPG_DEV_HELPER void un_nn_walk(const cpg_intvl *v, int N, int idx, int nb[2][2])
{ CPG_LOOP for (int h = 0; h < 2; h++)
    { const int s = h ? ST_D : ST_H;
      int l = idx-1;
      CPG_LOOP while (l >= 0 && !(v[l].asgn == s && v[l].is_rel)) l--;
      int r = idx+1;
      CPG_LOOP while (r < N && !(v[r].asgn == s && v[r].is_rel)) r++;
      nb[h][0] = l; nb[h][1] = (r >= N) ? -1 : r;
    }
}

/* the pure step for one interval, by one lane: neighbours and the ten task values as they are now */
CPG_DEV_NOINL void un_pre_interval(WCtx &W, const cpg_intvl *v, int N, int idx, cpg_upre *out)
{ const cpg_intvl I = v[idx];
  cpg_upre U;
  U.st = 0;
  if (imax(I.cb,I.ce) >= W.M->cov[ST_R])
    { U.nb[0] = U.nb[1] = U.nb[2] = U.nb[3] = -2;           /* forced R: never looked at */
      CPG_LOOP for (int q = 0; q < 10; q++) U.val[q] = 0.;
    }
  else
    { int nb[2][2], bad = 0;
      const int st_in = W.status;
      W.status = 0;
      un_nn_walk(v,N,idx,nb);
      U.nb[0] = nb[0][0]; U.nb[1] = nb[0][1]; U.nb[2] = nb[1][0]; U.nb[3] = nb[1][1];
      U.val[0] = un_lp_e(W,I);
      U.val[1] = un_lp_r(W,idx,v,nb);
      CPG_LOOP for (int t = 0; t < 8; t++)
        { int mkind, mk; double ma;
          un_task_args(W,I,v,nb,t,mkind,mk,ma);
          U.val[2+t] = (mkind == 0) ? -CPG_INF : un_task_eval(W,mkind,mk,ma,&bad);
        }
      if (bad) W.status |= CPG_ST_BINOM;
      U.st = W.status;                                        /* raised only if the values are used */
      W.status = st_in;
    }
  U.pad = 0;
  *out = U;
}

/* src/class_unrel.c:185-237 for interval idx, whose recorded neighbours and task values are *U */
/* dirty: a reliable interval has entered or left H / D since the pure step (only then can a recorded
   neighbour be out of date: the neighbours are reliable H / D intervals, and those that were so at the
   start are never visited) */
CPG_DEV_NOINL void un_update(WCtx &W, int idx, cpg_intvl *v, int N, const cpg_upre *U, int &dirty)
{ const cpg_intvl I = v[idx];
  int ns;
  if (imax(I.cb,I.ce) >= W.M->cov[ST_R]) ns = ST_R;
  else
    { double *term = W.ws->term;
      int nb[2][2];
      int same = 1;
      if (!dirty) { nb[0][0] = U->nb[0]; nb[0][1] = U->nb[1]; nb[1][0] = U->nb[2]; nb[1][1] = U->nb[3]; }
      /* the four walks, one per lane where there are lanes */
      else if (W.gsize >= 4)
        { int mine = -1;
          if (W.glane < 4)
            { const int s = (W.glane & 2) ? ST_D : ST_H;
              if (W.glane & 1) { int r = idx+1; CPG_LOOP while (r < N && !(v[r].asgn == s && v[r].is_rel)) r++; mine = (r >= N) ? -1 : r; }
              else             { int l = idx-1; CPG_LOOP while (l >= 0 && !(v[l].asgn == s && v[l].is_rel)) l--; mine = l; }
            }
          nb[0][0] = (int)cpg_gshfl(W,(unsigned)mine,0); nb[0][1] = (int)cpg_gshfl(W,(unsigned)mine,1);
          nb[1][0] = (int)cpg_gshfl(W,(unsigned)mine,2); nb[1][1] = (int)cpg_gshfl(W,(unsigned)mine,3);
        }
      else un_nn_walk(v,N,idx,nb);
      if (dirty) same = (U->nb[0] == nb[0][0] && U->nb[1] == nb[0][1] && U->nb[2] == nb[1][0] && U->nb[3] == nb[1][1]);
      if (same) W.status |= U->st;
      else
        { /* a neighbour changed since the pure step: the tasks again, lanes in parallel */
          int bad = 0;
          CPG_SYNCGROUP(W);
          CPG_LOOP for (int q = W.glane; q < 10; q += W.gsize)
            { double val;
              if (q == 0) val = un_lp_e(W,I);
              else if (q == 1) val = un_lp_r(W,idx,v,nb);
              else
                { int mkind, mk; double ma;
                  un_task_args(W,I,v,nb,q-2,mkind,mk,ma);
                  val = (mkind == 0) ? -CPG_INF : un_task_eval(W,mkind,mk,ma,&bad);
                }
              term[q] = val;
            }
          CPG_SYNCGROUP(W);
          if (bad) W.status |= CPG_ST_BINOM;
        }
      const double *tv = same ? U->val : term;
      double mx = -CPG_INF; int ms = -1;
      CPG_LOOP for (int s = ST_E; s <= ST_D; s++)
        { double lp;
          if (s == ST_E) lp = tv[0];
          else if (s == ST_R) lp = tv[1];
          else
            { const double *t = tv+2+(s == ST_D ? 4 : 0);
              double er_l = -CPG_INF, er_r = -CPG_INF;
              if (idx-1 >= 0 && v[idx-1].asgn == s) er_l = I.peob;
              if (idx+1 < N && v[idx+1].asgn == s) er_r = I.peoe;
              double lpl = dmax_ref(dmax_ref(er_l,t[0]),t[1]);
              double lpr = dmax_ref(dmax_ref(er_r,t[2]),t[3]);
              if (lpl == -CPG_INF && lpr == -CPG_INF)
                { lpl = cpg_lp_poisson(W,I.cb,W.M->cov[s]); lpr = cpg_lp_poisson(W,I.ce,W.M->cov[s]); }
              else if (lpl == -CPG_INF) lpl = lpr;
              else if (lpr == -CPG_INF) lpr = lpl;
              lp = lpl+lpr;
            }
          if (mx < lp) { mx = lp; ms = s; }
        }
      if (ms == -1) { W.status |= CPG_ST_NO_PROB; ms = ST_E; }
      ns = ms;
    }
  CPG_SYNCGROUP(W);
  if (I.asgn != ns)
    { if (W.glane == 0) v[idx].asgn = (int8_t)ns;
      if (I.is_rel && (I.asgn == ST_H || I.asgn == ST_D || ns == ST_H || ns == ST_D)) dirty = 1;
    }
  CPG_SYNCGROUP(W);
}

CPG_DEV int un_is_fixed(const cpg_intvl &I) { return I.is_rel && (I.asgn == ST_H || I.asgn == ST_D); }

/* the intervals the sweeps visit (not fixed: src/class_unrel.c:249-251), in index order -> R.S.ord; returns how many */
CPG_DEV_NOINL int un_list(ReadCtx &R, const WCtx &W)
{ const cpg_intvl *v = R.S.intvl;
  int nf = 0;
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int base = 0; base < R.N; base += W.gsize)
    { const int i = base+W.glane;
      const int open = (i < R.N) && !un_is_fixed(v[i]);
      const unsigned m = cpg_gballot(W,open);
      if (open) R.S.ord[nf+cpg_popc(m & ((1u << W.glane)-1u))] = i;
      nf += cpg_popc(m);
    }
  CPG_SYNCGROUP(W);
  return nf;
}

/* src/class_unrel.c:248-275 on the listed intervals with their recorded values U[0..nf): the stable sort by
   min(cb,ce) as lane-parallel ranks (ties by index = the order glibc's merge sort leaves, :244-258) -- only
   the intervals the sweeps visit need an order --, then the two sweeps. */
CPG_DEV_NOINL void un_sweeps(ReadCtx &R, WCtx &W, int nf, const cpg_upre *U)
{ cpg_intvl *v = R.S.intvl;
  const int N = R.N;
  const int32_t *lst = R.S.ord;
  uint32_t *key = R.S.key;
  int32_t *srt = R.S.srt;
  CPG_LOOP for (int p = W.glane; p < nf; p += W.gsize) { const cpg_intvl &I = v[lst[p]]; key[p] = (uint32_t)imin(I.cb,I.ce); }
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int p = W.glane; p < nf; p += W.gsize)
    { const uint32_t kp = key[p];
      int rank = 0;
      for (int q = 0; q < p; q++) rank += (key[q] <= kp);
      for (int q = p+1; q < nf; q++) rank += (key[q] < kp);
      srt[rank] = p;
    }
  CPG_SYNCGROUP(W);
  int dirty = 0;
  CPG_LOOP for (int i = nf-1; i >= 0; i--) { const int p = srt[i]; un_update(W,lst[p],v,N,U+p,dirty); }
  CPG_LOOP for (int i = 0; i < nf; i++)    { const int p = srt[i]; un_update(W,lst[p],v,N,U+p,dirty); }
}

/* one read, both steps back to back (retry launch, host tests) */
CPG_DEV_NOINL void classify_unreliable(ReadCtx &R, WCtx &W)
{ const int nf = un_list(R,W);
  CPG_LOOP for (int p = W.glane; p < nf; p += W.gsize) un_pre_interval(W,R.S.intvl,R.N,R.S.ord[p],R.S.upre+p);
  CPG_SYNCGROUP(W);
  un_sweeps(R,W,nf,R.S.upre);
}

/* class characters of a read: 'N' x (K-1), then one character per k-mer (src/ClassPro.c:114-117,265-271) */
CPG_DEV_NOINL void emit_classes(const ReadCtx &R, const WCtx &W, uint8_t *cls)
{ const int K = W.M->K;
  CPG_LOOP for (int j = W.glane; j < K-1; j += W.gsize) cls[j] = 'N';
  const cpg_intvl *v = R.S.intvl;
  CPG_LOOP for (int i = 0; i < R.N; i++)
    { const int a = v[i].asgn;
      const char c = (a == ST_E) ? 'E' : (a == ST_R) ? 'R' : (a == ST_H) ? 'H' : (a == ST_D) ? 'D' : '?';
      const int b = v[i].b, e = v[i].e;
      CPG_LOOP for (int j = b+W.glane; j < e; j += W.gsize) cls[K-1+j] = (uint8_t)c;
    }
}

/* ---- the whole read: src/ClassPro.c:229-271, in three phases.
 * The kernel runs the phases CTA-synchronously (all warps of a CTA do phase 1 on their reads, then
 * phase 2, then phase 3): the per-read code is large and branchy, and warps that sit in the same
 * phase share their instruction-cache lines instead of evicting each other's. ---- */
CPG_DEV_NOINL void classify_phase1(ReadCtx &R, WCtx &W)
{ find_walls_and_reliable(R,W); }

CPG_DEV_NOINL void classify_phase2(ReadCtx &R, WCtx &W, RelShared *sh)
{ if (!(W.status & CPG_ST_ABORT)) classify_reliable(R,W,sh);
}

CPG_DEV_NOINL int classify_phase3(ReadCtx &R, WCtx &W, uint8_t *cls)
{ if (!(W.status & CPG_ST_ABORT)) classify_unreliable(R,W);
  emit_classes(R,W,cls);
  CPG_SYNCGROUP(W);
  return W.status;
}

CPG_DEV int classify_read(ReadCtx &R, WCtx &W, RelShared *sh, uint8_t *cls)
{ classify_phase1(R,W);
  classify_phase2(R,W,sh);
  return classify_phase3(R,W,cls);
}

#endif
