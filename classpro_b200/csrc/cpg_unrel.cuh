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

/* src/class_unrel.c:11-25 for both states at once: nb[h][side] = nearest reliable interval
 * assigned H (h = 0) or D (h = 1) to the left (side 0) / right (side 1) of idx, or -1.  The
 * reference walks interval by interval, once per use; here the lanes of the group look at
 * consecutive intervals together and the result is shared by all the tasks of the update. */
CPG_DEV_HELPER void un_nn_group(const WCtx &W, int idx, const cpg_intvl *v, int N, int nb[2][2])
{ nb[0][0] = nb[0][1] = nb[1][0] = nb[1][1] = -1;
  int need = 3;
  CPG_LOOP for (int base = idx-1; base >= 0 && need; base -= W.gsize)
    { const int j = base-W.glane;
      int a = -1;
      if (j >= 0 && v[j].is_rel) a = v[j].asgn;
      if (need & 1) { unsigned m = cpg_gballot(W,a == ST_H); if (m) { nb[0][0] = base-(cpg_ffs(m)-1); need &= ~1; } }
      if (need & 2) { unsigned m = cpg_gballot(W,a == ST_D); if (m) { nb[1][0] = base-(cpg_ffs(m)-1); need &= ~2; } }
    }
  need = 3;
  CPG_LOOP for (int base = idx+1; base < N && need; base += W.gsize)
    { const int j = base+W.glane;
      int a = -1;
      if (j < N && v[j].is_rel) a = v[j].asgn;
      if (need & 1) { unsigned m = cpg_gballot(W,a == ST_H); if (m) { nb[0][1] = base+(cpg_ffs(m)-1); need &= ~1; } }
      if (need & 2) { unsigned m = cpg_gballot(W,a == ST_D); if (m) { nb[1][1] = base+(cpg_ffs(m)-1); need &= ~2; } }
    }
}

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
 * either side -- are tasks; the arg-max (first maximum wins, order E,R,H,D) is formed uniformly
 * afterwards.
 *   task 0: E     task 1: R     task 2+t, t = 4*h+2*side+kind (h: 0=H,1=D; side: 0=left,1=right;
 *   kind 0 = Skellam transition from/to the nearest fixed interval of that state,
 *   kind 1 = log binomial tail of the count against the interpolated coverage)
 * Tasks 2..9 are pure functions of a small argument triple (what, integer, double), which is how
 * they are memoised: the triple is formed by un_task_args, the value by un_task_eval. */
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

/* Before the sweeps: the tasks of MANY intervals at once, one interval per lane, on the states
 * as they are now.  The sweeps change few states, so most of their tasks then find their triple
 * in the memo; evaluated inside un_update the same tasks run on 2-3 lanes of the group (ncu).
 * A memo entry is (triple, value) and the value depends on nothing else, so an entry left by an
 * earlier read is as good as a fresh one. */
CPG_DEV_NOINL void un_precompute(WCtx &W, cpg_intvl *v, int N, const uint8_t *fixed, cpg_unmemo *memo)
{ const int rcov = W.M->cov[ST_R];
  const int n = imin(N,CPG_MEMO_CAP);
  int bad = 0;
  CPG_LOOP for (int idx = W.glane; idx < n; idx += W.gsize)
    { if (fixed[idx]) continue;
      const cpg_intvl I = v[idx];
      if (imax(I.cb,I.ce) >= rcov) continue;
      int nb[2][2];
      CPG_LOOP for (int h = 0; h < 2; h++)
        { const int s = h ? ST_D : ST_H;
          int l = idx-1;
          CPG_LOOP while (l >= 0 && !(v[l].asgn == s && v[l].is_rel)) l--;
          int r = idx+1;
          CPG_LOOP while (r < N && !(v[r].asgn == s && v[r].is_rel)) r++;
          nb[h][0] = l; nb[h][1] = (r >= N) ? -1 : r;
        }
      cpg_unmemo *mm = memo+(size_t)idx*8;
      CPG_LOOP for (int t = 0; t < 8; t++)
        { int mkind, mk; double ma;
          un_task_args(W,I,v,nb,t,mkind,mk,ma);
          if (mkind == 0 || (mm[t].kind == mkind && mm[t].k == mk && mm[t].a == ma)) continue;
          const double val = un_task_eval(W,mkind,mk,ma,&bad);
          mm[t].kind = mkind; mm[t].k = mk; mm[t].a = ma; mm[t].val = val;
        }
    }
  if (bad) W.status |= CPG_ST_BINOM;
  CPG_SYNCGROUP(W);
}

CPG_DEV_NOINL void un_update(WCtx &W, int idx, cpg_intvl *v, int N, cpg_unmemo *memo)
{ const cpg_intvl I = v[idx];
  int ns;
  if (imax(I.cb,I.ce) >= W.M->cov[ST_R]) ns = ST_R;
  else
    { double *term = W.ws->term;
      int bad = 0;
      cpg_unmemo *mm = (memo != 0 && idx < CPG_MEMO_CAP) ? memo+(size_t)idx*8 : 0;
      int nb[2][2];
      CPG_SYNCGROUP(W);
      un_nn_group(W,idx,v,N,nb);
      CPG_LOOP for (int q = W.glane; q < 10; q += W.gsize)
        { double val;
          if (q == 0) val = un_lp_e(W,I);
          else if (q == 1) val = un_lp_r(W,idx,v,nb);
          else
            { const int t = q-2;
              int mkind, mk; double ma;
              un_task_args(W,I,v,nb,t,mkind,mk,ma);
              if (mkind == 0) val = -CPG_INF;                      /* no task: never memoised */
              else if (mm != 0 && mm[t].kind == mkind && mm[t].k == mk && mm[t].a == ma) val = mm[t].val;
              else
                { val = un_task_eval(W,mkind,mk,ma,&bad);
                  if (mm != 0) { mm[t].kind = mkind; mm[t].k = mk; mm[t].a = ma; mm[t].val = val; }
                }
            }
          term[q] = val;
        }
      CPG_SYNCGROUP(W);
      if (bad) W.status |= CPG_ST_BINOM;
      double mx = -CPG_INF; int ms = -1;
      CPG_LOOP for (int s = ST_E; s <= ST_D; s++)
        { double lp;
          if (s == ST_E) lp = term[0];
          else if (s == ST_R) lp = term[1];
          else
            { const double *t = term+2+(s == ST_D ? 4 : 0);
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
  if (W.glane == 0 && I.asgn != ns) v[idx].asgn = (int8_t)ns;
  CPG_SYNCGROUP(W);
}

/* src/class_unrel.c:248-275 */
CPG_DEV_NOINL void classify_unreliable(ReadCtx &R, WCtx &W)
{ cpg_intvl *v = R.S.intvl;
  const int N = R.N;
  int32_t *ord = R.S.ord;
  uint8_t *fixed = R.S.fixed;
  /* keys in a compact array first, then the ranks */
  uint32_t *key = R.S.key;
  CPG_LOOP for (int i = W.glane; i < N; i += W.gsize) key[i] = (uint32_t)imin(v[i].cb,v[i].ce);
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int i = W.glane; i < N; i += W.gsize)
    { const uint32_t ki = key[i];
      int rank = 0;
      for (int j = 0; j < i; j++) rank += (key[j] <= ki);
      for (int j = i+1; j < N; j++) rank += (key[j] < ki);
      ord[rank] = i;
      fixed[i] = (uint8_t)(v[i].is_rel && (v[i].asgn == ST_H || v[i].asgn == ST_D));
    }
  CPG_SYNCGROUP(W);
  un_precompute(W,v,N,fixed,R.S.memo);
  CPG_LOOP for (int i = N-1; i >= 0; i--) { int x = ord[i]; if (!fixed[x]) un_update(W,x,v,N,R.S.memo); }
  CPG_LOOP for (int i = 0; i < N; i++)    { int x = ord[i]; if (!fixed[x]) un_update(W,x,v,N,R.S.memo); }
}

/* ---- the whole read: src/ClassPro.c:229-271, in three phases.
 * The kernel runs the phases CTA-synchronously (all warps of a CTA do phase 1 on their reads, then
 * phase 2, then phase 3): the per-read code is large and branchy, and warps that sit in the same
 * phase share their instruction-cache lines instead of evicting each other's. ---- */
CPG_DEV_NOINL void classify_phase1(ReadCtx &R, WCtx &W)
{ find_walls_and_reliable(R,W); }

CPG_DEV_NOINL void classify_phase2(ReadCtx &R, WCtx &W, RelShared *sh)
{ if (!(W.status & CPG_ST_ABORT)) classify_reliable(R,W,sh); }

CPG_DEV_NOINL int classify_phase3(ReadCtx &R, WCtx &W, uint8_t *cls)
{ const int K = W.M->K;
  if (!(W.status & CPG_ST_ABORT)) classify_unreliable(R,W);
  /* emit: 'N' x (K-1), then one class character per k-mer */
  CPG_LOOP for (int j = W.glane; j < K-1; j += W.gsize) cls[j] = 'N';
  const cpg_intvl *v = R.S.intvl;
  CPG_LOOP for (int i = 0; i < R.N; i++)
    { const int a = v[i].asgn;
      const char c = (a == ST_E) ? 'E' : (a == ST_R) ? 'R' : (a == ST_H) ? 'H' : (a == ST_D) ? 'D' : '?';
      const int b = v[i].b, e = v[i].e;
      CPG_LOOP for (int j = b+W.glane; j < e; j += W.gsize) cls[K-1+j] = (uint8_t)c;
    }
  CPG_SYNCGROUP(W);
  return W.status;
}

CPG_DEV int classify_read(ReadCtx &R, WCtx &W, RelShared *sh, uint8_t *cls)
{ classify_phase1(R,W);
  classify_phase2(R,W,sh);
  return classify_phase3(R,W,cls);
}

#endif
