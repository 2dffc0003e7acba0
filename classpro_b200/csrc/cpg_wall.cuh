/*******************************************************************************************
 *  cpg_wall.cuh -- wall detection and reliable-interval selection for one read, one warp.
 *
 *  Replaces find_wall (+ find_gain/find_drop/find_pair, update_perror, remove_duplicates,
 *  bs_eintvl) src/wall.c:264-958 and find_rel_intvl/correct_wall_cnt src/wall.c:960-1051.
 *
 *  Layout differences from the reference (results identical):
 *   - one flag byte per profile position (the reference's wall flags plus a "has slot" bit); the
 *     four error probabilities of a position live in a lazily allocated slot whose 16-bit index
 *     sits in a side array that never needs a reset (it is only read where the flag says so).
 *     The reference keeps a flag byte plus four doubles for every position (33 B/position reset
 *     per read, here 1 B/position);
 *   - wall candidates come as a bit map written by the profile decoder (cpg_decode.cuh: a
 *     candidate can only arise at a delta token); lone O-walls and interval boundaries are found by
 *     lane-parallel sweeps over the flag bytes, 16 positions per lane and load.  All of them are
 *     then handled in position order, because pairing, the first-writer-wins probability cache
 *     and the paired flags are order dependent (src/wall.c:310-315,639-640);
 *   - pairs explained by errors in others are not stored: their only use in the reference is to
 *     clear the O-wall flag of both ends (src/wall.c:722-726), which commutes with the rest of
 *     pass A and is done at pairing time;
 *   - index plen of the scratch is reset with the rest (the reference leaves stale state there,
 *     SURVEY A.5), and profile[plen], which src/wall.c:977-978 can read one past the end when a
 *     low-complexity run reaches the end of the read, is defined as profile[plen-1].
 *******************************************************************************************/
#ifndef CPG_WALL_CUH
#define CPG_WALL_CUH
#include "cpg_math.cuh"
#include "cpg_context.cuh"

/* ---- lane-group primitives.  A read is owned by a GROUP of lanes (a whole warp, or an aligned
 *      half / quarter of one: several reads then share a warp and their instruction streams
 *      interleave).  Ballots are returned relative to the group (bit j = group lane j).
 *      Width 1 in the host-side unit-test build. ---- */
#if defined(CPG_HOSTSIM) && CPG_HOSTSIM == 32
CPG_DEV unsigned cpg_gballot(const WCtx &W, int pred) { return cpg_sim_gballot(W.gmask,pred) >> W.gbase; }
CPG_DEV int      cpg_gsum(const WCtx &W, int v)       { return cpg_sim_gsum(W.gmask,v); }
CPG_DEV unsigned cpg_gshfl(const WCtx &W, unsigned v, int l) { return cpg_sim_gshfl(W.gmask,v,W.gbase+l); }
CPG_DEV int      cpg_ffs(unsigned m)  { return __builtin_ffs((int)m); }
#elif defined(CPG_HOSTSIM)
CPG_DEV unsigned cpg_gballot(const WCtx &W, int pred) { (void)W; return pred ? 1u : 0u; }
CPG_DEV int      cpg_gsum(const WCtx &W, int v)       { (void)W; return v; }
CPG_DEV unsigned cpg_gshfl(const WCtx &W, unsigned v, int l) { (void)W; (void)l; return v; }
CPG_DEV int      cpg_ffs(unsigned m)  { return __builtin_ffs((int)m); }
#else
CPG_DEV unsigned cpg_gballot(const WCtx &W, int pred) { return __ballot_sync(W.gmask,pred) >> W.gbase; }
CPG_DEV int      cpg_gsum(const WCtx &W, int v)       { return __reduce_add_sync(W.gmask,v); }
CPG_DEV unsigned cpg_gshfl(const WCtx &W, unsigned v, int l) { return __shfl_sync(W.gmask,v,W.gbase+l); }
CPG_DEV int      cpg_ffs(unsigned m)  { return __ffs((int)m); }
#endif

/* 16 flag bytes at a 16-byte aligned address, as four little-endian words */
CPG_DEV void cpg_ld16(const uint8_t *p, unsigned w[4])
{
#ifdef CPG_HOSTSIM
  memcpy(w,p,16);
#else
  const uint4 v = *reinterpret_cast<const uint4 *>(p);
  w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#endif
}
/* bit 0 of each of the four bytes of x -> bits 0..3 */
CPG_DEV unsigned cpg_pack4(unsigned x) { return ((x & 0x01010101u)*0x00204081u >> 21) & 0xfu; }

/* flag bits (src/wall.c:264-269) of a mark byte, plus MK_SLOT: slot[] holds this position's slot */
#define MK_BY_S       0x01u
#define MK_PAIR_S     0x02u
#define MK_SLOT       0x04u
#define MK_BY_O       0x10u
#define MK_PAIR_O     0x20u
#define MK_PAIR_MULT  0x40u
#define MK_ERROR      0x80u
#define MK_STALE_PROF 128     /* status bit: profile[plen] was read (reference reads stale memory) */

struct ReadCtx
  { const uint16_t *prof;
    int             plen, rlen;
    cpg_seq         seq;
    const uint32_t *cand;     /* wall-candidate bit map of the read (bit i of word i>>5 = position i) */
    cpg_scratch     S;
    int             nslots;
    int             N, M;
  };

CPG_DEV uint16_t rc_prof(const ReadCtx &R, WCtx &W, int p)
{ if (p >= R.plen) { W.status |= MK_STALE_PROF; p = R.plen-1; }
  return R.prof[p];
}

CPG_DEV unsigned mk_by(int e)   { return e == ET_SELF ? MK_BY_S : MK_BY_O; }
CPG_DEV unsigned mk_pair(int e) { return e == ET_SELF ? MK_PAIR_S : MK_PAIR_O; }

/* Lane 0 is the only writer of the scratch words below.  Each helper synchronises the warp BEFORE
 * the write (the other lanes may still be reading the old value: the CUDA memory model does not
 * promise lock-step execution) and AFTER it (so that every lane sees the new one). */
CPG_DEV_HELPER void mark_or(ReadCtx &R, const WCtx &W, int pos, unsigned bits)
{ CPG_SYNCGROUP(W);
  if (W.glane == 0) R.S.mark[pos] |= (uint8_t)bits;
  CPG_SYNCGROUP(W);
}
CPG_DEV_HELPER void mark_clear(ReadCtx &R, const WCtx &W, int pos, unsigned bits)
{ CPG_SYNCGROUP(W);
  if (W.glane == 0) R.S.mark[pos] &= (uint8_t)~bits;
  CPG_SYNCGROUP(W);
}

CPG_DEV_HELPER double perr_get(const ReadCtx &R, int pos, int e, int w)
{ if (!(R.S.mark[pos] & MK_SLOT)) return -CPG_INF;
  return R.S.perr[(size_t)R.S.slot[pos]*4+e*2+w];
}

/* src/wall.c:317-322 */
CPG_DEV_HELPER double lp_diff_pair(const ReadCtx &R, const WCtx &W, int i, int j)
{ const uint16_t *p = R.prof;
  int n_drop = (int)p[i-1]-p[i], n_gain = (int)p[j]-p[j-1];
  uint16_t cov = (uint16_t)imax(p[i-1],p[j]);
  return cpg_lp_trans(W,i,j,n_drop,n_gain,cov);
}

CPG_DEV int cthres_at(const WCtx &W, int t, int l, int cout, int s, int e)
{ return W.cthres[((CPG_LROW(t,l)*256+cout)*2+s)*2+e]; }

/* src/wall.c:324-329 (cin travels through an 8-bit parameter in the reference) */
CPG_DEV int thres_ng(int e, int cin, int ct)
{ cin &= 0xff; return (e == ET_SELF) ? (cin >= ct) : (cin < ct); }

/* store a freshly computed probability in the first-writer-wins cache (src/wall.c:310-315) */
CPG_DEV_NOINL void perr_store(ReadCtx &R, WCtx &W, int pos, int e, int w, double v)
{ const unsigned m = R.S.mark[pos];
  unsigned s = (m & MK_SLOT) ? R.S.slot[pos] : 0u;
  CPG_SYNCGROUP(W);
  if (!(m & MK_SLOT))
    { if (R.nslots >= R.S.capS) { W.status |= CPG_ST_RETRY; return; }       /* uniform in the group */
      s = (unsigned)(R.nslots++);
      if (W.glane == 0)
        { R.S.mark[pos] = (uint8_t)(m | MK_SLOT);
          R.S.slot[pos] = (uint16_t)s;
          double *q = R.S.perr+(size_t)s*4;
          q[0] = q[1] = q[2] = q[3] = -CPG_INF;
        }
      CPG_SYNCGROUP(W);
    }
  if (W.glane == 0) R.S.perr[(size_t)s*4+e*2+w] = v;
  CPG_SYNCGROUP(W);
}

/* Everything find_gain/find_drop (src/wall.c:331-507) need to know about one candidate.
 * fwd = 1: a DROP at i looks for its GAIN about K-1 positions ahead; fwd = 0: a GAIN at i looks
 * for its DROP behind.  Partner slot 0 is the low-complexity partner (context run walked by whole
 * units), slots 1..6 the high-complexity partners at 0..MAX_N_HC extra bases. */
struct PairGeom
  { int      fwd, i, t, l, lc_kind;       /* lc_kind: 0 = no partner (find_* returns false), 1 = read boundary, 2 = regular */
    int      lc_j;
    uint16_t cout, cin;
    double   erate;
  };

CPG_DEV int pg_hc_j(const PairGeom &G, int K, int n) { return G.fwd ? G.i+K-1+n : G.i-K+1-n; }
CPG_DEV int pg_in_range(const PairGeom &G, int plen, int j) { return G.fwd ? (j < plen) : (j > 0); }
CPG_DEV void pg_counts(const PairGeom &G, const uint16_t *prof, int j, uint16_t &cin_j, uint16_t &cout_j)
{ cin_j  = G.fwd ? prof[j-1] : prof[j];
  cout_j = G.fwd ? prof[j]   : prof[j-1];
}
/* count tests of the low-complexity partner (src/wall.c:364-365,455-456) */
CPG_DEV_HELPER int pg_lc_ok(const PairGeom &G, const WCtx &W, const uint16_t *prof, int e)
{ uint16_t cin_j, cout_j;
  pg_counts(G,prof,G.lc_j,cin_j,cout_j);
  return cin_j <= cout_j
         && !(cout_j < W.M->cmax && thres_ng(e,cin_j,cthres_at(W,G.t,G.l,cout_j,TH_FINAL,e)));
}
/* count tests of a high-complexity partner (src/wall.c:384-389,475-480) */
CPG_DEV_HELPER int pg_hc_ok(const PairGeom &G, const WCtx &W, const uint16_t *prof, int e, int j)
{ uint16_t cin_j, cout_j;
  pg_counts(G,prof,j,cin_j,cout_j);
  const int cmax = W.M->cmax;
  if (!(cin_j <= cout_j)) return 0;
  if ((G.cout < cmax && thres_ng(e,G.cin,cthres_at(W,CT_HP,1,G.cout,TH_FINAL,e)))
      || (cout_j < cmax && thres_ng(e,cin_j,cthres_at(W,CT_HP,1,cout_j,TH_FINAL,e))))
    return 0;
  return 1;
}

/* Layout of the per-candidate task results in the warp's exchange block:
 *   [e*8+0]      p_errorin of the low-complexity partner under this candidate's error rate
 *   [e*8+1+n]    p_errorin of high-complexity partner n under HC_ERATE
 *   [e*8+7]      p_errorin of the candidate itself under HC_ERATE
 *   [16+p]       log Skellam probability that candidate and partner p belong together (OTHERS) */

/* The decision part of find_gain/find_drop, replayed serially on the precomputed values. */
CPG_DEV_NOINL int pair_replay(ReadCtx &R, WCtx &W, const PairGeom &G, int e, cpg_eintvl *out)
{ const uint16_t *prof = R.prof;
  const int plen = R.plen, K = W.M->K;
  const int wi = G.fwd ? WT_DROP : WT_GAIN, wj = G.fwd ? WT_GAIN : WT_DROP;
  const double *term = W.ws->term;
  int max_j = -1; double max_pe = -CPG_INF, pe;
  if (G.lc_kind == 0) return 0;
  int j = G.lc_j;
  if (G.lc_kind == 1)
    { double pi = perr_get(R,G.i,e,wi);
      pe = pi*pi;
    }
  else
    { pe = -CPG_INF;
      if (pg_lc_ok(G,W,prof,e) && (e == ET_SELF || term[16] >= CPG_THRES_DIFF_EO))
        { if (perr_get(R,j,e,wj) == -CPG_INF) perr_store(R,W,j,e,wj,term[e*8]);
          pe = G.fwd ? perr_get(R,G.i,e,WT_DROP)*perr_get(R,j,e,WT_GAIN)
                     : perr_get(R,j,e,WT_DROP)*perr_get(R,G.i,e,WT_GAIN);
        }
    }
  if (max_pe < pe) { max_j = j; max_pe = pe; }
  CPG_LOOP for (int n = 0; n <= CPG_MAX_N_HC; n++)
    { j = pg_hc_j(G,K,n);
      if (!pg_in_range(G,plen,j)) break;
      if (!pg_hc_ok(G,W,prof,e,j)) continue;
      if (e == ET_OTHERS && term[17+n] < CPG_THRES_DIFF_EO) continue;
      pe = term[e*8+7]*term[e*8+1+n];
      if (max_pe < pe) { max_j = j; max_pe = pe; }
    }
  if (max_j == -1) return 0;
  if (G.fwd) { out->b = G.i; out->e = max_j; }
  else       { out->b = max_j; out->e = G.i; }
  out->pe = max_pe;
  return 1;
}

/* ---- E-interval list helpers: order of src/wall.c:519-528 under a stable sort is (b,e) then
 *      input order, since the (int) cast of a probability difference is 0 ---- */
CPG_DEV int ei_before(const cpg_eintvl &x, const cpg_eintvl &y)
{ if (x.b == y.b)
    { if (x.e == y.e) return ((int)(y.pe-x.pe)) < 0;
      return x.e < y.e;
    }
  return x.b < y.b;
}

/* stable insertion sort; the lists are produced almost in order */
CPG_DEV_NOINL void ei_sort(cpg_eintvl *a, int n, const WCtx &W)
{ CPG_SYNCGROUP(W);
  if (W.glane == 0)
    CPG_LOOP for (int i = 1; i < n; i++)
      { cpg_eintvl v = a[i];
        int j = i-1;
        CPG_LOOP while (j >= 0 && ei_before(v,a[j])) { a[j+1] = a[j]; j--; }
        a[j+1] = v;
      }
  CPG_SYNCGROUP(W);
}

/* src/wall.c:548-568 */
CPG_DEV_HELPER int ei_unique(cpg_eintvl *a, int n, const WCtx &W)
{ ei_sort(a,n,W);
  if (n >= 2)
    { int i = 1;
      CPG_LOOP while (i < n && !(a[i-1].b == a[i].b && a[i-1].e == a[i].e)) i++;
      /* every lane needs the new length: count first (read only), then lane 0 compacts */
      int keep_b = (i < n) ? a[i-1].b : 0, keep_e = (i < n) ? a[i-1].e : 0;
      int cnt = i;
      CPG_LOOP for (int j = i+1; j < n; j++)
        if (!(keep_b == a[j].b && keep_e == a[j].e))
          { keep_b = a[j].b; keep_e = a[j].e; cnt++; }
      CPG_SYNCGROUP(W);
      if (W.glane == 0)
        { int w = i;
          CPG_LOOP for (int j = i+1; j < n; j++)
            if (!(a[w-1].b == a[j].b && a[w-1].e == a[j].e))
              a[w++] = a[j];
        }
      CPG_SYNCGROUP(W);
      n = cnt;
    }
  return n;
}

/* src/wall.c:530-546 */
CPG_DEV_HELPER int ei_find(const cpg_eintvl *a, int l, int r, int b, int e)
{ CPG_LOOP while (l <= r)
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

CPG_DEV_HELPER void ei_put(ReadCtx &R, WCtx &W, int k, int b, int e, double pe)
{ if (k >= R.S.capE) { W.status |= CPG_ST_RETRY; return; }                  /* uniform in the group */
  CPG_SYNCGROUP(W);
  if (W.glane == 0) { R.S.eint[k].b = b; R.S.eint[k].e = e; R.S.eint[k].pe = pe; }
  CPG_SYNCGROUP(W);
}

/* clear bits on the open range (b,e), lanes striding */
CPG_DEV_HELPER void mark_clear_range(ReadCtx &R, const WCtx &W, int b, int e, unsigned bits)
{ CPG_SYNCGROUP(W);
  CPG_LOOP for (int j = b+1+W.glane; j < e; j += W.gsize) R.S.mark[j] &= (uint8_t)~bits;
  CPG_SYNCGROUP(W);
}

/* ---- pass A for one candidate position (src/wall.c:606-692) ----
 * The reference evaluates, one after the other, up to 18 binomial tails and 7 Skellam
 * probabilities per candidate.  They are pure functions of the profile, so here they are formed as
 * independent tasks, one per lane, and evaluated together (stage 1: the candidate's own
 * probabilities; stage 2: every partner of both error types); the order-dependent part -- the
 * first-writer-wins probability cache, the paired flags, the E-interval list -- is then replayed
 * serially on the results, in the reference's order. */
CPG_DEV_NOINL void wall_candidate(ReadCtx &R, WCtx &W, int i, int &eidx)
{ const cpg_dmodel *M = W.M;
  const uint16_t *prof = R.prof;
  const double *lf = M->logfact;
  const int plen = R.plen, K = M->K, cmax = M->cmax;
  const uint16_t cim1 = prof[i-1], ci = prof[i];
  const int cng = (cim1 > ci) ? cim1-ci : ci-cim1;
  int wtype; uint16_t cin, cout;
  if (cim1 > ci) { wtype = WT_DROP; cin = ci;   cout = cim1; }
  else           { wtype = WT_GAIN; cin = cim1; cout = ci;   }

  int maxt = -1, maxl = -1; double maxpe = -CPG_INF;
  { int cl[3];
    cpg_ctx3_at(R.seq,R.rlen,K,wtype,i,cl);
    if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) W.status |= CPG_ST_LONG_RUN;
    CPG_LOOP for (int t = 0; t < CT_N; t++)
      { int l = imin(cl[t],M->lmax[t]);
        double pe = M->pe[t][l];
        if (maxpe < pe) { maxpe = pe; maxt = t; maxl = l; }
      }
  }

  /* stage 0: how far does each error type get before any probability is needed */
  int reach[2] = {0,0}, o_wall_now = 0;
  const unsigned mi = R.S.mark[i];
  CPG_LOOP for (int e = ET_SELF; e <= ET_OTHERS; e++)
    { if (mi & mk_pair(e)) continue;
      int ct_final = 0;
      if (cout < cmax)
        { int ct_init = cthres_at(W,maxt,maxl,cout,TH_INIT,e);
          ct_final = cthres_at(W,maxt,maxl,cout,TH_FINAL,e);
          if (!(cng > CPG_MAX_CNT_CHANGE || cin < imax(ct_init,3))) continue;
        }
      if (e == ET_SELF)
        { if (cout < cmax && cin >= ct_final) continue;
          reach[e] = 1;
        }
      else
        { if (cng >= M->cov[ST_H] || (cout < cmax && cin < ct_final)) { o_wall_now = 1; continue; }
          reach[e] = 1;
        }
    }
  if (!reach[0] && !reach[1])
    { if (o_wall_now) mark_or(R,W,i,MK_BY_O);
      return;
    }

  /* stage 1: the candidate's own probabilities, lanes 0 and 1 */
  double *term = W.ws->term;
  int bad = 0;
  int fresh[2];
  CPG_LOOP for (int e = 0; e < 2; e++) fresh[e] = reach[e] && perr_get(R,i,e,wtype) == -CPG_INF;
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int q = W.glane; q < 2; q += W.gsize)
    if (fresh[q]) term[q] = cpg_p_errorin_lane(lf,q,cpg_rate_pe(M,maxt,maxl),cout,cin,&bad);
  CPG_SYNCGROUP(W);
  int go[2];
  CPG_LOOP for (int e = 0; e < 2; e++)
    { if (fresh[e]) perr_store(R,W,i,e,wtype,term[e]);
      go[e] = reach[e] && !(perr_get(R,i,e,wtype) < CPG_PE_FINAL);
    }

  cpg_eintvl I;
  if (go[0] || go[1])
    { /* stage 2: partner geometry (src/wall.c:344-357,432-450), shared by both error types */
      PairGeom G;
      G.fwd = (wtype == WT_DROP); G.i = i; G.t = maxt; G.l = maxl; G.cout = cout; G.cin = cin; G.erate = maxpe;
      { const int ulen = maxt+1, m = ulen*maxl;
        int n = 0;
        CPG_LOOP for (;;)
          { int idx = G.fwd ? i+ulen*(n+1) : i-ulen*(n+1);
            if (G.fwd) { if (idx >= plen) break; }
            else       { if (idx <= 0) break; }
            const int cx = cpg_ctx_at(R.seq,R.rlen,K,wtype,idx,maxt);
            if (cx >= 127) W.status |= CPG_ST_LONG_RUN;
            if (cx != m+n+1) break;
            n++;
          }
        int j = G.fwd ? i+K-1+n-m : i-K+1-n+m;
        if (G.fwd ? (j <= i) : (j >= i)) { G.lc_kind = 0; G.lc_j = j; }
        else if (G.fwd ? (j >= plen) : (j <= 0)) { G.lc_kind = 1; G.lc_j = G.fwd ? plen : 0; }
        else { G.lc_kind = 2; G.lc_j = j; }
      }
      const int wj = G.fwd ? WT_GAIN : WT_DROP;
      CPG_SYNCGROUP(W);
      CPG_LOOP for (int q = W.glane; q < 23; q += W.gsize)
        { double val = 0.;
          int need_b = 0, need_s = 0, be = 0, bco = 0, bci = 0, sj = 0, bhc = 1;
          if (G.lc_kind != 0)
            { if (q < 16)
                { const int e = q >> 3, p = q & 7;
                  if (go[e])
                    { if (p == 7) { need_b = 1; be = e; bhc = 1; bco = cout; bci = cin; }
                      else
                        { int j = (p == 0) ? G.lc_j : pg_hc_j(G,K,p-1);
                          int ok = (p == 0) ? (G.lc_kind == 2 && pg_lc_ok(G,W,prof,e) && perr_get(R,j,e,wj) == -CPG_INF)
                                            : (pg_in_range(G,plen,j) && pg_hc_ok(G,W,prof,e,j));
                          if (ok)
                            { uint16_t cin_j, cout_j;
                              pg_counts(G,prof,j,cin_j,cout_j);
                              need_b = 1; be = e; bhc = (p != 0); bco = cout_j; bci = cin_j;
                            }
                        }
                    }
                }
              else if (go[ET_OTHERS])
                { const int p = q-16;
                  int j = (p == 0) ? G.lc_j : pg_hc_j(G,K,p-1);
                  int ok = (p == 0) ? (G.lc_kind == 2 && pg_lc_ok(G,W,prof,ET_OTHERS))
                                    : (pg_in_range(G,plen,j) && pg_hc_ok(G,W,prof,ET_OTHERS,j));
                  if (ok) { need_s = 1; sj = j; }
                }
            }
          if (need_b) val = cpg_p_errorin_lane(lf,be,bhc ? cpg_rate_hc(M) : cpg_rate_pe(M,maxt,maxl),bco,bci,&bad);
          if (need_s) val = G.fwd ? lp_diff_pair(R,W,i,sj) : lp_diff_pair(R,W,sj,i);
          term[q] = val;
        }
      CPG_SYNCGROUP(W);

      if (go[ET_SELF] && pair_replay(R,W,G,ET_SELF,&I) && I.pe >= CPG_PE_FINAL)
        { mark_or(R,W,I.b,MK_BY_S|MK_PAIR_S);
          mark_or(R,W,I.e,MK_BY_S|MK_PAIR_S);
          ei_put(R,W,eidx,I.b,I.e,I.pe);
          eidx++;
        }
      if (go[ET_OTHERS] && pair_replay(R,W,G,ET_OTHERS,&I) && I.pe >= CPG_PE_FINAL)
        { /* paired O-walls stop being walls (src/wall.c:722-726), see header note */
          CPG_SYNCGROUP(W);
          if (W.glane == 0)
            { R.S.mark[I.b] = (uint8_t)((R.S.mark[I.b] | MK_PAIR_O) & ~MK_BY_O);
              R.S.mark[I.e] = (uint8_t)((R.S.mark[I.e] | MK_PAIR_O) & ~MK_BY_O);
            }
          CPG_SYNCGROUP(W);
          reach[ET_OTHERS] = 0;          /* explained by a pair: not a wall */
        }
    }
  if (bad) W.status |= CPG_ST_BINOM;
  /* OTHERS: whatever is not explained by a pair is a wall (src/wall.c:672-690) */
  if (o_wall_now || reach[ET_OTHERS]) mark_or(R,W,i,MK_BY_O);
}

/* ---- pass C for one lone O-wall (src/wall.c:763-860) ---- */
CPG_DEV_NOINL int wall_multi(ReadCtx &R, WCtx &W, int i, int NS, int midx)
{ const int plen = R.plen;
  cpg_eintvl *eint = R.S.eint;
  CPG_LOOP for (int w = WT_DROP; w <= WT_GAIN; w++)
    { double pe_i = perr_get(R,i,ET_SELF,w), pe;
      if (pe_i < CPG_PE_FINAL) continue;
      const int jend = (w == WT_DROP) ? imin(i+200,plen+1) : imax(i-200,0);   /* DROP: j < jend; GAIN: j >= jend */
      int done = 0;
      CPG_LOOP for (int jb = (w == WT_DROP) ? i+1 : i-1; !done && ((w == WT_DROP) ? (jb < jend) : (jb >= jend));
           jb += (w == WT_DROP) ? W.gsize : -W.gsize)
        { int j = (w == WT_DROP) ? jb+W.glane : jb-W.glane;
          int in = (w == WT_DROP) ? (j < jend) : (j >= jend);
          int edge = in && (j == ((w == WT_DROP) ? plen : 0));
          unsigned f = in ? (R.S.mark[j] & (MK_BY_S|MK_BY_O)) : 0u;
          unsigned mask = cpg_gballot(W,f != 0 || edge);
          CPG_LOOP while (mask)
            { int l = cpg_ffs(mask)-1; mask &= mask-1;
              j = (w == WT_DROP) ? jb+l : jb-l;
              if (j == ((w == WT_DROP) ? plen : 0))          /* boundary E-interval */
                { if ((pe = pe_i*pe_i) < CPG_PE_FINAL) continue;
                  if (w == WT_DROP) ei_put(R,W,midx,i,plen,pe); else ei_put(R,W,midx,0,i,pe);
                  mark_or(R,W,i,MK_PAIR_MULT);
                  midx++;
                  if (midx >= plen) { W.status |= CPG_ST_EINTVL_OVF; return midx; }
                  if (W.status & CPG_ST_RETRY) return midx;
                }
              unsigned mj = R.S.mark[j];
              if (!(mj & (MK_BY_S|MK_BY_O))) continue;
              int b = (w == WT_DROP) ? i : j, e = (w == WT_DROP) ? j : i;
              if (ei_find(eint,0,NS-1,b,e) == -1)
                { double pe_j = perr_get(R,j,ET_SELF,(w == WT_DROP) ? WT_GAIN : WT_DROP);
                  if ((pe = pe_i*pe_j) >= CPG_PE_FINAL)
                    { ei_put(R,W,midx,b,e,pe);
                      mark_or(R,W,i,MK_PAIR_MULT);
                      mark_or(R,W,j,MK_PAIR_MULT);
                      midx++;
                      if (midx >= plen) { W.status |= CPG_ST_EINTVL_OVF; return midx; }
                      if (W.status & CPG_ST_RETRY) return midx;
                    }
                }
              if (mj & MK_BY_O) { done = 1; break; }
            }
        }
    }
  return midx;
}

/* ---- src/wall.c:960-1014 ---- */
CPG_DEV_NOINL void correct_wall_cnt(ReadCtx &R, WCtx &W, int idx)
{ const int K = W.M->K;
  const cpg_intvl I = R.S.intvl[idx];
  const uint16_t *prof = R.prof;
  int n_gain = 0, n_drop = 0;

  /* The four sums of src/wall.c:966-996 -- gains over the first K-1 positions and drops over the
     last K-1, each minus the part explained by the low-complexity run at that end -- as two
     loops that work on both ends of the interval at once (twice the loads in flight: the kernel
     waits on DRAM), then one group reduction per sum. */
  { const int e1 = imin(I.b+K-1,I.e-1);              /* gains:  p in [I.b,e1)  */
    const int b3 = imax(I.e-K+1,I.b);                /* drops:  q in [b3,I.e-1) */
    int sg = 0, sd = 0;
    CPG_LOOP for (int o = W.glane; o < K-1; o += W.gsize)
      { const int p = I.b+o, q = b3+o;
        if (p < e1)    sg += imax((int)prof[p+1]-prof[p],0);
        if (q < I.e-1) sd += imax((int)prof[q]-prof[q+1],0);
      }
    n_gain += cpg_gsum(W,sg);
    n_drop += cpg_gsum(W,sd);
  }
  { int e2 = I.b, b4 = I.e-1;                        /* empty ranges unless the interval is longer than K-1 */
    if (I.b+K-1 < I.e)
      { int cl[3], lmax = 0;
        cpg_rctx3(R.seq,R.rlen,I.b+K-1,cl);
        if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) W.status |= CPG_ST_LONG_RUN;
        CPG_LOOP for (int t = 0; t < CT_N; t++) lmax = imax(lmax,cl[t]*(t+1));
        e2 = I.b+lmax;                               /* p in [I.b,e2)  */
      }
    if (I.b < I.e-K+1)
      { int cl[3], lmax = 0;
        cpg_lctx3(R.seq,R.rlen,I.e-K+1+K-2,cl);
        if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) W.status |= CPG_ST_LONG_RUN;
        CPG_LOOP for (int t = 0; t < CT_N; t++) lmax = imax(lmax,cl[t]*(t+1));
        b4 = I.e-lmax;                               /* q in [b4,I.e-1) */
      }
    const int len = imax(e2-I.b,I.e-1-b4);
    if (len > 0)
      { int sg = 0, sd = 0;
        CPG_LOOP for (int o = W.glane; o < len; o += W.gsize)
          { const int p = I.b+o, q = b4+o;
            if (p < e2)    sg += imax((int)prof[p]-rc_prof(R,W,p+1),0);
            if (q < I.e-1) sd += imax((int)prof[q+1]-prof[q],0);
          }
        n_gain -= cpg_gsum(W,sg);
        n_drop -= cpg_gsum(W,sd);
      }
  }
  uint16_t ccb = (uint16_t)imin(I.cb+imax(n_gain,0),CPG_MAX_CNT);
  uint16_t cce = (uint16_t)imin(I.ce+imax(n_drop,0),CPG_MAX_CNT);
  /* src/wall.c:999-1006 index intvl[] with a POSITION that hides the interval index; the only
     write that can land on this interval is the one at position I.b, when I.b == idx:
     ccb = max(ccb,profile[I.b]) is a no-op, cce = max(cce,profile[I.b]) happens iff the scan
     [max(I.e-2K,I.b),I.e) starts at I.b.  Writes to higher slots hit intervals that are either
     recomputed from scratch later or never read. */
  if (I.b == idx && I.e-2*K <= I.b && cce < I.cb) cce = I.cb;
  CPG_SYNCGROUP(W);
  if (W.glane == 0) { R.S.intvl[idx].ccb = ccb; R.S.intvl[idx].cce = cce; }
  CPG_SYNCGROUP(W);
}

/* ---- whole wall stage: fills R.S.intvl[0..N) and R.S.rint[0..M) ---- */
CPG_DEV_NOINL void find_walls_and_reliable(ReadCtx &R, WCtx &W)
{ const cpg_dmodel *M = W.M;
  const int plen = R.plen, K = M->K;
  const uint16_t *prof = R.prof;
  uint8_t *mark = R.S.mark;
  cpg_eintvl *eint = R.S.eint;

  /* flags of positions 0..plen, 16 per store (the array is padded to a multiple of 16) */
  CPG_LOOP for (int i = 16*W.glane; i <= plen; i += 16*W.gsize)
    {
#ifdef CPG_HOSTSIM
      memset(mark+i,0,16);
#else
      *reinterpret_cast<uint4 *>(mark+i) = make_uint4(0u,0u,0u,0u);
#endif
    }
  R.nslots = 0;
  CPG_SYNCGROUP(W);

  /* pass A: candidates in position order, from the decoder's bit map (32 positions per lane) */
  int eidx = 0;
  const int rcov = M->cov[ST_R];
  CPG_LOOP for (int base = 0; base < plen; base += 32*W.gsize)
    { const int p0 = base+32*W.glane;
      unsigned cw = (p0 < plen) ? R.cand[p0 >> 5] : 0u;
      if (p0+32 > plen && p0 < plen) cw &= (1u << (plen-p0))-1u;      /* bits past the profile: none are set, but do not rely on it */
      unsigned lanes = cpg_gballot(W,cw != 0u);
      CPG_LOOP while (lanes)
        { const int l = cpg_ffs(lanes)-1; lanes &= lanes-1;
          unsigned m = cpg_gshfl(W,cw,l);
          CPG_LOOP while (m)
            { const int b = cpg_ffs(m)-1; m &= m-1;
              wall_candidate(R,W,base+32*l+b,eidx);
              if (W.status & CPG_ST_RETRY) { R.N = 0; R.M = 0; return; }
            }
        }
    }
  int NS = eidx;

  /* pass B (src/wall.c:727-735) */
  CPG_LOOP for (int k = 0; k < NS; k++) mark_clear_range(R,W,eint[k].b,eint[k].e,MK_BY_O);
  NS = ei_unique(eint,eidx,W);

  /* pass C: lone O-walls, 16 flag bytes per lane */
  int midx = NS;
  CPG_LOOP for (int base = 0; base < plen && !(W.status & CPG_ST_ABORT); base += 16*W.gsize)
    { const int p0 = base+16*W.glane;
      unsigned hit = 0;
      if (p0 < plen)
        { unsigned w[4];
          cpg_ld16(mark+p0,w);
          CPG_LOOP for (int k = 0; k < 4; k++) hit |= cpg_pack4((w[k] >> 4) & ~w[k]) << (4*k);     /* BY_O and not BY_S */
          if (p0 == 0) hit &= ~1u;                                    /* positions 1..plen-1 */
          if (p0+16 > plen) hit &= (1u << (plen-p0))-1u;
        }
      unsigned lanes = cpg_gballot(W,hit != 0u);
      CPG_LOOP while (lanes)
        { const int l = cpg_ffs(lanes)-1; lanes &= lanes-1;
          unsigned m = cpg_gshfl(W,hit,l);
          CPG_LOOP while (m)
            { const int b = cpg_ffs(m)-1; m &= m-1;
              const int i = base+16*l+b;
              if (mark[i] & MK_PAIR_MULT) continue;
              midx = wall_multi(R,W,i,NS,midx);
              if (W.status & CPG_ST_ABORT) { lanes = 0; break; }
            }
        }
    }
  if (W.status & CPG_ST_ABORT) { R.N = 0; R.M = 0; return; }
  CPG_LOOP for (int k = NS; k < midx; k++) mark_clear_range(R,W,eint[k].b,eint[k].e,MK_BY_O);
  if (NS < midx) { NS = midx; ei_sort(eint,NS,W); }

  /* pass D (src/wall.c:877-909): hulls of chains of overlapping E-intervals are appended while
     the list is being walked, and the loop bound is re-read */
  { int i = 0;
    CPG_LOOP while (i < NS-1)
      { int max_e = eint[i].e; double max_pe = eint[i].pe;
        int j = i;
        CPG_LOOP while (j < NS-1 && eint[j+1].b <= eint[j].e)
          { max_e = imax(max_e,eint[j+1].e);
            max_pe = dmax_ref(max_pe,eint[j+1].pe);
            j++;
          }
        if (i < j)
          { ei_put(R,W,NS,eint[i].b,max_e,max_pe);
            NS++;
            if (NS >= plen) { W.status |= CPG_ST_EINTVL_OVF; R.N = 0; R.M = 0; return; }
            if (W.status & CPG_ST_RETRY) { R.N = 0; R.M = 0; return; }
          }
        i = j+1;
      }
  }
  ei_sort(eint,NS,W);
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int k = 0; k < NS; k++)
    { CPG_LOOP for (int j = eint[k].b+W.glane; j < eint[k].e; j += W.gsize) mark[j] |= (uint8_t)MK_ERROR;
      CPG_SYNCGROUP(W);
    }

  /* pass E (src/wall.c:921-948) */
  int N = 0, b = 0;
  cpg_intvl *intvl = R.S.intvl;
  CPG_LOOP for (int base = 0; base <= plen; base += 16*W.gsize)
    { const int p0 = base+16*W.glane;
      unsigned cut = 0;
      if (p0 <= plen)
        { unsigned w[4];
          cpg_ld16(mark+p0,w);
          unsigned carry = (p0 > 0) ? ((unsigned)mark[p0-1] >> 7) : 0u;       /* error bit of the position before */
          CPG_LOOP for (int k = 0; k < 4; k++)
            { const unsigned er = (w[k] >> 7) & 0x01010101u, ow = (w[k] >> 4) & 0x01010101u;
              const unsigned pv = (er << 8) | carry;
              cut |= cpg_pack4((er ^ pv) | (~er & ow)) << (4*k);
              carry = er >> 24;
            }
          if (p0 == 0) cut &= ~1u;                                    /* positions 1..plen */
          if (p0+16 > plen) { cut &= (2u << (plen-p0))-1u; cut |= 1u << (plen-p0); }
        }
      unsigned lanes = cpg_gballot(W,cut != 0u);
      CPG_LOOP while (lanes)
        { const int ll = cpg_ffs(lanes)-1; lanes &= lanes-1;
          unsigned mm = cpg_gshfl(W,cut,ll);
          CPG_LOOP while (mm)
            { const int bb = cpg_ffs(mm)-1; mm &= mm-1;
              const int e = base+16*ll+bb;
              int k = ei_find(eint,0,NS-1,b,e);
              double pe  = (k != -1) ? cpg_log(eint[k].pe) : -CPG_INF;
              double pob = dmax_ref(perr_get(R,b,ET_OTHERS,WT_DROP),perr_get(R,b,ET_OTHERS,WT_GAIN));
              double poe = dmax_ref(perr_get(R,e,ET_OTHERS,WT_DROP),perr_get(R,e,ET_OTHERS,WT_GAIN));
              double lpob = (pob != -CPG_INF) ? cpg_log(pob) : -CPG_INF;
              double lpoe = (poe != -CPG_INF) ? cpg_log(poe) : -CPG_INF;
              if (W.glane == 0 && N < R.S.capI)
                { cpg_intvl *I = intvl+N;
                  I->b = b; I->e = e; I->cb = prof[b]; I->ce = prof[e-1];
                  I->ccb = 0; I->cce = 0; I->is_rel = 0; I->asgn = ST_N;
                  I->pe = pe; I->peob = lpob; I->peoe = lpoe;
                }
              N++;
              b = e;
            }
        }
    }
  CPG_SYNCGROUP(W);
  if (N > R.S.capI) { W.status |= CPG_ST_RETRY; R.N = 0; R.M = 0; return; }
  R.N = N;

  /* reliable intervals (src/wall.c:1016-1037).  Three phases: corrected end counts of every
     interval that passes the cheap filters (warp-cooperative sums), then the Skellam plausibility
     test of all of them at once (one interval per lane), then the copies in interval order. */
  int ncand = 0;
  const double logpthres = cpg_log(CPG_PE_FINAL);
  int32_t *cand = R.S.ord;
  uint8_t *keep = R.S.fixed;
  CPG_LOOP for (int i = 0; i < N; i++)
    { const cpg_intvl I = intvl[i];
      if (I.e-I.b < K) continue;
      if (imax(I.cb,I.ce) >= rcov) continue;
      if (I.pe >= logpthres) continue;
      correct_wall_cnt(R,W,i);
      if (W.glane == 0) cand[ncand] = i;
      ncand++;
    }
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int q = W.glane; q < ncand; q += W.gsize)
    { const cpg_intvl I = intvl[cand[q]];
      const int ccb = I.ccb, cce = I.cce;
      int ok = !(cpg_lp_trans(W,I.b,I.e,ccb,cce,(uint16_t)((ccb+cce)/2)) < CPG_THRES_DIFF_REL);
      if (imax(ccb,cce) == CPG_MAX_CNT) ok = 0;
      keep[q] = (uint8_t)ok;
    }
  CPG_SYNCGROUP(W);
  int Mrel = 0;
  CPG_LOOP for (int q = 0; q < ncand; q++) if (keep[q]) Mrel++;
  if (W.glane == 0)
    { int m = 0;
      CPG_LOOP for (int q = 0; q < ncand; q++)
        if (keep[q])
          { const int i = cand[q];
            intvl[i].is_rel = 1;
            R.S.rint[m++] = intvl[i];
          }
    }
  CPG_SYNCGROUP(W);
  R.M = Mrel;
}

#endif
