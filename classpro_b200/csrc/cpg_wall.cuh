/*******************************************************************************************
 *  cpg_wall.cuh -- wall detection and reliable-interval selection for one read, one warp.
 *
 *  Replaces find_wall (+ find_gain/find_drop/find_pair, update_perror, remove_duplicates,
 *  bs_eintvl) src/wall.c:264-958 and find_rel_intvl/correct_wall_cnt src/wall.c:960-1051.
 *
 *  The stage is cut where its data dependences are (results identical to the reference):
 *   wa_  PURE, one wall candidate per thread: wall type, context, the count-threshold tests of
 *        src/wall.c:643-675 and -- for the one candidate in ten that gets past them -- every
 *        probability pass A can ask for (the candidate's own binomial tails, those of its low- and
 *        high-complexity partners, the Skellam terms of the OTHERS pairs).  All of them are functions
 *        of the counts, the sequence and the model only.  Output: a 16-byte header per candidate and
 *        a 216-byte record per candidate that needs one (cpg_common.h).
 *   wb_  the ORDER-DEPENDENT rest of find_wall, one read per lane group, replayed on those records in
 *        position order: the first-writer-wins probability cache, the paired flags, the E-interval
 *        list (src/wall.c:310-315,639-640), lone O-walls, hulls, and the interval cuts.  It never
 *        looks at a count or a base.  Its state is sparse: one flag byte per profile position that
 *        is zero between reads (the replay logs the positions it touches and zeroes exactly those),
 *        probability slots allocated on first store.  Nothing in it sweeps the profile: O-walls can
 *        only stand at candidates, so "in position order" means "down the header list", and the
 *        error flag of a position (the reference sets one per position, src/wall.c:911-919) is
 *        membership in the merged E-interval list.
 *   wc_  PURE again, one interval per thread: end counts, the corrected counts of
 *        correct_wall_cnt and the Skellam plausibility test of find_rel_intvl; then the reliable
 *        intervals are copied out in order.
 *  k_wall_a / k_wall_b / k_wall_c run the three steps as three launches over the whole batch;
 *  find_walls_and_reliable() below runs them back to back for one read (retry launch, host tests).
 *  Further layout differences from the reference:
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
CPG_DEV int      cpg_popc(unsigned m) { return __builtin_popcount(m); }
#elif defined(CPG_HOSTSIM)
CPG_DEV unsigned cpg_gballot(const WCtx &W, int pred) { (void)W; return pred ? 1u : 0u; }
CPG_DEV int      cpg_gsum(const WCtx &W, int v)       { (void)W; return v; }
CPG_DEV unsigned cpg_gshfl(const WCtx &W, unsigned v, int l) { (void)W; (void)l; return v; }
CPG_DEV int      cpg_ffs(unsigned m)  { return __builtin_ffs((int)m); }
CPG_DEV int      cpg_popc(unsigned m) { return __builtin_popcount(m); }
#else
CPG_DEV unsigned cpg_gballot(const WCtx &W, int pred) { return __ballot_sync(W.gmask,pred) >> W.gbase; }
CPG_DEV int      cpg_gsum(const WCtx &W, int v)       { return __reduce_add_sync(W.gmask,v); }
CPG_DEV unsigned cpg_gshfl(const WCtx &W, unsigned v, int l) { return __shfl_sync(W.gmask,v,W.gbase+l); }
CPG_DEV int      cpg_ffs(unsigned m)  { return __ffs((int)m); }
CPG_DEV int      cpg_popc(unsigned m) { return __popc(m); }
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
#define MK_STALE_PROF 128     /* status bit: profile[plen] was read (reference reads stale memory) */

struct ReadCtx
  { const uint16_t *prof;
    int             plen, rlen;
    cpg_seq         seq;
    const uint32_t *cand;     /* wall-candidate bit map of the read (bit i of word i>>5 = position i) */
    cpg_scratch     S;
    int             nslots;
    int             N, M;
    const cpg_chdr *hdr;      /* candidate headers of the read, in position order */
    const cpg_cbig *big;      /* base of the big records cpg_chdr.big indexes */
    int             ncand;
    int             ntlog;    /* entries of S.tlog; > S.capT: the log overflowed, clean the whole flag array */
    int             prune;    /* single-read path: record partner values as k_wall_a does (wa_tasks; host tests) */
  };

CPG_DEV unsigned mk_by(int e)   { return e == ET_SELF ? MK_BY_S : MK_BY_O; }
CPG_DEV unsigned mk_pair(int e) { return e == ET_SELF ? MK_PAIR_S : MK_PAIR_O; }

/* Lane 0 is the only writer of the scratch words below.  Each helper synchronises the warp BEFORE
 * the write (the other lanes may still be reading the old value: the CUDA memory model does not
 * promise lock-step execution) and AFTER it (so that every lane sees the new one). */
/* a flag byte leaves zero: remember the position (group-uniform; lane 0 writes) */
CPG_DEV_HELPER void mark_touch(ReadCtx &R, const WCtx &W, int pos)
{ if (R.ntlog < R.S.capT) { if (W.glane == 0) R.S.tlog[R.ntlog] = pos; }
  R.ntlog++;
}
CPG_DEV_HELPER void mark_or(ReadCtx &R, const WCtx &W, int pos, unsigned bits)
{ const unsigned m = R.S.mark[pos];
  CPG_SYNCGROUP(W);
  if (m == 0u) mark_touch(R,W,pos);
  if (W.glane == 0) R.S.mark[pos] = (uint8_t)(m | bits);
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

/* max of the two OTHERS probabilities of a position (src/wall.c:941-944), -inf where it has no slot */
CPG_DEV_HELPER double perr_max_o(const ReadCtx &R, int pos)
{ if (!(R.S.mark[pos] & MK_SLOT)) return -CPG_INF;
  const double *q = R.S.perr+(size_t)R.S.slot[pos]*4+ET_OTHERS*2;
  return dmax_ref(q[WT_DROP],q[WT_GAIN]);
}

/* src/wall.c:317-322 */
CPG_DEV_HELPER double lp_diff_pair(const uint16_t *p, const WCtx &W, int i, int j)
{ int n_drop = (int)p[i-1]-p[i], n_gain = (int)p[j]-p[j-1];
  uint16_t cov = (uint16_t)imax(p[i-1],p[j]);
  return cpg_lp_trans_thr(W,i,j,n_drop,n_gain,cov,CPG_THRES_DIFF_EO);       /* only ever compared with that threshold */
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
      if (m == 0u) mark_touch(R,W,pos);
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

/* Layout of cpg_cbig.term (the per-candidate results of the pure step):
 *   [e*8+0]      p_errorin of the low-complexity partner under this candidate's error rate
 *   [e*8+1+n]    p_errorin of high-complexity partner n under HC_ERATE
 *   [e*8+7]      p_errorin of the candidate itself under HC_ERATE
 *   [16+p]       log Skellam probability that candidate and partner p belong together (OTHERS) */
CPG_DEV int big_lc_ok(const cpg_cbig *B, int e)        { return (B->ok >> (e*7)) & 1; }
CPG_DEV int big_hc_ok(const cpg_cbig *B, int e, int n) { return (B->ok >> (e*7+1+n)) & 1; }

/* ==========================================================================================
 *  wa_: the pure step, one candidate per thread (no group collectives in here)
 * ========================================================================================== */
struct WaCand
  { int      wtype, t, l;
    uint16_t cout, cin;
    int      cng;
    double   erate;
  };

/* src/wall.c:606-648,651-652,672-675 without the paired flags: which error types get past the count
   thresholds, and whether OTHERS makes the position a wall outright.  Returns the CH_* bits. */
CPG_DEV_NOINL unsigned wa_stage0(const uint16_t *prof, const cpg_seq seq, int rlen, const WCtx &W, int i, WaCand &C)
{ const cpg_dmodel *M = W.M;
  const int K = M->K, cmax = M->cmax;
  const uint16_t cim1 = prof[i-1], ci = prof[i];
  unsigned info = 0;
  if (cim1 > ci) { C.wtype = WT_DROP; C.cin = ci;   C.cout = cim1; }
  else           { C.wtype = WT_GAIN; C.cin = cim1; C.cout = ci; info |= CH_GAIN; }
  C.cng = (int)C.cout-(int)C.cin;
  int maxt = -1, maxl = -1; double maxpe = -CPG_INF;
  { int cl[3];
    cpg_ctx3_at(seq,rlen,K,C.wtype,i,cl);
    if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) info |= CH_LONG;
    CPG_LOOP for (int t = 0; t < CT_N; t++)
      { int l = imin(cl[t],M->lmax[t]);
        double pe = M->pe[t][l];
        if (maxpe < pe) { maxpe = pe; maxt = t; maxl = l; }
      }
  }
  C.t = maxt; C.l = maxl; C.erate = maxpe;
  const int cout = C.cout, cin = C.cin, cng = C.cng;
  CPG_LOOP for (int e = ET_SELF; e <= ET_OTHERS; e++)
    { int ct_final = 0;
      if (cout < cmax)
        { int ct_init = cthres_at(W,maxt,maxl,cout,TH_INIT,e);
          ct_final = cthres_at(W,maxt,maxl,cout,TH_FINAL,e);
          if (!(cng > CPG_MAX_CNT_CHANGE || cin < imax(ct_init,3))) continue;
        }
      if (e == ET_SELF)
        { if (cout < cmax && cin >= ct_final) continue;
          info |= CH_REACH_S;
        }
      else
        { if (cng >= M->cov[ST_H] || (cout < cmax && cin < ct_final)) { info |= CH_ONOW; continue; }
          info |= CH_REACH_O;
        }
    }
  return info;
}

/* Everything pass A can ask for about a candidate that got past stage 0 (src/wall.c:331-507,651-690):
   the reference evaluates these one after the other and only as far as the order-dependent state lets
   it get; here all of them are evaluated (they are pure), and the replay picks what it needs.
   prune: an error type whose own probability is already below the threshold gets no partner values (pass A
   stops there too, src/wall.c:653,676 -- unless the probability it finds in the cache is another one, stored
   by an earlier candidate under its error rate: the replay notices that case and sends the read to the retry
   launch, which records everything).  It is what keeps the high-count candidates of repeat-rich profiles
   cheap: above CMAX there is no count threshold, every candidate gets here, and almost none is an error. */
CPG_DEV_NOINL void wa_tasks(const uint16_t *prof, int plen, const cpg_seq seq, int rlen, const WCtx &W, int i,
                            const WaCand &C, unsigned info, cpg_cbig *B, int prune)
{ const cpg_dmodel *M = W.M;
  const double *lf = M->logfact;
  const int K = M->K;
  const int reach[2] = { (int)(info & CH_REACH_S) != 0, (int)(info & CH_REACH_O) != 0 };
  int bad = 0, lr = 0;
  double term[23];
  CPG_LOOP for (int q = 0; q < 23; q++) term[q] = 0.;
  B->own[0] = B->own[1] = 0.;
  int want[2];
  unsigned ok = 0;
  CPG_LOOP for (int e = 0; e < 2; e++)
    { if (reach[e]) B->own[e] = cpg_p_errorin_lane(lf,e,cpg_rate_pe(M,C.t,C.l),C.cout,C.cin,&bad);
      want[e] = reach[e] && !(prune && B->own[e] < CPG_PE_FINAL);
      if (reach[e] && !want[e]) ok |= 0x4000u << e;                         /* no partner values for this type */
    }

  /* partner geometry (src/wall.c:344-357,432-450), shared by both error types */
  PairGeom G;
  G.fwd = (C.wtype == WT_DROP); G.i = i; G.t = C.t; G.l = C.l; G.cout = C.cout; G.cin = C.cin; G.erate = C.erate;
  G.lc_kind = 0; G.lc_j = 0;
  if (want[0] || want[1])
  { const int ulen = C.t+1, m = ulen*C.l;
    int n = 0;
    CPG_LOOP for (;;)
      { int idx = G.fwd ? i+ulen*(n+1) : i-ulen*(n+1);
        if (G.fwd) { if (idx >= plen) break; }
        else       { if (idx <= 0) break; }
        const int cx = cpg_ctx_at(seq,rlen,K,C.wtype,idx,C.t);
        if (cx >= 127) lr = 1;
        if (cx != m+n+1) break;
        n++;
      }
    int j = G.fwd ? i+K-1+n-m : i-K+1-n+m;
    if (G.fwd ? (j <= i) : (j >= i)) { G.lc_kind = 0; G.lc_j = j; }
    else if (G.fwd ? (j >= plen) : (j <= 0)) { G.lc_kind = 1; G.lc_j = G.fwd ? plen : 0; }
    else { G.lc_kind = 2; G.lc_j = j; }
  }
  int nhc = 0;
  if (G.lc_kind != 0 && (want[0] || want[1]))
    { CPG_LOOP for (int n = 0; n <= CPG_MAX_N_HC; n++) { if (!pg_in_range(G,plen,pg_hc_j(G,K,n))) break; nhc++; }
      CPG_LOOP for (int e = 0; e < 2; e++)
        { if (!want[e]) continue;
          term[e*8+7] = cpg_p_errorin_lane(lf,e,cpg_rate_hc(M),C.cout,C.cin,&bad);
          if (G.lc_kind == 2 && pg_lc_ok(G,W,prof,e))
            { uint16_t cin_j, cout_j;
              pg_counts(G,prof,G.lc_j,cin_j,cout_j);
              ok |= 1u << (e*7);
              term[e*8] = cpg_p_errorin_lane(lf,e,cpg_rate_pe(M,C.t,C.l),cout_j,cin_j,&bad);
              if (e == ET_OTHERS) term[16] = G.fwd ? lp_diff_pair(prof,W,i,G.lc_j) : lp_diff_pair(prof,W,G.lc_j,i);
            }
          CPG_LOOP for (int n = 0; n < nhc; n++)
            { const int j = pg_hc_j(G,K,n);
              if (!pg_hc_ok(G,W,prof,e,j)) continue;
              uint16_t cin_j, cout_j;
              pg_counts(G,prof,j,cin_j,cout_j);
              ok |= 1u << (e*7+1+n);
              term[e*8+1+n] = cpg_p_errorin_lane(lf,e,cpg_rate_hc(M),cout_j,cin_j,&bad);
              if (e == ET_OTHERS) term[17+n] = G.fwd ? lp_diff_pair(prof,W,i,j) : lp_diff_pair(prof,W,j,i);
            }
        }
    }
  CPG_LOOP for (int q = 0; q < 23; q++) B->term[q] = term[q];
  B->lc_j = G.lc_j; B->lc_kind = (uint8_t)G.lc_kind; B->nhc = (uint8_t)nhc;
  B->bad = (uint8_t)bad; B->lr_walk = (uint8_t)lr; B->ok = (uint16_t)ok;
  B->pad[0] = B->pad[1] = B->pad[2] = 0;
}

/* one candidate: header, and the big record at big_base[big_idx] if it needs one (the caller has a
   place ready: it is simply left unused otherwise) */
CPG_DEV void wa_candidate(const uint16_t *prof, int plen, const cpg_seq seq, int rlen, const WCtx &W, int i,
                          cpg_chdr *H, cpg_cbig *big_base, uint32_t big_idx, int prune)
{ WaCand C;
  const unsigned info = wa_stage0(prof,seq,rlen,W,i,C);
  H->pos = i; H->info = info; H->big = big_idx; H->pad = 0;
  if (info & (CH_REACH_S|CH_REACH_O)) wa_tasks(prof,plen,seq,rlen,W,i,C,info,big_base+big_idx,prune);
}

/* ==========================================================================================
 *  wb_: the order-dependent replay of one read (group-uniform control flow, lane 0 writes)
 * ========================================================================================== */

/* The decision part of find_gain/find_drop (src/wall.c:359-411,452-502) on the recorded values. */
CPG_DEV_NOINL int pair_replay(ReadCtx &R, WCtx &W, const cpg_cbig *B, int fwd, int i, int e, cpg_eintvl *out)
{ const int K = W.M->K;
  const int wi = fwd ? WT_DROP : WT_GAIN, wj = fwd ? WT_GAIN : WT_DROP;
  const double *term = B->term;
  int max_j = -1; double max_pe = -CPG_INF, pe;
  if (B->lc_kind == 0) return 0;
  int j = B->lc_j;
  if (B->lc_kind == 1)
    { double pi = perr_get(R,i,e,wi);
      pe = pi*pi;
    }
  else
    { pe = -CPG_INF;
      if (big_lc_ok(B,e) && (e == ET_SELF || term[16] >= CPG_THRES_DIFF_EO))
        { if (perr_get(R,j,e,wj) == -CPG_INF) perr_store(R,W,j,e,wj,term[e*8]);
          pe = fwd ? perr_get(R,i,e,WT_DROP)*perr_get(R,j,e,WT_GAIN)
                   : perr_get(R,j,e,WT_DROP)*perr_get(R,i,e,WT_GAIN);
        }
    }
  if (max_pe < pe) { max_j = j; max_pe = pe; }
  CPG_LOOP for (int n = 0; n < (int)B->nhc; n++)
    { j = fwd ? i+K-1+n : i-K+1-n;
      if (!big_hc_ok(B,e,n)) continue;
      if (e == ET_OTHERS && term[17+n] < CPG_THRES_DIFF_EO) continue;
      pe = term[e*8+7]*term[e*8+1+n];
      if (max_pe < pe) { max_j = j; max_pe = pe; }
    }
  if (max_j == -1) return 0;
  if (fwd) { out->b = i; out->e = max_j; }
  else     { out->b = max_j; out->e = i; }
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

/* first candidate of the read at a position > p (the headers are in position order) */
CPG_DEV_HELPER int hdr_after(const ReadCtx &R, int p)
{ int lo = 0, hi = R.ncand;
  CPG_LOOP while (lo < hi)
    { const int m = (lo+hi) >> 1;
      if (R.hdr[m].pos > p) hi = m; else lo = m+1;
    }
  return lo;
}

/* Down the header list from c: the first candidate whose flag byte has all of `want` and none of `not`,
   or the first at a position >= lim; R.ncand if neither.  Four headers, then their four flag bytes, are
   loaded together (the list is walked for every cut: one dependent pair of loads per candidate otherwise). */
CPG_DEV_HELPER int hdr_scan(const ReadCtx &R, int c, int lim, unsigned want, unsigned nots)
{ CPG_LOOP while (c < R.ncand)
    { int p[4]; unsigned m[4];
      CPG_UNROLL4 for (int k = 0; k < 4; k++) p[k] = (c+k < R.ncand) ? R.hdr[c+k].pos : 0x7fffffff;
      CPG_UNROLL4 for (int k = 0; k < 4; k++) m[k] = (p[k] < lim) ? R.S.mark[p[k]] : 0u;
      CPG_UNROLL4 for (int k = 0; k < 4; k++)
        { if (p[k] >= lim) return (c+k < R.ncand) ? c+k : R.ncand;
          if ((m[k] & want) == want && !(m[k] & nots)) return c+k;
        }
      c += 4;
    }
  return R.ncand;
}

/* O-walls strictly inside an E-interval stop being walls (src/wall.c:727-735,865-873).  An O-wall can
   only stand at a candidate, so the open range (b,e) is walked down the header list. */
CPG_DEV_HELPER void clear_o_range(ReadCtx &R, const WCtx &W, int b, int e)
{ CPG_LOOP for (int c = hdr_after(R,b); c < R.ncand; c++)
    { const int p = R.hdr[c].pos;
      if (p >= e) break;
      if (R.S.mark[p] & MK_BY_O) mark_clear(R,W,p,MK_BY_O);
    }
}

/* ---- pass A for one candidate position (src/wall.c:606-692), replayed on its records ---- */
CPG_DEV_NOINL void wb_candidate(ReadCtx &R, WCtx &W, const cpg_chdr H, int &eidx)
{ const int i = H.pos;
  const int wtype = (H.info & CH_GAIN) ? WT_GAIN : WT_DROP;
  if (H.info & CH_LONG) W.status |= CPG_ST_LONG_RUN;
  const unsigned mi = R.S.mark[i];
  /* a position that is already one end of a pair of error type e is not looked at again for e
     (src/wall.c:639-640) */
  int reach[2];
  reach[ET_SELF]   = (H.info & CH_REACH_S) && !(mi & MK_PAIR_S);
  reach[ET_OTHERS] = (H.info & CH_REACH_O) && !(mi & MK_PAIR_O);
  const int o_wall_now = (H.info & CH_ONOW) && !(mi & MK_PAIR_O);
  if (!reach[0] && !reach[1])
    { if (o_wall_now) mark_or(R,W,i,MK_BY_O);
      return;
    }
  const cpg_cbig *B = R.big+H.big;

  /* the candidate's own probabilities: first writer wins (src/wall.c:310-315) */
  int go[2];
  CPG_LOOP for (int e = 0; e < 2; e++)
    { if (reach[e] && perr_get(R,i,e,wtype) == -CPG_INF) perr_store(R,W,i,e,wtype,B->own[e]);
      go[e] = reach[e] && !(perr_get(R,i,e,wtype) < CPG_PE_FINAL);
    }
  if (B->bad) W.status |= CPG_ST_BINOM;

  cpg_eintvl I;
  if ((go[0] && (B->ok & 0x4000u)) || (go[1] && (B->ok & 0x8000u)))
    { W.status |= CPG_ST_RETRY; return; }                 /* values not recorded (see wa_tasks, prune) */
  if (go[0] || go[1])
    { const int fwd = (wtype == WT_DROP);
      if (B->lr_walk) W.status |= CPG_ST_LONG_RUN;
      if (go[ET_SELF] && pair_replay(R,W,B,fwd,i,ET_SELF,&I) && I.pe >= CPG_PE_FINAL)
        { mark_or(R,W,I.b,MK_BY_S|MK_PAIR_S);
          mark_or(R,W,I.e,MK_BY_S|MK_PAIR_S);
          ei_put(R,W,eidx,I.b,I.e,I.pe);
          eidx++;
        }
      if (go[ET_OTHERS] && pair_replay(R,W,B,fwd,i,ET_OTHERS,&I) && I.pe >= CPG_PE_FINAL)
        { /* paired O-walls stop being walls (src/wall.c:722-726), see header note */
          const unsigned mb = R.S.mark[I.b], me = R.S.mark[I.e];
          CPG_SYNCGROUP(W);
          if (mb == 0u) mark_touch(R,W,I.b);
          if (me == 0u) mark_touch(R,W,I.e);
          if (W.glane == 0)
            { R.S.mark[I.b] = (uint8_t)((mb | MK_PAIR_O) & ~MK_BY_O);
              R.S.mark[I.e] = (uint8_t)((me | MK_PAIR_O) & ~MK_BY_O);
            }
          CPG_SYNCGROUP(W);
          reach[ET_OTHERS] = 0;          /* explained by a pair: not a wall */
        }
    }
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

/* Passes A-D of find_wall for one read (src/wall.c:590-909).  Leaves the final E-interval list sorted in
   R.S.eint and returns its length (the flags and probability slots stay in place for wb_cuts). */
CPG_DEV_NOINL int wb_walls(ReadCtx &R, WCtx &W)
{ const int plen = R.plen;
  cpg_eintvl *eint = R.S.eint;
  R.nslots = 0; R.ntlog = 0;

  /* pass A: candidates in position order */
  int eidx = 0;
  CPG_LOOP for (int c = 0; c < R.ncand; c++)
    { wb_candidate(R,W,R.hdr[c],eidx);
      if (W.status & CPG_ST_RETRY) return 0;
    }
  int NS = eidx;

  /* pass B (src/wall.c:727-735) */
  CPG_LOOP for (int k = 0; k < NS; k++) clear_o_range(R,W,eint[k].b,eint[k].e);
  NS = ei_unique(eint,eidx,W);

  /* pass C: lone O-walls (by OTHERS, not by SELF), positions 1..plen-1; they stand at candidates */
  int midx = NS;
  CPG_LOOP for (int c = hdr_scan(R,0,plen,MK_BY_O,MK_BY_S|MK_PAIR_MULT); c < R.ncand; c = hdr_scan(R,c+1,plen,MK_BY_O,MK_BY_S|MK_PAIR_MULT))
    { midx = wall_multi(R,W,R.hdr[c].pos,NS,midx);
      if (W.status & CPG_ST_ABORT) return 0;
    }
  CPG_LOOP for (int k = NS; k < midx; k++) clear_o_range(R,W,eint[k].b,eint[k].e);
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
            if (NS >= plen) { W.status |= CPG_ST_EINTVL_OVF; return 0; }
            if (W.status & CPG_ST_RETRY) return 0;
          }
        i = j+1;
      }
  }
  ei_sort(eint,NS,W);
  return NS;
}

/* a fresh interval (counts are filled in by wc_interval), as three 16-byte stores */
CPG_DEV void intvl_put(cpg_intvl *dst, int b, int e, double pe, double peob, double peoe)
{
#ifdef CPG_HOSTSIM
  cpg_intvl I; memset(&I,0,sizeof(I));
  I.b = b; I.e = e; I.asgn = ST_N; I.pe = pe; I.peob = peob; I.peoe = peoe;
  *dst = I;
#else
  uint4 *q = reinterpret_cast<uint4 *>(dst);
  q[0] = make_uint4((unsigned)b,(unsigned)e,0u,0u);                               /* b, e, cb, ce, ccb, cce */
  const unsigned long long upe = (unsigned long long)__double_as_longlong(pe);
  q[1] = make_uint4((unsigned)ST_N << 8,0u,(unsigned)upe,(unsigned)(upe >> 32));  /* is_rel, asgn, pad, pe */
  const unsigned long long u1 = (unsigned long long)__double_as_longlong(peob), u2 = (unsigned long long)__double_as_longlong(peoe);
  q[2] = make_uint4((unsigned)u1,(unsigned)(u1 >> 32),(unsigned)u2,(unsigned)(u2 >> 32));
#endif
}

/* Pass E (src/wall.c:911-948).  The reference flags every position covered by an E-interval and cuts
   where the flag toggles, or at an O-wall outside the flagged stretches, or at plen.  Here the flagged
   stretches are the runs of the sorted E-interval list (intervals that touch or overlap merge) and the
   O-walls come down the header list, so the cuts are a merge of two sorted streams.
   dst == NULL: count only.  Returns the number of intervals; *mcap = how many are at least K long
   (an upper bound on the reliable ones). */
CPG_DEV_NOINL int wb_cuts(ReadCtx &R, WCtx &W, int NS, cpg_intvl *dst, int cap, int *mcap)
{ const int plen = R.plen, K = W.M->K;
  const cpg_eintvl *eint = R.S.eint;
  int N = 0, b = 0, c = 0, k = 0, nlong = 0;
  double lpob = -CPG_INF;                 /* log of the OTHERS probability at the current interval start */
  { const double pob = perr_max_o(R,0);
    if (dst && pob != -CPG_INF) lpob = cpg_log(pob);
  }
  CPG_LOOP for (;;)
    { int rb, re;
      if (k < NS)
        { rb = eint[k].b; re = eint[k].e; k++;
          CPG_LOOP while (k < NS && eint[k].b <= re) { re = imax(re,eint[k].e); k++; }
        }
      else { rb = plen; re = plen; }
      /* three kinds of cut before the next run is over: O-walls in front of it, its start, its end */
      CPG_LOOP for (int step = 0; ; )
        { int e, in_run = 0;
          if (step == 0)
            { /* next O-wall in front of the run */
              c = hdr_scan(R,c,rb,MK_BY_O,0u);
              const int pc = (c < R.ncand) ? R.hdr[c].pos : rb;
              if (pc < rb) { e = pc; c++; }
              else { step = 1; continue; }
            }
          else if (step == 1)
            { step = 2;
              if (rb > b && rb < plen) e = rb; else continue;      /* rb == 0: no cut at position 0; rb == b cannot happen twice */
            }
          else if (step == 2)
            { step = 3;
              if (rb >= plen) { if (b < plen) e = plen; else break; }
              else { e = re; in_run = 1; }
            }
          else break;
          /* interval [b,e) */
          if (e-b >= K) nlong++;
          if (dst)
            { double pe = -CPG_INF;
              if (in_run)
                { const int q = ei_find(eint,0,NS-1,b,e);
                  if (q != -1) pe = cpg_log(eint[q].pe);
                }
              const double poe = perr_max_o(R,e);
              const double lpoe = (poe != -CPG_INF) ? cpg_log(poe) : -CPG_INF;
              if (W.glane == 0 && N < cap) intvl_put(dst+N,b,e,pe,lpob,lpoe);
              lpob = lpoe;
            }
          N++;
          b = e;
          if (in_run)
            { /* candidates inside the run, or at its end, cut nothing more */
              CPG_LOOP while (c < R.ncand && R.hdr[c].pos <= re) c++;
            }
        }
      if (rb >= plen) break;
    }
  CPG_SYNCGROUP(W);
  if (mcap) *mcap = nlong;
  return N;
}

/* the flag bytes go back to zero: exactly the touched ones (or all of them if the log overflowed) */
CPG_DEV_NOINL void wb_clean(ReadCtx &R, const WCtx &W)
{ CPG_SYNCGROUP(W);
  if (R.ntlog <= R.S.capT)
    { CPG_LOOP for (int q = W.glane; q < R.ntlog; q += W.gsize) R.S.mark[R.S.tlog[q]] = 0; }
  else
    { CPG_LOOP for (int q = W.glane; q <= R.plen; q += W.gsize) R.S.mark[q] = 0; }
  R.ntlog = 0;
  CPG_SYNCGROUP(W);
}

/* ==========================================================================================
 *  wc_: pure again, one interval per thread: end counts, corrected end counts, reliability
 * ========================================================================================== */

/* ---- src/wall.c:960-1014 ---- */
CPG_DEV_NOINL void wc_correct(const uint16_t *prof, int plen, const cpg_seq seq, int rlen, WCtx &W, cpg_intvl &I, int idx)
{ const int K = W.M->K;
  int n_gain = 0, n_drop = 0;
  /* gains over the first K-1 positions and drops over the last K-1 (src/wall.c:966-996) */
  { const int e1 = imin(I.b+K-1,I.e-1);              /* gains:  p in [I.b,e1)  */
    const int b3 = imax(I.e-K+1,I.b);                /* drops:  q in [b3,I.e-1) */
    int prev = prof[I.b];
    CPG_UNROLL4 for (int p = I.b; p < e1; p++) { const int nx = prof[p+1]; n_gain += imax(nx-prev,0); prev = nx; }
    prev = (b3 < I.e-1) ? prof[b3] : 0;
    CPG_UNROLL4 for (int q = b3; q < I.e-1; q++) { const int nx = prof[q+1]; n_drop += imax(prev-nx,0); prev = nx; }
  }
  /* minus the part explained by the low-complexity run at that end */
  { int e2 = I.b, b4 = I.e-1;                        /* empty ranges unless the interval is longer than K-1 */
    if (I.b+K-1 < I.e)
      { int cl[3], lmax = 0;
        cpg_rctx3(seq,rlen,I.b+K-1,cl);
        if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) W.status |= CPG_ST_LONG_RUN;
        CPG_LOOP for (int t = 0; t < CT_N; t++) lmax = imax(lmax,cl[t]*(t+1));
        e2 = I.b+lmax;                               /* p in [I.b,e2)  */
      }
    if (I.b < I.e-K+1)
      { int cl[3], lmax = 0;
        cpg_lctx3(seq,rlen,I.e-K+1+K-2,cl);
        if (imax(imax(cl[0],cl[1]),cl[2]) >= 127) W.status |= CPG_ST_LONG_RUN;
        CPG_LOOP for (int t = 0; t < CT_N; t++) lmax = imax(lmax,cl[t]*(t+1));
        b4 = I.e-lmax;                               /* q in [b4,I.e-1) */
      }
    CPG_LOOP for (int p = I.b; p < e2; p++)
      { int nx = p+1;
        if (nx >= plen) { W.status |= MK_STALE_PROF; nx = plen-1; }       /* src/wall.c:977-978 reads profile[plen] */
        n_gain -= imax((int)prof[p]-(int)prof[nx],0);
      }
    CPG_LOOP for (int q = b4; q < I.e-1; q++) n_drop -= imax((int)prof[q+1]-(int)prof[q],0);
  }
  uint16_t ccb = (uint16_t)imin(I.cb+imax(n_gain,0),CPG_MAX_CNT);
  uint16_t cce = (uint16_t)imin(I.ce+imax(n_drop,0),CPG_MAX_CNT);
  /* src/wall.c:999-1006 index intvl[] with a POSITION that hides the interval index; the only
     write that can land on this interval is the one at position I.b, when I.b == idx:
     ccb = max(ccb,profile[I.b]) is a no-op, cce = max(cce,profile[I.b]) happens iff the scan
     [max(I.e-2K,I.b),I.e) starts at I.b.  Writes to higher slots hit intervals that are either
     recomputed from scratch later or never read. */
  if (I.b == idx && I.e-2*K <= I.b && cce < I.cb) cce = I.cb;
  I.ccb = ccb; I.cce = cce;
}

/* end counts of interval idx and the reliability test of src/wall.c:1016-1037; returns is_rel */
CPG_DEV_NOINL int wc_interval(const uint16_t *prof, int plen, const cpg_seq seq, int rlen, WCtx &W, cpg_intvl *v, int idx)
{ const cpg_dmodel *M = W.M;
  cpg_intvl I = v[idx];
  I.cb = prof[I.b]; I.ce = prof[I.e-1];
  int rel = 0;
  if (I.e-I.b >= M->K && imax(I.cb,I.ce) < M->cov[ST_R] && !(I.pe >= cpg_log(CPG_PE_FINAL)))
    { wc_correct(prof,plen,seq,rlen,W,I,idx);
      const int ccb = I.ccb, cce = I.cce;
      rel = !(cpg_lp_trans_thr(W,I.b,I.e,ccb,cce,(uint16_t)((ccb+cce)/2),CPG_THRES_DIFF_REL) < CPG_THRES_DIFF_REL);
      if (imax(ccb,cce) == CPG_MAX_CNT) rel = 0;
    }
  I.is_rel = (uint8_t)rel;
  v[idx] = I;
  return rel;
}

/* all intervals of a read, one per lane at a time; the reliable ones are copied to rint in order.
   Returns their number. */
CPG_DEV_NOINL int wc_read(const uint16_t *prof, int plen, const cpg_seq seq, int rlen, WCtx &W, cpg_intvl *v, int N, cpg_intvl *rint)
{ int Mrel = 0;
  CPG_SYNCGROUP(W);
  CPG_LOOP for (int base = 0; base < N; base += W.gsize)
    { const int i = base+W.glane;
      int rel = 0;
      if (i < N) rel = wc_interval(prof,plen,seq,rlen,W,v,i);
      const unsigned m = cpg_gballot(W,rel);
      if (rel) rint[Mrel+cpg_popc(m & ((1u << W.glane)-1u))] = v[i];
      Mrel += cpg_popc(m);
    }
  CPG_SYNCGROUP(W);
  return Mrel;
}

/* the pure step for the npend collected candidates, one per lane; 1 = the tables are full */
CPG_DEV_HELPER int wa_flush(ReadCtx &R, WCtx &W, int ncand, int npend, int mypos)
{ if (ncand+npend > R.S.capC) { W.status |= CPG_ST_RETRY; return 1; }
  if (W.glane < npend)
    wa_candidate(R.prof,R.plen,R.seq,R.rlen,W,mypos,R.S.hdr+ncand+W.glane,R.S.big,(uint32_t)(ncand+W.glane),R.prune);
  return 0;
}

/* ---- whole wall stage for one read, the three steps back to back (retry launch of the library, host
 *      tests): fills R.S.intvl[0..N) and R.S.rint[0..M).  The candidates come from the decoder's bit
 *      map, 32 positions per lane; each lane of the group runs the pure step for one of them. ---- */
CPG_DEV_NOINL void find_walls_and_reliable(ReadCtx &R, WCtx &W)
{ const int plen = R.plen;
  R.N = 0; R.M = 0;
  /* step 1: candidate positions are collected gsize at a time, then every lane takes one */
  int ncand = 0, npend = 0, mypos = 0;
  CPG_LOOP for (int base = 0; base < plen; base += 32*W.gsize)
    { const int p0 = base+32*W.glane;
      unsigned cw = (p0 < plen) ? R.cand[p0 >> 5] : 0u;
      if (p0+32 > plen && p0 < plen) cw &= (1u << (plen-p0))-1u;      /* bits past the profile: none are set, but do not rely on it */
      unsigned lanes = cpg_gballot(W,cw != 0u);
      CPG_LOOP while (lanes)
        { const int l = cpg_ffs(lanes)-1; lanes &= lanes-1;
          unsigned m = cpg_gshfl(W,cw,l);
          CPG_LOOP while (m)
            { const int bit = cpg_ffs(m)-1; m &= m-1;
              if (npend == W.glane) mypos = base+32*l+bit;
              if (++npend == W.gsize)
                { if (wa_flush(R,W,ncand,npend,mypos)) return;
                  ncand += npend; npend = 0;
                }
            }
        }
    }
  if (npend > 0)
    { if (wa_flush(R,W,ncand,npend,mypos)) return;
      ncand += npend;
    }
  CPG_SYNCGROUP(W);
  R.hdr = R.S.hdr; R.big = R.S.big; R.ncand = ncand;
  /* step 2 */
  const int NS = wb_walls(R,W);
  int N = 0;
  if (!(W.status & CPG_ST_ABORT))
    { N = wb_cuts(R,W,NS,R.S.intvl,R.S.capI,0);
      if (N > R.S.capI) W.status |= CPG_ST_RETRY;
    }
  wb_clean(R,W);
  if (W.status & CPG_ST_ABORT) return;
  /* step 3 */
  R.N = N;
  R.M = wc_read(R.prof,plen,R.seq,R.rlen,W,R.S.intvl,N,R.S.rint);
}

#endif
