/*******************************************************************************************
 *  cpsim -- harness data generator (NOT part of the product path).
 *
 *  Writes, from fixed seeds and with no network or external tool:
 *    <out>.fasta                  simulated HiFi-like reads of a synthetic diploid genome
 *    <out>.hist                   FastK histogram        (format read by libfastk.c:51-96)
 *    <out>.prof                   FastK profile stub     (libfastk.c:1283-1293)
 *    .<out>.pidx.<p>, .<out>.prof.<p>   per-part index + compressed profiles
 *                                 (libfastk.c:1299-1336, codec libfastk.c:1467-1535)
 *  FastK itself is not vendored in the reference, so the harness produces the same on-disk
 *  format from exact canonical k-mer counts of the simulated read set (mode "exact"), or from
 *  the simulator's ground-truth coverage (mode "fast", for bench-sized inputs where an exact
 *  count of 10^9..10^11 k-mers is out of reach of a harness tool).
 *
 *  Also usable as a library (libcpsim.so): see cpsim.h.
 *******************************************************************************************/
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <errno.h>
#include "cpsim.h"

typedef unsigned __int128 u128;

/* ---------- PRNG (splitmix64 seeding xoshiro256**) ---------- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t *x)
{ uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t *r, uint64_t seed)
{ for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64-k)); }
static inline uint64_t rng_next(rng_t *r)
{ uint64_t *s = r->s, res = rotl(s[1]*5,7)*9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3],45);
  return res;
}
static inline double rng_unif(rng_t *r) { return (rng_next(r) >> 11) * (1.0/9007199254740992.0); }
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return (uint64_t)(rng_unif(r)*n); }
static double rng_normal(rng_t *r)
{ double u1 = rng_unif(r), u2 = rng_unif(r);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0*log(u1))*cos(6.283185307179586*u2);
}

static const char BASES[4] = { 'A','C','G','T' };

static void *xmalloc(size_t n)
{ void *p = malloc(n ? n : 1);
  if (p == NULL) { fprintf(stderr,"cpsim: out of memory (%zu bytes)\n",n); exit(1); }
  return p;
}

/* ---------- growable byte string ---------- */
typedef struct { uint8_t *s; int64_t n, m; } bstr;
static void bs_push(bstr *b, uint8_t c)
{ if (b->n >= b->m) { b->m = b->m ? b->m*2 : 1024; b->s = realloc(b->s,b->m); if (!b->s) exit(1); }
  b->s[b->n++] = c;
}
static void bs_append(bstr *b, const uint8_t *p, int64_t n)
{ for (int64_t i = 0; i < n; i++) bs_push(b,p[i]); }

/* ---------- genome ---------- */
/* Haplotype 0 is built from i.i.d. bases with optional repeat content; haplotype 1 is a copy
 * with SNPs (and, unless snp_only, short indels) at rate `het`. Bases are codes 0..3. */
static void add_random(bstr *g, rng_t *r, int64_t n)
{ for (int64_t i = 0; i < n; i++) bs_push(g,(uint8_t)(rng_next(r)&3)); }

static void build_hap0(const cpsim_params *P, rng_t *r, bstr *g)
{ int64_t L = P->genome_len;
  /* library of interspersed repeat elements */
  enum { NLIB = 12 };
  bstr lib[NLIB]; memset(lib,0,sizeof(lib));
  for (int i = 0; i < NLIB; i++) add_random(&lib[i],r,300+(int64_t)rng_below(r,2700));

  while (g->n < L)
    { double u = rng_unif(r);
      if (u >= P->repeat_frac)
        { add_random(g,r,2000+(int64_t)rng_below(r,6000)); continue; }
      /* a repeat block; roughly as long as a unique block so repeat_frac ~ base fraction */
      double v = rng_unif(r);
      int64_t target = 2000+(int64_t)rng_below(r,6000);
      int64_t start = g->n;
      while (g->n-start < target)
        { if (v < 0.35)          /* tandem repeat: unit 2..60 bp, imperfect copies */
            { int ulen = 2+(int)rng_below(r,59);
              uint8_t unit[64];
              for (int k = 0; k < ulen; k++) unit[k] = (uint8_t)(rng_next(r)&3);
              int64_t span = 100+(int64_t)rng_below(r,1900);
              for (int64_t k = 0; k < span; k++)
                { uint8_t c = unit[k%ulen];
                  if (rng_unif(r) < 0.01) c = (uint8_t)(rng_next(r)&3);
                  bs_push(g,c);
                }
              add_random(g,r,50+(int64_t)rng_below(r,400));
            }
          else if (v < 0.85)     /* interspersed element, 0-5 % diverged copy */
            { int e = (int)rng_below(r,NLIB);
              double div = 0.05*rng_unif(r);
              for (int64_t k = 0; k < lib[e].n; k++)
                { uint8_t c = lib[e].s[k];
                  if (rng_unif(r) < div) c = (uint8_t)(rng_next(r)&3);
                  bs_push(g,c);
                }
              add_random(g,r,100+(int64_t)rng_below(r,600));
            }
          else                   /* low-complexity: homopolymer / dinucleotide / trinucleotide runs */
            { int ulen = 1+(int)rng_below(r,3);
              uint8_t unit[3];
              do { for (int k = 0; k < ulen; k++) unit[k] = (uint8_t)(rng_next(r)&3); }
              while (ulen > 1 && unit[0] == unit[1] && (ulen == 2 || unit[1] == unit[2]));
              int64_t span = (ulen == 1) ? 8+(int64_t)rng_below(r,50) : ulen*(5+(int64_t)rng_below(r,40));
              for (int64_t k = 0; k < span; k++) bs_push(g,unit[k%ulen]);
              add_random(g,r,40+(int64_t)rng_below(r,300));
            }
          v = rng_unif(r);
        }
    }
  g->n = L;
  /* segmental duplications: copy 5-20 kb blocks elsewhere (overwrite) */
  for (int d = 0; d < P->seg_dups && L > 60000; d++)
    { int64_t len = 5000+(int64_t)rng_below(r,15000);
      int64_t src = (int64_t)rng_below(r,L-len), dst = (int64_t)rng_below(r,L-len);
      if (llabs(src-dst) < len) continue;
      for (int64_t k = 0; k < len; k++)
        { uint8_t c = g->s[src+k];
          if (rng_unif(r) < 0.002) c = (uint8_t)(rng_next(r)&3);
          g->s[dst+k] = c;
        }
    }
  for (int i = 0; i < NLIB; i++) free(lib[i].s);
}

static void build_hap1(const cpsim_params *P, rng_t *r, const bstr *h0, bstr *h1)
{ for (int64_t i = 0; i < h0->n; i++)
    { if (rng_unif(r) < P->het)
        { double u = rng_unif(r);
          if (P->snp_only || u < 0.85)
            bs_push(h1,(uint8_t)((h0->s[i]+1+rng_below(r,3))&3));
          else if (u < 0.925)
            { int n = 1+(int)rng_below(r,3);          /* insertion */
              bs_push(h1,h0->s[i]);
              for (int k = 0; k < n; k++) bs_push(h1,(uint8_t)(rng_next(r)&3));
            }
          else
            i += (int64_t)rng_below(r,3);             /* deletion of 1-3 bases */
        }
      else
        bs_push(h1,h0->s[i]);
    }
}

/* ---------- read simulation ---------- */
/* HiFi-like: rare substitutions, indels concentrated in homopolymer runs.  For every emitted
 * base we also record whether it is a faithful copy (err==0) and, in fast mode, its genome
 * coordinate. */
typedef struct
  { uint8_t *seq;      /* codes 0..3 */
    uint8_t *err;      /* 1 if this base is an inserted/substituted base, or follows a deletion */
    int32_t *gpos;     /* genome coordinate of faithful bases (fast mode) */
    int      len, cap;
  } simread;

static void sr_push(simread *R, uint8_t c, uint8_t e, int32_t gp)
{ if (R->len >= R->cap)
    { R->cap = R->cap ? R->cap*2 : 32768;
      R->seq = realloc(R->seq,R->cap);
      R->err = realloc(R->err,R->cap);
      R->gpos = realloc(R->gpos,sizeof(int32_t)*(size_t)R->cap);
      if (!R->seq || !R->err || !R->gpos) exit(1);
    }
  R->seq[R->len] = c; R->err[R->len] = e; R->gpos[R->len] = gp; R->len++;
}

static void sim_one(const cpsim_params *P, rng_t *r, const bstr *hap, int64_t pos, int len, simread *R)
{ R->len = 0;
  int64_t i = pos, end = pos+len;
  uint8_t pending_del = 0;
  while (i < end)
    { /* homopolymer run starting at i (within template) */
      int64_t j = i+1;
      while (j < end && hap->s[j] == hap->s[i]) j++;
      int l = (int)(j-i);
      double pindel = P->err_indel_base + P->err_indel_hp*(double)l*l;
      if (pindel > 0.25) pindel = 0.25;
      int delta = 0;
      if (l >= 1 && rng_unif(r) < pindel)
        delta = (rng_unif(r) < 0.5) ? -1 : 1;
      int out = l+delta;
      for (int k = 0; k < out; k++)
        { uint8_t c = hap->s[i];
          uint8_t e = pending_del;
          pending_del = 0;
          if (k >= l) e = 1;                       /* inserted base */
          if (rng_unif(r) < P->err_sub)
            { c = (uint8_t)((c+1+rng_below(r,3))&3); e = 1; }
          sr_push(R,c,e,(k < l) ? (int32_t)(i+k) : -1);
        }
      if (delta < 0) pending_del = 1;             /* next emitted base follows a deletion */
      i = j;
    }
}

static inline uint8_t comp(uint8_t c) { return (uint8_t)(3-c); }

/* ---------- exact canonical k-mer counting ---------- */
typedef struct { u128 key; uint64_t idx; } kent;

static int cmp_kent(const void *a, const void *b)
{ const kent *x = a, *y = b;
  if (x->key < y->key) return -1;
  if (x->key > y->key) return 1;
  return 0;
}

/* LSD radix sort on the low `bits` bits of key, 11 bits per pass */
static void radix_sort(kent *a, kent *tmp, int64_t n, int bits)
{ const int RB = 11, NB = 1<<RB;
  int64_t *cnt = xmalloc(sizeof(int64_t)*NB);
  for (int sh = 0; sh < bits; sh += RB)
    { memset(cnt,0,sizeof(int64_t)*NB);
      for (int64_t i = 0; i < n; i++) cnt[(int)((a[i].key >> sh) & (NB-1))]++;
      int64_t s = 0;
      for (int b = 0; b < NB; b++) { int64_t c = cnt[b]; cnt[b] = s; s += c; }
      for (int64_t i = 0; i < n; i++) tmp[cnt[(int)((a[i].key >> sh) & (NB-1))]++] = a[i];
      kent *t = a; a = tmp; tmp = t;
    }
  /* number of passes */
  int passes = (bits+RB-1)/RB;
  if (passes & 1) memcpy(tmp,a,sizeof(kent)*(size_t)n);   /* result must end in original `a` */
  free(cnt);
}

/* ---------- FastK profile codec (encoder side of libfastk.c:1467-1535) ---------- */
static void encode_profile(const uint16_t *c, int n, bstr *out)
{ if (n <= 0) return;
  uint16_t d = c[0];
  if (d >= 128) { bs_push(out,(uint8_t)(0x80|(d>>8))); bs_push(out,(uint8_t)(d&0xff)); }
  else bs_push(out,(uint8_t)d);
  int i = 1;
  while (i < n)
    { if (c[i] == d)
        { int run = 1;
          while (i+run < n && c[i+run] == d && run < 63) run++;
          bs_push(out,(uint8_t)run);
          i += run;
          continue;
        }
      int diff = (int)c[i]-(int)d;
      if (diff >= -32 && diff <= 31)
        { if (diff >= 0) bs_push(out,(uint8_t)(0x40|diff));
          else           bs_push(out,(uint8_t)(0x40|0x20|((diff+32)&0x1f)));
        }
      else
        { uint16_t x = (uint16_t)(diff & 0x7fff);   /* 15-bit two's complement */
          bs_push(out,(uint8_t)(0x80|(x>>8)));
          bs_push(out,(uint8_t)(x&0xff));
        }
      d = c[i];
      i++;
    }
}

/* ---------- main generation ---------- */
void cpsim_default_params(cpsim_params *P)
{ memset(P,0,sizeof(*P));
  P->seed = 1; P->genome_len = 300000; P->het = 0.005; P->snp_only = 0;
  P->repeat_frac = 0.; P->seg_dups = 0;
  P->cov = 30.; P->len_mean = 12000; P->len_sd = 1500; P->len_min = 2000; P->len_max = 50000;
  P->err_sub = 0.0002; P->err_indel_base = 0.0002; P->err_indel_hp = 0.0006;
  P->kmer = 40; P->nparts = 2; P->exact = 1; P->short_reads = 0;
}

void cpsim_free(cpsim_data *D)
{ free(D->seq); free(D->seq_off); free(D->rlen); free(D->counts); free(D->cnt_off);
  free(D->prof); free(D->prof_off); free(D->hist); free(D->hdr); free(D->hdr_off);
  memset(D,0,sizeof(*D));
}

int cpsim_generate(const cpsim_params *P, cpsim_data *D)
{ rng_t rg, rr;
  rng_seed(&rg,P->seed*2654435761u+11);
  rng_seed(&rr,P->seed*40503u+977);
  const int K = P->kmer;
  if (K < 2 || K > 63) { fprintf(stderr,"cpsim: k must be in [2,63]\n"); return 1; }
  memset(D,0,sizeof(*D));
  D->kmer = K;

  bstr hap[2]; memset(hap,0,sizeof(hap));
  build_hap0(P,&rg,&hap[0]);
  build_hap1(P,&rg,&hap[0],&hap[1]);

  /* number of reads for the requested coverage of the diploid genome */
  int64_t target = (int64_t)(P->cov*(double)P->genome_len);
  int64_t nreads_cap = target/(P->len_min > 0 ? P->len_min : 1)+16;
  D->seq_off = xmalloc(sizeof(int64_t)*(size_t)(nreads_cap+1));
  D->rlen    = xmalloc(sizeof(int32_t)*(size_t)nreads_cap);
  D->hdr_off = xmalloc(sizeof(int64_t)*(size_t)(nreads_cap+1));
  bstr seq = {0}, hdr = {0};
  bstr errv = {0};
  int32_t *gposv = NULL; int64_t gpos_cap = 0;
  uint8_t *hapv = xmalloc((size_t)nreads_cap), *strandv = xmalloc((size_t)nreads_cap);

  simread R; memset(&R,0,sizeof(R));
  int64_t nreads = 0, tot = 0;
  D->seq_off[0] = 0; D->hdr_off[0] = 0;
  while (tot < target && nreads < nreads_cap)
    { int len = (int)(P->len_mean+P->len_sd*rng_normal(&rr));
      if (len < P->len_min) len = P->len_min;
      if (len > P->len_max) len = P->len_max;
      if (P->short_reads && (nreads % 97) == 13) len = 1+(int)rng_below(&rr,(uint64_t)(K+8));
      int h = (int)(rng_next(&rr)&1);
      if (len > hap[h].n) len = (int)hap[h].n;
      int64_t pos = (int64_t)rng_below(&rr,(uint64_t)(hap[h].n-len+1));
      int strand = (int)(rng_next(&rr)&1);
      sim_one(P,&rr,&hap[h],pos,len,&R);
      if (R.len > P->len_max+64) R.len = P->len_max+64;
      if (gpos_cap < seq.n+R.len)
        { gpos_cap = (seq.n+R.len)*2+1024; gposv = realloc(gposv,sizeof(int32_t)*(size_t)gpos_cap); }
      if (strand == 0)
        for (int i = 0; i < R.len; i++)
          { gposv[seq.n] = R.gpos[i]; bs_push(&seq,R.seq[i]); bs_push(&errv,R.err[i]); }
      else
        for (int i = R.len-1; i >= 0; i--)
          { gposv[seq.n] = R.gpos[i]; bs_push(&seq,comp(R.seq[i]));
            bs_push(&errv,R.err[i]);
          }
      char hb[128];
      int hl = snprintf(hb,sizeof(hb),"Sim %lld %d %d %lld %d",(long long)(nreads+1),h,strand,(long long)pos,R.len);
      bs_append(&hdr,(uint8_t*)hb,hl);
      hapv[nreads] = (uint8_t)h; strandv[nreads] = (uint8_t)strand;
      D->rlen[nreads] = R.len;
      nreads++;
      D->seq_off[nreads] = seq.n;
      D->hdr_off[nreads] = hdr.n;
      tot += R.len;
    }
  free(R.seq); free(R.err); free(R.gpos);
  D->nreads = nreads;
  D->seq = seq.s; D->hdr = (char*)hdr.s;
  D->total_bases = seq.n;

  /* profile offsets */
  D->cnt_off = xmalloc(sizeof(int64_t)*(size_t)(nreads+1));
  D->cnt_off[0] = 0;
  for (int64_t i = 0; i < nreads; i++)
    { int pl = D->rlen[i]-K+1; if (pl < 0) pl = 0;
      D->cnt_off[i+1] = D->cnt_off[i]+pl;
    }
  int64_t NK = D->cnt_off[nreads];
  D->total_kmers = NK;
  D->counts = xmalloc(sizeof(uint16_t)*(size_t)(NK+1));
  D->hist = calloc(32768+2,sizeof(int64_t));

  if (P->exact == 2)
    { /* reads only: the counts come from somewhere else (the GPU profile producer, classpro_b200/profiler) */
      memset(D->counts,0,sizeof(uint16_t)*(size_t)(NK+1));
    }
  else if (P->exact)
    { kent *a = xmalloc(sizeof(kent)*(size_t)(NK+1)), *tmp = xmalloc(sizeof(kent)*(size_t)(NK+1));
      const u128 mask = (K == 64) ? ~(u128)0 : (((u128)1 << (2*K))-1);
      int64_t m = 0;
      for (int64_t rd = 0; rd < nreads; rd++)
        { const uint8_t *s = D->seq+D->seq_off[rd];
          int L = D->rlen[rd];
          if (L < K) continue;
          u128 f = 0, rv = 0;
          for (int i = 0; i < L; i++)
            { f  = ((f << 2) | s[i]) & mask;
              rv = (rv >> 2) | ((u128)(3-s[i]) << (2*(K-1)));
              if (i >= K-1)
                { a[m].key = (f < rv) ? f : rv;
                  a[m].idx = (uint64_t)m;
                  m++;
                }
            }
        }
      if (m != NK) { fprintf(stderr,"cpsim: internal k-mer count mismatch\n"); return 1; }
      if (NK > 4000000) radix_sort(a,tmp,NK,2*K);
      else qsort(a,(size_t)NK,sizeof(kent),cmp_kent);
      for (int64_t i = 0; i < NK; )
        { int64_t j = i+1;
          while (j < NK && a[j].key == a[i].key) j++;
          int64_t c = j-i;
          uint16_t cc = (uint16_t)(c > 32767 ? 32767 : c);
          for (int64_t k = i; k < j; k++) D->counts[a[k].idx] = cc;
          D->hist[cc] += 1;
          if (cc == 32767) D->hist[32769] += c;   /* instance count of the top bin */
          i = j;
        }
      free(a); free(tmp);
    }
  else
    { /* ground-truth coverage: error-free k-mers inherit the number of error-free read k-mers
         covering the same genomic k-mer (both haplotypes if identical there; snp_only assumed) */
      int64_t GL = hap[0].n;
      int32_t *cov[2];
      for (int h = 0; h < 2; h++) cov[h] = calloc((size_t)GL+2,sizeof(int32_t));
      uint8_t *shared = xmalloc((size_t)GL+1);      /* k-mer at g identical on both haplotypes */
      {
        int64_t n = (hap[1].n < GL) ? hap[1].n : GL;
        /* shared[g] = no SNP in [g,g+K) */
        int64_t *nx = xmalloc(sizeof(int64_t)*(size_t)(GL+1));
        int64_t nxt = GL+K;
        for (int64_t g = GL-1; g >= 0; g--)
          { if (g >= n || hap[0].s[g] != hap[1].s[g]) nxt = g;
            nx[g] = nxt;
          }
        for (int64_t g = 0; g < GL; g++) shared[g] = (nx[g] >= g+K);
        free(nx);
      }
      /* pass 1: accumulate error-free k-mer coverage as difference arrays */
      for (int pass = 0; pass < 2; pass++)
        { if (pass == 1)
            for (int h = 0; h < 2; h++)
              { int64_t run = 0;
                for (int64_t g = 0; g <= GL; g++) { run += cov[h][g]; cov[h][g] = (int32_t)run; }
              }
          for (int64_t rd = 0; rd < nreads; rd++)
            { const int h = hapv[rd];
              const uint8_t *e = errv.s+D->seq_off[rd];
              const int32_t *gp = gposv+D->seq_off[rd];
              int L = D->rlen[rd];
              uint16_t *out = D->counts+D->cnt_off[rd];
              /* window [st,st+K) is clean iff none of its bases is flagged and its genome
                 coordinates are contiguous (a flag on the first base over-approximates a junction
                 before the window: harmless for a plausibility generator) */
              int last = -K-1;
              for (int i = 0; i < L; i++)
                { if (e[i]) last = i;
                  if (i >= K-1)
                    { int st = i-K+1;
                      int clean = (last < st);
                      int32_t g0 = gp[st], g1 = gp[i];
                      int64_t g = (g0 < g1) ? g0 : g1;
                      if (clean && g0 >= 0 && g1 >= 0 && llabs((int64_t)g1-g0) == K-1)
                        { if (pass == 0) { cov[h][g] += 1; cov[h][g+1] -= 1; }
                          else
                            { int64_t c = cov[h][g];
                              if (shared[g]) c += cov[1-h][g];
                              if (c < 1) c = 1;
                              out[st] = (uint16_t)(c > 32767 ? 32767 : c);
                            }
                        }
                      else if (pass == 1)
                        out[st] = 1;
                    }
                }
            }
        }
      /* approximate histogram of distinct k-mers from instance counts */
      for (int64_t i = 0; i < NK; i++) D->hist[D->counts[i]] += 1;
      for (int c = 2; c <= 32767; c++) D->hist[c] = (D->hist[c]+c/2)/c;
      for (int h = 0; h < 2; h++) free(cov[h]);
      free(shared);
    }
  D->hist[32768] = D->hist[1];   /* instance-mode value of the low bin (count 1) */
  if (!P->exact) D->hist[32769] = D->hist[32767]*32767;

  /* compressed profiles */
  { bstr pc = {0};
    D->prof_off = xmalloc(sizeof(int64_t)*(size_t)(nreads+1));
    D->prof_off[0] = 0;
    for (int64_t rd = 0; rd < nreads; rd++)
      { encode_profile(D->counts+D->cnt_off[rd],(int)(D->cnt_off[rd+1]-D->cnt_off[rd]),&pc);
        D->prof_off[rd+1] = pc.n;
      }
    D->prof = pc.s;
    D->prof_bytes = pc.n;
  }

  free(errv.s); free(gposv); free(hapv); free(strandv);
  free(hap[0].s); free(hap[1].s);
  return 0;
}

/* ---------- file writers ---------- */
static FILE *xopen(const char *path)
{ FILE *f = fopen(path,"wb");
  if (f == NULL) { fprintf(stderr,"cpsim: cannot write %s: %s\n",path,strerror(errno)); exit(1); }
  return f;
}

int cpsim_write_files(const cpsim_params *P, const cpsim_data *D, const char *dir, const char *root)
{ char path[4096];
  const int K = D->kmer;
  /* reads */
  snprintf(path,sizeof(path),"%s/%s.fasta",dir,root);
  FILE *f = xopen(path);
  char *line = xmalloc((size_t)P->len_max+256);
  for (int64_t rd = 0; rd < D->nreads; rd++)
    { fputc('>',f);
      fwrite(D->hdr+D->hdr_off[rd],1,(size_t)(D->hdr_off[rd+1]-D->hdr_off[rd]),f);
      fputc('\n',f);
      int L = D->rlen[rd];
      const uint8_t *s = D->seq+D->seq_off[rd];
      for (int i = 0; i < L; i++) line[i] = BASES[s[i]];
      line[L] = '\n';
      fwrite(line,1,(size_t)L+1,f);
    }
  free(line);
  fclose(f);
  if (P->exact == 2) return 0;                     /* reads only */
  /* histogram: kmer, low=1, high=32767, ilowcnt, ihighcnt, hist[1..32767] */
  snprintf(path,sizeof(path),"%s/%s.hist",dir,root);
  f = xopen(path);
  { int32_t k = K, low = 1, high = 32767;
    fwrite(&k,4,1,f); fwrite(&low,4,1,f); fwrite(&high,4,1,f);
    fwrite(&D->hist[32768],8,1,f); fwrite(&D->hist[32769],8,1,f);
    fwrite(D->hist+1,8,32767,f);
  }
  fclose(f);
  /* profile stub */
  int nparts = P->nparts < 1 ? 1 : P->nparts;
  if (nparts > D->nreads) nparts = (int)(D->nreads > 0 ? D->nreads : 1);
  snprintf(path,sizeof(path),"%s/%s.prof",dir,root);
  f = xopen(path);
  { int32_t k = K, np = nparts; fwrite(&k,4,1,f); fwrite(&np,4,1,f); }
  fclose(f);
  /* parts: contiguous read ranges */
  for (int p = 0; p < nparts; p++)
    { int64_t b = D->nreads*p/nparts, e = D->nreads*(p+1)/nparts;
      snprintf(path,sizeof(path),"%s/.%s.pidx.%d",dir,root,p+1);
      f = xopen(path);
      int32_t k = K; int64_t first = b, n = e-b;
      fwrite(&k,4,1,f); fwrite(&first,8,1,f); fwrite(&n,8,1,f);
      for (int64_t rd = b; rd < e; rd++)
        { int64_t off = D->prof_off[rd+1]-D->prof_off[b];
          fwrite(&off,8,1,f);
        }
      fclose(f);
      snprintf(path,sizeof(path),"%s/.%s.prof.%d",dir,root,p+1);
      f = xopen(path);
      fwrite(D->prof+D->prof_off[b],1,(size_t)(D->prof_off[e]-D->prof_off[b]),f);
      fclose(f);
    }
  return 0;
}

#ifdef CPSIM_MAIN
static void usage(void)
{ fprintf(stderr,
  "usage: cpsim [options] <out_dir> <root>\n"
  "  --seed N --genome-len N --het F --snp-only --repeat-frac F --seg-dups N\n"
  "  --cov F --len-mean N --len-sd N --len-min N --len-max N\n"
  "  --err-sub F --err-indel-base F --err-indel-hp F --kmer N --nparts N --fast --reads-only --short-reads\n");
  exit(1);
}

int main(int argc, char **argv)
{ cpsim_params P; cpsim_default_params(&P);
  const char *pos[2]; int np = 0;
  for (int i = 1; i < argc; i++)
    { const char *a = argv[i];
#define OPTI(name,field) if (!strcmp(a,name) && i+1 < argc) { P.field = atoll(argv[++i]); continue; }
#define OPTF(name,field) if (!strcmp(a,name) && i+1 < argc) { P.field = atof(argv[++i]); continue; }
      OPTI("--seed",seed) OPTI("--genome-len",genome_len) OPTF("--het",het)
      OPTF("--repeat-frac",repeat_frac) OPTI("--seg-dups",seg_dups)
      OPTF("--cov",cov) OPTI("--len-mean",len_mean) OPTI("--len-sd",len_sd)
      OPTI("--len-min",len_min) OPTI("--len-max",len_max)
      OPTF("--err-sub",err_sub) OPTF("--err-indel-base",err_indel_base) OPTF("--err-indel-hp",err_indel_hp)
      OPTI("--kmer",kmer) OPTI("--nparts",nparts)
      if (!strcmp(a,"--snp-only")) { P.snp_only = 1; continue; }
      if (!strcmp(a,"--fast")) { P.exact = 0; P.snp_only = 1; continue; }
      if (!strcmp(a,"--reads-only")) { P.exact = 2; continue; }
      if (!strcmp(a,"--short-reads")) { P.short_reads = 1; continue; }
      if (a[0] == '-') usage();
      if (np < 2) pos[np++] = a; else usage();
    }
  if (np != 2) usage();
  cpsim_data D;
  if (cpsim_generate(&P,&D)) return 1;
  cpsim_write_files(&P,&D,pos[0],pos[1]);
  fprintf(stderr,"cpsim: %lld reads, %lld bases, %lld k-mers, %lld profile bytes (%.3f B/k-mer)\n",
          (long long)D.nreads,(long long)D.total_bases,(long long)D.total_kmers,(long long)D.prof_bytes,
          D.total_kmers ? (double)D.prof_bytes/D.total_kmers : 0.);
  cpsim_free(&D);
  return 0;
}
#endif
