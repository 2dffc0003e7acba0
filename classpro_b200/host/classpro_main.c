/*******************************************************************************************
 *  classpro_main.c -- the ClassPro command line on top of libclasspro_b200.so.
 *
 *      ClassPro [-vs] [-T<int(4)>] [-c<int>] [-r<int(20000)>] [-P<tmp_dir(./)>] [-N<fastk_root>]
 *               [-M<model_path>] [-G<int>] [-B<int>] <source>[.f[ast][aq][.gz]]
 *
 *  Same options, inputs and output bytes as the reference program (src/ClassPro.c:348-631,
 *  usage string src/const.c:14-17): reads <source> plus the FastK <root>.hist / <root>.prof
 *  (+ hidden .pidx.N / .prof.N parts) and writes <dir of source>/<root>.class with one
 *  fastq-like record per read (src/ClassPro.c:289).
 *
 *  What changed is how the work is organised (the host stays C, the per-read path is CUDA):
 *    reader thread   parses the FASTX stream once (the reference re-parses it from byte 0 in every
 *                    thread, src/ClassPro.c:104-110) -- a plain file is mapped and parsed by the -T
 *                    threads at the same time, piece by piece, with the records of a serial kseq
 *                    pass (see "Reader" below) --, 2-bit packs the reads, reads the compressed
 *                    profiles of a whole batch with one pread per read, all on the -T threads (the
 *                    reference does one lseek + >= 2 read(4096) per read, src/libfastk.c:1444-1462);
 *    GPU workers     one host thread per GPU, each with its own cpg_ctx; batches are contiguous
 *                    read ranges handed out in order, no data ever moves between GPUs;
 *    writer thread   emits the records of finished batches in read order (the reference writes
 *                    per-thread part files and concatenates them, src/io.c:70-112): the -T threads
 *                    of its own pool format disjoint record ranges straight into a mapping of the
 *                    output file.
 *  -T<n> is the number of host threads of the reader's pool and of the writer's pool
 *  (the reference's default of 4), -P is accepted and unused (no part files), -G<n> limits the
 *  number of GPUs (default: all),
 *  -B<n> sets the batch size in megabases (default 64).
 *  Not supported, with a clear error: .db/.dam inputs (DAZZ_DB is out of scope), -M (needs GSL).
 *  Environment (tests / measurements): CPG_SERIAL_IO=1 one parser and write(), CPG_PARSE_PIECES=<n>
 *  pieces per window, CPG_BATCH_BASES=<n> batch size in bases.
 *  -s is accepted and ignored with a note: in the reference it never changes .class bytes and
 *  crashes on FASTX input (src/ClassPro.c:281-282, src/seed.c:548-572).
 *******************************************************************************************/
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <pthread.h>
#include <time.h>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <zlib.h>
#include <stdarg.h>
#include <strings.h>
#include "classpro_gpu.h"

static const char *PROG = "ClassPro";
static const char *USAGE = "[-vs] [-T<int(4)>] [-c<int>] [-r<int(20000)>] [-P<tmp_dir(./)>] [-N<fastk_root>] "
                           "[-M<model_path>] [-G<int>] [-B<int>] <source>[.f[ast][aq][.gz]]";
#define MAX_READ_LEN 60000       /* src/const.c:57 */

static double now_s(void);
static long   g_reparsed = 0, g_patched = 0;      /* pieces parsed again after a wrong guess, headers completed in flatten_chunks */
static double g_t_reader = 0., g_t_gpu_create = 0., g_t_gpu_busy = 0., g_t_writer = 0.;
#define MAX_GPUS 64
static long   g_gpu_batches[MAX_GPUS]; static long long g_gpu_kmers[MAX_GPUS];     /* per device: batches and k-mers classified */
/* wall-clock marks (seconds since start) printed with -v: model ready, first pinned buffer, GPU0
   context, first batch parsed, reader done, last batch collected, writer done */
static double g_tl_model = 0., g_tl_pinned = 0., g_tl_ctx = 0., g_tl_first = 0., g_tl_reader = 0., g_tl_collect = 0.,
              g_tl_writer = 0.;

#include "cpg_hostio.h"

/* ---------------------------------------------------------------------------------------
 *  Batches and the three-stage pipeline
 * --------------------------------------------------------------------------------------- */
/* the records one parser produced: header and sequence text back to back in one arena */
typedef struct
  { char    *arena; size_t len, cap;
    size_t  *hoff, *soff;        /* header "@name comment" / sequence of record j, both NUL terminated */
    int32_t *hlen, *rlen;
    uint8_t *nocm;               /* record j had no comment and no earlier record of this parser had one */
    int      n, n_cap;
    int64_t  start, next;        /* memory mode: marker position parsing started from / of the first record not parsed */
    int      status;             /* -2: a truncated quality string follows the n records */
    int      has_cm; size_t cm_off;   /* the comment a later comment-less record would repeat (src/ClassPro.c:188) */
  } chunk_t;

typedef struct batch
  { int64_t   first_id;
    int       n_all;             /* records in the batch, including reads shorter than K */
    int       n;                 /* reads sent to the GPU */
    /* per record; header and seq point into the arenas of ch[] (or into patch[]) */
    char    **header; char **seq; int32_t *hlen; int32_t *rlen_all; int32_t *slot_of;   /* slot_of[i] = index among the n, or -1 */
    int32_t  *lastc; int64_t *out_off;       /* writer: last classified record before i (-1: none), byte offset of record i */
    chunk_t  *ch; int nch;
    char    **patch; int npatch, patch_cap;  /* headers rebuilt with a comment carried over a chunk boundary */
    /* device-side inputs / outputs (pinned) */
    uint8_t  *pseq; int64_t *seq_off; int32_t *rlen; uint8_t *prof; int64_t *prof_off;
    /* result: the interval tables of the reads (cpg_result_ivl); the class strings are written straight into
       the output by the formatter (cpg_expand_intervals) */
    uint32_t *ivl; int64_t *ivl_at; int32_t *ivl_n; int64_t ivl_cap;
    int64_t  *cls_off; int32_t *status;
    size_t    pseq_cap, prof_cap; int rec_cap, n_cap;
    int       seq_bits;
    int64_t   kmers;
    struct batch *next;
  } batch_t;

typedef struct
  { pthread_mutex_t mu; pthread_cond_t cv;
    batch_t *head, *tail; int closed;
  } queue_t;

static void q_init(queue_t *q) { pthread_mutex_init(&q->mu,NULL); pthread_cond_init(&q->cv,NULL); q->head = q->tail = NULL; q->closed = 0; }
static void q_push(queue_t *q, batch_t *b)
{ pthread_mutex_lock(&q->mu);
  b->next = NULL;
  if (q->tail) q->tail->next = b; else q->head = b;
  q->tail = b;
  pthread_cond_broadcast(&q->cv);
  pthread_mutex_unlock(&q->mu);
}
static void q_close(queue_t *q)
{ pthread_mutex_lock(&q->mu); q->closed = 1; pthread_cond_broadcast(&q->cv); pthread_mutex_unlock(&q->mu); }
static batch_t *q_pop(queue_t *q)
{ pthread_mutex_lock(&q->mu);
  while (q->head == NULL && !q->closed) pthread_cond_wait(&q->cv,&q->mu);
  batch_t *b = q->head;
  if (b) { q->head = b->next; if (q->head == NULL) q->tail = NULL; }
  pthread_mutex_unlock(&q->mu);
  return b;
}

/* without waiting: NULL when nothing is queued right now */
static batch_t *q_trypop(queue_t *q)
{ pthread_mutex_lock(&q->mu);
  batch_t *b = q->head;
  if (b) { q->head = b->next; if (q->head == NULL) q->tail = NULL; }
  pthread_mutex_unlock(&q->mu);
  return b;
}

typedef struct
  { /* options */
    int verbose, find_seeds, nthreads, cov, read_len, ngpus; int64_t batch_bases;
    char *fk_root, *model_path, *src_path, *out_path;
    /* shared state */
    profidx_t  P;
    cpg_model *model;
    queue_t    q_free, q_ready;
    /* finished batches, handed to the writer in id order */
    pthread_mutex_t mu; pthread_cond_t cv;
    batch_t   *done;              /* unordered list */
    int64_t    next_id;           /* first_id the writer waits for */
    int        reader_done; int64_t total_reads;
    int64_t    kmers;
  } app_t;

static void *pinned(size_t n)
{ void *p = cpg_host_alloc(n);
  if (p == NULL) die("%s: cannot allocate %zu bytes of pinned memory",PROG,n);
  return p;
}

static void batch_reserve_records(batch_t *b, int nrec)
{ if (nrec > b->rec_cap)
    { int cap = nrec+nrec/2+64;
      b->header = xrealloc(b->header,sizeof(char *)*(size_t)cap);
      b->seq = xrealloc(b->seq,sizeof(char *)*(size_t)cap);
      b->hlen = xrealloc(b->hlen,sizeof(int32_t)*(size_t)cap);
      b->rlen_all = xrealloc(b->rlen_all,sizeof(int32_t)*(size_t)cap);
      b->slot_of = xrealloc(b->slot_of,sizeof(int32_t)*(size_t)cap);
      b->lastc = xrealloc(b->lastc,sizeof(int32_t)*(size_t)cap);
      b->out_off = xrealloc(b->out_off,sizeof(int64_t)*(size_t)(cap+1));
      b->rec_cap = cap;
    }
}

static void batch_reserve(batch_t *b, int nrec, size_t pseq, size_t prof, size_t cls)
{ batch_reserve_records(b,nrec);
  if (nrec > b->n_cap)
    { int cap = nrec+nrec/2+64;
      int64_t *so = pinned(sizeof(int64_t)*(size_t)(cap+1)), *po = pinned(sizeof(int64_t)*(size_t)(cap+1)),
              *co = pinned(sizeof(int64_t)*(size_t)(cap+1)), *ia = pinned(sizeof(int64_t)*(size_t)(cap+1));
      int32_t *in = pinned(sizeof(int32_t)*(size_t)(cap+1));
      int32_t *rl = pinned(sizeof(int32_t)*(size_t)(cap+1)), *st = pinned(sizeof(int32_t)*(size_t)(cap+1));
      if (b->n_cap)
        { memcpy(so,b->seq_off,sizeof(int64_t)*(size_t)(b->n_cap+1)); memcpy(po,b->prof_off,sizeof(int64_t)*(size_t)(b->n_cap+1));
          memcpy(co,b->cls_off,sizeof(int64_t)*(size_t)(b->n_cap+1)); memcpy(rl,b->rlen,sizeof(int32_t)*(size_t)(b->n_cap+1));
          cpg_host_free(b->seq_off); cpg_host_free(b->prof_off); cpg_host_free(b->cls_off); cpg_host_free(b->rlen); cpg_host_free(b->status);
          cpg_host_free(b->ivl_at); cpg_host_free(b->ivl_n);
        }
      b->seq_off = so; b->prof_off = po; b->cls_off = co; b->rlen = rl; b->status = st; b->ivl_at = ia; b->ivl_n = in;
      b->n_cap = cap;
    }
  if (pseq > b->pseq_cap)
    { size_t cap = pseq+pseq/2+4096; uint8_t *p = pinned(cap);
      if (b->pseq_cap) { memcpy(p,b->pseq,b->pseq_cap); cpg_host_free(b->pseq); }
      b->pseq = p; b->pseq_cap = cap;
    }
  if (prof > b->prof_cap)
    { size_t cap = prof+prof/2+4096; uint8_t *p = pinned(cap);
      if (b->prof_cap) { memcpy(p,b->prof,b->prof_cap); cpg_host_free(b->prof); }
      b->prof = p; b->prof_cap = cap;
    }
  (void)cls;
}

/* room for the interval tables of the batch in flight on `slot` (known once it is submitted) */
static void batch_reserve_ivl(batch_t *b, int64_t need)
{ if (need <= b->ivl_cap) return;
  const int64_t cap = need+need/4+4096;
  uint32_t *p = pinned(sizeof(uint32_t)*(size_t)cap);
  if (b->ivl_cap) cpg_host_free(b->ivl);
  b->ivl = p; b->ivl_cap = cap;
}

/* ---------------------------------------------------------------------------------------
 *  A small fork-join pool: the reader hands the records of a parsed batch to -T host threads
 *  (itself included) for 2-bit packing and profile reads; chunks are claimed with a counter.
 * --------------------------------------------------------------------------------------- */
typedef struct
  { pthread_mutex_t mu; pthread_cond_t cv_job, cv_done;
    pthread_t *th; int nth;
    void (*fn)(void *, int, int); void *arg;
    int total, chunk, next, busy, gen, stop;
  } pool_t;

static void pool_work(pool_t *P)                      /* called with P->mu held */
{ while (P->next < P->total)
    { int lo = P->next, hi = lo+P->chunk < P->total ? lo+P->chunk : P->total;
      P->next = hi;
      pthread_mutex_unlock(&P->mu);
      P->fn(P->arg,lo,hi);
      pthread_mutex_lock(&P->mu);
    }
}

static void *pool_main(void *arg)
{ pool_t *P = arg;
  int seen = 0;
  pthread_mutex_lock(&P->mu);
  for (;;)
    { while (!P->stop && P->gen == seen) pthread_cond_wait(&P->cv_job,&P->mu);
      if (P->stop) break;
      seen = P->gen;
      P->busy++;
      pool_work(P);
      if (--P->busy == 0) pthread_cond_broadcast(&P->cv_done);
    }
  pthread_mutex_unlock(&P->mu);
  return NULL;
}

static void pool_start(pool_t *P, int nthreads)
{ memset(P,0,sizeof(*P));
  pthread_mutex_init(&P->mu,NULL); pthread_cond_init(&P->cv_job,NULL); pthread_cond_init(&P->cv_done,NULL);
  P->nth = nthreads > 1 ? nthreads-1 : 0;
  P->th = xmalloc(sizeof(pthread_t)*(size_t)(P->nth+1));
  for (int i = 0; i < P->nth; i++) pthread_create(&P->th[i],NULL,pool_main,P);
}

static void pool_run(pool_t *P, void (*fn)(void *, int, int), void *arg, int total, int chunk)
{ pthread_mutex_lock(&P->mu);
  P->fn = fn; P->arg = arg; P->total = total; P->chunk = chunk > 0 ? chunk : 1; P->next = 0;
  P->gen++;
  pthread_cond_broadcast(&P->cv_job);
  P->busy++;
  pool_work(P);
  P->busy--;
  while (P->busy > 0) pthread_cond_wait(&P->cv_done,&P->mu);
  pthread_mutex_unlock(&P->mu);
}

static void pool_stop(pool_t *P)
{ pthread_mutex_lock(&P->mu); P->stop = 1; pthread_cond_broadcast(&P->cv_job); pthread_mutex_unlock(&P->mu);
  for (int i = 0; i < P->nth; i++) pthread_join(P->th[i],NULL);
  free(P->th);
}

typedef struct { app_t *A; batch_t *b; int bad; } packjob_t;

/* records lo..hi-1 of the batch: 2-bit pack the read, fetch its compressed profile */
static void pack_records(void *arg, int lo, int hi)
{ packjob_t *J = arg; app_t *A = J->A; batch_t *b = J->b;
  int bad = 0;
  for (int i = lo; i < hi; i++)
    { const int k = b->slot_of[i];
      if (k < 0) continue;
      const int rlen = b->rlen_all[i];
      bad |= cpg_pack_seq(b->seq[i],rlen,b->pseq+b->seq_off[k]);
      int part; int64_t off, len;
      prof_range(&A->P,b->first_id+i,&part,&off,&len);
      int64_t got = 0;
      while (got < len)
        { ssize_t r = pread(A->P.fd[part],b->prof+b->prof_off[k]+got,(size_t)(len-got),off+got);
          if (r <= 0) die("%s: cannot read profile of read %lld",PROG,(long long)(b->first_id+i+1));
          got += r;
        }
    }
  if (bad) __atomic_store_n(&J->bad,1,__ATOMIC_RELAXED);
}

/* ---------------------------------------------------------------------------------------
 *  Reader: FASTX -> batches.
 *  A plain (not gzipped) regular file is mapped and parsed by the -T pool threads at the same
 *  time: the next window of the file is cut into pieces at guessed record starts, every piece is
 *  parsed by its own kseq-semantics parser, and the pieces are then checked in order -- a piece is
 *  kept iff the parser of the piece before it stopped exactly at its start (the parser state
 *  between two records is the marker position, fx_next_marker, plus the last comment seen, which
 *  flatten_chunks carries over); otherwise it is parsed again from the true position.  So the
 *  records are those of one serial pass whatever the guesses were (a FASTQ quality line may begin
 *  with '@').  gzip input and non-mappable sources keep the single serial parser.
 * --------------------------------------------------------------------------------------- */
static void chunk_reset(chunk_t *c) { c->len = 0; c->n = 0; c->status = 0; c->has_cm = 0; c->start = c->next = 0; }

static size_t chunk_text(chunk_t *c, size_t need)          /* room for need more bytes; returns the offset they start at */
{ if (c->len+need > c->cap)
    { c->cap = (c->len+need)*2+4096;
      c->arena = xrealloc(c->arena,c->cap);
    }
  size_t o = c->len;
  c->len += need;
  return o;
}

static void chunk_add(chunk_t *c, const fastx_t *X, int rlen)
{ if (c->n >= c->n_cap)
    { c->n_cap = c->n_cap*2+256;
      c->hoff = xrealloc(c->hoff,sizeof(size_t)*(size_t)c->n_cap); c->soff = xrealloc(c->soff,sizeof(size_t)*(size_t)c->n_cap);
      c->hlen = xrealloc(c->hlen,sizeof(int32_t)*(size_t)c->n_cap); c->rlen = xrealloc(c->rlen,sizeof(int32_t)*(size_t)c->n_cap);
      c->nocm = xrealloc(c->nocm,(size_t)c->n_cap);
    }
  const char *cm = X->have_comment ? X->comment.s : "(null)";      /* src/ClassPro.c:188: printf("%s") of a NULL comment */
  const size_t nl = X->name.l, cl = strlen(cm), hl = nl+cl+2;
  const int j = c->n++;
  size_t o = chunk_text(c,hl+1+(size_t)rlen+1);
  char *h = c->arena+o;
  h[0] = '@'; memcpy(h+1,X->name.s,nl); h[1+nl] = ' '; memcpy(h+2+nl,cm,cl); h[hl] = 0;
  memcpy(h+hl+1,X->seq.s,(size_t)rlen); h[hl+1+(size_t)rlen] = 0;
  c->hoff[j] = o; c->hlen[j] = (int32_t)hl; c->soff[j] = o+hl+1; c->rlen[j] = rlen; c->nocm[j] = !X->have_comment;
}

static void chunk_keep_comment(chunk_t *c, const fastx_t *X)
{ c->has_cm = X->have_comment;
  if (c->has_cm)
    { size_t cl = strlen(X->comment.s)+1;
      c->cm_off = chunk_text(c,cl);
      memcpy(c->arena+c->cm_off,X->comment.s,cl);
    }
}

/* records whose marker lies in [start,limit) of the mapped file */
static void parse_range(fastx_t *X, chunk_t *c, const uint8_t *map, int64_t flen, int64_t start, int64_t limit)
{ chunk_reset(c);
  fx_set_memory(X,map,start,flen);
  X->have_comment = 0;
  c->start = start;
  for (;;)
    { int64_t m = fx_next_marker(X);
      if (m >= limit) { c->next = m; break; }
      int rlen = fx_read(X);
      if (rlen < 0) { c->next = flen; if (rlen == -2) c->status = -2; break; }
      chunk_add(c,X,rlen);
    }
  chunk_keep_comment(c,X);
}

/* first line start at or after p that looks like the start of a record */
static int64_t guess_record_start(const uint8_t *map, int64_t flen, int64_t p)
{ if (p <= 0) return 0;
  while (p < flen)
    { const uint8_t *q = memchr(map+p-1,'\n',(size_t)(flen-p+1));
      if (q == NULL) return flen;
      p = (int64_t)(q-map)+1;
      if (p >= flen) return flen;
      if (map[p] == '>') return p;
      if (map[p] == '@')
        { /* a FASTQ header has its '+' line two lines down (unwrapped records); a quality line that
             begins with '@' has the next header there */
          const uint8_t *l1 = memchr(map+p,'\n',(size_t)(flen-p));
          const uint8_t *l2 = l1 ? memchr(l1+1,'\n',(size_t)(flen-(l1+1-map))) : NULL;
          if (l2 == NULL || l2+1 >= map+flen || l2[1] == '+') return p;
        }
      p++;
    }
  return flen;
}

typedef struct { fastx_t *X; batch_t *b; const uint8_t *map; int64_t flen; int64_t *cut; } parsejob_t;

static void parse_pieces(void *arg, int lo, int hi)
{ parsejob_t *J = arg;
  for (int t = lo; t < hi; t++) parse_range(&J->X[t],&J->b->ch[t],J->map,J->flen,J->cut[t],J->cut[t+1]);
}

typedef struct
  { app_t *A; int64_t id; int eof;
    str_t  carry; int have_carry;          /* last comment seen so far in the stream */
  } rstate_t;

/* the records of b->ch[0..nch) become records first_id.. of the batch; sizes of the device buffers */
static void flatten_chunks(rstate_t *R, batch_t *b, size_t *pseq, size_t *prof, size_t *cls)
{ app_t *A = R->A; const int K = A->P.kmer;
  for (int i = 0; i < b->npatch; i++) free(b->patch[i]);
  b->npatch = 0;
  int total = 0;
  for (int t = 0; t < b->nch; t++) total += b->ch[t].n;
  batch_reserve_records(b,total+1);                       /* no pinned memory yet: CUDA may still be starting */
  for (int t = 0; t < b->nch && !R->eof; t++)
    { chunk_t *c = &b->ch[t];
      for (int j = 0; j < c->n; j++)
        { if (R->id >= A->P.nreads) { R->eof = 1; break; }
          const int rlen = c->rlen[j];
          if (rlen > MAX_READ_LEN)
            die("rlen (%d) > MAX_READ_LEN for FASTX inputs (%d)",rlen,MAX_READ_LEN);
          const int i = b->n_all++;
          b->header[i] = c->arena+c->hoff[j]; b->hlen[i] = c->hlen[j];
          b->seq[i] = c->arena+c->soff[j]; b->rlen_all[i] = rlen;
          if (c->nocm[j] && R->have_carry)
            { /* no comment of its own, and its parser had not seen the one an earlier piece ended with */
              const size_t nl = (size_t)c->hlen[j]-7, cl = R->carry.l;        /* header is "@name (null)" */
              char *h = xmalloc(nl+cl+2);
              memcpy(h,b->header[i],nl+1); memcpy(h+nl+1,R->carry.s,cl+1);
              if (b->npatch >= b->patch_cap)
                { b->patch_cap = b->patch_cap*2+16; b->patch = xrealloc(b->patch,sizeof(char *)*(size_t)b->patch_cap); }
              b->patch[b->npatch++] = h; g_patched++;
              b->header[i] = h; b->hlen[i] = (int32_t)(nl+1+cl);
            }
          if (rlen >= K)
            { int part; int64_t off, len;
              prof_range(&A->P,R->id,&part,&off,&len);
              b->slot_of[i] = b->n++;
              *pseq += (size_t)(rlen+3)/4; *prof += (size_t)len; *cls += (size_t)rlen;
              b->kmers += rlen-K+1;
            }
          else b->slot_of[i] = -1;
          R->id++;
        }
      if (R->eof) break;
      if (c->status == -2) die("%s: truncated quality string in %s",PROG,A->src_path);
      if (c->has_cm)
        { const char *cm = c->arena+c->cm_off; size_t cl = strlen(cm);
          R->carry.l = 0; str_reserve(&R->carry,cl); memcpy(R->carry.s,cm,cl+1); R->carry.l = cl;
          R->have_carry = 1;
        }
    }
}

static void batch_chunks(batch_t *b, int nch)
{ if (nch > b->nch)
    { b->ch = xrealloc(b->ch,sizeof(chunk_t)*(size_t)nch);
      memset(b->ch+b->nch,0,sizeof(chunk_t)*(size_t)(nch-b->nch));
      b->nch = nch;
    }
}

static int env_int(const char *name, int dflt)
{ const char *e = getenv(name);
  return (e && atoi(e) > 0) ? atoi(e) : dflt;
}

static void *reader_main(void *arg)
{ app_t *A = arg;
  pool_t pool;
  pool_start(&pool,A->nthreads);
  rstate_t R; memset(&R,0,sizeof(R)); R.A = A;

  /* source: a mapped plain file (parallel parse) or a gz stream (one parser) */
  const uint8_t *map = NULL; int64_t flen = 0, pos = 0;
  int npiece = env_int("CPG_PARSE_PIECES",A->nthreads);
  int fd = getenv("CPG_SERIAL_IO") ? -1 : open(A->src_path,O_RDONLY);
  if (fd >= 0)
    { struct stat st; uint8_t magic[2] = { 0, 0 };
      if (fstat(fd,&st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && pread(fd,magic,2,0) == 2
          && !(magic[0] == 0x1f && magic[1] == 0x8b))
        { void *m = mmap(NULL,(size_t)st.st_size,PROT_READ,MAP_PRIVATE,fd,0);
          if (m != MAP_FAILED) { map = m; flen = st.st_size; }
        }
      close(fd);
    }
  fastx_t *X = xmalloc(sizeof(fastx_t)*(size_t)(map ? npiece : 1));
  memset(X,0,sizeof(fastx_t)*(size_t)(map ? npiece : 1));
  int64_t *cut = xmalloc(sizeof(int64_t)*(size_t)(npiece+1));
  if (map == NULL)
    { X[0].f = gzopen(A->src_path,"r");
      if (X[0].f == NULL) die("%s: Cannot open %s",PROG,A->src_path);
      gzbuffer(X[0].f,1<<20);
      X[0].buf = xmalloc(FX_BUF);
    }
  else if (A->verbose) fprintf(stderr,"    Parsing %s with %d threads\n",A->src_path,A->nthreads);
  /* a window of the file holds about one batch: 1 byte per base in FASTA, 2 in FASTQ */
  int64_t window = A->batch_bases;
  if (map) { fastx_t P0; fx_set_memory(&P0,map,0,flen); int64_t m = fx_next_marker(&P0); if (m < flen && map[m] == '@') window *= 2; }

  while (!R.eof && R.id < A->P.nreads)
    { batch_t *b = q_pop(&A->q_free);
      if (b == NULL) break;
      const double t_r0 = now_s();
      b->first_id = R.id; b->n_all = 0; b->n = 0; b->kmers = 0; b->seq_bits = 2;
      size_t pseq = 0, prof = 0, cls = 0;
      /* pass 1: parse records, keep header + sequence text */
      if (map)
        { batch_chunks(b,npiece);
          cut[0] = pos;
          for (int t = 1; t <= npiece; t++)
            { int64_t p = pos+(int64_t)((double)window*t/npiece);
              cut[t] = p >= flen ? flen : guess_record_start(map,flen,p);
              if (cut[t] < cut[t-1]) cut[t] = cut[t-1];
            }
          parsejob_t J = { X, b, map, flen, cut };
          pool_run(&pool,parse_pieces,&J,npiece,1);
          for (int t = 1; t < npiece; t++)
            if (b->ch[t-1].next != b->ch[t].start)       /* the guess was not a record start: parse again from the true one */
              { parse_range(&X[t],&b->ch[t],map,flen,b->ch[t-1].next,cut[t+1]); g_reparsed++; }
          { /* the text of the records now lives in the arenas: let go of the pages of the window */
            const int64_t pg = sysconf(_SC_PAGESIZE), a = (pos+pg-1) & ~(pg-1), e = b->ch[npiece-1].next & ~(pg-1);
            if (e > a) madvise((void *)(map+a),(size_t)(e-a),MADV_DONTNEED);
          }
          pos = b->ch[npiece-1].next;
          flatten_chunks(&R,b,&pseq,&prof,&cls);
          if (pos >= flen) R.eof = 1;
        }
      else
        { batch_chunks(b,1);
          chunk_t *c = &b->ch[0];
          chunk_reset(c);
          int64_t bases = 0, n = 0;
          while (bases < A->batch_bases && R.id+n < A->P.nreads)
            { int rlen = fx_read(&X[0]);
              if (rlen < 0)
                { if (rlen == -2) c->status = -2;
                  R.eof = 1; break;
                }
              if (rlen > MAX_READ_LEN)
                die("rlen (%d) > MAX_READ_LEN for FASTX inputs (%d)",rlen,MAX_READ_LEN);
              chunk_add(c,&X[0],rlen);
              bases += rlen; n++;
            }
          chunk_keep_comment(c,&X[0]);
          int eof = R.eof; R.eof = 0;
          flatten_chunks(&R,b,&pseq,&prof,&cls);
          R.eof |= eof;
        }
      if (b->n_all == 0) { g_t_reader += now_s()-t_r0; q_push(&A->q_free,b); break; }     /* nothing left (or an empty input) */
      /* pass 2: pack + fetch profiles into pinned memory */
      if (g_tl_first == 0.) g_tl_first = now_s();
      batch_reserve(b,b->n_all,pseq+16,prof+16,cls+16);
      if (g_tl_pinned == 0.) g_tl_pinned = now_s();
      int64_t so = 0, po = 0, co = 0;
      int k = 0, bad = 0;
      for (int i = 0; i < b->n_all; i++)
        { if (b->slot_of[i] < 0) continue;
          const int rlen = b->rlen_all[i];
          int part; int64_t off, len;
          prof_range(&A->P,b->first_id+i,&part,&off,&len);
          b->seq_off[k] = so; b->prof_off[k] = po; b->cls_off[k] = co; b->rlen[k] = rlen;
          so += (rlen+3)/4; co += rlen; po += len;
          k++;
        }
      b->seq_off[k] = so; b->prof_off[k] = po; b->cls_off[k] = co;
      { packjob_t J = { A, b, 0 };
        pool_run(&pool,pack_records,&J,b->n_all,64);
        bad = J.bad;
      }
      if (bad)
        { /* a character outside ACGT: ship the raw bytes, compared as the reference compares them */
          size_t raw = 0;
          for (int i = 0; i < b->n_all; i++) if (b->slot_of[i] >= 0) raw += (size_t)b->rlen_all[i];
          batch_reserve(b,b->n_all,raw+16,0,0);
          so = 0; k = 0;
          for (int i = 0; i < b->n_all; i++)
            { if (b->slot_of[i] < 0) continue;
              b->seq_off[k++] = so;
              memcpy(b->pseq+so,b->seq[i],(size_t)b->rlen_all[i]);
              so += b->rlen_all[i];
            }
          b->seq_off[k] = so;
          b->seq_bits = 8;
        }
      g_t_reader += now_s()-t_r0;
      q_push(&A->q_ready,b);
    }
  pthread_mutex_lock(&A->mu);
  g_tl_reader = now_s();
  A->reader_done = 1; A->total_reads = R.id;
  pthread_cond_broadcast(&A->cv);
  pthread_mutex_unlock(&A->mu);
  q_close(&A->q_ready);
  pool_stop(&pool);
  if (map) munmap((void *)map,(size_t)flen); else gzclose(X[0].f);
  return NULL;
}

typedef struct { app_t *A; int device; } gpu_arg_t;

static void check_batch_status(batch_t *b)
{ for (int i = 0; i < b->n_all; i++)
    { int k = b->slot_of[i];
      if (k >= 0 && (b->status[k] & (1|2|4|8|64)))
        { if (b->status[k] & 1)
            die("Read %lld: rlen (%d) != plen+Km1",(long long)(b->first_id+i+1),b->rlen_all[i]);
          die("Read %lld: %s",(long long)(b->first_id+i+1),cpg_status_string(b->status[k]));
        }
    }
}

/* one worker per GPU: two batches in flight (slots 0/1), so the copies of one overlap the kernels
   of the other; finished batches go to the writer, which restores read order.  A worker never
   sleeps on the queue with a batch in flight: the writer may be waiting for exactly that batch while
   every other batch sits, out of order, behind it (with several workers and batches that take no
   time that is a deadlock -- found by the parser fuzz of the CPU suite). */
static void *gpu_main(void *arg)
{ gpu_arg_t *G = arg; app_t *A = G->A;
  cpg_ctx *ctx = NULL;
  const double t_c0 = now_s();
  if (cpg_create(&ctx,G->device,A->model,0,0) != CPG_OK)
    die("%s: %s",PROG,cpg_last_error(NULL));
  if (cpg_set_result_mode(ctx,CPG_RESULT_INTERVALS) != CPG_OK) die("%s: %s",PROG,cpg_last_error(ctx));
  if (G->device == 0) { g_t_gpu_create = now_s()-t_c0; g_tl_ctx = now_s(); }
  batch_t *fly[2] = { NULL, NULL };
  int slot = 0;
  for (;;)
    { const int in_flight = fly[slot ^ 1] != NULL;
      batch_t *b = in_flight ? q_trypop(&A->q_ready) : q_pop(&A->q_ready);
      const double t_g0 = now_s();
      if (b != NULL && b->n > 0)
        { cpg_batch in = { b->n, b->seq_bits, b->pseq, b->seq_off, b->rlen, b->prof, b->prof_off };
          if (cpg_submit(ctx,slot,&in) != CPG_OK) die("%s: %s",PROG,cpg_last_error(ctx));
          if (G->device < MAX_GPUS) { g_gpu_batches[G->device]++; g_gpu_kmers[G->device] += b->kmers; }
        }
      batch_t *done = fly[slot ^ 1];          /* the batch submitted before this one */
      fly[slot ^ 1] = NULL;
      if (b != NULL) fly[slot] = b;
      if (done != NULL)
        { if (done->n > 0)
            { batch_reserve_ivl(done,cpg_intervals_bound(ctx,slot ^ 1));
              cpg_result_ivl out = { done->ivl, done->ivl_cap, done->ivl_at, done->ivl_n, done->status, 0 };
              int rc = cpg_collect_intervals(ctx,slot ^ 1,&out);
              if (rc == CPG_EREAD) check_batch_status(done);
              else if (rc != CPG_OK) die("%s: %s",PROG,cpg_last_error(ctx));
            }
          pthread_mutex_lock(&A->mu);
          done->next = A->done; A->done = done;
          pthread_cond_broadcast(&A->cv);
          pthread_mutex_unlock(&A->mu);
        }
      if (G->device == 0) g_t_gpu_busy += now_s()-t_g0;
      if (b == NULL)
        { if (!in_flight) break;              /* queue closed and nothing in flight */
          continue;                           /* nothing was ready: the batch in flight is finished, now wait */
        }
      slot ^= 1;
    }
  if (G->device == 0) g_tl_collect = now_s();
  cpg_destroy(ctx);
  return NULL;
}

/* ---------------------------------------------------------------------------------------
 *  Writer: records of finished batches in read order (src/ClassPro.c:289).  The byte offset of
 *  every record of a batch is known before anything is formatted, so the file is grown by the
 *  size of the batch, that range is mapped, and the -T pool threads format disjoint record ranges
 *  straight into the page cache (one write() stream copies at ~0.7 GB/s on the test machine, four
 *  mapping threads at ~2.5 GB/s).  Where the file cannot be mapped the same formatter fills a heap
 *  buffer that is written with write().
 * --------------------------------------------------------------------------------------- */
typedef struct
  { batch_t *b; char *dst;               /* dst + b->out_off[i] = first byte of record i */
    const char *carry; int carry_len;    /* class string of the last classified read of earlier batches */
    int K;
  } fmtjob_t;

/* class string of classified record k of the batch (its rlen characters), from its interval table */
static void expand_class(const batch_t *b, int K, int k, char *out)
{ cpg_expand_intervals(K,b->rlen[k],b->ivl+b->ivl_at[k],b->ivl_n[k],out); }

static void format_records(void *arg, int lo, int hi)
{ fmtjob_t *J = arg; batch_t *b = J->b;
  for (int i = lo; i < hi; i++)
    { char *p = J->dst+b->out_off[i];
      const int rlen = b->rlen_all[i], k = b->slot_of[i], hl = b->hlen[i];
      memcpy(p,b->header[i],(size_t)hl); p += hl; *p++ = '\n';
      memcpy(p,b->seq[i],(size_t)rlen); p += rlen;
      memcpy(p,"\n+\n",3); p += 3;
      if (k >= 0) { expand_class(b,J->K,k,p); p += rlen; }
      else
        { /* a read shorter than K: the class string of the last classified read again, whole,
             right-aligned in at least rlen columns ("%*s" is a minimum width, src/ClassPro.c:215) */
          const int l = b->lastc[i];
          const int sl = l >= 0 ? b->rlen_all[l] : J->carry_len;
          if (rlen > sl) { memset(p,' ',(size_t)(rlen-sl)); p += rlen-sl; }
          if (l >= 0) expand_class(b,J->K,b->slot_of[l],p); else memcpy(p,J->carry,(size_t)sl);
          p += sl;
        }
      *p++ = '\n';
    }
}

static void *writer_main(void *arg)
{ app_t *A = arg;
  const int K = A->P.kmer;
  int fd = open(A->out_path,O_RDWR|O_CREAT|O_TRUNC,0666);
  if (fd < 0) die("Cannot open %s",A->out_path);
  int use_map = getenv("CPG_SERIAL_IO") == NULL;
  const int64_t page = sysconf(_SC_PAGESIZE);
  pool_t pool;
  pool_start(&pool,A->nthreads);
  char *rasgn = xmalloc(MAX_READ_LEN+2);
  int   rasgn_len = K-1;
  for (int i = 0; i < K-1; i++) rasgn[i] = 'N';
  char *heap = NULL; size_t heap_cap = 0;
  int64_t cur = 0;                                /* bytes written so far */
  for (;;)
    { pthread_mutex_lock(&A->mu);
      batch_t *b = NULL;
      for (;;)
        { batch_t **pp = &A->done;
          while (*pp && (*pp)->first_id != A->next_id) pp = &(*pp)->next;
          if (*pp) { b = *pp; *pp = b->next; break; }
          if (A->reader_done && A->next_id >= A->total_reads) break;
          pthread_cond_wait(&A->cv,&A->mu);
        }
      pthread_mutex_unlock(&A->mu);
      if (b == NULL) break;
      const double t_w0 = now_s();
      /* where every record goes */
      int last = -1, sl = rasgn_len; int64_t o = 0;
      for (int i = 0; i < b->n_all; i++)
        { const int rlen = b->rlen_all[i];
          b->out_off[i] = o; b->lastc[i] = last;
          if (b->slot_of[i] >= 0) { last = i; sl = rlen; o += b->hlen[i]+2*(int64_t)rlen+5; }
          else o += b->hlen[i]+(int64_t)rlen+5+(rlen > sl ? rlen : sl);
        }
      b->out_off[b->n_all] = o;
      fmtjob_t J = { b, NULL, rasgn, rasgn_len, K };
      char *m = MAP_FAILED; int64_t mstart = cur & ~(page-1);
      if (use_map)
        { int rc = posix_fallocate(fd,cur,o);
          if (rc == ENOSPC || rc == EFBIG) die("Cannot write %s",A->out_path);
          if (rc != 0 && ftruncate(fd,cur+o) != 0) use_map = 0;
          if (use_map) m = mmap(NULL,(size_t)(cur+o-mstart),PROT_READ|PROT_WRITE,MAP_SHARED,fd,mstart);
          if (m == MAP_FAILED)
            { use_map = 0;
              if (ftruncate(fd,cur) != 0 || lseek(fd,cur,SEEK_SET) < 0) die("Cannot write %s",A->out_path);
            }
        }
      if (m != MAP_FAILED) J.dst = m+(cur-mstart);
      else
        { if ((size_t)o > heap_cap) { free(heap); heap_cap = (size_t)o+(size_t)o/4; heap = xmalloc(heap_cap); }
          J.dst = heap;
        }
      pool_run(&pool,format_records,&J,b->n_all,32);
      if (m != MAP_FAILED) munmap(m,(size_t)(cur+o-mstart));
      else
        for (int64_t w = 0; w < o; )
          { ssize_t r = write(fd,heap+w,(size_t)(o-w));
            if (r <= 0) die("Cannot write %s",A->out_path);
            w += r;
          }
      cur += o;
      if (last >= 0)
        { rasgn_len = b->rlen_all[last];
          expand_class(b,K,b->slot_of[last],rasgn);
        }
      g_t_writer += now_s()-t_w0;
      pthread_mutex_lock(&A->mu);
      A->next_id = b->first_id+b->n_all;
      A->kmers += b->kmers;
      pthread_mutex_unlock(&A->mu);
      q_push(&A->q_free,b);
    }
  q_close(&A->q_free);
  pool_stop(&pool);
  if (close(fd) != 0) die("Cannot write %s",A->out_path);
  g_tl_writer = now_s();
  free(rasgn); free(heap);
  return NULL;
}

/* ---------------------------------------------------------------------------------------
 *  Command line (src/ClassPro.c:348-501) and timing lines (src/benchmark.c:12-96)
 * --------------------------------------------------------------------------------------- */
static struct timespec T0;
static struct rusage   R0;

static double now_s(void)
{ struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t);
  return (t.tv_sec-T0.tv_sec)+(t.tv_nsec-T0.tv_nsec)*1e-9;
}

static void time_line(FILE *f, const char *what)
{ struct timespec t; struct rusage r;
  clock_gettime(CLOCK_MONOTONIC,&t); getrusage(RUSAGE_SELF,&r);
  double wall = (t.tv_sec-T0.tv_sec)+(t.tv_nsec-T0.tv_nsec)*1e-9;
  double user = (r.ru_utime.tv_sec-R0.ru_utime.tv_sec)+(r.ru_utime.tv_usec-R0.ru_utime.tv_usec)*1e-6;
  double sys  = (r.ru_stime.tv_sec-R0.ru_stime.tv_sec)+(r.ru_stime.tv_usec-R0.ru_stime.tv_usec)*1e-6;
  fprintf(f,"%s  %.3f (s.ms) user  %.3f (s.ms) sys  %.3f (s.ms) wall  %.1f%%  %ld MB max rss\n",
          what,user,sys,wall,wall > 0 ? 100.*(user+sys)/wall : 0.,r.ru_maxrss/1024);
}

static const char *EXT[10] = { ".db",".dam",".fastq",".fasta",".fq",".fa",".fastq.gz",".fasta.gz",".fq.gz",".fa.gz" };

int main(int argc, char **argv)
{ clock_gettime(CLOCK_MONOTONIC,&T0); getrusage(RUSAGE_SELF,&R0);
  app_t *A = calloc(1,sizeof(app_t));
  A->nthreads = 4; A->read_len = 20000; A->ngpus = 0; A->batch_bases = 64000000;
  int npos = 0; char *pos = NULL;
  for (int i = 1; i < argc; i++)
    { char *a = argv[i];
      if (a[0] != '-') { pos = a; npos++; continue; }
      char *e = NULL;
      switch (a[1])
        { case 'T': A->nthreads = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0) die("%s: -T '%s' argument is not an integer",PROG,a+2);
                    if (A->nthreads <= 0) die("%s: Number of threads must be positive (%d)",PROG,A->nthreads);
                    break;
          case 'c': A->cov = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0) die("%s: -c '%s' argument is not an integer",PROG,a+2);
                    if (A->cov < 0) die("%s: Estimated k-mer coverage must be non-negative (%d)",PROG,A->cov);
                    break;
          case 'r': A->read_len = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0) die("%s: -r '%s' argument is not an integer",PROG,a+2);
                    if (A->read_len <= 0) die("%s: Average read length must be positive (%d)",PROG,A->read_len);
                    break;
          case 'G': A->ngpus = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0 || A->ngpus <= 0) die("%s: -G needs a positive integer",PROG);
                    break;
          case 'B': { long mb = strtol(a+2,&e,10);
                      if (*e || a[2] == 0 || mb <= 0) die("%s: -B needs a positive integer (megabases)",PROG);
                      A->batch_bases = mb*1000000L;
                    }
                    break;
          case 'N': A->fk_root = a+2; break;
          case 'P': break;                                   /* no part files any more */
          case 'M': A->model_path = a+2; break;
          default:
            for (char *p = a+1; *p; p++)
              { if (*p == 'v') A->verbose = 1;
                else if (*p == 's') A->find_seeds = 1;
                else die("%s: -%c is an illegal option",PROG,*p);
              }
        }
    }
  { const char *e = getenv("CPG_BATCH_BASES");               /* test knob: batches smaller than -B1 */
    if (e && atol(e) > 0) A->batch_bases = atol(e);
  }
  if (npos < 1) { fprintf(stderr,"Usage: %s %s\n",PROG,USAGE); return 1; }
  if (npos != 1) die("Currently only single file is accepted for FASTX input");
  if (A->model_path) die("%s: -M <model_path> is not supported (the polynomial fit needs GSL, absent from the reference tree)",PROG);
  if (A->find_seeds)
    fprintf(stderr,"%s: -s has no effect on the .class output and is not implemented for FASTX inputs; ignored\n",PROG);

  if (A->verbose) fprintf(stderr,"Info about inputs:\n");
  /* resolve <source> by trying the extensions in the reference's order (src/ClassPro.c:411-430) */
  char dir[4096], base[1024], root[1024], path[8192];
  split_path(pos,dir,sizeof(dir),base,sizeof(base));
  int idx;
  for (idx = 0; idx < 10; idx++)
    { size_t bl = strlen(base), el = strlen(EXT[idx]);
      snprintf(root,sizeof(root),"%s",base);
      if (bl > el && strcasecmp(base+bl-el,EXT[idx]) == 0) root[bl-el] = 0;
      snprintf(path,sizeof(path),"%s/%s%s",dir,root,EXT[idx]);
      int f = open(path,O_RDONLY);
      if (f >= 0) { close(f); break; }
    }
  if (idx == 10) die("Cannot open %s as a .db|.dam or .f{ast}[aq][.gz] file",pos);
  if (idx <= 1) die("%s: .db/.dam inputs are not supported by this build (DAZZ_DB is out of scope)",PROG);
  A->src_path = strdup(path);
  char fk[8192], outp[8192];
  if (A->fk_root == NULL) { snprintf(fk,sizeof(fk),"%s/%s",dir,root); A->fk_root = fk; }
  snprintf(outp,sizeof(outp),"%s/%s.class",dir,root);
  A->out_path = outp;
  if (A->verbose)
    { fprintf(stderr,"    # of sequence files   = %d\n",1);
      fprintf(stderr,"    First (path,root,ext) = (%s, %s, %s)\n",dir,root,EXT[idx]);
      fprintf(stderr,"    FASTK outputs' root   = %s\n",A->fk_root);
      fprintf(stderr,"    Output .class file    = %s/%s.class\n",dir,root);
    }

  if (profidx_open(&A->P,A->fk_root)) die("%s: Cannot open %s.prof",PROG,A->fk_root);
  if (A->verbose) fprintf(stderr,"    Total # of reads      = %lld\n",(long long)A->P.nreads);

  A->model = xmalloc(sizeof(cpg_model));
  int rc = cpg_model_load(A->model,A->fk_root,A->cov,A->read_len,A->verbose);
  if (rc == CPG_EIO) die("%s: Cannot open %s.hist",PROG,A->fk_root);
  if (rc != CPG_OK) exit(1);
  A->model->kmer = A->P.kmer;
  if (A->verbose) fprintf(stderr,"Error model not specified. Using the default error model.\n");
  g_tl_model = now_s();

  const char *env = getenv("CLASSPRO_GPUS");
  if (A->ngpus == 0 && env && atoi(env) > 0) A->ngpus = atoi(env);
  /* The CUDA runtime initialises every VISIBLE device when it starts (about 0.2 s each on an 8-GPU box),
     and nothing can be allocated before that: a run that was asked for fewer GPUs hides the others from it
     (before the first CUDA call; an explicit CUDA_VISIBLE_DEVICES of the caller is left alone). */
  if (A->ngpus > 0 && A->ngpus <= MAX_GPUS && getenv("CUDA_VISIBLE_DEVICES") == NULL)
    { char vis[8*MAX_GPUS+8]; int o = 0;
      for (int g = 0; g < A->ngpus; g++) o += snprintf(vis+o,sizeof(vis)-(size_t)o,g ? ",%d" : "%d",g);
      setenv("CUDA_VISIBLE_DEVICES",vis,1);
    }
  int ndev = cpg_device_count();
  if (ndev <= 0) die("%s: no CUDA device found: this program has no CPU fallback",PROG);
  if (A->ngpus == 0 || A->ngpus > ndev) A->ngpus = ndev;
  if (A->ngpus > MAX_GPUS) A->ngpus = MAX_GPUS;
  if (A->verbose)
    fprintf(stderr,"Classifying %d-mers on %d GPU%s...\n",A->P.kmer,A->ngpus,A->ngpus > 1 ? "s" : "");

  q_init(&A->q_free); q_init(&A->q_ready);
  pthread_mutex_init(&A->mu,NULL); pthread_cond_init(&A->cv,NULL);
  const int nbatch = 2*A->ngpus+2;
  for (int i = 0; i < nbatch; i++) q_push(&A->q_free,calloc(1,sizeof(batch_t)));

  pthread_t rd, wr, *gp = xmalloc(sizeof(pthread_t)*(size_t)A->ngpus);
  gpu_arg_t *ga = xmalloc(sizeof(gpu_arg_t)*(size_t)A->ngpus);
  pthread_create(&wr,NULL,writer_main,A);
  for (int g = 0; g < A->ngpus; g++) { ga[g].A = A; ga[g].device = g; pthread_create(&gp[g],NULL,gpu_main,&ga[g]); }
  pthread_create(&rd,NULL,reader_main,A);
  pthread_join(rd,NULL);
  for (int g = 0; g < A->ngpus; g++) pthread_join(gp[g],NULL);
  pthread_mutex_lock(&A->mu); pthread_cond_broadcast(&A->cv); pthread_mutex_unlock(&A->mu);
  pthread_join(wr,NULL);

  if (A->verbose)
    { time_line(stderr,"Resources for phase:");
      fprintf(stderr,"Classified %lld k-mers of %lld reads\n",(long long)A->kmers,(long long)A->total_reads);
      fprintf(stderr,"    stage seconds: reader %.3f (parse+pack+profile read), GPU0 context %.3f, GPU0 submit/collect %.3f, writer %.3f\n",
              g_t_reader,g_t_gpu_create,g_t_gpu_busy,g_t_writer);
      fprintf(stderr,"    per GPU (batches/k-mers):");
      for (int g = 0; g < A->ngpus && g < MAX_GPUS; g++) fprintf(stderr," %ld/%lld",g_gpu_batches[g],g_gpu_kmers[g]);
      fprintf(stderr,"\n");
      fprintf(stderr,"    parser: %ld pieces parsed again from their true start, %ld headers completed with a carried comment\n",
              g_reparsed,g_patched);
      fprintf(stderr,"    timeline (s): model %.3f, first batch parsed %.3f, first pinned buffer %.3f, GPU0 context %.3f, "
                     "reader done %.3f, last collect %.3f, writer done %.3f\n",
              g_tl_model,g_tl_first,g_tl_pinned,g_tl_ctx,g_tl_reader,g_tl_collect,g_tl_writer);
      time_line(stderr,"Total Resources:");
    }
  return 0;
}
