/*******************************************************************************************
 *  cpg_hostio.h -- host-side file plumbing shared by the command-line programs of this
 *  repository (ClassPro, prof2class, class2acc): error exit, a FASTA/FASTQ(.gz) record reader with
 *  the semantics of src/kseq.h:177-218, and the FastK profile index (src/libfastk.c:1267-1361,
 *  1444-1454).  Static functions: include it once per program, after defining PROG.
 *******************************************************************************************/
#ifndef CPG_HOSTIO_H
#define CPG_HOSTIO_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdarg.h>
#include <ctype.h>
#include <fcntl.h>
#include <unistd.h>
#include <strings.h>
#include <zlib.h>

__attribute__((noreturn,format(printf,1,2))) static void die(const char *fmt, ...)
{ va_list ap; va_start(ap,fmt); vfprintf(stderr,fmt,ap); va_end(ap); fputc('\n',stderr); exit(1); }

static void *xmalloc(size_t n)
{ void *p = malloc(n ? n : 1);
  if (p == NULL) die("%s: Out of memory",PROG);
  return p;
}
static void *xrealloc(void *p, size_t n)
{ p = realloc(p,n ? n : 1);
  if (p == NULL) die("%s: Out of memory",PROG);
  return p;
}

/* ---------------------------------------------------------------------------------------
 *  FASTA/FASTQ(.gz) stream with the record semantics of src/kseq.h:177-218
 * --------------------------------------------------------------------------------------- */
typedef struct { char *s; size_t l, m; } str_t;

typedef struct
  { gzFile  f;
    uint8_t *buf;
    int64_t  beg, end;              /* window of buf; int64: in memory mode buf is a whole mapped file */
    int      eof;
    int      last_char;
    str_t    name, comment, seq, qual;
    int      have_comment;          /* comment.s non-NULL in kseq terms */
  } fastx_t;

#define FX_BUF (1<<20)

static int fx_getc(fastx_t *x)
{ if (x->beg >= x->end)
    { if (x->eof) return -1;
      x->beg = 0;
      x->end = gzread(x->f,x->buf,FX_BUF);
      if (x->end <= 0) { x->eof = 1; x->end = 0; return -1; }
    }
  return x->buf[x->beg++];
}

static void str_reserve(str_t *s, size_t extra)
{ if (s->l+extra+1 > s->m)
    { s->m = (s->l+extra+1)*2;
      s->s = xrealloc(s->s,s->m);
    }
}

/* mode 0: stop at any isspace(); mode 1: stop at '\n'.  Returns -1 at EOF with nothing read. */
static int fx_getuntil(fastx_t *x, int line_mode, str_t *s, int *dret, int append)
{ int got = 0;
  if (dret) *dret = 0;
  if (!append) s->l = 0;
  for (;;)
    { if (x->beg >= x->end)
        { if (x->eof) break;
          x->beg = 0;
          x->end = gzread(x->f,x->buf,FX_BUF);
          if (x->end <= 0) { x->eof = 1; x->end = 0; break; }
        }
      int64_t i = x->beg;
      if (line_mode) { uint8_t *q = memchr(x->buf+i,'\n',(size_t)(x->end-i)); i = q ? (int64_t)(q-x->buf) : x->end; }
      else while (i < x->end && !isspace(x->buf[i])) i++;
      str_reserve(s,(size_t)(i-x->beg));
      got = 1;
      memcpy(s->s+s->l,x->buf+x->beg,(size_t)(i-x->beg));
      s->l += (size_t)(i-x->beg);
      x->beg = i+1;
      if (i < x->end) { if (dret) *dret = x->buf[i]; break; }
    }
  if (!got && x->eof) return -1;
  str_reserve(s,0);
  if (line_mode && s->l > 1 && s->s[s->l-1] == '\r') s->l--;
  s->s[s->l] = 0;
  return (int)s->l;
}

/* Memory mode: the stream is the byte range [pos,len) of mem (a mapped file), never refilled.
   Several fastx_t can parse different ranges of the same mapping at the same time. */
__attribute__((unused)) static void fx_set_memory(fastx_t *x, const uint8_t *mem, int64_t pos, int64_t len)
{ x->f = NULL; x->buf = (uint8_t *)mem; x->beg = pos; x->end = len; x->eof = 1; x->last_char = 0; }

/* Memory mode: position of the record marker ('>' or '@') the next fx_read starts from (= end at the
   end of the stream).  With the comment carried from earlier records this is the whole parser state
   between two records (src/kseq.h:181-186: either the marker was already consumed as the character
   that ended the previous sequence, or the reader skips to the next one). */
__attribute__((unused)) static int64_t fx_next_marker(const fastx_t *x)
{ if (x->last_char) return x->beg-1;
  int64_t i = x->beg;
  while (i < x->end && x->buf[i] != '>' && x->buf[i] != '@') i++;
  return i < x->end ? i : x->end;
}

/* >= 0 sequence length, -1 end of file, -2 truncated quality string */
static int fx_read(fastx_t *x)
{ int c;
  if (x->last_char == 0)
    { do c = fx_getc(x); while (c >= 0 && c != '>' && c != '@');
      if (c < 0) return -1;
      x->last_char = c;
    }
  x->seq.l = 0; x->qual.l = 0;
  if (fx_getuntil(x,0,&x->name,&c,0) < 0) return -1;
  if (c != '\n') { fx_getuntil(x,1,&x->comment,NULL,0); x->have_comment = 1; }
  str_reserve(&x->seq,256);
  while ((c = fx_getc(x)) >= 0 && c != '>' && c != '+' && c != '@')
    { if (c == '\n') continue;
      str_reserve(&x->seq,1);
      x->seq.s[x->seq.l++] = (char)c;
      fx_getuntil(x,1,&x->seq,NULL,1);
    }
  if (c == '>' || c == '@') x->last_char = c;
  str_reserve(&x->seq,0);
  x->seq.s[x->seq.l] = 0;
  if (c != '+') return (int)x->seq.l;
  while ((c = fx_getc(x)) >= 0 && c != '\n');          /* rest of the '+' line */
  if (c < 0) return -2;
  x->qual.l = 0;
  while (fx_getuntil(x,1,&x->qual,NULL,1) >= 0 && x->qual.l < x->seq.l);
  x->last_char = 0;
  if (x->qual.l != x->seq.l) return -2;
  return (int)x->seq.l;
}

/* ---------------------------------------------------------------------------------------
 *  FastK profile index (format of src/libfastk.c:1267-1361)
 * --------------------------------------------------------------------------------------- */
typedef struct
  { int      kmer, nparts;
    int64_t  nreads;
    int64_t *index;         /* [nreads+1]: index[i+1] = end offset of read i inside its part */
    int64_t *nbase;         /* [nparts]: reads before the end of part p */
    int     *fd;            /* [nparts] */
  } profidx_t;

static void split_path(const char *name, char *dir, size_t dn, char *base, size_t bn)
{ const char *sl = strrchr(name,'/');
  if (sl) { snprintf(dir,dn,"%.*s",(int)(sl-name),name); snprintf(base,bn,"%s",sl+1); }
  else    { snprintf(dir,dn,"."); snprintf(base,bn,"%s",name); }
}

__attribute__((unused)) static int profidx_open(profidx_t *P, const char *fk_root)
{ char dir[4096], root[1024], path[8192];
  split_path(fk_root,dir,sizeof(dir),root,sizeof(root));
  size_t rl = strlen(root);
  if (rl > 5 && strcasecmp(root+rl-5,".prof") == 0) root[rl-5] = 0;
  snprintf(path,sizeof(path),"%s/%s.prof",dir,root);
  int f = open(path,O_RDONLY);
  if (f < 0) return 1;
  int32_t smer, nthreads;
  if (read(f,&smer,4) != 4 || read(f,&nthreads,4) != 4) { close(f); return 1; }
  close(f);
  P->kmer = smer; P->nparts = nthreads;
  P->nbase = xmalloc(sizeof(int64_t)*(size_t)nthreads);
  P->fd = xmalloc(sizeof(int)*(size_t)nthreads);
  int64_t total = 0;
  for (int p = 0; p < nthreads; p++)
    { snprintf(path,sizeof(path),"%s/.%s.pidx.%d",dir,root,p+1);
      f = open(path,O_RDONLY);
      if (f < 0) die("Profile part %s is misssing ?",path);
      int32_t k; int64_t n;
      if (read(f,&k,4) != 4 || read(f,&n,8) != 8 || read(f,&n,8) != 8) die("Profile part %s is truncated",path);
      if (k != smer) die("Profile part %s does not have k-mer length matching stub ?",path);
      close(f);
      total += n;
    }
  P->index = xmalloc(sizeof(int64_t)*(size_t)(total+1));
  P->index[0] = 0;
  int64_t nr = 0;
  for (int p = 0; p < nthreads; p++)
    { snprintf(path,sizeof(path),"%s/.%s.pidx.%d",dir,root,p+1);
      f = open(path,O_RDONLY);
      int32_t k; int64_t n;
      if (read(f,&k,4) != 4 || read(f,&n,8) != 8 || read(f,&n,8) != 8) die("Profile part %s is truncated",path);
      size_t want = sizeof(int64_t)*(size_t)n, got = 0;
      while (got < want)
        { ssize_t r = read(f,(char *)(P->index+nr+1)+got,want-got);
          if (r <= 0) die("Profile part %s is truncated",path);
          got += (size_t)r;
        }
      close(f);
      nr += n;
      P->nbase[p] = nr;
      snprintf(path,sizeof(path),"%s/.%s.prof.%d",dir,root,p+1);
      P->fd[p] = open(path,O_RDONLY);
      if (P->fd[p] < 0) die("Profile part %s is misssing ?",path);
    }
  P->nreads = nr;
  return 0;
}

/* byte range of read id inside its part (src/libfastk.c:1444-1454) */
__attribute__((unused)) static void prof_range(const profidx_t *P, int64_t id, int *part, int64_t *off, int64_t *len)
{ int w = 0;
  while (w < P->nparts && id >= P->nbase[w]) w++;
  if (w >= P->nparts) die("Id %lld is out of range [1,%lld]",(long long)id,(long long)P->nbase[P->nparts-1]);
  *part = w;
  *off = (id == 0 || (w > 0 && id == P->nbase[w-1])) ? 0 : P->index[id];
  *len = P->index[id+1]-*off;
}

#endif
