/*******************************************************************************************
 *  cpg_profiler.c -- "FastK -k<K> -t1 -p <source>" for ClassPro on top of libclasspro_b200.so:
 *  writes the files the reference reads before it classifies (FastK itself is
 *  not part of the reference tree):
 *      <root>.hist                          histogram of distinct canonical k-mers by count
 *                                           (reader: src/libfastk.c:51-96; low = 1, high = 32767)
 *      <root>.prof, .<root>.pidx.<p>, .<root>.prof.<p>
 *                                           per-read count profiles, token-compressed
 *                                           (readers: src/libfastk.c:1238-1386, :1414-1562)
 *
 *      profiler [-v] [-k<int(40)>] [-p<parts(1)>] [-N<out_root>] <source>.f[ast][aq][.gz]
 *
 *  The host parses and 2-bit packs the reads; counting (canonical k-mer keys, radix sort, run
 *  lengths scattered back to read positions, histogram) and the profile encoder run on the GPU
 *  (cpg_count_kmers, cpg_encode_profiles).  No CPU fallback.  Reads with a character outside ACGT
 *  are refused: FastK's treatment of them is not pinned by anything in the reference tree.
 *******************************************************************************************/
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "classpro_gpu.h"

static const char *PROG = "profiler";
#include "cpg_hostio.h"

static FILE *xopen(const char *path)
{ FILE *f = fopen(path,"wb");
  if (f == NULL) die("%s: Cannot write %s",PROG,path);
  return f;
}

static void xwrite(const void *p, size_t size, size_t n, FILE *f, const char *path)
{ if (n > 0 && fwrite(p,size,n,f) != n) die("%s: Cannot write %s",PROG,path); }

int main(int argc, char **argv)
{ int K = 40, nparts = 1, verbose = 0; char *pos = NULL, *out_root = NULL; int npos = 0;
  for (int i = 1; i < argc; i++)
    { char *a = argv[i], *e = NULL;
      if (a[0] != '-') { pos = a; npos++; continue; }
      switch (a[1])
        { case 'k': K = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0 || K < 1 || K > 40) die("%s: -k needs an integer in [1,40]",PROG);
                    break;
          case 'p': nparts = (int)strtol(a+2,&e,10);
                    if (*e || a[2] == 0 || nparts < 1) die("%s: -p needs a positive integer",PROG);
                    break;
          case 'N': out_root = a+2; break;
          case 'v': verbose = 1; break;
          default:  die("%s: -%c is an illegal option",PROG,a[1]);
        }
    }
  if (npos != 1) { fprintf(stderr,"Usage: %s [-v] [-k<int(40)>] [-p<parts(1)>] [-N<out_root>] <source>.f[ast][aq][.gz]\n",PROG); return 1; }

  /* the reads, 2-bit packed, every read from a byte boundary (the cpg_batch layout) */
  fastx_t X; memset(&X,0,sizeof(X));
  X.f = gzopen(pos,"r");
  if (X.f == NULL) die("%s: Cannot open %s",PROG,pos);
  gzbuffer(X.f,1<<20);
  X.buf = xmalloc(FX_BUF);
  uint8_t *pseq = NULL; size_t pcap = 0; int64_t *seq_off = NULL; int32_t *rlen = NULL; size_t rcap = 0;
  int64_t nreads = 0, so = 0, bases = 0;
  for (;;)
    { int rl = fx_read(&X);
      if (rl == -1) break;
      if (rl < 0) die("%s: truncated quality string in %s",PROG,pos);
      if ((size_t)nreads+2 > rcap)
        { rcap = rcap*2+1024;
          seq_off = xrealloc(seq_off,sizeof(int64_t)*rcap); rlen = xrealloc(rlen,sizeof(int32_t)*rcap);
        }
      if ((size_t)so+(size_t)rl/4+64 > pcap) { pcap = ((size_t)so+(size_t)rl/4+64)*2; pseq = xrealloc(pseq,pcap); }
      if (cpg_pack_seq(X.seq.s,rl,pseq+so))
        die("%s: read %lld has a character outside ACGT",PROG,(long long)nreads+1);
      seq_off[nreads] = so; rlen[nreads] = rl;
      so += (rl+3)/4; bases += rl; nreads++;
    }
  gzclose(X.f);
  if (seq_off == NULL) { seq_off = xmalloc(sizeof(int64_t)); rlen = xmalloc(sizeof(int32_t)); pseq = xmalloc(64); }
  seq_off[nreads] = so;
  if (nreads > 0x7fffffff) die("%s: too many reads",PROG);
  if (verbose) fprintf(stderr,"%s: %lld reads, %lld bases\n",PROG,(long long)nreads,(long long)bases);

  /* counts, histogram, compressed profiles */
  int64_t *cnt_off = xmalloc(sizeof(int64_t)*(size_t)(nreads+1)), *prof_off = xmalloc(sizeof(int64_t)*(size_t)(nreads+1));
  int64_t *hist = xmalloc(sizeof(int64_t)*32770);
  int64_t nk = 0;
  for (int64_t i = 0; i < nreads; i++) nk += rlen[i] >= K ? rlen[i]-K+1 : 0;
  uint16_t *counts = xmalloc(sizeof(uint16_t)*(size_t)(nk+1));
  uint8_t  *prof = xmalloc(2*(size_t)nk+16);
  if (cpg_count_kmers(0,K,(int32_t)nreads,pseq,seq_off,rlen,cnt_off,counts,hist) != CPG_OK) die("%s: %s",PROG,cpg_count_error());
  if (cpg_encode_profiles(0,(int32_t)nreads,counts,cnt_off,prof,2*nk+16,prof_off) != CPG_OK) die("%s: %s",PROG,cpg_count_error());
  if (verbose)
    fprintf(stderr,"%s: %lld %d-mers, %lld profile bytes (%.3f per k-mer)\n",PROG,(long long)nk,K,(long long)prof_off[nreads],
            nk ? (double)prof_off[nreads]/(double)nk : 0.);

  /* files (layouts: SURVEY A.1) */
  char dir[4096], base[1024], root[1024], path[8192];
  if (out_root) split_path(out_root,dir,sizeof(dir),root,sizeof(root));
  else
    { static const char *EXT[8] = { ".fastq.gz",".fasta.gz",".fq.gz",".fa.gz",".fastq",".fasta",".fq",".fa" };
      split_path(pos,dir,sizeof(dir),base,sizeof(base));
      snprintf(root,sizeof(root),"%s",base);
      for (int i = 0; i < 8; i++)
        { size_t bl = strlen(base), el = strlen(EXT[i]);
          if (bl > el && strcasecmp(base+bl-el,EXT[i]) == 0) { root[bl-el] = 0; break; }
        }
    }
  snprintf(path,sizeof(path),"%s/%s.hist",dir,root);
  FILE *f = xopen(path);
  { int32_t k = K, low = 1, high = 32767;
    xwrite(&k,4,1,f,path); xwrite(&low,4,1,f,path); xwrite(&high,4,1,f,path);
    xwrite(&hist[32768],8,1,f,path); xwrite(&hist[32769],8,1,f,path);
    xwrite(hist+1,8,32767,f,path);
  }
  if (fclose(f) != 0) die("%s: Cannot write %s",PROG,path);
  if (nparts > nreads) nparts = nreads > 0 ? (int)nreads : 1;
  snprintf(path,sizeof(path),"%s/%s.prof",dir,root);
  f = xopen(path);
  { int32_t k = K, np = nparts; xwrite(&k,4,1,f,path); xwrite(&np,4,1,f,path); }
  if (fclose(f) != 0) die("%s: Cannot write %s",PROG,path);
  for (int p = 0; p < nparts; p++)                         /* parts: contiguous read ranges */
    { const int64_t b = nreads*p/nparts, e = nreads*(p+1)/nparts;
      snprintf(path,sizeof(path),"%s/.%s.pidx.%d",dir,root,p+1);
      f = xopen(path);
      int32_t k = K; int64_t first = b, n = e-b;
      xwrite(&k,4,1,f,path); xwrite(&first,8,1,f,path); xwrite(&n,8,1,f,path);
      for (int64_t rd = b; rd < e; rd++)
        { int64_t off = prof_off[rd+1]-prof_off[b];
          xwrite(&off,8,1,f,path);
        }
      if (fclose(f) != 0) die("%s: Cannot write %s",PROG,path);
      snprintf(path,sizeof(path),"%s/.%s.prof.%d",dir,root,p+1);
      f = xopen(path);
      xwrite(prof+prof_off[b],1,(size_t)(prof_off[e]-prof_off[b]),f,path);
      if (fclose(f) != 0) die("%s: Cannot write %s",PROG,path);
    }
  return 0;
}
