/*******************************************************************************************
 *  cpg_prof2class.c -- the prof2class program: ground-truth .class file from a relative profile.
 *
 *      prof2class [-G<device>] <relative_profile>[.prof] <source>[.f[ast][aq][.gz]]
 *
 *  Same inputs, output file and bytes as the reference tool (src/prof2class.c): the profile holds,
 *  for every k-mer of every read, its count in a genome / haplotype k-mer table (FastK -p:table);
 *  count 0 -> E, 1 -> H, 2 -> D, more -> R (src/prof2class.c:236-258); the result goes to
 *  <dir of profile>/<profile root>.class, one fastq-like record per read with K-1 leading 'N's.
 *  Profiles are decoded and mapped on the GPU (cpg_prof2class: k_decode + k_count2class), a batch
 *  of reads at a time.  .db/.dam sources are not supported (DAZZ_DB is out of scope).
 *******************************************************************************************/
#define _GNU_SOURCE
static const char *PROG = "prof2class";
#include "cpg_hostio.h"
#include "classpro_gpu.h"

#define MAX_READ_LEN 60000       /* src/prof2class.c:176 */
static const char *EXT[10] = { ".db",".dam",".fastq",".fasta",".fq",".fa",".fastq.gz",".fasta.gz",".fq.gz",".fa.gz" };

int main(int argc, char **argv)
{ int device = 0, npos = 0; char *pos[2];
  for (int i = 1; i < argc; i++)
    { if (argv[i][0] == '-')
        { if (argv[i][1] == 'G' && argv[i][2]) device = atoi(argv[i]+2);
          else die("%s: -%c is an illegal option",PROG,argv[i][1]);
        }
      else { if (npos < 2) pos[npos] = argv[i]; npos++; }
    }
  if (npos != 2) { fprintf(stderr,"Usage: %s <relative_profile>[.prof] <source>[.f[ast][aq][.gz]]\n",PROG); return 1; }

  char dir[4096], base[1024], root[1024], path[8192], outp[8192];
  split_path(pos[0],dir,sizeof(dir),base,sizeof(base));
  { size_t bl = strlen(base);
    snprintf(root,sizeof(root),"%s",base);
    if (bl > 5 && strcasecmp(base+bl-5,".prof") == 0) root[bl-5] = 0;
  }
  snprintf(outp,sizeof(outp),"%s/%s.class",dir,root);

  char sdir[4096], sbase[1024], sroot[1024];
  split_path(pos[1],sdir,sizeof(sdir),sbase,sizeof(sbase));
  int idx;
  for (idx = 0; idx < 10; idx++)
    { size_t bl = strlen(sbase), el = strlen(EXT[idx]);
      snprintf(sroot,sizeof(sroot),"%s",sbase);
      if (bl > el && strcasecmp(sbase+bl-el,EXT[idx]) == 0) sroot[bl-el] = 0;
      snprintf(path,sizeof(path),"%s/%s%s",sdir,sroot,EXT[idx]);
      int f = open(path,O_RDONLY);
      if (f >= 0) { close(f); break; }
    }
  if (idx == 10) die("Cannot open %s as a .db|.dam or .f{ast}[aq][.gz] file",pos[1]);
  if (idx <= 1) die("%s: .db/.dam inputs are not supported by this build (DAZZ_DB is out of scope)",PROG);

  profidx_t P;
  if (profidx_open(&P,pos[0])) die("%s: Cannot open %s as a .prof file",PROG,pos[0]);
  const int K = P.kmer;

  cpg_model *model = xmalloc(sizeof(cpg_model));
  if (cpg_model_from_cov(model,K,10,20,20000) != CPG_OK) die("%s: cannot set up the device context",PROG);
  if (cpg_device_count() <= 0) die("%s: no CUDA device found: this program has no CPU fallback",PROG);
  cpg_ctx *ctx = NULL;
  if (cpg_create(&ctx,device,model,0,0) != CPG_OK) die("%s: %s",PROG,cpg_last_error(NULL));

  fastx_t X; memset(&X,0,sizeof(X));
  X.f = gzopen(path,"r");
  if (X.f == NULL) die("%s: Cannot open %s",PROG,path);
  gzbuffer(X.f,1<<20);
  X.buf = xmalloc(FX_BUF);
  FILE *out = fopen(outp,"wb");
  if (out == NULL) die("Cannot open %s",outp);
  setvbuf(out,NULL,_IOFBF,1<<22);

  /* a batch: records kept as text, compressed profiles gathered in one buffer */
  const int64_t batch_bases = 256000000;
  int cap = 0, n = 0;
  char **hdr = NULL, **seq = NULL; int32_t *rlen = NULL; int64_t *poff = NULL;
  uint8_t *prof = NULL, *cls = NULL; size_t prof_cap = 0, cls_cap = 0;
  int64_t id = 0; int eof = 0;
  while (!eof && id < P.nreads)
    { int64_t bases = 0, pb = 0;
      n = 0;
      while (bases < batch_bases && id < P.nreads)
        { int rl = fx_read(&X);
          if (rl < 0) { if (rl == -1) { eof = 1; break; } die("%s: truncated quality string in %s",PROG,path); }
          if (rl > MAX_READ_LEN) die("rlen (%d) > rlen_max (%d)",rl,MAX_READ_LEN);
          if (n+1 > cap)
            { int nc = cap+cap/2+1024;
              hdr = xrealloc(hdr,sizeof(char *)*(size_t)nc); seq = xrealloc(seq,sizeof(char *)*(size_t)nc);
              for (int i = cap; i < nc; i++) { hdr[i] = NULL; seq[i] = NULL; }
              rlen = xrealloc(rlen,sizeof(int32_t)*(size_t)nc); poff = xrealloc(poff,sizeof(int64_t)*(size_t)(nc+1));
              cap = nc;
            }
          const char *cm = X.have_comment ? X.comment.s : "(null)";        /* src/prof2class.c:228 */
          size_t hl = strlen(X.name.s)+strlen(cm)+3;
          hdr[n] = xrealloc(hdr[n],hl);
          snprintf(hdr[n],hl,"@%s %s",X.name.s,cm);
          seq[n] = xrealloc(seq[n],(size_t)rl+1);
          memcpy(seq[n],X.seq.s,(size_t)rl+1);
          rlen[n] = rl;
          int part; int64_t off, len;
          prof_range(&P,id,&part,&off,&len);
          if ((size_t)(pb+len)+16 > prof_cap) { prof_cap = (size_t)(pb+len)*3/2+(1<<20); prof = xrealloc(prof,prof_cap); }
          poff[n] = pb;
          int64_t got = 0;
          while (got < len)
            { ssize_t r = pread(P.fd[part],prof+pb+got,(size_t)(len-got),off+got);
              if (r <= 0) die("%s: cannot read profile of read %lld",PROG,(long long)(id+1));
              got += r;
            }
          pb += len; bases += rl; n++; id++;
        }
      if (n == 0) break;
      poff[n] = pb;
      if ((size_t)bases+16 > cls_cap) { cls_cap = (size_t)bases+16; cls = xrealloc(cls,cls_cap); }
      int rc = cpg_prof2class(ctx,n,prof,poff,rlen,cls,NULL);
      if (rc != CPG_OK) die("%s: %s",PROG,cpg_last_error(ctx));
      int64_t co = 0;
      for (int i = 0; i < n; i++)
        { fputs(hdr[i],out); fputc('\n',out);
          fwrite(seq[i],1,(size_t)rlen[i],out);
          fputs("\n+\n",out);
          fwrite(cls+co,1,(size_t)rlen[i],out);
          fputc('\n',out);
          co += rlen[i];
        }
    }
  if (fclose(out) != 0) die("Cannot write %s",outp);
  gzclose(X.f);
  cpg_destroy(ctx);
  return 0;
}
