/*******************************************************************************************
 *  refstage.c -- stage-level dump driver around the UNMODIFIED reference sources.
 *
 *  TEST INFRASTRUCTURE.  This translation unit #includes the reference's own .c files from
 *  /root/reference/src (same order as src/ClassPro.c:16-25; nothing is copied into this repo),
 *  repeats the host one-shot setup of src/ClassPro.c:536-554 and runs the per-read stage
 *  functions, dumping for every read the final interval table
 *      <read#> N M
 *      b e cb ce is_rel ccb cce asgn pe pe_o.b pe_o.e      (N lines; ccb/cce only when is_rel)
 *  so that the restatement (oracle/classpro_oracle.c) and the CUDA path can be compared stage
 *  by stage, not only on the final .class bytes.
 *
 *  usage: refstage [-c<int>] [-r<int>] <reads.fasta> > dump.txt
 *******************************************************************************************/
#include <stdio.h>
#include <stdlib.h>
#include <stdbool.h>
#include <string.h>
#include <unistd.h>
#include <math.h>
#include <limits.h>
#include <float.h>
#include <fcntl.h>
#include <pthread.h>
#include <sys/stat.h>

#include "ClassPro.h"
#include "benchmark.h"

#include "const.c"
#include "io.c"
#include "prob.c"
#include "util.c"
#include "hist.c"
#include "context.c"
#include "wall.c"
#include "class_rel.c"
#include "class_unrel.c"
#include "seed.c"

bool  VERBOSE;
int   READ_LEN;
bool  IS_DB;
bool  IS_DAM;
bool  FIND_SEED;
cnt_t GLOBAL_COV[N_STATE];

int main(int argc, char *argv[])
{ int cov = 0;
  const char *fasta = NULL;
  READ_LEN = 20000;
  Prog_Name = Strdup("refstage","");
  for (int i = 1; i < argc; i++)
    { if (argv[i][0] == '-' && argv[i][1] == 'c') cov = atoi(argv[i]+2);
      else if (argv[i][0] == '-' && argv[i][1] == 'r') READ_LEN = atoi(argv[i]+2);
      else fasta = argv[i];
    }
  if (fasta == NULL) { fprintf(stderr,"usage: refstage [-c<int>] [-r<int>] <reads.fasta>\n"); return 1; }
  char *path = PathTo(fasta);
  char *root = Root((char *)fasta,".fasta");
  char *fk_root = Strdup(Catenate(path,"/",root,""),"fk");

  Profile_Index *P = Open_Profiles(fk_root);
  if (P == NULL) { fprintf(stderr,"refstage: cannot open %s.prof\n",fk_root); return 1; }
  precompute_logfact();
  process_global_hist(fk_root,cov);
  GLOBAL_COV[HAPLO] = lambda_prior[0];
  GLOBAL_COV[DIPLO] = lambda_prior[1];
  GLOBAL_COV[ERROR] = 1;
  GLOBAL_COV[REPEAT] = plus_sigma(GLOBAL_COV[DIPLO],N_SIGMA_RCOV);
  DR_RATIO = 1.+(double)N_SIGMA_R*(1./sqrt(GLOBAL_COV[DIPLO]));
  Error_Model *emodel = calc_init_thres(NULL);

  const int K = P->kmer, Km1 = K-1;
  const int rlen_max = MAX_READ_LEN;
  gzFile fp = gzopen(fasta,"r");
  kseq_t *ks = kseq_init(fp);

  Rel_Arg  *rel_arg = alloc_rel_arg(rlen_max);
  Wall_Arg *warg    = alloc_wall_arg(rlen_max);
  Intvl    *intvl   = Malloc(rlen_max*sizeof(Intvl),"i");
  Intvl    *rintvl  = Malloc(rlen_max*sizeof(Intvl),"r");
  cnt_t    *profile = Malloc(rlen_max*sizeof(cnt_t),"p");
  Seq_Ctx  *_lctx   = Malloc(rlen_max*sizeof(Seq_Ctx),"l");
  Seq_Ctx  *rctx    = Malloc(rlen_max*sizeof(Seq_Ctx),"r");
  Seq_Ctx  *ctx[N_WTYPE];
  _lctx[0][HP] = 1;
  _lctx[0][DS] = _lctx[0][TS] = _lctx[1][TS] = 0;
  ctx[DROP] = _lctx+Km1-1;
  ctx[GAIN] = rctx;

  for (int id = 0; id < P->nreads; id++)
    { if (kseq_read(ks) < 0) break;
      int rlen = ks->seq.l;
      char *seq = ks->seq.s;
      if (rlen <= Km1) { printf("%d 0 0\n",id+1); continue; }
      calc_seq_context(_lctx,rctx,seq,rlen);
      int plen = Fetch_Profile(P,(int64)id,rlen_max,profile);
      if (rlen != plen+Km1) { fprintf(stderr,"refstage: rlen mismatch at read %d\n",id+1); return 1; }
      int N = find_wall(warg,intvl,profile,plen,ctx,emodel,K);
      int M = find_rel_intvl(intvl,N,rintvl,profile,ctx,K);
      classify_rel(rel_arg,rintvl,M,intvl,N,plen);
      classify_unrel(intvl,N);
      printf("%d %d %d\n",id+1,N,M);
      for (int i = 0; i < N; i++)
        { Intvl I = intvl[i];
          printf("%d %d %d %d %d %d %d %d %.17g %.17g %.17g\n",I.b,I.e,I.cb,I.ce,I.is_rel ? 1 : 0,
                 I.is_rel ? I.ccb : 0,I.is_rel ? I.cce : 0,(int)I.asgn,I.pe,I.pe_o.b,
                 (i+1 < N) ? I.pe_o.e : 0.);
        }
    }
  return 0;
}
