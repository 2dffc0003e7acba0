/*******************************************************************************************
 *  cpg_class2acc.c -- the class2acc program: accuracy of an estimated .class file against a ground-truth .class file.
 *
 *      class2acc [-s] [-e<int>] [-f<int(100)>] [-m<int(0)>] [-n<int(100)>] [-r<int(0)>] [-w<int>] [-p<profile root>]
 *                <estimate>.class <truth>.class
 *
 *  Same comparison and the same report as the reference's evaluation tool (src/class2acc.c:
 *  per-read loop :152-295, report :300-316): confusion matrix Truth\Est over E,R,H,D, accuracy and
 *  false-negative error rate over all reads and split into "normal" / "repeat" reads (-r), reads
 *  with more than -f percent of true E-mers skipped, optional per-read lines (-e, -s, -m, -n), and with -p
 *  <FastK profile root> the H / D coverages of a read from the counts of its true H- / D-mers (in the -e lines)
 *  and, with -w, one line per window of that many k-mers (src/class2acc.c:98-104,174-185,228-247,272-279).
 *  Host-only evaluation tool: SURVEY section 8 row f2; no GPU work in it (the profiles of the handful of reads it
 *  looks at are decoded here, by a restatement of the codec of src/libfastk.c:1467-1535).
 *******************************************************************************************/
#define _GNU_SOURCE
static const char *PROG = "class2acc";
#include "cpg_hostio.h"

static const char STOC[4] = { 'E', 'R', 'H', 'D' };

static int ctos(int c) { return c == 'D' ? 3 : c == 'H' ? 2 : c == 'R' ? 1 : 0; }    /* src/class2acc.c:16-31 */

/* FastK profile codec, src/libfastk.c:1467-1535 (format: SURVEY A.1): decoded length, first min(length,cap) counts stored */
static int decode_profile(const uint8_t *p, int64_t len, uint16_t *out, int cap)
{ if (len == 0) return 0;
  const uint8_t *q = p+len;
  uint16_t x = *p++, d;
  if (x & 0x80) d = (uint16_t)(((x & 0x7f) << 8) | *p++); else d = x;
  int n = 1;
  if (cap > 0) out[0] = d;
  while (p < q)
    { x = *p++;
      if ((x & 0xc0) == 0)
        { for (int i = 0; i < x; i++, n++) if (n < cap) out[n] = d; }
      else
        { if (x & 0x80)
            { x = (x & 0x40) ? (uint16_t)(x << 8) : (uint16_t)((x << 8) & 0x7fff);
              x |= *p++;
              d = (uint16_t)((d+x) & 0x7fff);
            }
          else if (x & 0x20) d = (uint16_t)(d+((x & 0x1fu) | 0xffe0u));
          else               d = (uint16_t)(d+(x & 0x1fu));
          if (n < cap) out[n] = d;
          n++;
        }
    }
  return n;
}

static void open_class(fastx_t *x, const char *path)
{ memset(x,0,sizeof(*x));
  x->f = gzopen(path,"r");
  if (x->f == NULL) die("%s: Cannot open %s",PROG,path);
  gzbuffer(x->f,1<<20);
  x->buf = xmalloc(FX_BUF);
}

int main(int argc, char **argv)
{ int show_lq = 0, show_class = 0, min_r = 0, max_r = 100, thres_lq = -1, thres_r = 0, thres_e = 100, window = -1;
  char *pos[2]; int npos = 0;
  const char *prof_root = NULL;
  for (int i = 1; i < argc; i++)
    { char *a = argv[i], *e = NULL;
      if (a[0] != '-') { if (npos < 2) pos[npos] = a; npos++; continue; }
      switch (a[1])
        { case 'e': show_lq = 1; thres_lq = (int)strtol(a+2,&e,10); break;
          case 'f': thres_e = (int)strtol(a+2,&e,10); break;
          case 'm': min_r = (int)strtol(a+2,&e,10); break;
          case 'n': max_r = (int)strtol(a+2,&e,10); break;
          case 'r': thres_r = (int)strtol(a+2,&e,10); break;
          case 'w': window = (int)strtol(a+2,&e,10); break;
          case 'p': prof_root = a+2; continue;
          default:
            for (char *p = a+1; *p; p++)
              { if (*p == 's') show_class = 1;
                else die("%s: -%c is an illegal option",PROG,*p);
              }
            continue;
        }
      if (e == NULL || *e || a[2] == 0) die("%s: -%c '%s' argument is not an integer",PROG,a[1],a+2);
    }
  if (npos != 2)
    { fprintf(stderr,"Usage: %s [-s] [-e<int>] [-f<int(100)>] [-m<int(0)>] [-n<int(100)>] [-r<int(0)>] [-w<int>] "
                     "[-p<profile root>] <estimate>.class <truth>.class\n",PROG);
      return 1;
    }
  fastx_t E, T;
  open_class(&E,pos[0]); open_class(&T,pos[1]);
  profidx_t P; int have_p = 0, km1 = -1;
  uint16_t *profile = NULL; int pmax = 0; uint8_t *raw = NULL; int64_t raw_cap = 0;
  if (prof_root != NULL)
    { if (profidx_open(&P,prof_root) != 0) die("%s: Cannot open %s as a .prof file",PROG,prof_root);
      have_p = 1; km1 = P.kmer-1;
    }
  double cov[2] = { -1., -1. };            /* per-read (or per-window) haplo / diplo coverages; kept across reads as in the reference */

  long long ntot = 0, ncor = 0, nfne = 0, ntot_n = 0, ncor_n = 0, nfne_n = 0, ntot_r = 0, ncor_r = 0, nfne_r = 0;
  long long cfm[4][4];
  memset(cfm,0,sizeof(cfm));
  int id = 1;
  for (;;)
    { int le = fx_read(&E);
      if (le < 0) break;
      int lt = fx_read(&T);
      if (lt < 0) die("# seqs in %s > # seqs in %s",pos[0],pos[1]);
      if (strcmp(E.name.s,T.name.s) != 0)
        die("Read %d inconsistent names: %s (estimate) vs %s (truth)",id,E.name.s,T.name.s);
      if (!(E.seq.l == E.qual.l && T.seq.l == T.qual.l && E.seq.l == T.seq.l))
        die("Read %d inconsistent lengths",id);
      const char *qe = E.qual.s, *qt = T.qual.s;
      const int L = (int)T.qual.l;
      if (have_p)
        { int part; int64_t off, len;
          prof_range(&P,(int64_t)id-1,&part,&off,&len);
          if (len > raw_cap) { raw_cap = len+len/4+4096; raw = xrealloc(raw,(size_t)raw_cap); }
          if (len > 0 && pread(P.fd[part],raw,(size_t)len,(off_t)off) != (ssize_t)len) die("%s: cannot read the profile of read %d",PROG,id);
          if (L+1 > pmax) { pmax = L+L/4+1024; profile = xrealloc(profile,sizeof(uint16_t)*(size_t)pmax); }
          const int plen = decode_profile(raw,len,profile,pmax);
          if (plen+km1 != (int)E.qual.l) die("Read %d inconsist lengths: %ld (estimate) vs %d (profile)",id,(long)E.qual.l,plen+km1);
        }
      int i = 0;
      while (i < L && qe[i] == 'N')
        { if (qt[i] != 'N') die("Read %d inconsistent # of prefix Ns (= K-1)",id);
          i++;
        }
      const int rtot = L-i;
      int rcor = 0, rfne = 0, rcomp[4] = {0,0,0,0};
      int wcor = 0, wcomp[4] = {0,0,0,0};
      long long scnts[2] = {0,0};
      for (int c = 1; i < L; i++, c++)
        { if (qe[i] == qt[i]) { rcor++; wcor++; }
          if (qt[i] == 'E' && qe[i] != 'E') rfne++;
          cfm[ctos(qt[i])][ctos(qe[i])]++;
          switch (qt[i])
            { case 'E': rcomp[0]++; wcomp[0]++; break;
              case 'H': rcomp[1]++; wcomp[1]++; break;
              case 'D': rcomp[2]++; wcomp[2]++; break;
              case 'R': rcomp[3]++; wcomp[3]++; break;
              default:  fprintf(stderr,"Invalid class: %c\n",qt[i]); break;
            }
          if (have_p)
            { /* src/class2acc.c:228-247: the sums restart at every window, so a read's own coverages (below) are
                 those of its last, partial window over the H / D counts of the whole read -- kept as it is */
              if (qt[i] == 'H') scnts[0] += profile[i-km1];
              else if (qt[i] == 'D') scnts[1] += profile[i-km1];
              if (window > 0 && c % window == 0)
                { cov[0] = (wcomp[1] > 0) ? (double)scnts[0]/wcomp[1] : -1;
                  cov[1] = (wcomp[2] > 0) ? (double)scnts[1]/wcomp[2] : -1;
                  if (cov[0] == -1 || cov[1] == -1 || cov[0] > cov[1]) cov[0] = cov[1] = -1;
                  else cov[1] -= cov[0];
                  printf("%%error = %4.1lf [H1-cov=%.lf,H2-cov=%.lf]\n",(double)(window-wcor)/window*100,cov[0],cov[1]);
                  scnts[0] = scnts[1] = 0;
                  wcomp[0] = wcomp[1] = wcomp[2] = wcomp[3] = 0;
                  wcor = 0;
                }
            }
        }
      if ((double)rcomp[0]/rtot*100 > thres_e) { id++; continue; }
      ntot += rtot; ncor += rcor; nfne += rfne;
      if ((double)rcomp[3]/rtot*100 > thres_r) { ntot_r += rtot; ncor_r += rcor; nfne_r += rfne; }
      else                                     { ntot_n += rtot; ncor_n += rcor; nfne_n += rfne; }
      if (have_p)
        { cov[0] = (rcomp[1] > 0) ? (double)scnts[0]/rcomp[1] : -1;
          cov[1] = (rcomp[2] > 0) ? (double)scnts[1]/rcomp[2] : -1;
          if (cov[0] == -1 || cov[1] == -1 || cov[0] > cov[1]) cov[0] = cov[1] = -1;
          else cov[1] -= cov[0];
        }
      if (show_lq && (double)(rtot-rcor)/rtot*100 >= thres_lq
          && min_r <= (double)rcomp[3]/rtot*100 && (double)rcomp[3]/rtot*100 <= max_r)
        { printf("Read %6d (%ld bp, %d classes): %%error = %4.1lf [%%E=%4.1lf,%%H=%4.1lf,%%D=%4.1lf,%%R=%4.1lf] [H1-cov=%.lf,H2-cov=%.lf]\n",
                 id,(long)T.seq.l,rtot,(double)(rtot-rcor)/rtot*100,
                 (double)rcomp[0]/rtot*100,(double)rcomp[1]/rtot*100,(double)rcomp[2]/rtot*100,(double)rcomp[3]/rtot*100,cov[0],cov[1]);
          if (show_class)
            { printf("truth: %s\n  est: ",qt);
              for (int k = 0; k < L; k++) putchar(qt[k] != qe[k] ? qe[k] : '-');
              putchar('\n');
            }
        }
      id++;
    }
  if (fx_read(&T) >= 0) die("# seqs in %s < # seqs in %s",pos[0],pos[1]);

  printf("\nConfusion Matrix (Truth\\Est):\n  ");
  for (int i = 0; i < 4; i++) printf("%15c",STOC[i]);
  printf("\n");
  for (int i = 0; i < 4; i++)
    { printf("%c:",STOC[i]);
      for (int j = 0; j < 4; j++) printf("%15lld",cfm[i][j]);
      printf("\n");
    }
  printf("\nAccuracy = %4.2lf %% (= %lld / %lld), FN Error = %4.2lf %%\n",(double)ncor/ntot*100,ncor,ntot,(double)nfne/ntot*100);
  printf("[Normal] Accuracy = %4.2lf %% (= %lld / %lld), FN Error = %4.2lf %%\n",
         (double)ncor_n/ntot_n*100,ncor_n,ntot_n,(double)nfne_n/ntot_n*100);
  printf("[Repeat] Accuracy = %4.2lf %% (= %lld / %lld), FN Error = %4.2lf %%\n",
         (double)ncor_r/ntot_r*100,ncor_r,ntot_r,(double)nfne_r/ntot_r*100);
  gzclose(E.f); gzclose(T.f);
  return 0;
}
