/*******************************************************************************************
 *  cpg_model.c -- host one-shot model (stays on the host by design: it is a few milliseconds of
 *  scalar work on a 256 KB histogram and must use the host libm exactly as the reference does).
 *
 *  Replaces process_global_hist (src/hist.c:28-143) with Load_Histogram/Modify_Histogram
 *  (src/libfastk.c:22-147), the derived globals of src/ClassPro.c:543-548 (plus_sigma:
 *  src/util.c:9-11), load_emodel/calc_init_thres with the default error model
 *  (src/wall.c:120-244) and precompute_logfact (src/prob.c:14-19).
 *  The -M <model_path> error model needs GSL, which is absent from the reference tree; it is out
 *  of scope.
 *******************************************************************************************/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "classpro_gpu.h"

enum { S_E = 0, S_R = 1, S_H = 2, S_D = 3 };

static const double PE_THRES[2][2] = { {0.001, 0.05}, {1e-5, 1e-5} };   /* src/const.c:64-65 */

static int lrow(int t, int l) { return (t == 0 ? 0 : (t == 1 ? 20 : 30))+l-1; }

static int finish_model(cpg_model *m)
{ /* src/prob.c:14-19 */
  m->logfact[0] = 0.;
  for (int n = 1; n <= 32767; n++)
    m->logfact[n] = m->logfact[n-1]+log((double)n);

  /* src/ClassPro.c:544-548 */
  const int D = m->cov[S_D];
  m->cov[S_E] = 1;
  m->cov[S_R] = (uint16_t)(D+(uint16_t)(sqrt((double)D)*5));
  m->dr_ratio = 1.+(double)2*(1./sqrt((double)D));

  /* src/wall.c:174-180 */
  if (m->cov[S_R] > 255)
    { fprintf(stderr,"Too high REPEAT coverage (%d) > 255\n",m->cov[S_R]);
      return CPG_EMODEL;
    }
  m->cmax = (uint8_t)m->cov[S_R];
  for (int t = 0; t < 3; t++)
    { m->lmax[t] = (uint8_t)(20/(t+1));
      m->pe[t][0] = 0.;
      for (int l = 1; l <= m->lmax[t]; l++)
        m->pe[t][l] = 0.002*l*l+0.002;
    }
  m->hc_erate = m->pe[0][1];

  /* src/wall.c:190-224: for every context row and outside count, the inside counts at which the
     running upper tail 1 - sum_{c<=cin} Binom(c; cout, pe) first drops under each threshold */
  memset(m->cthres,0,sizeof(m->cthres));
  for (int t = 0; t < 3; t++)
    for (int l = 1; l <= m->lmax[t]; l++)
      { const double pe = m->pe[t][l], lpe = log(pe), l1mpe = log(1-pe);
        for (int cout = 1; cout < m->cmax; cout++)
          { uint8_t *cell = m->cthres+((size_t)lrow(t,l)*256+cout)*4;     /* [thresT][etype] */
            int found[2][2] = {{0,0},{0,0}};
            for (int s = 0; s < 2; s++) { cell[s*2+0] = (uint8_t)cout; cell[s*2+1] = 0; }
            double psum = 1.;
            for (int cin = 0; cin <= cout; cin++)
              { if (found[0][0] && found[1][0] && found[0][1] && found[1][1]) break;
                psum -= exp(m->logfact[cout]-m->logfact[cin]-m->logfact[cout-cin]+cin*lpe+(cout-cin)*l1mpe);
                for (int s = 0; s < 2; s++)
                  for (int e = 0; e < 2; e++)
                    if (!found[s][e] && psum < PE_THRES[s][e])
                      { cell[s*2+e] = (uint8_t)(e == 0 ? cin : cout-cin); found[s][e] = 1; }
              }
          }
      }
  return CPG_OK;
}

int cpg_model_from_cov(cpg_model *m, int kmer, int h, int d, int read_len)
{ if (m == NULL || d <= 0 || read_len <= 0) return CPG_EINVAL;
  memset(m,0,sizeof(*m));
  m->kmer = kmer; m->read_len = read_len;
  m->cov[S_D] = (uint16_t)d;
  m->cov[S_H] = (uint16_t)(h > 0 ? h : (d >> 1));
  return finish_model(m);
}

int cpg_model_from_hist(cpg_model *m, int kmer, int low, int high, int64_t ilowcnt, int64_t ihighcnt,
                        const int64_t *raw, int cov_opt, int read_len, int verbose)
{ if (m == NULL || raw == NULL || read_len <= 0 || high <= low) return CPG_EINVAL;
  memset(m,0,sizeof(*m));
  m->kmer = kmer; m->read_len = read_len;
  int H, D;
  if (verbose) fprintf(stderr,"Global histogram inspection:\n");
  if (cov_opt > 0)
    { D = cov_opt; H = cov_opt >> 1;                      /* src/hist.c:44-49 */
      if (verbose) fprintf(stderr,"    Specified (H,D) cov   = (%d,%d)\n",H,D);
    }
  else
    { /* distinct-k-mer bins -> k-mer instance bins: interior bins times their count, the two
         boundary bins replaced by the instance totals stored in the header */
      const int nb = high-low+1;
      int64_t *inst = malloc(sizeof(int64_t)*(size_t)(nb+3));
      if (inst == NULL) return CPG_ENOMEM;
      int64_t *h = inst-low;
      for (int c = low; c <= high; c++) h[c] = raw[c-low];
      for (int c = low+1; c < high; c++) h[c] *= c;
      h[low] = ilowcnt; h[high] = ihighcnt;
      h[high+1] = raw[0]; h[high+2] = raw[nb-1];     /* the toggled-out distinct counts */

      /* tallest interior local maximum below 1000 (src/hist.c:58-64) */
      int top = 0; int64_t toppk = 0;
      const int lo = low > 2 ? low : 2, hi = high < 1000 ? high : 1000;
      for (int c = lo; c < hi; c++)
        if (h[c-1] < h[c] && h[c] > h[c+1] && toppk < h[c]) { top = c; toppk = h[c]; }
      if (top < 10)
        { fprintf(stderr,"[ERROR] Could not find any peak count >= 10 in the histogram. Revise data and use the `-c` option.");
          free(inst);
          return CPG_EMODEL;
        }
      if (verbose)
        fprintf(stderr,"    Tallest peak count    = %d (# of k-mers = %lld)\n",top,(long long)toppk);

      /* best bin within one sigma of top/2 and of 2*top, and whether it is a local maximum */
      int cnt2[2], ispk[2]; int64_t pk2[2];
      for (int side = 0; side < 2; side++)
        { double mean = side == 0 ? (double)top/2 : (double)top*2, sd = sqrt(mean);
          cnt2[side] = 0; ispk[side] = 0; pk2[side] = 0;
          for (int c = (int)round(mean-sd); c <= (int)round(mean+sd); c++)
            if (pk2[side] < h[c])
              { cnt2[side] = c; pk2[side] = h[c];
                ispk[side] = (h[c-1] < h[c] && h[c] > h[c+1]) ? 1 : 0;
              }
        }
      if (pk2[0] > pk2[1]) { D = top; H = ispk[0] ? cnt2[0] : (top >> 1); }
      else                 { H = top; D = ispk[1] ? cnt2[1] : (top << 1); }
      if (verbose) fprintf(stderr,"    Estimated (H,D) cov   = (%d,%d)\n",H,D);
      free(inst);
    }
  m->cov[S_H] = (uint16_t)H;
  m->cov[S_D] = (uint16_t)D;
  int rc = finish_model(m);
  if (rc == CPG_OK && verbose) fprintf(stderr,"    Estimated R-threshold = %d\n",m->cov[S_R]);
  return rc;
}

int cpg_model_load(cpg_model *m, const char *fk_root, int cov_opt, int read_len, int verbose)
{ char path[4200];
  snprintf(path,sizeof(path),"%s.hist",fk_root);
  FILE *f = fopen(path,"rb");
  if (f == NULL) { fprintf(stderr,"Cannot open %s\n",path); return CPG_EIO; }
  int32_t kmer, low, high; int64_t il, ih;
  if (fread(&kmer,4,1,f) != 1 || fread(&low,4,1,f) != 1 || fread(&high,4,1,f) != 1
      || fread(&il,8,1,f) != 1 || fread(&ih,8,1,f) != 1 || high < low)
    { fclose(f); return CPG_EIO; }
  const size_t nb = (size_t)(high-low+1);
  int64_t *h = malloc(sizeof(int64_t)*nb);
  if (h == NULL) { fclose(f); return CPG_ENOMEM; }
  if (fread(h,8,nb,f) != nb) { fclose(f); free(h); return CPG_EIO; }
  fclose(f);
  int rc = cpg_model_from_hist(m,kmer,low,high,il,ih,h,cov_opt,read_len,verbose);
  free(h);
  return rc;
}
