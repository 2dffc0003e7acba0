/* cpsim.h -- harness data generator interface (test/bench infrastructure, not product). */
#ifndef CPSIM_H
#define CPSIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
  { int64_t seed;
    int64_t genome_len;
    double  het;             /* per-base variant rate between the two haplotypes */
    int     snp_only;        /* 1: SNPs only (haplotype coordinates stay aligned) */
    double  repeat_frac;     /* fraction of blocks that are tandem/interspersed/low-complexity */
    int64_t seg_dups;
    double  cov;             /* read bases / genome_len */
    int64_t len_mean, len_sd, len_min, len_max;
    double  err_sub, err_indel_base, err_indel_hp;
    int64_t kmer;
    int64_t nparts;          /* number of FastK profile parts written */
    int     exact;           /* 1: exact canonical k-mer counts; 0: ground-truth coverage; 2: reads only (no counts) */
    int     short_reads;     /* 1: sprinkle reads shorter than k (edge-case tests) */
  } cpsim_params;

typedef struct
  { int      kmer;
    int64_t  nreads, total_bases, total_kmers, prof_bytes;
    uint8_t *seq;            /* base codes 0..3 (A,C,G,T), reads concatenated */
    int64_t *seq_off;        /* [nreads+1] */
    int32_t *rlen;           /* [nreads] */
    uint16_t*counts;         /* uncompressed profiles, concatenated */
    int64_t *cnt_off;        /* [nreads+1] */
    uint8_t *prof;           /* FastK-compressed profiles, concatenated */
    int64_t *prof_off;       /* [nreads+1] */
    int64_t *hist;           /* [32770]: hist[c] distinct k-mers with count c; [32768],[32769] hidden bins */
    char    *hdr;            /* header text (without '>'), concatenated */
    int64_t *hdr_off;        /* [nreads+1] */
  } cpsim_data;

void cpsim_default_params(cpsim_params *P);
int  cpsim_generate(const cpsim_params *P, cpsim_data *D);
int  cpsim_write_files(const cpsim_params *P, const cpsim_data *D, const char *dir, const char *root);
void cpsim_free(cpsim_data *D);

#ifdef __cplusplus
}
#endif
#endif
