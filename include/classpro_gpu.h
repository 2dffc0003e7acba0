/*******************************************************************************************
 *  classpro_gpu.h -- C ABI of libclasspro_b200.so: ClassPro's per-read classification path
 *  (profile decode -> sequence context -> wall detection -> reliable-interval DP ->
 *  unreliable-interval assignment -> per-k-mer E/H/D/R string) on a B200.
 *
 *  Plain C: pointers and sizes only, no CUDA or torch types.  The reference has no library
 *  boundary for this path (one translation unit, src/ClassPro.c:16-25); every entry point below
 *  names the reference code it stands in for.  file:line are relative to the reference tree.
 *
 *  Threading: a cpg_ctx is bound to one GPU and must be driven by one host thread at a time; use
 *  one context per GPU (reads are independent, src/io.c:353-354, so GPUs never talk to each other).
 *  Errors: every call returns 0 on success or a CPG_E* code; cpg_last_error() gives the text.
 *  There is NO CPU fallback: without a usable CUDA device cpg_create fails.
 *******************************************************************************************/
#ifndef CLASSPRO_GPU_H
#define CLASSPRO_GPU_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CPG_OK          0
#define CPG_EINVAL      1      /* bad argument                                           */
#define CPG_ECUDA       2      /* CUDA runtime error (text in cpg_last_error)            */
#define CPG_ENOMEM      3      /* host or device allocation failed                       */
#define CPG_EMODEL      4      /* histogram/model error (hist.c:65-68, wall.c:174-177)   */
#define CPG_EREAD       5      /* at least one read failed; see cpg_result.status        */
#define CPG_EIO         6

/* ---- host one-shot model ---------------------------------------------------------------
 * Replaces process_global_hist (src/hist.c:28-143), the derived globals of
 * src/ClassPro.c:543-548, calc_init_thres/load_emodel (src/wall.c:120-244, default error model)
 * and precompute_logfact (src/prob.c:14-19).  Computed on the host with the host libm exactly as
 * the reference does (north-star item 3), uploaded once by cpg_create. */
typedef struct
  { int32_t  kmer;              /* K                                      */
    int32_t  read_len;          /* -r, READ_LEN (src/ClassPro.c:516)      */
    uint16_t cov[4];            /* GLOBAL_COV[E,R,H,D]                    */
    double   dr_ratio;          /* DR_RATIO                               */
    int32_t  cmax;              /* CMAX (src/wall.c:178)                  */
    double   hc_erate;          /* HC_ERATE (src/wall.c:180)              */
    int32_t  lmax[3];           /* per context type (src/wall.c:123)      */
    double   pe[3][21];         /* error rate by context type and length  */
    uint8_t  cthres[36*256*4];  /* [row(t,l)][cout][INIT|FINAL][SELF|OTHERS], rows 20+10+6 */
    double   logfact[32768];
  } cpg_model;

/* hist = the (high-low+1) int64 bins as stored in <root>.hist after the 28-byte header
 * (src/libfastk.c:72-83).  cov_opt = the -c value (0 = estimate from the histogram). */
int cpg_model_from_hist(cpg_model *m, int kmer, int low, int high, int64_t ilowcnt, int64_t ihighcnt,
                        const int64_t *hist, int cov_opt, int read_len, int verbose);
/* Same, reading <fk_root>.hist (src/libfastk.c:51-96). */
int cpg_model_load(cpg_model *m, const char *fk_root, int cov_opt, int read_len, int verbose);
/* Model for given (H,D) coverages (what -c<D> gives with h = 0 -> D>>1, src/hist.c:44-49). */
int cpg_model_from_cov(cpg_model *m, int kmer, int h, int d, int read_len);

/* ---- context ------------------------------------------------------------------------------ */
typedef struct cpg_ctx cpg_ctx;

int  cpg_device_count(void);
/* max_batch_bases / max_batch_reads size the device and pinned staging buffers (0 = defaults). */
int  cpg_create(cpg_ctx **ctx, int device, const cpg_model *model,
                int64_t max_batch_bases, int32_t max_batch_reads);
void cpg_destroy(cpg_ctx *ctx);
const char *cpg_last_error(const cpg_ctx *ctx);   /* ctx may be NULL: last creation error */

/* ---- batches ------------------------------------------------------------------------------
 * A batch is n_reads consecutive reads, every one with rlen >= K (shorter reads never reach the
 * stage functions in the reference either, src/ClassPro.c:209-226; the caller prints them).
 *   seq      : read sequences. seq_bits = 8: the raw characters (compared as bytes, exactly as
 *              src/context.c does); seq_bits = 2: bases packed 4 per byte, base i of a read in
 *              bits 2*(i&3) of byte i>>2, A,C,G,T = 0..3 (cpg_pack_seq), every read starting at
 *              a byte boundary.
 *   seq_off  : [n_reads+1] byte offsets of each read inside seq
 *   rlen     : [n_reads]   read lengths in bases
 *   prof     : FastK-compressed profiles (src/libfastk.c:1467-1535), read after read
 *   prof_off : [n_reads+1] byte offsets inside prof  (= the .pidx index of the part, rebased)
 * The arrays may live in pageable or pinned host memory. */
typedef struct
  { int32_t        n_reads;
    int32_t        seq_bits;
    const uint8_t *seq;
    const int64_t *seq_off;
    const int32_t *rlen;
    const uint8_t *prof;
    const int64_t *prof_off;
  } cpg_batch;

/* Result of a batch, owned by the caller.
 *   cls      : class strings, read after read: rlen characters per read, 'N' x (K-1) followed by
 *              one of E/H/D/R per k-mer -- the 4th line of the read's .class record without the
 *              newline (src/ClassPro.c:114-117,265-271,289)
 *   cls_off  : [n_reads+1] byte offsets inside cls; the caller fills it (normally the prefix sums
 *              of rlen) and must provide cls_off[n_reads] bytes
 *   status   : [n_reads] 0 = classified; otherwise CPG_ST_* bits (cpg_status_string) for the
 *              conditions on which the reference prints a message and exits */
typedef struct
  { uint8_t       *cls;
    const int64_t *cls_off;
    int32_t       *status;
  } cpg_result;

/* Pack an ASCII read into 2-bit codes.  Returns 0, or 1 if a character outside "ACGT" is met
 * (then ship the batch with seq_bits = 8). */
int cpg_pack_seq(const char *seq, int32_t rlen, uint8_t *out);

/* Host buffers in, host buffers out: H2D copies, kernels, D2H copy, synchronous.
 * Stands in for steps 2-6 of the per-read loop, src/ClassPro.c:229-271, for a whole batch. */
int cpg_classify(cpg_ctx *ctx, const cpg_batch *batch, cpg_result *result);

/* Asynchronous pair on one of two slots (double buffering: slot 0/1): cpg_submit enqueues the
 * host->device copies STRAIGHT FROM THE CALLER'S ARRAYS (no staging copy) and the kernels;
 * cpg_collect waits for that slot and copies the classes out.  The arrays of the batch must
 * therefore stay valid and unchanged until cpg_collect of that slot has returned (pinned arrays,
 * cpg_host_alloc, make the copies true DMA transfers).  A slot must be collected before it is
 * submitted again.  cpg_result.cls_off, when given, must be the prefix sums of rlen; a result that
 * cpg_collect rejects (CPG_EINVAL) leaves the batch in flight: call again with a valid one.
 * One host thread per context: the calls on one cpg_ctx are not re-entrant. */
int cpg_submit(cpg_ctx *ctx, int slot, const cpg_batch *batch);
int cpg_collect(cpg_ctx *ctx, int slot, cpg_result *result);

/* ---- compact results ---------------------------------------------------------------------------
 * A class string is piecewise constant: the read's interval table (about one interval per 70 k-mers on
 * HiFi profiles) says everything it says.  In CPG_RESULT_INTERVALS mode a batch comes back as that table
 * -- 4 bytes per interval instead of 1 byte per base, ~30 times fewer bytes over PCIe and no class-string
 * kernel -- and the caller expands it where the characters are needed (cpg_expand_intervals: the bytes are
 * those of CPG_RESULT_CLASSES mode, i.e. of src/ClassPro.c:114-117,265-271).  The mode is a property of the
 * context: set it before the first cpg_submit.
 *   ivl      : [ivl_cap] OUT packed intervals, (end position << 3) | class code (0 E, 1 R, 2 H, 3 D); the
 *              intervals of a read are consecutive, in position order, the first one starts at 0
 *   ivl_cap  : entries the caller provides; cpg_intervals_bound(ctx,slot) after cpg_submit is enough
 *   ivl_at   : [n_reads] OUT first entry of read r (reads are NOT in order inside ivl)
 *   ivl_n    : [n_reads] OUT number of intervals of read r
 *   status   : [n_reads] OUT as in cpg_result
 *   ivl_used : OUT entries of ivl that were filled */
#define CPG_RESULT_CLASSES   0
#define CPG_RESULT_INTERVALS 1
typedef struct
  { uint32_t *ivl;
    int64_t   ivl_cap;
    int64_t  *ivl_at;
    int32_t  *ivl_n;
    int32_t  *status;
    int64_t   ivl_used;
  } cpg_result_ivl;
int     cpg_set_result_mode(cpg_ctx *ctx, int mode);
int64_t cpg_intervals_bound(cpg_ctx *ctx, int slot);
int     cpg_collect_intervals(cpg_ctx *ctx, int slot, cpg_result_ivl *result);
/* rlen characters at out: 'N' x (K-1), then the class of every k-mer (host code, no device involved) */
void    cpg_expand_intervals(int32_t K, int32_t rlen, const uint32_t *ivl, int32_t n, char *out);

/* ---- stage access (tests, benchmarks) ------------------------------------------------------
 * cpg_decode_profiles: the Fetch_Profile replacement alone (src/libfastk.c:1414-1562): counts of
 * read r are written to counts[cnt_off[r] .. ) with capacity cnt_off[r+1]-cnt_off[r]; plen[r]
 * receives the decoded length (which may exceed the capacity, as Fetch_Profile's return does). */
int cpg_decode_profiles(cpg_ctx *ctx, int32_t n_reads, const uint8_t *prof, const int64_t *prof_off,
                        const int64_t *cnt_off, uint16_t *counts, int32_t *plen);

/* prof2class on the device (src/prof2class.c:165-258): decode the RELATIVE profiles of a batch (counts
 * of each read's k-mers in a genome / haplotype k-mer table) and map them to ground-truth class
 * strings, count 0 -> 'E', 1 -> 'H', 2 -> 'D', more -> 'R', behind K-1 'N's.  cls receives rlen[i]
 * characters per read, reads concatenated; reads shorter than K get rlen[i] 'N's.  K is the model's
 * k-mer length.  status (may be NULL) flags reads whose profile length is not rlen-K+1. */
int cpg_prof2class(cpg_ctx *ctx, int32_t n_reads, const uint8_t *prof, const int64_t *prof_off,
                   const int32_t *rlen, uint8_t *cls, int32_t *status);

/* Device-resident timing: upload once, run the kernels `iters` times on data already in HBM,
 * report the mean device time of each kernel in milliseconds (CUDA events on the context's
 * stream), then fetch the result of the last run. */
int cpg_upload(cpg_ctx *ctx, const cpg_batch *batch);
int cpg_run_resident(cpg_ctx *ctx, int iters, float *ms_decode, float *ms_classify, int *launches);
int cpg_download(cpg_ctx *ctx, cpg_result *result);
/* Device time of the classification phases in the last cpg_run_resident iteration, nanoseconds
 * (CUDA events between the kernels): [0] the three wall kernels (wall detection + reliable intervals), [1] k_rel
 * (reliable-interval DP), [2] k_unrel_a + k_unrel_b + k_emit (unreliable intervals + class strings), [3] the retry launch.
 * With CPG_FUSED=1 (single-kernel path): summed per-group clock cycles of the three phases and of
 * the waits at the CTA phase barriers. */
int cpg_phase_cycles(cpg_ctx *ctx, uint64_t out[4]);
/* ... and of the kernels inside those phases: [0] k_wall_a (pure, one candidate per lane), [1] k_wall_b
 * (order-dependent replay, one read per lane group), [2] k_wall_c (one interval per lane), [3] k_unrel_a (pure,
 * one interval per lane), [4] k_unrel_b (sweeps), [5] k_emit (class strings). */
int cpg_wall_ns(cpg_ctx *ctx, uint64_t out[6]);
/* Sums over the reads of the resident batch after cpg_run_resident: [0] wall candidates, [1] intervals,
 * [2] reliable intervals, [3] intervals the unreliable sweeps visit (the units of the per-kernel byte counts). */
int cpg_batch_stats(cpg_ctx *ctx, int64_t out[4]);

/* ---- profile producer (what FastK does before ClassPro runs) ----------------
 * FastK is not part of the reference tree; the reference only reads its files (src/libfastk.c:51-96
 * histogram, :1238-1386 profile index, :1414-1562 Fetch_Profile).  These two calls produce what
 * those readers expect, on the GPU, from the 2-bit packed reads; they need no model and no cpg_ctx.
 *
 * cpg_count_kmers: exact counts of the canonical k-mers (a k-mer and its reverse complement are
 * one) of the whole read set at every read position, saturated at 32767.
 *   seq, seq_off, rlen : as in cpg_batch with seq_bits = 2
 *   cnt_off  : [n_reads+1] OUT, prefix sums of max(rlen-K+1,0)
 *   counts   : [cnt_off[n_reads]] OUT, counts of read r at counts[cnt_off[r] ..)  (= what
 *              Fetch_Profile decodes, src/libfastk.c:1414-1562)
 *   hist     : [32770] OUT, hist[c] = number of DISTINCT k-mers with count c, 1 <= c <= 32767 (the
 *              bins of <root>.hist with low = 1, high = 32767, src/libfastk.c:72-83); hist[32768] and
 *              hist[32769] = the instance-mode values of the two boundary bins (the file's two
 *              hidden words, src/libfastk.c:91-93)
 * 1 <= K <= 40.  Needs ~40 bytes of device memory per k-mer; a read set that does not fit is counted in
 * key-range passes (every pass extracts the keys again and sorts the ones whose hash falls in its range:
 * equal k-mers always meet in the same pass), fewer than 2^32 k-mers per pass. */
int cpg_count_kmers(int device, int32_t kmer, int32_t n_reads, const uint8_t *seq, const int64_t *seq_off,
                    const int32_t *rlen, int64_t *cnt_off, uint16_t *counts, int64_t *hist);
/* cpg_encode_profiles: the encoder side of the profile codec (decoder: src/libfastk.c:1467-1535):
 * prof receives the token streams read after read, prof_off[n_reads+1] their byte offsets (= the
 * .pidx index of a part).  prof_cap = capacity of prof in bytes (2 bytes per count always suffice). */
int cpg_encode_profiles(int device, int32_t n_reads, const uint16_t *counts, const int64_t *cnt_off,
                        uint8_t *prof, int64_t prof_cap, int64_t *prof_off);
const char *cpg_count_error(void);      /* text of the last error of the two calls above (per thread) */

/* Pinned (page-locked) host memory, so that the copies of cpg_submit/cpg_collect are truly
 * asynchronous DMA transfers; pageable buffers work too but are staged by the driver. */
void *cpg_host_alloc(size_t bytes);
void  cpg_host_free(void *p);

const char *cpg_status_string(int32_t status);
const char *cpg_version(void);

#ifdef __cplusplus
}
#endif
#endif
