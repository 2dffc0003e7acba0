/* Stand-in for <gsl/gsl_multifit.h>, used ONLY to let the unmodified reference
 * (src/wall.c:9, polynomialfit at src/wall.c:11-42) compile without GSL 2.7,
 * whose tarball is absent from /root/reference (.MISSING_LARGE_BLOBS).
 * polynomialfit() is reached only with the -M <model_path> option
 * (src/wall.c:101 <- load_himodel <- load_emodel when name != NULL); every entry
 * point here aborts, so a -M run fails loudly instead of producing unpinned numbers.
 * This file is test infrastructure (oracle/), never part of the product. */
#ifndef ORACLE_GSL_MULTIFIT_STUB_H
#define ORACLE_GSL_MULTIFIT_STUB_H
#include <stdio.h>
#include <stdlib.h>

typedef struct { int unused; } gsl_matrix;
typedef struct { int unused; } gsl_vector;
typedef struct { int unused; } gsl_multifit_linear_workspace;

static inline void *oracle_gsl_unavailable(void)
{ fprintf(stderr,"[oracle] GSL is not available in this image: the -M option is unsupported\n");
  abort();
  return NULL;
}

static inline gsl_matrix *gsl_matrix_alloc(size_t a, size_t b)
{ (void)a; (void)b; return (gsl_matrix *)oracle_gsl_unavailable(); }
static inline gsl_vector *gsl_vector_alloc(size_t a)
{ (void)a; return (gsl_vector *)oracle_gsl_unavailable(); }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x)
{ (void)m; (void)i; (void)j; (void)x; oracle_gsl_unavailable(); }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x)
{ (void)v; (void)i; (void)x; oracle_gsl_unavailable(); }
static inline double gsl_vector_get(const gsl_vector *v, size_t i)
{ (void)v; (void)i; oracle_gsl_unavailable(); return 0.; }
static inline gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(size_t n, size_t p)
{ (void)n; (void)p; return (gsl_multifit_linear_workspace *)oracle_gsl_unavailable(); }
static inline int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c,
                                      gsl_matrix *cov, double *chisq,
                                      gsl_multifit_linear_workspace *w)
{ (void)X; (void)y; (void)c; (void)cov; (void)chisq; (void)w; oracle_gsl_unavailable(); return 0; }
static inline void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w) { (void)w; }
static inline void gsl_matrix_free(gsl_matrix *m) { (void)m; }
static inline void gsl_vector_free(gsl_vector *v) { (void)v; }
#endif
