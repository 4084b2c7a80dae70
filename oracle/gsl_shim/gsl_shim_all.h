/*
 * gsl_shim_all.h -- a minimal, self-written stand-in for the slice of the GNU GSL
 * C API that MADAIEmulator's libEmu uses.
 *
 * TEST INFRASTRUCTURE ONLY.  GSL is not installed in this image and is not vendored by
 * the reference, so the reference's own hot-path sources (/root/reference/src/libEmu/*.c
 * etc.) are compiled UNMODIFIED against this shim to obtain `oracle/_ref/libemu_ref.so`,
 * the parity oracle.  Nothing in the product path (madaiemulator_b200/) includes this.
 *
 * Struct layouts match real GSL (gsl_block/gsl_vector/gsl_matrix and their views) so
 * objects are interchangeable with a real GSL build.  All numerical routines are written
 * from their textbook definitions (row-major reference BLAS loops, Cholesky-Crout,
 * triangular inversion, cyclic Jacobi, MT19937, a Fletcher line-search BFGS), not taken
 * from GSL.
 */
#ifndef GSL_SHIM_ALL_H
#define GSL_SHIM_ALL_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errno ------------------------------------------------------------------------- */
enum {
	GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2,
	GSL_EDOM = 1, GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5,
	GSL_EFACTOR = 6, GSL_ESANITY = 7, GSL_ENOMEM = 8, GSL_EBADFUNC = 9, GSL_ERUNAWAY = 10,
	GSL_EMAXITER = 11, GSL_EZERODIV = 12, GSL_EBADTOL = 13, GSL_ETOL = 14, GSL_EUNDRFLW = 15,
	GSL_EOVRFLW = 16, GSL_ELOSS = 17, GSL_EROUND = 18, GSL_EBADLEN = 19, GSL_ENOTSQR = 20,
	GSL_ESING = 21, GSL_EDIVERGE = 22, GSL_EUNSUP = 23, GSL_EUNIMPL = 24, GSL_ECACHE = 25,
	GSL_ETABLE = 26, GSL_ENOPROG = 27, GSL_ENOPROGJ = 28, GSL_ETOLF = 29, GSL_ETOLX = 30,
	GSL_ETOLG = 31, GSL_EOF = 32
};
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);
gsl_error_handler_t *gsl_set_error_handler_off(void);
const char *gsl_strerror(const int gsl_errno);
void gsl_error(const char *reason, const char *file, int line, int gsl_errno);

/* ---- math -------------------------------------------------------------------------- */
#define GSL_NAN    (NAN)
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_MAX(a,b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a,b) ((a) < (b) ? (a) : (b))

/* ---- block / vector / matrix ---------------------------------------------------------- */
typedef struct { size_t size; double *data; } gsl_block;

typedef struct {
	size_t size; size_t stride; double *data; gsl_block *block; int owner;
} gsl_vector;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef gsl_vector_view gsl_vector_const_view;

typedef struct {
	size_t size1; size_t size2; size_t tda; double *data; gsl_block *block; int owner;
} gsl_matrix;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;

/* float vector only appears in a comment of the reference; declare the name anyway */
typedef struct { size_t size; size_t stride; float *data; void *block; int owner; } gsl_vector_float;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
static inline double gsl_vector_get(const gsl_vector *v, size_t i){ return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x){ v->data[i * v->stride] = x; }
static inline double *gsl_vector_ptr(gsl_vector *v, size_t i){ return v->data + i * v->stride; }
void gsl_vector_set_zero(gsl_vector *v);
void gsl_vector_set_all(gsl_vector *v, double x);
int gsl_vector_memcpy(gsl_vector *dst, const gsl_vector *src);
int gsl_vector_add(gsl_vector *a, const gsl_vector *b);
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b);
int gsl_vector_scale(gsl_vector *a, double x);
double gsl_vector_min(const gsl_vector *v);
double gsl_vector_max(const gsl_vector *v);
gsl_vector_view gsl_vector_view_array(double *base, size_t n);

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j){ return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x){ m->data[i * m->tda + j] = x; }
static inline double *gsl_matrix_ptr(gsl_matrix *m, size_t i, size_t j){ return m->data + i * m->tda + j; }
void gsl_matrix_set_zero(gsl_matrix *m);
void gsl_matrix_set_identity(gsl_matrix *m);
void gsl_matrix_set_all(gsl_matrix *m, double x);
int gsl_matrix_memcpy(gsl_matrix *dst, const gsl_matrix *src);
int gsl_matrix_scale(gsl_matrix *a, double x);
int gsl_matrix_set_row(gsl_matrix *m, size_t i, const gsl_vector *v);
int gsl_matrix_set_col(gsl_matrix *m, size_t j, const gsl_vector *v);
int gsl_matrix_get_row(gsl_vector *v, const gsl_matrix *m, size_t i);
int gsl_matrix_get_col(gsl_vector *v, const gsl_matrix *m, size_t j);
gsl_vector_view gsl_matrix_row(gsl_matrix *m, size_t i);
gsl_vector_view gsl_matrix_column(gsl_matrix *m, size_t j);
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2);

/* ---- BLAS ----------------------------------------------------------------------------- */
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER_t;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO_t;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG_t;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE_t;

int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result);
double gsl_blas_dnrm2(const gsl_vector *x);
int gsl_blas_daxpy(double alpha, const gsl_vector *x, gsl_vector *y);
int gsl_blas_dscal(double alpha, gsl_vector *x);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A,
                   const gsl_vector *x, double beta, gsl_vector *y);
int gsl_blas_dsymv(CBLAS_UPLO_t Uplo, double alpha, const gsl_matrix *A,
                   const gsl_vector *x, double beta, gsl_vector *y);
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix *A, const gsl_matrix *B, double beta, gsl_matrix *C);

/* ---- linalg ----------------------------------------------------------------------------- */
int gsl_linalg_cholesky_decomp(gsl_matrix *A);
int gsl_linalg_cholesky_invert(gsl_matrix *cholesky);
int gsl_linalg_cholesky_solve(const gsl_matrix *cholesky, const gsl_vector *b, gsl_vector *x);

/* ---- rng (MT19937 only) ---------------------------------------------------------------- */
typedef struct { const char *name; unsigned long max, min; size_t size; } gsl_rng_type;
typedef struct { const gsl_rng_type *type; void *state; } gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
extern const gsl_rng_type *gsl_rng_mt19937;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(const gsl_rng *r, unsigned long seed);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long gsl_rng_get(const gsl_rng *r);
void gsl_rng_free(gsl_rng *r);

/* ---- eigen (symmetric, Jacobi) ------------------------------------------------------------ */
typedef struct { size_t size; double *d; double *sd; double *gc; double *gs; } gsl_eigen_symmv_workspace;
typedef enum {
	GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC, GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC
} gsl_eigen_sort_t;
gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(const size_t n);
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w);
int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w);
int gsl_eigen_symmv_sort(gsl_vector *eval, gsl_matrix *evec, gsl_eigen_sort_t sort_type);

/* ---- multimin (fdf minimizers; only a BFGS with Fletcher line search is provided) ------------ */
typedef struct {
	double (*f)(const gsl_vector *x, void *params);
	void (*df)(const gsl_vector *x, void *params, gsl_vector *df);
	void (*fdf)(const gsl_vector *x, void *params, double *f, gsl_vector *df);
	size_t n;
	void *params;
} gsl_multimin_function_fdf;

typedef struct gsl_multimin_fdfminimizer_type_s {
	const char *name;
	size_t size;
	int (*alloc)(void *state, size_t n);
	int (*set)(void *state, gsl_multimin_function_fdf *fdf, const gsl_vector *x, double *f,
	           gsl_vector *gradient, double step_size, double tol);
	int (*iterate)(void *state, gsl_multimin_function_fdf *fdf, gsl_vector *x, double *f,
	               gsl_vector *gradient, gsl_vector *dx);
	int (*restart)(void *state);
	void (*free)(void *state);
} gsl_multimin_fdfminimizer_type;

typedef struct {
	const gsl_multimin_fdfminimizer_type *type;
	gsl_multimin_function_fdf *fdf;
	double f;
	gsl_vector *x;
	gsl_vector *gradient;
	gsl_vector *dx;
	void *state;
} gsl_multimin_fdfminimizer;

extern const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_vector_bfgs2;
extern const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_vector_bfgs;
extern const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_conjugate_fr;

gsl_multimin_fdfminimizer *gsl_multimin_fdfminimizer_alloc(const gsl_multimin_fdfminimizer_type *T, size_t n);
int gsl_multimin_fdfminimizer_set(gsl_multimin_fdfminimizer *s, gsl_multimin_function_fdf *fdf,
                                  const gsl_vector *x, double step_size, double tol);
int gsl_multimin_fdfminimizer_iterate(gsl_multimin_fdfminimizer *s);
int gsl_multimin_fdfminimizer_restart(gsl_multimin_fdfminimizer *s);
void gsl_multimin_fdfminimizer_free(gsl_multimin_fdfminimizer *s);
const char *gsl_multimin_fdfminimizer_name(const gsl_multimin_fdfminimizer *s);
gsl_vector *gsl_multimin_fdfminimizer_x(const gsl_multimin_fdfminimizer *s);
gsl_vector *gsl_multimin_fdfminimizer_dx(const gsl_multimin_fdfminimizer *s);
gsl_vector *gsl_multimin_fdfminimizer_gradient(const gsl_multimin_fdfminimizer *s);
double gsl_multimin_fdfminimizer_minimum(const gsl_multimin_fdfminimizer *s);
int gsl_multimin_test_gradient(const gsl_vector *g, double epsabs);

#ifdef __cplusplus
}
#endif
#endif /* GSL_SHIM_ALL_H */
