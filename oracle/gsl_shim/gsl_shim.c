/*
 * gsl_shim.c -- implementation of the minimal GSL-API shim declared in gsl_shim_all.h.
 *
 * TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline); never linked into the product.
 * Every routine is written from its textbook definition:
 *   - BLAS: straightforward row-major loops (the same cost model as the gslcblas reference
 *     BLAS the reference links, CMakeLists.txt:64),
 *   - Cholesky: row-wise Cholesky-Banachiewicz; inverse: forward substitution for L^-1 then
 *     L^-T L^-1, symmetric fill (semantics of gsl_linalg_cholesky_decomp/_invert as called at
 *     reference src/libEmu/maxmultimin.c:325,361),
 *   - RNG: MT19937 (Matsumoto & Nishimura 1998/2002 initialisation),
 *   - eigen: cyclic Jacobi,
 *   - multimin "vector_bfgs2": adaptor onto madaiemulator_b200/host/emub_bfgs.c.
 */
#include "gsl_shim_all.h"
#include "../../madaiemulator_b200/host/emub_bfgs.h"
#include <string.h>

/* ---- errors ------------------------------------------------------------------------------ */
static gsl_error_handler_t *g_handler = NULL;
static int g_handler_off = 0;
static void no_error_handler(const char *r, const char *f, int l, int e) { (void)r; (void)f; (void)l; (void)e; }

gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h)
{
	gsl_error_handler_t *prev = g_handler_off ? no_error_handler : g_handler;
	if (h == no_error_handler) { g_handler_off = 1; g_handler = NULL; }
	else { g_handler_off = 0; g_handler = h; }
	return prev;
}
gsl_error_handler_t *gsl_set_error_handler_off(void)
{
	gsl_error_handler_t *prev = g_handler_off ? no_error_handler : g_handler;
	g_handler_off = 1; g_handler = NULL;
	return prev;
}
void gsl_error(const char *reason, const char *file, int line, int gsl_errno)
{
	if (g_handler_off) return;
	if (g_handler) { g_handler(reason, file, line, gsl_errno); return; }
	fprintf(stderr, "gsl(shim): %s:%d: ERROR: %s\n", file, line, reason);
	abort();
}
const char *gsl_strerror(const int e)
{
	switch (e) {
	case GSL_SUCCESS: return "success";
	case GSL_FAILURE: return "failure";
	case GSL_CONTINUE: return "the iteration has not converged yet";
	case GSL_EDOM: return "input domain error";
	case GSL_ENOPROG: return "iteration is not making progress towards solution";
	case GSL_EINVAL: return "invalid argument supplied by user";
	default: return "unknown error code";
	}
}

/* ---- vector / matrix ------------------------------------------------------------------------ */
static gsl_block *block_alloc(size_t n)
{
	gsl_block *b = (gsl_block *)malloc(sizeof(gsl_block));
	b->size = n;
	b->data = (double *)malloc((n ? n : 1) * sizeof(double));
	return b;
}
gsl_vector *gsl_vector_alloc(size_t n)
{
	gsl_vector *v = (gsl_vector *)malloc(sizeof(gsl_vector));
	v->block = block_alloc(n);
	v->data = v->block->data; v->size = n; v->stride = 1; v->owner = 1;
	return v;
}
gsl_vector *gsl_vector_calloc(size_t n)
{
	gsl_vector *v = gsl_vector_alloc(n);
	memset(v->data, 0, n * sizeof(double));
	return v;
}
void gsl_vector_free(gsl_vector *v)
{
	if (!v) return;
	if (v->owner && v->block) { free(v->block->data); free(v->block); }
	free(v);
}
void gsl_vector_set_zero(gsl_vector *v) { for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = 0.0; }
void gsl_vector_set_all(gsl_vector *v, double x) { for (size_t i = 0; i < v->size; i++) v->data[i * v->stride] = x; }
int gsl_vector_memcpy(gsl_vector *d, const gsl_vector *s)
{
	if (d->size != s->size) { gsl_error("vector lengths are not equal", __FILE__, __LINE__, GSL_EBADLEN); return GSL_EBADLEN; }
	for (size_t i = 0; i < s->size; i++) d->data[i * d->stride] = s->data[i * s->stride];
	return GSL_SUCCESS;
}
int gsl_vector_add(gsl_vector *a, const gsl_vector *b)
{
	for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] += b->data[i * b->stride];
	return GSL_SUCCESS;
}
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b)
{
	for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] -= b->data[i * b->stride];
	return GSL_SUCCESS;
}
int gsl_vector_scale(gsl_vector *a, double x)
{
	for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] *= x;
	return GSL_SUCCESS;
}
double gsl_vector_min(const gsl_vector *v)
{
	double m = v->data[0];
	for (size_t i = 1; i < v->size; i++) { double x = v->data[i * v->stride]; if (x < m) m = x; if (isnan(x)) return x; }
	return m;
}
double gsl_vector_max(const gsl_vector *v)
{
	double m = v->data[0];
	for (size_t i = 1; i < v->size; i++) { double x = v->data[i * v->stride]; if (x > m) m = x; if (isnan(x)) return x; }
	return m;
}
gsl_vector_view gsl_vector_view_array(double *base, size_t n)
{
	gsl_vector_view w;
	w.vector.size = n; w.vector.stride = 1; w.vector.data = base; w.vector.block = NULL; w.vector.owner = 0;
	return w;
}

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{
	gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix));
	m->block = block_alloc(n1 * n2);
	m->data = m->block->data; m->size1 = n1; m->size2 = n2; m->tda = n2; m->owner = 1;
	return m;
}
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2)
{
	gsl_matrix *m = gsl_matrix_alloc(n1, n2);
	memset(m->data, 0, n1 * n2 * sizeof(double));
	return m;
}
void gsl_matrix_free(gsl_matrix *m)
{
	if (!m) return;
	if (m->owner && m->block) { free(m->block->data); free(m->block); }
	free(m);
}
void gsl_matrix_set_zero(gsl_matrix *m)
{
	for (size_t i = 0; i < m->size1; i++) memset(m->data + i * m->tda, 0, m->size2 * sizeof(double));
}
void gsl_matrix_set_all(gsl_matrix *m, double x)
{
	for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = x;
}
void gsl_matrix_set_identity(gsl_matrix *m)
{
	for (size_t i = 0; i < m->size1; i++) for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = (i == j) ? 1.0 : 0.0;
}
int gsl_matrix_memcpy(gsl_matrix *d, const gsl_matrix *s)
{
	if (d->size1 != s->size1 || d->size2 != s->size2) {
		gsl_error("matrix sizes are different", __FILE__, __LINE__, GSL_EBADLEN);
		return GSL_EBADLEN;
	}
	for (size_t i = 0; i < s->size1; i++) memcpy(d->data + i * d->tda, s->data + i * s->tda, s->size2 * sizeof(double));
	return GSL_SUCCESS;
}
int gsl_matrix_scale(gsl_matrix *a, double x)
{
	for (size_t i = 0; i < a->size1; i++) for (size_t j = 0; j < a->size2; j++) a->data[i * a->tda + j] *= x;
	return GSL_SUCCESS;
}
int gsl_matrix_set_row(gsl_matrix *m, size_t i, const gsl_vector *v)
{
	for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = v->data[j * v->stride];
	return GSL_SUCCESS;
}
int gsl_matrix_set_col(gsl_matrix *m, size_t j, const gsl_vector *v)
{
	for (size_t i = 0; i < m->size1; i++) m->data[i * m->tda + j] = v->data[i * v->stride];
	return GSL_SUCCESS;
}
int gsl_matrix_get_row(gsl_vector *v, const gsl_matrix *m, size_t i)
{
	for (size_t j = 0; j < m->size2; j++) v->data[j * v->stride] = m->data[i * m->tda + j];
	return GSL_SUCCESS;
}
int gsl_matrix_get_col(gsl_vector *v, const gsl_matrix *m, size_t j)
{
	for (size_t i = 0; i < m->size1; i++) v->data[i * v->stride] = m->data[i * m->tda + j];
	return GSL_SUCCESS;
}
gsl_vector_view gsl_matrix_row(gsl_matrix *m, size_t i)
{
	gsl_vector_view w;
	w.vector.size = m->size2; w.vector.stride = 1; w.vector.data = m->data + i * m->tda;
	w.vector.block = m->block; w.vector.owner = 0;
	return w;
}
gsl_vector_view gsl_matrix_column(gsl_matrix *m, size_t j)
{
	gsl_vector_view w;
	w.vector.size = m->size1; w.vector.stride = m->tda; w.vector.data = m->data + j;
	w.vector.block = m->block; w.vector.owner = 0;
	return w;
}
gsl_matrix_view gsl_matrix_view_array(double *base, size_t n1, size_t n2)
{
	gsl_matrix_view w;
	w.matrix.size1 = n1; w.matrix.size2 = n2; w.matrix.tda = n2; w.matrix.data = base;
	w.matrix.block = NULL; w.matrix.owner = 0;
	return w;
}

/* ---- BLAS ----------------------------------------------------------------------------------- */
int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result)
{
	double s = 0.0;
	if (x->size != y->size) { gsl_error("invalid length", __FILE__, __LINE__, GSL_EBADLEN); return GSL_EBADLEN; }
	for (size_t i = 0; i < x->size; i++) s += x->data[i * x->stride] * y->data[i * y->stride];
	*result = s;
	return GSL_SUCCESS;
}
double gsl_blas_dnrm2(const gsl_vector *x)
{
	double s = 0.0;
	for (size_t i = 0; i < x->size; i++) { double v = x->data[i * x->stride]; s += v * v; }
	return sqrt(s);
}
int gsl_blas_daxpy(double alpha, const gsl_vector *x, gsl_vector *y)
{
	for (size_t i = 0; i < x->size; i++) y->data[i * y->stride] += alpha * x->data[i * x->stride];
	return GSL_SUCCESS;
}
int gsl_blas_dscal(double alpha, gsl_vector *x) { return gsl_vector_scale(x, alpha); }

int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A, const gsl_vector *x,
                   double beta, gsl_vector *y)
{
	const size_t M = A->size1, N = A->size2;
	const size_t leny = (TransA == CblasNoTrans) ? M : N, lenx = (TransA == CblasNoTrans) ? N : M;
	if (x->size != lenx || y->size != leny) { gsl_error("invalid length", __FILE__, __LINE__, GSL_EBADLEN); return GSL_EBADLEN; }
	if (beta == 0.0) for (size_t i = 0; i < leny; i++) y->data[i * y->stride] = 0.0;
	else if (beta != 1.0) for (size_t i = 0; i < leny; i++) y->data[i * y->stride] *= beta;
	if (alpha == 0.0) return GSL_SUCCESS;
	if (TransA == CblasNoTrans) {
		for (size_t i = 0; i < M; i++) {
			const double *row = A->data + i * A->tda;
			double s = 0.0;
			for (size_t j = 0; j < N; j++) s += x->data[j * x->stride] * row[j];
			y->data[i * y->stride] += alpha * s;
		}
	} else {
		for (size_t j = 0; j < M; j++) {
			const double *row = A->data + j * A->tda;
			const double t = alpha * x->data[j * x->stride];
			if (t != 0.0) for (size_t i = 0; i < N; i++) y->data[i * y->stride] += t * row[i];
		}
	}
	return GSL_SUCCESS;
}
int gsl_blas_dsymv(CBLAS_UPLO_t Uplo, double alpha, const gsl_matrix *A, const gsl_vector *x, double beta, gsl_vector *y)
{
	const size_t N = A->size1;
	if (beta == 0.0) for (size_t i = 0; i < N; i++) y->data[i * y->stride] = 0.0;
	else if (beta != 1.0) for (size_t i = 0; i < N; i++) y->data[i * y->stride] *= beta;
	for (size_t i = 0; i < N; i++) {
		double s = 0.0;
		for (size_t j = 0; j < N; j++) {
			double a;
			if (Uplo == CblasUpper) a = (j >= i) ? A->data[i * A->tda + j] : A->data[j * A->tda + i];
			else a = (j <= i) ? A->data[i * A->tda + j] : A->data[j * A->tda + i];
			s += a * x->data[j * x->stride];
		}
		y->data[i * y->stride] += alpha * s;
	}
	return GSL_SUCCESS;
}
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha, const gsl_matrix *A,
                   const gsl_matrix *B, double beta, gsl_matrix *C)
{
	const size_t M = C->size1, N = C->size2;
	const size_t MA = (TransA == CblasNoTrans) ? A->size1 : A->size2;
	const size_t K = (TransA == CblasNoTrans) ? A->size2 : A->size1;
	const size_t KB = (TransB == CblasNoTrans) ? B->size1 : B->size2;
	const size_t NB = (TransB == CblasNoTrans) ? B->size2 : B->size1;
	if (M != MA || N != NB || K != KB) { gsl_error("invalid length", __FILE__, __LINE__, GSL_EBADLEN); return GSL_EBADLEN; }
	for (size_t i = 0; i < M; i++) {
		double *c = C->data + i * C->tda;
		if (beta == 0.0) for (size_t j = 0; j < N; j++) c[j] = 0.0;
		else if (beta != 1.0) for (size_t j = 0; j < N; j++) c[j] *= beta;
	}
	if (alpha == 0.0) return GSL_SUCCESS;
	if (TransA == CblasNoTrans && TransB == CblasNoTrans) {
		for (size_t i = 0; i < M; i++) {
			double *c = C->data + i * C->tda;
			for (size_t k = 0; k < K; k++) {
				const double t = alpha * A->data[i * A->tda + k];
				const double *b = B->data + k * B->tda;
				if (t != 0.0) for (size_t j = 0; j < N; j++) c[j] += t * b[j];
			}
		}
	} else if (TransA == CblasNoTrans && TransB == CblasTrans) {
		for (size_t i = 0; i < M; i++) {
			const double *a = A->data + i * A->tda;
			for (size_t j = 0; j < N; j++) {
				const double *b = B->data + j * B->tda;
				double s = 0.0;
				for (size_t k = 0; k < K; k++) s += a[k] * b[k];
				C->data[i * C->tda + j] += alpha * s;
			}
		}
	} else if (TransA == CblasTrans && TransB == CblasNoTrans) {
		for (size_t k = 0; k < K; k++) {
			const double *b = B->data + k * B->tda;
			for (size_t i = 0; i < M; i++) {
				const double t = alpha * A->data[k * A->tda + i];
				double *c = C->data + i * C->tda;
				if (t != 0.0) for (size_t j = 0; j < N; j++) c[j] += t * b[j];
			}
		}
	} else {
		for (size_t i = 0; i < M; i++)
			for (size_t j = 0; j < N; j++) {
				double s = 0.0;
				for (size_t k = 0; k < K; k++) s += A->data[k * A->tda + i] * B->data[j * B->tda + k];
				C->data[i * C->tda + j] += alpha * s;
			}
	}
	return GSL_SUCCESS;
}

/* ---- Cholesky --------------------------------------------------------------------------------- */
int gsl_linalg_cholesky_decomp(gsl_matrix *A)
{
	const size_t n = A->size1, ld = A->tda;
	double *a = A->data;
	int status = GSL_SUCCESS;
	if (A->size1 != A->size2) { gsl_error("cholesky decomposition requires square matrix", __FILE__, __LINE__, GSL_ENOTSQR); return GSL_ENOTSQR; }
	for (size_t i = 0; i < n; i++) {
		double *ri = a + i * ld;
		for (size_t j = 0; j <= i; j++) {
			const double *rj = a + j * ld;
			double s = ri[j];
			for (size_t k = 0; k < j; k++) s -= ri[k] * rj[k];
			if (j == i) {
				if (!(s > 0.0)) { status = GSL_EDOM; s = fabs(s); if (!(s > 0.0)) s = 1.0; }
				ri[i] = sqrt(s);
			} else {
				ri[j] = s / rj[j];
			}
		}
	}
	/* L^T into the upper triangle */
	for (size_t i = 0; i < n; i++) for (size_t j = i + 1; j < n; j++) a[i * ld + j] = a[j * ld + i];
	if (status == GSL_EDOM) gsl_error("matrix must be positive definite", __FILE__, __LINE__, GSL_EDOM);
	return status;
}

int gsl_linalg_cholesky_invert(gsl_matrix *LLT)
{
	const size_t n = LLT->size1, ld = LLT->tda;
	double *a = LLT->data;
	double *x = (double *)calloc(n * n, sizeof(double)); /* X = L^-1, lower, row-major */
	if (!x) return GSL_ENOMEM;
	for (size_t i = 0; i < n; i++) {
		double *xi = x + i * n;
		const double *li = a + i * ld;
		/* row i of X:  (e_i - sum_{k<i} L[i][k] X[k][:]) / L[i][i] */
		for (size_t k = 0; k < i; k++) {
			const double t = li[k];
			const double *xk = x + k * n;
			if (t != 0.0) for (size_t j = 0; j <= k; j++) xi[j] -= t * xk[j];
		}
		xi[i] = 1.0;
		const double inv = 1.0 / li[i];
		for (size_t j = 0; j <= i; j++) xi[j] *= inv;
	}
	/* A^-1 = X^T X, lower triangle by rank-one accumulation */
	for (size_t i = 0; i < n; i++) for (size_t j = 0; j <= i; j++) a[i * ld + j] = 0.0;
	for (size_t k = 0; k < n; k++) {
		const double *xk = x + k * n;
		for (size_t i = 0; i <= k; i++) {
			const double t = xk[i];
			double *ai = a + i * ld;
			if (t != 0.0) for (size_t j = 0; j <= i; j++) ai[j] += t * xk[j];
		}
	}
	for (size_t i = 0; i < n; i++) for (size_t j = i + 1; j < n; j++) a[i * ld + j] = a[j * ld + i];
	free(x);
	return GSL_SUCCESS;
}

int gsl_linalg_cholesky_solve(const gsl_matrix *LLT, const gsl_vector *b, gsl_vector *xv)
{
	const size_t n = LLT->size1, ld = LLT->tda;
	const double *a = LLT->data;
	gsl_vector_memcpy(xv, b);
	for (size_t i = 0; i < n; i++) {
		double s = xv->data[i * xv->stride];
		for (size_t k = 0; k < i; k++) s -= a[i * ld + k] * xv->data[k * xv->stride];
		xv->data[i * xv->stride] = s / a[i * ld + i];
	}
	for (size_t ii = n; ii-- > 0;) {
		double s = xv->data[ii * xv->stride];
		for (size_t k = ii + 1; k < n; k++) s -= a[k * ld + ii] * xv->data[k * xv->stride];
		xv->data[ii * xv->stride] = s / a[ii * ld + ii];
	}
	return GSL_SUCCESS;
}

/* ---- MT19937 ------------------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
typedef struct { unsigned long mt[MT_N]; int mti; } mt_state_t;
static const gsl_rng_type mt_type = { "mt19937", 0xffffffffUL, 0, sizeof(mt_state_t) };
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_default = &mt_type;

void gsl_rng_set(const gsl_rng *r, unsigned long s)
{
	mt_state_t *st = (mt_state_t *)r->state;
	if (s == 0) s = 4357;
	st->mt[0] = s & 0xffffffffUL;
	for (int i = 1; i < MT_N; i++)
		st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
	st->mti = MT_N;
}
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
	gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
	r->type = T;
	r->state = calloc(1, sizeof(mt_state_t));
	gsl_rng_set(r, 0);
	return r;
}
void gsl_rng_free(gsl_rng *r) { if (r) { free(r->state); free(r); } }
unsigned long gsl_rng_get(const gsl_rng *r)
{
	mt_state_t *st = (mt_state_t *)r->state;
	unsigned long *mt = st->mt, y;
	if (st->mti >= MT_N) {
		int kk;
		for (kk = 0; kk < MT_N - MT_M; kk++) {
			y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
			mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
		}
		for (; kk < MT_N - 1; kk++) {
			y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
			mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
		}
		y = (mt[MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
		mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
		st->mti = 0;
	}
	y = mt[st->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680UL;
	y ^= (y << 15) & 0xefc60000UL;
	y ^= (y >> 18);
	return y & 0xffffffffUL;
}
double gsl_rng_uniform(const gsl_rng *r) { return (double)gsl_rng_get(r) / 4294967296.0; }

/* ---- symmetric eigenproblem: cyclic Jacobi ----------------------------------------------------------- */
gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(const size_t n)
{
	gsl_eigen_symmv_workspace *w = (gsl_eigen_symmv_workspace *)calloc(1, sizeof(*w));
	w->size = n;
	return w;
}
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w) { free(w); }

int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w)
{
	const size_t n = A->size1;
	(void)w;
	gsl_matrix_set_identity(evec);
	for (int sweep = 0; sweep < 100; sweep++) {
		double off = 0.0;
		for (size_t p = 0; p < n; p++) for (size_t q = p + 1; q < n; q++) off += gsl_matrix_get(A, p, q) * gsl_matrix_get(A, p, q);
		if (off < 1e-300) break;
		for (size_t p = 0; p < n; p++) {
			for (size_t q = p + 1; q < n; q++) {
				const double apq = gsl_matrix_get(A, p, q);
				if (apq == 0.0) continue;
				const double app = gsl_matrix_get(A, p, p), aqq = gsl_matrix_get(A, q, q);
				const double tau = (aqq - app) / (2.0 * apq);
				const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
				const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
				for (size_t k = 0; k < n; k++) {
					const double akp = gsl_matrix_get(A, k, p), akq = gsl_matrix_get(A, k, q);
					gsl_matrix_set(A, k, p, c * akp - s * akq);
					gsl_matrix_set(A, k, q, s * akp + c * akq);
				}
				for (size_t k = 0; k < n; k++) {
					const double apk = gsl_matrix_get(A, p, k), aqk = gsl_matrix_get(A, q, k);
					gsl_matrix_set(A, p, k, c * apk - s * aqk);
					gsl_matrix_set(A, q, k, s * apk + c * aqk);
				}
				for (size_t k = 0; k < n; k++) {
					const double vkp = gsl_matrix_get(evec, k, p), vkq = gsl_matrix_get(evec, k, q);
					gsl_matrix_set(evec, k, p, c * vkp - s * vkq);
					gsl_matrix_set(evec, k, q, s * vkp + c * vkq);
				}
			}
		}
	}
	for (size_t i = 0; i < n; i++) gsl_vector_set(eval, i, gsl_matrix_get(A, i, i));
	return GSL_SUCCESS;
}

int gsl_eigen_symmv_sort(gsl_vector *eval, gsl_matrix *evec, gsl_eigen_sort_t sort_type)
{
	const size_t n = eval->size;
	for (size_t i = 0; i + 1 < n; i++) {
		size_t k = i;
		double ek = gsl_vector_get(eval, i);
		for (size_t j = i + 1; j < n; j++) {
			const double ej = gsl_vector_get(eval, j);
			int better;
			switch (sort_type) {
			case GSL_EIGEN_SORT_VAL_ASC: better = ej < ek; break;
			case GSL_EIGEN_SORT_VAL_DESC: better = ej > ek; break;
			case GSL_EIGEN_SORT_ABS_ASC: better = fabs(ej) < fabs(ek); break;
			default: better = fabs(ej) > fabs(ek); break;
			}
			if (better) { k = j; ek = ej; }
		}
		if (k != i) {
			const double t = gsl_vector_get(eval, i);
			gsl_vector_set(eval, i, gsl_vector_get(eval, k));
			gsl_vector_set(eval, k, t);
			for (size_t r = 0; r < evec->size1; r++) {
				const double u = gsl_matrix_get(evec, r, i);
				gsl_matrix_set(evec, r, i, gsl_matrix_get(evec, r, k));
				gsl_matrix_set(evec, r, k, u);
			}
		}
	}
	return GSL_SUCCESS;
}

/* ---- multimin: adaptor onto emub_bfgs ------------------------------------------------------------------ */
typedef struct {
	emub_bfgs *opt;
	gsl_multimin_function_fdf *fdf;
	size_t n;
} mm_state_t;

static double mm_f(const double *x, void *ctx)
{
	mm_state_t *st = (mm_state_t *)ctx;
	gsl_vector_view xv = gsl_vector_view_array((double *)x, st->n);
	return st->fdf->f(&xv.vector, st->fdf->params);
}
static void mm_df(const double *x, void *ctx, double *g)
{
	mm_state_t *st = (mm_state_t *)ctx;
	gsl_vector_view xv = gsl_vector_view_array((double *)x, st->n);
	gsl_vector_view gv = gsl_vector_view_array(g, st->n);
	st->fdf->df(&xv.vector, st->fdf->params, &gv.vector);
}
static void mm_fdf(const double *x, void *ctx, double *f, double *g)
{
	mm_state_t *st = (mm_state_t *)ctx;
	gsl_vector_view xv = gsl_vector_view_array((double *)x, st->n);
	gsl_vector_view gv = gsl_vector_view_array(g, st->n);
	st->fdf->fdf(&xv.vector, st->fdf->params, f, &gv.vector);
}

static const gsl_multimin_fdfminimizer_type bfgs2_type = { "vector_bfgs2", sizeof(mm_state_t), NULL, NULL, NULL, NULL, NULL };
const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_vector_bfgs2 = &bfgs2_type;
const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_vector_bfgs = &bfgs2_type;
const gsl_multimin_fdfminimizer_type *gsl_multimin_fdfminimizer_conjugate_fr = &bfgs2_type;

gsl_multimin_fdfminimizer *gsl_multimin_fdfminimizer_alloc(const gsl_multimin_fdfminimizer_type *T, size_t n)
{
	gsl_multimin_fdfminimizer *s = (gsl_multimin_fdfminimizer *)calloc(1, sizeof(*s));
	mm_state_t *st = (mm_state_t *)calloc(1, sizeof(*st));
	s->type = T;
	s->x = gsl_vector_calloc(n);
	s->gradient = gsl_vector_calloc(n);
	s->dx = gsl_vector_calloc(n);
	st->opt = emub_bfgs_alloc(n);
	st->n = n;
	s->state = st;
	return s;
}
static void mm_sync(gsl_multimin_fdfminimizer *s)
{
	mm_state_t *st = (mm_state_t *)s->state;
	memcpy(s->x->data, emub_bfgs_x(st->opt), st->n * sizeof(double));
	memcpy(s->gradient->data, emub_bfgs_gradient(st->opt), st->n * sizeof(double));
	memcpy(s->dx->data, emub_bfgs_dx(st->opt), st->n * sizeof(double));
	s->f = emub_bfgs_minimum(st->opt);
}
int gsl_multimin_fdfminimizer_set(gsl_multimin_fdfminimizer *s, gsl_multimin_function_fdf *fdf,
                                  const gsl_vector *x, double step_size, double tol)
{
	mm_state_t *st = (mm_state_t *)s->state;
	emub_bfgs_fn fn;
	double *x0 = (double *)malloc(st->n * sizeof(double));
	if (x->size != st->n || fdf->n != st->n) { free(x0); return GSL_EBADLEN; }
	for (size_t i = 0; i < st->n; i++) x0[i] = gsl_vector_get(x, i);
	s->fdf = fdf;
	st->fdf = fdf;
	fn.n = st->n; fn.f = mm_f; fn.df = mm_df; fn.fdf = mm_fdf; fn.ctx = st;
	emub_bfgs_set(st->opt, &fn, x0, step_size, tol);
	free(x0);
	mm_sync(s);
	return GSL_SUCCESS;
}
int gsl_multimin_fdfminimizer_iterate(gsl_multimin_fdfminimizer *s)
{
	mm_state_t *st = (mm_state_t *)s->state;
	int rc = emub_bfgs_iterate(st->opt);
	mm_sync(s);
	if (rc == EMUB_BFGS_OK) return GSL_SUCCESS;
	if (rc == EMUB_BFGS_ENOPROG) return GSL_ENOPROG;
	return GSL_FAILURE;
}
int gsl_multimin_fdfminimizer_restart(gsl_multimin_fdfminimizer *s) { (void)s; return GSL_SUCCESS; }
void gsl_multimin_fdfminimizer_free(gsl_multimin_fdfminimizer *s)
{
	if (!s) return;
	mm_state_t *st = (mm_state_t *)s->state;
	emub_bfgs_free(st->opt);
	free(st);
	gsl_vector_free(s->x);
	gsl_vector_free(s->gradient);
	gsl_vector_free(s->dx);
	free(s);
}
const char *gsl_multimin_fdfminimizer_name(const gsl_multimin_fdfminimizer *s) { return s->type->name; }
gsl_vector *gsl_multimin_fdfminimizer_x(const gsl_multimin_fdfminimizer *s) { return s->x; }
gsl_vector *gsl_multimin_fdfminimizer_dx(const gsl_multimin_fdfminimizer *s) { return s->dx; }
gsl_vector *gsl_multimin_fdfminimizer_gradient(const gsl_multimin_fdfminimizer *s) { return s->gradient; }
double gsl_multimin_fdfminimizer_minimum(const gsl_multimin_fdfminimizer *s) { return s->f; }
int gsl_multimin_test_gradient(const gsl_vector *g, double epsabs)
{
	if (epsabs < 0.0) return GSL_EBADTOL;
	return (gsl_blas_dnrm2(g) < epsabs) ? GSL_SUCCESS : GSL_CONTINUE;
}
