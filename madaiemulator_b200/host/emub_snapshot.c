/* emub_snapshot.c -- see emub_snapshot.h */
#include "emub_snapshot.h"
#include <stdlib.h>
#include <string.h>

/* whole file -> token stream (the reference reads every value with fscanf("%lf%*c") / "%d%*c") */
typedef struct { char *buf; char *p; int bad; } tok_t;

static double next_double(tok_t *t)
{
	char *end;
	double v = strtod(t->p, &end);
	if (end == t->p) { t->bad = 1; return 0.0; }
	t->p = end;
	return v;
}
static int next_int(tok_t *t)
{
	char *end;
	long v = strtol(t->p, &end, 10);
	if (end == t->p) { t->bad = 1; return 0; }
	t->p = end;
	return (int)v;
}
static double *read_doubles(tok_t *t, size_t n)
{
	double *a = (double *)malloc(sizeof(double) * (n ? n : 1));
	for (size_t i = 0; i < n && !t->bad; i++) a[i] = next_double(t);
	return a;
}

static void seterr(char *err, int errlen, const char *msg)
{
	if (err && errlen > 0) { strncpy(err, msg, (size_t)errlen - 1); err[errlen - 1] = 0; }
}

emub_snapshot *emub_snapshot_load(FILE *f, char *err, int errlen)
{
	size_t cap = 1 << 20, len = 0;
	char *buf = (char *)malloc(cap + 1);
	for (;;) {
		if (len == cap) { cap *= 2; buf = (char *)realloc(buf, cap + 1); }
		size_t r = fread(buf + len, 1, cap - len, f);
		if (r == 0) break;
		len += r;
	}
	buf[len] = 0;
	tok_t t = {buf, buf, 0};
	emub_snapshot *s = (emub_snapshot *)calloc(1, sizeof(*s));
	s->nt = next_int(&t); s->nr = next_int(&t); s->nparams = next_int(&t); s->nmodel_points = next_int(&t);
	s->cov_fn_index = next_int(&t); s->regression_order = next_int(&t);
	if (t.bad || s->nt < 1 || s->nr < 1 || s->nr > s->nt || s->nparams < 1 || s->nmodel_points < 1 ||
	    s->nt > 4096 || s->nparams > 4096 || s->nmodel_points > (1 << 24)) {
		seterr(err, errlen, "snapshot header is malformed");
		free(buf); free(s);
		return NULL;
	}
	const size_t n = (size_t)s->nmodel_points, d = (size_t)s->nparams, nt = (size_t)s->nt, nr = (size_t)s->nr;
	s->xmodel = read_doubles(&t, n * d);
	s->training_matrix = read_doubles(&t, n * nt);
	s->pca_evals_r = read_doubles(&t, nr);
	s->pca_evecs_r = read_doubles(&t, nt * nr);
	s->pca_zmatrix = read_doubles(&t, n * nr);
	s->components = (emub_snapshot_component *)calloc(nr, sizeof(emub_snapshot_component));
	const char *inconsistent = NULL;
	for (size_t c = 0; c < nr && !t.bad; c++) {
		emub_snapshot_component *m = &s->components[c];
		m->nthetas = next_int(&t); m->nparams = next_int(&t); m->nmodel_points = next_int(&t);
		m->nemulate_points = next_int(&t); m->regression_order = next_int(&t); m->nregression_fns = next_int(&t);
		m->fixed_nugget_mode = next_int(&t); m->fixed_nugget = next_double(&t);
		m->cov_fn_index = next_int(&t); m->use_data_scales = next_int(&t);
		if (t.bad || m->nthetas < 1 || m->nthetas > 4096 || m->nparams != s->nparams || m->nmodel_points != s->nmodel_points) {
			t.bad = 1;
			break;
		}
		/* The engine copies nthetas(kernel) values out of `thetas` and configures ONE model from component 0, so the
		 * block must agree with both.  Kernel and order are read the way the reference's loader does
		 * (set_global_ptrs, modelstruct.c:214-258: anything but MATERN32/52 is the power-exponential kernel, anything
		 * but 1..3 the trivial regression). */
		{
			const int kern = (m->cov_fn_index == 2 || m->cov_fn_index == 3) ? m->cov_fn_index : 1;
			const int ord = (m->regression_order >= 1 && m->regression_order <= 3) ? m->regression_order : 0;
			const emub_snapshot_component *m0 = &s->components[0];
			const int kern0 = (m0->cov_fn_index == 2 || m0->cov_fn_index == 3) ? m0->cov_fn_index : 1;
			const int ord0 = (m0->regression_order >= 1 && m0->regression_order <= 3) ? m0->regression_order : 0;
			if (m->nthetas != (kern == 1 ? s->nparams + 2 : 3)) { /* modelstruct.c:301-308 */
				inconsistent = "snapshot component: nthetas does not match its covariance function";
				t.bad = 1;
				break;
			}
			if (kern != kern0 || ord != ord0) {
				inconsistent = "snapshot components disagree on covariance function or regression order";
				t.bad = 1;
				break;
			}
		}
		m->grad_ranges = read_doubles(&t, 2 * (size_t)m->nthetas);
		m->xmodel = read_doubles(&t, n * d);
		m->training_vector = read_doubles(&t, n);
		m->thetas = read_doubles(&t, (size_t)m->nthetas);
		m->sample_scales = read_doubles(&t, d);
	}
	free(buf);
	if (t.bad) {
		seterr(err, errlen, inconsistent ? inconsistent : "snapshot body is truncated or malformed");
		emub_snapshot_free(s);
		return NULL;
	}
	/* column means of the training matrix, summed in row order (vector_elt_sum, multi_modelstruct.c:464-486) */
	s->training_mean = (double *)malloc(sizeof(double) * nt);
	for (size_t j = 0; j < nt; j++) {
		double sum = 0.0;
		for (size_t i = 0; i < n; i++) sum += s->training_matrix[i * nt + j];
		s->training_mean[j] = sum / ((double)s->nmodel_points);
	}
	return s;
}

emub_snapshot *emub_snapshot_load_path(const char *path, char *err, int errlen)
{
	FILE *f = fopen(path, "r");
	if (!f) { seterr(err, errlen, "cannot open snapshot file"); return NULL; }
	emub_snapshot *s = emub_snapshot_load(f, err, errlen);
	fclose(f);
	return s;
}

/* "%.17lf " per value, one row per line: the exact bytes of dump_multi_modelstruct / dump_modelstruct_2 */
static void put_rows(FILE *f, const double *a, size_t rows, size_t cols)
{
	for (size_t i = 0; i < rows; i++) {
		for (size_t j = 0; j < cols; j++) fprintf(f, "%.17lf ", a[i * cols + j]);
		fprintf(f, "\n");
	}
}

int emub_snapshot_save(const emub_snapshot *s, FILE *f)
{
	if (!s || !f) return -1;
	const size_t n = (size_t)s->nmodel_points, d = (size_t)s->nparams, nt = (size_t)s->nt, nr = (size_t)s->nr;
	/* multi_modelstruct.c:358-401 */
	fprintf(f, "%d\n%d\n%d\n%d\n%d\n%d\n", s->nt, s->nr, s->nparams, s->nmodel_points, s->cov_fn_index, s->regression_order);
	put_rows(f, s->xmodel, n, d);
	put_rows(f, s->training_matrix, n, nt);
	put_rows(f, s->pca_evals_r, 1, nr);
	put_rows(f, s->pca_evecs_r, nt, nr);
	put_rows(f, s->pca_zmatrix, n, nr);
	for (size_t c = 0; c < nr; c++) { /* modelstruct.c:381-408 */
		const emub_snapshot_component *m = &s->components[c];
		fprintf(f, "%d\n%d\n%d\n%d\n%d\n%d\n%d\n%.17lf\n%d\n%d\n", m->nthetas, m->nparams, m->nmodel_points, m->nemulate_points,
		        m->regression_order, m->nregression_fns, m->fixed_nugget_mode, m->fixed_nugget, m->cov_fn_index, m->use_data_scales);
		for (int i = 0; i < m->nthetas; i++) fprintf(f, "%.17lf %.17lf\n", m->grad_ranges[2 * i], m->grad_ranges[2 * i + 1]);
		put_rows(f, m->xmodel, n, d);
		put_rows(f, m->training_vector, 1, n);
		put_rows(f, m->thetas, 1, (size_t)m->nthetas);
		put_rows(f, m->sample_scales, 1, d);
	}
	return ferror(f) ? -1 : 0;
}

int emub_snapshot_save_path(const emub_snapshot *s, const char *path)
{
	FILE *f = fopen(path, "w");
	if (!f) return -1;
	int rc = emub_snapshot_save(s, f);
	if (fclose(f) != 0) rc = -1;
	return rc;
}

/* A snapshot for nr scalar GPs that share one design, from plain arrays (what estimate_multi + dump leave behind):
 * Z is n x nr (component training vectors), thetas nr x nthetas.  With nt == nr, unit eigenvalues and identity
 * eigenvectors the back-projection is the identity, i.e. the outputs are the components themselves. */
emub_snapshot *emub_snapshot_from_arrays(const double *X, int n, int d, const double *Z, int nr, const double *thetas, int nthetas,
                                         int cov_fn_index, int regression_order)
{
	if (!X || !Z || !thetas || n < 1 || d < 1 || nr < 1 || nthetas < 1) return NULL;
	emub_snapshot *s = (emub_snapshot *)calloc(1, sizeof(*s));
	s->nt = nr; s->nr = nr; s->nparams = d; s->nmodel_points = n; s->cov_fn_index = cov_fn_index; s->regression_order = regression_order;
	const size_t N = (size_t)n, D = (size_t)d, R = (size_t)nr;
	s->xmodel = (double *)malloc(sizeof(double) * N * D);
	memcpy(s->xmodel, X, sizeof(double) * N * D);
	/* training matrix = Z with its column means restored as zero: Y = Z, so ybar = column means of Z */
	s->training_matrix = (double *)malloc(sizeof(double) * N * R);
	memcpy(s->training_matrix, Z, sizeof(double) * N * R);
	s->training_mean = (double *)calloc(R, sizeof(double));
	for (size_t j = 0; j < R; j++) {
		double sum = 0.0;
		for (size_t i = 0; i < N; i++) sum += Z[i * R + j];
		s->training_mean[j] = sum / (double)n;
	}
	s->pca_evals_r = (double *)malloc(sizeof(double) * R);
	s->pca_evecs_r = (double *)calloc(R * R, sizeof(double));
	s->pca_zmatrix = (double *)malloc(sizeof(double) * N * R);
	for (size_t j = 0; j < R; j++) { s->pca_evals_r[j] = 1.0; s->pca_evecs_r[j * R + j] = 1.0; }
	/* z = y - ybar so that ybar + 1 * z reproduces y */
	for (size_t i = 0; i < N; i++)
		for (size_t j = 0; j < R; j++) s->pca_zmatrix[i * R + j] = Z[i * R + j] - s->training_mean[j];
	s->components = (emub_snapshot_component *)calloc(R, sizeof(emub_snapshot_component));
	for (size_t c = 0; c < R; c++) {
		emub_snapshot_component *m = &s->components[c];
		m->nthetas = nthetas; m->nparams = d; m->nmodel_points = n; m->nemulate_points = 0;
		m->regression_order = regression_order; m->nregression_fns = 1 + regression_order * d;
		m->fixed_nugget_mode = 0; m->fixed_nugget = 0.0; m->cov_fn_index = cov_fn_index; m->use_data_scales = 1;
		m->grad_ranges = (double *)calloc(2 * (size_t)nthetas, sizeof(double));
		m->xmodel = (double *)malloc(sizeof(double) * N * D);
		memcpy(m->xmodel, X, sizeof(double) * N * D);
		m->training_vector = (double *)malloc(sizeof(double) * N);
		for (size_t i = 0; i < N; i++) m->training_vector[i] = s->pca_zmatrix[i * R + c];
		m->thetas = (double *)malloc(sizeof(double) * (size_t)nthetas);
		memcpy(m->thetas, thetas + c * (size_t)nthetas, sizeof(double) * (size_t)nthetas);
		m->sample_scales = (double *)calloc(D, sizeof(double));
	}
	return s;
}

void emub_snapshot_free(emub_snapshot *s)
{
	if (!s) return;
	if (s->components)
		for (int c = 0; c < s->nr; c++) {
			emub_snapshot_component *m = &s->components[c];
			free(m->grad_ranges); free(m->xmodel); free(m->training_vector); free(m->thetas); free(m->sample_scales);
		}
	free(s->components);
	free(s->xmodel); free(s->training_matrix); free(s->training_mean); free(s->pca_evals_r); free(s->pca_evecs_r);
	free(s->pca_zmatrix);
	free(s);
}
