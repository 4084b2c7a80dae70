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
		m->grad_ranges = read_doubles(&t, 2 * (size_t)m->nthetas);
		m->xmodel = read_doubles(&t, n * d);
		m->training_vector = read_doubles(&t, n);
		m->thetas = read_doubles(&t, (size_t)m->nthetas);
		m->sample_scales = read_doubles(&t, d);
	}
	free(buf);
	if (t.bad) {
		seterr(err, errlen, "snapshot body is truncated or malformed");
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
