/*
 * multivar_glue.c -- the multivariate front doors of the reference (src/multivar_support.c), same names and
 * signatures, over the BATCHED entry points of the B200 engine.  Compiled with the reference's own
 * headers, next to libemu_glue.c; the reference's definitions are weakened or deleted.
 *
 *   estimate_multi            multivar_support.c:20   serial loop over the PCA components
 *                             -> emub_estimate_thetas_multi_devices: the restart chains of ALL components feed one
 *                                evaluation front (and the components are shared out over EMUB_DEVICES)
 *   alloc_multi_emulator      multivar_support.c:30   one emulator_struct (n x n inverse on the host) per component
 *                             -> one engine model carrying the nr training vectors + one cached factor per component,
 *                                replicated on every device of EMUB_DEVICES
 *   emulate_point_multi[_pca] multivar_support.c:78,103   nr emulate_point calls + back-projection on the host
 *                             -> one emub_predict_multi_few call (the latency path for single points): all components
 *                                and the back-projection on the device
 *   free_multi_emulator       multivar_support.c:57
 *
 * Environment: EMUB_DEVICES ("0,1,..", default EMUB_DEVICE or 0), EMUB_TRIES (restarts per component, default
 * 50 x ncpus as estimate_threaded.c:97-113), EMUB_SLOTS (chains in flight per component, default 8), EMUB_SEED,
 * EMUB_POLISH (iterations of the refinement run from each component's best restart, default 100, 0 = off).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "multivar_support.h"
#include "multi_modelstruct.h"
#include "modelstruct.h"
#include "optstruct.h"
#include "libEmu/estimate_threaded.h"

#include "emu_b200.h"
#include "emub_estimate.h"
#include "emub_interactive.h"

#define MV_MAX 64

static pthread_mutex_t mv_mu = PTHREAD_MUTEX_INITIALIZER;
static struct { const multi_emulator *key; emub_multi_emulator *me; } mv_emus[MV_MAX];
static int mv_nemus = 0;

static void mv_die(const char *what)
{
	fprintf(stderr, "multivar_glue: %s: %s\n", what, emub_last_error());
	exit(EXIT_FAILURE);
}

static int mv_env_int(const char *name, int dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? atoi(s) : dflt;
}

/* EMUB_DEVICES = "0,1,.." (else EMUB_DEVICE, else 0) */
static int mv_devices(int *devs, int max)
{
	const char *s = getenv("EMUB_DEVICES");
	int n = 0;
	if (s && *s) {
		char *copy = strdup(s);
		for (char *tok = strtok(copy, ","); tok && n < max; tok = strtok(NULL, ",")) devs[n++] = atoi(tok);
		free(copy);
	}
	if (n == 0) devs[n++] = mv_env_int("EMUB_DEVICE", 0);
	return n;
}

static double *mv_gather_matrix(const gsl_matrix *a)
{
	double *out = (double *)malloc(sizeof(double) * a->size1 * a->size2);
	for (size_t i = 0; i < a->size1; i++)
		for (size_t j = 0; j < a->size2; j++) out[i * a->size2 + j] = gsl_matrix_get(a, i, j);
	return out;
}

static double *mv_gather_vector(const gsl_vector *v)
{
	double *out = (double *)malloc(sizeof(double) * v->size);
	for (size_t i = 0; i < v->size; i++) out[i] = gsl_vector_get(v, i);
	return out;
}

/* multivar_support.c:20 */
void estimate_multi(multi_modelstruct *m, FILE *outfp)
{
	const int n = m->nmodel_points, d = m->nparams, nr = m->nr;
	optstruct *o0 = m->pca_model_array[0]->options;
	const int nth = o0->nthetas;
	/* one front needs one set of optimisation ranges; they come from the design and the options, which the components
	 * share (gen_pca_model_array, multi_modelstruct.c:121-148) -- if a caller made them differ, keep the serial loop */
	int same = 1;
	for (int c = 1; c < nr && same; c++) {
		optstruct *oc = m->pca_model_array[c]->options;
		if (oc->nthetas != nth || oc->cov_fn_index != o0->cov_fn_index || oc->regression_order != o0->regression_order) same = 0;
		for (int i = 0; i < nth && same; i++)
			if (gsl_matrix_get(oc->grad_ranges, i, 0) != gsl_matrix_get(o0->grad_ranges, i, 0) ||
			    gsl_matrix_get(oc->grad_ranges, i, 1) != gsl_matrix_get(o0->grad_ranges, i, 1)) same = 0;
	}
	if (!same) {
		for (int c = 0; c < nr; c++) estimate_thetas_threaded(m->pca_model_array[c], m->pca_model_array[c]->options);
		dump_multi_modelstruct(outfp, m);
		return;
	}
	double *X = mv_gather_matrix(m->xmodel);
	double *Z = (double *)malloc(sizeof(double) * (size_t)n * nr);
	for (int c = 0; c < nr; c++)
		for (int i = 0; i < n; i++) Z[(size_t)i * nr + c] = gsl_vector_get(m->pca_model_array[c]->training_vector, i);
	double *ranges = (double *)malloc(sizeof(double) * 2 * (size_t)nth);
	for (int i = 0; i < nth; i++) {
		ranges[2 * i] = gsl_matrix_get(o0->grad_ranges, i, 0);
		ranges[2 * i + 1] = gsl_matrix_get(o0->grad_ranges, i, 1);
	}
	emub_estimate_opts o;
	emub_estimate_default_opts(&o);
	long ncpus = sysconf(_SC_NPROCESSORS_ONLN);
	o.max_tries = mv_env_int("EMUB_TRIES", 50 * (int)(ncpus > 0 ? ncpus : 1)); /* estimate_threaded.c:97-113 */
	o.nchains = mv_env_int("EMUB_SLOTS", 8);
	o.polish_steps = mv_env_int("EMUB_POLISH", 100); /* refinement run from the best restart (emub_estimate.h); 0 = off */
	FILE *ur = fopen("/dev/urandom", "rb"); /* useful.c:49 */
	if (ur) { if (fread(&o.seed, sizeof(o.seed), 1, ur) != 1) o.seed = 1; fclose(ur); }
	if (getenv("EMUB_SEED")) o.seed = strtoull(getenv("EMUB_SEED"), NULL, 10);
	int devs[64];
	const int ndev = mv_devices(devs, 64);
	double *th = (double *)calloc((size_t)nr * nth, sizeof(double));
	double *best = (double *)calloc((size_t)nr, sizeof(double));
	emub_estimate_stats st;
	int rc = emub_estimate_thetas_multi_devices_ranges(devs, ndev, X, d, n, d, Z, nr, nr, o0->cov_fn_index, o0->regression_order, 0,
	                                                   ranges, &o, th, best, &st);
	if (rc != EMUB_OK && rc != EMUB_EDOM) mv_die("emub_estimate_thetas_multi_devices");
	if (rc == EMUB_EDOM) fprintf(stderr, "maximisation didn't work at all, relax your ranges\n"); /* maxmultimin.c:122 */
	for (int c = 0; c < nr; c++)
		for (int i = 0; i < nth; i++) gsl_vector_set(m->pca_model_array[c]->thetas, i, th[(size_t)c * nth + i]);
	free(X); free(Z); free(ranges); free(th); free(best);
	/* dump the trained modelstruct, as the reference does (:26) */
	dump_multi_modelstruct(outfp, m);
}

/* multivar_support.c:30 */
multi_emulator *alloc_multi_emulator(multi_modelstruct *m)
{
	multi_emulator *e = (multi_emulator *)malloc(sizeof(multi_emulator));
	e->nt = m->nt;
	e->nr = m->nr;
	e->nparams = m->nparams;
	e->nmodel_points = m->nmodel_points;
	e->nregression_fns = m->pca_model_array[0]->options->nregression_fns;
	e->nthetas = m->pca_model_array[0]->options->nthetas;
	e->model = m;
	/* the per-component host structs (n x n inverse each) are not built: the factors live on the device */
	e->emu_struct_array = (emulator_struct **)calloc((size_t)e->nr, sizeof(emulator_struct *));

	/* the engine's view of the model: the same arrays a snapshot carries */
	emub_snapshot s;
	memset(&s, 0, sizeof(s));
	s.nt = m->nt; s.nr = m->nr; s.nparams = m->nparams; s.nmodel_points = m->nmodel_points;
	s.cov_fn_index = m->cov_fn_index; s.regression_order = m->regression_order;
	s.training_mean = mv_gather_vector(m->training_mean);
	s.pca_evals_r = mv_gather_vector(m->pca_evals_r);
	s.pca_evecs_r = mv_gather_matrix(m->pca_evecs_r);
	s.components = (emub_snapshot_component *)calloc((size_t)m->nr, sizeof(emub_snapshot_component));
	for (int c = 0; c < m->nr; c++) {
		modelstruct *ms = m->pca_model_array[c];
		emub_snapshot_component *sc = &s.components[c];
		sc->nthetas = ms->options->nthetas; sc->nparams = m->nparams; sc->nmodel_points = m->nmodel_points;
		sc->regression_order = ms->options->regression_order; sc->nregression_fns = ms->options->nregression_fns;
		sc->cov_fn_index = ms->options->cov_fn_index;
		sc->xmodel = mv_gather_matrix(ms->xmodel);
		sc->training_vector = mv_gather_vector(ms->training_vector);
		sc->thetas = mv_gather_vector(ms->thetas);
	}
	int devs[64];
	const int ndev = mv_devices(devs, 64);
	emub_multi_emulator *me = NULL;
	/* a failed factorisation ends the process in the reference (emulate-fns.c:282-285) */
	if (emub_multi_emulator_from_snapshot_devices(devs, ndev, &s, &me) != EMUB_OK) mv_die("alloc_multi_emulator");
	for (int c = 0; c < m->nr; c++) { free(s.components[c].xmodel); free(s.components[c].training_vector); free(s.components[c].thetas); }
	free(s.components); free(s.training_mean); free(s.pca_evals_r); free(s.pca_evecs_r);
	pthread_mutex_lock(&mv_mu);
	if (mv_nemus == MV_MAX) { pthread_mutex_unlock(&mv_mu); fprintf(stderr, "multivar_glue: too many multi_emulators\n"); exit(EXIT_FAILURE); }
	mv_emus[mv_nemus].key = e; mv_emus[mv_nemus].me = me; mv_nemus++;
	pthread_mutex_unlock(&mv_mu);
	return e;
}

static emub_multi_emulator *mv_lookup(const multi_emulator *e)
{
	emub_multi_emulator *me = NULL;
	pthread_mutex_lock(&mv_mu);
	for (int i = 0; i < mv_nemus; i++)
		if (mv_emus[i].key == e) me = mv_emus[i].me;
	pthread_mutex_unlock(&mv_mu);
	if (!me) { fprintf(stderr, "multivar_glue: unknown multi_emulator\n"); exit(EXIT_FAILURE); }
	return me;
}

/* multivar_support.c:57 */
void free_multi_emulator(multi_emulator *e)
{
	pthread_mutex_lock(&mv_mu);
	for (int i = 0; i < mv_nemus; i++)
		if (mv_emus[i].key == e) {
			emub_multi_emulator_destroy(mv_emus[i].me);
			mv_emus[i] = mv_emus[--mv_nemus];
			break;
		}
	pthread_mutex_unlock(&mv_mu);
	free_multimodelstruct(e->model);
	free(e->emu_struct_array);
}

static void mv_emulate(multi_emulator *emu, gsl_vector *the_point, gsl_vector *the_mean, gsl_vector *the_variance, int pca)
{
	emub_multi_emulator *me = mv_lookup(emu);
	double pt[64], mean[4096], var[4096];
	const int d = emu->nparams, nt = emu->nt, nout = pca ? emu->nr : emu->nt;
	if (d > 64 || nt > 4096) { fprintf(stderr, "multivar_glue: model too wide\n"); exit(EXIT_FAILURE); }
	for (int k = 0; k < d; k++) pt[k] = gsl_vector_get(the_point, k);
	/* the reference's emulate_point_multi is re-entrant; an engine context serves one call at a time */
	static pthread_mutex_t call_mu = PTHREAD_MUTEX_INITIALIZER;
	pthread_mutex_lock(&call_mu);
	const int rc = emub_multi_emulator_predict_few(me, pt, 1, pca, mean, var);
	pthread_mutex_unlock(&call_mu);
	if (rc != EMUB_OK) mv_die("emulate_point_multi");
	for (int i = 0; i < nout; i++) {
		gsl_vector_set(the_mean, i, mean[i]);
		gsl_vector_set(the_variance, i, var[i]);
	}
}

/* multivar_support.c:103: values in the observable space (nt) */
void emulate_point_multi(multi_emulator *emu, gsl_vector *the_point, gsl_vector *the_mean, gsl_vector *the_variance)
{
	mv_emulate(emu, the_point, the_mean, the_variance, 0);
}

/* multivar_support.c:78: values in the PCA space (nr) */
void emulate_point_multi_pca(multi_emulator *emu, gsl_vector *the_point, gsl_vector *the_mean, gsl_vector *the_variance)
{
	mv_emulate(emu, the_point, the_mean, the_variance, 1);
}
