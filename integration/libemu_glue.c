/*
 * libemu_glue.c -- the reference-side binding of the B200 engine.
 *
 * This file is meant to be compiled INSIDE the reference tree (it includes the reference's own
 * headers, nothing of the reference is copied here) and defines the hot-path libEmu symbols with the
 * reference's exact signatures; every body only marshals gsl objects to (double*, ld) and calls the
 * C-ABI of include/emu_b200.h / madaiemulator_b200/host/emub_estimate.h.
 *
 *   evalFnMulti, gradFnMulti, evalFnGradMulti, estimateSigmaFull      src/libEmu/maxmultimin.c:288,416,615,148
 *   estimate_thetas_threaded                                         src/libEmu/estimate_threaded.c:78
 *   alloc_emulator_struct, free_emulator_struct, emulate_point       src/emulator_struct.c:13,43,124
 *   emulateAtPointList, emulateAtPoint                               src/libEmu/emulate-fns.c:73,138
 *   makeCovMatrix_fnptr                                              src/libEmu/emulator.c:636
 *
 * Engine handles are kept in side tables keyed by the reference's struct pointers, so no reference
 * struct changes.  Environment: EMUB_DEVICE (default 0), EMUB_SLOTS (in-flight evaluations per model,
 * default 8), EMUB_TRIES (restarts of estimate_thetas_threaded, default 50 x ncpus like the reference),
 * EMUB_POLISH (iterations of the refinement run from the best restart, default 100, 0 = the reference's stop rule only).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "modelstruct.h"
#include "optstruct.h"
#include "emulator_struct.h"
#include "libEmu/emulator.h"
#include "libEmu/maxmultimin.h"
#include "libEmu/estimate_threaded.h"
#include "libEmu/emulate-fns.h"

#include "emu_b200.h"
#include "emub_estimate.h"

#define GLUE_MAX 256

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
/* the reference's emulate_point is re-entrant (emulator_struct.c:124-143 touches only its arguments); an engine
 * context serves one call at a time, so concurrent callers take turns here */
static pthread_mutex_t g_call_mu = PTHREAD_MUTEX_INITIALIZER;
static emub_ctx *g_ctx = NULL;
static struct { const void *key_x, *key_y; int n, d, kernel, order; unsigned long long hash; emub_model *m; } g_models[GLUE_MAX];
static int g_nmodels = 0;
static struct { const emulator_struct *key; emub_emulator *e; emub_model *m; } g_emus[GLUE_MAX];
static int g_nemus = 0;

static int env_int(const char *name, int dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? atoi(s) : dflt;
}

static void glue_die(const char *what)
{
	fprintf(stderr, "libemu_glue: %s: %s\n", what, emub_last_error());
	exit(EXIT_FAILURE);
}

static emub_ctx *glue_ctx(void)
{
	if (!g_ctx && emub_ctx_create(env_int("EMUB_DEVICE", 0), &g_ctx) != EMUB_OK) glue_die("emub_ctx_create");
	return g_ctx;
}

/* FNV-1a over the design and the training vector: a modelstruct is recognised by its buffers AND their contents, so a
 * caller that frees a model and gets the same addresses back for another one (the R entry points build a model per
 * call) never meets a stale engine copy */
static unsigned long long glue_hash(modelstruct *ms)
{
	unsigned long long h = 1469598103934665603ull;
	const int n = ms->options->nmodel_points, d = ms->options->nparams;
	for (int i = 0; i < n; i++) {
		for (int k = 0; k <= d; k++) {
			const double v = k < d ? gsl_matrix_get(ms->xmodel, i, k) : gsl_vector_get(ms->training_vector, i);
			unsigned long long bits;
			memcpy(&bits, &v, sizeof(bits));
			h = (h ^ bits) * 1099511628211ull;
		}
	}
	return h;
}

/* the engine model that belongs to a reference modelstruct (created on first use) */
static emub_model *glue_model_for(modelstruct *ms)
{
	optstruct *o = ms->options;
	const unsigned long long hash = glue_hash(ms);
	pthread_mutex_lock(&g_mu);
	for (int i = 0; i < g_nmodels; i++)
		if (g_models[i].key_x == ms->xmodel->data && g_models[i].key_y == ms->training_vector->data &&
		    g_models[i].n == o->nmodel_points && g_models[i].d == o->nparams && g_models[i].kernel == o->cov_fn_index &&
		    g_models[i].order == o->regression_order) {
			if (g_models[i].hash == hash) {
				emub_model *m = g_models[i].m;
				pthread_mutex_unlock(&g_mu);
				return m;
			}
			/* same buffers, other contents: the old engine copy (and the emulators cached from it) is dead */
			for (int k = 0; k < g_nemus;)
				if (g_emus[k].m == g_models[i].m) { emub_emulator_destroy(g_emus[k].e); g_emus[k] = g_emus[--g_nemus]; }
				else k++;
			emub_model_destroy(g_models[i].m);
			g_models[i] = g_models[--g_nmodels];
			break;
		}
	if (g_nmodels == GLUE_MAX) { pthread_mutex_unlock(&g_mu); fprintf(stderr, "libemu_glue: too many models\n"); exit(EXIT_FAILURE); }
	/* training_vector may be a strided view: gather it */
	double *y = (double *)malloc(sizeof(double) * (size_t)o->nmodel_points);
	for (int i = 0; i < o->nmodel_points; i++) y[i] = gsl_vector_get(ms->training_vector, i);
	emub_model *m = NULL;
	if (emub_model_create(glue_ctx(), ms->xmodel->data, (int)ms->xmodel->tda, o->nmodel_points, o->nparams, y,
	                      o->cov_fn_index, o->regression_order, env_int("EMUB_SLOTS", 8), &m) != EMUB_OK)
		glue_die("emub_model_create");
	free(y);
	g_models[g_nmodels].key_x = ms->xmodel->data; g_models[g_nmodels].key_y = ms->training_vector->data;
	g_models[g_nmodels].n = o->nmodel_points; g_models[g_nmodels].d = o->nparams;
	g_models[g_nmodels].kernel = o->cov_fn_index; g_models[g_nmodels].order = o->regression_order;
	g_models[g_nmodels].m = m;
	g_models[g_nmodels].hash = hash;
	g_nmodels++;
	pthread_mutex_unlock(&g_mu);
	return m;
}

/* release every engine object (tests call this between cases) */
void libemu_glue_reset(void)
{
	pthread_mutex_lock(&g_mu);
	for (int i = 0; i < g_nemus; i++) emub_emulator_destroy(g_emus[i].e);
	for (int i = 0; i < g_nmodels; i++) emub_model_destroy(g_models[i].m);
	g_nemus = g_nmodels = 0;
	if (g_ctx) { emub_ctx_destroy(g_ctx); g_ctx = NULL; }
	pthread_mutex_unlock(&g_mu);
}

static int glue_eval(const gsl_vector *theta_vec, void *params_in, int want_grad, double *negL, double *grad, double *sigma2)
{
	struct estimate_thetas_params *p = (struct estimate_thetas_params *)params_in;
	emub_model *m = glue_model_for(p->the_model);
	double th[64];
	int status = 0;
	for (size_t i = 0; i < theta_vec->size; i++) th[i] = gsl_vector_get(theta_vec, i);
	if (emub_loglik_grad_batch(m, th, 1, want_grad, negL, grad, sigma2, &status) != EMUB_OK) glue_die("emub_loglik_grad_batch");
	return status;
}

/* maxmultimin.c:288 -- returns -L; NaN when the covariance matrix is not positive definite (:327-350) */
double evalFnMulti(const gsl_vector *theta_vec, void *params_in)
{
	double negL = 0.0;
	int status = glue_eval(theta_vec, params_in, 0, &negL, NULL, NULL);
	return status == EMUB_OK ? negL : GSL_NAN;
}

/* maxmultimin.c:416 -- the reference exits when the factorisation fails here (:495) */
void gradFnMulti(const gsl_vector *theta_vec, void *params_in, gsl_vector *grad_vec)
{
	double negL, g[64];
	int status = glue_eval(theta_vec, params_in, 1, &negL, g, NULL);
	if (status != EMUB_OK) {
		fprintf(stderr, "gradFnMulti: covariance matrix not positive definite\n");
		exit(EXIT_FAILURE);
	}
	for (size_t i = 0; i < grad_vec->size; i++) gsl_vector_set(grad_vec, i, g[i]);
}

/* maxmultimin.c:615 -- one evaluation instead of the reference's two factorisations */
void evalFnGradMulti(const gsl_vector *theta_vec, void *params_in, double *fnval, gsl_vector *grad_vec)
{
	double negL, g[64];
	int status = glue_eval(theta_vec, params_in, 1, &negL, g, NULL);
	*fnval = status == EMUB_OK ? negL : GSL_NAN;
	for (size_t i = 0; i < grad_vec->size; i++) gsl_vector_set(grad_vec, i, status == EMUB_OK ? g[i] : GSL_NAN);
}

/* maxmultimin.c:148 -- thetas is the vector without the amplitude */
double estimateSigmaFull(gsl_vector *thetas, void *params_in)
{
	double negL, sigma2 = GSL_NAN;
	int status = glue_eval(thetas, params_in, 0, &negL, NULL, &sigma2);
	return status == EMUB_OK ? sigma2 : GSL_NAN;
}

/* estimate_threaded.c:78 -- the restart machinery over the batched evaluator */
void estimate_thetas_threaded(modelstruct *the_model, optstruct *options)
{
	emub_model *m = glue_model_for(the_model);
	const int nth = options->nthetas;
	long ncpus = sysconf(_SC_NPROCESSORS_ONLN);
	emub_estimate_opts o;
	emub_estimate_default_opts(&o);
	o.max_tries = env_int("EMUB_TRIES", 50 * (int)(ncpus > 0 ? ncpus : 1)); /* estimate_threaded.c:97-113 */
	o.nchains = env_int("EMUB_SLOTS", 8);
	o.polish_steps = env_int("EMUB_POLISH", 100); /* refinement run from the best restart (emub_estimate.h); 0 = off */
	FILE *ur = fopen("/dev/urandom", "rb"); /* useful.c:49 */
	if (ur) { if (fread(&o.seed, sizeof(o.seed), 1, ur) != 1) o.seed = 1; fclose(ur); }
	if (getenv("EMUB_SEED")) o.seed = strtoull(getenv("EMUB_SEED"), NULL, 10);
	double *ranges = (double *)malloc(sizeof(double) * 2 * (size_t)nth);
	for (int i = 0; i < nth; i++) {
		ranges[2 * i] = gsl_matrix_get(options->grad_ranges, i, 0);
		ranges[2 * i + 1] = gsl_matrix_get(options->grad_ranges, i, 1);
	}
	double *th = (double *)calloc((size_t)nth, sizeof(double));
	double best = 0.0;
	emub_estimate_stats st;
	int rc = emub_estimate_thetas(m, ranges, &o, th, &best, &st);
	if (rc != EMUB_OK && rc != EMUB_EDOM) glue_die("emub_estimate_thetas");
	if (rc == EMUB_EDOM) fprintf(stderr, "maximisation didn't work at all, relax your ranges\n"); /* maxmultimin.c:122 */
	for (int i = 0; i < nth; i++) gsl_vector_set(the_model->thetas, i, th[i]);
	free(ranges);
	free(th);
}

/* emulator_struct.c:13 */
emulator_struct *alloc_emulator_struct(modelstruct *model)
{
	emulator_struct *e = (emulator_struct *)malloc(sizeof(emulator_struct));
	e->nparams = model->options->nparams;
	e->nmodel_points = model->options->nmodel_points;
	e->nregression_fns = model->options->nregression_fns;
	e->nthetas = model->options->nthetas;
	e->model = model;
	e->cinverse = NULL; /* lives on the device */
	e->h_matrix = NULL;
	e->beta_vector = gsl_vector_alloc(e->nregression_fns);
	emub_model *m = glue_model_for(model);
	double th[64], beta[64];
	for (int i = 0; i < e->nthetas; i++) th[i] = gsl_vector_get(model->thetas, i);
	emub_emulator *eh = NULL;
	int rc = emub_emulator_create(m, th, &eh);
	if (rc != EMUB_OK) { /* emulate-fns.c:282-285 exits on a failed factorisation */
		fprintf(stderr, "alloc_emulator_struct: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	emub_emulator_beta(eh, beta);
	for (int i = 0; i < e->nregression_fns; i++) gsl_vector_set(e->beta_vector, i, beta[i]);
	pthread_mutex_lock(&g_mu);
	if (g_nemus == GLUE_MAX) { fprintf(stderr, "libemu_glue: too many emulators\n"); exit(EXIT_FAILURE); }
	g_emus[g_nemus].key = e; g_emus[g_nemus].e = eh; g_emus[g_nemus].m = m; g_nemus++;
	pthread_mutex_unlock(&g_mu);
	return e;
}

static emub_emulator *glue_emulator_for(const emulator_struct *e, int remove)
{
	emub_emulator *r = NULL;
	pthread_mutex_lock(&g_mu);
	for (int i = 0; i < g_nemus; i++)
		if (g_emus[i].key == e) {
			r = g_emus[i].e;
			if (remove) { g_emus[i] = g_emus[g_nemus - 1]; g_nemus--; }
			break;
		}
	pthread_mutex_unlock(&g_mu);
	return r;
}

/* emulator_struct.c:43 */
void free_emulator_struct(emulator_struct *e)
{
	emub_emulator *eh = glue_emulator_for(e, 1);
	if (eh) emub_emulator_destroy(eh);
	gsl_vector_free(e->beta_vector);
	free((void *)e);
}

/* emulator_struct.c:124 */
void emulate_point(emulator_struct *e, gsl_vector *point, double *mean, double *variance)
{
	emub_emulator *eh = glue_emulator_for(e, 0);
	double x[64];
	for (int i = 0; i < e->nparams; i++) x[i] = gsl_vector_get(point, i);
	pthread_mutex_lock(&g_call_mu);
	const int rc = eh ? emub_predict_few(eh, x, e->nparams, 1, mean, variance) : EMUB_EINVAL;
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_predict_few");
}

/* emulate-fns.c:73 -- covariance, factorisation and regression once, then every point of the list: one emulator, one
 * batched prediction (options->nemulate_points rows of point_list) */
void emulateAtPointList(modelstruct *the_model, gsl_matrix *point_list, optstruct *options, double *the_mean, double *the_variance)
{
	emub_model *m = glue_model_for(the_model);
	double th[64];
	for (int i = 0; i < options->nthetas; i++) th[i] = gsl_vector_get(the_model->thetas, i);
	emub_emulator *eh = NULL;
	if (emub_emulator_create(m, th, &eh) != EMUB_OK) { /* chol_inverse_cov_matrix exits on a failed factorisation, :282-285 */
		fprintf(stderr, "emulateAtPointList: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	if (emub_predict_batch(eh, point_list->data, (int)point_list->tda, options->nemulate_points, the_mean, the_variance) != EMUB_OK)
		glue_die("emub_predict_batch");
	emub_emulator_destroy(eh);
}

/* emulate-fns.c:138 */
void emulateAtPoint(modelstruct *the_model, gsl_vector *the_point, optstruct *options, double *the_mean, double *the_variance)
{
	emub_model *m = glue_model_for(the_model);
	double th[64], x[64];
	for (int i = 0; i < options->nthetas; i++) th[i] = gsl_vector_get(the_model->thetas, i);
	for (int i = 0; i < options->nparams; i++) x[i] = gsl_vector_get(the_point, i);
	emub_emulator *eh = NULL;
	if (emub_emulator_create(m, th, &eh) != EMUB_OK) {
		fprintf(stderr, "emulateAtPoint: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	if (emub_predict_few(eh, x, options->nparams, 1, the_mean, the_variance) != EMUB_OK) glue_die("emub_predict_few");
	emub_emulator_destroy(eh);
}

/* emulator.c:636 -- the kernel is identified by the function pointer the caller passes */
void makeCovMatrix_fnptr(gsl_matrix *cov_matrix, gsl_matrix *xmodel, gsl_vector *thetas, int nmodel_points, int nthetas,
                         int nparams, double (*covariance_fn_ptr)(gsl_vector *, gsl_vector *, gsl_vector *, int, int))
{
	int kernel = EMUB_POWEREXP;
	if (covariance_fn_ptr == covariance_fn_matern_three) kernel = EMUB_MATERN32;
	else if (covariance_fn_ptr == covariance_fn_matern_five) kernel = EMUB_MATERN52;
	double th[64];
	for (int i = 0; i < nthetas; i++) th[i] = gsl_vector_get(thetas, i);
	double *zeros = (double *)calloc((size_t)nmodel_points, sizeof(double));
	emub_model *m = NULL;
	if (emub_model_create(glue_ctx(), xmodel->data, (int)xmodel->tda, nmodel_points, nparams, zeros, kernel, 0, 1, &m) != EMUB_OK)
		glue_die("emub_model_create");
	if (emub_cov_matrix(m, th, cov_matrix->data, (int)cov_matrix->tda) != EMUB_OK) glue_die("emub_cov_matrix");
	emub_model_destroy(m);
	free(zeros);
}
