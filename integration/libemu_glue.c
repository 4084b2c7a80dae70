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
 *   makeCovMatrix[_fnptr], makeKVector[_fnptr]                       src/libEmu/emulator.c:607,636,553,578
 *   emulateQuick, chol_inverse_cov_matrix                            src/libEmu/emulate-fns.c:201,275
 *   emulate_model_results, emulate_ith_location                      src/libEmu/emulate-fns.c:13,234
 *
 * Engine handles are kept in side tables keyed by the reference's struct pointers, so no reference
 * struct changes.  Environment: EMUB_DEVICE (default 0), EMUB_SLOTS (in-flight evaluations per model,
 * default 8), EMUB_GLUE_MODELS (engine models kept at once, least recently used dropped first, default 8), EMUB_TRIES (restarts of estimate_thetas_threaded, default 50 x ncpus like the reference),
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
#include "resultstruct.h"
#include "libEmu/emulator.h"
#include "libEmu/maxmultimin.h"
#include "libEmu/estimate_threaded.h"
#include "libEmu/emulate-fns.h"

#include "emu_b200.h"
#include "emub_estimate.h"

#define GLUE_MAX 256
#define GLUE_TH 128 /* thetas / coordinates marshalled on the stack */

/* g_call_mu serialises every use of the engine (one context, one set of streams and factorisation slots: it serves one
 * call at a time; the reference's emulate_point is re-entrant, emulator_struct.c:124-143, so concurrent callers take
 * turns here).  It is taken first, by every entry point; g_mu guards the side tables and is only taken inside it. */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_mutex_t g_call_mu = PTHREAD_MUTEX_INITIALIZER;
static emub_ctx *g_ctx = NULL;
/* engine models by content: key_x / key_y are the caller's buffers (key_y NULL: a design-only model, key_x NULL too: a
 * bare n x n factorisation workspace), hash their contents */
static struct {
	const void *key_x, *key_y;
	int n, d, kernel, order;
	unsigned long long hash, last_use;
	emub_model *m;
} g_models[GLUE_MAX];
static int g_nmodels = 0;
static unsigned long long g_clock = 0;
static struct { const emulator_struct *key; emub_emulator *e; emub_model *m; } g_emus[GLUE_MAX];
static int g_nemus = 0;
/* emulators made for emulateQuick: by model and theta contents */
#define GLUE_QUICK 32
static struct { emub_model *m; int nth; double th[GLUE_TH]; emub_emulator *e; unsigned long long last_use; } g_quick[GLUE_QUICK];
static int g_nquick = 0;

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

/* call with g_call_mu held */
static emub_ctx *glue_ctx(void)
{
	pthread_mutex_lock(&g_mu);
	if (!g_ctx && emub_ctx_create(env_int("EMUB_DEVICE", 0), &g_ctx) != EMUB_OK) glue_die("emub_ctx_create");
	emub_ctx *c = g_ctx;
	pthread_mutex_unlock(&g_mu);
	return c;
}

/* FNV-1a over a buffer's bit patterns, four interleaved lanes (the multiply chain is what costs: ~15 us for a
 * 4096 x 10 design).  A model is recognised by its buffers AND their contents, so a caller that frees a model and
 * gets the same addresses back for another one (the R entry points build a model per call) never meets a stale
 * engine copy. */
static unsigned long long hash_doubles(unsigned long long h, const double *a, size_t stride_rows, size_t rows, size_t cols)
{
	unsigned long long l[4] = {h, h ^ 0x9E3779B97F4A7C15ull, h ^ 0xBF58476D1CE4E5B9ull, h ^ 0x94D049BB133111EBull};
	size_t k = 0;
	for (size_t i = 0; i < rows; i++) {
		const double *row = a + i * stride_rows;
		for (size_t j = 0; j < cols; j++, k++) {
			unsigned long long bits;
			memcpy(&bits, &row[j], sizeof(bits));
			l[k & 3] = (l[k & 3] ^ bits) * 1099511628211ull;
		}
	}
	return ((l[0] * 31 + l[1]) * 31 + l[2]) * 31 + l[3];
}

static unsigned long long glue_hash_xy(const gsl_matrix *x, const gsl_vector *y, int n, int d)
{
	unsigned long long h = 1469598103934665603ull;
	if (x) h = hash_doubles(h, x->data, x->tda, (size_t)n, (size_t)d);
	if (y) h = hash_doubles(h, y->data, y->stride, (size_t)n, 1);
	return h;
}

/* drop entry i of the model table with everything cached from it (g_mu held) */
static void glue_drop_model(int i)
{
	emub_model *m = g_models[i].m;
	for (int k = 0; k < g_nemus;)
		if (g_emus[k].m == m) { emub_emulator_destroy(g_emus[k].e); g_emus[k] = g_emus[--g_nemus]; }
		else k++;
	for (int k = 0; k < g_nquick;)
		if (g_quick[k].m == m) { emub_emulator_destroy(g_quick[k].e); g_quick[k] = g_quick[--g_nquick]; }
		else k++;
	emub_model_destroy(m);
	g_models[i] = g_models[--g_nmodels];
}

/* least recently used model that no live emulator_struct depends on; -1 if there is none (g_mu held) */
static int glue_lru_victim(void)
{
	int best = -1;
	for (int i = 0; i < g_nmodels; i++) {
		int pinned = 0;
		for (int k = 0; k < g_nemus; k++)
			if (g_emus[k].m == g_models[i].m) { pinned = 1; break; }
		if (!pinned && (best < 0 || g_models[i].last_use < g_models[best].last_use)) best = i;
	}
	return best;
}

/* The engine model for (design, training vector, kernel, order), created on first use.  Every model keeps EMUB_SLOTS
 * factorisation slots on the device, so the table is an LRU cache: at most EMUB_GLUE_MODELS (default 8) models, and
 * when the device is full the least recently used ones go first (a caller that walks over many modelstructs -- the
 * per-component loop of estimate_multi, the R entry points -- no longer accumulates them until ENOMEM).
 * Call with g_call_mu held. */
static emub_model *glue_model_get(const gsl_matrix *x, const gsl_vector *y, int n, int d, int kernel, int order, int slots)
{
	const unsigned long long hash = glue_hash_xy(x, y, n, d);
	const void *kx = x ? (const void *)x->data : NULL, *ky = y ? (const void *)y->data : NULL;
	emub_ctx *ctx = glue_ctx();
	pthread_mutex_lock(&g_mu);
	for (int i = 0; i < g_nmodels; i++)
		if (g_models[i].key_x == kx && g_models[i].key_y == ky && g_models[i].n == n && g_models[i].d == d &&
		    g_models[i].kernel == kernel && g_models[i].order == order) {
			if (g_models[i].hash == hash) {
				emub_model *m = g_models[i].m;
				g_models[i].last_use = ++g_clock;
				pthread_mutex_unlock(&g_mu);
				return m;
			}
			glue_drop_model(i); /* same buffers, other contents: the old engine copy (and what was cached from it) is dead */
			break;
		}
	const int cap = env_int("EMUB_GLUE_MODELS", 8);
	while (g_nmodels >= (cap < GLUE_MAX ? (cap > 0 ? cap : 1) : GLUE_MAX)) {
		const int v = glue_lru_victim();
		if (v < 0) break;
		glue_drop_model(v);
	}
	if (g_nmodels == GLUE_MAX) { fprintf(stderr, "libemu_glue: too many models in use\n"); exit(EXIT_FAILURE); }
	/* row-major copies: xmodel may have a row stride, training_vector may be a strided view */
	double *X = (double *)calloc((size_t)n * (size_t)d, sizeof(double));
	double *yv = (double *)calloc((size_t)n, sizeof(double));
	if (x) for (int i = 0; i < n; i++) memcpy(X + (size_t)i * d, x->data + (size_t)i * x->tda, sizeof(double) * (size_t)d);
	if (y) for (int i = 0; i < n; i++) yv[i] = gsl_vector_get(y, i);
	emub_model *m = NULL;
	for (;;) {
		const int rc = emub_model_create(ctx, X, d, n, d, yv, kernel, order, slots, &m);
		if (rc == EMUB_OK) break;
		const int v = (rc == EMUB_ENOMEM) ? glue_lru_victim() : -1;
		if (v < 0) glue_die("emub_model_create");
		glue_drop_model(v);
	}
	free(X); free(yv);
	g_models[g_nmodels].key_x = kx; g_models[g_nmodels].key_y = ky;
	g_models[g_nmodels].n = n; g_models[g_nmodels].d = d; g_models[g_nmodels].kernel = kernel; g_models[g_nmodels].order = order;
	g_models[g_nmodels].m = m; g_models[g_nmodels].hash = hash; g_models[g_nmodels].last_use = ++g_clock;
	g_nmodels++;
	pthread_mutex_unlock(&g_mu);
	return m;
}

/* the engine model that belongs to a reference modelstruct.  o: the optstruct the caller was handed next to the model
 * (the R entry points build their modelstruct with alloc_modelstruct, which leaves the_model->options NULL,
 * modelstruct.c:12-19), else the model's own */
static emub_model *glue_model_for_opts(modelstruct *ms, optstruct *o)
{
	if (!o) o = ms->options;
	return glue_model_get(ms->xmodel, ms->training_vector, o->nmodel_points, o->nparams, o->cov_fn_index, o->regression_order,
	                      env_int("EMUB_SLOTS", 8));
}
static emub_model *glue_model_for(modelstruct *ms) { return glue_model_for_opts(ms, NULL); }

static int glue_kernel_of(double (*fn)(gsl_vector *, gsl_vector *, gsl_vector *, int, int))
{
	if (fn == covariance_fn_matern_three) return EMUB_MATERN32;
	if (fn == covariance_fn_matern_five) return EMUB_MATERN52;
	return EMUB_POWEREXP;
}

/* release every engine object (tests call this between cases) */
void libemu_glue_reset(void)
{
	pthread_mutex_lock(&g_call_mu);
	pthread_mutex_lock(&g_mu);
	while (g_nmodels > 0) glue_drop_model(g_nmodels - 1);
	g_nemus = g_nquick = 0;
	if (g_ctx) { emub_ctx_destroy(g_ctx); g_ctx = NULL; }
	pthread_mutex_unlock(&g_mu);
	pthread_mutex_unlock(&g_call_mu);
}

/* how many engine models the glue holds (tests) */
int libemu_glue_model_count(void)
{
	pthread_mutex_lock(&g_mu);
	const int n = g_nmodels;
	pthread_mutex_unlock(&g_mu);
	return n;
}

static int glue_eval(const gsl_vector *theta_vec, void *params_in, int want_grad, double *negL, double *grad, double *sigma2)
{
	struct estimate_thetas_params *p = (struct estimate_thetas_params *)params_in;
	double th[GLUE_TH];
	int status = 0;
	if (theta_vec->size > GLUE_TH) { fprintf(stderr, "libemu_glue: %zu thetas: more than the engine takes\n", theta_vec->size); exit(EXIT_FAILURE); }
	for (size_t i = 0; i < theta_vec->size; i++) th[i] = gsl_vector_get(theta_vec, i);
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_for_opts(p->the_model, p->options);
	const int rc = emub_loglik_grad_batch(m, th, 1, want_grad, negL, grad, sigma2, &status);
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_loglik_grad_batch");
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
	double negL, g[GLUE_TH];
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
	double negL, g[GLUE_TH];
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
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_for_opts(the_model, options);
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
	pthread_mutex_unlock(&g_call_mu);
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
	double th[GLUE_TH], beta[GLUE_TH];
	for (int i = 0; i < e->nthetas; i++) th[i] = gsl_vector_get(model->thetas, i);
	emub_emulator *eh = NULL;
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_for(model);
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
	pthread_mutex_unlock(&g_call_mu);
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
	pthread_mutex_lock(&g_call_mu);
	emub_emulator *eh = glue_emulator_for(e, 1);
	if (eh) emub_emulator_destroy(eh);
	pthread_mutex_unlock(&g_call_mu);
	gsl_vector_free(e->beta_vector);
	free((void *)e);
}

/* emulator_struct.c:124 */
void emulate_point(emulator_struct *e, gsl_vector *point, double *mean, double *variance)
{
	double x[GLUE_TH];
	for (int i = 0; i < e->nparams; i++) x[i] = gsl_vector_get(point, i);
	pthread_mutex_lock(&g_call_mu);
	emub_emulator *eh = glue_emulator_for(e, 0);
	const int rc = eh ? emub_predict_few(eh, x, e->nparams, 1, mean, variance) : EMUB_EINVAL;
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_predict_few");
}

/* emulate-fns.c:73 -- covariance, factorisation and regression once, then every point of the list: one emulator, one
 * batched prediction (options->nemulate_points rows of point_list) */
void emulateAtPointList(modelstruct *the_model, gsl_matrix *point_list, optstruct *options, double *the_mean, double *the_variance)
{
	double th[GLUE_TH];
	for (int i = 0; i < options->nthetas; i++) th[i] = gsl_vector_get(the_model->thetas, i);
	emub_emulator *eh = NULL;
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_for_opts(the_model, options);
	if (emub_emulator_create(m, th, &eh) != EMUB_OK) { /* chol_inverse_cov_matrix exits on a failed factorisation, :282-285 */
		fprintf(stderr, "emulateAtPointList: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	if (emub_predict_batch(eh, point_list->data, (int)point_list->tda, options->nemulate_points, the_mean, the_variance) != EMUB_OK)
		glue_die("emub_predict_batch");
	emub_emulator_destroy(eh);
	pthread_mutex_unlock(&g_call_mu);
}

/* the emulator of (model, thetas), kept between calls: what struct emulateMCData's host-side C^-1 is to the reference's
 * Monte-Carlo fast path (rbind.c:299-425).  g_call_mu held. */
static emub_emulator *glue_quick_emulator(modelstruct *the_model, optstruct *options)
{
	double th[GLUE_TH];
	const int nth = options->nthetas;
	for (int i = 0; i < nth; i++) th[i] = gsl_vector_get(the_model->thetas, i);
	emub_model *m = glue_model_for_opts(the_model, options);
	pthread_mutex_lock(&g_mu);
	for (int k = 0; k < g_nquick; k++)
		if (g_quick[k].m == m && g_quick[k].nth == nth && memcmp(g_quick[k].th, th, sizeof(double) * (size_t)nth) == 0) {
			g_quick[k].last_use = ++g_clock;
			emub_emulator *e = g_quick[k].e;
			pthread_mutex_unlock(&g_mu);
			return e;
		}
	if (g_nquick == GLUE_QUICK) {
		int v = 0;
		for (int k = 1; k < g_nquick; k++) if (g_quick[k].last_use < g_quick[v].last_use) v = k;
		emub_emulator_destroy(g_quick[v].e);
		g_quick[v] = g_quick[--g_nquick];
	}
	pthread_mutex_unlock(&g_mu);
	emub_emulator *e = NULL;
	if (emub_emulator_create(m, th, &e) != EMUB_OK) { /* emulate-fns.c:282-285 */
		fprintf(stderr, "libemu_glue: emulator set-up: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	pthread_mutex_lock(&g_mu);
	g_quick[g_nquick].m = m; g_quick[g_nquick].nth = nth; memcpy(g_quick[g_nquick].th, th, sizeof(double) * (size_t)nth);
	g_quick[g_nquick].e = e; g_quick[g_nquick].last_use = ++g_clock;
	g_nquick++;
	pthread_mutex_unlock(&g_mu);
	return e;
}

/* emulate-fns.c:138 -- the reference rebuilds C, its inverse and beta on every call; here the factor of (model, thetas)
 * is kept, so a caller that loops over points pays the n^3 set-up once */
void emulateAtPoint(modelstruct *the_model, gsl_vector *the_point, optstruct *options, double *the_mean, double *the_variance)
{
	double x[GLUE_TH];
	for (int i = 0; i < options->nparams; i++) x[i] = gsl_vector_get(the_point, i);
	pthread_mutex_lock(&g_call_mu);
	emub_emulator *eh = glue_quick_emulator(the_model, options);
	const int rc = emub_predict_few(eh, x, options->nparams, 1, the_mean, the_variance);
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_predict_few");
}

/* emulate-fns.c:201-232 -- the Monte-Carlo fast path of the R binding (callEmulateMC / callEmulateMCMulti,
 * rbind.c:403-425, :535-567): one point against precomputed set-up data.  The reference passes its host-side C^-1, H
 * and beta (struct emulateMCData); the engine's counterpart is the cached factor of (the_model, the_model->thetas),
 * built on the first call and reused for every later point -- the arguments h_matrix / cinverse / beta_vector hold the
 * same quantities (setupEmulateMCHelper computed them from the same model and thetas) and are not read. */
void emulateQuick(modelstruct *the_model, gsl_vector *the_point, optstruct *options, double *mean_out, double *var_out,
                  gsl_matrix *h_matrix, gsl_matrix *cinverse, gsl_vector *beta_vector)
{
	(void)h_matrix; (void)cinverse; (void)beta_vector;
	emulateAtPoint(the_model, the_point, options, mean_out, var_out);
}

/* emulate-fns.c:13 -- the reference builds C, C^-1 and beta on the host and walks emulate_ith_location over the rows of
 * results->new_x; here: one emulator, one batched prediction.  Its two debugging prints are kept (regression
 * coefficients to stderr, the first coordinate of every point to stdout, :39-50). */
void emulate_model_results(modelstruct *the_model, optstruct *options, resultstruct *results)
{
	const int mq = options->nemulate_points;
	double th[GLUE_TH], beta[GLUE_TH];
	for (int i = 0; i < options->nthetas; i++) th[i] = gsl_vector_get(the_model->thetas, i);
	double *mean = (double *)malloc(sizeof(double) * (size_t)(mq > 0 ? mq : 1));
	double *var = (double *)malloc(sizeof(double) * (size_t)(mq > 0 ? mq : 1));
	emub_emulator *eh = NULL;
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_for_opts(the_model, options);
	if (emub_emulator_create(m, th, &eh) != EMUB_OK) { /* chol_inverse_cov_matrix exits on a failed factorisation, :282-285 */
		fprintf(stderr, "emulate_model_results: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	emub_emulator_beta(eh, beta);
	const int rc = emub_predict_batch(eh, results->new_x->data, (int)results->new_x->tda, mq, mean, var);
	emub_emulator_destroy(eh);
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_predict_batch");
	fprintf(stderr, "regression cpts: ");
	for (int i = 0; i < options->nregression_fns; i++) fprintf(stderr, "%g ", beta[i]);
	fprintf(stderr, "\n");
	for (int i = 0; i < mq; i++) printf("%g\n", gsl_matrix_get(results->new_x, i, 0));
	for (int i = 0; i < mq; i++) {
		gsl_vector_set(results->emulated_mean, i, mean[i]);
		gsl_vector_set(results->emulated_var, i, var[i]);
	}
	free(mean); free(var);
}

/* emulate-fns.c:234 -- one row of results->new_x against precomputed host-side set-up data: like emulateQuick, the
 * engine's cached factor of (the_model, the_model->thetas) stands in for h_matrix / cinverse / beta_vector */
void emulate_ith_location(modelstruct *the_model, optstruct *options, resultstruct *results, int i, gsl_matrix *h_matrix,
                          gsl_matrix *cinverse, gsl_vector *beta_vector)
{
	(void)h_matrix; (void)cinverse; (void)beta_vector;
	gsl_vector_view row = gsl_matrix_row(results->new_x, (size_t)i);
	double mean = 0.0, var = 0.0;
	emulateAtPoint(the_model, &row.vector, options, &mean, &var);
	gsl_vector_set(results->emulated_mean, i, mean);
	gsl_vector_set(results->emulated_var, i, var);
}

/* emulator.c:636 -- the kernel is identified by the function pointer the caller passes */
void makeCovMatrix_fnptr(gsl_matrix *cov_matrix, gsl_matrix *xmodel, gsl_vector *thetas, int nmodel_points, int nthetas,
                         int nparams, double (*covariance_fn_ptr)(gsl_vector *, gsl_vector *, gsl_vector *, int, int))
{
	double th[GLUE_TH];
	for (int i = 0; i < nthetas; i++) th[i] = gsl_vector_get(thetas, i);
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_get(xmodel, NULL, nmodel_points, nparams, glue_kernel_of(covariance_fn_ptr), 0, 1);
	const int rc = emub_cov_matrix(m, th, cov_matrix->data, (int)cov_matrix->tda);
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_cov_matrix");
}

/* emulator.c:607 -- the same through the process-global kernel pointer (emulator.h:13) */
void makeCovMatrix(gsl_matrix *cov_matrix, gsl_matrix *xmodel, gsl_vector *thetas, int nmodel_points, int nthetas, int nparams)
{
	makeCovMatrix_fnptr(cov_matrix, xmodel, thetas, nmodel_points, nthetas, nparams, covariance_fn);
}

/* emulator.c:578 -- k_i = c(x_i, xnew), values below 1e-10 set to 0 (:588-590) */
void makeKVector_fnptr(gsl_vector *kvector, gsl_matrix *xmodel, gsl_vector *xnew, gsl_vector *thetas, int nmodel_points, int nthetas,
                       int nparams, double (*covariance_fn_ptr)(gsl_vector *, gsl_vector *, gsl_vector *, int, int))
{
	double th[GLUE_TH], x[GLUE_TH];
	for (int i = 0; i < nthetas; i++) th[i] = gsl_vector_get(thetas, i);
	for (int i = 0; i < nparams; i++) x[i] = gsl_vector_get(xnew, i);
	double *k = (double *)malloc(sizeof(double) * (size_t)nmodel_points);
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_get(xmodel, NULL, nmodel_points, nparams, glue_kernel_of(covariance_fn_ptr), 0, 1);
	const int rc = emub_k_vectors(m, th, x, nparams, 1, k, 1);
	pthread_mutex_unlock(&g_call_mu);
	if (rc != EMUB_OK) glue_die("emub_k_vectors");
	for (int i = 0; i < nmodel_points; i++) gsl_vector_set(kvector, i, k[i]);
	free(k);
}

/* emulator.c:553 */
void makeKVector(gsl_vector *kvector, gsl_matrix *xmodel, gsl_vector *xnew, gsl_vector *thetas, int nmodel_points, int nthetas, int nparams)
{
	makeKVector_fnptr(kvector, xmodel, xnew, thetas, nmodel_points, nthetas, nparams, covariance_fn);
}

/* emulate-fns.c:275-300 -- Cholesky, determinant and explicit inverse of the caller's covariance matrix: the n^3 step of
 * setupEmulateMCHelper (rbind.c:378) and emulate_model_results.  The determinant is returned as the reference forms
 * it, (prod L_ii)^2 (:290-293), here exp(2 sum log L_ii): 0 or inf where the running product under- or overflows. */
void chol_inverse_cov_matrix(optstruct *options, gsl_matrix *temp_matrix, gsl_matrix *result_matrix, double *final_determinant_c)
{
	const int n = options->nmodel_points;
	double logdet = 0.0;
	pthread_mutex_lock(&g_call_mu);
	emub_model *m = glue_model_get(NULL, NULL, n, 1, EMUB_POWEREXP, 0, 1);
	const int rc = emub_spd_inverse(m, temp_matrix->data, (int)temp_matrix->tda, result_matrix->data, (int)result_matrix->tda, &logdet);
	pthread_mutex_unlock(&g_call_mu);
	if (rc == EMUB_EDOM) {
		fprintf(stderr, "trying to cholesky a non postive def matrix, in emulate-fns.c sorry...\n"); /* :283 */
		exit(1);
	}
	if (rc != EMUB_OK) glue_die("emub_spd_inverse");
	/* the reference leaves the inverse in temp_matrix too (its in-place GSL inversion, :296-297) */
	for (int i = 0; i < n; i++)
		memcpy(temp_matrix->data + (size_t)i * temp_matrix->tda, result_matrix->data + (size_t)i * result_matrix->tda, sizeof(double) * (size_t)n);
	*final_determinant_c = exp(logdet);
}
