/*
 * ref_driver.c -- flat C entry points (ctypes-friendly: plain pointers and ints) around the
 * reference's OWN hot-path functions, which are compiled unmodified from /root/reference/src
 * against oracle/gsl_shim (see oracle/Makefile) into oracle/_ref/libemu_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (to pin the restatement oracle and to generate
 * tests/golden/), by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * leg.  Nothing under madaiemulator_b200/ links or loads this.
 *
 * Every function here only marshals arguments and calls the reference symbol named in its
 * comment; no numerics are restated in this file except ref_eval_logsum's determinant term
 * (deviation D-1).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <time.h>

#include "modelstruct.h"
#include "optstruct.h"
#include "emulator_struct.h"
#include "resultstruct.h"
#include "multi_modelstruct.h"
#include "multivar_support.h"
#include "libEmu/emulator.h"
#include "libEmu/regression.h"
#include "libEmu/maxmultimin.h"
#include "libEmu/estimator-fns.h"
#include "libEmu/emulate-fns.h"
#include "libEmu/estimate_threaded.h"

typedef struct {
	modelstruct *model;
	gsl_matrix *xmodel_in;
	gsl_vector *training;
	struct estimate_thetas_params params;
	int have_h;
} ref_model;

/* alloc_modelstruct_2 (src/modelstruct.c:282) prints the ranges to stdout; silence it so the
 * oracle can run inside pytest / bench without noise. */
static int silence_stdout(void)
{
	fflush(stdout);
	int saved = dup(1);
	FILE *devnull = fopen("/dev/null", "w");
	if (devnull) { dup2(fileno(devnull), 1); fclose(devnull); }
	return saved;
}
static void restore_stdout(int saved)
{
	fflush(stdout);
	if (saved >= 0) { dup2(saved, 1); close(saved); }
}

/* -> alloc_modelstruct_2, src/modelstruct.c:282 */
void *ref_model_create(const double *X, int n, int d, const double *y, int cov_fn_index, int regression_order)
{
	ref_model *r = (ref_model *)calloc(1, sizeof(ref_model));
	r->xmodel_in = gsl_matrix_alloc(n, d);
	r->training = gsl_vector_alloc(n);
	memcpy(r->xmodel_in->data, X, sizeof(double) * (size_t)n * d);
	memcpy(r->training->data, y, sizeof(double) * (size_t)n);
	int saved = silence_stdout();
	r->model = alloc_modelstruct_2(r->xmodel_in, r->training, cov_fn_index, regression_order);
	restore_stdout(saved);
	gsl_vector_set_zero(r->model->thetas);
	r->params.options = r->model->options;
	r->params.the_model = r->model;
	r->params.random_number = NULL;
	r->params.max_tries = 1;
	r->params.h_matrix = gsl_matrix_alloc(n, r->model->options->nregression_fns);
	/* -> makeHMatrix_fnptr, src/libEmu/regression.c:100 (as maxmultimin.c:78 does) */
	makeHMatrix_fnptr(r->params.h_matrix, r->model->xmodel, n, d, r->model->options->nregression_fns,
	                  r->model->makeHVector);
	r->have_h = 1;
	return r;
}

void ref_model_free(void *h)
{
	ref_model *r = (ref_model *)h;
	if (!r) return;
	gsl_matrix_free(r->params.h_matrix);
	gsl_matrix_free(r->model->xmodel);
	free_modelstruct_2(r->model);
	gsl_matrix_free(r->xmodel_in);
	gsl_vector_free(r->training);
	free(r);
}

int ref_model_nthetas(void *h) { return ((ref_model *)h)->model->options->nthetas; }
int ref_model_nregression_fns(void *h) { return ((ref_model *)h)->model->options->nregression_fns; }

/* grad_ranges as set by setup_optimization_ranges, src/optstruct.c:142; out is nthetas x 2 */
void ref_model_ranges(void *h, double *out)
{
	ref_model *r = (ref_model *)h;
	int nt = r->model->options->nthetas;
	for (int i = 0; i < nt; i++) {
		out[2 * i] = gsl_matrix_get(r->model->options->grad_ranges, i, 0);
		out[2 * i + 1] = gsl_matrix_get(r->model->options->grad_ranges, i, 1);
	}
}
/* -> setup_optimization_ranges, src/optstruct.c:142, with the optstruct's two switches set by the caller (what
 * callEstimate's use_fixed_nugget does, src/libRbind/rbind.c:43-60): the model's options are modified for the call and
 * restored */
void ref_model_ranges_ex(void *h, int use_data_scales, int fixed_nugget_mode, double fixed_nugget, double *out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	const int s_uds = o->use_data_scales, s_fnm = o->fixed_nugget_mode;
	const double s_fn = o->fixed_nugget;
	gsl_matrix *s_ranges = o->grad_ranges;
	o->use_data_scales = use_data_scales; o->fixed_nugget_mode = fixed_nugget_mode; o->fixed_nugget = fixed_nugget;
	int saved = silence_stdout();
	setup_optimization_ranges(o, r->model); /* allocates a fresh grad_ranges */
	restore_stdout(saved);
	for (int i = 0; i < o->nthetas; i++) {
		out[2 * i] = gsl_matrix_get(o->grad_ranges, i, 0);
		out[2 * i + 1] = gsl_matrix_get(o->grad_ranges, i, 1);
	}
	gsl_matrix_free(o->grad_ranges);
	o->grad_ranges = s_ranges; o->use_data_scales = s_uds; o->fixed_nugget_mode = s_fnm; o->fixed_nugget = s_fn;
}
void ref_model_sample_scales(void *h, double *out)
{
	ref_model *r = (ref_model *)h;
	for (int i = 0; i < r->model->options->nparams; i++) out[i] = gsl_vector_get(r->model->sample_scales, i);
}

/* -> makeCovMatrix_fnptr, src/libEmu/emulator.c:636; thetas is the FULL vector (nthetas) */
void ref_cov_matrix(void *h, const double *thetas, double *C_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	gsl_vector_view tv = gsl_vector_view_array((double *)thetas, o->nthetas);
	gsl_matrix_view cv = gsl_matrix_view_array(C_out, o->nmodel_points, o->nmodel_points);
	makeCovMatrix_fnptr(&cv.matrix, r->model->xmodel, &tv.vector, o->nmodel_points, o->nthetas, o->nparams,
	                    r->model->covariance_fn);
}

/* -> makeKVector_fnptr, src/libEmu/emulator.c:578 (1e-10 clamp :588-590); k_out has nmodel_points entries */
void ref_k_vector(void *h, const double *thetas, const double *xnew, double *k_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	gsl_vector_view tv = gsl_vector_view_array((double *)thetas, o->nthetas);
	gsl_vector_view xv = gsl_vector_view_array((double *)xnew, o->nparams);
	gsl_vector_view kv = gsl_vector_view_array(k_out, o->nmodel_points);
	makeKVector_fnptr(&kv.vector, r->model->xmodel, &xv.vector, &tv.vector, o->nmodel_points, o->nthetas, o->nparams,
	                  r->model->covariance_fn);
}

/* -> chol_inverse_cov_matrix, src/libEmu/emulate-fns.c:275: A (n x n, destroyed by the reference: pass a copy) ->
 * inverse and determinant */
void ref_chol_inverse(void *h, const double *A, double *Ainv_out, double *det_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	const int n = o->nmodel_points;
	gsl_matrix *tmp = gsl_matrix_alloc(n, n);
	memcpy(tmp->data, A, sizeof(double) * (size_t)n * n);
	gsl_matrix_view out = gsl_matrix_view_array(Ainv_out, n, n);
	chol_inverse_cov_matrix(o, tmp, &out.matrix, det_out);
	gsl_matrix_free(tmp);
}

/* one covariance value -> model->covariance_fn (covariance_fn_gaussian :101 / matern :344,:438) */
double ref_cov_pair(void *h, const double *xa, const double *xb, const double *thetas)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	gsl_vector_view a = gsl_vector_view_array((double *)xa, o->nparams);
	gsl_vector_view b = gsl_vector_view_array((double *)xb, o->nparams);
	gsl_vector_view tv = gsl_vector_view_array((double *)thetas, o->nthetas);
	return r->model->covariance_fn(&a.vector, &b.vector, &tv.vector, o->nthetas, o->nparams);
}

/* -> model->makeGradMatLength (derivative_l_gauss, emulator.c:173); index in [2, nthetas) */
void ref_deriv_matrix(void *h, double theta_length, int index, double *D_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	gsl_matrix_view dv = gsl_matrix_view_array(D_out, o->nmodel_points, o->nmodel_points);
	r->model->makeGradMatLength(&dv.matrix, r->model->xmodel, theta_length, index, o->nmodel_points, o->nparams);
}

void ref_h_matrix(void *h, double *H_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	for (int i = 0; i < o->nmodel_points; i++)
		for (int j = 0; j < o->nregression_fns; j++)
			H_out[(size_t)i * o->nregression_fns + j] = gsl_matrix_get(r->params.h_matrix, i, j);
}

/* -> evalFnMulti, src/libEmu/maxmultimin.c:288 ; returns -L (literal, incl. the determinant product) */
double ref_eval(void *h, const double *theta_less_amp)
{
	ref_model *r = (ref_model *)h;
	gsl_vector_view tv = gsl_vector_view_array((double *)theta_less_amp, r->model->options->nthetas - 1);
	return evalFnMulti(&tv.vector, &r->params);
}

/* -> gradFnMulti, src/libEmu/maxmultimin.c:416 (exits the process on a non-PD matrix, like the reference) */
void ref_grad(void *h, const double *theta_less_amp, double *g_out)
{
	ref_model *r = (ref_model *)h;
	int nto = r->model->options->nthetas - 1;
	gsl_vector_view tv = gsl_vector_view_array((double *)theta_less_amp, nto);
	gsl_vector_view gv = gsl_vector_view_array(g_out, nto);
	gradFnMulti(&tv.vector, &r->params, &gv.vector);
}

/* -> evalFnGradMulti, src/libEmu/maxmultimin.c:615 */
void ref_eval_grad(void *h, const double *theta_less_amp, double *f_out, double *g_out)
{
	ref_model *r = (ref_model *)h;
	int nto = r->model->options->nthetas - 1;
	gsl_vector_view tv = gsl_vector_view_array((double *)theta_less_amp, nto);
	gsl_vector_view gv = gsl_vector_view_array(g_out, nto);
	evalFnGradMulti(&tv.vector, &r->params, f_out, &gv.vector);
}

/* -> estimateSigmaFull, src/libEmu/maxmultimin.c:148 ; argument is theta without the amplitude */
double ref_sigma_full(void *h, const double *theta_less_amp)
{
	ref_model *r = (ref_model *)h;
	gsl_vector_view tv = gsl_vector_view_array((double *)theta_less_amp, r->model->options->nthetas - 1);
	return estimateSigmaFull(&tv.vector, &r->params);
}

/*
 * The literal evalFnMulti with ONE change (deviation D-1): the determinant is accumulated as
 * sum 2 log L_ii instead of the running product (maxmultimin.c:355-358) which underflows to 0
 * for n >~ 700.  Everything else goes through the reference's functions.
 * Outputs: *negL, and optionally logdet, sigma2, beta (nregression_fns).
 * Returns 0, or 1 when the matrix is not positive definite (negL = NAN, like :349).
 */
int ref_eval_logsum(void *h, const double *theta_less_amp, double *negL, double *logdet_out,
                    double *sigma2_out, double *beta_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	int n = o->nmodel_points, nt = o->nthetas;
	gsl_matrix *c = gsl_matrix_alloc(n, n);
	gsl_vector *theta_local = gsl_vector_alloc(nt);
	gsl_vector_set(theta_local, 0, 0.0);
	for (int i = 1; i < nt; i++) gsl_vector_set(theta_local, i, theta_less_amp[i - 1]);
	makeCovMatrix_fnptr(c, r->model->xmodel, theta_local, n, nt, o->nparams, r->model->covariance_fn);
	gsl_error_handler_t *old = gsl_set_error_handler_off();
	int rc = gsl_linalg_cholesky_decomp(c);
	gsl_set_error_handler(old);
	if (rc == GSL_EDOM) {
		*negL = NAN;
		gsl_matrix_free(c); gsl_vector_free(theta_local);
		return 1;
	}
	double logdet = 0.0;
	for (int i = 0; i < n; i++) logdet += 2.0 * log(gsl_matrix_get(c, i, i));
	gsl_linalg_cholesky_invert(c);
	double sigma2 = estimateSigma(c, &r->params);
	/* det = 1 -> log det = 0 inside getLogLikelyhood (estimator-fns.c:56); add the real term here */
	double L = getLogLikelyhood(c, 1.0, r->model->xmodel, r->model->training_vector, theta_local,
	                            r->params.h_matrix, n, nt, o->nparams, o->nregression_fns, r->model->makeHVector);
	L += -0.5 * logdet;
	*negL = -L;
	if (logdet_out) *logdet_out = logdet;
	if (sigma2_out) *sigma2_out = sigma2;
	if (beta_out) {
		gsl_vector *beta = gsl_vector_alloc(o->nregression_fns);
		estimateBeta(beta, r->params.h_matrix, c, r->model->training_vector, n, o->nregression_fns);
		for (int i = 0; i < o->nregression_fns; i++) beta_out[i] = gsl_vector_get(beta, i);
		gsl_vector_free(beta);
	}
	gsl_matrix_free(c);
	gsl_vector_free(theta_local);
	return 0;
}

/* C^-1 at theta_local = (0, theta') through gsl_linalg_cholesky_decomp/_invert as evalFnMulti does */
int ref_cinverse(void *h, const double *theta_less_amp, double *cinv_out)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	int n = o->nmodel_points, nt = o->nthetas;
	gsl_matrix_view cv = gsl_matrix_view_array(cinv_out, n, n);
	gsl_vector *theta_local = gsl_vector_alloc(nt);
	gsl_vector_set(theta_local, 0, 0.0);
	for (int i = 1; i < nt; i++) gsl_vector_set(theta_local, i, theta_less_amp[i - 1]);
	makeCovMatrix_fnptr(&cv.matrix, r->model->xmodel, theta_local, n, nt, o->nparams, r->model->covariance_fn);
	gsl_vector_free(theta_local);
	gsl_error_handler_t *old = gsl_set_error_handler_off();
	int rc = gsl_linalg_cholesky_decomp(&cv.matrix);
	gsl_set_error_handler(old);
	if (rc == GSL_EDOM) return 1;
	gsl_linalg_cholesky_invert(&cv.matrix);
	return 0;
}

/* ---- prediction ------------------------------------------------------------------------------ */
/* -> alloc_emulator_struct, src/emulator_struct.c:13 ; thetas is the FULL vector */
void *ref_emulator_create(void *h, const double *thetas)
{
	ref_model *r = (ref_model *)h;
	for (int i = 0; i < r->model->options->nthetas; i++) gsl_vector_set(r->model->thetas, i, thetas[i]);
	return alloc_emulator_struct(r->model);
}
void ref_emulator_free(void *e) { free_emulator_struct((emulator_struct *)e); }

/* -> emulate_point, src/emulator_struct.c:124 ; pts is m x nparams row-major */
void ref_emulate(void *e, const double *pts, int m, double *mean, double *var)
{
	emulator_struct *es = (emulator_struct *)e;
	for (int i = 0; i < m; i++) {
		gsl_vector_view pv = gsl_vector_view_array((double *)pts + (size_t)i * es->nparams, es->nparams);
		emulate_point(es, &pv.vector, &mean[i], &var[i]);
	}
}
/* -> emulateAtPointList, src/libEmu/emulate-fns.c:73 (the R binding's list entry point, rbind.c:178) and, with
 * single != 0, emulateAtPoint (:138) point by point; thetas is the FULL vector, pts m x nparams row-major */
void ref_emulate_at_point_list(void *h, const double *thetas, const double *pts, int m, int single, double *mean, double *var)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	for (int i = 0; i < o->nthetas; i++) gsl_vector_set(r->model->thetas, i, thetas[i]);
	const int saved = o->nemulate_points;
	int so = silence_stdout();
	if (single) {
		for (int i = 0; i < m; i++) {
			gsl_vector_view pv = gsl_vector_view_array((double *)pts + (size_t)i * o->nparams, o->nparams);
			emulateAtPoint(r->model, &pv.vector, o, &mean[i], &var[i]);
		}
	} else {
		o->nemulate_points = m;
		gsl_matrix_view pl = gsl_matrix_view_array((double *)pts, m, o->nparams);
		emulateAtPointList(r->model, &pl.matrix, o, mean, var);
		o->nemulate_points = saved;
	}
	restore_stdout(so);
}

/* -> emulate_model_results, src/libEmu/emulate-fns.c:13: (mean, variance) at the rows of results->new_x (the model's
 * global covariance / regression function pointers are the ones alloc_modelstruct_2 set) */
void ref_emulate_model_results(void *h, const double *thetas, const double *pts, int m, double *mean, double *var)
{
	ref_model *r = (ref_model *)h;
	optstruct *o = r->model->options;
	for (int i = 0; i < o->nthetas; i++) gsl_vector_set(r->model->thetas, i, thetas[i]);
	const int saved = o->nemulate_points;
	o->nemulate_points = m;
	resultstruct res;
	gsl_matrix_view nx = gsl_matrix_view_array((double *)pts, m, o->nparams);
	gsl_vector_view mv = gsl_vector_view_array(mean, m), vv = gsl_vector_view_array(var, m);
	res.new_x = &nx.matrix; res.emulated_mean = &mv.vector; res.emulated_var = &vv.vector; res.options = o; res.model = r->model;
	int so = silence_stdout(); /* it prints the first coordinate of every point */
	emulate_model_results(r->model, o, &res);
	restore_stdout(so);
	o->nemulate_points = saved;
}

void ref_emulator_beta(void *e, double *beta_out)
{
	emulator_struct *es = (emulator_struct *)e;
	for (int i = 0; i < es->nregression_fns; i++) beta_out[i] = gsl_vector_get(es->beta_vector, i);
}

/* ---- training (restart loop) --------------------------------------------------------------------- */
/* -> maxWithMultiMin, src/libEmu/maxmultimin.c:47 with a seeded mt19937 (the reference seeds
 * from /dev/urandom, estimate_threaded.c:155-161).  thetas_out: nthetas.  Returns best log-lik. */
double ref_max_with_multimin(void *h, int max_tries, unsigned long seed, double *thetas_out)
{
	ref_model *r = (ref_model *)h;
	struct estimate_thetas_params p = r->params;
	p.random_number = gsl_rng_alloc(gsl_rng_default);
	gsl_rng_set(p.random_number, seed);
	p.max_tries = max_tries;
	p.h_matrix = NULL; /* maxWithMultiMin allocates its own (maxmultimin.c:75) */
	maxWithMultiMin(&p);
	for (int i = 0; i < r->model->options->nthetas; i++) thetas_out[i] = gsl_vector_get(r->model->thetas, i);
	gsl_rng_free(p.random_number);
	gsl_matrix_free(p.h_matrix);
	return p.lhood_current;
}

/* -> set_random_init_value, maxmultimin.c:789 : draws `count` start vectors (nthetas each) */
void ref_random_inits(void *h, unsigned long seed, int count, double *out)
{
	ref_model *r = (ref_model *)h;
	int nt = r->model->options->nthetas;
	gsl_rng *rng = gsl_rng_alloc(gsl_rng_default);
	gsl_rng_set(rng, seed);
	gsl_vector *x = gsl_vector_alloc(nt);
	for (int c = 0; c < count; c++) {
		set_random_init_value(rng, x, r->model->options->grad_ranges, nt);
		for (int i = 0; i < nt; i++) out[(size_t)c * nt + i] = gsl_vector_get(x, i);
	}
	gsl_vector_free(x);
	gsl_rng_free(rng);
}

/* ---- timing helpers for the CPU baseline (bench.py) ----------------------------------------------- */
typedef struct {
	ref_model *model;
	const double *theta;
	int reps;
	double fsum;
} timing_job;

static void *eval_grad_worker(void *arg)
{
	timing_job *j = (timing_job *)arg;
	int nto = j->model->model->options->nthetas - 1;
	double *g = (double *)malloc(sizeof(double) * nto);
	for (int i = 0; i < j->reps; i++) {
		double f;
		ref_eval_grad(j->model, j->theta, &f, g);
		j->fsum += g[0];
	}
	free(g);
	return NULL;
}

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/*
 * Time evalFnGradMulti with `nthreads` threads, each running `reps` independent evaluations on its
 * own model copy -- the reference's own parallel model (one restart chain per pthread,
 * estimate_threaded.c:97,172).  models: array of nthreads handles from ref_model_create.
 * Returns wall seconds for nthreads*reps evaluations.
 */
double ref_time_eval_grad(void **models, int nthreads, const double *theta_less_amp, int reps)
{
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
	timing_job *jobs = (timing_job *)calloc(nthreads, sizeof(timing_job));
	double t0 = now_s();
	for (int i = 0; i < nthreads; i++) {
		jobs[i].model = (ref_model *)models[i];
		jobs[i].theta = theta_less_amp;
		jobs[i].reps = reps;
		pthread_create(&th[i], NULL, eval_grad_worker, &jobs[i]);
	}
	for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
	double t1 = now_s();
	free(th);
	free(jobs);
	return t1 - t0;
}

typedef struct {
	emulator_struct *emu;
	const double *pts;
	int m;
	double *mean, *var;
} emu_job;
static void *emulate_worker(void *arg)
{
	emu_job *j = (emu_job *)arg;
	ref_emulate(j->emu, j->pts, j->m, j->mean, j->var);
	return NULL;
}
/* emulate_point over m points split across nthreads threads sharing one emulator_struct
 * (emulate_point is documented thread-safe, emulator_struct.c:120-143).  Returns wall seconds. */
double ref_time_emulate(void *e, const double *pts, int m, int nthreads, double *mean, double *var)
{
	emulator_struct *es = (emulator_struct *)e;
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
	emu_job *jobs = (emu_job *)calloc(nthreads, sizeof(emu_job));
	int per = (m + nthreads - 1) / nthreads;
	double t0 = now_s();
	for (int i = 0; i < nthreads; i++) {
		int lo = i * per, hi = lo + per > m ? m : lo + per;
		if (lo > m) lo = m;
		if (hi < lo) hi = lo;
		jobs[i].emu = es;
		jobs[i].pts = pts + (size_t)lo * es->nparams;
		jobs[i].m = hi - lo;
		jobs[i].mean = mean + lo;
		jobs[i].var = var + lo;
		pthread_create(&th[i], NULL, emulate_worker, &jobs[i]);
	}
	for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
	double t1 = now_s();
	free(th);
	free(jobs);
	return t1 - t0;
}

int ref_ncpus(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n < 1 ? 1 : (int)n; }
