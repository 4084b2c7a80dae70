/* emub_estimate.c -- see emub_estimate.h.  Citations are file:line under the reference's src/. */
#include "emub_estimate.h"
#include "emub_bfgs.h"
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SCREWUPVALUE -2000000 /* maxmultimin.h */

void emub_estimate_default_opts(emub_estimate_opts *o)
{
	o->max_tries = 50;
	o->nchains = 0;
	o->seed = 1;
	o->step_size = 1.5;
	o->tol = 0.5;
	o->eps_abs = 0.1;
	o->step_max = 30;
	o->first_component = 0;
	o->component_stride = 1;
	o->polish_steps = 0;
	o->polish_eps = 1e-3;
	o->value_policy = EMUB_VALUE_ADAPTIVE;
}

/* modelstruct.c:188-213 */
void emub_sample_scales(const double *X, int ldx, int n, int d, double *scales)
{
	for (int i = 0; i < d; i++) {
		double min_value = (n > 1) ? fabs(X[(size_t)1 * ldx + i] - X[i]) : 1.0;
		for (int j = 1; j < n - 1; j++) {
			double value = fabs(X[(size_t)(j + 1) * ldx + i] - X[(size_t)j * ldx + i]);
			if (value < min_value) min_value = value;
		}
		if (min_value < 1.0e-5) min_value = 1.0e-5;
		scales[i] = min_value;
	}
}

/* optstruct.c:142-226 */
void emub_optimization_ranges_ex(int kernel, const double *X, int ldx, int n, int d, int use_data_scales, int fixed_nugget_mode,
                                 double fixed_nugget, double *ranges)
{
	const int nthetas = (kernel == EMUB_POWEREXP) ? d + 2 : 3;
	double *scales = (double *)malloc(sizeof(double) * (size_t)d);
	double range_min, range_max;
	emub_sample_scales(X, ldx, n, d, scales);
	if (kernel == EMUB_POWEREXP) { range_min = 0.0001; range_max = 5; }
	else { range_min = 0; range_max = 10.0; }
	ranges[0] = 0.0001; ranges[1] = range_max; /* amplitude (not optimised, kept for layout) */
	ranges[2] = -5.0; ranges[3] = -2.0;        /* nugget, optstruct.c:153-154 */
	for (int i = 2; i < nthetas; i++) {
		if (use_data_scales) { /* :179-203; otherwise the defaults above, :205-211 */
			if (kernel == EMUB_POWEREXP) {
				range_min = 0.5 * log(scales[i - 2]);
				range_max = log(25 * exp(range_min));
			} else {
				range_min = 0.5 * scales[i - 2];
			}
		}
		ranges[2 * i] = range_min;
		ranges[2 * i + 1] = range_max;
	}
	if (fixed_nugget_mode == 1) { /* :217-225: the lower end stays, the upper end is the fixed nugget + 20% */
		ranges[2] = -5.0;
		ranges[3] = fixed_nugget + 0.20 * fixed_nugget;
	}
	free(scales);
}

void emub_optimization_ranges(int kernel, const double *X, int ldx, int n, int d, double *ranges)
{
	emub_optimization_ranges_ex(kernel, X, ldx, n, d, 1, 0, 0.0, ranges);
}

static uint64_t splitmix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

void emub_random_init(unsigned long long seed, int try_index, const double *ranges, int nthetas, double *x)
{
	for (int i = 0; i < nthetas; i++) {
		uint64_t z = splitmix64(splitmix64(seed) + (uint64_t)try_index * 1024u + (uint64_t)i);
		double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
		x[i] = u * (ranges[2 * i + 1] - ranges[2 * i]) + ranges[2 * i];
	}
}

/* ---- the evaluation front ------------------------------------------------------------------------ */
enum { REQ_IDLE = 0, REQ_PENDING = 1, REQ_DONE = 2 };

struct front;
typedef struct {
	struct front *fr;
	int id;
	int comp;             /* training vector (PCA component) this chain optimises */
	int first_try, try_stride;
	unsigned long long seed;
	int state;
	int want_grad;        /* the pending request needs the gradient */
	double *x, *g;        /* request point (nth1), returned gradient */
	double f, sigma2;
	int status;
	/* last evaluated point, so that f(x) followed by df(x) costs one evaluation when the gradient is already there */
	double *cx, *cg;
	double cf, csigma2;
	int cstatus, cvalid, chas_grad, cgrad_used;
	/* value-policy bookkeeping (per chain, so independent of thread timing) */
	long long n_value_requests, n_followed;
	long long value_evals, repeated, unused_grads;
	/* result of this chain's restarts */
	double best_lhood;
	double *best_thetas;
	int success_count, finite_count;
} chain_t;

typedef struct front {
	emub_model *model;
	const emub_estimate_opts *opts;
	const double *ranges;
	const double *starts; /* optional max_tries x nthetas explicit start points */
	int starts_per_comp;  /* starts holds max_tries rows for every component, component-major */
	int nth, nth1, nchains;
	pthread_mutex_t mu;
	pthread_cond_t cv_disp, cv_done;
	int nactive, npending;
	chain_t *chains;
	long long evaluations, batches;
	int failed;
} front_t;

/* value_request: the optimiser asked for f alone (it may or may not come back for the gradient at this point) */
static void chain_request(chain_t *c, const double *x, int need_grad, int value_request)
{
	front_t *fr = c->fr;
	const int same = c->cvalid && memcmp(c->cx, x, sizeof(double) * (size_t)fr->nth1) == 0;
	if (same && need_grad && !value_request && c->chas_grad) c->cgrad_used = 1;
	if (same && (c->chas_grad || !need_grad)) return;
	int want_grad = need_grad;
	if (value_request) {
		const int policy = fr->opts->value_policy;
		c->n_value_requests++;
		if (policy == EMUB_VALUE_ALWAYS_GRADIENT) want_grad = 1;
		else if (policy == EMUB_VALUE_ONLY) want_grad = 0;
		else want_grad = c->n_value_requests <= 4 || 100 * c->n_followed >= 62 * (c->n_value_requests - 1);
	} else if (same) {
		/* value-only point the optimiser now wants the gradient of: evaluated twice */
		c->repeated++;
	}
	if (c->cvalid && c->chas_grad && !c->cgrad_used) c->unused_grads++; /* the cached gradient is about to be dropped unread */
	pthread_mutex_lock(&fr->mu);
	memcpy(c->x, x, sizeof(double) * (size_t)fr->nth1);
	c->want_grad = want_grad;
	c->state = REQ_PENDING;
	fr->npending++;
	if (fr->npending >= fr->nactive) pthread_cond_signal(&fr->cv_disp);
	while (c->state != REQ_DONE) pthread_cond_wait(&fr->cv_done, &fr->mu);
	c->state = REQ_IDLE;
	pthread_mutex_unlock(&fr->mu);
	memcpy(c->cx, x, sizeof(double) * (size_t)fr->nth1);
	memcpy(c->cg, c->g, sizeof(double) * (size_t)fr->nth1);
	c->cf = c->f;
	c->csigma2 = c->sigma2;
	c->cstatus = c->status;
	c->cvalid = 1;
	c->chas_grad = want_grad;
	c->cgrad_used = need_grad && !value_request;
	if (!want_grad) c->value_evals++;
}

/* evalFnMulti / gradFnMulti / evalFnGradMulti as the optimiser sees them (maxmultimin.c:288, :416, :615) */
static double cb_f(const double *x, void *ctx)
{
	chain_t *c = (chain_t *)ctx;
	chain_request(c, x, 0, 1);
	return c->cf;
}
static void note_followed(chain_t *c, const double *x)
{
	/* a gradient request at the point of the last value request: the line search accepted that trial point */
	if (c->cvalid && !c->cgrad_used && memcmp(c->cx, x, sizeof(double) * (size_t)c->fr->nth1) == 0) c->n_followed++;
}
static void cb_df(const double *x, void *ctx, double *g)
{
	chain_t *c = (chain_t *)ctx;
	note_followed(c, x);
	chain_request(c, x, 1, 0);
	memcpy(g, c->cg, sizeof(double) * (size_t)c->fr->nth1);
}
static void cb_fdf(const double *x, void *ctx, double *f, double *g)
{
	chain_t *c = (chain_t *)ctx;
	note_followed(c, x);
	chain_request(c, x, 1, 0);
	*f = c->cf;
	memcpy(g, c->cg, sizeof(double) * (size_t)c->fr->nth1);
}

/* one restart: doOptimizeMultiMin (maxmultimin.c:633-778) + scoring (maxmultimin.c:98-115) */
static void run_restart(chain_t *c, int try_index, emub_bfgs *bf)
{
	front_t *fr = c->fr;
	const emub_estimate_opts *o = fr->opts;
	const int nth = fr->nth, nth1 = fr->nth1;
	double *x_init = (double *)malloc(sizeof(double) * (size_t)nth);
	double *x_final = (double *)malloc(sizeof(double) * (size_t)nth1);
	if (fr->starts)
		memcpy(x_init, fr->starts + ((size_t)(fr->starts_per_comp ? c->comp * o->max_tries : 0) + (size_t)try_index) * nth,
		       sizeof(double) * (size_t)nth);
	else emub_random_init(c->seed, try_index, fr->ranges, nth, x_init);
	emub_bfgs_fn fn = {(size_t)nth1, cb_f, cb_df, cb_fdf, c};
	int status = emub_bfgs_set(bf, &fn, x_init + 1, o->step_size, o->tol); /* skip the amplitude, :665-668 */
	int stepcount = 0;
	do {
		status = emub_bfgs_iterate(bf);
		if (status == EMUB_BFGS_ENOPROG && stepcount > 0) break; /* :704-708 */
		status = emub_bfgs_test_gradient(emub_bfgs_gradient(bf), (size_t)nth1, o->eps_abs);
		stepcount++;
	} while (status == EMUB_BFGS_CONTINUE && stepcount < o->step_max);
	if (status == EMUB_BFGS_OK) c->success_count++;
	memcpy(x_final, emub_bfgs_x(bf), sizeof(double) * (size_t)nth1);
	/* sigma re-estimate and score at the final point: one evaluation gives both (:757, :103) */
	chain_request(c, x_final, 0, 0);
	const double likelihood = -1.0 * c->cf;
	if (c->cstatus == 0 && isfinite(likelihood) && isfinite(c->csigma2) && c->csigma2 > 0.0) {
		c->finite_count++;
		if (likelihood > c->best_lhood) {
			c->best_lhood = likelihood;
			c->best_thetas[0] = log(c->csigma2);
			memcpy(c->best_thetas + 1, x_final, sizeof(double) * (size_t)nth1);
		}
	}
	free(x_init);
	free(x_final);
}

static void *chain_main(void *arg)
{
	chain_t *c = (chain_t *)arg;
	front_t *fr = c->fr;
	emub_bfgs *bf = emub_bfgs_alloc((size_t)fr->nth1);
	for (int t = c->first_try; t < fr->opts->max_tries; t += c->try_stride) {
		c->cvalid = 0;
		run_restart(c, t, bf);
	}
	emub_bfgs_free(bf);
	pthread_mutex_lock(&fr->mu);
	fr->nactive--;
	pthread_cond_signal(&fr->cv_disp);
	pthread_mutex_unlock(&fr->mu);
	return NULL;
}

static int estimate_impl(emub_model *model, int ncomp, const double *ranges, const double *starts,
                         const emub_estimate_opts *opts_in, double *thetas_out, double *best_lhood,
                         emub_estimate_stats *stats);
static int estimate_front(emub_model *model, int ncomp, const double *ranges, const double *starts, int starts_per_comp,
                          const emub_estimate_opts *opts_in, double *thetas_out, double *best_lhood,
                          emub_estimate_stats *stats);

int emub_estimate_thetas(emub_model *model, const double *ranges, const emub_estimate_opts *opts_in,
                         double *thetas_out, double *best_lhood, emub_estimate_stats *stats)
{
	return emub_estimate_thetas_from(model, ranges, NULL, opts_in, thetas_out, best_lhood, stats);
}

int emub_estimate_thetas_from(emub_model *model, const double *ranges, const double *starts,
                              const emub_estimate_opts *opts_in, double *thetas_out, double *best_lhood,
                              emub_estimate_stats *stats)
{
	return estimate_impl(model, 1, ranges, starts, opts_in, thetas_out, best_lhood, stats);
}

int emub_estimate_thetas_multi(emub_model *model, int ncomp, const double *ranges, const emub_estimate_opts *opts_in,
                               double *thetas_out, double *best_lhood, emub_estimate_stats *stats)
{
	if (ncomp < 1 || ncomp > emub_model_ncomponents(model)) return EMUB_EINVAL;
	return estimate_impl(model, ncomp, ranges, NULL, opts_in, thetas_out, best_lhood, stats);
}

/* the restarts, then the optional refinement of every component's best point (one chain per component, one front) */
static int estimate_impl(emub_model *model, int ncomp, const double *ranges, const double *starts,
                         const emub_estimate_opts *opts_in, double *thetas_out, double *best_lhood,
                         emub_estimate_stats *stats)
{
	emub_estimate_opts o;
	if (opts_in) o = *opts_in; else emub_estimate_default_opts(&o);
	const int nth = emub_model_nthetas(model);
	double *best = (double *)malloc(sizeof(double) * (size_t)(ncomp > 0 ? ncomp : 1));
	int rc = estimate_front(model, ncomp, ranges, starts, 0, &o, thetas_out, best, stats);
	if ((rc == EMUB_OK || rc == EMUB_EDOM) && o.polish_steps > 0) {
		emub_estimate_opts po = o;
		po.max_tries = 1; po.nchains = 1; po.step_max = o.polish_steps; po.eps_abs = o.polish_eps;
		double *th2 = (double *)malloc(sizeof(double) * (size_t)ncomp * nth);
		double *best2 = (double *)malloc(sizeof(double) * (size_t)ncomp);
		emub_estimate_stats st2;
		/* the refinement follows the TRUE gradient of the objective (EMUB_GRAD_EXACT): the reference's formula is not the
		 * gradient of what evalFnMulti returns, so its zero is not the likelihood's maximum.  A component without any
		 * finite restart has nothing to refine: its start stays at zeros and is rejected again. */
		const int prev_mode = emub_model_gradient_mode(model);
		emub_model_set_gradient_mode(model, EMUB_GRAD_EXACT);
		int rc2 = estimate_front(model, ncomp, ranges, thetas_out, 1, &po, th2, best2, &st2);
		emub_model_set_gradient_mode(model, prev_mode);
		if (rc2 == EMUB_OK || rc2 == EMUB_EDOM) {
			for (int k = 0; k < ncomp; k++)
				if (best[k] != SCREWUPVALUE && best2[k] != SCREWUPVALUE && best2[k] > best[k]) {
					best[k] = best2[k];
					memcpy(thetas_out + (size_t)k * nth, th2 + (size_t)k * nth, sizeof(double) * (size_t)nth);
				}
			if (stats) {
				stats->evaluations += st2.evaluations; stats->batches += st2.batches;
				stats->value_evaluations += st2.value_evaluations; stats->repeated_points += st2.repeated_points;
				stats->unused_gradients += st2.unused_gradients;
			}
		} else
			rc = rc2;
		free(th2); free(best2);
	}
	/* Matern kernels: hand the result over in the convention the covariance FUNCTIONS read (emulator.c:355-356,
	 * :448-449: amplitude and nugget raw, log rho), which is what emub_emulator_create / alloc_emulator_struct, a
	 * MODEL_SNAPSHOT_FILE and interactive_mode consume.  The chains work on (log sigma^2 | log nugget, log rho)
	 * (deviation D-2: unit amplitude and e^theta_1 nugget during training, like the power-exponential kernel), so
	 * amp = sigma^2 and nugget = e^theta_1 -- the same C = sigma^2 c + e^theta_1 delta the power-exponential
	 * emulator builds from (log sigma^2, theta_1, ..) (emulator_struct.c:28, emulator.c:116-117). */
	if (emub_model_kernel(model) != EMUB_POWEREXP)
		for (int k = 0; k < ncomp; k++)
			if (best[k] != SCREWUPVALUE) {
				thetas_out[(size_t)k * nth] = exp(thetas_out[(size_t)k * nth]);
				thetas_out[(size_t)k * nth + 1] = exp(thetas_out[(size_t)k * nth + 1]);
			}
	if (best_lhood) memcpy(best_lhood, best, sizeof(double) * (size_t)ncomp);
	free(best);
	return rc;
}

/* ncomp components x nchains chains each, all in one evaluation front */
static int estimate_front(emub_model *model, int ncomp, const double *ranges, const double *starts, int starts_per_comp,
                          const emub_estimate_opts *opts_in, double *thetas_out, double *best_lhood,
                          emub_estimate_stats *stats)
{
	if (!model || (!ranges && !starts) || !thetas_out) return EMUB_EINVAL;
	emub_estimate_opts o;
	if (opts_in) o = *opts_in; else emub_estimate_default_opts(&o);
	if (o.max_tries < 1) o.max_tries = 1;
	int per_comp = o.nchains > 0 ? o.nchains : (o.max_tries < 64 ? o.max_tries : 64);
	if (per_comp > o.max_tries) per_comp = o.max_tries;
	int nchains = per_comp * ncomp;
	front_t fr;
	memset(&fr, 0, sizeof(fr));
	fr.model = model; fr.opts = &o; fr.ranges = ranges; fr.starts = starts; fr.starts_per_comp = starts_per_comp;
	fr.nth = emub_model_nthetas(model); fr.nth1 = fr.nth - 1; fr.nchains = nchains;
	pthread_mutex_init(&fr.mu, NULL);
	pthread_cond_init(&fr.cv_disp, NULL);
	pthread_cond_init(&fr.cv_done, NULL);
	fr.nactive = nchains;
	fr.chains = (chain_t *)calloc((size_t)nchains, sizeof(chain_t));
	const size_t vb = sizeof(double) * (size_t)fr.nth;
	for (int i = 0; i < nchains; i++) {
		chain_t *c = &fr.chains[i];
		c->fr = &fr; c->id = i; c->best_lhood = SCREWUPVALUE;
		c->comp = i / per_comp; c->first_try = i % per_comp; c->try_stride = per_comp;
		c->seed = o.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(o.first_component + c->comp * (o.component_stride > 0 ? o.component_stride : 1));
		c->x = (double *)calloc(1, vb); c->g = (double *)calloc(1, vb);
		c->cx = (double *)calloc(1, vb); c->cg = (double *)calloc(1, vb);
		c->best_thetas = (double *)calloc(1, vb);
	}
	pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nchains);
	int nstarted = 0, thread_failed = 0;
	{
		/* a chain parks in the evaluation front until every live chain has asked, so all of them need a thread of their
		 * own; if one cannot be created the front is shrunk to the chains that run (their restarts are then incomplete
		 * and the call reports the failure) */
		pthread_attr_t at;
		pthread_attr_init(&at);
		pthread_attr_setstacksize(&at, 1 << 20);
		for (int i = 0; i < nchains; i++) {
			if (pthread_create(&tids[i], &at, chain_main, &fr.chains[i]) != 0) { thread_failed = 1; break; }
			nstarted++;
		}
		pthread_attr_destroy(&at);
		if (thread_failed) {
			pthread_mutex_lock(&fr.mu);
			fr.nactive -= nchains - nstarted;
			pthread_cond_signal(&fr.cv_disp);
			pthread_mutex_unlock(&fr.mu);
		}
	}

	/* dispatcher: whenever every live chain is waiting, evaluate the whole front in one GPU call */
	double *bx = (double *)malloc(sizeof(double) * (size_t)nchains * fr.nth1);
	double *bg = (double *)malloc(sizeof(double) * (size_t)nchains * fr.nth1);
	double *bf = (double *)malloc(sizeof(double) * (size_t)nchains);
	double *bs = (double *)malloc(sizeof(double) * (size_t)nchains);
	int *bst = (int *)malloc(sizeof(int) * (size_t)nchains);
	int *who = (int *)malloc(sizeof(int) * (size_t)nchains);
	int *bcomp = (int *)malloc(sizeof(int) * (size_t)nchains);
	int *bwant = (int *)malloc(sizeof(int) * (size_t)nchains);
	int rc = EMUB_OK;
	FILE *front_log = getenv("EMUB_FRONT_LOG") ? fopen(getenv("EMUB_FRONT_LOG"), "a") : NULL;
	double *log_buf = NULL;
	size_t log_len = 0, log_cap = 0;
	pthread_mutex_lock(&fr.mu);
	for (;;) {
		while (fr.nactive > 0 && fr.npending < fr.nactive) pthread_cond_wait(&fr.cv_disp, &fr.mu);
		if (fr.nactive == 0) break;
		/* the whole front in ONE batched call; every point says whether it needs the gradient (the value-only points
		 * skip the inverse, emub_loglik_grad_batch_mixed) */
		int B = 0;
		for (int i = 0; i < nchains; i++)
			if (fr.chains[i].state == REQ_PENDING) {
				memcpy(bx + (size_t)B * fr.nth1, fr.chains[i].x, sizeof(double) * (size_t)fr.nth1);
				bcomp[B] = fr.chains[i].comp;
				bwant[B] = fr.chains[i].want_grad;
				who[B++] = i;
			}
		pthread_mutex_unlock(&fr.mu);
		int call = emub_loglik_grad_batch_mixed(model, bx, bcomp, bwant, B, bf, bg, bs, bst);
		fr.batches++;
		if (front_log) { /* EMUB_FRONT_LOG (debugging aid): every batched call kept in memory, written as text when the front ends */
			const size_t rec = 1 + (size_t)B * (size_t)(fr.nth1 + 5);
			if (log_len + rec > log_cap) { log_cap = (log_len + rec) * 2; log_buf = (double *)realloc(log_buf, sizeof(double) * log_cap); }
			log_buf[log_len++] = (double)B;
			for (int k = 0; k < B; k++) {
				log_buf[log_len++] = (double)bcomp[k]; log_buf[log_len++] = (double)bwant[k];
				for (int j = 0; j < fr.nth1; j++) log_buf[log_len++] = bx[(size_t)k * fr.nth1 + j];
				log_buf[log_len++] = bf[k]; log_buf[log_len++] = bs[k]; log_buf[log_len++] = (double)bst[k];
			}
		}
		pthread_mutex_lock(&fr.mu);
		if (call != EMUB_OK) { rc = call; fr.failed = 1; }
		fr.evaluations += B;
		for (int k = 0; k < B; k++) {
			chain_t *c = &fr.chains[who[k]];
			if (call == EMUB_OK) {
				c->f = bf[k]; c->sigma2 = bs[k]; c->status = bst[k];
				memcpy(c->g, bg + (size_t)k * fr.nth1, sizeof(double) * (size_t)fr.nth1);
			} else {
				c->f = NAN; c->sigma2 = NAN; c->status = EMUB_ECUDA;
				for (int j = 0; j < fr.nth1; j++) c->g[j] = NAN;
			}
			c->state = REQ_DONE;
		}
		fr.npending = 0;
		pthread_cond_broadcast(&fr.cv_done);
	}
	pthread_mutex_unlock(&fr.mu);
	if (front_log) {
		for (size_t i = 0; i < log_len;) {
			const int B = (int)log_buf[i++];
			fprintf(front_log, "%d\n", B);
			for (int k = 0; k < B; k++) {
				fprintf(front_log, "%d %d", (int)log_buf[i], (int)log_buf[i + 1]);
				i += 2;
				for (int j = 0; j < fr.nth1; j++) fprintf(front_log, " %.17g", log_buf[i++]);
				fprintf(front_log, " | %.17g %.17g %d\n", log_buf[i], log_buf[i + 1], (int)log_buf[i + 2]);
				i += 3;
			}
		}
		fclose(front_log);
		free(log_buf);
	}
	for (int i = 0; i < nstarted; i++) pthread_join(tids[i], NULL);
	if (thread_failed && rc == EMUB_OK) rc = EMUB_ENOMEM;

	/* best per component over its chains, in chain order (deterministic) -- estimate_threaded.c:294-323 */
	int succ = 0, fin = 0, nfailed_comp = 0;
	for (int k = 0; k < ncomp; k++) {
		double best = SCREWUPVALUE;
		for (int i = k * per_comp; i < (k + 1) * per_comp; i++) {
			chain_t *c = &fr.chains[i];
			succ += c->success_count;
			fin += c->finite_count;
			if (c->best_lhood > best) {
				best = c->best_lhood;
				memcpy(thetas_out + (size_t)k * fr.nth, c->best_thetas, vb);
			}
		}
		if (best_lhood) best_lhood[k] = best;
		if (best == SCREWUPVALUE) {
			nfailed_comp++;
			for (int i = 0; i < fr.nth; i++) thetas_out[(size_t)k * fr.nth + i] = 0.0;
		}
	}
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->evaluations = fr.evaluations; stats->batches = fr.batches; stats->success_count = succ; stats->finite_count = fin;
		for (int i = 0; i < nchains; i++) {
			stats->value_evaluations += fr.chains[i].value_evals;
			stats->repeated_points += fr.chains[i].repeated;
			stats->unused_gradients += fr.chains[i].unused_grads;
		}
	}
	for (int i = 0; i < nchains; i++) {
		chain_t *c = &fr.chains[i];
		free(c->x); free(c->g); free(c->cx); free(c->cg); free(c->best_thetas);
	}
	free(fr.chains); free(tids); free(bx); free(bg); free(bf); free(bs); free(bst); free(who); free(bcomp); free(bwant);
	pthread_mutex_destroy(&fr.mu); pthread_cond_destroy(&fr.cv_disp); pthread_cond_destroy(&fr.cv_done);
	if (rc != EMUB_OK) return rc;
	return nfailed_comp ? EMUB_EDOM : EMUB_OK;
}

/* ---- components sharded over devices ------------------------------------------------------------------ */
typedef struct {
	int device, g, ndev;
	const double *X, *Z;
	int ldx, n, d, ldz, ncomp, kernel, order, max_slots;
	emub_estimate_opts opts;
	const double *ranges_in; /* NULL: emub_optimization_ranges of the design */
	double *thetas_out, *best;
	emub_estimate_stats stats;
	int rc;
} dev_job;

static void *dev_main(void *arg)
{
	dev_job *j = (dev_job *)arg;
	int nloc = 0;
	for (int c = j->g; c < j->ncomp; c += j->ndev) nloc++;
	if (nloc == 0) { j->rc = EMUB_OK; return NULL; }
	const int nth = (j->kernel == EMUB_POWEREXP) ? j->d + 2 : 3;
	emub_ctx *ctx = NULL;
	emub_model *m = NULL;
	double *Y = (double *)malloc(sizeof(double) * (size_t)j->n * nloc);
	for (int i = 0; i < j->n; i++)
		for (int k = 0; k < nloc; k++) Y[(size_t)i * nloc + k] = j->Z[(size_t)i * j->ldz + j->g + k * j->ndev];
	double *ranges = (double *)malloc(sizeof(double) * 2 * (size_t)nth);
	double *th = (double *)calloc((size_t)nloc * nth, sizeof(double));
	double *best = (double *)calloc((size_t)nloc, sizeof(double));
	if (j->ranges_in) memcpy(ranges, j->ranges_in, sizeof(double) * 2 * (size_t)nth);
	else emub_optimization_ranges(j->kernel, j->X, j->ldx, j->n, j->d, ranges);
	j->rc = emub_ctx_create(j->device, &ctx);
	if (j->rc == EMUB_OK) j->rc = emub_model_create(ctx, j->X, j->ldx, j->n, j->d, Y, j->kernel, j->order, j->max_slots, &m);
	if (j->rc == EMUB_OK) j->rc = emub_model_set_training_multi(m, Y, nloc, nloc);
	if (j->rc == EMUB_OK) {
		j->opts.first_component = j->g;
		j->opts.component_stride = j->ndev;
		j->rc = emub_estimate_thetas_multi(m, nloc, ranges, &j->opts, th, best, &j->stats);
		for (int k = 0; k < nloc; k++) {
			memcpy(j->thetas_out + (size_t)(j->g + k * j->ndev) * nth, th + (size_t)k * nth, sizeof(double) * (size_t)nth);
			j->best[j->g + k * j->ndev] = best[k];
		}
	}
	if (m) emub_model_destroy(m);
	if (ctx) emub_ctx_destroy(ctx);
	free(Y); free(ranges); free(th); free(best);
	return NULL;
}

int emub_estimate_thetas_multi_devices(const int *devices, int ndev, const double *X, int ldx, int n, int d,
                                       const double *Z, int ldz, int ncomp, int kernel, int regression_order,
                                       int max_slots, const emub_estimate_opts *opts_in, double *thetas_out,
                                       double *best_lhood, emub_estimate_stats *stats)
{
	return emub_estimate_thetas_multi_devices_ranges(devices, ndev, X, ldx, n, d, Z, ldz, ncomp, kernel, regression_order, max_slots,
	                                                 NULL, opts_in, thetas_out, best_lhood, stats);
}

int emub_estimate_thetas_multi_devices_ranges(const int *devices, int ndev, const double *X, int ldx, int n, int d,
                                              const double *Z, int ldz, int ncomp, int kernel, int regression_order,
                                              int max_slots, const double *ranges, const emub_estimate_opts *opts_in,
                                              double *thetas_out, double *best_lhood, emub_estimate_stats *stats)
{
	if (!devices || ndev < 1 || ndev > 64 || !X || !Z || ncomp < 1 || !thetas_out || !best_lhood) return EMUB_EINVAL;
	dev_job jobs[64];
	pthread_t th[64];
	int started[64];
	emub_estimate_opts o;
	if (opts_in) o = *opts_in; else emub_estimate_default_opts(&o);
	for (int g = 0; g < ndev; g++) {
		dev_job *j = &jobs[g];
		memset(j, 0, sizeof(*j));
		j->device = devices[g]; j->g = g; j->ndev = ndev; j->X = X; j->Z = Z; j->ldx = ldx; j->n = n; j->d = d; j->ldz = ldz;
		j->ncomp = ncomp; j->kernel = kernel; j->order = regression_order; j->max_slots = max_slots; j->opts = o;
		j->thetas_out = thetas_out; j->best = best_lhood; j->ranges_in = ranges;
		started[g] = g > 0 && pthread_create(&th[g], NULL, dev_main, j) == 0;
	}
	int rc = EMUB_OK;
	if (stats) memset(stats, 0, sizeof(*stats));
	for (int g = 0; g < ndev; g++) {
		if (started[g]) pthread_join(th[g], NULL);
		else dev_main(&jobs[g]); /* device 0, and any device whose thread could not be created, on the caller's thread */
		if (jobs[g].rc != EMUB_OK && rc == EMUB_OK) rc = jobs[g].rc;
		if (stats) {
			stats->evaluations += jobs[g].stats.evaluations; stats->batches += jobs[g].stats.batches;
			stats->success_count += jobs[g].stats.success_count; stats->finite_count += jobs[g].stats.finite_count;
			stats->value_evaluations += jobs[g].stats.value_evaluations; stats->repeated_points += jobs[g].stats.repeated_points;
			stats->unused_gradients += jobs[g].stats.unused_gradients;
		}
	}
	return rc;
}
