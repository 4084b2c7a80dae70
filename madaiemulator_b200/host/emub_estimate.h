/*
 * emub_estimate.h -- host-side (plain C) hyper-parameter estimation over the batched GPU evaluator.
 *
 * Replaces the reference's restart machinery for the hot path: estimate_thetas_threaded
 * (src/libEmu/estimate_threaded.c:78) -> maxWithMultiMin (src/libEmu/maxmultimin.c:47) ->
 * doOptimizeMultiMin (:633).  The optimiser stays on the host exactly as in the reference -- one BFGS
 * chain per pthread, the reference's own parallel model (estimate_threaded.c:172) -- but no chain ever
 * evaluates a likelihood itself: every evalFnMulti / gradFnMulti request is parked in a queue, and when
 * all live chains are waiting the dispatcher issues ONE emub_loglik_grad_batch call for the whole front
 * (include/emu_b200.h).  Results do not depend on thread timing: start points come from a counter-based
 * generator keyed by (seed, try index) and the batched evaluator is bit-wise independent of the batch
 * composition.
 * Debugging aid: EMUB_FRONT_LOG=<file> appends every batched call of a front as text (B, then per point: component,
 * gradient flag, theta | -L, sigma^2, status), kept in memory and written when the front ends.
 */
#ifndef EMUB_ESTIMATE_H
#define EMUB_ESTIMATE_H
#include "../../include/emu_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int max_tries;            /* random restarts; reference: 50 per job x ncpus jobs (estimate_threaded.c:101,113) */
	int nchains;              /* concurrent BFGS chains = width of the evaluation front (0: min(max_tries, 64)) */
	unsigned long long seed;  /* reference: /dev/urandom (useful.c:49); here explicit and reproducible */
	double step_size;         /* 1.5   maxmultimin.c:644 */
	double tol;               /* 0.5   maxmultimin.c:645 */
	double eps_abs;           /* 0.1   maxmultimin.c:650 */
	int step_max;             /* 30    maxmultimin.c:641 */
	/* start-point stream of local component k: seed + (first_component + k * component_stride) * 0x9E3779B97F4A7C15,
	 * so that components keep their streams when they are sharded over devices (defaults 0 and 1) */
	int first_component, component_stride;
	/* Optional refinement, off by default (0): after the restarts, one more BFGS run per component from its best
	 * point, on the TRUE gradient of the objective (EMUB_GRAD_EXACT, include/emu_b200.h), with the tighter gradient
	 * tolerance polish_eps and at most polish_steps iterations; kept only if the likelihood improves.  The reference's
	 * gradient formula is not the gradient of its own objective and its chains stop at |g| < 0.1, so where a restart
	 * ends is partly chance; the extra run costs a handful of batched evaluations and can only raise the likelihood. */
	int polish_steps;
	double polish_eps;
	/* What a chain asks the GPU for when its optimiser wants the VALUE alone (evalFnMulti, maxmultimin.c:288 -- the
	 * trial points of the line search).  The value-only evaluation costs 0.38 n^3 flops against n^3 with the gradient
	 * (gradFnMulti, :416), but a trial point the line search accepts is asked for its gradient right away, and then
	 * costs 1.38 n^3.  EMUB_VALUE_ADAPTIVE (default): per chain, value-only while fewer than 62% of its value requests
	 * were followed by a gradient request at the same point, otherwise the gradient speculatively.  The optimiser sees
	 * the same bits under every policy (the two evaluations return identical -L), so the result does not depend on it. */
	int value_policy;
} emub_estimate_opts;
enum { EMUB_VALUE_ADAPTIVE = 0, EMUB_VALUE_ALWAYS_GRADIENT = 1, EMUB_VALUE_ONLY = 2 };

typedef struct {
	long long evaluations;    /* likelihood(+gradient) points evaluated on the GPU */
	long long batches;        /* emub_loglik_grad_batch calls */
	int success_count;        /* chains that stopped on |g| < eps_abs (maxmultimin.c:91-92) */
	int finite_count;         /* restarts with a finite final likelihood (maxmultimin.c:110) */
	long long value_evaluations;  /* of `evaluations`: value-only points (no triangular inverse, no C^-1, no gradient) */
	long long repeated_points;    /* value-only points evaluated again with the gradient (the line search accepted them) */
	long long unused_gradients;   /* gradients computed speculatively and never read by the optimiser */
} emub_estimate_stats;

void emub_estimate_default_opts(emub_estimate_opts *o);

/* sample scales (modelstruct.c:188-213) and optimiser search ranges (optstruct.c:142-226, use_data_scales = 1,
 * no fixed nugget); ranges is nthetas x 2 row-major */
void emub_sample_scales(const double *X, int ldx, int n, int d, double *scales);
void emub_optimization_ranges(int kernel, const double *X, int ldx, int n, int d, double *ranges);
/* the same with the two switches of the optstruct: use_data_scales = 0 keeps the default length ranges (optstruct.c:205-211),
 * fixed_nugget_mode = 1 caps the nugget range at fixed_nugget + 20% (optstruct.c:217-225; callEstimate's use_fixed_nugget,
 * rbind.c:43-60) */
void emub_optimization_ranges_ex(int kernel, const double *X, int ldx, int n, int d, int use_data_scales, int fixed_nugget_mode,
                                 double fixed_nugget, double *ranges);

/* start point of restart `try_index`: uniform in ranges (set_random_init_value, maxmultimin.c:789-804) */
void emub_random_init(unsigned long long seed, int try_index, const double *ranges, int nthetas, double *x);

/*
 * maxWithMultiMin over the batched evaluator.  thetas_out (nthetas): best point, thetas_out[0] =
 * log(sigma^2) re-estimated at the optimum (maxmultimin.c:757-769).  best_lhood: its log-likelihood
 * (maxmultimin.c:103).  Returns EMUB_OK, or EMUB_EDOM when no restart produced a finite likelihood
 * ("maximisation didn't work at all", maxmultimin.c:121-123).
 * thetas_out is always in the convention emub_emulator_create / the covariance functions read:
 * power-exponential (log sigma^2, log nugget, log lengths); Matern (sigma^2, nugget, log rho) -- amplitude and
 * nugget RAW (emulator.c:355-356, :448-449), converted from the chain's log-scale working point.
 */
int emub_estimate_thetas(emub_model *model, const double *ranges, const emub_estimate_opts *opts,
                         double *thetas_out, double *best_lhood, emub_estimate_stats *stats);

/* same, but restart t starts from starts[t * nthetas ..] (max_tries rows; entry 0 of each row, the amplitude, is
 * ignored as in maxmultimin.c:665-668) instead of a random point; ranges may then be NULL */
int emub_estimate_thetas_from(emub_model *model, const double *ranges, const double *starts,
                              const emub_estimate_opts *opts, double *thetas_out, double *best_lhood,
                              emub_estimate_stats *stats);

/*
 * estimate_multi (multivar_support.c:20-27) without its serial loop over the PCA components: the model carries ncomp
 * training vectors (emub_model_set_training_multi) and the restart chains of ALL components feed one evaluation
 * front.  opts->max_tries and opts->nchains are per component.  thetas_out: ncomp x nthetas, best_lhood: ncomp.
 * Component k uses the start-point stream seed + k * 0x9E3779B97F4A7C15, so its result equals a single-component run
 * with that seed.
 */
int emub_estimate_thetas_multi(emub_model *model, int ncomp, const double *ranges, const emub_estimate_opts *opts,
                               double *thetas_out, double *best_lhood, emub_estimate_stats *stats);

/*
 * The same over several GPUs of one box: PCA components are independent scalar GPs on a shared design,
 * component c goes to device devices[c % ndev]; one host thread, context and model per device, no exchange
 * between devices, the thetas are gathered on the host.  Z is the n x ncomp training matrix (row stride ldz).
 * Results are identical to a single-device emub_estimate_thetas_multi with the same seed.
 */
int emub_estimate_thetas_multi_devices(const int *devices, int ndev, const double *X, int ldx, int n, int d,
                                       const double *Z, int ldz, int ncomp, int kernel, int regression_order,
                                       int max_slots, const emub_estimate_opts *opts, double *thetas_out,
                                       double *best_lhood, emub_estimate_stats *stats);

/* the same with explicit optimisation ranges (nthetas x 2, e.g. a reference optstruct's grad_ranges); NULL = the
 * ranges emub_optimization_ranges derives from the design */
int emub_estimate_thetas_multi_devices_ranges(const int *devices, int ndev, const double *X, int ldx, int n, int d,
                                              const double *Z, int ldz, int ncomp, int kernel, int regression_order,
                                              int max_slots, const double *ranges, const emub_estimate_opts *opts,
                                              double *thetas_out, double *best_lhood, emub_estimate_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
