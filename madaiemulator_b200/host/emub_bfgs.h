/*
 * emub_bfgs.h -- memoryless ("vector") BFGS minimiser with a Fletcher bracketing /
 * sectioning line search (cubic interpolation), plain C, callback driven.
 *
 * Role: the reference drives its likelihood maximisation with GSL's
 * gsl_multimin_fdfminimizer_vector_bfgs2 (reference src/libEmu/maxmultimin.c:683-735,
 * step 1.5, tol 0.5, |g| < 0.1, <= 30 iterations).  GSL is an un-vendored third-party
 * dependency that is absent here, so this file restates the published algorithm
 * (R. Fletcher, "Practical Methods of Optimization", 2nd ed., algorithms 2.6.2 and 2.6.4,
 * plus the one-step BFGS direction update) in our own code.  It is used by the host
 * restart driver (emub_estimate.c) and, through oracle/gsl_shim, it also stands in for
 * gsl_multimin when the reference's own maxmultimin.c is compiled as the CPU oracle.
 */
#ifndef EMUB_BFGS_H
#define EMUB_BFGS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	size_t n;
	double (*f)(const double *x, void *ctx);
	void (*df)(const double *x, void *ctx, double *g);
	void (*fdf)(const double *x, void *ctx, double *f, double *g);
	void *ctx;
} emub_bfgs_fn;

enum { EMUB_BFGS_OK = 0, EMUB_BFGS_CONTINUE = -2, EMUB_BFGS_ENOPROG = 27, EMUB_BFGS_EINVAL = 4 };

typedef struct emub_bfgs emub_bfgs;

emub_bfgs *emub_bfgs_alloc(size_t n);
void emub_bfgs_free(emub_bfgs *s);
/* evaluates f and g at x0 */
int emub_bfgs_set(emub_bfgs *s, const emub_bfgs_fn *fn, const double *x0, double step_size, double tol);
/* one quasi-Newton iteration (one line search); EMUB_BFGS_ENOPROG when stuck */
int emub_bfgs_iterate(emub_bfgs *s);
const double *emub_bfgs_x(const emub_bfgs *s);
const double *emub_bfgs_gradient(const emub_bfgs *s);
const double *emub_bfgs_dx(const emub_bfgs *s);
double emub_bfgs_minimum(const emub_bfgs *s);
/* EMUB_BFGS_OK if |g|_2 < epsabs else EMUB_BFGS_CONTINUE */
int emub_bfgs_test_gradient(const double *g, size_t n, double epsabs);

#ifdef __cplusplus
}
#endif
#endif
