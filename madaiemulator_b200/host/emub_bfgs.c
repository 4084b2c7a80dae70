/*
 * emub_bfgs.c -- see emub_bfgs.h.  Memoryless BFGS + Fletcher line search.
 *
 * Constants of the line search follow Fletcher's recommendations as used by the optimiser
 * the reference selects (maxmultimin.c:683): rho = 0.01 (sufficient decrease), sigma = tol
 * (curvature), tau1 = 9 (bracket extrapolation), tau2 = 0.05 / tau3 = 0.5 (section
 * shrinkage), cubic interpolation.
 */
#include "emub_bfgs.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

struct emub_bfgs {
	size_t n;
	emub_bfgs_fn fn;
	int iter;
	double step, tol;
	double f;            /* current minimum value */
	double *x, *g, *dx;  /* current point, gradient, last step */
	double *x0, *g0, *p; /* line-search origin, its gradient, unit search direction */
	double *dx0, *dg0;
	double g0norm, pnorm, fp0, delta_f;
	/* line-function cache: phi(alpha) = f(x0 + alpha p) */
	double *x_alpha, *g_alpha;
	double f_alpha, df_alpha;
	double f_cache_alpha, g_cache_alpha, x_cache_alpha; /* NAN = invalid */
};

static double dot(const double *a, const double *b, size_t n)
{
	double s = 0.0;
	for (size_t i = 0; i < n; i++) s += a[i] * b[i];
	return s;
}
static double nrm2(const double *a, size_t n) { return sqrt(dot(a, a, n)); }

emub_bfgs *emub_bfgs_alloc(size_t n)
{
	emub_bfgs *s = (emub_bfgs *)calloc(1, sizeof(*s));
	if (!s) return NULL;
	s->n = n;
	double *blk = (double *)calloc(10 * n, sizeof(double));
	if (!blk) { free(s); return NULL; }
	s->x = blk; s->g = blk + n; s->dx = blk + 2 * n; s->x0 = blk + 3 * n; s->g0 = blk + 4 * n;
	s->p = blk + 5 * n; s->dx0 = blk + 6 * n; s->dg0 = blk + 7 * n; s->x_alpha = blk + 8 * n;
	s->g_alpha = blk + 9 * n;
	return s;
}

void emub_bfgs_free(emub_bfgs *s)
{
	if (!s) return;
	free(s->x);
	free(s);
}

/* ---- line function with caching ------------------------------------------------------- */
static void move_to(emub_bfgs *s, double alpha)
{
	if (alpha == s->x_cache_alpha) return;
	for (size_t i = 0; i < s->n; i++) s->x_alpha[i] = s->x0[i] + alpha * s->p[i];
	s->x_cache_alpha = alpha;
}
static double phi_f(emub_bfgs *s, double alpha)
{
	if (alpha == s->f_cache_alpha) return s->f_alpha;
	move_to(s, alpha);
	s->f_alpha = s->fn.f(s->x_alpha, s->fn.ctx);
	s->f_cache_alpha = alpha;
	return s->f_alpha;
}
static double phi_df(emub_bfgs *s, double alpha)
{
	if (alpha == s->g_cache_alpha) return s->df_alpha;
	move_to(s, alpha);
	s->fn.df(s->x_alpha, s->fn.ctx, s->g_alpha);
	s->df_alpha = dot(s->g_alpha, s->p, s->n);
	s->g_cache_alpha = alpha;
	return s->df_alpha;
}
static void phi_update_position(emub_bfgs *s, double alpha, double *x, double *f, double *g)
{
	if (alpha == s->f_cache_alpha && alpha == s->g_cache_alpha) {
		/* both cached */
	} else {
		move_to(s, alpha);
		s->fn.fdf(s->x_alpha, s->fn.ctx, &s->f_alpha, s->g_alpha);
		s->f_cache_alpha = alpha;
		s->g_cache_alpha = alpha;
		s->df_alpha = dot(s->g_alpha, s->p, s->n);
	}
	*f = s->f_alpha;
	memcpy(x, s->x_alpha, s->n * sizeof(double));
	memcpy(g, s->g_alpha, s->n * sizeof(double));
}
static void phi_change_direction(emub_bfgs *s)
{
	/* the new origin is the current point: keep its f and g as alpha = 0 */
	memcpy(s->x_alpha, s->x0, s->n * sizeof(double));
	s->x_cache_alpha = 0.0;
	s->f_alpha = s->f;
	s->f_cache_alpha = 0.0;
	memcpy(s->g_alpha, s->g0, s->n * sizeof(double));
	s->df_alpha = dot(s->g_alpha, s->p, s->n);
	s->g_cache_alpha = 0.0;
}

/* ---- interpolation on the normalised interval ------------------------------------------- */
/* minimum over [zl, zh] of q(z) = f0 + fp0 z + (f1 - f0 - fp0) z^2 */
static double interp_quad(double f0, double fp0, double f1, double zl, double zh)
{
	double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
	double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
	double c = 2.0 * (f1 - f0 - fp0); /* curvature */
	double zmin = zl, fmin = fl;
	if (fh < fmin) { zmin = zh; fmin = fh; }
	if (c > 0.0) {
		double z = -fp0 / c;
		if (z > zl && z < zh) {
			double fz = f0 + z * (fp0 + z * (f1 - f0 - fp0));
			if (fz < fmin) { zmin = z; fmin = fz; }
		}
	}
	return zmin;
}
static double cubic_val(double c0, double c1, double c2, double c3, double z)
{
	return c0 + z * (c1 + z * (c2 + z * c3));
}
static void check_extremum(double c0, double c1, double c2, double c3, double z, double *zmin, double *fmin)
{
	double y = cubic_val(c0, c1, c2, c3, z);
	if (y < *fmin) { *zmin = z; *fmin = y; }
}
/* minimum over [zl, zh] of the Hermite cubic through (0,f0,fp0), (1,f1,fp1) */
static double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh)
{
	double eta = 3.0 * (f1 - f0) - 2.0 * fp0 - fp1;
	double xi = fp0 + fp1 - 2.0 * (f1 - f0);
	double c0 = f0, c1 = fp0, c2 = eta, c3 = xi;
	double zmin = zl, fmin = cubic_val(c0, c1, c2, c3, zl);
	check_extremum(c0, c1, c2, c3, zh, &zmin, &fmin);
	/* stationary points of c1 + 2 c2 z + 3 c3 z^2 */
	double a = 3.0 * c3, b = 2.0 * c2, c = c1;
	if (a == 0.0) {
		if (b != 0.0) {
			double z = -c / b;
			if (b > 0.0 && z > zl && z < zh) check_extremum(c0, c1, c2, c3, z, &zmin, &fmin);
		}
	} else {
		double disc = b * b - 4.0 * a * c;
		if (disc > 0.0) {
			double sq = sqrt(disc);
			double q = -0.5 * (b + (b > 0 ? sq : -sq));
			double z0 = q / a, z1 = (q != 0.0) ? c / q : z0;
			/* the minimum is the root with positive second derivative b + 2 a z */
			if (b + 2.0 * a * z0 > 0.0 && z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
			if (b + 2.0 * a * z1 > 0.0 && z1 > zl && z1 < zh) check_extremum(c0, c1, c2, c3, z1, &zmin, &fmin);
		} else if (disc == 0.0) {
			/* inflection: no interior minimum */
		}
	}
	return zmin;
}
static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb,
                          double xmin, double xmax)
{
	double zl = (xmin - a) / (b - a), zh = (xmax - a) / (b - a), z;
	if (zl > zh) { double t = zl; zl = zh; zh = t; }
	if (isfinite(fpb))
		z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zl, zh);
	else
		z = interp_quad(fa, fpa * (b - a), fb, zl, zh);
	return a + z * (b - a);
}

/* Fletcher's line search: find alpha satisfying the strong Wolfe conditions
 *   phi(alpha) <= phi(0) + rho alpha phi'(0),  |phi'(alpha)| <= -sigma phi'(0). */
static int line_minimize(emub_bfgs *s, double rho, double sigma, double tau1, double tau2, double tau3,
                         double alpha1, double *alpha_new)
{
	double f0 = phi_f(s, 0.0), fp0 = phi_df(s, 0.0);
	double falpha, fpalpha, delta, alpha_next;
	double alpha = alpha1, alpha_prev = 0.0;
	double a = 0.0, b = alpha, fa = f0, fb = 0.0, fpa = fp0, fpb = 0.0;
	double falpha_prev = f0, fpalpha_prev = fp0;
	const int bracket_iters = 100, section_iters = 100;
	int i = 0;

	/* bracketing */
	while (i++ < bracket_iters) {
		falpha = phi_f(s, alpha);
		if (falpha > f0 + alpha * rho * fp0 || falpha >= falpha_prev) {
			a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
			b = alpha; fb = falpha; fpb = NAN;
			break;
		}
		fpalpha = phi_df(s, alpha);
		if (fabs(fpalpha) <= -sigma * fp0) { *alpha_new = alpha; return EMUB_BFGS_OK; }
		if (fpalpha >= 0.0) {
			a = alpha; fa = falpha; fpa = fpalpha;
			b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
			break;
		}
		delta = alpha - alpha_prev;
		{
			double lower = alpha + delta, upper = alpha + tau1 * delta;
			alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, lower, upper);
		}
		alpha_prev = alpha; falpha_prev = falpha; fpalpha_prev = fpalpha;
		alpha = alpha_next;
	}
	/* sectioning */
	while (i++ < section_iters) {
		delta = b - a;
		{
			double lower = a + tau2 * delta, upper = b - tau3 * delta;
			alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper);
		}
		falpha = phi_f(s, alpha);
		if ((a - alpha) * fpa <= DBL_EPSILON) {
			/* round-off prevents further progress */
			return EMUB_BFGS_ENOPROG;
		}
		if (falpha > f0 + rho * alpha * fp0 || falpha >= fa) {
			b = alpha; fb = falpha; fpb = NAN;
		} else {
			fpalpha = phi_df(s, alpha);
			if (fabs(fpalpha) <= -sigma * fp0) { *alpha_new = alpha; return EMUB_BFGS_OK; }
			if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
				b = a; fb = fa; fpb = fpa;
				a = alpha; fa = falpha; fpa = fpalpha;
			} else {
				a = alpha; fa = falpha; fpa = fpalpha;
			}
		}
	}
	*alpha_new = alpha;
	return EMUB_BFGS_OK;
}

int emub_bfgs_set(emub_bfgs *s, const emub_bfgs_fn *fn, const double *x0, double step_size, double tol)
{
	size_t n = s->n;
	if (fn->n != n) return EMUB_BFGS_EINVAL;
	s->fn = *fn;
	s->iter = 0;
	s->step = step_size;
	s->tol = tol;
	s->delta_f = 0.0;
	memcpy(s->x, x0, n * sizeof(double));
	s->fn.fdf(s->x, s->fn.ctx, &s->f, s->g);
	memset(s->dx, 0, n * sizeof(double));
	memcpy(s->x0, s->x, n * sizeof(double));
	memcpy(s->g0, s->g, n * sizeof(double));
	s->g0norm = nrm2(s->g0, n);
	/* first direction: steepest descent, unit length */
	for (size_t i = 0; i < n; i++) s->p[i] = (s->g0norm > 0.0) ? -s->g[i] / s->g0norm : 0.0;
	s->pnorm = nrm2(s->p, n);
	s->fp0 = -s->g0norm;
	s->x_cache_alpha = s->f_cache_alpha = s->g_cache_alpha = NAN;
	phi_change_direction(s);
	return EMUB_BFGS_OK;
}

int emub_bfgs_iterate(emub_bfgs *s)
{
	size_t n = s->n;
	double alpha = 0.0, alpha1;
	double f0 = s->f;
	int status;

	if (s->pnorm == 0.0 || s->g0norm == 0.0 || s->fp0 == 0.0 || !isfinite(s->fp0) || !isfinite(f0)) {
		memset(s->dx, 0, n * sizeof(double));
		return EMUB_BFGS_ENOPROG;
	}
	if (s->delta_f < 0.0) {
		double del = fmax(-s->delta_f, 10.0 * DBL_EPSILON * fabs(f0));
		alpha1 = fmin(1.0, 2.0 * del / (-s->fp0));
	} else {
		alpha1 = fabs(s->step);
	}
	status = line_minimize(s, 0.01, s->tol, 9.0, 0.05, 0.5, alpha1, &alpha);
	if (status != EMUB_BFGS_OK) return status;

	phi_update_position(s, alpha, s->x, &s->f, s->g);
	s->delta_f = s->f - f0;

	/* one-step (memoryless) BFGS direction:  p' = g - A dx - B dg */
	for (size_t i = 0; i < n; i++) {
		s->dx0[i] = s->x[i] - s->x0[i];
		s->dx[i] = s->dx0[i];
		s->dg0[i] = s->g[i] - s->g0[i];
	}
	{
		double dxg = dot(s->dx0, s->g, n), dgg = dot(s->dg0, s->g, n);
		double dxdg = dot(s->dx0, s->dg0, n), dgnorm = nrm2(s->dg0, n);
		double A, B;
		if (dxdg != 0.0) {
			B = dxg / dxdg;
			A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
		} else {
			B = 0.0; A = 0.0;
		}
		for (size_t i = 0; i < n; i++) s->p[i] = s->g[i] - A * s->dx0[i] - B * s->dg0[i];
	}
	memcpy(s->g0, s->g, n * sizeof(double));
	memcpy(s->x0, s->x, n * sizeof(double));
	s->g0norm = nrm2(s->g0, n);
	s->pnorm = nrm2(s->p, n);
	{
		/* make it a unit-length descent direction */
		double pg = dot(s->p, s->g, n);
		double dir = (pg >= 0.0) ? -1.0 : +1.0;
		if (s->pnorm > 0.0)
			for (size_t i = 0; i < n; i++) s->p[i] *= dir / s->pnorm;
		s->pnorm = nrm2(s->p, n);
		s->fp0 = dot(s->p, s->g0, n);
	}
	phi_change_direction(s);
	s->iter++;
	return EMUB_BFGS_OK;
}

const double *emub_bfgs_x(const emub_bfgs *s) { return s->x; }
const double *emub_bfgs_gradient(const emub_bfgs *s) { return s->g; }
const double *emub_bfgs_dx(const emub_bfgs *s) { return s->dx; }
double emub_bfgs_minimum(const emub_bfgs *s) { return s->f; }

int emub_bfgs_test_gradient(const double *g, size_t n, double epsabs)
{
	if (epsabs < 0.0) return EMUB_BFGS_EINVAL;
	return (nrm2(g, n) < epsabs) ? EMUB_BFGS_OK : EMUB_BFGS_CONTINUE;
}
