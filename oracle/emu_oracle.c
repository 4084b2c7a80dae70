/*
 * emu_oracle.c -- CPU restatement of the libEmu hot path; see emu_oracle.h for role and status.
 * TEST INFRASTRUCTURE ONLY.  Citations are file:line under /root/reference/src.
 *
 * The arithmetic of the covariance functions, regression basis, beta/sigma estimates, likelihood
 * and prediction follows the reference literally (same operation order, same truncated literals).
 * Dense linear algebra (the role GSL plays for the reference) is restated from textbook
 * definitions.  Where the reference has no well-defined behaviour the deviation register of
 * SURVEY.md section 8a applies:
 *   D-1  log det C = sum 2 log L_ii (det_mode = EMUO_DET_LOGSUM); the literal running product is
 *        available as EMUO_DET_PRODUCT.
 *   D-2  Matern training covariance: the literal kernel called with theta = (1, exp(theta_1),
 *        theta_2), i.e. unit amplitude and exp-scaled nugget like the power-exponential path (the
 *        reference passes amp = 0 raw and cannot factorise, Q6).
 *   D-3  Matern length derivative: d c / d log(rho) of the literal kernel with r recomputed per
 *        pair and rho = exp(theta_2) (the reference accumulates r across pairs, Q8):
 *        Matern32: t^2 exp(-t), t = root3 r / rho;  Matern52: (t^2/3)(1 + t) exp(-t), t = root5 r / rho.
 */
#include "emu_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

int emuo_nthetas(int kernel, int d)
{
	return (kernel == EMUO_MATERN32 || kernel == EMUO_MATERN52) ? 3 : d + 2;
}
int emuo_nregression_fns(int order, int d)
{
	if (order < 0 || order > 3) order = 0;
	return 1 + order * d;
}

/* emulator.c:101-152 */
static double cov_gaussian(const double *xm, const double *xn, const double *thetas, int d)
{
	int truecount = 0;
	double exponent = 0.0;
	double amp = exp(thetas[0]);
	double nug = exp(thetas[1]);
	for (int i = 0; i < d; i++) {
		double r_temp = exp(thetas[i + 2]);
		r_temp = r_temp * r_temp;
		double dist_temp = fabs(xm[i] - xn[i]);
		exponent += (-1.0 / 2.0) * dist_temp * dist_temp / (r_temp);
		if (dist_temp < 0.0000000001) truecount++;
	}
	double covariance = exp(exponent) * amp;
	if (truecount == d) covariance += nug;
	return covariance;
}
/* emulator.c:344-386 */
static double cov_matern32(const double *xm, const double *xn, const double *thetas, int d)
{
	double covariance, distance = 0.0;
	int truecount = 0;
	double amp = thetas[0], nugget = thetas[1], rho = exp(thetas[2]);
	double root3 = 1.732050808;
	for (int i = 0; i < d; i++) {
		double temp_dist = fabs(xm[i] - xn[i]);
		distance += temp_dist * temp_dist;
		if (temp_dist < 0.0000000000000001) truecount++;
	}
	distance = sqrt(distance);
	if (distance > 0.0) covariance = amp * (1 + root3 * (distance / rho)) * exp(-root3 * (distance / rho));
	else covariance = amp;
	if (truecount == d) covariance += nugget;
	return covariance;
}
/* emulator.c:438-480 */
static double cov_matern52(const double *xm, const double *xn, const double *thetas, int d)
{
	double covariance = 0.0, distance = 0.0;
	int truecount = 0;
	double amp = thetas[0], nugget = thetas[1], rho = exp(thetas[2]);
	double root5 = 2.236067978;
	for (int i = 0; i < d; i++) {
		distance += pow(fabs(xm[i] - xn[i]), 2.0);
		if (fabs(xm[i] - xn[i]) < 0.0000000000000001) truecount++;
	}
	distance = sqrt(distance);
	double d_over_r = distance / rho;
	if (distance > 0.0)
		covariance = amp * (1 + root5 * (d_over_r) + (5.0 / 3.0) * (d_over_r) * (d_over_r)) * exp(-root5 * (d_over_r));
	else if (distance == 0)
		covariance = amp;
	if (truecount == d) covariance += nugget;
	return covariance;
}

double emuo_cov_pair(int kernel, const double *xa, const double *xb, const double *thetas, int d)
{
	switch (kernel) {
	case EMUO_MATERN32: return cov_matern32(xa, xb, thetas, d);
	case EMUO_MATERN52: return cov_matern52(xa, xb, thetas, d);
	default: return cov_gaussian(xa, xb, thetas, d);
	}
}

/* emulator.c:636-653 */
void emuo_cov_matrix(int kernel, const double *X, int n, int d, const double *thetas, double *C)
{
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++)
			C[(size_t)i * n + j] = emuo_cov_pair(kernel, X + (size_t)i * d, X + (size_t)j * d, thetas, d);
}

void emuo_deriv_matrix(int kernel, const double *X, int n, int d, double theta_length, int index, double *D)
{
	if (kernel == EMUO_MATERN32 || kernel == EMUO_MATERN52) {
		/* deviation D-3 */
		const double rho = exp(theta_length);
		const double root = (kernel == EMUO_MATERN32) ? 1.732050808 : 2.236067978;
		for (int i = 0; i < n; i++)
			for (int j = 0; j < n; j++) {
				double r2 = 0.0;
				for (int k = 0; k < d; k++) {
					double t = X[(size_t)i * d + k] - X[(size_t)j * d + k];
					r2 += t * t;
				}
				double t = root * (sqrt(r2) / rho);
				double v = (kernel == EMUO_MATERN32) ? t * t * exp(-t) : (t * t / 3.0) * (1.0 + t) * exp(-t);
				D[(size_t)i * n + j] = v;
			}
		return;
	}
	/* emulator.c:173-209 */
	const int indexScaled = index - 2;
	const double expTheta = exp(-2.0 * theta_length);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) {
			double rtemp = X[(size_t)i * d + indexScaled] - X[(size_t)j * d + indexScaled];
			D[(size_t)i * n + j] = exp(-0.5 * expTheta * rtemp * rtemp - 2 * theta_length) * rtemp * rtemp;
		}
}

/* emulator.c:578-593 */
void emuo_k_vector(int kernel, const double *X, int n, int d, const double *xnew, const double *thetas, double *k)
{
	for (int i = 0; i < n; i++) {
		double cov = emuo_cov_pair(kernel, X + (size_t)i * d, xnew, thetas, d);
		if (cov < 1E-10) cov = 0.0;
		k[i] = cov;
	}
}

/* regression.c:9-67 */
void emuo_h_vector(int order, const double *x, int d, double *h)
{
	if (order < 0 || order > 3) order = 0;
	h[0] = 1;
	if (order >= 1) for (int i = 0; i < d; i++) h[i + 1] = x[i];
	if (order >= 2) for (int i = 0; i < d; i++) h[d + i + 1] = x[i] * x[i];
	if (order >= 3) for (int i = 0; i < d; i++) h[2 * d + i + 1] = x[i] * x[i] * x[i];
}
/* regression.c:100-112 */
void emuo_h_matrix(int order, const double *X, int n, int d, double *H)
{
	int p = emuo_nregression_fns(order, d);
	for (int i = 0; i < n; i++) emuo_h_vector(order, X + (size_t)i * d, d, H + (size_t)i * p);
}

/* ---- dense linear algebra (GSL's role) --------------------------------------------------------- */
int emuo_cholesky(double *a, int n)
{
	int bad = 0;
	for (int i = 0; i < n; i++) {
		double *ri = a + (size_t)i * n;
		for (int j = 0; j <= i; j++) {
			const double *rj = a + (size_t)j * n;
			double s = ri[j];
			for (int k = 0; k < j; k++) s -= ri[k] * rj[k];
			if (j == i) {
				if (!(s > 0.0)) { bad = 1; s = fabs(s); if (!(s > 0.0)) s = 1.0; }
				ri[i] = sqrt(s);
			} else
				ri[j] = s / rj[j];
		}
	}
	for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) a[(size_t)i * n + j] = a[(size_t)j * n + i];
	return bad;
}

void emuo_cholesky_invert(double *a, int n)
{
	double *x = (double *)calloc((size_t)n * n, sizeof(double));
	for (int i = 0; i < n; i++) {
		double *xi = x + (size_t)i * n;
		const double *li = a + (size_t)i * n;
		for (int k = 0; k < i; k++) {
			const double t = li[k];
			const double *xk = x + (size_t)k * n;
			if (t != 0.0) for (int j = 0; j <= k; j++) xi[j] -= t * xk[j];
		}
		xi[i] = 1.0;
		const double inv = 1.0 / li[i];
		for (int j = 0; j <= i; j++) xi[j] *= inv;
	}
	for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) a[(size_t)i * n + j] = 0.0;
	for (int k = 0; k < n; k++) {
		const double *xk = x + (size_t)k * n;
		for (int i = 0; i <= k; i++) {
			const double t = xk[i];
			double *ai = a + (size_t)i * n;
			if (t != 0.0) for (int j = 0; j <= i; j++) ai[j] += t * xk[j];
		}
	}
	for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) a[(size_t)i * n + j] = a[(size_t)j * n + i];
	free(x);
}

static void matvec(const double *A, const double *x, int n, int m, double *y)
{ /* y (n) = A (n x m) x */
	for (int i = 0; i < n; i++) {
		const double *r = A + (size_t)i * m;
		double s = 0.0;
		for (int j = 0; j < m; j++) s += x[j] * r[j];
		y[i] = s;
	}
}
static double dotp(const double *a, const double *b, int n)
{
	double s = 0.0;
	for (int i = 0; i < n; i++) s += a[i] * b[i];
	return s;
}

/* regression.c:120-176.  HtCinv = H^T Cinv (p x n); denom = HtCinv H; beta = denom^-1 (HtCinv y) */
static int beta_parts(const double *H, const double *Cinv, const double *y, int n, int p, double *beta,
                      double *HtCinv_out /* p x n or NULL */, double *denom_inv_out /* p x p or NULL */)
{
	double *HtCinv = (double *)calloc((size_t)p * n, sizeof(double));
	double *den = (double *)calloc((size_t)p * p, sizeof(double));
	double *num = (double *)calloc(p, sizeof(double));
	/* dgemm(Trans, NoTrans): HtCinv[i][j] = sum_k H[k][i] Cinv[k][j] */
	for (int k = 0; k < n; k++)
		for (int i = 0; i < p; i++) {
			const double t = H[(size_t)k * p + i];
			if (t != 0.0) for (int j = 0; j < n; j++) HtCinv[(size_t)i * n + j] += t * Cinv[(size_t)k * n + j];
		}
	for (int i = 0; i < p; i++)
		for (int k = 0; k < n; k++) {
			const double t = HtCinv[(size_t)i * n + k];
			if (t != 0.0) for (int j = 0; j < p; j++) den[(size_t)i * p + j] += t * H[(size_t)k * p + j];
		}
	int bad = emuo_cholesky(den, p);
	if (!bad) {
		emuo_cholesky_invert(den, p);
		matvec(HtCinv, y, p, n, num);
		matvec(den, num, p, p, beta);
	}
	if (HtCinv_out) memcpy(HtCinv_out, HtCinv, sizeof(double) * (size_t)p * n);
	if (denom_inv_out) memcpy(denom_inv_out, den, sizeof(double) * (size_t)p * p);
	free(HtCinv); free(den); free(num);
	return bad;
}
int emuo_estimate_beta(const double *H, const double *Cinv, const double *y, int n, int p, double *beta)
{
	return beta_parts(H, Cinv, y, n, p, beta, NULL, NULL);
}

/* modelstruct.c:188-213 */
void emuo_sample_scales(const double *X, int n, int d, double *scales)
{
	for (int i = 0; i < d; i++) {
		double min_value = fabs(X[(size_t)1 * d + i] - X[i]);
		for (int j = 1; j < n - 1; j++) {
			double value = fabs(X[(size_t)(j + 1) * d + i] - X[(size_t)j * d + i]);
			if (value < min_value) min_value = value;
		}
		if (min_value < 1.0e-5) min_value = 1.0e-5;
		scales[i] = min_value;
	}
}
/* optstruct.c:142-226 with use_data_scales = 1, fixed_nugget_mode = 0 (modelstruct.c:322-326) */
void emuo_ranges(int kernel, const double *X, int n, int d, double *ranges)
{
	int nthetas = emuo_nthetas(kernel, d);
	double *scales = (double *)malloc(sizeof(double) * d);
	double rangeMin, rangeMax;
	emuo_sample_scales(X, n, d, scales);
	if (kernel == EMUO_POWEREXP) { rangeMin = 0.0001; rangeMax = 5; }
	else { rangeMin = 0; rangeMax = 10.0; }
	ranges[0] = 0.0001; ranges[1] = rangeMax;
	ranges[2] = -5.0; ranges[3] = -2.0;
	for (int i = 2; i < nthetas; i++) {
		if (kernel == EMUO_POWEREXP) {
			rangeMin = 0.5 * log(scales[i - 2]);
			rangeMax = log(25 * exp(rangeMin));
		} else {
			rangeMin = 0.5 * (scales[i - 2]);
		}
		ranges[2 * i] = rangeMin;
		ranges[2 * i + 1] = rangeMax;
	}
	free(scales);
}

/* maxmultimin.c:288-394 (value), :416-608 (gradient), :215-273 (sigma), estimator-fns.c:38-103 */
int emuo_loglik_grad(int kernel, int order, const double *X, int n, int d, const double *y,
                     const double *theta_less_amp, int det_mode,
                     double *negL, double *grad, double *sigma2_out, double *logdet_out, double *beta_out)
{
	const int nthetas = emuo_nthetas(kernel, d);
	const int p = emuo_nregression_fns(order, d);
	const size_t nn = (size_t)n * n;
	int rc = 0;
	double *theta_local = (double *)malloc(sizeof(double) * nthetas);
	double *C = (double *)malloc(sizeof(double) * nn);
	double *H = (double *)malloc(sizeof(double) * (size_t)n * p);
	double *beta = (double *)malloc(sizeof(double) * p);
	double *mean = (double *)malloc(sizeof(double) * n);
	double *resid = (double *)malloc(sizeof(double) * n);
	double *tmp = (double *)malloc(sizeof(double) * n);
	double *alpha = (double *)malloc(sizeof(double) * n);
	double *D = NULL;

	/* maxmultimin.c:311-313 (power-exp: log-amp 0) ; Matern: deviation D-2 */
	if (kernel == EMUO_POWEREXP) {
		theta_local[0] = 0.0;
		for (int i = 1; i < nthetas; i++) theta_local[i] = theta_less_amp[i - 1];
	} else {
		theta_local[0] = 1.0;
		theta_local[1] = exp(theta_less_amp[0]);
		theta_local[2] = theta_less_amp[1];
	}
	emuo_cov_matrix(kernel, X, n, d, theta_local, C);
	emuo_h_matrix(order, X, n, d, H);
	if (emuo_cholesky(C, n)) {
		if (negL) *negL = NAN;
		rc = 1;
		goto done;
	}
	double logdet;
	if (det_mode == EMUO_DET_PRODUCT) {
		double det = 1.0; /* maxmultimin.c:355-358 */
		for (int i = 0; i < n; i++) det *= C[(size_t)i * n + i];
		det = det * det;
		logdet = log(det); /* estimator-fns.c:56 */
	} else {
		logdet = 0.0;
		for (int i = 0; i < n; i++) logdet += 2.0 * log(C[(size_t)i * n + i]);
	}
	emuo_cholesky_invert(C, n); /* C is now C1^-1 */

	if (emuo_estimate_beta(H, C, y, n, p, beta)) { if (negL) *negL = NAN; rc = 2; goto done; }
	/* estimateSigma, maxmultimin.c:245-263 */
	for (int i = 0; i < n; i++) mean[i] = dotp(beta, H + (size_t)i * p, p);
	for (int i = 0; i < n; i++) resid[i] = y[i] - mean[i];
	matvec(C, resid, n, n, tmp);
	double sigma2 = dotp(y, tmp, n) / (double)n;
	/* getLogLikelyhood, estimator-fns.c:48,85-95 */
	const double log_2_pi = 1.83788;
	double L = -(1.0 / 2.0) * logdet - (n / 2.0) * log_2_pi;
	L += dotp(resid, tmp, n) * (-1.0 / 2.0);
	if (negL) *negL = -1 * L;
	if (sigma2_out) *sigma2_out = sigma2;
	if (logdet_out) *logdet_out = logdet;
	if (beta_out) memcpy(beta_out, beta, sizeof(double) * p);

	if (grad) {
		/* gradFnMulti, maxmultimin.c:503-538 ; getGradientCn :571-608 restated in its O(n^2)
		 * algebraically identical form: grad = 0.5 a^T dC a - 0.5 sum_ij Cinv_ij dC_ji, a = Cinv y */
		const double amp = exp(log(sigma2));
		const double nug = (kernel == EMUO_POWEREXP) ? exp(theta_local[1]) : theta_local[1];
		matvec(C, y, n, n, alpha);
		double tr = 0.0;
		for (int i = 0; i < n; i++) tr += C[(size_t)i * n + i];
		grad[0] = -1.0 * (-0.5 * nug * tr + 0.5 * nug * dotp(alpha, alpha, n));
		D = (double *)malloc(sizeof(double) * nn);
		for (int t = 2; t < nthetas; t++) {
			emuo_deriv_matrix(kernel, X, n, d, theta_local[t], t, D);
			double trace = 0.0, quad = 0.0;
			for (int i = 0; i < n; i++) {
				double rowq = 0.0;
				for (int j = 0; j < n; j++) {
					const double dij = amp * D[(size_t)i * n + j];
					trace += C[(size_t)i * n + j] * (amp * D[(size_t)j * n + i]);
					rowq += dij * alpha[j];
				}
				quad += alpha[i] * rowq;
			}
			grad[t - 1] = -1.0 * (-0.5 * trace + 0.5 * quad);
		}
	}
done:
	free(theta_local); free(C); free(H); free(beta); free(mean); free(resid); free(tmp); free(alpha); free(D);
	return rc;
}

/* ---- prediction ------------------------------------------------------------------------------------ */
struct emuo_emulator {
	int kernel, order, n, d, p, nthetas;
	double *X, *y, *thetas;
	double *Cinv, *H, *beta;
	/* hoisted per-point invariants of makeEmulatedMean / makeEmulatedVariance (emulator.c:680-695,734-769) */
	double *Cinv_y, *Cinv_Hbeta, *Cinv_H /* n x p */, *HtCinvH_inv /* p x p */;
};

emuo_emulator *emuo_emulator_create(int kernel, int order, const double *X, int n, int d, const double *y,
                                    const double *thetas)
{
	emuo_emulator *e = (emuo_emulator *)calloc(1, sizeof(*e));
	const int p = emuo_nregression_fns(order, d), nt = emuo_nthetas(kernel, d);
	e->kernel = kernel; e->order = (order < 0 || order > 3) ? 0 : order; e->n = n; e->d = d; e->p = p; e->nthetas = nt;
	e->X = (double *)malloc(sizeof(double) * (size_t)n * d); memcpy(e->X, X, sizeof(double) * (size_t)n * d);
	e->y = (double *)malloc(sizeof(double) * n); memcpy(e->y, y, sizeof(double) * n);
	e->thetas = (double *)malloc(sizeof(double) * nt); memcpy(e->thetas, thetas, sizeof(double) * nt);
	e->Cinv = (double *)malloc(sizeof(double) * (size_t)n * n);
	e->H = (double *)malloc(sizeof(double) * (size_t)n * p);
	e->beta = (double *)calloc(p, sizeof(double));
	/* emulator_struct.c:28-32 */
	emuo_cov_matrix(kernel, X, n, d, thetas, e->Cinv);
	if (emuo_cholesky(e->Cinv, n)) { emuo_emulator_free(e); return NULL; } /* reference: exit(1), emulate-fns.c:282 */
	emuo_cholesky_invert(e->Cinv, n);
	emuo_h_matrix(order, X, n, d, e->H);
	if (emuo_estimate_beta(e->H, e->Cinv, y, n, p, e->beta)) { emuo_emulator_free(e); return NULL; }

	e->Cinv_y = (double *)malloc(sizeof(double) * n);
	e->Cinv_Hbeta = (double *)malloc(sizeof(double) * n);
	e->Cinv_H = (double *)calloc((size_t)n * p, sizeof(double));
	e->HtCinvH_inv = (double *)calloc((size_t)p * p, sizeof(double));
	double *Hbeta = (double *)malloc(sizeof(double) * n);
	matvec(e->Cinv, y, n, n, e->Cinv_y);       /* emulator.c:680 */
	matvec(e->H, e->beta, n, p, Hbeta);        /* emulator.c:691 */
	matvec(e->Cinv, Hbeta, n, n, e->Cinv_Hbeta); /* emulator.c:693 */
	free(Hbeta);
	/* emulator.c:734  Cinv . H */
	for (int i = 0; i < n; i++)
		for (int k = 0; k < n; k++) {
			const double t = e->Cinv[(size_t)i * n + k];
			if (t != 0.0) for (int j = 0; j < p; j++) e->Cinv_H[(size_t)i * p + j] += t * e->H[(size_t)k * p + j];
		}
	/* emulator.c:745  H^T (Cinv H) ; :751,769 Cholesky + invert */
	for (int k = 0; k < n; k++)
		for (int i = 0; i < p; i++) {
			const double t = e->H[(size_t)k * p + i];
			if (t != 0.0) for (int j = 0; j < p; j++) e->HtCinvH_inv[(size_t)i * p + j] += t * e->Cinv_H[(size_t)k * p + j];
		}
	if (emuo_cholesky(e->HtCinvH_inv, p)) { emuo_emulator_free(e); return NULL; }
	emuo_cholesky_invert(e->HtCinvH_inv, p);
	return e;
}

void emuo_emulator_free(emuo_emulator *e)
{
	if (!e) return;
	free(e->X); free(e->y); free(e->thetas); free(e->Cinv); free(e->H); free(e->beta);
	free(e->Cinv_y); free(e->Cinv_Hbeta); free(e->Cinv_H); free(e->HtCinvH_inv);
	free(e);
}
void emuo_emulator_beta(const emuo_emulator *e, double *beta) { memcpy(beta, e->beta, sizeof(double) * e->p); }

/* emulator_struct.c:124-143 -> emulator.c:578 (k), regression.c (h), emulator.c:672 (mean), :720 (variance) */
void emuo_emulate(const emuo_emulator *e, const double *pts, int m, double *mean, double *var)
{
	const int n = e->n, p = e->p, d = e->d;
	double *k = (double *)malloc(sizeof(double) * n);
	double *h = (double *)malloc(sizeof(double) * p);
	double *rho = (double *)malloc(sizeof(double) * p);
	double *rho2 = (double *)malloc(sizeof(double) * p);
	double *holder = (double *)malloc(sizeof(double) * n);
	for (int q = 0; q < m; q++) {
		const double *x = pts + (size_t)q * d;
		emuo_k_vector(e->kernel, e->X, n, d, x, e->thetas, k);
		emuo_h_vector(e->order, x, d, h);
		/* makeEmulatedMean, emulator.c:680-703 */
		double emulated_mean = dotp(k, e->Cinv_y, n);
		double regression_cpt = dotp(h, e->beta, p);
		double residual_cpt = dotp(k, e->Cinv_Hbeta, n);
		mean[q] = regression_cpt + emulated_mean - residual_cpt;
		/* kappa, emulator_struct.c:135 */
		double kappa = emuo_cov_pair(e->kernel, x, x, e->thetas, d);
		/* makeEmulatedVariance, emulator.c:738-784 : rho = h - (Cinv H)^T k */
		for (int j = 0; j < p; j++) rho[j] = 0.0;
		for (int i = 0; i < n; i++) {
			const double t = -1.0 * k[i];
			if (t != 0.0) for (int j = 0; j < p; j++) rho[j] += t * e->Cinv_H[(size_t)i * p + j];
		}
		for (int j = 0; j < p; j++) rho[j] += h[j];
		matvec(e->HtCinvH_inv, rho, p, p, rho2);
		double reg = dotp(rho, rho2, p);
		matvec(e->Cinv, k, n, n, holder);
		double ev = dotp(k, holder, n);
		var[q] = kappa - ev + reg;
	}
	free(k); free(h); free(rho); free(rho2); free(holder);
}

/* multivar_support.c:126-151 */
void emuo_backproject(int nt, int nr, const double *training_mean, const double *evecs, const double *evals,
                      const double *mean_pca, const double *var_pca, double *mean_out, double *var_out)
{
	for (int i = 0; i < nt; i++) {
		double s = 0.0, v = 0.0;
		for (int j = 0; j < nr; j++) {
			s += evecs[(size_t)i * nr + j] * sqrt(evals[j]) * mean_pca[j];
			v += pow(evecs[(size_t)i * nr + j], 2.0) * evals[j] * var_pca[j];
		}
		mean_out[i] = training_mean[i] + s;
		var_out[i] = v;
	}
}
