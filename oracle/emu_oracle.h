/*
 * emu_oracle.h -- CPU restatement (plain C, no GSL types) of MADAIEmulator libEmu's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this; the product (madaiemulator_b200/) never does.
 *
 * Parity status: PINNED.  The reference ships no golden vectors for this path, so
 * this restatement is pinned against outputs of the reference's own sources compiled here
 * (oracle/_ref/libemu_ref.so, built by oracle/Makefile) -- tests/test_oracle_vs_ref.py -- and against
 * the fixtures those sources generated (tests/golden/ JSON files, script tests/golden/make_golden.py).
 *
 * kernel ids follow src/optstruct.h:12-14 (POWEREXPCOVFN 1, MATERN32 2, MATERN52 3).
 */
#ifndef EMU_ORACLE_H
#define EMU_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { EMUO_POWEREXP = 1, EMUO_MATERN32 = 2, EMUO_MATERN52 = 3 };
/* determinant handling in the likelihood */
enum { EMUO_DET_LOGSUM = 0 /* deviation D-1 */, EMUO_DET_PRODUCT = 1 /* literal maxmultimin.c:355-358 */ };

int emuo_nthetas(int kernel, int d);        /* modelstruct.c:301-308 */
int emuo_nregression_fns(int order, int d); /* modelstruct.c:319 */

/* emulator.c:101 / :344 / :438 -- thetas is the FULL vector */
double emuo_cov_pair(int kernel, const double *xa, const double *xb, const double *thetas, int d);
/* emulator.c:636 -- C is n x n row-major */
void emuo_cov_matrix(int kernel, const double *X, int n, int d, const double *thetas, double *C);
/* emulator.c:173 (power-exp, literal).  Matern: deviation D-3 (see emu_oracle.c). index in [2,nthetas) */
void emuo_deriv_matrix(int kernel, const double *X, int n, int d, double theta_length, int index, double *D);
/* emulator.c:578 -- includes the 1e-10 clamp */
void emuo_k_vector(int kernel, const double *X, int n, int d, const double *xnew, const double *thetas, double *k);
/* regression.c:9-67 */
void emuo_h_vector(int order, const double *x, int d, double *h);
void emuo_h_matrix(int order, const double *X, int n, int d, double *H);

/* lower Cholesky in place (upper gets L^T); returns 1 if not positive definite */
int emuo_cholesky(double *A, int n);
/* in: emuo_cholesky output; out: full symmetric inverse */
void emuo_cholesky_invert(double *A, int n);
/* regression.c:120 ; returns 1 if H^T C^-1 H is not PD (the reference exits) */
int emuo_estimate_beta(const double *H, const double *Cinv, const double *y, int n, int p, double *beta);

/* modelstruct.c:188 / optstruct.c:142 (use_data_scales = 1, no fixed nugget) ; ranges is nthetas x 2 */
void emuo_sample_scales(const double *X, int n, int d, double *scales);
void emuo_ranges(int kernel, const double *X, int n, int d, double *ranges);

/*
 * evalFnMulti + gradFnMulti (maxmultimin.c:288, :416) at theta' = theta without the amplitude.
 * Training-covariance convention for the Matern kernels: deviation D-2.
 * Outputs (any may be NULL): negL, grad[nthetas-1], sigma2, logdet, beta[p].
 * Returns 0 ok, 1 covariance not PD (negL = NAN), 2 regression matrix not PD.
 */
int emuo_loglik_grad(int kernel, int order, const double *X, int n, int d, const double *y,
                     const double *theta_less_amp, int det_mode,
                     double *negL, double *grad, double *sigma2, double *logdet, double *beta);

/* emulator_struct.c:13 / :124 */
typedef struct emuo_emulator emuo_emulator;
emuo_emulator *emuo_emulator_create(int kernel, int order, const double *X, int n, int d, const double *y,
                                    const double *thetas);
void emuo_emulator_free(emuo_emulator *e);
void emuo_emulate(const emuo_emulator *e, const double *pts, int m, double *mean, double *var);
void emuo_emulator_beta(const emuo_emulator *e, double *beta);

/* multivar_support.c:126-151: back-projection PCA -> observable space */
void emuo_backproject(int nt, int nr, const double *training_mean, const double *evecs /* nt x nr */,
                      const double *evals /* nr */, const double *mean_pca, const double *var_pca,
                      double *mean_out, double *var_out);

#ifdef __cplusplus
}
#endif
#endif
