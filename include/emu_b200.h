/*
 * emu_b200.h -- thin C-ABI of the B200-native Gaussian-process engine behind MADAIEmulator's libEmu
 * hot path.  Plain C: opaque handles, plain pointers and sizes, int status codes.  No GSL types, no
 * C++ and no torch types cross this boundary.  The shared library is
 * madaiemulator_b200/csrc/libemub.so (built by madaiemulator_b200/csrc/Makefile, sm_100a only).
 *
 * Every entry point names the reference interface it replaces (file:line under the reference's
 * src/).  All matrices are row-major doubles, like gsl_matrix (`ld` = row stride = gsl `tda`).
 * kernel ids follow optstruct.h:12-14; theta conventions follow the reference literally:
 *   power-exponential: thetas = (log amp, log nugget, log length_1 .. log length_d)   emulator.c:101-152
 *   Matern 3/2, 5/2  : thetas = (amp, nugget, log rho)   [amp, nugget raw]            emulator.c:344-480
 *
 * There is NO CPU fallback: every call runs CUDA kernels on the context's device and returns
 * EMUB_ECUDA when that is impossible.
 *
 * Limits: nparams <= 64, 1 + nregression_fns <= 104 (regression_order * nparams <= 102), at most 1024
 * observables / components in emub_predict_multi; anything larger is refused with EMUB_EINVAL.
 * Environment (all optional): EMUB_NO_GRAPHS=1 (direct launches instead of CUDA-graph replay),
 * EMUB_SMALL_CTAS (products with at most this many 64 x 64 sub-tiles use 32 x 32 ones, default 296),
 * EMUB_AUX_MAX (use the factorisation's side streams when a stream group holds at most this many
 * matrices, default 8, 0 = never).
 */
#ifndef EMU_B200_H
#define EMU_B200_H
#ifdef __cplusplus
extern "C" {
#endif

enum { EMUB_POWEREXP = 1, EMUB_MATERN32 = 2, EMUB_MATERN52 = 3 }; /* optstruct.h:12-14 */

enum {
	EMUB_OK = 0,
	EMUB_EDOM = 1,    /* covariance matrix not positive definite (reference: GSL_EDOM -> NaN, maxmultimin.c:327-350) */
	EMUB_EREG = 2,    /* H^T C^-1 H not positive definite (reference: exit(1), regression.c:159) */
	EMUB_EINVAL = 3,  /* bad argument */
	EMUB_ECUDA = 4,   /* CUDA runtime failure; see emub_last_error() */
	EMUB_ENOMEM = 5
};

typedef struct emub_ctx emub_ctx;           /* one device + its streams */
typedef struct emub_model emub_model;       /* design, training vector, kernel, regression order (modelstruct.h:28-98) */
typedef struct emub_emulator emub_emulator; /* cached factor for prediction (emulator_struct.h:20-29) */

/* ---- context ---------------------------------------------------------------------------------- */
int emub_ctx_create(int device, emub_ctx **out);
void emub_ctx_destroy(emub_ctx *ctx);
const char *emub_last_error(void);
const char *emub_version(void);
/* The cudaStream_t (as void*) that group-0 kernels are launched on, so a caller can record its own
 * CUDA events around library calls. */
void *emub_ctx_stream(emub_ctx *ctx);
/* number of concurrently scheduled matrix groups (streams); 1..4, default 4 */
int emub_ctx_set_groups(emub_ctx *ctx, int ngroups);
/* the launch sequence of a batch chunk (~160 dependent kernels per group) is captured once into a CUDA graph and
 * replayed; on by default (EMUB_NO_GRAPHS=1 or on = 0 issues the kernels one by one) */
int emub_ctx_use_graphs(emub_ctx *ctx, int on);

/* ---- model ------------------------------------------------------------------------------------ */
/* replaces alloc_modelstruct_2 (modelstruct.c:282) + makeHMatrix_fnptr (regression.c:100):
 * uploads X (n x d, row stride ldx) and the training vector y (n), builds H on the device.
 * max_slots bounds how many n x n problems are in flight at once (0 = choose from free memory). */
int emub_model_create(emub_ctx *ctx, const double *X, int ldx, int n, int d, const double *y,
                      int kernel, int regression_order, int max_slots, emub_model **out);
void emub_model_destroy(emub_model *m);
int emub_model_nthetas(const emub_model *m);         /* modelstruct.c:301-308 */
int emub_model_nregression_fns(const emub_model *m); /* modelstruct.c:319 */
int emub_model_slots(const emub_model *m);
int emub_model_kernel(const emub_model *m);          /* EMUB_POWEREXP / EMUB_MATERN32 / EMUB_MATERN52 */
/* replace the training vector (same design): the PCA components of one multivariate model share X
 * (multi_modelstruct.c:121-148) */
int emub_model_set_training(emub_model *m, const double *y);
/* several training vectors on one design: Y is n x ncomp row-major (row stride ldy), e.g. the pca_zmatrix of a
 * multivariate model (multi_modelstruct.c:295-316); component c is column c */
int emub_model_set_training_multi(emub_model *m, const double *Y, int ldy, int ncomp);
int emub_model_ncomponents(const emub_model *m);
/* Which gradient the batch entry points return.  EMUB_GRAD_LITERAL (default): the reference's formula, gradFnMulti /
 * getGradientCn (maxmultimin.c:416-608) -- raw y instead of the residual, the k-th factor of the kernel only, a
 * sigma^2 factor; it is NOT the gradient of what evalFnMulti returns.  EMUB_GRAD_EXACT (deviation
 * D-4, optional): the true gradient of that objective, d(-L)/dtheta = 1/2 tr(C^-1 dC) - 1/2 z^T dC z with
 * z = C^-1 (y - H beta); used by the optional refinement run of the restart driver (emub_estimate.h).  The nugget
 * derivative is taken on the diagonal only, like the reference (:520-522). */
enum { EMUB_GRAD_LITERAL = 0, EMUB_GRAD_EXACT = 1 };
int emub_model_set_gradient_mode(emub_model *m, int mode);
int emub_model_gradient_mode(const emub_model *m);

/* makeCovMatrix_fnptr (emulator.c:636): C (n x n, row stride ldc) at the FULL theta vector */
int emub_cov_matrix(emub_model *m, const double *thetas, double *C, int ldc);
/* makeHMatrix_fnptr (regression.c:100): H (n x p, row stride ldh) */
int emub_h_matrix(emub_model *m, double *H, int ldh);
/* makeKVector_fnptr (emulator.c:578) for mq points at once: K (n x mq, row stride ldk),
 * K[i][q] = c(x_i, pts_q) with the 1e-10 clamp */
int emub_k_vectors(emub_model *m, const double *thetas, const double *pts, int ldp, int mq, double *K, int ldk);

/* chol_inverse_cov_matrix (emulate-fns.c:275-300) for a caller-owned symmetric positive definite n x n matrix
 * (n = the model's nmodel_points; A row stride lda; only the model's factorisation workspace is used, not its design):
 * Ainv (row stride ldi) <- A^-1, *logdet <- log det A = 2 sum log L_ii (the reference forms (prod L_ii)^2, :290-293).
 * Returns EMUB_EDOM where the reference exits ("trying to cholesky a non postive def matrix", :282-285). */
int emub_spd_inverse(emub_model *m, const double *A, int lda, double *Ainv, int ldi, double *logdet);

/*
 * evalFnMulti / gradFnMulti / evalFnGradMulti (maxmultimin.c:288 / :416 / :615) for B points at
 * once.  thetas is B x (nthetas-1) row-major: theta without the amplitude, exactly the vector the
 * optimizer sees.  Outputs (host, any may be NULL): negL[B] (= -loglikelihood, what evalFnMulti
 * returns), grad[B x (nthetas-1)], sigma2[B] (estimateSigma, maxmultimin.c:215), status[B]
 * (EMUB_OK / EMUB_EDOM / EMUB_EREG; on failure negL and grad are NaN).  The determinant term is
 * sum 2 log L_ii (deviation D-1).  Returns EMUB_OK if the batch ran, whatever the per-point status.
 */
int emub_loglik_grad_batch(emub_model *m, const double *thetas, int B, int want_grad,
                           double *negL, double *grad, double *sigma2, int *status);
/* same with a training-vector (PCA component) index per point: comp[B] (NULL = component 0).  This is how the restart
 * fronts of all components of a multivariate model (estimate_multi, multivar_support.c:20-27) share one batch */
int emub_loglik_grad_batch_comp(emub_model *m, const double *thetas, const int *comp, int B, int want_grad,
                                double *negL, double *grad, double *sigma2, int *status);
/* same with a gradient request per point: want_grad[b] != 0 evaluates the gradient of point b (evalFnGradMulti), 0 the
 * value alone (evalFnMulti: the factor without the inverse, about 0.38 of the work).  The two kinds share every launch
 * they have in common, so a front that mixes line-search trial points with accepted ones stays one batched call;
 * grad rows of value-only points are zero.  Values are bit-identical to the uniform calls. */
int emub_loglik_grad_batch_mixed(emub_model *m, const double *thetas, const int *comp, const int *want_grad, int B,
                                 double *negL, double *grad, double *sigma2, int *status);
/* same, but thetas / outputs are DEVICE pointers (out: B x (nthetas+3) doubles per point:
 * negL, sigma2, status, logdet, grad[nthetas-1]); asynchronous on the context's streams until
 * emub_ctx_synchronize. */
int emub_loglik_grad_batch_dev(emub_model *m, const double *d_thetas, int B, int want_grad, double *d_out);
int emub_ctx_synchronize(emub_ctx *ctx);
/* last batch's extra diagnostics for point b (b < slots): logdet, beta[p] (estimateBeta, regression.c:120) */
int emub_loglik_extras(emub_model *m, int b, double *logdet, double *beta);

/* ---- prediction ------------------------------------------------------------------------------- */
/* alloc_emulator_struct (emulator_struct.c:13): covariance at the FULL thetas, Cholesky, the
 * cached C^-1-derived quantities and beta.  Returns EMUB_EDOM / EMUB_EREG where the reference exits. */
int emub_emulator_create(emub_model *m, const double *thetas, emub_emulator **out);
int emub_emulator_create_comp(emub_model *m, int comp, const double *thetas, emub_emulator **out);
void emub_emulator_destroy(emub_emulator *e);
int emub_emulator_beta(emub_emulator *e, double *beta); /* p values */
/* emulate_point (emulator_struct.c:124) for mq points: pts (mq x d, row stride ldp) -> mean[mq], var[mq] */
int emub_predict_batch(emub_emulator *e, const double *pts, int ldp, int mq, double *mean, double *var);
/* emulate_point for a handful of points (mq <= 8) -- the call pattern of an MCMC driver (EmuPlusPlus::QueryEmulator,
 * EmuPlusPlus.cpp:137-179): a latency path that streams the cached factor once instead of running the batched tensor
 * pass on a 128-column block (n = 4096: 340 -> 70 us per call for one point).  Same quantities as emub_predict_batch; the
 * sums run in another order, so the two agree to rounding (1e-15), not bit for bit. */
int emub_predict_few(emub_emulator *e, const double *pts, int ldp, int mq, double *mean, double *var);
/* emulate_point_multi (multivar_support.c:103-157) for mq points: all nr PCA components (emulators of one model) and
 * the back-projection  mean_i = ybar_i + sum_j U_ij sqrt(lambda_j) m_j,  var_i = sum_j U_ij^2 lambda_j v_j  on the
 * device; evecs is nt x nr row-major.  mean, var: mq x nt (nt, nr <= 1024).  nt = 0: PCA-space output mq x nr
 * (emulate_point_multi_pca, multivar_support.c:78) */
int emub_predict_multi(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                       const double *training_mean, const double *evecs, const double *evals, double *mean, double *var);
/* emub_predict_multi for a handful of points (mq <= 8) on the latency path (see emub_predict_few) */
int emub_predict_multi_few(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                           const double *training_mean, const double *evecs, const double *evals, double *mean, double *var);
/* device-pointer variant: d_pts is mq x d contiguous; asynchronous */
int emub_predict_batch_dev(emub_emulator *e, const double *d_pts, int mq, double *d_mean, double *d_var);

/* ---- instrumentation -------------------------------------------------------------------------- */
/* When enabled, every kernel family launch is bracketed by CUDA events on its stream (forces one
 * group); emub_profile_read returns accumulated milliseconds, launches and algorithmic flops/bytes
 * per family since the last reset. */
enum { EMUB_K_COV = 0, EMUB_K_POTF2, EMUB_K_GEMM_CHOL, EMUB_K_GEMM_TRTRI, EMUB_K_GEMM_LAUUM, EMUB_K_SKINNY,
       EMUB_K_SMALL, EMUB_K_GRAD, EMUB_K_KCROSS, EMUB_K_GEMM_PRED, EMUB_K_PRED_FINAL, EMUB_K_NFAMILIES };
int emub_profile_enable(emub_ctx *ctx, int on);
int emub_profile_reset(emub_ctx *ctx);
int emub_profile_read(emub_ctx *ctx, int family, double *ms, long long *launches, double *work);
const char *emub_profile_name(int family);
/* total kernel launches issued by this context since creation */
long long emub_launch_count(emub_ctx *ctx);

/* ---- debugging / tests ------------------------------------------------------------------------ */
/* copy an internal n x n matrix of slot b of the last batch to the host: which = 0 Cinv (lower
 * triangle + diagonal blocks valid), 1 W = L^-1 (lower) */
int emub_debug_fetch(emub_model *m, int b, int which, double *out, int ldo);
/* the device exp used by the covariance / gradient kernels (x <= 0), for accuracy tests */
int emub_debug_exp(emub_ctx *ctx, const double *x, int n, double *out);
/* the pre-scaled variant the power-exponential kernels use (argument in units of ln2/64, replicated table): e^x */
int emub_debug_exp_scaled(emub_ctx *ctx, const double *x, int n, double *out);
/* factor only: runs covariance + Cholesky at theta-less-amp and returns L (lower, n x n) */
int emub_debug_cholesky(emub_model *m, const double *theta_less_amp, double *L, int ldl, double *logdet);

#ifdef __cplusplus
}
#endif
#endif
