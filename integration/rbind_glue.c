/*
 * rbind_glue.c -- batched replacements for the list-shaped entry points of the reference's R binding
 * (src/libRbind/rbind.c), with the reference's exact `.C()` signatures (all arguments are pointers, matrices
 * arrive column-major and are transposed as convertDoubleToMatrix does: m[j][i] = in[j + ny*i], rbind.c:840-855).
 *
 *   callEvalLhoodList   rbind.c:626-724   loops evalFnMulti over the rows of a theta list
 *                                         -> ONE emub_loglik_grad_batch over the whole list
 *   callEmulateAtList   rbind.c:121-187   loops the emulator over a list of points (emulateAtPointList)
 *                                         -> one emub_emulator_create + one emub_predict_batch
 *
 *   setupEmulateMC / callEmulateMC / freeEmulateMC                 rbind.c:299-455   host-side C^-1 + emulateQuick per point
 *   setupEmulateMCMulti / callEmulateMCMulti / freeEmulateMCMulti  rbind.c:483-590   the same, looped over the outputs
 *                                         -> cached factors on the device + the latency path for one point, all outputs
 *                                            of the multivariate case in one set of launches
 *
 * Pure pointer marshalling: no R headers, no GSL.  Link into libRBIND in place of the reference bodies
 * (the other entry points -- callEstimate, callEmulateAtPt -- keep working through integration/libemu_glue.c, which also
 * makes the UNCHANGED rbind.c fast: emulateQuick, chol_inverse_cov_matrix, makeCovMatrix are bound there).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "emu_b200.h"

static emub_ctx *rb_ctx = NULL;

static emub_ctx *rbind_ctx(void)
{
	if (!rb_ctx) {
		const char *dev = getenv("EMUB_DEVICE");
		if (emub_ctx_create(dev ? atoi(dev) : 0, &rb_ctx) != EMUB_OK) {
			fprintf(stderr, "rbind_glue: %s\n", emub_last_error());
			return NULL;
		}
	}
	return rb_ctx;
}

/* column-major R array (nrow x ncol stored as in[j + nrow*i], j the row) -> row-major nrow x ncol */
static double *to_row_major(const double *in, int ncol, int nrow)
{
	double *out = (double *)malloc(sizeof(double) * (size_t)ncol * (size_t)nrow);
	for (int i = 0; i < ncol; i++)
		for (int j = 0; j < nrow; j++) out[(size_t)j * ncol + i] = in[j + (size_t)nrow * i];
	return out;
}

void callEvalLhoodList(double *xmodel_in, int *nparams_in, double *pointList_in, int *nevalPoints_in, double *training_in,
                       int *nmodelPoints_in, int *nthetas_in, double *answer, int *cov_fn_index_in, int *regression_order_in)
{
	const int d = *nparams_in, n = *nmodelPoints_in, nth = *nthetas_in, B = *nevalPoints_in;
	emub_ctx *ctx = rbind_ctx();
	double *X = to_row_major(xmodel_in, d, n);
	double *P = to_row_major(pointList_in, nth, B); /* B rows of nthetas values; evalFnMulti reads the first nthetas-1 (maxmultimin.c:311-313) */
	double *th = (double *)malloc(sizeof(double) * (size_t)B * (size_t)(nth - 1));
	for (int b = 0; b < B; b++)
		for (int k = 0; k < nth - 1; k++) th[(size_t)b * (nth - 1) + k] = P[(size_t)b * nth + k];
	emub_model *m = NULL;
	int *status = (int *)calloc((size_t)B, sizeof(int));
	if (!ctx || emub_model_create(ctx, X, d, n, d, training_in, *cov_fn_index_in, *regression_order_in, 0, &m) != EMUB_OK ||
	    emub_loglik_grad_batch(m, th, B, 0, answer, NULL, NULL, status) != EMUB_OK) {
		fprintf(stderr, "callEvalLhoodList: %s\n", emub_last_error());
		for (int b = 0; b < B; b++) answer[b] = NAN;
	}
	/* evalFnMulti returns NaN for a non-PD matrix (maxmultimin.c:327-350): already NaN in answer */
	if (m) emub_model_destroy(m);
	free(status); free(th); free(P); free(X);
}

void callEmulateAtList(double *xmodel_in, int *nparams_in, double *points_in, int *nemupoints, double *training_in,
                       int *nmodelpts, double *thetas_in, int *nthetas_in, double *final_emulated_y,
                       double *final_emulated_variance, int *cov_fn_index_in, int *regression_order_in)
{
	const int d = *nparams_in, n = *nmodelpts, mq = *nemupoints;
	(void)nthetas_in;
	emub_ctx *ctx = rbind_ctx();
	double *X = to_row_major(xmodel_in, d, n);
	double *Q = to_row_major(points_in, d, mq);
	emub_model *m = NULL;
	emub_emulator *e = NULL;
	if (!ctx || emub_model_create(ctx, X, d, n, d, training_in, *cov_fn_index_in, *regression_order_in, 1, &m) != EMUB_OK ||
	    emub_emulator_create(m, thetas_in, &e) != EMUB_OK ||
	    emub_predict_batch(e, Q, d, mq, final_emulated_y, final_emulated_variance) != EMUB_OK) {
		/* the reference exits when the factorisation fails (emulate-fns.c:282-285) */
		fprintf(stderr, "callEmulateAtList: %s\n", emub_last_error());
		exit(EXIT_FAILURE);
	}
	emub_emulator_destroy(e);
	emub_model_destroy(m);
	free(Q); free(X);
}

/* ---- the Monte-Carlo fast path: setupEmulateMC / callEmulateMC / freeEmulateMC (rbind.c:299-455) and their
 * multivariate twins (rbind.c:483-590) --------------------------------------------------------------------------
 * The reference's set-up builds C, C^-1 (n^3, on the host), H and beta into `struct emulateMCData` and every
 * callEmulateMC then runs emulateQuick against that host-side inverse ((6 + 2p) n^2 flops per point, emulate-fns.c:201).
 * Here the set-up is one engine model + one cached factor per output dimension (emub_emulator_create_comp), and a
 * call is the latency path for one point (emub_predict_few / emub_predict_multi_few: all output dimensions in one set
 * of launches).  Same signatures, same pointer-only .C() convention; the handles replace the reference's globals
 * emuMCData / emuMCDataMulti (rbind.h:55,57). */
static emub_model *mc_model = NULL;
static emub_emulator *mc_emu = NULL;
static int mc_d = 0;
static emub_model *mcm_model = NULL;
static emub_emulator **mcm_emus = NULL;
static int mcm_n = 0, mcm_d = 0;

static void mc_die(const char *who)
{
	/* the reference exits when the factorisation fails (emulate-fns.c:282-285) */
	fprintf(stderr, "%s: %s\n", who, emub_last_error());
	exit(EXIT_FAILURE);
}

void freeEmulateMC(void)
{
	if (mc_emu) emub_emulator_destroy(mc_emu);
	if (mc_model) emub_model_destroy(mc_model);
	mc_emu = NULL; mc_model = NULL; mc_d = 0;
}

void setupEmulateMC(double *xmodel_in, int *nparams_in, double *training_in, int *nmodelpts, double *thetas_in, int *nthetas_in,
                    int *cov_fn_index_in, int *regression_order_in)
{
	const int d = *nparams_in, n = *nmodelpts;
	(void)nthetas_in;
	freeEmulateMC(); /* a second set-up replaces the first (the reference leaks it) */
	emub_ctx *ctx = rbind_ctx();
	double *X = to_row_major(xmodel_in, d, n);
	if (!ctx || emub_model_create(ctx, X, d, n, d, training_in, *cov_fn_index_in, *regression_order_in, 1, &mc_model) != EMUB_OK ||
	    emub_emulator_create(mc_model, thetas_in, &mc_emu) != EMUB_OK)
		mc_die("setupEmulateMC");
	mc_d = d;
	free(X);
}

void callEmulateMC(double *point_in, double *mean_out, double *var_out)
{
	if (!mc_emu) { fprintf(stderr, "callEmulateMC: setupEmulateMC has not been called\n"); exit(EXIT_FAILURE); } /* the reference asserts, rbind.c:415-418 */
	if (emub_predict_few(mc_emu, point_in, mc_d, 1, mean_out, var_out) != EMUB_OK) mc_die("callEmulateMC");
}

void freeEmulateMCMulti(int *nydims_in)
{
	(void)nydims_in;
	for (int i = 0; i < mcm_n; i++)
		if (mcm_emus && mcm_emus[i]) emub_emulator_destroy(mcm_emus[i]);
	free(mcm_emus);
	if (mcm_model) emub_model_destroy(mcm_model);
	mcm_emus = NULL; mcm_model = NULL; mcm_n = 0; mcm_d = 0;
}

/* training_in: nmodelpts x nydims, thetas_in: nydims x nthetas, both column-major as R hands them over
 * (convertDoubleToMatrix, rbind.c:506-507); output dimension `index` trains on column index of the training matrix with
 * row index of the thetas (rbind.c:510-522) */
void setupEmulateMCMulti(double *xmodel_in, int *nparams_in, double *training_in, int *nydims_in, int *nmodelpts_in, double *thetas_in,
                         int *nthetas_in, int *cov_fn_index_in, int *regression_order_in)
{
	const int d = *nparams_in, n = *nmodelpts_in, ny = *nydims_in, nth = *nthetas_in;
	freeEmulateMCMulti(nydims_in);
	emub_ctx *ctx = rbind_ctx();
	double *X = to_row_major(xmodel_in, d, n);
	double *Y = to_row_major(training_in, ny, n);  /* n rows of ny values */
	double *TH = to_row_major(thetas_in, nth, ny); /* ny rows of nthetas values */
	if (!ctx || emub_model_create(ctx, X, d, n, d, Y, *cov_fn_index_in, *regression_order_in, 1, &mcm_model) != EMUB_OK ||
	    emub_model_set_training_multi(mcm_model, Y, ny, ny) != EMUB_OK)
		mc_die("setupEmulateMCMulti");
	mcm_emus = (emub_emulator **)calloc((size_t)(ny > 0 ? ny : 1), sizeof(emub_emulator *));
	if (!mcm_emus) { fprintf(stderr, "setupEmulateMCMulti: out of memory\n"); exit(EXIT_FAILURE); }
	mcm_n = ny; mcm_d = d;
	for (int i = 0; i < ny; i++)
		if (emub_emulator_create_comp(mcm_model, i, TH + (size_t)i * nth, &mcm_emus[i]) != EMUB_OK) mc_die("setupEmulateMCMulti");
	free(X); free(Y); free(TH);
}

void callEmulateMCMulti(double *point_in, int *nydims_in, double *final_mean, double *final_var)
{
	const int ny = *nydims_in;
	if (!mcm_emus || ny != mcm_n) { fprintf(stderr, "callEmulateMCMulti: setupEmulateMCMulti has not been called for %d outputs\n", ny); exit(EXIT_FAILURE); }
	/* nt = 0: the per-emulator values, no back-projection (the R side works on the outputs themselves) */
	if (emub_predict_multi_few((emub_emulator *const *)mcm_emus, ny, point_in, mcm_d, 1, 0, NULL, NULL, NULL, final_mean, final_var) != EMUB_OK)
		mc_die("callEmulateMCMulti");
}

void rbind_glue_reset(void)
{
	freeEmulateMC();
	freeEmulateMCMulti(NULL);
	if (rb_ctx) { emub_ctx_destroy(rb_ctx); rb_ctx = NULL; }
}
