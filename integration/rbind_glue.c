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
 * Pure pointer marshalling: no R headers, no GSL.  Link into libRBIND in place of the two reference bodies
 * (the other entry points keep working through integration/libemu_glue.c).
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

void rbind_glue_reset(void)
{
	if (rb_ctx) { emub_ctx_destroy(rb_ctx); rb_ctx = NULL; }
}
