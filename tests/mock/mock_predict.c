/*
 * mock_predict.c -- a CPU stand-in for the C-ABI entry points (include/emu_b200.h) that the streaming interactive_mode
 * layer (madaiemulator_b200/host/emub_interactive.c, emub_cli.c) calls, so that the text / binary protocol, the block
 * pipeline and the multi-device split are tested without a GPU (tests/test_interactive_cpu.py).
 *
 * TEST INFRASTRUCTURE ONLY; never linked into the product.  The "emulator" is an analytic function of the point that
 * the test recomputes in the same operation order:  m_j = (sum_k x_k (k+1)) (j+1) + theta_j1,  v_j = 0.5 (j+1) + x_0^2,
 * then the reference's back-projection (multivar_support.c:126-151).
 */
#include <math.h>
#include <stdlib.h>
#include "../../include/emu_b200.h"

struct emub_ctx { int device; };
struct emub_model { int d, n, ncomp, nth; };
struct emub_emulator { int comp, d; double th1; };

const char *emub_last_error(void) { return "mock"; }
int emub_ctx_create(int device, emub_ctx **out) { *out = (emub_ctx *)calloc(1, sizeof(emub_ctx)); (*out)->device = device; return EMUB_OK; }
void emub_ctx_destroy(emub_ctx *c) { free(c); }
int emub_model_create(emub_ctx *ctx, const double *X, int ldx, int n, int d, const double *y, int kernel, int order, int max_slots,
                      emub_model **out)
{
	(void)ctx; (void)X; (void)ldx; (void)y; (void)order; (void)max_slots;
	emub_model *m = (emub_model *)calloc(1, sizeof(emub_model));
	m->d = d; m->n = n; m->ncomp = 1; m->nth = (kernel == EMUB_POWEREXP) ? d + 2 : 3;
	*out = m;
	return EMUB_OK;
}
void emub_model_destroy(emub_model *m) { free(m); }
int emub_model_nthetas(const emub_model *m) { return m->nth; }
int emub_model_set_training_multi(emub_model *m, const double *Y, int ldy, int ncomp) { (void)Y; (void)ldy; m->ncomp = ncomp; return EMUB_OK; }
int emub_emulator_create_comp(emub_model *m, int comp, const double *thetas, emub_emulator **out)
{
	emub_emulator *e = (emub_emulator *)calloc(1, sizeof(emub_emulator));
	e->comp = comp; e->d = m->d; e->th1 = thetas[1];
	*out = e;
	return EMUB_OK;
}
void emub_emulator_destroy(emub_emulator *e) { free(e); }

int emub_predict_multi(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt, const double *training_mean,
                       const double *evecs, const double *evals, double *mean, double *var)
{
	const int d = emus[0]->d;
	double *pm = (double *)malloc(sizeof(double) * (size_t)nr), *pv = (double *)malloc(sizeof(double) * (size_t)nr);
	for (int q = 0; q < mq; q++) {
		const double *x = pts + (size_t)q * ldp;
		double s = 0.0;
		for (int k = 0; k < d; k++) s += x[k] * (double)(k + 1);
		for (int j = 0; j < nr; j++) {
			pm[j] = s * (double)(j + 1) + emus[j]->th1;
			pv[j] = 0.5 * (double)(j + 1) + x[0] * x[0];
		}
		if (nt == 0) {
			for (int j = 0; j < nr; j++) { mean[(size_t)q * nr + j] = pm[j]; var[(size_t)q * nr + j] = pv[j]; }
			continue;
		}
		for (int i = 0; i < nt; i++) {
			double a = training_mean[i], b = 0.0;
			for (int j = 0; j < nr; j++) {
				a += evecs[(size_t)i * nr + j] * sqrt(evals[j]) * pm[j];
				b += evecs[(size_t)i * nr + j] * evecs[(size_t)i * nr + j] * evals[j] * pv[j];
			}
			mean[(size_t)q * nt + i] = a;
			var[(size_t)q * nt + i] = b;
		}
	}
	free(pm); free(pv);
	return EMUB_OK;
}

int emub_predict_multi_few(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt, const double *training_mean,
                           const double *evecs, const double *evals, double *mean, double *var)
{
	if (mq > 8) return EMUB_EINVAL;
	return emub_predict_multi(emus, nr, pts, ldp, mq, nt, training_mean, evecs, evals, mean, var);
}
