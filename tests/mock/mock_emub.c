/*
 * mock_emub.c -- a CPU stand-in for the few C-ABI entry points (include/emu_b200.h) that the host restart driver
 * (madaiemulator_b200/host/emub_estimate.c) calls, so that the driver's own logic -- evaluation fronts, value policy,
 * restart bookkeeping, component sharding -- is tested without a GPU (tests/test_host_front_cpu.py).
 *
 * TEST INFRASTRUCTURE ONLY; never linked into the product.  The "likelihood" is an analytic function with a known
 * minimum per component:   f_c(x) = sum_i w_i (x_i - a_ci)^2 + 0.1 sum_i (x_i - a_ci)^4,   sigma2 = 1 + f.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/emu_b200.h"

struct emub_ctx { int device; };
struct emub_model { int nth, ncomp, kernel, mode, device; };
struct emub_emulator { int unused; };

static long long g_calls = 0, g_points = 0, g_value_points = 0, g_max_b = 0;
static int g_models_alive = 0;

long long mock_calls(void) { return g_calls; }
long long mock_points(void) { return g_points; }
long long mock_value_points(void) { return g_value_points; }
long long mock_max_batch(void) { return g_max_b; }
int mock_models_alive(void) { return g_models_alive; }
void mock_reset(void) { g_calls = g_points = g_value_points = g_max_b = 0; }

static double target(int comp, int i) { return -3.0 + 0.37 * comp + 0.21 * i; }

const char *emub_last_error(void) { return "mock"; }
int emub_ctx_create(int device, emub_ctx **out) { *out = (emub_ctx *)calloc(1, sizeof(emub_ctx)); (*out)->device = device; return EMUB_OK; }
void emub_ctx_destroy(emub_ctx *c) { free(c); }
int emub_model_create(emub_ctx *ctx, const double *X, int ldx, int n, int d, const double *y, int kernel, int order, int max_slots,
                      emub_model **out)
{
	(void)X; (void)ldx; (void)n; (void)y; (void)order; (void)max_slots;
	emub_model *m = (emub_model *)calloc(1, sizeof(emub_model));
	m->kernel = kernel; m->nth = (kernel == EMUB_POWEREXP) ? d + 2 : 3; m->ncomp = 1; m->device = ctx->device;
	__sync_fetch_and_add(&g_models_alive, 1);
	*out = m;
	return EMUB_OK;
}
void emub_model_destroy(emub_model *m) { if (m) { __sync_fetch_and_sub(&g_models_alive, 1); free(m); } }
int emub_model_set_training_multi(emub_model *m, const double *Y, int ldy, int ncomp) { (void)Y; (void)ldy; m->ncomp = ncomp; return EMUB_OK; }
int emub_model_nthetas(const emub_model *m) { return m->nth; }
int emub_model_ncomponents(const emub_model *m) { return m->ncomp; }
int emub_model_kernel(const emub_model *m) { return m->kernel; }
int emub_model_gradient_mode(const emub_model *m) { return m->mode; }
int emub_model_set_gradient_mode(emub_model *m, int mode) { m->mode = mode; return EMUB_OK; }

/* component ids are GLOBAL for the mock: a sharded model is told its first component through the device id trick below,
 * so tests pass an explicit offset table instead */
static int g_comp_offset[64], g_comp_stride[64];
void mock_set_component_map(int device, int first, int stride) { g_comp_offset[device & 63] = first; g_comp_stride[device & 63] = stride; }

int emub_loglik_grad_batch_mixed(emub_model *m, const double *thetas, const int *comp, const int *want_grad, int B, double *negL,
                                 double *grad, double *sigma2, int *status)
{
	const int nth1 = m->nth - 1;
	__sync_fetch_and_add(&g_calls, 1);
	__sync_fetch_and_add(&g_points, B);
	if (B > g_max_b) g_max_b = B;
	const int first = g_comp_offset[m->device & 63], stride = g_comp_stride[m->device & 63] ? g_comp_stride[m->device & 63] : 1;
	for (int b = 0; b < B; b++) {
		const int c = first + (comp ? comp[b] : 0) * stride;
		double f = 0.0;
		for (int i = 0; i < nth1; i++) {
			const double dx = thetas[(size_t)b * nth1 + i] - target(c, i), w = 1.0 + 0.5 * i;
			f += w * dx * dx + 0.1 * dx * dx * dx * dx;
			grad[(size_t)b * nth1 + i] = want_grad[b] ? 2.0 * w * dx + 0.4 * dx * dx * dx : 0.0;
		}
		if (!want_grad[b]) __sync_fetch_and_add(&g_value_points, 1);
		negL[b] = f; sigma2[b] = 1.0 + f; status[b] = EMUB_OK;
	}
	return EMUB_OK;
}
