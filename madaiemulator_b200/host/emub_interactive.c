/* emub_interactive.c -- see emub_interactive.h.  Citations are file:line under the reference's src/. */
#include "emub_interactive.h"
#include <poll.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

struct emub_multi_emulator {
	emub_model *model;
	emub_emulator **emus;
	int nt, nr, d;
	double *training_mean, *evecs, *evals;
	/* replicas on further devices (query blocks are independent: SURVEY 8e); replica 0 is this struct */
	int nreplicas;
	struct emub_multi_emulator **replicas; /* [nreplicas - 1] */
	emub_ctx **owned_ctx;                  /* contexts created by _from_snapshot_devices */
	int nowned;
};

int emub_multi_emulator_from_snapshot(emub_ctx *ctx, const emub_snapshot *s, emub_multi_emulator **out)
{
	if (!ctx || !s || !out) return EMUB_EINVAL;
	const int n = s->nmodel_points, d = s->nparams, nr = s->nr, nt = s->nt;
	emub_multi_emulator *me = (emub_multi_emulator *)calloc(1, sizeof(*me));
	me->nt = nt; me->nr = nr; me->d = d;
	/* every component block repeats the design; the reference emulates from the block's own copy (modelstruct.c:437-441) */
	const emub_snapshot_component *c0 = &s->components[0];
	int rc = emub_model_create(ctx, c0->xmodel, d, n, d, c0->training_vector, c0->cov_fn_index, c0->regression_order, 1, &me->model);
	if (rc != EMUB_OK) { free(me); return rc; }
	double *Y = (double *)malloc(sizeof(double) * (size_t)n * nr);
	for (int c = 0; c < nr; c++)
		for (int i = 0; i < n; i++) Y[(size_t)i * nr + c] = s->components[c].training_vector[i];
	rc = emub_model_set_training_multi(me->model, Y, nr, nr);
	free(Y);
	if (rc != EMUB_OK) { emub_model_destroy(me->model); free(me); return rc; }
	me->emus = (emub_emulator **)calloc((size_t)nr, sizeof(emub_emulator *));
	for (int c = 0; c < nr; c++) {
		rc = emub_emulator_create_comp(me->model, c, s->components[c].thetas, &me->emus[c]);
		if (rc != EMUB_OK) { emub_multi_emulator_destroy(me); return rc; }
	}
	me->training_mean = (double *)malloc(sizeof(double) * (size_t)nt);
	me->evecs = (double *)malloc(sizeof(double) * (size_t)nt * nr);
	me->evals = (double *)malloc(sizeof(double) * (size_t)nr);
	memcpy(me->training_mean, s->training_mean, sizeof(double) * (size_t)nt);
	memcpy(me->evecs, s->pca_evecs_r, sizeof(double) * (size_t)nt * nr);
	memcpy(me->evals, s->pca_evals_r, sizeof(double) * (size_t)nr);
	*out = me;
	return EMUB_OK;
}

int emub_multi_emulator_from_snapshot_devices(const int *devices, int ndev, const emub_snapshot *s, emub_multi_emulator **out)
{
	if (!devices || ndev < 1 || !s || !out) return EMUB_EINVAL;
	emub_ctx **ctxs = (emub_ctx **)calloc((size_t)ndev, sizeof(emub_ctx *));
	emub_multi_emulator **mes = (emub_multi_emulator **)calloc((size_t)ndev, sizeof(emub_multi_emulator *));
	int rc = EMUB_OK;
	for (int g = 0; g < ndev && rc == EMUB_OK; g++) {
		rc = emub_ctx_create(devices[g], &ctxs[g]);
		if (rc == EMUB_OK) rc = emub_multi_emulator_from_snapshot(ctxs[g], s, &mes[g]);
	}
	if (rc != EMUB_OK) {
		for (int g = 0; g < ndev; g++) { emub_multi_emulator_destroy(mes[g]); if (ctxs[g]) emub_ctx_destroy(ctxs[g]); }
		free(ctxs); free(mes);
		return rc;
	}
	emub_multi_emulator *me = mes[0];
	me->nreplicas = ndev;
	me->replicas = (emub_multi_emulator **)calloc((size_t)ndev, sizeof(emub_multi_emulator *));
	for (int g = 1; g < ndev; g++) me->replicas[g - 1] = mes[g];
	me->owned_ctx = ctxs;
	me->nowned = ndev;
	free(mes);
	*out = me;
	return EMUB_OK;
}

void emub_multi_emulator_destroy(emub_multi_emulator *me)
{
	if (!me) return;
	for (int g = 1; g < me->nreplicas; g++) emub_multi_emulator_destroy(me->replicas[g - 1]);
	free(me->replicas);
	emub_ctx **owned = me->owned_ctx;
	const int nowned = me->nowned;
	if (me->emus)
		for (int c = 0; c < me->nr; c++) emub_emulator_destroy(me->emus[c]);
	free(me->emus);
	emub_model_destroy(me->model);
	free(me->training_mean); free(me->evecs); free(me->evals);
	free(me);
	for (int g = 0; g < nowned; g++) emub_ctx_destroy(owned[g]);
	free(owned);
}
int emub_multi_emulator_nt(const emub_multi_emulator *me) { return me ? me->nt : 0; }
int emub_multi_emulator_nr(const emub_multi_emulator *me) { return me ? me->nr : 0; }
int emub_multi_emulator_nparams(const emub_multi_emulator *me) { return me ? me->d : 0; }

static int predict_one(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var);

typedef struct { emub_multi_emulator *me; const double *pts; int m, pca; double *mean, *var; int rc; } shard_job;
static void *shard_main(void *arg)
{
	shard_job *j = (shard_job *)arg;
	j->rc = predict_one(j->me, j->pts, j->m, j->pca, j->mean, j->var);
	return NULL;
}

/* contiguous blocks of the query list, one per device replica, run concurrently (one host thread per GPU) */
int emub_multi_emulator_predict(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var)
{
	if (!me || !pts || !mean || !var || m < 0) return EMUB_EINVAL;
	const int G = me->nreplicas > 1 ? me->nreplicas : 1;
	if (G == 1 || m < 2 * G) return predict_one(me, pts, m, pca_output, mean, var);
	shard_job jobs[64];
	pthread_t th[64];
	const int base = m / G, rem = m % G;
	int lo = 0;
	for (int g = 0; g < G; g++) {
		const int cnt = base + (g < rem ? 1 : 0);
		jobs[g].me = g == 0 ? me : me->replicas[g - 1];
		jobs[g].pts = pts + (size_t)lo * me->d; jobs[g].m = cnt; jobs[g].pca = pca_output;
		jobs[g].mean = mean + (size_t)lo * me->nt; jobs[g].var = var + (size_t)lo * me->nt; jobs[g].rc = EMUB_OK;
		lo += cnt;
		pthread_create(&th[g], NULL, shard_main, &jobs[g]);
	}
	int rc = EMUB_OK;
	for (int g = 0; g < G; g++) { pthread_join(th[g], NULL); if (jobs[g].rc != EMUB_OK) rc = jobs[g].rc; }
	return rc;
}

static int predict_one(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var)
{
	if (!pca_output)
		return emub_predict_multi(me->emus, me->nr, pts, me->d, m, me->nt, me->training_mean, me->evecs, me->evals, mean, var);
	/* PCA space (multivar_support.c:78-101): nr values per point, laid out in rows of nt like the reference's vectors */
	double *pm = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1) * me->nr);
	double *pv = (double *)malloc(sizeof(double) * (size_t)(m ? m : 1) * me->nr);
	int rc = emub_predict_multi(me->emus, me->nr, pts, me->d, m, 0, NULL, NULL, NULL, pm, pv);
	if (rc == EMUB_OK)
		for (int q = 0; q < m; q++)
			for (int i = 0; i < me->nt; i++) {
				mean[(size_t)q * me->nt + i] = i < me->nr ? pm[(size_t)q * me->nr + i] : 0.0;
				var[(size_t)q * me->nt + i] = i < me->nr ? pv[(size_t)q * me->nr + i] : 0.0;
			}
	free(pm); free(pv);
	return rc;
}

/* parse as many doubles as are complete in [p, end); returns the number parsed and advances *pp.  A token that
 * touches `end` may be incomplete unless at_eof. */
static size_t parse_doubles(char **pp, char *end, int at_eof, double *out, size_t max)
{
	char *p = *pp;
	size_t k = 0;
	while (k < max) {
		while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == ',')) p++;
		if (p >= end) break;
		char *q = p;
		while (q < end && !(*q == ' ' || *q == '\n' || *q == '\t' || *q == '\r' || *q == ',')) q++;
		if (q == end && !at_eof) break; /* token may continue in the next block */
		char *stop;
		double v = strtod(p, &stop);
		if (stop == p) { p = end; break; } /* not a number: the reference's fscanf would stop here too */
		out[k++] = v;
		p = stop > q ? stop : q;
	}
	*pp = p;
	return k;
}

int emub_interactive_stream(emub_multi_emulator *me, FILE *in, FILE *out, int quiet, int pca_output, int binary,
                            int block_points, long long *npoints)
{
	if (!me || !in || !out) return EMUB_EINVAL;
	const int d = me->d, nt = me->nt;
	if (block_points <= 0) block_points = 16384 * (me->nreplicas > 1 ? me->nreplicas : 1);
	long long total = 0;
	if (!quiet) { /* interactive_emulator.c:398-414 */
		fprintf(out, "%d\n", d);
		for (int i = 0; i < d; i++) fprintf(out, "%s%d\n", "param_", i);
		fprintf(out, "%d\n", 2 * nt);
		for (int i = 0; i < nt; i++) fprintf(out, "%s_%d\n%s_%d\n", "mean", i, "variance", i);
		fflush(out);
	}
	const size_t maxvals = (size_t)block_points * d;
	double *pts = (double *)malloc(sizeof(double) * maxvals);
	double *mean = (double *)malloc(sizeof(double) * (size_t)block_points * nt);
	double *var = (double *)malloc(sizeof(double) * (size_t)block_points * nt);
	const size_t obuf_cap = (size_t)block_points * nt * 2 * 48 + 64;
	char *obuf = (char *)malloc(obuf_cap);
	const size_t ibuf_cap = 1 << 22;
	char *ibuf = (char *)malloc(ibuf_cap + 1);
	size_t ilen = 0, nvals = 0, bin_bytes = 0;
	int eof = 0, rc = EMUB_OK;
	const int fd = fileno(in);
	while (rc == EMUB_OK) {
		/* Fill the point block.  Input is taken with read(2) as it becomes available: a block is closed as soon as it
		 * is full, the input ends, or at least one complete point is in hand and nothing more is waiting -- so a
		 * client that sends one point and waits for its answer (the reference flushes per point) is still served. */
		for (;;) {
			if (!binary) {
				char *p = ibuf;
				nvals += parse_doubles(&p, ibuf + ilen, eof, pts + nvals, maxvals - nvals);
				const size_t rest = (size_t)(ibuf + ilen - p);
				memmove(ibuf, p, rest);
				ilen = rest;
			}
			if (nvals >= maxvals || eof) break;
			if (!binary && ilen == ibuf_cap) break; /* a single token longer than the buffer: give up on it */
			if (nvals >= (size_t)d) {
				struct pollfd pfd = {fd, POLLIN, 0};
				if (poll(&pfd, 1, 0) <= 0) break;
			}
			if (binary) {
				ssize_t r = read(fd, (char *)pts + bin_bytes, maxvals * sizeof(double) - bin_bytes);
				if (r <= 0) eof = 1; else bin_bytes += (size_t)r;
				nvals = bin_bytes / sizeof(double);
			} else {
				ssize_t r = read(fd, ibuf + ilen, ibuf_cap - ilen);
				if (r <= 0) eof = 1; else ilen += (size_t)r;
			}
		}
		const int m = (int)(nvals / (size_t)d); /* an incomplete trailing point is dropped, like the reference (:427-428) */
		if (m > 0) {
			rc = emub_multi_emulator_predict(me, pts, m, pca_output, mean, var);
			if (rc != EMUB_OK) break;
			if (binary) {
				for (int q = 0; q < m; q++)
					for (int i = 0; i < nt; i++) {
						fwrite(&mean[(size_t)q * nt + i], sizeof(double), 1, out);
						fwrite(&var[(size_t)q * nt + i], sizeof(double), 1, out);
					}
			} else {
				size_t o = 0;
				for (int q = 0; q < m; q++)
					for (int i = 0; i < nt; i++) { /* interactive_emulator.c:431-437 */
						o += (size_t)snprintf(obuf + o, obuf_cap - o, "%.17f\n", mean[(size_t)q * nt + i]);
						o += (size_t)snprintf(obuf + o, obuf_cap - o, "%.17f\n", var[(size_t)q * nt + i]);
						if (obuf_cap - o < 1024) { fwrite(obuf, 1, o, out); o = 0; }
					}
				fwrite(obuf, 1, o, out);
			}
			fflush(out);
			total += m;
		}
		const size_t used = (size_t)m * d;
		if (binary) {
			memmove(pts, (char *)pts + used * sizeof(double), bin_bytes - used * sizeof(double));
			bin_bytes -= used * sizeof(double);
			nvals = bin_bytes / sizeof(double);
		} else {
			memmove(pts, pts + used, sizeof(double) * (nvals - used));
			nvals -= used;
		}
		if (eof && (binary || ilen == 0 || m == 0)) break;
	}
	free(pts); free(mean); free(var); free(obuf); free(ibuf);
	if (npoints) *npoints = total;
	return rc;
}
