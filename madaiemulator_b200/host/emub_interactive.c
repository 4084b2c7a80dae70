/* emub_interactive.c -- see emub_interactive.h.  Citations are file:line under the reference's src/. */
#define _GNU_SOURCE
#include "emub_interactive.h"
#include "emub_fastfloat.h"
#include <fcntl.h>
#include <poll.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
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
	if (!ctx || !s || !out || s->nr < 1 || s->nt < 1) return EMUB_EINVAL;
	const int n = s->nmodel_points, d = s->nparams, nr = s->nr, nt = s->nt;
	emub_multi_emulator *me = (emub_multi_emulator *)calloc(1, sizeof(*me));
	me->nt = nt; me->nr = nr; me->d = d;
	/* every component block repeats the design; the reference emulates from the block's own copy (modelstruct.c:437-441) */
	const emub_snapshot_component *c0 = &s->components[0];
	int rc = emub_model_create(ctx, c0->xmodel, d, n, d, c0->training_vector, c0->cov_fn_index, c0->regression_order, 1, &me->model);
	if (rc != EMUB_OK) { free(me); return rc; }
	/* emub_emulator_create_comp copies nthetas(kernel) values out of every component's thetas */
	for (int c = 0; c < nr; c++)
		if (s->components[c].nthetas != emub_model_nthetas(me->model)) { emub_model_destroy(me->model); free(me); return EMUB_EINVAL; }
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

typedef struct { int device; const emub_snapshot *s; emub_ctx *ctx; emub_multi_emulator *me; int rc; } replica_job;
static void *replica_main(void *arg)
{
	replica_job *j = (replica_job *)arg;
	j->rc = emub_ctx_create(j->device, &j->ctx);
	if (j->rc == EMUB_OK) j->rc = emub_multi_emulator_from_snapshot(j->ctx, j->s, &j->me);
	return NULL;
}

int emub_multi_emulator_from_snapshot_devices(const int *devices, int ndev, const emub_snapshot *s, emub_multi_emulator **out)
{
	if (!devices || ndev < 1 || ndev > 64 || !s || !out) return EMUB_EINVAL;
	emub_ctx **ctxs = (emub_ctx **)calloc((size_t)ndev, sizeof(emub_ctx *));
	emub_multi_emulator **mes = (emub_multi_emulator **)calloc((size_t)ndev, sizeof(emub_multi_emulator *));
	int rc = EMUB_OK;
	/* the replicas are set up concurrently, one host thread per device (context creation + factorisation each) */
	replica_job jobs[64];
	pthread_t th[64];
	int started[64];
	for (int g = 0; g < ndev; g++) {
		jobs[g].device = devices[g]; jobs[g].s = s; jobs[g].ctx = NULL; jobs[g].me = NULL; jobs[g].rc = EMUB_OK;
		started[g] = g > 0 && pthread_create(&th[g], NULL, replica_main, &jobs[g]) == 0;
	}
	for (int g = 0; g < ndev; g++) {
		if (started[g]) pthread_join(th[g], NULL);
		else replica_main(&jobs[g]);
	}
	for (int g = 0; g < ndev; g++) {
		ctxs[g] = jobs[g].ctx; mes[g] = jobs[g].me;
		if (jobs[g].rc != EMUB_OK && rc == EMUB_OK) rc = jobs[g].rc;
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
	int started[64];
	const int base = m / G, rem = m % G;
	int lo = 0;
	for (int g = 0; g < G; g++) {
		const int cnt = base + (g < rem ? 1 : 0);
		jobs[g].me = g == 0 ? me : me->replicas[g - 1];
		jobs[g].pts = pts + (size_t)lo * me->d; jobs[g].m = cnt; jobs[g].pca = pca_output;
		jobs[g].mean = mean + (size_t)lo * me->nt; jobs[g].var = var + (size_t)lo * me->nt; jobs[g].rc = EMUB_OK;
		lo += cnt;
		started[g] = g > 0 && pthread_create(&th[g], NULL, shard_main, &jobs[g]) == 0;
	}
	int rc = EMUB_OK;
	for (int g = 0; g < G; g++) {
		if (started[g]) pthread_join(th[g], NULL);
		else shard_main(&jobs[g]); /* share 0, and any share whose thread could not be created, on the caller's thread */
		if (jobs[g].rc != EMUB_OK) rc = jobs[g].rc;
	}
	return rc;
}

/* emulate_point_multi[_pca] for a handful of points (m <= 8): the latency path on the first device */
int emub_multi_emulator_predict_few(emub_multi_emulator *me, const double *pts, int m, int pca_output, double *mean, double *var)
{
	if (!me || !pts || !mean || !var || m < 0 || m > 8) return EMUB_EINVAL;
	if (!pca_output)
		return emub_predict_multi_few(me->emus, me->nr, pts, me->d, m, me->nt, me->training_mean, me->evecs, me->evals, mean, var);
	double pm[8 * 64], pv[8 * 64];
	if (me->nr > 64) return emub_multi_emulator_predict(me, pts, m, pca_output, mean, var);
	int rc = emub_predict_multi_few(me->emus, me->nr, pts, me->d, m, 0, NULL, NULL, NULL, pm, pv);
	if (rc == EMUB_OK)
		for (int q = 0; q < m; q++)
			for (int i = 0; i < me->nt; i++) {
				mean[(size_t)q * me->nt + i] = i < me->nr ? pm[(size_t)q * me->nr + i] : 0.0;
				var[(size_t)q * me->nt + i] = i < me->nr ? pv[(size_t)q * me->nr + i] : 0.0;
			}
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

/* ---- streaming interactive_mode ----------------------------------------------------------------------------
 * Three pipeline stages over a ring of block slots, each on its own host thread:
 *   reader (caller's thread): read(2) + text -> doubles, the conversion spread over IO worker threads
 *   device thread           : emub_multi_emulator_predict on the block (all replicas)
 *   writer thread           : "%.17f\n" formatting spread over IO worker threads, ordered write, flush
 * so the strtod / printf cost (about 0.3 us per value on one core, i.e. more than the GPU's time per point at
 * n = 4096) is both parallel and hidden behind the device stage.  Block order is the input order. */
static int is_sep(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == ','; }

typedef struct { void (*fn)(void *, int); void *arg; int part; } par_job;
static void *par_main(void *a) { par_job *j = (par_job *)a; j->fn(j->arg, j->part); return NULL; }
/* fn(arg, part) for part in [0, nparts) on nparts threads (the caller runs part 0) */
static void parallel_for(int nparts, void (*fn)(void *, int), void *arg)
{
	pthread_t th[64];
	par_job jobs[64];
	int started[64];
	if (nparts > 64) nparts = 64;
	for (int p = 1; p < nparts; p++) {
		jobs[p].fn = fn; jobs[p].arg = arg; jobs[p].part = p;
		started[p] = pthread_create(&th[p], NULL, par_main, &jobs[p]) == 0;
	}
	fn(arg, 0);
	for (int p = 1; p < nparts; p++) {
		if (started[p]) pthread_join(th[p], NULL);
		else fn(arg, p);
	}
}

static int io_threads(void)
{
	const char *e = getenv("EMUB_IO_THREADS");
	long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN) - 2;
	if (t < 1) t = 1;
	if (t > 32) t = 32;
	return (int)t;
}

/* text -> doubles.  buf[0, len) ends on a separator (or is the end of the input); buf[len] is readable.
 * Two passes over nseg segments cut at separators: count the tokens, then convert token g into out[g] for g < max. */
typedef struct {
	char *buf;
	size_t seg[65];    /* segment boundaries */
	size_t count[64];  /* tokens in the segment */
	size_t first[64];  /* global index of its first token */
	size_t stop[64];   /* offset of the first byte the segment did not consume */
	size_t bad[64];    /* global index of a token that is not a number, or (size_t)-1 */
	double *out;
	size_t max;
} parse_job;

static void parse_count(void *arg, int s)
{
	parse_job *j = (parse_job *)arg;
	const char *p = j->buf + j->seg[s], *end = j->buf + j->seg[s + 1];
	size_t k = 0;
	while (p < end) {
		while (p < end && is_sep(*p)) p++;
		if (p >= end) break;
		k++;
		while (p < end && !is_sep(*p)) p++;
	}
	j->count[s] = k;
}

static void parse_convert(void *arg, int s)
{
	parse_job *j = (parse_job *)arg;
	char *p = j->buf + j->seg[s], *end = j->buf + j->seg[s + 1];
	size_t g = j->first[s];
	j->bad[s] = (size_t)-1;
	while (p < end && g < j->max) {
		while (p < end && is_sep(*p)) p++;
		if (p >= end) break;
		char *q = p;
		while (q < end && !is_sep(*q)) q++;
		double v;
		if (!emub_fast_strtod(p, q, &v)) { /* plain decimal tokens take the exact fast conversion, the rest strtod */
			char *stop;
			v = strtod(p, &stop);
			/* not a number, or a number with a tail ("1.5abc"): the reference's fscanf("%lf%*c") takes 1.5, eats one
			 * character and stops the stream at the next conversion; here the stream ends at the token */
			if (stop != q) { j->bad[s] = g; break; }
		}
		j->out[g++] = v;
		p = q;
	}
	j->stop[s] = (size_t)(p - j->buf);
}

/* converts up to max values from buf[0, len); returns how many, *consumed = bytes used up, *bad = 1 if a token
 * that is not a number ended the conversion (everything from it on is dropped by the caller) */
size_t emub_parse_doubles(char *buf, size_t len, double *out, size_t max, int threads, size_t *consumed, int *bad)
{
	const char keep = buf[len];
	buf[len] = 0; /* strtod must not run past the text */
	parse_job j;
	int nseg = len >= ((size_t)1 << 16) ? threads : 1;
	if (nseg > 64) nseg = 64;
	if (nseg < 1) nseg = 1;
	j.buf = buf; j.out = out; j.max = max;
	j.seg[0] = 0;
	for (int s = 1; s < nseg; s++) {
		size_t c = len * (size_t)s / (size_t)nseg;
		if (c < j.seg[s - 1]) c = j.seg[s - 1];
		while (c < len && !is_sep(buf[c])) c++; /* cut on a separator so that no token straddles two segments */
		j.seg[s] = c;
	}
	j.seg[nseg] = len;
	parallel_for(nseg, parse_count, &j);
	size_t total = 0;
	for (int s = 0; s < nseg; s++) { j.first[s] = total; total += j.count[s]; }
	parallel_for(nseg, parse_convert, &j);
	*bad = 0;
	size_t nvals = total < max ? total : max;
	*consumed = len;
	for (int s = 0; s < nseg; s++) {
		if (j.bad[s] != (size_t)-1) { *bad = 1; if (j.bad[s] < nvals) nvals = j.bad[s]; *consumed = len; break; }
		if (j.first[s] + j.count[s] > max) { *consumed = j.stop[s]; break; } /* the block filled up inside this segment */
	}
	buf[len] = keep;
	return nvals;
}

/* "%.17f\n" lines (interactive_emulator.c:431-437) for points [q0, q1) of a block, into a growable buffer per part */
typedef struct {
	const double *mean, *var;
	int m, nt, nparts;
	char *buf[64];
	size_t len[64], cap[64];
} format_job;

static void format_part(void *arg, int part)
{
	format_job *j = (format_job *)arg;
	const int q0 = (int)((long long)j->m * part / j->nparts), q1 = (int)((long long)j->m * (part + 1) / j->nparts);
	size_t o = 0, cap = j->cap[part];
	char *b = j->buf[part];
	for (int q = q0; q < q1; q++)
		for (int i = 0; i < j->nt; i++) {
			if (cap - o < 1024) { /* a double prints in at most 1 + 309 + 1 + 17 + 1 characters */
				cap = cap * 2 + 4096;
				b = (char *)realloc(b, cap);
			}
			const double pair[2] = {j->mean[(size_t)q * j->nt + i], j->var[(size_t)q * j->nt + i]};
			for (int h = 0; h < 2; h++) { /* exact fast formatter, snprintf for what it declines (inf, nan, |x| >= 2^63) */
				int len = emub_fast_format17(pair[h], b + o);
				if (len == 0) len = snprintf(b + o, cap - o, "%.17f\n", pair[h]);
				o += (size_t)len;
			}
		}
	j->buf[part] = b; j->cap[part] = cap; j->len[part] = o;
}

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

enum { SLOT_FREE = 0, SLOT_PARSED = 1, SLOT_PREDICTED = 2, NSLOT = 3 };
typedef struct {
	double *pts, *mean, *var;
	int m, last, state;
} stream_slot;
typedef struct {
	emub_multi_emulator *me;
	FILE *out;
	int pca_output, binary, threads;
	stream_slot slot[NSLOT];
	pthread_mutex_t mu;
	pthread_cond_t cv;
	int rc;
	long long total;
	/* seconds spent per stage (EMUB_STREAM_STATS=1 prints them to stderr) */
	double t_read, t_parse, t_reader_wait, t_predict, t_device_wait, t_format, t_write, t_writer_wait;
} stream_pipe;

static void slot_wait(stream_pipe *sp, stream_slot *s, int state)
{
	pthread_mutex_lock(&sp->mu);
	while (s->state != state) pthread_cond_wait(&sp->cv, &sp->mu);
	pthread_mutex_unlock(&sp->mu);
}
static int pipe_ok(stream_pipe *sp)
{
	pthread_mutex_lock(&sp->mu);
	const int ok = sp->rc == EMUB_OK;
	pthread_mutex_unlock(&sp->mu);
	return ok;
}
static void slot_set(stream_pipe *sp, stream_slot *s, int state)
{
	pthread_mutex_lock(&sp->mu);
	s->state = state;
	pthread_cond_broadcast(&sp->cv);
	pthread_mutex_unlock(&sp->mu);
}

static void *device_stage(void *arg)
{
	stream_pipe *sp = (stream_pipe *)arg;
	for (int i = 0;; i = (i + 1) % NSLOT) {
		stream_slot *s = &sp->slot[i];
		double t0 = now_s();
		slot_wait(sp, s, SLOT_PARSED);
		sp->t_device_wait += now_s() - t0;
		if (s->m > 0 && pipe_ok(sp)) {
			t0 = now_s();
			const int rc = emub_multi_emulator_predict(sp->me, s->pts, s->m, sp->pca_output, s->mean, s->var);
			sp->t_predict += now_s() - t0;
			if (rc != EMUB_OK) {
				/* the library's message is per thread: report it from the thread that has it */
				fprintf(stderr, "emub_interactive_stream: prediction failed: %s\n", emub_last_error());
				pthread_mutex_lock(&sp->mu); sp->rc = rc; pthread_mutex_unlock(&sp->mu);
			}
		}
		const int last = s->last;
		slot_set(sp, s, SLOT_PREDICTED);
		if (last) return NULL;
	}
}

static void *writer_stage(void *arg)
{
	stream_pipe *sp = (stream_pipe *)arg;
	const int nt = sp->me->nt;
	format_job fj;
	memset(&fj, 0, sizeof(fj));
	double *inter = NULL; /* binary: mean / variance interleaved */
	size_t inter_cap = 0;
	for (int i = 0;; i = (i + 1) % NSLOT) {
		stream_slot *s = &sp->slot[i];
		double t0 = now_s();
		slot_wait(sp, s, SLOT_PREDICTED);
		sp->t_writer_wait += now_s() - t0;
		if (s->m > 0 && pipe_ok(sp)) {
			t0 = now_s();
			if (sp->binary) {
				const size_t cnt = (size_t)s->m * nt;
				if (cnt * 2 > inter_cap) { inter_cap = cnt * 2; inter = (double *)realloc(inter, sizeof(double) * inter_cap); }
				for (size_t k = 0; k < cnt; k++) { inter[2 * k] = s->mean[k]; inter[2 * k + 1] = s->var[k]; }
				sp->t_format += now_s() - t0;
				t0 = now_s();
				fwrite(inter, sizeof(double), cnt * 2, sp->out);
			} else {
				fj.mean = s->mean; fj.var = s->var; fj.m = s->m; fj.nt = nt;
				fj.nparts = (size_t)s->m * nt >= 4096 ? sp->threads : 1;
				parallel_for(fj.nparts, format_part, &fj);
				sp->t_format += now_s() - t0;
				t0 = now_s();
				for (int p = 0; p < fj.nparts; p++) fwrite(fj.buf[p], 1, fj.len[p], sp->out);
			}
			fflush(sp->out);
			sp->t_write += now_s() - t0;
			sp->total += s->m;
		}
		const int last = s->last;
		slot_set(sp, s, SLOT_FREE);
		if (last) break;
	}
	for (int p = 0; p < 64; p++) free(fj.buf[p]);
	free(inter);
	return NULL;
}

static void stream_pipe_free(stream_pipe *sp)
{
	for (int i = 0; i < NSLOT; i++) { free(sp->slot[i].pts); free(sp->slot[i].mean); free(sp->slot[i].var); }
	pthread_mutex_destroy(&sp->mu);
	pthread_cond_destroy(&sp->cv);
}

int emub_interactive_stream(emub_multi_emulator *me, FILE *in, FILE *out, int quiet, int pca_output, int binary,
                            int block_points, long long *npoints)
{
	if (!me || !in || !out) return EMUB_EINVAL;
	const int d = me->d, nt = me->nt;
	if (block_points <= 0) {
		/* four device chunks per GPU, fewer when the model has so many observables that the answer blocks get large */
		long per_dev = (4L << 20) / (nt > 0 ? nt : 1);
		if (per_dev > 4 * 16384) per_dev = 4 * 16384;
		if (per_dev < 16384) per_dev = 16384;
		block_points = (int)per_dev * (me->nreplicas > 1 ? me->nreplicas : 1);
	}
	if (!quiet) { /* interactive_emulator.c:398-414 */
		fprintf(out, "%d\n", d);
		for (int i = 0; i < d; i++) fprintf(out, "%s%d\n", "param_", i);
		fprintf(out, "%d\n", 2 * nt);
		for (int i = 0; i < nt; i++) fprintf(out, "%s_%d\n%s_%d\n", "mean", i, "variance", i);
		fflush(out);
	}
	const size_t maxvals = (size_t)block_points * d;
	stream_pipe sp;
	memset(&sp, 0, sizeof(sp));
	sp.me = me; sp.out = out; sp.pca_output = pca_output; sp.binary = binary; sp.threads = io_threads(); sp.rc = EMUB_OK;
	pthread_mutex_init(&sp.mu, NULL);
	pthread_cond_init(&sp.cv, NULL);
	for (int i = 0; i < NSLOT; i++) {
		sp.slot[i].pts = (double *)malloc(sizeof(double) * maxvals);
		sp.slot[i].mean = (double *)malloc(sizeof(double) * (size_t)block_points * nt);
		sp.slot[i].var = (double *)malloc(sizeof(double) * (size_t)block_points * nt);
	}
	pthread_t dev_th, wr_th;
	if (pthread_create(&dev_th, NULL, device_stage, &sp) != 0) {
		stream_pipe_free(&sp);
		return EMUB_ENOMEM;
	}
	if (pthread_create(&wr_th, NULL, writer_stage, &sp) != 0) {
		/* let the device stage run out on an empty last block, then give up */
		sp.slot[0].m = 0; sp.slot[0].last = 1;
		slot_set(&sp, &sp.slot[0], SLOT_PARSED);
		pthread_join(dev_th, NULL);
		stream_pipe_free(&sp);
		return EMUB_ENOMEM;
	}

	size_t ibuf_cap = maxvals * 24; /* about one block of "%.17g" text */
	if (ibuf_cap < ((size_t)1 << 20)) ibuf_cap = (size_t)1 << 20;
	if (ibuf_cap > ((size_t)1 << 28)) ibuf_cap = (size_t)1 << 28;
	char *ibuf = (char *)malloc(ibuf_cap + 1);
	size_t ilen = 0, nvals = 0, bin_bytes = 0;
	int eof = 0;
	const int fd = fileno(in);
#ifdef F_SETPIPE_SZ
	(void)fcntl(fd, F_SETPIPE_SZ, 1 << 20); /* fewer, larger reads when stdin is a pipe; fails harmlessly otherwise */
#endif
	for (int si = 0;; si = (si + 1) % NSLOT) {
		stream_slot *s = &sp.slot[si];
		double t0 = now_s();
		slot_wait(&sp, s, SLOT_FREE);
		sp.t_reader_wait += now_s() - t0;
		/* values of an incomplete trailing point of the previous block lead this one */
		stream_slot *prev = &sp.slot[(si + NSLOT - 1) % NSLOT];
		if (binary) { if (bin_bytes) memcpy(s->pts, (char *)prev->pts + (size_t)prev->m * d * sizeof(double), bin_bytes); }
		else if (nvals) memcpy(s->pts, prev->pts + (size_t)prev->m * d, sizeof(double) * nvals);
		/* Fill the point block.  Input is taken with read(2) as it becomes available: a block is closed as soon as it
		 * is full, the input ends, or at least one complete point is in hand and nothing more is waiting -- so a
		 * client that sends one point and waits for its answer (the reference flushes per point) is still served. */
		int must_block = 0;
		for (;;) {
			/* gather what is waiting, without blocking once something complete is in hand */
			for (;;) {
				const size_t room = binary ? maxvals * sizeof(double) - bin_bytes : ibuf_cap - ilen;
				if (eof || room == 0) break;
				const int have = binary ? bin_bytes >= (size_t)d * sizeof(double) : (ilen > 0 || nvals >= (size_t)d);
				if (have && !must_block) {
					/* a producer that is still writing refills a pipe within microseconds: give it 1 ms before taking
					 * the silence for a client that waits for its answers */
					struct pollfd pfd = {fd, POLLIN, 0};
					if (poll(&pfd, 1, 1) <= 0) break;
				}
				t0 = now_s();
				const ssize_t r = binary ? read(fd, (char *)s->pts + bin_bytes, room) : read(fd, ibuf + ilen, room);
				if (r <= 0) eof = 1;
				else if (binary) bin_bytes += (size_t)r;
				else ilen += (size_t)r;
				must_block = 0;
				sp.t_read += now_s() - t0;
			}
			size_t progress = 0;
			if (binary) nvals = bin_bytes / sizeof(double);
			else if (ilen > 0) {
				/* only text up to the last separator is complete, unless the input has ended */
				size_t usable = ilen;
				if (!eof) while (usable > 0 && !is_sep(ibuf[usable - 1])) usable--;
				if (usable > 0) {
					size_t consumed = 0;
					int bad = 0;
					t0 = now_s();
					nvals += emub_parse_doubles(ibuf, usable, s->pts + nvals, maxvals - nvals, sp.threads, &consumed, &bad);
					if (bad) { eof = 1; consumed = ilen; }
					memmove(ibuf, ibuf + consumed, ilen - consumed);
					ilen -= consumed;
					progress = consumed;
					sp.t_parse += now_s() - t0;
				}
			}
			if (nvals >= maxvals || eof) break;
			if (!binary && ilen == ibuf_cap && progress == 0) { eof = 1; break; } /* a single token longer than the buffer: give up */
			if (nvals >= (size_t)d) {
				struct pollfd pfd = {fd, POLLIN, 0};
				if (poll(&pfd, 1, 0) <= 0) break; /* complete points and nothing waiting: answer them now */
			} else
				must_block = 1;
		}
		const int m = (int)(nvals / (size_t)d); /* an incomplete trailing point is dropped at the end, like the reference (:427-428) */
		const size_t used = (size_t)m * d;
		if (binary) { bin_bytes -= used * sizeof(double); nvals = bin_bytes / sizeof(double); }
		else nvals -= used;
		pthread_mutex_lock(&sp.mu);
		const int failed = sp.rc != EMUB_OK;
		pthread_mutex_unlock(&sp.mu);
		s->m = m;
		s->last = failed || (eof && (binary || ilen == 0 || m == 0));
		const int last = s->last;
		slot_set(&sp, s, SLOT_PARSED);
		if (last) break;
	}
	pthread_join(dev_th, NULL);
	pthread_join(wr_th, NULL);
	free(ibuf);
	stream_pipe_free(&sp);
	if (getenv("EMUB_STREAM_STATS"))
		fprintf(stderr, "emub_interactive_stream: %lld points, %d io threads | reader: read %.3f s, parse %.3f s, wait %.3f s | device: predict %.3f s, "
		        "wait %.3f s | writer: format %.3f s, write %.3f s, wait %.3f s\n", sp.total, sp.threads, sp.t_read, sp.t_parse,
		        sp.t_reader_wait, sp.t_predict, sp.t_device_wait, sp.t_format, sp.t_write, sp.t_writer_wait);
	if (npoints) *npoints = sp.total;
	return sp.rc;
}
