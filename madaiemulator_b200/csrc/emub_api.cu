// emub_api.cu -- host orchestration + C-ABI (include/emu_b200.h) of the B200 GP engine.
//
// Data layout in HBM, per model (npad = n rounded up to 128, the padding is an identity block so
// that no kernel needs edge predication; nblk = npad / 128):
//   X   (npad x d)            design, padded rows zero
//   Yh  (npad x ncp)          [ y | H | 0 ]  (ncp = p + 1 rounded up to 8)
//   per slot (one in-flight likelihood evaluation):
//     bufA (npad x npad)      C (lower) -> scratch T = L21 W11 of the inverse merge -> Cinv = W^T W (lower tiles)
//     bufT (npad x npad)      L (lower)
//     bufW (npad x npad)      W = L^-1 (lower; diagonal blocks from POTF2, the rest by recursive merge)
//     UG (npad x ncp)         W [y | H]      AB (npad x ncp)  W^T UG = [alpha | C^-1 H]
//     small per-slot arrays (constants, log-det partials, Gram partials, gradient partials, results)
//
// One likelihood + gradient evaluation (reference: evalFnGradMulti, maxmultimin.c:615) is
//   K1 covariance (lower) -> recursive Cholesky + triangular inverse (register-resident POTF2 on the 128 x 128
//   diagonal blocks; DMMA TRSM-as-GEMM, SYRK and inverse-merge products with K = half the node size) ->
//   skinny products + p x p regression algebra -> DMMA W^T W -> fused gradient reduction.
// Slots are processed in lock step per stream group, groups run concurrently on their own streams.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/emu_b200.h"
#include "emub_gemm.cuh"
#include "emub_kernels.cuh"

using namespace emub;

static thread_local char g_err[512] = "";
static int set_err(int code, const char *fmt, const char *a = "", int line = 0)
{
	snprintf(g_err, sizeof(g_err), fmt, a, line);
	return code;
}
#define CUDA_TRY(expr)                                                                             \
	do {                                                                                           \
		cudaError_t e__ = (expr);                                                                  \
		if (e__ != cudaSuccess) return set_err(EMUB_ECUDA, "CUDA: %s (emub_api.cu:%d)", cudaGetErrorString(e__), __LINE__); \
	} while (0)

// Host -> device upload that the kernels of stream `st` may read afterwards.  A plain cudaMemcpy from pageable memory
// returns once the data sits in the driver's staging buffer (copies up to 64 KB), the DMA itself runs on the legacy
// default stream -- which the engine's non-blocking streams do not wait for.  A kernel launched right after such a
// copy can therefore read the destination before the data lands (seen as a model that trained on the previous
// model's training vector).  Every upload goes through the consumer's stream instead and is completed before the call
// returns, so that the caller's buffer may be released.
static cudaError_t upload(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
	cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
	return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}
static cudaError_t upload2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t st)
{
	cudaError_t e = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, st);
	return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}

// device scratch that lives for one call: released on every return path
struct ScopedDev {
	double *p = nullptr;
	~ScopedDev() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t doubles) { return cudaMalloc(&p, sizeof(double) * (doubles ? doubles : 1)); }
	ScopedDev() = default;
	ScopedDev(const ScopedDev &) = delete;
	ScopedDev &operator=(const ScopedDev &) = delete;
};

static const char *k_family_names[EMUB_K_NFAMILIES] = {"cov", "potf2", "gemm_chol", "gemm_trtri", "gemm_lauum", "skinny",
                                                       "small", "grad", "kcross", "gemm_pred", "pred_final"};

struct ProfAcc { double ms; long long launches; double work; };
constexpr int AUX_DEPTH = 12;

struct emub_ctx {
	int device;
	int ngroups;
	int use_graphs;  // replay captured CUDA graphs for the per-chunk launch sequences (default on; EMUB_NO_GRAPHS=1 disables)
	cudaStream_t streams[4];
	// side streams of the factorisation, per group and recursion depth: the first product of a node's inverse merge
	// (T = L21 W11) is independent of everything its right child does, so with few matrices in flight it runs beside it
	cudaStream_t aux[4][AUX_DEPTH];
	cudaEvent_t aux_fork[4][AUX_DEPTH], aux_join[4][AUX_DEPTH];
	int aux_max_count;  // use the side streams when a group holds at most this many matrices (EMUB_AUX_MAX, default 8; 0 = never)
	int profile;
	ProfAcc prof[EMUB_K_NFAMILIES];
	long long launches;
	cudaEvent_t ev0, ev1;
	cudaEvent_t gev[4];
	long long small_launch_ctas;  // products with at most this many 64 x 64 sub-tiles run with 32 x 32 ones (EMUB_SMALL_CTAS)
};

// launch bookkeeping: in profile mode every launch is timed on its own (events + sync)
struct LaunchScope {
	emub_ctx *c; int fam; double work; cudaStream_t st;
	LaunchScope(emub_ctx *c_, int fam_, double work_, cudaStream_t st_) : c(c_), fam(fam_), work(work_), st(st_)
	{
		c->launches++;
		if (c->profile) cudaEventRecord(c->ev0, st);
	}
	~LaunchScope()
	{
		if (c->profile) {
			cudaEventRecord(c->ev1, st);
			cudaEventSynchronize(c->ev1);
			float ms = 0.f;
			cudaEventElapsedTime(&ms, c->ev0, c->ev1);
			c->prof[fam].ms += ms;
			c->prof[fam].launches++;
			c->prof[fam].work += work;
		}
	}
};

// one kernel launch of the factorisation sequence (depth-first order of the recursion)
enum { STEP_POTF2 = 0, STEP_TRSM, STEP_SYRK, STEP_TMUL, STEP_WMUL };
struct FactorStep { int type; int off, cnt; int kblk; double flops; int full_only; int depth; };
// a node [lo, hi) of the recursion on its right spine (root, right child of the root, ...), split at mid
struct SpineNode { int lo, mid, hi; };

struct emub_model {
	emub_ctx *ctx;
	int n, d, p, ncp, kernel, order, nth, npad, nblk, nslots;
	size_t mat;  // npad * npad
	double *dX, *dy, *dYh;  // dYh: ncomp x npad x ncp
	int ncomp;              // training vectors sharing the design (PCA components, multi_modelstruct.c:121-148)
	int exact_grad;         // 0: the reference's literal gradient formula; 1: the true gradient of the objective (D-4)
	int *dComp, *hComp;     // component evaluated by each slot of the current chunk
	struct QueryWs *qws;    // query workspace shared by every emulator of this model (lazily allocated)
	GemmTask *dTasks;
	std::vector<FactorStep> steps;
	std::vector<SpineNode> spine;
	int lauum_off, lauum_cnt;
	double lauum_flops;
	double *bufA, *bufW, *bufT;
	double *dUG, *dAB, *dConsts, *dLogdet, *dGramPart, *dRes, *dGradPart, *dMinv, *dThetas;
	int *dInfo;
	double *hThetas, *hRes;  // pinned
	int last_count;
	std::vector<std::pair<unsigned long long, cudaGraphExec_t>> graphs;   // captured chunk sequences, by configuration
	std::vector<std::pair<unsigned long long, long long>> graph_launches;  // kernels per replay (for emub_launch_count)
};

// query workspace: one chunk of up to mqc points
struct QueryWs {
	int mqc;      // query chunk (multiple of 128)
	int ntcap;    // observables the output / projection buffers hold (grown on demand, ensure_output_capacity)
	int ncomp;    // rows of dMean / dVar
	double *dQ, *dK, *dVsq, *dKA, *dMean, *dVar;  // dMean, dVar: ncomp x mqc
	double *dFew;                                 // partial products of the few-points path: [ceil(npad/FEW_JC)][npad][8]
	unsigned long long proj_hash;                 // contents of dProj (0: none), so a point-by-point caller uploads it once
	// several emulators answered by one set of launches on the few-points path (predict_few_set)
	FewSet *dFewSet;
	double *dFewConsts;                           // [few_cap][CONST_STRIDE], contiguous copy of the emulators' constants
	int few_cap;                                  // emulators dFewSet / dFewConsts / dFew hold
	unsigned long long few_hash;                  // which emulators the tables describe (0: none)
	// the few-points call sequences (copy in, 4-5 kernels, copies out) replayed as CUDA graphs; every pointer in them
	// belongs to this workspace or to the emulators named by the key
	struct { unsigned long long key; cudaGraphExec_t exec; long long launches; } few_graphs[8];
	int few_graph_next;
	double *dOutM, *dOutV, *dProj;                // back-projected outputs (mqc x ntmax), projection data
	GemmTask *dTasks;
	double *hQ, *hOut;  // pinned
	// second set for the double-buffered chunk walk (walk_query_chunks): [0] aliases dQ / hQ / hOut
	double *dQb[2], *hQb[2], *hOutb[2];
	cudaStream_t copy_stream;
	cudaEvent_t evIn[2], evOut[2];
};
constexpr int NTMAX = 64;     // observables (or components) the back-projection buffers start out with
constexpr int NTLIMIT = 1024; // ... and the most they grow to

struct emub_emulator {
	emub_model *m;
	double *W, *AB, *beta, *Minv, *consts;
	double kappa;
	double hbeta[MAXNCP];
	unsigned long long uid;  // never reused: device addresses are, so caches of emulator sets key on this
};
static std::atomic<unsigned long long> g_emulator_uid{1};

extern "C" const char *emub_last_error(void) { return g_err; }
extern "C" const char *emub_version(void) { return "emub 0.1 (sm_100a, FP64 DMMA)"; }
extern "C" const char *emub_profile_name(int f) { return (f >= 0 && f < EMUB_K_NFAMILIES) ? k_family_names[f] : "?"; }

extern "C" int emub_ctx_create(int device, emub_ctx **out)
{
	if (!out) return set_err(EMUB_EINVAL, "emub_ctx_create: null out%s");
	int ndev = 0;
	CUDA_TRY(cudaGetDeviceCount(&ndev));
	if (device < 0 || device >= ndev) return set_err(EMUB_EINVAL, "emub_ctx_create: no such device%s");
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) return set_err(EMUB_ECUDA, "emub: device is not sm_100 class (%s)", prop.name);
	emub_ctx *c = new emub_ctx();
	memset(c, 0, sizeof(*c));
	c->device = device;
	c->ngroups = 4; // measured best from B = 4 upwards (profiles/r02_groups_sweep.txt); a call uses min(ngroups, points)
	c->use_graphs = getenv("EMUB_NO_GRAPHS") ? 0 : 1;
	// half of what the default configuration keeps resident (148 SMs x 4 CTAs)
	c->small_launch_ctas = getenv("EMUB_SMALL_CTAS") ? atoll(getenv("EMUB_SMALL_CTAS")) : 296;
	c->aux_max_count = getenv("EMUB_AUX_MAX") ? atoi(getenv("EMUB_AUX_MAX")) : 8;
	int prio_lo = 0, prio_hi = 0;
	CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	for (int g = 0; g < 4; g++) {
		// the main stream of a group carries the critical path (leaf factorisations, TRSM, SYRK): highest priority
		CUDA_TRY(cudaStreamCreateWithPriority(&c->streams[g], cudaStreamNonBlocking, prio_hi));
		CUDA_TRY(cudaEventCreateWithFlags(&c->gev[g], cudaEventDisableTiming));
		for (int k = 0; k < AUX_DEPTH; k++) {
			CUDA_TRY(cudaStreamCreateWithPriority(&c->aux[g][k], cudaStreamNonBlocking, prio_lo));
			CUDA_TRY(cudaEventCreateWithFlags(&c->aux_fork[g][k], cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&c->aux_join[g][k], cudaEventDisableTiming));
		}
	}
	CUDA_TRY(cudaEventCreate(&c->ev0));
	CUDA_TRY(cudaEventCreate(&c->ev1));
	// opt in to large dynamic shared memory once
	CUDA_TRY(cudaFuncSetAttribute(k_gemm<KMAJOR, KMAJOR, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_gemm<KMAJOR, KMAJOR, EPI_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_gemm<KMAJOR, RMAJOR, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_gemm<RMAJOR, RMAJOR, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_gemm<KMAJOR, RMAJOR, EPI_COLSUMSQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
	CUDA_TRY(cudaFuncSetAttribute(k_potf2, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM_BYTES));
	// kernels whose dynamic shared memory grows with nparams / nregression_fns past the 48 KB default
	{
		const int stage = (2 * MAXD * CT + 2 * CT) * (int)sizeof(double);
		CUDA_TRY(cudaFuncSetAttribute(k_cov<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_cov<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_cov<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_cov<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_cov<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_cov<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_grad_tiles<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, stage));
		CUDA_TRY(cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * MAXNCP * (MAXNCP + 1) * (int)sizeof(double)));
		CUDA_TRY(cudaFuncSetAttribute(k_pred_final, cudaFuncAttributeMaxDynamicSharedMemorySize, MAXNCP * MAXNCP * (int)sizeof(double)));
		CUDA_TRY(cudaFuncSetAttribute(k_gram_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, TB * MAXNCP * (int)sizeof(double)));
	}
	*out = c;
	return EMUB_OK;
}

extern "C" void emub_ctx_destroy(emub_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaDeviceSynchronize();
	for (int g = 0; g < 4; g++) {
		cudaStreamDestroy(c->streams[g]);
		cudaEventDestroy(c->gev[g]);
		for (int k = 0; k < AUX_DEPTH; k++) {
			cudaStreamDestroy(c->aux[g][k]);
			cudaEventDestroy(c->aux_fork[g][k]);
			cudaEventDestroy(c->aux_join[g][k]);
		}
	}
	cudaEventDestroy(c->ev0);
	cudaEventDestroy(c->ev1);
	delete c;
}
extern "C" void *emub_ctx_stream(emub_ctx *c) { return c ? (void *)c->streams[0] : nullptr; }
extern "C" int emub_ctx_set_groups(emub_ctx *c, int ng)
{
	if (!c || ng < 1 || ng > 4) return set_err(EMUB_EINVAL, "emub_ctx_set_groups: 1..4%s");
	c->ngroups = ng;
	return EMUB_OK;
}
extern "C" int emub_ctx_use_graphs(emub_ctx *c, int on)
{
	if (!c) return EMUB_EINVAL;
	c->use_graphs = on ? 1 : 0;
	return EMUB_OK;
}
extern "C" int emub_ctx_synchronize(emub_ctx *c)
{
	if (!c) return EMUB_EINVAL;
	CUDA_TRY(cudaSetDevice(c->device));
	for (int g = 0; g < 4; g++) CUDA_TRY(cudaStreamSynchronize(c->streams[g]));  // the side streams join their main stream
	return EMUB_OK;
}
extern "C" int emub_profile_enable(emub_ctx *c, int on) { if (!c) return EMUB_EINVAL; c->profile = on ? 1 : 0; return EMUB_OK; }
extern "C" int emub_profile_reset(emub_ctx *c) { if (!c) return EMUB_EINVAL; memset(c->prof, 0, sizeof(c->prof)); return EMUB_OK; }
extern "C" int emub_profile_read(emub_ctx *c, int f, double *ms, long long *launches, double *work)
{
	if (!c || f < 0 || f >= EMUB_K_NFAMILIES) return EMUB_EINVAL;
	if (ms) *ms = c->prof[f].ms;
	if (launches) *launches = c->prof[f].launches;
	if (work) *work = c->prof[f].work;
	return EMUB_OK;
}
extern "C" long long emub_launch_count(emub_ctx *c) { return c ? c->launches : 0; }

// ---- schedules --------------------------------------------------------------------------------------
// Recursive factorisation of the block range [lo, hi):  A = L L^T and W = L^-1 together.
//   rec(lo, mid);                                   L11 (bufT), W11 (bufW)
//   L21 = A21 W11^T                                 TRSM as a GEMM against the inverse; bufA -> bufT
//   A22 -= L21 L21^T                                SYRK, in place in bufA
//   rec(mid, hi);                                   L22, W22
//   T = L21 W11 ;  W21 = -W22 T                     triangular-inverse merge; T lives in the dead A21 region
// Every GEMM therefore runs with K = the size of the half it multiplies against (94% of the flops at
// K >= 1024 for n = 4096), instead of the K = 128 rank updates of a panel-by-panel sweep.
//
// The Cholesky factor itself only needs the inverses of LEFT halves (they feed the TRSM of their parent): the merge of
// a node on the right spine of the recursion (the root, its right child, ...) produces blocks of W that nothing but the
// full inverse uses.  Those steps are marked full_only: a value-only likelihood (evalFnMulti, maxmultimin.c:288-394)
// skips them and costs n^3/3 + n^3/21 flops instead of 2n^3/3 (+ n^3/3 for W^T W); L and the W blocks it does build are
// the same bits either way.
static void build_factor(emub_model *m, int lo, int hi, std::vector<GemmTask> &tasks, bool on_spine, int depth = 0)
{
	const long long ld = m->npad;
	auto off = [&](int bi, int bj) { return (long long)bi * TB * ld + (long long)bj * TB; };
	auto by_k = [](const GemmTask &a, const GemmTask &b) { return a.klen > b.klen; };
	auto emit = [&](int type, std::vector<GemmTask> &t, int full_only) {
		std::stable_sort(t.begin(), t.end(), by_k);
		FactorStep st{type, (int)tasks.size(), (int)t.size(), 0, 0.0, full_only, depth};
		for (auto &x : t) { tasks.push_back(x); st.flops += task_flops(x); }
		m->steps.push_back(st);
	};
	if (hi - lo == 1) {
		m->steps.push_back({STEP_POTF2, 0, 0, lo, 2.0 * TB * TB * TB / 3.0, 0, depth});
		if (on_spine) m->spine.push_back({lo, hi, hi});
		return;
	}
	const int mid = lo + (hi - lo + 1) / 2;
	if (on_spine) m->spine.push_back({lo, mid, hi});
	build_factor(m, lo, mid, tasks, false, depth + 1);
	std::vector<GemmTask> t;
	// L(i,j) = sum_{k in [lo, j]} A(i,k) W(j,k)^T          A = bufA KMAJOR, B = bufW KMAJOR -> bufT
	for (int i = mid; i < hi; i++)
		for (int j = lo; j < mid; j++) t.push_back({off(i, lo), off(j, lo), off(i, j), (j - lo + 1) * TB, TASK_TRIM_END_SC0});
	emit(STEP_TRSM, t, 0);
	t.clear();
	// T(i,j) = sum_{k in [j, mid)} L(i,k) W(k,j)           A = bufT KMAJOR, B = bufW RMAJOR -> bufA
	// Issued right after the TRSM: it needs L21 and W11 only and writes the A21 region of bufA, which nothing reads any
	// more, so it is independent of the SYRK and of the whole right child -- run_factor puts it on a side stream.
	for (int i = mid; i < hi; i++)
		for (int j = lo; j < mid; j++) t.push_back({off(i, j), off(j, j), off(i, j), (mid - j) * TB, TASK_TRIM_BEGIN_SC1});
	emit(STEP_TMUL, t, on_spine ? 1 : 0);
	t.clear();
	// A(i,j) -= sum_{k in [lo, mid)} L(i,k) L(j,k)^T      A, B = bufT KMAJOR -> bufA
	for (int i = mid; i < hi; i++)
		for (int j = mid; j <= i; j++) t.push_back({off(i, lo), off(j, lo), off(i, j), (mid - lo) * TB, (i == j) ? TASK_LOWER : 0});
	emit(STEP_SYRK, t, 0);
	build_factor(m, mid, hi, tasks, on_spine, depth + 1);
	t.clear();
	// W(i,j) = - sum_{k in [mid, i]} W(i,k) T(k,j)         A = bufW KMAJOR, B = bufA RMAJOR -> bufW
	for (int i = mid; i < hi; i++)
		for (int j = lo; j < mid; j++) t.push_back({off(i, mid), off(mid, j), off(i, j), (i + 1 - mid) * TB, TASK_TRIM_END_SR0});
	emit(STEP_WMUL, t, on_spine ? 1 : 0);
}

static void build_schedules(emub_model *m, std::vector<GemmTask> &tasks)
{
	const long long ld = m->npad;
	const int nb = m->nblk;
	auto off = [&](int bi, int bj) { return (long long)bi * TB * ld + (long long)bj * TB; };
	build_factor(m, 0, nb, tasks, true);
	// Cinv(i,j) = sum_{k >= i} W(k,i)^T W(k,j), i >= j     A = W RMAJOR, B = W RMAJOR
	m->lauum_off = (int)tasks.size();
	m->lauum_flops = 0;
	for (int i = 0; i < nb; i++)
		for (int j = 0; j <= i; j++) {
			tasks.push_back({off(i, i), off(i, j), off(i, j), (nb - i) * TB, TASK_TRIM_BEGIN_SR1 | ((i == j) ? TASK_LOWER : 0)});
			m->lauum_flops += task_flops(tasks.back());
		}
	m->lauum_cnt = (int)tasks.size() - m->lauum_off;
}

static void free_query_ws(emub_model *m);
extern "C" void emub_model_destroy(emub_model *m);
static int regression_fns(int order, int d) { if (order < 0 || order > 3) order = 0; return 1 + order * d; }

extern "C" int emub_model_create(emub_ctx *ctx, const double *X, int ldx, int n, int d, const double *y, int kernel,
                                 int order, int max_slots, emub_model **out)
{
	if (!ctx || !X || !y || !out) return set_err(EMUB_EINVAL, "emub_model_create: null argument%s");
	if (n < 1 || d < 1 || d > MAXD || ldx < d) return set_err(EMUB_EINVAL, "emub_model_create: need 1 <= nparams <= 64, n >= 1%s");
	if (kernel < 1 || kernel > 3) kernel = EMUB_POWEREXP;  // multi_modelstruct.c:67-74 falls back to power-exp
	if (order < 0 || order > 3) order = 0;
	const int p = regression_fns(order, d);
	if (p + 1 > MAXNCP) return set_err(EMUB_EINVAL, "emub_model_create: 1 + regression_order * nparams + 1 must be <= 104%s");
	CUDA_TRY(cudaSetDevice(ctx->device));
	emub_model *m = new emub_model();  // value-initialised: every pointer starts out null, so a failed set-up can be torn down
	m->ctx = ctx; m->n = n; m->d = d; m->p = p; m->kernel = kernel; m->order = order;
	m->ncp = ((p + 1) + 7) / 8 * 8;
	m->nth = (kernel == EMUB_POWEREXP) ? d + 2 : 3;
	m->npad = (n + TB - 1) / TB * TB;
	m->nblk = m->npad / TB;
	m->mat = (size_t)m->npad * m->npad;
	m->last_count = 0;
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { delete m; return set_err(EMUB_ECUDA, "emub_model_create: cudaMemGetInfo failed%s"); }
	const size_t per_slot = 3 * m->mat * sizeof(double) + (size_t)m->npad * m->ncp * 16 + (1 << 20);
	int slots = (int)std::min<size_t>(64, (size_t)(0.6 * (double)free_b) / per_slot);
	if (max_slots > 0) slots = std::min(slots, max_slots);
	if (slots < 1) { delete m; return set_err(EMUB_ENOMEM, "emub_model_create: not enough device memory for one slot%s"); }
#define MODEL_TRY(expr)                                                                                                  \
	do {                                                                                                                 \
		cudaError_t e__ = (expr);                                                                                        \
		if (e__ != cudaSuccess) {                                                                                        \
			emub_model_destroy(m);                                                                                       \
			return set_err(e__ == cudaErrorMemoryAllocation ? EMUB_ENOMEM : EMUB_ECUDA, "CUDA: %s (emub_model_create:%d)", \
			               cudaGetErrorString(e__), __LINE__);                                                           \
		}                                                                                                                \
	} while (0)
	m->nslots = slots;

	std::vector<double> Xp((size_t)m->npad * d, 0.0);
	for (int i = 0; i < n; i++) memcpy(&Xp[(size_t)i * d], X + (size_t)i * ldx, sizeof(double) * d);
	MODEL_TRY(cudaMalloc(&m->dX, Xp.size() * sizeof(double)));
	MODEL_TRY(upload(m->dX, Xp.data(), Xp.size() * sizeof(double), ctx->streams[0]));
	MODEL_TRY(cudaMalloc(&m->dy, sizeof(double) * n));
	MODEL_TRY(upload(m->dy, y, sizeof(double) * n, ctx->streams[0]));
	MODEL_TRY(cudaMalloc(&m->dYh, sizeof(double) * (size_t)m->npad * m->ncp));
	m->ncomp = 1;
	m->qws = nullptr;
	std::vector<GemmTask> tasks;
	build_schedules(m, tasks);
	if (tasks.empty()) tasks.push_back({0, 0, 0, 0, 0});
	MODEL_TRY(cudaMalloc(&m->dTasks, tasks.size() * sizeof(GemmTask)));
	MODEL_TRY(upload(m->dTasks, tasks.data(), tasks.size() * sizeof(GemmTask), ctx->streams[0]));
	const size_t S = slots;
	MODEL_TRY(cudaMalloc(&m->bufA, S * m->mat * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->bufW, S * m->mat * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->bufT, S * m->mat * sizeof(double)));
	// stream-ordered like every other initialisation (a cudaMemset on the legacy stream is not ordered against the
	// engine's non-blocking streams and may still run when the first evaluation starts)
	MODEL_TRY(cudaMemsetAsync(m->bufW, 0, S * m->mat * sizeof(double), ctx->streams[0]));
	MODEL_TRY(cudaMemsetAsync(m->bufT, 0, S * m->mat * sizeof(double), ctx->streams[0]));
	MODEL_TRY(cudaMalloc(&m->dUG, S * m->npad * m->ncp * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dAB, S * m->npad * m->ncp * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dConsts, S * CONST_STRIDE * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dLogdet, S * m->nblk * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dGramPart, S * m->nblk * (size_t)m->ncp * m->ncp * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dRes, S * RES_STRIDE * sizeof(double)));
	const size_t nt64 = m->npad / CT, ntl = nt64 * (nt64 + 1) / 2;
	MODEL_TRY(cudaMalloc(&m->dGradPart, S * ntl * (size_t)m->d * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dMinv, S * MAXNCP * MAXNCP * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dThetas, S * (MAXD + 2) * sizeof(double)));
	MODEL_TRY(cudaMalloc(&m->dInfo, S * sizeof(int)));
	MODEL_TRY(cudaMalloc(&m->dComp, S * sizeof(int)));
	MODEL_TRY(cudaMemsetAsync(m->dComp, 0, S * sizeof(int), ctx->streams[0]));
	MODEL_TRY(cudaMallocHost(&m->hComp, S * sizeof(int)));
	MODEL_TRY(cudaMallocHost(&m->hThetas, S * (MAXD + 2) * sizeof(double)));
	MODEL_TRY(cudaMallocHost(&m->hRes, S * RES_STRIDE * sizeof(double)));
	cudaStream_t st = ctx->streams[0];
	{
		LaunchScope ls(ctx, EMUB_K_SMALL, 0, st);
		k_build_yh<<<(m->npad + 127) / 128, 128, 0, st>>>(m->dX, m->dy, 1, n, m->npad, d, order, m->ncp, m->dYh);
	}
	MODEL_TRY(cudaStreamSynchronize(st));
	MODEL_TRY(cudaGetLastError());
#undef MODEL_TRY
	*out = m;
	return EMUB_OK;
}

extern "C" void emub_model_destroy(emub_model *m)
{
	if (!m) return;
	cudaSetDevice(m->ctx->device);
	cudaDeviceSynchronize();
	cudaFree(m->dX); cudaFree(m->dy); cudaFree(m->dYh); cudaFree(m->dTasks);
	cudaFree(m->bufA); cudaFree(m->bufW); cudaFree(m->bufT);
	cudaFree(m->dUG); cudaFree(m->dAB); cudaFree(m->dConsts); cudaFree(m->dLogdet); cudaFree(m->dGramPart);
	cudaFree(m->dRes); cudaFree(m->dGradPart); cudaFree(m->dMinv); cudaFree(m->dThetas); cudaFree(m->dInfo);
	cudaFree(m->dComp);
	cudaFreeHost(m->hThetas); cudaFreeHost(m->hRes); cudaFreeHost(m->hComp);
	for (auto &g : m->graphs) cudaGraphExecDestroy(g.second);
	free_query_ws(m);
	delete m;
}
extern "C" int emub_model_nthetas(const emub_model *m) { return m ? m->nth : 0; }
extern "C" int emub_model_nregression_fns(const emub_model *m) { return m ? m->p : 0; }
extern "C" int emub_model_slots(const emub_model *m) { return m ? m->nslots : 0; }
extern "C" int emub_model_kernel(const emub_model *m) { return m ? m->kernel : 0; }

static void few_graphs_clear(QueryWs *w)
{
	for (auto &g : w->few_graphs) {
		if (g.exec) cudaGraphExecDestroy(g.exec);
		g.exec = nullptr;
		g.key = 0;
	}
}

static void free_query_ws(emub_model *m)
{
	QueryWs *w = m->qws;
	if (!w) return;
	few_graphs_clear(w);
	cudaFree(w->dQ); cudaFree(w->dK); cudaFree(w->dVsq); cudaFree(w->dKA); cudaFree(w->dMean); cudaFree(w->dVar);
	cudaFree(w->dOutM); cudaFree(w->dOutV); cudaFree(w->dProj); cudaFree(w->dTasks); cudaFree(w->dFew);
	cudaFree(w->dFewSet); cudaFree(w->dFewConsts);
	cudaFreeHost(w->hQ); cudaFreeHost(w->hOut);
	cudaFree(w->dQb[1]); cudaFreeHost(w->hQb[1]); cudaFreeHost(w->hOutb[1]);
	if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
	for (int i = 0; i < 2; i++) { if (w->evIn[i]) cudaEventDestroy(w->evIn[i]); if (w->evOut[i]) cudaEventDestroy(w->evOut[i]); }
	delete w;
	m->qws = nullptr;
}

extern "C" int emub_model_set_training_multi(emub_model *m, const double *Y, int ldy, int ncomp)
{
	if (!m || !Y || ncomp < 1 || ldy < ncomp) return set_err(EMUB_EINVAL, "emub_model_set_training_multi: bad argument%s");
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	CUDA_TRY(cudaDeviceSynchronize());
	cudaStream_t st = m->ctx->streams[0];
	ScopedDev sY;
	CUDA_TRY(sY.alloc((size_t)m->n * ncomp));
	double *dY = sY.p;
	CUDA_TRY(upload2d(dY, sizeof(double) * ncomp, Y, sizeof(double) * ldy, sizeof(double) * ncomp, m->n, st));
	// captured graphs hold the address of dYh: drop them whenever the training data is rebuilt
	for (auto &g : m->graphs) cudaGraphExecDestroy(g.second);
	m->graphs.clear();
	m->graph_launches.clear();
	if (ncomp != m->ncomp) {
		cudaFree(m->dYh);
		m->dYh = nullptr;
		m->ncomp = 0;
		free_query_ws(m);
		CUDA_TRY(cudaMalloc(&m->dYh, sizeof(double) * (size_t)ncomp * m->npad * m->ncp));
		m->ncomp = ncomp;
	}
	for (int c = 0; c < ncomp; c++) {
		LaunchScope ls(m->ctx, EMUB_K_SMALL, 0, st);
		k_build_yh<<<(m->npad + 127) / 128, 128, 0, st>>>(m->dX, dY + c, ncomp, m->n, m->npad, m->d, m->order, m->ncp,
		                                                   m->dYh + (size_t)c * m->npad * m->ncp);
	}
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

extern "C" int emub_model_ncomponents(const emub_model *m) { return m ? m->ncomp : 0; }

extern "C" int emub_model_set_gradient_mode(emub_model *m, int mode)
{
	if (!m || (mode != EMUB_GRAD_LITERAL && mode != EMUB_GRAD_EXACT)) return set_err(EMUB_EINVAL, "emub_model_set_gradient_mode: bad argument%s");
	m->exact_grad = mode == EMUB_GRAD_EXACT;
	return EMUB_OK;
}
extern "C" int emub_model_gradient_mode(const emub_model *m) { return (m && m->exact_grad) ? EMUB_GRAD_EXACT : EMUB_GRAD_LITERAL; }

extern "C" int emub_model_set_training(emub_model *m, const double *y)
{
	if (!m || !y) return set_err(EMUB_EINVAL, "emub_model_set_training: null%s");
	return emub_model_set_training_multi(m, y, 1, 1);
}

// ---- kernel launch helpers ----------------------------------------------------------------------------
template <int AL, int BL, int EPI>
static void launch_gemm(emub_ctx *c, int fam, double flops, cudaStream_t st, const GemmTask *tasks, int ntasks, int batch,
                        const double *A, long long sA, int lda, const double *B, long long sB, int ldb, double *C,
                        long long sC, int ldc, double alpha)
{
	if (ntasks <= 0 || batch <= 0) return;
	GemmArgs a{tasks, A, B, C, sA, sB, sC, lda, ldb, ldc, alpha};
	LaunchScope ls(c, fam, flops * batch, st);
	// the column-sum epilogue lays its output out by sub-tile row, so it keeps the default configuration
	if (EPI != EPI_COLSUMSQ && (long long)ntasks * DefaultCfg::SUBS * batch <= c->small_launch_ctas)
		k_gemm<AL, BL, EPI, SmallCfg><<<dim3(ntasks * SmallCfg::SUBS, batch), SmallCfg::THREADS, SmallCfg::SMEM_BYTES, st>>>(a);
	else
		k_gemm<AL, BL, EPI><<<dim3(ntasks * DefaultCfg::SUBS, batch), GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(a);
}

static void launch_cov(emub_model *m, cudaStream_t st, int count, const double *consts, double *out, long long ostride,
                       int lower_only)
{
	const int nt = m->npad / CT;
	const size_t smem = 2 * (size_t)m->d * CT * sizeof(double);
	const double bytes = (lower_only ? 0.5 : 1.0) * 8.0 * (double)m->npad * m->npad + 8.0 * m->npad * m->d;
	LaunchScope ls(m->ctx, EMUB_K_COV, bytes * count, st);
	dim3 grid(nt, nt, count);
	switch (m->kernel) {
	case 2: k_cov<2, false><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, nullptr, 0, consts, CONST_STRIDE, out, ostride, m->npad, lower_only); break;
	case 3: k_cov<3, false><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, nullptr, 0, consts, CONST_STRIDE, out, ostride, m->npad, lower_only); break;
	default: k_cov<1, false><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, nullptr, 0, consts, CONST_STRIDE, out, ostride, m->npad, lower_only); break;
	}
}

static void launch_kcross(emub_model *m, cudaStream_t st, const double *consts, const double *dQ, int mq, int mq_pad,
                          double *dK, int ldk)
{
	const size_t smem = 2 * (size_t)m->d * CT * sizeof(double);
	LaunchScope ls(m->ctx, EMUB_K_KCROSS, 8.0 * (double)m->npad * mq_pad, st);
	dim3 grid(mq_pad / CT, m->npad / CT, 1);
	switch (m->kernel) {
	case 2: k_cov<2, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, consts, 0, dK, 0, ldk, 0); break;
	case 3: k_cov<3, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, consts, 0, dK, 0, ldk, 0); break;
	default: k_cov<1, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, consts, 0, dK, 0, ldk, 0); break;
	}
}

// L (bufT) and W = L^-1 (bufW) from C (bufA) for `count` slots starting at s0
// full = false: the factor only (the merges of the right spine of the recursion are skipped, see build_factor)
// The first nfull of the slots also get the merges of the right spine (the whole W = L^-1); the others the factor only.
static void run_factor(emub_model *m, cudaStream_t st, int s0, int nslots, int nfull)
{
	emub_ctx *c = m->ctx;
	const int ld = m->npad;
	const long long ms = (long long)m->mat;
	double *A = m->bufA + (size_t)s0 * m->mat, *W = m->bufW + (size_t)s0 * m->mat, *T = m->bufT + (size_t)s0 * m->mat;
	int g = 0;
	for (int k = 0; k < 4; k++) if (c->streams[k] == st) g = k;
	const bool side = !c->profile && nslots <= c->aux_max_count;
	for (const FactorStep &s : m->steps) {
		const int count = s.full_only ? nfull : nslots;
		if (count <= 0) continue;
		const GemmTask *tk = m->dTasks + s.off;
		const bool aux = side && s.depth < AUX_DEPTH;
		switch (s.type) {
		case STEP_POTF2: {
			LaunchScope ls(c, EMUB_K_POTF2, count * s.flops, st);
			k_potf2<<<count, POTF2_THREADS, POTF2_SMEM_BYTES, st>>>(A, ms, T, ms, W, ms, ld, s.kblk, m->nblk,
			                                                        m->dLogdet + (size_t)s0 * m->nblk, m->dInfo + s0);
			break;
		}
		case STEP_TRSM:
			launch_gemm<KMAJOR, KMAJOR, EPI_STORE>(c, EMUB_K_GEMM_CHOL, s.flops, st, tk, s.cnt, count, A, ms, ld, W, ms, ld, T, ms, ld, 1.0);
			break;
		case STEP_SYRK:
			launch_gemm<KMAJOR, KMAJOR, EPI_SUB>(c, EMUB_K_GEMM_CHOL, s.flops, st, tk, s.cnt, count, T, ms, ld, T, ms, ld, A, ms, ld, 1.0);
			break;
		case STEP_TMUL:
			if (aux) {
				cudaEventRecord(c->aux_fork[g][s.depth], st);  // after the node's TRSM
				cudaStreamWaitEvent(c->aux[g][s.depth], c->aux_fork[g][s.depth], 0);
				launch_gemm<KMAJOR, RMAJOR, EPI_STORE>(c, EMUB_K_GEMM_TRTRI, s.flops, c->aux[g][s.depth], tk, s.cnt, count, T, ms, ld, W, ms, ld, A, ms, ld, 1.0);
				cudaEventRecord(c->aux_join[g][s.depth], c->aux[g][s.depth]);
			} else
				launch_gemm<KMAJOR, RMAJOR, EPI_STORE>(c, EMUB_K_GEMM_TRTRI, s.flops, st, tk, s.cnt, count, T, ms, ld, W, ms, ld, A, ms, ld, 1.0);
			break;
		case STEP_WMUL:
			if (aux) cudaStreamWaitEvent(st, c->aux_join[g][s.depth], 0);  // T of this node
			launch_gemm<KMAJOR, RMAJOR, EPI_STORE>(c, EMUB_K_GEMM_TRTRI, s.flops, st, tk, s.cnt, count, W, ms, ld, A, ms, ld, W, ms, ld, -1.0);
			break;
		}
	}
}

static void run_lauum(emub_model *m, cudaStream_t st, int s0, int count)
{
	const int ld = m->npad;
	double *A = m->bufA + (size_t)s0 * m->mat, *W = m->bufW + (size_t)s0 * m->mat;
	launch_gemm<RMAJOR, RMAJOR, EPI_STORE>(m->ctx, EMUB_K_GEMM_LAUUM, m->lauum_flops, st, m->dTasks + m->lauum_off, m->lauum_cnt,
	                                       count, W, m->mat, ld, W, m->mat, ld, A, m->mat, ld, 1.0);
}

// UG = W [y | H]; Gram partials; beta / sigma2 / -L
static void run_regression(emub_model *m, cudaStream_t st, int s0, int count, int emulator_mode)
{
	emub_ctx *c = m->ctx;
	const long long sUG = (long long)m->npad * m->ncp;
	const int nchunk_cols = (m->p + 1 + 7) / 8;
	double *W = m->bufW + (size_t)s0 * m->mat, *L = m->bufT + (size_t)s0 * m->mat;
	double *UG = m->dUG + (size_t)s0 * sUG, *AB = m->dAB + (size_t)s0 * sUG;
	// UG = L^-1 [y | H] by block forward substitution down the right spine of the recursion: for a spine node
	// [lo, hi) split at mid,  u[lo:mid] = W11 b[lo:mid]  (W11 = inverse of the left half, built by every factorisation),
	// b[mid:hi] -= L21 u[lo:mid].  b starts as Yh and lives in AB (free until the gradient stage).  The same launches
	// serve the value-only and the gradient path, so both return the same bits.
	const int *comp = m->dComp + s0;
	const bool two_cols = m->p + 1 <= 2;  // regression order 0: y and the constant column
	for (size_t lv = 0; lv < m->spine.size(); lv++) {
		const SpineNode nd = m->spine[lv];
		const int r_lo = nd.lo * TB, r_mid = nd.mid * TB, r_hi = nd.hi * TB;
		const double *bsrc = lv == 0 ? m->dYh : AB;
		const long long bstride = sUG;
		const int *bcomp = lv == 0 ? comp : nullptr;
		{
			LaunchScope ls(c, EMUB_K_SKINNY, count * 4.0 * (double)(r_mid - r_lo) * (r_mid - r_lo) * nchunk_cols, st);
			if (two_cols)
				k_rows_times_range<true, 2><<<dim3((r_mid - r_lo) / 4, 1, count), 256, 0, st>>>(
				    W, (long long)m->mat, m->npad, r_lo, r_lo, r_mid, bsrc, bstride, bcomp, nullptr, 0, nullptr, m->ncp, UG, sUG);
			else
				k_rows_times_range<true, 8><<<dim3((r_mid - r_lo) / 4, nchunk_cols, count), 256, 0, st>>>(
				    W, (long long)m->mat, m->npad, r_lo, r_lo, r_mid, bsrc, bstride, bcomp, nullptr, 0, nullptr, m->ncp, UG, sUG);
		}
		if (r_hi > r_mid) {
			LaunchScope ls(c, EMUB_K_SKINNY, count * 8.0 * (double)(r_hi - r_mid) * (r_mid - r_lo) * nchunk_cols, st);
			if (two_cols)
				k_rows_times_range<false, 2><<<dim3((r_hi - r_mid) / 4, 1, count), 256, 0, st>>>(
				    L, (long long)m->mat, m->npad, r_mid, r_lo, r_mid, UG, sUG, nullptr, bsrc, bstride, bcomp, m->ncp, AB, sUG);
			else
				k_rows_times_range<false, 8><<<dim3((r_hi - r_mid) / 4, nchunk_cols, count), 256, 0, st>>>(
				    L, (long long)m->mat, m->npad, r_mid, r_lo, r_mid, UG, sUG, nullptr, bsrc, bstride, bcomp, m->ncp, AB, sUG);
		}
	}
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_gram_partial<<<dim3(m->nblk, count), 256, (size_t)TB * m->ncp * sizeof(double), st>>>(
		    UG, sUG, m->ncp, m->p + 1, m->dGramPart + (size_t)s0 * m->nblk * m->ncp * m->ncp, m->nblk);
	}
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_small<<<count, 256, 2 * (size_t)(m->p + 1) * (m->p + 2) * sizeof(double), st>>>(m->dGramPart + (size_t)s0 * m->nblk * m->ncp * m->ncp, m->nblk, UG, sUG, m->ncp, m->p, m->n,
		                               m->npad, m->dLogdet + (size_t)s0 * m->nblk, m->nblk, m->dInfo + s0,
		                               m->dRes + (size_t)s0 * RES_STRIDE, m->dConsts + (size_t)s0 * CONST_STRIDE, emulator_mode,
		                               m->dMinv + (size_t)s0 * MAXNCP * MAXNCP);
	}
}

// AB[:, first ncols8 chunks] = W^T UG
// nchunk_cols = 0: column 0 only (alpha = W^T u)
static void run_wt_times(emub_model *m, cudaStream_t st, int s0, int count, int nchunk_cols)
{
	const long long sUG = (long long)m->npad * m->ncp;
	LaunchScope ls(m->ctx, EMUB_K_SKINNY, count * 4.0 * (double)m->mat * (nchunk_cols ? nchunk_cols : 1), st);
	if (nchunk_cols == 0) {
		k_cols_times<true, 1><<<dim3(m->npad / 32, 1, count), COLS_TIMES_WARPS(1) * 32, 0, st>>>(
		    m->bufW + (size_t)s0 * m->mat, (long long)m->mat, m->npad, m->npad, m->dUG + (size_t)s0 * sUG, sUG, m->ncp,
		    m->dAB + (size_t)s0 * sUG, sUG);
		return;
	}
	k_cols_times<true><<<dim3(m->npad / 32, nchunk_cols, count), 256, 0, st>>>(
	    m->bufW + (size_t)s0 * m->mat, (long long)m->mat, m->npad, m->npad, m->dUG + (size_t)s0 * sUG, sUG, m->ncp,
	    m->dAB + (size_t)s0 * sUG, sUG);
}

static void run_gradient(emub_model *m, cudaStream_t st, int s0, int count)
{
	emub_ctx *c = m->ctx;
	const long long sUG = (long long)m->npad * m->ncp;
	const int nt64 = m->npad / CT;
	const size_t ntl = (size_t)nt64 * (nt64 + 1) / 2;
	const size_t smem = (2 * (size_t)m->d * CT + 2 * CT) * sizeof(double);
	double *Cinv = m->bufA + (size_t)s0 * m->mat;
	double *AB = m->dAB + (size_t)s0 * sUG;
	double *part = m->dGradPart + (size_t)s0 * ntl * m->d;
	const double *consts = m->dConsts + (size_t)s0 * CONST_STRIDE;
	{
		LaunchScope ls(c, EMUB_K_GRAD, count * 4.0 * (double)m->mat, st);
		dim3 grid(nt64, nt64, count);
		const double *res = m->dRes + (size_t)s0 * RES_STRIDE;
#define EMUB_GRAD_LAUNCH(KN, EX) k_grad_tiles<KN, EX><<<grid, 256, smem, st>>>(Cinv, (long long)m->mat, m->npad, AB, sUG, m->ncp, m->dX, m->n, m->d, consts, part, nt64, res, m->p)
		if (m->exact_grad) {
			switch (m->kernel) {
			case 2: EMUB_GRAD_LAUNCH(2, true); break;
			case 3: EMUB_GRAD_LAUNCH(3, true); break;
			default: EMUB_GRAD_LAUNCH(1, true); break;
			}
		} else {
			switch (m->kernel) {
			case 2: EMUB_GRAD_LAUNCH(2, false); break;
			case 3: EMUB_GRAD_LAUNCH(3, false); break;
			default: EMUB_GRAD_LAUNCH(1, false); break;
			}
		}
#undef EMUB_GRAD_LAUNCH
	}
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_grad_final<<<count, 256, 0, st>>>(part, nt64, Cinv, (long long)m->mat, m->npad, AB, sUG, m->ncp, m->n, m->d, m->kernel, consts,
		                                    m->dRes + (size_t)s0 * RES_STRIDE, m->exact_grad, m->p);
	}
}

// factorise + likelihood for slots [s0, s0+count) whose thetas are already in dThetas; the first ngrad of them also get
// the gradient (whole inverse, C^-1 = W^T W, fused reduction)
static void run_group(emub_model *m, cudaStream_t st, int s0, int count, int nth_in, int mode, int ngrad, int emulator_mode)
{
	emub_ctx *c = m->ctx;
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_theta_prep<<<(count + 63) / 64, 64, 0, st>>>(m->dThetas + (size_t)s0 * nth_in, count, nth_in, m->kernel, m->d, mode,
		                                               m->dConsts + (size_t)s0 * CONST_STRIDE);
	}
	cudaMemsetAsync(m->dInfo + s0, 0, sizeof(int) * count, st);
	launch_cov(m, st, count, m->dConsts + (size_t)s0 * CONST_STRIDE, m->bufA + (size_t)s0 * m->mat, (long long)m->mat, 1);
	// the whole triangular inverse is only needed for C^-1 (gradient) and for prediction (W K)
	run_factor(m, st, s0, count, emulator_mode ? count : ngrad);
	run_regression(m, st, s0, count, emulator_mode);
	if (ngrad > 0) {
		run_lauum(m, st, s0, ngrad);
		// alpha = W^T u; the exact-gradient mode also reads C^-1 H = W^T G (every column chunk)
		run_wt_times(m, st, s0, ngrad, m->exact_grad ? (m->p + 1 + 7) / 8 : 0);
		run_gradient(m, st, s0, ngrad);
	}
}

// how the slots of a chunk are shared out: group g owns cnt(g) consecutive slots, the first ngl(g) of which evaluate
// the gradient (the host front orders the points of a chunk accordingly, chunk_slot_order)
static inline int group_count(int total, int ng, int g) { return total / ng + (g < total % ng ? 1 : 0); }

// split `count` slots over the context's stream groups and run them concurrently
// the launch sequence of one chunk: fork stream 0 to the other groups, run them, join back
static int issue_chunk(emub_model *m, int count, int nth_in, int mode, int ngrad, int emulator_mode, int ng)
{
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaEventRecord(c->gev[0], c->streams[0]));
	int s0 = 0;
	for (int g = 0; g < ng; g++) {
		const int cnt = group_count(count, ng, g);
		if (g > 0) CUDA_TRY(cudaStreamWaitEvent(c->streams[g], c->gev[0], 0));
		run_group(m, c->streams[g], s0, cnt, nth_in, mode, group_count(ngrad, ng, g), emulator_mode);
		s0 += cnt;
		if (g > 0) {
			CUDA_TRY(cudaEventRecord(c->gev[g], c->streams[g]));
			CUDA_TRY(cudaStreamWaitEvent(c->streams[0], c->gev[g], 0));
		}
	}
	return EMUB_OK;
}

// One chunk = a fixed sequence of ~160 dependent launches per stream group whose arguments (device buffers,
// sizes) depend only on (count, theta layout, gradient / emulator flags, groups): it is captured ONCE into a CUDA
// graph (the fork/join events between the group streams become graph edges) and replayed afterwards, so the
// host pays one launch per chunk and the small dependent kernels of the recursion run back to back.
static int chunk_groups(const emub_model *m, int count) { return m->ctx->profile ? 1 : std::min(m->ctx->ngroups, count); }

// ngrad: how many of the count points evaluate the gradient (they sit first in every group's slot range)
static int run_chunk(emub_model *m, int count, int nth_in, int mode, int ngrad, int emulator_mode)
{
	emub_ctx *c = m->ctx;
	const int ng = chunk_groups(m, count);
	m->last_count = count;
	if (c->profile || !c->use_graphs) {
		int rc = issue_chunk(m, count, nth_in, mode, ngrad, emulator_mode, ng);
		if (rc) return rc;
		CUDA_TRY(cudaGetLastError());
		return EMUB_OK;
	}
	const unsigned long long key = ((unsigned long long)count << 48) | ((unsigned long long)ngrad << 32) | ((unsigned)nth_in << 16) |
	                               ((unsigned)ng << 8) | ((unsigned)(m->exact_grad != 0) << 4) | ((unsigned)mode << 2) |
	                               (unsigned)(emulator_mode != 0);
	cudaGraphExec_t exec = nullptr;
	long long nlaunch = 0;
	for (size_t i = 0; i < m->graphs.size(); i++)
		if (m->graphs[i].first == key) {
			exec = m->graphs[i].second;
			nlaunch = m->graph_launches[i].second;
			if (i + 1 < m->graphs.size()) {  // most recently used last
				std::rotate(m->graphs.begin() + i, m->graphs.begin() + i + 1, m->graphs.end());
				std::rotate(m->graph_launches.begin() + i, m->graph_launches.begin() + i + 1, m->graph_launches.end());
			}
			break;
		}
	if (!exec) {
		// a front of restart chains produces many (count, ngrad) shapes: keep the most recently used ones
		constexpr size_t GRAPH_CACHE = 96;
		if (m->graphs.size() >= GRAPH_CACHE) {
			CUDA_TRY(cudaStreamSynchronize(c->streams[0]));
			cudaGraphExecDestroy(m->graphs.front().second);
			m->graphs.erase(m->graphs.begin());
			m->graph_launches.erase(m->graph_launches.begin());
		}
		const long long launches_before = c->launches;
		cudaGraph_t graph = nullptr;
		CUDA_TRY(cudaStreamBeginCapture(c->streams[0], cudaStreamCaptureModeThreadLocal));
		int rc = issue_chunk(m, count, nth_in, mode, ngrad, emulator_mode, ng);
		cudaError_t ce = cudaStreamEndCapture(c->streams[0], &graph);
		if (rc != EMUB_OK || ce != cudaSuccess || !graph) {
			if (graph) cudaGraphDestroy(graph);
			cudaGetLastError();
			return rc != EMUB_OK ? rc : set_err(EMUB_ECUDA, "CUDA graph capture failed: %s", cudaGetErrorString(ce));
		}
		ce = cudaGraphInstantiate(&exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ce != cudaSuccess) return set_err(EMUB_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce));
		nlaunch = c->launches - launches_before;
		m->graphs.push_back({key, exec});
		m->graph_launches.push_back({key, nlaunch});
		c->launches = launches_before;  // counted per replay below
	}
	c->launches += nlaunch;
	CUDA_TRY(cudaGraphLaunch(exec, c->streams[0]));
	return EMUB_OK;
}

// component of every slot of the chunk -> dComp (NULL: component 0)
static int upload_components(emub_model *m, const int *comp, int done, int count, cudaStream_t st)
{
	for (int b = 0; b < count; b++) {
		const int cidx = comp ? comp[done + b] : 0;
		if (cidx < 0 || cidx >= m->ncomp) return set_err(EMUB_EINVAL, "emub: component index out of range%s");
		m->hComp[b] = cidx;
	}
	CUDA_TRY(cudaMemcpyAsync(m->dComp, m->hComp, sizeof(int) * count, cudaMemcpyHostToDevice, st));
	return EMUB_OK;
}

extern "C" int emub_loglik_grad_batch_dev(emub_model *m, const double *d_thetas, int B, int want_grad, double *d_out)
{
	if (!m || !d_thetas || !d_out || B < 0) return set_err(EMUB_EINVAL, "emub_loglik_grad_batch_dev: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const int nth1 = m->nth - 1;
	cudaStream_t st = c->streams[0];
	for (int done = 0; done < B; done += m->nslots) {
		const int count = std::min(m->nslots, B - done);
		int rc = upload_components(m, nullptr, done, count, st);
		if (rc) return rc;
		CUDA_TRY(cudaMemcpyAsync(m->dThetas, d_thetas + (size_t)done * nth1, sizeof(double) * count * nth1, cudaMemcpyDeviceToDevice, st));
		rc = run_chunk(m, count, nth1, THETA_LIK, want_grad ? count : 0, 0);
		if (rc) return rc;
		{
			LaunchScope ls(c, EMUB_K_SMALL, 0, st);
			k_pack_results<<<(count + 63) / 64, 64, 0, st>>>(m->dRes, count, nth1, d_out + (size_t)done * (nth1 + 4));
		}
	}
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

// The host-pointer evaluation: want[b] != 0 asks for the gradient of point b (want == NULL: want_all for every point).
// Within a chunk the gradient points are placed first in every stream group's slot range, so that the launches of the
// gradient-only stages (spine merges of the inverse, W^T W, the fused reduction) cover a contiguous run of slots; the
// results go back to the caller's order.  Values do not depend on where a point sits.
static int loglik_batch_host(emub_model *m, const double *thetas, const int *comp, const int *want, int want_all, int B,
                             double *negL, double *grad, double *sigma2, int *status)
{
	if (!m || !thetas || B < 0) return set_err(EMUB_EINVAL, "emub_loglik_grad_batch: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	const int nth1 = m->nth - 1;
	cudaStream_t st = c->streams[0];
	std::vector<int> perm((size_t)std::max(1, std::min(B, m->nslots)));
	for (int done = 0; done < B; done += m->nslots) {
		const int count = std::min(m->nslots, B - done);
		const int ng = chunk_groups(m, count);
		int ngrad = 0;
		for (int b = 0; b < count; b++) ngrad += (want ? want[done + b] != 0 : want_all != 0);
		// slot -> point: group g takes its share of the gradient points, then its share of the others
		{
			int ig = 0, iv = 0, slot = 0;
			auto next = [&](bool g) {
				int &i = g ? ig : iv;
				while ((want ? want[done + i] != 0 : want_all != 0) != g) i++;
				return i++;
			};
			for (int g = 0; g < ng; g++) {
				const int cnt = group_count(count, ng, g), ngl = group_count(ngrad, ng, g);
				for (int k = 0; k < cnt; k++) perm[slot++] = next(k < ngl);
			}
		}
		for (int s = 0; s < count; s++) {
			const int b = done + perm[s];
			const int cidx = comp ? comp[b] : 0;
			if (cidx < 0 || cidx >= m->ncomp) return set_err(EMUB_EINVAL, "emub: component index out of range%s");
			m->hComp[s] = cidx;
			memcpy(m->hThetas + (size_t)s * nth1, thetas + (size_t)b * nth1, sizeof(double) * nth1);
		}
		CUDA_TRY(cudaMemcpyAsync(m->dComp, m->hComp, sizeof(int) * count, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(m->dThetas, m->hThetas, sizeof(double) * count * nth1, cudaMemcpyHostToDevice, st));
		int rc = run_chunk(m, count, nth1, THETA_LIK, ngrad, 0);
		if (rc) return rc;
		CUDA_TRY(cudaMemcpyAsync(m->hRes, m->dRes, sizeof(double) * count * RES_STRIDE, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		for (int s = 0; s < count; s++) {
			const int b = done + perm[s];
			const double *r = m->hRes + (size_t)s * RES_STRIDE;
			const int stt = (int)r[2];
			const bool wg = want ? want[b] != 0 : want_all != 0;
			if (negL) negL[b] = r[0];
			if (sigma2) sigma2[b] = r[1];
			if (status) status[b] = stt;
			if (grad)
				for (int k = 0; k < nth1; k++) grad[(size_t)b * nth1 + k] = wg ? r[RES_GRAD + k] : (stt ? nan("") : 0.0);
		}
		// emub_loglik_extras addresses the last chunk by point index: undo the slot order in place
		if (count > 1) {
			std::vector<double> tmp((size_t)count * RES_STRIDE);
			for (int s = 0; s < count; s++) memcpy(&tmp[(size_t)perm[s] * RES_STRIDE], m->hRes + (size_t)s * RES_STRIDE, sizeof(double) * RES_STRIDE);
			memcpy(m->hRes, tmp.data(), sizeof(double) * tmp.size());
		}
	}
	return EMUB_OK;
}

extern "C" int emub_loglik_grad_batch_comp(emub_model *m, const double *thetas, const int *comp, int B, int want_grad,
                                           double *negL, double *grad, double *sigma2, int *status)
{
	return loglik_batch_host(m, thetas, comp, nullptr, want_grad, B, negL, grad, sigma2, status);
}

extern "C" int emub_loglik_grad_batch_mixed(emub_model *m, const double *thetas, const int *comp, const int *want_grad, int B,
                                            double *negL, double *grad, double *sigma2, int *status)
{
	if (!want_grad) return set_err(EMUB_EINVAL, "emub_loglik_grad_batch_mixed: null want_grad%s");
	return loglik_batch_host(m, thetas, comp, want_grad, 0, B, negL, grad, sigma2, status);
}

extern "C" int emub_loglik_grad_batch(emub_model *m, const double *thetas, int B, int want_grad, double *negL, double *grad,
                                      double *sigma2, int *status)
{
	return emub_loglik_grad_batch_comp(m, thetas, nullptr, B, want_grad, negL, grad, sigma2, status);
}

extern "C" int emub_loglik_extras(emub_model *m, int b, double *logdet, double *beta)
{
	if (!m || b < 0 || b >= m->last_count) return set_err(EMUB_EINVAL, "emub_loglik_extras: bad slot%s");
	const double *r = m->hRes + (size_t)b * RES_STRIDE;
	if (logdet) *logdet = r[3];
	if (beta) for (int c = 0; c < m->p; c++) beta[c] = r[RES_BETA + c];
	return EMUB_OK;
}

// ---- compatibility entry points (not on the hot path) -------------------------------------------------
extern "C" int emub_cov_matrix(emub_model *m, const double *thetas, double *C, int ldc)
{
	if (!m || !thetas || !C || ldc < m->n) return set_err(EMUB_EINVAL, "emub_cov_matrix: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t st = c->streams[0];
	CUDA_TRY(cudaMemcpyAsync(m->dThetas, thetas, sizeof(double) * m->nth, cudaMemcpyHostToDevice, st));
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_theta_prep<<<1, 64, 0, st>>>(m->dThetas, 1, m->nth, m->kernel, m->d, THETA_FULL, m->dConsts);
	}
	launch_cov(m, st, 1, m->dConsts, m->bufT, (long long)m->mat, 0);
	CUDA_TRY(cudaMemcpy2DAsync(C, sizeof(double) * ldc, m->bufT, sizeof(double) * m->npad, sizeof(double) * m->n, m->n,
	                           cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

extern "C" int emub_h_matrix(emub_model *m, double *H, int ldh)
{
	if (!m || !H || ldh < m->p) return set_err(EMUB_EINVAL, "emub_h_matrix: bad argument%s");
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	CUDA_TRY(cudaMemcpy2D(H, sizeof(double) * ldh, m->dYh + 1, sizeof(double) * m->ncp, sizeof(double) * m->p, m->n, cudaMemcpyDeviceToHost));
	return EMUB_OK;
}

extern "C" int emub_k_vectors(emub_model *m, const double *thetas, const double *pts, int ldp, int mq, double *K, int ldk)
{
	if (!m || !thetas || !pts || !K || mq < 1 || ldp < m->d || ldk < mq) return set_err(EMUB_EINVAL, "emub_k_vectors: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t st = c->streams[0];
	const int mq_pad = (mq + TB - 1) / TB * TB;
	ScopedDev sQ, sK;
	CUDA_TRY(sQ.alloc((size_t)mq_pad * m->d));
	CUDA_TRY(sK.alloc((size_t)m->npad * mq_pad));
	double *dQ = sQ.p, *dK = sK.p;
	CUDA_TRY(cudaMemcpy2DAsync(dQ, sizeof(double) * m->d, pts, sizeof(double) * ldp, sizeof(double) * m->d, mq, cudaMemcpyHostToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(m->dThetas, thetas, sizeof(double) * m->nth, cudaMemcpyHostToDevice, st));
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_theta_prep<<<1, 64, 0, st>>>(m->dThetas, 1, m->nth, m->kernel, m->d, THETA_FULL, m->dConsts);
	}
	launch_kcross(m, st, m->dConsts, dQ, mq, mq_pad, dK, mq_pad);
	CUDA_TRY(cudaMemcpy2DAsync(K, sizeof(double) * ldk, dK, sizeof(double) * mq_pad, sizeof(double) * mq, m->n, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

// chol_inverse_cov_matrix (emulate-fns.c:275-300) for a caller-owned n x n matrix (n = the model's nmodel_points; the
// design of the model is not used): Cholesky, log-determinant, explicit inverse on the engine's factorisation slot 0.
extern "C" int emub_spd_inverse(emub_model *m, const double *A, int lda, double *Ainv, int ldi, double *logdet)
{
	if (!m || !A || !Ainv || lda < m->n || ldi < m->n) return set_err(EMUB_EINVAL, "emub_spd_inverse: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t st = c->streams[0];
	const int n = m->n, npad = m->npad;
	CUDA_TRY(cudaMemsetAsync(m->bufA, 0, m->mat * sizeof(double), st));
	CUDA_TRY(cudaMemcpy2DAsync(m->bufA, sizeof(double) * npad, A, sizeof(double) * lda, sizeof(double) * n, n, cudaMemcpyHostToDevice, st));
	if (npad > n) {
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_pad_identity<<<(npad - n + 127) / 128, 128, 0, st>>>(m->bufA, npad, n);
	}
	CUDA_TRY(cudaMemsetAsync(m->dInfo, 0, sizeof(int), st));
	run_factor(m, st, 0, 1, 1);
	run_lauum(m, st, 0, 1);
	CUDA_TRY(cudaMemcpy2DAsync(Ainv, sizeof(double) * ldi, m->bufA, sizeof(double) * npad, sizeof(double) * n, n, cudaMemcpyDeviceToHost, st));
	std::vector<double> parts(m->nblk);
	int info = 0;
	CUDA_TRY(cudaMemcpyAsync(parts.data(), m->dLogdet, sizeof(double) * m->nblk, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaMemcpyAsync(&info, m->dInfo, sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	if (info) return set_err(EMUB_EDOM, "emub_spd_inverse: matrix not positive definite%s");
	double sum = 0.0;
	for (double v : parts) sum += v;
	if (logdet) *logdet = 2.0 * sum;
	// the engine keeps the lower tiles: mirror them (tiles above the diagonal were never written)
	for (int i = 0; i < n; i++)
		for (int j = i + 1; j < n; j++) Ainv[(size_t)i * ldi + j] = Ainv[(size_t)j * ldi + i];
	return EMUB_OK;
}

extern "C" int emub_debug_fetch(emub_model *m, int b, int which, double *out, int ldo)
{
	if (!m || !out || b < 0 || b >= m->nslots || ldo < m->n) return set_err(EMUB_EINVAL, "emub_debug_fetch: bad argument%s");
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	CUDA_TRY(cudaDeviceSynchronize());
	const double *src = (which == 0 ? m->bufA : (which == 1 ? m->bufW : m->bufT)) + (size_t)b * m->mat;
	CUDA_TRY(cudaMemcpy2D(out, sizeof(double) * ldo, src, sizeof(double) * m->npad, sizeof(double) * m->n, m->n, cudaMemcpyDeviceToHost));
	return EMUB_OK;
}

extern "C" int emub_debug_exp(emub_ctx *c, const double *x, int n, double *out)
{
	if (!c || !x || !out || n < 1) return set_err(EMUB_EINVAL, "emub_debug_exp: bad argument%s");
	CUDA_TRY(cudaSetDevice(c->device));
	ScopedDev sx, sout;
	CUDA_TRY(sx.alloc((size_t)n));
	CUDA_TRY(sout.alloc((size_t)n));
	double *dx = sx.p, *dout = sout.p;
	CUDA_TRY(upload(dx, x, sizeof(double) * n, c->streams[0]));
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, c->streams[0]);
		k_debug_exp<<<(n + 255) / 256, 256, 0, c->streams[0]>>>(dx, n, dout);
	}
	CUDA_TRY(cudaStreamSynchronize(c->streams[0]));
	CUDA_TRY(cudaMemcpy(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost));
	return EMUB_OK;
}

extern "C" int emub_debug_exp_scaled(emub_ctx *c, const double *x, int n, double *out)
{
	if (!c || !x || !out || n < 1) return set_err(EMUB_EINVAL, "emub_debug_exp_scaled: bad argument%s");
	CUDA_TRY(cudaSetDevice(c->device));
	ScopedDev sx, sout;
	CUDA_TRY(sx.alloc((size_t)n));
	CUDA_TRY(sout.alloc((size_t)n));
	CUDA_TRY(upload(sx.p, x, sizeof(double) * n, c->streams[0]));
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, c->streams[0]);
		k_debug_exp_scaled<<<(n + 255) / 256, 256, 0, c->streams[0]>>>(sx.p, n, sout.p);
	}
	CUDA_TRY(cudaStreamSynchronize(c->streams[0]));
	CUDA_TRY(cudaMemcpy(out, sout.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
	return EMUB_OK;
}

extern "C" int emub_debug_cholesky(emub_model *m, const double *theta_less_amp, double *L, int ldl, double *logdet)
{
	if (!m || !theta_less_amp || !L || ldl < m->n) return set_err(EMUB_EINVAL, "emub_debug_cholesky: bad argument%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t st = c->streams[0];
	const int nth1 = m->nth - 1;
	CUDA_TRY(cudaMemcpyAsync(m->dThetas, theta_less_amp, sizeof(double) * nth1, cudaMemcpyHostToDevice, st));
	{
		LaunchScope ls(c, EMUB_K_SMALL, 0, st);
		k_theta_prep<<<1, 64, 0, st>>>(m->dThetas, 1, nth1, m->kernel, m->d, THETA_LIK, m->dConsts);
	}
	CUDA_TRY(cudaMemsetAsync(m->dInfo, 0, sizeof(int), st));
	launch_cov(m, st, 1, m->dConsts, m->bufA, (long long)m->mat, 1);
	run_factor(m, st, 0, 1, 1);
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpy2D(L, sizeof(double) * ldl, m->bufT, sizeof(double) * m->npad, sizeof(double) * m->n, m->n, cudaMemcpyDeviceToHost));
	std::vector<double> parts(m->nblk);
	int info = 0;
	CUDA_TRY(cudaMemcpy(parts.data(), m->dLogdet, sizeof(double) * m->nblk, cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemcpy(&info, m->dInfo, sizeof(int), cudaMemcpyDeviceToHost));
	double s = 0;
	for (double v : parts) s += v;
	if (logdet) *logdet = 2.0 * s;
	// zero the strictly upper part for the caller (upper off-diagonal tiles are never written)
	for (int i = 0; i < m->n; i++) for (int j = i + 1; j < m->n; j++) L[(size_t)i * ldl + j] = 0.0;
	return info ? EMUB_EDOM : EMUB_OK;
}

// ---- prediction ------------------------------------------------------------------------------------------
extern "C" void emub_emulator_destroy(emub_emulator *e);
// every allocation of the workspace; on a failure the caller (ensure_query_ws) tears the half-built workspace down
static int build_query_ws(emub_model *m)
{
	QueryWs *w = new QueryWs();
	memset(w, 0, sizeof(*w));
	m->qws = w;
	// chunk of up to 16384 points, K (npad x mqc) bounded to ~1 GiB
	int mqc = 16384;
	while (mqc > TB && (size_t)m->npad * mqc * sizeof(double) > ((size_t)1 << 30)) mqc /= 2;
	w->mqc = mqc;
	w->ncomp = m->ncomp;
	CUDA_TRY(cudaMalloc(&w->dQ, sizeof(double) * (size_t)mqc * m->d));
	CUDA_TRY(cudaMalloc(&w->dK, sizeof(double) * (size_t)m->npad * mqc));
	CUDA_TRY(cudaMalloc(&w->dVsq, sizeof(double) * (size_t)m->nblk * DefaultCfg::SUBM * mqc));
	CUDA_TRY(cudaMalloc(&w->dKA, sizeof(double) * (size_t)mqc * m->ncp));
	CUDA_TRY(cudaMalloc(&w->dMean, sizeof(double) * (size_t)w->ncomp * mqc));
	CUDA_TRY(cudaMalloc(&w->dVar, sizeof(double) * (size_t)w->ncomp * mqc));
	CUDA_TRY(cudaMalloc(&w->dFew, sizeof(double) * (size_t)((m->npad + FEW_JC - 1) / FEW_JC) * m->npad * 8));
	w->few_cap = 1;
	w->ntcap = NTMAX;
	CUDA_TRY(cudaMalloc(&w->dOutM, sizeof(double) * (size_t)NTMAX * mqc));
	CUDA_TRY(cudaMalloc(&w->dOutV, sizeof(double) * (size_t)NTMAX * mqc));
	CUDA_TRY(cudaMalloc(&w->dProj, sizeof(double) * (size_t)(NTMAX + NTMAX * NTMAX + NTMAX)));
	CUDA_TRY(cudaMallocHost(&w->hQ, sizeof(double) * (size_t)mqc * m->d));
	CUDA_TRY(cudaMallocHost(&w->hOut, sizeof(double) * 2 * (size_t)std::max(NTMAX, 1) * mqc));
	CUDA_TRY(cudaMalloc(&w->dQb[1], sizeof(double) * (size_t)mqc * m->d));
	CUDA_TRY(cudaMallocHost(&w->hQb[1], sizeof(double) * (size_t)mqc * m->d));
	CUDA_TRY(cudaMallocHost(&w->hOutb[1], sizeof(double) * 2 * (size_t)std::max(NTMAX, 1) * mqc));
	w->dQb[0] = w->dQ; w->hQb[0] = w->hQ; w->hOutb[0] = w->hOut;
	CUDA_TRY(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		CUDA_TRY(cudaEventCreateWithFlags(&w->evIn[i], cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&w->evOut[i], cudaEventDisableTiming));
	}
	// V = W K, one task per row block (longest K first); the batch dimension walks the query blocks
	std::vector<GemmTask> tasks;
	for (int i = m->nblk - 1; i >= 0; i--) tasks.push_back({(long long)i * TB * m->npad, 0, 0, (i + 1) * TB, i | TASK_TRIM_END_SR0});
	CUDA_TRY(cudaMalloc(&w->dTasks, tasks.size() * sizeof(GemmTask)));
	CUDA_TRY(upload(w->dTasks, tasks.data(), tasks.size() * sizeof(GemmTask), m->ctx->streams[0]));
	return EMUB_OK;
}

// The workspace is installed only when every allocation succeeded: a failed cudaMalloc (dK alone is up to 1 GiB)
// must not leave a workspace with null device pointers behind for the next predict call to launch kernels on.
static int ensure_query_ws(emub_model *m)
{
	if (m->qws) return EMUB_OK;
	const int rc = build_query_ws(m);
	if (rc != EMUB_OK) {
		free_query_ws(m);  // frees whatever was allocated and resets m->qws
		cudaGetLastError();
	}
	return rc;
}

// back-projected outputs of a chunk are mqc x nt on the device and twice that in each pinned host buffer: grow them
// when a model has more observables (or components) than the buffers were sized for
static int ensure_output_capacity(emub_model *m, int need)
{
	QueryWs *w = m->qws;
	if (need <= w->ntcap) return EMUB_OK;
	const int cap = (need + 63) / 64 * 64;
	CUDA_TRY(cudaStreamSynchronize(m->ctx->streams[0]));
	cudaFree(w->dOutM); cudaFree(w->dOutV); cudaFree(w->dProj);
	cudaFreeHost(w->hOutb[0]); cudaFreeHost(w->hOutb[1]);
	w->dOutM = w->dOutV = w->dProj = nullptr;
	w->proj_hash = 0;
	few_graphs_clear(w);
	w->hOut = w->hOutb[0] = w->hOutb[1] = nullptr;
	w->ntcap = 0;
	CUDA_TRY(cudaMalloc(&w->dOutM, sizeof(double) * (size_t)cap * w->mqc));
	CUDA_TRY(cudaMalloc(&w->dOutV, sizeof(double) * (size_t)cap * w->mqc));
	CUDA_TRY(cudaMalloc(&w->dProj, sizeof(double) * ((size_t)cap + (size_t)cap * cap + cap)));
	CUDA_TRY(cudaMallocHost(&w->hOutb[0], sizeof(double) * 2 * (size_t)cap * w->mqc));
	CUDA_TRY(cudaMallocHost(&w->hOutb[1], sizeof(double) * 2 * (size_t)cap * w->mqc));
	w->hOut = w->hOutb[0];
	w->ntcap = cap;
	return EMUB_OK;
}

extern "C" int emub_emulator_create_comp(emub_model *m, int comp, const double *thetas, emub_emulator **out)
{
	if (!m || !thetas || !out) return set_err(EMUB_EINVAL, "emub_emulator_create: null%s");
	if (comp < 0 || comp >= m->ncomp) return set_err(EMUB_EINVAL, "emub_emulator_create: component index out of range%s");
	emub_ctx *c = m->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	cudaStream_t st = c->streams[0];
	int rc = upload_components(m, &comp, 0, 1, st);
	if (rc) return rc;
	CUDA_TRY(cudaMemcpyAsync(m->dThetas, thetas, sizeof(double) * m->nth, cudaMemcpyHostToDevice, st));
	// factorise in slot 0 on stream 0 only
	const int saved_groups = c->ngroups;
	c->ngroups = 1;
	rc = run_chunk(m, 1, m->nth, THETA_FULL, 0, 1);
	c->ngroups = saved_groups;
	if (rc) return rc;
	run_wt_times(m, st, 0, 1, (m->p + 1 + 7) / 8);
	CUDA_TRY(cudaMemcpyAsync(m->hRes, m->dRes, sizeof(double) * RES_STRIDE, cudaMemcpyDeviceToHost, st));
	double hc[4];
	CUDA_TRY(cudaMemcpyAsync(hc, m->dConsts, sizeof(hc), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	const int status = (int)m->hRes[2];
	if (status == 1) return set_err(EMUB_EDOM, "emub_emulator_create: covariance matrix not positive definite%s");
	if (status == 2) return set_err(EMUB_EREG, "emub_emulator_create: regression matrix not positive definite%s");
	rc = ensure_query_ws(m);
	if (rc) return rc;
	emub_emulator *e = new emub_emulator();
	memset(e, 0, sizeof(*e));
	e->m = m;
	e->kappa = hc[0] + hc[1];  // c(x*, x*) = amp + nugget   (emulator_struct.c:135)
	e->uid = g_emulator_uid.fetch_add(1);
	for (int i = 0; i < m->p; i++) e->hbeta[i] = m->hRes[RES_BETA + i];
	const size_t sUG = (size_t)m->npad * m->ncp;
	if (cudaMalloc(&e->W, m->mat * sizeof(double)) != cudaSuccess || cudaMalloc(&e->AB, sUG * sizeof(double)) != cudaSuccess ||
	    cudaMalloc(&e->beta, MAXNCP * sizeof(double)) != cudaSuccess || cudaMalloc(&e->Minv, MAXNCP * MAXNCP * sizeof(double)) != cudaSuccess ||
	    cudaMalloc(&e->consts, CONST_STRIDE * sizeof(double)) != cudaSuccess) {
		emub_emulator_destroy(e);
		return set_err(EMUB_ENOMEM, "emub_emulator_create: out of device memory%s");
	}
	CUDA_TRY(cudaMemcpyAsync(e->W, m->bufW, m->mat * sizeof(double), cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(e->AB, m->dAB, sUG * sizeof(double), cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(e->beta, m->dRes + RES_BETA, MAXNCP * sizeof(double), cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(e->Minv, m->dMinv, MAXNCP * MAXNCP * sizeof(double), cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaMemcpyAsync(e->consts, m->dConsts, CONST_STRIDE * sizeof(double), cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	*out = e;
	return EMUB_OK;
}

extern "C" int emub_emulator_create(emub_model *m, const double *thetas, emub_emulator **out)
{
	return emub_emulator_create_comp(m, 0, thetas, out);
}

extern "C" void emub_emulator_destroy(emub_emulator *e)
{
	if (!e) return;
	cudaSetDevice(e->m->ctx->device);
	cudaDeviceSynchronize();
	cudaFree(e->W); cudaFree(e->AB); cudaFree(e->beta); cudaFree(e->Minv); cudaFree(e->consts);
	delete e;
}

extern "C" int emub_emulator_beta(emub_emulator *e, double *beta)
{
	if (!e || !beta) return EMUB_EINVAL;
	for (int i = 0; i < e->m->p; i++) beta[i] = e->hbeta[i];
	return EMUB_OK;
}

// one chunk (mq <= mqc) whose points are already at dQ (contiguous mq x d)
static int predict_chunk(emub_emulator *e, cudaStream_t st, const double *dQ, int mq, double *dMean, double *dVar, bool few = false)
{
	emub_model *m = e->m;
	emub_ctx *c = m->ctx;
	QueryWs *w = m->qws;
	const int mq_pad = (mq + TB - 1) / TB * TB;
	const int ldk = w->mqc;
	launch_kcross(m, st, e->consts, dQ, mq, mq_pad, w->dK, ldk);
	if (few && mq <= FEW_MAX && FEW_MAX * (m->npad / FEW_ROWS) <= w->mqc) {
		// latency path: W K[:, 0:8] as a row- and column-split skinny product, one CTA for the reductions
		{
			LaunchScope ls(c, EMUB_K_GEMM_PRED, (double)m->npad * m->npad * 8, st);
			const dim3 grid(m->npad / 32, (m->npad + FEW_JC - 1) / FEW_JC);
			if (mq == 1) k_few_wk<1><<<grid, 256, 0, st>>>(e->W, m->npad, w->dK, ldk, w->dFew, m->npad, nullptr, 0, 0);
			else k_few_wk<8><<<grid, 256, 0, st>>>(e->W, m->npad, w->dK, ldk, w->dFew, m->npad, nullptr, 0, 0);
		}
		{
			LaunchScope ls(c, EMUB_K_SKINNY, 8.0 * (double)m->npad * (8 + m->ncp), st);
			k_few_finish<<<m->npad / FEW_ROWS, FEW_ROWS, 0, st>>>(w->dFew, m->npad, w->dK, ldk, e->AB, m->ncp, w->dVsq, ldk, w->dKA, nullptr, 0, 0, 0);
		}
		{
			LaunchScope ls(c, EMUB_K_PRED_FINAL, 0, st);
			k_pred_final<<<1, 128, (size_t)m->p * m->p * sizeof(double), st>>>(dQ, mq, m->d, m->order, m->p, w->dKA, m->ncp, w->dVsq, m->npad / FEW_ROWS, ldk, e->beta, e->Minv,
			                                e->kappa, dMean, dVar, m->npad / FEW_ROWS, (long long)FEW_MAX * m->ncp, nullptr, 0, 0);
		}
		CUDA_TRY(cudaGetLastError());
		return EMUB_OK;
	}
	const int nqb = mq_pad / TB;
	// |W k|^2 partials: tile (row block i, query block qb)
	launch_gemm<KMAJOR, RMAJOR, EPI_COLSUMSQ>(c, EMUB_K_GEMM_PRED, (double)m->npad * (m->npad + TB / 4) * TB, st, w->dTasks, m->nblk, nqb,
	                                          e->W, 0, m->npad, w->dK, TB, ldk, w->dVsq, TB, ldk, 1.0);
	{
		const int nchunk_cols = (m->p + 1 + 7) / 8;
		LaunchScope ls(c, EMUB_K_SKINNY, 8.0 * (double)m->npad * mq_pad * nchunk_cols, st);
		if (m->p + 1 <= 2)  // regression order 0: K^T [a | C^-1 1]; k_pred_final reads the first p + 1 columns of KA only
			k_cols_times<false, 2><<<dim3(mq_pad / 32, 1, 1), COLS_TIMES_WARPS(2) * 32, 0, st>>>(w->dK, 0, ldk, m->npad, e->AB, 0, m->ncp, w->dKA, 0);
		else
			k_cols_times<false><<<dim3(mq_pad / 32, nchunk_cols, 1), 256, 0, st>>>(w->dK, 0, ldk, m->npad, e->AB, 0, m->ncp, w->dKA, 0);
	}
	{
		LaunchScope ls(c, EMUB_K_PRED_FINAL, 0, st);
		k_pred_final<<<(mq + 127) / 128, 128, (size_t)m->p * m->p * sizeof(double), st>>>(dQ, mq, m->d, m->order, m->p, w->dKA, m->ncp, w->dVsq, m->nblk * DefaultCfg::SUBM, ldk, e->beta,
		                                               e->Minv, e->kappa, dMean, dVar, 1, 0, nullptr, 0, 0);
	}
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

// The few-points path for nr emulators of one model at once: every kernel carries the component in its grid, so a
// point of a multivariate model costs 4 launches instead of 4 per component.  dMean / dVar: [nr][mqc].
static bool few_set_fits(const emub_model *m, int nr)
{
	const QueryWs *w = m->qws;
	const long long nparts = m->npad / FEW_ROWS;
	return (long long)nr * TB <= w->mqc && (long long)nr * nparts * FEW_MAX <= w->mqc;
}

// tables of the set (device buffers of every emulator, a contiguous copy of their constants): rebuilt when the set
// changes; *hash_out identifies the set
static int few_set_prepare(emub_emulator *const *emus, int nr, cudaStream_t st, unsigned long long *hash_out)
{
	emub_model *m = emus[0]->m;
	QueryWs *w = m->qws;
	const int nsplit = (m->npad + FEW_JC - 1) / FEW_JC;
	const long long tstride = (long long)nsplit * m->npad * 8;
	unsigned long long h = 1469598103934665603ull ^ (unsigned long long)nr;
	std::vector<FewSet> hs((size_t)nr);
	for (int j = 0; j < nr; j++) {
		hs[j] = FewSet{emus[j]->W, emus[j]->AB, emus[j]->beta, emus[j]->Minv, emus[j]->kappa};
		h = (h ^ emus[j]->uid) * 1099511628211ull;
	}
	if (h == 0) h = 1;
	if (h != w->few_hash) {
		w->few_hash = 0;
		if (nr > w->few_cap || !w->dFewSet) {
			CUDA_TRY(cudaStreamSynchronize(st));
			few_graphs_clear(w);
			cudaFree(w->dFewSet); cudaFree(w->dFewConsts); cudaFree(w->dFew);
			w->dFewSet = nullptr; w->dFewConsts = nullptr; w->dFew = nullptr;
			w->few_cap = 0;
			if (cudaMalloc(&w->dFewSet, sizeof(FewSet) * (size_t)nr) != cudaSuccess ||
			    cudaMalloc(&w->dFewConsts, sizeof(double) * (size_t)nr * CONST_STRIDE) != cudaSuccess ||
			    cudaMalloc(&w->dFew, sizeof(double) * (size_t)nr * tstride) != cudaSuccess) {
				// back to the state build_query_ws leaves: partial products for ONE emulator (emub_predict_few uses
				// dFew without the tables); if even that does not fit, drop the workspace so that the next call rebuilds it
				cudaGetLastError();
				cudaFree(w->dFewSet); cudaFree(w->dFewConsts); cudaFree(w->dFew);
				w->dFewSet = nullptr; w->dFewConsts = nullptr; w->dFew = nullptr;
				if (cudaMalloc(&w->dFew, sizeof(double) * (size_t)tstride) == cudaSuccess) w->few_cap = 1;
				else { cudaGetLastError(); free_query_ws(m); }
				return set_err(EMUB_ENOMEM, "emub_predict_multi_few: out of device memory for the component tables%s");
			}
			w->few_cap = nr;
		}
		CUDA_TRY(cudaMemcpyAsync(w->dFewSet, hs.data(), sizeof(FewSet) * (size_t)nr, cudaMemcpyHostToDevice, st));
		for (int j = 0; j < nr; j++)
			CUDA_TRY(cudaMemcpyAsync(w->dFewConsts + (size_t)j * CONST_STRIDE, emus[j]->consts, sizeof(double) * CONST_STRIDE,
			                         cudaMemcpyDeviceToDevice, st));
		CUDA_TRY(cudaStreamSynchronize(st));  // hs is a local
		w->few_hash = h;
	}
	*hash_out = h;
	return EMUB_OK;
}

// the kernels of the set path (capturable: no synchronisation, no host data)
static int few_set_launch(emub_emulator *const *emus, int nr, cudaStream_t st, const double *dQ, int mq, double *dMean, double *dVar)
{
	emub_model *m = emus[0]->m;
	emub_ctx *c = m->ctx;
	QueryWs *w = m->qws;
	const int ldk = w->mqc;
	const int nsplit = (m->npad + FEW_JC - 1) / FEW_JC, nparts = m->npad / FEW_ROWS;
	const long long tstride = (long long)nsplit * m->npad * 8, kastride = (long long)nparts * FEW_MAX * m->ncp;
	{
		// cross covariances of every component: component z at the columns [128 z, 128 z + 128) of dK
		const size_t smem = 2 * (size_t)m->d * CT * sizeof(double);
		LaunchScope ls(c, EMUB_K_KCROSS, 8.0 * (double)m->npad * TB * nr, st);
		dim3 grid(TB / CT, m->npad / CT, nr);
		switch (m->kernel) {
		case 2: k_cov<2, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, w->dFewConsts, CONST_STRIDE, w->dK, TB, ldk, 0); break;
		case 3: k_cov<3, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, w->dFewConsts, CONST_STRIDE, w->dK, TB, ldk, 0); break;
		default: k_cov<1, true><<<grid, 256, smem, st>>>(m->dX, m->n, m->d, dQ, mq, w->dFewConsts, CONST_STRIDE, w->dK, TB, ldk, 0); break;
		}
	}
	{
		LaunchScope ls(c, EMUB_K_GEMM_PRED, (double)m->npad * m->npad * 8 * nr, st);
		const dim3 grid(m->npad / 32, nsplit, nr);
		if (mq == 1) k_few_wk<1><<<grid, 256, 0, st>>>(nullptr, m->npad, w->dK, ldk, w->dFew, m->npad, w->dFewSet, TB, tstride);
		else k_few_wk<8><<<grid, 256, 0, st>>>(nullptr, m->npad, w->dK, ldk, w->dFew, m->npad, w->dFewSet, TB, tstride);
	}
	{
		LaunchScope ls(c, EMUB_K_SKINNY, 8.0 * (double)m->npad * (8 + m->ncp) * nr, st);
		k_few_finish<<<dim3(nparts, nr), FEW_ROWS, 0, st>>>(w->dFew, m->npad, w->dK, ldk, nullptr, m->ncp, w->dVsq, ldk, w->dKA, w->dFewSet, TB,
		                                                    tstride, kastride);
	}
	{
		LaunchScope ls(c, EMUB_K_PRED_FINAL, 0, st);
		k_pred_final<<<dim3(1, nr), 128, (size_t)m->p * m->p * sizeof(double), st>>>(dQ, mq, m->d, m->order, m->p, w->dKA, m->ncp, w->dVsq, nparts, ldk, nullptr, nullptr, 0.0, dMean,
		                                          dVar, nparts, (long long)FEW_MAX * m->ncp, w->dFewSet, kastride, w->mqc);
	}
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

extern "C" int emub_predict_batch_dev(emub_emulator *e, const double *d_pts, int mq, double *d_mean, double *d_var)
{
	if (!e || !d_pts || !d_mean || !d_var || mq < 0) return set_err(EMUB_EINVAL, "emub_predict_batch_dev: bad argument%s");
	emub_model *m = e->m;
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	{ int rc0 = ensure_query_ws(m); if (rc0) return rc0; }
	cudaStream_t st = m->ctx->streams[0];
	const int mqc = m->qws->mqc;
	for (int done = 0; done < mq; done += mqc) {
		const int cnt = std::min(mqc, mq - done);
		int rc = predict_chunk(e, st, d_pts + (size_t)done * m->d, cnt, d_mean + done, d_var + done);
		if (rc) return rc;
	}
	return EMUB_OK;
}

// Chunked, double-buffered walk over mq query points: packing + host-to-device copy of chunk i+1 (copy stream) and the
// copy-out of chunk i-1 (host) overlap the kernels of chunk i.  compute(cnt, dQ, hOut) enqueues the chunk's kernels and
// its device-to-host copies on streams[0]; collect(done, cnt, hOut) moves a finished chunk into the caller's arrays.
template <class Compute, class Collect>
static int walk_query_chunks(emub_model *m, const double *pts, int ldp, int mq, Compute compute, Collect collect)
{
	QueryWs *w = m->qws;
	cudaStream_t st = m->ctx->streams[0], cp = w->copy_stream;
	const int nchunks = (mq + w->mqc - 1) / w->mqc;
	auto stage = [&](int i) -> int {
		const int b = i & 1, done = i * w->mqc, cnt = std::min(w->mqc, mq - done);
		if (i >= 2) {
			CUDA_TRY(cudaEventSynchronize(w->evIn[b]));          // hQb[b] has left the host
			CUDA_TRY(cudaStreamWaitEvent(cp, w->evOut[b], 0));   // chunk i-2 no longer reads dQb[b]
		}
		if (ldp == m->d) memcpy(w->hQb[b], pts + (size_t)done * ldp, sizeof(double) * (size_t)cnt * m->d);
		else for (int q = 0; q < cnt; q++) memcpy(w->hQb[b] + (size_t)q * m->d, pts + (size_t)(done + q) * ldp, sizeof(double) * m->d);
		CUDA_TRY(cudaMemcpyAsync(w->dQb[b], w->hQb[b], sizeof(double) * (size_t)cnt * m->d, cudaMemcpyHostToDevice, cp));
		CUDA_TRY(cudaEventRecord(w->evIn[b], cp));
		return EMUB_OK;
	};
	if (nchunks > 0) { int rc = stage(0); if (rc) return rc; }
	for (int i = 0; i < nchunks; i++) {
		const int b = i & 1, cnt = std::min(w->mqc, mq - i * w->mqc);
		CUDA_TRY(cudaStreamWaitEvent(st, w->evIn[b], 0));
		int rc = compute(cnt, w->dQb[b], w->hOutb[b]);
		if (rc) { cudaStreamSynchronize(st); cudaStreamSynchronize(cp); return rc; }
		CUDA_TRY(cudaEventRecord(w->evOut[b], st));
		if (i + 1 < nchunks) { rc = stage(i + 1); if (rc) return rc; }
		if (i > 0) {
			CUDA_TRY(cudaEventSynchronize(w->evOut[b ^ 1]));
			collect((i - 1) * w->mqc, w->mqc, w->hOutb[b ^ 1]);
		}
	}
	if (nchunks > 0) {
		const int i = nchunks - 1;
		CUDA_TRY(cudaEventSynchronize(w->evOut[i & 1]));
		collect(i * w->mqc, mq - i * w->mqc, w->hOutb[i & 1]);
	}
	CUDA_TRY(cudaGetLastError());
	return EMUB_OK;
}

extern "C" int emub_predict_batch(emub_emulator *e, const double *pts, int ldp, int mq, double *mean, double *var)
{
	if (!e || !pts || !mean || !var || mq < 0 || ldp < e->m->d) return set_err(EMUB_EINVAL, "emub_predict_batch: bad argument%s");
	emub_model *m = e->m;
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	{ int rc0 = ensure_query_ws(m); if (rc0) return rc0; }
	QueryWs *w = m->qws;
	cudaStream_t st = m->ctx->streams[0];
	return walk_query_chunks(
	    m, pts, ldp, mq,
	    [&](int cnt, const double *dQ, double *hOut) -> int {
		    int rc = predict_chunk(e, st, dQ, cnt, w->dMean, w->dVar);
		    if (rc) return rc;
		    CUDA_TRY(cudaMemcpyAsync(hOut, w->dMean, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
		    CUDA_TRY(cudaMemcpyAsync(hOut + w->mqc, w->dVar, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
		    return EMUB_OK;
	    },
	    [&](int done, int cnt, const double *hOut) {
		    memcpy(mean + done, hOut, sizeof(double) * cnt);
		    memcpy(var + done, hOut + w->mqc, sizeof(double) * cnt);
	    });
}

// emulate_point for a handful of points (mq <= 8), the call pattern of an MCMC driver: no chunk walk, the skinny
// latency path of predict_chunk
// Runs `enqueue` (copies and kernels on st, nothing else) -- as a replayed CUDA graph when graphs are on: one launch per
// call instead of six to eight.  key must change whenever any pointer or size inside the sequence does.
template <class Enqueue>
static int few_run(emub_model *m, unsigned long long key, cudaStream_t st, Enqueue enqueue)
{
	emub_ctx *c = m->ctx;
	QueryWs *w = m->qws;
	if (c->profile || !c->use_graphs) {
		int rc = enqueue();
		if (rc) return rc;
		CUDA_TRY(cudaStreamSynchronize(st));
		return EMUB_OK;
	}
	if (key == 0) key = 1;
	int slot = -1;
	for (int i = 0; i < 8; i++)
		if (w->few_graphs[i].key == key && w->few_graphs[i].exec) slot = i;
	if (slot < 0) {
		slot = w->few_graph_next;
		w->few_graph_next = (w->few_graph_next + 1) % 8;
		if (w->few_graphs[slot].exec) { cudaGraphExecDestroy(w->few_graphs[slot].exec); w->few_graphs[slot].exec = nullptr; }
		w->few_graphs[slot].key = 0;
		const long long before = c->launches;
		cudaGraph_t graph = nullptr;
		CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
		int rc = enqueue();
		cudaError_t ce = cudaStreamEndCapture(st, &graph);
		if (rc != EMUB_OK || ce != cudaSuccess || !graph) {
			if (graph) cudaGraphDestroy(graph);
			cudaGetLastError();
			return rc != EMUB_OK ? rc : set_err(EMUB_ECUDA, "CUDA graph capture failed: %s", cudaGetErrorString(ce));
		}
		ce = cudaGraphInstantiate(&w->few_graphs[slot].exec, graph, 0);
		cudaGraphDestroy(graph);
		if (ce != cudaSuccess) { w->few_graphs[slot].exec = nullptr; return set_err(EMUB_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
		w->few_graphs[slot].key = key;
		w->few_graphs[slot].launches = c->launches - before;
		c->launches = before;
	}
	c->launches += w->few_graphs[slot].launches;
	CUDA_TRY(cudaGraphLaunch(w->few_graphs[slot].exec, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	return EMUB_OK;
}

static unsigned long long few_key(unsigned long long h, unsigned long long a, unsigned long long b, unsigned long long c2)
{
	const unsigned long long words[3] = {a, b, c2};
	for (unsigned long long v : words) h = (h ^ v) * 1099511628211ull;
	return h;
}

extern "C" int emub_predict_few(emub_emulator *e, const double *pts, int ldp, int mq, double *mean, double *var)
{
	if (!e || !pts || !mean || !var || mq < 0 || mq > FEW_MAX || ldp < e->m->d) return set_err(EMUB_EINVAL, "emub_predict_few: bad argument (at most 8 points)%s");
	if (mq == 0) return EMUB_OK;
	emub_model *m = e->m;
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	{ int rc0 = ensure_query_ws(m); if (rc0) return rc0; }
	QueryWs *w = m->qws;
	cudaStream_t st = m->ctx->streams[0];
	for (int q = 0; q < mq; q++) memcpy(w->hQ + (size_t)q * m->d, pts + (size_t)q * ldp, sizeof(double) * m->d);
	int rc = few_run(m, few_key(1469598103934665603ull, e->uid, (unsigned long long)mq, 0x53ull), st, [&]() -> int {
		CUDA_TRY(cudaMemcpyAsync(w->dQ, w->hQ, sizeof(double) * (size_t)mq * m->d, cudaMemcpyHostToDevice, st));
		int r = predict_chunk(e, st, w->dQ, mq, w->dMean, w->dVar, true);
		if (r) return r;
		// mean and variance sit mqc apart: one copy of the first FEW_MAX of each
		CUDA_TRY(cudaMemcpyAsync(w->hOut, w->dMean, sizeof(double) * mq, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaMemcpyAsync(w->hOut + FEW_MAX, w->dVar, sizeof(double) * mq, cudaMemcpyDeviceToHost, st));
		return EMUB_OK;
	});
	if (rc) return rc;
	memcpy(mean, w->hOut, sizeof(double) * mq);
	memcpy(var, w->hOut + FEW_MAX, sizeof(double) * mq);
	return EMUB_OK;
}

// emulate_point_multi (multivar_support.c:103-157) for mq points: every PCA component's (mean, var) for the whole
// chunk, then the back-projection to the nt observables on the device.  emus: nr emulators of ONE model.
// mean / var: mq x nt row-major.  With nt = 0 the PCA-space values are returned instead (emulate_point_multi_pca,
// multivar_support.c:78): mean / var are then mq x nr.
static int predict_multi_impl(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                              const double *training_mean, const double *evecs, const double *evals, double *mean, double *var,
                              bool few);

extern "C" int emub_predict_multi(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                                  const double *training_mean, const double *evecs, const double *evals, double *mean,
                                  double *var)
{
	return predict_multi_impl(emus, nr, pts, ldp, mq, nt, training_mean, evecs, evals, mean, var, false);
}

// the same for a handful of points (mq <= 8) on the latency path of every component (see emub_predict_few)
extern "C" int emub_predict_multi_few(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                                      const double *training_mean, const double *evecs, const double *evals, double *mean,
                                      double *var)
{
	if (mq > FEW_MAX) return set_err(EMUB_EINVAL, "emub_predict_multi_few: at most 8 points%s");
	return predict_multi_impl(emus, nr, pts, ldp, mq, nt, training_mean, evecs, evals, mean, var, true);
}

static int predict_multi_impl(emub_emulator *const *emus, int nr, const double *pts, int ldp, int mq, int nt,
                              const double *training_mean, const double *evecs, const double *evals, double *mean, double *var,
                              bool few)
{
	if (!emus || nr < 1 || !pts || !mean || !var || mq < 0) return set_err(EMUB_EINVAL, "emub_predict_multi: bad argument%s");
	emub_model *m = emus[0]->m;
	for (int j = 0; j < nr; j++)
		if (!emus[j] || emus[j]->m != m) return set_err(EMUB_EINVAL, "emub_predict_multi: emulators must share one model%s");
	if (nt > NTLIMIT || nr > NTLIMIT || (nt > 0 && (!training_mean || !evecs || !evals)) || ldp < m->d)
		return set_err(EMUB_EINVAL, "emub_predict_multi: bad projection arguments (at most 1024 observables / components)%s");
	CUDA_TRY(cudaSetDevice(m->ctx->device));
	{ int rc0 = ensure_query_ws(m); if (rc0) return rc0; }
	QueryWs *w = m->qws;
	if (nr > w->ncomp) return set_err(EMUB_EINVAL, "emub_predict_multi: more emulators than model components%s");
	{ int rc1 = ensure_output_capacity(m, std::max(nt, nr)); if (rc1) return rc1; }
	cudaStream_t st = m->ctx->streams[0];
	if (nt > 0) {
		// projection data: ybar[nt] | evecs[nt x nr] | evals[nr]; uploaded when it differs from what dProj holds
		std::vector<double> proj((size_t)nt + (size_t)nt * nr + nr);
		memcpy(proj.data(), training_mean, sizeof(double) * nt);
		memcpy(proj.data() + nt, evecs, sizeof(double) * (size_t)nt * nr);
		memcpy(proj.data() + nt + (size_t)nt * nr, evals, sizeof(double) * nr);
		unsigned long long h = 1469598103934665603ull ^ ((unsigned long long)nt << 32 | (unsigned)nr);
		for (double v : proj) {
			unsigned long long bits;
			memcpy(&bits, &v, sizeof(bits));
			h = (h ^ bits) * 1099511628211ull;
		}
		if (h == 0) h = 1;
		if (h != w->proj_hash) {
			w->proj_hash = 0;
			CUDA_TRY(upload(w->dProj, proj.data(), sizeof(double) * proj.size(), st));
			w->proj_hash = h;
		}
	}
	const size_t vofs = (size_t)w->ntcap * w->mqc;  // variances start here in a host output buffer
	const bool few_set = few && nr > 1 && mq <= FEW_MAX && few_set_fits(m, nr);
	auto compute = [&](int cnt, const double *dQ, double *hOut) -> int {
		if (few_set) {  // tables prepared by the caller of this lambda
			int rc = few_set_launch(emus, nr, st, dQ, cnt, w->dMean, w->dVar);
			if (rc) return rc;
		} else {
			for (int j = 0; j < nr; j++) {
				int rc = predict_chunk(emus[j], st, dQ, cnt, w->dMean + (size_t)j * w->mqc, w->dVar + (size_t)j * w->mqc, few);
				if (rc) return rc;
			}
		}
		if (nt > 0) {
			{
				LaunchScope ls(m->ctx, EMUB_K_PRED_FINAL, 0, st);
				k_backproject<<<(cnt + 127) / 128, 128, 0, st>>>(w->dMean, w->dVar, w->mqc, cnt, nt, nr, w->dProj, w->dProj + nt,
				                                                 w->dProj + nt + (size_t)nt * nr, w->dOutM, w->dOutV);
			}
			CUDA_TRY(cudaMemcpyAsync(hOut, w->dOutM, sizeof(double) * (size_t)cnt * nt, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaMemcpyAsync(hOut + vofs, w->dOutV, sizeof(double) * (size_t)cnt * nt, cudaMemcpyDeviceToHost, st));
		} else {
			CUDA_TRY(cudaMemcpy2DAsync(hOut, sizeof(double) * cnt, w->dMean, sizeof(double) * w->mqc, sizeof(double) * cnt, nr, cudaMemcpyDeviceToHost, st));
			CUDA_TRY(cudaMemcpy2DAsync(hOut + vofs, sizeof(double) * cnt, w->dVar, sizeof(double) * w->mqc, sizeof(double) * cnt, nr, cudaMemcpyDeviceToHost, st));
		}
		return EMUB_OK;
	};
	auto collect = [&](int done, int cnt, const double *hOut) {
		if (nt > 0) {
			memcpy(mean + (size_t)done * nt, hOut, sizeof(double) * (size_t)cnt * nt);
			memcpy(var + (size_t)done * nt, hOut + vofs, sizeof(double) * (size_t)cnt * nt);
		} else {
			for (int q = 0; q < cnt; q++)
				for (int j = 0; j < nr; j++) {
					mean[(size_t)(done + q) * nr + j] = hOut[(size_t)j * cnt + q];
					var[(size_t)(done + q) * nr + j] = hOut[vofs + (size_t)j * cnt + q];
				}
		}
	};
	if (few) {
		// a handful of points: no chunk walk; the whole sequence (copy in, kernels, copies out) is one replayed graph
		if (mq == 0) return EMUB_OK;
		unsigned long long hset = 1469598103934665603ull;
		if (few_set) {
			int rc = few_set_prepare(emus, nr, st, &hset);
			if (rc) return rc;
		} else
			for (int j = 0; j < nr; j++) hset = (hset ^ emus[j]->uid) * 1099511628211ull;
		for (int q = 0; q < mq; q++) memcpy(w->hQ + (size_t)q * m->d, pts + (size_t)q * ldp, sizeof(double) * m->d);
		const unsigned long long shape = (unsigned long long)mq | ((unsigned long long)nt << 8) | ((unsigned long long)nr << 24) |
		                                 ((unsigned long long)few_set << 40);
		int rc = few_run(m, few_key(hset, shape, nt > 0 ? w->proj_hash : 0, 0x4Dull), st, [&]() -> int {
			CUDA_TRY(cudaMemcpyAsync(w->dQ, w->hQ, sizeof(double) * (size_t)mq * m->d, cudaMemcpyHostToDevice, st));
			return compute(mq, w->dQ, w->hOut);
		});
		if (rc) return rc;
		collect(0, mq, w->hOut);
		return EMUB_OK;
	}
	return walk_query_chunks(m, pts, ldp, mq, compute, collect);
}
