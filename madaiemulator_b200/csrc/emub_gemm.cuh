// emub_gemm.cuh -- FP64 tensor-core (DMMA, mma.sync.m16n8k4.f64 -> SASS DMMA.8x8x4) tile engine.
//
// One generic batched kernel, driven by a task list: every CTA computes one 128 x 128 output tile
//     acc(m, n) = sum_k A(m, k) * B(n, k)
// over a task-specific K range, then applies an epilogue.  Triangular structure (Cholesky trailing
// updates, triangular inverse, W^T W, W K) is expressed purely through the K range and the tile
// list, built once per model on the host (emub_schedule.cpp).  tcgen05/TMEM has no FP64 kind, so
// on sm_100a FP64 tensor work goes through mma.sync; operands are staged global -> shared with a
// multi-stage cp.async pipeline, padded so that every fragment load is bank-conflict free.
//
// Operand layouts (per operand):
//   KMAJOR : elem(r, k) = P[r * ld + k]   (k contiguous)   smem tile [rows][BK + 4]
//   RMAJOR : elem(r, k) = P[k * ld + r]   (r contiguous)   smem tile [BK][rows + 4]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emub {

constexpr int TB = 128;       // tile edge (BM = BN = NB)

// Tile-engine configuration: the task list addresses 128 x 128 output tiles; a CTA computes one BM x BN
// sub-tile of it with WM x WN warps (warp tile (BM/WM) x (BN/WN)), BK k-columns per pipeline stage and
// STAGES cp.async stages; MINB CTAs are resident per SM so that the prologue, barriers and epilogue of one
// CTA overlap the DMMA stream of the others (cuBLAS reaches 95% tensor-pipe activity the same way with
// 64 x 64 tiles).  Shared-memory strides are padded to 4 (mod 16) doubles so that the 16 lanes of a half
// warp (4 rows x 4 k) hit 16 different 8-byte banks.
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_, int MINB_>
struct GemmCfg {
	static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_, MINB = MINB_;
	static constexpr int THREADS = WM * WN * 32;
	static constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
	static constexpr int MI = WTM / 16, NI = WTN / 8;   // m16n8k4 tiles per warp
	static constexpr int SUBM = TB / BM, SUBN = TB / BN, SUBS = SUBM * SUBN;
	static constexpr int KM_STRIDE = BK + 4;
	static constexpr int RMA_STRIDE = BM + 4, RMB_STRIDE = BN + 4;
	static constexpr int A_DOUBLES = (BM * KM_STRIDE > BK * RMA_STRIDE) ? BM * KM_STRIDE : BK * RMA_STRIDE;
	static constexpr int B_DOUBLES = (BN * KM_STRIDE > BK * RMB_STRIDE) ? BN * KM_STRIDE : BK * RMB_STRIDE;
	static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
	static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
};
using DefaultCfg = GemmCfg<64, 64, 2, 2, 8, 4, 4>;  // 4 CTAs / SM: 35.3 TFLOP/s on 4096^3 (cuBLAS: 35.2)
// Launches that would not even half fill the GPU with 64 x 64 sub-tiles (the deep levels of the recursion when few
// matrices are in flight) use 32 x 32 sub-tiles instead: 4x the CTAs, each with a quarter of the DMMAs per k-step.
// Every output element still sees the same k-ordered sequence of DMMAs, so results do not depend on the choice (the
// K-range trimming of the default configuration only skips products with structural zeros).
using SmallCfg = GemmCfg<32, 32, 2, 2, 8, 4, 8>;
constexpr int GEMM_THREADS = DefaultCfg::THREADS;
constexpr int GEMM_SMEM_BYTES = DefaultCfg::SMEM_BYTES;

enum { KMAJOR = 0, RMAJOR = 1 };
enum { EPI_STORE = 0, EPI_SUB = 1, EPI_COLSUMSQ = 2 };

struct GemmTask {
	long long a_off;  // element offsets inside one slot's matrices
	long long b_off;
	long long c_off;
	int klen;         // multiple of BK
	int aux;          // epilogue specific (COLSUMSQ: row-block index)
};

struct GemmArgs {
	const GemmTask *tasks;
	const double *A;
	const double *B;
	double *C;
	long long strideA, strideB, strideC;  // per batch element (blockIdx.y)
	int lda, ldb, ldc;
	double alpha;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma_16x8x4(double (&c)[4], double a0, double a1, double b0)
{
	asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
	             : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
	             : "d"(a0), "d"(a1), "d"(b0));
}

// ROWS x BK operand tile, global -> shared, 16-byte cp.async chunks
template <class Cfg, int LAYOUT, int ROWS>
__device__ __forceinline__ void load_operand_stage(double *s, const double *g, int ld, int tid)
{
	constexpr int CHUNKS = ROWS * Cfg::BK / 2;
	constexpr int ITERS = (CHUNKS + Cfg::THREADS - 1) / Cfg::THREADS;
	if (LAYOUT == KMAJOR) {
		constexpr int CPR = Cfg::BK / 2;  // chunks per row
#pragma unroll
		for (int c = 0; c < ITERS; c++) {
			int chunk = tid + c * Cfg::THREADS;
			if (CHUNKS % Cfg::THREADS == 0 || chunk < CHUNKS) {
				int row = chunk / CPR, kc = chunk % CPR;
				cp_async16(s + row * Cfg::KM_STRIDE + kc * 2, g + (long long)row * ld + kc * 2);
			}
		}
	} else {
		constexpr int CPK = ROWS / 2;  // chunks per k row
		constexpr int STRIDE = ROWS + 4;
#pragma unroll
		for (int c = 0; c < ITERS; c++) {
			int chunk = tid + c * Cfg::THREADS;
			if (CHUNKS % Cfg::THREADS == 0 || chunk < CHUNKS) {
				int kr = chunk / CPK, rc = chunk % CPK;
				cp_async16(s + kr * STRIDE + rc * 2, g + (long long)kr * ld + rc * 2);
			}
		}
	}
}

template <class Cfg, int LAYOUT, int ROWS>
__device__ __forceinline__ double frag(const double *s, int r, int k)
{
	return (LAYOUT == KMAJOR) ? s[r * Cfg::KM_STRIDE + k] : s[k * (ROWS + 4) + r];
}

// acc[mi][ni][4]: warp tile WTM (m) x WTN (n)
template <class Cfg, int AL, int BL>
// [w_lo, w_hi): k-tiles for which THIS warp issues DMMAs (warp-uniform).  Outside it the warp still takes part in
// the loads and barriers; its operands there are structural zeros of a triangular block.
__device__ __forceinline__ void gemm_mainloop(const double *__restrict__ gA, const double *__restrict__ gB, int lda,
                                              int ldb, int klen, int w_lo, int w_hi, double *smem,
                                              double (&acc)[Cfg::MI][Cfg::NI][4])
{
	constexpr int BK = Cfg::BK, STAGES = Cfg::STAGES, STG = Cfg::STAGE_DOUBLES;
	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int wm = (warp / Cfg::WN) * Cfg::WTM, wn = (warp % Cfg::WN) * Cfg::WTN;
	const int ktiles = klen / BK;
	const long long a_step = (AL == KMAJOR) ? BK : (long long)BK * lda;
	const long long b_step = (BL == KMAJOR) ? BK : (long long)BK * ldb;

#pragma unroll
	for (int s = 0; s < STAGES - 1; s++) {
		if (s < ktiles) {
			load_operand_stage<Cfg, AL, Cfg::BM>(smem + s * STG, gA + s * a_step, lda, tid);
			load_operand_stage<Cfg, BL, Cfg::BN>(smem + s * STG + Cfg::A_DOUBLES, gB + s * b_step, ldb, tid);
		}
		cp_async_commit();
	}
	for (int kt = 0; kt < ktiles; kt++) {
		cp_async_wait<STAGES - 2>();
		__syncthreads();
		{
			int nk = kt + STAGES - 1;
			if (nk < ktiles) {
				int slot = nk % STAGES;
				load_operand_stage<Cfg, AL, Cfg::BM>(smem + slot * STG, gA + nk * a_step, lda, tid);
				load_operand_stage<Cfg, BL, Cfg::BN>(smem + slot * STG + Cfg::A_DOUBLES, gB + nk * b_step, ldb, tid);
			}
			cp_async_commit();
		}
		if (kt < w_lo || kt >= w_hi) continue;
		const double *sA = smem + (kt % STAGES) * STG;
		const double *sB = sA + Cfg::A_DOUBLES;
#pragma unroll
		for (int kk = 0; kk < BK; kk += 4) {
			double a[Cfg::MI][2], b[Cfg::NI];
#pragma unroll
			for (int mi = 0; mi < Cfg::MI; mi++) {
				a[mi][0] = frag<Cfg, AL, Cfg::BM>(sA, wm + mi * 16 + g, kk + t);
				a[mi][1] = frag<Cfg, AL, Cfg::BM>(sA, wm + mi * 16 + g + 8, kk + t);
			}
#pragma unroll
			for (int ni = 0; ni < Cfg::NI; ni++) b[ni] = frag<Cfg, BL, Cfg::BN>(sB, wn + ni * 8 + g, kk + t);
#pragma unroll
			for (int mi = 0; mi < Cfg::MI; mi++)
#pragma unroll
				for (int ni = 0; ni < Cfg::NI; ni++) dmma_16x8x4(acc[mi][ni], a[mi][0], a[mi][1], b[ni]);
		}
	}
	cp_async_wait<0>();
}

// grid: (ntasks * SUBS, batch).  TASK_LOWER in task.aux (EPI_STORE / EPI_SUB): the tile is a diagonal tile of a
// symmetric / triangular result, sub-tiles strictly above the diagonal are skipped.
constexpr int TASK_LOWER = 1 << 30;
// K-range trimming at sub-tile granularity for products against a triangular diagonal block (valid for 64 x 64
// sub-tiles): the first / last 64 k of the range multiply structural zeros for one half of the sub-tiles.
constexpr int TASK_TRIM_END_SC0 = 1 << 29;    // B's last k-block is lower triangular in (n, k): columns n < 64 need k < 64
constexpr int TASK_TRIM_BEGIN_SC1 = 1 << 28;  // B's first k-block is upper triangular in (n, k): columns n >= 64 need k >= 64
constexpr int TASK_TRIM_END_SR0 = 1 << 27;    // same for A and the rows m
constexpr int TASK_TRIM_BEGIN_SR1 = 1 << 26;
constexpr int TASK_FLAGS = TASK_LOWER | TASK_TRIM_END_SC0 | TASK_TRIM_BEGIN_SC1 | TASK_TRIM_END_SR0 | TASK_TRIM_BEGIN_SR1;

// executed flops of one task under the default 64 x 64 sub-tiling (host side bookkeeping)
inline double task_flops(const GemmTask &t, bool colsumsq = false)
{
	double f = 0.0;
	for (int sr = 0; sr < 2; sr++)
		for (int sc = 0; sc < 2; sc++) {
			if (!colsumsq && (t.aux & TASK_LOWER) && sc > sr) continue;
			int k = t.klen;
			if ((t.aux & TASK_TRIM_END_SC0) && sc == 0) k -= 64;
			if ((t.aux & TASK_TRIM_BEGIN_SC1) && sc == 1) k -= 64;
			if ((t.aux & TASK_TRIM_END_SR0) && sr == 0) k -= 64;
			if ((t.aux & TASK_TRIM_BEGIN_SR1) && sr == 1) k -= 64;
			for (int wr = 0; wr < 2; wr++)
				for (int wc = 0; wc < 2; wc++) {
					int kw = k;
					if (((t.aux & TASK_TRIM_END_SC0) && wc == 0) || ((t.aux & TASK_TRIM_END_SR0) && wr == 0)) kw -= 32;
					if (((t.aux & TASK_TRIM_BEGIN_SC1) && wc == 1) || ((t.aux & TASK_TRIM_BEGIN_SR1) && wr == 1)) kw -= 32;
					if (!colsumsq && (t.aux & TASK_LOWER) && sr == sc && wr == 0 && wc == 1) kw = 0;
					f += 2.0 * 32 * 32 * (double)(kw > 0 ? kw : 0);
				}
		}
	return f;
}

template <int AL, int BL, int EPI, class Cfg = DefaultCfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) k_gemm(GemmArgs p)
{
	constexpr int MI = Cfg::MI, NI = Cfg::NI;
	extern __shared__ __align__(16) double smem[];
	const GemmTask task = p.tasks[blockIdx.x / Cfg::SUBS];
	const int sub = blockIdx.x % Cfg::SUBS;
	const int sr = sub / Cfg::SUBN, sc = sub % Cfg::SUBN;
	if (EPI != EPI_COLSUMSQ && (task.aux & TASK_LOWER) && sc * Cfg::BN >= (sr + 1) * Cfg::BM) return;
	const int b = blockIdx.y;
	const double *gA = p.A + b * p.strideA + task.a_off + ((AL == KMAJOR) ? (long long)sr * Cfg::BM * p.lda : (long long)sr * Cfg::BM);
	const double *gB = p.B + b * p.strideB + task.b_off + ((BL == KMAJOR) ? (long long)sc * Cfg::BN * p.ldb : (long long)sc * Cfg::BN);
	int klen = task.klen;
	if (Cfg::BM == 64 && Cfg::BN == 64) {
		const bool trim_end = ((task.aux & TASK_TRIM_END_SC0) && sc == 0) || ((task.aux & TASK_TRIM_END_SR0) && sr == 0);
		const bool trim_begin = ((task.aux & TASK_TRIM_BEGIN_SC1) && sc == 1) || ((task.aux & TASK_TRIM_BEGIN_SR1) && sr == 1);
		if (trim_end) klen -= 64;
		if (trim_begin) {
			klen -= 64;
			gA += (AL == KMAJOR) ? 64 : 64 * (long long)p.lda;
			gB += (BL == KMAJOR) ? 64 : 64 * (long long)p.ldb;
		}
	}

	double acc[MI][NI][4];
#pragma unroll
	for (int i = 0; i < MI; i++)
#pragma unroll
		for (int j = 0; j < NI; j++)
#pragma unroll
			for (int r = 0; r < 4; r++) acc[i][j][r] = 0.0;

	// warp-level window: a 32-row / 32-column warp tile needs even less of a triangular diagonal block
	int w_lo = 0, w_hi = klen / Cfg::BK;
	if (Cfg::BM == 64 && Cfg::BN == 64 && Cfg::WTM == 32 && Cfg::WTN == 32) {
		const int w = threadIdx.x >> 5;
		const int wr = w / Cfg::WN, wc = w % Cfg::WN;
		constexpr int T32 = 32 / Cfg::BK;
		if (((task.aux & TASK_TRIM_END_SC0) && wc == 0) || ((task.aux & TASK_TRIM_END_SR0) && wr == 0)) w_hi -= T32;
		if (((task.aux & TASK_TRIM_BEGIN_SC1) && wc == 1) || ((task.aux & TASK_TRIM_BEGIN_SR1) && wr == 1)) w_lo += T32;
		if (EPI != EPI_COLSUMSQ && (task.aux & TASK_LOWER) && sr == sc && wr == 0 && wc == 1) w_hi = 0;  // strictly upper warp tile
	}
	gemm_mainloop<Cfg, AL, BL>(gA, gB, p.lda, p.ldb, klen, w_lo, w_hi, smem, acc);

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int wmi = warp / Cfg::WN;
	const int wm = wmi * Cfg::WTM, wn = (warp % Cfg::WN) * Cfg::WTN;

	if (EPI == EPI_STORE || EPI == EPI_SUB) {
		double *gC = p.C + b * p.strideC + task.c_off + (long long)sr * Cfg::BM * p.ldc + sc * Cfg::BN;
#pragma unroll
		for (int mi = 0; mi < MI; mi++)
#pragma unroll
			for (int ni = 0; ni < NI; ni++)
#pragma unroll
				for (int h = 0; h < 2; h++) {
					int row = wm + mi * 16 + g + 8 * h;
					int col = wn + ni * 8 + 2 * t;
					double2 *ptr = reinterpret_cast<double2 *>(gC + (long long)row * p.ldc + col);
					double2 v;
					if (EPI == EPI_SUB) {
						v = *ptr;
						v.x -= acc[mi][ni][2 * h];
						v.y -= acc[mi][ni][2 * h + 1];
					} else {
						v.x = p.alpha * acc[mi][ni][2 * h];
						v.y = p.alpha * acc[mi][ni][2 * h + 1];
					}
					*ptr = v;
				}
	} else {
		// column sums of squares of the BM x BN sub-tile -> C[(aux * SUBM + sr) * ldc + c_off + sc * BN + col]
		__syncthreads();  // everyone is done with the pipeline buffers
		double *red = smem;  // [WM][BN]
#pragma unroll
		for (int ni = 0; ni < NI; ni++) {
			double s0 = 0.0, s1 = 0.0;
#pragma unroll
			for (int mi = 0; mi < MI; mi++) {
				s0 += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][2] * acc[mi][ni][2];
				s1 += acc[mi][ni][1] * acc[mi][ni][1] + acc[mi][ni][3] * acc[mi][ni][3];
			}
#pragma unroll
			for (int o = 4; o < 32; o <<= 1) {
				s0 += __shfl_xor_sync(0xffffffffu, s0, o);
				s1 += __shfl_xor_sync(0xffffffffu, s1, o);
			}
			if (g == 0) {
				red[wmi * Cfg::BN + wn + ni * 8 + 2 * t] = s0;
				red[wmi * Cfg::BN + wn + ni * 8 + 2 * t + 1] = s1;
			}
		}
		__syncthreads();
		if (tid < Cfg::BN) {
			double *gC = p.C + b * p.strideC + (long long)((task.aux & 0xffff) * Cfg::SUBM + sr) * p.ldc + task.c_off + sc * Cfg::BN;
			double sum = 0.0;
#pragma unroll
			for (int w = 0; w < Cfg::WM; w++) sum += red[w * Cfg::BN + tid];
			gC[tid] = sum;
		}
	}
}

}  // namespace emub
