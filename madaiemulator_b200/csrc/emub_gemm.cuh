// emub_gemm.cuh -- FP64 tensor-core (DMMA, mma.sync.m16n8k4.f64 -> SASS DMMA.8x8x4) tile engine.
//
// One generic batched kernel, driven by a task list: every CTA computes one 128 x 128 output tile
//     acc(m, n) = sum_k A(m, k) * B(n, k)
// over a task-specific K range, then applies an epilogue.  Triangular structure (Cholesky trailing
// updates, triangular inverse, W^T W, W K) is expressed purely through the K range and the tile
// list, built once per model on the host (emub_schedule.cpp).  tcgen05/TMEM has no FP64 kind, so
// on sm_100a FP64 tensor work goes through mma.sync; operands are staged global -> shared with a
// 4-stage cp.async pipeline, padded so that every fragment load is bank-conflict free.
//
// Operand layouts (per operand):
//   KMAJOR : elem(r, k) = P[r * ld + k]   (k contiguous)   smem tile [128][BK + 4]
//   RMAJOR : elem(r, k) = P[k * ld + r]   (r contiguous)   smem tile [BK][128 + 4]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emub {

constexpr int TB = 128;       // tile edge (BM = BN = NB)
constexpr int BK = 16;        // k per pipeline stage
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int KM_STRIDE = BK + 4;   // 20 doubles: (g*20 + t) mod 16 distinct over a half warp
constexpr int RM_STRIDE = TB + 4;   // 132 doubles: (t*132 + g) mod 16 distinct over a half warp
constexpr int OPER_DOUBLES = TB * KM_STRIDE;  // 2560 >= BK * RM_STRIDE (2112)
constexpr int GEMM_SMEM_BYTES = STAGES * 2 * OPER_DOUBLES * 8;  // 163840

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

template <int LAYOUT>
__device__ __forceinline__ void load_operand_stage(double *s, const double *g, int ld, int tid)
{
	if (LAYOUT == KMAJOR) {
		// 128 rows x 16 k: 8 chunks (16 B) per row
#pragma unroll
		for (int c = 0; c < 4; c++) {
			int chunk = tid + c * GEMM_THREADS;
			int row = chunk >> 3, kc = chunk & 7;
			cp_async16(s + row * KM_STRIDE + kc * 2, g + (long long)row * ld + kc * 2);
		}
	} else {
		// 16 k rows x 128 r: 64 chunks per k row
#pragma unroll
		for (int c = 0; c < 4; c++) {
			int chunk = tid + c * GEMM_THREADS;
			int kr = chunk >> 6, rc = chunk & 63;
			cp_async16(s + kr * RM_STRIDE + rc * 2, g + (long long)kr * ld + rc * 2);
		}
	}
}

template <int LAYOUT>
__device__ __forceinline__ double frag(const double *s, int r, int k)
{
	return (LAYOUT == KMAJOR) ? s[r * KM_STRIDE + k] : s[k * RM_STRIDE + r];
}

// acc[mi][ni][4]: warp tile 64 (m) x 32 (n); 8 warps as 2 (m) x 4 (n)
template <int AL, int BL>
__device__ __forceinline__ void gemm_mainloop(const double *__restrict__ gA, const double *__restrict__ gB, int lda,
                                              int ldb, int klen, double *smem, double (&acc)[4][4][4])
{
	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
	const int ktiles = klen / BK;
	const long long a_step = (AL == KMAJOR) ? BK : (long long)BK * lda;
	const long long b_step = (BL == KMAJOR) ? BK : (long long)BK * ldb;

#pragma unroll
	for (int s = 0; s < STAGES - 1; s++) {
		if (s < ktiles) {
			load_operand_stage<AL>(smem + (2 * s) * OPER_DOUBLES, gA + s * a_step, lda, tid);
			load_operand_stage<BL>(smem + (2 * s + 1) * OPER_DOUBLES, gB + s * b_step, ldb, tid);
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
				load_operand_stage<AL>(smem + (2 * slot) * OPER_DOUBLES, gA + nk * a_step, lda, tid);
				load_operand_stage<BL>(smem + (2 * slot + 1) * OPER_DOUBLES, gB + nk * b_step, ldb, tid);
			}
			cp_async_commit();
		}
		const double *sA = smem + (2 * (kt % STAGES)) * OPER_DOUBLES;
		const double *sB = sA + OPER_DOUBLES;
#pragma unroll
		for (int kk = 0; kk < BK; kk += 4) {
			double a[4][2], b[4];
#pragma unroll
			for (int mi = 0; mi < 4; mi++) {
				a[mi][0] = frag<AL>(sA, wm + mi * 16 + g, kk + t);
				a[mi][1] = frag<AL>(sA, wm + mi * 16 + g + 8, kk + t);
			}
#pragma unroll
			for (int ni = 0; ni < 4; ni++) b[ni] = frag<BL>(sB, wn + ni * 8 + g, kk + t);
#pragma unroll
			for (int mi = 0; mi < 4; mi++)
#pragma unroll
				for (int ni = 0; ni < 4; ni++) dmma_16x8x4(acc[mi][ni], a[mi][0], a[mi][1], b[ni]);
		}
	}
	cp_async_wait<0>();
}

// grid: (ntasks, batch)
template <int AL, int BL, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_gemm(GemmArgs p)
{
	extern __shared__ __align__(16) double smem[];
	const GemmTask task = p.tasks[blockIdx.x];
	const int b = blockIdx.y;
	const double *gA = p.A + b * p.strideA + task.a_off;
	const double *gB = p.B + b * p.strideB + task.b_off;

	double acc[4][4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++)
#pragma unroll
			for (int r = 0; r < 4; r++) acc[i][j][r] = 0.0;

	gemm_mainloop<AL, BL>(gA, gB, p.lda, p.ldb, task.klen, smem, acc);

	const int tid = threadIdx.x;
	const int warp = tid >> 5, lane = tid & 31;
	const int g = lane >> 2, t = lane & 3;
	const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;

	if (EPI == EPI_STORE || EPI == EPI_SUB) {
		double *gC = p.C + b * p.strideC + task.c_off;
#pragma unroll
		for (int mi = 0; mi < 4; mi++)
#pragma unroll
			for (int ni = 0; ni < 4; ni++)
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
		// column sums of squares of the 128 x 128 tile -> C[aux * ldc + c_off + col]
		__syncthreads();  // everyone is done with the pipeline buffers
		double *red = smem;  // [2][128]
#pragma unroll
		for (int ni = 0; ni < 4; ni++) {
			double s0 = 0.0, s1 = 0.0;
#pragma unroll
			for (int mi = 0; mi < 4; mi++) {
				s0 += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][2] * acc[mi][ni][2];
				s1 += acc[mi][ni][1] * acc[mi][ni][1] + acc[mi][ni][3] * acc[mi][ni][3];
			}
#pragma unroll
			for (int o = 4; o < 32; o <<= 1) {
				s0 += __shfl_xor_sync(0xffffffffu, s0, o);
				s1 += __shfl_xor_sync(0xffffffffu, s1, o);
			}
			if (g == 0) {
				red[(warp >> 2) * TB + wn + ni * 8 + 2 * t] = s0;
				red[(warp >> 2) * TB + wn + ni * 8 + 2 * t + 1] = s1;
			}
		}
		__syncthreads();
		if (tid < TB) {
			double *gC = p.C + b * p.strideC + (long long)task.aux * p.ldc + task.c_off;
			gC[tid] = red[tid] + red[TB + tid];
		}
	}
}

}  // namespace emub
