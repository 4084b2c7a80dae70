// emub_kernels.cuh -- the non-GEMM kernels of the engine: theta preparation, covariance
// construction (K1), regression basis, diagonal-block factorisation (POTF2 + in-place triangular
// inverse), skinny products against W = L^-1, the small p x p regression algebra (K6), the fused
// likelihood-gradient reduction (K4) and the prediction epilogue.
// Reference formulas are cited as file:line under the reference's src/.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "emub_gemm.cuh"
#include "emub_exp.cuh"
#include "emub_potf2.cuh"

namespace emub {

constexpr int MAXD = 64;     // nparams limit (smem staging of design rows: 2 * d * 64 doubles per covariance / gradient CTA)
constexpr int MAXNCP = 104;  // 1 + nregression_fns, padded to a multiple of 8 (cubic regression up to d = 34, linear up to d = 64)
constexpr int CT = 64;       // covariance / gradient tile edge

// few-points prediction path (k_few_*, k_pred_final): one entry per emulator (PCA component) when several are answered
// by the same launches (component = blockIdx.z / .y)
struct FewSet {
	const double *W, *AB, *beta, *Minv;
	double kappa;
};

// per-point constants derived from theta, layout (stride CONST_STRIDE doubles):
//  [0] amp  [1] nugget  [2] rho (Matern)  [3] sigma2 slot (filled later)
//  [4 .. 4+d)       1 / (exp(theta_k))^2               emulator.c:123-127
//  [4+d .. 4+2d)    0.5 * exp(-2 theta_k)              emulator.c:181,203
//  [4+2d .. 4+3d)   exp(-2 theta_k)
//  [4+3d .. 4+4d)   sqrt(0.5 / l_k^2 * 64/ln2)         covariance: coordinates are staged times this, so that the exponent
//  [4+4d .. 4+5d)   sqrt(0.5 exp(-2 theta_k) * 64/ln2)  ... gradient                 is -sum_k (x'_ik - x'_jk)^2 in exp_scaled's units
constexpr int CONST_STRIDE = 4 + 5 * MAXD;

enum { THETA_LIK = 0 /* theta without amplitude, unit amplitude (maxmultimin.c:311-313; Matern: D-2) */,
       THETA_FULL = 1 /* literal full vector (emulator_struct.c:28) */ };

__global__ void k_theta_prep(const double *__restrict__ thetas, int B, int nth_in, int kernel, int d, int mode,
                             double *__restrict__ consts)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	const double *th = thetas + (size_t)b * nth_in;
	double *c = consts + (size_t)b * CONST_STRIDE;
	if (kernel == 1) {
		const double *len;
		if (mode == THETA_LIK) { c[0] = 1.0; c[1] = exp(th[0]); len = th + 1; }
		else { c[0] = exp(th[0]); c[1] = exp(th[1]); len = th + 2; }
		c[2] = 0.0;
		for (int k = 0; k < d; k++) {
			double r = exp(len[k]);
			r = r * r;
			c[4 + k] = 1.0 / r;
			double e2 = exp(-2.0 * len[k]);
			c[4 + d + k] = 0.5 * e2;
			c[4 + 2 * d + k] = e2;
			c[4 + 3 * d + k] = sqrt((0.5 * c[4 + k]) * EXP_SCALE);
			c[4 + 4 * d + k] = sqrt(c[4 + d + k] * EXP_SCALE);
		}
	} else {
		if (mode == THETA_LIK) { c[0] = 1.0; c[1] = exp(th[0]); c[2] = exp(th[1]); }
		else { c[0] = th[0]; c[1] = th[1]; c[2] = exp(th[2]); }
	}
	c[3] = 0.0;
}

// K1.  grid (ceil(ncols/64), ceil(nrows/64), B), 256 threads, dynamic smem 2*d*64 doubles.
// Square mode (CROSS=false): rows and columns are design points; entries outside n x n are the identity
// (padding); lower_only skips tiles strictly above the diagonal.
// Cross mode: columns are query points (mq valid), clamp < 1e-10 -> 0 (emulator.c:588-590), padding is 0.
// Each thread owns a 4 x 4 register tile of pairs; per parameter k it loads 4 + 4 coordinates once and
// updates the 16 exponents (3 FP64 ops per pair and parameter), so the kernel is FP64-pipe bound, not
// shared-memory bound.  "All |dist_k| below the threshold" (the nugget condition, emulator.c:136-150) is
// tested on k = 0 for every pair and on k >= 1 only for the pairs that are still candidates.
template <int KERNEL, bool CROSS>
__global__ void __launch_bounds__(256) k_cov(const double *__restrict__ X, int n, int d, const double *__restrict__ Q,
                                             int mq, const double *__restrict__ consts, long long const_stride,
                                             double *__restrict__ out, long long out_stride, int ld, int lower_only)
{
	const int bj = blockIdx.x, bi = blockIdx.y, b = blockIdx.z;
	if (!CROSS && lower_only && bj > bi) return;
	extern __shared__ double sm[];
	double *sXi = sm;            // [d][64]
	double *sXj = sm + d * CT;   // [d][64]
	__shared__ double sc[CONST_STRIDE];
	__shared__ double sh[MAXD];  // coincidence threshold per parameter, in staged units
	__shared__ double stab[64 * EXP_REP];
	exp_table_load<EXP_REP>(stab);
	const int tid = threadIdx.x;
	const double *cg = consts + b * const_stride;
	for (int i = tid; i < 4 + 3 * d; i += 256) sc[i] = cg[i];
	// Power-exponential kernel: the coordinates are staged already multiplied by s_k = sqrt(0.5 / l_k^2 * 64/ln2), so the
	// exponent (-1/2 dist) dist / l_k^2 (emulator.c:132, D-5) summed over k is -sum_k (x'_ik - x'_jk)^2 in exp_scaled's
	// units: one subtraction and one fused multiply-add per pair and parameter.  sh[k] = 1e-10 * s_k is the
	// coincidence threshold (emulator.c:136-150) in the same units.
	if (KERNEL == 1) for (int i = tid; i < d; i += 256) sh[i] = 0.0000000001 * cg[4 + 3 * d + i];
	const int i0 = bi * CT, j0 = bj * CT;
	const int ncols = CROSS ? mq : n;
	const double *XJ = CROSS ? Q : X;
	for (int idx = tid; idx < CT * d; idx += 256) {
		int r = idx / d, k = idx - r * d;
		const double sk = (KERNEL == 1) ? cg[4 + 3 * d + k] : 1.0;
		sXi[k * CT + r] = (i0 + r < n) ? X[(size_t)(i0 + r) * d + k] * sk : 0.0;
		sXj[k * CT + r] = (j0 + r < ncols) ? XJ[(size_t)(j0 + r) * d + k] * sk : 0.0;
	}
	__syncthreads();
	const int tx = tid & 15, ty = tid >> 4;
	const double thr0 = (KERNEL == 1) ? sh[0] : 0.0000000000000001;
	double e[4][4];
	unsigned same = 0;
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < 4; c++) e[r][c] = 0.0;
	{
		double xi[4], xj[4];
#pragma unroll
		for (int r = 0; r < 4; r++) xi[r] = sXi[ty + 16 * r];
#pragma unroll
		for (int c = 0; c < 4; c++) xj[c] = sXj[tx + 16 * c];
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) {
				const double dist = fabs(xi[r] - xj[c]);
				e[r][c] = dist * dist;
				if (dist < thr0) same |= 1u << (r * 4 + c);
			}
	}
	for (int k = 1; k < d; k++) {
		double xi[4], xj[4];
#pragma unroll
		for (int r = 0; r < 4; r++) xi[r] = sXi[k * CT + ty + 16 * r];
#pragma unroll
		for (int c = 0; c < 4; c++) xj[c] = sXj[k * CT + tx + 16 * c];
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) {
				const double dl = xi[r] - xj[c];
				e[r][c] = fma(dl, dl, e[r][c]);
			}
		if (same) {  // rare: only pairs that coincide in every parameter seen so far
			const double thr = (KERNEL == 1) ? sh[k] : 0.0000000000000001;
#pragma unroll
			for (int r = 0; r < 4; r++)
#pragma unroll
				for (int c = 0; c < 4; c++)
					if (!(fabs(xi[r] - xj[c]) < thr)) same &= ~(1u << (r * 4 + c));
		}
	}
	double *o = out + b * out_stride;
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const int gi = i0 + ty + 16 * r;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const int gj = j0 + tx + 16 * c;
			double v;
			if (gi < n && gj < ncols) {
				if (KERNEL == 1) {
					v = exp_scaled<EXP_REP>(-e[r][c], stab) * sc[0];  // emulator.c:134
				} else {
					// emulator.c:344-386 (Matern 3/2), :438-480 (Matern 5/2); e holds the squared distance
					const double dist = sqrt(e[r][c]);
					if (KERNEL == 2) {
						const double root3 = 1.732050808;
						v = (dist > 0.0) ? sc[0] * (1 + root3 * (dist / sc[2])) * exp_neg<EXP_REP>(-root3 * (dist / sc[2]), stab) : sc[0];
					} else {
						const double root5 = 2.236067978;
						const double dr = dist / sc[2];
						v = (dist > 0.0) ? sc[0] * (1 + root5 * dr + (5.0 / 3.0) * dr * dr) * exp_neg<EXP_REP>(-root5 * dr, stab) : sc[0];
					}
				}
				if (same & (1u << (r * 4 + c))) v += sc[1];
				if (CROSS && v < 1E-10) v = 0.0;
			} else {
				v = (!CROSS && gi == gj) ? 1.0 : 0.0;
			}
			o[(size_t)gi * ld + gj] = v;
		}
	}
}

// Yh = [ y | H | 0 ] (npad x ncp), H per regression.c:9-67; rows >= n are zero.
// y is read with stride ystride (component c of an n x ncomp training matrix: y + c, stride ncomp)
__global__ void k_build_yh(const double *__restrict__ X, const double *__restrict__ y, int ystride, int n, int npad, int d,
                           int order, int ncp, double *__restrict__ Yh)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= npad) return;
	double *row = Yh + (size_t)i * ncp;
	for (int c = 0; c < ncp; c++) row[c] = 0.0;
	if (i >= n) return;
	row[0] = y[(size_t)i * ystride];
	row[1] = 1.0;
	const double *x = X + (size_t)i * d;
	if (order >= 1) for (int k = 0; k < d; k++) row[2 + k] = x[k];
	if (order >= 2) for (int k = 0; k < d; k++) row[2 + d + k] = x[k] * x[k];
	if (order >= 3) for (int k = 0; k < d; k++) row[2 + 2 * d + k] = x[k] * x[k] * x[k];
}

// identity on the padding of a caller-supplied matrix (rows / columns [n, npad); the rest of the padding is already 0)
__global__ void k_pad_identity(double *__restrict__ A, int npad, int n)
{
	const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
	if (i < npad) A[(size_t)i * npad + i] = 1.0;
}

// ---- skinny products ---------------------------------------------------------------------------------
// (a) rows [r_lo, r_lo + 4 gridDim.x) of   OUT = [S -] M[:, j_lo:j_hi) V[j_lo:j_hi)   on columns c0.. of V:
//     TRI:  acc(i) = sum_{j in [j_lo, min(i, j_hi - 1)]} M[i][j] V[j]     (M = W lower triangular; OUT = acc)
//     else: acc(i) = sum_{j in [j_lo, j_hi)} M[i][j] V[j]                  (M = L21; OUT = S - acc when S is given)
// These are the two steps of the block forward substitution L u = [y | H] (emub_api.cu: run_regression).
// grid (nrows/4, ncp/8, B), 256 threads: a CTA owns 4 rows, its 8 warps take the columns in interleaved slices of 32
// (warp w: j_lo + 32 w + lane + 256 t), the 8 partial sums meet in shared memory and are added in warp order -- a fixed
// order, so the result does not depend on the batch or on timing.  Four rows per CTA keeps a single matrix (B = 1, the
// reference's own restart loop through the glue) spread over the whole GPU: 512 CTAs for the 2048-row step at n = 4096.
// V (and S) of slot b: base + comp[b] * stride when comp is given (the training vector / PCA component the slot
// evaluates, shared by all slots), else base + b * stride (per-slot buffer).
// NC = 8: eight columns per CTA column chunk (blockIdx.y); NC = 2: the first two columns only -- y and the constant
// regression function, all there is for regression order 0 -- a quarter of the V traffic and of the multiply-adds.
template <bool TRI, int NC>
__global__ void __launch_bounds__(256) k_rows_times_range(const double *__restrict__ Mbase, long long strideM, int ld, int r_lo,
                                                          int j_lo, int j_hi, const double *__restrict__ Vbase, long long strideV,
                                                          const int *__restrict__ compV, const double *Sbase, long long strideS,
                                                          const int *__restrict__ compS, int ncp, double *Obase, long long strideO)
{
	__shared__ double red[8][4][NC];
	const int b = blockIdx.z, c0 = blockIdx.y * 8;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r0 = r_lo + blockIdx.x * 4;
	const double *M = Mbase + b * strideM;
	const double *V = Vbase + (compV ? compV[b] : b) * strideV;
	double acc[4][NC];
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < NC; c++) acc[r][c] = 0.0;
	const int jmax = TRI ? min(r0 + 3, j_hi - 1) : j_hi - 1;
#pragma unroll 2
	for (int j = j_lo + warp * 32 + lane; j <= jmax; j += 256) {
		double v[NC];
		if (NC == 2) {
			const double2 v0 = *reinterpret_cast<const double2 *>(V + (size_t)j * ncp + c0);
			v[0] = v0.x; v[1] = v0.y;
		} else {
			const double4 v0 = *reinterpret_cast<const double4 *>(V + (size_t)j * ncp + c0);
			const double4 v1 = *reinterpret_cast<const double4 *>(V + (size_t)j * ncp + c0 + 4);
			v[0] = v0.x; v[1] = v0.y; v[2 % NC] = v0.z; v[3 % NC] = v0.w;
			v[4 % NC] = v1.x; v[5 % NC] = v1.y; v[6 % NC] = v1.z; v[7 % NC] = v1.w;
		}
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const double w = (!TRI || j <= r0 + r) ? M[(size_t)(r0 + r) * ld + j] : 0.0;
#pragma unroll
			for (int c = 0; c < NC; c++) acc[r][c] += w * v[c];
		}
	}
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < NC; c++) {
			double s = acc[r][c];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if (lane == 0) red[warp][r][c] = s;
		}
	__syncthreads();
	if (threadIdx.x < 4 * NC) {
		const int r = threadIdx.x / NC, c = threadIdx.x % NC;
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < 8; w++) s += red[w][r][c];
		double *O = Obase + b * strideO + (size_t)(r0 + r) * ncp + c0;
		if (Sbase) s = Sbase[(compS ? compS[b] : b) * strideS + (size_t)(r0 + r) * ncp + c0 + c] - s;
		O[c] = s;
		// the columns this instantiation skips are structural zeros of Yh
		if (NC < 8 && !Sbase && c == 0)
#pragma unroll
			for (int cc = NC; cc < 8; cc++) O[cc] = 0.0;
	}
}

// (b) OUT[j][c0..c0+8) = sum_{i >= ibegin(j)} M[i][j] * V[i][c0..c0+8)
// TRI: M = W lower triangular (i >= j); otherwise all nrows rows (M = K, n x mq).
// grid (ncols/32, ncp/8, B), 256 threads: lanes = 32 consecutive columns, warps stride over rows.
// NC = 8: the eight columns c0..c0+8 of V, 8 warps; NC = 1: column 0 only (alpha = W^T u, all the literal gradient
// reads of AB); NC = 2: columns 0 and 1 ([a | C^-1 1], all there is for regression order 0) -- 16 warps each: twice the
// rows in flight per column block, which is what a single matrix (B = 1) needs.
// Launch with COLS_TIMES_WARPS(NC) * 32 threads.
#define COLS_TIMES_WARPS(NC) ((NC) <= 2 ? 16 : 8)
template <bool TRI, int NC = 8>
__global__ void __launch_bounds__(COLS_TIMES_WARPS(NC) * 32) k_cols_times(const double *__restrict__ Mbase, long long strideM, int ld,
                                                                         int nrows, const double *__restrict__ Vbase,
                                                                         long long strideV, int ncp, double *__restrict__ Obase,
                                                                         long long strideO)
{
	constexpr int NW = COLS_TIMES_WARPS(NC);
	__shared__ double red[NW][32][NC + 1];
	const int b = blockIdx.z, c0 = blockIdx.y * 8;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int j0 = blockIdx.x * 32, j = j0 + lane;
	const double *M = Mbase + b * strideM;
	const double *V = Vbase + b * strideV;
	double acc[NC];
#pragma unroll
	for (int c = 0; c < NC; c++) acc[c] = 0.0;
	const int ibeg = TRI ? j0 : 0;
	if (NC == 1) {
#pragma unroll 8
		for (int i = ibeg + warp; i < nrows; i += NW) {
			double m = M[(size_t)i * ld + j];
			if (TRI && i < j) m = 0.0;
			acc[0] += m * V[(size_t)i * ncp + c0];
		}
	} else if (NC == 2) {
#pragma unroll 8
		for (int i = ibeg + warp; i < nrows; i += NW) {
			double m = M[(size_t)i * ld + j];
			if (TRI && i < j) m = 0.0;
			const double2 v = *reinterpret_cast<const double2 *>(V + (size_t)i * ncp + c0);
			acc[0] += m * v.x; acc[1 % NC] += m * v.y;
		}
	} else {
#pragma unroll 4
		for (int i = ibeg + warp; i < nrows; i += NW) {
			double m = M[(size_t)i * ld + j];
			if (TRI && i < j) m = 0.0;
			const double4 v0 = *reinterpret_cast<const double4 *>(V + (size_t)i * ncp + c0);
			const double4 v1 = *reinterpret_cast<const double4 *>(V + (size_t)i * ncp + c0 + 4);
			acc[0] += m * v0.x; acc[1 % NC] += m * v0.y; acc[2 % NC] += m * v0.z; acc[3 % NC] += m * v0.w;
			acc[4 % NC] += m * v1.x; acc[5 % NC] += m * v1.y; acc[6 % NC] += m * v1.z; acc[7 % NC] += m * v1.w;
		}
	}
#pragma unroll
	for (int c = 0; c < NC; c++) red[warp][lane][c] = acc[c];
	__syncthreads();
	if (NC == 1) {
		if (threadIdx.x < 32) {
			double s = 0.0;
#pragma unroll
			for (int w = 0; w < NW; w++) s += red[w][threadIdx.x][0];
			Obase[b * strideO + (size_t)(j0 + threadIdx.x) * ncp + c0] = s;
		}
	} else if (NC == 2) {
		if (threadIdx.x < 64) {
			const int l = threadIdx.x >> 1, c = threadIdx.x & 1;
			double s = 0.0;
#pragma unroll
			for (int w = 0; w < NW; w++) s += red[w][l][c % NC];
			Obase[b * strideO + (size_t)(j0 + l) * ncp + c0 + c] = s;
		}
	} else {
		const int l = threadIdx.x >> 3, c = threadIdx.x & 7;
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < NW; w++) s += red[w][l][c % NC];
		Obase[b * strideO + (size_t)(j0 + l) * ncp + c0 + c] = s;
	}
}

// ---- Gram partials: G_part[b][chunk][c1][c2] = sum_{r in chunk} UG[r][c1] UG[r][c2] ----------------
// grid (npad/128, B), 256 threads, dynamic smem 128*ncp doubles
__global__ void __launch_bounds__(256) k_gram_partial(const double *__restrict__ UGbase, long long strideUG, int ncp, int nc,
                                                      double *__restrict__ part, int nchunks)
{
	extern __shared__ double sm[];
	const int b = blockIdx.y, ch = blockIdx.x;
	const double *UG = UGbase + b * strideUG + (size_t)ch * TB * ncp;
	for (int idx = threadIdx.x; idx < TB * ncp; idx += 256) sm[idx] = UG[idx];
	__syncthreads();
	double *o = part + ((size_t)b * nchunks + ch) * ((size_t)ncp * ncp);
	for (int idx = threadIdx.x; idx < nc * nc; idx += 256) {
		int c1 = idx / nc, c2 = idx - c1 * nc;
		double s = 0.0;
		for (int r = 0; r < TB; r++) s += sm[r * ncp + c1] * sm[r * ncp + c2];
		o[c1 * ncp + c2] = s;
	}
}

__device__ __forceinline__ double block_sum_256(double v, double *scratch /* >= 8 */)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0.0;
#pragma unroll
	for (int w = 0; w < 8; w++) s += scratch[w];
	return s;
}

// per-point results, stride RES_STRIDE doubles:
//  [0] negL [1] sigma2 [2] status [3] logdet [4] u.z [5] z.z [6] tr(Cinv) [7] alpha.alpha
//  [8 .. 8+MAXNCP) beta   [8+MAXNCP ..) gradient (nthetas-1)
constexpr int RES_BETA = 8;
constexpr int RES_GRAD = 8 + MAXNCP;
constexpr int RES_STRIDE = RES_GRAD + MAXD + 2;

// K6 + likelihood scalar.  One CTA (256 threads) per point.
//  S = sum of Gram partials;  beta = (G^T G)^-1 G^T u  (regression.c:120-176 with C^-1 = W^T W);
//  z = u - G beta = W (y - H beta);  sigma2 = u.z / n (maxmultimin.c:259-263);
//  -L = 0.5 logdet + (n/2) 1.83788 + 0.5 z.z  (estimator-fns.c:48,85-95; logdet per D-1)
// emulator mode: additionally Minv = (G^T G)^-1 (emulator.c:745-769) and UG[:,0] <- z.
__global__ void __launch_bounds__(256) k_small(const double *__restrict__ part, int nchunks, double *UGbase, long long strideUG,
                                               int ncp, int p, int n, int npad, const double *__restrict__ logdet_parts, int nblk,
                                               const int *__restrict__ info, double *__restrict__ res,
                                               double *__restrict__ consts, int emulator_mode, double *__restrict__ Minv)
{
	// S and Lm: (p + 1) x (p + 2) doubles each, in dynamic shared memory (2 * (p + 1) * (p + 2) doubles)
	extern __shared__ double sm_small[];
	const int sld = p + 2;
	double *S = sm_small, *Lm = sm_small + (size_t)(p + 1) * sld;
#define S_(i, j) S[(i) * sld + (j)]
#define L_(i, j) Lm[(i) * sld + (j)]
	__shared__ double beta[MAXNCP];
	__shared__ double scratch[8];
	__shared__ int regbad;
	const int b = blockIdx.x, tid = threadIdx.x;
	const int nc = p + 1;
	double *UG = UGbase + b * strideUG;
	double *r = res + (size_t)b * RES_STRIDE;
	for (int idx = tid; idx < nc * nc; idx += 256) {
		int c1 = idx / nc, c2 = idx - c1 * nc;
		double s = 0.0;
		for (int ch = 0; ch < nchunks; ch++) s += part[((size_t)b * nchunks + ch) * ((size_t)ncp * ncp) + c1 * ncp + c2];
		S_(c1, c2) = s;
	}
	if (tid == 0) regbad = 0;
	__syncthreads();
	// Cholesky of D = S[1..p][1..p] by warp 0 (lanes own rows i, i + 32, ...)
	if (tid < 32) {
		for (int i = tid; i < p; i += 32)
			for (int c = 0; c < p; c++) L_(i, c) = S_(1 + i, 1 + c);
		__syncwarp();
		for (int j = 0; j < p; j++) {
			double djj = L_(j, j);
			if (!(djj > 0.0)) { if (tid == 0) regbad = 1; djj = 1.0; }
			double ljj = sqrt(djj);
			__syncwarp();
			for (int i = tid; i < p; i += 32) {
				if (i == j) L_(j, j) = ljj;
				else if (i > j) L_(i, j) = L_(i, j) / ljj;
			}
			__syncwarp();
			for (int i = j + 1 + tid; i < p; i += 32) {
				double lij = L_(i, j);
				for (int c = j + 1; c <= i; c++) L_(i, c) -= lij * L_(c, j);
			}
			__syncwarp();
		}
		// solve D beta = G^T u  (S[1+i][0])
		if (tid == 0) {
			for (int i = 0; i < p; i++) {
				double s = S_(1 + i, 0);
				for (int c = 0; c < i; c++) s -= L_(i, c) * beta[c];
				beta[i] = s / L_(i, i);
			}
			for (int i = p - 1; i >= 0; i--) {
				double s = beta[i];
				for (int c = i + 1; c < p; c++) s -= L_(c, i) * beta[c];
				beta[i] = s / L_(i, i);
			}
		}
		if (emulator_mode) {
			// Minv column by column: lane c solves D x = e_c
			__syncwarp();
			for (int c = tid; c < p; c += 32) {
				double x[MAXNCP];
				for (int i = 0; i < p; i++) {
					double s = (i == c) ? 1.0 : 0.0;
					for (int k = 0; k < i; k++) s -= L_(i, k) * x[k];
					x[i] = s / L_(i, i);
				}
				for (int i = p - 1; i >= 0; i--) {
					double s = x[i];
					for (int k = i + 1; k < p; k++) s -= L_(k, i) * x[k];
					x[i] = s / L_(i, i);
				}
				for (int i = 0; i < p; i++) Minv[(size_t)b * MAXNCP * MAXNCP + i * MAXNCP + c] = x[i];
			}
		}
	}
	__syncthreads();
	double uz = 0.0, zz = 0.0;
	for (int i = tid; i < npad; i += 256) {
		const double *row = UG + (size_t)i * ncp;
		double u = row[0];
		double z = u;
		for (int c = 0; c < p; c++) z -= row[1 + c] * beta[c];
		uz += u * z;
		zz += z * z;
		if (emulator_mode) UG[(size_t)i * ncp] = z;
	}
	uz = block_sum_256(uz, scratch);
	zz = block_sum_256(zz, scratch);
	if (tid == 0) {
		double logdet = 0.0;
		for (int k = 0; k < nblk; k++) logdet += logdet_parts[(size_t)b * nblk + k];
		logdet *= 2.0;
		const double sigma2 = uz / (double)n;
		const double log_2_pi = 1.83788;  // estimator-fns.c:48 (truncated literal)
		double L = -(1.0 / 2.0) * logdet - (n / 2.0) * log_2_pi;
		L += zz * (-1.0 / 2.0);
		int status = info[b] ? 1 : (regbad ? 2 : 0);
		r[0] = status ? nan("") : -1.0 * L;
		r[1] = status ? nan("") : sigma2;
		r[2] = (double)status;
		r[3] = logdet;
		r[4] = uz;
		r[5] = zz;
		for (int c = 0; c < p; c++) r[RES_BETA + c] = beta[c];
		consts[(size_t)b * CONST_STRIDE + 3] = sigma2;
	}
}
#undef S_
#undef L_

// ---- K4: fused gradient reduction ---------------------------------------------------------------------
// part[b][tile][k] = sum over strictly-lower pairs (i > j) of the 64 x 64 tile of
//      (alpha_i alpha_j - Cinv_ij) * q_k * exp(-a_k q_k),   q_k = (x_ik - x_jk)^2      (power-exp;
//      emulator.c:173-209 with the constant exp(-2 theta_k) factored out, maxmultimin.c:530-538,583-602)
// Matern (deviation D-3): one slot, (alpha_i alpha_j - Cinv_ij) * f(t), t = root * r / rho.
// dC/dtheta is never stored.  grid (npad/64, npad/64, B) (upper tiles exit), 256 threads,
// dynamic smem (2*d*64 + 2*64) doubles.
// EXACT (deviation D-4, optional): the true gradient of the objective evalFnMulti returns,
//   d(-L)/dtheta = 1/2 tr(C^-1 dC) - 1/2 z^T dC z,  z = C^-1 (y - H beta),  dC_ij/dtheta_k = c_ij q_k exp(-2 theta_k)
// i.e. the same sums with z for alpha (res: beta from k_small) and the full product kernel c_ij for its k-th factor --
// one exp per pair instead of d.  The literal formula stays the default.
template <int KERNEL, bool EXACT>
__global__ void __launch_bounds__(256) k_grad_tiles(const double *__restrict__ Cbase, long long strideC, int ld,
                                                    const double *__restrict__ ABbase, long long strideAB, int ncp,
                                                    const double *__restrict__ X, int n, int d,
                                                    const double *__restrict__ consts, double *__restrict__ part, int ntiles64,
                                                    const double *__restrict__ res, int p)
{
	const int bj = blockIdx.x, bi = blockIdx.y, b = blockIdx.z;
	if (bj > bi) return;
	extern __shared__ double sm[];
	double *sXi = sm, *sXj = sm + d * CT;
	double *sAi = sm + 2 * d * CT, *sAj = sAi + CT;
	__shared__ double sc[CONST_STRIDE];
	__shared__ double red[8][MAXD];
	__shared__ double stab[64 * EXP_REP];
	exp_table_load<EXP_REP>(stab);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const double *cg = consts + (size_t)b * CONST_STRIDE;
	for (int i = tid; i < 4 + 5 * d; i += 256) sc[i] = cg[i];
	const int i0 = bi * CT, j0 = bj * CT;
	// power-exponential: coordinates staged times s_k = sqrt(a_k 64/ln2), a_k = 0.5 exp(-2 theta_k), so that the exponent
	// of the k-th factor, -a_k q_k, is z = -(x'_ik - x'_jk)^2 in exp_scaled's units and q_k = -z / s_k^2
	for (int idx = tid; idx < CT * d; idx += 256) {
		int r = idx / d, k = idx - r * d;
		const double sk = (KERNEL == 1) ? cg[4 + 4 * d + k] : 1.0;
		sXi[k * CT + r] = (i0 + r < n) ? X[(size_t)(i0 + r) * d + k] * sk : 0.0;
		sXj[k * CT + r] = (j0 + r < n) ? X[(size_t)(j0 + r) * d + k] * sk : 0.0;
	}
	const double *AB = ABbase + b * strideAB;
	if (tid < 2 * CT) {
		const double *row = AB + (size_t)((tid < CT) ? i0 + tid : j0 + tid - CT) * ncp;
		double a = row[0];
		if (EXACT) {  // z = alpha - (C^-1 H) beta
			const double *beta = res + (size_t)b * RES_STRIDE + RES_BETA;
			for (int c = 0; c < p; c++) a -= row[1 + c] * beta[c];
		}
		if (tid < CT) sAi[tid] = a; else sAj[tid - CT] = a;
	}
	__syncthreads();
	const int tx = tid & 15, ty = tid >> 4;
	const double *C = Cbase + b * strideC;
	double w[4][4];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const int li = ty + 16 * r, gi = i0 + li;
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const int lj = tx + 16 * c, gj = j0 + lj;
			double v = 0.0;
			if (gi > gj && gi < n) v = sAi[li] * sAj[lj] - C[(size_t)gi * ld + gj];
			w[r][c] = v;
		}
	}
	const int nslots = (KERNEL == 1) ? d : 1;
	if (KERNEL == 1 && EXACT) {
		// w <- w c_ij, the unit-amplitude kernel from its exponent sum_k a_k q_k; then one weighted sum of q_k per theta_k
		double ex[4][4];
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) ex[r][c] = 0.0;
		for (int k = 0; k < d; k++) {
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const double xi = sXi[k * CT + ty + 16 * r];
#pragma unroll
				for (int c = 0; c < 4; c++) {
					const double dl = xi - sXj[k * CT + tx + 16 * c];
					ex[r][c] = fma(dl, dl, ex[r][c]);
				}
			}
		}
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) w[r][c] *= exp_scaled<EXP_REP>(-ex[r][c], stab);
		for (int k = 0; k < d; k++) {
			double s = 0.0;
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const double xi = sXi[k * CT + ty + 16 * r];
#pragma unroll
				for (int c = 0; c < 4; c++) {
					const double dl = xi - sXj[k * CT + tx + 16 * c];
					s += w[r][c] * (dl * dl);
				}
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if (lane == 0) red[warp][k] = s / (sc[4 + 4 * d + k] * sc[4 + 4 * d + k]);  // q_k = (staged difference)^2 / s_k^2
		}
	} else if (KERNEL == 1) {
		for (int k = 0; k < d; k++) {
			double s = 0.0;
#pragma unroll
			for (int r = 0; r < 4; r++) {
				const double xi = sXi[k * CT + ty + 16 * r];
#pragma unroll
				for (int c = 0; c < 4; c++) {
					const double dl = xi - sXj[k * CT + tx + 16 * c];
					const double qs = dl * dl;  // q_k s_k^2 = a_k q_k in exp_scaled's units
					s += w[r][c] * (qs * exp_scaled<EXP_REP>(-qs, stab));
				}
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			if (lane == 0) red[warp][k] = s / (sc[4 + 4 * d + k] * sc[4 + 4 * d + k]);  // back to q_k exp(-a_k q_k)
		}
	} else {
		const double root = (KERNEL == 2) ? 1.732050808 : 2.236067978;
		const double rho = sc[2];
		double s = 0.0;
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int c = 0; c < 4; c++) {
				double r2 = 0.0;
				for (int k = 0; k < d; k++) {
					const double dl = sXi[k * CT + ty + 16 * r] - sXj[k * CT + tx + 16 * c];
					r2 += dl * dl;
				}
				const double t = root * (sqrt(r2) / rho);
				const double f = (KERNEL == 2) ? t * t * exp_neg<EXP_REP>(-t, stab) : (t * t / 3.0) * (1.0 + t) * exp_neg<EXP_REP>(-t, stab);
				s += w[r][c] * f;
			}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (lane == 0) red[warp][0] = s;
	}
	__syncthreads();
	if (tid < nslots) {
		double s = 0.0;
#pragma unroll
		for (int wv = 0; wv < 8; wv++) s += red[wv][tid];
		// lower-triangular tile index
		const size_t tile = (size_t)bi * (bi + 1) / 2 + bj;
		const size_t ntl = (size_t)ntiles64 * (ntiles64 + 1) / 2;
		part[((size_t)b * ntl + tile) * nslots + tid] = s;
	}
}

// final gradient assembly, one CTA per point (maxmultimin.c:514-538):
//  g_nugget = -( -0.5 nug tr(Cinv) + 0.5 nug alpha.alpha )
//  g_len_k  = -sigma2 * exp(-2 theta_k) * sum_{i>j} (...)      (the i<j half is the mirror image)
__global__ void __launch_bounds__(256) k_grad_final(const double *__restrict__ part, int ntiles64, const double *__restrict__ Cbase,
                                                    long long strideC, int ld, const double *__restrict__ ABbase, long long strideAB,
                                                    int ncp, int n, int d, int kernel, const double *__restrict__ consts,
                                                    double *__restrict__ res, int exact, int p)
{
	__shared__ double scratch[8];
	const int b = blockIdx.x, tid = threadIdx.x;
	const double *C = Cbase + b * strideC;
	const double *AB = ABbase + b * strideAB;
	const double *cg = consts + (size_t)b * CONST_STRIDE;
	double *r = res + (size_t)b * RES_STRIDE;
	double tr = 0.0, aa = 0.0;
	for (int i = tid; i < n; i += 256) {
		tr += C[(size_t)i * ld + i];
		double a = AB[(size_t)i * ncp];
		if (exact)
			for (int c = 0; c < p; c++) a -= AB[(size_t)i * ncp + 1 + c] * r[RES_BETA + c];
		aa += a * a;
	}
	tr = block_sum_256(tr, scratch);
	aa = block_sum_256(aa, scratch);
	const size_t ntl = (size_t)ntiles64 * (ntiles64 + 1) / 2;
	const int nslots = (kernel == 1) ? d : 1;
	const bool failed = r[2] != 0.0;
	const double sigma2 = cg[3];
	const double amp = exact ? 1.0 : exp(log(sigma2));  // maxmultimin.c:514; the objective itself has unit amplitude
	for (int k = 0; k < nslots; k++) {
		double s = 0.0;
		for (size_t tix = tid; tix < ntl; tix += 256) s += part[((size_t)b * ntl + tix) * nslots + k];
		s = block_sum_256(s, scratch);
		if (tid == 0) {
			const double e2 = (kernel == 1) ? cg[4 + 2 * d + k] : 1.0;
			r[RES_GRAD + 1 + k] = failed ? nan("") : -1.0 * (amp * e2 * s);
		}
	}
	if (tid == 0) {
		const double nug = cg[1];
		r[6] = tr;
		r[7] = aa;
		r[RES_GRAD] = failed ? nan("") : -1.0 * (-0.5 * nug * tr + 0.5 * nug * aa);
	}
}

// test hook: the device exp on an array
__global__ void k_debug_exp(const double *__restrict__ x, int n, double *__restrict__ out)
{
	__shared__ double stab[64];
	exp_table_load(stab);
	__syncthreads();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = exp_neg(x[i], stab);
}

// test hook: exp_scaled with the replicated table, z = x * 64 / ln 2 formed on the device like the kernels do
__global__ void k_debug_exp_scaled(const double *__restrict__ x, int n, double *__restrict__ out)
{
	__shared__ double stab[64 * EXP_REP];
	exp_table_load<EXP_REP>(stab);
	__syncthreads();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = exp_scaled<EXP_REP>(x[i] * EXP_SCALE, stab);
}

// pack results for the device-pointer API: out[b] = (negL, sigma2, status, logdet, grad[nth1])
__global__ void k_pack_results(const double *__restrict__ res, int B, int nth1, double *__restrict__ out)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= B) return;
	const double *r = res + (size_t)b * RES_STRIDE;
	double *o = out + (size_t)b * (nth1 + 4);
	o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = r[3];
	for (int k = 0; k < nth1; k++) o[4 + k] = r[RES_GRAD + k];
}

// ---- prediction epilogue (emulator.c:672-704, :720-785; emulator_struct.c:124-143) ---------------------
//  mean = h.beta + k.a               a = C^-1 (y - H beta)
//  var  = kappa - |L^-1 k|^2 + rho^T Minv rho,   rho = h - (C^-1 H)^T k
// KA[q][0] = k.a, KA[q][1+c] = ((C^-1 H)^T k)_c ;  vsq_part[blk][q] = partial |W k|^2
__global__ void __launch_bounds__(128) k_pred_final(const double *__restrict__ Q, int mq, int d, int order, int p,
                                                    const double *__restrict__ KA, int ncp, const double *__restrict__ vsq_part,
                                                    int nblk, int ldq, const double *__restrict__ beta,
                                                    const double *__restrict__ Minv, double kappa,
                                                    double *__restrict__ mean, double *__restrict__ var, int nka, long long ka_stride,
                                                    const FewSet *__restrict__ set, long long kastride_z, int mstride)
{
	if (set) {  // component z = blockIdx.y of a few-points call over several emulators
		const int z = blockIdx.y;
		beta = set[z].beta;
		Minv = set[z].Minv;
		kappa = set[z].kappa;
		KA += (size_t)z * kastride_z;
		vsq_part += (size_t)z * 8;
		mean += (size_t)z * mstride;
		var += (size_t)z * mstride;
	}
	extern __shared__ double sM[];  // p x p
	__shared__ double sb[MAXNCP];
	for (int i = threadIdx.x; i < p * p; i += blockDim.x) sM[i] = Minv[(i / p) * MAXNCP + (i % p)];
	for (int i = threadIdx.x; i < p; i += blockDim.x) sb[i] = beta[i];
	__syncthreads();
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= mq) return;
	const double *x = Q + (size_t)q * d;
	// KA comes in nka partial sums (1 on the batched path)
	double ka[MAXNCP + 1];
	for (int c = 0; c <= p; c++) {
		double s = 0.0;
		for (int r = 0; r < nka; r++) s += KA[(size_t)r * ka_stride + (size_t)q * ncp + c];
		ka[c] = s;
	}
	double rho[MAXNCP];
	double hb = 0.0;
	for (int c = 0; c < p; c++) {
		double h;
		if (c == 0) h = 1.0;
		else {
			int k = (c - 1) % d, pw = (c - 1) / d;
			double xv = x[k];
			h = (pw == 0) ? xv : (pw == 1 ? xv * xv : xv * xv * xv);
		}
		hb += h * sb[c];
		rho[c] = h - ka[1 + c];
	}
	double reg = 0.0;
	for (int i = 0; i < p; i++) {
		double s = 0.0;
		for (int c = 0; c < p; c++) s += sM[i * p + c] * rho[c];
		reg += rho[i] * s;
	}
	double vs = 0.0;
	for (int k = 0; k < nblk; k++) vs += vsq_part[(size_t)k * ldq + q];
	mean[q] = hb + ka[0];
	var[q] = kappa - vs + reg;
}

// ---- a handful of query points (<= 8): latency path ---------------------------------------------------------
// The batched pass multiplies W by a 128-column block of K: for one MCMC-style query that is a single wave of CTAs
// each walking a K range of up to n, 220 us at n = 4096.  Here W K[:, 0:8] is a triangular matrix - skinny product
// split over rows AND over chunks of FEW_JC columns of W (enough CTAs to stream W at memory speed), followed by
// one CTA that adds the chunks up, squares, and forms the two small reductions of the prediction.
constexpr int FEW_MAX = 8;      // query points per call on this path
constexpr int FEW_JC = 512;     // columns of W per partial product

// T[js][i][0..NCOL) = sum_{j in chunk js, j <= i} W[i][j] K[j][0..NCOL)     grid (npad/32, ceil(npad/FEW_JC)), 256 threads
// NCOL = 8, or 1 for the single query point of an MCMC step: few registers, so twice the loads in flight.
// set != null: component z = blockIdx.z uses set[z].W, the columns [z * kstride, ..) of K and T + z * tstride
template <int NCOL>
__global__ void __launch_bounds__(256) k_few_wk(const double *__restrict__ W, int ld, const double *__restrict__ K, int ldk,
                                                double *__restrict__ T, int npad, const FewSet *__restrict__ set, int kstride,
                                                long long tstride)
{
	constexpr int UNROLL = (NCOL == 1) ? 8 : 4;
	if (set) {
		W = set[blockIdx.z].W;
		K += (size_t)blockIdx.z * kstride;
		T += (size_t)blockIdx.z * tstride;
	}
	const int js = blockIdx.y, j_lo = js * FEW_JC;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int r0 = blockIdx.x * 32 + warp * 4;
	if (j_lo > r0 + 3) return;  // the chunk lies right of the diagonal for these rows
	const int jend = min(j_lo + FEW_JC - 1, r0 + 3);
	double acc[4][NCOL];
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < NCOL; c++) acc[r][c] = 0.0;
	// UNROLL steps of 32 columns at a time, every load of the group issued before the first use (the loop bound is
	// not a compile-time constant, so the compiler would not overlap the iterations on its own)
	for (int jb = j_lo + lane; jb <= jend; jb += 32 * UNROLL) {
		double wv[UNROLL][4];
		double v[UNROLL][NCOL];
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const int j = jb + 32 * u;
			const bool in = j <= jend;
#pragma unroll
			for (int r = 0; r < 4; r++) wv[u][r] = (in && j <= r0 + r) ? W[(size_t)(r0 + r) * ld + j] : 0.0;
			if (NCOL == 1) v[u][0] = in ? K[(size_t)j * ldk] : 0.0;
			else {
#pragma unroll
				for (int c = 0; c < NCOL; c += 4) {
					const double4 t = in ? *reinterpret_cast<const double4 *>(K + (size_t)j * ldk + c) : make_double4(0.0, 0.0, 0.0, 0.0);
					v[u][c] = t.x; v[u][c + 1] = t.y; v[u][c + 2] = t.z; v[u][c + 3] = t.w;
				}
			}
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++)
#pragma unroll
			for (int r = 0; r < 4; r++)
#pragma unroll
				for (int c = 0; c < NCOL; c++) acc[r][c] += wv[u][r] * v[u][c];
	}
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < NCOL; c++) {
			double s = acc[r][c];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
			acc[r][c] = s;
		}
	if (lane < 4) {
		double *o = T + ((size_t)js * npad + r0 + lane) * 8;
#pragma unroll
		for (int r = 0; r < 4; r++)
			if (lane == r) {
#pragma unroll
				for (int c = 0; c < NCOL; c++) o[c] = acc[r][c];
#pragma unroll
				for (int c = NCOL; c < 8; c++) o[c] = 0.0;
			}
	}
}

// FEW_ROWS rows per CTA, one row per thread: partial vsq[blk][c] = sum_i (sum_js T[js][i][c])^2 and partial
// KA[blk][c][pp] = sum_i K[i][c] AB[i][pp]   (c < 8, pp < ncp); k_pred_final adds the partials up.  grid npad / FEW_ROWS
constexpr int FEW_ROWS = 128;
// set != null: component z = blockIdx.y; its vsq partials go to the columns [8 z, 8 z + 8), its KA partials to KA + z * kastride
__global__ void __launch_bounds__(FEW_ROWS) k_few_finish(const double *__restrict__ T, int npad, const double *__restrict__ K, int ldk,
                                                         const double *__restrict__ AB, int ncp, double *__restrict__ vsq, int ldv,
                                                         double *__restrict__ KA, const FewSet *__restrict__ set, int kstride,
                                                         long long tstride, long long kastride)
{
	if (set) {
		const int z = blockIdx.y;
		AB = set[z].AB;
		K += (size_t)z * kstride;
		T += (size_t)z * tstride;
		vsq += (size_t)z * 8;
		KA += (size_t)z * kastride;
	}
	__shared__ double red[FEW_ROWS / 32][65];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int i = blockIdx.x * FEW_ROWS + tid;
	double t[8];
#pragma unroll
	for (int c = 0; c < 8; c++) t[c] = 0.0;
	const int nch = i / FEW_JC + 1;
	for (int js = 0; js < nch; js++) {
		const double4 a = *reinterpret_cast<const double4 *>(T + ((size_t)js * npad + i) * 8);
		const double4 b = *reinterpret_cast<const double4 *>(T + ((size_t)js * npad + i) * 8 + 4);
		t[0] += a.x; t[1] += a.y; t[2] += a.z; t[3] += a.w; t[4] += b.x; t[5] += b.y; t[6] += b.z; t[7] += b.w;
	}
#pragma unroll
	for (int c = 0; c < 8; c++) {
		double s = t[c] * t[c];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		if (lane == 0) red[warp][c] = s;
	}
	__syncthreads();
	if (tid < 8) {
		double s = 0.0;
		for (int w = 0; w < FEW_ROWS / 32; w++) s += red[w][tid];
		vsq[(size_t)blockIdx.x * ldv + tid] = s;
	}
	const double4 k0 = *reinterpret_cast<const double4 *>(K + (size_t)i * ldk);
	const double4 k1 = *reinterpret_cast<const double4 *>(K + (size_t)i * ldk + 4);
	const double kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
	for (int p0 = 0; p0 < ncp; p0 += 8) {
		const double4 a0 = *reinterpret_cast<const double4 *>(AB + (size_t)i * ncp + p0);
		const double4 a1 = *reinterpret_cast<const double4 *>(AB + (size_t)i * ncp + p0 + 4);
		const double av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
		__syncthreads();
#pragma unroll
		for (int c = 0; c < 8; c++)
#pragma unroll
			for (int q = 0; q < 8; q++) {
				double s = kv[c] * av[q];
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
				if (lane == 0) red[warp][c * 8 + q] = s;
			}
		__syncthreads();
		if (tid < 64) {
			double s = 0.0;
			for (int w = 0; w < FEW_ROWS / 32; w++) s += red[w][tid];
			const int c = tid >> 3, q = tid & 7;
			KA[((size_t)blockIdx.x * 8 + c) * ncp + p0 + q] = s;
		}
	}
}

// PCA back-projection (multivar_support.c:126-151): mean_i = ybar_i + sum_j U_ij sqrt(lambda_j) m_j,
// var_i = sum_j U_ij^2 lambda_j v_j.  pm / pv: [nr][ldq] per-component predictions; out: [mq][nt].
__global__ void __launch_bounds__(128) k_backproject(const double *__restrict__ pm, const double *__restrict__ pv, int ldq,
                                                     int mq, int nt, int nr, const double *__restrict__ ybar,
                                                     const double *__restrict__ evecs, const double *__restrict__ evals,
                                                     double *__restrict__ mean_out, double *__restrict__ var_out)
{
	const int q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= mq) return;
	for (int i = 0; i < nt; i++) {
		double s = 0.0, v = 0.0;
		for (int j = 0; j < nr; j++) {
			const double u = evecs[i * nr + j];
			s += u * sqrt(evals[j]) * pm[(size_t)j * ldq + q];
			v += (u * u) * evals[j] * pv[(size_t)j * ldq + q];
		}
		mean_out[(size_t)q * nt + i] = ybar[i] + s;
		var_out[(size_t)q * nt + i] = v;
	}
}

}  // namespace emub
