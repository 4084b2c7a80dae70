// emub_potf2.cuh -- leaf of the recursive factorisation: Cholesky of one 128 x 128 diagonal block together with
// its triangular inverse, register resident, blocked by panels of 8 columns.
//
// The block is cut into a 16 x 16 grid of 8 x 8 sub-blocks; the 136 lower ones live in the registers of 136
// threads (64 doubles each).  For every panel p (columns J = [8p, 8p+8)):
//   A  a helper warp factorises the diagonal sub-block (p,p), D = L_D L_D^T, inverts L_D (W_D = L_D^-1) and
//      publishes both -- one panel AHEAD: it works from a parked copy of the sub-block plus the rank-8 term of the
//      previous panel, while the other warps are still in step C of that panel, so the 8 dependent reciprocal
//      square roots are off the critical path
//   B  all threads share the two small transforms against W_D, one panel row / one column each (36 FMA):
//      L_R = P W_D^T for the raw panel rows P that their owners parked in shared memory (-> the columns J of L),
//      and E_J <- W_D E_J for the rows J of the inverse left of the panel (-> final rows J of W)
//   C  every thread with rows below the panel applies the rank-8 update
//          a(r,c) -= sum_k L(r, 8p+k) V(k, c),   V(k,c) = L(c, 8p+k) right of the panel (Cholesky update; kept as a
//                                                         transposed copy so that all threads run the same code),
//                                                         E_J(k, c) up to and including it (the same row operations
//                                                         applied to the identity, kept in the eliminated slots)
//      512 independent DFMA per thread.
// Two barriers per panel (32 per block) instead of one per column (128), no per-column role logic, every register
// index is a compile-time constant.  When the sweep ends the registers hold W = L^-1; L was collected in a packed
// shared array on the way and is written out coalesced.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "emub_gemm.cuh"

namespace emub {

// 8 warps = 2 per warp scheduler (scheduler = warp id mod 4).  The 136 sub-block owners fill warps 0-3 and a quarter
// of warp 4; warp 7 factorises the diagonal sub-blocks: it shares its scheduler with warp 3, whose sub-blocks (block
// rows 3-8) retire first, so that for the second half of the sweep the dependent FP64 chain of step A does not queue
// behind an owner's DFMA stream (8.4 -> 13 cycles per dependent DFMA with a streaming warp on the same scheduler,
// tools/fp64_contention.cu).  Warps 5 and 6 only take part in step B.
constexpr int POTF2_THREADS = 256;
constexpr int POTF2_NBLOCKS = 136;
constexpr int POTF2_HELPER0 = 224;    // first thread of the helper warp (warp 7)
constexpr int POTF2_LPACK = TB * (TB + 1) / 2;
// Shared-memory layouts of the two panels, chosen so that the sub-block owners of one warp (same k, up to 16
// different block rows / block columns) hit different banks:
//   L panel: row i at i * 9 + (i >> 3); rows of consecutive block rows are 73 doubles apart (146 words = 18 mod 32)
//   V panel: row k at k * 160, column g at (g >> 3) * 10 + (g & 7); the 8-column groups of consecutive block columns
//            are 80 bytes apart, so the 16-byte loads of a quarter warp cover all 32 banks once
constexpr int POTF2_LS = 9;
constexpr int POTF2_LPANEL = TB * POTF2_LS + TB / 8;
constexpr int POTF2_VS = (TB / 8) * 10;
__device__ __forceinline__ int potf2_lrow(int i) { return i * POTF2_LS + (i >> 3); }
__device__ __forceinline__ int potf2_vcol(int g) { return (g >> 3) * 10 + (g & 7); }
constexpr int POTF2_SMEM_DOUBLES = POTF2_LPACK + 2 * POTF2_LPANEL + 2 * 8 * POTF2_VS + 6 * 64 + TB;
constexpr int POTF2_SMEM_BYTES = POTF2_SMEM_DOUBLES * 8;

// 1 / sqrt(x) for a normal, positive x without the library routine's special-case branch: hardware seed (2^-22) and two
// Newton steps, about one ulp
__device__ __forceinline__ double potf2_rsqrt(double x)
{
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
	for (int it = 0; it < 2; it++) {
		const double e = fma(-(x * y), y, 1.0);
		y = fma(0.5 * y, e, y);
	}
	return y;
}

// Step A for panel q, by the helper warp: D = parked diagonal sub-block (updated through panel q - 2) minus the
// rank-8 contribution of panel q - 1 (its rows of L are final in LPprev), then D = L_D L_D^T, W_D = L_D^-1.
// Publishes W_D (WD, and the columns J of VR), L_D (packed L) and the pivots.
__device__ __forceinline__ void potf2_diag_block(int q, int lane, const double *Dpark, const double *LPprev, double *Dwork,
                                                 double *WD, double *VR, double *Lp, double *pivots, int *s_bad)
{
	// 64 elements over 32 lanes
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const int e = lane + 32 * h, r = e >> 3, c = e & 7;
		double v = Dpark[e];
		if (LPprev) {
			const double *lr = LPprev + potf2_lrow(q * 8 + r), *lc = LPprev + potf2_lrow(q * 8 + c);
#pragma unroll
			for (int k = 0; k < 8; k++) v -= lr[k] * lc[k];
		}
		Dwork[e] = v;
	}
	__syncwarp();
	if (lane == 0) {
		// One thread, everything in registers, no branch inside: the only long dependent chain is pivot -> rsqrt ->
		// column scale -> update -> next pivot.  Row j of the inverse (forward substitution,
		// W(j,c) = -W(j,j) sum_{c <= k < j} L(j,k) W(k,c)) needs nothing from step j but the final scaling, so its
		// sums sit in the shadow of the reciprocal square root.
		double a[8][8], w[8][8];
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int c = 0; c <= r; c++) a[r][c] = Dwork[r * 8 + c];
#pragma unroll
		for (int j = 0; j < 8; j++) {
			double pv = a[j][j];
			// gsl_linalg_cholesky_decomp fails on a pivot <= 0 (GSL_EDOM, maxmultimin.c:325-327); so does this, plus the
			// pivots the reciprocal square root below cannot take: subnormal (< 2.3e-308), inf or NaN
			const bool ok = (pv >= 2.2250738585072014e-308) && (pv < 1.0e307);
			if (!ok) { *s_bad = 1; pv = 1.0; }
			pivots[q * 8 + j] = pv;
			const double isq = potf2_rsqrt(pv);
			double t[8];
#pragma unroll
			for (int c = 0; c < j; c++) {
				double sum = 0.0;
#pragma unroll
				for (int k = c; k < j; k++) sum += a[j][k] * w[k][c];
				t[c] = sum;
			}
			a[j][j] = pv * isq;
#pragma unroll
			for (int i = j + 1; i < 8; i++) a[i][j] *= isq;
#pragma unroll
			for (int i = j + 1; i < 8; i++)
#pragma unroll
				for (int c = j + 1; c <= i; c++) a[i][c] -= a[i][j] * a[c][j];
			w[j][j] = isq;
#pragma unroll
			for (int c = 0; c < j; c++) w[j][c] = -t[c] * isq;
		}
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int c = 0; c <= r; c++) {
				Dwork[64 + r * 8 + c] = a[r][c];  // L_D
				Dwork[r * 8 + c] = w[r][c];       // W_D
			}
	}
	__syncwarp();
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const int e = lane + 32 * h, r = e >> 3, c = e & 7;
		const double w = (c <= r) ? Dwork[e] : 0.0;
		WD[e] = w;
		VR[r * POTF2_VS + q * 10 + c] = w;
		if (c <= r) {
			const int gi = q * 8 + r;
			Lp[gi * (gi + 1) / 2 + q * 8 + c] = Dwork[64 + e];
		}
	}
}

__global__ void __launch_bounds__(POTF2_THREADS) k_potf2(const double *Abase, long long strideA, double *Lbase, long long strideL,
                                                         double *Wbase, long long strideW, int ld, int kblk, int nblk,
                                                         double *__restrict__ logdet_parts, int *__restrict__ info)
{
	extern __shared__ __align__(16) double sm[];
	double *Lp = sm;                                   // packed lower triangle of L: (i, c) at i (i + 1) / 2 + c
	double *Lpanel = Lp + POTF2_LPACK;                 // [2][128][9]   L(i, 8p + k)
	double *Vrow = Lpanel + 2 * POTF2_LPANEL;          // [2][8][160]   rows J of the inverse in progress
	double *Dbuf = Vrow + 2 * 8 * POTF2_VS;            // [2][8][8]     W_D
	double *Dpark = Dbuf + 2 * 64;                     // [2][8][8]     diagonal sub-block q, updated through panel q - 2
	double *Dwork = Dpark + 2 * 64;                    // [2][8][8]     helper scratch
	double *pivots = Dwork + 2 * 64;                   // [128] a_jj at elimination time (log-determinant)
	__shared__ int s_bad;
	const int tid = threadIdx.x, b = blockIdx.x;
	const size_t blk = (size_t)kblk * TB * ld + (size_t)kblk * TB;
	const double *A = Abase + b * strideA + blk;
	double *Lg = Lbase + b * strideL + blk;
	double *Wg = Wbase + b * strideW + blk;
	const bool helper = tid >= POTF2_HELPER0;
	const int oslot = tid;
	const bool active = oslot < POTF2_NBLOCKS;
	// sub-block t = 135 - slot in row-major triangular order: the first slots hold the last block rows (live until the
	// end), the last ones the sub-blocks of the first rows (retired after a few panels)
	int bi = 0, bj = 0;
	if (active) {
		const int t = POTF2_NBLOCKS - 1 - oslot;
		while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
		bj = t - bi * (bi + 1) / 2;
	}
	if (tid == 0) s_bad = 0;
	double a[8][8];
	if (active) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const double *src = A + (size_t)(bi * 8 + r) * ld + bj * 8;
#pragma unroll
			for (int c = 0; c < 8; c += 2) {
				double2 v = *reinterpret_cast<const double2 *>(src + c);
				a[r][c] = v.x;
				a[r][c + 1] = v.y;
			}
		}
		// park the first two diagonal sub-blocks and the raw panel 0
		if (bi == bj && bi < 2) {
#pragma unroll
			for (int r = 0; r < 8; r++)
#pragma unroll
				for (int c = 0; c < 8; c++) Dpark[bi * 64 + r * 8 + c] = a[r][c];
		}
		if (bj == 0 && bi > 0) {
#pragma unroll
			for (int r = 0; r < 8; r++)
#pragma unroll
				for (int k = 0; k < 8; k++) { Lpanel[potf2_lrow(bi * 8 + r) + k] = a[r][k]; a[r][k] = 0.0; }
		}
	}
	// strictly upper sub-blocks of both outputs are zero
	for (int idx = tid; idx < TB * TB / 2; idx += POTF2_THREADS) {
		const int i = idx >> 6, c = (idx & 63) * 2;
		if ((c >> 3) > (i >> 3)) {
			*reinterpret_cast<double2 *>(Lg + (size_t)i * ld + c) = make_double2(0.0, 0.0);
			*reinterpret_cast<double2 *>(Wg + (size_t)i * ld + c) = make_double2(0.0, 0.0);
		}
	}
	__syncthreads();
	if (helper) potf2_diag_block(0, tid - POTF2_HELPER0, Dpark, nullptr, Dwork, Dbuf, Vrow, Lp, pivots, &s_bad);

#ifdef EMUB_POTF2_TIMING
	// per panel and warp: clock at arrival at barrier 1 and at barrier 2 (lane 0 of every warp; ptxas hoists a clock
	// read above the barrier that precedes it, so "after the barrier" cannot be stamped)
	__shared__ long long s_clk[16][8][2];
#define POTF2_STAMP(slot) do { if ((tid & 31) == 0) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); s_clk[p][tid >> 5][slot] = t_; } } while (0)
#else
#define POTF2_STAMP(slot) do { } while (0)
#endif
	for (int p = 0; p < 16; p++) {
		double *LP = Lpanel + (p & 1) * POTF2_LPANEL;
		double *VR = Vrow + (p & 1) * 8 * POTF2_VS;
		double *WD = Dbuf + (p & 1) * 64;
		POTF2_STAMP(0);
		__syncthreads();  // W_D(p), L_D(p) are there (helper warp, during the previous rank-8 update)
		// ---- B: all threads share the two small transforms against W_D ------------------------------------------------
		{
			// the 36 entries of W_D first, as independent loads (otherwise every FMA below waits for its own load)
			double w[36];
#pragma unroll
			for (int k = 0; k < 8; k++)
#pragma unroll
				for (int c = 0; c <= k; c++) w[k * (k + 1) / 2 + c] = WD[k * 8 + c];
			// one panel row each:  L_R(i, k) = sum_{c <= k} P(i, c) W_D(k, c)   (raw P was parked in LP by its owners)
			const int gi = p * 8 + 8 + tid;
			if (gi < TB) {
				double pr[8], lr[8];
#pragma unroll
				for (int c = 0; c < 8; c++) pr[c] = LP[potf2_lrow(gi) + c];
#pragma unroll
				for (int k = 0; k < 8; k++) {
					double s = 0.0;
#pragma unroll
					for (int c = 0; c <= k; c++) s += pr[c] * w[k * (k + 1) / 2 + c];
					lr[k] = s;
				}
				const int base = gi * (gi + 1) / 2 + p * 8;
#pragma unroll
				for (int k = 0; k < 8; k++) {
					LP[potf2_lrow(gi) + k] = lr[k];
					VR[k * POTF2_VS + potf2_vcol(gi)] = lr[k];  // the same column of L as the right factor V(k, gi) of the rank-8 update
					Lp[base + k] = lr[k];
				}
			}
			// one column of the rows J of the inverse each:  E_J(r, c) <- sum_{k <= r} W_D(r, k) E_J(k, c)
			const int gc = POTF2_THREADS - 1 - tid;
			if (gc < p * 8) {
				double er[8];
#pragma unroll
				for (int k = 0; k < 8; k++) er[k] = VR[k * POTF2_VS + potf2_vcol(gc)];
#pragma unroll
				for (int r = 0; r < 8; r++) {
					double s = 0.0;
#pragma unroll
					for (int k = 0; k <= r; k++) s += w[r * (r + 1) / 2 + k] * er[k];
					VR[r * POTF2_VS + potf2_vcol(gc)] = s;
				}
			}
		}
		POTF2_STAMP(1);
		__syncthreads();
		// ---- C: rank-8 update of every row below the panel; the owners of rows J pick up their final values; the
		//         helper warp factorises the next diagonal sub-block at the same time ------------------------------------
		if (helper) {
			if (p + 1 < 16)
				potf2_diag_block(p + 1, tid - POTF2_HELPER0, Dpark + ((p + 1) & 1) * 64, LP, Dwork, Dbuf + ((p + 1) & 1) * 64,
				                 Vrow + ((p + 1) & 1) * 8 * POTF2_VS, Lp, pivots, &s_bad);
		} else if (active && bi > p) {
			const double *lrow = LP + potf2_lrow(bi * 8);
			const double *vcol = VR + bj * 10;
#pragma unroll
			for (int k = 0; k < 8; k++) {
				double l[8], v[8];
#pragma unroll
				for (int r = 0; r < 8; r++) l[r] = lrow[r * POTF2_LS + k];
#pragma unroll
				for (int c = 0; c < 8; c += 2) {
					const double2 t = *reinterpret_cast<const double2 *>(vcol + k * POTF2_VS + c);
					v[c] = t.x;
					v[c + 1] = t.y;
				}
#pragma unroll
				for (int r = 0; r < 8; r++)
#pragma unroll
					for (int c = 0; c < 8; c++) a[r][c] -= l[r] * v[c];
			}
			// park the raw data of the coming panels in the other halves of the buffers (their last readers are two
			// barriers behind) and clear the slots that start to hold the inverse
			if (p + 1 < 16) {
				double *LPn = Lpanel + ((p + 1) & 1) * POTF2_LPANEL;
				double *VRn = Vrow + ((p + 1) & 1) * 8 * POTF2_VS;
				if (bj == p + 1 && bi > p + 1) {
#pragma unroll
					for (int r = 0; r < 8; r++)
#pragma unroll
						for (int k = 0; k < 8; k++) { LPn[potf2_lrow(bi * 8 + r) + k] = a[r][k]; a[r][k] = 0.0; }
				}
				if (bi == p + 1 && bj <= p) {
#pragma unroll
					for (int k = 0; k < 8; k++)
#pragma unroll
						for (int c = 0; c < 8; c++) VRn[k * POTF2_VS + bj * 10 + c] = a[k][c];
				}
				if (bi == p + 2 && bj == p + 2) {
#pragma unroll
					for (int r = 0; r < 8; r++)
#pragma unroll
						for (int c = 0; c < 8; c++) Dpark[(p & 1) * 64 + r * 8 + c] = a[r][c];
				}
			}
		} else if (active && bi == p) {
			// rows J are final: from VR (left of the panel) or W_D (the diagonal sub-block)
			const double *src = (bj < p) ? (VR + bj * 10) : nullptr;
#pragma unroll
			for (int k = 0; k < 8; k++)
#pragma unroll
				for (int c = 0; c < 8; c++) a[k][c] = src ? src[k * POTF2_VS + c] : WD[k * 8 + c];
		}
		// no barrier here: the next panel works on the other halves of the double buffers
	}
	__syncthreads();
#ifdef EMUB_POTF2_TIMING
	if (b == 0)
		for (int idx = tid; idx < 16 * 8 * 2; idx += POTF2_THREADS) logdet_parts[128 + idx] = (double)((&s_clk[0][0][0])[idx] - s_clk[0][0][0]);
#endif
	double logsum = 0.0;
	if (tid < 32) {
		// sum_j log L_jj = 0.5 sum_j log a_jj, fixed order (a failed pivot counts as 1, the point is flagged)
		double s = log(pivots[tid]) + log(pivots[tid + 32]) + log(pivots[tid + 64]) + log(pivots[tid + 96]);
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
		logsum = 0.5 * s;
	}
	if (tid == 0) {
		logdet_parts[(size_t)b * nblk + kblk] = logsum;
		if (s_bad) info[b] = 1;
	}
	if (active) {
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const int gi = bi * 8 + r;
			double *dst = Wg + (size_t)gi * ld + bj * 8;
#pragma unroll
			for (int c = 0; c < 8; c += 2) {
				double2 v;
				v.x = (bj * 8 + c <= gi) ? a[r][c] : 0.0;
				v.y = (bj * 8 + c + 1 <= gi) ? a[r][c + 1] : 0.0;
				*reinterpret_cast<double2 *>(dst + c) = v;
			}
		}
	}
	for (int idx = tid; idx < TB * TB / 2; idx += POTF2_THREADS) {
		const int i = idx >> 6, c = (idx & 63) * 2;
		if ((c >> 3) <= (i >> 3)) {
			double2 v;
			v.x = (c <= i) ? Lp[i * (i + 1) / 2 + c] : 0.0;
			v.y = (c + 1 <= i) ? Lp[i * (i + 1) / 2 + c + 1] : 0.0;
			*reinterpret_cast<double2 *>(Lg + (size_t)i * ld + c) = v;
		}
	}
}

}  // namespace emub
