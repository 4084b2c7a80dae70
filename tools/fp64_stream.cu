// fp64_stream.cu -- how many warps per scheduler does it take to saturate the FP64 pipe with INDEPENDENT DFMAs of the
// k_potf2 step-C shape (64 accumulators per thread, a[r][c] -= l[r] * v[c], operands in registers)?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int NACC>
__global__ void k_stream(double *out, long long *cyc, double seed, int iters)
{
	double a[NACC], l[8], v[8];
	for (int k = 0; k < NACC; k++) a[k] = seed + k + threadIdx.x;
	for (int k = 0; k < 8; k++) { l[k] = seed * (1 + k) * 1e-3; v[k] = seed * (2 + k) * 1e-3; }
	__syncthreads();
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 8; r++)
#pragma unroll
			for (int c = 0; c < NACC / 8; c++) a[r * (NACC / 8) + c] -= l[r] * v[c];
	}
	long long t1 = clock64();
	double s = 0;
	for (int k = 0; k < NACC; k++) s += a[k];
	out[threadIdx.x] = s;
	if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

int main()
{
	double *out; long long *cyc;
	CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&cyc, 64 * 8));
	const int iters = 2000;
	for (int warps : {1, 4, 8, 12, 16}) {
		for (int nacc : {64, 32}) {
			if (nacc == 64) k_stream<64><<<1, warps * 32>>>(out, cyc, 1.000001, iters);
			else k_stream<32><<<1, warps * 32>>>(out, cyc, 1.000001, iters);
			CK(cudaDeviceSynchronize());
			long long h[32];
			CK(cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost));
			long long mx = 0;
			for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
			const double dfma_per_warp = (double)iters * nacc;
			printf("%2d warps (%.1f per scheduler), %2d accumulators: %.2f cycles per warp-DFMA per warp, %.2f cycles per DFMA per scheduler\n",
			       warps, warps / 4.0, nacc, mx / dfma_per_warp, mx / (dfma_per_warp * (warps < 4 ? 1 : warps / 4.0)));
		}
	}
	return 0;
}
