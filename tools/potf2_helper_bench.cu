// potf2_helper_bench.cu -- latency of step A of k_potf2 (potf2_diag_block: 8 x 8 Cholesky + inverse by one warp) alone.
#include <cstdio>
#include <cuda_runtime.h>
#include "../madaiemulator_b200/csrc/emub_kernels.cuh"
using namespace emub;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__global__ void k_helper(double *out, long long *cyc, int reps, int loaders)
{
	__shared__ double Lp[POTF2_LPACK];
	__shared__ double LP[TB * POTF2_LS], VR[8 * POTF2_VS], WD[64], Dpark[64], Dwork[128], pivots[TB];
	__shared__ int s_bad;
	__shared__ volatile int stop;
	const int tid = threadIdx.x;
	if (tid == 0) { s_bad = 0; stop = 0; }
	for (int i = tid; i < TB * POTF2_LS; i += blockDim.x) LP[i] = 0.01 * ((i * 7) % 13);
	for (int i = tid; i < 64; i += blockDim.x) { const int r = i >> 3, c = i & 7; Dpark[i] = (r == c ? 4.0 : 0.0) + 0.1 * (r + c); }
	__syncthreads();
	if (tid < 32) {
		long long t0 = clock64();
		for (int it = 0; it < reps; it++) {
			potf2_diag_block(1, tid, Dpark, LP, Dwork, WD, VR, Lp, pivots, &s_bad);
			__syncwarp();
		}
		long long t1 = clock64();
		if (tid == 0) { cyc[0] = (t1 - t0) / reps; stop = 1; }
		out[tid] = WD[tid] + WD[tid + 32];
	} else if (tid < 32 * (1 + loaders)) {
		// streaming DFMA warps, like the sub-block owners during step C
		double a[16];
		for (int k = 0; k < 16; k++) a[k] = 1.0 + k + tid;
		while (!stop) {
#pragma unroll
			for (int r = 0; r < 8; r++)
#pragma unroll
				for (int k = 0; k < 16; k++) a[k] = fma(a[k], 0.999, 1e-9);
		}
		double s = 0;
		for (int k = 0; k < 16; k++) s += a[k];
		out[tid] = s;
	}
}

int main()
{
	double *out; long long *cyc;
	CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&cyc, 8));
	for (int loaders : {0, 3, 4, 5}) {
		k_helper<<<1, 192>>>(out, cyc, 200, loaders);
		CK(cudaDeviceSynchronize());
		long long h;
		CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
		double w[64];
		CK(cudaMemcpy(w, out, 32 * 8, cudaMemcpyDeviceToHost));
		printf("potf2_diag_block with %d streaming warps beside it: %lld cycles per call (checksum %.6f)\n", loaders, h, w[0] + w[9]);
	}
	return 0;
}
