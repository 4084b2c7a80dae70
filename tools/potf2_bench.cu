// potf2_bench.cu -- times emub::k_potf2 alone (one 128 x 128 SPD block per CTA) and checks L L^T = A, W L = I.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../madaiemulator_b200/csrc/emub_kernels.cuh"
using namespace emub;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
int main(int argc, char **argv)
{
	const int n = 128, B = 8;
	std::vector<double> A((size_t)B * n * n);
	for (int b = 0; b < B; b++)
		for (int i = 0; i < n; i++)
			for (int j = 0; j < n; j++) {
				double d = (i - j) * 0.05;
				A[(size_t)b * n * n + i * n + j] = exp(-d * d) + (i == j ? 0.05 + 0.01 * b : 0.0);
			}
	double *dA, *dL, *dW, *dlog; int *dinfo;
	size_t bytes = A.size() * 8;
	CK(cudaMalloc(&dA, bytes)); CK(cudaMalloc(&dL, bytes)); CK(cudaMalloc(&dW, bytes)); CK(cudaMalloc(&dlog, (B + 1024) * 8)); CK(cudaMalloc(&dinfo, B * 4));
	CK(cudaMemcpy(dA, A.data(), bytes, cudaMemcpyHostToDevice));
	CK(cudaMemset(dinfo, 0, B * 4));
	CK(cudaFuncSetAttribute(k_potf2, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM_BYTES));
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	if (argc > 1) {  // profiling mode: 148 CTAs on the same block (stride 0), one launch
		k_potf2<<<148, POTF2_THREADS, POTF2_SMEM_BYTES>>>(dA, 0, dL, 0, dW, 0, n, 0, 1, dlog, dinfo);
		CK(cudaDeviceSynchronize());
		k_potf2<<<148, POTF2_THREADS, POTF2_SMEM_BYTES>>>(dA, 0, dL, 0, dW, 0, n, 0, 1, dlog, dinfo);
		CK(cudaDeviceSynchronize());
		printf("profiled launch done\n");
		return 0;
	}
	for (int nb : {1, 8}) {
		for (int w = 0; w < 3; w++) k_potf2<<<nb, POTF2_THREADS, POTF2_SMEM_BYTES>>>(dA, n * n, dL, n * n, dW, n * n, n, 0, 1, dlog, dinfo);
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0));
		for (int r = 0; r < 50; r++) k_potf2<<<nb, POTF2_THREADS, POTF2_SMEM_BYTES>>>(dA, n * n, dL, n * n, dW, n * n, n, 0, 1, dlog, dinfo);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		printf("k_potf2 x%d CTAs: %.2f us per launch (50 back-to-back launches)\n", nb, ms * 1000 / 50);
	}
#ifdef EMUB_POTF2_TIMING
	{
		const int NW = POTF2_THREADS / 32;
		std::vector<double> clk(16 * 8 * 2);
		CK(cudaMemcpy(clk.data(), dlog + 128, clk.size() * 8, cudaMemcpyDeviceToHost));
		printf("cycles since the first stamp, per warp: arrival at barrier 1 (previous C / A done) | arrival at barrier 2 (B done)\n");
		for (int p = 0; p < 16; p++) {
			printf("panel %2d:", p);
			for (int w = 0; w < NW; w++) printf("  w%d %6.0f %6.0f", w, clk[(p * 8 + w) * 2], clk[(p * 8 + w) * 2 + 1]);
			printf("\n");
		}
	}
#endif
	std::vector<double> L(A.size()), W(A.size());
	CK(cudaMemcpy(L.data(), dL, bytes, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(W.data(), dW, bytes, cudaMemcpyDeviceToHost));
	double e1m = 0, e2m = 0;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) {
			double s = 0, t = 0;
			for (int k = 0; k < n; k++) { s += L[i * n + k] * L[j * n + k]; t += W[i * n + k] * L[k * n + j]; }
			e1m = fmax(e1m, fabs(s - A[i * n + j]));
			e2m = fmax(e2m, fabs(t - (i == j ? 1.0 : 0.0)));
		}
	printf("max |L L^T - A| = %.2e   max |W L - I| = %.2e\n", e1m, e2m);
	return 0;
}
