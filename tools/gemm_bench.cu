// gemm_bench.cu -- tuning harness for the FP64 DMMA tile engine (madaiemulator_b200/csrc/emub_gemm.cuh):
// times tile-engine configurations on full-square task lists against cuBLAS DGEMM, and checks one tile.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include "../madaiemulator_b200/csrc/emub_gemm.cuh"
using namespace emub;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int AL, int BL, class Cfg>
static void run_cfg(const char *name, int n, int K, const double *A, const double *B, double *C, const double *Cref)
{
	if (Cfg::SMEM_BYTES > 227 * 1024) { printf("%-28s smem %d too large\n", name, Cfg::SMEM_BYTES); return; }
	const int nb = n / TB;
	std::vector<GemmTask> tasks;
	for (int i = 0; i < nb; i++)
		for (int j = 0; j < nb; j++) {
			long long a_off = (AL == KMAJOR) ? (long long)i * TB * n : (long long)i * TB;
			long long b_off = (BL == KMAJOR) ? (long long)j * TB * n : (long long)j * TB;
			tasks.push_back({a_off, b_off, (long long)i * TB * n + (long long)j * TB, K, 0});
		}
	GemmTask *dT;
	CK(cudaMalloc(&dT, tasks.size() * sizeof(GemmTask)));
	CK(cudaMemcpy(dT, tasks.data(), tasks.size() * sizeof(GemmTask), cudaMemcpyHostToDevice));
	CK(cudaFuncSetAttribute(k_gemm<AL, BL, EPI_STORE, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
	GemmArgs a{dT, A, B, C, 0, 0, 0, n, n, n, 1.0};
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	k_gemm<AL, BL, EPI_STORE, Cfg><<<dim3((unsigned)tasks.size() * Cfg::SUBS, 1), Cfg::THREADS, Cfg::SMEM_BYTES>>>(a);
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		CK(cudaEventRecord(e0));
		k_gemm<AL, BL, EPI_STORE, Cfg><<<dim3((unsigned)tasks.size() * Cfg::SUBS, 1), Cfg::THREADS, Cfg::SMEM_BYTES>>>(a);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	double err = -1;
	if (Cref) {
		std::vector<double> h(1024), hr(1024);
		CK(cudaMemcpy(h.data(), C + 5 * n + 128, 1024 * 8, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(hr.data(), Cref + 5 * n + 128, 1024 * 8, cudaMemcpyDeviceToHost));
		err = 0;
		for (int i = 0; i < 1024; i++) err = fmax(err, fabs(h[i] - hr[i]));
	}
	printf("%-28s n=%d K=%d threads=%d smem=%dKB : %8.3f ms %7.2f TFLOP/s  maxerr=%.2e\n", name, n, K, Cfg::THREADS, Cfg::SMEM_BYTES / 1024,
	       best, 2.0 * n * n * K / best * 1e-9, err);
	cudaFree(dT);
}

int main()
{
	const int n = 4096;
	size_t bytes = (size_t)n * n * 8;
	double *A, *B, *C, *Cref;
	CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes)); CK(cudaMalloc(&Cref, bytes));
	std::vector<double> h((size_t)n * n);
	for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
	CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice));
	for (size_t i = 0; i < h.size(); i++) h[i] = (double)((i * 40503u + 17) % 1000) / 1000.0 - 0.5;
	CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
	cublasHandle_t hd; cublasCreate(&hd);
	double one = 1, zero = 0;
	cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int K : {4096, 512}) {
		// row-major C = A B^T (A, B row-major n x K with ld n)  ==  col-major C^T = B A^T : op(B)=T? use gemm(T, N): C^T(n x n) = (B^T)^T ...
		// col-major view: A_rm is (K x n) col-major with ld n.  C_rm^T = B_rm A_rm^T -> cublas: C_cm = op(Bcm)^T ... C_cm(j,i) = sum_k Bcm(k,j) Acm(k,i) = gemm(T, N, n, n, K, Bcm, Acm)
		for (int w = 0; w < 2; w++) cublasDgemm(hd, CUBLAS_OP_T, CUBLAS_OP_N, n, n, K, &one, B, n, A, n, &zero, Cref, n);
		CK(cudaDeviceSynchronize());
		float best = 1e30f;
		for (int r = 0; r < 3; r++) {
			CK(cudaEventRecord(e0));
			cublasDgemm(hd, CUBLAS_OP_T, CUBLAS_OP_N, n, n, K, &one, B, n, A, n, &zero, Cref, n);
			CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
			if (ms < best) best = ms;
		}
		printf("cublasDgemm (NT rm)          n=%d K=%d : %8.3f ms %7.2f TFLOP/s\n", n, K, best, 2.0 * n * n * K / best * 1e-9);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<128, 128, 2, 4, 16, 4, 1>>("NT 128x128 2x4 bk16 s4 x1", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 64, 2, 2, 16, 3, 3>>("NT 64x64 2x2 bk16 s3 x3", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 64, 2, 2, 16, 4, 2>>("NT 64x64 2x2 bk16 s4 x2", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 64, 2, 2, 8, 4, 4>>("NT 64x64 2x2 bk8 s4 x4", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 64, 2, 2, 8, 5, 3>>("NT 64x64 2x2 bk8 s5 x3", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 64, 1, 4, 16, 3, 3>>("NT 64x64 1x4 bk16 s3 x3", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<128, 64, 4, 2, 16, 3, 2>>("NT 128x64 4x2 bk16 s3 x2", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<64, 128, 2, 4, 16, 3, 2>>("NT 64x128 2x4 bk16 s3 x2", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<128, 64, 2, 2, 16, 3, 2>>("NT 128x64 2x2(64x32) bk16 s3 x2", n, K, A, B, C, Cref);
		run_cfg<KMAJOR, KMAJOR, GemmCfg<128, 128, 4, 4, 16, 3, 1>>("NT 128x128 4x4 bk16 s3 x1", n, K, A, B, C, Cref);
		run_cfg<RMAJOR, RMAJOR, GemmCfg<64, 64, 2, 2, 16, 3, 3>>("TN 64x64 2x2 bk16 s3 x3", n, K, A, B, C, nullptr);
		run_cfg<KMAJOR, RMAJOR, GemmCfg<64, 64, 2, 2, 16, 3, 3>>("NN 64x64 2x2 bk16 s3 x3", n, K, A, B, C, nullptr);
		run_cfg<RMAJOR, RMAJOR, GemmCfg<128, 64, 4, 2, 16, 3, 2>>("TN 128x64 4x2 bk16 s3 x2", n, K, A, B, C, nullptr);
		run_cfg<KMAJOR, RMAJOR, GemmCfg<128, 64, 4, 2, 16, 3, 2>>("NN 128x64 4x2 bk16 s3 x2", n, K, A, B, C, nullptr);
	}
	return 0;
}
