// dmma_probe.cu -- microbenchmark: FP64 DMMA (mma.sync f64) shapes vs DFMA vs cuBLAS DGEMM on B200.
// Output feeds DESIGN.md's roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double seed)
{
	double c[NACC][4];
	for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = seed * (i + j);
	double a[8], b[4];
	for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-9 + i;
	for (int i = 0; i < 4; i++) b[i] = seed - threadIdx.x * 1e-9 + i;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int i = 0; i < NACC; i++) {
			if (SHAPE == 0) {
				// m8n8k4: c has 2 regs
				asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
				             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
			} else if (SHAPE == 1) {
				asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
				             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
			} else if (SHAPE == 2) {
				asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
				             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
				             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
			} else if (SHAPE == 3) {
				asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
				             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
				             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
				               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
			} else {
				// DFMA: 4 independent per acc
				c[i][0] = fma(a[0], b[0], c[i][0]); c[i][1] = fma(a[1], b[1], c[i][1]);
				c[i][2] = fma(a[2], b[2], c[i][2]); c[i][3] = fma(a[3], b[3], c[i][3]);
			}
		}
	}
	double s = 0;
	for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
static void run(const char *name, double flop_per_warp_instr, int warps_per_cta, int ctas_per_sm, double *d_out)
{
	int nsm = 148;
	int iters = 4096;
	dim3 grid(nsm * ctas_per_sm), block(warps_per_cta * 32);
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	k_dmma<SHAPE, NACC><<<grid, block>>>(d_out, 64, 1.0);
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		CK(cudaEventRecord(e0));
		k_dmma<SHAPE, NACC><<<grid, block>>>(d_out, iters, 1.0);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
		if (ms < best) best = ms;
	}
	double total = (double)grid.x * warps_per_cta * iters * NACC * flop_per_warp_instr;
	printf("%-12s warps/cta=%2d ctas/sm=%d nacc=%d : %8.3f ms  %8.2f TFLOP/s\n", name, warps_per_cta, ctas_per_sm, NACC, best,
	       total / best * 1e-9);
}

int main()
{
	double *d_out;
	CK(cudaMalloc(&d_out, 148 * 8 * 1024 * sizeof(double)));
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
	for (int w = 4; w <= 16; w *= 2) {
		run<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, w, 1, d_out);
		run<1, 8>("m16n8k4", 2.0 * 16 * 8 * 4, w, 1, d_out);
		run<2, 8>("m16n8k8", 2.0 * 16 * 8 * 8, w, 1, d_out);
		run<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, w, 1, d_out);
		run<4, 8>("dfma", 2.0 * 32 * 4, w, 1, d_out);
	}
	run<1, 2>("m16n8k4", 2.0 * 16 * 8 * 4, 8, 1, d_out);
	run<3, 2>("m16n8k16", 2.0 * 16 * 8 * 16, 8, 1, d_out);
	run<3, 1>("m16n8k16", 2.0 * 16 * 8 * 16, 4, 1, d_out);

	// cuBLAS DGEMM peak
	cublasHandle_t h; cublasCreate(&h);
	for (int n : {4096, 8192}) {
		double *A, *B, *C;
		size_t bytes = (size_t)n * n * sizeof(double);
		CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
		CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
		double one = 1.0, zero = 0.0;
		cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
		for (int op = 0; op < 2; op++) {
			cublasOperation_t tb = op ? CUBLAS_OP_T : CUBLAS_OP_N;
			for (int w = 0; w < 2; w++) cublasDgemm(h, CUBLAS_OP_N, tb, n, n, n, &one, A, n, B, n, &zero, C, n);
			CK(cudaDeviceSynchronize());
			float best = 1e30f;
			for (int r = 0; r < 5; r++) {
				CK(cudaEventRecord(e0));
				cublasDgemm(h, CUBLAS_OP_N, tb, n, n, n, &one, A, n, B, n, &zero, C, n);
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
				if (ms < best) best = ms;
			}
			printf("cublasDgemm N%c n=%d burst: %.3f ms %.2f TFLOP/s\n", op ? 'T' : 'N', n, best, 2.0 * n * n * n / best * 1e-9);
			// sustained 3 s
			CK(cudaEventRecord(e0));
			int cnt = 0; float ms = 0;
			while (ms < 3000.f) {
				for (int r = 0; r < 10; r++) cublasDgemm(h, CUBLAS_OP_N, tb, n, n, n, &one, A, n, B, n, &zero, C, n);
				cnt += 10;
				CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
				CK(cudaEventElapsedTime(&ms, e0, e1));
			}
			printf("cublasDgemm N%c n=%d sustained: %d calls %.1f ms %.2f TFLOP/s\n", op ? 'T' : 'N', n, cnt, ms, 2.0 * n * n * n * cnt / ms * 1e-9);
		}
		// DSYRK & DPOTRF-ish reference: dsyrk n x 128
		cudaFree(A); cudaFree(B); cudaFree(C);
	}
	return 0;
}
