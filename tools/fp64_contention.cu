// fp64_contention.cu -- dependent-DFMA latency of one warp while other warps of the CTA stream independent DFMAs:
// which warps (same scheduler = warp id mod 4, or the other three) slow the chain down, and by how much.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__global__ void k_contend(double *out, long long *cyc, unsigned loadmask, double seed)
{
	const int w = threadIdx.x >> 5;
	__shared__ volatile int stop;
	if (threadIdx.x == 0) stop = 0;
	__syncthreads();
	if (w == 0) {
		double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
		long long t0 = clock64();
#pragma unroll 16
		for (int i = 0; i < 4096; i++) x = fma(x, y, 1e-9);
		long long t1 = clock64();
		if (threadIdx.x == 0) { cyc[0] = t1 - t0; stop = 1; }
		out[threadIdx.x] = x;
	} else if ((loadmask >> w) & 1) {
		double a[16];
		for (int k = 0; k < 16; k++) a[k] = seed + k + threadIdx.x;
		const double y = seed * 0.25;
		long long n = 0;
		while (!stop) {
#pragma unroll
			for (int r = 0; r < 8; r++)
#pragma unroll
				for (int k = 0; k < 16; k++) a[k] = fma(a[k], y, 1e-9);
			n += 128;
		}
		double s = 0;
		for (int k = 0; k < 16; k++) s += a[k];
		out[threadIdx.x] = s;
		if ((threadIdx.x & 31) == 0) cyc[w] = n;
	}
}

int main()
{
	double *out; long long *cyc;
	CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMalloc(&cyc, 64 * 8));
	struct { const char *name; unsigned mask; } cases[] = {
	    {"chain alone", 0u}, {"+ warp 4 (same scheduler)", 1u << 4}, {"+ warps 1,2,3 (other schedulers)", 0xEu},
	    {"+ warps 1..4", 0x1Eu}, {"+ warps 1..7", 0xFEu}, {"+ warps 4 (same) and 1", (1u << 4) | 2u}};
	for (auto &c : cases) {
		CK(cudaMemset(cyc, 0, 64 * 8));
		k_contend<<<1, 256>>>(out, cyc, c.mask, 1.000001);
		CK(cudaDeviceSynchronize());
		long long h[8];
		CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
		long long tot = 0;
		for (int w = 1; w < 8; w++) tot += h[w];
		printf("%-36s dependent DFMA %.1f cycles each | loaders issued %.2f warp-DFMA per cycle in total\n", c.name, h[0] / 4096.0, (double)tot / h[0]);
	}
	return 0;
}
