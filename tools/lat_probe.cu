// lat_probe.cu -- latencies that bound the register-resident POTF2 column step on B200:
// dependent DFMA, rsqrt(double), 1/sqrt chain, LDS round trip, __syncthreads with 5 warps.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

__global__ void k_lat(double *out, long long *cyc, double seed)
{
	__shared__ double sm[256];
	const int N = 1024;
	double x = seed + threadIdx.x * 1e-3, y = seed * 0.5;
	long long t0, t1;
	// dependent DFMA
	t0 = clock64();
#pragma unroll 16
	for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
	t1 = clock64();
	if (threadIdx.x == 0) cyc[0] = (t1 - t0);
	// independent DFMA x8
	double a[8];
	for (int k = 0; k < 8; k++) a[k] = x + k;
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++)
#pragma unroll
		for (int k = 0; k < 8; k++) a[k] = fma(a[k], y, 1e-9);
	t1 = clock64();
	if (threadIdx.x == 0) cyc[1] = (t1 - t0);
	for (int k = 0; k < 8; k++) x += a[k];
	// rsqrt chain
	x = fabs(x) + 1.0;
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++) x = rsqrt(x) + 1.0;
	t1 = clock64();
	if (threadIdx.x == 0) cyc[2] = (t1 - t0);
	// sqrt + div chain
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++) x = 1.0 / sqrt(x) + 1.0;
	t1 = clock64();
	if (threadIdx.x == 0) cyc[3] = (t1 - t0);
	// LDS round trip (store, load dependent)
	sm[threadIdx.x] = x;
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++) { sm[threadIdx.x] = x; x = sm[(threadIdx.x + 1) & 255] + 1e-9; }
	t1 = clock64();
	if (threadIdx.x == 0) cyc[4] = (t1 - t0);
	// syncthreads
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++) { __syncthreads(); }
	t1 = clock64();
	if (threadIdx.x == 0) cyc[5] = (t1 - t0);
	// publish -> barrier -> load chain
	t0 = clock64();
#pragma unroll 4
	for (int i = 0; i < N; i++) { sm[threadIdx.x] = x; __syncthreads(); x = sm[(threadIdx.x + 7) & 127] * 1.0000001; }
	t1 = clock64();
	if (threadIdx.x == 0) cyc[6] = (t1 - t0);
	out[threadIdx.x] = x;
}

int main()
{
	double *d_out; long long *d_c, h[8];
	CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_c, 64));
	for (int threads : {32, 160, 256}) {
		k_lat<<<1, threads>>>(d_out, d_c, 1.0001);
		CK(cudaDeviceSynchronize());
		CK(cudaMemcpy(h, d_c, 56, cudaMemcpyDeviceToHost));
		printf("threads=%3d cycles/iter: dep DFMA %.1f | 8 indep DFMA %.1f | rsqrt chain %.1f | 1/sqrt chain %.1f | STS+LDS %.1f | syncthreads %.1f | publish-barrier-load %.1f\n",
		       threads, h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0, h[6] / 1024.0);
	}
	return 0;
}
