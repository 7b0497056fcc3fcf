// tools/ubench/fma_rate.cu - issue rate of FFMA (3-register), FFMA2 (packed, 64-bit registers) and FADD2 on sm_100a:
// what one scheduler sustains per cycle decides whether the FCCH correlation (one packed FFMA2 per tap and output) sits
// at the fp32 pipe's limit.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, int iters, float s)
{
	float2 a[8];
	for (int i = 0; i < 8; i++)
		a[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
	float2 b = make_float2(s, s * 0.5f), c = make_float2(s * 0.25f, -s);
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int r = 0; r < 4; r++)
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) {
					a[i].x = fmaf(a[i].x, b.x, c.x);
					a[i].y = fmaf(a[i].y, b.y, c.y);
				} else if (MODE == 1) {
					unsigned long long &aa = *reinterpret_cast<unsigned long long *>(&a[i]);
					asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(aa) : "l"(*reinterpret_cast<unsigned long long *>(&b)),
					             "l"(*reinterpret_cast<unsigned long long *>(&c)));
				} else {
					unsigned long long &aa = *reinterpret_cast<unsigned long long *>(&a[i]);
					asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(aa) : "l"(*reinterpret_cast<unsigned long long *>(&b)));
				}
			}
	}
	float r = 0;
	for (int i = 0; i < 8; i++)
		r += a[i].x + a[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount, blocks = sms * 8, iters = 4096;
	float *out;
	cudaMalloc(&out, blocks * 256 * sizeof(float));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const char *name[3] = {"FFMA x2 (scalar)", "FFMA2 (packed)", "FADD2 (packed)"};
	for (int mode = 0; mode < 3; mode++) {
		for (int rep = 0; rep < 2; rep++) {
			cudaEventRecord(e0);
			if (mode == 0) k<0><<<blocks, 256>>>(out, iters, 1.0001f);
			if (mode == 1) k<1><<<blocks, 256>>>(out, iters, 1.0001f);
			if (mode == 2) k<2><<<blocks, 256>>>(out, iters, 1.0001f);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
		}
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		const double lane_ops = (double)blocks * 256 * iters * 32 * 2;       // scalar fp32 operations (FMA = 1) per launch
		const double clk = p.clockRate * 1e3;
		printf("%-18s %.3f ms  %.1f G lane-op/s  %.1f lane-op/clk/SM  (%.2f TFLOP/s as FMA)\n", name[mode], ms, lane_ops / ms / 1e6,
		       lane_ops / (ms * 1e-3) / clk / sms, 2 * lane_ops / ms / 1e9);
	}
	return 0;
}
