// ubench_ffma.cu - FP32 issue-rate microbenchmark for sm_100a: scalar FFMA vs packed FFMA2
// (fma.rn.f32x2).  Prints lane-FMAs per clock per SM for both.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float r0, int iters)
{
	float2 c[8];
	float2 w[8];
#pragma unroll
	for (int i = 0; i < 8; i++) {
		c[i] = make_float2(threadIdx.x * 1e-3f, i);
		w[i] = make_float2(1.0f + i * 1e-3f, 1.0f - i * 1e-3f);
	}
	float r = r0;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
#pragma unroll
			for (int i = 0; i < 8; i++) {
				if (MODE == 0) {
					c[i].x = fmaf(r, w[(i + u) & 7].x, c[i].x);
					c[i].y = fmaf(r, w[(i + u) & 7].y, c[i].y);
				} else {
					unsigned long long cc = *reinterpret_cast<unsigned long long *>(&c[i]);
					const float2 rr = make_float2(r, r);
					asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc)
					    : "l"(*reinterpret_cast<const unsigned long long *>(&rr)),
					      "l"(*reinterpret_cast<const unsigned long long *>(&w[(i + u) & 7])));
					c[i] = *reinterpret_cast<float2 *>(&cc);
				}
			}
		}
		r = -r;
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; i++)
		s += c[i].x + c[i].y;
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount, ctas = sms * 8, iters = 4096;
	float *out;
	cudaMalloc(&out, sizeof(float) * ctas * 256);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	for (int mode = 0; mode < 2; mode++) {
		cudaEvent_t a, b;
		cudaEventCreate(&a); cudaEventCreate(&b);
		for (int rep = 0; rep < 3; rep++) {
			cudaEventRecord(a);
			if (mode == 0) k<0><<<ctas, 256>>>(out, 1e-6f, iters); else k<1><<<ctas, 256>>>(out, 1e-6f, iters);
			cudaEventRecord(b);
			cudaEventSynchronize(b);
		}
		float ms;
		cudaEventElapsedTime(&ms, a, b);
		const double fma = (double)ctas * 256 * iters * 128.0;
		printf("%s: %.3f ms, %.1f lane-FMA/clk/SM at %d MHz (max clock)\n", mode ? "FFMA2" : "FFMA ", ms,
		       fma / (ms * 1e-3) / (khz * 1e3) / sms, khz / 1000);
	}
	return 0;
}
