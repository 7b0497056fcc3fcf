// tools/ubench_acs.cu - A/B of the Viterbi add-compare-select inner loop in isolation: the
// shipped form (one codeword per thread, 32-bit metrics, decision = sign of b - a shifted in with a funnel shift) against
// two codewords per thread with 16-bit metrics packed in one register (VIADD.16x2 / VIMNMX.U16x2 with predicate outputs).
// Same trellis (K5 r1/2, 212 steps x 16 states), decisions stored step-major to a global scratch array exactly as the
// product kernel does, the same shared-memory footprint per CTA as the product kernel (54 KB of soft-bit rows per 128
// codewords -> 4 CTAs per SM), >= 1 M codewords (several full waves).  The soft-bit gather and the traceback are left
// out on both sides: they cost the same per codeword in either form, so this is the most the packed form can win.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_acs tools/ubench_acs.cu && tools/ubench_acs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

static constexpr int STEPS = 212, NS = 16, H = 8;
__host__ __device__ constexpr unsigned parity(unsigned x) { x ^= x >> 4; x ^= x >> 2; x ^= x >> 1; return x & 1u; }
__host__ __device__ constexpr unsigned outp(unsigned s, unsigned b)
{
	const unsigned reg = (s << 1) | b;
	return (parity(reg & 0x19) << 1) | parity(reg & 0x17);
}

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; return x; }

// A: one codeword per thread, 32-bit metrics
__global__ void __launch_bounds__(128) acs32(uint16_t *dec, uint32_t *out, int n)
{
	extern __shared__ uint8_t smem[];
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n) return;
	uint32_t ae[NS], na[NS];
#pragma unroll
	for (int s = 0; s < NS; s++) ae[s] = s ? 0xffffffu : 0u;
	uint32_t seed = hash(t + 1);
	for (int i = 0; i < STEPS; i += 2) {
#pragma unroll
		for (int half = 0; half < 2; half++) {
			seed = seed * 1664525u + 1013904223u;
			const uint32_t m0 = (seed >> 8) & 127, m1 = (seed >> 16) & 127, d0 = (seed >> 3) & 63, d1 = (seed >> 24) & 63;
			const uint32_t bm[4] = {m0 + m1, m0 + m1 + d1, m0 + d0 + m1, m0 + d0 + m1 + d1};
			const uint32_t (&src)[NS] = half ? na : ae;
			uint32_t (&dst)[NS] = half ? ae : na;
			uint32_t acc_hi = 0, acc_lo = 0;
#pragma unroll
			for (int s = NS - 1; s >= 0; s--) {
				const int k = s >> 1, bit = s & 1;
				const uint32_t a = src[k] + bm[outp(k, bit)], b = src[k + H] + bm[outp(k + H, bit)];
				const uint32_t diff = b - a;
				dst[s] = b < a ? b : a;
				if (s >= 8) acc_hi = __funnelshift_l(diff, acc_hi, 1);
				else        acc_lo = __funnelshift_l(diff, acc_lo, 1);
			}
			dec[(size_t)(i + half) * n + t] = (uint16_t)((acc_hi << 8) | acc_lo);
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int s = 0; s < NS; s++) r ^= ae[s];
	out[t] = r + smem[threadIdx.x & 15];
}

// B: two codewords per thread, 16-bit metrics packed (low half: codeword t, high half: codeword t + n/2)
__global__ void __launch_bounds__(64) acs16x2(uint32_t *dec, uint32_t *out, int n2)
{
	extern __shared__ uint8_t smem[];
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n2) return;
	uint32_t ae[NS], na[NS];
#pragma unroll
	for (int s = 0; s < NS; s++) ae[s] = s ? 0x40004000u : 0u;
	uint32_t seed = hash(t + 1), seed2 = hash(t + 77777);
	for (int i = 0; i < STEPS; i += 2) {
#pragma unroll
		for (int half = 0; half < 2; half++) {
			seed = seed * 1664525u + 1013904223u;
			seed2 = seed2 * 1664525u + 1013904223u;
			// packed branch metrics of the two codewords
			const uint32_t m0 = ((seed >> 8) & 127) | (((seed2 >> 8) & 127) << 16), m1 = ((seed >> 16) & 127) | (((seed2 >> 16) & 127) << 16);
			const uint32_t d0 = ((seed >> 3) & 63) | (((seed2 >> 3) & 63) << 16), d1 = ((seed >> 24) & 63) | (((seed2 >> 24) & 63) << 16);
			const uint32_t b00 = __vadd2(m0, m1);
			const uint32_t bm[4] = {b00, __vadd2(b00, d1), __vadd2(b00, d0), __vadd2(__vadd2(b00, d0), d1)};
			const uint32_t (&src)[NS] = half ? na : ae;
			uint32_t (&dst)[NS] = half ? ae : na;
			uint32_t acc_lo = 0, acc_hi = 0;       // per 8 states: bit s of the low half = codeword A, of the high half = B
#pragma unroll
			for (int s = NS - 1; s >= 0; s--) {
				const int k = s >> 1, bit = s & 1;
				const uint32_t a = __vadd2(src[k], bm[outp(k, bit)]), b = __vadd2(src[k + H], bm[outp(k + H, bit)]);
				bool ph, pl;
				dst[s] = __vibmin_u16x2(a, b, &ph, &pl);      // pred = (a <= b): decision "b < a" is its negation
				const uint32_t d = (ph ? 0u : 0x10000u) | (pl ? 0u : 1u);
				if (s >= 8) acc_hi = (acc_hi << 1) | d;
				else        acc_lo = (acc_lo << 1) | d;
			}
			dec[(size_t)(i + half) * n2 + t] = (acc_hi << 8) | acc_lo;
		}
		if ((i & 63) == 62) {                      // 16-bit metrics: renormalise now and then (the product would not need to
			uint32_t mn = ae[0];                   // for 212 steps of rate 1/2; here the synthetic metrics are larger)
#pragma unroll
			for (int s = 1; s < NS; s++) mn = __vminu2(mn, ae[s]);
#pragma unroll
			for (int s = 0; s < NS; s++) ae[s] = __vsub2(ae[s], mn);
		}
	}
	uint32_t r = 0;
#pragma unroll
	for (int s = 0; s < NS; s++) r ^= ae[s];
	out[t] = r + smem[threadIdx.x & 15];
}

int main()
{
	const int n = 1 << 20;                         // codewords
	uint8_t *dec;
	uint32_t *out;
	cudaMalloc(&dec, (size_t)STEPS * n * 2);
	cudaMalloc(&out, (size_t)n * 4);
	const int smem = 54 * 1024 + 272;              // what 128 codewords of soft-bit rows take in the product kernel
	cudaFuncSetAttribute(acs32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaFuncSetAttribute(acs16x2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int variant = 0; variant < 2; variant++) {
		float best = 1e9f;
		for (int rep = 0; rep < 5; rep++) {
			cudaEventRecord(e0);
			if (variant == 0)
				acs32<<<n / 128, 128, smem>>>((uint16_t *)dec, out, n);
			else
				acs16x2<<<n / 128, 64, smem>>>((uint32_t *)dec, out, n / 2);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			if (rep && ms < best) best = ms;
		}
		printf("{\"variant\": \"%s\", \"codewords\": %d, \"ms\": %.4f, \"acs_per_s\": %.4g, \"err\": \"%s\"}\n",
		       variant ? "16-bit x2 packed (VIADD.16x2 / VIMNMX.U16x2), 64 threads per CTA" : "32-bit, funnel-shift decisions, 128 threads per CTA",
		       n, best, (double)n * STEPS * NS / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
