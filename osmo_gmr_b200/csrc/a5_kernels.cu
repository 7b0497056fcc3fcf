// a5_kernels.cu - GMR-1 A5/1 keystream for a whole batch: one thread per (Kc, frame number).
// Replaces, per launch, n calls of gmr1_a5 / gmr1_a5_1 (reference src/l1/a5.c:57-282): key setup
// (byte-swapped pairs, frame number folded into the key, 64 clock-all steps, bit 0 of R1..R4 forced,
// 250 warm-up steps under the majority clocking rule of R4), then nbits downlink and nbits uplink
// bits.  Integer work, bit-exact.  The streams are the sign masks of the ciphered channel decoders
// (TCH3 208 bits, FACCH3 4 x 96, FACCH9 / TCH9 658), so producing them on the device keeps the
// ciphered path free of per-burst host work and of the H2D copy of the masks.
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.h"

namespace gmr1 {

namespace {

struct A51 {
	uint32_t r0, r1, r2, r3;
};

__device__ __forceinline__ uint32_t par(uint32_t x) { return (uint32_t)__popc(x) & 1u; }
__device__ __forceinline__ uint32_t step19(uint32_t v) { return ((v << 1) & 0x07ffffu) | par(v & 0x072000u); }
__device__ __forceinline__ uint32_t step22(uint32_t v) { return ((v << 1) & 0x3fffffu) | par(v & 0x311000u); }
__device__ __forceinline__ uint32_t step23(uint32_t v) { return ((v << 1) & 0x7fffffu) | par(v & 0x660000u); }
__device__ __forceinline__ uint32_t step17(uint32_t v) { return ((v << 1) & 0x01ffffu) | par(v & 0x013100u); }

__device__ __forceinline__ void clock_rule(A51 &s)
{
	const uint32_t c0 = (s.r3 >> 15) & 1u, c1 = (s.r3 >> 6) & 1u, c2 = (s.r3 >> 1) & 1u;
	const uint32_t m = (c0 + c1 + c2) >> 1;
	s.r0 = c0 == m ? step19(s.r0) : s.r0;
	s.r1 = c1 == m ? step22(s.r1) : s.r1;
	s.r2 = c2 == m ? step23(s.r2) : s.r2;
	s.r3 = step17(s.r3);
}

__device__ __forceinline__ uint32_t maj(uint32_t v, int a, int b, int c)
{
	return (((v >> a) & 1u) + ((v >> b) & 1u) + ((v >> c) & 1u)) >> 1;
}

__device__ __forceinline__ uint32_t output(const A51 &s)
{
	return (maj(s.r0, 1, 6, 15) ^ ((s.r0 >> 11) & 1u)) ^ (maj(s.r1, 3, 8, 14) ^ ((s.r1 >> 1) & 1u)) ^
	       (maj(s.r2, 4, 15, 19) ^ (s.r2 & 1u));
}

// nbits keystream bits as ubits; 4 per 32-bit store when the row allows it
__device__ __forceinline__ void emit(A51 &s, uint8_t *row, int nbits, bool word_ok)
{
	int i = 0;
	if (row && word_ok)
		for (; i + 4 <= nbits; i += 4) {
			uint32_t w = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				clock_rule(s);
				w |= output(s) << (8 * k);
			}
			*reinterpret_cast<uint32_t *>(row + i) = w;
		}
	for (; i < nbits; i++) {
		clock_rule(s);
		if (row)
			row[i] = (uint8_t)output(s);
	}
}

}  // namespace

__global__ void __launch_bounds__(128) a5_kernel(const A5Args a)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (a.n_dev ? min(a.n, *a.n_dev * (a.n_dev_mul ? a.n_dev_mul : 1)) : a.n))
		return;
	uint8_t *dl = a.dl ? a.dl + (size_t)t * a.stride : nullptr;
	uint8_t *ul = a.ul ? a.ul + (size_t)t * a.stride : nullptr;
	const bool word_ok = ((a.stride & 3) == 0) && ((((uintptr_t)a.dl) | ((uintptr_t)a.ul)) & 3) == 0;
	const int alg = a.alg ? a.alg[t] : a.alg0;
	if (alg != 1) {                  // A5/0: all-zero streams; A5/2..7 do not exist for GMR-1 (a5.c:73-76)
		if (alg == 0)
			for (int i = 0; i < a.nbits; i++) {
				if (dl) dl[i] = 0;
				if (ul) ul[i] = 0;
			}
		return;
	}
	const uint8_t *key = a.key + (size_t)t * 8;
	const uint32_t fn = a.fn[t];
	uint8_t k[8];
#pragma unroll
	for (int i = 0; i < 8; i++)
		k[i] = key[i ^ 1];
	k[6] ^= (uint8_t)((fn & 0x0000fu) << 4);
	k[3] ^= (uint8_t)((fn & 0x00030u) << 2);
	k[1] ^= (uint8_t)((fn & 0x007c0u) >> 3);
	k[0] ^= (uint8_t)((fn & 0x0f800u) >> 11);
	k[0] ^= (uint8_t)((fn & 0x70000u) >> 11);
	A51 s = {0, 0, 0, 0};
#pragma unroll
	for (int j = 0; j < 8; j++)
#pragma unroll 1
		for (int i = 7; i >= 0; i--) {
			const uint32_t b = (k[j] >> i) & 1u;
			s.r0 = step19(s.r0) ^ b;
			s.r1 = step22(s.r1) ^ b;
			s.r2 = step23(s.r2) ^ b;
			s.r3 = step17(s.r3) ^ b;
		}
	s.r0 |= 1u; s.r1 |= 1u; s.r2 |= 1u; s.r3 |= 1u;
#pragma unroll 1
	for (int i = 0; i < 250; i++)
		clock_rule(s);
	emit(s, dl, a.nbits, word_ok);
	if (ul)
		emit(s, ul, a.nbits, word_ok);
}

cudaError_t launch_a5(const A5Args &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	a5_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a);
	return cudaGetLastError();
}

}  // namespace gmr1
