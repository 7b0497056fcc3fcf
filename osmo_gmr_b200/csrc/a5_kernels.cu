// a5_kernels.cu - GMR-1 A5/1 keystream for a whole batch: one thread per (Kc, frame number).
// Replaces, per launch, n calls of gmr1_a5 / gmr1_a5_1 (reference src/l1/a5.c:57-282): key setup
// (byte-swapped pairs, frame number folded into the key, 64 clock-all steps, bit 0 of R1..R4 forced,
// 250 warm-up steps under the majority clocking rule of R4), then nbits downlink and nbits uplink
// bits.  Integer work, bit-exact.  The streams are the sign masks of the ciphered channel decoders
// (TCH3 208 bits, FACCH3 4 x 96, FACCH9 / TCH9 658), so producing them on the device keeps the
// ciphered path free of per-burst host work and of the H2D copy of the masks.
#include <cuda_runtime.h>
#include <stdint.h>

#include "a5_bitslice.cuh"
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

// ---- 32 streams per thread (a5_bitslice.cuh): the form for large batches -----------------------------------------------
// A warp takes 1024 consecutive units: bit l of lane t's words is unit base + 32 l + t.  The key bits are transposed
// into that layout with warp votes (lane l holds the folded key of unit base + 32 l + tt, the vote over bit q is lane
// tt's word of set-up step q), 32 set-up steps at a time so that the key words and the 81 state words fit the register
// file.  The keystream leaves 32 clocks at a time: a 32 x 32 bit transpose turns "one word per clock" into "32 bits per
// unit", which are spread to ubit bytes four at a time (one multiply) and stored as 32-bit words.
constexpr int A5S_THREADS = 64;

__global__ void __launch_bounds__(A5S_THREADS) a5_slice_kernel(const A5Args a)
{
	using namespace a5s;
	const int lane = threadIdx.x & 31;
	const int64_t base = ((int64_t)blockIdx.x * (A5S_THREADS / 32) + (threadIdx.x >> 5)) * 1024;
	const int n_eff = a.n_dev ? min(a.n, *a.n_dev * (a.n_dev_mul ? a.n_dev_mul : 1)) : a.n;
	if (base >= n_eff)
		return;
	const bool word_ok = ((a.stride & 3) == 0) && ((((uintptr_t)a.dl) | ((uintptr_t)a.ul)) & 3) == 0;
	__shared__ uint32_t sw[32 * A5S_THREADS];
	__shared__ uint64_t sk[A5S_THREADS / 32][1024 + 32];     // folded keys of the warp's units, one pad slot per 32
	uint64_t *wk = sk[threadIdx.x >> 5];
	// coalesced pass over the warp's 1024 units: fold the frame number into the key, park it in shared memory
#pragma unroll 4
	for (int i = 0; i < 32; i++) {
		const int64_t u = base + 32 * i + lane;
		uint64_t fk = 0;
		if (u < n_eff && (a.alg ? a.alg[u] : a.alg0) == 1)
			fk = folded_key(a.key + (size_t)u * 8, a.fn[u]);
		wk[33 * i + lane] = fk;
	}
	__syncwarp();
	State s;
	init(s);
#pragma unroll 1
	for (int half = 0; half < 2; half++) {
		uint32_t kb[32];
#pragma unroll
		for (int q = 0; q < 32; q++)
			kb[q] = 0;
#pragma unroll 2
		for (int tt = 0; tt < 32; tt++) {
			// lane l holds the key of unit base + 32 l + tt: the vote over bit q is lane tt's word of set-up step q
			const uint32_t part = (uint32_t)(wk[33 * lane + tt] >> (32 * half));
#pragma unroll
			for (int q = 0; q < 32; q++) {
				const uint32_t w = __ballot_sync(0xffffffffu, (part >> q) & 1u);
				kb[q] = lane == tt ? w : kb[q];
			}
		}
#pragma unroll
		for (int q = 0; q < 32; q++)
			key_step(s, kb[q]);
	}
	force_bit0(s);
#pragma unroll 1
	for (int i = 0; i < 250; i++)
		clock(s);
#pragma unroll 1
	for (int dir = 0; dir < 2; dir++) {
		uint8_t *dst = dir ? a.ul : a.dl;
		if (dir && !a.ul)
			break;
#pragma unroll 1
		for (int b0 = 0; b0 < a.nbits; b0 += 32) {
			uint32_t w[32];
#pragma unroll
			for (int c = 0; c < 32; c++) {
				w[c] = 0;
				if (b0 + c < a.nbits) {
					clock(s);
					w[c] = output(s);
				}
			}
			if (!dst)
				continue;
			transpose32(w);
			// the 32 words go through shared memory so that the store loop can run over the units at run time (unrolled
			// 32 times it made the compiler hoist and spill 32 row addresses)
			__syncwarp();
#pragma unroll
			for (int l = 0; l < 32; l++)
				sw[l * A5S_THREADS + threadIdx.x] = w[l];
			__syncwarp();
#pragma unroll 1
			for (int l = 0; l < 32; l++) {
				const int64_t u = base + 32 * l + lane;
				if (u >= n_eff)
					break;
				const int alg = a.alg ? a.alg[u] : a.alg0;
				if (alg != 0 && alg != 1)          // A5/2..7 do not exist for GMR-1 (a5.c:73-76): rows left alone
					continue;
				const uint32_t bits = alg == 1 ? sw[l * A5S_THREADS + threadIdx.x] : 0u;
				uint8_t *row = dst + (size_t)u * a.stride + b0;
#pragma unroll
				for (int b = 0; b < 32; b += 4) {
					const uint32_t v = spread4(bits, b);
					if (word_ok && b0 + b + 4 <= a.nbits)
						*reinterpret_cast<uint32_t *>(row + b) = v;
					else
						for (int k = 0; k < 4 && b0 + b + k < a.nbits; k++)
							row[b + k] = (uint8_t)(v >> (8 * k));
				}
			}
		}
	}
}

static std::atomic<int> g_a5_mode{-1};             // -1 by batch size, 0 one unit per thread, 1 bitsliced
void a5_force_mode(int mode) { g_a5_mode.store(mode); }

cudaError_t launch_a5(const A5Args &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	// The bitsliced form executes 9 x fewer instructions per unit (36 M against 320 M warp-instructions for 393 216
	// units) but a warp carries 1024 units through 94 k dependent-ish instructions: 0.147 ms however few units there
	// are, and only 384 warps for 393 216 units.  Measured: 0.221 vs 0.401 ms at 393 216 units, 0.147 vs 0.085 at
	// 65 536 - the one-unit-per-thread kernel (1.0 ns per unit) wins below ~150 k units.
	const int mode = g_a5_mode.load();
	const bool slice = mode < 0 ? a.n >= 160 * 1024 : mode == 1;
	if (slice) {
		const int warps = (a.n + 1023) / 1024, per = A5S_THREADS / 32;
		a5_slice_kernel<<<(warps + per - 1) / per, A5S_THREADS, 0, st>>>(a);
	} else {
		a5_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a);
	}
	return cudaGetLastError();
}

}  // namespace gmr1
