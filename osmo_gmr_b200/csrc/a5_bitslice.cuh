// a5_bitslice.cuh - GMR-1 A5/1 (reference src/l1/a5.c:57-282) for 32 keystreams at once: bit l of every word belongs
// to stream l.  The four shift registers (19 + 22 + 23 + 17 bits) become 81 words; a register "clocks where its
// control bit agrees with the majority" is one select per word under a 32-stream mask, a feedback parity is three
// XORs of whole words, the output bit three word-wide majorities.  ~110 word operations per clock for 32 streams
// against ~50 instructions per clock and stream in the one-stream-per-thread form (a5_kernels.cu: a5_kernel, four
// POPC parities per clock on the quarter-rate pipe).
//
// __host__ __device__: a5_slice_kernel inlines it; tests/emu/chan_emu.cpp runs it on the CPU against the reference
// (tests/test_a5_bitslice_cpu.py).
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__)
struct uint2 { unsigned int x, y; };
#endif
#if defined(__CUDACC__)
#define A5_HD __host__ __device__ __forceinline__
#else
#define A5_HD inline
#endif

namespace gmr1 {
namespace a5s {

struct State {
	uint32_t r1[19], r2[22], r3[23], r4[17];
};

A5_HD uint32_t sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
A5_HD uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { return (a & b) | (a & c) | (b & c); }

template <int L> A5_HD void shift_all(uint32_t (&r)[L], uint32_t in)
{
#pragma unroll
	for (int i = L - 1; i >= 1; i--)
		r[i] = r[i - 1];
	r[0] = in;
}

template <int L> A5_HD void shift_where(uint32_t (&r)[L], uint32_t in, uint32_t m)
{
#pragma unroll
	for (int i = L - 1; i >= 1; i--)
		r[i] = sel(m, r[i - 1], r[i]);
	r[0] = sel(m, in, r[0]);
}

// feedback taps (a5.c: 0x072000, 0x311000, 0x660000, 0x013100)
A5_HD uint32_t fb1(const State &s) { return s.r1[18] ^ s.r1[17] ^ s.r1[16] ^ s.r1[13]; }
A5_HD uint32_t fb2(const State &s) { return s.r2[21] ^ s.r2[20] ^ s.r2[16] ^ s.r2[12]; }
A5_HD uint32_t fb3(const State &s) { return s.r3[22] ^ s.r3[21] ^ s.r3[18] ^ s.r3[17]; }
A5_HD uint32_t fb4(const State &s) { return s.r4[16] ^ s.r4[13] ^ s.r4[12] ^ s.r4[8]; }

A5_HD void init(State &s)
{
#pragma unroll
	for (int i = 0; i < 19; i++) s.r1[i] = 0;
#pragma unroll
	for (int i = 0; i < 22; i++) s.r2[i] = 0;
#pragma unroll
	for (int i = 0; i < 23; i++) s.r3[i] = 0;
#pragma unroll
	for (int i = 0; i < 17; i++) s.r4[i] = 0;
}

// key setup step: all four registers clock, the key bit (one per stream) is XORed into their bit 0
A5_HD void key_step(State &s, uint32_t kb)
{
	const uint32_t f1 = fb1(s), f2 = fb2(s), f3 = fb3(s), f4 = fb4(s);
	shift_all(s.r1, f1 ^ kb);
	shift_all(s.r2, f2 ^ kb);
	shift_all(s.r3, f3 ^ kb);
	shift_all(s.r4, f4 ^ kb);
}

A5_HD void force_bit0(State &s)
{
	s.r1[0] = s.r2[0] = s.r3[0] = s.r4[0] = 0xffffffffu;
}

// one clock under the majority rule of R4's bits 15, 6, 1; R4 itself always clocks
A5_HD void clock(State &s)
{
	const uint32_t c0 = s.r4[15], c1 = s.r4[6], c2 = s.r4[1];
	const uint32_t m = maj3(c0, c1, c2);
	const uint32_t f1 = fb1(s), f2 = fb2(s), f3 = fb3(s), f4 = fb4(s);
	shift_where(s.r1, f1, ~(c0 ^ m));
	shift_where(s.r2, f2, ~(c1 ^ m));
	shift_where(s.r3, f3, ~(c2 ^ m));
	shift_all(s.r4, f4);
}

A5_HD uint32_t output(const State &s)
{
	return maj3(s.r1[1], s.r1[6], s.r1[15]) ^ s.r1[11] ^ maj3(s.r2[3], s.r2[8], s.r2[14]) ^ s.r2[1] ^
	       maj3(s.r3[4], s.r3[15], s.r3[19]) ^ s.r3[0];
}

// the 64 key bits in the order the set-up consumes them: bytes swapped in pairs, the frame number folded in
// (a5.c:232-247), byte j from its bit 7 down; bit q of the result = key bit of set-up step q
A5_HD uint32_t rev_bits_in_bytes(uint32_t x)
{
	x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
	x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
	return ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
}

A5_HD uint64_t folded_key(const uint8_t *key, uint32_t fn)
{
	uint32_t lo, hi;                        // bytes k[0..3], k[4..7] with k[i] = key[i ^ 1]
	if ((((uintptr_t)key) & 7) == 0) {
		const uint2 raw = *reinterpret_cast<const uint2 *>(key);
		lo = ((raw.x >> 8) & 0x00ff00ffu) | ((raw.x & 0x00ff00ffu) << 8);
		hi = ((raw.y >> 8) & 0x00ff00ffu) | ((raw.y & 0x00ff00ffu) << 8);
	} else {
		lo = (uint32_t)key[1] | ((uint32_t)key[0] << 8) | ((uint32_t)key[3] << 16) | ((uint32_t)key[2] << 24);
		hi = (uint32_t)key[5] | ((uint32_t)key[4] << 8) | ((uint32_t)key[7] << 16) | ((uint32_t)key[6] << 24);
	}
	hi ^= (((fn & 0x0000fu) << 4) & 0xffu) << 16;                       // k[6]
	lo ^= (((fn & 0x00030u) << 2) & 0xffu) << 24;                       // k[3]
	lo ^= (((fn & 0x007c0u) >> 3) & 0xffu) << 8;                        // k[1]
	lo ^= (((fn & 0x0f800u) >> 11) ^ ((fn & 0x70000u) >> 11)) & 0xffu;  // k[0]
	return (uint64_t)rev_bits_in_bytes(lo) | ((uint64_t)rev_bits_in_bytes(hi) << 32);
}

// 32 x 32 bit-matrix transpose in place: afterwards bit c of a[l] is what bit l of a[c] was
template <int J> A5_HD void transpose_stage(uint32_t (&a)[32])
{
	constexpr uint32_t M = J == 16 ? 0x0000ffffu : J == 8 ? 0x00ff00ffu : J == 4 ? 0x0f0f0f0fu : J == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
	for (int k = 0; k < 32; k++)
		if ((k & J) == 0) {
			const uint32_t t = ((a[k] >> J) ^ a[k + J]) & M;
			a[k] ^= t << J;
			a[k + J] ^= t;
		}
}

A5_HD void transpose32(uint32_t (&a)[32])
{
	transpose_stage<16>(a);
	transpose_stage<8>(a);
	transpose_stage<4>(a);
	transpose_stage<2>(a);
	transpose_stage<1>(a);
}

// four keystream bits (bits b .. b+3 of v) as four ubit bytes
A5_HD uint32_t spread4(uint32_t v, int b)
{
	return (((v >> b) & 0xfu) * 0x00204081u) & 0x01010101u;
}

}  // namespace a5s
}  // namespace gmr1
