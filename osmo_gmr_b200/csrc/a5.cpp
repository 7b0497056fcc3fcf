// a5.cpp - GMR-1 A5/1 keystream generator (host).  Replaces gmr1_a5 / gmr1_a5_1, reference
// src/l1/a5.c:57-282.  Four LFSRs R1..R4 (19, 22, 23, 17 bits; feedback polynomials
// x^19+x^18+x^17+x^14+1, x^22+x^21+x^17+x^13+1, x^23+x^22+x^19+x^18+1, x^17+x^14+x^13+x^9+1);
// R4 clocks every step and its bits 15/6/1 decide by majority which of R1..R3 step; the output is
// the xor of one plain tap and one majority of three taps per register.  The keystream is an
// input of the ciphered channel decoders (sign mask); generating it is O(nbits) integer work per
// burst and is done on the host next to the frame-number bookkeeping it depends on.
#include "../../include/gmr1_b200.h"
#include <string.h>

namespace {

inline uint32_t parity32(uint32_t x) { return (uint32_t)__builtin_parity(x); }

struct A51 {
	uint32_t r[4];
	static constexpr uint32_t LEN[4]  = {19, 22, 23, 17};
	static constexpr uint32_t TAPS[4] = {0x072000, 0x311000, 0x660000, 0x013100};

	static uint32_t step(uint32_t v, int i) { return ((v << 1) & ((1u << LEN[i]) - 1u)) | parity32(v & TAPS[i]); }
	void clock_all() { for (int i = 0; i < 4; i++) r[i] = step(r[i], i); }
	void clock_rule()
	{
		const int c0 = (r[3] >> 15) & 1, c1 = (r[3] >> 6) & 1, c2 = (r[3] >> 1) & 1;
		const int m = (c0 + c1 + c2) >= 2;
		if (c0 == m) r[0] = step(r[0], 0);
		if (c1 == m) r[1] = step(r[1], 1);
		if (c2 == m) r[2] = step(r[2], 2);
		r[3] = step(r[3], 3);
	}
	static int maj(uint32_t v, int a, int b, int c) { return (((v >> a) & 1) + ((v >> b) & 1) + ((v >> c) & 1)) >= 2; }
	int output() const
	{
		const int m0 = maj(r[0], 1, 6, 15) ^ (int)((r[0] >> 11) & 1);
		const int m1 = maj(r[1], 3, 8, 14) ^ (int)((r[1] >> 1) & 1);
		const int m2 = maj(r[2], 4, 15, 19) ^ (int)(r[2] & 1);
		return m0 ^ m1 ^ m2;
	}
};
constexpr uint32_t A51::LEN[4];
constexpr uint32_t A51::TAPS[4];

}  // namespace

extern "C" void gmr1b200_a5(int n, const uint8_t *key, uint32_t fn, int nbits, uint8_t *dl, uint8_t *ul)
{
	if (n == 0) {
		if (dl) memset(dl, 0, nbits);
		if (ul) memset(ul, 0, nbits);
		return;
	}
	if (n != 1)
		return;                       // A5/2..7 do not exist for GMR-1 (a5.c:73-76)
	uint8_t k[8];
	for (int i = 0; i < 8; i++)       // byte-swapped pairs, then the frame number folded in
		k[i] = key[i ^ 1];
	k[6] ^= (uint8_t)((fn & 0x0000f) << 4);
	k[3] ^= (uint8_t)((fn & 0x00030) << 2);
	k[1] ^= (uint8_t)((fn & 0x007c0) >> 3);
	k[0] ^= (uint8_t)((fn & 0x0f800) >> 11);
	k[0] ^= (uint8_t)((fn & 0x70000) >> 11);
	A51 s;
	s.r[0] = s.r[1] = s.r[2] = s.r[3] = 0;
	for (int i = 0; i < 64; i++) {
		const uint32_t b = (k[i >> 3] >> (7 - (i & 7))) & 1;
		s.clock_all();
		for (int j = 0; j < 4; j++)
			s.r[j] ^= b;
	}
	for (int j = 0; j < 4; j++)
		s.r[j] |= 1;
	for (int i = 0; i < 250; i++)
		s.clock_rule();
	for (int i = 0; i < nbits; i++) {
		s.clock_rule();
		if (dl) dl[i] = (uint8_t)s.output();
	}
	if (!ul)
		return;
	for (int i = 0; i < nbits; i++) {
		s.clock_rule();
		ul[i] = (uint8_t)s.output();
	}
}
