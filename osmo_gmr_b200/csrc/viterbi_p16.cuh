// viterbi_p16.cuh - the thread-per-codeword Viterbi of viterbi_tpc.cuh with TWO codewords per thread: the path
// metrics of codeword A in the low and of codeword B in the high 16 bits of one register.  One instruction stream
// decides for two codewords.
//
// The metrics are kept below 2^15, so that
//   * adding a (non-negative) branch sum never carries from the low into the high half: the adds are ORDINARY 32-bit
//     adds, which the compiler spreads over the integer ALU and the multiply-add pipe (IMAD.IADD) - the packed
//     VIADD.16x2 exists on the ALU only, and a first version built on it was no faster than one codeword per thread,
//     because both forms were bound by that one pipe;
//   * bit 15 / bit 31 are free for a guard: with G = 0x80008000 added to b, t = (b + G) - a is one ordinary subtract
//     that cannot borrow across the halves, and bit 15 of each half of t says b >= a.  The decision "b < a" (strict:
//     ties keep a, as the reference does) is the inverted guard bit;
//   * the minimum is the packed unsigned VIMNMX.U16x2 - the only packed instruction left.
//
// Why 16 (15) bits are exact (the reference, osmo_conv_decode as called from src/l1/*.c, keeps 32-bit metrics that
// are never renormalised):
//   * a branch metric per soft bit is ((is -+ 127)^2) >> 9 <= 127, so a step adds at most 127 N to any metric;
//   * only DECISIONS (b < a) and the final metric leave the forward pass.  Taking the same amount off all states
//     changes no decision, so the metrics are renormalised (minimum taken off, the sum kept per codeword in 32 bits and
//     added back to the reported metric) every 64 (N <= 2) or 32 steps: after a renormalisation the spread between
//     states is at most (K-1) 127 N (any state is K-1 steps from the best one), and the steps until the next one add
//     at most 64 x 254 resp. 32 x 635: below 23 000 in every case;
//   * the reference starts with state 0 at 0 and all others at MAX_AE = 0xffffff: "unreachable", which loses against
//     every real path and ties against another unreachable one up to the branch metrics.  0x4000 does the same as long
//     as real metrics stay below it for the K-1 steps the sentinel lives (they are <= (K-1) 127 N <= 3 810);
//   * FLUSH steps: the reference sets the odd states to MAX_AE and keeps deciding for states that only unreachable
//     paths enter.  The traceback from end state 0 never visits those, so only the reachable states (low j bits zero
//     after j flush steps) are computed - no sentinel, and 8 + 4 + 2 + 1 instead of 4 x 8 updates for K = 5.
// Results are bit-identical to viterbi_tpc.cuh (and to the reference): tests/test_decode_emu.py runs both on the CPU.
#pragma once
#include "viterbi_tpc.cuh"

namespace gmr1 {

GMR1_HD uint32_t p16_minu(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
	return __vminu2(a, b);
#else
	const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
	return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
// low halves of lo and hi side by side
GMR1_HD uint32_t p16_pack(uint32_t lo, uint32_t hi)
{
#ifdef __CUDA_ARCH__
	return __byte_perm(lo, hi, 0x5410);
#else
	return (lo & 0xffffu) | (hi << 16);
#endif
}

static constexpr uint32_t P16_UNREACHABLE = 0x40004000u;
static constexpr uint32_t P16_GUARD = 0x80008000u;
// steps between renormalisations (a power of two, even) for a rate 1/N code
GMR1_HD constexpr int p16_renorm_every(int n) { return n <= 2 ? 64 : 32; }
// does a forward pass of n_steps steps need renormalising at all?
GMR1_HD constexpr bool p16_needs_renorm(int n, int k, int n_steps) { return 0x4000 + 127 * n * (n_steps + k) > 0x7fff; }

// ---- soft bit -> branch metrics through a table ----------------------------------------------------------------
// One 32-bit word per soft-bit byte value: low half m0 = ((is - 127)^2) >> 9, high half m1 = ((is + 127)^2) >> 9
// (both 0 for an erased soft bit, soft_metrics).  `flipped` is the same table for the negated soft bit (the gather program's
// descrambling flips), so a flip costs nothing: the table base is picked by a uniform select.  The kernel keeps the
// two tables in shared memory (2 KB per CTA); lanes that hold the same value read the same word (broadcast).
using P16Lut = MetricLut;      // viterbi_tpc.cuh
GMR1_HD uint32_t p16_lut_entry(int is)
{
	uint32_t m0, m1;
	soft_metrics(is, m0, m1);
	return m0 | (m1 << 16);
}
// entry i of the 512 words of a P16Lut
GMR1_HD uint32_t p16_lut_word(int i)
{
	const int v = (int)(int8_t)(i & 0xff);
	return p16_lut_entry(i < 256 ? v : sbit_neg(v));
}

GMR1_HD uint32_t p16_ld_u32(const uint32_t *base, unsigned idx)
{
#ifdef __CUDA_ARCH__
	uint32_t v;
	asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(base) + idx * 4u));
	return v;
#else
	return base[idx];
#endif
}
GMR1_HD unsigned p16_row_u8(const int8_t *row, unsigned idx)
{
#ifdef __CUDA_ARCH__
	unsigned v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(row) + idx));
	return v;
#else
	return (uint8_t)row[idx];
#endif
}
GMR1_HD uint32_t p16_pack_hi(uint32_t lo, uint32_t hi)      // high halves of lo and hi side by side
{
#ifdef __CUDA_ARCH__
	return __byte_perm(lo, hi, 0x7632);
#else
	return (lo >> 16) | (hi & 0xffff0000u);
#endif
}

// packed metrics of step i for the two codewords: m0[j] and m1[j] per soft bit j
template <class C, bool HAS_G2, bool ERASE>
GMR1_HD void p16_fetch(uint32_t (&m0)[C::N], uint32_t (&m1)[C::N], const P16Lut *lut, const int8_t *rowA, const int8_t *rowB,
                       const uint16_t *g, const uint16_t *g2, int i)
{
#pragma unroll
	for (int j = 0; j < C::N; j++) {
		const uint16_t w = g[i * C::N + j];
		uint32_t la, lb;
		if (HAS_G2) {                             // RACH: two sources averaged (rach.c:159-160), then the table
			int sa = gather_sbit<ERASE>(rowA, w), sb = gather_sbit<ERASE>(rowB, w);
			const uint16_t w2 = g2[i * C::N + j];
			if (w2 != G_ERASED) {
				sa = (sa + gather_sbit(rowA, w2)) >> 1;
				sb = (sb + gather_sbit(rowB, w2)) >> 1;
			}
			la = p16_ld_u32(lut->plain, (unsigned)sa & 0xffu);
			lb = p16_ld_u32(lut->plain, (unsigned)sb & 0xffu);
		} else if (ERASE && (w & 0x8000u)) {      // punctured position: no metric
			la = lb = 0;
		} else {
			const uint32_t *t = (w & G_FLIP) ? lut->flipped : lut->plain;
			la = p16_ld_u32(t, p16_row_u8(rowA, w & G_IDX));
			lb = p16_ld_u32(t, p16_row_u8(rowB, w & G_IDX));
		}
		m0[j] = p16_pack(la, lb);
		m1[j] = p16_pack_hi(la, lb);
	}
}

// packed branch sums of one step for the two codewords: bm[o] for every N-bit output o (MSB = first generator);
// ordinary adds (no half can overflow)
template <class C>
GMR1_HD void p16_branch_sums(const uint32_t (&m0)[C::N], const uint32_t (&m1)[C::N], uint32_t (&bm)[1 << C::N])
{
	bm[0] = 0;
#pragma unroll
	for (int j = 0; j < C::N; j++) {
#pragma unroll
		for (int o = (1 << j) - 1; o >= 0; o--) {
			const uint32_t base = bm[o];
			bm[2 * o + 1] = base + m1[j];
			bm[2 * o] = base + m0[j];
		}
	}
}

// one step.  dec[w]: decisions of states 16 w .. 16 w + 15, codeword A in bits 0..15, codeword B in bits 16..31.
// FLUSH >= 0: flush step number FLUSH (0-based), only the 0-input branches into the still reachable states.
template <class C, int FLUSH>
GMR1_HD void p16_step(const uint32_t (&ae)[C::NS], uint32_t (&nae)[C::NS], const uint32_t (&m0)[C::N],
                      const uint32_t (&m1)[C::N], uint32_t (&dec)[C::NS / 16])
{
	constexpr int NS = C::NS, H = NS / 2;
	uint32_t bm[1 << C::N];
	p16_branch_sums<C>(m0, m1, bm);
	// four partial words per 16 states: one OR chain over all 16 would serialise the step (each OR waits for the last)
	uint32_t part[NS / 16][4];
#pragma unroll
	for (int w = 0; w < NS / 16; w++)
#pragma unroll
		for (int q = 0; q < 4; q++)
			part[w][q] = 0;
#pragma unroll
	for (int s = NS - 1; s >= 0; s--) {
		const int k = s >> 1, bit = s & 1;
		if (FLUSH >= 0 && (bit || (k & ((1 << (FLUSH < 0 ? 0 : FLUSH)) - 1))))
			continue;
		const uint32_t a = ae[k] + bm[C::out(k, bit)], b = ae[k + H] + bm[C::out(k + H, bit)];
		const uint32_t t = (b + P16_GUARD) - a;          // guard bit of a half still set <=> b >= a there
		nae[s] = p16_minu(a, b);
		part[s >> 4][s & 3] |= ~(t >> (15 - (s & 15))) & (0x10001u << (s & 15));
	}
#pragma unroll
	for (int w = 0; w < NS / 16; w++)
		dec[w] = (part[w][0] | part[w][1]) | (part[w][2] | part[w][3]);
}

template <class C>
GMR1_HD void p16_store_dec(const uint32_t (&dec)[C::NS / 16], uint32_t *dec_base, int T, int t, int i)
{
#pragma unroll
	for (int w = 0; w < C::NS / 16; w++)
		dec_base[(size_t)(i * (C::NS / 16) + w) * T + t] = dec[w];
}

// minimum of every half taken off all states; the two minima go to offA / offB
template <class C>
GMR1_HD void p16_renorm(uint32_t (&ae)[C::NS], uint32_t &offA, uint32_t &offB)
{
	uint32_t mn = ae[0];
#pragma unroll
	for (int s = 1; s < C::NS; s++)
		mn = p16_minu(mn, ae[s]);
#pragma unroll
	for (int s = 0; s < C::NS; s++)
		ae[s] -= mn;                                     // no half borrows: mn is the minimum of each
	offA += mn & 0xffffu;
	offB += mn >> 16;
}

// forward pass over steps step0 .. step0 + nsteps - 1 (data steps), decisions stored when STORE
template <class C, bool STORE, bool HAS_G2, bool ERASE, bool RENORM>
GMR1_HD void p16_forward(uint32_t (&ae)[C::NS], const P16Lut *lut, const int8_t *rowA, const int8_t *rowB,
                         const uint16_t *g, const uint16_t *g2, int step0, int nsteps, uint32_t *dec_base, int T, int t,
                         uint32_t &offA, uint32_t &offB)
{
	uint32_t tmp[C::NS];
	int i = step0;
	const int end = step0 + nsteps;
	for (; i + 1 < end; i += 2) {
		uint32_t m0[C::N], m1[C::N];
		uint32_t dec[C::NS / 16];
		p16_fetch<C, HAS_G2, ERASE>(m0, m1, lut, rowA, rowB, g, g2, i);
		p16_step<C, -1>(ae, tmp, m0, m1, dec);
		if (STORE)
			p16_store_dec<C>(dec, dec_base, T, t, i);
		p16_fetch<C, HAS_G2, ERASE>(m0, m1, lut, rowA, rowB, g, g2, i + 1);
		p16_step<C, -1>(tmp, ae, m0, m1, dec);
		if (STORE)
			p16_store_dec<C>(dec, dec_base, T, t, i + 1);
		if (RENORM && ((i - step0) & (p16_renorm_every(C::N) - 1)) == p16_renorm_every(C::N) - 2)
			p16_renorm<C>(ae, offA, offB);
	}
	if (i < end) {
		uint32_t m0[C::N], m1[C::N];
		uint32_t dec[C::NS / 16];
		p16_fetch<C, HAS_G2, ERASE>(m0, m1, lut, rowA, rowB, g, g2, i);
		p16_step<C, -1>(ae, tmp, m0, m1, dec);
		if (STORE)
			p16_store_dec<C>(dec, dec_base, T, t, i);
#pragma unroll
		for (int s = 0; s < C::NS; s++)
			ae[s] = tmp[s];
	}
}

// one flush step (K = 5 codes: four of them, FJ = 0 .. 3)
template <class C, int FJ, bool HAS_G2, bool ERASE>
GMR1_HD void p16_flush_step(const uint32_t (&src)[C::NS], uint32_t (&dst)[C::NS], const P16Lut *lut, const int8_t *rowA,
                            const int8_t *rowB, const uint16_t *g, const uint16_t *g2, int i, uint32_t *dec_base, int T, int t)
{
	uint32_t m0[C::N], m1[C::N];
	uint32_t dec[C::NS / 16];
	p16_fetch<C, HAS_G2, ERASE>(m0, m1, lut, rowA, rowB, g, g2, i);
	p16_step<C, FJ>(src, dst, m0, m1, dec);
	p16_store_dec<C>(dec, dec_base, T, t, i);
}

}  // namespace gmr1
