// viterbi_tpc.cuh - thread-per-codeword soft Viterbi for the 16-state (K5) and 64-state (K7)
// GMR-1 codes, fused with the gather program (demux/decipher/descramble/deinterleave/
// depuncture), traceback, CRC and L2 packing.
//
// Semantics follow osmo_conv_decode as the reference calls it (src/l1/bcch.c:94 and the
// other seven call sites; algorithm restated in SURVEY.md Appendix A.1):
//   * branch metric per soft bit  ((is - (+-127))^2) >> 9, 0 for an erased (0) soft bit
//   * path metrics 32-bit, start state 0, all others MAX_AE = 0xffffff, never renormalised
//   * survivor: predecessor st>>1 is evaluated first and kept on ties (strict '<' replaces)
//   * FLUSH: K-1 extra steps with the 0-input branch only, end state 0
//   * TAIL_BITING: one seeding pass, subtract min, second pass, end state = first minimum
//   * output bit i = LSB of the state after step i
//
// Why one thread per codeword (and not one warp): all path metrics live in registers, the
// trellis is fully unrolled with compile-time branch-metric indices, and no shuffle/ballot
// traffic is needed - ~6 integer ops per state update, 32 codewords per warp in lock-step.
// The functions are __host__ __device__ so tests can run the identical code on the CPU
// (tests/emu) - that build is a test harness, not a product path.
#pragma once
#include <stdint.h>
#include "gmr1_tables.h"

#ifdef __CUDACC__
#define GMR1_HD __host__ __device__ __forceinline__
#else
#define GMR1_HD inline
#endif

namespace gmr1 {

static constexpr uint32_t MAX_AE = 0x00ffffffu;

// ---- compile-time trellis ---------------------------------------------------------------
template <int N_, int K_, unsigned G0, unsigned G1 = 0, unsigned G2 = 0, unsigned G3 = 0, unsigned G4 = 0>
struct Code {
	static constexpr int N = N_, K = K_, NS = 1 << (K_ - 1);
	static constexpr unsigned poly(int j) { return j == 0 ? G0 : j == 1 ? G1 : j == 2 ? G2 : j == 3 ? G3 : G4; }
	static constexpr unsigned parity(unsigned x)
	{
		x ^= x >> 16; x ^= x >> 8; x ^= x >> 4; x ^= x >> 2; x ^= x >> 1;
		return x & 1u;
	}
	// next_output[s][b], MSB = g0 (reference src/l1/conv.c tables)
	static constexpr unsigned out(unsigned s, unsigned b)
	{
		unsigned reg = (s << 1) | b, ov = 0;
		for (int j = 0; j < N_; j++)
			ov = (ov << 1) | parity(reg & poly(j));
		return ov;
	}
};

using CodeK5_12 = Code<2, 5, 0x19, 0x17>;
using CodeK5_13 = Code<3, 5, 0x15, 0x1b, 0x1f>;
using CodeK5_14 = Code<4, 5, 0x19, 0x17, 0x15, 0x1f>;
using CodeK5_15 = Code<5, 5, 0x15, 0x1b, 0x1f, 0x1d, 0x17>;
using CodeK7_12 = Code<2, 7, 0x6d, 0x4f>;
using CodeK9_13 = Code<3, 9, 0x1ed, 0x19b, 0x127>;

// ---- soft-bit fetch through the gather program -------------------------------------------
GMR1_HD int sbit_neg(int v) { return (int)(int8_t)(-v); }   // int8 negate, -128 stays -128

// one staged soft bit.  On the device the rows live in shared memory and are read through their 32-bit shared
// address: with a generic pointer the compiler re-derives the shared window base (S2UR CgaCtaId, ULEA, a move and a
// three-input add) in front of every one of the 2 x 212 loads of a codeword.
GMR1_HD int row_ld(const int8_t *row, unsigned idx)
{
#ifdef __CUDA_ARCH__
	int v;
	asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(row) + idx));
	return v;
#else
	return row[idx];
#endif
}

// ERASE = false: the program of this channel has no erased (punctured) position - checked when the tables are
// built, gmr1_tables.cpp - so the test (a uniform branch in front of every load) is compiled out
template <bool ERASE = true>
GMR1_HD int gather_sbit(const int8_t *row, uint16_t w)
{
	if (ERASE && (w & 0x8000u))                  // G_ERASED is the only program word with bit 15 set
		return 0;
	int v = row_ld(row, w & G_IDX);
	return (w & G_FLIP) ? sbit_neg(v) : v;
}

// ---- one trellis step --------------------------------------------------------------------
// DW = number of 32-bit decision words per step (NS/32 rounded up)
// branch metric of one soft bit against a transmitted 0 (+127) / 1 (-127): ((is -+ 127)^2) >> 9,
// nothing for an erased (0) soft bit
GMR1_HD void soft_metrics(int is, uint32_t &m0, uint32_t &m1)
{
	const int d0 = is - 127, d1 = is + 127;
	m0 = is ? (uint32_t)((d0 * d0) >> 9) : 0u;
	m1 = is ? (uint32_t)((d1 * d1) >> 9) : 0u;
}

// the same as (m0, m1 - m0): (is + 127)^2 = (is - 127)^2 + 508 is, so the second square is one multiply-add on the
// first, and the difference is 0 for an erased soft bit by itself (no second select)
GMR1_HD void soft_metrics_rel(int is, uint32_t &m0, uint32_t &d)
{
	const int t = is - 127, a = t * t, r0 = a >> 9;
	d = (uint32_t)(((a + 508 * is) >> 9) - r0);
	m0 = is ? (uint32_t)r0 : 0u;
}

// ---- branch metrics through a table ----------------------------------------------------------------------------
// One entry per soft-bit byte value, `flipped` the same for the negated soft bit (the gather program's descrambling
// flips cost nothing: the table base is picked by a uniform select).  Lanes that hold the same value read the same
// word (broadcast).  The kernel keeps the tables in shared memory.
//   two codewords per thread (viterbi_p16.cuh, p16_lut_word): MetricLut, 32-bit words, low half m0, high half m1;
//   one codeword per thread: RelLut, BYTE tables of m0 and d = m1 - m0 (soft_metrics_rel; both fit a byte:
//       m0 <= 127, -127 <= d <= 126), read with sign- / zero-extending byte loads at [table + value]: no index
//       scaling, no field extraction.  The arithmetic form costs ~13 integer-ALU instructions per soft bit (compare,
//       selects, shifts) on the pipe that bounds the kernel; a first table of packed 32-bit words moved the
//       extraction onto the same pipe (LEA.HI.SX32) and won nothing there.
struct MetricLut {
	uint32_t plain[256], flipped[256];
};
struct RelLut {
	int8_t  d[2][256];       // [flipped][value]
	uint8_t m0[2][256];
};
static_assert(sizeof(RelLut) <= sizeof(MetricLut), "the kernels reserve sizeof(MetricLut) for either table");
// entry i (0..511: 256 + value = flipped) of the two byte tables
GMR1_HD void rel_lut_fill(RelLut *lut, int i)
{
	const int v = (int)(int8_t)(i & 0xff);
	uint32_t m0, d;
	soft_metrics_rel(i < 256 ? v : sbit_neg(v), m0, d);
	lut->d[i >> 8][i & 0xff] = (int8_t)(int32_t)d;
	lut->m0[i >> 8][i & 0xff] = (uint8_t)m0;
}
GMR1_HD int lut_ld_s8(const int8_t *base, unsigned idx)
{
#ifdef __CUDA_ARCH__
	int v;
	asm("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(base) + idx));
	return v;
#else
	return base[idx];
#endif
}
GMR1_HD unsigned row_ld_u8(const int8_t *row, unsigned idx)
{
#ifdef __CUDA_ARCH__
	unsigned v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(row) + idx));
	return v;
#else
	return (uint8_t)row[idx];
#endif
}

// (hi << 1) | (lo >> 31): shifts the sign of lo into hi
GMR1_HD uint32_t funnel_l1(uint32_t lo, uint32_t hi)
{
#ifdef __CUDA_ARCH__
	return __funnelshift_l(lo, hi, 1);
#else
	return (hi << 1) | (lo >> 31);
#endif
}

// One trellis step from `ae` into `nae` (callers ping-pong the two arrays so that no register
// copies are needed).  DW = number of 32-bit decision words per step (NS/32 rounded up).
//
// REL: branch metrics relative to the all-zero branch.  Every path takes exactly one branch per step, so taking
// the same amount (the metric of output 0...0, sum of m0[j]) off all 2^N branch sums of a step changes no
// decision; the path metrics then run `off` below their reference values (off = running sum of what was taken
// off, handed back to the caller for conv_rv and the MAX_AE sentinel of the flush steps).  bm[0] becomes the
// constant 0: the 2^(K-1-... ) transitions with an all-zero output need no add, and the branch sums need N - 1 adds
// less.  Relative metrics can be negative: comparisons are signed (|values| < 2^25).
template <class C, bool FLUSH_STEP, bool REL>
GMR1_HD void acs_core(const uint32_t (&ae)[C::NS], uint32_t (&nae)[C::NS], const uint32_t (&m0)[C::N], const uint32_t (&m1)[C::N],
                      uint32_t (&dec)[(C::NS + 31) / 32], uint32_t &off);

template <class C, bool FLUSH_STEP, bool REL = false>
GMR1_HD void acs_step(const uint32_t (&ae)[C::NS], uint32_t (&nae)[C::NS], const int (&v)[C::N],
                      uint32_t (&dec)[(C::NS + 31) / 32], uint32_t &off)
{
	uint32_t m0[C::N], m1[C::N];                 // REL: m1 holds m1 - m0
#pragma unroll
	for (int j = 0; j < C::N; j++) {
		if (REL)
			soft_metrics_rel(v[j], m0[j], m1[j]);
		else
			soft_metrics(v[j], m0[j], m1[j]);
	}
	acs_core<C, FLUSH_STEP, REL>(ae, nae, m0, m1, dec, off);
}

// the step proper, from the per-bit metrics (REL: m1 holds m1 - m0)
template <class C, bool FLUSH_STEP, bool REL>
GMR1_HD void acs_core(const uint32_t (&ae)[C::NS], uint32_t (&nae)[C::NS], const uint32_t (&m0)[C::N], const uint32_t (&m1)[C::N],
                      uint32_t (&dec)[(C::NS + 31) / 32], uint32_t &off)
{
	constexpr int N = C::N, NS = C::NS, H = NS / 2;
	// all 2^N branch sums, built by doubling (entries that no transition uses are dead code)
	uint32_t bm[1 << N];
	bm[0] = 0;
	if (REL) {
#pragma unroll
		for (int j = 0; j < N; j++) {
			const uint32_t dj = m1[j];
			off += m0[j];
#pragma unroll
			for (int o = (1 << j) - 1; o >= 0; o--) {
				const uint32_t base = bm[o];
				bm[2 * o + 1] = base + dj;
				bm[2 * o]     = base;
			}
		}
	} else {
#pragma unroll
		for (int j = 0; j < N; j++) {
#pragma unroll
			for (int o = (1 << j) - 1; o >= 0; o--) {
				uint32_t base = bm[o];
				bm[2 * o + 1] = base + m1[j];
				bm[2 * o]     = base + m0[j];
			}
		}
	}
	if (FLUSH_STEP) {
		// only the 0-input branches; odd states become unreachable
#pragma unroll
		for (int i = 0; i < (NS + 31) / 32; i++)
			dec[i] = 0;
#pragma unroll
		for (int k = 0; k < H; k++) {
			const uint32_t a = ae[k] + bm[C::out(k, 0)], b = ae[k + H] + bm[C::out(k + H, 0)];
			const bool d = REL ? (int32_t)(b - a) < 0 : b < a;
			nae[2 * k] = d ? b : a;
			dec[(2 * k) >> 5] |= d ? (1u << ((2 * k) & 31)) : 0u;
			nae[2 * k + 1] = REL ? MAX_AE - off : MAX_AE;
		}
	} else {
		// States from the highest down: the decision "b < a" is the sign of b - a (both below 2^25), shifted
		// into the decision word from the right - one funnel shift per state, and the subtraction can issue
		// on the multiply pipe (IMAD), which takes load off the integer ALU this kernel is bound by.
		// Two partial words per 32 states keep the shift chains short.
		constexpr int DW = (NS + 31) / 32, PER = NS / DW;          // states per decision word
#pragma unroll
		for (int w = DW - 1; w >= 0; w--) {
			uint32_t acc_hi = 0, acc_lo = 0;
#pragma unroll
			for (int q = PER - 1; q >= 0; q--) {
				const int s = w * PER + q, k = s >> 1, bit = s & 1;
				const uint32_t a = ae[k] + bm[C::out(k, bit)], b = ae[k + H] + bm[C::out(k + H, bit)];
				const uint32_t diff = b - a;                       // negative <=> b < a (strict: ties keep a)
				if (REL)
					nae[s] = (uint32_t)((int32_t)b < (int32_t)a ? (int32_t)b : (int32_t)a);
				else
					nae[s] = b < a ? b : a;
				if (q >= PER / 2)
					acc_hi = funnel_l1(diff, acc_hi);
				else
					acc_lo = funnel_l1(diff, acc_lo);
			}
			dec[w] = (acc_hi << (PER / 2)) | acc_lo;
		}
	}
}

// decision storage: word w of step i of thread t lives at dec[(i*DW + w)*T + t] (T = threads
// per CTA) so a warp's accesses are consecutive.  16-state codes use 16-bit words.
template <int NS> struct DecWord { using type = uint32_t; };
template <> struct DecWord<16> { using type = uint16_t; };

// ---- forward pass over n steps ------------------------------------------------------------
// g: gather program (N words per step), g2: optional second source averaged in (RACH)
template <class C, bool HAS_G2, bool ERASE = true>
GMR1_HD void fetch_inputs(int (&v)[C::N], const int8_t *row, const uint16_t *g, const uint16_t *g2, int i)
{
#pragma unroll
	for (int j = 0; j < C::N; j++) {
		int s = gather_sbit<ERASE>(row, g[i * C::N + j]);
		if (HAS_G2) {
			const uint16_t w2 = g2[i * C::N + j];
			if (w2 != G_ERASED)
				s = (s + gather_sbit(row, w2)) >> 1;    // rach.c:159-160
		}
		v[j] = s;
	}
}

// relative metrics (m0, d = m1 - m0) of the N soft bits of step i through the byte tables (RelLut)
template <class C, bool HAS_G2, bool ERASE = true>
GMR1_HD void fetch_metrics_rel(uint32_t (&m0)[C::N], uint32_t (&d)[C::N], const RelLut *lut, const int8_t *row,
                               const uint16_t *g, const uint16_t *g2, int i)
{
#pragma unroll
	for (int j = 0; j < C::N; j++) {
		const uint16_t w = g[i * C::N + j];
		if (HAS_G2) {                             // RACH: two sources averaged (rach.c:159-160), then the table
			int s = gather_sbit<ERASE>(row, w);
			const uint16_t w2 = g2[i * C::N + j];
			if (w2 != G_ERASED)
				s = (s + gather_sbit(row, w2)) >> 1;
			m0[j] = row_ld_u8((const int8_t *)lut->m0[0], (unsigned)s & 0xffu);
			d[j] = (uint32_t)lut_ld_s8(lut->d[0], (unsigned)s & 0xffu);
		} else if (ERASE && (w & 0x8000u)) {      // punctured position: no metric
			m0[j] = 0;
			d[j] = 0;
		} else {
			const int f = (w & G_FLIP) ? 1 : 0;
			const unsigned val = row_ld_u8(row, w & G_IDX);
			m0[j] = row_ld_u8((const int8_t *)lut->m0[f], val);
			d[j] = (uint32_t)lut_ld_s8(lut->d[f], val);
		}
	}
}

template <class C, bool STORE>
GMR1_HD void store_dec(const uint32_t (&dec)[(C::NS + 31) / 32], typename DecWord<C::NS>::type *dec_base, int T, int t, int i)
{
	constexpr int DW = (C::NS + 31) / 32;
	if (STORE) {
#pragma unroll
		for (int w = 0; w < DW; w++)
			dec_base[(size_t)(i * DW + w) * T + t] = (typename DecWord<C::NS>::type)dec[w];
	}
}

// LUT (REL only): the per-bit metrics come from the table `lut` instead of the arithmetic of soft_metrics_rel
template <class C, bool FLUSH_STEP, bool STORE, bool HAS_G2, bool REL = false, bool ERASE = true, bool LUT = false>
GMR1_HD void forward(uint32_t (&ae)[C::NS], const int8_t *row, const uint16_t *g, const uint16_t *g2,
                     int step0, int nsteps, typename DecWord<C::NS>::type *dec_base, int T, int t, uint32_t &off,
                     const RelLut *lut = nullptr)
{
	constexpr int DW = (C::NS + 31) / 32;
	uint32_t tmp[C::NS];
	int i = step0;
	const int end = step0 + nsteps;
	if constexpr (LUT) {
		static_assert(REL, "the metric table holds relative metrics");
		for (; i + 1 < end; i += 2) {
			uint32_t m0[C::N], d[C::N];
			uint32_t dec[DW];
			fetch_metrics_rel<C, HAS_G2, ERASE>(m0, d, lut, row, g, g2, i);
			acs_core<C, FLUSH_STEP, true>(ae, tmp, m0, d, dec, off);
			store_dec<C, STORE>(dec, dec_base, T, t, i);
			fetch_metrics_rel<C, HAS_G2, ERASE>(m0, d, lut, row, g, g2, i + 1);
			acs_core<C, FLUSH_STEP, true>(tmp, ae, m0, d, dec, off);
			store_dec<C, STORE>(dec, dec_base, T, t, i + 1);
		}
		if (i < end) {
			uint32_t m0[C::N], d[C::N];
			uint32_t dec[DW];
			fetch_metrics_rel<C, HAS_G2, ERASE>(m0, d, lut, row, g, g2, i);
			acs_core<C, FLUSH_STEP, true>(ae, tmp, m0, d, dec, off);
			store_dec<C, STORE>(dec, dec_base, T, t, i);
#pragma unroll
			for (int s = 0; s < C::NS; s++)
				ae[s] = tmp[s];
		}
	} else {
	for (; i + 1 < end; i += 2) {           // two steps per iteration: ae -> tmp -> ae
		int v[C::N];
		uint32_t dec[DW];
		fetch_inputs<C, HAS_G2, ERASE>(v, row, g, g2, i);
		acs_step<C, FLUSH_STEP, REL>(ae, tmp, v, dec, off);
		store_dec<C, STORE>(dec, dec_base, T, t, i);
		fetch_inputs<C, HAS_G2, ERASE>(v, row, g, g2, i + 1);
		acs_step<C, FLUSH_STEP, REL>(tmp, ae, v, dec, off);
		store_dec<C, STORE>(dec, dec_base, T, t, i + 1);
	}
	if (i < end) {
		int v[C::N];
		uint32_t dec[DW];
		fetch_inputs<C, HAS_G2, ERASE>(v, row, g, g2, i);
		acs_step<C, FLUSH_STEP, REL>(ae, tmp, v, dec, off);
		store_dec<C, STORE>(dec, dec_base, T, t, i);
#pragma unroll
		for (int s = 0; s < C::NS; s++)
			ae[s] = tmp[s];
	}
	}
}

// ---- traceback ------------------------------------------------------------------------------
// Walks steps n_steps-1 .. 0 from end_state, emits bit i (< len) = LSB of the state after
// step i through emit(i, bit).
template <class C, class Emit>
GMR1_HD void traceback(const typename DecWord<C::NS>::type *dec_base, int T, int t,
                       int n_steps, int len, unsigned end_state, Emit emit)
{
	constexpr int DW = (C::NS + 31) / 32;
	unsigned st = end_state;
	for (int i = n_steps - 1; i >= 0; i--) {
		const uint32_t d = dec_base[(size_t)(i * DW + (DW > 1 ? (st >> 5) : 0)) * T + t];
		const unsigned bit = (d >> (st & 31)) & 1u;
		if (i < len)
			emit(i, st & 1u);
		st = (st >> 1) | (bit << (C::K - 2));
	}
}

// ---- bit-serial CRC over LSB-first packed bytes (osmo_crc{8,16}gen_*) --------
// returns 0 when the `bits` CRC bits that follow the data match, 1 otherwise
GMR1_HD int crc_check_packed(const uint8_t *p, int bit0, int n_data, unsigned poly, int bits)
{
	const unsigned top = 1u << (bits - 1), mask = (1u << bits) - 1u;
	unsigned crc = 0;
	for (int i = 0; i < n_data; i++) {
		const int q = bit0 + i;
		const unsigned b = (p[q >> 3] >> (q & 7)) & 1u;
		crc ^= b << (bits - 1);
		crc = (crc & top) ? ((crc << 1) ^ poly) : (crc << 1);
	}
	crc &= mask;
	int bad = 0;
	for (int i = 0; i < bits; i++) {
		const int q = bit0 + n_data + i;
		const unsigned b = (p[q >> 3] >> (q & 7)) & 1u;
		bad |= (int)(b ^ ((crc >> (bits - 1 - i)) & 1u));
	}
	return bad;
}


// ---- CRC16 (poly 0x1021, init 0) a byte at a time -------------------------------------------------------
// Same register as the bit-serial loop above after every 8 bits: crc' = (crc << 8) ^ T[(crc >> 8) ^ b] with b = the
// next 8 stream bits, first bit in the MSB.  The stream is LSB-first inside the packed bytes, hence the bit reversal.
// 192 data bits cost 24 table steps instead of 192 bit steps (the check was 5 % of the BCCH decode kernel).
struct Crc16Tab {
	uint16_t v[256];
	constexpr Crc16Tab() : v()
	{
		for (int i = 0; i < 256; i++) {
			unsigned c = (unsigned)i << 8;
			for (int k = 0; k < 8; k++)
				c = (c & 0x8000u) ? ((c << 1) ^ 0x1021u) : (c << 1);
			v[i] = (uint16_t)c;
		}
	}
};
#ifdef __CUDACC__
static __device__ const Crc16Tab d_crc16_tab = Crc16Tab();      // global memory: per-lane indices (a __constant__ table would serialise)
#endif
static constexpr Crc16Tab h_crc16_tab = Crc16Tab();

GMR1_HD unsigned crc16_tab(unsigned i)
{
#ifdef __CUDA_ARCH__
	return __ldg(&d_crc16_tab.v[i]);
#else
	return h_crc16_tab.v[i];
#endif
}

GMR1_HD unsigned rev8(unsigned b)
{
#ifdef __CUDA_ARCH__
	return __brev(b) >> 24;
#else
	b = ((b & 0xf0u) >> 4) | ((b & 0x0fu) << 4);
	b = ((b & 0xccu) >> 2) | ((b & 0x33u) << 2);
	return ((b & 0xaau) >> 1) | ((b & 0x55u) << 1);
#endif
}

// crc_check_packed(p, 0, n_data, 0x1021, 16) for data that starts on a byte boundary
GMR1_HD int crc16_check_packed(const uint8_t *p, int n_data)
{
	unsigned crc = 0;
	const int nb = n_data >> 3;
	for (int i = 0; i < nb; i++)
		crc = ((crc << 8) ^ crc16_tab(((crc >> 8) ^ rev8(p[i])) & 0xffu)) & 0xffffu;
	for (int q = 8 * nb; q < n_data; q++) {                     // ragged tail, bit by bit
		const unsigned b = (p[q >> 3] >> (q & 7)) & 1u;
		crc ^= b << 15;
		crc = ((crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) : (crc << 1)) & 0xffffu;
	}
	// the 16 bits that follow the data, first one against the MSB of the register
	unsigned rx = 0;
	if ((n_data & 7) == 0) {
		rx = (rev8(p[nb]) << 8) | rev8(p[nb + 1]);
	} else {
		for (int i = 0; i < 16; i++) {
			const int q = n_data + i;
			rx |= ((p[q >> 3] >> (q & 7)) & 1u) << (15 - i);
		}
	}
	return rx != crc;
}

}  // namespace gmr1
