// decode_unit.cuh - per-channel decode of ONE unit (codeword / burst) by ONE thread:
// staged soft-bit row -> Viterbi -> traceback -> CRC -> packed L2 + side outputs.
// Mirrors the reference decode entry points (return value = CRC result, *conv_rv = Viterbi
// metric): gmr1_bcch_decode (src/l1/bcch.c:84), gmr1_ccch_decode (ccch.c:88),
// gmr1_facch3_decode (facch3.c:122), gmr1_facch9_decode (facch9.c:107), gmr1_tch9_decode
// (tch9.c:140), gmr1_rach_decode (rach.c:137), gmr1_tch3_decode (tch3.c:124).
#pragma once
#include "viterbi_tpc.cuh"
#include "viterbi_p16.cuh"

namespace gmr1 {

// Batched decode arguments. All pointers are device pointers on the GPU path. Unused ones
// are NULL. Layouts are unit-major, exactly n consecutive copies of the reference's arrays.
struct DecodeArgs {
	const int8_t  *ebits;     // [n][n_in]
	const uint8_t *ciph;      // [n][n_ciph] ubit (one byte per bit) or NULL
	int32_t        n;
	uint8_t       *l2;        // [n][l2_bytes]   (tch3: frame0 [n][10])
	uint8_t       *l2b;       // tch3: frame1 [n][10]
	int32_t       *conv;      // [n] or NULL      (tch3: conv0)
	int32_t       *conv1;     // tch3: conv1 [n] or NULL
	int32_t       *crc;       // [n] (return value of the reference function) or NULL
	int32_t       *crc2;      // rach: crc_rv [n][2] or NULL
	uint8_t       *bits_s;    // facch3: [n][32] ubit, tch3: [n][4] ubit, or NULL
	int8_t        *sacch;     // facch9/tch9: [n][10] sbit or NULL
	int8_t        *status;    // facch9/tch9: [n][4] sbit or NULL
	const int32_t *prev1;     // tch9: index of the previous burst of the same channel, -1 = none
	const int32_t *prev2;     // tch9: index of the burst before that, -1 = none
	const uint8_t *sb_mask;   // rach: [n] or NULL (then sb_mask0 for all)
	int32_t        sb_mask0;
	int32_t        tch3_m;    // tch3 multiplexing mode (0 / 1)
	int32_t        t9_rows;   // tch9: ebits is [n][648] already deciphered / descrambled / inter-burst
	                          //       de-interleaved (what gmr1_deinterleave_inter returns); no side outputs
	const int32_t *n_dev;     // optional device-side unit count (<= n): units beyond it are skipped (rx scheduler)
	uint8_t       *dec_scratch;   // optional device scratch of decode_scratch_bytes(ch, n): the survivor decisions go
	                          // there (step-major, coalesced) instead of shared memory, which doubles the resident CTAs
};

// device-resident (or host, in the emulation) tables one channel needs
struct TabRef {
	const uint16_t *g;        // gather program            (uniform access: __constant__)
	const uint16_t *g2;       // RACH second source or NULL (uniform)
	const int16_t  *cmap;     // ebit -> cipher index       (per-lane access: global)
	const uint16_t *t9_src;   // TCH9 staging map           (per-lane access: global)
	int32_t n_in, n_row, n_ciph, n_steps, len;
};

template <int CH> struct ChanCode { using type = CodeK5_12; };
template <> struct ChanCode<CH_FACCH3>   { using type = CodeK5_14; };
template <> struct ChanCode<CH_RACH>     { using type = CodeK5_14; };
template <> struct ChanCode<CH_TCH9_2K4> { using type = CodeK5_15; };
template <> struct ChanCode<CH_TCH9_4K8> { using type = CodeK5_13; };
template <> struct ChanCode<CH_TCH3>     { using type = CodeK7_12; };
template <> struct ChanCode<CH_DC12>     { using type = CodeK9_13; };

// bytes of packed L2 each unit produces
// channels whose gather program has erased (punctured) positions; the others are verified to have none when the
// tables are built (gmr1_tables.cpp: check_no_erasures)
GMR1_HD constexpr bool chan_has_erasures(int ch)
{
	return !(ch == CH_BCCH || ch == CH_CCCH || ch == CH_FACCH3 || ch == CH_FACCH9);
}

// info bits per codeword (the data steps of the forward pass)
GMR1_HD constexpr int chan_len_of(int ch)
{
	return ch == CH_BCCH || ch == CH_CCCH || ch == CH_DC12 ? 208 : ch == CH_FACCH3 ? 92 :
	       ch == CH_FACCH9 ? 316 : ch == CH_TCH9_2K4 ? 144 : ch == CH_TCH9_4K8 ? 240 :
	       ch == CH_TCH9_9K6 ? 480 : ch == CH_RACH ? 159 : 48;
}

GMR1_HD constexpr int chan_l2_bytes(int ch)
{
	return ch == CH_BCCH || ch == CH_CCCH || ch == CH_DC12 ? 24 :
	       ch == CH_FACCH3 ? 10 : ch == CH_FACCH9 ? 38 :
	       ch == CH_TCH9_2K4 ? 18 : ch == CH_TCH9_4K8 ? 30 : ch == CH_TCH9_9K6 ? 60 :
	       ch == CH_RACH ? 18 : 10;
}

// ---- staging: value of staged-row element r of `unit` ---------------------------------------
// Default channels: the row is the ebits row with the cipher sign already applied
// (reference: "if (ciph[i]) bits[i] *= -1", e.g. facch9.c:121-125).
// TCH9: the row is the inter-burst-deinterleaved, descrambled vector (tch9.c:163-166).

// TCH9, element r of the staged row of `unit`: p1 / p2 = index of the burst received one / two bursts earlier on
// the same channel (-1: none), already looked up by the caller (the kernel keeps them in shared memory so that an
// element costs one dependent global load, not two)
GMR1_HD int8_t stage_elem_t9(const TabRef &tb, const DecodeArgs &a, int unit, int p1, int p2, int r)
{
	const uint16_t w = tb.t9_src[r];
	const int age = (w >> 10) & 3, s = w & G_IDX;
	const int u = age == 0 ? unit : (age == 1 ? p1 : p2);
	if (u < 0)
		return 0;
	int v = a.ebits[(size_t)u * tb.n_in + s];
	if (a.ciph && a.ciph[(size_t)u * tb.n_ciph + tb.cmap[s]])
		v = sbit_neg(v);
	if (w & G_FLIP)
		v = sbit_neg(v);
	return (int8_t)v;
}

template <int CH>
GMR1_HD int8_t stage_elem(const TabRef &tb, const DecodeArgs &a, int unit, int r)
{
	constexpr bool T9 = (CH == CH_TCH9_2K4 || CH == CH_TCH9_4K8 || CH == CH_TCH9_9K6);
	if (T9 && a.t9_rows) {
		return a.ebits[(size_t)unit * 648 + r];
	} else if (T9) {
		return stage_elem_t9(tb, a, unit, a.prev1 ? a.prev1[unit] : -1, a.prev2 ? a.prev2[unit] : -1, r);
	} else {
		int v = a.ebits[(size_t)unit * tb.n_in + r];
		if (a.ciph) {
			const int c = tb.cmap[r];
			if (c >= 0 && a.ciph[(size_t)unit * tb.n_ciph + c])
				v = sbit_neg(v);
		}
		return (int8_t)v;
	}
}

// ---- per-unit decode, flush-terminated K5 channels ------------------------------------------
// row: staged soft bits of this unit (n_row bytes, overwritten with the packed output bits),
// dec: decision storage [n_steps][T] words, t: this thread's slot
// side outputs that are plain copies of (deciphered) soft bits
template <int CH>
GMR1_HD void k5_side_outputs(const DecodeArgs &a, int unit)
{
	constexpr bool T9 = (CH == CH_TCH9_2K4 || CH == CH_TCH9_4K8 || CH == CH_TCH9_9K6);
	constexpr bool F9 = (CH == CH_FACCH9);
	if (CH == CH_FACCH3 && a.bits_s) {
		for (int b = 0; b < 4; b++)
			for (int j = 0; j < 8; j++)     // facch3.c:141-142 (status bits are not ciphered)
				a.bits_s[(size_t)unit * 32 + 8 * b + j] = a.ebits[(size_t)unit * 416 + 104 * b + 22 + j] < 0;
	}
	if ((T9 && !a.t9_rows) || F9) {
		const int8_t *e = a.ebits + (size_t)unit * 662;
		if (a.status)
			for (int i = 0; i < 4; i++)     // facch9.c:118
				a.status[(size_t)unit * 4 + i] = e[52 + i];
		if (a.sacch)
			for (int i = 0; i < 10; i++) {  // facch9.c:121-128: my[52+i] after deciphering
				int v = e[56 + i];
				if (a.ciph && a.ciph[(size_t)unit * 658 + 52 + i])
					v = sbit_neg(v);
				a.sacch[(size_t)unit * 10 + i] = (int8_t)v;
			}
	}
}

template <int CH>
GMR1_HD void k5_outputs(const DecodeArgs &a, int unit, uint8_t *out);

// lut != nullptr: branch metrics through the byte tables (RelLut) instead of the arithmetic - same numbers
template <int CH, bool LUT = false>
GMR1_HD void decode_unit_k5(const TabRef &tb, const DecodeArgs &a, int unit,
                            int8_t *row, uint16_t *dec, int T, int t, const RelLut *lut = nullptr)
{
	using C = typename ChanCode<CH>::type;
	k5_side_outputs<CH>(a, unit);

	uint32_t ae[C::NS];
#pragma unroll
	for (int s = 0; s < C::NS; s++)
		ae[s] = s ? MAX_AE : 0u;

	// path metrics relative to the all-zero branch of every step (acs_step REL): `off` is what they lack
	uint32_t off = 0;
	constexpr bool ER = chan_has_erasures(CH);
	forward<C, false, true, CH == CH_RACH, true, ER, LUT>(ae, row, tb.g, tb.g2, 0, tb.len, dec, T, t, off, lut);
	forward<C, true,  true, CH == CH_RACH, true, ER, LUT>(ae, row, tb.g, tb.g2, tb.len, C::K - 1, dec, T, t, off, lut);

	if (a.conv)
		a.conv[unit] = (int32_t)(ae[0] + off);

	// traceback into LSB-first packed bytes, reusing the (now dead) staged row.  Flush steps first
	// (no output), then the data steps in groups of 8 = one output byte each.
	uint8_t *out = (uint8_t *)row;
	{
		unsigned st = 0;
		{
			uint16_t fw[C::K - 1];               // the K - 1 flush steps (n_steps = len + K - 1), requested together
#pragma unroll
			for (int j = 0; j < C::K - 1; j++)
				fw[j] = dec[(size_t)(tb.len + j) * T + t];
#pragma unroll
			for (int j = C::K - 2; j >= 0; j--) {
				const unsigned bit = ((unsigned)fw[j] >> st) & 1u;
				st = (st >> 1) | (bit << (C::K - 2));
			}
		}
		int i = tb.len - 1;
		unsigned acc = 0;
		for (; (i & 7) != 7; i--) {          // ragged top byte (len not a multiple of 8)
			acc |= (st & 1u) << (i & 7);
			const unsigned bit = (dec[(size_t)i * T + t] >> st) & 1u;
			st = (st >> 1) | (bit << (C::K - 2));
		}
		if ((tb.len & 7) != 0)
			out[tb.len >> 3] = (uint8_t)acc;
		// The decision words live in global memory (L2) and their addresses do not depend on the state: 32 steps'
		// words are requested together, then walked (four output bytes).  With 8 per batch the walk waited for L2
		// 26 times per codeword on a kernel with four resident warps per scheduler: 19 % of its stall samples.
		for (; i >= 31; i -= 32) {
			uint16_t dw[32];
#pragma unroll
			for (int b = 0; b < 32; b++)
				dw[b] = dec[(size_t)(i - 31 + b) * T + t];
#pragma unroll
			for (int q = 3; q >= 0; q--) {
				acc = 0;
#pragma unroll
				for (int b = 7; b >= 0; b--) {
					acc |= (st & 1u) << b;
					const unsigned bit = ((unsigned)dw[8 * q + b] >> st) & 1u;
					st = (st >> 1) | (bit << (C::K - 2));
				}
				out[(i >> 3) - 3 + q] = (uint8_t)acc;
			}
		}
		for (; i >= 7; i -= 8) {
			acc = 0;
#pragma unroll
			for (int b = 7; b >= 0; b--) {
				acc |= (st & 1u) << b;
				const unsigned bit = (dec[(size_t)(i - 7 + b) * T + t] >> st) & 1u;
				st = (st >> 1) | (bit << (C::K - 2));
			}
			out[i >> 3] = (uint8_t)acc;
		}
	}
	k5_outputs<CH>(a, unit, out);
}

// CRC(s) and packed L2 of one unit from the decoded bits (LSB-first packed bytes in `out`)
template <int CH>
GMR1_HD void k5_outputs(const DecodeArgs &a, int unit, uint8_t *out)
{
	constexpr bool T9 = (CH == CH_TCH9_2K4 || CH == CH_TCH9_4K8 || CH == CH_TCH9_9K6);
	constexpr int NB = chan_l2_bytes(CH);
	uint8_t *l2 = a.l2 + (size_t)unit * NB;

	if (CH == CH_RACH) {
		// bits_u[0..134] = class 2 (123 data + CRC12), bits_u[135..158] = class 1 (16 data + CRC8)
		int c0 = crc_check_packed(out, 135, 16, 0x9b, 8);
		const int c1 = crc_check_packed(out, 0, 123, 0x80f, 12);
		if (c0) {   // retry with the SB mask applied to the CRC8 bits (rach.c:178-182)
			const unsigned mask = a.sb_mask ? a.sb_mask[unit] : (unsigned)a.sb_mask0;
			for (int i = 0; i < 8; i++) {
				const int q = 135 + 16 + i;
				out[q >> 3] ^= (uint8_t)(((mask >> (7 - i)) & 1u) << (q & 7));
			}
			c0 = crc_check_packed(out, 135, 16, 0x9b, 8);
		}
		if (a.crc2) {
			a.crc2[(size_t)unit * 2] = c0;
			a.crc2[(size_t)unit * 2 + 1] = c1;
		}
		if (a.crc)
			a.crc[unit] = (c0 || c1) ? 1 : 0;
		// rach[0..15] = class-1 data, rach[16..138] = class-2 data, LSB first (rach.c:190-193)
		for (int i = 0; i < 18; i++) {
			unsigned byte = 0;
			for (int b = 0; b < 8; b++) {
				const int o = 8 * i + b;
				int q = -1;
				if (o < 16)       q = 135 + o;
				else if (o < 139) q = o - 16;
				if (q >= 0)
					byte |= ((out[q >> 3] >> (q & 7)) & 1u) << b;
			}
			l2[i] = (uint8_t)byte;
		}
	} else {
		constexpr int n_data = CH == CH_FACCH3 ? 76 : CH == CH_FACCH9 ? 300 :
		                       CH == CH_TCH9_2K4 ? 144 : CH == CH_TCH9_4K8 ? 240 :
		                       CH == CH_TCH9_9K6 ? 480 : 192;
		if (!T9) {
			const int c = crc16_check_packed(out, n_data);
			if (a.crc)
				a.crc[unit] = c;
		}
		for (int i = 0; i < NB; i++) {
			unsigned byte = out[i];
			if (8 * i + 8 > n_data)         // last partial byte: upper bits stay 0
				byte &= (1u << (n_data - 8 * i)) - 1u;
			l2[i] = (uint8_t)byte;
		}
	}
}

// ---- TWO units per thread, flush-terminated K5 channels (viterbi_p16.cuh) -------------------------------------
// unit A in the low, unit B in the high halves.  okB = false: there is no unit B (ragged tail of the batch; its staged
// row is all zeros), nothing is written for it.  dec: [n_steps][T] 32-bit words, t: this thread's slot.
template <int CH>
GMR1_HD void decode_pair_k5(const TabRef &tb, const DecodeArgs &a, const P16Lut *lut, int unitA, int unitB, bool okB,
                            int8_t *rowA, int8_t *rowB, uint32_t *dec, int T, int t)
{
	using C = typename ChanCode<CH>::type;
	static_assert(C::NS == 16, "K = 5 codes");
	k5_side_outputs<CH>(a, unitA);
	if (okB)
		k5_side_outputs<CH>(a, unitB);

	uint32_t ae[C::NS], tmp[C::NS];
#pragma unroll
	for (int s = 0; s < C::NS; s++)
		ae[s] = s ? P16_UNREACHABLE : 0u;
	uint32_t offA = 0, offB = 0;
	constexpr bool ER = chan_has_erasures(CH), G2 = CH == CH_RACH;
	constexpr bool RN = p16_needs_renorm(C::N, C::K, chan_len_of(CH) + C::K - 1);
	p16_forward<C, true, G2, ER, RN>(ae, lut, rowA, rowB, tb.g, tb.g2, 0, tb.len, dec, T, t, offA, offB);
	if (RN)
		p16_renorm<C>(ae, offA, offB);
	p16_flush_step<C, 0, G2, ER>(ae, tmp, lut, rowA, rowB, tb.g, tb.g2, tb.len, dec, T, t);
	p16_flush_step<C, 1, G2, ER>(tmp, ae, lut, rowA, rowB, tb.g, tb.g2, tb.len + 1, dec, T, t);
	p16_flush_step<C, 2, G2, ER>(ae, tmp, lut, rowA, rowB, tb.g, tb.g2, tb.len + 2, dec, T, t);
	p16_flush_step<C, 3, G2, ER>(tmp, ae, lut, rowA, rowB, tb.g, tb.g2, tb.len + 3, dec, T, t);

	if (a.conv) {
		a.conv[unitA] = (int32_t)((ae[0] & 0xffffu) + offA);
		if (okB)
			a.conv[unitB] = (int32_t)((ae[0] >> 16) + offB);
	}

	// traceback of both codewords in one walk (two independent chains of load -> shift -> state)
	uint8_t *outA = (uint8_t *)rowA, *outB = (uint8_t *)rowB;
	{
		unsigned sa = 0, sb = 0;
		auto back = [&](int i) {
			const uint32_t d = dec[(size_t)i * T + t];
			const unsigned ba = (d >> sa) & 1u, bb = (d >> (16 + sb)) & 1u;
			sa = (sa >> 1) | (ba << (C::K - 2));
			sb = (sb >> 1) | (bb << (C::K - 2));
		};
		for (int i = tb.n_steps - 1; i >= tb.len; i--)
			back(i);
		int i = tb.len - 1;
		unsigned acca = 0, accb = 0;
		for (; (i & 7) != 7; i--) {          // ragged top byte (len not a multiple of 8)
			acca |= (sa & 1u) << (i & 7);
			accb |= (sb & 1u) << (i & 7);
			back(i);
		}
		if ((tb.len & 7) != 0) {
			outA[tb.len >> 3] = (uint8_t)acca;
			outB[tb.len >> 3] = (uint8_t)accb;
		}
		// whole bytes: the decision word of a step does not depend on the state (one word per step for 16 states), so
		// the eight words of the NEXT byte are fetched while the dependent shift chain of this one runs
		uint32_t cur[8], nxt[8];
		if (i >= 7) {
#pragma unroll
			for (int b = 0; b < 8; b++)
				cur[b] = dec[(size_t)(i - 7 + b) * T + t];
		}
		for (; i >= 7; i -= 8) {
			if (i >= 15) {
#pragma unroll
				for (int b = 0; b < 8; b++)
					nxt[b] = dec[(size_t)(i - 15 + b) * T + t];
			}
			acca = accb = 0;
#pragma unroll
			for (int b = 7; b >= 0; b--) {
				acca |= (sa & 1u) << b;
				accb |= (sb & 1u) << b;
				const uint32_t d = cur[b];
				const unsigned ba = (d >> sa) & 1u, bb = (d >> (16 + sb)) & 1u;
				sa = (sa >> 1) | (ba << (C::K - 2));
				sb = (sb >> 1) | (bb << (C::K - 2));
			}
			outA[i >> 3] = (uint8_t)acca;
			outB[i >> 3] = (uint8_t)accb;
#pragma unroll
			for (int b = 0; b < 8; b++)
				cur[b] = nxt[b];
		}
	}
	k5_outputs<CH>(a, unitA, outA);
	if (okB)
		k5_outputs<CH>(a, unitB, outB);
}

// ---- per-unit decode, TCH3 (two tail-biting K7 frames + class-2 bits) --------------------------
template <bool LUT = false>
GMR1_HD void decode_unit_tch3(const TabRef &tb, const DecodeArgs &a, int unit,
                              const int8_t *row, uint32_t *dec, int T, int t, const RelLut *lut = nullptr)
{
	using C = CodeK7_12;
	if (a.bits_s)
		for (int i = 0; i < 4; i++)         // tch3.c:134-135
			a.bits_s[(size_t)unit * 4 + i] = a.ebits[(size_t)unit * 212 + 52 + i] < 0;

	for (int f = 0; f < 2; f++) {
		const uint16_t *g = tb.g + (2 * (a.tch3_m ? 1 : 0) + f) * 128;
		uint32_t ae[C::NS];
#pragma unroll
		for (int s = 0; s < C::NS; s++)
			ae[s] = s ? MAX_AE : 0u;
		// seeding pass, no history kept
		// Path metrics relative to the all-zero branch of every step (acs_step REL, signed comparisons): the
		// offset of the seeding pass drops out in the normalisation, the one of the second pass goes back into
		// the reported metric.  (After the normalisation the smallest metric is 0 < MAX_AE, so an end state always
		// exists: the reference's "no state" return cannot occur here.)
		uint32_t off = 0;
		forward<C, false, false, false, true, true, LUT>(ae, row, g, nullptr, 0, 48, dec, T, t, off, lut);
		int32_t mn = (int32_t)ae[0];
#pragma unroll
		for (int s = 1; s < C::NS; s++)
			mn = (int32_t)ae[s] < mn ? (int32_t)ae[s] : mn;
#pragma unroll
		for (int s = 0; s < C::NS; s++)
			ae[s] -= (uint32_t)mn;
		off = 0;
		forward<C, false, true, false, true, true, LUT>(ae, row, g, nullptr, 0, 48, dec, T, t, off, lut);
		// end state: first state with the minimal metric
		int32_t best = (int32_t)ae[0];
		unsigned end = 0;
#pragma unroll
		for (int s = 1; s < C::NS; s++)
			if ((int32_t)ae[s] < best) {
				best = (int32_t)ae[s];
				end = (unsigned)s;
			}
		int32_t *cv = f ? a.conv1 : a.conv;
		if (cv)
			cv[unit] = best + (int32_t)off;

		// frame bits 0..47 from the decoder, 48..79 = sign of c[72..103]; MSB-first packing
		uint32_t w0 = 0, w1 = 0, w2 = 0;        // bits 0..31, 32..63, 64..79
		if (end != 0xff) {
			auto emit = [&](int i, unsigned bit) {
				if (i < 32) w0 |= bit << (31 - i);
				else        w1 |= bit << (63 - i);
			};
			traceback<C>(dec, T, t, 48, 48, end, emit);
		}
		for (int j = 48; j < 80; j++) {
			const unsigned bit = gather_sbit(row, g[96 + (j - 48)]) < 0 ? 1u : 0u;
			if (j < 64) w1 |= bit << (63 - j);
			else        w2 |= bit << (95 - j);
		}
		uint8_t *fr = (f ? a.l2b : a.l2) + (size_t)unit * 10;
		fr[0] = (uint8_t)(w0 >> 24); fr[1] = (uint8_t)(w0 >> 16); fr[2] = (uint8_t)(w0 >> 8); fr[3] = (uint8_t)w0;
		fr[4] = (uint8_t)(w1 >> 24); fr[5] = (uint8_t)(w1 >> 16); fr[6] = (uint8_t)(w1 >> 8); fr[7] = (uint8_t)w1;
		fr[8] = (uint8_t)(w2 >> 24); fr[9] = (uint8_t)(w2 >> 16);
	}
}

// ---- TWO units per thread, TCH3 (viterbi_p16.cuh; 64 packed path metrics per thread) -------------------------------
// The one-unit form is bound by registers (143 per thread: 12 resident warps per SM); two units in the same registers
// double the work of every resident warp.  dec: [48][4][T] 32-bit words (word w of a step: states 16 w .. 16 w + 15).
GMR1_HD void decode_pair_tch3(const TabRef &tb, const DecodeArgs &a, const P16Lut *lut, int unitA, int unitB, bool okB,
                              const int8_t *rowA, const int8_t *rowB, uint32_t *dec, int T, int t)
{
	using C = CodeK7_12;
	if (a.bits_s)
		for (int i = 0; i < 4; i++) {       // tch3.c:134-135
			a.bits_s[(size_t)unitA * 4 + i] = a.ebits[(size_t)unitA * 212 + 52 + i] < 0;
			if (okB)
				a.bits_s[(size_t)unitB * 4 + i] = a.ebits[(size_t)unitB * 212 + 52 + i] < 0;
		}

	for (int f = 0; f < 2; f++) {
		const uint16_t *g = tb.g + (2 * (a.tch3_m ? 1 : 0) + f) * 128;
		uint32_t ae[C::NS];
#pragma unroll
		for (int s = 0; s < C::NS; s++)
			ae[s] = s ? P16_UNREACHABLE : 0u;
		uint32_t offA = 0, offB = 0;
		// seeding pass, no history kept; 48 steps add at most 12 192 to the sentinel: no renormalisation inside a pass
		p16_forward<C, false, false, true, false>(ae, lut, rowA, rowB, g, nullptr, 0, 48, dec, T, t, offA, offB);
		p16_renorm<C>(ae, offA, offB);          // osmo_conv_decode_rewind: the minimum comes off (its value is not reported)
		p16_forward<C, true, false, true, false>(ae, lut, rowA, rowB, g, nullptr, 0, 48, dec, T, t, offA, offB);
		// end state: FIRST state with the minimal metric, per codeword
		uint32_t mn = ae[0];
#pragma unroll
		for (int s = 1; s < C::NS; s++)
			mn = p16_minu(mn, ae[s]);
		const uint32_t bestA = mn & 0xffffu, bestB = mn >> 16;
		unsigned sa = 0, sb = 0;
#pragma unroll
		for (int s = C::NS - 1; s >= 0; s--) {
			// halves equal to the minimum: x = ae ^ mn has a zero half there
			const uint32_t x = ae[s] ^ mn;
			if ((x & 0xffffu) == 0)
				sa = (unsigned)s;
			if ((x >> 16) == 0)
				sb = (unsigned)s;
		}
		int32_t *cv = f ? a.conv1 : a.conv;
		if (cv) {
			cv[unitA] = (int32_t)bestA;
			if (okB)
				cv[unitB] = (int32_t)bestB;
		}

		// frame bits 0..47 from the decoder, 48..79 = sign of c[72..103]; MSB-first packing
		uint32_t wa0 = 0, wa1 = 0, wa2 = 0, wb0 = 0, wb1 = 0, wb2 = 0;
		// Which of a step's four decision words a codeword needs depends on its state, so a load per step would be a
		// chain of 48 dependent global loads: all four words of eight steps are requested together instead (32 loads
		// in flight, six waits per frame) and the word is picked by selects.
		for (int i0 = 40; i0 >= 0; i0 -= 8) {
			uint32_t dw[8][4];
#pragma unroll
			for (int b = 0; b < 8; b++)
#pragma unroll
				for (int w = 0; w < 4; w++)
					dw[b][w] = dec[(size_t)((i0 + b) * 4 + w) * T + t];
#pragma unroll
			for (int b = 7; b >= 0; b--) {
				const int i = i0 + b;
				const unsigned qa = sa >> 4, qb = sb >> 4;
				const uint32_t da = (qa & 2u) ? ((qa & 1u) ? dw[b][3] : dw[b][2]) : ((qa & 1u) ? dw[b][1] : dw[b][0]);
				const uint32_t db = (qb & 2u) ? ((qb & 1u) ? dw[b][3] : dw[b][2]) : ((qb & 1u) ? dw[b][1] : dw[b][0]);
				if (i < 32) {
					wa0 |= (sa & 1u) << (31 - i);
					wb0 |= (sb & 1u) << (31 - i);
				} else {
					wa1 |= (sa & 1u) << (63 - i);
					wb1 |= (sb & 1u) << (63 - i);
				}
				const unsigned ba = (da >> (sa & 15u)) & 1u, bb = (db >> (16u + (sb & 15u))) & 1u;
				sa = (sa >> 1) | (ba << (C::K - 2));
				sb = (sb >> 1) | (bb << (C::K - 2));
			}
		}
		for (int j = 48; j < 80; j++) {
			const uint16_t w = g[96 + (j - 48)];
			const unsigned bitA = gather_sbit(rowA, w) < 0 ? 1u : 0u, bitB = gather_sbit(rowB, w) < 0 ? 1u : 0u;
			if (j < 64) {
				wa1 |= bitA << (63 - j);
				wb1 |= bitB << (63 - j);
			} else {
				wa2 |= bitA << (95 - j);
				wb2 |= bitB << (95 - j);
			}
		}
		auto put = [&](int unit, uint32_t w0, uint32_t w1, uint32_t w2) {
			uint8_t *fr = (f ? a.l2b : a.l2) + (size_t)unit * 10;
			fr[0] = (uint8_t)(w0 >> 24); fr[1] = (uint8_t)(w0 >> 16); fr[2] = (uint8_t)(w0 >> 8); fr[3] = (uint8_t)w0;
			fr[4] = (uint8_t)(w1 >> 24); fr[5] = (uint8_t)(w1 >> 16); fr[6] = (uint8_t)(w1 >> 8); fr[7] = (uint8_t)w1;
			fr[8] = (uint8_t)(w2 >> 24); fr[9] = (uint8_t)(w2 >> 16);
		};
		put(unitA, wa0, wa1, wa2);
		if (okB)
			put(unitB, wb0, wb1, wb2);
	}
}

}  // namespace gmr1
