// tests/emu/decode_emu.cpp - TEST HARNESS: runs the product's __host__ __device__ decode code
// (osmo_gmr_b200/csrc/decode_unit.cuh, viterbi_tpc.cuh - the exact functions the CUDA kernels
// inline) on the CPU, one "thread" at a time, so that the kernel logic can be checked against
// the oracle without a GPU.  It is NOT part of the product library and nothing in the product
// falls back to it; it only exists so `pytest -m "not gpu"` can catch logic errors early.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "decode_unit.cuh"

using namespace gmr1;

// lut: one codeword per "thread" with the branch metrics through the byte tables (RelLut), as the kernels run it
template <int CH, bool LUT = false>
static void run(const DecodeArgs &a)
{
	static RelLut lut;
	for (int i = 0; i < 512; i++)
		rel_lut_fill(&lut, i);
	const ChanTab &t = chan_tab(CH);
	TabRef tb;
	tb.g = t.g; tb.g2 = (CH == CH_RACH) ? t.g2 : nullptr; tb.cmap = t.cmap; tb.t9_src = t.t9_src;
	tb.n_in = t.n_in; tb.n_row = t.n_row; tb.n_ciph = t.n_ciph; tb.n_steps = t.n_steps; tb.len = t.len;
	std::vector<int8_t> row(t.n_row + 16);
	std::vector<uint32_t> dec(2 * 1024);
	for (int u = 0; u < a.n; u++) {
		for (int r = 0; r < t.n_row; r++)
			row[r] = stage_elem<CH>(tb, a, u, r);
		if constexpr (CH == CH_TCH3)
			decode_unit_tch3<LUT>(tb, a, u, row.data(), dec.data(), 1, 0, &lut);
		else
			decode_unit_k5<CH, LUT>(tb, a, u, row.data(), (uint16_t *)dec.data(), 1, 0, &lut);
	}
}

// two units per "thread" (viterbi_p16.cuh): units (u, u + 1), a ragged last unit runs with an all-zero partner row
template <int CH>
static void run_p16(const DecodeArgs &a)
{
	const ChanTab &t = chan_tab(CH);
	TabRef tb;
	tb.g = t.g; tb.g2 = (CH == CH_RACH) ? t.g2 : nullptr; tb.cmap = t.cmap; tb.t9_src = t.t9_src;
	tb.n_in = t.n_in; tb.n_row = t.n_row; tb.n_ciph = t.n_ciph; tb.n_steps = t.n_steps; tb.len = t.len;
	std::vector<int8_t> rowA(t.n_row + 16), rowB(t.n_row + 16);
	std::vector<uint32_t> dec(2 * 1024);
	static P16Lut lut;
	for (int i = 0; i < 512; i++)
		(&lut.plain[0])[i] = p16_lut_word(i);
	for (int u = 0; u < a.n; u += 2) {
		const bool okB = u + 1 < a.n;
		for (int r = 0; r < t.n_row; r++) {
			rowA[r] = stage_elem<CH>(tb, a, u, r);
			rowB[r] = okB ? stage_elem<CH>(tb, a, u + 1, r) : (int8_t)0;
		}
		if constexpr (CH == CH_TCH3)
			decode_pair_tch3(tb, a, &lut, u, u + 1, okB, rowA.data(), rowB.data(), dec.data(), 1, 0);
		else
			decode_pair_k5<CH>(tb, a, &lut, u, u + 1, okB, rowA.data(), rowB.data(), dec.data(), 1, 0);
	}
}

extern "C" int gmr1_emu_decode_p16(int ch, const DecodeArgs *a)
{
	switch (ch) {
	case CH_BCCH:     run_p16<CH_BCCH>(*a); break;
	case CH_CCCH:     run_p16<CH_CCCH>(*a); break;
	case CH_FACCH3:   run_p16<CH_FACCH3>(*a); break;
	case CH_FACCH9:   run_p16<CH_FACCH9>(*a); break;
	case CH_TCH9_2K4: run_p16<CH_TCH9_2K4>(*a); break;
	case CH_TCH9_4K8: run_p16<CH_TCH9_4K8>(*a); break;
	case CH_TCH9_9K6: run_p16<CH_TCH9_9K6>(*a); break;
	case CH_RACH:     run_p16<CH_RACH>(*a); break;
	case CH_TCH3:     run_p16<CH_TCH3>(*a); break;
	default: return -1;
	}
	return 0;
}

extern "C" int gmr1_emu_decode_lut(int ch, const DecodeArgs *a)
{
	switch (ch) {
	case CH_BCCH:     run<CH_BCCH, true>(*a); break;
	case CH_CCCH:     run<CH_CCCH, true>(*a); break;
	case CH_FACCH3:   run<CH_FACCH3, true>(*a); break;
	case CH_FACCH9:   run<CH_FACCH9, true>(*a); break;
	case CH_TCH9_2K4: run<CH_TCH9_2K4, true>(*a); break;
	case CH_TCH9_4K8: run<CH_TCH9_4K8, true>(*a); break;
	case CH_TCH9_9K6: run<CH_TCH9_9K6, true>(*a); break;
	case CH_RACH:     run<CH_RACH, true>(*a); break;
	case CH_TCH3:     run<CH_TCH3, true>(*a); break;
	default: return -1;
	}
	return 0;
}

extern "C" int gmr1_emu_decode(int ch, const DecodeArgs *a)
{
	switch (ch) {
	case CH_BCCH:     run<CH_BCCH>(*a); break;
	case CH_CCCH:     run<CH_CCCH>(*a); break;
	case CH_FACCH3:   run<CH_FACCH3>(*a); break;
	case CH_FACCH9:   run<CH_FACCH9>(*a); break;
	case CH_TCH9_2K4: run<CH_TCH9_2K4>(*a); break;
	case CH_TCH9_4K8: run<CH_TCH9_4K8>(*a); break;
	case CH_TCH9_9K6: run<CH_TCH9_9K6>(*a); break;
	case CH_RACH:     run<CH_RACH>(*a); break;
	case CH_TCH3:     run<CH_TCH3>(*a); break;
	default: return -1;
	}
	return 0;
}

extern "C" int gmr1_emu_keep_mask(int ch, uint8_t *mask, int max) { return chan_keep_mask(ch, mask, max); }
extern "C" int gmr1_emu_code_output(int ch, int state, int bit) { return code_output(chan_code(ch), state, bit); }
extern "C" int gmr1_emu_sizeof_args() { return (int)sizeof(DecodeArgs); }

// ---- TCH3 burst-loop state machine (osmo_gmr_b200/csrc/tch3_state.cuh) -------------------------------
#include "tch3_state.cuh"
extern "C" int gmr1_emu_tch3_sizeof() { return (int)sizeof(Tch3State); }
extern "C" void gmr1_emu_tch3_init(Tch3State *s, int8_t *eb, const uint8_t *imm_ass, float ref) { tch3_init(*s, eb, imm_ass, ref); }
extern "C" int gmr1_emu_tch3_gate(Tch3State *s, float be) { return tch3_gate(*s, be); }
extern "C" int gmr1_emu_tch3_dkab_result(Tch3State *s, float be, int rv) { return tch3_dkab_result(*s, be, rv); }
extern "C" int gmr1_emu_tch3_facch_flush_before(const Tch3State *s, int sync_id) { return tch3_facch_flush_before(*s, sync_id); }
extern "C" int gmr1_emu_tch3_facch_store(Tch3State *s, int8_t *eb, const int8_t *b, int sync_id, uint32_t fn) { return tch3_facch_store(*s, eb, b, sync_id, fn); }
extern "C" int gmr1_emu_tch3_flush_first_try_ciphered(const Tch3State *s) { return tch3_flush_first_try_ciphered(*s); }
extern "C" int gmr1_emu_tch3_flush_wants_retry(const Tch3State *s, int crc) { return tch3_flush_wants_retry(*s, crc); }
extern "C" int gmr1_emu_tch3_flush_done(Tch3State *s, int8_t *eb, int crc, int retried) { return tch3_flush_done(*s, eb, crc, retried != 0); }
extern "C" int gmr1_emu_tch9_init_from_facch3(Tch9State *s, const uint8_t *l2, int crc_ok) { return tch9_init_from_facch3(*s, l2, crc_ok != 0); }
extern "C" int gmr1_emu_tch9_is_facch9(int sync_id) { return tch9_is_facch9(sync_id); }
extern "C" int gmr1_emu_tch9_avg_magnitude(const int8_t *eb) { return tch9_avg_magnitude(eb); }
