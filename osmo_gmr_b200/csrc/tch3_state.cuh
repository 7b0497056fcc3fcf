// tch3_state.cuh - per-channel state machine of the reference receiver's TCH3 burst loop, as __host__ __device__
// functions.  Replaces the scalar decisions of
//   rx_tch3_init          src/gmr1_rx.c:362-381   state after an IMMEDIATE ASSIGNMENT
//   rx_tch3               :538-600                energy gate DKAB / burst, running energy averages, release
//   _rx_tch3_facch        :455-494                FACCH3 codeword assembly over four frames
//   _rx_tch3_facch_flush  :394-452                plain / ciphered decode attempts, cipher discovery, group reset
//   rx_tch9_init, rx_tch9 :264-355                hand-off to the TCH9 loop by an ASSIGNMENT COMMAND 1, FACCH9 / TCH9 split
// The signal processing between the decisions (burst_energy, gmr1_dkab_demod, gmr1_pi4cxpsk_detect / _demod, gmr1_a5,
// gmr1_facch3_decode, gmr1_tch3_decode) is the batched kernels of this library.  These functions are what the state
// kernels of the device-side loop run per channel; tests/test_tch3_state_emu.py runs this very code on the CPU, next to
// the reference's signal-processing functions, against the reference application on a recorded call.
#pragma once
#include <stdint.h>

#ifndef GMR1_HD
#ifdef __CUDACC__
#define GMR1_HD __host__ __device__ __forceinline__
#else
#define GMR1_HD inline
#endif
#endif

namespace gmr1 {

struct Tch3State {                     // struct tch3_state, gmr1_rx.c:59-79 (the 4 x 104 soft bits live beside it)
	int32_t  active, tn, p, ciph;
	float    energy_dkab, energy_burst;
	int32_t  weak_cnt, sync_id, burst_cnt;
	uint32_t bi_fn[4];
};

enum { TCH3_GATE_DKAB = 1, TCH3_GATE_BURST = 2 };

// a * x + b * y with each operation rounded (the reference is compiled without contraction; on the device the
// compiler would fuse the second product into an FMA)
GMR1_HD float tch3_mix(float a, float x, float b, float y)
{
#ifdef __CUDA_ARCH__
	return __fadd_rn(__fmul_rn(a, x), __fmul_rn(b, y));
#else
	volatile float u = a * x, v = b * y;
	return u + v;
#endif
}

// rx_tch3_init(cd, imm_ass, ref_energy), :362-381.  `ciph` is not touched (the reference leaves it as it is).
// ebits [4][104] is the channel's FACCH3 soft-bit store.
GMR1_HD void tch3_init(Tch3State &s, int8_t *ebits, const uint8_t *imm_ass, float ref_energy)
{
	s.active = 1;
	s.p = (imm_ass[8] & 0xfc) >> 2;                                   // ccch_imm_ass_parse :240-245
	s.tn = ((imm_ass[8] & 0x03) << 3) | (imm_ass[9] >> 5);
	s.energy_burst = ref_energy * 0.75f;
	s.energy_dkab = s.energy_burst / 8.0f;
	s.weak_cnt = 0;
	s.sync_id = 0;
	for (int i = 0; i < 4 * 104; i++)
		ebits[i] = 0;
}

// the energy gate of rx_tch3, :552-585: a window weaker than (energy_dkab + energy_burst) / 4 is looked at as a DKAB,
// anything else is a burst and feeds the burst average
GMR1_HD int tch3_gate(Tch3State &s, float be)
{
	const float det = (s.energy_dkab + s.energy_burst) / 4.0f;
	if (be < det)
		return TCH3_GATE_DKAB;
	s.weak_cnt = 0;
	s.energy_burst = tch3_mix(0.1f, be, 0.9f, s.energy_burst);
	return TCH3_GATE_BURST;
}

// after gmr1_dkab_demod on a weak window (:558-576): rv 1 = nothing there, the tenth such frame in a row releases the
// channel (returns 1); rv 0 = a DKAB, which feeds the DKAB average
GMR1_HD int tch3_dkab_result(Tch3State &s, float be, int rv)
{
	if (rv < 0)
		return 0;
	if (rv == 1) {
		if (s.weak_cnt++ > 8) {
			s.active = 0;
			return 1;
		}
	} else {
		s.energy_dkab = tch3_mix(0.1f, be, 0.9f, s.energy_dkab);
	}
	return 0;
}

// _rx_tch3_facch :479-481: a FACCH3 burst whose sync sequence differs from the group's closes the group first
GMR1_HD bool tch3_facch_flush_before(const Tch3State &s, int sync_id) { return sync_id != s.sync_id; }

// _rx_tch3_facch :483-491: store the burst at its place in the codeword; true when the codeword is complete
GMR1_HD bool tch3_facch_store(Tch3State &s, int8_t *ebits, const int8_t *burst_ebits, int sync_id, uint32_t fn)
{
	const int bi = (int)(fn & 3u);
	for (int i = 0; i < 104; i++)
		ebits[104 * bi + i] = burst_ebits[i];
	s.sync_id = sync_id;
	s.bi_fn[bi] = fn;
	s.burst_cnt += 1;
	return s.burst_cnt == 4;
}

// _rx_tch3_facch_flush :417-430: a failed plain attempt on a channel not known to be ciphered is retried with the
// A5 masks of the four frame numbers in bi_fn
GMR1_HD bool tch3_flush_first_try_ciphered(const Tch3State &s) { return s.ciph != 0; }
GMR1_HD bool tch3_flush_wants_retry(const Tch3State &s, int crc_first) { return !s.ciph && crc_first != 0; }

// the end of _rx_tch3_facch_flush (:427-428, :445-449): crc = result of the last attempt, retried = whether that
// was the ciphered retry.  Returns true when the L2 message is good (to GSMTAP, ASS.CMD check).
GMR1_HD bool tch3_flush_done(Tch3State &s, int8_t *ebits, int crc, bool retried)
{
	if (retried && !crc)
		s.ciph = 1;
	s.sync_id ^= 1;
	s.burst_cnt = 0;
	for (int i = 0; i < 4; i++)
		s.bi_fn[i] = 0xffffffffu;
	for (int i = 0; i < 4 * 104; i++)
		ebits[i] = 0;
	return crc == 0;
}

// ---- hand-off to the TCH9 burst loop (rx_tch9_init :264-275, facch3_is_ass_cmd_1 / facch3_ass_cmd_1_parse :247-257) ----
struct Tch9State {                     // struct tch9_state, gmr1_rx.c:81-90 (the interleaver history lives beside it)
	int32_t active, tn;
};

GMR1_HD bool facch3_is_ass_cmd_1(const uint8_t *l2) { return l2[3] == 0x06 && l2[4] == 0x2e; }

// a good FACCH3 message that is an ASSIGNMENT COMMAND 1 starts the TCH9 loop on the timeslot it names (the caller
// also resets the channel's depth-3 interleaver history, gmr1_interleaver_init :273)
GMR1_HD bool tch9_init_from_facch3(Tch9State &s, const uint8_t *l2, bool crc_ok)
{
	if (!crc_ok || !facch3_is_ass_cmd_1(l2))
		return false;
	s.active = 1;
	s.tn = ((l2[5] & 0x03) << 3) | (l2[6] >> 5);
	return true;
}

// rx_tch9 :305-352: sync sequence 0 of the NT9 burst marks a FACCH9 message, anything else a TCH9 block; `avg` is the
// mean soft-bit magnitude the reference prints with a TCH9 block
GMR1_HD bool tch9_is_facch9(int sync_id) { return sync_id == 0; }

GMR1_HD int tch9_avg_magnitude(const int8_t *ebits)
{
	int s = 0;
	for (int i = 0; i < 662; i++)
		s += ebits[i] < 0 ? -(int)ebits[i] : (int)ebits[i];
	return s / 662;
}

}  // namespace gmr1
