// burst_formats.h - the ten pi/4-CxPSK burst formats of the reference (src/sdr/nb.c:34-377, ETSI TS 101 376-5-2
// section 7.4) as ONE constexpr table: gmr1_tables.cpp flattens it into the run-time BurstTab descriptors, and the
// per-format demodulation kernels (demod_fast.cu) read it at compile time, so chunk positions, tap counts and symbol
// lists are immediates in their code.
#pragma once
#include "gmr1_tables.h"

namespace gmr1 {

struct SyncDef { int pos; const char *syms; };          // syms: one digit per symbol, "" ends list
struct DataDef { int pos, len; };
struct BurstDef {
	float rot_div;  // rotation = pi / rot_div
	int nbits, len, ebits;
	SyncDef sync[MAX_SYNC][MAX_SYNC_CHUNK];
	DataDef data[MAX_DATA_CHUNK];
};

static constexpr char S32x2[] = "22222222222222222222222222222222";

static constexpr BurstDef BURSTS[BT_COUNT] = {
	/* BCCH  */ {4, 2, 234, 424, {{{28, "02200020222"}, {119, "220"}, {197, "220"}}},
	             {{2, 26}, {39, 80}, {122, 75}, {200, 31}}},
	/* DC2   */ {4, 2, 78, 132, {{{28, "0123030"}}}, {{2, 26}, {35, 40}}},
	/* DC6   */ {4, 2, 234, 432, {{{28, "0002202"}, {119, "030"}, {197, "311"}}},
	             {{2, 26}, {35, 84}, {122, 75}, {200, 31}}},
	/* DC12  */ {2, 1, 468, 432, {{{10, "0010001111"}, {228, "00100011101"}, {447, "0010001111"}}},
	             {{2, 8}, {20, 208}, {239, 208}, {457, 8}}},
	/* NT3 S */ {4, 2, 117, 212, {{{28, "033123"}}}, {{2, 26}, {34, 80}}},
	/* NT3 F */ {4, 1, 117, 104, {{{28, "10101010"}}, {{28, "11001001"}}}, {{2, 26}, {36, 78}}},
	/* NT6   */ {4, 2, 234, 434,
	             {{{28, "022323"}, {119, "010"}, {197, "230"}}, {{28, "000220"}, {119, "130"}, {197, "213"}}},
	             {{2, 26}, {34, 85}, {122, 75}, {200, 31}}},
	/* NT9   */ {4, 2, 351, 662,
	             {{{28, "022323"}, {119, "122"}, {197, "010"}, {275, "230"}},
	              {{28, "000220"}, {119, "020"}, {197, "130"}, {275, "213"}}},
	             {{2, 26}, {34, 85}, {122, 75}, {200, 75}, {278, 70}}},
	/* RACH  */ {4, 2, 351, 494,
	             {{{78, "02200020222220220"}, {127, S32x2}, {191, S32x2}, {255, "02200020222220220"}, {347, "0"}}},
	             {{2, 76}, {95, 32}, {159, 32}, {223, 32}, {272, 75}}},
	/* SDCCH */ {4, 1, 234, 208,
	             {{{28, "0101010"}, {115, "1010101"}, {197, "0101011"}},
	              {{28, "0011001"}, {115, "1001100"}, {197, "1100111"}},
	              {{28, "0000111"}, {115, "1000011"}, {197, "1100001"}},
	              {{28, "0110100"}, {115, "1011010"}, {197, "0101101"}}},
	             {{2, 26}, {35, 80}, {122, 75}, {204, 27}}},
};


// ---- compile-time accessors (usable in device code as constant expressions) ----
constexpr int bf_strlen(const char *s) { int n = 0; if (s) while (s[n]) n++; return n; }
constexpr int bf_n_sync(int bt)
{
	int n = 0;
	for (int i = 0; i < MAX_SYNC; i++)
		if (BURSTS[bt].sync[i][0].syms) n = i + 1; else break;
	return n;
}
constexpr int bf_n_chunk(int bt, int s)
{
	int n = 0;
	for (int c = 0; c < MAX_SYNC_CHUNK; c++)
		if (BURSTS[bt].sync[s][c].syms) n = c + 1; else break;
	return n;
}
constexpr int bf_s_pos(int bt, int s, int c) { return BURSTS[bt].sync[s][c].pos; }
constexpr int bf_s_len(int bt, int s, int c) { return bf_strlen(BURSTS[bt].sync[s][c].syms); }
// symbol index 0..3 (phase k*pi/2); 1 bit/symbol formats list bits: bit 1 is the phase-pi point
constexpr int bf_s_sym(int bt, int s, int c, int k)
{
	const int v = BURSTS[bt].sync[s][c].syms[k] - '0';
	return BURSTS[bt].nbits == 1 ? 2 * v : v;
}
constexpr int bf_n_data(int bt)
{
	int n = 0;
	for (int c = 0; c < MAX_DATA_CHUNK; c++)
		if (BURSTS[bt].data[c].len) n = c + 1; else break;
	return n;
}

}  // namespace gmr1
