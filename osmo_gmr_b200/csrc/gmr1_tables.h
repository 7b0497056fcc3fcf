// gmr1_tables.h - constant tables of the GMR-1 receive path, laid out for the GPU.
//
// Everything a decode kernel needs besides the soft bits is a *gather program*: the
// reference's chain  demux -> decipher -> descramble -> deinterleave -> depuncture
// (src/l1/bcch.c:84-103, ccch.c:88-107, facch3.c:122-170, facch9.c:107-144, tch3.c:124-183,
// tch9.c:140-175, rach.c:137-196, xch_dc12.c:87-106 of the reference) is a fixed permutation
// with sign flips, so it is flattened at library load into one uint16 per coded bit.  The
// kernels then never run those stages as separate passes.
#pragma once
#include <stdint.h>

namespace gmr1 {

// ---- channel (decode program) ids -------------------------------------------------------
enum Chan : int {
	CH_BCCH = 0,      // K5 r1/2 flush  208 bits, CRC16, 424 ebits   (bcch.c)
	CH_CCCH,          // same, 432 ebits with 4+4 pad               (ccch.c)
	CH_FACCH3,        // K5 r1/4 flush   92 bits, CRC16, 4x104 ebits (facch3.c)
	CH_FACCH9,        // K5 r1/2 flush  316 bits, CRC16, 662 ebits   (facch9.c)
	CH_TCH9_2K4,      // K5 r1/5 flush  144 bits punctured, 662 ebits x3 bursts (tch9.c)
	CH_TCH9_4K8,      // K5 r1/3 flush  240 bits punctured
	CH_TCH9_9K6,      // K5 r1/2 flush  480 bits punctured
	CH_RACH,          // K5 r1/4 flush  159 bits punctured, CRC8+CRC12, 494 ebits (rach.c)
	CH_TCH3,          // K7 r1/2 tail-biting 48 bits punctured, x2 frames, 212 ebits (tch3.c)
	CH_DC12,          // K9 r1/3 tail-biting 208 bits punctured, CRC16, 432 ebits (xch_dc12.c)
	CH_COUNT
};

// gather word: [9:0] index into the staged row, [14] sign flip (scrambler), 0xffff = erased
// (punctured) position that feeds a 0 into the Viterbi.
static constexpr uint16_t G_ERASED = 0xffff;
static constexpr uint16_t G_FLIP   = 0x4000;
static constexpr uint16_t G_IDX    = 0x03ff;

static constexpr int MAX_CODED = 1024;   // >= 968 (TCH9 9k6), 652 (RACH), 640 (FACCH9)
static constexpr int MAX_EBITS = 672;    // >= 662 (NT9)

struct ChanTab {
	int32_t n_in;        // soft bits per decode unit as the reference API takes them
	int32_t n_row;       // bytes per staged row in shared memory (== n_in except TCH9: 648)
	int32_t N, K, len;   // code rate 1/N, constraint length, data bits
	int32_t flush;       // 1: CONV_TERM_FLUSH, 0: CONV_TERM_TAIL_BITING
	int32_t n_steps;     // len (+K-1 when flush)
	int32_t n_ciph;      // cipher bits per unit (0 = channel is never ciphered)
	uint16_t g[MAX_CODED];    // gather program, one word per unpunctured coded bit
	uint16_t g2[MAX_CODED];   // second source (RACH class-1 soft averaging), G_ERASED if none
	int16_t  cmap[MAX_EBITS]; // ebit index -> cipher bit index, -1 = not ciphered
	// TCH9 only: staged row r (0..647, scrambled/deinterleaved-inter domain):
	uint16_t t9_src[648];     // [9:0] ebit index, [11:10] burst age (0 = current, 1, 2), [14] flip
};

// generator polynomials (bit i = D^i) per code, rate order g0..g(N-1)
struct CodePoly { int N, K; uint16_t g[5]; };

// ---- pi/4-CxPSK burst descriptors, flattened -------------------------------------------
static constexpr int MAX_SYNC       = 4;    // GMR1_MAX_SYNC   (reference sdr/pi4cxpsk.h:39)
static constexpr int MAX_SYNC_CHUNK = 6;
static constexpr int MAX_SYNC_SYMS  = 32;   // GMR1_MAX_SYNC_SYMS
static constexpr int MAX_DATA_CHUNK = 6;

struct BurstTab {
	float   rotation;     // per-symbol rotation (pi/4 or pi/2)
	int32_t nbits;        // bits per symbol (1 or 2)
	int32_t len;          // symbols incl. guard
	int32_t ebits;        // soft bits produced
	int32_t n_sync;       // number of alternative sync sequences
	int32_t n_chunk[MAX_SYNC];
	int16_t s_pos[MAX_SYNC][MAX_SYNC_CHUNK];
	int16_t s_len[MAX_SYNC][MAX_SYNC_CHUNK];
	uint8_t s_sym[MAX_SYNC][MAX_SYNC_CHUNK][MAX_SYNC_SYMS];  // symbol index 0..3 (phase k*pi/2)
	int32_t n_data;
	int16_t d_pos[MAX_DATA_CHUNK];
	int16_t d_len[MAX_DATA_CHUNK];
};

enum BurstId : int {
	BT_BCCH = 0, BT_DC2, BT_DC6, BT_DC12, BT_NT3_SPEECH, BT_NT3_FACCH, BT_NT6, BT_NT9,
	BT_RACH, BT_SDCCH, BT_COUNT
};

// ---- host-side builders (gmr1_tables.cpp) ----------------------------------------------
const ChanTab  &chan_tab(int ch);
const CodePoly &chan_code(int ch);
const BurstTab &burst_tab(int bt);
// next_output[s][b] of a code, reference convention (src/l1/conv.c): MSB = g0
uint8_t code_output(const CodePoly &c, int state, int bit);
// scrambler sequence (src/l1/scramb.c:39-52): bit i, 1 = flip
const uint8_t *scramble_seq();   // 1024 entries
// keep-mask (1 = transmitted) over the unpunctured coded bits of a channel
int chan_keep_mask(int ch, uint8_t *mask, int max);

}  // namespace gmr1
