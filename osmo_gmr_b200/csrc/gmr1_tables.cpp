// gmr1_tables.cpp - builds the flattened gather programs / code tables / burst descriptors.
//
// Sources of the facts encoded here (all in the reference, paths relative to its root):
//   generator polynomials ........ src/l1/conv.c:124-127,149-153,175-180,202-208,346-350,521-522
//   puncturing masks ............. src/l1/punct.c:137-173,239-247,389-427,448-481,1105-1125
//   puncture expansion rule ...... src/l1/punct.c:48-133 (pre / repeated main / post blocks)
//   scrambler LFSR ............... src/l1/scramb.c:39-52
//   intra-burst interleaver ...... src/l1/interleave.c:49-90
//   inter-burst interleaver ...... src/l1/interleave.c:168-190 (closed form)
//   per-channel bit plumbing ..... src/l1/{bcch,ccch,facch3,facch9,tch3,tch9,rach,xch_dc12}.c
//   burst formats ................ src/sdr/nb.c:34-377 (ETSI TS 101 376-5-2 section 7.4)
// Nothing here is executed per burst: tables are built once and uploaded to __constant__.
#include "gmr1_tables.h"
#include "burst_formats.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace gmr1 {

std::mutex &init_mutex();
std::mutex &init_mutex()
{
	static std::mutex m;
	return m;
}

// ---------------------------------------------------------------------------- codes

static const CodePoly POLY_K5_12 = {2, 5, {0x19, 0x17}};                   // 1+D3+D4 ; 1+D+D2+D4
static const CodePoly POLY_K5_13 = {3, 5, {0x15, 0x1b, 0x1f}};             // 1+D2+D4 ; 1+D+D3+D4 ; 1+D+D2+D3+D4
static const CodePoly POLY_K5_14 = {4, 5, {0x19, 0x17, 0x15, 0x1f}};
static const CodePoly POLY_K5_15 = {5, 5, {0x15, 0x1b, 0x1f, 0x1d, 0x17}};
static const CodePoly POLY_K7_12 = {2, 7, {0x6d, 0x4f}};                   // 1+D2+D3+D5+D6 ; 1+D+D2+D3+D6
static const CodePoly POLY_K9_13 = {3, 9, {0x1ed, 0x19b, 0x127}};

uint8_t code_output(const CodePoly &c, int state, int bit)
{
	unsigned reg = ((unsigned)state << 1) | (unsigned)bit, ov = 0;
	for (int j = 0; j < c.N; j++)
		ov = (ov << 1) | (unsigned)(__builtin_popcount(reg & c.g[j]) & 1);
	return (uint8_t)ov;
}

const CodePoly &chan_code(int ch)
{
	switch (ch) {
	case CH_FACCH3: case CH_RACH: return POLY_K5_14;
	case CH_TCH9_2K4:             return POLY_K5_15;
	case CH_TCH9_4K8:             return POLY_K5_13;
	case CH_TCH3:                 return POLY_K7_12;
	case CH_DC12:                 return POLY_K9_13;
	default:                      return POLY_K5_12;
	}
}

// ---------------------------------------------------------------------------- primitives

static uint8_t g_scr[1024];

static void build_scrambler()
{
	uint16_t r = 0x4d4b;
	for (int i = 0; i < 1024; i++) {
		int b = ((r >> 14) ^ r) & 1;
		r = (uint16_t)((r << 1) | b);
		g_scr[i] = (uint8_t)b;
	}
}

static inline int kep_intra(int n, int kc) { return n * ((5 * kc) & 7) + (kc >> 3); }

// puncture block: mask over L*N coded bits, '1' = kept
struct Punct { int L; const char *mask; };

static const Punct P_K5_12_P23  = {3, "01" "10" "11"};
static const Punct P_K5_12_P25  = {5, "10" "11" "10" "11" "11"};
static const Punct P_K5_12_PS25 = {5, "11" "11" "10" "11" "10"};
static const Punct P_K5_12_P12  = {2, "11" "10"};
static const Punct P_K5_13_P25  = {5, "111" "111" "101" "111" "101"};
static const Punct P_K5_13_P15  = {5, "101" "111" "111" "111" "111"};
static const Punct P_K5_13_PS15 = {5, "111" "111" "111" "111" "101"};
static const Punct P_K5_15_P23  = {3, "11111" "11011" "11110"};
static const Punct P_K5_15_P53  = {3, "11101" "10011" "11100"};
static const Punct P_K5_15_PS53 = {3, "11100" "10011" "11101"};
static const Punct P_K9_13_P1213 = {13, "110" "101" "011" "110" "101" "011" "110" "101" "011" "110" "101" "011" "111"};

// keep[] over `total` coded bits: optional pre block, main block repeated (`repeat` times, 0 =
// until the end), optional post block on the last positions
static void expand_punct(std::vector<uint8_t> &keep, int N, int total,
                         const Punct *pre, const Punct *mainp, const Punct *post, int repeat)
{
	keep.assign(total, 1);
	int ii = 0, end = total;
	if (pre)
		for (int ip = 0; ii < total && ip < pre->L * N; ii++, ip++)
			keep[ii] = pre->mask[ip] == '1';
	if (post)
		end -= post->L * N;
	int d = mainp->L * N;
	if (!repeat) {
		int cl = total - (pre ? pre->L * N : 0) - (post ? post->L * N : 0);
		repeat = (cl + d - 1) / d;
	}
	for (int i = 0; i < repeat; i++)
		for (int ip = 0; ii < end && ip < d; ii++, ip++)
			keep[ii] = mainp->mask[ip] == '1';
	if (post) {
		ii = end;
		for (int ip = 0; ip < post->L * N && ii < total; ii++, ip++)
			keep[ii] = post->mask[ip] == '1';
	}
}

// ---------------------------------------------------------------------------- channels

static ChanTab g_chan[CH_COUNT];
static std::vector<uint8_t> g_keep[CH_COUNT];
static std::once_flag g_once;

static void init_tab(ChanTab &t, int n_in, int n_row, const CodePoly &c, int len, int flush, int n_ciph)
{
	memset(&t, 0, sizeof(t));
	t.n_in = n_in;
	t.n_row = n_row;
	t.N = c.N;
	t.K = c.K;
	t.len = len;
	t.flush = flush;
	t.n_steps = len + (flush ? c.K - 1 : 0);
	t.n_ciph = n_ciph;
	for (int i = 0; i < MAX_CODED; i++)
		t.g[i] = t.g2[i] = G_ERASED;
	for (int i = 0; i < MAX_EBITS; i++)
		t.cmap[i] = -1;
}

static inline uint16_t gw(int idx, int flip) { return (uint16_t)(idx | (flip ? G_FLIP : 0)); }

// map the k-th *received* coded bit through keep[] onto the unpunctured positions
template <class F>
static void fill_punctured(ChanTab &t, const std::vector<uint8_t> &keep, F recv_word)
{
	int q = 0;
	for (size_t k = 0; k < keep.size(); k++)
		if (keep[k])
			t.g[k] = recv_word(q++);
}

// TCH3 soft bit kc (0..103) of frame f under mux mode m: index into e[212] + scrambler flip
static uint16_t tch3_gather(int m, int f, int kc)
{
	int ii = kc % 24, ij = kc / 24;
	int kep = (ii < 8) ? (ij + 5 * ii) : (ij + 4 * ii + 8);
	int x = m ? (104 * f + kep) : (2 * kep + f);        // index into bits_epp / bits_xmy
	int s = x < 52 ? x : x + 4;
	return gw(s, g_scr[x]);
}

static void build_all()
{
	build_scrambler();
	const uint8_t *scr = g_scr;

	// BCCH: e[424] -descramble-> -deinterleave(53)-> K5 r1/2 (bcch.c:91-94)
	{
		ChanTab &t = g_chan[CH_BCCH];
		init_tab(t, 424, 424, POLY_K5_12, 208, 1, 0);
		for (int kc = 0; kc < 424; kc++) {
			int s = kep_intra(53, kc);
			t.g[kc] = gw(s, scr[s]);
		}
		g_keep[CH_BCCH].assign(424, 1);
	}
	// CCCH: 432 ebits, 4 pad bits either side of the interleaved block (ccch.c:95-98)
	{
		ChanTab &t = g_chan[CH_CCCH];
		init_tab(t, 432, 432, POLY_K5_12, 208, 1, 0);
		for (int kc = 0; kc < 424; kc++) {
			int s = 4 + kep_intra(53, kc);
			t.g[kc] = gw(s, scr[s]);
		}
		g_keep[CH_CCCH].assign(424, 1);
	}
	// FACCH3: 4 bursts x 104; status bits 22..29 removed; per-burst scrambler restart;
	// per-burst deinterleave(12); 4-way mux c[i] = cp[(i&3)*96 + (i>>2)] (facch3.c:132-158)
	{
		ChanTab &t = g_chan[CH_FACCH3];
		init_tab(t, 416, 416, POLY_K5_14, 92, 1, 384);
		for (int i = 0; i < 384; i++) {
			int b = i & 3, kc = i >> 2;
			int j = kep_intra(12, kc);               // index inside xmy / ep of burst b
			int s = 104 * b + (j < 22 ? j : j + 8);
			t.g[i] = gw(s, scr[j]);
		}
		for (int b = 0; b < 4; b++)
			for (int j = 0; j < 96; j++)
				t.cmap[104 * b + (j < 22 ? j : j + 8)] = (int16_t)(96 * b + j);
		g_keep[CH_FACCH3].assign(384, 1);
	}
	// FACCH9: status e[52..55] out, decipher 658, sacch my[52..61] out, descramble 648,
	// deinterleave(80) of epp_x+4 (facch9.c:117-134)
	{
		ChanTab &t = g_chan[CH_FACCH9];
		init_tab(t, 662, 662, POLY_K5_12, 316, 1, 658);
		for (int kc = 0; kc < 640; kc++) {
			int x = 4 + kep_intra(80, kc);
			int m = x < 52 ? x : x + 10;
			int s = m < 52 ? m : m + 4;
			t.g[kc] = gw(s, scr[x]);
		}
		for (int m = 0; m < 658; m++)
			t.cmap[m < 52 ? m : m + 4] = (int16_t)m;
		g_keep[CH_FACCH9].assign(640, 1);
	}
	// TCH9 (3 rates): same demux as FACCH9, then inter-burst deinterleave (depth 3, width 648)
	// out_n[x] = in_{n-(2-(x%3))}[x], intra deinterleave(81), punctured K5 code (tch9.c:55-79,
	// 150-170).  Staged row r is the inter-deinterleaved vector; g[] = deinterleave(81)+depuncture.
	{
		struct { int ch; const CodePoly *c; int len; const Punct *pre, *mainp, *post; int rep; } v[3] = {
			{CH_TCH9_2K4, &POLY_K5_15, 144, &P_K5_15_P53, &P_K5_15_P23, &P_K5_15_PS53, 41},
			{CH_TCH9_4K8, &POLY_K5_13, 240, &P_K5_13_P15, &P_K5_13_P25, &P_K5_13_PS15, 41},
			{CH_TCH9_9K6, &POLY_K5_12, 480, &P_K5_12_P25, &P_K5_12_P23, &P_K5_12_PS25, 158},
		};
		for (auto &e : v) {
			ChanTab &t = g_chan[e.ch];
			init_tab(t, 662, 648, *e.c, e.len, 1, 658);
			expand_punct(g_keep[e.ch], e.c->N, e.c->N * (e.len + 4), e.pre, e.mainp, e.post, e.rep);
			fill_punctured(t, g_keep[e.ch], [](int q) { return gw(kep_intra(81, q), 0); });
			for (int x = 0; x < 648; x++) {
				int age = 2 - (x % 3);
				int m = x < 52 ? x : x + 10;
				int s = m < 52 ? m : m + 4;
				t.t9_src[x] = (uint16_t)(s | (age << 10) | (scr[x] ? G_FLIP : 0));
			}
			for (int m = 0; m < 658; m++)
				t.cmap[m < 52 ? m : m + 4] = (int16_t)m;
		}
	}
	// RACH: demux into x[494], descramble, class-1 block transmitted twice and soft-averaged,
	// deinterleave(14) / deinterleave(33)+6, K5 r1/4 with b[4i+2], b[4i+3] (i<135) punctured
	// (rach.c:58-63, 148-167)
	{
		ChanTab &t = g_chan[CH_RACH];
		init_tab(t, 494, 494, POLY_K5_14, 159, 1, 0);
		auto x2e = [](int x) {           // x index -> ebit index (inverse of rach.c:148-151)
			if (x < 112) return x + 136;
			if (x < 248) return x - 112;
			if (x < 382) return x + 112;
			return x - 134;
		};
		std::vector<uint8_t> &keep = g_keep[CH_RACH];
		keep.assign(652, 1);
		for (int i = 0; i < 135; i++)
			keep[4 * i + 2] = keep[4 * i + 3] = 0;
		int q = 0;
		for (int k = 0; k < 652; k++) {
			if (!keep[k])
				continue;
			int c = q++;                 // index into bits_c[382]
			if (c < 264) {               // class 2, deinterleave(33)
				int x = 112 + kep_intra(33, c);
				t.g[k] = gw(x2e(x), scr[x]);
			} else if (c < 270) {        // class 2 tail passes through
				int x = 112 + c;
				t.g[k] = gw(x2e(x), scr[x]);
			} else {                     // class 1: average of the two copies
				int i1 = kep_intra(14, c - 270), i2 = i1 + 382;
				t.g[k]  = gw(x2e(i1), scr[i1]);
				t.g2[k] = gw(x2e(i2), scr[i2]);
			}
		}
	}
	// TCH3: status e[52..55] out, decipher/descramble 208, de-mux two 104-bit frames, 104-bit
	// permutation, K7 r1/2 tail-biting on the first 72 bits with every 4th coded bit punctured
	// (tch3.c:42-48, 134-174).  Four programs (mux mode m, frame f) of 128 words each at
	// g[(2m+f)*128]: words 0..95 = coded bits (punctured -> erased), 96..127 = the 32
	// unprotected class-2 bits c[72..103] that are sliced by sign (tch3.c:178-179).
	{
		ChanTab &t = g_chan[CH_TCH3];
		init_tab(t, 212, 212, POLY_K7_12, 48, 0, 208);
		expand_punct(g_keep[CH_TCH3], 2, 96, nullptr, &P_K5_12_P12, nullptr, 0);
		for (int m = 0; m < 208; m++)
			t.cmap[m < 52 ? m : m + 4] = (int16_t)m;
		for (int m = 0; m < 2; m++)
			for (int f = 0; f < 2; f++) {
				uint16_t *g = &t.g[(2 * m + f) * 128];
				int q = 0;
				for (int k = 0; k < 96; k++)
					if (g_keep[CH_TCH3][k])
						g[k] = tch3_gather(m, f, q++);
				for (int j = 0; j < 32; j++)
					g[96 + j] = tch3_gather(m, f, 72 + j);
			}
	}
	// DC12: descramble 432, deinterleave(54), K9 r1/3 tail-biting, P(12;13) (xch_dc12.c:45-52,94-97)
	{
		ChanTab &t = g_chan[CH_DC12];
		init_tab(t, 432, 432, POLY_K9_13, 208, 0, 0);
		expand_punct(g_keep[CH_DC12], 3, 624, nullptr, &P_K9_13_P1213, nullptr, 0);
		fill_punctured(t, g_keep[CH_DC12], [&](int q) { int s = kep_intra(54, q); return gw(s, scr[s]); });
	}
}

// The decode kernels of these four channels are compiled without the erased-position test (decode_unit.cuh:
// chan_has_erasures): their programs must cover every coded bit of every trellis step.
static void check_no_erasures()
{
	const int chans[4] = {CH_BCCH, CH_CCCH, CH_FACCH3, CH_FACCH9};
	for (int c : chans) {
		const ChanTab &t = g_chan[c];
		for (int i = 0; i < t.n_steps * t.N; i++)
			if (t.g[i] == G_ERASED) {
				fprintf(stderr, "gmr1_tables: channel %d has an erased position at coded bit %d\n", c, i);
				abort();
			}
	}
}

static void ensure()
{
	std::call_once(g_once, [] {
		build_all();
		check_no_erasures();
	});
}

const ChanTab &chan_tab(int ch) { ensure(); return g_chan[ch]; }
const uint8_t *scramble_seq() { ensure(); return g_scr; }

int chan_keep_mask(int ch, uint8_t *mask, int max)
{
	ensure();
	const std::vector<uint8_t> &k = g_keep[ch];
	if ((int)k.size() > max)
		return -1;
	memcpy(mask, k.data(), k.size());
	return (int)k.size();
}

// ---------------------------------------------------------------------------- burst formats

static BurstTab g_burst[BT_COUNT];
static std::once_flag g_burst_once;

static void build_bursts()
{
	for (int b = 0; b < BT_COUNT; b++) {
		const BurstDef &d = BURSTS[b];
		BurstTab &t = g_burst[b];
		memset(&t, 0, sizeof(t));
		t.rotation = 3.14159265358979323846264338327f / d.rot_div;
		t.nbits = d.nbits;
		t.len = d.len;
		t.ebits = d.ebits;
		for (int i = 0; i < MAX_SYNC; i++) {
			if (!d.sync[i][0].syms)
				break;
			t.n_sync = i + 1;
			for (int c = 0; c < MAX_SYNC_CHUNK && d.sync[i][c].syms; c++) {
				const char *s = d.sync[i][c].syms;
				int l = (int)strlen(s);
				t.n_chunk[i] = c + 1;
				t.s_pos[i][c] = (int16_t)d.sync[i][c].pos;
				t.s_len[i][c] = (int16_t)l;
				for (int k = 0; k < l; k++) {
					int v = s[k] - '0';
					// 1 bit/symbol formats list bits: symbol idx 1 is the phase-pi point
					t.s_sym[i][c][k] = (uint8_t)(d.nbits == 1 ? 2 * v : v);
				}
			}
		}
		for (int c = 0; c < MAX_DATA_CHUNK && d.data[c].len; c++) {
			t.n_data = c + 1;
			t.d_pos[c] = (int16_t)d.data[c].pos;
			t.d_len[c] = (int16_t)d.data[c].len;
		}
	}
}

const BurstTab &burst_tab(int bt)
{
	std::call_once(g_burst_once, build_bursts);
	return g_burst[bt];
}

}  // namespace gmr1
