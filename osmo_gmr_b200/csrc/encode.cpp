// encode.cpp - host-side GMR-1 channel encoders (the mirror of the decode gather programs).
//
// API surface of the reference's libgmr1-l1 that the receive path's tooling needs: the
// synthetic-signal generator (bench, tests) encodes L2 payloads with these, and the compat layer
// exports them under the reference's names.  They run on the host on purpose: encoding is the
// transmit side, not the receive hot path.  Reference: gmr1_bcch_encode src/l1/bcch.c:59,
// gmr1_ccch_encode ccch.c:59, gmr1_facch3_encode facch3.c:64, gmr1_facch9_encode facch9.c:57,
// gmr1_tch3_encode tch3.c:60 (whose osmo_conv_encode call has its arguments swapped at :81 -
// this implementation encodes what gmr1_tch3_decode decodes), gmr1_tch9_encode tch9.c:93,
// gmr1_rach_encode rach.c:76, gmr1_xch_dc12_encode xch_dc12.c:63.
#include "encode.h"

#include <string.h>
#include <vector>

namespace gmr1 {

static void crc_bits(const uint8_t *in, int len, unsigned poly, int bits, uint8_t *out)
{
	const unsigned top = 1u << (bits - 1), mask = (1u << bits) - 1u;
	unsigned crc = 0;
	for (int i = 0; i < len; i++) {
		crc ^= (unsigned)(in[i] & 1) << (bits - 1);
		crc = (crc & top) ? ((crc << 1) ^ poly) : (crc << 1);
	}
	crc &= mask;
	for (int i = 0; i < bits; i++)
		out[i] = (crc >> (bits - 1 - i)) & 1;
}

static void unpack_lsb(uint8_t *out, const uint8_t *in, int in_ofs, int n)
{
	for (int i = 0; i < n; i++) {
		const int p = in_ofs + i;
		out[i] = (in[p >> 3] >> (p & 7)) & 1;
	}
}

// unpunctured coded bits of `ch` for data bits u[len]; c_full has N*n_steps entries
static void conv_encode_full(int ch, const uint8_t *u, uint8_t *c_full)
{
	const ChanTab &t = chan_tab(ch);
	const CodePoly &c = chan_code(ch);
	unsigned state = 0;
	if (!t.flush)
		for (int i = 0; i < c.K - 1; i++)
			state = (state << 1) | u[t.len - (c.K - 1) + i];
	const unsigned smask = (1u << (c.K - 1)) - 1u;
	for (int i = 0; i < t.n_steps; i++) {
		const unsigned bit = i < t.len ? u[i] : 0u;
		const unsigned ov = code_output(c, (int)state, (int)bit);
		state = ((state << 1) | bit) & smask;
		for (int j = 0; j < c.N; j++)
			c_full[i * c.N + j] = (ov >> (c.N - 1 - j)) & 1;
	}
}

// scatter coded bits through a gather program: e[idx] = c ^ scramble-flip
static void scatter(const uint16_t *g, const uint16_t *g2, int n_coded, const uint8_t *c_full, uint8_t *e)
{
	for (int k = 0; k < n_coded; k++) {
		if (g[k] != G_ERASED)
			e[g[k] & G_IDX] = c_full[k] ^ ((g[k] & G_FLIP) ? 1 : 0);
		if (g2 && g2[k] != G_ERASED)
			e[g2[k] & G_IDX] = c_full[k] ^ ((g2[k] & G_FLIP) ? 1 : 0);
	}
}

static void apply_cipher(const ChanTab &t, uint8_t *e, const uint8_t *ciph)
{
	if (!ciph)
		return;
	for (int s = 0; s < t.n_in; s++)
		if (t.cmap[s] >= 0)
			e[s] ^= ciph[t.cmap[s]] & 1;
}

void encode_simple(int ch, uint8_t *bits_e, const uint8_t *l2)
{
	const ChanTab &t = chan_tab(ch);
	const uint8_t *scr = scramble_seq();
	uint8_t u[208], c[MAX_CODED];
	unpack_lsb(u, l2, 0, 192);
	crc_bits(u, 192, 0x1021, 16, u + 192);
	conv_encode_full(ch, u, c);
	for (int i = 0; i < t.n_in; i++)      // padding positions carry scrambled zeros (ccch.c:66-67,75)
		bits_e[i] = scr[i];
	scatter(t.g, nullptr, t.N * t.n_steps, c, bits_e);
}

void encode_facch3(uint8_t *bits_e, const uint8_t *l2, const uint8_t *bits_s, const uint8_t *ciph)
{
	const ChanTab &t = chan_tab(CH_FACCH3);
	uint8_t u[92], c[MAX_CODED];
	unpack_lsb(u, l2, 0, 76);
	crc_bits(u, 76, 0x1021, 16, u + 76);
	conv_encode_full(CH_FACCH3, u, c);
	memset(bits_e, 0, 416);
	scatter(t.g, nullptr, 384, c, bits_e);
	apply_cipher(t, bits_e, ciph);
	for (int b = 0; b < 4; b++)
		for (int j = 0; j < 8; j++)
			bits_e[104 * b + 22 + j] = bits_s[8 * b + j] & 1;
}

// common NT9 framing of FACCH9 / TCH9: x[648] (scrambled domain) + sacch + status -> e[662]
static void nt9_frame(uint8_t *bits_e, const uint8_t *x_scrambled, const uint8_t *sacch, const uint8_t *status,
                      const uint8_t *ciph)
{
	uint8_t my[658];
	memcpy(my, x_scrambled, 52);
	for (int i = 0; i < 10; i++)
		my[52 + i] = sacch[i] & 1;
	memcpy(my + 62, x_scrambled + 52, 596);
	if (ciph)
		for (int i = 0; i < 658; i++)
			my[i] ^= ciph[i] & 1;
	memcpy(bits_e, my, 52);
	for (int i = 0; i < 4; i++)
		bits_e[52 + i] = status[i] & 1;
	memcpy(bits_e + 56, my + 52, 606);
}

void encode_facch9(uint8_t *bits_e, const uint8_t *l2, const uint8_t *sacch, const uint8_t *status, const uint8_t *ciph)
{
	const uint8_t *scr = scramble_seq();
	uint8_t u[316], c[MAX_CODED], x[648];
	unpack_lsb(u, l2, 0, 300);
	crc_bits(u, 300, 0x1021, 16, u + 300);
	conv_encode_full(CH_FACCH9, u, c);
	memset(x, 0, sizeof(x));
	for (int kc = 0; kc < 640; kc++)          // interleave(80) into x[4..643], facch9.c:78-80
		x[4 + 80 * ((5 * kc) & 7) + (kc >> 3)] = c[kc];
	for (int i = 0; i < 648; i++)
		x[i] ^= scr[i];
	nt9_frame(bits_e, x, sacch, status, ciph);
}

void interleaver_init(Interleaver *il)
{
	memset(il, 0, sizeof(*il));
}

// conv-encode + puncture + intra-burst interleave(81): the 648 bits that enter the inter-burst interleaver
void encode_tch9_ep(uint8_t *ep, const uint8_t *l2, int mode)
{
	const int ch = CH_TCH9_2K4 + mode;
	const ChanTab &t = chan_tab(ch);
	uint8_t u[480], c[MAX_CODED], rx[648];
	std::vector<uint8_t> keep(MAX_CODED);
	const int n_coded = chan_keep_mask(ch, keep.data(), MAX_CODED);
	unpack_lsb(u, l2, 0, t.len);
	conv_encode_full(ch, u, c);
	int q = 0;
	for (int k = 0; k < n_coded; k++)
		if (keep[k])
			rx[q++] = c[k];
	for (int kc = 0; kc < 648; kc++)
		ep[81 * ((5 * kc) & 7) + (kc >> 3)] = rx[kc];
}

void encode_tch9(uint8_t *bits_e, const uint8_t *l2, int mode, const uint8_t *sacch, const uint8_t *status,
                 const uint8_t *ciph, Interleaver *il)
{
	const int ch = CH_TCH9_2K4 + mode;
	const ChanTab &t = chan_tab(ch);
	const uint8_t *scr = scramble_seq();
	uint8_t u[480], c[MAX_CODED], rx[648], ep[648], x[648];
	std::vector<uint8_t> keep(MAX_CODED);
	const int n_coded = chan_keep_mask(ch, keep.data(), MAX_CODED);
	unpack_lsb(u, l2, 0, t.len);
	conv_encode_full(ch, u, c);
	int q = 0;
	for (int k = 0; k < n_coded; k++)
		if (keep[k])
			rx[q++] = c[k];                    // 648 transmitted coded bits
	for (int kc = 0; kc < 648; kc++)           // intra interleave(81)
		ep[81 * ((5 * kc) & 7) + (kc >> 3)] = rx[kc];
	// inter-burst interleaver, depth 3 (interleave.c:138-160): column x of burst n carries the
	// bit of burst n - (x % 3)
	memcpy(il->hist[il->n % 3], ep, 648);
	for (int jk = 0; jk < 648; jk++) {
		const int age = jk % 3;
		x[jk] = (il->n - age >= 0) ? il->hist[(il->n - age) % 3][jk] : 0;
	}
	il->n++;
	for (int i = 0; i < 648; i++)
		x[i] ^= scr[i];
	nt9_frame(bits_e, x, sacch, status, ciph);
}

void encode_rach(uint8_t *bits_e, const uint8_t *rach, uint8_t sb_mask)
{
	const ChanTab &t = chan_tab(CH_RACH);
	uint8_t u[159], c[MAX_CODED];
	uint8_t *u1 = u + 135, *u2 = u;
	unpack_lsb(u1, rach, 0, 16);
	unpack_lsb(u2, rach, 16, 123);
	crc_bits(u1, 16, 0x9b, 8, u1 + 16);
	crc_bits(u2, 123, 0x80f, 12, u2 + 123);
	for (int i = 0; i < 8; i++)
		u1[16 + i] ^= (sb_mask >> (7 - i)) & 1;
	conv_encode_full(CH_RACH, u, c);
	memset(bits_e, 0, 494);
	scatter(t.g, t.g2, 652, c, bits_e);
}

void encode_tch3(uint8_t *bits_e, const uint8_t *frame0, const uint8_t *frame1, const uint8_t *bits_s,
                 const uint8_t *ciph, int m)
{
	const ChanTab &t = chan_tab(CH_TCH3);
	memset(bits_e, 0, 212);
	for (int f = 0; f < 2; f++) {
		const uint8_t *fr = f ? frame1 : frame0;
		uint8_t d[80], c[128];
		for (int i = 0; i < 80; i++)
			d[i] = (fr[i >> 3] >> (7 - (i & 7))) & 1;       // MSB first (osmo_pbit2ubit)
		conv_encode_full(CH_TCH3, d, c);                    // 96 coded bits of d[0..47]
		const uint16_t *g = &t.g[(2 * (m ? 1 : 0) + f) * 128];
		scatter(g, nullptr, 96, c, bits_e);
		scatter(g + 96, nullptr, 32, d + 48, bits_e);       // class-2 bits ride unprotected
	}
	apply_cipher(t, bits_e, ciph);
	for (int i = 0; i < 4; i++)
		bits_e[52 + i] = bits_s[i] & 1;
}

}  // namespace gmr1
