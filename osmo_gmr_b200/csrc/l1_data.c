/* l1_data.c - the reference's L1 data symbols and the puncturing-array generator (host side, API surface).
 *
 *   gmr1_conv_*              src/l1/conv.c:138-571   nine trellises as struct osmo_conv_code
 *   gmr1_crc8 / 12 / 16      src/l1/crc.c:38-63
 *   gmr1_punct_*             src/l1/punct.c:137-end  51 puncturing masks (ETSI TS 101 376-5-3)
 *   gmr1_puncturer_generate  src/l1/punct.c:48-135
 *
 * The trellis tables are not literals: they are generated from the generator polynomials when the library is
 * loaded (next_state[s][b] = (2s + b) mod 2^(K-1), next_output[s][b] = parity of (2s + b) & g_j, g_0 in the most
 * significant output bit), the same rule csrc/gmr1_tables.cpp uses for the decode kernels.
 * tests/test_l1_data_cpu.py compares every symbol with the reference build (oracle/_ref), field by field. */
#include "../../include/gmr1_b200_compat.h"

#include <errno.h>
#include <stdlib.h>

/* generator polynomials, bit i = D^i */
struct poly_def { int N, K; uint16_t g[5]; };
static const struct poly_def P_K5_12 = {2, 5, {0x19, 0x17}};                  /* 1+D3+D4 ; 1+D+D2+D4 */
static const struct poly_def P_K5_13 = {3, 5, {0x15, 0x1b, 0x1f}};
static const struct poly_def P_K5_14 = {4, 5, {0x19, 0x17, 0x15, 0x1f}};
static const struct poly_def P_K5_15 = {5, 5, {0x15, 0x1b, 0x1f, 0x1d, 0x17}};
static const struct poly_def P_K6_14 = {4, 6, {0x25, 0x2d, 0x3b, 0x3f}};      /* 1+D2+D5 ; 1+D2+D3+D5 ; 1+D+D3+D4+D5 ; all */
static const struct poly_def P_K9_12 = {2, 9, {0x11d, 0x1af}};                /* 1+D2+D3+D4+D8 ; 1+D+D2+D3+D5+D7+D8 */
static const struct poly_def P_K9_13 = {3, 9, {0x1ed, 0x19b, 0x127}};
/* g3 as the reference's TABLE has it (1+D+D2+D3+D4+D6+D8); the comment above that table (conv.c:428-438) also lists D5 */
static const struct poly_def P_K9_14 = {4, 9, {0x1b9, 0x1a5, 0x13b, 0x15f}};
static const struct poly_def P_K7_12 = {2, 7, {0x6d, 0x4f}};                  /* TCH3: 1+D2+D3+D5+D6 ; 1+D+D2+D3+D6 */

static uint8_t ns_k5[16][2], ns_k6[32][2], ns_k7[64][2], ns_k9[256][2];
static uint8_t no_k5_12[16][2], no_k5_13[16][2], no_k5_14[16][2], no_k5_15[16][2], no_k6_14[32][2],
               no_k9_12[256][2], no_k9_13[256][2], no_k9_14[256][2], no_k7_12[64][2];

static void fill_state(uint8_t (*ns)[2], int K)
{
	const int n = 1 << (K - 1);
	for (int s = 0; s < n; s++)
		for (int b = 0; b < 2; b++)
			ns[s][b] = (uint8_t)(((s << 1) | b) & (n - 1));
}

static void fill_output(uint8_t (*no)[2], const struct poly_def *p)
{
	const int n = 1 << (p->K - 1);
	for (int s = 0; s < n; s++)
		for (int b = 0; b < 2; b++) {
			const unsigned reg = ((unsigned)s << 1) | (unsigned)b;
			unsigned o = 0;
			for (int j = 0; j < p->N; j++)
				o = (o << 1) | (unsigned)(__builtin_popcount(reg & p->g[j]) & 1);
			no[s][b] = (uint8_t)o;
		}
}

__attribute__((constructor(101))) static void l1_data_init(void)
{
	fill_state(ns_k5, 5);
	fill_state(ns_k6, 6);
	fill_state(ns_k7, 7);
	fill_state(ns_k9, 9);
	fill_output(no_k5_12, &P_K5_12);
	fill_output(no_k5_13, &P_K5_13);
	fill_output(no_k5_14, &P_K5_14);
	fill_output(no_k5_15, &P_K5_15);
	fill_output(no_k6_14, &P_K6_14);
	fill_output(no_k9_12, &P_K9_12);
	fill_output(no_k9_13, &P_K9_13);
	fill_output(no_k9_14, &P_K9_14);
	fill_output(no_k7_12, &P_K7_12);
}

/* len is filled in and term overridden when a channel specialises its copy (bcch.c:44-52 and friends) */
#define CONV(name, n, k, t, out, st) \
	const struct osmo_conv_code name = {.N = n, .K = k, .len = 0, .term = t, .next_output = out, .next_state = st}
CONV(gmr1_conv_k5_12, 2, 5, CONV_TERM_FLUSH, no_k5_12, ns_k5);
CONV(gmr1_conv_k5_13, 3, 5, CONV_TERM_FLUSH, no_k5_13, ns_k5);
CONV(gmr1_conv_k5_14, 4, 5, CONV_TERM_FLUSH, no_k5_14, ns_k5);
CONV(gmr1_conv_k5_15, 5, 5, CONV_TERM_FLUSH, no_k5_15, ns_k5);
CONV(gmr1_conv_k6_14, 4, 6, CONV_TERM_FLUSH, no_k6_14, ns_k6);
CONV(gmr1_conv_k9_12, 2, 9, CONV_TERM_FLUSH, no_k9_12, ns_k9);
CONV(gmr1_conv_k9_13, 3, 9, CONV_TERM_FLUSH, no_k9_13, ns_k9);
CONV(gmr1_conv_k9_14, 4, 9, CONV_TERM_FLUSH, no_k9_14, ns_k9);
CONV(gmr1_conv_tch3,  2, 7, CONV_TERM_TAIL_BITING, no_k7_12, ns_k7);

const struct osmo_crc8gen_code  gmr1_crc8  = {.bits = 8,  .poly = 0x9b,   .init = 0, .remainder = 0};
const struct osmo_crc16gen_code gmr1_crc12 = {.bits = 12, .poly = 0x80f,  .init = 0, .remainder = 0};
const struct osmo_crc16gen_code gmr1_crc16 = {.bits = 16, .poly = 0x1021, .init = 0, .remainder = 0};

#define PUNCT(name, r_, L_, N_, ...) \
	const struct gmr1_puncturer gmr1_punct_##name = {.r = r_, .L = L_, .N = N_, .mask = {__VA_ARGS__}};
#include "../../include/gmr1_punct_masks.inc"
#undef PUNCT

/* coded bits of one block before puncturing (osmo_conv_get_output_length(code, 0) of libosmocore for a code
 * whose puncture array is still unset - the state every caller of the generator is in) */
static int coded_len(const struct osmo_conv_code *code)
{
	int n = code->len * code->N;
	if (code->term == CONV_TERM_FLUSH)
		n += code->N * (code->K - 1);
	if (code->puncture)
		for (const int *p = code->puncture; *p >= 0; p++)
			if (*p < code->len * code->N + (code->term == CONV_TERM_FLUSH ? code->N * (code->K - 1) : 0))
				n--;
	return n;
}

int gmr1_puncturer_generate(struct osmo_conv_code *code, const struct gmr1_puncturer *punct_pre,
                            const struct gmr1_puncturer *punct_main, const struct gmr1_puncturer *punct_post,
                            int repeat)
{
	const int N = code->N;
	if ((punct_pre && punct_pre->N != N) || punct_main->N != N || (punct_post && punct_post->N != N))
		return -EINVAL;

	/* room: every zero of the first / last block's mask, `repeat` passes of the main mask, the terminator */
	const int total = coded_len(code);
	int body = total, room = 1;
	if (punct_pre) {
		body -= punct_pre->L * N;
		room += punct_pre->r;
	}
	if (punct_post) {
		body -= punct_post->L * N;
		room += punct_post->r;
	}
	const int span = punct_main->L * N;
	if (!repeat)
		repeat = (body + span - 1) / span;
	room += repeat * punct_main->r;

	int *p = malloc((size_t)(room > 0 ? room : 1) * sizeof(int));
	if (!p)
		return -ENOMEM;

	int pos = 0, n = 0;                       /* coded-bit position, entries written */
	if (punct_pre)
		for (int k = 0; pos < total && k < punct_pre->L * N; pos++, k++)
			if (!punct_pre->mask[k])
				p[n++] = pos;
	const int main_end = punct_post ? total - punct_post->L * N : total;
	for (int i = 0; i < repeat; i++)
		for (int k = 0; pos < main_end && k < span; pos++, k++)
			if (!punct_main->mask[k])
				p[n++] = pos;
	if (punct_post) {
		pos = main_end;                       /* the last block sits at the very end (punct.c:119-125) */
		for (int k = 0; pos > 0 && k < punct_post->L * N; pos++, k++)
			if (!punct_post->mask[k])
				p[n++] = pos;
	}
	p[n] = -1;
	code->puncture = p;
	return 0;
}
