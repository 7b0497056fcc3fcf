/* oracle/port/l1_port.c - plain-C restatement of the reference's L1 channel coding (src/l1/).
 *
 * TEST INFRASTRUCTURE ONLY: the oracle is the checker for the CUDA path, never the thing shipped
 * or measured (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call it).
 * Pinned against the reference itself: tests/test_oracle_cpu.py runs this file and the reference's
 * own C sources (oracle/_ref, built from /root/reference + shim) on the same inputs and requires
 * identical outputs; the committed fixtures under tests/golden/ were generated from oracle/_ref.
 * Third-party semantics (osmo_conv_*, osmo_crc*, bit packing) come from oracle/shim, which restates
 * un-vendored libosmocore and is "parity unpinned" (see shim_core.c header).
 *
 * Exports the reference's function names so one ctypes wrapper drives either library.  Each
 * function cites the reference lines it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <osmocom/core/bits.h>
#include <osmocom/core/conv.h>
#include <osmocom/core/crcgen.h>

#include "port.h"

/* ------------------------------------------------------------------ CRCs: src/l1/crc.c:38-63 */
const struct osmo_crc8gen_code  gmr1_crc8  = { .bits = 8,  .poly = 0x9b,   .init = 0, .remainder = 0 };
const struct osmo_crc16gen_code gmr1_crc12 = { .bits = 12, .poly = 0x80f,  .init = 0, .remainder = 0 };
const struct osmo_crc16gen_code gmr1_crc16 = { .bits = 16, .poly = 0x1021, .init = 0, .remainder = 0 };

/* ------------------------------------------------------------------ trellises: src/l1/conv.c
 * The reference stores next_output / next_state tables; they are the shift-register trellis of the
 * generator polynomials given in its comments (conv.c:124-127, 149-153, 175-180, 202-208, 230-235,
 * 261-264, 346-350, 432-437, 519-522), so the tables are generated here. */
#define MAX_STATES 256
struct trellis { uint8_t out[MAX_STATES][2], nxt[MAX_STATES][2]; };
static struct trellis t_k5_12, t_k5_13, t_k5_14, t_k5_15, t_k7_12, t_k9_13;

static void trellis_build(struct trellis *t, int N, int K, const unsigned *g)
{
	int ns = 1 << (K - 1), s, b, j;
	for (s = 0; s < ns; s++)
		for (b = 0; b < 2; b++) {
			unsigned reg = ((unsigned)s << 1) | (unsigned)b, ov = 0;
			for (j = 0; j < N; j++)
				ov = (ov << 1) | (unsigned)__builtin_parity(reg & g[j]);
			t->out[s][b] = (uint8_t)ov;
			t->nxt[s][b] = (uint8_t)(reg & (unsigned)(ns - 1));
		}
}

static struct osmo_conv_code c_bcch, c_ccch, c_facch3, c_facch9, c_tch9[3], c_rach, c_tch3, c_dc12;

static void code_set(struct osmo_conv_code *c, int N, int K, int len, enum osmo_conv_term term, struct trellis *t)
{
	memset(c, 0, sizeof(*c));
	c->N = N; c->K = K; c->len = len; c->term = term;
	c->next_output = t->out;
	c->next_state = t->nxt;
}

/* ------------------------------------------------------------------ puncturing: src/l1/punct.c:48-133
 * mask strings: '1' transmitted, '0' punctured, L*N entries (punct.c:137-173, 239-247, 389-427,
 * 448-481, 1105-1125) */
struct punct { int L; const char *m; };
static const struct punct P_12_P23 = {3, "011011"}, P_12_P25 = {5, "1011101111"}, P_12_Ps25 = {5, "1111101110"},
	P_12_P12 = {2, "1110"},
	P_13_P25 = {5, "111111101111101"}, P_13_P15 = {5, "101111111111111"}, P_13_Ps15 = {5, "111111111111101"},
	P_15_P23 = {3, "111111101111110"}, P_15_P53 = {3, "111011001111100"}, P_15_Ps53 = {3, "111001001111101"},
	P_93_P1213 = {13, "110101011110101011110101011110101011111"};

static int *punct_generate(const struct osmo_conv_code *code, const struct punct *pre, const struct punct *mainp,
                           const struct punct *post, int repeat)
{
	int N = code->N, cl = osmo_conv_get_output_length(code, 0), end = cl;
	int *p = malloc(sizeof(int) * (cl + 1)), io = 0, ii = 0, ip, i, d;
	if (pre)
		for (ip = 0; ii < cl && ip < pre->L * N; ii++, ip++)
			if (pre->m[ip] == '0')
				p[io++] = ii;
	if (post)
		end -= post->L * N;
	d = mainp->L * N;
	if (!repeat) {
		int span = cl - (pre ? pre->L * N : 0) - (post ? post->L * N : 0);
		repeat = (span + d - 1) / d;
	}
	for (i = 0; i < repeat; i++)
		for (ip = 0; ii < end && ip < d; ii++, ip++)
			if (mainp->m[ip] == '0')
				p[io++] = ii;
	if (post)
		for (ii = end, ip = 0; ip < post->L * N; ii++, ip++)
			if (post->m[ip] == '0')
				p[io++] = ii;
	p[io] = -1;
	return p;
}

static void __attribute__((constructor)) l1_port_init(void)
{
	static const unsigned g512[] = {0x19, 0x17}, g513[] = {0x15, 0x1b, 0x1f}, g514[] = {0x19, 0x17, 0x15, 0x1f},
		g515[] = {0x15, 0x1b, 0x1f, 0x1d, 0x17}, g712[] = {0x6d, 0x4f}, g913[] = {0x1ed, 0x19b, 0x127};
	int i, *p;
	trellis_build(&t_k5_12, 2, 5, g512); trellis_build(&t_k5_13, 3, 5, g513); trellis_build(&t_k5_14, 4, 5, g514);
	trellis_build(&t_k5_15, 5, 5, g515); trellis_build(&t_k7_12, 2, 7, g712); trellis_build(&t_k9_13, 3, 9, g913);

	code_set(&c_bcch, 2, 5, 208, CONV_TERM_FLUSH, &t_k5_12);            /* bcch.c:44-50 */
	code_set(&c_ccch, 2, 5, 208, CONV_TERM_FLUSH, &t_k5_12);            /* ccch.c:44-50 */
	code_set(&c_facch3, 4, 5, 92, CONV_TERM_FLUSH, &t_k5_14);           /* facch3.c:44-50 */
	code_set(&c_facch9, 2, 5, 316, CONV_TERM_FLUSH, &t_k5_12);          /* facch9.c:44-50 */
	code_set(&c_tch9[0], 5, 5, 144, CONV_TERM_FLUSH, &t_k5_15);         /* tch9.c:55-79 */
	c_tch9[0].puncture = punct_generate(&c_tch9[0], &P_15_P53, &P_15_P23, &P_15_Ps53, 41);
	code_set(&c_tch9[1], 3, 5, 240, CONV_TERM_FLUSH, &t_k5_13);
	c_tch9[1].puncture = punct_generate(&c_tch9[1], &P_13_P15, &P_13_P25, &P_13_Ps15, 41);
	code_set(&c_tch9[2], 2, 5, 480, CONV_TERM_FLUSH, &t_k5_12);
	c_tch9[2].puncture = punct_generate(&c_tch9[2], &P_12_P25, &P_12_P23, &P_12_Ps25, 158);
	code_set(&c_rach, 4, 5, 159, CONV_TERM_FLUSH, &t_k5_14);            /* rach.c:44-66 */
	p = malloc(sizeof(int) * 271);
	for (i = 0; i < 135; i++) { p[2 * i] = 4 * i + 2; p[2 * i + 1] = 4 * i + 3; }
	p[270] = -1;
	c_rach.puncture = p;
	code_set(&c_tch3, 2, 7, 48, CONV_TERM_TAIL_BITING, &t_k7_12);       /* tch3.c:42-48 */
	c_tch3.puncture = punct_generate(&c_tch3, NULL, &P_12_P12, NULL, 0);
	code_set(&c_dc12, 3, 9, 208, CONV_TERM_TAIL_BITING, &t_k9_13);      /* xch_dc12.c:45-52 */
	c_dc12.puncture = punct_generate(&c_dc12, NULL, &P_93_P1213, NULL, 0);
}

/* ------------------------------------------------------------------ scrambler: src/l1/scramb.c:39-92 */
static inline int scr_step(uint16_t *r)
{
	int b = ((*r >> 14) ^ *r) & 1;
	*r = (uint16_t)((*r << 1) | b);
	return b;
}

void gmr1_scramble_sbit(sbit_t *out, const sbit_t *in, int len)
{
	uint16_t r = 0x4d4b;
	int i;
	for (i = 0; i < len; i++) {
		sbit_t v = in[i];
		out[i] = scr_step(&r) ? -v : v;
	}
}

void gmr1_scramble_ubit(ubit_t *out, const ubit_t *in, int len)
{
	uint16_t r = 0x4d4b;
	int i;
	for (i = 0; i < len; i++)
		out[i] = in[i] ^ scr_step(&r);
}

/* ------------------------------------------------------------------ interleavers: src/l1/interleave.c */
void gmr1_interleave_intra(void *out, const void *in, int N)        /* :49-62 */
{
	const uint8_t *i8 = in; uint8_t *o8 = out; int kc;
	for (kc = 0; kc < (N << 3); kc++)
		o8[N * ((5 * kc) & 7) + (kc >> 3)] = i8[kc];
}

void gmr1_deinterleave_intra(void *out, const void *in, int N)      /* :74-87 */
{
	const uint8_t *i8 = in; uint8_t *o8 = out; int kc;
	for (kc = 0; kc < (N << 3); kc++)
		o8[kc] = i8[N * ((5 * kc) & 7) + (kc >> 3)];
}

int gmr1_interleaver_init(struct gmr1_interleaver *il, int N, int K)   /* :95-117 */
{
	memset(il, 0, sizeof(*il));
	il->bits_cpp = calloc((size_t)N * K, 1);
	if (!il->bits_cpp)
		return -12;
	il->N = N; il->K = K;
	return 0;
}

void gmr1_interleaver_fini(struct gmr1_interleaver *il) { free(il->bits_cpp); memset(il, 0, sizeof(*il)); }

void gmr1_interleave_inter(struct gmr1_interleaver *il, void *bits_epp, void *bits_ep)    /* :138-160 */
{
	uint8_t *d = bits_epp; int jk;
	memcpy(&il->bits_cpp[(il->n % il->N) * il->K], bits_ep, il->K);
	for (jk = 0; jk < il->K; jk++)
		d[jk] = il->bits_cpp[(((il->n % il->N) - (jk % il->N) + il->N) % il->N) * il->K + jk];
	il->n++;
}

void gmr1_deinterleave_inter(struct gmr1_interleaver *il, void *bits_ep, void *bits_epp)  /* :168-190 */
{
	const uint8_t *s = bits_epp; int jk;
	for (jk = 0; jk < il->K; jk++)
		il->bits_cpp[(((il->n % il->N) - (jk % il->N) + il->N) % il->N) * il->K + jk] = s[jk];
	memcpy(bits_ep, &il->bits_cpp[((il->n + 1) % il->N) * il->K], il->K);
	il->n++;
}

/* ------------------------------------------------------------------ BCCH / CCCH / DC12 */
static void xcch_encode(const struct osmo_conv_code *code, ubit_t *bits_e, const uint8_t *l2, int ileave, int pad)
{
	ubit_t u[208], c[432], ep[432];
	int n = 8 * ileave + 2 * pad;
	memset(ep, 0, sizeof(ep));
	osmo_pbit2ubit_ext(u, 0, l2, 0, 192, 1);
	osmo_crc16gen_set_bits(&gmr1_crc16, u, 192, u + 192);
	osmo_conv_encode(code, u, c);
	gmr1_interleave_intra(ep + pad, c, ileave);
	gmr1_scramble_ubit(bits_e, ep, n);
}

static int xcch_decode(const struct osmo_conv_code *code, uint8_t *l2, const sbit_t *bits_e, int *conv_rv, int ileave, int pad)
{
	sbit_t ep[432], c[432];
	ubit_t u[208];
	int rv;
	gmr1_scramble_sbit(ep, bits_e, 8 * ileave + 2 * pad);
	gmr1_deinterleave_intra(c, ep + pad, ileave);
	rv = osmo_conv_decode(code, c, u);
	if (conv_rv)
		*conv_rv = rv;
	rv = osmo_crc16gen_check_bits(&gmr1_crc16, u, 192, u + 192);
	osmo_ubit2pbit_ext(l2, 0, u, 0, 192, 1);
	return rv;
}

void gmr1_bcch_encode(ubit_t *e, const uint8_t *l2) { xcch_encode(&c_bcch, e, l2, 53, 0); }               /* bcch.c:59-72 */
int  gmr1_bcch_decode(uint8_t *l2, const sbit_t *e, int *cv) { return xcch_decode(&c_bcch, l2, e, cv, 53, 0); } /* :84-103 */
void gmr1_ccch_encode(ubit_t *e, const uint8_t *l2) { xcch_encode(&c_ccch, e, l2, 53, 4); }               /* ccch.c:59-76 */
int  gmr1_ccch_decode(uint8_t *l2, const sbit_t *e, int *cv) { return xcch_decode(&c_ccch, l2, e, cv, 53, 4); } /* :88-107 */
void gmr1_xch_dc12_encode(ubit_t *e, const uint8_t *l2) { xcch_encode(&c_dc12, e, l2, 54, 0); }           /* xch_dc12.c:63-75 */
int  gmr1_xch_dc12_decode(uint8_t *l2, const sbit_t *e, int *cv) { return xcch_decode(&c_dc12, l2, e, cv, 54, 0); } /* :87-106 */

/* ------------------------------------------------------------------ FACCH3: src/l1/facch3.c */
void gmr1_facch3_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *bits_s, const ubit_t *ciph)   /* :64-106 */
{
	ubit_t u[92], c[384], cp[384], ep[96], xmy[96];
	int i, j;
	osmo_pbit2ubit_ext(u, 0, l2, 0, 76, 1);
	osmo_crc16gen_set_bits(&gmr1_crc16, u, 76, u + 76);
	osmo_conv_encode(&c_facch3, u, c);
	for (i = 0; i < 384; i++)
		cp[(i & 3) * 96 + (i >> 2)] = c[i];
	for (i = 0; i < 4; i++) {
		gmr1_interleave_intra(ep, cp + 96 * i, 12);
		gmr1_scramble_ubit(xmy, ep, 96);
		if (ciph)
			for (j = 0; j < 96; j++)
				xmy[j] ^= ciph[96 * i + j];
		memcpy(bits_e + 104 * i, xmy, 22);
		memcpy(bits_e + 104 * i + 22, bits_s + 8 * i, 8);
		memcpy(bits_e + 104 * i + 30, xmy + 22, 74);
	}
}

int gmr1_facch3_decode(uint8_t *l2, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph, int *conv_rv)  /* :122-170 */
{
	sbit_t xmy[96], ep[96], cp[384], c[384];
	ubit_t u[92];
	int i, j, rv;
	for (i = 0; i < 4; i++) {
		const sbit_t *e = bits_e + 104 * i;
		for (j = 0; j < 8; j++)
			bits_s[8 * i + j] = e[22 + j] < 0;
		memcpy(xmy, e, 22);
		memcpy(xmy + 22, e + 30, 74);
		if (ciph)
			for (j = 0; j < 96; j++)
				if (ciph[96 * i + j])
					xmy[j] *= -1;
		gmr1_scramble_sbit(ep, xmy, 96);
		gmr1_deinterleave_intra(cp + 96 * i, ep, 12);
	}
	for (i = 0; i < 384; i++)
		c[i] = cp[(i & 3) * 96 + (i >> 2)];
	rv = osmo_conv_decode(&c_facch3, c, u);
	if (conv_rv)
		*conv_rv = rv;
	rv = osmo_crc16gen_check_bits(&gmr1_crc16, u, 76, u + 76);
	l2[9] = 0;
	osmo_ubit2pbit_ext(l2, 0, u, 0, 76, 1);
	return rv;
}

/* ------------------------------------------------------------------ NT9 framing shared by FACCH9 / TCH9 */
static void nt9_mux(ubit_t *bits_e, const ubit_t *x, const ubit_t *sacch, const ubit_t *status, const ubit_t *ciph)
{
	ubit_t my[658];
	int i;
	memcpy(my, x, 52); memcpy(my + 52, sacch, 10); memcpy(my + 62, x + 52, 596);
	if (ciph)
		for (i = 0; i < 658; i++)
			my[i] ^= ciph[i];
	memcpy(bits_e, my, 52); memcpy(bits_e + 52, status, 4); memcpy(bits_e + 56, my + 52, 606);
}

static void nt9_demux(sbit_t *x, sbit_t *sacch, sbit_t *status, const sbit_t *bits_e, const ubit_t *ciph)
{
	sbit_t my[658];
	int i;
	memcpy(my, bits_e, 52); memcpy(status, bits_e + 52, 4); memcpy(my + 52, bits_e + 56, 606);
	if (ciph)
		for (i = 0; i < 658; i++)
			if (ciph[i])
				my[i] *= -1;
	memcpy(x, my, 52); memcpy(sacch, my + 52, 10); memcpy(x + 52, my + 62, 596);
}

void gmr1_facch9_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *sacch, const ubit_t *status, const ubit_t *ciph) /* facch9.c:57-93 */
{
	ubit_t u[316], c[640], x[648];
	osmo_pbit2ubit_ext(u, 0, l2, 0, 300, 1);
	osmo_crc16gen_set_bits(&gmr1_crc16, u, 300, u + 300);
	osmo_conv_encode(&c_facch9, u, c);
	memset(x, 0, sizeof(x));
	gmr1_interleave_intra(x + 4, c, 80);
	gmr1_scramble_ubit(x, x, 648);
	nt9_mux(bits_e, x, sacch, status, ciph);
}

int gmr1_facch9_decode(uint8_t *l2, sbit_t *sacch, sbit_t *status, const sbit_t *bits_e, const ubit_t *ciph, int *conv_rv) /* :107-144 */
{
	sbit_t x[648], c[640];
	ubit_t u[316];
	int rv;
	nt9_demux(x, sacch, status, bits_e, ciph);
	gmr1_scramble_sbit(x, x, 648);
	gmr1_deinterleave_intra(c, x + 4, 80);
	rv = osmo_conv_decode(&c_facch9, c, u);
	if (conv_rv)
		*conv_rv = rv;
	rv = osmo_crc16gen_check_bits(&gmr1_crc16, u, 300, u + 300);
	l2[37] = 0;
	osmo_ubit2pbit_ext(l2, 0, u, 0, 300, 1);
	return rv;
}

/* ------------------------------------------------------------------ TCH9: src/l1/tch9.c */
void gmr1_tch9_encode(ubit_t *bits_e, const uint8_t *l2, int mode, const ubit_t *sacch, const ubit_t *status,
                      const ubit_t *ciph, struct gmr1_interleaver *il)                                /* :93-128 */
{
	const struct osmo_conv_code *cc = &c_tch9[mode];
	ubit_t u[480], c[648], x[648];
	osmo_pbit2ubit_ext(u, 0, l2, 0, cc->len, 1);
	osmo_conv_encode(cc, u, c);
	gmr1_interleave_intra(x, c, 81);
	gmr1_interleave_inter(il, x, x);
	gmr1_scramble_ubit(x, x, 648);
	nt9_mux(bits_e, x, sacch, status, ciph);
}

void gmr1_tch9_decode(uint8_t *l2, sbit_t *sacch, sbit_t *status, const sbit_t *bits_e, int mode, const ubit_t *ciph,
                      struct gmr1_interleaver *il, int *conv_rv)                                      /* :140-175 */
{
	const struct osmo_conv_code *cc = &c_tch9[mode];
	sbit_t x[648], c[648];
	ubit_t u[480];
	int rv;
	nt9_demux(x, sacch, status, bits_e, ciph);
	gmr1_scramble_sbit(x, x, 648);
	gmr1_deinterleave_inter(il, x, x);
	gmr1_deinterleave_intra(c, x, 81);
	rv = osmo_conv_decode(cc, c, u);
	if (conv_rv)
		*conv_rv = rv;
	osmo_ubit2pbit_ext(l2, 0, u, 0, cc->len, 1);
}

/* ------------------------------------------------------------------ TCH3: src/l1/tch3.c:124-183 */
void gmr1_tch3_decode(uint8_t *frame0, uint8_t *frame1, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph,
                      int m, int *conv0_rv, int *conv1_rv)
{
	sbit_t xmy[208], epp[208], ep[104], c[104];
	ubit_t d[80];
	int i, j, kc, rv;
	for (i = 0; i < 4; i++)
		bits_s[i] = bits_e[52 + i] < 0;
	memcpy(xmy, bits_e, 52);
	memcpy(xmy + 52, bits_e + 56, 156);
	if (ciph)
		for (i = 0; i < 208; i++)
			if (ciph[i])
				xmy[i] *= -1;
	gmr1_scramble_sbit(epp, xmy, 208);
	for (i = 0; i < 2; i++) {
		for (j = 0; j < 104; j++)
			ep[j] = m ? epp[104 * i + j] : epp[2 * j + i];
		for (kc = 0; kc < 104; kc++) {
			int ii = kc % 24, ij = kc / 24;
			c[kc] = ep[ii < 8 ? ij + 5 * ii : ij + 4 * ii + 8];
		}
		rv = osmo_conv_decode(&c_tch3, c, d);
		if (i ? conv1_rv != NULL : conv0_rv != NULL)
			*(i ? conv1_rv : conv0_rv) = rv;
		for (j = 48; j < 80; j++)
			d[j] = c[j + 24] < 0;
		osmo_ubit2pbit(i ? frame1 : frame0, d, 80);
	}
}

/* ------------------------------------------------------------------ RACH: src/l1/rach.c */
void gmr1_rach_encode(ubit_t *bits_e, const uint8_t *rach, uint8_t sb_mask)      /* :76-122 */
{
	ubit_t u[159], c[382], e1p[112], e2p[270], ep[494], x[494];
	ubit_t *u1 = u + 135, *u2 = u;
	int i;
	osmo_pbit2ubit_ext(u1, 0, rach, 0, 16, 1);
	osmo_pbit2ubit_ext(u2, 0, rach, 16, 123, 1);
	osmo_crc8gen_set_bits(&gmr1_crc8, u1, 16, u1 + 16);
	osmo_crc16gen_set_bits(&gmr1_crc12, u2, 123, u2 + 123);
	for (i = 0; i < 8; i++)
		u1[16 + i] ^= (sb_mask >> (7 - i)) & 1;
	osmo_conv_encode(&c_rach, u, c);
	gmr1_interleave_intra(e1p, c + 270, 14);
	gmr1_interleave_intra(e2p, c, 33);
	memcpy(e2p + 264, c + 264, 6);
	memcpy(ep, e1p, 112); memcpy(ep + 112, e2p, 270); memcpy(ep + 382, e1p, 112);
	gmr1_scramble_ubit(x, ep, 494);
	memcpy(bits_e, x + 112, 136); memcpy(bits_e + 136, x, 112);
	memcpy(bits_e + 248, x + 382, 112); memcpy(bits_e + 360, x + 248, 134);
}

int gmr1_rach_decode(uint8_t *rach, const sbit_t *bits_e, uint8_t sb_mask, int *conv_rv, int *crc_rv)   /* :137-196 */
{
	sbit_t x[494], ep[494], e1p[112], e2p[270], c[382];
	ubit_t u[159], *u1 = u + 135, *u2 = u;
	int i, rv, crc[2];
	memcpy(x, bits_e + 136, 112); memcpy(x + 112, bits_e, 136);
	memcpy(x + 248, bits_e + 360, 134); memcpy(x + 382, bits_e + 248, 112);
	gmr1_scramble_sbit(ep, x, 494);
	memcpy(e2p, ep + 112, 270);
	for (i = 0; i < 112; i++)
		e1p[i] = (sbit_t)(((int)ep[i] + (int)ep[i + 382]) >> 1);
	gmr1_deinterleave_intra(c + 270, e1p, 14);
	gmr1_deinterleave_intra(c, e2p, 33);
	memcpy(c + 264, e2p + 264, 6);
	rv = osmo_conv_decode(&c_rach, c, u);
	if (conv_rv)
		*conv_rv = rv;
	crc[0] = osmo_crc8gen_check_bits(&gmr1_crc8, u1, 16, u1 + 16);
	crc[1] = osmo_crc16gen_check_bits(&gmr1_crc12, u2, 123, u2 + 123);
	if (crc[0]) {
		for (i = 0; i < 8; i++)
			u1[16 + i] ^= (sb_mask >> (7 - i)) & 1;
		crc[0] = osmo_crc8gen_check_bits(&gmr1_crc8, u1, 16, u1 + 16);
	}
	if (crc_rv) { crc_rv[0] = crc[0]; crc_rv[1] = crc[1]; }
	rach[17] = 0;
	osmo_ubit2pbit_ext(rach, 0, u1, 0, 16, 1);
	osmo_ubit2pbit_ext(rach, 16, u2, 0, 123, 1);
	return crc[0] || crc[1];
}

/* ------------------------------------------------------------------ A5: src/l1/a5.c */
static inline uint32_t a5_step(uint32_t r, int len, uint32_t taps)
{
	return ((r << 1) & ((1u << len) - 1u)) | (uint32_t)__builtin_parity(r & taps);
}

static void a5_clock(uint32_t *r, int force)                         /* :139-177 */
{
	static const int len[4] = {19, 22, 23, 17};
	static const uint32_t taps[4] = {0x072000, 0x311000, 0x660000, 0x013100};
	int cb[3] = { !!(r[3] & (1 << 15)), !!(r[3] & (1 << 6)), !!(r[3] & (1 << 1)) };
	int m = (cb[0] + cb[1] + cb[2]) >= 2, i;
	for (i = 0; i < 3; i++)
		if (force || cb[i] == m)
			r[i] = a5_step(r[i], len[i], taps[i]);
	r[3] = a5_step(r[3], len[3], taps[3]);
}

static ubit_t a5_out(const uint32_t *r)                              /* :183-207 */
{
#define B(x, n) (((x) >> (n)) & 1)
#define MAJ(x, a, b, c) ((B(x, a) + B(x, b) + B(x, c)) >= 2)
	return (ubit_t)((MAJ(r[0], 1, 6, 15) ^ B(r[0], 11)) ^ (MAJ(r[1], 3, 8, 14) ^ B(r[1], 1)) ^
	                (MAJ(r[2], 4, 15, 19) ^ B(r[2], 0)));
}

void gmr1_a5_1(uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul)    /* :226-282 */
{
	uint32_t r[4] = {0, 0, 0, 0};
	uint8_t k[8];
	int i, j;
	for (i = 0; i < 8; i++)
		k[i] = key[i ^ 1];
	k[6] ^= (fn & 0x0000f) << 4; k[3] ^= (fn & 0x00030) << 2; k[1] ^= (fn & 0x007c0) >> 3;
	k[0] ^= (fn & 0x0f800) >> 11; k[0] ^= (fn & 0x70000) >> 11;
	for (i = 0; i < 64; i++) {
		uint32_t b = (k[i >> 3] >> (7 - (i & 7))) & 1;
		a5_clock(r, 1);
		for (j = 0; j < 4; j++)
			r[j] ^= b;
	}
	for (j = 0; j < 4; j++)
		r[j] |= 1;
	for (i = 0; i < 250; i++)
		a5_clock(r, 0);
	for (i = 0; i < nbits; i++) {
		a5_clock(r, 0);
		if (dl) dl[i] = a5_out(r);
	}
	if (!ul)
		return;
	for (i = 0; i < nbits; i++) {
		a5_clock(r, 0);
		ul[i] = a5_out(r);
	}
}

void gmr1_a5(int n, uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul)   /* :57-80 */
{
	if (n == 0) {
		if (dl) memset(dl, 0, nbits);
		if (ul) memset(ul, 0, nbits);
	} else if (n == 1)
		gmr1_a5_1(key, fn, nbits, dl, ul);
}

/* ---- GSMTAP record (src/gsmtap.c:44-71): gsmtap_hdr + L2 in a libosmocore msgb -------------------------
 * header: version GSMTAP_VERSION, hdr_len in 32-bit words, type GSMTAP_TYPE_GMR1_UM, timeslot, frame number in
 * network byte order, sub_type = channel type; arfcn / signal / snr / antenna / sub-slot stay 0 */
#include <arpa/inet.h>
#include <osmocom/core/msgb.h>
#include <osmocom/core/gsmtap.h>

struct msgb *gmr1_gsmtap_makemsg(uint8_t chan_type, uint32_t fn, uint8_t tn, const uint8_t *l2, int len)
{
	struct msgb *m = msgb_alloc(sizeof(struct gsmtap_hdr) + len, "gmr1_gsmtap_tx");
	if (!m)
		return NULL;
	struct gsmtap_hdr *h = (struct gsmtap_hdr *)msgb_put(m, sizeof(*h));
	memset(h, 0, sizeof(*h));
	h->version = GSMTAP_VERSION;
	h->hdr_len = sizeof(*h) / 4;
	h->type = GSMTAP_TYPE_GMR1_UM;
	h->timeslot = tn;
	h->frame_number = htonl(fn);
	h->sub_type = chan_type;
	memcpy(msgb_put(m, len), l2, len);
	return m;
}
