/* compat.c - the reference's C API (libgmr1-sdr + libgmr1-l1 symbols) as n = 1 wrappers over the
 * batched CUDA entry points, so that osmo-gmr's src/gmr1_rx.c links against libgmr1_b200.so
 * unchanged (see include/gmr1_b200_compat.h).  Plain C because the reference's structs carry C99
 * complex members.  Nothing here computes on the CPU except the host-side primitives that are API
 * surface but not hot path (encoders, scrambler, interleaver bookkeeping, A5 keystream, the 1 sps
 * modulator); every demod / decode / FCCH call goes to the GPU and fails with -errno if it cannot.
 */
#include "../../include/gmr1_b200.h"
#include "../../include/gmr1_b200_compat.h"

#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Every pointer the reference's API hands us is host memory: say so for the duration of the batched call, so that the
 * library stages through its per-thread page-locked arena instead of classifying and allocating per pointer. */
#define HOSTCALL(call) __extension__({ const int hint_ = gmr1b200_host_hint(1); const int rv_ = (call); \
                                       gmr1b200_host_hint(hint_); rv_; })

#define PI_F 3.14159265358979323846264338327f

/* ------------------------------------------------------------------ burst format data symbols */

static struct gmr1_pi4cxpsk_symbol bpsk_syms[2], qpsk_syms[4], qpsk_bits[4];
struct gmr1_pi4cxpsk_modulation gmr1_pi2cbpsk, gmr1_pi4cbpsk, gmr1_pi4cqpsk;

struct gmr1_pi4cxpsk_burst gmr1_bcch_burst, gmr1_dc2_burst, gmr1_dc6_burst, gmr1_dc12_burst,
	gmr1_nt3_speech_burst, gmr1_nt3_facch_burst, gmr1_nt6_burst, gmr1_nt9_burst, gmr1_rach_burst,
	gmr1_sdcch_burst;

const struct gmr1_fcch_burst gmr1_fcch_burst        = { 0.32f, 3 * 39 };
const struct gmr1_fcch_burst gmr1_fcch3_lband_burst = { 0.32f, 12 * 39 };
const struct gmr1_fcch_burst gmr1_fcch3_sband_burst = { 0.16f, 12 * 39 };

static struct gmr1_pi4cxpsk_burst *const std_bursts[GMR1B200_BT_COUNT] = {
	&gmr1_bcch_burst, &gmr1_dc2_burst, &gmr1_dc6_burst, &gmr1_dc12_burst, &gmr1_nt3_speech_burst,
	&gmr1_nt3_facch_burst, &gmr1_nt6_burst, &gmr1_nt9_burst, &gmr1_rach_burst, &gmr1_sdcch_burst,
};
static struct gmr1_pi4cxpsk_sync std_sync[GMR1B200_BT_COUNT][GMR1B200_MAX_SYNC][GMR1B200_MAX_SYNC_CHUNK + 1];
static struct gmr1_pi4cxpsk_data std_data[GMR1B200_BT_COUNT][GMR1B200_MAX_DATA_CHUNK + 1];

static void sym_set(struct gmr1_pi4cxpsk_symbol *s, int idx, int b0, int b1, int quarter)
{
	static const float re[4] = { 1, 0, -1, 0 }, im[4] = { 0, 1, 0, -1 };
	s->idx = (short)idx;
	s->data[0] = (ubit_t)b0;
	s->data[1] = (ubit_t)b1;
	s->mod_phase = (float)quarter * PI_F / 2;
	s->mod_val = re[quarter] + im[quarter] * I;
}

/* the data symbols are filled at load time from the library's own descriptor tables */
static void __attribute__((constructor)) compat_init(void)
{
	int b, i, c, k;

	sym_set(&bpsk_syms[0], 0, 0, 0, 0);
	sym_set(&bpsk_syms[1], 1, 1, 0, 2);
	sym_set(&qpsk_syms[0], 0, 0, 0, 0);        /* symbol order: 00 01 11 10 */
	sym_set(&qpsk_syms[1], 1, 0, 1, 1);
	sym_set(&qpsk_syms[2], 2, 1, 1, 2);
	sym_set(&qpsk_syms[3], 3, 1, 0, 3);
	sym_set(&qpsk_bits[0], 0, 0, 0, 0);        /* bit order: 00 01 10 11 */
	sym_set(&qpsk_bits[1], 1, 0, 1, 1);
	sym_set(&qpsk_bits[2], 3, 1, 0, 3);
	sym_set(&qpsk_bits[3], 2, 1, 1, 2);
	gmr1_pi2cbpsk.rotation = PI_F / 2; gmr1_pi2cbpsk.nbits = 1; gmr1_pi2cbpsk.syms = gmr1_pi2cbpsk.bits = bpsk_syms;
	gmr1_pi4cbpsk.rotation = PI_F / 4; gmr1_pi4cbpsk.nbits = 1; gmr1_pi4cbpsk.syms = gmr1_pi4cbpsk.bits = bpsk_syms;
	gmr1_pi4cqpsk.rotation = PI_F / 4; gmr1_pi4cqpsk.nbits = 2; gmr1_pi4cqpsk.syms = qpsk_syms;
	gmr1_pi4cqpsk.bits = qpsk_bits;

	for (b = 0; b < GMR1B200_BT_COUNT; b++) {
		struct gmr1b200_burst_desc d;
		struct gmr1_pi4cxpsk_burst *bt = std_bursts[b];
		gmr1b200_burst_desc_get(b, &d);
		bt->mod = d.nbits == 2 ? &gmr1_pi4cqpsk : (b == GMR1B200_BT_DC12 ? &gmr1_pi2cbpsk : &gmr1_pi4cbpsk);
		bt->guard_pre = 2;
		bt->guard_post = 3;
		bt->len = d.len;
		bt->ebits = d.ebits;
		for (i = 0; i < GMR1_MAX_SYNC; i++) {
			bt->sync[i] = NULL;
			if (i >= d.n_sync)
				continue;
			bt->sync[i] = std_sync[b][i];
			for (c = 0; c < d.n_chunk[i]; c++) {
				struct gmr1_pi4cxpsk_sync *s = &std_sync[b][i][c];
				s->pos = d.s_pos[i][c];
				s->len = d.s_len[i][c];
				for (k = 0; k < s->len; k++)    /* symbol index into mod->syms (BPSK: phase index / 2) */
					s->syms[k] = d.nbits == 1 ? (d.s_sym[i][c][k] >> 1) : d.s_sym[i][c][k];
				s->_ref = NULL;
			}
			std_sync[b][i][c].pos = -1;
			std_sync[b][i][c].len = 0;
		}
		for (c = 0; c < d.n_data; c++) {
			std_data[b][c].pos = d.d_pos[c];
			std_data[b][c].len = d.d_len[c];
		}
		std_data[b][c].pos = -1;
		std_data[b][c].len = 0;
		bt->data = std_data[b];
	}
}

/* struct gmr1_pi4cxpsk_burst (any, also caller-defined) -> flattened descriptor */
static int desc_from_burst(const struct gmr1_pi4cxpsk_burst *bt, struct gmr1b200_burst_desc *d)
{
	int i, c, k;
	const struct gmr1_pi4cxpsk_sync *s;
	const struct gmr1_pi4cxpsk_data *dc;

	if (!bt || !bt->mod || !bt->data)
		return -EINVAL;
	memset(d, 0, sizeof(*d));
	d->rotation = bt->mod->rotation;
	d->nbits = bt->mod->nbits;
	d->len = bt->len;
	d->ebits = bt->ebits;
	for (i = 0; i < GMR1_MAX_SYNC && bt->sync[i]; i++) {
		for (c = 0, s = bt->sync[i]; s->pos >= 0; s++, c++) {
			if (c >= GMR1B200_MAX_SYNC_CHUNK || s->len > GMR1_MAX_SYNC_SYMS)
				return -EINVAL;
			d->s_pos[i][c] = (int16_t)s->pos;
			d->s_len[i][c] = (int16_t)s->len;
			for (k = 0; k < s->len; k++) {
				const float ph = bt->mod->syms[s->syms[k]].mod_phase;
				d->s_sym[i][c][k] = (uint8_t)(((int)lroundf(ph / (PI_F / 2))) & 3);
			}
		}
		d->n_chunk[i] = c;
	}
	d->n_sync = i;
	for (c = 0, dc = bt->data; dc->pos >= 0; dc++, c++) {
		if (c >= GMR1B200_MAX_DATA_CHUNK)
			return -EINVAL;
		d->d_pos[c] = (int16_t)dc->pos;
		d->d_len[c] = (int16_t)dc->len;
	}
	d->n_data = c;
	return 0;
}

/* ------------------------------------------------------------------ sdr/pi4cxpsk.h */

int gmr1_pi4cxpsk_demod(struct gmr1_pi4cxpsk_burst *burst_type, struct osmo_cxvec *burst_in, int sps,
                        float freq_shift, sbit_t *ebits, int *sync_id_p, float *toa_p, float *freq_err_p)
{
	struct gmr1b200_burst_desc d;
	int32_t sid = -1;
	float toa = 0.0f, fe = 0.0f;
	int rv = desc_from_burst(burst_type, &d);
	if (rv)
		return rv;
	rv = HOSTCALL(gmr1b200_pi4cxpsk_demod_desc_batch(&d, (const float *)burst_in->data, burst_in->len, NULL, 0,
	                                        burst_in->len, sps, NULL, freq_shift, ebits, d.ebits,
	                                        &sid, &toa, &fe, NULL, 1, NULL));
	if (rv)
		return rv;
	if (sid < 0)
		return sid;                    /* the reference returns the (negative) sync id, pi4cxpsk.c:549-552 */
	if (sync_id_p) *sync_id_p = sid;
	if (toa_p) *toa_p = toa;
	if (freq_err_p) *freq_err_p = fe;
	return 0;
}

int gmr1_pi4cxpsk_detect(struct gmr1_pi4cxpsk_burst **burst_types, float e_toa, struct osmo_cxvec *burst_in,
                         int sps, float freq_shift, int *bt_id_p, int *sync_id_p, float *toa_p)
{
	struct gmr1b200_burst_desc d[8];
	int32_t bt = -1, sid = -1;
	float toa = 0.0f;
	int n, rv;
	for (n = 0; burst_types[n]; n++) {
		if (n >= 8)
			return -EINVAL;
		if ((rv = desc_from_burst(burst_types[n], &d[n])))
			return rv;
	}
	rv = HOSTCALL(gmr1b200_pi4cxpsk_detect_desc_batch(d, n, NULL, e_toa, (const float *)burst_in->data, burst_in->len,
	                                         NULL, 0, burst_in->len, sps, NULL, freq_shift, &bt, &sid, &toa, 1, NULL));
	if (rv)
		return rv;
	if (bt_id_p) *bt_id_p = bt;
	if (sync_id_p) *sync_id_p = sid;
	if (toa_p) *toa_p = toa;
	return 0;
}

int gmr1_pi4cxpsk_mod_order(struct osmo_cxvec *burst_in, int sps, float freq_shift)
{
	int32_t order = 0;
	int rv = HOSTCALL(gmr1b200_pi4cxpsk_mod_order_batch((const float *)burst_in->data, burst_in->len, NULL, 0, burst_in->len,
	                                           sps, NULL, freq_shift, &order, 1, NULL));
	return rv ? rv : order;
}

/* 1 sample/symbol modulator (pi4cxpsk.c:741-799): guard silent, training symbols, Gray-mapped data,
 * continuous rotation */
int gmr1_pi4cxpsk_mod(struct gmr1_pi4cxpsk_burst *burst_type, ubit_t *ebits, int sync_id, struct osmo_cxvec *burst_out)
{
	const struct gmr1_pi4cxpsk_modulation *mod = burst_type->mod;
	const struct gmr1_pi4cxpsk_sync *s;
	const struct gmr1_pi4cxpsk_data *dc;
	int i, j, k = 0;

	if (burst_out->max_len < burst_type->len)
		return -ENOMEM;
	burst_out->len = burst_type->len;
	for (i = 0; i < burst_type->len; i++)
		burst_out->data[i] = 0.0f;
	for (s = burst_type->sync[sync_id]; s->pos >= 0 && s->len; s++)
		for (i = 0; i < s->len; i++)
			burst_out->data[s->pos + i] = mod->syms[s->syms[i]].mod_val;
	for (dc = burst_type->data; dc->pos >= 0 && dc->len; dc++)
		for (i = 0; i < dc->len; i++) {
			int sym = 0;
			for (j = 0; j < mod->nbits; j++)
				sym = (sym << 1) | ebits[k++];
			burst_out->data[dc->pos + i] = mod->bits[sym].mod_val;
		}
	for (i = 0; i < burst_out->len; i++)
		burst_out->data[i] *= cexpf(I * (mod->rotation * (float)i));
	burst_out->flags = 0;
	return 0;
}

/* ------------------------------------------------------------------ sdr/fcch.h, sdr/dkab.h */

static int fcch_type_of(const struct gmr1_fcch_burst *bt)
{
	if (!bt)
		return -1;
	if (bt->len == 117 && bt->freq == 0.32f) return 0;
	if (bt->len == 468 && bt->freq == 0.32f) return 1;
	if (bt->len == 468 && bt->freq == 0.16f) return 2;
	return -1;
}

int gmr1_fcch_rough(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *win, int sps, float freq_shift, int *toa)
{
	int32_t t = 0;
	int rv = HOSTCALL(gmr1b200_fcch_rough_batch(fcch_type_of(burst_type), (const float *)win->data, win->len, NULL, 0,
	                                   win->len, sps, NULL, freq_shift, &t, NULL, 1, NULL));
	if (!rv)
		*toa = t;
	return rv;
}

int gmr1_fcch_rough_multi(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *win, int sps,
                          float freq_shift, int *toa, int N)
{
	return HOSTCALL(gmr1b200_fcch_rough_multi(fcch_type_of(burst_type), (const float *)win->data, win->len, sps, freq_shift,
	                                 (int32_t *)toa, N, NULL));
}

int gmr1_fcch_fine(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                   int *toa, float *freq_error)
{
	int32_t t = 0;
	float fe = 0.0f;
	int rv;
	if (!burst_type || burst_in->len / sps != burst_type->len)     /* fcch.c:546-551 */
		return -EINVAL;
	rv = HOSTCALL(gmr1b200_fcch_fine_batch(fcch_type_of(burst_type), (const float *)burst_in->data, burst_in->len, NULL, 0,
	                              sps, NULL, freq_shift, &t, &fe, 1, NULL));
	if (!rv) {
		*toa = t;
		*freq_error = fe;
	}
	return rv;
}

int gmr1_fcch_snr(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                  float *snr)
{
	if (!burst_type || burst_in->len / sps != burst_type->len)
		return -EINVAL;
	return HOSTCALL(gmr1b200_fcch_snr_batch(fcch_type_of(burst_type), (const float *)burst_in->data, burst_in->len, NULL, 0,
	                               sps, NULL, freq_shift, snr, 1, NULL));
}

int gmr1_dkab_demod(struct osmo_cxvec *burst_in, int sps, float freq_shift, int p, sbit_t *ebits, float *toa_p)
{
	int32_t rv = 0;
	int rc = HOSTCALL(gmr1b200_dkab_demod_batch((const float *)burst_in->data, burst_in->len, NULL, 0, burst_in->len, sps,
	                                   NULL, freq_shift, NULL, p, ebits, toa_p, &rv, 1, NULL));
	return rc ? rc : rv;
}

/* ------------------------------------------------------------------ l1 decoders / encoders */

int gmr1_bcch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv)
{
	int32_t crc = 1, cv = 0;
	int rc = HOSTCALL(gmr1b200_bcch_decode_batch(l2, bits_e, &cv, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	return rc ? rc : crc;
}

int gmr1_ccch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv)
{
	int32_t crc = 1, cv = 0;
	int rc = HOSTCALL(gmr1b200_ccch_decode_batch(l2, bits_e, &cv, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	return rc ? rc : crc;
}

int gmr1_xch_dc12_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv)
{
	int32_t crc = 1, cv = 0;
	int rc = HOSTCALL(gmr1b200_xch_dc12_decode_batch(l2, bits_e, &cv, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	return rc ? rc : crc;
}

int gmr1_facch3_decode(uint8_t *l2, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph, int *conv_rv)
{
	int32_t crc = 1, cv = 0;
	int rc = HOSTCALL(gmr1b200_facch3_decode_batch(l2, bits_s, bits_e, ciph, &cv, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	return rc ? rc : crc;
}

int gmr1_facch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e,
                       const ubit_t *ciph, int *conv_rv)
{
	int32_t crc = 1, cv = 0;
	int rc = HOSTCALL(gmr1b200_facch9_decode_batch(l2, bits_sacch, bits_status, bits_e, ciph, &cv, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	return rc ? rc : crc;
}

void gmr1_tch3_decode(uint8_t *frame0, uint8_t *frame1, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph,
                      int m, int *conv0_rv, int *conv1_rv)
{
	int32_t c0 = 0, c1 = 0;
	HOSTCALL(gmr1b200_tch3_decode_batch(frame0, frame1, bits_s, bits_e, ciph, m, &c0, &c1, 1, NULL));
	if (conv0_rv) *conv0_rv = c0;
	if (conv1_rv) *conv1_rv = c1;
}

int gmr1_rach_decode(uint8_t *rach, const sbit_t *bits_e, uint8_t sb_mask, int *conv_rv, int *crc_rv)
{
	int32_t crc = 1, cv = 0, c2[2] = { 1, 1 };
	int rc = HOSTCALL(gmr1b200_rach_decode_batch(rach, bits_e, NULL, sb_mask, &cv, c2, &crc, 1, NULL));
	if (conv_rv) *conv_rv = cv;
	if (crc_rv) { crc_rv[0] = c2[0]; crc_rv[1] = c2[1]; }
	return rc ? rc : crc;
}

/* TCH9 keeps the reference's stateful interleaver object: the per-burst byte plumbing and the
 * inter-burst de-interleaver run here exactly as in tch9.c:150-165 (they ARE the state), the
 * intra de-interleave + de-puncture + Viterbi run on the GPU. */
void gmr1_tch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e,
                      enum gmr1_tch9_mode mode, const ubit_t *ciph, struct gmr1_interleaver *il, int *conv_rv)
{
	sbit_t my[658], x[648];
	int32_t cv = 0;
	int i;

	memcpy(my, bits_e, 52);
	memcpy(bits_status, bits_e + 52, 4);
	memcpy(my + 52, bits_e + 56, 606);
	if (ciph)
		for (i = 0; i < 658; i++)
			if (ciph[i])
				my[i] = (sbit_t)(-my[i]);
	memcpy(x, my, 52);
	memcpy(bits_sacch, my + 52, 10);
	memcpy(x + 52, my + 62, 596);
	gmr1_scramble_sbit(x, x, 648);
	gmr1_deinterleave_inter(il, x, x);
	HOSTCALL(gmr1b200_tch9_decode_rows_batch(l2, x, (int)mode, &cv, 1, NULL));
	if (conv_rv) *conv_rv = cv;
}

void gmr1_bcch_encode(ubit_t *bits_e, const uint8_t *l2) { gmr1b200_xcch_encode_batch(0, bits_e, l2, 1); }
void gmr1_ccch_encode(ubit_t *bits_e, const uint8_t *l2) { gmr1b200_xcch_encode_batch(1, bits_e, l2, 1); }
int  gmr1_xch_dc12_encode(ubit_t *bits_e, const uint8_t *l2) { return gmr1b200_xcch_encode_batch(2, bits_e, l2, 1); }
void gmr1_facch3_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *bits_s, const ubit_t *ciph)
{
	gmr1b200_facch3_encode(bits_e, l2, bits_s, ciph);
}
void gmr1_facch9_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *sa, const ubit_t *st, const ubit_t *ciph)
{
	gmr1b200_facch9_encode(bits_e, l2, sa, st, ciph);
}
void gmr1_tch3_encode(ubit_t *bits_e, const uint8_t *f0, const uint8_t *f1, const ubit_t *bits_s, const ubit_t *ciph, int m)
{
	gmr1b200_tch3_encode(bits_e, f0, f1, bits_s, ciph, m);
}
void gmr1_rach_encode(ubit_t *bits_e, const uint8_t *rach, uint8_t sb_mask) { gmr1b200_rach_encode(bits_e, rach, sb_mask); }

/* TCH9 encode with the caller's interleaver object (tch9.c:93-128): the library produces this burst's
 * intra-interleaved bits, the reference's inter-burst interleaver state machine runs on the caller's
 * object, then scrambling and the NT9 framing */
void gmr1_tch9_encode(ubit_t *bits_e, const uint8_t *l2, enum gmr1_tch9_mode mode, const ubit_t *bits_sacch,
                      const ubit_t *bits_status, const ubit_t *ciph, struct gmr1_interleaver *il)
{
	ubit_t x[648], my[658];
	int i;
	gmr1b200_tch9_encode_ep(x, l2, (int)mode);
	gmr1_interleave_inter(il, x, x);
	gmr1_scramble_ubit(x, x, 648);
	memcpy(my, x, 52);
	memcpy(my + 52, bits_sacch, 10);
	memcpy(my + 62, x + 52, 596);
	if (ciph)
		for (i = 0; i < 658; i++)
			my[i] ^= ciph[i];
	memcpy(bits_e, my, 52);
	memcpy(bits_e + 52, bits_status, 4);
	memcpy(bits_e + 56, my + 52, 606);
}

/* ------------------------------------------------------------------ l1 primitives (host) */

void gmr1_interleave_intra(void *out, const void *in, int N)
{
	const uint8_t *i8 = in;
	uint8_t *o8 = out;
	int kc;
	for (kc = 0; kc < 8 * N; kc++)
		o8[N * ((5 * kc) & 7) + (kc >> 3)] = i8[kc];
}

void gmr1_deinterleave_intra(void *out, const void *in, int N)
{
	const uint8_t *i8 = in;
	uint8_t *o8 = out;
	int kc;
	for (kc = 0; kc < 8 * N; kc++)
		o8[kc] = i8[N * ((5 * kc) & 7) + (kc >> 3)];
}

int gmr1_interleaver_init(struct gmr1_interleaver *il, int N, int K)
{
	memset(il, 0, sizeof(*il));
	il->bits_cpp = calloc((size_t)N * K, 1);
	if (!il->bits_cpp)
		return -ENOMEM;
	il->N = N;
	il->K = K;
	return 0;
}

void gmr1_interleaver_fini(struct gmr1_interleaver *il)
{
	free(il->bits_cpp);
	memset(il, 0, sizeof(*il));
}

/* row (n mod N) holds burst n; column jk of the output comes from the burst (jk mod N) back */
void gmr1_interleave_inter(struct gmr1_interleaver *il, void *bits_epp, void *bits_ep)
{
	uint8_t *d = bits_epp;
	int jk;
	memcpy(&il->bits_cpp[(il->n % il->N) * il->K], bits_ep, il->K);
	for (jk = 0; jk < il->K; jk++) {
		const int row = ((il->n % il->N) - (jk % il->N) + il->N) % il->N;
		d[jk] = il->bits_cpp[row * il->K + jk];
	}
	il->n++;
}

void gmr1_deinterleave_inter(struct gmr1_interleaver *il, void *bits_ep, void *bits_epp)
{
	const uint8_t *s = bits_epp;
	int jk;
	for (jk = 0; jk < il->K; jk++) {
		const int row = ((il->n % il->N) - (jk % il->N) + il->N) % il->N;
		il->bits_cpp[row * il->K + jk] = s[jk];
	}
	memcpy(bits_ep, &il->bits_cpp[((il->n + 1) % il->N) * il->K], il->K);
	il->n++;
}

static inline int scr_next(uint16_t *r)
{
	const int b = ((*r >> 14) ^ *r) & 1;
	*r = (uint16_t)((*r << 1) | b);
	return b;
}

void gmr1_scramble_sbit(sbit_t *out, const sbit_t *in, int len)
{
	uint16_t r = 0x4d4b;
	int i;
	for (i = 0; i < len; i++) {
		const sbit_t v = in[i];
		out[i] = scr_next(&r) ? (sbit_t)(-v) : v;
	}
}

void gmr1_scramble_ubit(ubit_t *out, const ubit_t *in, int len)
{
	uint16_t r = 0x4d4b;
	int i;
	for (i = 0; i < len; i++)
		out[i] = in[i] ^ (ubit_t)scr_next(&r);
}

void gmr1_a5(int n, uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul) { gmr1b200_a5(n, key, fn, nbits, dl, ul); }
void gmr1_a5_1(uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul) { gmr1b200_a5(1, key, fn, nbits, dl, ul); }
