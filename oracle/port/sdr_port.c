/* oracle/port/sdr_port.c - plain-C restatement of the reference's SDR/PHY layer (src/sdr/).
 *
 * TEST INFRASTRUCTURE ONLY (see l1_port.c header).  Same float expressions in the same order as the
 * reference so that, built with the same flags (-ffp-contract=off), it reproduces oracle/_ref bit for
 * bit - tests/test_oracle_cpu.py checks exactly that.  libosmo-dsp / FFTW semantics come from
 * oracle/shim (restated, un-vendored third party: parity unpinned there).
 */
#include <complex.h>
#include <errno.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <fftw3.h>
#include <osmocom/core/bits.h>
#include <osmocom/dsp/cxvec.h>
#include <osmocom/dsp/cxvec_math.h>

#include "port.h"

#define SYM_RATE 23400      /* include/osmocom/gmr1/sdr/defs.h:33 */

/* ------------------------------------------------------------------ modulations: pi4cxpsk.c:71-115 */
static struct gmr1_pi4cxpsk_symbol bpsk_syms[] = {
	{ 0, {0}, 0*M_PIf/2,  1+0*I }, { 1, {1}, 2*M_PIf/2, -1+0*I },
};
static struct gmr1_pi4cxpsk_symbol qpsk_syms[] = {
	{ 0, {0,0}, 0*M_PIf/2,  1+0*I }, { 1, {0,1}, 1*M_PIf/2,  0+1*I },
	{ 2, {1,1}, 2*M_PIf/2, -1+0*I }, { 3, {1,0}, 3*M_PIf/2,  0-1*I },
};
static struct gmr1_pi4cxpsk_symbol qpsk_bits[] = {
	{ 0, {0,0}, 0*M_PIf/2,  1+0*I }, { 1, {0,1}, 1*M_PIf/2,  0+1*I },
	{ 3, {1,0}, 3*M_PIf/2,  0-1*I }, { 2, {1,1}, 2*M_PIf/2, -1+0*I },
};
struct gmr1_pi4cxpsk_modulation gmr1_pi2cbpsk = { M_PIf/2, 1, bpsk_syms, bpsk_syms };
struct gmr1_pi4cxpsk_modulation gmr1_pi4cbpsk = { M_PIf/4, 1, bpsk_syms, bpsk_syms };
struct gmr1_pi4cxpsk_modulation gmr1_pi4cqpsk = { M_PIf/4, 2, qpsk_syms, qpsk_bits };

/* ------------------------------------------------------------------ burst formats: nb.c:34-377
 * (ETSI TS 101 376-5-2 section 7.4), written as "position:symbols" strings and expanded at load */
struct gmr1_pi4cxpsk_burst gmr1_bcch_burst, gmr1_dc2_burst, gmr1_dc6_burst, gmr1_dc12_burst,
	gmr1_nt3_speech_burst, gmr1_nt3_facch_burst, gmr1_nt6_burst, gmr1_nt9_burst, gmr1_rach_burst, gmr1_sdcch_burst;

struct fmt { struct gmr1_pi4cxpsk_burst *bt; struct gmr1_pi4cxpsk_modulation *mod; int len, ebits;
             const char *sync[4]; const char *data; };
#define S32 "22222222222222222222222222222222"
static const struct fmt FORMATS[] = {
	{ &gmr1_bcch_burst, &gmr1_pi4cqpsk, 234, 424, { "28:02200020222 119:220 197:220" }, "2:26 39:80 122:75 200:31" },
	{ &gmr1_dc2_burst, &gmr1_pi4cqpsk, 78, 132, { "28:0123030" }, "2:26 35:40" },
	{ &gmr1_dc6_burst, &gmr1_pi4cqpsk, 234, 432, { "28:0002202 119:030 197:311" }, "2:26 35:84 122:75 200:31" },
	{ &gmr1_dc12_burst, &gmr1_pi2cbpsk, 468, 432, { "10:0010001111 228:00100011101 447:0010001111" }, "2:8 20:208 239:208 457:8" },
	{ &gmr1_nt3_speech_burst, &gmr1_pi4cqpsk, 117, 212, { "28:033123" }, "2:26 34:80" },
	{ &gmr1_nt3_facch_burst, &gmr1_pi4cbpsk, 117, 104, { "28:10101010", "28:11001001" }, "2:26 36:78" },
	{ &gmr1_nt6_burst, &gmr1_pi4cqpsk, 234, 434, { "28:022323 119:010 197:230", "28:000220 119:130 197:213" }, "2:26 34:85 122:75 200:31" },
	{ &gmr1_nt9_burst, &gmr1_pi4cqpsk, 351, 662, { "28:022323 119:122 197:010 275:230", "28:000220 119:020 197:130 275:213" },
	  "2:26 34:85 122:75 200:75 278:70" },
	{ &gmr1_rach_burst, &gmr1_pi4cqpsk, 351, 494, { "78:02200020222220220 127:" S32 " 191:" S32 " 255:02200020222220220 347:0" },
	  "2:76 95:32 159:32 223:32 272:75" },
	{ &gmr1_sdcch_burst, &gmr1_pi4cbpsk, 234, 208, { "28:0101010 115:1010101 197:0101011", "28:0011001 115:1001100 197:1100111",
	  "28:0000111 115:1000011 197:1100001", "28:0110100 115:1011010 197:0101101" }, "2:26 35:80 122:75 204:27" },
};

static void __attribute__((constructor)) sdr_port_init(void)
{
	unsigned f;
	for (f = 0; f < sizeof(FORMATS) / sizeof(FORMATS[0]); f++) {
		const struct fmt *F = &FORMATS[f];
		struct gmr1_pi4cxpsk_burst *bt = F->bt;
		int i, n;
		const char *p;
		bt->mod = F->mod; bt->guard_pre = 2; bt->guard_post = 3; bt->len = F->len; bt->ebits = F->ebits;
		for (i = 0; i < 4; i++) {
			struct gmr1_pi4cxpsk_sync *s;
			bt->sync[i] = NULL;
			if (!F->sync[i])
				continue;
			s = bt->sync[i] = calloc(8, sizeof(*s));
			for (p = F->sync[i], n = 0; *p; n++) {
				s[n].pos = (int)strtol(p, (char **)&p, 10);
				p++;                                    /* ':' */
				for (s[n].len = 0; *p && *p != ' '; p++)
					s[n].syms[s[n].len++] = (uint8_t)(*p - '0');
				while (*p == ' ')
					p++;
			}
			s[n].pos = -1;
		}
		bt->data = calloc(8, sizeof(*bt->data));
		for (p = F->data, n = 0; *p; n++) {
			bt->data[n].pos = (int)strtol(p, (char **)&p, 10);
			p++;
			bt->data[n].len = (int)strtol(p, (char **)&p, 10);
			while (*p == ' ')
				p++;
		}
		bt->data[n].pos = -1;
	}
}

/* ------------------------------------------------------------------ pi4cxpsk.c:126-171 */
static int sync_gen_ref(struct gmr1_pi4cxpsk_burst *bt)
{
	int i, j;
	for (i = 0; i < 4 && bt->sync[i]; i++) {
		struct gmr1_pi4cxpsk_sync *cs;
		for (cs = bt->sync[i]; cs->pos >= 0; cs++) {
			int is_real = 1;
			if (cs->_ref)
				continue;
			cs->_ref = osmo_cxvec_alloc(cs->len);
			if (!cs->_ref)
				return -ENOMEM;
			for (j = 0; j < cs->len; j++) {
				float complex mv = bt->mod->syms[cs->syms[j]].mod_val;
				if (cimagf(mv) != 0.0f)
					is_real = 0;
				cs->_ref->data[j] = mv;
			}
			cs->_ref->len = cs->len;
			if (is_real)
				cs->_ref->flags |= CXVEC_FLG_REAL_ONLY;
		}
	}
	return 0;
}

/* pi4cxpsk.c:184-268 - note the accumulator is cleared once, before the loop over sequences (:207) */
static int sync_find(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst, int sps, float *toa, float *pwr)
{
	struct osmo_cxvec win, *corr, *tmp;
	int i, j, w = burst->len - (bt->len * sps) + 1;
	float p_toa = 0.0f, p_pwr = 0.0f, p_idx = -1;
	corr = osmo_cxvec_alloc(w);
	tmp = osmo_cxvec_alloc(w);
	if (!corr || !tmp) {
		osmo_cxvec_free(tmp); osmo_cxvec_free(corr);
		return -ENOMEM;
	}
	memset(corr->data, 0, sizeof(float complex) * corr->max_len);
	corr->len = w;
	for (i = 0; i < 4 && bt->sync[i]; i++) {
		struct gmr1_pi4cxpsk_sync *cs;
		float s_toa, s_pwr;
		float complex s_peak;
		int tl = 0;
		if (i > 0 && getenv("GMR1_ORACLE_SYNC_RESET"))     /* opt-in, NOT the reference: see gmr1_b200.h */
			memset(corr->data, 0, sizeof(float complex) * corr->max_len);
		for (cs = bt->sync[i]; cs->pos >= 0; cs++) {
			osmo_cxvec_init_from_data(&win, &burst->data[cs->pos * sps], (cs->len * sps) + w - 1);
			osmo_cxvec_correlate(cs->_ref, &win, sps, tmp);
			for (j = 0; j < w; j++)
				corr->data[j] += cabsf(tmp->data[j]);
			tl += cs->_ref->len;
		}
		s_toa = osmo_cxvec_peak_energy_find(corr, 3, PEAK_EARLY_LATE, &s_peak);
		s_peak /= (float)tl;
		s_pwr = osmo_normsqf(s_peak);
		if (s_pwr > p_pwr) {
			p_pwr = s_pwr; p_toa = s_toa; p_idx = i;
		}
	}
	if (toa) *toa = p_toa;
	if (pwr) *pwr = p_pwr;
	osmo_cxvec_free(tmp); osmo_cxvec_free(corr);
	return p_idx;
}

/* pi4cxpsk.c:280-348 */
static void align(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst, int sps, float toa)
{
	int i;
	if (sps >= 4) {
		int d = roundf(toa);
		for (i = 0; i < bt->len; i++)
			burst->data[i] = burst->data[i * sps + d];
	} else {
		struct osmo_cxvec *conv = NULL, *src = burst;
		int ofs_int = roundf(toa);
		float ofs_frac = toa - ofs_int;
		if (fabs(ofs_frac) > 0.1f) {
			float complex taps[21];
			struct osmo_cxvec pulse;
			for (i = 0; i < 21; i++)
				taps[i] = osmo_sinc(M_PIf * ((float)(i - 10) + ofs_frac));
			osmo_cxvec_init_from_data(&pulse, taps, 21);
			pulse.flags |= CXVEC_FLG_REAL_ONLY;
			src = conv = osmo_cxvec_convolve(&pulse, burst, CONV_NO_DELAY, NULL);
		}
		for (i = 0; i < bt->len; i++) {
			int j = (i * sps) + ofs_int;
			burst->data[i] = (j < 0 || j >= src->len) ? 0.0f : src->data[j];
		}
		if (conv)
			osmo_cxvec_free(conv);
	}
	burst->len = bt->len;
}

/* pi4cxpsk.c:360-406 */
static float freq_err(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst, int sync_id)
{
	struct gmr1_pi4cxpsk_sync *cs;
	int n, i, j;
	for (n = 0, cs = bt->sync[sync_id]; cs->pos >= 0; n++, cs++);
	if (n <= 1)
		return 0.0f;
	{
		float complex corr[n];
		float pos[n], f = 0.0f;
		for (i = 0; i < n; i++) {
			cs = &bt->sync[sync_id][i];
			corr[i] = 0.0f;
			pos[i] = (float)cs->pos + (float)cs->len / 2.0f;
			for (j = 0; j < cs->len; j++)
				corr[i] += conjf(cs->_ref->data[j]) * burst->data[cs->pos + j];
		}
		for (i = 1; i < n; i++)
			f += cargf(corr[i] * conjf(corr[i - 1])) / (pos[i] - pos[i - 1]);
		return f / (n - 1);
	}
}

/* pi4cxpsk.c:415-433 */
static float complex phase_ref(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst, int sync_id)
{
	struct gmr1_pi4cxpsk_sync *cs;
	float complex corr = 0.0f;
	int i;
	for (cs = bt->sync[sync_id]; cs->pos >= 0; cs++)
		for (i = 0; i < cs->len; i++)
			corr += conjf(cs->_ref->data[i]) * burst->data[cs->pos + i];
	return corr / cabsf(corr);
}

/* pi4cxpsk.c:442-503 */
static void soft_bits(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst, sbit_t *ebits)
{
	struct gmr1_pi4cxpsk_modulation *mod = bt->mod;
	struct gmr1_pi4cxpsk_data *dc;
	int mask = (1 << mod->nbits) - 1, i, j, k = 0;
	float d = (2.0f * M_PIf) / (1 << mod->nbits);
	for (dc = bt->data; dc->pos >= 0; dc++)
		for (i = dc->pos; i < dc->pos + dc->len; i++) {
			float sv = cargf(burst->data[i]) / d, svr = roundf(sv);
			int sp = (int)svr & mask, ss = (svr > sv ? (sp - 1) : (sp + 1)) & mask;
			int dq = roundf((2.0f * fabs(svr - sv)) * 64.0f);
			for (j = 0; j < mod->nbits; j++) {
				uint8_t vp = mod->syms[sp].data[j], vs = mod->syms[ss].data[j];
				sbit_t v = 127 - ((vp ^ vs) ? dq : (dq >> 1));
				ebits[k++] = vp ? -v : v;
			}
		}
}

/* pi4cxpsk.c:520-602 */
int gmr1_pi4cxpsk_demod(struct gmr1_pi4cxpsk_burst *bt, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                        sbit_t *ebits, int *sync_id_p, float *toa_p, float *freq_err_p)
{
	struct osmo_cxvec *burst;
	float toa, ferr;
	float complex phasor;
	int sync_id, rv;
	if ((rv = sync_gen_ref(bt)))
		return rv;
	burst = osmo_cxvec_sig_normalize(burst_in, 1, (freq_shift - bt->mod->rotation) / sps, NULL);
	if (!burst)
		return -ENOMEM;
	sync_id = sync_find(bt, burst, sps, &toa, NULL);
	if (sync_id < 0) {
		osmo_cxvec_free(burst);
		return sync_id;
	}
	if (sync_id_p) *sync_id_p = sync_id;
	if (toa_p) *toa_p = toa;
	align(bt, burst, sps, toa);
	ferr = freq_err(bt, burst, sync_id);
	if (freq_err_p) *freq_err_p = ferr;
	if (ferr != 0.0f)
		osmo_cxvec_rotate(burst, -ferr, burst);
	phasor = phase_ref(bt, burst, sync_id);
	osmo_cxvec_scale(burst, conjf(phasor), burst);
	soft_bits(bt, burst, ebits);
	osmo_cxvec_free(burst);
	return 0;
}

/* pi4cxpsk.c:617-682 */
int gmr1_pi4cxpsk_detect(struct gmr1_pi4cxpsk_burst **bts, float e_toa, struct osmo_cxvec *burst_in, int sps,
                         float freq_shift, int *bt_id_p, int *sync_id_p, float *toa_p)
{
	struct osmo_cxvec *burst;
	int id, p_id = -1, p_sid = -1, rv = 0;
	float p_toa = 0.0f, p_pwr = 0.0f;
	burst = osmo_cxvec_sig_normalize(burst_in, 1, (freq_shift - bts[0]->mod->rotation) / sps, NULL);
	if (!burst)
		return -ENOMEM;
	for (id = 0; bts[id]; id++) {
		float toa, pwr;
		int sid;
		if ((rv = sync_gen_ref(bts[id])))
			break;
		sid = sync_find(bts[id], burst, sps, &toa, &pwr);
		if (sid < 0) {
			rv = sid;
			break;
		}
		if (e_toa >= 0.0f)
			pwr /= fabs(e_toa - toa);
		if (pwr > p_pwr) {
			p_id = id; p_sid = sid; p_pwr = pwr; p_toa = toa;
		}
	}
	if (!rv) {
		if (bt_id_p) *bt_id_p = p_id;
		if (sync_id_p) *sync_id_p = p_sid;
		if (toa_p) *toa_p = p_toa;
	}
	osmo_cxvec_free(burst);
	return rv;
}

/* pi4cxpsk.c:693-729 */
int gmr1_pi4cxpsk_mod_order(struct osmo_cxvec *burst_in, int sps, float freq_shift)
{
	struct osmo_cxvec *burst = osmo_cxvec_sig_normalize(burst_in, 1, (freq_shift - (M_PIf/4)) / sps, NULL);
	float complex sb = 0.0f, sq = 0.0f;
	int i, rv;
	if (!burst)
		return -ENOMEM;
	for (i = 0; i < burst->len; i++) {
		float complex v = burst->data[i];
		v = (v * v) / osmo_normsqf(v);
		sb += v;
		sq += v * v;
	}
	rv = osmo_normsqf(sb) < (osmo_normsqf(sq) / 2.0f) ? 4 : 2;
	osmo_cxvec_free(burst);
	return rv;
}

/* pi4cxpsk.c:741-799 */
int gmr1_pi4cxpsk_mod(struct gmr1_pi4cxpsk_burst *bt, ubit_t *ebits, int sync_id, struct osmo_cxvec *out)
{
	struct gmr1_pi4cxpsk_sync *s;
	struct gmr1_pi4cxpsk_data *d;
	int i, j, k = 0, rv;
	if (out->max_len < bt->len)
		return -ENOMEM;
	out->len = bt->len;
	if ((rv = sync_gen_ref(bt)))
		return rv;
	for (i = 0; i < bt->guard_pre; i++)
		out->data[i] = 0.0f;
	for (i = 0; i < bt->guard_post; i++)
		out->data[out->len - i - 1] = 0.0f;
	for (s = bt->sync[sync_id]; s->len; s++)
		for (i = 0; i < s->len; i++)
			out->data[s->pos + i] = s->_ref->data[i];
	for (d = bt->data; d->len; d++)
		for (i = 0; i < d->len; i++) {
			int sym = 0;
			for (j = 0; j < bt->mod->nbits; j++)
				sym = (sym << 1) | ebits[k++];
			out->data[d->pos + i] = bt->mod->bits[sym].mod_val;
		}
	osmo_cxvec_rotate(out, bt->mod->rotation, out);
	return 0;
}

/* ------------------------------------------------------------------ FCCH: fcch.c */
const struct gmr1_fcch_burst gmr1_fcch_burst = { 0.32f, 3 * 39 };           /* :50-70 */
const struct gmr1_fcch_burst gmr1_fcch3_lband_burst = { 0.32f, 12 * 39 };
const struct gmr1_fcch_burst gmr1_fcch3_sband_burst = { 0.16f, 12 * 39 };

/* kind 0: dual (real) chirp, +1: up, -1: down  (:92-193) */
static struct osmo_cxvec *gen_chirp(const struct gmr1_fcch_burst *bt, int sps, int kind)
{
	int i, l = bt->len * sps;
	struct osmo_cxvec *cv = osmo_cxvec_alloc(l);
	float phase_base = bt->freq * 2.0f * M_PIf / (float)(bt->len), halfpos = ((float)(bt->len)) / 2.0f;
	if (!cv)
		return NULL;
	cv->len = l;
	if (kind == 0)
		cv->flags |= CXVEC_FLG_REAL_ONLY;
	if (kind < 0)
		phase_base *= -1.0f;
	for (i = 0; i < l; i++) {
		float pos = ((float)i / (float)sps) - halfpos;
		if (kind == 0)
			cv->data[i] = sqrtf(2.0f) * cosf(phase_base * (pos * pos));
		else
			cv->data[i] = (sqrtf(2.0f) / 2.0f) * cexpf(I * phase_base * (pos * pos));
	}
	return cv;
}

int gmr1_fcch_rough(const struct gmr1_fcch_burst *bt, struct osmo_cxvec *win_in, int sps, float freq_shift, int *toa)  /* :211-250 */
{
	struct osmo_cxvec *ref = gen_chirp(bt, 1, 0), *win, *corr;
	float pos;
	if (!ref)
		return -ENOMEM;
	win = osmo_cxvec_sig_normalize(win_in, sps, freq_shift, NULL);
	corr = osmo_cxvec_correlate(ref, win, 1, NULL);
	pos = osmo_cxvec_peak_energy_find(corr, 5, PEAK_WEIGH_WIN, NULL);
	*toa = (int)round(pos * sps);
	osmo_cxvec_free(corr); osmo_cxvec_free(win); osmo_cxvec_free(ref);
	return 0;
}

static void peak_record(const struct gmr1_fcch_burst *bt, int *toa, float *pwr, int *n, int N, int Lp, int sps,
                        int peak_toa, float peak_pwr)                                    /* :264-326 */
{
	int i, j, has_dupe = 0;
	for (i = 0; i < *n; i++) {
		int th = (bt->len * sps) >> 1, d = (toa[i] % Lp) - (peak_toa % Lp);
		if (abs(d) > th)
			continue;
		if (pwr[i] > peak_pwr) {
			if (!has_dupe)
				has_dupe = 1;
			continue;
		}
		for (j = i; j < (*n) - 1; j++) { toa[j] = toa[j + 1]; pwr[j] = pwr[j + 1]; }
		*n = *n - 1;
		has_dupe = -1;
	}
	if (has_dupe > 0)
		return;
	for (i = 0; i < *n; i++)
		if (peak_pwr > pwr[i])
			break;
	if (i == N)
		return;
	for (j = N - 1; j > i; j--) { toa[j] = toa[j - 1]; pwr[j] = pwr[j - 1]; }
	toa[i] = peak_toa; pwr[i] = peak_pwr;
	if (*n != N)
		*n = *n + 1;
}

int gmr1_fcch_rough_multi(const struct gmr1_fcch_burst *bt, struct osmo_cxvec *win_in, int sps, float freq_shift,
                          int *peaks_toa, int N)                                         /* :341-496 */
{
	struct osmo_cxvec *ref, *win, *corr;
	float *cp, pwr_max = 0.0f, pwrs[2] = {0, 0}, peaks[2] = {0, 0}, avg = 0.0f, stddev = 0.0f, th, peaks_pwr[N];
	int Lw, Lp, nLp, i, a, pwr_max_idx = 0, cnt = 0, rv;
	if (win_in->len < ((650 * SYM_RATE * sps) / 1000))
		return -EINVAL;
	ref = gen_chirp(bt, 1, 0);
	win = osmo_cxvec_sig_normalize(win_in, sps, freq_shift, NULL);
	corr = osmo_cxvec_correlate(ref, win, 1, NULL);
	cp = malloc(sizeof(float) * corr->len);
	Lw = (320 * SYM_RATE) / 1000 + bt->len;
	Lp = (320 * SYM_RATE) / 1000;
	for (i = 0; i < corr->len; i++) {
		float e = osmo_normsqf(corr->data[i]);
		cp[i] = e;
		if ((e > pwr_max) && (i < Lw)) { pwr_max = e; pwr_max_idx = i; }
	}
	for (i = -10; i <= 10; i++) {
		int j = pwr_max_idx + i;
		if ((j > 0) && (j < corr->len)) { pwrs[0] += cp[j]; peaks[0] += cp[j] * j; }
		j += Lp;
		if ((j > 0) && (j < corr->len)) { pwrs[1] += cp[j]; peaks[1] += cp[j] * j; }
	}
	peaks[0] /= pwrs[0];
	peaks[1] /= pwrs[1];
	nLp = (int)round(peaks[1] - peaks[0]);
	if (abs(nLp - Lp) > 10) {
		rv = -EINVAL;
		goto done;
	}
	Lp = nLp;
	for (i = 0; i < Lw; i++) {
		float v = sqrtf(cp[i] * cp[i + Lp]);
		cp[i] = v;
		avg += v;
	}
	avg /= Lw;
	for (i = 0; i < Lw; i++) {
		float v = cp[i] - avg;
		stddev += v * v;
	}
	stddev = sqrtf(stddev / Lw);
	th = avg + 3.0f * stddev;
	for (i = 1, a = 0; i < Lw - 1; i++) {
		if (cp[i] > th) {
			float p_pwr, p_fpos;
			if (a)
				continue;
			a = 1;
			p_pwr = cp[i - 1] + cp[i] + cp[i + 1];
			p_fpos = (-cp[i - 1] + cp[i + 1]) / p_pwr;
			peak_record(bt, peaks_toa, peaks_pwr, &cnt, N, Lp, sps, (int)round((i + p_fpos) * sps), p_pwr);
		} else
			a = 0;
	}
	rv = cnt;
done:
	free(cp);
	osmo_cxvec_free(corr); osmo_cxvec_free(win); osmo_cxvec_free(ref);
	return rv;
}

int gmr1_fcch_fine(const struct gmr1_fcch_burst *bt, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                   int *toa, float *freq_error)                                          /* :512-628 */
{
	struct osmo_cxvec *up = gen_chirp(bt, 1, +1), *down = gen_chirp(bt, 1, -1), *burst, *mu, *md;
	fftwf_plan plan;
	float bin_hz, peak_up, peak_down, freq_err_hz, chirp_rate, toa_ms, toa_samples;
	int len = bt->len, mid, i, rv = 0;
	burst = osmo_cxvec_sig_normalize(burst_in, sps, freq_shift, NULL);
	mu = osmo_cxvec_alloc(len);
	md = osmo_cxvec_alloc(len);
	if (len != burst->len) {
		rv = -EINVAL;
		goto done;
	}
	for (i = 0; i < len; i++) {
		mu->data[i] = burst->data[i] * up->data[i];
		md->data[i] = burst->data[i] * down->data[i];
	}
	mu->len = md->len = len;
	mid = (float)(len >> 1);
	for (i = 0; i < len; i++) {
		float complex fs = cexp(I * 2.0f * M_PIf * mid / (float)(len) * i);
		mu->data[i] *= fs;
		md->data[i] *= fs;
	}
	plan = fftwf_plan_dft_1d(len, mu->data, mu->data, FFTW_FORWARD, FFTW_ESTIMATE);
	fftwf_execute(plan); fftwf_destroy_plan(plan);
	plan = fftwf_plan_dft_1d(len, md->data, md->data, FFTW_FORWARD, FFTW_ESTIMATE);
	fftwf_execute(plan); fftwf_destroy_plan(plan);
	peak_up = osmo_cxvec_peak_energy_find(mu, 5, PEAK_WEIGH_WIN, NULL);
	peak_down = osmo_cxvec_peak_energy_find(md, 5, PEAK_WEIGH_WIN, NULL);
	bin_hz = (float)SYM_RATE / (float)len;
	peak_up = (peak_up - mid) * bin_hz;
	peak_down = (peak_down - mid) * bin_hz;
	freq_err_hz = (peak_up + peak_down) / 2.0f;
	*freq_error = (2.0f * M_PIf * freq_err_hz) / SYM_RATE;
	chirp_rate = (2.0f * bt->freq * SYM_RATE * SYM_RATE) / (float)(bt->len * 1000);
	toa_ms = ((peak_up - peak_down) / 2.0f) / chirp_rate;
	toa_samples = (toa_ms * SYM_RATE * sps) / 1000.0f;
	*toa = (int)round(toa_samples);
done:
	osmo_cxvec_free(md); osmo_cxvec_free(mu); osmo_cxvec_free(burst); osmo_cxvec_free(down); osmo_cxvec_free(up);
	return rv;
}

int gmr1_fcch_snr(const struct gmr1_fcch_burst *bt, struct osmo_cxvec *burst_in, int sps, float freq_shift, float *snr)  /* :643-708 */
{
	struct osmo_cxvec *ref = gen_chirp(bt, 1, 0), *burst = osmo_cxvec_sig_normalize(burst_in, sps, freq_shift, NULL);
	fftwf_plan plan;
	int pk[6], len = bt->len, i, rv = 0;
	if (len != burst->len) {
		rv = -EINVAL;
		goto done;
	}
	for (i = 0; i < len; i++)
		burst->data[i] *= crealf(ref->data[i]);
	plan = fftwf_plan_dft_1d(len, burst->data, burst->data, FFTW_FORWARD, FFTW_ESTIMATE);
	fftwf_execute(plan); fftwf_destroy_plan(plan);
	osmo_cxvec_peaks_scan(burst, pk, 6);
	*snr = (osmo_normsqf(burst->data[pk[0]]) + osmo_normsqf(burst->data[pk[1]])) /
	       (osmo_normsqf(burst->data[pk[4]]) + osmo_normsqf(burst->data[pk[5]]));
done:
	osmo_cxvec_free(burst); osmo_cxvec_free(ref);
	return rv;
}

/* ------------------------------------------------------------------ DKAB: dkab.c:57-214 */
int gmr1_dkab_demod(struct osmo_cxvec *burst_in, int sps, float freq_shift, int p, sbit_t *ebits, float *toa_p)
{
	struct osmo_cxvec *burst = osmo_cxvec_sig_normalize(burst_in, 1, (freq_shift - (M_PIf/4)) / sps, NULL);
	int w, i, ofs[2], d, mi, toa_i, l_valley, rv;
	float mp, toa, egy_peak = 0.0f, egy_valley = 0.0f, *pwr;
	if (!burst)
		return -ENOMEM;
	w = burst->len - (39 * 3 * sps) + 1;
	if (w <= 0) {
		osmo_cxvec_free(burst);
		return -EINVAL;
	}
	pwr = malloc(sizeof(float) * w);
	ofs[0] = sps * (2 + p); ofs[1] = sps * (2 + p + 59); d = sps * 5;
	pwr[0] = 0.0f;
	for (i = 0; i < d; i++)
		pwr[0] += osmo_normsqf(burst->data[ofs[0] + i]) + osmo_normsqf(burst->data[ofs[1] + i]);
	mi = 0; mp = pwr[0];
	for (i = 0; i < w - 1; i++) {
		float np = pwr[i] - osmo_normsqf(burst->data[ofs[0] + i]) - osmo_normsqf(burst->data[ofs[1] + i])
		                  + osmo_normsqf(burst->data[ofs[0] + d + i]) + osmo_normsqf(burst->data[ofs[1] + d + i]);
		pwr[i + 1] = np;
		if (np > mp) { mi = i + 1; mp = np; }
	}
	toa = (float)mi;
	if ((mi > 0) && (mi < (w - 1)))
		toa += 0.5f * (-pwr[mi - 1] + pwr[mi + 1]) / (-pwr[mi - 1] + 2.0f * pwr[mi] - pwr[mi + 1]);
	toa += ((float)(sps - 1)) / 2.0f;
	*toa_p = toa;
	toa_i = (int)roundf(toa);
	for (i = 0; i < d; i++)
		egy_peak += osmo_normsqf(burst->data[toa_i + ofs[0] + i]) + osmo_normsqf(burst->data[toa_i + ofs[1] + i]);
	egy_peak /= d * 2;
	l_valley = ofs[1] - ofs[0] - d;
	for (i = 0; i < l_valley; i++)
		egy_valley += osmo_normsqf(burst->data[toa_i + ofs[0] + d + i]);
	egy_valley /= l_valley;
	rv = ((egy_peak / egy_valley) > 10.0f) ? 0 : 1;
	free(pwr);
	if (!rv) {
		int o0 = toa_i + sps * (2 + p), o1 = toa_i + sps * (2 + p + 59);
		for (i = 0; i < 8; i++) {
			int o = (i >> 2 ? o1 : o0) + sps * (i & 3);
			float pd = cargf(burst->data[o] * conjf(burst->data[o + sps]));
			ebits[i] = (sbit_t)roundf((0.5f - (fabsf(pd) / M_PIf)) * 254.0f);
		}
	}
	osmo_cxvec_free(burst);
	return rv;
}
