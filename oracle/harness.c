/* oracle/harness.c - batch loops around the reference's per-burst C functions.  TEST INFRASTRUCTURE ONLY.
 *
 * Linked into BOTH oracle/_ref/libgmr1_ref.so (the reference's own sources) and oracle/liboracle.so (the port), it
 * calls nothing but the reference's public entry points - gmr1_pi4cxpsk_demod (src/sdr/pi4cxpsk.c:520),
 * gmr1_{bcch,ccch,facch3,facch9,tch3,tch9,rach}_decode (src/l1/), gmr1_fcch_rough / _fine (src/sdr/fcch.c:211,512),
 * gmr1_interleaver_init (src/l1/interleave.c) - on n windows per call, the way gmr1_rx.c strings them together
 * (rx_bcch :747-783, rx_ccch :797-851, rx_tch3 :538-600, rx_tch9 :276-355, fcch_single_init :606-639).
 * It exists so that the CPU baseline of bench.py and the full-size parity tests run compiled loops (no Python per
 * burst).  Contains no reference code. */
#include <complex.h>
#include <stdint.h>
#include <string.h>

#include <osmocom/core/bits.h>
#include <osmocom/dsp/cxvec.h>

struct gmr1_pi4cxpsk_burst;
struct gmr1_fcch_burst;
struct gmr1_interleaver_h { int N, K, n; uint8_t *bits_cpp; };      /* include/osmocom/gmr1/l1/interleave.h:43-49 */

extern struct gmr1_pi4cxpsk_burst gmr1_bcch_burst, gmr1_dc6_burst, gmr1_nt3_speech_burst, gmr1_nt3_facch_burst,
	gmr1_nt9_burst, gmr1_rach_burst;
extern struct gmr1_fcch_burst gmr1_fcch_burst;

int gmr1_pi4cxpsk_demod(struct gmr1_pi4cxpsk_burst *burst_type, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                        sbit_t *ebits, int *sync_id_p, float *toa_p, float *freq_err_p);
int gmr1_fcch_rough(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *search_win_in, int sps,
                    float freq_shift, int *toa);
int gmr1_fcch_fine(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *burst_in, int sps, float freq_shift,
                   int *toa, float *freq_error);
int gmr1_bcch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv);
int gmr1_ccch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv);
int gmr1_facch3_decode(uint8_t *l2, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph, int *conv_rv);
int gmr1_facch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e, const ubit_t *ciph,
                       int *conv_rv);
void gmr1_tch3_decode(uint8_t *frame0, uint8_t *frame1, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph, int m,
                      int *conv0_rv, int *conv1_rv);
void gmr1_tch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e, int mode,
                      const ubit_t *ciph, struct gmr1_interleaver_h *il, int *conv_rv);
int gmr1_rach_decode(uint8_t *rach, const sbit_t *bits_e, uint8_t sb_mask, int *conv_rv, int *crc_rv);
int gmr1_interleaver_init(struct gmr1_interleaver_h *il, int N, int K);
void gmr1_interleaver_fini(struct gmr1_interleaver_h *il);

static void view(struct osmo_cxvec *cv, const float complex *x, int len)
{
	cv->len = cv->max_len = len;
	cv->flags = 0;
	cv->data = (float complex *)x;
}

/* BCCH (is_ccch = 0) / DC6-CCCH windows -> L2, CRC, sync id, TOA */
void oh_xcch(int is_ccch, const float complex *iq, int n, int wl, int sps, float freq_shift, uint8_t *l2, int32_t *crc,
             int32_t *conv, float *toa)
{
	for (int i = 0; i < n; i++) {
		struct osmo_cxvec cv;
		sbit_t eb[432];
		int sid = -1, cr = 0;
		float t = 0.0f;
		view(&cv, iq + (size_t)i * wl, wl);
		int rv = gmr1_pi4cxpsk_demod(is_ccch ? &gmr1_dc6_burst : &gmr1_bcch_burst, &cv, sps, freq_shift, eb, &sid, &t, NULL);
		if (rv)
			memset(eb, 0, sizeof(eb));
		crc[i] = is_ccch ? gmr1_ccch_decode(l2 + 24 * (size_t)i, eb, &cr) : gmr1_bcch_decode(l2 + 24 * (size_t)i, eb, &cr);
		if (conv) conv[i] = cr;
		if (toa) toa[i] = t;
	}
}

/* NT3 speech windows -> two TCH3 frames + status bits (ciph: n x 208 cipher bits or NULL) */
void oh_tch3(const float complex *iq, int n, int wl, int sps, const uint8_t *ciph, int m, uint8_t *f0, uint8_t *f1,
             uint8_t *bits_s, int32_t *conv)
{
	for (int i = 0; i < n; i++) {
		struct osmo_cxvec cv;
		sbit_t eb[212];
		int sid = -1, c0 = 0, c1 = 0;
		view(&cv, iq + (size_t)i * wl, wl);
		if (gmr1_pi4cxpsk_demod(&gmr1_nt3_speech_burst, &cv, sps, 0.0f, eb, &sid, NULL, NULL))
			memset(eb, 0, sizeof(eb));
		gmr1_tch3_decode(f0 + 10 * (size_t)i, f1 + 10 * (size_t)i, bits_s + 4 * (size_t)i, eb,
		                 ciph ? ciph + 208 * (size_t)i : NULL, m, &c0, &c1);
		if (conv) {
			conv[2 * i] = c0;
			conv[2 * i + 1] = c1;
		}
	}
}

/* NT3 FACCH windows in groups of 4 -> FACCH3 L2 (10 bytes), status bits (32), CRC; sync id of every burst */
void oh_facch3(const float complex *iq, int n_groups, int wl, int sps, uint8_t *l2, uint8_t *bits_s, int32_t *crc,
               int32_t *sync_id)
{
	for (int g = 0; g < n_groups; g++) {
		sbit_t eb[416];
		int cr = 0;
		for (int k = 0; k < 4; k++) {
			struct osmo_cxvec cv;
			int sid = -1;
			view(&cv, iq + ((size_t)g * 4 + k) * wl, wl);
			if (gmr1_pi4cxpsk_demod(&gmr1_nt3_facch_burst, &cv, sps, 0.0f, eb + 104 * k, &sid, NULL, NULL))
				memset(eb + 104 * k, 0, 104);
			if (sync_id) sync_id[4 * g + k] = sid;
		}
		crc[g] = gmr1_facch3_decode(l2 + 10 * (size_t)g, bits_s + 32 * (size_t)g, eb, NULL, &cr);
	}
}

/* NT9 windows carrying FACCH9 -> L2 (38 bytes), CRC, sync id */
void oh_facch9(const float complex *iq, int n, int wl, int sps, uint8_t *l2, int32_t *crc, int32_t *sync_id)
{
	for (int i = 0; i < n; i++) {
		struct osmo_cxvec cv;
		sbit_t eb[662], sa[10], st[4];
		int sid = -1, cr = 0;
		view(&cv, iq + (size_t)i * wl, wl);
		if (gmr1_pi4cxpsk_demod(&gmr1_nt9_burst, &cv, sps, 0.0f, eb, &sid, NULL, NULL))
			memset(eb, 0, sizeof(eb));
		crc[i] = gmr1_facch9_decode(l2 + 38 * (size_t)i, sa, st, eb, NULL, &cr);
		if (sync_id) sync_id[i] = sid;
	}
}

/* NT9 windows carrying TCH9: n_chan channels x n_burst consecutive bursts (channel-major), one depth-3 interleaver
 * per channel as rx_tch9_init / rx_tch9 keep it -> L2 blocks of l2_bytes (60 / 30 / 18 for mode 2 / 1 / 0) */
void oh_tch9(const float complex *iq, int n_chan, int n_burst, int wl, int sps, int mode, uint8_t *l2, int32_t *conv)
{
	const int l2_bytes = mode == 2 ? 60 : (mode == 1 ? 30 : 18);
	for (int c = 0; c < n_chan; c++) {
		struct gmr1_interleaver_h il;
		gmr1_interleaver_init(&il, 3, 648);
		for (int k = 0; k < n_burst; k++) {
			const size_t i = (size_t)c * n_burst + k;
			struct osmo_cxvec cv;
			sbit_t eb[662], sa[10], st[4];
			int sid = -1, cr = 0;
			view(&cv, iq + i * wl, wl);
			if (gmr1_pi4cxpsk_demod(&gmr1_nt9_burst, &cv, sps, 0.0f, eb, &sid, NULL, NULL))
				memset(eb, 0, sizeof(eb));
			gmr1_tch9_decode(l2 + l2_bytes * i, sa, st, eb, mode, NULL, &il, &cr);
			if (conv) conv[i] = cr;
		}
		gmr1_interleaver_fini(&il);
	}
}

/* RACH windows -> 18 bytes, CRC result */
void oh_rach(const float complex *iq, int n, int wl, int sps, const uint8_t *sb_mask, uint8_t *rach, int32_t *crc)
{
	for (int i = 0; i < n; i++) {
		struct osmo_cxvec cv;
		sbit_t eb[494];
		int sid = -1, cr = 0, c2[2] = {0, 0};
		view(&cv, iq + (size_t)i * wl, wl);
		if (gmr1_pi4cxpsk_demod(&gmr1_rach_burst, &cv, sps, 0.0f, eb, &sid, NULL, NULL))
			memset(eb, 0, sizeof(eb));
		crc[i] = gmr1_rach_decode(rach + 18 * (size_t)i, eb, sb_mask ? sb_mask[i] : 0, &cr, c2);
	}
}

/* fcch_single_init (gmr1_rx.c:606-639): rough TOA over the window, fine TOA + frequency error on the burst found.
 * align[i] = rough + fine (samples), ferr[i] rad/symbol */
void oh_fcch_acquire(const float complex *iq, int n, int wl, int sps, float freq_shift, int32_t *rough, int32_t *align,
                     float *ferr)
{
	const int bl = 117 * sps;
	for (int i = 0; i < n; i++) {
		struct osmo_cxvec cv;
		int toa = 0, ftoa = 0;
		float fe = 0.0f;
		view(&cv, iq + (size_t)i * wl, wl);
		gmr1_fcch_rough(&gmr1_fcch_burst, &cv, sps, freq_shift, &toa);
		int a = toa < 0 ? 0 : (toa > wl - bl ? wl - bl : toa);
		view(&cv, iq + (size_t)i * wl + a, bl);
		gmr1_fcch_fine(&gmr1_fcch_burst, &cv, sps, freq_shift, &ftoa, &fe);
		if (rough) rough[i] = toa;
		align[i] = toa + ftoa;
		ferr[i] = fe;
	}
}

/* the "+-frequency-offset FCCH search" of BASELINE config 4: gmr1_fcch_rough once per frequency shift of a grid */
void oh_fcch_grid(const float complex *iq, int n, int wl, int sps, const float *shifts, int k, int32_t *toa)
{
	for (int i = 0; i < n; i++)
		for (int j = 0; j < k; j++) {
			struct osmo_cxvec cv;
			int t = 0;
			view(&cv, iq + (size_t)i * wl, wl);
			gmr1_fcch_rough(&gmr1_fcch_burst, &cv, sps, shifts[j], &t);
			toa[(size_t)j * n + i] = t;
		}
}
