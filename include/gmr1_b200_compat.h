/* gmr1_b200_compat.h - the reference's own C API, exported by libgmr1_b200.so.
 *
 * libgmr1_b200.so defines the symbols of osmo-gmr's libgmr1-sdr / libgmr1-l1 with the same names,
 * signatures, struct layouts and error conventions, as n = 1 wrappers over the batched CUDA entry
 * points of gmr1_b200.h, so that the reference's src/gmr1_rx.c (and gmr1_rach_gen.c) link against
 * it UNCHANGED.  A caller that already has the reference's headers (include/osmocom/gmr1/...) keeps
 * using those and only swaps the libraries on the link line; this header is for callers that do
 * not have them.  Layouts below are ABI: gmr1_rx.c takes the address of the burst descriptors and
 * reads .len (src/gmr1_rx.c:162), and builds struct osmo_cxvec views itself (:144,167).
 *
 * Each declaration cites the reference header it mirrors.  Calls are synchronous (copy in, one
 * kernel, copy out) and therefore latency-bound: use them for bring-up and for linking existing
 * code, use the *_batch entry points for throughput.
 */
#ifndef GMR1_B200_COMPAT_H
#define GMR1_B200_COMPAT_H

#include <stdint.h>
#include <complex.h>

#ifdef __cplusplus
#error "C header (uses C99 complex); C++ callers use gmr1_b200.h"
#endif

/* ---- foreign types at the boundary (libosmocore / libosmo-dsp) ---- */
typedef int8_t  sbit_t;
typedef uint8_t ubit_t;
typedef uint8_t pbit_t;

struct osmo_cxvec {                     /* osmocom/dsp/cxvec.h */
	int len;
	int max_len;
	int flags;
	float complex *data;
	float complex _data[0];
};

/* ---- sdr/pi4cxpsk.h:37-117 ---- */
#define GMR1_MAX_SYM_EBITS	2
#define GMR1_MAX_SYNC		4
#define GMR1_MAX_SYNC_SYMS	32

struct gmr1_pi4cxpsk_symbol {
	short  idx;
	ubit_t data[GMR1_MAX_SYM_EBITS];
	float  mod_phase;
	float complex mod_val;
};

struct gmr1_pi4cxpsk_modulation {
	float rotation;
	int nbits;
	struct gmr1_pi4cxpsk_symbol *syms;
	struct gmr1_pi4cxpsk_symbol *bits;
};

struct gmr1_pi4cxpsk_sync {
	int pos;
	int len;
	uint8_t syms[GMR1_MAX_SYNC_SYMS];
	struct osmo_cxvec *_ref;
};

struct gmr1_pi4cxpsk_data {
	int pos;
	int len;
};

struct gmr1_pi4cxpsk_burst {
	struct gmr1_pi4cxpsk_modulation *mod;
	int guard_pre;
	int guard_post;
	int len;
	int ebits;
	struct gmr1_pi4cxpsk_sync *sync[GMR1_MAX_SYNC];
	struct gmr1_pi4cxpsk_data *data;
};

extern struct gmr1_pi4cxpsk_modulation gmr1_pi2cbpsk, gmr1_pi4cbpsk, gmr1_pi4cqpsk;

int gmr1_pi4cxpsk_demod(struct gmr1_pi4cxpsk_burst *burst_type, struct osmo_cxvec *burst_in, int sps,
                        float freq_shift, sbit_t *ebits, int *sync_id_p, float *toa_p, float *freq_err_p);
int gmr1_pi4cxpsk_detect(struct gmr1_pi4cxpsk_burst **burst_types, float e_toa, struct osmo_cxvec *burst_in,
                         int sps, float freq_shift, int *bt_id_p, int *sync_id_p, float *toa_p);
int gmr1_pi4cxpsk_mod_order(struct osmo_cxvec *burst_in, int sps, float freq_shift);
int gmr1_pi4cxpsk_mod(struct gmr1_pi4cxpsk_burst *burst_type, ubit_t *ebits, int sync_id,
                      struct osmo_cxvec *burst_out);

/* ---- sdr/nb.h:37-46 ---- */
extern struct gmr1_pi4cxpsk_burst gmr1_bcch_burst, gmr1_dc2_burst, gmr1_dc6_burst, gmr1_dc12_burst,
	gmr1_nt3_speech_burst, gmr1_nt3_facch_burst, gmr1_nt6_burst, gmr1_nt9_burst, gmr1_rach_burst,
	gmr1_sdcch_burst;

/* ---- sdr/fcch.h:36-61 ---- */
struct gmr1_fcch_burst {
	float freq;
	int len;
};
extern const struct gmr1_fcch_burst gmr1_fcch_burst, gmr1_fcch3_lband_burst, gmr1_fcch3_sband_burst;

int gmr1_fcch_rough(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *search_win_in, int sps,
                    float freq_shift, int *toa);
int gmr1_fcch_rough_multi(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *search_win_in, int sps,
                          float freq_shift, int *toa, int N);
int gmr1_fcch_fine(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *burst_in, int sps,
                   float freq_shift, int *toa, float *freq_error);
int gmr1_fcch_snr(const struct gmr1_fcch_burst *burst_type, struct osmo_cxvec *burst_in, int sps,
                  float freq_shift, float *snr);

/* ---- sdr/dkab.h:38-42 ---- */
#define GMR1_DKAB_SYMS (39*3)
int gmr1_dkab_demod(struct osmo_cxvec *burst_in, int sps, float freq_shift, int p, sbit_t *ebits, float *toa_p);

/* ---- l1/{bcch,ccch,facch3,facch9,tch3,tch9,rach,xch_dc12}.h ---- */
void gmr1_bcch_encode(ubit_t *bits_e, const uint8_t *l2);
int  gmr1_bcch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv);
void gmr1_ccch_encode(ubit_t *bits_e, const uint8_t *l2);
int  gmr1_ccch_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv);
void gmr1_facch3_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *bits_s, const ubit_t *ciph);
int  gmr1_facch3_decode(uint8_t *l2, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph, int *conv_rv);
void gmr1_facch9_encode(ubit_t *bits_e, const uint8_t *l2, const ubit_t *bits_sacch, const ubit_t *bits_status,
                        const ubit_t *ciph);
int  gmr1_facch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e,
                        const ubit_t *ciph, int *conv_rv);
void gmr1_tch3_encode(ubit_t *bits_e, const uint8_t *frame0, const uint8_t *frame1, const ubit_t *bits_s,
                      const ubit_t *ciph, int m);
void gmr1_tch3_decode(uint8_t *frame0, uint8_t *frame1, ubit_t *bits_s, const sbit_t *bits_e, const ubit_t *ciph,
                      int m, int *conv0_rv, int *conv1_rv);

enum gmr1_tch9_mode { GMR1_TCH9_2k4, GMR1_TCH9_4k8, GMR1_TCH9_9k6, GMR1_TCH9_MAX };
struct gmr1_interleaver {               /* l1/interleave.h:43-49 */
	int N;
	int K;
	int n;
	uint8_t *bits_cpp;
};
void gmr1_tch9_encode(ubit_t *bits_e, const uint8_t *l2, enum gmr1_tch9_mode mode, const ubit_t *bits_sacch,
                      const ubit_t *bits_status, const ubit_t *ciph, struct gmr1_interleaver *il);
void gmr1_tch9_decode(uint8_t *l2, sbit_t *bits_sacch, sbit_t *bits_status, const sbit_t *bits_e,
                      enum gmr1_tch9_mode mode, const ubit_t *ciph, struct gmr1_interleaver *il, int *conv_rv);
void gmr1_rach_encode(ubit_t *bits_e, const uint8_t *rach, uint8_t sb_mask);
int  gmr1_rach_decode(uint8_t *rach, const sbit_t *bits_e, uint8_t sb_mask, int *conv_rv, int *crc_rv);
int  gmr1_xch_dc12_encode(ubit_t *bits_e, const uint8_t *l2);   /* declared int, xch_dc12.h:37 */
int  gmr1_xch_dc12_decode(uint8_t *l2, const sbit_t *bits_e, int *conv_rv);

/* ---- l1/interleave.h, l1/scramb.h, l1/a5.h ---- */
void gmr1_interleave_intra(void *out, const void *in, int N);
void gmr1_deinterleave_intra(void *out, const void *in, int N);
int  gmr1_interleaver_init(struct gmr1_interleaver *il, int N, int K);
void gmr1_interleaver_fini(struct gmr1_interleaver *il);
void gmr1_interleave_inter(struct gmr1_interleaver *il, void *bits_epp, void *bits_ep);
void gmr1_deinterleave_inter(struct gmr1_interleaver *il, void *bits_ep, void *bits_epp);
void gmr1_scramble_sbit(sbit_t *out, const sbit_t *in, int len);
void gmr1_scramble_ubit(ubit_t *out, const ubit_t *in, int len);
void gmr1_a5(int n, uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul);
void gmr1_a5_1(uint8_t *key, uint32_t fn, int nbits, ubit_t *dl, ubit_t *ul);

/* ---- l1/conv.h:36-44, l1/crc.h:36-38, l1/punct.h:38-106: the code tables as data symbols ----
 * (libosmocore types as the reference's headers see them: osmocom/core/conv.h, osmocom/core/crcgen.h) */
enum osmo_conv_term { CONV_TERM_FLUSH = 0, CONV_TERM_TRUNCATION, CONV_TERM_TAIL_BITING };
struct osmo_conv_code {
	int N;
	int K;
	int len;
	enum osmo_conv_term term;
	const uint8_t (*next_output)[2];
	const uint8_t (*next_state)[2];
	const uint8_t *next_term_output;
	const uint8_t *next_term_state;
	const int *puncture;
};
struct osmo_crc8gen_code  { int bits; uint8_t  poly, init, remainder; };
struct osmo_crc16gen_code { int bits; uint16_t poly, init, remainder; };

extern const struct osmo_conv_code gmr1_conv_k5_12, gmr1_conv_k5_13, gmr1_conv_k5_14, gmr1_conv_k5_15,
                                   gmr1_conv_k6_14, gmr1_conv_k9_12, gmr1_conv_k9_13, gmr1_conv_k9_14,
                                   gmr1_conv_tch3;
extern const struct osmo_crc8gen_code  gmr1_crc8;
extern const struct osmo_crc16gen_code gmr1_crc12, gmr1_crc16;

struct gmr1_puncturer {                 /* l1/punct.h:38-43 */
	int r;                              /* punctured bits */
	int L;                              /* mask length (input bits) */
	int N;                              /* code rate 1/N */
	const uint8_t mask[];               /* L*N entries, 0 = punctured */
};
/* fills code->puncture (malloc'ed, -1 terminated, owned by the caller); -EINVAL on a rate mismatch */
int gmr1_puncturer_generate(struct osmo_conv_code *code, const struct gmr1_puncturer *punct_pre,
                            const struct gmr1_puncturer *punct_main, const struct gmr1_puncturer *punct_post,
                            int repeat);
/* the 51 masks gmr1_punct_<code>_<name> of l1/punct.h:56-106 */
#define PUNCT(name, r, L, N, ...) extern const struct gmr1_puncturer gmr1_punct_##name;
#include "gmr1_punct_masks.inc"
#undef PUNCT

#endif /* GMR1_B200_COMPAT_H */
