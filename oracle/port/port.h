/* oracle/port/port.h - types shared by the oracle's plain-C restatement (TEST INFRASTRUCTURE).
 * Layouts follow the reference's public headers (include/osmocom/gmr1/sdr/pi4cxpsk.h:44-98,
 * sdr/fcch.h:36-40, l1/interleave.h:43-49) so the same ctypes wrapper drives oracle/_ref and this. */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <complex.h>
#include <stdint.h>
#include <osmocom/core/bits.h>
#include <osmocom/dsp/cxvec.h>

struct gmr1_interleaver { int N, K, n; uint8_t *bits_cpp; };

struct gmr1_pi4cxpsk_symbol { short idx; ubit_t data[2]; float mod_phase; float complex mod_val; };
struct gmr1_pi4cxpsk_modulation { float rotation; int nbits; struct gmr1_pi4cxpsk_symbol *syms, *bits; };
struct gmr1_pi4cxpsk_sync { int pos, len; uint8_t syms[32]; struct osmo_cxvec *_ref; };
struct gmr1_pi4cxpsk_data { int pos, len; };
struct gmr1_pi4cxpsk_burst {
	struct gmr1_pi4cxpsk_modulation *mod;
	int guard_pre, guard_post, len, ebits;
	struct gmr1_pi4cxpsk_sync *sync[4];
	struct gmr1_pi4cxpsk_data *data;
};
struct gmr1_fcch_burst { float freq; int len; };

void gmr1_scramble_sbit(sbit_t *out, const sbit_t *in, int len);
void gmr1_scramble_ubit(ubit_t *out, const ubit_t *in, int len);
#endif
