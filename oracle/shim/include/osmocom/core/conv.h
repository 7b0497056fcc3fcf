/* oracle shim: libosmocore <osmocom/core/conv.h> (generic soft Viterbi), SURVEY.md A.1.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_OSMO_CORE_CONV_H
#define SHIM_OSMO_CORE_CONV_H
#include <stdint.h>
#include <osmocom/core/bits.h>

enum osmo_conv_term {
	CONV_TERM_FLUSH = 0,
	CONV_TERM_TRUNCATION,
	CONV_TERM_TAIL_BITING,
};

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

int osmo_conv_get_input_length(const struct osmo_conv_code *code, int len);
int osmo_conv_get_output_length(const struct osmo_conv_code *code, int len);
int osmo_conv_encode(const struct osmo_conv_code *code, const ubit_t *input, ubit_t *output);
int osmo_conv_decode(const struct osmo_conv_code *code, const sbit_t *input, ubit_t *output);
#endif
