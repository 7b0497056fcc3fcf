/* oracle shim: libosmocore generic bit-wise CRC (crc8gen / crc16gen), SURVEY.md A.3.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_OSMO_CORE_CRCGEN_H
#define SHIM_OSMO_CORE_CRCGEN_H
#include <stdint.h>
#include <osmocom/core/bits.h>

struct osmo_crc8gen_code  { int bits; uint8_t  poly, init, remainder; };
struct osmo_crc16gen_code { int bits; uint16_t poly, init, remainder; };

uint8_t osmo_crc8gen_compute_bits(const struct osmo_crc8gen_code *c, const ubit_t *in, int len);
int  osmo_crc8gen_check_bits(const struct osmo_crc8gen_code *c, const ubit_t *in, int len, const ubit_t *crc_bits);
void osmo_crc8gen_set_bits(const struct osmo_crc8gen_code *c, const ubit_t *in, int len, ubit_t *crc_bits);

uint16_t osmo_crc16gen_compute_bits(const struct osmo_crc16gen_code *c, const ubit_t *in, int len);
int  osmo_crc16gen_check_bits(const struct osmo_crc16gen_code *c, const ubit_t *in, int len, const ubit_t *crc_bits);
void osmo_crc16gen_set_bits(const struct osmo_crc16gen_code *c, const ubit_t *in, int len, ubit_t *crc_bits);
#endif
