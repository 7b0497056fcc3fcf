/* oracle shim: libosmocore <osmocom/core/bits.h> subset.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Written from the published
 * libosmocore API as used by osmo-gmr; libosmocore itself
 * is not vendored in the reference (configure.ac:23 pins only ">= 0.4.1"). */
#ifndef SHIM_OSMO_CORE_BITS_H
#define SHIM_OSMO_CORE_BITS_H
#include <stdint.h>

typedef int8_t  sbit_t;	/* soft bit: +127 strong 0, -127 strong 1, 0 erased */
typedef uint8_t ubit_t;	/* unpacked hard bit */
typedef uint8_t pbit_t;	/* packed bits */

int osmo_pbit2ubit(ubit_t *out, const pbit_t *in, unsigned int num_bits);
int osmo_ubit2pbit(pbit_t *out, const ubit_t *in, unsigned int num_bits);
int osmo_pbit2ubit_ext(ubit_t *out, unsigned int out_ofs,
                       const pbit_t *in, unsigned int in_ofs,
                       unsigned int num_bits, int lsb_mode);
int osmo_ubit2pbit_ext(pbit_t *out, unsigned int out_ofs,
                       const ubit_t *in, unsigned int in_ofs,
                       unsigned int num_bits, int lsb_mode);
#endif
