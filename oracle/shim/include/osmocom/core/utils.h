/* oracle shim: libosmocore <osmocom/core/utils.h> subset used by gmr1_rx.c. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_OSMO_CORE_UTILS_H
#define SHIM_OSMO_CORE_UTILS_H
#include <stdint.h>
int osmo_hexparse(const char *str, uint8_t *b, int max_len);
char *osmo_hexdump_nospc(const unsigned char *buf, int len);
#endif
