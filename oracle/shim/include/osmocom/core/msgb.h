/* oracle shim: minimal msgb so that gmr1_rx.c links; GSMTAP output is out of scope. */
#ifndef SHIM_OSMO_CORE_MSGB_H
#define SHIM_OSMO_CORE_MSGB_H
#include <stdint.h>
struct msgb { uint16_t len, alloc; uint8_t data[0]; };
struct msgb *msgb_alloc(uint16_t size, const char *name);
uint8_t *msgb_put(struct msgb *m, unsigned int len);
void msgb_free(struct msgb *m);
#endif
