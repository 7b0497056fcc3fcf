#ifndef SHIM_OSMO_CORE_GSMTAP_UTIL_H
#define SHIM_OSMO_CORE_GSMTAP_UTIL_H
#include <stdint.h>
struct msgb;
struct gsmtap_inst;
struct gsmtap_inst *gsmtap_source_init(const char *host, uint16_t port, int ofd_wq_mode);
int gsmtap_source_add_sink(struct gsmtap_inst *gti);
int gsmtap_sendmsg(struct gsmtap_inst *gti, struct msgb *msg);
#endif
