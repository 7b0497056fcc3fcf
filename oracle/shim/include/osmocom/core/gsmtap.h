/* oracle shim: GSMTAP constants used by gmr1_rx.c / gsmtap.c. Wire output is out of scope. */
#ifndef SHIM_OSMO_CORE_GSMTAP_H
#define SHIM_OSMO_CORE_GSMTAP_H
#include <stdint.h>
#define GSMTAP_VERSION		0x02
#define GSMTAP_UDP_PORT		4729
#define GSMTAP_TYPE_GMR1_UM	0x0a
#define GSMTAP_GMR1_UNKNOWN	0x00
#define GSMTAP_GMR1_BCCH	0x01
#define GSMTAP_GMR1_CCCH	0x02
#define GSMTAP_GMR1_PCH		0x03
#define GSMTAP_GMR1_AGCH	0x04
#define GSMTAP_GMR1_BACH	0x05
#define GSMTAP_GMR1_RACH	0x06
#define GSMTAP_GMR1_CBCH	0x07
#define GSMTAP_GMR1_SDCCH	0x08
#define GSMTAP_GMR1_TACCH	0x09
#define GSMTAP_GMR1_GBCH	0x0a
#define GSMTAP_GMR1_SACCH	0x01
#define GSMTAP_GMR1_FACCH	0x02
#define GSMTAP_GMR1_DKAB	0x03
#define GSMTAP_GMR1_TCH3	0x10
#define GSMTAP_GMR1_TCH6	0x14
#define GSMTAP_GMR1_TCH9	0x18
struct gsmtap_hdr {
	uint8_t version, hdr_len, type, timeslot;
	uint16_t arfcn;
	int8_t signal_dbm, snr_db;
	uint32_t frame_number;
	uint8_t sub_type, antenna_nr, sub_slot, res;
} __attribute__((packed));
#endif
