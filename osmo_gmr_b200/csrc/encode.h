// encode.h - host-side channel encoders (see encode.cpp)
#pragma once
#include <stdint.h>
#include "gmr1_tables.h"

namespace gmr1 {

struct Interleaver {          // TCH9 inter-burst interleaver memory (depth 3, width 648)
	int n;
	uint8_t hist[3][648];
};

void encode_simple(int ch, uint8_t *bits_e, const uint8_t *l2);      // CH_BCCH, CH_CCCH, CH_DC12
void encode_facch3(uint8_t *bits_e, const uint8_t *l2, const uint8_t *bits_s, const uint8_t *ciph);
void encode_facch9(uint8_t *bits_e, const uint8_t *l2, const uint8_t *sacch, const uint8_t *status, const uint8_t *ciph);
void interleaver_init(Interleaver *il);
void encode_tch9(uint8_t *bits_e, const uint8_t *l2, int mode, const uint8_t *sacch, const uint8_t *status,
                 const uint8_t *ciph, Interleaver *il);
void encode_tch9_ep(uint8_t *ep, const uint8_t *l2, int mode);
void encode_rach(uint8_t *bits_e, const uint8_t *rach, uint8_t sb_mask);
void encode_tch3(uint8_t *bits_e, const uint8_t *frame0, const uint8_t *frame1, const uint8_t *bits_s,
                 const uint8_t *ciph, int m);

}  // namespace gmr1
