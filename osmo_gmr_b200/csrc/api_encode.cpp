// api_encode.cpp - C ABI: host-side channel encoders (include/gmr1_b200.h, "transmit side").
#include "../../include/gmr1_b200.h"
#include "encode.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>

using namespace gmr1;

extern "C" {

int gmr1b200_xcch_encode_batch(int chan, uint8_t *bits_e, const uint8_t *l2, int n)
{
	int ch, n_in;
	switch (chan) {
	case 0: ch = CH_BCCH; n_in = 424; break;
	case 1: ch = CH_CCCH; n_in = 432; break;
	case 2: ch = CH_DC12; n_in = 432; break;
	default: return -EINVAL;
	}
	if (!bits_e || !l2 || n < 0)
		return -EINVAL;
	for (int i = 0; i < n; i++)
		encode_simple(ch, bits_e + (size_t)i * n_in, l2 + (size_t)i * 24);
	return 0;
}

int gmr1b200_facch3_encode(uint8_t *bits_e, const uint8_t *l2, const uint8_t *bits_s, const uint8_t *ciph)
{
	if (!bits_e || !l2 || !bits_s)
		return -EINVAL;
	encode_facch3(bits_e, l2, bits_s, ciph);
	return 0;
}

int gmr1b200_facch9_encode(uint8_t *bits_e, const uint8_t *l2, const uint8_t *bits_sacch,
                           const uint8_t *bits_status, const uint8_t *ciph)
{
	if (!bits_e || !l2 || !bits_sacch || !bits_status)
		return -EINVAL;
	encode_facch9(bits_e, l2, bits_sacch, bits_status, ciph);
	return 0;
}

void *gmr1b200_tch9_interleaver_new(void)
{
	Interleaver *il = (Interleaver *)malloc(sizeof(Interleaver));
	if (il)
		interleaver_init(il);
	return il;
}

void gmr1b200_tch9_interleaver_free(void *il) { free(il); }

int gmr1b200_tch9_encode(uint8_t *bits_e, const uint8_t *l2, int mode, const uint8_t *bits_sacch,
                         const uint8_t *bits_status, const uint8_t *ciph, void *interleaver)
{
	if (!bits_e || !l2 || !bits_sacch || !bits_status || !interleaver || mode < 0 || mode > 2)
		return -EINVAL;
	encode_tch9(bits_e, l2, mode, bits_sacch, bits_status, ciph, (Interleaver *)interleaver);
	return 0;
}

int gmr1b200_tch9_encode_ep(uint8_t *ep, const uint8_t *l2, int mode)
{
	if (!ep || !l2 || mode < 0 || mode > 2)
		return -EINVAL;
	encode_tch9_ep(ep, l2, mode);
	return 0;
}

int gmr1b200_rach_encode(uint8_t *bits_e, const uint8_t *rach, int sb_mask)
{
	if (!bits_e || !rach)
		return -EINVAL;
	encode_rach(bits_e, rach, (uint8_t)sb_mask);
	return 0;
}

int gmr1b200_tch3_encode(uint8_t *bits_e, const uint8_t *frame0, const uint8_t *frame1, const uint8_t *bits_s,
                         const uint8_t *ciph, int m)
{
	if (!bits_e || !frame0 || !frame1 || !bits_s)
		return -EINVAL;
	encode_tch3(bits_e, frame0, frame1, bits_s, ciph, m);
	return 0;
}

}  // extern "C"
