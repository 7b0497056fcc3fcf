/* oracle/ref_glue.c - tiny accessors linked into _ref/libgmr1_ref.so so that tests can read
 * data the reference keeps in file-static objects (the specialised osmo_conv_code of each
 * channel are `static`, e.g. src/l1/bcch.c:42).  TEST INFRASTRUCTURE ONLY; contains no
 * reference code. */
#include <osmocom/core/conv.h>

/* Expand a code's puncture list into a 0/1 "kept" mask over the unpunctured coded bits. */
int oracle_puncture_mask(const struct osmo_conv_code *code, unsigned char *mask, int max)
{
	int n = code->len * code->N, i, p = 0;
	if (code->term == CONV_TERM_FLUSH)
		n += code->N * (code->K - 1);
	if (n > max)
		return -1;
	for (i = 0; i < n; i++) {
		if (code->puncture && code->puncture[p] == i) {
			mask[i] = 0;
			p++;
		} else
			mask[i] = 1;
	}
	return n;
}
