/* oracle shim: libosmo-dsp <osmocom/dsp/cxvec.h>, SURVEY.md A.2. TEST INFRASTRUCTURE ONLY.
 * libosmo-dsp is an un-vendored, unversioned dependency of the reference (configure.ac:24). */
#ifndef SHIM_OSMO_DSP_CXVEC_H
#define SHIM_OSMO_DSP_CXVEC_H
#include <complex.h>

#define CXVEC_FLG_REAL_ONLY	(1<<0)

struct osmo_cxvec {
	int len;
	int max_len;
	int flags;
	float complex *data;
	float complex _data[0];
};

void osmo_cxvec_init_from_data(struct osmo_cxvec *cv, float complex *data, int len);
struct osmo_cxvec *osmo_cxvec_alloc_from_data(float complex *data, int len);
struct osmo_cxvec *osmo_cxvec_alloc(int max_len);
void osmo_cxvec_free(struct osmo_cxvec *cv);
void osmo_cxvec_dbg_dump(struct osmo_cxvec *cv, const char *fname);
#endif
