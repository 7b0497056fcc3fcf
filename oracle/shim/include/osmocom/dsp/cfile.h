/* oracle shim: libosmo-dsp <osmocom/dsp/cfile.h>. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_OSMO_DSP_CFILE_H
#define SHIM_OSMO_DSP_CFILE_H
#include <complex.h>
struct cfile { float complex *data; unsigned int len; unsigned int _blen; };
struct cfile *cfile_load(const char *filename);
void cfile_release(struct cfile *cf);
#endif
