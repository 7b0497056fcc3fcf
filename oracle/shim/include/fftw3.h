/* oracle shim: the three FFTW3f entry points fcch.c uses. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_FFTW3_H
#define SHIM_FFTW3_H
#include <complex.h>
typedef float complex fftwf_complex;
typedef struct shim_fftwf_plan_s *fftwf_plan;
#define FFTW_FORWARD	(-1)
#define FFTW_BACKWARD	(+1)
#define FFTW_ESTIMATE	(1U << 6)
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
#endif
