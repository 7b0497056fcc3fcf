/* oracle shim: libosmo-dsp <osmocom/dsp/cxvec_math.h>, SURVEY.md A.2. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_OSMO_DSP_CXVEC_MATH_H
#define SHIM_OSMO_DSP_CXVEC_MATH_H
#include <complex.h>
#include <math.h>
#include <osmocom/dsp/cxvec.h>

#define M_PIf (3.14159265358979323846264338327f)

static inline float osmo_normsqf(float complex c)
{
	return crealf(c) * crealf(c) + cimagf(c) * cimagf(c);
}

static inline float osmo_sinc(float x)
{
	if ((x >= 0.01f) || (x <= -0.01f))
		return sinf(x) / x;
	return 1.0f;
}

enum osmo_cxvec_conv_type { CONV_FULL_SPAN, CONV_OVERLAP_ONLY, CONV_NO_DELAY };
enum osmo_cxvec_peak_alg  { PEAK_WEIGH_WIN, PEAK_WEIGH_WIN_CENTER, PEAK_EARLY_LATE };

struct osmo_cxvec *osmo_cxvec_scale(const struct osmo_cxvec *in, float complex scale, struct osmo_cxvec *out);
struct osmo_cxvec *osmo_cxvec_rotate(const struct osmo_cxvec *in, float rps, struct osmo_cxvec *out);
struct osmo_cxvec *osmo_cxvec_delay(const struct osmo_cxvec *in, float delay, struct osmo_cxvec *out);
struct osmo_cxvec *osmo_cxvec_convolve(const struct osmo_cxvec *f, const struct osmo_cxvec *g,
                                       enum osmo_cxvec_conv_type type, struct osmo_cxvec *out);
struct osmo_cxvec *osmo_cxvec_correlate(const struct osmo_cxvec *f, const struct osmo_cxvec *g,
                                        int g_corr_step, struct osmo_cxvec *out);
float complex osmo_cxvec_interpolate_point(const struct osmo_cxvec *cv, float pos);
float osmo_cxvec_peak_energy_find(const struct osmo_cxvec *cv, int win_size,
                                  enum osmo_cxvec_peak_alg alg, float complex *peak_val_p);
int osmo_cxvec_peaks_scan(const struct osmo_cxvec *cv, int *peaks_idx, int N);
struct osmo_cxvec *osmo_cxvec_sig_normalize(const struct osmo_cxvec *sig, int decim, float freq_shift,
                                            struct osmo_cxvec *out);
#endif
