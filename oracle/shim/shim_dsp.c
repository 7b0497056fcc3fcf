/* oracle shim: libosmo-dsp (cxvec / cxvec_math / cfile) and the FFTW3f calls of fcch.c.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * libosmo-dsp (configure.ac:24, unversioned) and FFTW3f (configure.ac:25) are third-party
 * dependencies that are NOT vendored under /root/reference.  This file restates their
 * published semantics as SURVEY.md Appendix A.2 / A.4 records them, constrained by how the
 * reference calls them (osmo_cxvec_sig_normalize: fcch.c:230,366,537,662, pi4cxpsk.c:539,
 * 629,702, dkab.c:195; osmo_cxvec_correlate: fcch.c:233,369, pi4cxpsk.c:229;
 * osmo_cxvec_peak_energy_find: fcch.c:238,596,597, pi4cxpsk.c:240; fftwf_*: fcch.c:583-589,
 * 684-686).  The reference has no tests at this boundary: PARITY IS UNPINNED, and this file
 * is the normative definition for the repo.  Edge cases the upstream source would decide
 * are frozen here and listed in oracle/README.md.
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

#include <osmocom/dsp/cxvec.h>
#include <osmocom/dsp/cxvec_math.h>
#include <osmocom/dsp/cfile.h>
#include <fftw3.h>

/* ---------------------------------------------------------------- cxvec.h */

void osmo_cxvec_init_from_data(struct osmo_cxvec *cv, float complex *data, int len)
{
	cv->len = cv->max_len = len;
	cv->flags = 0;
	cv->data = data;
}

struct osmo_cxvec *osmo_cxvec_alloc_from_data(float complex *data, int len)
{
	struct osmo_cxvec *cv = malloc(sizeof(*cv));
	if (cv)
		osmo_cxvec_init_from_data(cv, data, len);
	return cv;
}

struct osmo_cxvec *osmo_cxvec_alloc(int max_len)
{
	struct osmo_cxvec *cv = malloc(sizeof(*cv) + max_len * sizeof(float complex));
	if (!cv)
		return NULL;
	cv->len = 0;
	cv->max_len = max_len;
	cv->flags = 0;
	cv->data = &cv->_data[0];
	return cv;
}

void osmo_cxvec_free(struct osmo_cxvec *cv) { free(cv); }

void osmo_cxvec_dbg_dump(struct osmo_cxvec *cv, const char *fname)
{
	FILE *f = fopen(fname, "wb");
	if (!f)
		return;
	fwrite(cv->data, sizeof(float complex), cv->len, f);
	fclose(f);
}

/* ---------------------------------------------------------------- cxvec_math.h */

static struct osmo_cxvec *out_prep(struct osmo_cxvec *out, int l)
{
	if (!out)
		out = osmo_cxvec_alloc(l);
	else if (out->max_len < l)
		return NULL;
	return out;
}

struct osmo_cxvec *osmo_cxvec_scale(const struct osmo_cxvec *in, float complex scale, struct osmo_cxvec *out)
{
	int i, seq_real = !!(in->flags & CXVEC_FLG_REAL_ONLY);
	int scale_real = (cimagf(scale) == 0.0f);

	out = out_prep(out, in->len);
	if (!out)
		return NULL;

	if (scale_real) {
		float s = crealf(scale);
		for (i = 0; i < in->len; i++)
			out->data[i] = in->data[i] * s;
	} else {
		for (i = 0; i < in->len; i++)
			out->data[i] = in->data[i] * scale;
	}
	out->len = in->len;
	out->flags = in->flags;
	if (!(seq_real && scale_real))
		out->flags &= ~CXVEC_FLG_REAL_ONLY;
	return out;
}

struct osmo_cxvec *osmo_cxvec_rotate(const struct osmo_cxvec *in, float rps, struct osmo_cxvec *out)
{
	int i;

	out = out_prep(out, in->len);
	if (!out)
		return NULL;

	for (i = 0; i < in->len; i++)
		out->data[i] = in->data[i] * cexpf(I * (rps * (float)i));

	out->len = in->len;
	out->flags = in->flags & ~CXVEC_FLG_REAL_ONLY;
	return out;
}

/* CONV_NO_DELAY: output has the length of g and is aligned with g (the filter centre tap
 * maps onto the same index); CONV_FULL_SPAN / CONV_OVERLAP_ONLY are the usual 'full' and
 * 'valid' spans.  Samples outside g count as zero. */
struct osmo_cxvec *osmo_cxvec_convolve(const struct osmo_cxvec *f, const struct osmo_cxvec *g,
                                       enum osmo_cxvec_conv_type type, struct osmo_cxvec *out)
{
	int Lf = f->len, Lg = g->len, Lo, si, i, j;
	int f_real = !!(f->flags & CXVEC_FLG_REAL_ONLY);

	switch (type) {
	case CONV_FULL_SPAN:    Lo = Lf + Lg - 1; si = 0; break;
	case CONV_OVERLAP_ONLY: Lo = abs(Lf - Lg) + 1; si = (Lf < Lg ? Lf : Lg) - 1; break;
	case CONV_NO_DELAY:     Lo = Lg; si = (Lf >> 1) - ((Lf & 1) ^ 1); break;
	default: return NULL;
	}

	out = out_prep(out, Lo);
	if (!out)
		return NULL;

	for (i = 0; i < Lo; i++) {
		float complex acc = 0.0f;
		int n = i + si;
		for (j = 0; j < Lf; j++) {
			int k = n - j;
			if (k < 0 || k >= Lg)
				continue;
			if (f_real)
				acc += crealf(f->data[j]) * g->data[k];
			else
				acc += f->data[j] * g->data[k];
		}
		out->data[i] = acc;
	}
	out->len = Lo;
	out->flags = (f->flags & g->flags) & CXVEC_FLG_REAL_ONLY;
	return out;
}

/* out[m] = sum_n conj(f[n]) * g[m + n*step],  m in [0, g->len - f->len*step] */
struct osmo_cxvec *osmo_cxvec_correlate(const struct osmo_cxvec *f, const struct osmo_cxvec *g,
                                        int g_corr_step, struct osmo_cxvec *out)
{
	int l = g->len - f->len * g_corr_step + 1, m, n;
	int f_real = !!(f->flags & CXVEC_FLG_REAL_ONLY);

	if (l < 0)
		l = 0;
	out = out_prep(out, l);
	if (!out)
		return NULL;

	for (m = 0; m < l; m++) {
		float complex acc = 0.0f;
		const float complex *gp = &g->data[m];
		if (f_real) {
			for (n = 0; n < f->len; n++)
				acc += crealf(f->data[n]) * gp[n * g_corr_step];
		} else {
			for (n = 0; n < f->len; n++)
				acc += conjf(f->data[n]) * gp[n * g_corr_step];
		}
		out->data[m] = acc;
	}
	out->len = l;
	out->flags = 0;
	return out;
}

/* windowed-sinc interpolation, 10 taps either side, window clipped to the vector */
float complex osmo_cxvec_interpolate_point(const struct osmo_cxvec *cv, float pos)
{
	const int N = 10;
	int b, e, i;
	float complex val = 0.0f;

	b = (int)floorf(pos) - N;
	e = b + 2 * N + 1;
	if (b < 0)
		b = 0;
	if (e > cv->len)
		e = cv->len;

	for (i = b; i < e; i++)
		val += cv->data[i] * osmo_sinc(M_PIf * ((float)i - pos));

	return val;
}

float osmo_cxvec_peak_energy_find(const struct osmo_cxvec *cv, int win_size,
                                  enum osmo_cxvec_peak_alg alg, float complex *peak_val_p)
{
	float val, max_val, peak_pos = 0.0f;
	int idx, max_idx, hi;

	if (win_size > cv->len)
		win_size = cv->len;
	if (win_size <= 0)
		return 0.0f;

	/* Energy of the win_size samples ending at idx (samples before the start count as 0, i.e.
	 * the window slides in from the left); a new maximum is taken only on a strictly greater
	 * sum, so the first of equal windows wins.
	 * Frozen choice: upstream keeps a running sum (subtract the oldest, add the newest); here
	 * every window is summed afresh, oldest sample first.  The two are the same function up to
	 * float rounding drift of the running sum; the fresh sum has no history, which is what lets
	 * a GPU evaluate all windows in parallel and still agree bit for bit. */
	max_val = 0.0f;
	max_idx = 0;
	for (idx = 0; idx < cv->len; idx++) {
		val = 0.0f;
		for (hi = idx - win_size + 1; hi <= idx; hi++)
			if (hi >= 0)
				val += osmo_normsqf(cv->data[hi]);
		if (val > max_val) {
			max_val = val;
			max_idx = idx - win_size + 1;
		}
	}
	if (max_idx < 0)	/* frozen edge case: best window hangs over the start */
		max_idx = 0;

	if (alg == PEAK_WEIGH_WIN || alg == PEAK_WEIGH_WIN_CENTER) {
		float mw = 0.0f, sw = 0.0f;
		for (idx = max_idx; idx < max_idx + win_size; idx++) {
			float e = osmo_normsqf(cv->data[idx]);
			sw += e;
			mw += e * (float)idx;
		}
		peak_pos = (sw > 0.0f) ? (mw / sw) : (float)max_idx;
	} else {	/* PEAK_EARLY_LATE */
		float early_idx, late_idx, incr;
		float early, late, mv = -1.0f;
		int mwi = max_idx;

		for (idx = max_idx; idx < max_idx + win_size; idx++) {
			float e = osmo_normsqf(cv->data[idx]);
			if (e > mv) {
				mv = e;
				mwi = idx;
			}
		}

		early_idx = (float)(mwi - 1);
		late_idx  = (float)(mwi + 1);
		incr = 0.5f;

		while (incr > (1.0f / 1024.0f)) {
			early = osmo_normsqf(osmo_cxvec_interpolate_point(cv, early_idx));
			late  = osmo_normsqf(osmo_cxvec_interpolate_point(cv, late_idx));
			if (early < late)
				early_idx += incr;
			else if (early > late)
				early_idx -= incr;
			else
				break;
			incr /= 2.0f;
			late_idx = early_idx + 2.0f;
		}
		peak_pos = early_idx + 1.0f;
	}

	if (peak_val_p)
		*peak_val_p = osmo_cxvec_interpolate_point(cv, peak_pos);

	return peak_pos;
}

/* indices of the N largest |cv|^2, descending (ties: lower index first) */
int osmo_cxvec_peaks_scan(const struct osmo_cxvec *cv, int *peaks_idx, int N)
{
	float val[N];
	int i, j, k, n = 0;

	for (i = 0; i < N; i++) {
		peaks_idx[i] = 0;
		val[i] = -1.0f;
	}
	for (i = 0; i < cv->len; i++) {
		float e = osmo_normsqf(cv->data[i]);
		for (j = 0; j < n; j++)
			if (e > val[j])
				break;
		if (j == N)
			continue;
		for (k = (n < N ? n : N - 1); k > j; k--) {
			val[k] = val[k - 1];
			peaks_idx[k] = peaks_idx[k - 1];
		}
		val[j] = e;
		peaks_idx[j] = i;
		if (n < N)
			n++;
	}
	return n;
}

struct osmo_cxvec *osmo_cxvec_sig_normalize(const struct osmo_cxvec *sig, int decim, float freq_shift,
                                            struct osmo_cxvec *out)
{
	float complex avg = 0.0f;
	float sigma = 0.0f, stddev;
	int l, i, j;

	l = sig->len / decim;
	out = out_prep(out, l);
	if (!out)
		return NULL;

	for (i = 0; i < sig->len; i++)
		avg += sig->data[i];
	avg /= (float)sig->len;

	for (i = 0; i < sig->len; i++)
		sigma += osmo_normsqf(sig->data[i] - avg);
	sigma /= (float)sig->len;

	stddev = sqrtf(sigma);
	if (stddev == 0.0f)
		stddev = 1.0f;

	for (i = 0, j = 0; i < l; i++, j += decim)
		out->data[i] = (sig->data[j] - avg) / stddev;

	out->len = l;
	out->flags = 0;

	if (freq_shift != 0.0f)
		for (i = 0; i < l; i++)
			out->data[i] *= cexpf(I * (freq_shift * (float)i));

	return out;
}

/* ---------------------------------------------------------------- cfile.h */

struct cfile *cfile_load(const char *filename)
{
	struct cfile *cf;
	struct stat st;
	int fd;

	cf = calloc(1, sizeof(*cf));
	if (!cf)
		return NULL;
	fd = open(filename, O_RDONLY);
	if (fd < 0 || fstat(fd, &st) < 0 || st.st_size == 0)
		goto err;
	cf->_blen = st.st_size;
	cf->len = st.st_size / sizeof(float complex);
	cf->data = mmap(NULL, cf->_blen, PROT_READ, MAP_SHARED, fd, 0);
	if (cf->data == MAP_FAILED)
		goto err;
	close(fd);
	return cf;
err:
	if (fd >= 0)
		close(fd);
	free(cf);
	return NULL;
}

void cfile_release(struct cfile *cf)
{
	if (!cf)
		return;
	munmap(cf->data, cf->_blen);
	free(cf);
}

/* ---------------------------------------------------------------- fftw3.h
 * Unnormalised DFT X[k] = sum_n x[n] e^{sign*2*pi*i*k*n/N}, in or out of place, direct O(N^2)
 * form with double accumulation (N = 117 in every call gmr1_rx makes). */

struct shim_fftwf_plan_s { int n, sign; fftwf_complex *in, *out; };

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{
	fftwf_plan p = malloc(sizeof(*p));
	(void)flags;
	p->n = n; p->sign = sign; p->in = in; p->out = out;
	return p;
}

void fftwf_execute(const fftwf_plan p)
{
	int n = p->n, k, t;
	double complex *tw = malloc(sizeof(double complex) * n);
	double complex *res = malloc(sizeof(double complex) * n);

	for (k = 0; k < n; k++)
		tw[k] = cexp(I * (double)p->sign * 2.0 * M_PI * (double)k / (double)n);
	for (k = 0; k < n; k++) {
		double complex acc = 0.0;
		for (t = 0; t < n; t++)
			acc += (double complex)p->in[t] * tw[(int)(((long)k * t) % n)];
		res[k] = acc;
	}
	for (k = 0; k < n; k++)
		p->out[k] = (float complex)res[k];
	free(res);
	free(tw);
}

void fftwf_destroy_plan(fftwf_plan p) { free(p); }
