// api_sdr.cu - C ABI: FCCH acquisition, DKAB demodulation, modulation-order test (include/gmr1_b200.h)
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "launch.h"

#include <math.h>
#include <stdlib.h>
#include <vector>

using namespace gmr1;

// FCCH burst formats (reference src/sdr/fcch.c:50-70): sweep range, symbols
static const struct { float freq; int len; } FCCH_TYPES[3] = {{0.32f, 117}, {0.32f, 468}, {0.16f, 468}};

static int fcch_common(int type, FcchArgs &a, Stage &s, int64_t iq_len, const char *what)
{
	if (type < 0 || type > 2 || a.n < 0 || !a.iq || a.sps < 1 || a.sps > 16)
		return set_err(-EINVAL, what);
	a.freq = FCCH_TYPES[type].freq;
	a.len = FCCH_TYPES[type].len;
	if (a.n && !a.ofs && (a.stride < 0 || (int64_t)(a.n - 1) * a.stride + a.win_len > iq_len))
		return set_err(-EINVAL, "fcch batch: windows exceed iq_len");
	const size_t n = (size_t)a.n;
	a.iq = (const float2 *)s.in((const float *)a.iq, (size_t)iq_len * 2);
	a.ofs = s.in(a.ofs, n);
	a.freq_shift = s.in(a.freq_shift, n);
	a.toa = s.out(a.toa, n);
	a.freq_error = s.out(a.freq_error, n);
	a.snr = s.out(a.snr, n);
	a.peak = s.out(a.peak, n);
	return 0;
}

extern "C" {

int gmr1b200_fcch_rough_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                              int64_t win_stride, int win_len, int sps, const float *freq_shift, float freq_shift0,
                              int32_t *toa, float *peak, int n, void *stream)
{
	if (!toa)
		return set_err(-EINVAL, "fcch_rough_batch: toa NULL");
	FcchArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.toa = toa; a.peak = peak;
	Stage s(stream);
	int rc = fcch_common(fcch_type, a, s, iq_len, "fcch_rough_batch: bad argument");
	if (rc)
		return rc;
	if (win_len / sps < a.len)
		return set_err(-EINVAL, "fcch_rough_batch: window shorter than the FCCH burst");
	if (n == 0)
		return 0;
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_fcch_rough(a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "fcch_rough kernel");
}

// ---- multi-FCCH acquisition: GPU correlation + the reference's scalar peak bookkeeping ----------
// (src/sdr/fcch.c:264-326 _peak_record, :373-483)
static void peak_record(int burst_len, int *toa, float *pwr, int *n, int N, int Lp, int sps, int peak_toa, float peak_pwr)
{
	int has_dupe = 0;
	const int th = (burst_len * sps) >> 1;
	for (int i = 0; i < *n; i++) {
		const int d = (toa[i] % Lp) - (peak_toa % Lp);
		if (abs(d) > th)
			continue;
		if (pwr[i] > peak_pwr) {            // an equal-or-stronger twin is already recorded
			if (!has_dupe)
				has_dupe = 1;
			continue;
		}
		for (int j = i; j < *n - 1; j++) {  // drop the weaker twin
			toa[j] = toa[j + 1];
			pwr[j] = pwr[j + 1];
		}
		*n -= 1;
		has_dupe = -1;
	}
	if (has_dupe > 0)
		return;
	int i;
	for (i = 0; i < *n; i++)
		if (peak_pwr > pwr[i])
			break;
	if (i == N)
		return;
	for (int j = N - 1; j > i; j--) {
		toa[j] = toa[j - 1];
		pwr[j] = pwr[j - 1];
	}
	toa[i] = peak_toa;
	pwr[i] = peak_pwr;
	if (*n != N)
		*n += 1;
}

int gmr1b200_fcch_rough_multi(int fcch_type, const float *iq, int64_t win_len, int sps, float freq_shift,
                              int32_t *peaks_toa, int N, void *stream)
{
	if (fcch_type < 0 || fcch_type > 2 || !iq || !peaks_toa || N < 1 || sps < 1 || sps > 16)
		return set_err(-EINVAL, "fcch_rough_multi: bad argument");
	const int sym_rate = 23400;
	if (win_len < ((int64_t)650 * sym_rate * sps) / 1000)      // fcch.c:355
		return set_err(-EINVAL, "fcch_rough_multi: needs 650 ms of signal");
	const int blen = FCCH_TYPES[fcch_type].len;
	const int l = (int)(win_len / sps), nc = l - blen + 1;
	std::vector<float> pw((size_t)nc);
	{
		FcchArgs a = {};
		a.iq = (const float2 *)iq; a.stride = 0; a.n = 1; a.win_len = (int)win_len; a.sps = sps;
		a.freq_shift0 = freq_shift;
		Stage s(stream);
		int dummy_toa;
		a.toa = &dummy_toa;
		int rc = fcch_common(fcch_type, a, s, win_len, "fcch_rough_multi: bad argument");
		if (rc)
			return rc;
		a.en_out = s.out(pw.data(), (size_t)nc);
		cudaError_t e = cudaSuccess;
		if (!s.failed()) {
			e = launch_fcch_rough(a, (cudaStream_t)stream);
			if (e == cudaSuccess)
				g_launches.fetch_add(1);
		}
		rc = s.finish(e, "fcch_rough kernel");
		if (rc)
			return rc;
	}
	float *corr_pwr = pw.data();
	int Lw = (320 * sym_rate) / 1000 + blen, Lp = (320 * sym_rate) / 1000;
	int pwr_max_idx = 0;
	float pwr_max = 0.0f;
	for (int i = 0; i < nc; i++)
		if (corr_pwr[i] > pwr_max && i < Lw) {
			pwr_max = corr_pwr[i];
			pwr_max_idx = i;
		}
	// the twin one period later, +-10 symbols (fcch.c:397-430)
	float pwrs[2] = {0.0f, 0.0f}, peaks[2] = {0.0f, 0.0f};
	for (int i = -10; i <= 10; i++) {
		int j = pwr_max_idx + i;
		if (j > 0 && j < nc) {
			pwrs[0] += corr_pwr[j];
			peaks[0] += corr_pwr[j] * j;
		}
		j += Lp;
		if (j > 0 && j < nc) {
			pwrs[1] += corr_pwr[j];
			peaks[1] += corr_pwr[j] * j;
		}
	}
	peaks[0] /= pwrs[0];
	peaks[1] /= pwrs[1];
	const int nLp = (int)round(peaks[1] - peaks[0]);
	if (abs(nLp - Lp) > 10)
		return set_err(-EINVAL, "fcch_rough_multi: FCCH period mismatch");
	Lp = nLp;
	if (Lw + Lp > nc)
		Lw = nc - Lp;
	float avg = 0.0f;
	for (int i = 0; i < Lw; i++) {          // geometric mix of the two cycles (fcch.c:435-441)
		const float v = sqrtf(corr_pwr[i] * corr_pwr[i + Lp]);
		corr_pwr[i] = v;
		avg += v;
	}
	avg /= Lw;
	float stddev = 0.0f;
	for (int i = 0; i < Lw; i++) {
		const float v = corr_pwr[i] - avg;
		stddev += v * v;
	}
	stddev = sqrtf(stddev / Lw);
	const float th = avg + 3.0f * stddev;
	std::vector<float> peaks_pwr((size_t)N, 0.0f);
	int peaks_cnt = 0;
	for (int i = 1, in_peak = 0; i < Lw - 1; i++) {
		if (corr_pwr[i] > th) {
			if (in_peak)
				continue;
			in_peak = 1;
			const float p_pwr = corr_pwr[i - 1] + corr_pwr[i] + corr_pwr[i + 1];
			const float p_fpos = (-corr_pwr[i - 1] + corr_pwr[i + 1]) / p_pwr;
			const int p_pos = (int)round((i + p_fpos) * sps);
			peak_record(blen, peaks_toa, peaks_pwr.data(), &peaks_cnt, N, Lp, sps, p_pos, p_pwr);
		} else {
			in_peak = 0;
		}
	}
	return peaks_cnt;
}

static int fine_or_snr(int mode, int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                       int64_t win_stride, int sps, const float *freq_shift, float freq_shift0,
                       int32_t *toa, float *freq_error, float *snr, int n, void *stream)
{
	FcchArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.toa = toa; a.freq_error = freq_error; a.snr = snr;
	if (fcch_type >= 0 && fcch_type <= 2)
		a.win_len = FCCH_TYPES[fcch_type].len * sps;      // the reference insists on exactly this (fcch.c:546-551)
	Stage s(stream);
	int rc = fcch_common(fcch_type, a, s, iq_len, "fcch fine/snr batch: bad argument");
	if (rc)
		return rc;
	if (n == 0)
		return 0;
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_fcch_fine(a, mode, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "fcch_fine kernel");
}

int gmr1b200_fcch_fine_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                             int64_t win_stride, int sps, const float *freq_shift, float freq_shift0,
                             int32_t *toa, float *freq_error, int n, void *stream)
{
	if (!toa || !freq_error)
		return set_err(-EINVAL, "fcch_fine_batch: NULL output");
	return fine_or_snr(0, fcch_type, iq, iq_len, win_ofs, win_stride, sps, freq_shift, freq_shift0,
	                   toa, freq_error, nullptr, n, stream);
}

int gmr1b200_fcch_acquire_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                int64_t win_stride, int win_len, int sps, int32_t *rough_toa, int32_t *align,
                                float *freq_error, int n, void *stream)
{
	if (!align || !freq_error)
		return set_err(-EINVAL, "fcch_acquire_batch: NULL output");
	FcchArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.toa = rough_toa;
	Stage s(stream);
	int rc = fcch_common(fcch_type, a, s, iq_len, "fcch_acquire_batch: bad argument");
	if (rc)
		return rc;
	if (win_len / sps < a.len)
		return set_err(-EINVAL, "fcch_acquire_batch: window shorter than the FCCH burst");
	if (n == 0)
		return 0;
	if (!a.toa)
		a.toa = s.tmp<int32_t>((size_t)n);
	FcchArgs f = a;                      // fine stage on the burst the rough stage found (gmr1_rx.c:626-636)
	f.rel = a.toa;
	f.rel_max = win_len - a.len * sps;
	f.rel_add = 1;
	f.win_len = a.len * sps;
	f.toa = s.out(align, (size_t)n);
	f.freq_error = s.out(freq_error, (size_t)n);
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_fcch_rough(a, (cudaStream_t)stream);
		if (e == cudaSuccess) {
			g_launches.fetch_add(1);
			e = launch_fcch_fine(f, 0, (cudaStream_t)stream);
			if (e == cudaSuccess)
				g_launches.fetch_add(1);
		}
	}
	return s.finish(e, "fcch_acquire kernels");
}

int gmr1b200_fcch_snr_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                            int64_t win_stride, int sps, const float *freq_shift, float freq_shift0,
                            float *snr, int n, void *stream)
{
	if (!snr)
		return set_err(-EINVAL, "fcch_snr_batch: NULL output");
	return fine_or_snr(1, fcch_type, iq, iq_len, win_ofs, win_stride, sps, freq_shift, freq_shift0,
	                   nullptr, nullptr, snr, n, stream);
}

static int misc_common(MiscArgs &a, Stage &s, int64_t iq_len)
{
	if (a.n < 0 || !a.iq || a.sps < 1 || a.sps > 16 || a.win_len < 1)
		return set_err(-EINVAL, "sdr batch: bad argument");
	if (a.n && !a.ofs && (a.stride < 0 || (int64_t)(a.n - 1) * a.stride + a.win_len > iq_len))
		return set_err(-EINVAL, "sdr batch: windows exceed iq_len");
	const size_t n = (size_t)a.n;
	a.iq = (const float2 *)s.in((const float *)a.iq, (size_t)iq_len * 2);
	a.ofs = s.in(a.ofs, n);
	a.freq_shift = s.in(a.freq_shift, n);
	a.dkab_p = s.in(a.dkab_p, n);
	a.ebits = s.out(a.ebits, n * 8);
	a.toa = s.out(a.toa, n);
	a.rv = s.out(a.rv, n);
	return 0;
}

int gmr1b200_dkab_demod_batch(const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                              int win_len, int sps, const float *freq_shift, float freq_shift0,
                              const int32_t *p, int p0, int8_t *ebits, float *toa, int32_t *rv, int n, void *stream)
{
	if (!rv)
		return set_err(-EINVAL, "dkab_demod_batch: rv NULL");
	MiscArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.dkab_p = p; a.dkab_p0 = p0;
	a.ebits = ebits; a.toa = toa; a.rv = rv;
	Stage s(stream);
	int rc = misc_common(a, s, iq_len);
	if (rc || n == 0)
		return rc;
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_dkab(a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "dkab kernel");
}

int gmr1b200_pi4cxpsk_mod_order_batch(const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                                      int win_len, int sps, const float *freq_shift, float freq_shift0,
                                      int32_t *order, int n, void *stream)
{
	if (!order)
		return set_err(-EINVAL, "mod_order_batch: order NULL");
	MiscArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.rv = order;
	Stage s(stream);
	int rc = misc_common(a, s, iq_len);
	if (rc || n == 0)
		return rc;
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_mod_order(a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "mod_order kernel");
}

}  // extern "C"

extern "C" int gmr1b200_a5_batch(const int32_t *alg, int alg0, const uint8_t *key, const uint32_t *fn, int nbits,
                                 int stride, uint8_t *dl, uint8_t *ul, int n, void *stream)
{
	if (n < 0 || nbits < 0 || stride < nbits || (!dl && !ul) || (n && (!key || !fn)))
		return set_err(-EINVAL, "a5_batch: bad argument");
	if (n == 0 || nbits == 0)
		return 0;
	Stage s(stream);
	A5Args a = {};
	a.alg = s.in(alg, (size_t)n); a.alg0 = alg0;
	a.key = s.in(key, (size_t)n * 8); a.fn = s.in(fn, (size_t)n);
	a.n = n; a.nbits = nbits; a.stride = stride;
	a.dl = s.out(dl, (size_t)n * stride); a.ul = s.out(ul, (size_t)n * stride);
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_a5(a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "a5 kernel");
}

extern "C" int gmr1b200_gsmtap_batch(const uint8_t *chan_type, int chan_type0, const uint32_t *fn, uint32_t fn0,
                                     const uint8_t *tn, int tn0, const uint8_t *l2, int l2_stride, int len,
                                     uint8_t *out, int out_stride, int n, void *stream)
{
	if (n < 0 || len < 0 || l2_stride < len || out_stride < 16 + len || (n && (!out || (len && !l2))))
		return set_err(-EINVAL, "gsmtap_batch: bad argument");
	if (n == 0)
		return 0;
	Stage s(stream);
	GsmtapArgs a = {};
	a.chan_type = s.in(chan_type, (size_t)n); a.chan_type0 = (uint8_t)chan_type0;
	a.fn = s.in(fn, (size_t)n); a.fn0 = fn0;
	a.tn = s.in(tn, (size_t)n); a.tn0 = (uint8_t)tn0;
	a.l2 = s.in(l2, (size_t)n * l2_stride);
	a.l2_stride = l2_stride; a.len = len; a.n = n; a.out_stride = out_stride;
	a.out = s.out(out, (size_t)n * out_stride);
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_gsmtap(a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "gsmtap kernel");
}
