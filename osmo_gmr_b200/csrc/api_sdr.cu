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

// everything of gmr1_fcch_rough_multi behind the correlation (fcch.c:385-483): strongest peak of the first
// cycle, its twin one period later, geometric mix of the two cycles, avg + 3 sigma threshold, sorted
// de-duplicated insert.  corr_pwr [nc] is overwritten.  Returns the number of FCCHs or -EINVAL.
static int rough_multi_peaks(float *corr_pwr, int nc, int blen, int sps, int32_t *peaks_toa, int N)
{
	const int sym_rate = 23400;
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
		return -EINVAL;
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
	const int cnt = rough_multi_peaks(pw.data(), nc, blen, sps, peaks_toa, N);
	if (cnt < 0)
		return set_err(cnt, "fcch_rough_multi: FCCH period mismatch");
	return cnt;
}

// ---- fcch_multi_process (src/gmr1_rx.c:643-741) up to its callback, for n recordings -----------------
// Per recording: all FCCHs of the 650 ms behind the primary one (rough_multi with the primary's frequency
// error), fine TOA / frequency error and SNR of each, the reference's plausibility filter.  Three kernel
// launches per chunk of recordings (correlation, fine, SNR); the scalar bookkeeping between them is the
// reference's, on the host.
int gmr1b200_fcch_multi_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *rec_ofs,
                              const int32_t *rec_len, const int32_t *align, const float *freq_err, int sps, int n,
                              int max_cand, int32_t *n_fcch, int32_t *cand_align, float *cand_snr,
                              float *cand_freq_err, void *stream)
{
	if (fcch_type < 0 || fcch_type > 2 || !iq || !rec_ofs || !rec_len || !align || n < 0 || sps < 1 || sps > 16 ||
	    max_cand < 1 || max_cand > 16 || !n_fcch || !cand_align)
		return set_err(-EINVAL, "fcch_multi_batch: bad argument");
	if (n == 0)
		return 0;
	const int sym_rate = 23400, NPK = 16;                          // mtoa[16], gmr1_rx.c:647
	const int blen = FCCH_TYPES[fcch_type].len;
	const int W = (650 * sym_rate * sps) / 1000;                   // :661
	const int nc = W / sps - blen + 1;
	cudaStream_t cs = (cudaStream_t)stream;
	Stage s(stream);
	const float2 *d_iq = (const float2 *)s.in(iq, (size_t)iq_len * 2);
	const int CH = 256;                                            // recordings per chunk (15 MB of correlation power)
	int64_t *d_ofs = s.tmp<int64_t>((size_t)CH * NPK);
	float *d_fs = s.tmp<float>((size_t)CH * NPK);
	float *d_pw = s.tmp<float>((size_t)CH * nc);
	int32_t *d_toa = s.tmp<int32_t>((size_t)CH * NPK);
	float *d_fe = s.tmp<float>((size_t)CH * NPK), *d_snr = s.tmp<float>((size_t)CH * NPK);
	if (s.failed())
		return s.finish(cudaSuccess, "fcch_multi_batch: staging");
	std::vector<float> pw((size_t)CH * nc), fs, fe, snr;
	std::vector<int64_t> ofs;
	std::vector<int32_t> toa, chan, base, mtoa, first, bad;
	cudaError_t e = cudaSuccess;
	uint64_t launches = 0;
	auto up = [&](void *d, const void *h, size_t bytes) {
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, cs);
	};
	auto down = [&](void *h, const void *d, size_t bytes) {
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, cs);
	};
	for (int c0 = 0; c0 < n && e == cudaSuccess; c0 += CH) {
		const int c1 = c0 + CH < n ? c0 + CH : n;
		// 650 ms window one FCCH burst ahead of the primary alignment (:655-666)
		chan.clear(); base.clear(); ofs.clear(); fs.clear();
		for (int i = c0; i < c1; i++) {
			int b = align[i] - blen * sps;
			if (b < 0)
				b = 0;
			if (b + W > rec_len[i] || rec_ofs[i] < 0 || rec_ofs[i] + rec_len[i] > iq_len) {
				n_fcch[i] = -EINVAL;                               // "Not enough samples"
				continue;
			}
			chan.push_back(i); base.push_back(b);
			ofs.push_back(rec_ofs[i] + b);
			fs.push_back(-(freq_err ? freq_err[i] : 0.0f));        // :668
		}
		const int m = (int)chan.size();
		if (!m)
			continue;
		up(d_ofs, ofs.data(), (size_t)m * sizeof(int64_t));
		up(d_fs, fs.data(), (size_t)m * sizeof(float));
		FcchArgs a = {};
		a.iq = d_iq; a.ofs = d_ofs; a.n = m; a.win_len = W; a.sps = sps;
		a.freq = FCCH_TYPES[fcch_type].freq; a.len = blen;
		a.freq_shift = d_fs; a.toa = d_toa; a.en_out = d_pw;
		if (e == cudaSuccess) {
			e = launch_fcch_rough(a, cs);
			launches++;
		}
		down(pw.data(), d_pw, (size_t)m * nc * sizeof(float));
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(cs);
		if (e != cudaSuccess)
			break;
		// candidates of every recording of the chunk, flattened
		mtoa.clear(); first.assign((size_t)m + 1, 0); bad.assign((size_t)m, 0);
		std::vector<int64_t> cofs;
		std::vector<float> cfs;
		for (int k = 0; k < m; k++) {
			int32_t t[NPK];
			const int cnt = rough_multi_peaks(&pw[(size_t)k * nc], nc, blen, sps, t, NPK);
			bad[k] = cnt < 0 ? cnt : 0;                            // -EINVAL: the two cycles do not line up
			for (int j = 0; j < cnt; j++) {
				mtoa.push_back(t[j]);
				cofs.push_back(ofs[k] + t[j]);                     // :682
				cfs.push_back(fs[k]);
			}
			first[k + 1] = (int32_t)mtoa.size();
		}
		const int nk = (int)mtoa.size();
		for (int k = 0; k < m; k++)
			n_fcch[chan[k]] = bad[k];                              // 0 found so far, or the error
		if (!nk)
			continue;
		toa.resize(nk); fe.resize(nk); snr.resize(nk);
		up(d_ofs, cofs.data(), (size_t)nk * sizeof(int64_t));
		up(d_fs, cfs.data(), (size_t)nk * sizeof(float));
		FcchArgs f = {};
		f.iq = d_iq; f.ofs = d_ofs; f.n = nk; f.win_len = blen * sps; f.sps = sps;
		f.freq = a.freq; f.len = blen; f.freq_shift = d_fs; f.toa = d_toa; f.freq_error = d_fe;
		if (e == cudaSuccess) {
			e = launch_fcch_fine(f, 0, cs);                        // :684
			launches++;
		}
		down(toa.data(), d_toa, (size_t)nk * sizeof(int32_t));
		down(fe.data(), d_fe, (size_t)nk * sizeof(float));
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(cs);
		if (e != cudaSuccess)
			break;
		for (int j = 0; j < nk; j++) {                             // SNR at the fine position (:691-696)
			cofs[j] += toa[j];
			cfs[j] = -(-cfs[j] + fe[j]);
		}
		up(d_ofs, cofs.data(), (size_t)nk * sizeof(int64_t));
		up(d_fs, cfs.data(), (size_t)nk * sizeof(float));
		f.toa = nullptr; f.freq_error = nullptr; f.snr = d_snr;
		if (e == cudaSuccess) {
			e = launch_fcch_fine(f, 1, cs);
			launches++;
		}
		down(snr.data(), d_snr, (size_t)nk * sizeof(float));
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(cs);
		if (e != cudaSuccess)
			break;
		// the strongest is the reference; the others must be strong enough and close in frequency (:702-717)
		for (int k = 0; k < m; k++) {
			const int i = chan[k];
			if (bad[k])
				continue;
			float ref_snr = 0.0f, ref_fe = 0.0f;
			int cnt = 0;
			for (int j = first[k]; j < first[k + 1]; j++) {
				if (j == first[k]) {
					ref_snr = snr[j];
					ref_fe = fe[j];
				} else {
					if (snr[j] < 2.0f)
						continue;
					if (snr[j] < ref_snr / 6.0f)
						continue;
					const float d = (float)fabs(ref_fe - fe[j]);
					if ((sym_rate * d) / (2.0f * (float)M_PI) > 500.0f)
						continue;
				}
				if (cnt < max_cand) {
					const size_t o = (size_t)i * max_cand + cnt;
					cand_align[o] = base[k] + mtoa[j] + toa[j];     // :738, where process_bcch starts
					if (cand_snr)
						cand_snr[o] = snr[j];
					if (cand_freq_err)
						cand_freq_err[o] = fe[j];
				}
				cnt++;
			}
			n_fcch[i] = cnt < max_cand ? cnt : max_cand;
		}
	}
	g_launches.fetch_add(launches);
	if (e != cudaSuccess)
		cudaStreamSynchronize(cs);
	return s.finish(e, "fcch_multi_batch kernels");
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
