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

// ---- fcch_multi_process (src/gmr1_rx.c:643-741) up to its callback, for n recordings, on the device -----------------
// Per recording: all FCCHs of the 650 ms behind the primary one (gmr1_fcch_rough_multi with the primary's frequency
// error: the 117-tap correlation by fcch_rough_kernel, the peak bookkeeping of src/sdr/fcch.c:264-326, 385-483 by
// multi_peaks_kernel), fine TOA / frequency error and SNR of each candidate (fcch_fine_kernel over the [n][16]
// candidate slots), the reference's plausibility filter (multi_filter_kernel).  Everything is enqueued on the
// caller's stream; there is no host round trip inside the call.
namespace {

constexpr int NPK = 16;                                            // mtoa[16], gmr1_rx.c:647

struct MultiSt {                                                   // device memory
	int64_t *w_ofs;                                                // [n] 650 ms window of a recording
	float   *w_fs;                                                 // [n] -freq_err of the primary FCCH
	int32_t *w_skip, *base;                                        // [n]
	float   *pw;                                                   // [n][nc] correlation power
	int32_t *cnt;                                                  // [n] peaks found (or -EINVAL)
	int32_t *mtoa;                                                 // [n][16]
	int64_t *c_ofs;                                                // [n][16] candidate burst
	float   *c_fs;
	int32_t *c_skip, *c_toa;
	float   *c_fe, *c_snr;
};

// window one FCCH burst ahead of the primary alignment (gmr1_rx.c:655-668)
__global__ void multi_prep_kernel(MultiSt m, const int64_t *rec_ofs, const int32_t *rec_len, const int32_t *align,
                                  const float *freq_err, int64_t iq_len, int blen, int sps, int W, int32_t *n_fcch, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int b = align[i] - blen * sps;
	if (b < 0)
		b = 0;
	const bool bad = b + W > rec_len[i] || rec_ofs[i] < 0 || rec_ofs[i] + rec_len[i] > iq_len;
	m.w_skip[i] = bad;
	m.base[i] = b;
	m.w_ofs[i] = bad ? 0 : rec_ofs[i] + b;
	m.w_fs[i] = -(freq_err ? freq_err[i] : 0.0f);
	n_fcch[i] = bad ? -EINVAL : 0;                                 // "Not enough samples"
}

// _peak_record, src/sdr/fcch.c:264-326
__device__ void peak_record_dev(int burst_len, int *toa, float *pwr, int *n, int N, int Lp, int sps, int peak_toa, float peak_pwr)
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

// everything of gmr1_fcch_rough_multi behind the correlation (fcch.c:385-483), one warp per recording: the lanes bring
// the correlation power through shared memory in coalesced tiles, lane 0 does the arithmetic in the reference's order
// (its running sums are sequential float additions; the threshold is compared against them)
constexpr int MP_TILE = 1024;
__global__ void __launch_bounds__(32) multi_peaks_kernel(MultiSt m, int nc, int blen, int sps, int n)
{
	__shared__ float tile[MP_TILE + 2];
	__shared__ int s_toa[NPK];
	__shared__ float s_pwr[NPK];
	const int r = blockIdx.x, lane = threadIdx.x;
	if (r >= n || m.w_skip[r])
		return;
	float *pw = m.pw + (size_t)r * nc;
	const int sym_rate = 23400;
	int Lw = (320 * sym_rate) / 1000 + blen, Lp = (320 * sym_rate) / 1000;
	// strongest sample of the first cycle, first one on ties (:385-392)
	float best = 0.0f;
	int best_i = 0x7fffffff;
	for (int i = lane; i < nc && i < Lw; i += 32)
		if (pw[i] > best) {
			best = pw[i];
			best_i = i;
		}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		const float ov = __shfl_xor_sync(0xffffffffu, best, o);
		const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
		if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
	}
	const int pwr_max_idx = best > 0.0f ? best_i : 0;
	int cnt = 0, err = 0;
	if (lane == 0) {
		// the twin one period later, +-10 symbols (:397-430)
		float pwrs[2] = {0.0f, 0.0f}, peaks[2] = {0.0f, 0.0f};
		for (int i = -10; i <= 10; i++) {
			int j = pwr_max_idx + i;
			if (j > 0 && j < nc) {
				pwrs[0] += pw[j];
				peaks[0] += pw[j] * j;
			}
			j += Lp;
			if (j > 0 && j < nc) {
				pwrs[1] += pw[j];
				peaks[1] += pw[j] * j;
			}
		}
		peaks[0] /= pwrs[0];
		peaks[1] /= pwrs[1];
		const int nLp = (int)round((double)(peaks[1] - peaks[0]));
		if (abs(nLp - Lp) > 10)
			err = 1;
		Lp = nLp;
	}
	err = __shfl_sync(0xffffffffu, err, 0);
	Lp = __shfl_sync(0xffffffffu, Lp, 0);
	if (err) {
		if (lane == 0)
			m.cnt[r] = -EINVAL;
		return;
	}
	if (Lw + Lp > nc)
		Lw = nc - Lp;
	// geometric mix of the two cycles (:435-441), in place; its mean, sequentially
	float avg = 0.0f;
	for (int t0 = 0; t0 < Lw; t0 += MP_TILE) {
		const int tn = min(MP_TILE, Lw - t0);
		for (int i = lane; i < tn; i += 32) {
			const float v = sqrtf(pw[t0 + i] * pw[t0 + i + Lp]);
			tile[i] = v;
		}
		__syncwarp();
		for (int i = lane; i < tn; i += 32)
			pw[t0 + i] = tile[i];
		if (lane == 0)
			for (int i = 0; i < tn; i++)
				avg += tile[i];
		__syncwarp();
	}
	avg = __shfl_sync(0xffffffffu, avg, 0) / Lw;
	float stddev = 0.0f;
	for (int t0 = 0; t0 < Lw; t0 += MP_TILE) {
		const int tn = min(MP_TILE, Lw - t0);
		for (int i = lane; i < tn; i += 32)
			tile[i] = pw[t0 + i];
		__syncwarp();
		if (lane == 0)
			for (int i = 0; i < tn; i++) {
				const float v = tile[i] - avg;
				stddev += v * v;
			}
		__syncwarp();
	}
	const float th = avg + 3.0f * sqrtf(__shfl_sync(0xffffffffu, stddev, 0) / Lw);
	// peaks above the threshold, 3-point centroid, sorted de-duplicated insert (:454-481)
	if (lane < NPK) {
		s_toa[lane] = 0;
		s_pwr[lane] = 0.0f;
	}
	__syncwarp();
	int in_peak = 0;
	for (int t0 = 1; t0 < Lw - 1; t0 += MP_TILE) {
		const int tn = min(MP_TILE, Lw - 1 - t0);
		for (int i = lane; i < tn + 2; i += 32)                    // tile[k] = pw[t0 - 1 + k]
			tile[i] = pw[t0 - 1 + i];
		__syncwarp();
		if (lane == 0)
			for (int k = 0; k < tn; k++) {
				const int i = t0 + k;
				if (tile[k + 1] > th) {
					if (in_peak)
						continue;
					in_peak = 1;
					const float p_pwr = tile[k] + tile[k + 1] + tile[k + 2];
					const float p_fpos = (-tile[k] + tile[k + 2]) / p_pwr;
					const int p_pos = (int)round((double)(((float)i + p_fpos) * (float)sps));
					peak_record_dev(blen, s_toa, s_pwr, &cnt, NPK, Lp, sps, p_pos, p_pwr);
				} else {
					in_peak = 0;
				}
			}
		__syncwarp();
	}
	if (lane == 0)
		m.cnt[r] = cnt;
	__syncwarp();
	if (lane < NPK)
		m.mtoa[r * NPK + lane] = s_toa[lane];
}

// candidate slots of every recording: the burst at each rough position (gmr1_rx.c:680-684)
__global__ void multi_cand_kernel(MultiSt m, int n)
{
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n * NPK)
		return;
	const int r = e / NPK, j = e % NPK;
	const bool live = !m.w_skip[r] && j < m.cnt[r];
	m.c_skip[e] = !live;
	m.c_ofs[e] = live ? m.w_ofs[r] + m.mtoa[e] : 0;
	m.c_fs[e] = m.w_fs[r];
}

// SNR at the fine position, with the fine frequency error taken off (:691-696)
__global__ void multi_snr_prep_kernel(MultiSt m, int n)
{
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n * NPK || m.c_skip[e])
		return;
	m.c_ofs[e] += m.c_toa[e];
	m.c_fs[e] = -(-m.c_fs[e] + m.c_fe[e]);
}

// the strongest is the reference; the others must be strong enough and close in frequency (:702-738)
__global__ void multi_filter_kernel(MultiSt m, int max_cand, int32_t *n_fcch, int32_t *cand_align, float *cand_snr,
                                    float *cand_freq_err, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || m.w_skip[i])
		return;
	if (m.cnt[i] < 0) {
		n_fcch[i] = m.cnt[i];
		return;
	}
	float ref_snr = 0.0f, ref_fe = 0.0f;
	int cnt = 0;
	for (int j = 0; j < m.cnt[i]; j++) {
		const int e = i * NPK + j;
		const float snr = m.c_snr[e], fe = m.c_fe[e];
		if (j == 0) {
			ref_snr = snr;
			ref_fe = fe;
		} else {
			if (snr < 2.0f)
				continue;
			if (snr < ref_snr / 6.0f)
				continue;
			const float d = fabsf(ref_fe - fe);
			if ((23400 * d) / (2.0f * (float)M_PI) > 500.0f)
				continue;
		}
		if (cnt < max_cand) {
			const size_t o = (size_t)i * max_cand + cnt;
			cand_align[o] = m.base[i] + m.mtoa[e] + m.c_toa[e];     // :738, where process_bcch starts
			if (cand_snr)
				cand_snr[o] = snr;
			if (cand_freq_err)
				cand_freq_err[o] = fe;
		}
		cnt++;
	}
	n_fcch[i] = cnt < max_cand ? cnt : max_cand;
}

}  // namespace

// replaces gmr1_fcch_rough_multi for one window: the same two kernels with n = 1
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
	cudaStream_t cs = (cudaStream_t)stream;
	int32_t cnt = 0, mtoa[NPK] = {0};
	int rc;
	{
		Stage s(stream);
		const float2 *d_iq = (const float2 *)s.in(iq, (size_t)win_len * 2);
		MultiSt m = {};
		m.w_skip = s.tmp<int32_t>(1); m.pw = s.tmp<float>((size_t)nc); m.c_toa = s.tmp<int32_t>(1);
		m.cnt = s.out(&cnt, 1); m.mtoa = s.out(mtoa, NPK);
		cudaError_t e = cudaSuccess;
		if (!s.failed()) {
			cudaMemsetAsync(m.w_skip, 0, sizeof(int32_t), cs);
			cudaMemsetAsync(m.cnt, 0, sizeof(int32_t), cs);
			FcchArgs a = {};
			a.iq = d_iq; a.stride = 0; a.n = 1; a.win_len = (int)win_len; a.sps = sps;
			a.freq = FCCH_TYPES[fcch_type].freq; a.len = blen; a.freq_shift0 = freq_shift;
			a.toa = m.c_toa; a.en_out = m.pw;
			e = launch_fcch_rough(a, cs);
			if (e == cudaSuccess) {
				multi_peaks_kernel<<<1, 32, 0, cs>>>(m, nc, blen, sps, 1);
				e = cudaGetLastError();
				g_launches.fetch_add(2);
			}
		}
		rc = s.finish(e, "fcch_rough_multi kernels");       // cnt / mtoa are host memory: finish() synchronises
	}
	if (rc)
		return rc;
	if (cnt < 0)
		return set_err(cnt, "fcch_rough_multi: FCCH period mismatch");
	for (int i = 0; i < cnt && i < N; i++)
		peaks_toa[i] = mtoa[i];
	return cnt < N ? cnt : N;
}

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
	const int sym_rate = 23400;
	const int blen = FCCH_TYPES[fcch_type].len;
	const int W = (650 * sym_rate * sps) / 1000;                   // :661
	const int nc = W / sps - blen + 1;
	cudaStream_t cs = (cudaStream_t)stream;
	Stage s(stream);
	const size_t N = (size_t)n;
	const float2 *d_iq = (const float2 *)s.in(iq, (size_t)iq_len * 2);
	const int64_t *d_rec_ofs = s.in(rec_ofs, N);
	const int32_t *d_rec_len = s.in(rec_len, N), *d_align = s.in(align, N);
	const float *d_ferr = s.in(freq_err, N);
	int32_t *d_nf = s.out(n_fcch, N), *d_ca = s.out(cand_align, N * max_cand);
	float *d_cs = s.out(cand_snr, N * max_cand), *d_cf = s.out(cand_freq_err, N * max_cand);
	const int CH = 2048;                                           // recordings per chunk (124 MB of correlation power)
	const size_t C = (size_t)(n < CH ? n : CH);
	MultiSt m = {};
	m.w_ofs = s.tmp<int64_t>(N); m.w_fs = s.tmp<float>(N); m.w_skip = s.tmp<int32_t>(N); m.base = s.tmp<int32_t>(N);
	m.pw = s.tmp<float>(C * nc);
	m.cnt = s.tmp<int32_t>(N); m.mtoa = s.tmp<int32_t>(N * NPK);
	m.c_ofs = s.tmp<int64_t>(N * NPK); m.c_fs = s.tmp<float>(N * NPK); m.c_skip = s.tmp<int32_t>(N * NPK);
	m.c_toa = s.tmp<int32_t>(N * NPK); m.c_fe = s.tmp<float>(N * NPK); m.c_snr = s.tmp<float>(N * NPK);
	if (s.failed())
		return s.finish(cudaSuccess, "fcch_multi_batch: staging");
	const int tb = 128;
	uint64_t launches = 0;
	multi_prep_kernel<<<(n + tb - 1) / tb, tb, 0, cs>>>(m, d_rec_ofs, d_rec_len, d_align, d_ferr, iq_len, blen, sps, W, d_nf, n);
	cudaMemsetAsync(m.cnt, 0, N * sizeof(int32_t), cs);
	cudaError_t e = cudaGetLastError();
	launches++;
	for (int c0 = 0; c0 < n && e == cudaSuccess; c0 += CH) {       // correlation + peak bookkeeping, chunk by chunk
		const int c = n - c0 < CH ? n - c0 : CH;
		FcchArgs a = {};
		a.iq = d_iq; a.ofs = m.w_ofs + c0; a.n = c; a.win_len = W; a.sps = sps;
		a.freq = FCCH_TYPES[fcch_type].freq; a.len = blen;
		a.freq_shift = m.w_fs + c0; a.toa = m.c_toa; a.en_out = m.pw; a.skip = m.w_skip + c0;
		if ((e = launch_fcch_rough(a, cs)) != cudaSuccess)
			break;
		MultiSt mc = m;
		mc.w_skip += c0; mc.cnt += c0; mc.mtoa += (size_t)c0 * NPK;
		multi_peaks_kernel<<<c, 32, 0, cs>>>(mc, nc, blen, sps, c);
		e = cudaGetLastError();
		launches += 2;
	}
	if (e == cudaSuccess) {
		const int ne = n * NPK;
		multi_cand_kernel<<<(ne + tb - 1) / tb, tb, 0, cs>>>(m, n);
		FcchArgs f = {};
		f.iq = d_iq; f.ofs = m.c_ofs; f.n = ne; f.win_len = blen * sps; f.sps = sps;
		f.freq = FCCH_TYPES[fcch_type].freq; f.len = blen; f.freq_shift = m.c_fs; f.skip = m.c_skip;
		f.toa = m.c_toa; f.freq_error = m.c_fe;
		e = launch_fcch_fine(f, 0, cs);                            // :684
		if (e == cudaSuccess) {
			multi_snr_prep_kernel<<<(ne + tb - 1) / tb, tb, 0, cs>>>(m, n);
			f.toa = nullptr; f.freq_error = nullptr; f.snr = m.c_snr;
			e = launch_fcch_fine(f, 1, cs);                        // :694
		}
		if (e == cudaSuccess) {
			multi_filter_kernel<<<(n + tb - 1) / tb, tb, 0, cs>>>(m, max_cand, d_nf, d_ca, d_cs, d_cf, n);
			e = cudaGetLastError();
		}
		launches += 5;
	}
	g_launches.fetch_add(launches);
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

int gmr1b200_set_fcch_fft(int on)
{
	static std::atomic<int> cur{1};
	const int mode = on == 2 ? 2 : (on ? 1 : 0);
	const int prev = cur.exchange(mode);
	fcch_fft_enable(mode);
	return prev;
}

int gmr1b200_fcch_rough_grid_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                   int64_t win_stride, int win_len, int sps, const float *shifts, int n_shifts,
                                   int32_t *toa, float *peak, int n, void *stream)
{
	if (!toa || !shifts || n_shifts < 1 || n_shifts > 16)
		return set_err(-EINVAL, "fcch_rough_grid_batch: bad argument");
	FcchArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	Stage s(stream);
	int rc = fcch_common(fcch_type, a, s, iq_len, "fcch_rough_grid_batch: bad argument");
	if (rc)
		return rc;
	if (win_len / sps < a.len)
		return set_err(-EINVAL, "fcch_rough_grid_batch: window shorter than the FCCH burst");
	if (n == 0)
		return 0;
	const size_t N = (size_t)n;
	int32_t *d_toa = s.out(toa, N * n_shifts);
	float *d_peak = s.out(peak, N * n_shifts);
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_fcch_fft(a, shifts, n_shifts, d_toa, d_peak, (cudaStream_t)stream);
		if (e == cudaErrorNotSupported) {
			cudaGetLastError();
			e = launch_fcch_grid(a, shifts, n_shifts, d_toa, d_peak, (cudaStream_t)stream);
		}
		if (e == cudaSuccess) {
			g_launches.fetch_add(1);
		} else if (e == cudaErrorNotSupported) {       // geometry outside the grid kernel: one search per shift
			cudaGetLastError();
			e = cudaSuccess;
			for (int k = 0; k < n_shifts && e == cudaSuccess; k++) {
				FcchArgs b = a;
				b.freq_shift0 = shifts[k];
				b.toa = d_toa + N * k;
				b.peak = d_peak ? d_peak + N * k : nullptr;
				e = launch_fcch_rough(b, (cudaStream_t)stream);
				g_launches.fetch_add(1);
			}
		}
	}
	return s.finish(e, "fcch_rough_grid kernel");
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

extern "C" int gmr1b200_set_a5_bitslice(int mode)
{
	static std::atomic<int> cur{-1};
	const int m = mode < 0 ? -1 : (mode ? 1 : 0);
	const int prev = cur.exchange(m);
	a5_force_mode(m);
	return prev;
}

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
