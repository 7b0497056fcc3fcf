// fcch_kernels.cu - stage 1 of the receive path on the GPU: FCCH chirp acquisition.
//
//   fcch_rough_kernel   one CTA per search window (330 ms = 30 888 samples at sps 4, 247 KB read
//                       once).  Window statistics + decimation to 1 sample/symbol into shared
//                       memory, 117-tap real x complex sliding correlation with the dual-chirp
//                       reference (8 consecutive outputs per thread, samples slid through
//                       registers, padded shared layout), |corr|^2, 5-sample energy window argmax
//                       and centroid.     Replaces gmr1_fcch_rough, src/sdr/fcch.c:211-250.
//   fcch_fine_kernel    one warp per 117-symbol burst: normalise + decimate, multiply with the up /
//                       down chirp (or the real dual chirp for the SNR estimate), centre the
//                       spectrum, DFT-117 by direct evaluation from a twiddle table, energy-window
//                       peak(s).          Replaces gmr1_fcch_fine (:512-628), gmr1_fcch_snr (:643-708).
//
// libosmo-dsp / FFTW semantics as restated in SURVEY.md Appendix A.2 / A.4 (sig_normalize,
// correlate, peak_energy_find with PEAK_WEIGH_WIN, peaks_scan, forward DFT).
// Float contract: integer TOAs equal to the C path except at exact rounding ties (tests allow
// +-1 sample and require >= 99 % equal); freq_error within 1e-5 rad/symbol; SNR within 1e-3 relative.
#include <cuda_runtime.h>
#include <math.h>

#include "launch.h"

namespace gmr1 {

static constexpr float PI_F = 3.14159265358979323846264338327f;
static constexpr int FR_T = 512;            // threads per rough CTA
static constexpr int FR_TILE = 8;           // outputs per thread
static constexpr int MAX_FCCH_LEN = 468;    // FCCH3 bursts are 12 slots

__device__ __forceinline__ float warp_sum_f(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// block-wide sum of up to 3 values; result valid in all threads
__device__ __forceinline__ void block_sum3(float &a, float &b, float &c, float *scratch /* [3*32] */)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	a = warp_sum_f(a); b = warp_sum_f(b); c = warp_sum_f(c);
	__syncthreads();
	if (lane == 0) {
		scratch[warp] = a; scratch[32 + warp] = b; scratch[64 + warp] = c;
	}
	__syncthreads();
	a = lane < nw ? scratch[lane] : 0.0f;
	b = lane < nw ? scratch[32 + lane] : 0.0f;
	c = lane < nw ? scratch[64 + lane] : 0.0f;
	a = warp_sum_f(a); b = warp_sum_f(b); c = warp_sum_f(c);
}

// padded index: one spare slot every 8 so that threads 8 outputs apart hit different banks
__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(FR_T) fcch_rough_kernel(const FcchArgs a)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int tid = threadIdx.x, b = blockIdx.x;
	const int sps = a.sps, L = a.win_len, len = a.len;
	const int l = L / sps;                 // decimated length
	const int nc = l - len + 1;            // correlation outputs
	float *sre = (float *)smem;            // [pad8(l)+8] decimated, normalised samples (SoA)
	float *sim = sre + pad8(l) + 8;
	float *ref = sim + pad8(l) + 8;        // [len]
	float *red = ref + ((len + 3) & ~3);   // [96] reduction scratch
	float *en  = red + 96;                 // [nc] |corr|^2

	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;

	// dual-chirp reference at 1 sample/symbol (fcch.c:167-193)
	{
		const float phase_base = a.freq * 2.0f * PI_F / (float)len, halfpos = (float)len / 2.0f;
		for (int i = tid; i < len; i += FR_T) {
			const float pos = (float)i - halfpos;
			ref[i] = sqrtf(2.0f) * cosf(phase_base * (pos * pos));
		}
	}

	// statistics over ALL samples (sig_normalize averages before decimating); keep every sps-th
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	for (int i = tid; i < L; i += FR_T) {
		const float2 v = __ldg(&x[i]);
		sr += v.x;
		si += v.y;
		sq = fmaf(v.x, v.x, sq);
		sq = fmaf(v.y, v.y, sq);
		if (i % sps == 0 && i / sps < l) {
			sre[pad8(i / sps)] = v.x;
			sim[pad8(i / sps)] = v.y;
		}
	}
	block_sum3(sr, si, sq, red);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	for (int i = tid; i < l; i += FR_T) {
		float yr = (sre[pad8(i)] - ar) / sd, yi = (sim[pad8(i)] - ai) / sd;
		if (freq_shift != 0.0f) {
			float sn, cs;
			sincosf(freq_shift * (float)i, &sn, &cs);
			const float tr = yr * cs - yi * sn;
			yi = yr * sn + yi * cs;
			yr = tr;
		}
		sre[pad8(i)] = yr;
		sim[pad8(i)] = yi;
	}
	__syncthreads();

	// sliding correlation: thread -> FR_TILE consecutive outputs, samples slide through registers
	for (int m0 = tid * FR_TILE; m0 < nc; m0 += FR_T * FR_TILE) {
		float cr[FR_TILE], ci[FR_TILE], wr[FR_TILE], wi[FR_TILE];
#pragma unroll
		for (int t = 0; t < FR_TILE; t++) {
			cr[t] = ci[t] = 0.0f;
			const int i = min(m0 + t, l - 1);
			wr[t] = sre[pad8(i)];
			wi[t] = sim[pad8(i)];
		}
		for (int n0 = 0; n0 < len; n0 += FR_TILE) {
#pragma unroll
			for (int u = 0; u < FR_TILE; u++) {
				const int n = n0 + u;
				if (n < len) {
					const float r = ref[n];
#pragma unroll
					for (int t = 0; t < FR_TILE; t++) {
						// window register (u + t) % FR_TILE holds sample m0 + n + t
						cr[t] = fmaf(r, wr[(u + t) % FR_TILE], cr[t]);
						ci[t] = fmaf(r, wi[(u + t) % FR_TILE], ci[t]);
					}
					const int i = min(m0 + n + FR_TILE, l - 1);
					wr[u] = sre[pad8(i)];
					wi[u] = sim[pad8(i)];
				}
			}
		}
#pragma unroll
		for (int t = 0; t < FR_TILE; t++)
			if (m0 + t < nc)
				en[m0 + t] = cr[t] * cr[t] + ci[t] * ci[t];
	}
	__syncthreads();

	if (a.en_out) {        // multi-FCCH: hand the correlation power to the host-side peak bookkeeping
		float *o = a.en_out + (size_t)b * nc;
		for (int i = tid; i < nc; i += FR_T)
			o[i] = en[i];
		return;
	}

	// highest-energy window of 5 (osmo_cxvec_peak_energy_find, PEAK_WEIGH_WIN), fcch.c:238
	const int win = nc < 5 ? nc : 5;
	float best = 0.0f;
	int best_idx = 0x7fffffff;
	for (int idx = tid; idx < nc; idx += FR_T) {
		float val = 0.0f;
		for (int hi = idx - win + 1; hi <= idx; hi++)
			if (hi >= 0)
				val += en[hi];
		if (val > best) {
			best = val;
			best_idx = idx;
		}
	}
	// block argmax (value, then lowest index)
	{
		const int lane = tid & 31, warp = tid >> 5, nw = FR_T >> 5;
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
			if (ov > best || (ov == best && oi < best_idx)) { best = ov; best_idx = oi; }
		}
		__syncthreads();
		int *redi = (int *)(red + 32);
		if (lane == 0) { red[warp] = best; redi[warp] = best_idx; }
		__syncthreads();
		best = lane < nw ? red[lane] : 0.0f;
		best_idx = lane < nw ? redi[lane] : 0x7fffffff;
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, best, o);
			const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
			if (ov > best || (ov == best && oi < best_idx)) { best = ov; best_idx = oi; }
		}
	}
	if (tid == 0) {
		int max_idx = best > 0.0f ? best_idx - win + 1 : 0;
		if (max_idx < 0)
			max_idx = 0;
		float mw = 0.0f, sw = 0.0f;
		for (int idx = max_idx; idx < max_idx + win; idx++) {
			const float e = en[idx];
			sw += e;
			mw += e * (float)idx;
		}
		const float pos = sw > 0.0f ? mw / sw : (float)max_idx;
		a.toa[b] = (int)round((double)(pos * (float)sps));
		if (a.peak)
			a.peak[b] = best;
	}
}

// ---- fine / snr ---------------------------------------------------------------------------------
// DFT-len twiddles e^{-2*pi*i*k/len}, uploaded per len
__constant__ float2 c_tw[MAX_FCCH_LEN];

// peak_energy_find(cv, 5, PEAK_WEIGH_WIN) over re/im arrays in shared memory; all lanes return pos
__device__ float peak_weigh5(const float *re, const float *im, int n, int lane)
{
	const int win = n < 5 ? n : 5;
	float best = 0.0f;
	int best_idx = 0x7fffffff;
	for (int idx = lane; idx < n; idx += 32) {
		float val = 0.0f;
		for (int hi = idx - win + 1; hi <= idx; hi++)
			if (hi >= 0)
				val += re[hi] * re[hi] + im[hi] * im[hi];
		if (val > best) {
			best = val;
			best_idx = idx;
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		const float ov = __shfl_xor_sync(0xffffffffu, best, o);
		const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
		if (ov > best || (ov == best && oi < best_idx)) { best = ov; best_idx = oi; }
	}
	int max_idx = best > 0.0f ? best_idx - win + 1 : 0;
	if (max_idx < 0)
		max_idx = 0;
	float mw = 0.0f, sw = 0.0f;
	for (int idx = max_idx; idx < max_idx + win; idx++) {
		const float e = re[idx] * re[idx] + im[idx] * im[idx];
		sw += e;
		mw += e * (float)idx;
	}
	return sw > 0.0f ? mw / sw : (float)max_idx;
}

static constexpr int FF_WARPS = 4;

// mode 0: fine (toa + freq_error), mode 1: snr
__global__ void __launch_bounds__(FF_WARPS * 32) fcch_fine_kernel(const FcchArgs a, int mode)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int b = blockIdx.x * FF_WARPS + warp;
	if (b >= a.n)
		return;
	const int len = a.len, sps = a.sps, L = a.win_len;
	const int lp = (len + 3) & ~3;
	float *base = (float *)smem + (size_t)warp * lp * 6;
	float *ur = base, *ui = base + lp;            // mix with up chirp (or dual chirp) -> spectrum in
	float *dr = base + 2 * lp, *di = base + 3 * lp;
	float *Ur = base + 4 * lp, *Ui = base + 5 * lp;   // spectrum out (reused for both DFTs)

	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;

	// sig_normalize(burst_in, sps, freq_shift): statistics over all samples, keep every sps-th
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	for (int i = lane; i < L; i += 32) {
		const float2 v = __ldg(&x[i]);
		sr += v.x;
		si += v.y;
		sq = fmaf(v.x, v.x, sq);
		sq = fmaf(v.y, v.y, sq);
	}
	sr = warp_sum_f(sr); si = warp_sum_f(si); sq = warp_sum_f(sq);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;

	const float sq2d2 = sqrtf(2.0f) / 2.0f;
	const float phase_base = a.freq * 2.0f * PI_F / (float)len, halfpos = (float)len / 2.0f;
	const int mid = len >> 1;
	for (int i = lane; i < len; i += 32) {
		const float2 v = __ldg(&x[i * sps]);
		float yr = (v.x - ar) / sd, yi = (v.y - ai) / sd;
		if (freq_shift != 0.0f) {
			float sn, cs;
			sincosf(freq_shift * (float)i, &sn, &cs);
			const float tr = yr * cs - yi * sn;
			yi = yr * sn + yi * cs;
			yr = tr;
		}
		const float pos = (float)i - halfpos;
		const float ph = phase_base * (pos * pos);
		if (mode == 0) {
			// reference up / down chirps (fcch.c:92-121) and the centring shift (:575-580)
			float sn, cs;
			sincosf(ph, &sn, &cs);
			const float upr = sq2d2 * cs, upi = sq2d2 * sn;          // e^{+j ph}
			float mur = yr * upr - yi * upi, mui = yr * upi + yi * upr;
			float mdr = yr * upr + yi * upi, mdi = yi * upr - yr * upi;   // times e^{-j ph}
			const float ang = 2.0f * PI_F * (float)mid / (float)len * (float)i;
			double dsn, dcs;
			sincos((double)ang, &dsn, &dcs);
			const float fr = (float)dcs, fi = (float)dsn;
			ur[i] = mur * fr - mui * fi; ui[i] = mur * fi + mui * fr;
			dr[i] = mdr * fr - mdi * fi; di[i] = mdr * fi + mdi * fr;
		} else {
			const float dual = sqrtf(2.0f) * cosf(ph);               // fcch.c:183-189, :678-679
			ur[i] = yr * dual;
			ui[i] = yi * dual;
		}
	}
	__syncwarp();

	// forward DFT by direct evaluation: X[k] = sum_n x[n] W^(k*n mod len)
	float peak[2] = {0.0f, 0.0f};
	const int ndft = mode == 0 ? 2 : 1;
	for (int t = 0; t < ndft; t++) {
		const float *xr = t ? dr : ur, *xi = t ? di : ui;
		for (int k = lane; k < len; k += 32) {
			float accr = 0.0f, acci = 0.0f;
			int idx = 0;
			for (int n = 0; n < len; n++) {
				const float2 w = c_tw[idx];
				const float vr = xr[n], vi = xi[n];
				accr = fmaf(vr, w.x, accr);
				accr = fmaf(-vi, w.y, accr);
				acci = fmaf(vr, w.y, acci);
				acci = fmaf(vi, w.x, acci);
				idx += k;
				idx -= idx >= len ? len : 0;
			}
			Ur[k] = accr;
			Ui[k] = acci;
		}
		__syncwarp();
		if (mode == 0)
			peak[t] = peak_weigh5(Ur, Ui, len, lane);
		__syncwarp();
	}

	if (mode == 0) {
		if (lane == 0) {
			// fcch.c:599-615
			const float sym_rate = 23400.0f;
			const float bin_hz = sym_rate / (float)len;
			const float peak_up = (peak[0] - (float)mid) * bin_hz, peak_down = (peak[1] - (float)mid) * bin_hz;
			const float freq_err_hz = (peak_up + peak_down) / 2.0f;
			a.freq_error[b] = (2.0f * PI_F * freq_err_hz) / sym_rate;
			const float chirp_rate = (2.0f * a.freq * sym_rate * sym_rate) / (float)(len * 1000);
			const float toa_ms = ((peak_up - peak_down) / 2.0f) / chirp_rate;
			const float toa_samples = (toa_ms * sym_rate * (float)sps) / 1000.0f;
			a.toa[b] = (int)round((double)toa_samples);
		}
	} else {
		// six strongest bins, descending (osmo_cxvec_peaks_scan); ratio of top 2 over bins 5, 6
		float top[6];
		int prev_idx = -1;
		float prev_val = 3.4e38f;
		for (int r = 0; r < 6; r++) {
			float bv = -1.0f;
			int bi = 0x7fffffff;
			for (int k = lane; k < len; k += 32) {
				const float e = Ur[k] * Ur[k] + Ui[k] * Ui[k];
				// strictly after (prev_val, prev_idx) in (value desc, index asc) order
				const bool after = e < prev_val || (e == prev_val && k > prev_idx);
				if (after && (e > bv || (e == bv && k < bi))) {
					bv = e;
					bi = k;
				}
			}
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
			}
			top[r] = bv;
			prev_val = bv;
			prev_idx = bi;
		}
		if (lane == 0)
			a.snr[b] = (top[0] + top[1]) / (top[4] + top[5]);
	}
}

// ---- launchers ----------------------------------------------------------------------------------------
cudaError_t launch_fcch_rough(const FcchArgs &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	const int l = a.win_len / a.sps, nc = l - a.len + 1;
	if (nc < 1 || a.len > MAX_FCCH_LEN)
		return cudaErrorInvalidValue;
	const int pl = l + (l >> 3) + 8;
	const size_t smem = sizeof(float) * ((size_t)2 * pl + ((a.len + 3) & ~3) + 96 + nc);
	if (smem > 227 * 1024)
		return cudaErrorInvalidValue;
	static size_t attr_set[64] = {0};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 64 || attr_set[dev] < smem) {
		cudaError_t e = cudaFuncSetAttribute(fcch_rough_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_set[dev] = smem;
	}
	fcch_rough_kernel<<<a.n, FR_T, smem, st>>>(a);
	return cudaGetLastError();
}

cudaError_t launch_fcch_fine(const FcchArgs &a, int mode, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	if (a.len > MAX_FCCH_LEN || a.win_len != a.len * a.sps)
		return cudaErrorInvalidValue;
	static int tw_len[64] = {0};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 64 || tw_len[dev] != a.len) {
		float2 h[MAX_FCCH_LEN];
		for (int k = 0; k < a.len; k++) {
			const double ang = -2.0 * 3.14159265358979323846 * (double)k / (double)a.len;
			h[k] = make_float2((float)cos(ang), (float)sin(ang));
		}
		cudaError_t e = cudaMemcpyToSymbolAsync(c_tw, h, sizeof(float2) * a.len, 0, cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess)
			return e;
		cudaStreamSynchronize(st);      // h is on the stack
		if (dev < 64)
			tw_len[dev] = a.len;
	}
	const size_t smem = sizeof(float) * 6 * ((a.len + 3) & ~3) * FF_WARPS;
	fcch_fine_kernel<<<(a.n + FF_WARPS - 1) / FF_WARPS, FF_WARPS * 32, smem, st>>>(a, mode);
	return cudaGetLastError();
}

}  // namespace gmr1
