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
#include <stdlib.h>

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

// Shared layout of the decimated window: interleaved (re, im), every group of 8 samples (16 floats)
// padded to 20 floats.  A thread owns 8 consecutive outputs and pulls whole groups with four
// LDS.128; lanes are 20 words apart, so each quarter-warp covers all 32 banks exactly once.
static constexpr int FR_GRP = 20;
__device__ __forceinline__ int widx(int i) { return (i >> 3) * FR_GRP + (i & 7) * 2; }

// (cr, ci) += r * (wr, wi): one packed FFMA2 on sm_100a
__device__ __forceinline__ void fma2(float2 &c, float r, const float2 w)
{
	unsigned long long cc = *reinterpret_cast<unsigned long long *>(&c);
	const float2 rr2 = make_float2(r, r);
	asm("fma.rn.f32x2 %0, %1, %2, %0;"
	    : "+l"(cc)
	    : "l"(*reinterpret_cast<const unsigned long long *>(&rr2)), "l"(*reinterpret_cast<const unsigned long long *>(&w)));
	c = *reinterpret_cast<float2 *>(&cc);
}

__global__ void __launch_bounds__(FR_T, 2) fcch_rough_kernel(const FcchArgs a)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int tid = threadIdx.x, b = blockIdx.x;
	if (a.skip && a.skip[b])
		return;
	const int sps = a.sps, L = a.win_len, len = a.len;
	const int l = L / sps;                 // decimated length
	const int nc = l - len + 1;            // correlation outputs
	const int lenp = (len + 7) & ~7;       // taps, zero-padded to whole groups
	const int ng = (l + lenp + 15) >> 3;   // sample groups incl. the zero tail the padded taps touch
	float *w   = (float *)smem;            // [ng * FR_GRP] decimated, normalised samples
	float *ref = w + ng * FR_GRP;          // [lenp]
	float *red = ref + lenp;               // [96] reduction scratch
	float *en  = red + 96 + 4;             // [nc] |corr|^2, en[-4..-1] = 0 (energy windows need no edge cases)

	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;

#ifndef FR_PREFETCH
#define FR_PREFETCH 1
#endif
#if FR_PREFETCH
	// the whole search window (247 KB at sps 4) requested into L2 up front, 16 KB per bulk prefetch: the statistics
	// pass below has 64 bytes per thread in flight and would otherwise run at DRAM latency
	if ((((uintptr_t)x) & 15) == 0) {
		const int chunk = 16384, bytes = (L * 8) & ~15;
		for (int o = tid * chunk; o < bytes; o += FR_T * chunk)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)x + o), "r"(min(chunk, bytes - o))
			             : "memory");
	}
#endif
	// dual-chirp reference at 1 sample/symbol (fcch.c:167-193)
	{
		const float phase_base = a.freq * 2.0f * PI_F / (float)len, halfpos = (float)len / 2.0f;
		for (int i = tid; i < lenp; i += FR_T) {
			const float pos = (float)i - halfpos;
			ref[i] = i < len ? sqrtf(2.0f) * cosf(phase_base * (pos * pos)) : 0.0f;
		}
	}
	if (tid < 4)
		en[tid - 4] = 0.0f;
	for (int i = l + tid; i < ng * 8; i += FR_T) {
		w[widx(i)] = 0.0f;
		w[widx(i) + 1] = 0.0f;
	}

	// statistics over ALL samples (sig_normalize averages before decimating); keep every sps-th.
	// (symbol, phase) of sample i are carried along instead of dividing by the runtime sps.
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	if (sps == 4 && (((uintptr_t)x) & 15) == 0) {
		// four samples (one symbol) per thread and pass: two 16-byte loads, packed adds, the first sample is kept
		const float4 *x4 = reinterpret_cast<const float4 *>(x);
		const int nq = L >> 2;
		float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
#pragma unroll 2
		for (int i = tid; i < nq; i += FR_T) {
			const float4 v0 = __ldg(&x4[2 * i]), v1 = __ldg(&x4[2 * i + 1]);
			const float2 p0 = make_float2(v0.x, v0.y), p1 = make_float2(v0.z, v0.w), p2 = make_float2(v1.x, v1.y),
			             p3 = make_float2(v1.z, v1.w);
			s2 = __fadd2_rn(s2, __fadd2_rn(__fadd2_rn(p0, p1), __fadd2_rn(p2, p3)));
			q2 = __ffma2_rn(p0, p0, q2);
			q2 = __ffma2_rn(p1, p1, q2);
			q2 = __ffma2_rn(p2, p2, q2);
			q2 = __ffma2_rn(p3, p3, q2);
			*reinterpret_cast<float2 *>(&w[widx(i)]) = p0;        // i < l = L / 4
		}
		sr = s2.x;
		si = s2.y;
		sq = q2.x + q2.y;
		for (int i = 4 * nq + tid; i < L; i += FR_T) {           // L % 4 trailing samples (none is kept)
			const float2 v = __ldg(&x[i]);
			sr += v.x;
			si += v.y;
			sq = fmaf(v.x, v.x, sq);
			sq = fmaf(v.y, v.y, sq);
		}
	} else {
		const int dq = FR_T / sps, dr = FR_T % sps;
		int sym = tid / sps, ph = tid % sps;
#pragma unroll 1
		for (int i = tid; i < L; i += FR_T) {
			const float2 v = __ldg(&x[i]);
			sr += v.x;
			si += v.y;
			sq = fmaf(v.x, v.x, sq);
			sq = fmaf(v.y, v.y, sq);
			if (ph == 0 && sym < l)
				*reinterpret_cast<float2 *>(&w[widx(sym)]) = v;
			sym += dq;
			ph += dr;
			if (ph >= sps) {
				ph -= sps;
				sym++;
			}
		}
	}
	block_sum3(sr, si, sq, red);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	const float inv_sd = 1.0f / sd;        // a common scale: one reciprocal instead of two divisions per sample
	for (int i = tid; i < l; i += FR_T) {
		float2 *p = reinterpret_cast<float2 *>(&w[widx(i)]);
		float yr = (p->x - ar) * inv_sd, yi = (p->y - ai) * inv_sd;
		if (freq_shift != 0.0f) {
			float sn, cs;
			sincosf(freq_shift * (float)i, &sn, &cs);
			const float tr = yr * cs - yi * sn;
			yi = yr * sn + yi * cs;
			yr = tr;
		}
		*p = make_float2(yr, yi);
	}
	__syncthreads();

	// sliding correlation: thread -> 8 consecutive outputs; the 8-sample register window slides one
	// whole group per iteration (after tap u the slot u is refilled with sample u of the next group)
	for (int m0 = tid * FR_TILE; m0 < nc; m0 += FR_T * FR_TILE) {
		float2 c[FR_TILE], win[FR_TILE];
		const float4 *gp = reinterpret_cast<const float4 *>(w + (m0 >> 3) * FR_GRP);
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const float4 v = gp[q];
			win[2 * q] = make_float2(v.x, v.y);
			win[2 * q + 1] = make_float2(v.z, v.w);
			c[2 * q] = c[2 * q + 1] = make_float2(0.0f, 0.0f);
		}
		const float4 *rp = reinterpret_cast<const float4 *>(ref);
		// one group of 8 taps: `cur` holds samples m0 + 8g .. + 7, `nxt` is loaded with the following 8; output t
		// of tap u reads sample u + t of the 16.  Two groups per iteration with the roles of the two register
		// windows swapped, so no register is ever copied.
		auto group = [&](int g, const float2 (&cur)[FR_TILE], float2 (&nxt)[FR_TILE]) {
			gp += FR_GRP / 4;
#pragma unroll
			for (int q = 0; q < 4; q++) {
				const float4 v = gp[q];
				nxt[2 * q] = make_float2(v.x, v.y);
				nxt[2 * q + 1] = make_float2(v.z, v.w);
			}
			const float4 r0 = rp[2 * g], r1 = rp[2 * g + 1];
			const float r[FR_TILE] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
			for (int u = 0; u < FR_TILE; u++)
#pragma unroll
				for (int t = 0; t < FR_TILE; t++)
					fma2(c[t], r[u], u + t < FR_TILE ? cur[u + t] : nxt[u + t - FR_TILE]);
		};
		float2 alt[FR_TILE];
		const int G = lenp >> 3;
		int g = 0;
#pragma unroll 1
		for (; g + 1 < G; g += 2) {
			group(g, win, alt);
			group(g + 1, alt, win);
		}
		if (g < G)
			group(g, win, alt);
#pragma unroll
		for (int t = 0; t < FR_TILE; t++)
			if (m0 + t < nc)
				en[m0 + t] = c[t].x * c[t].x + c[t].y * c[t].y;
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
		float val;
		if (win == 5) {          // oldest first, as the reference sums; the zeros in front of en[0] stand in for hi < 0
			val = ((((0.0f + en[idx - 4]) + en[idx - 3]) + en[idx - 2]) + en[idx - 1]) + en[idx];
		} else {
			val = 0.0f;
			for (int hi = idx - win + 1; hi <= idx; hi++)
				if (hi >= 0)
					val += en[hi];
		}
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

static constexpr int FF_T = 128;            // threads per fine CTA (one CTA per burst)

// mode 0: fine (toa + freq_error), mode 1: snr
// One CTA per burst.  Warp 0 does the window statistics (lane-strided, the summation order the
// parity tests were pinned with), all threads mix and run the DFT (one bin per thread, both chirp
// directions in the same pass over the twiddle table held in shared memory - the table index differs
// per lane, which a __constant__ table would serialise), warps 0 / 1 search the two spectra.
__global__ void __launch_bounds__(FF_T) fcch_fine_kernel(const FcchArgs a, int mode)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int b = blockIdx.x;
	if (a.skip && a.skip[b])
		return;
	const int len = a.len, sps = a.sps, L = a.win_len;
	const int lp = (len + 3) & ~3;
	float4 *mix = (float4 *)smem;                    // [lp] (up.re, up.im, down.re, down.im) spectra in
	float2 *tw = (float2 *)(mix + lp);               // [lp] e^{-2 pi i k / len}
	float *Ur = (float *)(tw + lp), *Ui = Ur + lp;   // up (or dual) spectrum
	float *Dr = Ui + lp, *Di = Dr + lp;              // down spectrum
	float *stat = Di + lp;                           // [4] ar, ai, sd, -; [4..5] peaks

	int64_t base = a.ofs ? a.ofs[b] : (int64_t)b * a.stride;
	int rel = 0;
	if (a.rel) {                                     // chained after the rough stage (fcch_single_init)
		rel = min(max(a.rel[b], 0), a.rel_max);
		base += rel;
	}
	const float2 *x = a.iq + base;
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;

	for (int k = tid; k < len; k += FF_T)
		tw[k] = c_tw[k];

	// sig_normalize(burst_in, sps, freq_shift): statistics over all samples, keep every sps-th
	if (warp == 0) {
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
		if (lane == 0) {
			stat[0] = ar; stat[1] = ai; stat[2] = sd;
		}
	}
	__syncthreads();
	const float ar = stat[0], ai = stat[1], sd = stat[2];

	const float sq2d2 = sqrtf(2.0f) / 2.0f;
	const float phase_base = a.freq * 2.0f * PI_F / (float)len, halfpos = (float)len / 2.0f;
	const int mid = len >> 1;
	for (int i = tid; i < len; i += FF_T) {
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
			mix[i] = make_float4(mur * fr - mui * fi, mur * fi + mui * fr,
			                     mdr * fr - mdi * fi, mdr * fi + mdi * fr);
		} else {
			const float dual = sqrtf(2.0f) * cosf(ph);               // fcch.c:183-189, :678-679
			mix[i] = make_float4(yr * dual, yi * dual, 0.0f, 0.0f);
		}
	}
	__syncthreads();

	// forward DFT by direct evaluation: X[k] = sum_n x[n] W^(k*n mod len)
	for (int k = tid; k < len; k += FF_T) {
		float ur = 0.0f, ui = 0.0f, dr = 0.0f, di = 0.0f;
		int idx = 0;
		for (int n = 0; n < len; n++) {
			const float2 t = tw[idx];
			const float4 v = mix[n];
			ur = fmaf(v.x, t.x, ur);
			ur = fmaf(-v.y, t.y, ur);
			ui = fmaf(v.x, t.y, ui);
			ui = fmaf(v.y, t.x, ui);
			dr = fmaf(v.z, t.x, dr);
			dr = fmaf(-v.w, t.y, dr);
			di = fmaf(v.z, t.y, di);
			di = fmaf(v.w, t.x, di);
			idx += k;
			idx -= idx >= len ? len : 0;
		}
		Ur[k] = ur; Ui[k] = ui;
		Dr[k] = dr; Di[k] = di;
	}
	__syncthreads();
	float peak[2] = {0.0f, 0.0f};
	if (mode == 0) {
		if (warp < 2) {
			const float p = peak_weigh5(warp ? Dr : Ur, warp ? Di : Ui, len, lane);
			if (lane == 0)
				stat[4 + warp] = p;
		}
		__syncthreads();
		peak[0] = stat[4];
		peak[1] = stat[5];
	}
	if (warp != 0)
		return;

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
			a.toa[b] = (int)round((double)toa_samples) + (a.rel_add ? rel : 0);
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
	static const bool old_only = [] { const char *e = getenv("GMR1B200_FCCH_OLD"); return e && atoi(e) != 0; }();
	if (!a.en_out && !old_only) {          // the search itself: frequency-domain kernel, else the direct grid kernel
		const cudaError_t fe = launch_fcch_fft(a, nullptr, 0, nullptr, nullptr, st);
		if (fe != cudaErrorNotSupported)
			return fe;
		cudaGetLastError();
		const cudaError_t ge = launch_fcch_grid(a, nullptr, 0, nullptr, nullptr, st);
		if (ge != cudaErrorNotSupported)
			return ge;
		cudaGetLastError();
	}
	const int l = a.win_len / a.sps, nc = l - a.len + 1;
	if (nc < 1 || a.len > MAX_FCCH_LEN)
		return cudaErrorInvalidValue;
	const int lenp = (a.len + 7) & ~7, ng = (l + lenp + 15) >> 3;
	const size_t smem = sizeof(float) * ((size_t)ng * FR_GRP + lenp + 96 + 4 + nc);
	if (smem > 227 * 1024)
		return cudaErrorInvalidValue;
	GMR1_INIT_LOCK();
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
	GMR1_INIT_LOCK();       // (the twiddle table in __constant__ memory is per FCCH length: one length per device at a time)
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
	const size_t smem = sizeof(float) * (10 * ((a.len + 3) & ~3) + 8);
	fcch_fine_kernel<<<a.n, FF_T, smem, st>>>(a, mode);
	return cudaGetLastError();
}

}  // namespace gmr1
