// demod_kernels.cu - stage 2 of the receive path on the GPU: batched pi/4-CxPSK burst demodulation.
//
// One warp per burst.  The burst window (len*sps + search-window complex float samples, 4-11 KB
// at sps 4) is read from HBM exactly once into the warp's slice of shared memory; every later
// pass (statistics, training-sequence correlation over all search offsets, early/late peak
// search, frequency / phase estimation, soft bits) works out of shared memory and registers with
// warp-shuffle reductions.  Output per burst: ebits (int8), sync id, fractional TOA, frequency
// error, sync power.
//
// Replaces, for a whole batch per launch, the reference's
//   gmr1_pi4cxpsk_demod       src/sdr/pi4cxpsk.c:520-602
//   _gmr1_pi4cxpsk_sync_find  :184-268   (incl. the never-reset accumulator quirk, :207/:232)
//   _gmr1_pi4cxpsk_align      :280-348   (sps >= 4 path)
//   _gmr1_pi4cxpsk_freq_err   :360-406
//   _gmr1_pi4cxpsk_phase      :415-433
//   _gmr1_pi4cxpsk_soft_symbols / _soft_bits  :442-503
//   gmr1_pi4cxpsk_detect      :617-682
// and the libosmo-dsp primitives they call (sig_normalize, correlate, peak_energy_find with
// PEAK_EARLY_LATE, interpolate_point, rotate, scale) as restated in SURVEY.md Appendix A.2.
//
// The kernel computes the same quantities as the C path but not with the same instruction
// sequence; it is issue-bound, so the work is restructured to need ~3x fewer instructions:
//   * normalise + derotate is never applied to the 1000+ samples of the window.  Only magnitudes
//     of correlations are used, so the rotation moves onto the <= 32 reference taps and the
//     mean / scale become a per-chunk correction (see sync_find);
//   * window statistics are one pass (E|x|^2 - |E x|^2), tree sums;
//   * the sinc interpolation shares one sine per position: sin(pi*(k-pos)) = -(-1)^j sin(pi*frac);
//   * data symbols are sliced in the angle domain: arg(x-avg) + fs*idx - ferr*i - arg(phasor),
//     accumulated in double, instead of three complex rotations and an atan2 per symbol.
// Float contract (tests/test_demod_gpu.py): sync_id identical, TOA within 0.01 sample, freq_err
// within 2e-5 rad/symbol, soft bits within +-1 LSB (>= 99.5 % identical), and identical L2 / CRC
// after stage 3.  Integer stages (stage 3) are bit-exact.
#include <cuda_runtime.h>
#include <math.h>

#include "gmr1_tables.h"
#include "launch.h"

namespace gmr1 {

static constexpr int DM_WARPS = 4;
static constexpr float PI_F = 3.14159265358979323846264338327f;

// sin(pi * k / 512), k = 0..512: every position the early/late search visits is a multiple of
// 1/512 (start integer, steps 1/2 .. 1/512), so the one sine an interpolation needs is a lookup
__constant__ float c_sinpi512[513];

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// conj(ref) * g for ref in {1, j, -1, -j} (symbol index 0..3): exact component shuffles
__device__ __forceinline__ float2 mul_conj_sym(int sym, float2 g)
{
	const float a = (sym & 1) ? g.y : g.x, b = (sym & 1) ? -g.x : g.y;
	return (sym & 2) ? make_float2(-a, -b) : make_float2(a, b);
}

// atan2f with ~1e-7 rad absolute error: octant reduction + degree-8 minimax polynomial in a^2
__device__ __forceinline__ float fast_atan2f(float y, float x)
{
	const float ax = fabsf(x), ay = fabsf(y);
	const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
	const float a = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
	const float t = a * a;
	float p = 2.4464237603e-03f;
	p = fmaf(p, t, -1.4352691414e-02f);
	p = fmaf(p, t, 3.9685524385e-02f);
	p = fmaf(p, t, -7.2247349056e-02f);
	p = fmaf(p, t, 1.0492738066e-01f);
	p = fmaf(p, t, -1.4159015534e-01f);
	p = fmaf(p, t, 1.9985472684e-01f);
	p = fmaf(p, t, -3.3332556455e-01f);
	p = fmaf(p, t, 9.9999987390e-01f);
	float r = p * a;
	r = ay > ax ? (0.5f * PI_F - r) : r;
	r = x < 0.0f ? (PI_F - r) : r;
	return y < 0.0f ? -r : r;
}

// Sinc interpolation (osmo_cxvec_interpolate_point, 10 taps either side) of the real vector
// acc[0..len) at `pos` and, when LATE, also at `pos + 2` (the late gate).  One tap per lane,
// shuffle-tree sums.  The 21 sinc values of both gates are identical ((i+2)-(pos+2) == i-pos
// exactly in fp32 here) and share one sine: sin(pi*(j-frac)) = -(-1)^j sin(pi*frac).
template <bool LATE>
__device__ __forceinline__ void interp(const float *acc, int len, float pos, int lane, float &ev, float &lv)
{
	const float fl = floorf(pos);
	const int fe = (int)fl;
	const float frac = pos - fl;                      // exact, a multiple of 1/512
	const float S = c_sinpi512[(int)(frac * 512.0f)];
	float te = 0.0f, tl = 0.0f;
	if (lane < 21) {
		const int j = lane - 10, k = fe + j;
		const float x = PI_F * ((float)j - frac);
		float s = __fdividef((j & 1) ? S : -S, x);
		s = (x >= 0.01f || x <= -0.01f) ? s : 1.0f;   // osmo_sinc
		te = (k >= 0 && k < len) ? acc[k] * s : 0.0f;
		if (LATE)
			tl = (k + 2 >= 0 && k + 2 < len) ? acc[k + 2] * s : 0.0f;
	}
	if (LATE) {
		// fold both sums into one tree: after the first exchange the lower half-warp carries the
		// early terms, the upper half the late terms
		const bool up = lane & 16;
		float v = (up ? tl : te) + __shfl_xor_sync(0xffffffffu, up ? te : tl, 16);
#pragma unroll
		for (int o = 8; o; o >>= 1)
			v += __shfl_xor_sync(0xffffffffu, v, o);
		ev = __shfl_sync(0xffffffffu, v, 0);
		lv = __shfl_sync(0xffffffffu, v, 16);
	} else {
		ev = warp_sum(te);
	}
}

// osmo_cxvec_peak_energy_find(acc, 3, PEAK_EARLY_LATE, &peak) on a real vector; all lanes
// return the same position / peak value
__device__ float peak_early_late(const float *acc, int w, int lane, float &peak_val)
{
	const int win = w < 3 ? w : 3;
	float best = 0.0f;
	int best_idx = 0x7fffffff;
	for (int idx = lane; idx < w; idx += 32) {
		float val = 0.0f;
		for (int hi = idx - win + 1; hi <= idx; hi++)
			if (hi >= 0) {
				const float a = acc[hi];
				val += a * a;
			}
		if (val > best) {
			best = val;
			best_idx = idx;
		}
	}
#pragma unroll
	for (int o = 16; o; o >>= 1) {
		const float ov = __shfl_xor_sync(0xffffffffu, best, o);
		const int oi = __shfl_xor_sync(0xffffffffu, best_idx, o);
		if (ov > best || (ov == best && oi < best_idx)) {
			best = ov;
			best_idx = oi;
		}
	}
	int max_idx = (best > 0.0f) ? best_idx - win + 1 : 0;
	if (max_idx < 0)
		max_idx = 0;

	int mwi = max_idx;
	float mv = -1.0f;
	for (int idx = max_idx; idx < max_idx + win; idx++) {
		const float a = acc[idx], e = a * a;
		if (e > mv) {
			mv = e;
			mwi = idx;
		}
	}

	float early = (float)(mwi - 1), incr = 0.5f;
#pragma unroll 1
	while (incr > (1.0f / 1024.0f)) {
		float ev, lv;
		interp<true>(acc, w, early, lane, ev, lv);
		const float e2 = ev * ev, l2 = lv * lv;
		if (e2 < l2)
			early += incr;
		else if (e2 > l2)
			early -= incr;
		else
			break;
		incr *= 0.5f;
	}
	const float pos = early + 1.0f;
	float dummy;
	interp<false>(acc, w, pos, lane, peak_val, dummy);
	return pos;
}

// per-warp shared-memory slice
struct WarpSmem {
	float2 *win;     // [L]   raw window (never rewritten)
	float2 *taps;    // [32]  rotated reference taps of the chunk being correlated
	float  *accv;    // [w]   correlation magnitude accumulator
};

__device__ __forceinline__ WarpSmem carve(uint8_t *base, int L, int w)
{
	WarpSmem s;
	s.win = (float2 *)base;
	base += (size_t)((L + 1) & ~1) * 8;
	s.taps = (float2 *)base;
	base += 32 * 8;
	s.accv = (float *)base;
	return s;
}

static inline size_t warp_smem_bytes(int L, int w)
{
	return (size_t)((L + 1) & ~1) * 8 + 32 * 8 + (size_t)((w + 3) & ~3) * 4;
}

// window statistics of osmo_cxvec_sig_normalize: mean and 1/stddev.  One pass: the variance is
// E|x|^2 - |E x|^2 with per-lane partial sums and a shuffle tree (the C path sums sequentially;
// both are fp32 approximations of the same quantity).
struct Norm { float ar, ai, inv_sd; };

__device__ __forceinline__ Norm load_stats(const float2 *__restrict__ x, int L, float2 *win, int lane, bool want_sd)
{
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	if ((((uintptr_t)x) & 15) == 0) {
		// 16-byte aligned window: two samples per lane per load
		const float4 *x4 = reinterpret_cast<const float4 *>(x);
		float4 *w4 = reinterpret_cast<float4 *>(win);
		const int L2 = L >> 1;
#pragma unroll 4
		for (int i = lane; i < L2; i += 32) {
			const float4 v = __ldg(&x4[i]);
			w4[i] = v;
			sr += v.x + v.z;
			si += v.y + v.w;
			if (want_sd) {
				sq = fmaf(v.x, v.x, sq);
				sq = fmaf(v.y, v.y, sq);
				sq = fmaf(v.z, v.z, sq);
				sq = fmaf(v.w, v.w, sq);
			}
		}
		if ((L & 1) && lane == 0) {
			const float2 v = __ldg(&x[L - 1]);
			win[L - 1] = v;
			sr += v.x;
			si += v.y;
			sq += v.x * v.x + v.y * v.y;
		}
	} else {
#pragma unroll 4
		for (int i = lane; i < L; i += 32) {
			const float2 v = __ldg(&x[i]);
			win[i] = v;
			sr += v.x;
			si += v.y;
			sq += v.x * v.x + v.y * v.y;
		}
	}
	sr = warp_sum(sr);
	si = warp_sum(si);
	Norm n;
	n.ar = sr / (float)L;
	n.ai = si / (float)L;
	n.inv_sd = 1.0f;
	if (want_sd) {
		// The scale 1/stddev changes no decision and no soft bit (peak positions, phases and
		// angles are scale invariant); it only sets the absolute value of the reported sync
		// power, so the extra pass is skipped unless that output is requested.
		sq = warp_sum(sq);
		const float var = sq / (float)L - (n.ar * n.ar + n.ai * n.ai);
		float sd = var > 0.0f ? sqrtf(var) : 0.0f;
		if (sd == 0.0f)
			sd = 1.0f;
		n.inv_sd = 1.0f / sd;
	}
	__syncwarp();
	return n;
}

// Search all sync sequences of one burst type (pi4cxpsk.c:184-268) on the RAW window.
//
// The reference normalises and derotates every sample, y[i] = (x[i]-avg)/sd * e^{j*fs*i}, then
// correlates y with the +-1/+-j training symbols and keeps |corr|.  Since only the magnitude is
// used, the common factor e^{j*fs*(b0+m)} drops out and the rotation moves onto the <= 32
// reference taps:  |corr[m]| = |sum_n t_n x[b0+m+n*sps] - avg*sum_n t_n| / sd,
// t_n = conj(ref_n) e^{j*fs*sps*n}.  Same quantity, 60x fewer sincos.
// accv is NOT cleared between sequences - the reference clears it once per call (:207) and
// keeps adding (:232-233); tl restarts per sequence (:216).
__device__ int sync_find(const BurstTab &bt, const WarpSmem &sm, const Norm &nm, float fs, int sps, int w,
                         int lane, float &toa, float &pwr)
{
	for (int m = lane; m < w; m += 32)
		sm.accv[m] = 0.0f;
	float p_toa = 0.0f, p_pwr = 0.0f;
	int p_idx = -1;
	float rot_s, rot_c;                               // e^{j*fs*sps*lane}: tap n of every chunk
	sincosf((fs * (float)sps) * (float)lane, &rot_s, &rot_c);
	for (int s = 0; s < bt.n_sync; s++) {
		int tl = 0;
		for (int c = 0; c < bt.n_chunk[s]; c++) {
			const int b0 = bt.s_pos[s][c] * sps, cl = bt.s_len[s][c];
			// rotated taps + their sum
			float tr = 0.0f, ti = 0.0f;
			if (lane < cl) {
				const float2 t = mul_conj_sym(bt.s_sym[s][c][lane], make_float2(rot_c, rot_s));
				tr = t.x;
				ti = t.y;
			}
			__syncwarp();
			sm.taps[lane] = make_float2(tr, ti);
			const float Rr = warp_sum(tr), Ri = warp_sum(ti);
			const float cr0 = nm.ar * Rr - nm.ai * Ri, ci0 = nm.ar * Ri + nm.ai * Rr;   // avg * sum(taps)
			__syncwarp();
			for (int m = lane; m < w; m += 32) {
				float cr = 0.0f, ci = 0.0f;
				const float2 *g = sm.win + b0 + m;
				const float2 *tp = sm.taps;
#pragma unroll 4
				for (int n = 0; n < cl; n++, g += sps, tp++) {
					const float2 t = *tp, v = *g;
					cr = fmaf(t.x, v.x, cr);
					cr = fmaf(-t.y, v.y, cr);
					ci = fmaf(t.x, v.y, ci);
					ci = fmaf(t.y, v.x, ci);
				}
				cr = (cr - cr0) * nm.inv_sd;
				ci = (ci - ci0) * nm.inv_sd;
				sm.accv[m] += sqrtf(cr * cr + ci * ci);
			}
			tl += cl;
		}
		__syncwarp();
		float peak;
		const float s_toa = peak_early_late(sm.accv, w, lane, peak);
		peak /= (float)tl;
		const float s_pwr = peak * peak;
		if (s_pwr > p_pwr) {
			p_pwr = s_pwr;
			p_toa = s_toa;
			p_idx = s;
		}
	}
	toa = p_toa;
	pwr = p_pwr;
	return p_idx;
}

// ---- kernel ---------------------------------------------------------------------------------------
// mode 0: demod (bts[0] only).  mode 1: detect among n_bt burst types (pi4cxpsk.c:617-682).
__global__ void __launch_bounds__(DM_WARPS * 32)
demod_kernel(const DemodArgs a, const BurstTab *__restrict__ bts, int n_bt, int mode, int warp_bytes)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int b = blockIdx.x * DM_WARPS + warp;
	if (b >= a.n)
		return;

	const BurstTab &bt = bts[0];
	const int sps = a.sps, L = a.win_len;
	const int w = L - bt.len * sps + 1;
	const WarpSmem sm = carve(smem + (size_t)warp * warp_bytes, L, w);

	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;
	const float fs = (freq_shift - bt.rotation) / (float)sps;

	const Norm nm = load_stats(x, L, sm.win, lane, a.pwr != nullptr || (mode == 1 && (a.e_toa != nullptr || a.e_toa0 >= 0.0f)));

	if (mode == 1) {
		const float e_toa = a.e_toa ? a.e_toa[b] : a.e_toa0;
		int p_id = -1, p_sid = -1;
		float p_toa = 0.0f, p_pwr = 0.0f;
		for (int id = 0; id < n_bt; id++) {
			float toa, pwr;
			const int sid = sync_find(bts[id], sm, nm, fs, sps, w, lane, toa, pwr);
			if (e_toa >= 0.0f)     // the reference divides by fabs() in double (pi4cxpsk.c:658-659)
				pwr = (float)((double)pwr / fabs((double)(e_toa - toa)));
			if (pwr > p_pwr) {
				p_id = id;
				p_sid = sid;
				p_pwr = pwr;
				p_toa = toa;
			}
		}
		if (lane == 0) {
			if (a.bt_id) a.bt_id[b] = p_id;
			if (a.sync_id) a.sync_id[b] = p_sid;
			if (a.toa) a.toa[b] = p_toa;
			if (a.pwr) a.pwr[b] = p_pwr;
		}
		return;
	}

	float toa, pwr;
	const int sync_id = sync_find(bt, sm, nm, fs, sps, w, lane, toa, pwr);
	if (lane == 0) {
		if (a.sync_id) a.sync_id[b] = sync_id;
		if (a.toa) a.toa[b] = toa;
		if (a.pwr) a.pwr[b] = pwr;
	}
	int8_t *eb = a.ebits + (size_t)b * a.ebits_stride;
	if (sync_id < 0) {          // nothing correlated (all-zero input): the reference returns -errno
		if (lane == 0 && a.freq_err) a.freq_err[b] = 0.0f;
		for (int k = lane; k < bt.ebits; k += 32)
			eb[k] = 0;
		return;
	}

	// symbol i sits at sample i*sps + d (sps >= 4 path of _gmr1_pi4cxpsk_align, :286-297)
	const int d = (int)roundf(toa);
	auto sample_of = [&](int i) {
		const int q = i * sps + d;     // d >= -1; index -1 would read before the window: clamp
		return min(max(q, 0), L - 1);
	};

	// ---- training symbols (<= 32 per chunk, one per lane), derotated as the reference derotates
	//      every sample: z = (x - avg)/sd * e^{j*fl32(fs*idx)}.  Per chunk: correlation sum ->
	//      fine frequency error from the chunk-to-chunk phase slope (:360-406).  The derotated
	//      products are kept in registers (chunk c -> zr[c], zi[c]) for the phase reference.
	const int nch = bt.n_chunk[sync_id];
	float zr[MAX_SYNC_CHUNK], zi[MAX_SYNC_CHUNK];
	float ferr = 0.0f, f = 0.0f, prev_r = 0.0f, prev_i = 0.0f, prev_pos = 0.0f;
#pragma unroll
	for (int c = 0; c < MAX_SYNC_CHUNK; c++) {
		zr[c] = zi[c] = 0.0f;
		if (c < nch) {
			const int p0 = bt.s_pos[sync_id][c], cl = bt.s_len[sync_id][c];
			if (lane < cl) {
				const int q = sample_of(p0 + lane);
				const float2 v = sm.win[q];
				float sn, cs;
				sincosf(fs * (float)q, &sn, &cs);
				const float yr = (v.x - nm.ar) * nm.inv_sd, yi = (v.y - nm.ai) * nm.inv_sd;
				const float2 p = mul_conj_sym(bt.s_sym[sync_id][c][lane],
				                              make_float2(yr * cs - yi * sn, yr * sn + yi * cs));
				zr[c] = p.x;
				zi[c] = p.y;
			}
			if (nch > 1) {
				const float cr = warp_sum(zr[c]), ci = warp_sum(zi[c]);
				const float pos = (float)p0 + (float)cl / 2.0f;
				if (c > 0) {   // arg(corr[c] * conj(corr[c-1])) / (pos[c] - pos[c-1])
					const float re = cr * prev_r + ci * prev_i, im = ci * prev_r - cr * prev_i;
					f += fast_atan2f(im, re) / (pos - prev_pos);
				}
				prev_r = cr;
				prev_i = ci;
				prev_pos = pos;
			}
		}
	}
	if (nch > 1)
		ferr = f / (float)(nch - 1);
	if (lane == 0 && a.freq_err) a.freq_err[b] = ferr;

	// ---- phase reference: all training symbols after the -ferr rotation (:415-433, :574)
	float phi0;
	{
		float pr = 0.0f, pi = 0.0f;
#pragma unroll
		for (int c = 0; c < MAX_SYNC_CHUNK; c++) {
			if (c < nch) {
				float r = zr[c], i_ = zi[c];
				if (ferr != 0.0f) {
					float sn, cs;
					sincosf((-ferr) * (float)(bt.s_pos[sync_id][c] + lane), &sn, &cs);
					const float t = r * cs - i_ * sn;
					i_ = r * sn + i_ * cs;
					r = t;
				}
				pr += r;
				pi += i_;
			}
		}
		pr = warp_sum(pr);
		pi = warp_sum(pi);
		phi0 = fast_atan2f(pi, pr);
	}

	// ---- data symbols in the angle domain.  The reference rotates each sample three times
	// (e^{j*fs*idx}, e^{-j*ferr*i}, conj(phasor)) and takes cargf(); the argument of that product
	// is  arg(x - avg) + fl32(fs*idx) + fl32(-ferr*i) - arg(phasor)  (mod 2*pi), accumulated here
	// in double so that the only float rounding left is the atan2 of the raw sample.
	const int nbits = bt.nbits, mask = (1 << nbits) - 1;
	const double inv_dd = (double)(1 << nbits) / (2.0 * 3.14159265358979323846);
	const double period = (double)(1 << nbits), inv_period = 1.0 / period;
	const double c0 = -(double)phi0 * inv_dd;
	int kbase = 0;
	for (int c = 0; c < bt.n_data; c++) {
		const int p0 = bt.d_pos[c], cl = bt.d_len[c];
		for (int j = lane; j < cl; j += 32) {
			const int i = p0 + j, q = sample_of(i);
			const float2 v = sm.win[q];
			const float th = fast_atan2f(v.y - nm.ai, v.x - nm.ar);
			const float a1 = fs * (float)q;
			const float a2 = ferr != 0.0f ? (-ferr) * (float)i : 0.0f;
			double svd = fma((double)th + (double)a1 + (double)a2, inv_dd, c0);
			svd -= period * rint(svd * inv_period);        // -> [-period/2, period/2]
			const float sv = (float)svd;
			const float svr = roundf(sv);
			const int sp = (int)svr & mask;
			const int ss = (svr > sv ? (sp - 1) : (sp + 1)) & mask;
			const int dq = (int)roundf((2.0f * fabsf(svr - sv)) * 64.0f);
			if (nbits == 2) {
				// Gray map of the symbol index {00, 01, 11, 10}, MSB first
				const int gp = sp ^ (sp >> 1), gx = gp ^ ss ^ (ss >> 1);
				const int v0 = 127 - ((gx & 2) ? dq : (dq >> 1)), v1 = 127 - ((gx & 1) ? dq : (dq >> 1));
				const int b0 = (gp & 2) ? -v0 : v0, b1 = (gp & 1) ? -v1 : v1;
				int8_t *o = eb + kbase + 2 * j;
				if ((((uintptr_t)o) & 1) == 0)
					*reinterpret_cast<uint16_t *>(o) = (uint16_t)((b0 & 0xff) | ((b1 & 0xff) << 8));
				else {
					o[0] = (int8_t)b0;
					o[1] = (int8_t)b1;
				}
			} else {
				const int v0 = 127 - (((sp ^ ss) & 1) ? dq : (dq >> 1));
				eb[kbase + j] = (int8_t)(sp ? -v0 : v0);
			}
		}
		kbase += cl * nbits;
	}
}

// ---- launcher --------------------------------------------------------------------------------------
cudaError_t launch_demod(const DemodArgs &a, const BurstTab *d_bts, const BurstTab *h_bts, int n_bt, int mode,
                         cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	int maxlen = 0, minlen = 1 << 30;
	for (int i = 0; i < n_bt; i++) {
		maxlen = h_bts[i].len > maxlen ? h_bts[i].len : maxlen;
		minlen = h_bts[i].len < minlen ? h_bts[i].len : minlen;
	}
	if (maxlen != minlen)
		return cudaErrorInvalidValue;      // detect needs length-compatible burst types
	const int w = a.win_len - maxlen * a.sps + 1;
	if (w < 1 || a.sps < 4 || a.sps > 16)
		return cudaErrorInvalidValue;
	const size_t wb = (warp_smem_bytes(a.win_len, w) + 15) & ~(size_t)15;
	const size_t smem = wb * DM_WARPS;
	if (smem > 227 * 1024)
		return cudaErrorInvalidValue;
	static size_t attr_set[64] = {0};
	static bool tab_up[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 64 || !tab_up[dev]) {
		float h[513];
		for (int k = 0; k <= 512; k++)
			h[k] = (float)sin(3.14159265358979323846 * (double)k / 512.0);
		cudaError_t e = cudaMemcpyToSymbol(c_sinpi512, h, sizeof(h));
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			tab_up[dev] = true;
	}
	if (dev >= 64 || attr_set[dev] < smem) {
		cudaError_t e = cudaFuncSetAttribute(demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_set[dev] = smem;
	}
	demod_kernel<<<(a.n + DM_WARPS - 1) / DM_WARPS, DM_WARPS * 32, smem, st>>>(a, d_bts, n_bt, mode, (int)wb);
	return cudaGetLastError();
}

}  // namespace gmr1
