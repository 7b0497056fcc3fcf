// demod_kernels.cu - stage 2 of the receive path on the GPU: batched pi/4-CxPSK burst demodulation.
//
// One warp per burst, persistent warps (148 SMs x 8 CTAs x 4 warps stride over the batch).  The burst window
// (len*sps + search-window complex float samples, 4-11 KB at sps 4) is requested into L2 with one bulk prefetch and
// read from HBM exactly once (16-byte loads of the statistics pass).  Only the samples the training-sequence search
// touches ("correlation regions", ~1/3 of the window) are kept in the warp's ~5 KB slice of shared memory, filled
// from those same loads; the data-symbol samples are picked from L2 afterwards.  Statistics, correlation over all
// search offsets, early/late peak search, frequency / phase estimation and soft bits run on registers, shared
// memory and warp shuffles.  Output per burst: ebits (int8), sync id, fractional TOA, frequency error, sync power.
//
// Replaces, for a whole batch per launch, the reference's
//   gmr1_pi4cxpsk_demod       src/sdr/pi4cxpsk.c:520-602
//   _gmr1_pi4cxpsk_sync_find  :184-268   (incl. the never-reset accumulator quirk, :207/:232)
//   _gmr1_pi4cxpsk_align      :280-348
//   _gmr1_pi4cxpsk_freq_err   :360-406
//   _gmr1_pi4cxpsk_phase      :415-433
//   _gmr1_pi4cxpsk_soft_symbols / _soft_bits  :442-503
//   gmr1_pi4cxpsk_detect      :617-682
// and the libosmo-dsp primitives they call (sig_normalize, correlate, peak_energy_find with
// PEAK_EARLY_LATE, interpolate_point, rotate, scale) as restated in SURVEY.md Appendix A.2.
//
// The kernel computes the same quantities as the C path but not with the same instruction sequence; it is bound
// by instruction issue, so the work is restructured to need ~4x fewer instructions:
//   * normalise + derotate is never applied to the 1000+ samples of the window.  Only magnitudes
//     of correlations are used, so the rotation moves onto the <= 32 reference taps and the
//     mean / scale become a per-chunk correction (corr_block);
//   * window statistics are one pass, tree sums;
//   * the early/late search compares two sinc interpolations that share their fractional position (no sine,
//     one reciprocal per tap) and takes three of its bisection steps per round (8 grid points in parallel);
//   * data symbols are sliced in the angle domain: arg(x-avg) + fs*idx - ferr*i - arg(phasor) in fp32 with a
//     Cody-Waite reduction of fs*idx, instead of three complex rotations and an atan2 per symbol; the soft bits
//     come from a table over the symbol value in steps of 1/256.
// LAYOUT RULE: the per-burst loop must fit the 32 KB instruction cache.  Variants that a launch does not execute
// (sync power, unaligned windows, > 96 search offsets, > 32 training symbols) live in __noinline__ functions or in
// other template instantiations (ROWS, SYMB), never inline in the hot path (profiles/README.md has the numbers).
// Float contract (tests/test_demod_gpu.py): sync_id identical, TOA within 0.01 sample, freq_err
// within 2e-5 rad/symbol, soft bits within +-1 LSB (>= 99.5 % identical), and identical L2 / CRC
// after stage 3.  Integer stages (stage 3) are bit-exact.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "gmr1_tables.h"
#include "launch.h"


#include "demod_common.cuh"

namespace gmr1 {

// Correlation regions: the only samples that are read more than once are those the training-
// sequence search touches (chunk position .. + (len-1)*sps + search offsets).  They are the union of
// a few short intervals (~300 of the 1016 samples of a BCCH window) and are the only part of the
// window kept in shared memory, which is what lets 32 warps share an SM.  Symbol samples are read
// once each, straight from global memory (L2 hits: the warp streamed the window moments before).
static constexpr int MAX_REGIONS = 8;
static constexpr int MAX_BT = 8;            // burst types one detect call may choose from
struct Regions {
	int32_t n, total;                       // number of regions, samples in all of them
	int32_t start[MAX_REGIONS], len[MAX_REGIONS], off[MAX_REGIONS];   // window sample, length, smem offset
	int32_t n_slot;                         // training chunks of all types / sequences, numbered densely:
	uint8_t slot[MAX_BT][MAX_SYNC][MAX_SYNC_CHUNK];   // their rotated taps are cached per warp (build_taps)
	// what the per-burst code needs of the burst descriptors, in the constant bank (uniform loads, no LSU)
	uint8_t  n_sync[MAX_BT], n_chunk[MAX_BT][MAX_SYNC];
	uint8_t  cl[MAX_BT][MAX_SYNC][MAX_SYNC_CHUNK];     // chunk length in symbols
	uint16_t roff[MAX_BT][MAX_SYNC][MAX_SYNC_CHUNK];   // offset (samples) of the chunk's first sample in the region buffer
	float    cpos[MAX_SYNC][MAX_SYNC_CHUNK];           // type 0: centre of the chunk in symbols (s_pos + s_len / 2)
	float    rotation0;                                // type 0: per-symbol rotation
};

// per-warp shared-memory slice
struct WarpSmem {
	float2 *reg;     // [regions.total] raw samples of the correlation regions
	float2 *taps;    // [n_slot][32]  rotated reference taps of every training chunk (zero beyond its length)
	float2 *tsum;    // [n_slot]      sum of each chunk's taps
	float  *accv;    // [w]   correlation magnitude accumulator
	float2 *zbuf;    // [MAX_TRAIN] derotated training symbols x conj(reference)
};

static constexpr int MAX_TRAIN = 104;     // RACH: 17 + 32 + 32 + 17 + 1 = 99 training symbols

__device__ __forceinline__ WarpSmem carve(uint8_t *base, int nreg, int w, int nslot)
{
	WarpSmem s;
	s.reg = (float2 *)base;
	base += (size_t)((nreg + 1) & ~1) * 8;
	s.taps = (float2 *)base;
	base += (size_t)nslot * 32 * 8;
	s.tsum = (float2 *)base;
	base += (size_t)((nslot + 1) & ~1) * 8;
	s.accv = (float *)base + 4;            // accv[-4..-1] and everything from accv[w] on stay zero (peak search padding)
	base += (size_t)(((w + 31) & ~31) + 8) * 4;
	s.zbuf = (float2 *)base;
	return s;
}

static inline size_t warp_smem_bytes(int nreg, int w, int nslot)
{
	return (size_t)((nreg + 1) & ~1) * 8 + (size_t)nslot * 32 * 8 + (size_t)((nslot + 1) & ~1) * 8 +
	       (size_t)(((w + 31) & ~31) + 8) * 4 + MAX_TRAIN * 8;
}

// window statistics of osmo_cxvec_sig_normalize: mean and 1/stddev.  One pass: the variance is
// E|x|^2 - |E x|^2 with per-lane partial sums and a shuffle tree (the C path sums sequentially;
// both are fp32 approximations of the same quantity).
struct Norm { float ar, ai, inv_sd; };

// dst4 (optional): for every pair of samples of the window, the float4 slot of the warp's region buffer
// it belongs to (0xffff: none) - the correlation regions are then filled from the same loads and the
// separate load_regions pass (a second trip to L2) is not needed.
template <bool WANT_SD, bool ALIGNED>       // ALIGNED: the caller has checked the 16-byte alignment (hot path)
__device__ __forceinline__ Norm load_stats_t(const float2 *__restrict__ x, int L, int lane,
                                             const bool FILL, const uint16_t *dst4, float2 *reg)
{
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	float2 s2 = make_float2(0.0f, 0.0f);
	if (ALIGNED || (((uintptr_t)x) & 15) == 0) {
		// 16-byte aligned window: two samples per lane per load
		const float4 *x4 = reinterpret_cast<const float4 *>(x);
		float4 *reg4 = reinterpret_cast<float4 *>(reg);
		const int L2 = L >> 1;
DM_UNROLL(DM_STATS_UNROLL)
		for (int i = lane; i < L2; i += 32) {
			const float4 v = __ldg(&x4[i]);
			if (FILL) {
				const unsigned d = dst4[i];
				if (d != 0xffffu)
					reg4[d] = v;
			}
			s2 = __fadd2_rn(s2, __fadd2_rn(make_float2(v.x, v.y), make_float2(v.z, v.w)));     // two FADD2
			if (WANT_SD) {
				sq = fmaf(v.x, v.x, sq);
				sq = fmaf(v.y, v.y, sq);
				sq = fmaf(v.z, v.z, sq);
				sq = fmaf(v.w, v.w, sq);
			}
		}
		if ((L & 1) && lane == 0) {
			const float2 v = __ldg(&x[L - 1]);
			if (FILL) {
				const unsigned d = dst4[L2];
				if (d != 0xffffu)
					reg[2 * d] = v;
			}
			sr += v.x;
			si += v.y;
			if (WANT_SD)
				sq += v.x * v.x + v.y * v.y;
		}
	} else {
#pragma unroll 1
		for (int i = lane; i < L; i += 32) {        // 8-byte aligned window (odd sample offset): cold path
			const float2 v = __ldg(&x[i]);
			sr += v.x;
			si += v.y;
			if (WANT_SD)
				sq += v.x * v.x + v.y * v.y;
		}
	}
	sr = warp_sum(sr + s2.x);
	si = warp_sum(si + s2.y);
	Norm n;
	if (ALIGNED) {          // hot path: one reciprocal (the IEEE division is a 12-instruction sequence, twice per burst)
		const float inv_l = __frcp_rn((float)L);
		n.ar = sr * inv_l;
		n.ai = si * inv_l;
	} else {
		n.ar = sr / (float)L;
		n.ai = si / (float)L;
	}
	n.inv_sd = 1.0f;
	if (WANT_SD) {
		// The scale 1/stddev changes no decision and no soft bit (peak positions, phases and
		// angles are scale invariant); it only sets the absolute value of the reported sync
		// power, so the extra work is skipped unless that output is requested.
		sq = warp_sum(sq);
		const float var = sq / (float)L - (n.ar * n.ar + n.ai * n.ai);
		float sd = var > 0.0f ? sqrtf(var) : 0.0f;
		if (sd == 0.0f)
			sd = 1.0f;
		n.inv_sd = 1.0f / sd;
	}
	return n;
}

// copy the correlation regions of this window into the warp's shared memory
// (fallback for windows that are not 16-byte aligned or too long for the pair table; the usual path
// fills the regions from the loads of the statistics pass)
__device__ __noinline__ void load_regions(const float2 *__restrict__ x, int L, const Regions &rg, float2 *reg, int lane)
{
	for (int r = 0; r < rg.n; r++) {
		const float2 *src = x + rg.start[r];
		float2 *dst = reg + rg.off[r];
		const int n = min(rg.len[r], L - rg.start[r]);      // lengths are rounded up to even
#pragma unroll 1
		for (int i = lane; i < n; i += 32)
			dst[i] = __ldg(&src[i]);
	}
	__syncwarp();
}

// everything but the usual case (16-byte aligned window, regions filled on the fly, no sync power wanted),
// out of line: the per-burst loop is instruction-cache sensitive and must stay compact
__device__ __noinline__ Norm load_stats_cold(const float2 *__restrict__ x, int L, int lane, bool want_sd, bool fill,
                                             const uint16_t *dst4, const Regions &rg, float2 *reg)
{
	const Norm n = want_sd ? load_stats_t<true, false>(x, L, lane, fill, dst4, reg)
	                       : load_stats_t<false, false>(x, L, lane, fill, dst4, reg);
	if (!fill)
		load_regions(x, L, rg, reg, lane);
	return n;
}

// complex multiply-accumulate of the (zero-padded to a multiple of 4) rotated taps against R search
// offsets 32 samples apart; one tap load serves all R.  Accumulates P = sum t.re * x and Q = sum t.im * x
// (two FFMA2 per tap and offset); the correlation is (P.re - Q.im, P.im + Q.re).
template <int R>
__device__ __forceinline__ void corr_taps(const float2 *g, const float2 *tp, int cl4, int sps,
                                          float2 (&P)[3], float2 (&Q)[3])
{
#pragma unroll 1
	for (int n = 0; n < cl4; n += 4, tp += 4, g += 4 * sps) {
		const float4 t01 = *reinterpret_cast<const float4 *>(tp), t23 = *reinterpret_cast<const float4 *>(tp + 2);
		const float tr[4] = {t01.x, t01.z, t23.x, t23.z}, ti[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
		for (int u = 0; u < 4; u++) {
#pragma unroll
			for (int r = 0; r < R; r++) {
				const float2 v = g[u * sps + 32 * r];
				fma2s(P[r], tr[u], v);
				fma2s(Q[r], ti[u], v);
			}
		}
	}
}

// Rotated reference taps of every training chunk for the frequency shift fs: t_n = conj(ref_n) e^{j*fs*sps*n},
// and their sums.  Depends on the burst only through fs, so each warp rebuilds them only when fs changes
// (never, when the batch shares one freq_shift).
__device__ void build_taps(const BurstTab *__restrict__ bts, int n_bt, const Regions &rg, const WarpSmem &sm,
                           float fs, int sps, int lane)
{
	const float2 rot = sincos_acc((fs * (float)sps) * (float)lane);     // e^{j*fs*sps*lane}: tap n of every chunk
	__syncwarp();
	for (int ty = 0; ty < n_bt; ty++)
		for (int s = 0; s < bts[ty].n_sync; s++)
			for (int c = 0; c < bts[ty].n_chunk[s]; c++) {
				const int slot = rg.slot[ty][s][c];
				float2 t = make_float2(0.0f, 0.0f);
				if (lane < bts[ty].s_len[s][c])
					t = mul_conj_sym(bts[ty].s_sym[s][c][lane], rot);
				sm.taps[slot * 32 + lane] = t;
				const float2 Rs = warp_sum2(t.x, t.y, lane);
				if (lane == 0)
					sm.tsum[slot] = Rs;
			}
	__syncwarp();
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
// one block of up to 3 rows of 32 search offsets (m0, m0+32, m0+64), R of them in use: |corr| of every chunk of
// sequence s added onto the accumulator rows
template <int R, int SPS>
__device__ __forceinline__ int corr_block(const Regions &rg, int id, int s, const WarpSmem &sm, const Norm &nm, int sps_rt,
                                          int w, int m0, bool fresh)
{
	const int sps = SPS > 0 ? SPS : sps_rt;
	const int n_chunk = rg.n_chunk[id][s];
	float acc[R];
#pragma unroll
	for (int r = 0; r < R; r++)
		acc[r] = (fresh || m0 + 32 * r >= w) ? 0.0f : sm.accv[m0 + 32 * r];
	int tl = 0;
#pragma unroll 1
	for (int c = 0; c < n_chunk; c++) {
		const int cl = rg.cl[id][s][c];
		const int slot = rg.slot[id][s][c];
		const float2 Rs = sm.tsum[slot];
		const float cr0 = nm.ar * Rs.x - nm.ai * Rs.y, ci0 = nm.ar * Rs.y + nm.ai * Rs.x;   // avg * sum(taps)
		// taps beyond cl are zero, so the tap loop runs in whole groups of 4.  The up to 3 zero taps
		// may read up to 3*sps samples past their region: that lands in this warp's other regions /
		// taps / accv / zbuf, which only ever hold finite floats, and 0 * finite adds nothing.
		const int cl4 = (cl + 3) & ~3;
		float2 P[3], Q[3];
#pragma unroll
		for (int r = 0; r < 3; r++)
			P[r] = Q[r] = make_float2(0.0f, 0.0f);
		corr_taps<R>(sm.reg + rg.roff[id][s][c] + m0, sm.taps + slot * 32, cl4, sps, P, Q);
#pragma unroll
		for (int r = 0; r < R; r++) {
			// |corr| / sd: the scale is applied to the magnitude (one multiply instead of two), the square
			// root is the approximate one (MUFU.SQRT, 0 -> 0)
			const float xr = (P[r].x - Q[r].y) - cr0, xi = (P[r].y + Q[r].x) - ci0;
			float mag;
			asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(xr, xr, xi * xi)));
			acc[r] = fmaf(mag, nm.inv_sd, acc[r]);
		}
		tl += cl;
	}
#pragma unroll
	for (int r = 0; r < R; r++)
		if (m0 + 32 * r < w)            // entries from w on stay zero
			sm.accv[m0 + 32 * r] = acc[r];
	return tl;
}

// any number of search offsets, 96 per pass (cold: the formats of the reference have w <= 81 at sps 4)
template <int SPS>
__device__ __noinline__ int corr_generic(const Regions &rg, int id, int s, const WarpSmem &sm, const Norm &nm, int sps_rt,
                                         int w, int lane, bool fresh)
{
	int tl = 0;
#pragma unroll 1
	for (int mb = 0; mb < w; mb += 96) {
		if (mb + 64 < w)
			tl = corr_block<3, SPS>(rg, id, s, sm, nm, sps_rt, w, mb + lane, fresh);
		else if (mb + 32 < w)
			tl = corr_block<2, SPS>(rg, id, s, sm, nm, sps_rt, w, mb + lane, fresh);
		else
			tl = corr_block<1, SPS>(rg, id, s, sm, nm, sps_rt, w, mb + lane, fresh);
	}
	return tl;
}

// SPS > 0: compile-time samples per symbol (4 is the fast path), <= 0: run-time.
// ROWS > 0: the search has at most 32 * ROWS offsets and exactly ROWS rows are computed (compile-time: the hot
// path then holds one copy of the correlation loop - the kernel is instruction-cache sensitive); 0: any w.
template <int SPS, int ROWS>
__device__ __forceinline__ int sync_find(const Regions &rg, int id, const WarpSmem &sm, const Norm &nm, int sps_rt, int w,
                         const TapLane &tpl, int lane, bool sync_reset, float &toa, float &pwr)
{
	float p_toa = 0.0f, p_pwr = 0.0f;
	int p_idx = -1;
	const int n_sync = rg.n_sync[id];
#pragma unroll 1
	for (int s = 0; s < n_sync; s++) {
		const bool fresh = s == 0 || sync_reset;      // sync_reset (opt-in): score every candidate on its own correlation
		__syncwarp();
		// up to three search offsets per lane (m, m+32, m+64) share every tap load; their |corr| sums over
		// the chunks stay in registers.  Offsets >= w are computed too (the samples they read are whatever
		// finite values follow the region) and never stored.
		int tl;
		if (ROWS > 0)
			tl = corr_block<(ROWS > 0 ? ROWS : 1), SPS>(rg, id, s, sm, nm, sps_rt, w, lane, fresh);
		else
			tl = corr_generic<SPS>(rg, id, s, sm, nm, sps_rt, w, lane, fresh);
		__syncwarp();
		float peak;
		const float s_toa = peak_early_late<ROWS>(sm.accv, reinterpret_cast<float *>(sm.zbuf), w, tpl, lane, peak);
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

// sps < 4 (pi4cxpsk.c:298-343): symbols are not picked but interpolated.  y[k] = normalised, derotated
// sample k (0 outside the window, as osmo_cxvec_convolve treats it); with a fractional offset of more
// than 0.1 sample the symbol is the 21-tap sinc interpolation sum_j sinc(pi*((j-10)+frac)) * y[q+10-j].
__device__ __forceinline__ float2 norm_rot_sample(const float2 *__restrict__ x, int L, int k, const Norm &nm, float fs)
{
	if (k < 0 || k >= L)
		return make_float2(0.0f, 0.0f);
	const float2 v = __ldg(&x[k]);
	const float2 e = sincos_acc(fs * (float)k);
	const float yr = (v.x - nm.ar) * nm.inv_sd, yi = (v.y - nm.ai) * nm.inv_sd;
	return make_float2(yr * e.x - yi * e.y, yr * e.y + yi * e.x);
}

__device__ float2 lowsps_symbol(const float2 *__restrict__ x, int L, int q, bool interp, float frac, const Norm &nm, float fs)
{
	if (!interp)
		return norm_rot_sample(x, L, q, nm, fs);
	float sr = 0.0f, si = 0.0f;
#pragma unroll 1
	for (int j = 0; j < 21; j++) {
		const float xa = PI_F * ((float)(j - 10) + frac);
		const float tap = (xa >= 0.01f || xa <= -0.01f) ? sinf(xa) / xa : 1.0f;
		const float2 y = norm_rot_sample(x, L, q + 10 - j, nm, fs);
		sr = fmaf(tap, y.x, sr);
		si = fmaf(tap, y.y, si);
	}
	return make_float2(sr, si);
}

// ---- kernel ---------------------------------------------------------------------------------------
// Flattened symbol lists of the burst format, built once per CTA in shared memory: the training
// symbols of every sync sequence (position, reference symbol, chunk) and the data symbols in
// output order.  One lane per symbol then needs no per-chunk control flow.
static constexpr int MAX_DST4 = 1280;     // windows up to 2560 samples fill their regions in the statistics pass
struct FlatTab {
	uint16_t d_pos[480];
	uint16_t t_pos[MAX_SYNC][MAX_TRAIN];
	uint16_t t_off[MAX_SYNC][MAX_TRAIN];   // training symbol -> its sample in the region buffer, for TOA 0 (sps >= 4)
	uint8_t  t_sym[MAX_SYNC][MAX_TRAIN];
	uint8_t  t_chunk[MAX_SYNC][MAX_TRAIN];
	int32_t  n_train[MAX_SYNC];
	uint8_t  t_end[MAX_SYNC][32];              // training symbol t < 32 -> index after the last symbol of its chunk
	uint8_t  c_start[MAX_SYNC][MAX_SYNC_CHUNK];   // chunk -> its first training symbol
	int32_t  n_dsym;
	int32_t  d_lo, d_hi;             // symbol-pick offsets (samples) for which every data symbol lies inside the window
	int32_t  dst_ok;                 // dst4 is valid (the window has at most 2 * MAX_DST4 samples)
	uint16_t dst4[MAX_DST4 + 1];     // pair of samples -> float4 slot of the region buffer, 0xffff: not in a region
};

__device__ void build_flat(const BurstTab &bt, FlatTab &ft, const Regions &rg, int sps, int L)
{
	for (int t = threadIdx.x; t < 480; t += blockDim.x) {
		int acc = 0, pos = 0;
		for (int c = 0; c < bt.n_data; c++) {
			if (t >= acc && t < acc + bt.d_len[c])
				pos = bt.d_pos[c] + (t - acc);
			acc += bt.d_len[c];
		}
		ft.d_pos[t] = (uint16_t)pos;
		if (t == 0)
			ft.n_dsym = acc;
	}
	if (threadIdx.x == 0) {
		int lo = 1 << 30, hi = 0;
		for (int c = 0; c < bt.n_data; c++) {
			lo = min(lo, (int)bt.d_pos[c]);
			hi = max(hi, bt.d_pos[c] + bt.d_len[c] - 1);
		}
		ft.d_lo = -lo * sps;
		ft.d_hi = L - 1 - hi * sps;
	}
	for (int s = 0; s < bt.n_sync; s++)
		for (int t = threadIdx.x; t < MAX_TRAIN; t += blockDim.x) {
			int acc = 0, pos = 0, sym = 0, ch = 0, off = 0;
			for (int c = 0; c < bt.n_chunk[s]; c++) {
				const int cl = bt.s_len[s][c];
				if (t >= acc && t < acc + cl) {
					pos = bt.s_pos[s][c] + (t - acc);
					sym = bt.s_sym[s][c][t - acc];
					ch = c;
					off = rg.roff[0][s][c] + (t - acc) * sps;
				}
				acc += cl;
			}
			ft.t_pos[s][t] = (uint16_t)pos;
			ft.t_off[s][t] = (uint16_t)off;
			ft.t_sym[s][t] = (uint8_t)sym;
			ft.t_chunk[s][t] = (uint8_t)ch;
			if (t < 32) {
				int e = 0, a0 = 0;
				for (int c = 0; c < bt.n_chunk[s]; c++) {
					if (t >= a0 && t < a0 + bt.s_len[s][c])
						e = a0 + bt.s_len[s][c];
					if (t == 0)
						ft.c_start[s][c] = (uint8_t)min(a0, 255);
					a0 += bt.s_len[s][c];
				}
				ft.t_end[s][t] = (uint8_t)e;
			}
			if (t == 0)
				ft.n_train[s] = acc;
		}
}

// frequency error from the chunk-to-chunk phase slope (pi4cxpsk.c:360-406), any number of training symbols
// (cold: only RACH has more than 32; the usual case is inlined in the kernel)
__device__ __noinline__ float ferr_generic(const FlatTab &ft, const Regions &rg, const float2 *zbuf, int sync_id, int nch,
                                           int ntr, float2 z0, int ch0, int lane)
{
	float f = 0.0f, prev_r = 0.0f, prev_i = 0.0f, prev_pos = 0.0f;
#pragma unroll 1
	for (int c = 0; c < nch; c++) {
		float cr = ch0 == c ? z0.x : 0.0f, ci = ch0 == c ? z0.y : 0.0f;
#pragma unroll 1
		for (int t = 32 + lane; t < ntr; t += 32)
			if (ft.t_chunk[sync_id][t] == c) {
				const float2 z = zbuf[t];
				cr += z.x;
				ci += z.y;
			}
		const float2 sum = warp_sum2(cr, ci, lane);
		const float pos = rg.cpos[sync_id][c];
		if (c > 0) {   // arg(corr[c] * conj(corr[c-1])) / (pos[c] - pos[c-1])
			const float re = sum.x * prev_r + sum.y * prev_i, im = sum.y * prev_r - sum.x * prev_i;
			f += fast_atan2f(im, re) / (pos - prev_pos);
		}
		prev_r = sum.x;
		prev_i = sum.y;
		prev_pos = pos;
	}
	return f / (float)(nch - 1);
}

// mode 0: demod (bts[0] only).  mode 1: detect among n_bt burst types (pi4cxpsk.c:617-682).
// NB = bits per symbol of bts[0] (compile time: the soft-bit mapping is straight-line code).
// Persistent: each warp strides over the bursts of the batch.
template <int MODE, int SPS, int NB, int ROWS, int SYMB = DM_SYM_BATCH>
__global__ void __launch_bounds__(DM_WARPS * 32, DM_MIN_CTAS)
demod_kernel(const DemodArgs a, const BurstTab *__restrict__ bts, int n_bt, int warp_bytes,
             const __grid_constant__ Regions rg)
{
	extern __shared__ __align__(16) uint8_t smem[];
	__shared__ FlatTab ft;
#if DM_OPT_LUT
	__shared__ uint16_t soft_lut[LUT_CELLS << NB];
#endif
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const BurstTab &bt = bts[0];
	const int sps = SPS > 0 ? SPS : a.sps, L = a.win_len;
	const int w = L - bt.len * sps + 1;
	const WarpSmem sm = carve(smem + (size_t)warp * warp_bytes, rg.total, w, rg.n_slot);

	if (MODE == 0) {
		build_flat(bt, ft, rg, sps, L);
#if DM_OPT_LUT
		for (int k = threadIdx.x; k < (LUT_CELLS << NB); k += blockDim.x)
			soft_lut[k] = (uint16_t)soft_word<NB>(((float)k + 0.5f) * (1.0f / LUT_CELLS));
#endif
	}
	if (threadIdx.x == 0)
		ft.dst_ok = (L >> 1) <= MAX_DST4;
	if ((L >> 1) <= MAX_DST4)
		for (int i = threadIdx.x; i <= (L >> 1); i += blockDim.x) {      // region starts / lengths are even
			unsigned d = 0xffffu;
			for (int r = 0; r < rg.n; r++)
				if (2 * i >= rg.start[r] && 2 * i < rg.start[r] + rg.len[r])
					d = (unsigned)((rg.off[r] + 2 * i - rg.start[r]) >> 1);
			ft.dst4[i] = (uint16_t)d;
		}
	// the area behind the window must only ever hold finite values (see sync_find)
	for (int i = lane; i < ((rg.total + 1) & ~1); i += 32)
		sm.reg[i] = make_float2(0.0f, 0.0f);
	for (int i = lane; i < rg.n_slot * 32; i += 32)
		sm.taps[i] = make_float2(0.0f, 0.0f);
	for (int i = lane; i < ((rg.n_slot + 1) & ~1); i += 32)
		sm.tsum[i] = make_float2(0.0f, 0.0f);
	for (int i = lane; i < ((w + 31) & ~31) + 8; i += 32)
		sm.accv[i - 4] = 0.0f;
	for (int i = lane; i < MAX_TRAIN; i += 32)
		sm.zbuf[i] = make_float2(0.0f, 0.0f);
	__syncthreads();

	TapLane tpl;
	tpl.j = lane - 10;
	tpl.xj = PI_F * (float)(lane - 10);
	tpl.sgn = lane < 21 ? (((lane - 10) & 1) ? 1.0f : -1.0f) : 0.0f;

#if DM_OPT_LUT
	const unsigned lut_s = (unsigned)__cvta_generic_to_shared(soft_lut);
#endif
	const bool want_sd = a.pwr != nullptr || (MODE == 1 && (a.e_toa != nullptr || a.e_toa0 >= 0.0f));
	constexpr float inv_dd256 = (float)(LUT_CELLS << NB) * 0.15915494309189533577f;     // table cells per radian
	// 2*pi = TWO_PI_HI + TWO_PI_LO, TWO_PI_HI with 9 significant bits: k * TWO_PI_HI is exact for |k| < 2^15
	constexpr float TWO_PI_HI = 6.28125f, TWO_PI_LO = 1.9353071795864769e-3f, INV_2PI = 0.15915494309189533577f;
	float fs_taps = __int_as_float(0x7fc00000);     // frequency shift the cached taps were built for (NaN: none)

	const int n_eff = a.n_dev ? min(a.n, *a.n_dev) : a.n;
	for (int b = blockIdx.x * DM_WARPS + warp; b < n_eff; b += gridDim.x * DM_WARPS) {
		const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
		const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;
		const float fs = (freq_shift - rg.rotation0) / (float)sps;

#if DM_PREFETCH == 1 || DM_PREFETCH == 3
		// The statistics pass has 4 x 16 bytes per lane in flight (register budget) and would meet the DRAM
		// latency four times per window; with the whole window requested up front, batches 2-4 find their
		// lines in L2 or on the way.  No extra L2 footprint: it is this warp's own, current window.
		if (lane == 0)
			prefetch_window_l2(x, L * 8);
#endif
		__syncwarp();
		if (fs != fs_taps) {
			build_taps(bts, n_bt, rg, sm, fs, sps, lane);
			fs_taps = fs;
		}
		// aligned windows fill the correlation regions from the loads of the statistics pass
		const bool fill = ft.dst_ok && (((uintptr_t)x) & 15) == 0;
		const Norm nm = (fill && !want_sd) ? load_stats_t<false, true>(x, L, lane, true, ft.dst4, sm.reg)
		                                   : load_stats_cold(x, L, lane, want_sd, fill, ft.dst4, rg, sm.reg);
		__syncwarp();
		if (MODE == 1) {
			const float e_toa = a.e_toa ? a.e_toa[b] : a.e_toa0;
			int p_id = -1, p_sid = -1;
			float p_toa = 0.0f, p_pwr = 0.0f;
			for (int id = 0; id < n_bt; id++) {
				float toa, pwr;
				const int sid = sync_find<SPS, ROWS>(rg, id, sm, nm, sps, w, tpl, lane, a.sync_reset != 0, toa, pwr);
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
			continue;
		}

		float toa, pwr;
		const int sync_id = sync_find<SPS, ROWS>(rg, 0, sm, nm, sps, w, tpl, lane, a.sync_reset != 0, toa, pwr);
		if (lane == 0) {
			if (a.sync_id) a.sync_id[b] = sync_id;
			if (a.toa) a.toa[b] = toa;
			if (a.pwr) a.pwr[b] = pwr;
		}
		int8_t *eb = a.ebits + (size_t)b * a.ebits_stride;
		if (sync_id < 0) {          // nothing correlated (all-zero input): the reference returns -errno
			if (lane == 0 && a.freq_err) a.freq_err[b] = 0.0f;
#pragma unroll 1
			for (int k = lane; k < bt.ebits; k += 32)
				eb[k] = 0;
			continue;
		}

		// symbol i sits at sample i*sps + d (sps >= 4 path of _gmr1_pi4cxpsk_align, :286-297); for
		// sps < 4 (SPS == -1) the fractional part is interpolated when it exceeds 0.1 sample (:298-343)
		const int d = (int)roundf(toa);
		const float ofs_frac = toa - (float)d;
		const bool interp = SPS < 0 && fabsf(ofs_frac) > 0.1f;
		auto sample_of = [&](int i) {
			const int q = i * sps + d;     // d >= -1; index -1 would read before the window: clamp
			return min(max(q, 0), L - 1);
		};

		// ---- training symbols, one per lane, derotated as the reference derotates every sample:
		//      z = (x - avg)/sd * e^{j*fl32(fs*idx)}, times conj(reference symbol).  Per-chunk sums ->
		//      fine frequency error from the chunk-to-chunk phase slope (:360-406).
		const int nch = rg.n_chunk[0][sync_id], ntr = ft.n_train[sync_id];
		float2 z0 = make_float2(0.0f, 0.0f);     // training symbol `lane` (round 0) stays in registers
		int ch0 = -1;
		__syncwarp();        // zbuf doubles as the peak search's window buffer (last read there: the peak value)
#pragma unroll 1
		for (int t0 = 0; t0 < ntr; t0 += 32) {
			const int t = t0 + lane;
			if (t < ntr) {
				const int pos = ft.t_pos[sync_id][t];
				float2 y;
				if (SPS < 0) {
					y = lowsps_symbol(x, L, pos * sps + d, interp, ofs_frac, nm, fs);
				} else {
					// the training symbols lie inside the correlation regions: shared memory, not L2.  (d = -1
					// reads the sample in front of the chunk: regions start one pair early for that.)
					const int q = sample_of(pos);
					const float2 v = sm.reg[(int)ft.t_off[sync_id][t] + d];
					const float2 e = sincos_red(fs * (float)q);
					const float yr = (v.x - nm.ar) * nm.inv_sd, yi = (v.y - nm.ai) * nm.inv_sd;
					y = make_float2(yr * e.x - yi * e.y, yr * e.y + yi * e.x);
				}
				const float2 z = mul_conj_sym(ft.t_sym[sync_id][t], y);
				sm.zbuf[t] = z;
				if (t0 == 0) {
					z0 = z;
					ch0 = ft.t_chunk[sync_id][t];
				}
			}
		}
		__syncwarp();
		float ferr = 0.0f;
		if (nch > 1 && ntr <= 32) {
			// all chunk sums at once: segmented shuffle reduction over the training symbols in lanes 0..ntr-1
			// (a lane adds the value `o` lanes up while that lane is still inside its chunk); the sum of chunk c
			// ends in the chunk's first lane
			float2 v = z0;
			const int end = lane < ntr ? (int)ft.t_end[sync_id][lane] : 0;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const float ur = __shfl_down_sync(0xffffffffu, v.x, o), ui = __shfl_down_sync(0xffffffffu, v.y, o);
				if (lane + o < end) {
					v.x += ur;
					v.y += ui;
				}
			}
			// lane c < nch: chunk c and chunk c-1, arg(corr[c] * conj(corr[c-1])) / (pos[c] - pos[c-1]) - the
			// nch-1 angles in parallel lanes
			const int c = min(lane, nch - 1), src = ft.c_start[sync_id][c], srcp = ft.c_start[sync_id][max(c - 1, 0)];
			const float sr = __shfl_sync(0xffffffffu, v.x, src), si = __shfl_sync(0xffffffffu, v.y, src);
			const float qr = __shfl_sync(0xffffffffu, v.x, srcp), qi = __shfl_sync(0xffffffffu, v.y, srcp);
			const float re = sr * qr + si * qi, im = si * qr - sr * qi;
			const float part = fast_atan2f_inl(im, re) / (rg.cpos[sync_id][c] - rg.cpos[sync_id][max(c - 1, 0)]);
			float f = 0.0f;
#pragma unroll 1
			for (int k = 1; k < nch; k++)
				f += __shfl_sync(0xffffffffu, part, k);
			ferr = f / (float)(nch - 1);
		} else if (nch > 1) {
			ferr = ferr_generic(ft, rg, sm.zbuf, sync_id, nch, ntr, z0, ch0, lane);
		}
		if (lane == 0 && a.freq_err) a.freq_err[b] = ferr;

		// ---- phase reference: all training symbols after the -ferr rotation (:415-433, :574)
		float phi0;
		{
			float pr = 0.0f, pi = 0.0f;
#pragma unroll 1
			for (int t0 = 0; t0 < ntr; t0 += 32) {
				const int t = t0 + lane;
				if (t < ntr) {
					float2 z = sm.zbuf[t];
					if (ferr != 0.0f) {
						const float2 e = sincos_red((-ferr) * (float)ft.t_pos[sync_id][t]);
						z = make_float2(z.x * e.x - z.y * e.y, z.x * e.y + z.y * e.x);
					}
					pr += z.x;
					pi += z.y;
				}
			}
			const float2 sum = warp_sum2(pr, pi, lane);
			phi0 = fast_atan2f(sum.y, sum.x);
		}

		// ---- data symbols in the angle domain, one per lane in output order.  The reference rotates
		// each sample three times (e^{j*fs*idx}, e^{-j*ferr*i}, conj(phasor)) and takes cargf(); the
		// argument of that product is  arg(x - avg) + fl32(fs*idx) + fl32(-ferr*i) - arg(phasor)
		// (mod 2*pi).  fl32(fs*idx) reaches a few hundred radians, so it is reduced mod 2*pi first
		// (Cody-Waite, exact to ~1e-8 rad); the remaining three float additions of values below 2*pi
		// keep the sum within ~5e-7 rad, 4e-5 of a soft-bit step.
		const int nds = ft.n_dsym;
		const float nferr = -ferr, nphi0 = -phi0;
		const bool eb_even = (((uintptr_t)eb) & 1) == 0;
		// soft bits of data symbol t (burst position i) whose derotated angle, before the -ferr*i and
		// -phase corrections, is ang
		auto emit = [&](int t, int i, float ang) {
#if DM_OPT_LUT
			const float sv = ((ang + nferr * (float)i) + nphi0) * inv_dd256;
			const unsigned v = soft_lut[__float2int_rd(sv) & ((LUT_CELLS << NB) - 1)];
#else
			const unsigned v = soft_word<NB>(((ang + nferr * (float)i) + nphi0) * (inv_dd256 * (1.0f / LUT_CELLS)));
#endif
			if (NB == 2) {
				int8_t *o = eb + 2 * t;
				if (eb_even)
					*reinterpret_cast<uint16_t *>(o) = (uint16_t)v;
				else {
					o[0] = (int8_t)(v & 0xff);
					o[1] = (int8_t)(v >> 8);
				}
			} else {
				eb[t] = (int8_t)v;
			}
		};
		if (SPS < 0) {               // interpolated symbols already carry the derotation
			for (int t = lane; t < nds; t += 32) {
				const int i = ft.d_pos[t];
				const float2 z = lowsps_symbol(x, L, i * sps + d, interp, ofs_frac, nm, fs);
				emit(t, i, fast_atan2f_inl(z.y, z.x));
			}
		} else {
#if DM_PREFETCH >= 2
			{	// the next window of this warp starts its way to L2 while the data symbols are sliced
				const int bn = b + gridDim.x * DM_WARPS;
				if (lane == 0 && bn < n_eff)
					prefetch_window_l2(a.iq + (a.ofs ? a.ofs[bn] : (int64_t)bn * a.stride),
					                   DM_PREFETCH == 3 ? min(L * 8, DM_PF_NEXT) : L * 8);
			}
#endif
			// DM_SYM_BATCH symbols per lane and pass: their sample loads (L2 hits) are in flight together and the
			// angle arithmetic of the batch is straight-line code (indices past the end are clamped and only
			// the store is predicated), so the independent chains interleave.
			const int dc = min(max(d, ft.d_lo), ft.d_hi);        // no data symbol leaves the window (never binds for the standard formats)
			const float2 *xd = x + dc;
			const bool fast_store = NB == 1 || eb_even;
			auto data_pass = [&](auto bsel) {
			constexpr int B = decltype(bsel)::value;
#pragma unroll 1
			for (int t0 = lane; t0 < nds; t0 += 32 * B) {
				int ii[B];
				float2 v[B];
#pragma unroll
				for (int u = 0; u < B; u++) {
					ii[u] = ft.d_pos[min(t0 + 32 * u, nds - 1)];
					v[u] = __ldg(&xd[ii[u] * sps]);
				}
				unsigned sw[B];
#pragma unroll
				for (int u = 0; u < B; u++) {
					const float th = fast_atan2f_inl(v[u].y - nm.ai, v[u].x - nm.ar);
					const float a1 = fs * (float)(ii[u] * sps + dc);
					const float k = rintf(a1 * INV_2PI);
					float r = fmaf(k, -TWO_PI_HI, a1);
					r = fmaf(k, -TWO_PI_LO, r);
					const float sv = (((th + r) + nferr * (float)ii[u]) + nphi0) * inv_dd256;
#if DM_OPT_LUT
					const unsigned cell = (unsigned)__float2int_rd(sv) & ((LUT_CELLS << NB) - 1);
					asm("ld.shared.u16 %0, [%1];" : "=r"(sw[u]) : "r"(lut_s + 2 * cell));
#else
					sw[u] = soft_word<NB>(sv * (1.0f / LUT_CELLS));
#endif
				}
				if (fast_store) {
#pragma unroll
					for (int u = 0; u < B; u++)
						if (t0 + 32 * u < nds) {
							if (NB == 2)
								reinterpret_cast<uint16_t *>(eb)[t0 + 32 * u] = (uint16_t)sw[u];
							else
								eb[t0 + 32 * u] = (int8_t)sw[u];
						}
				} else {
#pragma unroll 1
					for (int u = 0; u < B; u++) {     // odd output address: byte stores (cold)
						unsigned val = sw[0];
#pragma unroll
						for (int k = 1; k < B; k++)
							val = u == k ? sw[k] : val;
						if (t0 + 32 * u < nds) {
							eb[2 * (t0 + 32 * u)] = (int8_t)(val & 0xff);
							eb[2 * (t0 + 32 * u) + 1] = (int8_t)(val >> 8);
						}
					}
				}
			}
			};
			// SYMB (template parameter, chosen by the launcher): formats with 7 rows of 32 data symbols (BCCH, DC6,
			// NT6, SDCCH; DC12 has 14) go in batches of 7, everything else in batches of DM_SYM_BATCH - whichever
			// leaves fewer empty rows.  One batch size per kernel: two copies of the pass do not fit the hot path.
			data_pass(std::integral_constant<int, SYMB>());
		}
	}
}

// ---- launcher --------------------------------------------------------------------------------------
std::atomic<int> g_sync_reset{0};

cudaError_t launch_demod(const DemodArgs &a_in, const BurstTab *d_bts, const BurstTab *h_bts, int n_bt, int mode,
                         cudaStream_t st)
{
	if (a_in.n <= 0)
		return cudaSuccess;
	DemodArgs a = a_in;
	a.sync_reset = g_sync_reset.load(std::memory_order_relaxed);
	if (mode == 0 && n_bt == 1 && a.sps == 4)       // standard format at its standard search width: per-format kernel
		for (int i = 0; i < BT_COUNT; i++)
			if (!memcmp(&h_bts[0], &burst_tab(i), sizeof(BurstTab))) {
				cudaError_t fe;
				if (launch_demod_fast(a, i, st, &fe))
					return fe;
				break;
			}
	int maxlen = 0, minlen = 1 << 30;
	for (int i = 0; i < n_bt; i++) {
		maxlen = h_bts[i].len > maxlen ? h_bts[i].len : maxlen;
		minlen = h_bts[i].len < minlen ? h_bts[i].len : minlen;
	}
	if (maxlen != minlen)
		return cudaErrorInvalidValue;      // detect needs length-compatible burst types
	const int w = a.win_len - maxlen * a.sps + 1;
	if (w < 1 || a.sps < 1 || a.sps > 16)
		return cudaErrorInvalidValue;
	// union of the intervals the training-sequence search reads, over all types / sequences / chunks
	Regions rg;
	memset(&rg, 0, sizeof(rg));
	{
		struct Iv { int lo, hi; } iv[64];
		int ni = 0;
		for (int i = 0; i < n_bt; i++)
			for (int s = 0; s < h_bts[i].n_sync; s++)
				for (int c = 0; c < h_bts[i].n_chunk[s]; c++) {
					int lo = h_bts[i].s_pos[s][c] * a.sps;
					int hi = lo + (h_bts[i].s_len[s][c] - 1 + 3) * a.sps + w;     // + 3 zero-padded taps
					lo = lo >= 2 ? lo - 2 : 0;    // the symbol pick for TOA -0.5 .. 0 reads one sample in front
					hi = hi > a.win_len ? a.win_len : hi;
					lo &= ~1;                     // whole pairs of samples (16-byte stores into the region buffer)
					hi = (hi + 1) & ~1;
					if (ni < 64 && lo < hi)
						iv[ni++] = {lo, hi};
				}
		for (int i = 1; i < ni; i++)                      // insertion sort by start
			for (int j = i; j > 0 && iv[j].lo < iv[j - 1].lo; j--) {
				Iv t = iv[j]; iv[j] = iv[j - 1]; iv[j - 1] = t;
			}
		for (int i = 0; i < ni; i++) {
			if (rg.n && iv[i].lo <= rg.start[rg.n - 1] + rg.len[rg.n - 1]) {
				const int hi = rg.start[rg.n - 1] + rg.len[rg.n - 1];
				if (iv[i].hi > hi)
					rg.len[rg.n - 1] = iv[i].hi - rg.start[rg.n - 1];
			} else {
				if (rg.n == MAX_REGIONS)
					return cudaErrorInvalidValue;
				rg.start[rg.n] = iv[i].lo;
				rg.len[rg.n] = iv[i].hi - iv[i].lo;
				rg.n++;
			}
		}
		for (int r = 0; r < rg.n; r++) {
			rg.off[r] = rg.total;
			rg.total += (rg.len[r] + 1) & ~1;
		}
		if (n_bt > MAX_BT)
			return cudaErrorInvalidValue;
		for (int i = 0; i < n_bt; i++) {
			rg.n_sync[i] = (uint8_t)h_bts[i].n_sync;
			for (int s = 0; s < h_bts[i].n_sync; s++) {
				rg.n_chunk[i][s] = (uint8_t)h_bts[i].n_chunk[s];
				for (int c = 0; c < h_bts[i].n_chunk[s]; c++) {
					rg.slot[i][s][c] = (uint8_t)rg.n_slot++;
					rg.cl[i][s][c] = (uint8_t)h_bts[i].s_len[s][c];
					const int b0 = h_bts[i].s_pos[s][c] * a.sps;
					for (int r = 0; r < rg.n; r++)
						if (b0 >= rg.start[r] && b0 < rg.start[r] + rg.len[r])
							rg.roff[i][s][c] = (uint16_t)(rg.off[r] + (b0 - rg.start[r]));
					if (i == 0)
						rg.cpos[s][c] = (float)h_bts[0].s_pos[s][c] + (float)h_bts[0].s_len[s][c] / 2.0f;
				}
			}
		}
		rg.rotation0 = h_bts[0].rotation;
	}
	const size_t wb = (warp_smem_bytes(rg.total, w, rg.n_slot) + 15) & ~(size_t)15;
	const size_t smem = wb * DM_WARPS;
	if (smem > 227 * 1024)
		return cudaErrorInvalidValue;
	GMR1_INIT_LOCK();
	static size_t attr_set[64] = {0};
	static bool tab_up[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 64 || !tab_up[dev]) {
		cudaError_t e = upload_sinpi512();
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			tab_up[dev] = true;
	}
	if (dev >= 64 || attr_set[dev] < smem) {
		cudaError_t e = cudaSuccess;
		const void *fns[] = {(const void *)demod_kernel<0, 4, 1, 0>, (const void *)demod_kernel<0, 4, 2, 0>,
		                     (const void *)demod_kernel<0, 4, 1, 1>, (const void *)demod_kernel<0, 4, 2, 1>,
		                     (const void *)demod_kernel<0, 4, 1, 2>, (const void *)demod_kernel<0, 4, 2, 2>,
		                     (const void *)demod_kernel<0, 4, 1, 3>, (const void *)demod_kernel<0, 4, 2, 3>,
		                     (const void *)demod_kernel<0, 4, 1, 0, 7>, (const void *)demod_kernel<0, 4, 2, 0, 7>,
		                     (const void *)demod_kernel<0, 4, 1, 1, 7>, (const void *)demod_kernel<0, 4, 2, 1, 7>,
		                     (const void *)demod_kernel<0, 4, 1, 2, 7>, (const void *)demod_kernel<0, 4, 2, 2, 7>,
		                     (const void *)demod_kernel<0, 4, 1, 3, 7>, (const void *)demod_kernel<0, 4, 2, 3, 7>,
		                     (const void *)demod_kernel<0, 0, 1, 0>, (const void *)demod_kernel<0, 0, 2, 0>,
		                     (const void *)demod_kernel<0, -1, 1, 0>, (const void *)demod_kernel<0, -1, 2, 0>,
		                     (const void *)demod_kernel<1, 4, 2, 0>, (const void *)demod_kernel<1, 0, 2, 0>};
		for (size_t i = 0; i < sizeof(fns) / sizeof(fns[0]) && e == cudaSuccess; i++)
			e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_set[dev] = smem;
	}
	// persistent warps: enough CTAs to fill the machine at the occupancy shared memory allows
	static int n_sm[64] = {0};
	if (dev >= 64 || !n_sm[dev]) {
		int v = 148;
		cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
		if (dev < 64)
			n_sm[dev] = v;
	}
	const int sms = dev < 64 ? n_sm[dev] : 148;
	int per_sm = (int)((228 * 1024) / (smem + sizeof(FlatTab) + 2 * (LUT_CELLS << 2) + 1024 + 64));
	per_sm = per_sm < 1 ? 1 : (per_sm > DM_MIN_CTAS ? DM_MIN_CTAS : per_sm);
	if (const char *e = getenv("GMR1B200_DEMOD_CTAS")) {     // tuning knob: resident CTAs per SM
		const int v = atoi(e);
		if (v >= 1 && v < per_sm)
			per_sm = v;
	}
	int grid = (a.n + DM_WARPS - 1) / DM_WARPS;
	if (grid > sms * per_sm)
		grid = sms * per_sm;
	const int nb = h_bts[0].nbits;
	if (nb != 1 && nb != 2)
		return cudaErrorInvalidValue;
#define DM_LAUNCH(M, S, B, R) demod_kernel<M, S, B, R><<<grid, DM_WARPS * 32, smem, st>>>(a, d_bts, n_bt, (int)wb, rg)
#define DM_LAUNCH7(M, S, B, R) demod_kernel<M, S, B, R, 7><<<grid, DM_WARPS * 32, smem, st>>>(a, d_bts, n_bt, (int)wb, rg)
#define DM_LAUNCH_NB(M, S, R) do { \
		if (S == 4 && batch7) { if (nb == 1) DM_LAUNCH7(M, S, 1, R); else DM_LAUNCH7(M, S, 2, R); } \
		else { if (nb == 1) DM_LAUNCH(M, S, 1, R); else DM_LAUNCH(M, S, 2, R); } \
	} while (0)
	// data symbols per lane and pass: batches of 7 when that leaves fewer empty rows of 32 symbols than batches of 4
	bool batch7;
	{
		int nds = 0;
		for (int c = 0; c < h_bts[0].n_data; c++)
			nds += h_bts[0].d_len[c];
		const int rows = (nds + 31) / 32;
		const int waste7 = (7 - rows % 7) % 7, waste4 = (DM_SYM_BATCH - rows % DM_SYM_BATCH) % DM_SYM_BATCH;
		batch7 = rows >= 7 && waste7 <= waste4;
	}
	if (mode == 0 && a.sps < 4) {
		DM_LAUNCH_NB(0, -1, 0);
	} else if (mode == 0 && a.sps == 4) {
		// the number of rows of 32 search offsets is a compile-time constant on the fast path
		if (w <= 32)      DM_LAUNCH_NB(0, 4, 1);
		else if (w <= 64) DM_LAUNCH_NB(0, 4, 2);
		else if (w <= 96) DM_LAUNCH_NB(0, 4, 3);
		else              DM_LAUNCH_NB(0, 4, 0);
	} else if (mode == 0) {
		DM_LAUNCH_NB(0, 0, 0);
	} else if (a.sps == 4)
		DM_LAUNCH(1, 4, 2, 0);
	else
		DM_LAUNCH(1, 0, 2, 0);
#undef DM_LAUNCH_NB
#undef DM_LAUNCH7
#undef DM_LAUNCH
	return cudaGetLastError();
}

}  // namespace gmr1
