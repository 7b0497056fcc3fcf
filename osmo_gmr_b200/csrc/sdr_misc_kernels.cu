// sdr_misc_kernels.cu - the two small SDR entry points beside the pi/4-CxPSK demodulator:
//   dkab_kernel       one warp per NT3-sized window: keep-alive burst (DKAB) detection, TOA and 8
//                     differential soft bits.   Replaces gmr1_dkab_demod, src/sdr/dkab.c:187-214
//                     (_gmr1_dkab_find_toa :57-144, _gmr1_dkab_soft_bits :154-172).
//   mod_order_kernel  one warp per burst: BPSK vs QPSK from |sum x^2|^2 vs |sum x^4|^2 / 2.
//                     Replaces gmr1_pi4cxpsk_mod_order, src/sdr/pi4cxpsk.c:693-729.
#include <cuda_runtime.h>
#include <math.h>

#include "launch.h"

namespace gmr1 {

static constexpr float PI_F = 3.14159265358979323846264338327f;
static constexpr int MW = 4;      // warps per CTA

__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// osmo_cxvec_sig_normalize(in, 1, fs) into shared memory (whole window, rotated)
__device__ void normalize_window(const float2 *__restrict__ x, int L, float fs, float2 *y, int lane)
{
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	for (int i = lane; i < L; i += 32) {
		const float2 v = __ldg(&x[i]);
		y[i] = v;
		sr += v.x;
		si += v.y;
		sq = fmaf(v.x, v.x, sq);
		sq = fmaf(v.y, v.y, sq);
	}
	sr = wsum(sr); si = wsum(si); sq = wsum(sq);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	for (int i = lane; i < L; i += 32) {
		const float2 v = y[i];
		float yr = (v.x - ar) / sd, yi = (v.y - ai) / sd;
		if (fs != 0.0f) {
			float sn, cs;
			sincosf(fs * (float)i, &sn, &cs);
			const float tr = yr * cs - yi * sn;
			yi = yr * sn + yi * cs;
			yr = tr;
		}
		y[i] = make_float2(yr, yi);
	}
	__syncwarp();
}

__device__ __forceinline__ float nsq(float2 v) { return v.x * v.x + v.y * v.y; }

__global__ void __launch_bounds__(MW * 32) dkab_kernel(const MiscArgs a)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int b = blockIdx.x * MW + warp;
	if (b >= (a.n_dev ? min(a.n, *a.n_dev) : a.n))
		return;
	const int sps = a.sps, L = a.win_len;
	float2 *y = (float2 *)smem + (size_t)warp * ((L + 1) & ~1);
	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;
	const int p = a.dkab_p ? a.dkab_p[b] : a.dkab_p0;

	normalize_window(x, L, (freq_shift - PI_F / 4.0f) / (float)sps, y, lane);

	const int w = L - (39 * 3 * sps) + 1;
	int rv = 0;
	float toa = 0.0f;
	if (w <= 0) {
		rv = -22;      // -EINVAL
	} else if (lane == 0) {
		// sliding energy of the two 5-symbol KAB pulses; the running sum is kept in the same
		// (sequential) order as the C code, w is small (7 for gmr1_rx's windows)
		const int o0 = sps * (2 + p), o1 = sps * (2 + p + 59), d = sps * 5;
		float cur = 0.0f;
		for (int i = 0; i < d; i++)
			cur += nsq(y[o0 + i]) + nsq(y[o1 + i]);
		int mi = 0;
		float mp = cur, prev = 0.0f, at_m1 = 0.0f, at_p1 = 0.0f;
		bool want_next = false;
		for (int i = 0; i < w - 1; i++) {
			const float np = cur - nsq(y[o0 + i]) - nsq(y[o1 + i]) + nsq(y[o0 + d + i]) + nsq(y[o1 + d + i]);
			if (want_next) {
				at_p1 = np;
				want_next = false;
			}
			if (np > mp) {
				mi = i + 1;
				mp = np;
				at_m1 = cur;
				want_next = true;
				at_p1 = 0.0f;
			}
			prev = cur;
			cur = np;
		}
		(void)prev;
		toa = (float)mi;
		if (mi > 0 && mi < w - 1)      // parabolic refinement around the maximum (dkab.c:107-110)
			toa += 0.5f * (-at_m1 + at_p1) / (-at_m1 + 2.0f * mp - at_p1);
		toa += (float)(sps - 1) / 2.0f;
		const int toa_i = (int)roundf(toa);
		float egy_peak = 0.0f, egy_valley = 0.0f;
		for (int i = 0; i < d; i++)
			egy_peak += nsq(y[min(toa_i + o0 + i, L - 1)]) + nsq(y[min(toa_i + o1 + i, L - 1)]);
		egy_peak /= (float)(d * 2);
		const int l_valley = o1 - o0 - d;
		for (int i = 0; i < l_valley; i++)
			egy_valley += nsq(y[min(toa_i + o0 + d + i, L - 1)]);
		egy_valley /= (float)l_valley;
		rv = (egy_peak / egy_valley) > 10.0f ? 0 : 1;       // DKAB_PWR_RATIO_THRESHOLD
	}
	rv = __shfl_sync(0xffffffffu, rv, 0);
	toa = __shfl_sync(0xffffffffu, toa, 0);
	if (lane == 0) {
		if (a.rv) a.rv[b] = rv;
		if (a.toa) a.toa[b] = toa;
	}
	if (rv == 0 && lane < 8 && a.ebits) {
		// differential phase between consecutive KAB symbols (dkab.c:154-172)
		const int toa_i = (int)roundf(toa);
		const int o = toa_i + sps * (2 + p + (lane >> 2) * 59) + sps * (lane & 3);
		const float2 u = y[min(o, L - 1)], v = y[min(o + sps, L - 1)];
		const float re = u.x * v.x + u.y * v.y, im = u.y * v.x - u.x * v.y;   // u * conj(v)
		const float pd = atan2f(im, re);
		a.ebits[(size_t)b * 8 + lane] = (int8_t)roundf((0.5f - (fabsf(pd) / PI_F)) * 254.0f);
	}
}

__global__ void __launch_bounds__(MW * 32) mod_order_kernel(const MiscArgs a)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int b = blockIdx.x * MW + warp;
	if (b >= a.n)
		return;
	const int sps = a.sps, L = a.win_len;
	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;
	const float fs = (freq_shift - PI_F / 4.0f) / (float)sps;

	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	for (int i = lane; i < L; i += 32) {
		const float2 v = __ldg(&x[i]);
		sr += v.x;
		si += v.y;
		sq = fmaf(v.x, v.x, sq);
		sq = fmaf(v.y, v.y, sq);
	}
	sr = wsum(sr); si = wsum(si); sq = wsum(sq);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	float br = 0.0f, bi = 0.0f, qr = 0.0f, qi = 0.0f;
	for (int i = lane; i < L; i += 32) {
		const float2 v = __ldg(&x[i]);
		float yr = (v.x - ar) / sd, yi = (v.y - ai) / sd;
		if (fs != 0.0f) {
			float sn, cs;
			sincosf(fs * (float)i, &sn, &cs);
			const float tr = yr * cs - yi * sn;
			yi = yr * sn + yi * cs;
			yr = tr;
		}
		const float n2 = yr * yr + yi * yi;
		const float vr = (yr * yr - yi * yi) / n2, vi = (2.0f * yr * yi) / n2;   // v*v / |v|^2
		if (n2 > 0.0f) {
			br += vr;
			bi += vi;
			qr += vr * vr - vi * vi;
			qi += 2.0f * vr * vi;
		}
	}
	br = wsum(br); bi = wsum(bi); qr = wsum(qr); qi = wsum(qi);
	if (lane == 0)
		a.rv[b] = (br * br + bi * bi) < ((qr * qr + qi * qi) / 2.0f) ? 4 : 2;
}

cudaError_t launch_dkab(const MiscArgs &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	const size_t smem = (size_t)MW * ((a.win_len + 1) & ~1) * sizeof(float2);
	if (smem > 227 * 1024)
		return cudaErrorInvalidValue;
	GMR1_INIT_LOCK();
	static size_t attr_set[64] = {0};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev >= 64 || attr_set[dev] < smem) {
		cudaError_t e = cudaFuncSetAttribute(dkab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_set[dev] = smem;
	}
	dkab_kernel<<<(a.n + MW - 1) / MW, MW * 32, smem, st>>>(a);
	return cudaGetLastError();
}

cudaError_t launch_mod_order(const MiscArgs &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	mod_order_kernel<<<(a.n + MW - 1) / MW, MW * 32, 0, st>>>(a);
	return cudaGetLastError();
}

}  // namespace gmr1
