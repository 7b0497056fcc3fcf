// synth_kernels.cu - synthetic pi/4-CxPSK burst generator on the GPU.
//
// The reference has no general modulator (gmr1_pi4cxpsk_mod, src/sdr/pi4cxpsk.c:741-799, is one
// sample per symbol without pulse shaping or channel), so benchmark / test recordings of the
// named shapes are synthesised here: symbol mapping per the burst descriptor (sync chunks + Gray
// mapped data chunks, guard symbols silent), continuous pi/4 (pi/2) rotation, raised-cosine
// pulse (= RRC 0.35 transmit filter x RRC 0.35 matched filter, what gmr1_rx expects to have
// happened upstream, utils/gmr1_rx_sdr.py:523-529), fractional timing offset, carrier frequency
// offset, phase, amplitude and AWGN from a counter-based generator (Philox4x32-10, seeded per
// (seed, burst, sample) so output does not depend on launch geometry).
// One CTA per burst window.  This is workload construction, not the receive hot path.
#include <cuda_runtime.h>

#include "gmr1_tables.h"
#include "launch.h"

namespace gmr1 {

static constexpr int SY_T = 128;
static constexpr int RC_TAPS = 13;          // symbols -6 .. +6 around the current one

__device__ __forceinline__ void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                       uint32_t (&out)[4])
{
#pragma unroll
	for (int r = 0; r < 10; r++) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		c0 = hi1 ^ c1 ^ k0;
		c1 = lo1;
		c2 = hi0 ^ c3 ^ k1;
		c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double raised_cosine(double t)
{
	const double alpha = 0.35, pi = 3.14159265358979323846;
	const double den = 1.0 - (2.0 * alpha * t) * (2.0 * alpha * t);
	const double x = pi * t;
	const double snc = fabs(t) < 1e-9 ? 1.0 : sin(x) / x;
	if (fabs(den) < 1e-9) {
		const double y = pi / (2.0 * alpha);
		return (pi / 4.0) * (sin(y) / y);
	}
	return snc * cos(pi * alpha * t) / den;
}

// root raised cosine, unit energy per symbol (t in symbols): the transmit half of the pulse above
__device__ __forceinline__ double root_raised_cosine(double t)
{
	const double alpha = 0.35, pi = 3.14159265358979323846;
	if (fabs(t) < 1e-9)
		return 1.0 - alpha + 4.0 * alpha / pi;
	const double q = 4.0 * alpha * t;
	if (fabs(fabs(q) - 1.0) < 1e-9)
		return alpha / sqrt(2.0) * ((1.0 + 2.0 / pi) * sin(pi / (4.0 * alpha)) + (1.0 - 2.0 / pi) * cos(pi / (4.0 * alpha)));
	return (sin(pi * t * (1.0 - alpha)) + q * cos(pi * t * (1.0 + alpha))) / (pi * t * (1.0 - q * q));
}

__global__ void __launch_bounds__(SY_T) synth_kernel(const SynthArgs a, const BurstTab *__restrict__ btp)
{
	__shared__ float2 sym[480];
	__shared__ float rcw[16][RC_TAPS];
	const BurstTab &bt = *btp;
	const int b = blockIdx.x, tid = threadIdx.x;
	const int len = bt.len, sps = a.sps, L = a.win_len;
	const double pi = 3.14159265358979323846;

	// 1. symbols
	for (int i = tid; i < len; i += SY_T)
		sym[i] = make_float2(0.0f, 0.0f);
	__syncthreads();
	const int sid = a.sync_id ? a.sync_id[b] : 0;
	const uint8_t *eb = a.ebits + (size_t)b * a.ebits_stride;
	for (int c = 0; c < bt.n_chunk[sid]; c++) {
		const int p0 = bt.s_pos[sid][c], cl = bt.s_len[sid][c];
		for (int j = tid; j < cl; j += SY_T) {
			const int k = p0 + j;
			double sn, cs;
			sincos((pi / 2.0) * bt.s_sym[sid][c][j] + (double)bt.rotation * k, &sn, &cs);
			sym[k] = make_float2((float)cs, (float)sn);
		}
	}
	int kbase = 0;
	for (int c = 0; c < bt.n_data; c++) {
		const int p0 = bt.d_pos[c], cl = bt.d_len[c];
		for (int j = tid; j < cl; j += SY_T) {
			const int k = p0 + j;
			int ph;
			if (bt.nbits == 2) {
				const int v = ((eb[kbase + 2 * j] & 1) << 1) | (eb[kbase + 2 * j + 1] & 1);
				ph = v == 0 ? 0 : v == 1 ? 1 : v == 3 ? 2 : 3;        // Gray: 00 01 11 10
			} else {
				ph = 2 * (eb[kbase + j] & 1);
			}
			double sn, cs;
			sincos((pi / 2.0) * ph + (double)bt.rotation * k, &sn, &cs);
			sym[k] = make_float2((float)cs, (float)sn);
		}
		kbase += cl * bt.nbits;
	}

	// 2. pulse taps per sample phase
	const float toa = a.toa ? a.toa[b] : a.toa0;
	const float cfo = a.cfo ? a.cfo[b] : a.cfo0;
	const float phase = a.phase ? a.phase[b] : a.phase0;
	const float esn0 = a.esn0_db ? a.esn0_db[b] : a.esn0_db0;
	const float amp = a.amp ? a.amp[b] : a.amp0;
	const int ti = (int)floorf(toa);
	const double tf = (double)toa - (double)ti;
	for (int i = tid; i < sps * RC_TAPS; i += SY_T) {
		const int r = i / RC_TAPS, j = i % RC_TAPS - 6;
		const double tt = ((double)r - tf) / sps - (double)j;
		rcw[r][j + 6] = (float)(a.tx_pulse ? root_raised_cosine(tt) : raised_cosine(tt));
	}
	__syncthreads();

	// 3. samples
	const float sigma = esn0 >= 100.0f ? 0.0f : amp * exp10f(-esn0 / 20.0f) * 0.70710678118654752f;
	float2 *out = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	for (int nn = tid; nn < L; nn += SY_T) {
		const int dnn = nn - ti;
		int q = dnn / sps, r = dnn % sps;
		if (r < 0) {
			r += sps;
			q -= 1;
		}
		float xr = 0.0f, xi = 0.0f;
#pragma unroll
		for (int j = -6; j <= 6; j++) {
			const int k = q + j;
			if (k >= 0 && k < len) {
				const float wgt = rcw[r][j + 6];
				xr += sym[k].x * wgt;
				xi += sym[k].y * wgt;
			}
		}
		const double t = (double)q + ((double)r - tf) / sps;       // symbols since symbol 0
		double sn, cs;
		sincos((double)cfo * t + (double)phase, &sn, &cs);
		float yr = amp * (float)(xr * cs - xi * sn), yi = amp * (float)(xr * sn + xi * cs);
		uint32_t rnd[4];
		philox((uint32_t)nn, (uint32_t)b, (uint32_t)(a.seed >> 32), 0x67d1u, (uint32_t)a.seed, 0x3c6ef372u, rnd);
		const float u1 = ((float)rnd[0] + 0.5f) * 2.3283064365386963e-10f;
		const float u2 = ((float)rnd[1] + 0.5f) * 2.3283064365386963e-10f;
		const float rad = sqrtf(-2.0f * logf(fmaxf(u1, 1e-30f)));
		float ns, nc;
		sincospif(2.0f * u2, &ns, &nc);
		yr += sigma * rad * nc;
		yi += sigma * rad * ns;
		out[nn] = make_float2(yr, yi);
	}
}

cudaError_t launch_synth(const SynthArgs &a, const BurstTab *d_bt, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	synth_kernel<<<a.n, SY_T, 0, st>>>(a, d_bt);
	return cudaGetLastError();
}

}  // namespace gmr1
