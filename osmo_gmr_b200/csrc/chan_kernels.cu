// chan_kernels.cu - wideband channeliser: one wideband recording -> one sps-oversampled stream per
// ARFCN, the step in front of the receive path.  Replaces the "PFB Channelizer mode" of utils/gmr1_rx_sdr.py
// (:391-604), which wires GNU Radio blocks: pfb.channelizer_ccf(n_chans, low-pass taps, oversample 2) (:465-470) and,
// per ARFCN, pfb.arb_resampler_ccf(93.6 / 62.5, RRC 0.35 taps, 32 phases) (:591-596).
//
//   pfb_kernel      2x oversampled polyphase analysis bank.  Output step m, channel k:
//                     y_k[m] = (-1)^(k m) * sum_p e^{+j 2 pi k p / N} * u_p[m],   u_p[m] = sum_q h[p + q N] x[m N/2 - p - q N]
//                   One CTA makes T = 8 G steps: the branch sums u from a sliding register window over the samples
//                   (each sample is loaded once per thread although 2 P (step, tap) pairs use it), then T reverse FFTs
//                   of size N in shared memory (Stockham, mixed radix 4 / 2 / odd primes), written time-major [m][k].
//                   HBM: reads the recording once (neighbouring CTAs share the P N-sample history through L2), writes
//                   8 N bytes per step.
//   resamp_kernel   32-phase arbitrary resampler with the RRC matched filter: out[n] = sum_t (f_j[t] + acc df_j[t]) x[i - t]
//                   with (i, j, acc) from the phase walk, which is the same for all channels (table built once per
//                   plan on the host).  One CTA: 64 channels x 64 outputs; input rows staged in shared memory
//                   (coalesced 512-byte rows of the time-major bank output), two channels per lane, results transposed
//                   through shared memory into channel-major streams - the layout the receive path reads.
//   wide_synth_kernel  the inverse direction for tests and the benchmark: per-ARFCN streams -> one wideband recording
//                   (cubic interpolation to the wideband rate, mixed to +k 31.25 kHz, summed, AWGN).  Workload
//                   construction, not the hot path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "chan.h"
#include "chan_fft.cuh"

namespace gmr1 {

namespace {

constexpr int PFB_TM = 8;                  // steps per work item (register window)
constexpr int PFB_THREADS = 256;

template <int FMT> __device__ __forceinline__ float2 load_wide(const void *x, int64_t i, int64_t n)
{
	if (i < 0 || i >= n)
		return make_float2(0.0f, 0.0f);
	if (FMT == 0)
		return __ldg(reinterpret_cast<const float2 *>(x) + i);
	if (FMT == 2) {
		const char2 v = __ldg(reinterpret_cast<const char2 *>(x) + i);
		return make_float2((float)v.x * (1.0f / 128.0f), (float)v.y * (1.0f / 128.0f));
	}
	const short2 v = __ldg(reinterpret_cast<const short2 *>(x) + i);
	return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
	return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int FMT>
__global__ void __launch_bounds__(PFB_THREADS) pfb_kernel(const PfbArgs a)
{
	extern __shared__ __align__(16) float2 sm[];
	const int N = a.n_chans, D = N >> 1, P = a.taps_per_branch, G = a.groups, T = G * PFB_TM;
	const int tid = threadIdx.x;
	float2 *buf0 = sm, *buf1 = sm + (size_t)T * N;
	const int64_t m_cta = a.m_begin + (int64_t)blockIdx.x * T;

	// ---- branch sums u_p[m] for the CTA's T steps
	for (int item = tid; item < N * G; item += PFB_THREADS) {
		const int p = item % N, g = item / N;
		const int64_t m0 = m_cta + g * PFB_TM;
		float2 xw[PFB_TM], acc[PFB_TM];
#pragma unroll
		for (int i = 0; i < PFB_TM; i++) {
			xw[i] = load_wide<FMT>(a.wide, (m0 + i) * D - p, a.n_wide);
			acc[i] = make_float2(0.0f, 0.0f);
		}
#pragma unroll 1
		for (int q = 0; q < P; q++) {
			const float h = __ldg(&a.taps[p + q * N]);
#pragma unroll
			for (int i = 0; i < PFB_TM; i++) {
				acc[i].x = fmaf(h, xw[i].x, acc[i].x);
				acc[i].y = fmaf(h, xw[i].y, acc[i].y);
			}
			if (q + 1 < P) {               // window of step offsets i - 2 (q + 1): shift by two, two new samples
#pragma unroll
				for (int i = PFB_TM - 1; i >= 2; i--)
					xw[i] = xw[i - 2];
				xw[0] = load_wide<FMT>(a.wide, (m0 - 2 * (q + 1)) * D - p, a.n_wide);
				xw[1] = load_wide<FMT>(a.wide, (m0 + 1 - 2 * (q + 1)) * D - p, a.n_wide);
			}
		}
#pragma unroll
		for (int i = 0; i < PFB_TM; i++)
			buf0[(size_t)(g * PFB_TM + i) * N + p] = acc[i];
	}
	__syncthreads();

	// ---- T reverse FFTs of size N (Stockham autosort; stage radix r, Ns = product of the radices before it)
	float2 *src = buf0, *dst = buf1;
	int Ns = 1;
	for (int st = 0; st < a.n_stage; st++) {
		const int r = a.radix[st], nb = N / r, tw_step = N / (Ns * r);
		for (int item = tid; item < T * nb; item += PFB_THREADS) {
			const int f = item / nb, j = item - f * nb;
			const int k = j % Ns, j0 = (j / Ns) * Ns * r + k;
			const float2 *s = src + (size_t)f * N;
			float2 *d = dst + (size_t)f * N;
			if (r == 4) {
				float2 v0 = s[j], v1 = s[j + nb], v2 = s[j + 2 * nb], v3 = s[j + 3 * nb];
				if (k) {
					v1 = cmul(v1, __ldg(&a.twiddle[k * tw_step]));
					v2 = cmul(v2, __ldg(&a.twiddle[2 * k * tw_step]));
					v3 = cmul(v3, __ldg(&a.twiddle[3 * k * tw_step]));
				}
				const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
				const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
				d[j0] = make_float2(s02.x + s13.x, s02.y + s13.y);
				d[j0 + Ns] = make_float2(d02.x - d13.y, d02.y + d13.x);          // + j d13 (reverse transform)
				d[j0 + 2 * Ns] = make_float2(s02.x - s13.x, s02.y - s13.y);
				d[j0 + 3 * Ns] = make_float2(d02.x + d13.y, d02.y - d13.x);      // - j d13
			} else if (r == 2) {
				float2 v0 = s[j], v1 = s[j + nb];
				if (k)
					v1 = cmul(v1, __ldg(&a.twiddle[k * tw_step]));
				d[j0] = make_float2(v0.x + v1.x, v0.y + v1.y);
				d[j0 + Ns] = make_float2(v0.x - v1.x, v0.y - v1.y);
			} else {                       // odd prime radix (<= CHAN_MAX_RADIX): direct r-point transform
				float2 v[CHAN_MAX_RADIX];
				for (int q = 0; q < r; q++) {
					v[q] = s[j + q * nb];
					if (k && q)
						v[q] = cmul(v[q], __ldg(&a.twiddle[q * k * tw_step]));
				}
				for (int q2 = 0; q2 < r; q2++) {
					float2 o = v[0];
					for (int q = 1; q < r; q++) {
						const float2 w = __ldg(&a.twiddle[((q * q2) % r) * nb]);
						o.x += v[q].x * w.x - v[q].y * w.y;
						o.y += v[q].x * w.y + v[q].y * w.x;
					}
					d[j0 + q2 * Ns] = o;
				}
			}
		}
		__syncthreads();
		float2 *t = src;
		src = dst;
		dst = t;
		Ns *= r;
	}

	// ---- time-major output, odd channels of odd steps negated
	for (int item = tid; item < T * N; item += PFB_THREADS) {
		const int f = item / N, k = item - f * N;
		const int64_t m = m_cta + f;
		if (m >= a.m_end)
			break;
		float2 v = src[item];
		if ((m & 1) && (k & 1))
			v = make_float2(-v.x, -v.y);
		a.mid[m * N + k] = v;
	}
}

// ---- the bank for power-of-two channel counts (64 .. 2048) and 10 taps per branch -----------------------------------
// Same arithmetic as pfb_kernel, a fifth of its instructions (10.8 k -> ~2 k warp-instructions per step at N = 1024):
//   * branch sums: P = 10 is a compile-time constant (firdes.low_pass gives 9.64 N taps for every N), so the
//     26-sample register window of 8 consecutive steps is loaded once and the 80 complex x real MACs are 80 packed
//     FFMA2 with no register moves; the int16 scale rides on the taps; bounds are tested once per CTA;
//   * FFT: register radix-16 butterflies (chan_fft.cuh), N / 16 threads per transform, in place in ONE padded
//     shared-memory buffer (barrier between the loads and the stores of a round of rows), compile-time index
//     arithmetic, the closing stage stores straight to HBM.  [T][N + N / 16] complex floats: 68 KB at N = 1024, three
//     CTAs per SM (the two-buffer generic kernel: one).
constexpr int PFB_P = 10;

__device__ __forceinline__ void fma2p(float2 &c, float h, const float2 x)
{
	unsigned long long cc = *reinterpret_cast<unsigned long long *>(&c);
	const float2 hh = make_float2(h, h);
	asm("fma.rn.f32x2 %0, %1, %2, %0;"
	    : "+l"(cc)
	    : "l"(*reinterpret_cast<const unsigned long long *>(&hh)), "l"(*reinterpret_cast<const unsigned long long *>(&x)));
	c = *reinterpret_cast<float2 *>(&cc);
}

template <int FMT> __device__ __forceinline__ float2 load_wide_raw(const void *x, int64_t i)
{
	if (FMT == 0)
		return __ldg(reinterpret_cast<const float2 *>(x) + i);
	if (FMT == 2) {
		const char2 v = __ldg(reinterpret_cast<const char2 *>(x) + i);
		return make_float2((float)v.x, (float)v.y);      // 1 / 128 is on the taps
	}
	const short2 v = __ldg(reinterpret_cast<const short2 *>(x) + i);
	return make_float2((float)v.x, (float)v.y);          // 1 / 32768 is on the taps
}

template <int LOG2N> struct PfbFast {
	typedef cfft::Plan<LOG2N> FP;
	static constexpr int N = FP::N, D = N / 2;
	static constexpr int T = 4096 / N > PFB_TM ? 4096 / N : PFB_TM;   // steps per CTA
	static constexpr int G = T / PFB_TM;
	static constexpr int RS = cfft::RowStride<N>::value;
	static constexpr int RPR = PFB_THREADS / FP::TPR;   // rows of one round (PFB_THREADS / TPR threads-per-row)
	static constexpr int ROUNDS = T / RPR;
	static constexpr size_t SMEM = (size_t)T * RS * sizeof(float2);
	static_assert(T % RPR == 0 && ROUNDS >= 1, "rows per round");
};

struct TwLdg {                             // e^{+j 2 pi t / N} from the plan's table (L1-resident)
	const float2 *t;
	__device__ __forceinline__ float2 operator()(int i) const { return __ldg(&t[i]); }
};

// one radix-16 stage S (NS = 16^S) over all T rows of the CTA, in place
template <int LOG2N, int S> __device__ __forceinline__ void pfb_stage16(float2 *buf, const TwLdg tw, int tid)
{
	typedef PfbFast<LOG2N> K;
	typedef cfft::Stage<K::N, cfft::Plan<LOG2N>::pow16(S), 16> St;
	const int t = tid % K::FP::TPR, r0 = tid / K::FP::TPR;
#pragma unroll 1
	for (int rd = 0; rd < K::ROUNDS; rd++) {
		float2 *row = buf + (size_t)(rd * K::RPR + r0) * K::RS;
		float2 v[16];
		St::read(row, t, tw, v);
		__syncthreads();                   // every load of this round (and every store of the previous one) is done
		St::write(row, t, v);
	}
	__syncthreads();
}

// closing stage: radix RL (16 / RL butterflies per thread), or the last radix-16 stage; stores to HBM, time-major,
// odd channels of odd steps negated
template <int LOG2N, int R, int NS> __device__ __forceinline__ void pfb_stage_out(const float2 *buf, const TwLdg tw, int tid,
                                                                               const PfbArgs &a, int64_t m_cta)
{
	typedef PfbFast<LOG2N> K;
	typedef cfft::Stage<K::N, NS, R> St;
	const int t = tid % K::FP::TPR, r0 = tid / K::FP::TPR;
#pragma unroll 1
	for (int rd = 0; rd < K::ROUNDS; rd++) {
		const int f = rd * K::RPR + r0;
		const int64_t m = m_cta + f;
		const float2 *row = buf + (size_t)f * K::RS;
		if (m >= a.m_end)
			continue;
		float2 *dst = a.mid + m * K::N;
		const bool odd = m & 1;
#pragma unroll
		for (int i = 0; i < 16 / R; i++) {
			const int j = t + i * K::FP::TPR;
			float2 v[R];
			St::read(row, j, tw, v);
#pragma unroll
			for (int q2 = 0; q2 < R; q2++) {
				const int k = St::out_index(j, q2);
				dst[k] = (odd && (k & 1)) ? make_float2(-v[q2].x, -v[q2].y) : v[q2];
			}
		}
	}
}

template <int LOG2N, int FMT>
__global__ void __launch_bounds__(PFB_THREADS) pfb_fast_kernel(const PfbArgs a)
{
	typedef PfbFast<LOG2N> K;
	constexpr int N = K::N, D = K::D, T = K::T;
	constexpr int W = PFB_TM + 2 * (PFB_P - 1);           // samples of one branch that 8 steps touch
	extern __shared__ __align__(16) float2 buf[];
	const int tid = threadIdx.x;
	const int64_t m_cta = a.m_begin + (int64_t)blockIdx.x * T;
	const bool interior = (m_cta - 2 * (PFB_P - 1)) * D - (N - 1) >= 0 && (m_cta + T - 1) * D < a.n_wide;
	const float scale = FMT == 0 ? 1.0f : FMT == 2 ? 1.0f / 128.0f : 1.0f / 32768.0f;

	// ---- branch sums: item = (branch p, group of 8 steps)
#pragma unroll 1
	for (int item = tid; item < N * K::G; item += PFB_THREADS) {
		const int p = item & (N - 1), g = item >> LOG2N;
		const int64_t m0 = m_cta + g * PFB_TM;
		const int64_t base = (m0 - 2 * (PFB_P - 1)) * D - p;   // sample of window slot 0; slot r is r D further
		float2 xs[W];
		if (interior) {
#pragma unroll
			for (int r = 0; r < W; r++)
				xs[r] = load_wide_raw<FMT>(a.wide, base + (int64_t)r * D);
		} else {
#pragma unroll
			for (int r = 0; r < W; r++) {
				const int64_t i = base + (int64_t)r * D;
				xs[r] = (i >= 0 && i < a.n_wide) ? load_wide_raw<FMT>(a.wide, i) : make_float2(0.0f, 0.0f);
			}
		}
		float2 acc[PFB_TM];
#pragma unroll
		for (int i = 0; i < PFB_TM; i++)
			acc[i] = make_float2(0.0f, 0.0f);
#pragma unroll
		for (int q = 0; q < PFB_P; q++) {
			const float h = __ldg(&a.taps[p + q * N]) * scale;
#pragma unroll
			for (int i = 0; i < PFB_TM; i++)
				fma2p(acc[i], h, xs[i - 2 * q + 2 * (PFB_P - 1)]);
		}
#pragma unroll
		for (int i = 0; i < PFB_TM; i++)
			buf[(size_t)(g * PFB_TM + i) * K::RS + cfft::pad(p)] = acc[i];
	}
	__syncthreads();

	// ---- T reverse FFTs
	const TwLdg tw = {a.twiddle};
	constexpr int N16 = K::FP::N16, RL = K::FP::RLAST;
	if constexpr (RL > 1) {
		if constexpr (N16 >= 1) pfb_stage16<LOG2N, 0>(buf, tw, tid);
		if constexpr (N16 >= 2) pfb_stage16<LOG2N, 1>(buf, tw, tid);
		constexpr int NS = cfft::Plan<LOG2N>::pow16(N16);
		pfb_stage_out<LOG2N, RL, NS>(buf, tw, tid, a, m_cta);
	} else {
		if constexpr (N16 >= 2) pfb_stage16<LOG2N, 0>(buf, tw, tid);
		if constexpr (N16 >= 3) pfb_stage16<LOG2N, 1>(buf, tw, tid);
		constexpr int NS = cfft::Plan<LOG2N>::pow16(N16 - 1);
		pfb_stage_out<LOG2N, 16, NS>(buf, tw, tid, a, m_cta);
	}
}

// ---- arbitrary resampler -------------------------------------------------------------------------------------------
// Tile: 128 channels x 64 outputs per CTA.  A warp takes EIGHT consecutive outputs at a time for four channels per lane:
// the outputs' input spans overlap almost completely (1.5 outputs per input step at sps 4), so two 16-byte loads of an
// input row serve 32 complex x real MACs.  The eight outputs' taps - filter phase j, derivative weight acc, each
// shifted by the output's own newest input step - are merged per group into one table [row][output] in shared memory,
// each value stored twice so that a 16-byte load is two ready-made FFMA2 operands: 6 loads (12 shared-memory
// wavefronts) + 32 FFMA2 per row.  (The first version - one output at a time, two channels per lane - needed 64
// wavefronts for the same MACs and was bound by shared-memory bandwidth; four outputs x two channels: 24.)
// The output tile is transposed through the memory of the input tile.
constexpr int RS_CH = 128, RS_TO = 64, RS_K = 8;   // RS_TO: outputs of the standard tile; plans with long input spans use 32

__device__ __forceinline__ void fma2u(unsigned long long &c, unsigned long long k, unsigned long long x)
{
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(k), "l"(x));
}

__device__ __forceinline__ float2 unpack2(unsigned long long v)
{
	float2 r;
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
	return r;
}

template <int TO>
__global__ void __launch_bounds__(TO / RS_K * 32, 2) resamp_kernel(const ResampArgs a)
{
	constexpr int RS_T = TO / RS_K * 32, RS_TO = TO;       // one warp per group of RS_K outputs
	extern __shared__ __align__(16) float2 sm[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tpf = a.tpf;
	float2 *in_tile = sm;                                   // [rows_max][RS_CH]; later the output tile [RS_CH][RS_TO + 1]
	float2 *out_tile = sm;
	const size_t tile_elems = max((size_t)a.rows_max * RS_CH, (size_t)RS_CH * (RS_TO + 1));
	float *filt = (float *)(sm + ((tile_elems + 1) & ~(size_t)1)); // [32][tpf], then dfilt [32][tpf]
	float *dfilt = filt + 32 * tpf;
	float2 *comb = (float2 *)(dfilt + 32 * tpf) + (size_t)warp * a.span_max * RS_K;   // [warps][span_max][RS_K] (k, k)
	__shared__ int ch_idx[RS_CH];

	const int64_t n0 = a.n_begin + (int64_t)blockIdx.x * RS_TO;
	const int c0 = blockIdx.y * RS_CH;
	const int n_here = (int)min((int64_t)RS_TO, a.n_end - n0);
	const int64_t row_hi = __ldg(&a.sched_i[n0 + n_here - 1]);
	const int64_t row_lo = (int64_t)__ldg(&a.sched_i[n0]) - (tpf - 1);
	const int rows = (int)(row_hi - row_lo + 1);
	const bool whole = a.chan_idx == nullptr && c0 + RS_CH <= a.n_wanted && row_lo >= 0 && row_hi < a.n_steps;
	if (whole) {
		// whole rows of 128 consecutive channels: asynchronous 16-byte copies straight into shared memory (LDGSTS), in
		// flight while the filters are fetched and the tap tables merged; collected in front of the MAC loop
		const float4 *src = reinterpret_cast<const float4 *>(a.mid + row_lo * a.n_chans + c0);
		const unsigned dst = (unsigned)__cvta_generic_to_shared(in_tile);
		const int stride4 = a.n_chans >> 1;
		for (int item = tid; item < rows * (RS_CH / 2); item += RS_T) {
			const int rr = item / (RS_CH / 2), c = item - rr * (RS_CH / 2);
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * item), "l"(src + (size_t)rr * stride4 + c)
			             : "memory");
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	}
	for (int i = tid; i < 32 * tpf; i += RS_T) {
		filt[i] = __ldg(&a.filt[i]);
		dfilt[i] = __ldg(&a.dfilt[i]);
	}
	if (tid < RS_CH)
		ch_idx[tid] = c0 + tid < a.n_wanted ? (a.chan_idx ? __ldg(&a.chan_idx[c0 + tid]) : c0 + tid) : -1;
	__syncthreads();
	if (!whole) {
		for (int item = tid; item < rows * RS_CH; item += RS_T) {
			const int rr = item / RS_CH, c = item - rr * RS_CH;
			const int64_t row = row_lo + rr;
			const int k = ch_idx[c];
			in_tile[item] = (row >= 0 && row < a.n_steps && k >= 0) ? __ldg(&a.mid[row * a.n_chans + k]) : make_float2(0.0f, 0.0f);
		}
	}

	// one group of RS_K outputs per warp (RS_TO / RS_K == warps): channels 2 lane, 2 lane + 1, 64 + 2 lane, 65 + 2 lane
	const int o0 = warp * RS_K;
	const int oo = lane & (RS_K - 1);                       // the output of the group whose taps this lane merges
	unsigned long long s[RS_K][4];
#pragma unroll
	for (int o = 0; o < RS_K; o++)
		s[o][0] = s[o][1] = s[o][2] = s[o][3] = 0ull;
	const int cnt = min(RS_K, n_here - o0);
	if (cnt > 0) {
		const int64_t n = n0 + o0;
		const int b_last = (int)(__ldg(&a.sched_i[n + cnt - 1]) - row_lo);   // newest row any of the outputs reads
		const int span = b_last - (int)(__ldg(&a.sched_i[n]) - row_lo) + tpf; // rows b_last, b_last - 1, ... the group touches
		{
			const bool have = oo < cnt;
			const int j = have ? a.sched_j[n + oo] : 0;
			const float acc = have ? a.sched_acc[n + oo] : 0.0f;
			const int shift = have ? b_last - (int)(__ldg(&a.sched_i[n + oo]) - row_lo) : 0;
			const float *fj = filt + j * tpf, *dj = dfilt + j * tpf;
			for (int e = lane; e < span * RS_K; e += 32) {      // e % RS_K == oo
				const int t = e / RS_K - shift;                 // tap of output oo that meets row b_last - e / RS_K
				const float k = (have && t >= 0 && t < tpf) ? fmaf(acc, dj[t], fj[t]) : 0.0f;
				comb[e] = make_float2(k, k);
			}
			__syncwarp();
		}
	}
	asm volatile("cp.async.wait_all;" ::: "memory");
	__syncthreads();                                            // the input tile is complete
	if (cnt > 0) {
		const int64_t n = n0 + o0;
		const int b_last = (int)(__ldg(&a.sched_i[n + cnt - 1]) - row_lo);
		const int span = b_last - (int)(__ldg(&a.sched_i[n]) - row_lo) + tpf;
		const ulonglong2 *xp = reinterpret_cast<const ulonglong2 *>(in_tile) + (size_t)b_last * (RS_CH / 2) + lane;
		const ulonglong2 *cp = reinterpret_cast<const ulonglong2 *>(comb);
#pragma unroll 2
		for (int r = 0; r < span; r++) {
			const ulonglong2 xa = xp[0], xb = xp[32];
			xp -= RS_CH / 2;
#pragma unroll
			for (int h = 0; h < RS_K / 2; h++) {
				const ulonglong2 k = cp[h];
				fma2u(s[2 * h][0], k.x, xa.x);
				fma2u(s[2 * h][1], k.x, xa.y);
				fma2u(s[2 * h][2], k.x, xb.x);
				fma2u(s[2 * h][3], k.x, xb.y);
				fma2u(s[2 * h + 1][0], k.y, xa.x);
				fma2u(s[2 * h + 1][1], k.y, xa.y);
				fma2u(s[2 * h + 1][2], k.y, xb.x);
				fma2u(s[2 * h + 1][3], k.y, xb.y);
			}
			cp += RS_K / 2;
		}
	}
	__syncthreads();                                            // every warp is done with the input tile
#pragma unroll
	for (int o = 0; o < RS_K; o++)
		if (o < cnt) {
			out_tile[(2 * lane) * (RS_TO + 1) + o0 + o] = unpack2(s[o][0]);
			out_tile[(2 * lane + 1) * (RS_TO + 1) + o0 + o] = unpack2(s[o][1]);
			out_tile[(64 + 2 * lane) * (RS_TO + 1) + o0 + o] = unpack2(s[o][2]);
			out_tile[(65 + 2 * lane) * (RS_TO + 1) + o0 + o] = unpack2(s[o][3]);
		}
	__syncthreads();
	for (int item = tid; item < RS_CH * RS_TO; item += RS_T) {
		const int c = item / RS_TO, o = item - c * RS_TO;
		if (o < n_here && c0 + c < a.n_wanted)
			a.out[(int64_t)(c0 + c) * a.out_stride + n0 + o] = out_tile[c * (RS_TO + 1) + o];
	}
}

// ---- wideband test-signal generator ------------------------------------------------------------------------------
__device__ __forceinline__ void philox2(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
	uint32_t c2 = 0x243f6a88u, c3 = 0x85a308d3u;
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

template <int FMT>
__global__ void __launch_bounds__(256) wide_synth_kernel(const WideSynthArgs a)
{
	extern __shared__ __align__(16) float2 tw[];            // e^{+j 2 pi t / N}
	const int N = a.n_chans;
	for (int i = threadIdx.x; i < N; i += blockDim.x)
		tw[i] = __ldg(&a.twiddle[i]);
	__syncthreads();
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n_wide; t += (int64_t)gridDim.x * blockDim.x) {
		const int64_t pos = t * a.num;
		const int64_t i = pos / a.den;
		const float f = (float)((double)(pos - i * a.den) / (double)a.den);
		const float w0 = -f * (f - 1.0f) * (f - 2.0f) / 6.0f, w1 = (f + 1.0f) * (f - 1.0f) * (f - 2.0f) / 2.0f;
		const float w2 = -(f + 1.0f) * f * (f - 2.0f) / 2.0f, w3 = (f + 1.0f) * f * (f - 1.0f) / 6.0f;
		const int tm = (int)(t % N);
		float xr = 0.0f, xi = 0.0f;
		for (int c = 0; c < a.n_streams; c++) {
			const float2 *s = a.streams + (int64_t)c * a.stream_stride;
			const int k = a.chan_idx ? __ldg(&a.chan_idx[c]) : c;
			float2 v = make_float2(0.0f, 0.0f);
			if (i >= 1 && i + 2 < a.stream_len) {
				const float2 s0 = __ldg(&s[i - 1]), s1 = __ldg(&s[i]), s2 = __ldg(&s[i + 1]), s3 = __ldg(&s[i + 2]);
				v.x = w0 * s0.x + w1 * s1.x + w2 * s2.x + w3 * s3.x;
				v.y = w0 * s0.y + w1 * s1.y + w2 * s2.y + w3 * s3.y;
			} else {
				for (int d = -1; d <= 2; d++) {
					const int64_t ii = i + d;
					if (ii >= 0 && ii < a.stream_len) {
						const float w = d == -1 ? w0 : d == 0 ? w1 : d == 1 ? w2 : w3;
						v.x += w * s[ii].x;
						v.y += w * s[ii].y;
					}
				}
			}
			const float2 ph = tw[(k * tm) % N];
			xr += v.x * ph.x - v.y * ph.y;
			xi += v.x * ph.y + v.y * ph.x;
		}
		if (a.sigma > 0.0f) {
			uint32_t rnd[4];
			philox2((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)a.seed, (uint32_t)(a.seed >> 32), rnd);
			const float u1 = ((float)rnd[0] + 0.5f) * 2.3283064365386963e-10f;
			const float u2 = ((float)rnd[1] + 0.5f) * 2.3283064365386963e-10f;
			const float rad = sqrtf(-2.0f * logf(fmaxf(u1, 1e-30f)));
			float ns, nc;
			sincospif(2.0f * u2, &ns, &nc);
			xr += a.sigma * rad * nc;
			xi += a.sigma * rad * ns;
		}
		xr *= a.gain;
		xi *= a.gain;
		if (FMT == 0) {
			reinterpret_cast<float2 *>(a.wide)[t] = make_float2(xr, xi);
		} else if (FMT == 2) {
			const float sr = fminf(fmaxf(rintf(xr * 128.0f), -128.0f), 127.0f);
			const float si = fminf(fmaxf(rintf(xi * 128.0f), -128.0f), 127.0f);
			reinterpret_cast<char2 *>(a.wide)[t] = make_char2((signed char)sr, (signed char)si);
		} else {
			const float sr = fminf(fmaxf(rintf(xr * 32768.0f), -32768.0f), 32767.0f);
			const float si = fminf(fmaxf(rintf(xi * 32768.0f), -32768.0f), 32767.0f);
			reinterpret_cast<short2 *>(a.wide)[t] = make_short2((short)sr, (short)si);
		}
	}
}

}  // namespace

int pfb_groups(int n_chans)
{
	// T = 8 G steps per CTA; two [T][N] complex buffers in shared memory: aim at <= 64 KB so three CTAs fit an SM
	int g = 4096 / (PFB_TM * n_chans);
	if (g < 1)
		g = 1;
	if (g > 16)
		g = 16;
	return g;
}

template <int LOG2N> cudaError_t launch_pfb_fast(const PfbArgs &a, int fmt, cudaStream_t st, int dev)
{
	typedef PfbFast<LOG2N> K;
	static std::atomic<int> attr_done[64][3];
	auto *fn = fmt == 0 ? pfb_fast_kernel<LOG2N, 0> : fmt == 2 ? pfb_fast_kernel<LOG2N, 2> : pfb_fast_kernel<LOG2N, 1>;
	{
		GMR1_INIT_LOCK();
		if (dev >= 64 || !attr_done[dev][fmt].load()) {
			cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM);
			if (e != cudaSuccess)
				return e;
			if (dev < 64)
				attr_done[dev][fmt].store(1);
		}
	}
	const int64_t grid = (a.m_end - a.m_begin + K::T - 1) / K::T;
	fn<<<(unsigned)grid, PFB_THREADS, K::SMEM, st>>>(a);
	return cudaGetLastError();
}

// which kernel a plan gets: 1 = pfb_fast_kernel (power-of-two bank, 64 .. 2048 channels, 10 taps per branch)
int pfb_is_fast(int n_chans, int taps_per_branch)
{
	return taps_per_branch == PFB_P && n_chans >= 64 && n_chans <= 2048 && (n_chans & (n_chans - 1)) == 0;
}

static std::atomic<int> g_pfb_force_generic{0};
void pfb_force_generic(int on) { g_pfb_force_generic.store(on); }

cudaError_t launch_pfb(const PfbArgs &a0, int fmt, cudaStream_t st)
{
	PfbArgs a = a0;
	if (a.m_end > a.n_steps)
		a.m_end = a.n_steps;
	if (a.m_end <= a.m_begin)
		return cudaSuccess;
	int dev = 0;
	cudaGetDevice(&dev);
	if (pfb_is_fast(a.n_chans, a.taps_per_branch) && !g_pfb_force_generic.load()) {
		switch (a.n_chans) {
		case 64:   return launch_pfb_fast<6>(a, fmt, st, dev);
		case 128:  return launch_pfb_fast<7>(a, fmt, st, dev);
		case 256:  return launch_pfb_fast<8>(a, fmt, st, dev);
		case 512:  return launch_pfb_fast<9>(a, fmt, st, dev);
		case 1024: return launch_pfb_fast<10>(a, fmt, st, dev);
		default:   return launch_pfb_fast<11>(a, fmt, st, dev);
		}
	}
	a.groups = pfb_groups(a.n_chans);
	const int T = a.groups * PFB_TM;
	const size_t smem = 2 * (size_t)T * a.n_chans * sizeof(float2);
	if (smem > 200 * 1024)
		return cudaErrorNotSupported;
	static std::atomic<size_t> attr_max[64][3];
	auto *fn = fmt == 0 ? pfb_kernel<0> : fmt == 2 ? pfb_kernel<2> : pfb_kernel<1>;
	{
		GMR1_INIT_LOCK();
		if (dev < 64 && attr_max[dev][fmt].load() < smem) {
			cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess)
				return e;
			attr_max[dev][fmt].store(smem);
		}
	}
	const int64_t grid = (a.m_end - a.m_begin + T - 1) / T;
	fn<<<(unsigned)grid, PFB_THREADS, smem, st>>>(a);
	return cudaGetLastError();
}

static size_t resamp_smem_to(int to, int rows_max, int span_max, int tpf)
{
	size_t tile = (size_t)rows_max * RS_CH > (size_t)RS_CH * (to + 1) ? (size_t)rows_max * RS_CH : (size_t)RS_CH * (to + 1);
	tile = (tile + 1) & ~(size_t)1;
	return tile * sizeof(float2) + (size_t)64 * tpf * sizeof(float) + (size_t)(to / RS_K) * span_max * RS_K * sizeof(float2);
}

constexpr size_t RS_SMEM_MAX = 110 * 1024;                  // two CTAs per SM

// outputs per tile for a plan whose 64-output tiles span rows64 input rows
int resamp_tile_outputs(int rows64, int span_max, int tpf)
{
	return resamp_smem_to(RS_TO, rows64, span_max, tpf) <= RS_SMEM_MAX ? RS_TO : RS_TO / 2;
}
int resamp_group_outputs() { return RS_K; }

cudaError_t launch_resamp(const ResampArgs &a0, cudaStream_t st)
{
	ResampArgs a = a0;
	if (a.n_end > a.n_out)
		a.n_end = a.n_out;
	if (a.n_end <= a.n_begin || a.n_wanted <= 0)
		return cudaSuccess;
	const int to = a.tile_out;
	if ((to != RS_TO && to != RS_TO / 2) || a.n_begin % to)
		return cudaErrorInvalidValue;
	const size_t smem = resamp_smem_to(to, a.rows_max, a.span_max, a.tpf);
	if (smem > 220 * 1024)
		return cudaErrorNotSupported;
	static std::atomic<size_t> attr_max[64][2];
	int dev = 0;
	cudaGetDevice(&dev);
	auto *fn = to == RS_TO ? resamp_kernel<RS_TO> : resamp_kernel<RS_TO / 2>;
	{
		GMR1_INIT_LOCK();
		if (dev >= 64 || attr_max[dev][to == RS_TO].load() < smem) {
			cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess)
				return e;
			if (dev < 64)
				attr_max[dev][to == RS_TO].store(smem);
		}
	}
	dim3 grid((unsigned)((a.n_end - a.n_begin + to - 1) / to), (unsigned)((a.n_wanted + RS_CH - 1) / RS_CH));
	fn<<<grid, to / RS_K * 32, smem, st>>>(a);
	return cudaGetLastError();
}

cudaError_t launch_wide_synth(const WideSynthArgs &a, int fmt, cudaStream_t st)
{
	if (a.n_wide <= 0)
		return cudaSuccess;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t smem = (size_t)a.n_chans * sizeof(float2);
	const int64_t want = (a.n_wide + 255) / 256;
	const unsigned grid = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
	if (fmt == 0)
		wide_synth_kernel<0><<<grid, 256, smem, st>>>(a);
	else if (fmt == 2)
		wide_synth_kernel<2><<<grid, 256, smem, st>>>(a);
	else
		wide_synth_kernel<1><<<grid, 256, smem, st>>>(a);
	return cudaGetLastError();
}

}  // namespace gmr1
