// fcch_fft.cu - stage 1, third generation of the coarse FCCH search: gmr1_fcch_rough (src/sdr/fcch.c:211-250) for one
// frequency shift or a grid of shifts per window, with the 117-tap correlation done in the frequency domain.
//
// The direct kernels (fcch_kernels.cu, fcch_grid.cu) are bound by arithmetic: 7 606 outputs x 117 taps x 2 FMAs per
// window and shift, 3.6 MFLOP for 247 KB of samples, 37 % of the fp32 rate at best.  The correlation
//     corr_s[m] = sum_n c_s[n] y[m + n],   c_s[n] = r[n] e^{j f_s n}   (r = the real dual chirp, fcch.c:167-193)
// over the normalised, decimated window y (l = 7 722 samples for a 330 ms window) is a circular one of length
// N = 8192 >= l, because every wanted output m < l - 116 ends inside the window:
//     corr_s = IDFT( Y . T_s ),   Y = DFT(y),   T_s[k] = (1 / N) sum_n c_s[n] e^{+2 pi i k n / N}
// T_s depends on the FCCH format and the shift only: built on the host in double, cached on the device.  Per window:
// ONE forward transform, then per shift a product and a reverse transform - 2 transforms instead of 117 taps for a
// single shift (1.1 vs 3.6 MFLOP), 6 instead of 5 x 117 for config 4's grid (3.2 vs 18 MFLOP).
//
// One CTA of 512 threads per window (persistent CTAs, two per SM).  The transforms are the register radix-16 butterflies of chan_fft.cuh (Stockham,
// 16 x 16 x 16 x 2, one radix-16 butterfly per thread and stage, in place in a padded shared-memory row with a barrier
// between the loads and the stores of a stage; the stage twiddles are powers of one table value).  Only reverse transforms are needed: DFT(y) = conj(reverse(conj y)),
// the conjugations ride on the store of the decimated samples and on the product.  The product is fused into the
// loads of the first reverse stage, |corr|^2 into the stores of the last one; the 5-sample energy-window argmax and
// its centroid (osmo_cxvec_peak_energy_find, PEAK_WEIGH_WIN) then run over the energies in shared memory exactly as
// in fcch_grid.cu (oldest-first sums, strict maximum, lowest index on ties).
// Float contract as for the direct kernels: integer TOA equal to the C path except at rounding ties (the transform's
// rounding error is ~1e-6 of the correlation peak, the direct sum's ~1e-6 as well, in a different order).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "chan_fft.cuh"
#include "launch.h"

namespace gmr1 {

namespace {

constexpr int FF_LOG2N = 13, FF_N = 1 << FF_LOG2N;
constexpr int FF_T = FF_N / 16;                    // 512 threads: one radix-16 butterfly each
constexpr int FF_RS = cfft::RowStride<FF_N>::value;
constexpr int FF_PER = FF_N / FF_T;                // 16 consecutive outputs per thread in the peak search
constexpr int FF_MAXLEN = 128;
// Two-block form (overlap-save): the window as two 4096-point blocks, block 1 starting FF_HOP samples into the window,
// each half of the CTA transforming one of them.  4096 = 16^3: three radix-16 stages per transform and no closing
// radix-2 stage - three shared-memory passes instead of four - and 8 % fewer butterflies than one 8192-point transform.
// Block b delivers the correlation outputs [b FF_HOP, b FF_HOP + 4096 - len]; two rows of 4096 + 256 points take
// exactly the shared memory of one row of 8192 + 512.
constexpr int FF_N2 = 4096, FF_HOP = 3712;

struct TwLdg {
	const float2 *t;
	__device__ __forceinline__ float2 operator()(int i) const { return __ldg(&t[i]); }
};

__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// stage S (radix 16, NS = 16^S) of the reverse transform of TN points, in place; u = butterfly of this thread
template <int TN, int S> __device__ __forceinline__ void stage16(float2 *row, const TwLdg tw, int u)
{
	typedef cfft::Stage<TN, cfft::Plan<13>::pow16(S), 16> St;
	float2 v[16];
	St::read(row, u, tw, v);
	__syncthreads();
	St::write(row, u, v);
	__syncthreads();
}

// the same with the outputs handed to f(point index, value) instead of stored to the row
template <int TN, int S, class F> __device__ __forceinline__ void stage16_out(const float2 *row, const TwLdg tw, int u, F f)
{
	typedef cfft::Stage<TN, cfft::Plan<13>::pow16(S), 16> St;
	float2 v[16];
	St::read(row, u, tw, v);
	__syncthreads();                       // the outputs may land in the memory the inputs came from
#pragma unroll
	for (int q = 0; q < 16; q++)
		f(St::out_index(u, q), v[q]);
	__syncthreads();
}

// closing radix-2 stage (NS = 4096): 8 butterflies per thread; f(point index, value) takes the outputs
template <class F> __device__ __forceinline__ void stage2_out(const float2 *row, const TwLdg tw, int tid, F f)
{
	typedef cfft::Stage<FF_N, FF_N / 2, 2> St;
	float2 v[8][2];
#pragma unroll
	for (int i = 0; i < 8; i++)
		St::read(row, tid + i * FF_T, tw, v[i]);
	__syncthreads();                       // the outputs may land in the memory the inputs came from
#pragma unroll
	for (int i = 0; i < 8; i++) {
		f(St::out_index(tid + i * FF_T, 0), v[i][0]);
		f(St::out_index(tid + i * FF_T, 1), v[i][1]);
	}
	__syncthreads();
}

struct FftPlan {
	int32_t n_shifts;
	const float2 *T;                       // [n_shifts][N] tap spectra / N (N = 8192, two-block form: 4096)
	const float2 *tw;                      // [N] e^{+2 pi i t / N}
	float2 *spec;                          // MULTI: [gridDim.x][FF_N] the window's spectrum between the shifts
};

// Persistent CTAs, two per SM: CTA c takes windows c, c + gridDim.x, ...  One padded shared-memory row A for everything.
// MULTI (several shifts per window): the spectrum is parked in the CTA's own 64 KB of global scratch (written once,
// read once per shift with coalesced loads, L2-resident: 296 CTAs x 64 KB) - a second shared-memory row would halve
// the CTAs per SM.  Single shift: the product is taken in place.
template <bool MULTI, bool SPLIT>
__global__ void __launch_bounds__(FF_T, 2)
fcch_fft_kernel(const FcchArgs a, const FftPlan fp, int32_t *toa_out, float *peak_out)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	float2 *A = (float2 *)smem;
	float2 *B = A;
	// SPLIT: block h of the window in row Ah, butterfly u of its transform
	constexpr int TN = SPLIT ? FF_N2 : FF_N, RSH = cfft::RowStride<FF_N2>::value;
	const int h = SPLIT ? tid >> 8 : 0, u = SPLIT ? tid & 255 : tid;
	float2 *Ah = A + h * RSH;
	float *red = (float *)(A + FF_RS);     // [96] reduction scratch
	float2 *spec = MULTI ? fp.spec + (size_t)blockIdx.x * FF_N : nullptr;
	const int L = a.win_len, len = a.len;
	const int l = L >> 2;                  // decimated length (sps 4), <= FF_N
	const int nc = l - len + 1;
	const TwLdg tw = {fp.tw};

#pragma unroll 1
	for (int b = blockIdx.x; b < a.n; b += gridDim.x) {
	if (a.skip && a.skip[b])
		continue;
	__syncthreads();                       // the previous window's energies and reduction scratch are done with
	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	const bool al16 = (((uintptr_t)x) & 15) == 0;
	if (al16) {                            // the whole window -> L2 up front
		const int chunk = 16384, bytes = (L * 8) & ~15;
		for (int o = tid * chunk; o < bytes; o += FF_T * chunk)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)x + o), "r"(min(chunk, bytes - o))
			             : "memory");
	}
	// ---- statistics over ALL samples (sig_normalize averages before decimating); every 4th sample is kept
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	{
		float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
		if (al16) {
			const float4 *x4 = reinterpret_cast<const float4 *>(x);
#pragma unroll 4
			for (int i = tid; i < l; i += FF_T) {
				const float4 v0 = __ldg(&x4[2 * i]), v1 = __ldg(&x4[2 * i + 1]);
				const float2 p0 = make_float2(v0.x, v0.y), p1 = make_float2(v0.z, v0.w), p2 = make_float2(v1.x, v1.y),
				             p3 = make_float2(v1.z, v1.w);
				s2 = __fadd2_rn(s2, __fadd2_rn(__fadd2_rn(p0, p1), __fadd2_rn(p2, p3)));
				q2 = __ffma2_rn(p0, p0, q2);
				q2 = __ffma2_rn(p1, p1, q2);
				q2 = __ffma2_rn(p2, p2, q2);
				q2 = __ffma2_rn(p3, p3, q2);
				// SPLIT: one store per sample as well (a second, predicated one for the overlap kept the compiler from
				// issuing the loads of the unrolled iterations together); the overlap is copied below
				A[SPLIT && i >= FF_N2 ? RSH + cfft::pad(i - FF_HOP) : cfft::pad(i)] = p0;
			}
		} else {
#pragma unroll 1
			for (int i = tid; i < l; i += FF_T)
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const float2 v = __ldg(&x[4 * i + k]);
					s2 = __fadd2_rn(s2, v);
					q2 = __ffma2_rn(v, v, q2);
					if (k == 0)
						A[SPLIT && i >= FF_N2 ? RSH + cfft::pad(i - FF_HOP) : cfft::pad(i)] = v;
				}
		}
		sr = s2.x;
		si = s2.y;
		sq = q2.x + q2.y;
		for (int i = 4 * l + tid; i < L; i += FF_T) {           // L % 4 trailing samples (none is kept)
			const float2 v = __ldg(&x[i]);
			sr += v.x;
			si += v.y;
			sq = fmaf(v.x, v.x, sq);
			sq = fmaf(v.y, v.y, sq);
		}
	}
	sr = wsum(sr);
	si = wsum(si);
	sq = wsum(sq);
	if (lane == 0) {
		red[warp] = sr;
		red[16 + warp] = si;
		red[32 + warp] = sq;
	}
	__syncthreads();
	if (SPLIT) {                           // samples FF_HOP .. 4095 belong to both blocks
		for (int j = tid; j < FF_N2 - FF_HOP; j += FF_T)
			A[RSH + cfft::pad(j)] = A[cfft::pad(FF_HOP + j)];
		__syncthreads();
	}
	sr = wsum(lane < FF_T / 32 ? red[lane] : 0.0f);
	si = wsum(lane < FF_T / 32 ? red[16 + lane] : 0.0f);
	sq = wsum(lane < FF_T / 32 ? red[32 + lane] : 0.0f);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	const float inv_sd = 1.0f / sd;
	// normalised samples, conjugated (the forward transform is run as conj . reverse . conj), zeros behind the window
	for (int i = tid; i < FF_N; i += FF_T) {
		// SPLIT: point i & 4095 of block i >> 12 is sample (i >> 12) FF_HOP + (i & 4095) of the window
		float2 *p = SPLIT ? &A[(i >> 12) * RSH + cfft::pad(i & (FF_N2 - 1))] : &A[cfft::pad(i)];
		const int g = SPLIT ? (i >> 12) * FF_HOP + (i & (FF_N2 - 1)) : i;
		*p = g < l ? make_float2((p->x - ar) * inv_sd, -((p->y - ai) * inv_sd)) : make_float2(0.0f, 0.0f);
	}
	__syncthreads();

	// ---- R = reverse(conj y) = conj(DFT(y)), in place in A (SPLIT: of each block, in its row)
	{
		typedef cfft::Stage<TN, 1, 16> St0;
		float2 v[16];
		St0::read(Ah, u, tw, v);
		__syncthreads();
		St0::write(Ah, u, v);
		__syncthreads();
	}
	stage16<TN, 1>(Ah, tw, u);
	if (SPLIT) {
		if (MULTI)
			stage16_out<TN, 2>(Ah, tw, u, [spec, h](int k, float2 v) { spec[h * FF_N2 + k] = v; });
		else
			stage16_out<TN, 2>(Ah, tw, u, [Ah](int k, float2 v) { Ah[cfft::pad(k)] = v; });
	} else {
		stage16<TN, 2>(Ah, tw, u);
		if (MULTI)
			stage2_out(A, tw, tid, [spec](int k, float2 v) { spec[k] = v; });
		else
			stage2_out(A, tw, tid, [A](int k, float2 v) { A[cfft::pad(k)] = v; });
	}

	// ---- per shift: corr = reverse(conj(R) . T_s), energies, peak
	float *en = (float *)B;                // [4 zeros][FF_N] energies, over the first half of B
#pragma unroll 1
	for (int s = 0; s < fp.n_shifts; s++) {
		const float2 *T = fp.T + (size_t)s * TN;
		float2 *Bh = B + h * RSH;
		{
			typedef cfft::Stage<TN, 1, 16> St0;
			float2 v[16];
			St0::read_ld([Ah, T, spec, h](int i) {
				const float2 r = MULTI ? spec[h * FF_N2 + i] : Ah[cfft::pad(i)], t = __ldg(&T[i]);
				return make_float2(r.x * t.x + r.y * t.y, r.x * t.y - r.y * t.x);      // conj(r) t
			}, u, tw, v);
			if (!MULTI)
				__syncthreads();
			St0::write(Bh, u, v);
			__syncthreads();
		}
		stage16<TN, 1>(Bh, tw, u);
		if (SPLIT) {
			// block 0 delivers the outputs [0, FF_HOP), block 1 the rest; everything from nc on is 0 for the peak search
			const int len1 = FF_N2 - len + 1;              // valid (non-wrapped) outputs of a block
			stage16_out<TN, 2>(Bh, tw, u, [en, nc, h, len1](int k, float2 v) {
				const int m = h * FF_HOP + k;
				if (h ? true : k < FF_HOP)
					en[4 + m] = (m < nc && k < len1) ? v.x * v.x + v.y * v.y : 0.0f;
			});
			for (int m = FF_HOP + FF_N2 + tid; m < FF_N; m += FF_T)
				en[4 + m] = 0.0f;
		} else {
			stage16<TN, 2>(Bh, tw, u);
			stage2_out(B, tw, tid, [en, nc](int k, float2 v) { en[4 + k] = k < nc ? v.x * v.x + v.y * v.y : 0.0f; });
		}
		if (tid < 4)
			en[tid] = 0.0f;
		__syncthreads();

		// 5-sample energy windows ending at m = 16 tid .. 16 tid + 15: oldest first, as the reference sums
		float xw[FF_PER + 4];
		{
			const float4 *e4 = reinterpret_cast<const float4 *>(en + FF_PER * tid);
#pragma unroll
			for (int q = 0; q < (FF_PER + 4) / 4; q++) {
				const float4 v = e4[q];
				xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
			}
		}
		float bv = 0.0f;
		int bi = 0x7fffffff;
		const int m0 = FF_PER * tid;
#pragma unroll
		for (int j = 0; j < FF_PER; j++) {
			const float val = ((((0.0f + xw[j]) + xw[j + 1]) + xw[j + 2]) + xw[j + 3]) + xw[j + 4];
			if (val > bv && m0 + j < nc) {
				bv = val;
				bi = m0 + j;
			}
		}
		const float my_v = bv;
		const int my_i = bi;
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
		}
		int *redi = (int *)(red + 48);
		if (lane == 0) { red[warp] = bv; redi[warp] = bi; }
		__syncthreads();
		bv = lane < FF_T / 32 ? red[lane] : 0.0f;
		bi = lane < FF_T / 32 ? redi[lane] : 0x7fffffff;
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
		}
		const size_t o = (size_t)s * a.n + b;
		if (bv <= 0.0f) {                  // nothing correlated: position 0 (max_idx = 0, empty centroid)
			if (tid == 0) {
				toa_out[o] = 0;
				if (peak_out) peak_out[o] = bv;
			}
		} else if (my_i == bi && my_v == bv) {                 // exactly one thread owns the winning window
			// centroid window: [idx-4, idx], or [0, 5) when that would start in front of the vector
			const int max_idx = bi - 4 < 0 ? 0 : bi - 4;
			float mw = 0.0f, sw = 0.0f;
#pragma unroll
			for (int k2 = 0; k2 < 5; k2++) {
				const float e = en[4 + max_idx + k2];
				sw += e;
				mw += e * (float)(max_idx + k2);
			}
			const float pos = sw > 0.0f ? mw / sw : (float)max_idx;
			toa_out[o] = (int)round((double)(pos * 4.0f));
			if (peak_out) peak_out[o] = bv;
		}
		__syncthreads();                   // the energies are read before the next shift overwrites B
	}
	}
}

// ---- tap spectra, cached per device -------------------------------------------------------------------------------
struct SpecKey {
	int dev, len, n_shifts, tn;             // tn: transform size (8192, two-block form 4096)
	float freq;
	float shifts[16];
	bool operator==(const SpecKey &o) const
	{
		if (dev != o.dev || len != o.len || n_shifts != o.n_shifts || tn != o.tn || freq != o.freq)
			return false;
		for (int i = 0; i < n_shifts; i++)
			if (shifts[i] != o.shifts[i])
				return false;
		return true;
	}
};
struct SpecEntry {
	SpecKey key;
	float2 *T;
};
std::vector<SpecEntry> g_spec;             // under GMR1_INIT_LOCK
float2 *g_tw[2][64];                      // [two-block form][device]
int     g_ctas[64];                        // persistent CTAs per launch: 2 x SMs

cudaError_t spectra(const SpecKey &key, const float2 **T, const float2 **tw)
{
	GMR1_INIT_LOCK();
	if (key.dev < 0 || key.dev >= 64)
		return cudaErrorInvalidDevice;
	const int TN = key.tn, w = TN == FF_N ? 0 : 1;
	if (!g_tw[w][key.dev]) {
		std::vector<float2> h(TN);
		for (int t = 0; t < TN; t++)
			h[t] = make_float2((float)cos(2.0 * M_PI * t / TN), (float)sin(2.0 * M_PI * t / TN));
		cudaError_t e = cudaMalloc((void **)&g_tw[w][key.dev], TN * sizeof(float2));
		if (e != cudaSuccess)
			return e;
		if ((e = cudaMemcpy(g_tw[w][key.dev], h.data(), TN * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess)
			return e;
	}
	*tw = g_tw[w][key.dev];
	for (const SpecEntry &e : g_spec)
		if (e.key == key) {
			*T = e.T;
			return cudaSuccess;
		}
	// dual-chirp reference at 1 sample/symbol as the reference computes it in float (fcch.c:182-190), times e^{j f n},
	// transformed in double
	std::vector<double> wr(TN), wi(TN);
	for (int t = 0; t < TN; t++) {
		wr[t] = cos(2.0 * M_PI * t / TN);
		wi[t] = sin(2.0 * M_PI * t / TN);
	}
	std::vector<float2> h((size_t)key.n_shifts * TN);
	const float phase_base = key.freq * 2.0f * 3.14159265358979323846264338327f / (float)key.len;
	const float halfpos = (float)key.len / 2.0f;
	for (int s = 0; s < key.n_shifts; s++) {
		std::vector<double> cr(key.len), ci(key.len);
		for (int n = 0; n < key.len; n++) {
			const float pos = (float)n - halfpos;
			const float r = sqrtf(2.0f) * cosf(phase_base * (pos * pos));
			const float ang = key.shifts[s] * (float)n;         // float product, as the sample rotation forms it
			cr[n] = (double)r * cos((double)ang);
			ci[n] = (double)r * sin((double)ang);
		}
		for (int k = 0; k < TN; k++) {
			double tr = 0.0, ti = 0.0;
			for (int n = 0; n < key.len; n++) {
				const int idx = (int)(((int64_t)k * n) & (TN - 1));
				tr += cr[n] * wr[idx] - ci[n] * wi[idx];
				ti += cr[n] * wi[idx] + ci[n] * wr[idx];
			}
			h[(size_t)s * TN + k] = make_float2((float)(tr / TN), (float)(ti / TN));
		}
	}
	float2 *d = nullptr;
	cudaError_t e = cudaMalloc((void **)&d, h.size() * sizeof(float2));
	if (e != cudaSuccess)
		return e;
	if ((e = cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) {
		cudaFree(d);
		return e;
	}
	if (g_spec.size() >= 32) {             // a caller sweeping shifts: drop the oldest table
		cudaFree(g_spec.front().T);
		g_spec.erase(g_spec.begin());
	}
	g_spec.push_back({key, d});
	*T = d;
	return cudaSuccess;
}

std::atomic<int> g_fft_off{0}, g_fft_one_block{0};

}  // namespace

// 0: off (direct kernels), 1: on, 2: on, always one 8192-point block
void fcch_fft_enable(int on)
{
	g_fft_off.store(on ? 0 : 1);
	g_fft_one_block.store(on == 2 ? 1 : 0);
}

// shifts == NULL: one search per window with the uniform shift a.freq_shift0, results to a.toa / a.peak.  Else n_shifts
// searches per window, results to toa / peak [n_shifts][n].  cudaErrorNotSupported: geometry or arguments outside what
// this kernel covers (per-window shifts, sps != 4, windows beyond 8192 symbols) - the caller falls back to the
// direct kernels.
cudaError_t launch_fcch_fft(const FcchArgs &a, const float *shifts, int n_shifts, int32_t *toa, float *peak, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	const int l = a.win_len / 4, nc = l - a.len + 1;
	static const bool env_off = [] { const char *e = getenv("GMR1B200_FCCH_FFT"); return e && atoi(e) == 0; }();   // A/B knob
	if (env_off || g_fft_off.load() || a.sps != 4 || l > FF_N || a.len > FF_MAXLEN || nc < 8 || a.en_out || a.freq_shift)
		return cudaErrorNotSupported;
	SpecKey key = {};
	cudaError_t e = cudaGetDevice(&key.dev);
	if (e != cudaSuccess)
		return e;
	// two-block form when the window fits two overlapping 4096-point blocks (the standard 330 ms window does)
	static const bool env_one = [] { const char *e = getenv("GMR1B200_FCCH_FFT_SPLIT"); return e && atoi(e) == 0; }();   // A/B knob
	const bool split = !env_one && !g_fft_one_block.load() && l <= FF_HOP + FF_N2 && l > FF_N2 && nc <= FF_HOP + (FF_N2 - a.len + 1);
	key.tn = split ? FF_N2 : FF_N;
	key.len = a.len;
	key.freq = a.freq;
	if (!shifts) {
		key.n_shifts = 1;
		key.shifts[0] = a.freq_shift0;
		toa = a.toa;
		peak = a.peak;
	} else {
		if (n_shifts < 1 || n_shifts > 16 || a.freq_shift0 != 0.0f)
			return cudaErrorNotSupported;
		key.n_shifts = n_shifts;
		for (int k = 0; k < n_shifts; k++)
			key.shifts[k] = shifts[k];
	}
	FftPlan fp = {};
	fp.n_shifts = key.n_shifts;
	if ((e = spectra(key, &fp.T, &fp.tw)) != cudaSuccess)
		return e;
	const bool multi = key.n_shifts > 1;
	const size_t smem = (size_t)FF_RS * sizeof(float2) + 96 * sizeof(float);
	int ctas = 0;
	{
		GMR1_INIT_LOCK();
		if (!g_ctas[key.dev]) {
			int sms = 148;
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, key.dev);
			g_ctas[key.dev] = 2 * sms;
		}
		ctas = g_ctas[key.dev];
		static bool attr_set[64][4];
		const void *fns[4] = {(const void *)fcch_fft_kernel<false, false>, (const void *)fcch_fft_kernel<true, false>,
		                      (const void *)fcch_fft_kernel<false, true>, (const void *)fcch_fft_kernel<true, true>};
		const int fi = (multi ? 1 : 0) + (split ? 2 : 0);
		if (key.dev >= 64 || !attr_set[key.dev][fi]) {
			if ((e = cudaFuncSetAttribute(fns[fi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
				return e;
			if (key.dev < 64)
				attr_set[key.dev][fi] = true;
		}
	}
	// several shifts per window: persistent CTAs (the spectrum parking is per CTA); one shift: a CTA per window, the
	// hardware balances the tail (1024 windows over 296 persistent CTAs would run at 3.46 / 4)
	const int grid = (multi && a.n > ctas) ? ctas : a.n;
	if (multi) {                           // spectrum parking of this launch: stream-ordered, back to the pool right behind it
		if ((e = cudaMallocAsync((void **)&fp.spec, (size_t)grid * FF_N * sizeof(float2), st)) != cudaSuccess)
			return e;
		if (split)
			fcch_fft_kernel<true, true><<<grid, FF_T, smem, st>>>(a, fp, toa, peak);
		else
			fcch_fft_kernel<true, false><<<grid, FF_T, smem, st>>>(a, fp, toa, peak);
		e = cudaGetLastError();
		cudaFreeAsync(fp.spec, st);
		return e;
	}
	if (split)
		fcch_fft_kernel<false, true><<<grid, FF_T, smem, st>>>(a, fp, toa, peak);
	else
		fcch_fft_kernel<false, false><<<grid, FF_T, smem, st>>>(a, fp, toa, peak);
	return cudaGetLastError();
}

}  // namespace gmr1
