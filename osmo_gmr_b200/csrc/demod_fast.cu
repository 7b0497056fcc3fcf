// demod_fast.cu - stage 2 of the receive path, the per-format kernels: batched pi/4-CxPSK burst demodulation with the
// burst format (chunk positions, tap counts, reference symbols, data-symbol layout) and the search width as
// COMPILE-TIME constants.  One warp per burst, persistent warps, the same restructured algorithm as the generic kernel
// (demod_kernels.cu, which stays the path for detect, sps != 4, non-standard search widths and RACH) - what changes is
// the instruction count: the generic kernel spends ~2 090 warp-instructions on a BCCH burst, more than half of them on
// descriptor loads, loop control, padded taps and per-chunk bookkeeping; here every loop over chunks / taps / rows /
// data-symbol rows is unrolled against the constexpr format table (burst_formats.h) and the per-lane roles (which
// training symbol, which region slot, which segment of the chunk sums) are computed once per kernel, not per burst.
//
// Replaces, for a whole batch per launch, the reference's
//   gmr1_pi4cxpsk_demod       src/sdr/pi4cxpsk.c:520-602
//   _gmr1_pi4cxpsk_sync_find  :184-268   (incl. the never-reset accumulator quirk, :207/:232)
//   _gmr1_pi4cxpsk_align      :280-297   (sps >= 4 branch)
//   _gmr1_pi4cxpsk_freq_err   :360-406
//   _gmr1_pi4cxpsk_phase      :415-433
//   _gmr1_pi4cxpsk_soft_symbols / _soft_bits  :442-503
// Same float contract as the generic kernel (tests/test_demod_gpu.py).
//
// Per burst (BCCH numbers: 1016-sample window, 81 search offsets, 17 training symbols in 3 chunks, 212 data symbols):
//   1. the window is requested into L2 with one bulk prefetch; the statistics pass reads it once with 16-byte loads
//      (8 in flight per lane), sums it with packed adds and stores the samples of the three correlation regions
//      (~310 samples) to the warp's shared-memory slice - which pairs go where is decided at compile time;
//   2. training-sequence correlation over all search offsets: lanes <-> offsets (m, m+32, m+64), exact tap counts,
//      rotated taps cached per warp, two packed FFMA2 per tap and offset, |corr| per chunk summed in registers;
//   3. coarse peak (redux argmax) + early/late search (radix-8, demod_common.cuh);
//   4. training symbols one per lane from the regions, chunk sums by one segmented shuffle reduction, frequency
//      error from the chunk-to-chunk angles, phase reference;
//   5. data symbols in the angle domain, one batch of DB rows of 32 symbols with all sample loads (L2 hits) in
//      flight, soft bits from the 1/256-cell table.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <type_traits>
#include <utility>

#include "burst_formats.h"
#include "demod_common.cuh"
#include "launch.h"

#ifndef DMF_PREFETCH
#define DMF_PREFETCH 1         // 0: none, 1: bulk L2 prefetch of the window at burst start, 2: per-lane L2 prefetches
#endif
#ifndef DMF_PF_FROM
#define DMF_PF_FROM 4096       // DMF_PREFETCH 2: first byte of the window that is prefetched
#endif
#ifndef DMF_CACHE_HINTS
#define DMF_CACHE_HINTS 1      // bit 0: data-symbol loads (last use of the window) streaming, bit 1: soft-bit stores streaming
#endif
#ifndef DMF_CORR_PIPE
#define DMF_CORR_PIPE 0        // correlation: samples of the next tap requested before the multiply-adds of this one
#endif
#ifndef DMF_MIN_CTAS
#define DMF_MIN_CTAS 8         // resident CTAs per SM the register allocation is capped for
#endif
#ifndef DMF_STATS_BATCH
#define DMF_STATS_BATCH 4      // 16-byte loads of the statistics pass issued together
#endif

namespace gmr1 {

// ---- compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N-1>)
template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
	(f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
	static_for_impl(f, std::make_integer_sequence<int, N>{});
}

// ---- format geometry at sps 4 (all constexpr; bt = BurstId, w = number of search offsets) -------------------------
constexpr int FG_SPS = 4;
constexpr int fg_cpos(int bt, int c) { return bf_s_pos(bt, 0, c); }
constexpr int fg_clen(int bt, int c) { return bf_s_len(bt, 0, c); }
constexpr int fg_nch(int bt) { return bf_n_chunk(bt, 0); }
constexpr int fg_cstart(int bt, int c)          // first training symbol of chunk c (c == nch: their number)
{
	int a = 0;
	for (int i = 0; i < c; i++)
		a += fg_clen(bt, i);
	return a;
}
constexpr int fg_L(int bt, int w) { return BURSTS[bt].len * FG_SPS + w - 1; }
// correlation region of chunk c: window samples [lo, hi) - everything the search (offsets 0..w-1), the symbol pick
// (d = round(toa) in -1..w) read of the chunk; whole pairs of samples
constexpr int fg_rlo(int bt, int c) { return fg_cpos(bt, c) * FG_SPS - 2; }
constexpr int fg_rhi(int bt, int w, int c)
{
	int h = fg_cpos(bt, c) * FG_SPS + (fg_clen(bt, c) - 1) * FG_SPS + w + 1;
	h = (h + 1) & ~1;
	return h > fg_L(bt, w) ? fg_L(bt, w) : h;
}
constexpr int fg_roff(int bt, int w, int c)      // offset of the region in the warp's buffer (c == nch: total)
{
	int a = 0;
	for (int i = 0; i < c; i++)
		a += fg_rhi(bt, w, i) - fg_rlo(bt, i);
	return a;
}
constexpr int fg_toff(int bt, int c)             // first tap of chunk c in a sequence's tap array (even; c == nch: total)
{
	int a = 0;
	for (int i = 0; i < c; i++)
		a += (fg_clen(bt, i) + 1) & ~1;
	return a;
}
constexpr int fg_nds(int bt)
{
	int a = 0;
	for (int c = 0; c < bf_n_data(bt); c++)
		a += BURSTS[bt].data[c].len;
	return a;
}
constexpr int fg_dpos(int bt, int t)             // burst position of data symbol t (clamped to the last one)
{
	int a = 0, pos = 0;
	for (int c = 0; c < bf_n_data(bt); c++) {
		if (t >= a)
			pos = BURSTS[bt].data[c].pos + (t - a < BURSTS[bt].data[c].len ? t - a : BURSTS[bt].data[c].len - 1);
		a += BURSTS[bt].data[c].len;
	}
	return pos;
}
constexpr int fg_dmax(int bt)
{
	int m = 0;
	for (int c = 0; c < bf_n_data(bt); c++)
		if (BURSTS[bt].data[c].pos + BURSTS[bt].data[c].len - 1 > m)
			m = BURSTS[bt].data[c].pos + BURSTS[bt].data[c].len - 1;
	return m;
}
// the per-format kernels require: every sync sequence has the chunk layout of sequence 0, at most 32 training
// symbols (or one sequence of chunks of at most 32 symbols and at most 8 search offsets: RACH), regions disjoint, in
// ascending order and inside the window, every symbol pick inside the window
constexpr bool fg_ok(int bt, int w)
{
	const int nch = fg_nch(bt);
	if (w < 1 || w > 96 || (fg_L(bt, w) & 1) || fg_cstart(bt, nch) < 1)
		return false;
	if (fg_cstart(bt, nch) > 32) {          // RACH: one row of lanes per chunk, quarter-chunks per lane in the search
		if (bf_n_sync(bt) != 1 || w > 8)
			return false;
		for (int c = 0; c < nch; c++)
			if (fg_clen(bt, c) > 32)
				return false;
	}
	for (int s = 1; s < bf_n_sync(bt); s++) {
		if (bf_n_chunk(bt, s) != nch)
			return false;
		for (int c = 0; c < nch; c++)
			if (bf_s_pos(bt, s, c) != fg_cpos(bt, c) || bf_s_len(bt, s, c) != fg_clen(bt, c))
				return false;
	}
	for (int c = 0; c < nch; c++) {
		if (fg_rlo(bt, c) < 0 || fg_rhi(bt, w, c) <= fg_rlo(bt, c))
			return false;
		if (c && fg_rlo(bt, c) < fg_rhi(bt, w, c - 1))
			return false;
		if (fg_cpos(bt, c) * FG_SPS + (fg_clen(bt, c) - 1) * FG_SPS + w >= fg_L(bt, w))
			return false;
	}
	return fg_dmax(bt) * FG_SPS + w < fg_L(bt, w) && fg_nds(bt) * BURSTS[bt].nbits == BURSTS[bt].ebits;
}

template <int BT, int W_>
struct Geo {
	static constexpr int W = W_, NB = BURSTS[BT].nbits, EBITS = BURSTS[BT].ebits;
	static constexpr int L = fg_L(BT, W_);
	static constexpr int ROWS = (W_ + 31) / 32;
	static constexpr int NSYNC = bf_n_sync(BT), NCH = fg_nch(BT), NTR = fg_cstart(BT, fg_nch(BT));
	// more than 32 training symbols (RACH, 99 in five chunks): the training-symbol phase takes one row of lanes per
	// chunk, and the search splits every chunk over the four lanes of a quad (8 quads <-> up to 8 search offsets)
	static constexpr bool TROWS = fg_cstart(BT, fg_nch(BT)) > 32;
	static constexpr int TR = TROWS ? fg_nch(BT) : 1;
	static constexpr int REG_TOTAL = fg_roff(BT, W_, fg_nch(BT));
	// the never-stored search offsets W .. 32*ROWS-1 read this far past the end of a region
	static constexpr int REG_ALLOC = (REG_TOTAL + 32 * ROWS - W_ + 2 + 1) & ~1;
	static constexpr int TAPS = fg_toff(BT, fg_nch(BT));
	static constexpr int NDS = fg_nds(BT), DROWS = (fg_nds(BT) + 31) / 32;
	static constexpr int DITER = (DROWS + 7) / 8;                   // data symbols: DITER passes of DB rows
	static constexpr int DB = (DROWS + DITER - 1) / DITER;
	static constexpr int ACC_ALLOC = 4 + 32 * ROWS + 8;             // accv[-4 .. 32*ROWS+3]
	static constexpr int WARP_BYTES = (REG_ALLOC * 8 + NSYNC * TAPS * 8 + ((NSYNC * NCH + 1) & ~1) * 8 + ACC_ALLOC * 4 +
	                                   32 * 4 + 15) & ~15;
	static_assert(fg_ok(BT, W_), "burst format / search width not eligible for the per-format kernel");
};

struct FastSmem {
	float2 *reg;     // [REG_ALLOC] raw samples of the correlation regions, chunk after chunk
	float2 *taps;    // [NSYNC][TAPS] rotated reference taps
	float2 *tsum;    // [NSYNC][NCH]  sum of each chunk's taps
	float  *accv;    // [-4 .. 32*ROWS+3] correlation magnitude accumulator, zero outside [0, W)
	float  *aw;      // [28] early/late search window
	float  *fs_taps; // frequency shift the cached taps were built for
};

template <class G>
__device__ __forceinline__ FastSmem fast_carve(uint8_t *base)
{
	FastSmem s;
	s.reg = (float2 *)base;
	base += G::REG_ALLOC * 8;
	s.taps = (float2 *)base;
	base += G::NSYNC * G::TAPS * 8;
	s.tsum = (float2 *)base;
	base += ((G::NSYNC * G::NCH + 1) & ~1) * 8;
	s.accv = (float *)base + 4;
	base += G::ACC_ALLOC * 4;
	s.aw = (float *)base;
	s.fs_taps = s.aw + 28;
	return s;
}

// ---- 1. window statistics + correlation-region fill (16-byte aligned window) ---------------------------------------
struct FNorm { float ar, ai, inv_sd; };

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int BT, class G, bool WANT_SD>
__device__ __forceinline__ FNorm stats_fill(const float2 *__restrict__ x, int lane, float2 *reg)
{
	constexpr int NP = G::L / 2, NIT = (NP + 31) / 32, BATCH = DMF_STATS_BATCH;
	const float4 *x4 = reinterpret_cast<const float4 *>(x) + lane;
	float4 *reg4 = reinterpret_cast<float4 *>(reg) + lane;
	float2 s0 = make_float2(0.0f, 0.0f), s1 = s0, q0 = s0, q1 = s0;
	static_for<(NIT + BATCH - 1) / BATCH>([&](auto BI) {
		constexpr int k0 = decltype(BI)::value * BATCH;
		constexpr int nb = NIT - k0 < BATCH ? NIT - k0 : BATCH;
		float4 v[nb];
		static_for<nb>([&](auto J) {
			constexpr int j = decltype(J)::value, k = k0 + j;
			if constexpr (32 * k + 32 <= NP)
				v[j] = __ldg(x4 + 32 * k);
			else
				v[j] = lane < NP - 32 * k ? __ldg(x4 + 32 * k) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		});
		static_for<nb>([&](auto J) {
			constexpr int j = decltype(J)::value, k = k0 + j;
			// which correlation regions do pairs 32k .. 32k+31 belong to?  (compile time)
			static_for<G::NCH>([&](auto C) {
				constexpr int c = decltype(C)::value;
				constexpr int plo = fg_rlo(BT, c) / 2, phi = fg_rhi(BT, G::W, c) / 2;
				constexpr int lo = plo - 32 * k > 0 ? plo - 32 * k : 0, hi = phi - 32 * k < 32 ? phi - 32 * k : 32;
				if constexpr (lo < hi) {
					constexpr int dst = fg_roff(BT, G::W, c) / 2 + 32 * k - plo;
					if constexpr (lo == 0 && hi == 32)
						reg4[dst] = v[j];
					else if constexpr (lo == 0) {
						if (lane < hi) reg4[dst] = v[j];
					} else if constexpr (hi == 32) {
						if (lane >= lo) reg4[dst] = v[j];
					} else {
						if (lane >= lo && lane < hi) reg4[dst] = v[j];
					}
				}
			});
			s0 = fadd2(s0, make_float2(v[j].x, v[j].y));
			s1 = fadd2(s1, make_float2(v[j].z, v[j].w));
			if constexpr (WANT_SD) {
				q0 = ffma2(make_float2(v[j].x, v[j].y), make_float2(v[j].x, v[j].y), q0);
				q1 = ffma2(make_float2(v[j].z, v[j].w), make_float2(v[j].z, v[j].w), q1);
			}
		});
	});
	s0 = fadd2(s0, s1);
	const float2 sum = warp_sum2(s0.x, s0.y, lane);
	constexpr float inv_l = 1.0f / (float)G::L;
	FNorm n;
	n.ar = sum.x * inv_l;
	n.ai = sum.y * inv_l;
	n.inv_sd = 1.0f;
	if constexpr (WANT_SD) {
		// sync power only: the scale 1/stddev changes no decision and no soft bit
		q0 = fadd2(q0, q1);
		const float sq = warp_sum(q0.x + q0.y);
		const float var = sq * inv_l - (n.ar * n.ar + n.ai * n.ai);
		float sd = var > 0.0f ? sqrtf(var) : 0.0f;
		if (sd == 0.0f)
			sd = 1.0f;
		n.inv_sd = 1.0f / sd;
	}
	return n;
}

// window at an odd sample offset (8-byte aligned only): cold
template <int BT, class G>
__device__ __noinline__ FNorm stats_fill_unaligned(const float2 *__restrict__ x, int lane, float2 *reg, bool want_sd)
{
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
#pragma unroll 1
	for (int i = lane; i < G::L; i += 32) {
		const float2 v = __ldg(&x[i]);
		sr += v.x;
		si += v.y;
		sq += v.x * v.x + v.y * v.y;
		static_for<G::NCH>([&](auto C) {
			constexpr int c = decltype(C)::value;
			if (i >= fg_rlo(BT, c) && i < fg_rhi(BT, G::W, c))
				reg[fg_roff(BT, G::W, c) + i - fg_rlo(BT, c)] = v;
		});
	}
	const float2 sum = warp_sum2(sr, si, lane);
	FNorm n;
	n.ar = sum.x / (float)G::L;
	n.ai = sum.y / (float)G::L;
	n.inv_sd = 1.0f;
	if (want_sd) {
		sq = warp_sum(sq);
		const float var = sq / (float)G::L - (n.ar * n.ar + n.ai * n.ai);
		float sd = var > 0.0f ? sqrtf(var) : 0.0f;
		if (sd == 0.0f)
			sd = 1.0f;
		n.inv_sd = 1.0f / sd;
	}
	return n;
}

// ---- rotated reference taps for the frequency shift fs (rad/sample): t_n = conj(ref_n) e^{j*fs*sps*n}, per chunk,
// and their sums.  Rebuilt only when fs changes (never, when the batch shares one freq_shift): cold.
template <int BT, class G>
__device__ __noinline__ void fast_build_taps(const FastSmem sm, float fs, int lane)
{
	__syncwarp();
	const float2 rot = sincos_acc((fs * (float)FG_SPS) * (float)lane);      // e^{j*fs*sps*lane}: tap `lane` of every chunk
	static_for<G::NSYNC>([&](auto S) {
		constexpr int s = decltype(S)::value;
		static_for<G::NCH>([&](auto C) {
			constexpr int c = decltype(C)::value, cl = fg_clen(BT, c);
			// symbols of this chunk, 2 bits each
			unsigned long long symw = 0;
			static_for<cl>([&](auto K) { symw |= (unsigned long long)bf_s_sym(BT, s, c, decltype(K)::value) << (2 * decltype(K)::value); });
			float2 t = make_float2(0.0f, 0.0f);
			if (lane < cl)
				t = mul_conj_sym((int)((symw >> (2 * lane)) & 3), rot);
			if (lane < ((cl + 1) & ~1))
				sm.taps[s * G::TAPS + fg_toff(BT, c) + lane] = t;
			const float2 Rs = warp_sum2(t.x, t.y, lane);
			if (lane == 0)
				sm.tsum[s * G::NCH + c] = Rs;
		});
	});
	__syncwarp();
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
template <int BT, int W, bool WANT_SD>
__global__ void __launch_bounds__(DM_WARPS * 32, DMF_MIN_CTAS)
demod_fast_kernel(const DemodArgs a)
{
	using G = Geo<BT, W>;
	constexpr int NB = G::NB, ROWS = G::ROWS, NCH = G::NCH, NTR = G::NTR, DB = G::DB;
	extern __shared__ __align__(16) uint8_t smem[];
	__shared__ uint16_t soft_lut[LUT_CELLS << NB];
	__shared__ uint2 dtab[G::DROWS * 32];          // data symbol -> (byte offset of its sample for TOA 0, position as float)
	__shared__ uint4 lane_tab[G::TR * 32];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const FastSmem sm = fast_carve<G>(smem + (size_t)warp * G::WARP_BYTES);

	for (int k = threadIdx.x; k < (LUT_CELLS << NB); k += blockDim.x)
		soft_lut[k] = (uint16_t)soft_word<NB>(((float)k + 0.5f) * (1.0f / LUT_CELLS));
	for (int t = threadIdx.x; t < G::DROWS * 32; t += blockDim.x) {
		int acc = 0, pos = 0;
		static_for<bf_n_data(BT)>([&](auto C) {
			constexpr int c = decltype(C)::value;
			constexpr int dp = BURSTS[BT].data[c].pos, dl = BURSTS[BT].data[c].len;
			if (t >= acc)
				pos = dp + min(t - acc, dl - 1);       // symbols past the last one repeat it (loaded, never stored)
			acc += dl;
		});
		dtab[t] = make_uint2((unsigned)(pos * FG_SPS * 8), __float_as_uint((float)pos));
	}
	// the warp's slice only ever holds finite values (the search offsets >= W read whatever follows a region)
	for (int i = lane; i < G::WARP_BYTES / 4; i += 32)
		reinterpret_cast<float *>(smem + (size_t)warp * G::WARP_BYTES)[i] = 0.0f;
	__syncthreads();

	// ---- the role of a lane in the training-symbol phase (training symbol t = lane), computed once per CTA into shared
	// memory and read back with one 16-byte load per burst: kept in registers across the burst loop these values get
	// spilled or rematerialised (~60 instructions per burst, measured)
	//   .x  byte offset of the symbol's sample for d = 0 in the region buffer
	//   .y  burst position of the symbol (float)
	//   .z  [7:0] end of the symbol's chunk (segmented sum)  [15:8] / [23:16] lane c >= 1: first symbol of chunk c / c-1
	//       [24 + 2s +: 2] reference symbol of sequence s
	//   .w  lane c >= 1: distance between the centres of chunks c and c-1 (float)
	if constexpr (G::TROWS) {
		// one row of lanes per chunk: entry [c][lane] = symbol `lane` of chunk c (.x, .y as above, .z its reference symbol)
		for (int t = threadIdx.x; t < G::TR * 32; t += blockDim.x) {
			const int cc = t >> 5, l = t & 31;
			uint4 ent = make_uint4(0u, 0u, 0u, 0u);
			static_for<NCH>([&](auto C) {
				constexpr int c = decltype(C)::value, cl = fg_clen(BT, c);
				unsigned long long w64 = 0;
				static_for<cl>([&](auto K) { w64 |= (unsigned long long)bf_s_sym(BT, 0, c, decltype(K)::value) << (2 * decltype(K)::value); });
				if (cc == c && l < cl)
					ent = make_uint4((unsigned)((fg_roff(BT, W, c) + 2 + l * FG_SPS) * 8), __float_as_uint((float)(fg_cpos(BT, c) + l)),
					                 (unsigned)((w64 >> (2 * l)) & 3), 0u);
			});
			lane_tab[t] = ent;
		}
	} else if (warp == 0) {
		int t_roff = 0, t_end = 0, c_src = 0, c_srcp = 0, syms = 0;
		float t_posf = 0.0f, c_dist = 1.0f;
		static_for<NCH>([&](auto C) {
			constexpr int c = decltype(C)::value, st = fg_cstart(BT, c), cl = fg_clen(BT, c);
			if (lane >= st && lane < st + cl) {
				t_roff = fg_roff(BT, W, c) + 2 + (lane - st) * FG_SPS;     // sample of the symbol for d = 0
				t_posf = (float)(fg_cpos(BT, c) + lane - st);
				t_end = st + cl;
			}
			if constexpr (c > 0) {
				// lane c: angle between chunk c and chunk c-1 over the distance of their centres (pi4cxpsk.c:390-394)
				constexpr float pc = (float)fg_cpos(BT, c) + (float)fg_clen(BT, c) / 2.0f;
				constexpr float pp = (float)fg_cpos(BT, c - 1) + (float)fg_clen(BT, c - 1) / 2.0f;
				if (lane == c) {
					c_src = st;
					c_srcp = fg_cstart(BT, c - 1);
					c_dist = pc - pp;
				}
			}
		});
		static_for<G::NSYNC>([&](auto S) {
			constexpr int s = decltype(S)::value;
			unsigned long long w64 = 0;
			static_for<NCH>([&](auto C) {
				constexpr int c = decltype(C)::value;
				static_for<fg_clen(BT, c)>([&](auto K) {
					w64 |= (unsigned long long)bf_s_sym(BT, s, c, decltype(K)::value) << (2 * (fg_cstart(BT, c) + decltype(K)::value));
				});
			});
			syms |= (int)((w64 >> (2 * lane)) & 3) << (2 * s);
		});
		lane_tab[lane] = make_uint4((unsigned)(t_roff * 8), __float_as_uint(t_posf),
		                            (unsigned)(t_end | (c_src << 8) | (c_srcp << 16) | (syms << 24)), __float_as_uint(c_dist));
	}
	__syncthreads();
	const unsigned lane_tab_s = (unsigned)__cvta_generic_to_shared(&lane_tab[lane]);
	const unsigned reg_s = (unsigned)__cvta_generic_to_shared(sm.reg);

	constexpr float inv_dd256x2 = 2.0f * (float)(LUT_CELLS << NB) * 0.15915494309189533577f;   // table BYTES per radian
	constexpr float TWO_PI_HI = 6.28125f, TWO_PI_LO = 1.9353071795864769e-3f, INV_2PI = 0.15915494309189533577f;
	constexpr float rotation = 3.14159265358979323846264338327f / BURSTS[BT].rot_div;
	// rotated taps: built once when the batch shares one frequency shift, else cached per warp and rebuilt when the
	// shift changes (the shift they were built for sits in the warp's shared-memory slice, not in a register)
	if (!a.freq_shift)
		fast_build_taps<BT, G>(sm, (a.freq_shift0 - rotation) / (float)FG_SPS, lane);
	else if (lane == 0)
		*sm.fs_taps = __int_as_float(0x7fc00000);       // NaN: none
	__syncwarp();

	const int n_eff = a.n_dev ? min(a.n, *a.n_dev) : a.n;
	const int b_step = gridDim.x * DM_WARPS;
	for (int b = blockIdx.x * DM_WARPS + warp; b < n_eff; b += b_step) {
		const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
		const float freq_shift = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;
		const float fs = (freq_shift - rotation) / (float)FG_SPS;
		const bool aligned = (((uintptr_t)x) & 15) == 0;
#if DMF_PREFETCH == 4
		{	// one burst ahead: at the start of burst b the window of burst b + step is requested (the first burst: both)
			if (lane == 0 && aligned && b < b_step)
				asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "n"(G::L * 8) : "memory");
			const int bn = b + b_step;
			if (lane == 0 && bn < n_eff) {
				const float2 *xn = a.iq + (a.ofs ? a.ofs[bn] : (int64_t)bn * a.stride);
				if ((((uintptr_t)xn) & 15) == 0)
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xn), "n"(G::L * 8) : "memory");
			}
		}
#elif DMF_PREFETCH == 3
		// first burst of this warp only: later windows were requested while the previous burst was sliced
		if (lane == 0 && aligned && b < b_step)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "n"(G::L * 8) : "memory");
#elif DMF_PREFETCH == 1
		// whole window -> L2 with one bulk prefetch (TMA unit, no registers, no completion to wait for)
		if (lane == 0 && aligned)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "n"(G::L * 8) : "memory");
#elif DMF_PREFETCH == 2
		// the part of the window the first batch of loads does not touch, one 128-byte line per lane and request
		static_for<(G::L * 8 - DMF_PF_FROM + 4095) / 4096>([&](auto K) {
			constexpr int o = DMF_PF_FROM + decltype(K)::value * 4096;
			if (o + lane * 128 < G::L * 8)
				asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(x) + o + lane * 128));
		});
#endif
		__syncwarp();        // the previous burst's training symbols were read from the regions
		if (a.freq_shift && fs != *sm.fs_taps) {
			fast_build_taps<BT, G>(sm, fs, lane);
			if (lane == 0)
				*sm.fs_taps = fs;
			__syncwarp();
		}
		FNorm nm;
		if (aligned)
			nm = stats_fill<BT, G, WANT_SD>(x, lane, sm.reg);
		else
			nm = stats_fill_unaligned<BT, G>(x, lane, sm.reg, WANT_SD);
		__syncwarp();

		// ---- 2./3. sync search: all sequences, |corr| of every chunk summed per search offset (accumulator never
		// cleared between sequences unless sync_reset, pi4cxpsk.c:207,232), peak + early/late per sequence
		float toa = 0.0f, pwr = 0.0f;
		int sync_id = -1;
		{
			float acc[ROWS];
#pragma unroll
			for (int r = 0; r < ROWS; r++)
				acc[r] = 0.0f;
#pragma unroll 1
			for (int s = 0; s < G::NSYNC; s++) {
				if (s > 0 && a.sync_reset) {
#pragma unroll
					for (int r = 0; r < ROWS; r++)
						acc[r] = 0.0f;
				}
				const float2 *tapb = sm.taps + s * G::TAPS;
				static_for<NCH>([&](auto C) {
					constexpr int c = decltype(C)::value, cl = fg_clen(BT, c);
					const float2 Rs = sm.tsum[s * NCH + c];
					// corr = sum t_n x - avg * sum t_n: the accumulators start at -avg * sum t_n
					const float2 init = make_float2(-(nm.ar * Rs.x - nm.ai * Rs.y), -(nm.ar * Rs.y + nm.ai * Rs.x));
					if constexpr (G::TROWS) {
						// quad m <-> search offset m, lane q of the quad takes taps q, q + 4, ...: a quarter of the
						// multiply-adds of the lane-per-offset form, whose lanes >= W idle (99 taps at 7 offsets)
						const int q = lane & 3;
						float2 P = q == 0 ? init : make_float2(0.0f, 0.0f), Q = make_float2(0.0f, 0.0f);
						const float2 *g = sm.reg + fg_roff(BT, W, c) + 2 + (lane >> 2) + FG_SPS * q;
						const float2 *tq = tapb + fg_toff(BT, c) + q;
						static_for<(cl + 3) / 4>([&](auto K) {
							constexpr int k = decltype(K)::value;
							float2 t = tq[4 * k];
							if constexpr (4 * k + 3 >= cl) {
								if (q >= cl - 4 * k)
									t = make_float2(0.0f, 0.0f);
							}
							const float2 v = g[4 * FG_SPS * k];
							fma2s(P, t.x, v);
							fma2s(Q, t.y, v);
						});
						float xr = P.x - Q.y, xi = P.y + Q.x;
						xr += __shfl_xor_sync(0xffffffffu, xr, 1);
						xi += __shfl_xor_sync(0xffffffffu, xi, 1);
						xr += __shfl_xor_sync(0xffffffffu, xr, 2);
						xi += __shfl_xor_sync(0xffffffffu, xi, 2);
						float mag;
						asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(xr, xr, xi * xi)));
						acc[0] += mag;
					} else {
					float2 P[ROWS], Q[ROWS];
#pragma unroll
					for (int r = 0; r < ROWS; r++) {
						P[r] = init;
						Q[r] = make_float2(0.0f, 0.0f);
					}
					const float2 *g = sm.reg + fg_roff(BT, W, c) + 2 + lane;        // sample of tap 0 for offset `lane`
					const float4 *tp4 = reinterpret_cast<const float4 *>(tapb + fg_toff(BT, c));
#if DMF_CORR_PIPE
					// software pipeline: the samples of tap n+1 (and the taps of the next pair) are requested before the
					// multiply-adds of tap n, so a shared-memory latency is covered by the warp's own arithmetic
					float2 vb[2][ROWS];
					float4 tb[2];
					tb[0] = tp4[0];
#pragma unroll
					for (int r = 0; r < ROWS; r++)
						vb[0][r] = g[32 * r];
					static_for<cl>([&](auto N) {
						constexpr int n = decltype(N)::value;
						if constexpr (n + 1 < cl) {
#pragma unroll
							for (int r = 0; r < ROWS; r++)
								vb[(n + 1) & 1][r] = g[32 * r + FG_SPS * (n + 1)];
							if constexpr (((n + 1) & 1) == 0)
								tb[((n + 1) >> 1) & 1] = tp4[(n + 1) >> 1];
						}
						const float4 t = tb[(n >> 1) & 1];
						const float tr = (n & 1) ? t.z : t.x, ti = (n & 1) ? t.w : t.y;
#pragma unroll
						for (int r = 0; r < ROWS; r++) {
							fma2s(P[r], tr, vb[n & 1][r]);
							fma2s(Q[r], ti, vb[n & 1][r]);
						}
					});
#else
					static_for<(cl + 1) / 2>([&](auto N2) {
						constexpr int n2 = decltype(N2)::value;
						const float4 t = tp4[n2];                 // taps 2*n2, 2*n2+1
#pragma unroll
						for (int r = 0; r < ROWS; r++) {
							const float2 v = g[32 * r + FG_SPS * (2 * n2)];
							fma2s(P[r], t.x, v);
							fma2s(Q[r], t.y, v);
						}
						if constexpr (2 * n2 + 1 < cl) {
#pragma unroll
							for (int r = 0; r < ROWS; r++) {
								const float2 v = g[32 * r + FG_SPS * (2 * n2 + 1)];
								fma2s(P[r], t.z, v);
								fma2s(Q[r], t.w, v);
							}
						}
					});
#endif
#pragma unroll
					for (int r = 0; r < ROWS; r++) {
						const float xr = P[r].x - Q[r].y, xi = P[r].y + Q[r].x;
						float mag;
						asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(xr, xr, xi * xi)));
						acc[r] += mag;      // unscaled: 1/stddev changes no decision, it only scales the reported power
					}
					}
				});
				__syncwarp();
				if constexpr (G::TROWS) {
					const float av = __shfl_sync(0xffffffffu, acc[0], (4 * lane) & 31);     // offset `lane` sits in quad `lane`
					sm.accv[lane] = lane < W ? av : 0.0f;
				} else {
#pragma unroll
				for (int r = 0; r < ROWS; r++)
					sm.accv[lane + 32 * r] = (32 * r + 31 < W || lane + 32 * r < W) ? acc[r] : 0.0f;
				}
				__syncwarp();
				float peak;
				TapLane tpl;
				tpl.j = lane - 10;
				tpl.xj = PI_F * (float)(lane - 10);
				tpl.sgn = lane < 21 ? ((lane & 1) ? 1.0f : -1.0f) : 0.0f;
				const float s_toa = peak_early_late<ROWS>(sm.accv, sm.aw, W, tpl, lane, peak);
				peak *= 1.0f / (float)NTR;
				const float s_pwr = peak * peak;
				if (s_pwr > pwr) {
					pwr = s_pwr;
					toa = s_toa;
					sync_id = s;
				}
			}
		}
		if (lane == 0) {
			if (a.sync_id) a.sync_id[b] = sync_id;
			if (a.toa) a.toa[b] = toa;
			if (WANT_SD && a.pwr) a.pwr[b] = pwr * (nm.inv_sd * nm.inv_sd);
		}
		int8_t *eb = a.ebits + (size_t)b * a.ebits_stride;
		if (sync_id < 0) {          // nothing correlated (all-zero input): the reference returns -errno
			if (lane == 0 && a.freq_err) a.freq_err[b] = 0.0f;
#pragma unroll 1
			for (int k = lane; k < G::EBITS; k += 32)
				eb[k] = 0;
			continue;
		}

#if DMF_PREFETCH == 3
		{	// the next window of this warp starts its way to L2 while this burst's symbols are sliced
			const int bn = b + b_step;
			if (lane == 0 && bn < n_eff) {
				const float2 *xn = a.iq + (a.ofs ? a.ofs[bn] : (int64_t)bn * a.stride);
				if ((((uintptr_t)xn) & 15) == 0)
					asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xn), "n"(G::L * 8) : "memory");
			}
		}
#endif
		// symbol i sits at sample i*sps + d (sps >= 4 branch of _gmr1_pi4cxpsk_align, :286-297)
		const float df = roundf(toa);
		const int d = (int)df;

		float ferr, phi0;
		if constexpr (G::TROWS) {
			// ---- 4. (more than 32 training symbols) one row of lanes per chunk: symbol `lane` of chunk c, chunk sums by
			// one folded shuffle tree each, the angle between chunks c and c-1 in lane c, phase reference from the
			// per-lane sums of the rotated symbols
			const unsigned lt_s = (unsigned)__cvta_generic_to_shared(&lane_tab[lane]);
			float2 zc[NCH];
			float posc[NCH];
			float2 S[NCH];
			static_for<NCH>([&](auto C) {
				constexpr int c = decltype(C)::value, cl = fg_clen(BT, c);
				uint4 lt;
				asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lt.x), "=r"(lt.y), "=r"(lt.z), "=r"(lt.w) : "r"(lt_s + c * 512));
				posc[c] = __uint_as_float(lt.y);
				zc[c] = make_float2(0.0f, 0.0f);
				if (lane < cl) {
					float2 v;
					asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(reg_s + lt.x + (unsigned)(d * 8)));
					const float2 e = sincos_red(fs * fmaf(posc[c], (float)FG_SPS, df));
					const float yr = v.x - nm.ar, yi = v.y - nm.ai;
					zc[c] = mul_conj_sym((int)lt.z, make_float2(yr * e.x - yi * e.y, yr * e.y + yi * e.x));
				}
				if constexpr (cl == 1)
					S[c] = make_float2(__shfl_sync(0xffffffffu, zc[c].x, 0), __shfl_sync(0xffffffffu, zc[c].y, 0));
				else
					S[c] = warp_sum2(zc[c].x, zc[c].y, lane);
			});
			float2 sa = S[0], sb = S[0];
			float dist = 1.0f;
			static_for<NCH - 1>([&](auto C) {
				constexpr int c = decltype(C)::value + 1;
				constexpr float pc = (float)fg_cpos(BT, c) + (float)fg_clen(BT, c) / 2.0f;
				constexpr float pp = (float)fg_cpos(BT, c - 1) + (float)fg_clen(BT, c - 1) / 2.0f;
				if (lane == c) {
					sa = S[c];
					sb = S[c - 1];
					dist = pc - pp;
				}
			});
			const float part = fast_atan2f_inl(sa.y * sb.x - sa.x * sb.y, sa.x * sb.x + sa.y * sb.y) / dist;
			float f = 0.0f;
#pragma unroll
			for (int k = 1; k < NCH; k++)
				f += __shfl_sync(0xffffffffu, part, k);
			ferr = f / (float)(NCH - 1);
			if (lane == 0 && a.freq_err) a.freq_err[b] = ferr;
			float2 acc2 = make_float2(0.0f, 0.0f);
			static_for<NCH>([&](auto C) {
				constexpr int c = decltype(C)::value;
				const float2 e = sincos_red((-ferr) * posc[c]);
				acc2.x += zc[c].x * e.x - zc[c].y * e.y;       // lanes past the chunk hold z = 0
				acc2.y += zc[c].x * e.y + zc[c].y * e.x;
			});
			const float2 sum = warp_sum2(acc2.x, acc2.y, lane);
			phi0 = fast_atan2f_inl(sum.y, sum.x);
		} else {
		// ---- 4. training symbols, one per lane, derotated as the reference derotates every sample:
		//      z = (x - avg) * e^{j*fl32(fs*idx)}, times conj(reference symbol)
		uint4 lt;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lt.x), "=r"(lt.y), "=r"(lt.z), "=r"(lt.w) : "r"(lane_tab_s));
		const float t_posf = __uint_as_float(lt.y);
		const int t_end = (int)(lt.z & 0xffu), c_src = (int)((lt.z >> 8) & 0xffu), c_srcp = (int)((lt.z >> 16) & 0xffu);
		float2 z = make_float2(0.0f, 0.0f);
		if (lane < NTR) {
			const int sym = (int)((lt.z >> (24 + 2 * sync_id)) & 3u);
			float2 v;
			asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(reg_s + lt.x + (unsigned)(d * 8)));
			const float2 e = sincos_red(fs * fmaf(t_posf, (float)FG_SPS, df));
			const float yr = v.x - nm.ar, yi = v.y - nm.ai;      // (the scale 1/sd does not change an angle)
			z = mul_conj_sym(sym, make_float2(yr * e.x - yi * e.y, yr * e.y + yi * e.x));
		}
		ferr = 0.0f;
		if constexpr (NCH > 1) {
			// all chunk sums at once: segmented shuffle reduction (a lane adds the value `o` lanes up while that lane
			// is still inside its chunk); the sum of a chunk ends in the chunk's first lane
			float2 v = z;
#pragma unroll
			for (int o = 1; o < 32 && o < NTR; o <<= 1) {
				const float ur = __shfl_down_sync(0xffffffffu, v.x, o), ui = __shfl_down_sync(0xffffffffu, v.y, o);
				if (lane + o < t_end) {
					v.x += ur;
					v.y += ui;
				}
			}
			// lane c (1 .. NCH-1): arg(corr[c] * conj(corr[c-1])) / (pos[c] - pos[c-1]) (:390-394)
			const float sr = __shfl_sync(0xffffffffu, v.x, c_src), si = __shfl_sync(0xffffffffu, v.y, c_src);
			const float qr = __shfl_sync(0xffffffffu, v.x, c_srcp), qi = __shfl_sync(0xffffffffu, v.y, c_srcp);
			const float re = sr * qr + si * qi, im = si * qr - sr * qi;
			const float part = fast_atan2f_inl(im, re) / __uint_as_float(lt.w);
			float f = 0.0f;
#pragma unroll
			for (int k = 1; k < NCH; k++)
				f += __shfl_sync(0xffffffffu, part, k);
			ferr = f / (float)(NCH - 1);
		}
		if (lane == 0 && a.freq_err) a.freq_err[b] = ferr;

		// ---- phase reference: all training symbols after the -ferr rotation (:415-433, :574)
		{
			float2 zr = z;
			if constexpr (NCH > 1) {
				const float2 e = sincos_red((-ferr) * t_posf);
				zr = make_float2(z.x * e.x - z.y * e.y, z.x * e.y + z.y * e.x);
			}
			const float2 sum = warp_sum2(zr.x, zr.y, lane);
			phi0 = fast_atan2f_inl(sum.y, sum.x);
		}

		}

		// ---- 5. data symbols in the angle domain: arg(x - avg) + fl32(fs*idx) + fl32(-ferr*i) - arg(phasor), the
		// rotation term reduced mod 2*pi first (Cody-Waite); soft bits from the table over the symbol value
		{
			const char *xd = reinterpret_cast<const char *>(x + d);
			const float nferr = -ferr, nphi0 = -phi0;
			const float2 navg = make_float2(-nm.ar, -nm.ai);
			const bool eb_even = NB == 1 || (((uintptr_t)eb) & 1) == 0;
#pragma unroll 1
			for (int it = 0; it < G::DITER; it++) {
				const int row0 = it * DB;
				uint2 e[DB];
				float2 v[DB];
#pragma unroll
				for (int u = 0; u < DB; u++) {
					e[u] = dtab[min(row0 + u, G::DROWS - 1) * 32 + lane];
#if DMF_CACHE_HINTS & 1
					v[u] = __ldcs(reinterpret_cast<const float2 *>(xd + e[u].x));     // last use of the window: evict first
#else
					v[u] = __ldg(reinterpret_cast<const float2 *>(xd + e[u].x));
#endif
				}
				unsigned sw[DB];
#pragma unroll
				for (int u = 0; u < DB; u++) {
					const float posf = __uint_as_float(e[u].y);
					const float2 y = fadd2(v[u], navg);
					const float th = fast_atan2f_inl(y.y, y.x);
					const float a1 = fs * fmaf(posf, (float)FG_SPS, df);
					const float k = rintf(a1 * INV_2PI);
					float r = fmaf(k, -TWO_PI_HI, a1);
					r = fmaf(k, -TWO_PI_LO, r);
					const float sv = (((th + r) + nferr * posf) + nphi0) * inv_dd256x2;
					const unsigned cell = (unsigned)__float2int_rd(sv) & (2u * (LUT_CELLS << NB) - 2u);
					sw[u] = *reinterpret_cast<const uint16_t *>(reinterpret_cast<const char *>(soft_lut) + cell);
				}
				if (eb_even) {
#pragma unroll
					for (int u = 0; u < DB; u++) {
						const int t = (row0 + u) * 32 + lane;
						if (t < G::NDS) {
#if DMF_CACHE_HINTS & 2
							if (NB == 2)
								__stcs(reinterpret_cast<unsigned short *>(eb) + t, (unsigned short)sw[u]);
							else
								__stcs(reinterpret_cast<signed char *>(eb) + t, (signed char)sw[u]);
#else
							if (NB == 2)
								reinterpret_cast<uint16_t *>(eb)[t] = (uint16_t)sw[u];
							else
								eb[t] = (int8_t)sw[u];
#endif
						}
					}
				} else {
#pragma unroll
					for (int u = 0; u < DB; u++) {     // odd output address: byte stores
						const int t = (row0 + u) * 32 + lane;
						if (t < G::NDS) {
							eb[2 * t] = (int8_t)(sw[u] & 0xff);
							eb[2 * t + 1] = (int8_t)(sw[u] >> 8);
						}
					}
				}
			}
		}
	}
}

// ---- launcher ---------------------------------------------------------------------------------------------------
namespace {

struct FastEntry {
	int bt, w;
	const void *fn[2];      // [want_sd]
	size_t warp_bytes, static_bytes;
};

template <int BT, int W>
FastEntry make_entry()
{
	using G = Geo<BT, W>;
	return {BT, W, {(const void *)demod_fast_kernel<BT, W, false>, (const void *)demod_fast_kernel<BT, W, true>},
	        (size_t)G::WARP_BYTES, (size_t)(2 * (LUT_CELLS << G::NB) + 8 * G::DROWS * 32 + 512 * G::TR)};
}

// the standard search widths: what gmr1_rx cuts (BCCH 20*sps, DC6 10*sps, NT3 / NT9 sps + sps/2, gmr1_rx.c:290,549,
// 759,809) and the natural ones for the formats it does not use
const FastEntry *fast_entries(int *n)
{
	static const FastEntry tab[] = {
		make_entry<BT_BCCH, 81>(), make_entry<BT_DC6, 41>(), make_entry<BT_NT3_SPEECH, 7>(), make_entry<BT_NT3_FACCH, 7>(),
		make_entry<BT_NT9, 7>(),   make_entry<BT_NT6, 7>(),  make_entry<BT_SDCCH, 41>(),     make_entry<BT_DC2, 25>(),
		make_entry<BT_DC12, 41>(), make_entry<BT_RACH, 7>(),
	};
	*n = (int)(sizeof(tab) / sizeof(tab[0]));
	return tab;
}

}  // namespace

std::atomic<int> g_demod_generic{0};

// Launches the per-format kernel when (burst type, sps, search width) is one of the compiled combinations;
// returns false when the generic kernel has to take the batch.
bool launch_demod_fast(const DemodArgs &a, int bt, cudaStream_t st, cudaError_t *err)
{
	*err = cudaSuccess;
	if (a.sps != FG_SPS || bt < 0 || bt >= BT_COUNT)
		return false;
	if (g_demod_generic.load(std::memory_order_relaxed))
		return false;
	const int w = a.win_len - BURSTS[bt].len * FG_SPS + 1;
	int n_ent = 0;
	const FastEntry *ent = fast_entries(&n_ent), *e = nullptr;
	for (int i = 0; i < n_ent; i++)
		if (ent[i].bt == bt && ent[i].w == w)
			e = &ent[i];
	if (!e)
		return false;
	if (a.n <= 0)
		return true;
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64)
		return false;
	struct DevState { std::once_flag once; cudaError_t err; int sms; };
	static DevState ds[64];
	std::call_once(ds[dev].once, [&] {
		cudaError_t r = upload_sinpi512();
		int ne = 0;
		const FastEntry *t = fast_entries(&ne);
		for (int i = 0; i < ne && r == cudaSuccess; i++)
			for (int k = 0; k < 2 && r == cudaSuccess; k++)
				r = cudaFuncSetAttribute(t[i].fn[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(t[i].warp_bytes * DM_WARPS));
		int v = 148;
		cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
		ds[dev].sms = v;
		ds[dev].err = r;
	});
	if (ds[dev].err != cudaSuccess) {
		*err = ds[dev].err;
		return true;
	}
	const size_t smem = e->warp_bytes * DM_WARPS;
	int per_sm = (int)((228 * 1024) / (smem + e->static_bytes + 1024 + 64));
	per_sm = per_sm < 1 ? 1 : (per_sm > DMF_MIN_CTAS ? DMF_MIN_CTAS : per_sm);
	if (const char *ev = getenv("GMR1B200_DEMOD_CTAS")) {     // tuning knob: resident CTAs per SM
		const int v = atoi(ev);
		if (v >= 1 && v < per_sm)
			per_sm = v;
	}
	int grid = (a.n + DM_WARPS - 1) / DM_WARPS;
	if (grid > ds[dev].sms * per_sm)
		grid = ds[dev].sms * per_sm;
	const bool want_sd = a.pwr != nullptr;
	void *args[] = {(void *)&a};
	*err = cudaLaunchKernel(e->fn[want_sd ? 1 : 0], dim3(grid), dim3(DM_WARPS * 32), args, smem, st);
	return true;
}

}  // namespace gmr1
