// fcch_grid.cu - stage 1, second generation of the coarse FCCH search: gmr1_fcch_rough (src/sdr/fcch.c:211-250) for
// one frequency shift or for a whole GRID of shifts per window ("+-frequency-offset FCCH search", BASELINE config 4),
// with the search window read, averaged, decimated and normalised ONCE.
//
// What changes against fcch_rough_kernel (fcch_kernels.cu), which stays the path for the multi-FCCH correlation power:
//   * the frequency shift never touches the samples.  The reference rotates them, y[i] e^{j f i}, and correlates with
//     the real dual chirp r[n]; only |corr|^2 is used, and  sum_n r[n] y[m+n] e^{j f (m+n)} = e^{j f m} (A[m] + j B[m])
//     with A = sum_n r[n] cos(f n) y[m+n], B = sum_n r[n] sin(f n) y[m+n]: two real-tap correlations of the SAME
//     samples.  The shift -f is A - j B: a symmetric pair of shifts costs what two separate searches cost, a grid
//     {0, +-f1, +-f2} five real-tap correlations - but one pass over the window instead of five, and no sine / cosine
//     per sample (7 722 of them per shift in the rotating form);
//   * the 5-sample energy-window argmax and its centroid (osmo_cxvec_peak_energy_find, PEAK_WEIGH_WIN) are taken on the
//     fly from the outputs in registers (a thread's eight consecutive outputs + the last four of its left neighbour):
//     no |corr|^2 array, 30 KB of shared memory less;
//   * the decimated window sits unpadded in shared memory (62 KB at sps 4) with an XOR swizzle of the 16-byte chunks
//     instead of the 20-of-16 padding: a quarter-warp's 16-byte loads still cover all banks once.  With 256 threads
//     per CTA that makes three windows per SM resident (single shift) instead of two.
// Float contract as for fcch_rough_kernel: integer TOA equal to the C path except at rounding ties.
#include <cuda_runtime.h>
#include <math.h>

#include "launch.h"

namespace gmr1 {

namespace {

constexpr float PI_F = 3.14159265358979323846264338327f;
constexpr int FG_T = 256;                 // threads per CTA
constexpr int FG_TILE = 8;                // consecutive outputs per thread and round
constexpr int FG_MAXLEN = 120;            // taps, padded to whole groups of 8 (gmr1_fcch_burst: 117)

__device__ __forceinline__ float wsum(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// float index of sample i (re; im follows) in the swizzled window: group of 8 samples = 4 chunks of 16 bytes,
// chunk c of group g stored at chunk (c ^ ((g >> 1) & 3))
__device__ __forceinline__ int sidx(int i)
{
	const int g = i >> 3, u = i & 7;
	return g * 16 + ((((u >> 1) ^ (g >> 1)) & 3) << 2) + ((u & 1) << 1);
}

// (cr, ci) += r * (wr, wi): one packed FFMA2
__device__ __forceinline__ void fma2(float2 &c, float r, const float2 w)
{
	unsigned long long cc = *reinterpret_cast<unsigned long long *>(&c);
	const float2 rr2 = make_float2(r, r);
	asm("fma.rn.f32x2 %0, %1, %2, %0;"
	    : "+l"(cc)
	    : "l"(*reinterpret_cast<const unsigned long long *>(&rr2)), "l"(*reinterpret_cast<const unsigned long long *>(&w)));
	c = *reinterpret_cast<float2 *>(&cc);
}

struct Best {                              // best 5-sample energy window seen by a thread for one shift
	float val;
	int   idx;                             // last sample of the window
	float e[5];                            // the five energies the centroid is taken over
};

__device__ __forceinline__ void best_init(Best &b)
{
	b.val = 0.0f;
	b.idx = 0x7fffffff;
#pragma unroll
	for (int k = 0; k < 5; k++)
		b.e[k] = 0.0f;
}

// x[0..3] = the four energies in front of this thread's outputs, x[4..11] = its eight outputs m0 .. m0+7
__device__ __forceinline__ void best_scan(Best &b, const float (&x)[12], int m0, int nc)
{
#pragma unroll
	for (int j = 0; j < FG_TILE; j++) {
		// oldest first, as the reference sums (the zeros in front of output 0 stand in for indices < 0)
		const float val = ((((0.0f + x[j]) + x[j + 1]) + x[j + 2]) + x[j + 3]) + x[j + 4];
		if (val > b.val && m0 + j < nc) {
			b.val = val;
			b.idx = m0 + j;
			// centroid window: [idx-4, idx], or [0, 5) when that would start in front of the vector (only for m0 = 0)
			const bool head = m0 + j < 4;
#pragma unroll
			for (int k = 0; k < 5; k++)
				b.e[k] = head ? x[4 + k] : x[j + k];
		}
	}
}

struct GridPlan {
	int32_t n_pass;                        // distinct |shift| values
	float   f[8];                          // |shift| of a pass, rad/symbol
	int32_t out_p[8], out_m[8];            // output slot of the shift +f / -f of a pass, -1 = not wanted
};

struct BarCta {                            // all threads of the CTA
	static __device__ __forceinline__ void sync() { __syncthreads(); }
};
struct BarCompute {                        // the FG_T compute threads of the pipelined kernel (named barrier 1)
	static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;" ::"n"(FG_T) : "memory"); }
};

// The search over one normalised, decimated window `w` (swizzled, zero tail behind sample l) by FG_T threads:
// every pass of the plan, results to toa_out / peak_out [slot][n_total].  Bar synchronises those FG_T threads.
template <bool PAIR, class Bar>
__device__ __forceinline__ void search_window(const FcchArgs &a, const GridPlan &gp, int32_t *toa_out, float *peak_out,
                                              const float *w, float *refc, float *refs, float *red, float *tail,
                                              int tid, int b, int n_total, int l, int len, int ng, float fbase)
{
	const int lane = tid & 31, warp = tid >> 5;
	const int nc = l - len + 1;
	const int lenp = (len + 7) & ~7;
	const int rounds = (nc + FG_T * FG_TILE - 1) / (FG_T * FG_TILE);
#pragma unroll 1
	for (int pass = 0; pass < gp.n_pass; pass++) {
		const float f = gp.n_pass == 1 && gp.f[0] < 0.0f ? fabsf(fbase) : gp.f[pass];
		Bar::sync();                   // normalised samples ready / previous pass done with the taps
		{	// dual-chirp reference at 1 sample/symbol (fcch.c:167-193) times e^{j f n}
			const float phase_base = a.freq * 2.0f * PI_F / (float)len, halfpos = (float)len / 2.0f;
			for (int i = tid; i < FG_MAXLEN; i += FG_T) {
				const float pos = (float)i - halfpos;
				const float r = i < len ? sqrtf(2.0f) * cosf(phase_base * (pos * pos)) : 0.0f;
				float sn, cs;
				sincosf(f * (float)i, &sn, &cs);
				refc[i] = r * cs;
				refs[i] = r * sn;
			}
			if (tid < 8)                   // carry slots: nothing in front of output 0
				tail[(tid >> 2) * (FG_T + 1) * 4 + (tid & 3)] = 0.0f;
		}
		Bar::sync();
		Best bp, bm;
		best_init(bp);
		best_init(bm);
#pragma unroll 1
		for (int rd = 0; rd < rounds; rd++) {
			const int m0 = (rd * FG_T + tid) * FG_TILE;
			float2 c[FG_TILE], d[FG_TILE], win[FG_TILE], alt[FG_TILE];
			const float4 *w4 = reinterpret_cast<const float4 *>(w);
			int g = m0 >> 3;                                   // m0 < ng * 8 always (zero tail)
			const bool act = g < ng - (lenp >> 3);             // a whole tap span fits behind this group
			if (!act)
				g = 0;
			auto load = [&](float2 (&dst)[FG_TILE], int gg) {
				const int sw = (gg >> 1) & 3;
#pragma unroll
				for (int q = 0; q < 4; q++) {
					const float4 v = w4[gg * 4 + (q ^ sw)];
					dst[2 * q] = make_float2(v.x, v.y);
					dst[2 * q + 1] = make_float2(v.z, v.w);
				}
			};
			load(win, g);
#pragma unroll
			for (int t = 0; t < FG_TILE; t++)
				c[t] = d[t] = make_float2(0.0f, 0.0f);
			// one group of 8 taps: `cur` holds samples m0 + 8k .. + 7, `nxt` the following 8; output t of tap u
			// reads sample u + t of the 16.  Two groups per iteration with the roles of the two register windows
			// swapped, so no register is ever copied.
			auto group = [&](int k, const float2 (&cur)[FG_TILE], float2 (&nxt)[FG_TILE]) {
				load(nxt, g + k + 1);
				const float4 c0 = reinterpret_cast<const float4 *>(refc)[2 * k], c1 = reinterpret_cast<const float4 *>(refc)[2 * k + 1];
				const float rc[FG_TILE] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
				for (int u = 0; u < FG_TILE; u++)
#pragma unroll
					for (int t = 0; t < FG_TILE; t++)
						fma2(c[t], rc[u], u + t < FG_TILE ? cur[u + t] : nxt[u + t - FG_TILE]);
				if (PAIR) {
					const float4 s0 = reinterpret_cast<const float4 *>(refs)[2 * k], s1 = reinterpret_cast<const float4 *>(refs)[2 * k + 1];
					const float rs[FG_TILE] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
					for (int u = 0; u < FG_TILE; u++)
#pragma unroll
						for (int t = 0; t < FG_TILE; t++)
							fma2(d[t], rs[u], u + t < FG_TILE ? cur[u + t] : nxt[u + t - FG_TILE]);
				}
			};
			const int G = lenp >> 3;
			int k = 0;
#pragma unroll 1
			for (; k + 1 < G; k += 2) {
				group(k, win, alt);
				group(k + 1, alt, win);
			}
			if (k < G)
				group(k, win, alt);
			// energies of the two shifts of this pass, window sums with the left neighbour's last four
			float xp[12], xm[12];
#pragma unroll
			for (int t = 0; t < FG_TILE; t++) {
				const bool ok = act && m0 + t < nc;
				if (PAIR) {
					const float pr = c[t].x - d[t].y, pi = c[t].y + d[t].x;      // A + jB
					const float mr = c[t].x + d[t].y, mi = c[t].y - d[t].x;      // A - jB
					xp[4 + t] = ok ? pr * pr + pi * pi : 0.0f;
					xm[4 + t] = ok ? mr * mr + mi * mi : 0.0f;
				} else {
					xp[4 + t] = ok ? c[t].x * c[t].x + c[t].y * c[t].y : 0.0f;
					xm[4 + t] = 0.0f;
				}
			}
			float *tp = tail, *tm = tail + (FG_T + 1) * 4;
			*reinterpret_cast<float4 *>(&tp[(tid + 1) * 4]) = make_float4(xp[8], xp[9], xp[10], xp[11]);
			if (PAIR)
				*reinterpret_cast<float4 *>(&tm[(tid + 1) * 4]) = make_float4(xm[8], xm[9], xm[10], xm[11]);
			Bar::sync();
			{
				const float4 v = *reinterpret_cast<const float4 *>(&tp[tid * 4]);
				xp[0] = v.x; xp[1] = v.y; xp[2] = v.z; xp[3] = v.w;
				if (PAIR) {
					const float4 u = *reinterpret_cast<const float4 *>(&tm[tid * 4]);
					xm[0] = u.x; xm[1] = u.y; xm[2] = u.z; xm[3] = u.w;
				}
			}
			Bar::sync();
			if (tid == FG_T - 1) {         // carry into the next round
				*reinterpret_cast<float4 *>(&tp[0]) = make_float4(xp[8], xp[9], xp[10], xp[11]);
				if (PAIR)
					*reinterpret_cast<float4 *>(&tm[0]) = make_float4(xm[8], xm[9], xm[10], xm[11]);
			}
			best_scan(bp, xp, m0, nc);
			if (PAIR)
				best_scan(bm, xm, m0, nc);
		}
		// block argmax (largest value, lowest index on ties) per shift; the winner writes TOA and peak
#pragma unroll
		for (int sgn = 0; sgn < (PAIR ? 2 : 1); sgn++) {
			const Best &bb = sgn ? bm : bp;
			const bool single = gp.n_pass == 1 && gp.f[0] < 0.0f;
			int slot = sgn ? gp.out_m[pass] : gp.out_p[pass];
			if (single)                    // single-shift use: the sign of the window's own shift picks the branch
				slot = (fbase < 0.0f) == (sgn == 1) ? 0 : -1;
			float bv = bb.val;
			int bi = bb.idx;
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
			}
			Bar::sync();
			int *redi = (int *)(red + 32);
			if (lane == 0) { red[warp] = bv; redi[warp] = bi; }
			Bar::sync();
			bv = lane < FG_T / 32 ? red[lane] : 0.0f;
			bi = lane < FG_T / 32 ? redi[lane] : 0x7fffffff;
#pragma unroll
			for (int o = 16; o; o >>= 1) {
				const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
				const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
			}
			if (slot < 0)
				continue;
			const size_t o = (size_t)slot * n_total + b;
			if (bv <= 0.0f) {              // nothing correlated: position 0 (max_idx = 0, empty centroid)
				if (tid == 0) {
					toa_out[o] = 0;
					if (peak_out) peak_out[o] = bv;
				}
			} else if (bb.idx == bi && bb.val == bv) {         // exactly one thread owns the winning window
				const int max_idx = bi - 4 < 0 ? 0 : bi - 4;
				float mw = 0.0f, sw = 0.0f;
#pragma unroll
				for (int k2 = 0; k2 < 5; k2++) {
					sw += bb.e[k2];
					mw += bb.e[k2] * (float)(max_idx + k2);
				}
				const float pos = sw > 0.0f ? mw / sw : (float)max_idx;
				toa_out[o] = (int)round((double)(pos * 4.0f));
				if (peak_out) peak_out[o] = bv;
			}
		}
	}
}

// PAIR = false: every pass has f = 0 (real taps only).  One CTA per window.
template <bool PAIR>
__global__ void __launch_bounds__(FG_T, PAIR ? 2 : 3)
fcch_grid_kernel(const FcchArgs a, const GridPlan gp, int32_t *toa_out, float *peak_out)
{
	extern __shared__ __align__(16) uint8_t smem[];
	const int tid = threadIdx.x, b = blockIdx.x, lane = tid & 31, warp = tid >> 5;
	if (a.skip && a.skip[b])
		return;
	const int L = a.win_len, len = a.len;
	const int l = L >> 2;                  // decimated length (sps 4)
	const int nc = l - len + 1;
	const int lenp = (len + 7) & ~7;
	const int ng = (l + lenp + 15) >> 3;   // sample groups incl. the zero tail the padded taps touch
	float *w = (float *)smem;              // [ng * 16] decimated, normalised samples (swizzled)
	float *refc = w + ng * 16;             // [FG_MAXLEN] r[n] cos(f n)
	float *refs = refc + FG_MAXLEN;        // [FG_MAXLEN] r[n] sin(f n)
	float *red = refs + FG_MAXLEN;         // [64] reduction scratch
	float *tail = red + 64;                // [2][FG_T + 1][4] last four energies of every thread (+ carry slot 0)

	const float2 *x = a.iq + (a.ofs ? a.ofs[b] : (int64_t)b * a.stride);
	if ((((uintptr_t)x) & 15) == 0) {      // the whole window -> L2 up front
		const int chunk = 16384, bytes = (L * 8) & ~15;
		for (int o = tid * chunk; o < bytes; o += FG_T * chunk)
			asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char *)x + o), "r"(min(chunk, bytes - o))
			             : "memory");
	}
	for (int i = l + tid; i < ng * 8; i += FG_T)
		*reinterpret_cast<float2 *>(&w[sidx(i)]) = make_float2(0.0f, 0.0f);

	// ---- statistics over ALL samples (sig_normalize averages before decimating); every 4th sample is kept
	float sr = 0.0f, si = 0.0f, sq = 0.0f;
	{
		const int nq = L >> 2;
		float2 s2 = make_float2(0.0f, 0.0f), q2 = make_float2(0.0f, 0.0f);
		if ((((uintptr_t)x) & 15) == 0) {
			const float4 *x4 = reinterpret_cast<const float4 *>(x);
#pragma unroll 4
			for (int i = tid; i < nq; i += FG_T) {
				const float4 v0 = __ldg(&x4[2 * i]), v1 = __ldg(&x4[2 * i + 1]);
				const float2 p0 = make_float2(v0.x, v0.y), p1 = make_float2(v0.z, v0.w), p2 = make_float2(v1.x, v1.y),
				             p3 = make_float2(v1.z, v1.w);
				s2 = __fadd2_rn(s2, __fadd2_rn(__fadd2_rn(p0, p1), __fadd2_rn(p2, p3)));
				q2 = __ffma2_rn(p0, p0, q2);
				q2 = __ffma2_rn(p1, p1, q2);
				q2 = __ffma2_rn(p2, p2, q2);
				q2 = __ffma2_rn(p3, p3, q2);
				*reinterpret_cast<float2 *>(&w[sidx(i)]) = p0;
			}
		} else {
#pragma unroll 1
			for (int i = tid; i < nq; i += FG_T)
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const float2 v = __ldg(&x[4 * i + k]);
					s2 = __fadd2_rn(s2, v);
					q2 = __ffma2_rn(v, v, q2);
					if (k == 0)
						*reinterpret_cast<float2 *>(&w[sidx(i)]) = v;
				}
		}
		sr = s2.x;
		si = s2.y;
		sq = q2.x + q2.y;
		for (int i = 4 * nq + tid; i < L; i += FG_T) {           // L % 4 trailing samples (none is kept)
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
		red[8 + warp] = si;
		red[16 + warp] = sq;
	}
	__syncthreads();
	sr = wsum(lane < FG_T / 32 ? red[lane] : 0.0f);
	si = wsum(lane < FG_T / 32 ? red[8 + lane] : 0.0f);
	sq = wsum(lane < FG_T / 32 ? red[16 + lane] : 0.0f);
	const float ar = sr / (float)L, ai = si / (float)L;
	const float var = sq / (float)L - (ar * ar + ai * ai);
	float sd = var > 0.0f ? sqrtf(var) : 0.0f;
	if (sd == 0.0f)
		sd = 1.0f;
	const float inv_sd = 1.0f / sd;
	for (int i = tid; i < l; i += FG_T) {
		float2 *p = reinterpret_cast<float2 *>(&w[sidx(i)]);
		*p = make_float2((p->x - ar) * inv_sd, (p->y - ai) * inv_sd);
	}

	const float fbase = a.freq_shift ? a.freq_shift[b] : a.freq_shift0;     // single-shift use: the pass is |fbase|
	search_window<PAIR, BarCta>(a, gp, toa_out, peak_out, w, refc, refs, red, tail, tid, b, (int)gridDim.x, l, len, ng, fbase);
}

}  // namespace

// shifts == NULL: one search per window with the window's own shift (a.freq_shift / a.freq_shift0), results to a.toa /
// a.peak.  Else: n_shifts searches per window, results to toa [n_shifts][n] / peak [n_shifts][n] (peak may be NULL).
// Returns cudaErrorNotSupported when the geometry is outside what this kernel covers (the caller falls back).
cudaError_t launch_fcch_grid(const FcchArgs &a, const float *shifts, int n_shifts, int32_t *toa, float *peak, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	const int l = a.win_len / a.sps, nc = l - a.len + 1;
	const int lenp = (a.len + 7) & ~7;
	if (a.sps != 4 || a.len > FG_MAXLEN - 3 || nc < 8 || a.en_out)
		return cudaErrorNotSupported;
	const int ng = (l + lenp + 15) >> 3;
	const size_t smem = sizeof(float) * ((size_t)ng * 16 + 2 * FG_MAXLEN + 64 + 2 * (FG_T + 1) * 4);
	if (smem > 227 * 1024)
		return cudaErrorNotSupported;
	GridPlan gp = {};
	bool pair = false;
	if (!shifts) {
		gp.n_pass = 1;
		gp.f[0] = -1.0f;                   // marker: the pass is |the window's own shift|
		gp.out_p[0] = gp.out_m[0] = 0;
		pair = a.freq_shift != nullptr || a.freq_shift0 != 0.0f;
		toa = a.toa;
		peak = a.peak;
	} else {
		if (n_shifts < 1 || n_shifts > 16 || a.freq_shift || a.freq_shift0 != 0.0f)
			return cudaErrorNotSupported;
		for (int k = 0; k < n_shifts; k++) {
			const float f = fabsf(shifts[k]);
			int p = -1;
			for (int q = 0; q < gp.n_pass; q++)
				if (gp.f[q] == f)
					p = q;
			if (p < 0) {
				if (gp.n_pass == 8)
					return cudaErrorNotSupported;
				p = gp.n_pass++;
				gp.f[p] = f;
				gp.out_p[p] = gp.out_m[p] = -1;
			}
			int32_t &slot = (shifts[k] < 0.0f) ? gp.out_m[p] : gp.out_p[p];
			if (slot >= 0)
				return cudaErrorNotSupported;      // the same shift twice
			slot = k;
			pair = pair || f != 0.0f;
		}
	}
	GMR1_INIT_LOCK();
	static size_t attr_set[64][2] = {{0}};
	int dev = 0;
	cudaGetDevice(&dev);
	const void *fn = pair ? (const void *)fcch_grid_kernel<true> : (const void *)fcch_grid_kernel<false>;
	if (dev >= 64 || attr_set[dev][pair] < smem) {
		cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_set[dev][pair] = smem;
	}
	if (pair)
		fcch_grid_kernel<true><<<a.n, FG_T, smem, st>>>(a, gp, toa, peak);
	else
		fcch_grid_kernel<false><<<a.n, FG_T, smem, st>>>(a, gp, toa, peak);
	return cudaGetLastError();
}

}  // namespace gmr1
