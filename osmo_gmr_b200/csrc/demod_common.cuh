// demod_common.cuh - device helpers shared by the generic demodulation kernel (demod_kernels.cu) and the per-format
// kernels (demod_fast.cu): warp reductions, reduced-argument sine / cosine, polynomial atan2, the early/late peak
// search of osmo_cxvec_peak_energy_find, packed FMA, the soft-bit rule of pi4cxpsk.c:468-503.
// Every translation unit gets its own copy of the __constant__ sine table (no relocatable device code).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gmr1 {
static constexpr int DM_WARPS = 4;
#define DM_PRAGMA(x) _Pragma(#x)
#define DM_UNROLL(n) DM_PRAGMA(unroll n)
#ifndef DM_STATS_UNROLL
#define DM_STATS_UNROLL 4      // 16-byte loads of the statistics pass in flight per lane
#endif
#ifndef DM_SYM_BATCH
#define DM_SYM_BATCH 4         // data symbols per lane whose sample loads are issued together
#endif
#ifndef DM_OPT_ATAN
#define DM_OPT_ATAN 1
#endif
#ifndef DM_OPT_LUT
#define DM_OPT_LUT 1
#endif
#ifndef DM_OPT_RADIX
#define DM_OPT_RADIX 1         // early/late search: three 8-point evaluations instead of eight sequential steps
#endif
#ifndef DM_OPT_WALK
#define DM_OPT_WALK 1          // early/late search: rounds unrolled, tie-free decision walk on the fast path
#endif
#ifndef DM_OPT_FSC
#define DM_OPT_FSC 1           // training symbols / phase reference: reduced-argument hardware sincos
#endif
#ifndef DM_PREFETCH
#define DM_PREFETCH 1          // 1: own window into L2 at burst start, 2: next window at the start of the data symbols,
#endif                         // 3: 1 + the first DM_PF_NEXT bytes of the next window at the start of the data symbols
#ifndef DM_PF_NEXT
#define DM_PF_NEXT 2048
#endif
#ifndef DM_MIN_CTAS
#define DM_MIN_CTAS 8          // resident CTAs per SM the register allocation is capped for
#endif
static constexpr float PI_F = 3.14159265358979323846264338327f;

// sin(pi * k / 512), k = 0..512: every position the early/late search visits is a multiple of
// 1/512 (start integer, steps 1/2 .. 1/512), so the one sine an interpolation needs is a lookup
static __constant__ float c_sinpi512[513];

// whole window -> L2 with one bulk prefetch (TMA unit, no registers, no completion to wait for)
__device__ __forceinline__ void prefetch_window_l2(const float2 *x, int bytes)
{
	if ((((uintptr_t)x) & 15) == 0)
		asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x), "r"(bytes & ~15) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
	for (int o = 16; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// folded shuffle tree for a (re, im) pair: returns both sums in all lanes
__device__ __forceinline__ float2 warp_sum2(float a, float b, int lane)
{
	const bool up = lane & 16;
	float v = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 16);
#pragma unroll
	for (int o = 8; o; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return make_float2(__shfl_sync(0xffffffffu, v, 0), __shfl_sync(0xffffffffu, v, 16));
}

// One out-of-line copy of the accurate sincosf (its large-argument slow path is ~150 instructions;
// inlined at every call site it pushed the kernel past the instruction cache).
static __device__ __noinline__ float2 sincos_acc(float x)
{
	float sn, cs;
	sincosf(x, &sn, &cs);
	return make_float2(cs, sn);
}

// e^{jx} for the per-symbol derotations of the hot path: x (up to a few hundred radians, fl32(fs * idx) as the
// reference forms it) is reduced mod 2*pi with a two-constant Cody-Waite step (2*pi = HI + LO, HI has 9
// significant bits, so k * HI is exact), then the hardware sine / cosine: ~5e-7 absolute error, against the
// 3e-5 rad that move one soft bit by one LSB in 0.5 % of the symbols.
__device__ __forceinline__ float2 sincos_red(float x)
{
#if DM_OPT_FSC
	const float k = rintf(x * 0.15915494309189533577f);
	float r = fmaf(k, -6.28125f, x);
	r = fmaf(k, -1.9353071795864769e-3f, r);
	float sn, cs;
	__sincosf(r, &sn, &cs);
	return make_float2(cs, sn);
#else
	return sincos_acc(x);
#endif
}

// conj(ref) * g for ref in {1, j, -1, -j} (symbol index 0..3): exact component shuffles
__device__ __forceinline__ float2 mul_conj_sym(int sym, float2 g)
{
	const float a = (sym & 1) ? g.y : g.x, b = (sym & 1) ? -g.x : g.y;
	return (sym & 2) ? make_float2(-a, -b) : make_float2(a, b);
}

// atan2f with ~1e-7 rad absolute error: octant reduction + degree-8 minimax polynomial in a^2
__device__ __forceinline__ float fast_atan2f_inl(float y, float x)
{
	const float ax = fabsf(x), ay = fabsf(y);
	const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
	// mn / mx by one approximate reciprocal (__fdividef carries a denormal-scaling sequence, 9 instructions);
	// mx == 0 implies mn == 0 and the product is 0
#if DM_OPT_ATAN
	float inv;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(fmaxf(mx, 1e-30f)));
	const float a = mn * inv;
#else
	const float a = mx > 0.0f ? __fdividef(mn, mx) : 0.0f;
#endif
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

static __device__ __noinline__ float fast_atan2f(float y, float x) { return fast_atan2f_inl(y, x); }

// Sinc interpolation (osmo_cxvec_interpolate_point, 10 taps either side) of the real correlation
// accumulator, inside the early/late search.  One tap per lane (lane j+10 <-> tap j = -10..10), one
// folded shuffle tree for the early and the late gate.  The 21 sinc values of both gates are
// identical ((i+2)-(pos+2) == i-pos exactly in fp32 here) and share one sine:
// sin(pi*(j-frac)) = -(-1)^j sin(pi*frac), looked up on the 1/512 grid the search visits.
struct TapLane { float xj, sgn; int j; };      // per-lane constants: pi*j, -(-1)^j (0 for lanes >= 21)

__device__ __forceinline__ float ld_acc(const float *acc, int k, int len)
{
	return ((unsigned)k < (unsigned)len) ? acc[k] : 0.0f;
}

// osmo_cxvec_peak_energy_find(acc, 3, PEAK_EARLY_LATE, &peak) on a real vector; all lanes
// return the same position / peak value
template <int ROWS>
__device__ __forceinline__ float peak_early_late(const float *acc, float *aw, int w, const TapLane &tp, int lane, float &peak_val)
{
	// acc[-2], acc[-1] are zero and the row of 32 that holds acc[w-1] is zero beyond it (sync_find), so the
	// energy windows need no edge cases: val[idx] = acc[idx-2]^2 + acc[idx-1]^2 + acc[idx]^2
	const int win = w < 3 ? w : 3;
	float best = 0.0f;
	int best_idx = 0x7fffffff;
	auto scan = [&](int idx, bool check) {
		const float a0 = acc[idx], a1 = acc[idx - 1], a2 = acc[idx - 2];
		// oldest sample first, products rounded separately as the C path does (no FMA contraction)
		const float val = __fadd_rn(__fadd_rn(__fmul_rn(a2, a2), __fmul_rn(a1, a1)), __fmul_rn(a0, a0));
		if (val > best && (!check || idx < w)) {
			best = val;
			best_idx = idx;
		}
	};
	if (ROWS > 0) {
#pragma unroll
		for (int r = 0; r < ROWS; r++)
			scan(lane + 32 * r, r == ROWS - 1);
	} else {
#pragma unroll 1
		for (int idx = lane; idx < w; idx += 32)
			scan(idx, false);
	}
	{	// warp argmax (largest value, lowest index on ties): energies are >= +0, so their bit patterns
		// order like unsigned integers and two redux instructions replace the shuffle tree
		const unsigned vb = __float_as_uint(best);
		const unsigned mx = __reduce_max_sync(0xffffffffu, vb);
		best_idx = (int)__reduce_min_sync(0xffffffffu, vb == mx ? (unsigned)best_idx : 0xffffffffu);
		best = __uint_as_float(mx);
	}
	int max_idx = (best > 0.0f) ? best_idx - win + 1 : 0;
	if (max_idx < 0)
		max_idx = 0;

	// strongest sample of the winning window (first one on ties); reads past w find the zero padding
	int mwi = max_idx;
	{
		const float b0 = acc[max_idx], b1 = acc[max_idx + 1], b2 = acc[max_idx + 2];
		const float e0 = b0 * b0, e1 = b1 * b1, e2 = b2 * b2;
		float mv = e0;
		if (win > 1 && e1 > mv) {
			mv = e1;
			mwi = max_idx + 1;
		}
		if (win > 2 && e2 > mv)
			mwi = max_idx + 2;
	}

#if DM_OPT_RADIX
	// The search starts at mwi-1 and moves by less than 1 in total, so floor(early) is mwi-2 or mwi-1 (mwi-1 .. mwi
	// for the final interpolation) and only acc[mwi-12 .. mwi+12] is ever read: copy that window (zero outside
	// the vector, as the interpolation treats it) to aw[0..24], aw[25..27] = 0.
	__syncwarp();
	if (lane < 28)
		aw[lane] = lane < 25 ? ld_acc(acc, mwi - 12 + lane, w) : 0.0f;
	__syncwarp();
	const float fbase = (float)(mwi - 2);

	// Step 1 sits on an integer position: the interpolation there is the sample itself.
	float early = (float)(mwi - 1);
	bool live;
	{
		const float e = aw[11], l = aw[13];
		const float e2 = e * e, l2 = l * l;
		live = e2 != l2;
		early += e2 < l2 ? 0.5f : (live ? -0.5f : 0.0f);
	}
	// Steps 2..9 (incr = 1/4 .. 1/512, the reference stops when incr <= 1/1024) visit positions with a
	// fractional part f in (0, 1): every sinc weight is sin(pi f) * -(-1)^j / (pi (j - f)), and the early and
	// the late gate share f, so the comparison e^2 < l^2 only needs  sum_j -(-1)^j acc[.+j] / (j - f)
	// for the two gates - no sine, one reciprocal per tap.
	// Three steps at a time: steps 2-4 can only visit early + m/8, m in {0, +-2, +-1, +-3} (then +- 1/16), steps
	// 5-7 the same on a grid of 1/64, steps 8-9 on 1/512.  Each round evaluates the 8 grid points in parallel -
	// one quad of lanes per point, 5-6 taps per lane - and then walks the three decisions on the ballots: same
	// comparisons, same result, three dependent rounds instead of eight.
	const int q = lane & 3;
	const float fm = (float)((lane >> 2) - 3);                  // grid point of this quad (m = 4 is never visited)
	const float fq = (float)(q - 10), sgn = (q & 1) ? 1.0f : -1.0f;     // first tap of this lane, -(-1)^j (j = q - 10 + 4k)
	float h = 0.125f;
	DM_UNROLL(DM_OPT_WALK ? 3 : 1)
	for (int round = 0; round < 3; round++, h *= 0.125f) {
		if (!live)
			break;
		const float pe = fmaf(fm, h, early);
		const float fl = floorf(pe), f = pe - fl;
		const float *ap = aw + q + (fl != fbase ? 1 : 0);       // early gate reads acc[floor(pe) + j], late gate + 2
		float te = 0.0f, tl = 0.0f;
#pragma unroll
		for (int k = 0; k < 6; k++) {
			float r;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((fq + (float)(4 * k)) - f));
			if (k == 5)
				r = q == 0 ? r : 0.0f;                          // j = q + 10: only tap 10 exists
			te = fmaf(ap[4 * k], r, te);
			tl = fmaf(ap[4 * k + 2], r, tl);
		}
		te *= sgn;
		tl *= sgn;
		te += __shfl_xor_sync(0xffffffffu, te, 1);
		tl += __shfl_xor_sync(0xffffffffu, tl, 1);
		te += __shfl_xor_sync(0xffffffffu, te, 2);
		tl += __shfl_xor_sync(0xffffffffu, tl, 2);
		const float e2 = te * te, l2 = tl * tl;
		const unsigned right = __ballot_sync(0xffffffffu, e2 < l2), dead = __ballot_sync(0xffffffffu, e2 == l2);
		// walk: bit 4*(m+3) of the ballots belongs to grid point m.
#if DM_OPT_WALK
		if (dead == 0u) {
			// usual case, no exact tie anywhere on the grid: step a moves +-2, step b +-1, step c +-1/2 (the last
			// round has only steps a and b)
			const unsigned s1 = (right & 0x1000u) ? 20u : 4u;
			float mf = (right & 0x1000u) ? 2.0f : -2.0f;
			const bool r1 = (right >> s1) & 1u;
			const unsigned s2 = r1 ? s1 + 4u : s1 - 4u;
			mf += r1 ? 1.0f : -1.0f;
			if (round < 2)
				mf += ((right >> s2) & 1u) ? 0.5f : -0.5f;
			early = fmaf(mf, h, early);
			continue;
		}
#endif
		// stop = first level (0, 1, 2) whose point is dead (3: none); moves of levels past it are dropped.
		const unsigned r0 = (right >> 12) & 1u, d0 = (dead >> 12) & 1u;
		const int m1 = r0 ? 2 : -2;
		const unsigned s1 = 4u * (unsigned)(m1 + 3);
		const unsigned r1 = (right >> s1) & 1u, d1 = (dead >> s1) & 1u;
		const int m2 = m1 + (r1 ? 1 : -1);
		const unsigned s2 = 4u * (unsigned)(m2 + 3);
		const unsigned r2 = (right >> s2) & 1u, d2 = ((dead >> s2) & 1u) | (round == 2 ? 1u : 0u);
		const int m = d0 ? 0 : (d1 ? m1 : m2);
		const float half = (d0 | d1 | d2) ? 0.0f : (r2 ? 0.5f : -0.5f);
		live = !(d0 | d1 | (round < 2 ? d2 : 0u));
		early = fmaf((float)m + half, h, early);
	}
	const float pos = early + 1.0f;
	{
		// value at the peak: the full sinc weights (osmo_sinc: 1 within |x| < 0.01), one sine from the table
		const float fl = floorf(pos), frac = pos - fl;
		const float S = c_sinpi512[(int)(frac * 512.0f)];
		const float x = fmaf(-PI_F, frac, tp.xj);         // pi*(j - frac)
		const float wgt = fabsf(x) >= 0.01f ? __fdividef(S, x) : tp.sgn;
		const int sel = (int)(fl - fbase);            // 1, 2 (or 3 when early ended on mwi exactly)
		const float cv = lane < 21 ? aw[lane + sel] : 0.0f;
		peak_val = warp_sum(tp.sgn * cv * wgt);
	}
	return pos;
#else
	// The search starts at mwi-1 and moves by less than 1 in total, so floor(early) is mwi-2 or
	// mwi-1 (mwi-1 .. mwi for the final interpolation): tap j of this lane only ever reads
	// acc[mwi-2+j .. mwi+2+j].  Preload those five values once, with the sign -(-1)^j of the tap folded in.
	const int kb = mwi - 2 + tp.j;
	const float c0 = tp.sgn * ld_acc(acc, kb, w), c1 = tp.sgn * ld_acc(acc, kb + 1, w), c2 = tp.sgn * ld_acc(acc, kb + 2, w),
	            c3 = tp.sgn * ld_acc(acc, kb + 3, w), c4 = tp.sgn * ld_acc(acc, kb + 4, w);
	const float fbase = (float)(mwi - 2), fj = (float)tp.j;
	const bool up = lane & 16;

	// Step 1 sits on an integer position: the interpolation there is the sample itself.
	float early = (float)(mwi - 1), incr = 0.25f;
	bool live;
	{
		const float e = ld_acc(acc, mwi - 1, w), l = ld_acc(acc, mwi + 1, w);
		const float e2 = e * e, l2 = l * l;
		live = e2 != l2;
		early += e2 < l2 ? 0.5f : (live ? -0.5f : 0.0f);
	}
	// Steps 2..9 (incr = 1/4 .. 1/512, the reference stops when incr <= 1/1024) visit positions with a
	// fractional part f in (0, 1): every sinc weight is sin(pi f) * -(-1)^j / (pi (j - f)), and the early and
	// the late gate share f, so the comparison e^2 < l^2 only needs  sum_j -(-1)^j acc[.+j] / (j - f)
	// for the two gates - no sine, one reciprocal per tap.
	if (live) {
#pragma unroll 1
		for (int it = 1; it < 9; it++) {
			const float fl = floorf(early);
			float r;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fj - (early - fl)));
			const bool hi = fl != fbase;                  // floor(early) == mwi-1
			const float te = (hi ? c1 : c0) * r, tl = (hi ? c3 : c2) * r;      // early gate, late gate (+2)
			// one folded shuffle tree for both sums: lower half-warp ends with early, upper with late
			float v = (up ? tl : te) + __shfl_xor_sync(0xffffffffu, up ? te : tl, 16);
#pragma unroll
			for (int o = 8; o; o >>= 1)
				v += __shfl_xor_sync(0xffffffffu, v, o);
			const float ov = __shfl_xor_sync(0xffffffffu, v, 16);     // the other gate
			const float e2 = up ? ov * ov : v * v, l2 = up ? v * v : ov * ov;
			if (e2 == l2)
				break;
			early += e2 < l2 ? incr : -incr;
			incr *= 0.5f;
		}
	}
	const float pos = early + 1.0f;
	{
		// value at the peak: the full sinc weights (osmo_sinc: 1 within |x| < 0.01), one sine from the table
		const float fl = floorf(pos), frac = pos - fl;
		const float S = c_sinpi512[(int)(frac * 512.0f)];
		const float x = fmaf(-PI_F, frac, tp.xj);         // pi*(j - frac)
		const float wgt = fabsf(x) >= 0.01f ? __fdividef(S, x) : tp.sgn;
		const int sel = (int)(fl - fbase);            // 1, 2 (or 3 when early ended on mwi exactly)
		const float cv = sel <= 1 ? c1 : (sel == 2 ? c2 : (sel == 3 ? c3 : c4));
		peak_val = warp_sum(cv * wgt);
	}
	return pos;
#endif
}

// c += s * v on the packed FP32 pipe (one FFMA2 instead of two FFMA)
__device__ __forceinline__ void fma2s(float2 &c, float sc, const float2 v)
{
	unsigned long long cc = *reinterpret_cast<unsigned long long *>(&c);
	const float2 ss = make_float2(sc, sc);
	asm("fma.rn.f32x2 %0, %1, %2, %0;"
	    : "+l"(cc)
	    : "l"(*reinterpret_cast<const unsigned long long *>(&ss)), "l"(*reinterpret_cast<const unsigned long long *>(&v)));
	c = *reinterpret_cast<float2 *>(&cc);
}

// Soft bits of one data symbol (pi4cxpsk.c:468-503) for the symbol value sv = angle / (2*pi / 2^NB):
// nearest symbol sp, distance to it d = round(128 * |round(sv) - sv|), each bit 127 - d (the bit that flips
// towards the second-nearest symbol) or 127 - d/2 (the others), sign by the Gray bit.  NB == 2: first bit in
// the low byte, second in the high byte.
template <int NB>
__device__ __forceinline__ unsigned soft_word(float sv)
{
	constexpr int mask = (1 << NB) - 1;
	constexpr float period = (float)(1 << NB), inv_period = 1.0f / period;
	sv = fmaf(-period, rintf(sv * inv_period), sv);        // -> [-period/2, period/2]
	const float svr = rintf(sv);
	const int sp = (int)svr & mask;
	const bool below = svr > sv;                   // second-nearest symbol is sp-1, else sp+1
	const int dq = __float2int_rn(128.0f * fabsf(svr - sv));
	const int v_far = 127 - dq, v_near = 127 - (dq >> 1);
	if (NB == 2) {
		// Gray map {00, 01, 11, 10}, MSB first.  sp -> sp+1 flips the LSB when sp is even, the MSB
		// when sp is odd; sp -> sp-1 the other way round.
		const int gp = sp ^ (sp >> 1);
		const bool msb_flips = ((sp & 1) != 0) != below;
		const int m1 = msb_flips ? v_far : v_near, m0 = msb_flips ? v_near : v_far;
		const int b1 = (gp & 2) ? -m1 : m1, b0 = (gp & 1) ? -m0 : m0;
		return (unsigned)((b1 & 0xff) | ((b0 & 0xff) << 8));
	}
	return (unsigned)((sp ? -v_far : v_far) & 0xff);       // one bit per symbol: both neighbours flip it
}

// The soft word is piecewise constant in sv with every breakpoint on a multiple of 1/256 (symbol decision at
// k + 1/2, `below` at k, the rounding of d at k +- (2n+1)/256), and periodic with 2^NB.  So it is a table
// over floor(256 * sv) mod 256 * 2^NB, filled once per CTA from soft_word() at the cell centres; per symbol
// that leaves a multiply, a float-to-int, a mask and a shared-memory load.
static constexpr int LUT_CELLS = 256;

// upload this translation unit's copy of the sine table (once per device; the caller serialises)
static inline cudaError_t upload_sinpi512()
{
	float h[513];
	for (int k = 0; k <= 512; k++)
		h[k] = (float)sin(3.14159265358979323846 * (double)k / 512.0);
	return cudaMemcpyToSymbol(c_sinpi512, h, sizeof(h));
}

}  // namespace gmr1
