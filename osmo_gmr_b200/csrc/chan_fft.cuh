// chan_fft.cuh - the reverse FFT of the channeliser's analysis bank for power-of-two channel counts, as register
// butterflies: a "thread" holds 16 points, does a radix-16 butterfly (4 x 4) in registers and exchanges through one
// row of shared memory between stages (Stockham autosort: natural order in, natural order out, no bit reversal).
// N = 16^a * r, r in {1, 2, 4, 8}: a radix-16 stages, then one stage of 16 / r radix-r butterflies per thread.
// N / 16 threads work on one transform.
//
// Everything here is __host__ __device__: pfb_fast_kernel (chan_kernels.cu) inlines it, and tests/emu/chan_emu.cpp runs
// the same functions on the CPU, one "thread" after the other with the kernel's barrier structure, against numpy
// (tests/test_chan_cpu.py) - index arithmetic and butterflies are checked without a GPU.
//
// Transform: X[k] = sum_n x[n] e^{+j 2 pi n k / N}  (the sign utils/gmr1_rx_sdr.py's pfb.channelizer_ccf uses: GNU
// Radio's polyphase channeliser runs a reverse FFT over the branch outputs).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define CF_HD __host__ __device__ __forceinline__
namespace gmr1 { namespace cfft { typedef float2 cfl; } }
#else
#define CF_HD inline
namespace gmr1 { namespace cfft { struct cfl { float x, y; }; } }
#endif

namespace gmr1 {
namespace cfft {

CF_HD cfl cf(float x, float y) { cfl r; r.x = x; r.y = y; return r; }
// complex add / subtract: on the device ONE packed instruction each (FADD2; a - b as the exact FFMA2 b * (-1, -1) + a) -
// half of a radix-16 butterfly's instructions are these; the host build (tests/emu) computes the same IEEE results
#if defined(__CUDA_ARCH__)
CF_HD cfl cadd(cfl a, cfl b) { return __fadd2_rn(a, b); }
CF_HD cfl csub(cfl a, cfl b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
#else
CF_HD cfl cadd(cfl a, cfl b) { return cf(a.x + b.x, a.y + b.y); }
CF_HD cfl csub(cfl a, cfl b) { return cf(a.x - b.x, a.y - b.y); }
#endif
CF_HD cfl cmul(cfl a, cfl b) { return cf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
CF_HD cfl csqr(cfl a) { return cf(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }

// shared-memory index of point i of a row: one pad slot per 16 points, so that the stride-16 stores of the first
// stage (thread j writes points 16 j + q) fall on 17 j + q - all banks, no conflict
CF_HD int pad(int i) { return i + (i >> 4); }
template <int N> struct RowStride { static constexpr int value = N + (N >> 4); };

// 4-point reverse DFT in place: a, b, c, d <- X0, X1, X2, X3
CF_HD void dft4(cfl &a, cfl &b, cfl &c, cfl &d)
{
	const cfl s02 = cadd(a, c), d02 = csub(a, c), s13 = cadd(b, d), d13 = csub(b, d);
	a = cadd(s02, s13);
	b = cf(d02.x - d13.y, d02.y + d13.x);          // d02 + j d13
	c = csub(s02, s13);
	d = cf(d02.x + d13.y, d02.y - d13.x);          // d02 - j d13
}

// multiply by e^{+j 2 pi M / 16}
template <int M> CF_HD cfl mul_w16(cfl v)
{
	constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, H = 0.70710678118654752440f;
	if (M == 0) return v;
	if (M == 4) return cf(-v.y, v.x);
	if (M == 2) return cf((v.x - v.y) * H, (v.x + v.y) * H);
	if (M == 6) return cf((-v.x - v.y) * H, (v.x - v.y) * H);
	if (M == 1) return cmul(v, cf(C1, S1));
	if (M == 3) return cmul(v, cf(S1, C1));
	if (M == 9) return cmul(v, cf(-C1, -S1));
	return v;                                      // not used
}

template <int R> struct Dft;                       // in place, natural order out

template <> struct Dft<2> {
	CF_HD static void run(cfl (&v)[2])
	{
		const cfl a = v[0], b = v[1];
		v[0] = cadd(a, b);
		v[1] = csub(a, b);
	}
};

template <> struct Dft<4> {
	CF_HD static void run(cfl (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
};

template <> struct Dft<8> {                       // n = a + 2 c, k = b + 4 d
	CF_HD static void run(cfl (&v)[8])
	{
		dft4(v[0], v[2], v[4], v[6]);              // t[0][b] at v[2 b]
		dft4(v[1], v[3], v[5], v[7]);              // t[1][b] at v[2 b + 1]
		v[3] = mul_w16<2>(v[3]);
		v[5] = mul_w16<4>(v[5]);
		v[7] = mul_w16<6>(v[7]);
		const cfl x0 = cadd(v[0], v[1]), x4 = csub(v[0], v[1]), x1 = cadd(v[2], v[3]), x5 = csub(v[2], v[3]);
		const cfl x2 = cadd(v[4], v[5]), x6 = csub(v[4], v[5]), x3 = cadd(v[6], v[7]), x7 = csub(v[6], v[7]);
		v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
	}
};

template <> struct Dft<16> {                      // n = a + 4 c, k = b + 4 d
	CF_HD static void run(cfl (&v)[16])
	{
		dft4(v[0], v[4], v[8], v[12]);             // t[a][b] at v[a + 4 b]
		dft4(v[1], v[5], v[9], v[13]);
		dft4(v[2], v[6], v[10], v[14]);
		dft4(v[3], v[7], v[11], v[15]);
		v[5] = mul_w16<1>(v[5]);   v[6] = mul_w16<2>(v[6]);   v[7] = mul_w16<3>(v[7]);
		v[9] = mul_w16<2>(v[9]);   v[10] = mul_w16<4>(v[10]); v[11] = mul_w16<6>(v[11]);
		v[13] = mul_w16<3>(v[13]); v[14] = mul_w16<6>(v[14]); v[15] = mul_w16<9>(v[15]);
		dft4(v[0], v[1], v[2], v[3]);              // X[b + 4 d] at v[4 b + d]
		dft4(v[4], v[5], v[6], v[7]);
		dft4(v[8], v[9], v[10], v[11]);
		dft4(v[12], v[13], v[14], v[15]);
		cfl w[16];
#pragma unroll
		for (int b = 0; b < 4; b++)
#pragma unroll
			for (int d = 0; d < 4; d++)
				w[b + 4 * d] = v[4 * b + d];
#pragma unroll
		for (int i = 0; i < 16; i++)
			v[i] = w[i];
	}
};

// One Stockham stage of radix R on a transform of size N, NS = product of the radices before it.  Butterfly j
// (0 <= j < N / R) reads the points j + q N / R, multiplies point q by e^{+j 2 pi q k / (NS R)}, k = j mod NS, and
// its output q2 is point (j - k) R + k + q2 NS of the next stage.  tw = e^{+j 2 pi t / N}, t < N.
template <int N, int NS, int R> struct Stage {
	static constexpr int NB = N / R;
	template <class TW> CF_HD static void read(const cfl *row, int j, TW tw, cfl (&v)[R])
	{
		read_ld([row](int i) { return row[pad(i)]; }, j, tw, v);
	}
	// the same with the points fetched by ld(point index) - a stage fused with whatever produces its input
	template <class LD, class TW> CF_HD static void read_ld(LD ld, int j, TW tw, cfl (&v)[R])
	{
#pragma unroll
		for (int q = 0; q < R; q++)
			v[q] = ld(j + q * NB);
		if (NS > 1) {
			// stage twiddles w^q, w = e^{+j 2 pi k / (NS R)}: ONE table load per butterfly, the powers by squaring and
			// multiplying (depth <= 4: ~3e-7 relative).  Fifteen loads of tw(q k STEP) would be fifteen gathers over up
			// to 32 cache lines each - measured: that, not arithmetic, bound the first version of these kernels.
			const int k = j & (NS - 1);
			constexpr int STEP = N / (NS * R);
			cfl w[R];
			w[1] = tw(k * STEP);
#pragma unroll
			for (int q = 2; q < R; q++)
				w[q] = (q & 1) ? cmul(w[q - 1], w[1]) : csqr(w[q >> 1]);
#pragma unroll
			for (int q = 1; q < R; q++)
				v[q] = cmul(v[q], w[q]);
		}
		Dft<R>::run(v);
	}
	CF_HD static int out_index(int j, int q2)
	{
		const int k = j & (NS - 1);
		return (j - k) * R + k + q2 * NS;
	}
	CF_HD static void write(cfl *row, int j, const cfl (&v)[R])
	{
#pragma unroll
		for (int q2 = 0; q2 < R; q2++)
			row[pad(out_index(j, q2))] = v[q2];
	}
};

// The stage plan of a 2^LOG2N-point transform
template <int LOG2N> struct Plan {
	static constexpr int N = 1 << LOG2N;
	static constexpr int N16 = LOG2N / 4;          // radix-16 stages
	static constexpr int RLAST = 1 << (LOG2N % 4); // radix of the closing stage (1 = none)
	static constexpr int TPR = N / 16;             // threads per transform
	static constexpr int STAGES = N16 + (RLAST > 1 ? 1 : 0);
	static constexpr int pow16(int e) { return e <= 0 ? 1 : 16 * pow16(e - 1); }
};

}  // namespace cfft
}  // namespace gmr1
