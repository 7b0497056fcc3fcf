// tests/emu/chan_emu.cpp - TEST HARNESS: runs the product's register-butterfly FFT (osmo_gmr_b200/csrc/chan_fft.cuh,
// the exact functions pfb_fast_kernel and the FCCH search inline) on the CPU, one "thread" after the other with the
// kernels' barrier structure (all loads of a stage, then all its stores), so that butterflies and index arithmetic
// can be checked against numpy without a GPU.  Not part of the product library; nothing in the product calls it.
#include <math.h>
#include <vector>
#include "chan_fft.cuh"

using namespace gmr1::cfft;

namespace {

struct Tw {
	const cfl *t;
	cfl operator()(int i) const { return t[i]; }
};

template <int LOG2N, int S> void stage16(cfl *row, Tw tw)
{
	typedef Plan<LOG2N> P;
	typedef Stage<P::N, P::pow16(S), 16> St;
	std::vector<cfl> keep(P::N);
	for (int t = 0; t < P::TPR; t++) {             // every thread loads and computes ...
		cfl v[16];
		St::read(row, t, tw, v);
		for (int q = 0; q < 16; q++)
			keep[t * 16 + q] = v[q];
	}
	for (int t = 0; t < P::TPR; t++) {             // ... barrier ... then stores
		cfl v[16];
		for (int q = 0; q < 16; q++)
			v[q] = keep[t * 16 + q];
		St::write(row, t, v);
	}
}

template <int LOG2N, int R, int NS> void stage_out(const cfl *row, Tw tw, cfl *out)
{
	typedef Plan<LOG2N> P;
	typedef Stage<P::N, NS, R> St;
	for (int t = 0; t < P::TPR; t++)
		for (int i = 0; i < 16 / R; i++) {
			const int j = t + i * P::TPR;
			cfl v[R];
			St::read(row, j, tw, v);
			for (int q2 = 0; q2 < R; q2++)
				out[St::out_index(j, q2)] = v[q2];
		}
}

template <int LOG2N> void fft(const float *in, float *out, int sign)
{
	typedef Plan<LOG2N> P;
	constexpr int N = P::N, N16 = P::N16, RL = P::RLAST;
	std::vector<cfl> tw(N), row(RowStride<N>::value), res(N);
	for (int t = 0; t < N; t++)
		tw[t] = cf((float)cos(2.0 * M_PI * t / N), (float)(sign * sin(2.0 * M_PI * t / N)));
	for (int i = 0; i < N; i++)
		row[pad(i)] = cf(in[2 * i], sign * in[2 * i + 1]);       // forward transform = conj(reverse(conj x))
	Tw w = {tw.data()};
	if constexpr (RL > 1) {
		if constexpr (N16 >= 1) stage16<LOG2N, 0>(row.data(), w);
		if constexpr (N16 >= 2) stage16<LOG2N, 1>(row.data(), w);
		if constexpr (N16 >= 3) stage16<LOG2N, 2>(row.data(), w);
		stage_out<LOG2N, RL, P::pow16(N16)>(row.data(), w, res.data());
	} else {
		if constexpr (N16 >= 2) stage16<LOG2N, 0>(row.data(), w);
		if constexpr (N16 >= 3) stage16<LOG2N, 1>(row.data(), w);
		if constexpr (N16 >= 4) stage16<LOG2N, 2>(row.data(), w);
		stage_out<LOG2N, 16, P::pow16(N16 - 1)>(row.data(), w, res.data());
	}
	for (int i = 0; i < N; i++) {
		out[2 * i] = res[i].x;
		out[2 * i + 1] = sign * res[i].y;
	}
}

}  // namespace

// reverse (sign = +1: e^{+j...}, what the bank runs) transform of 2^log2n complex floats; 0 / -1 (size not built)
extern "C" int chan_emu_fft(int log2n, const float *in, float *out)
{
	switch (log2n) {
	case 4: fft<4>(in, out, 1); return 0;
	case 5: fft<5>(in, out, 1); return 0;
	case 6: fft<6>(in, out, 1); return 0;
	case 7: fft<7>(in, out, 1); return 0;
	case 8: fft<8>(in, out, 1); return 0;
	case 9: fft<9>(in, out, 1); return 0;
	case 10: fft<10>(in, out, 1); return 0;
	case 11: fft<11>(in, out, 1); return 0;
	case 12: fft<12>(in, out, 1); return 0;
	case 13: fft<13>(in, out, 1); return 0;
	}
	return -1;
}

// ---- A5/1, 32 streams per "thread" (osmo_gmr_b200/csrc/a5_bitslice.cuh) on the CPU -----------------------------------
#include <stdint.h>
#include <string.h>
#include "a5_bitslice.cuh"

// keys [32][8], fn [32] -> dl / ul [32][nbits] ubits (ul may be NULL), exactly as a5_slice_kernel produces them
extern "C" void a5_emu_slice(const uint8_t *keys, const uint32_t *fn, int nbits, uint8_t *dl, uint8_t *ul)
{
	using namespace gmr1::a5s;
	State s;
	init(s);
	uint64_t fk[32];
	for (int l = 0; l < 32; l++)
		fk[l] = folded_key(keys + 8 * l, fn[l]);
	for (int q = 0; q < 64; q++) {
		uint32_t kb = 0;
		for (int l = 0; l < 32; l++)
			kb |= (uint32_t)((fk[l] >> q) & 1u) << l;
		key_step(s, kb);
	}
	force_bit0(s);
	for (int i = 0; i < 250; i++)
		clock(s);
	for (int dir = 0; dir < 2; dir++) {
		uint8_t *dst = dir ? ul : dl;
		for (int b0 = 0; b0 < nbits; b0 += 32) {
			uint32_t w[32];
			for (int c = 0; c < 32; c++) {
				if (b0 + c < nbits) {
					clock(s);
					w[c] = output(s);
				} else
					w[c] = 0;
			}
			transpose32(w);
			if (dst)
				for (int l = 0; l < 32; l++)
					for (int b = 0; b < 32 && b0 + b < nbits; b += 4) {
						const uint32_t v = spread4(w[l], b);
						for (int k = 0; k < 4 && b0 + b + k < nbits; k++)
							dst[(size_t)l * nbits + b0 + b + k] = (uint8_t)(v >> (8 * k));
					}
		}
		if (!ul)
			break;
	}
}
