// decode_kernels.cu - stage 3 of the receive path on the GPU: batched channel decode.
//
//   decode_tpc_kernel<CH>  K5 / K7 codes, one thread per codeword, 128 codewords per CTA.
//       phase 1  soft-bit rows of the CTA's 128 units -> shared memory
//                (un-ciphered channels: one TMA bulk copy of the contiguous tile; ciphered /
//                 TCH9: cooperative coalesced loop that applies the cipher sign / gathers the
//                 three bursts the inter-burst deinterleaver spans)
//       phase 2  per thread: gather program -> ACS in registers -> decisions to shared memory
//                -> traceback -> CRC -> packed L2 (decode_unit.cuh)
//   decode_dc12_kernel     K9 (256 states), one warp per codeword, path metrics and
//                survivor bits in shared memory, 8 states per lane.
//
// Replaces gmr1_{bcch,ccch,facch3,facch9,tch3,tch9,rach,xch_dc12}_decode + osmo_conv_decode +
// osmo_crc*gen_check_bits of the reference (see decode_unit.cuh for file:line).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>

#include "decode_unit.cuh"
#include "tma.cuh"
#include "launch.h"
#ifndef T9_E
#define T9_E 16        // TCH9 gather: elements per thread whose loads are issued together (8: 0.79, 16: 0.75 ms per 157 284 bursts)
#endif
#ifndef TPC_LUT
#define TPC_LUT 1       // one codeword per thread: branch metrics through the shared-memory table (0: arithmetic)
#endif

namespace gmr1 {

static constexpr int TPC_T = 128;    // codewords per CTA (upper bound; tpc_tile(ch) is the tile of a channel)
#ifndef TPC_T_XCCH
#define TPC_T_XCCH 128
#endif
#ifndef TPC_T_RACH
#define TPC_T_RACH 128
#endif
#ifndef TPC_T_BIG
#define TPC_T_BIG 32                 // tile of the channels with 648 / 662-byte rows
#endif
// threads per CTA: one per codeword, or - PAIR - one per two codewords (viterbi_p16.cuh: thread i takes codewords
// i and i + 64 of the tile)
__host__ __device__ constexpr int tpc_tile(int ch);
__host__ __device__ constexpr int tpc_threads(bool pair, int ch) { return pair ? tpc_tile(ch) / 2 : tpc_tile(ch); }

// ---- device tables --------------------------------------------------------------------------
__constant__ uint16_t c_g[CH_COUNT][MAX_CODED];     // gather programs (uniform access)
__constant__ uint16_t c_rach_g2[MAX_CODED];
__device__ int16_t  d_cmap[CH_COUNT][MAX_EBITS];     // per-lane access -> global
__device__ uint16_t d_t9_src[648];

static bool g_tables_up[64] = {false};

static cudaError_t upload_tables()
{
	GMR1_INIT_LOCK();
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev < 64 && g_tables_up[dev])
		return cudaSuccess;
	for (int ch = 0; ch < CH_COUNT; ch++) {
		const ChanTab &t = chan_tab(ch);
		if ((e = cudaMemcpyToSymbol(c_g, t.g, sizeof(t.g), sizeof(t.g) * ch)) != cudaSuccess) return e;
		if ((e = cudaMemcpyToSymbol(d_cmap, t.cmap, sizeof(t.cmap), sizeof(t.cmap) * ch)) != cudaSuccess) return e;
	}
	if ((e = cudaMemcpyToSymbol(c_rach_g2, chan_tab(CH_RACH).g2, sizeof(c_rach_g2))) != cudaSuccess) return e;
	if ((e = cudaMemcpyToSymbol(d_t9_src, chan_tab(CH_TCH9_9K6).t9_src, sizeof(d_t9_src))) != cudaSuccess) return e;
	if (dev < 64)
		g_tables_up[dev] = true;
	return cudaSuccess;
}

// ---- compile-time geometry per channel ---------------------------------------------------------
__host__ __device__ constexpr int chan_n_in(int ch)
{
	return ch == CH_BCCH ? 424 : ch == CH_CCCH ? 432 : ch == CH_FACCH3 ? 416 :
	       ch == CH_RACH ? 494 : ch == CH_TCH3 ? 212 : ch == CH_DC12 ? 432 : 662;
}
__host__ __device__ constexpr bool chan_is_t9(int ch)
{
	return ch == CH_TCH9_2K4 || ch == CH_TCH9_4K8 || ch == CH_TCH9_9K6;
}
__host__ __device__ constexpr int chan_n_row(int ch) { return chan_is_t9(ch) ? 648 : chan_n_in(ch); }
__host__ __device__ constexpr int chan_len(int ch)
{
	return ch == CH_BCCH || ch == CH_CCCH || ch == CH_DC12 ? 208 : ch == CH_FACCH3 ? 92 :
	       ch == CH_FACCH9 ? 316 : ch == CH_TCH9_2K4 ? 144 : ch == CH_TCH9_4K8 ? 240 :
	       ch == CH_TCH9_9K6 ? 480 : ch == CH_RACH ? 159 : 48;
}
__host__ __device__ constexpr int chan_n_ciph(int ch)
{
	return ch == CH_FACCH3 ? 384 : (ch == CH_FACCH9 || chan_is_t9(ch)) ? 658 : ch == CH_TCH3 ? 208 : 0;
}
__host__ __device__ constexpr int chan_n_steps(int ch)
{
	return chan_len(ch) + ((ch == CH_TCH3 || ch == CH_DC12) ? 0 : 4);
}
// decision bytes per thread
__host__ __device__ constexpr int chan_dec_bytes(int ch)
{
	return ch == CH_TCH3 ? 48 * 8 : chan_n_steps(ch) * 2;
}
// Codewords per CTA.  FACCH9 / TCH9 rows are 662 / 648 bytes: 128 of them (83 KB) leave two CTAs = 8 warps per SM,
// tiles of 32 (21 KB) ten CTAs = 10 warps and a finer interleave of the staging and the trellis phases of
// neighbouring CTAs: FACCH9 0.352 -> 0.293 (64) -> 0.255 ms (32), TCH9-9k6 0.751 -> 0.647 -> 0.634 ms per 157 284
// bursts.  The 424 / 432-byte BCCH / CCCH rows and RACH (494) are best at 128 (A/B on one box: 64 and 32 lose 1-5 %).
__host__ __device__ constexpr int tpc_tile(int ch)
{
	return (ch == CH_FACCH9 || chan_is_t9(ch)) ? TPC_T_BIG : ch == CH_RACH ? TPC_T_RACH :
	       (ch == CH_BCCH || ch == CH_CCCH) ? TPC_T_XCCH : TPC_T;
}
__host__ __device__ constexpr int tpc_rows_bytes(int ch) { return (tpc_tile(ch) * chan_n_row(ch) + 15) & ~15; }
__host__ __device__ constexpr int tpc_smem_bytes(int ch)
{
	return tpc_rows_bytes(ch) + tpc_tile(ch) * chan_dec_bytes(ch) + 16 + (int)sizeof(P16Lut);
}

// ---- thread-per-codeword kernel ------------------------------------------------------------------
template <int CH, bool PAIR>
__global__ void __launch_bounds__(tpc_threads(PAIR, CH)) decode_tpc_kernel(const DecodeArgs a)
{
	extern __shared__ __align__(16) uint8_t smem[];
	constexpr int NIN = chan_n_in(CH), NROW = chan_n_row(CH), NT = tpc_threads(PAIR, CH), TT = tpc_tile(CH);

	int8_t *rows = (int8_t *)smem;                                    // [TT][NROW], unpadded
	// survivor decisions [steps][words][slots]: in the caller's scratch (slot = unit, all units of the launch
	// side by side: coalesced 64-byte stores per warp and step) or, without scratch, behind the rows in
	// shared memory (slot = thread)
	const bool gdec = a.dec_scratch != nullptr;
	uint8_t *dec = gdec ? a.dec_scratch : smem + tpc_rows_bytes(CH);
	uint64_t *bar = (uint64_t *)(smem + tpc_rows_bytes(CH) + (gdec ? 0 : TT * chan_dec_bytes(CH)));
	P16Lut *lut = (P16Lut *)(bar + 2);                                // PAIR: soft bit -> packed metrics

	TabRef tb;
	tb.g = c_g[CH];
	tb.g2 = (CH == CH_RACH) ? c_rach_g2 : nullptr;
	tb.cmap = d_cmap[CH];
	tb.t9_src = d_t9_src;
	tb.n_in = NIN; tb.n_row = NROW; tb.n_ciph = chan_n_ciph(CH);
	tb.n_steps = chan_n_steps(CH); tb.len = chan_len(CH);

	const int tid = threadIdx.x;
	const int base = blockIdx.x * TT;
	const int n_eff = a.n_dev ? min(a.n, *a.n_dev) : a.n;
	if (base >= n_eff)
		return;                  // whole CTA, before any barrier
	const int cnt = min(TT, n_eff - base);
	if constexpr (PAIR) {
		for (int i = tid; i < 512; i += NT)
			(&lut->plain[0])[i] = p16_lut_word(i);                    // visible after the barrier that ends phase 1
	} else if constexpr (TPC_LUT) {
		for (int i = tid; i < 512; i += NT)
			rel_lut_fill((RelLut *)lut, i);
	}

	// ---- phase 1: stage the tile
	// ciphered FACCH3 / TCH3 / FACCH9: the rows come in like plain ones, the cipher signs are applied in shared memory
	// by a second pass (below)
	constexpr bool CAN_CIPH = !chan_is_t9(CH) && chan_n_ciph(CH) > 0;
	const bool ciphered = CAN_CIPH && a.ciph != nullptr;
	const bool bulk = !chan_is_t9(CH) && cnt == TT && ((((uintptr_t)a.ebits) & 15) == 0);
	if (bulk) {
		// the tile is one contiguous, 16-byte aligned span of TT*NIN bytes: a single TMA
		// bulk copy brings it in while no LSU instruction is spent on it
		if (tid == 0) {
			mbar_init(bar, 1);
			mbar_fence_init();
			mbar_arrive_expect_tx(bar, (uint32_t)(TT * NIN));
			tma_load_1d(rows, a.ebits + (size_t)base * NIN, (uint32_t)(TT * NIN), bar);
		}
		__syncthreads();
		mbar_wait(bar, 0);
	} else if (chan_is_t9(CH) && !a.t9_rows) {
		// TCH9 gathers three bursts per codeword through two index arrays: the predecessors of the tile's units
		// go to shared memory first, so that an element costs one dependent global load instead of two, and
		// four elements per thread are in flight (this staging, not the Viterbi, bounded the kernel: 8 resident
		// warps per SM and a chain of three dependent loads per byte)
		__shared__ int s_prev[2][TT];
		__shared__ uint16_t s_src[648];
		for (int u = tid; u < TT; u += NT) {
			s_prev[0][u] = (u < cnt && a.prev1) ? a.prev1[base + u] : -1;
			s_prev[1][u] = (u < cnt && a.prev2) ? a.prev2[base + u] : -1;
		}
		for (int r = tid; r < 648; r += NT)
			s_src[r] = tb.t9_src[r];
		__syncthreads();
		// T9_E elements per thread and pass: source addresses first (shared-memory lookups only), then the eight
		// byte loads back to back (clamped, so that they are unconditional), then the stores
		// A ragged last tile takes a second copy of the loop whose unit index is checked against the tile's count
		// (base + tt may lie behind the batch: found by compute-sanitizer on a 6-burst TCH9 batch).  The check is kept
		// out of the full-tile copy: one more select in front of the index lookups cost 0.85 -> 1.18 ms per 157 284
		// bursts (A/B on one box, tools/gpu_r2_t9ab.sh)
		constexpr int E = T9_E;
		auto gather = [&](auto FULL) {
		for (int idx0 = tid; idx0 < TT * NROW; idx0 += NT * E) {
			const int8_t *src[E];
			const uint8_t *csrc[E];
			bool ok[E], flip[E];
#pragma unroll
			for (int e = 0; e < E; e++) {
				const int idx = min(idx0 + e * NT, TT * NROW - 1);
				const int tt = idx / NROW, r = idx - tt * NROW;
				const uint16_t w = s_src[r];
				const int age = (w >> 10) & 3, sidx = w & G_IDX;
				int u;
				if constexpr (decltype(FULL)::value)
					u = age == 0 ? base + tt : s_prev[age - 1][tt];
				else
					u = tt >= cnt ? -1 : age == 0 ? base + tt : s_prev[age - 1][tt];
				ok[e] = tt < cnt && u >= 0;
				flip[e] = (w & G_FLIP) != 0;
				const size_t uu = (size_t)max(u, 0);
				src[e] = a.ebits + uu * NIN + sidx;
				csrc[e] = a.ciph ? a.ciph + uu * tb.n_ciph + tb.cmap[sidx] : nullptr;
			}
			int v[E];
			unsigned c[E];
#pragma unroll
			for (int e = 0; e < E; e++) {
				v[e] = *src[e];
				c[e] = csrc[e] ? *csrc[e] : 0u;
			}
#pragma unroll
			for (int e = 0; e < E; e++) {
				int x = v[e];
				if (c[e])
					x = sbit_neg(x);
				if (flip[e])
					x = sbit_neg(x);
				if (idx0 + e * NT < TT * NROW)
					rows[idx0 + e * NT] = ok[e] ? (int8_t)x : (int8_t)0;
			}
		}
		};
		if (cnt == TT)
			gather(std::true_type{});
		else
			gather(std::false_type{});
		__syncthreads();
	} else if (chan_is_t9(CH)) {
#pragma unroll 4
		for (int idx = tid; idx < TT * NROW; idx += NT) {
			const int tt = idx / NROW, r = idx - tt * NROW;
			rows[idx] = (tt < cnt) ? stage_elem<CH>(tb, a, base + tt, r) : (int8_t)0;
		}
		__syncthreads();
	} else {
		// ragged last tile or unaligned batch: plain coalesced copy (NROW == NIN: the tile is one span of bytes)
#pragma unroll 4
		for (int idx = tid; idx < TT * NROW; idx += NT)
			rows[idx] = idx < cnt * NROW ? a.ebits[(size_t)base * NIN + idx] : (int8_t)0;
		__syncthreads();
	}
	if constexpr (CAN_CIPH) {
		// Cipher signs (tch3.c:137-139, facch3.c:145-153, facch9.c:121-128: a set cipher bit negates the soft bit),
		// applied to the staged rows: cipher position of every soft bit from a shared-memory copy of the map, eight
		// cipher bytes per thread requested together (unconditional, clamped index), then the negations.  The first
		// form did this inside the staging loop, one element at a time with the map read from global memory: two
		// dependent global loads per soft bit, 55 % of the TCH3 kernel's stall samples (ncu, config 3).
		if (ciphered) {
			__shared__ int16_t s_cmap[NIN];
			for (int r = tid; r < NIN; r += NT)
				s_cmap[r] = tb.cmap[r];
			__syncthreads();
			// four soft bits (one word of the staged tile) per step: when their cipher bytes are four consecutive,
			// word-aligned bytes - all of TCH3's but the word of the four unciphered status bits - they come with one
			// 32-bit load, eight of those in flight per thread; other words take byte loads.  rows is 16-byte aligned.
			constexpr int E = 8, NC = chan_n_ciph(CH);
			const int nwords = (cnt * NROW + 3) / 4, total = cnt * NROW;
			const uint8_t *ctile = a.ciph + (size_t)base * NC;
			const uint8_t *safe = (const uint8_t *)(((uintptr_t)ctile + 3) & ~(uintptr_t)3);    // a valid aligned word
			uint32_t *rows4 = (uint32_t *)rows;
			for (int w0 = tid; w0 < nwords; w0 += NT * E) {
				uint32_t cw[E];
				int adr[E][4];
				bool fast[E];
#pragma unroll
				for (int e = 0; e < E; e++) {
					const int w = min(w0 + e * NT, nwords - 1);
#pragma unroll
					for (int k = 0; k < 4; k++) {
						const int idx = 4 * w + k;
						const int tt = idx / NROW, r = idx - tt * NROW;
						const int c = idx < total ? (int)s_cmap[r] : -1;
						adr[e][k] = c >= 0 ? tt * NC + c : -1;
					}
					fast[e] = adr[e][0] >= 0 && adr[e][1] == adr[e][0] + 1 && adr[e][2] == adr[e][0] + 2 &&
					          adr[e][3] == adr[e][0] + 3 && ((((uintptr_t)ctile + (unsigned)adr[e][0]) & 3) == 0);
					cw[e] = *(const uint32_t *)(fast[e] ? ctile + adr[e][0] : safe);
				}
#pragma unroll
				for (int e = 0; e < E; e++) {
					const int w = w0 + e * NT;
					if (w >= nwords)
						continue;
					uint32_t m = cw[e];
					if (!fast[e]) {
						m = 0;
#pragma unroll
						for (int k = 0; k < 4; k++)
							if (adr[e][k] >= 0)
								m |= (uint32_t)ctile[adr[e][k]] << (8 * k);
					}
					if (m) {
						const uint32_t v = rows4[w], sel = __vcmpne4(m, 0u);      // 0xff where the cipher byte is set
						rows4[w] = (__vneg4(v) & sel) | (v & ~sel);              // bytewise negate, -128 stays -128
					}
				}
			}
			__syncthreads();
		}
	}

	// ---- phase 2: one codeword (PAIR: two) per thread
	if constexpr (PAIR) {
		if (tid < cnt) {
			const int T = gdec ? (int)gridDim.x * NT : NT, t = gdec ? (int)blockIdx.x * NT + tid : tid;
			if constexpr (CH == CH_TCH3)
				decode_pair_tch3(tb, a, lut, base + tid, base + tid + NT, tid + NT < cnt, rows + tid * NROW,
				                 rows + (tid + NT) * NROW, (uint32_t *)dec, T, t);
			else
				decode_pair_k5<CH>(tb, a, lut, base + tid, base + tid + NT, tid + NT < cnt, rows + tid * NROW,
				                   rows + (tid + NT) * NROW, (uint32_t *)dec, T, t);
		}
	} else if (tid < cnt) {
		const int T = gdec ? (int)gridDim.x * TT : TT, t = gdec ? base + tid : tid;
		if constexpr (CH == CH_TCH3)
			decode_unit_tch3<TPC_LUT != 0>(tb, a, base + tid, rows + tid * NROW, (uint32_t *)dec, T, t, (const RelLut *)lut);
		else
			decode_unit_k5<CH, TPC_LUT != 0>(tb, a, base + tid, rows + tid * NROW, (uint16_t *)dec, T, t, (const RelLut *)lut);
	}
}

// Two codewords per thread (viterbi_p16.cuh) pay where REGISTERS bound the resident warps: TCH3 (64 path metrics per
// thread, 143 registers, 12 warps per SM).  The K5 channels are bound by the shared memory their soft-bit rows take
// (54 .. 85 KB per 128 codewords): two codewords per thread halve the resident warps there and win nothing (A/B in
// profiles/README.md), so they stay on one codeword per thread.  GMR1B200_DECODE_P16 = 0: one per thread everywhere,
// = 1: two per thread everywhere (results are identical in every setting).
static int decode_p16_mode()
{
	static const int m = [] { const char *e = getenv("GMR1B200_DECODE_P16"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
	return m;
}

template <int CH, bool PAIR = false>
static cudaError_t launch_tpc(const DecodeArgs &a, cudaStream_t st)
{
	if constexpr (!PAIR) {
		const int m = decode_p16_mode();
		if (m == 1 || (m < 0 && CH == CH_TCH3))
			return launch_tpc<CH, true>(a, st);
	}
	GMR1_INIT_LOCK();
	static bool attr_done[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	constexpr int smem = tpc_smem_bytes(CH);
	if (dev >= 64 || !attr_done[dev]) {
		cudaError_t e = cudaFuncSetAttribute(decode_tpc_kernel<CH, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_done[dev] = true;
	}
	const int grid = (a.n + tpc_tile(CH) - 1) / tpc_tile(CH);
	const int smem_used = (a.dec_scratch ? tpc_rows_bytes(CH) + 16 : smem - (int)sizeof(P16Lut)) +
	                      (PAIR ? (int)sizeof(P16Lut) : TPC_LUT ? (int)sizeof(RelLut) : 0);
	decode_tpc_kernel<CH, PAIR><<<grid, tpc_threads(PAIR, CH), smem_used, st>>>(a);
	return cudaGetLastError();
}

// ---- DC12: K9 r1/3 tail-biting, one warp per codeword ------------------------------------------------
// Lane L owns new states 8L..8L+7 = butterflies 4L..4L+3 (old states 4L+i and 128+4L+i).  For
// these generator polynomials (constant and D^8 term in every g) the four branch outputs of a
// butterfly are x, ~x, ~x, x, so one 3-bit x per butterfly is all a lane keeps.
static constexpr int DC12_WARPS = 4;
struct Dc12Smem {
	uint32_t ae[2][256];
	uint8_t  dec[208][32];
	int8_t   row[432];
	uint8_t  out[32];
};

__global__ void __launch_bounds__(DC12_WARPS * 32) decode_dc12_kernel(const DecodeArgs a)
{
	using C = CodeK9_13;
	extern __shared__ __align__(16) uint8_t smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int unit = blockIdx.x * DC12_WARPS + warp;
	if (unit >= (a.n_dev ? min(a.n, *a.n_dev) : a.n))
		return;
	Dc12Smem &s = reinterpret_cast<Dc12Smem *>(smem)[warp];
	const uint16_t *g = c_g[CH_DC12];

	for (int r = lane; r < 432; r += 32)
		s.row[r] = a.ebits[(size_t)unit * 432 + r];
	for (int i = lane; i < 256; i += 32)
		s.ae[0][i] = i ? MAX_AE : 0u;

	unsigned x[4];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		unsigned p = 4 * lane + i, reg = p << 1, ov = 0;
		for (int j = 0; j < 3; j++)
			ov = (ov << 1) | (__popc(reg & C::poly(j)) & 1);
		x[i] = ov;
	}
	__syncwarp();

	int cur = 0;
	for (int pass = 0; pass < 2; pass++) {
		for (int i = 0; i < 208; i++) {
			uint32_t m0[3], m1[3], tot = 0;
#pragma unroll
			for (int j = 0; j < 3; j++) {
				const int is = gather_sbit(s.row, g[i * 3 + j]);
				const int d0 = is - 127, d1 = is + 127;
				m0[j] = is ? (uint32_t)((d0 * d0) >> 9) : 0u;
				m1[j] = is ? (uint32_t)((d1 * d1) >> 9) : 0u;
				tot += m0[j] + m1[j];
			}
			const uint4 lo = *reinterpret_cast<const uint4 *>(&s.ae[cur][4 * lane]);
			const uint4 hi = *reinterpret_cast<const uint4 *>(&s.ae[cur][128 + 4 * lane]);
			const uint32_t lov[4] = {lo.x, lo.y, lo.z, lo.w}, hiv[4] = {hi.x, hi.y, hi.z, hi.w};
			uint32_t nv[8];
			unsigned d = 0;
#pragma unroll
			for (int b = 0; b < 4; b++) {
				const uint32_t A = ((x[b] & 4) ? m1[0] : m0[0]) + ((x[b] & 2) ? m1[1] : m0[1]) +
				                   ((x[b] & 1) ? m1[2] : m0[2]);
				const uint32_t B = tot - A;
				const uint32_t a0 = lov[b] + A, b0 = hiv[b] + B;   // -> state 8L+2b
				const uint32_t a1 = lov[b] + B, b1 = hiv[b] + A;   // -> state 8L+2b+1
				const bool d0 = b0 < a0, d1 = b1 < a1;
				nv[2 * b] = d0 ? b0 : a0;
				nv[2 * b + 1] = d1 ? b1 : a1;
				d |= (d0 ? 1u : 0u) << (2 * b);
				d |= (d1 ? 1u : 0u) << (2 * b + 1);
			}
			uint4 *dst = reinterpret_cast<uint4 *>(&s.ae[cur ^ 1][8 * lane]);
			dst[0] = make_uint4(nv[0], nv[1], nv[2], nv[3]);
			dst[1] = make_uint4(nv[4], nv[5], nv[6], nv[7]);
			if (pass)
				s.dec[i][lane] = (uint8_t)d;
			cur ^= 1;
			__syncwarp();
		}
		if (pass == 0) {   // rewind: subtract the minimum (osmo_conv_decode_rewind)
			uint32_t mn = MAX_AE;
			for (int i = 0; i < 8; i++)
				mn = min(mn, s.ae[cur][8 * lane + i]);
			for (int o = 16; o; o >>= 1)
				mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
			for (int i = 0; i < 8; i++)
				s.ae[cur][8 * lane + i] -= mn;
			__syncwarp();
		}
	}

	// end state = first state with the minimal metric: min over (metric << 8 | state)
	uint32_t key = 0xffffffffu;
	for (int i = 0; i < 8; i++) {
		const uint32_t v = s.ae[cur][8 * lane + i];
		if (v < MAX_AE)
			key = min(key, (v << 8) | (uint32_t)(8 * lane + i));
	}
	for (int o = 16; o; o >>= 1)
		key = min(key, __shfl_xor_sync(0xffffffffu, key, o));

	if (lane == 0) {
		if (a.conv)
			a.conv[unit] = key == 0xffffffffu ? -1 : (int32_t)(key >> 8);
		for (int i = 0; i < 32; i++)
			s.out[i] = 0;
		unsigned st = key & 0xff;
		if (key != 0xffffffffu)
			for (int i = 207; i >= 0; i--) {
				const unsigned bit = (s.dec[i][st >> 3] >> (st & 7)) & 1u;
				s.out[i >> 3] |= (uint8_t)((st & 1u) << (i & 7));
				st = (st >> 1) | (bit << 7);
			}
		if (a.crc)
			a.crc[unit] = crc16_check_packed(s.out, 192);
		for (int i = 0; i < 24; i++)
			a.l2[(size_t)unit * 24 + i] = s.out[i];
	}
}

static cudaError_t launch_dc12(const DecodeArgs &a, cudaStream_t st)
{
	GMR1_INIT_LOCK();
	static bool attr_done[64] = {false};
	int dev = 0;
	cudaGetDevice(&dev);
	const int smem = (int)sizeof(Dc12Smem) * DC12_WARPS;
	if (dev >= 64 || !attr_done[dev]) {
		cudaError_t e = cudaFuncSetAttribute(decode_dc12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if (e != cudaSuccess)
			return e;
		if (dev < 64)
			attr_done[dev] = true;
	}
	decode_dc12_kernel<<<(a.n + DC12_WARPS - 1) / DC12_WARPS, DC12_WARPS * 32, smem, st>>>(a);
	return cudaGetLastError();
}

size_t decode_scratch_bytes(int ch, int n)
{
	if (n <= 0 || ch == CH_DC12 || ch < 0 || ch >= CH_COUNT)
		return 0;
	const size_t slots = (size_t)((n + TPC_T - 1) / TPC_T) * TPC_T;
	switch (ch) {
	case CH_TCH3: return slots * chan_dec_bytes(CH_TCH3);
	default:      return slots * (size_t)(chan_n_steps(ch) * 2);
	}
}

// ---- dispatch ------------------------------------------------------------------------------------
cudaError_t launch_decode(int ch, const DecodeArgs &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	cudaError_t e = upload_tables();
	if (e != cudaSuccess)
		return e;
	switch (ch) {
	case CH_BCCH:     return launch_tpc<CH_BCCH>(a, st);
	case CH_CCCH:     return launch_tpc<CH_CCCH>(a, st);
	case CH_FACCH3:   return launch_tpc<CH_FACCH3>(a, st);
	case CH_FACCH9:   return launch_tpc<CH_FACCH9>(a, st);
	case CH_TCH9_2K4: return launch_tpc<CH_TCH9_2K4>(a, st);
	case CH_TCH9_4K8: return launch_tpc<CH_TCH9_4K8>(a, st);
	case CH_TCH9_9K6: return launch_tpc<CH_TCH9_9K6>(a, st);
	case CH_RACH:     return launch_tpc<CH_RACH>(a, st);
	case CH_TCH3:     return launch_tpc<CH_TCH3>(a, st);
	case CH_DC12:     return launch_dc12(a, st);
	default:          return cudaErrorInvalidValue;
	}
}

}  // namespace gmr1
