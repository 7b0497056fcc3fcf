// rx_sched.cu - the BCCH/CCCH frame loop of the reference receiver for N channels in lock step
// (SURVEY 8f N1).  Replaces, for a whole batch of channels per call,
//   process_bcch      src/gmr1_rx.c:853-895   frame walk, which burst a frame carries
//   rx_bcch           :747-803                BCCH burst: demod, decode, alignment / frequency tracking
//   rx_ccch           :805-851                CCCH burst behind an energy gate; an IMM.ASS on it fills the
//   rx_tch3_init / ccch_imm_ass_parse  :236-245,362-381   channel's TCH3 hand-off record (the TCH3 burst loop
//                                             itself, rx_tch3 :538-600, stays with the caller)
//   bcch_tdma_align   :194-236                SI1 / segment 2Abis -> frame number, SA_SIRFN_DELAY, SA_BCCH_STN
//   burst_map / burst_energy  :149-182        window geometry, mean energy of the inner 30/32 of a window
//
// A channel's frames depend on each other (its alignment, frequency error, frame number and energy gate
// are updated by every good BCCH burst), channels do not: so the loop runs over frames on the host and
// over channels on the device.  All channel state lives in device memory and the whole walk is enqueued
// on one stream without a host round trip.  Per frame:
//   rx_prep_kernel    (one warp per channel)  what the frame carries, window position, energy, gate
//   rx_compact_kernel (one CTA)               ordered lists of the channels with a BCCH / a CCCH burst
//   demod_kernel x 2, decode_tpc_kernel x 2   on those lists (device-side counts: n_dev)
//   rx_update_kernel  (one thread per channel) results, tracking feedback, next frame
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "decode_unit.cuh"
#include "launch.h"

using namespace gmr1;

namespace {

constexpr int SYM_PER_SLOT = 39, SLOTS_PER_FRAME = 24;
constexpr int BURST_SYMS = 234;              // BCCH and DC6 bursts are 6 slots
constexpr int KIND_NONE = 0, KIND_BCCH = 1, KIND_CCCH = 2;

struct RxState {                             // SoA, [n] each, device memory
	const int64_t *rec_ofs;
	const int32_t *rec_len;
	int32_t *align;
	float   *freq_err;
	int32_t *fn, *delay, *stn;
	float   *bcch_energy;                    // energy of the last BCCH window (NaN before the first)
	int32_t *done;
	// per frame
	int32_t *kind;                           // what this frame carries for the channel
	int32_t *begin;                          // window start within the recording
	float   *energy;                         // energy of that window
	int32_t *slot;                           // index in the BCCH / CCCH list
};

struct RxLists {                             // device memory
	int32_t *count;                          // [2] entries in the BCCH / CCCH list
	int64_t *ofs[2];                         // [n] absolute window start within iq
	float   *fs[2];                          // [n] freq_shift = -freq_err
};

struct RxOut {                               // [n][max_frames] device memory (l2: [n][max_frames][24])
	int32_t *kind, *fn, *crc, *conv;
	uint8_t *l2;
	int32_t *n_frames;                       // [n]
	int32_t *tch3;                           // [n][4] active, tn, p, frame of the assignment; or NULL
	float   *tch3_energy;                    // [n][2] energy_burst, energy_dkab; or NULL
};

// mean energy of the inner 30/32 of a window (burst_energy, gmr1_rx.c:172-182); warp-parallel partial sums
__device__ float window_energy(const float2 *__restrict__ x, int len, int lane)
{
	const int b = len >> 5;
	float e = 0.0f;
	for (int i = b + lane; i < len - b; i += 32) {
		const float2 v = __ldg(&x[i]);
		e += v.x * v.x + v.y * v.y;
	}
#pragma unroll
	for (int o = 16; o; o >>= 1)
		e += __shfl_xor_sync(0xffffffffu, e, o);
	return e / (float)len;
}

__global__ void __launch_bounds__(128) rx_prep_kernel(const float2 *__restrict__ iq, RxState st, int n, int sps)
{
	const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (i >= n)
		return;
	int kind = KIND_NONE, begin = 0;
	float energy = 0.0f;
	if (!st.done[i]) {
		const int sirfn = (st.fn[i] - st.delay[i]) & 63, m = sirfn & 7;       // process_bcch :870-878
		const int win = m == 2 ? 20 * sps : 10 * sps;                         // rx_bcch :760, rx_ccch :815
		if (m != 0) {
			const int etoa = win >> 1;                                        // burst_map :158-165
			begin = st.align[i] + sps * st.stn[i] * SYM_PER_SLOT - etoa;
			const int len = BURST_SYMS * sps + win;
			if (begin >= 0 && begin + len <= st.rec_len[i]) {
				kind = m == 2 ? KIND_BCCH : KIND_CCCH;
				energy = window_energy(iq + st.rec_ofs[i] + begin, len, lane);
				// energy gate of the CCCH (:819-820); a NaN threshold (no BCCH seen yet) lets everything pass
				if (kind == KIND_CCCH && energy < st.bcch_energy[i] / 2.0f)
					kind = KIND_NONE;
			}
		}
	}
	if (lane == 0) {
		st.kind[i] = kind;
		st.begin[i] = begin;
		st.energy[i] = energy;
	}
}

// ordered compaction of the channels that carry a BCCH / a CCCH burst in this frame (one CTA)
__global__ void __launch_bounds__(1024) rx_compact_kernel(RxState st, RxLists ls, int n)
{
	__shared__ int wtot[2][32], wbase[2][32], base[2];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < 2)
		base[tid] = 0;
	__syncthreads();
	for (int i0 = 0; i0 < n; i0 += 1024) {
		const int i = i0 + tid;
		const int kind = i < n ? st.kind[i] : KIND_NONE;
		int pos[2];
#pragma unroll
		for (int k = 0; k < 2; k++) {
			const unsigned m = __ballot_sync(0xffffffffu, kind == k + 1);
			pos[k] = __popc(m & ((1u << lane) - 1u));
			if (lane == 0)
				wtot[k][warp] = __popc(m);
		}
		__syncthreads();
		if (warp == 0) {
#pragma unroll
			for (int k = 0; k < 2; k++) {
				const int v = wtot[k][lane];
				int s = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int t = __shfl_up_sync(0xffffffffu, s, o);
					if (lane >= o)
						s += t;
				}
				wbase[k][lane] = base[k] + s - v;        // first list position of each warp's channels
			}
		}
		__syncthreads();
		if (kind != KIND_NONE) {
			const int k = kind - 1, p = wbase[k][warp] + pos[k];
			st.slot[i] = p;
			ls.ofs[k][p] = st.rec_ofs[i] + st.begin[i];
			ls.fs[k][p] = -st.freq_err[i];                // rx_bcch :765, rx_ccch :829
		}
		__syncthreads();
		if (tid < 2)
			base[tid] = wbase[tid][31] + wtot[tid][31];
		__syncthreads();
	}
	if (tid < 2)
		ls.count[tid] = base[tid];
}

struct RxBurstOut {                          // outputs of the two demod + decode chains, [n] per list
	const float   *toa[2], *ferr[2];
	const int32_t *crc[2], *conv[2];
	const uint8_t *l2[2];
};

__global__ void __launch_bounds__(128) rx_update_kernel(RxState st, RxBurstOut bo, RxOut out, int n, int sps,
                                                        int frame, int max_frames)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || st.done[i])
		return;
	const int kind = st.kind[i];
	const size_t rec = (size_t)i * max_frames + frame;
	out.kind[rec] = kind;
	out.fn[rec] = st.fn[i];
	int crc = -1, conv = 0;
	if (kind != KIND_NONE) {
		const int k = kind - 1, p = st.slot[i];
		crc = bo.crc[k][p];
		conv = bo.conv[k][p];
		const uint8_t *l2 = bo.l2[k] + (size_t)p * 24;
		uint8_t *o = out.l2 + rec * 24;
		for (int b = 0; b < 24; b++)
			o[b] = l2[b];
		if (kind == KIND_BCCH) {
			st.bcch_energy[i] = st.energy[i];                                 // rx_bcch :773-774
			if (crc == 0) {
				const int etoa = (20 * sps) >> 1;
				int align = st.align[i] + ((int)roundf(bo.toa[k][p]) - etoa);     // :784
				st.freq_err[i] += bo.ferr[k][p];                                  // :785
				// bcch_tdma_align :194-236: SI1 carrying segment 2Abis
				if ((l2[0] & 0xf8) == 0x08 && (l2[9] & 0xfc) == 0x80) {
					const int delay = (l2[10] >> 3) & 0x0f;
					const int stn = ((l2[10] << 2) & 0x1c) | (l2[11] >> 6);
					const int sf = ((l2[11] & 0x3f) << 7) | (l2[12] >> 1);
					const int mf = ((l2[12] & 0x01) << 1) | (l2[13] >> 7);
					const int hi = (l2[13] & 0x40) >> 6;
					align += (st.stn[i] - stn) * SYM_PER_SLOT * sps;
					st.fn[i] = (sf << 6) | (mf << 4) | (hi << 3) | ((2 + delay) & 7);
					st.delay[i] = delay;
					st.stn[i] = stn;
				}
				st.align[i] = align;
			}
		} else if (crc == 0 && out.tch3 && l2[1] == 0x06 && l2[2] == 0x3f) {      // ccch_is_imm_ass :236-239
			// rx_tch3_init(cd, l2, min_energy) with min_energy = bcch_energy / 2 (:836-838, :878); a later
			// IMM.ASS re-initialises the state as in the reference
			const float ref = st.bcch_energy[i] / 2.0f;
			const float eb = ref * 0.75f;
			out.tch3[4 * i + 0] = 1;
			out.tch3[4 * i + 1] = ((l2[8] & 0x03) << 3) | (l2[9] >> 5);            // ccch_imm_ass_parse :240-245
			out.tch3[4 * i + 2] = (l2[8] & 0xfc) >> 2;
			out.tch3[4 * i + 3] = frame;
			if (out.tch3_energy) {
				out.tch3_energy[2 * i + 0] = eb;                                   // :372-373
				out.tch3_energy[2 * i + 1] = eb / 8.0f;
			}
		}
	}
	out.crc[rec] = crc;
	out.conv[rec] = conv;
	// next frame (process_bcch :884-891)
	const int frame_len = sps * SLOTS_PER_FRAME * SYM_PER_SLOT;
	st.fn[i] += 1;
	st.align[i] += frame_len;
	out.n_frames[i] = frame + 1;
	if (st.align[i] + 2 * frame_len > st.rec_len[i])
		st.done[i] = 1;
}

__global__ void rx_init_kernel(RxState st, const int32_t *align0, const float *freq_err0, int32_t *n_frames,
                               int32_t *tch3, float *tch3_energy, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	if (tch3) {
		tch3[4 * i + 0] = 0; tch3[4 * i + 1] = 0; tch3[4 * i + 2] = 0; tch3[4 * i + 3] = -1;
	}
	if (tch3_energy) {
		tch3_energy[2 * i + 0] = 0.0f; tch3_energy[2 * i + 1] = 0.0f;
	}
	st.align[i] = align0[i];
	st.freq_err[i] = freq_err0 ? freq_err0[i] : 0.0f;
	st.fn[i] = 0;                                  // chan_desc is zeroed in main(), gmr1_rx.c:906
	st.delay[i] = 0;
	st.stn[i] = 0;
	st.bcch_energy[i] = __int_as_float(0x7fc00000);    // nan("inf"), :859
	st.done[i] = 0;
	n_frames[i] = 0;
}

__global__ void rx_final_kernel(RxState st, int32_t *align, float *freq_err, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	if (align)
		align[i] = st.align[i];
	if (freq_err)
		freq_err[i] = st.freq_err[i];
}

}  // namespace

static int rx_bcch_walk(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                        const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                        int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                        int32_t *n_frames, int32_t *align_out, float *freq_err_out,
                        int32_t *tch3, float *tch3_energy, void *stream)
{
	if (!iq || !rec_ofs || !rec_len || !align0 || n < 0 || max_frames < 1 || sps < 1 || sps > 16 || !kind || !fn ||
	    !crc || !conv || !l2 || !n_frames)
		return set_err(-EINVAL, "rx_bcch_batch: bad argument");
	if (n == 0)
		return 0;
	cudaStream_t cs = (cudaStream_t)stream;
	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");

	Stage s(stream);
	const size_t N = (size_t)n, NF = N * (size_t)max_frames;
	const float2 *d_iq = (const float2 *)s.in(iq, (size_t)iq_len * 2);
	RxState st = {};
	st.rec_ofs = s.in(rec_ofs, N);
	st.rec_len = s.in(rec_len, N);
	const int32_t *d_align0 = s.in(align0, N);
	const float *d_ferr0 = s.in(freq_err0, N);
	RxOut out = {};
	out.kind = s.out(kind, NF); out.fn = s.out(fn, NF); out.crc = s.out(crc, NF); out.conv = s.out(conv, NF);
	out.l2 = s.out(l2, NF * 24); out.n_frames = s.out(n_frames, N);
	out.tch3 = s.out(tch3, N * 4); out.tch3_energy = s.out(tch3_energy, N * 2);
	int32_t *d_align_out = s.out(align_out, N);
	float *d_ferr_out = s.out(freq_err_out, N);

	st.align = s.tmp<int32_t>(N); st.freq_err = s.tmp<float>(N);
	st.fn = s.tmp<int32_t>(N); st.delay = s.tmp<int32_t>(N); st.stn = s.tmp<int32_t>(N);
	st.bcch_energy = s.tmp<float>(N); st.done = s.tmp<int32_t>(N);
	st.kind = s.tmp<int32_t>(N); st.begin = s.tmp<int32_t>(N); st.energy = s.tmp<float>(N); st.slot = s.tmp<int32_t>(N);
	RxLists ls = {};
	ls.count = s.tmp<int32_t>(2);
	// per list: window offsets, frequency shifts, soft bits, demod and decode outputs
	int8_t *eb[2]; float *toa[2], *ferr[2]; int32_t *dcrc[2], *dconv[2]; uint8_t *dl2[2], *dscr[2];
	const int ebits[2] = {424, 432}, bt[2] = {BT_BCCH, BT_DC6}, ch[2] = {CH_BCCH, CH_CCCH};
	for (int k = 0; k < 2; k++) {
		ls.ofs[k] = s.tmp<int64_t>(N); ls.fs[k] = s.tmp<float>(N);
		eb[k] = s.tmp<int8_t>(N * ebits[k]); toa[k] = s.tmp<float>(N); ferr[k] = s.tmp<float>(N);
		dcrc[k] = s.tmp<int32_t>(N); dconv[k] = s.tmp<int32_t>(N); dl2[k] = s.tmp<uint8_t>(N * 24);
		dscr[k] = s.tmp<uint8_t>(decode_scratch_bytes(ch[k], n));
	}
	if (s.failed())
		return s.finish(cudaSuccess, "rx_bcch_batch: staging");

	// every recording must lie inside iq (checked on the host copy of the descriptors when they are host memory
	// is not possible in general: the kernels bound every window by rec_len, the caller vouches for rec_ofs)
	const int tb = 128, grid = (n + tb - 1) / tb;
	rx_init_kernel<<<grid, tb, 0, cs>>>(st, d_align0, d_ferr0, out.n_frames, out.tch3, out.tch3_energy, n);
	cudaMemsetAsync(out.kind, 0, NF * sizeof(int32_t), cs);
	cudaMemsetAsync(out.crc, 0xff, NF * sizeof(int32_t), cs);
	uint64_t launches = 1;
	RxBurstOut bo = {};
	for (int k = 0; k < 2; k++) {
		bo.toa[k] = toa[k]; bo.ferr[k] = ferr[k]; bo.crc[k] = dcrc[k]; bo.conv[k] = dconv[k]; bo.l2[k] = dl2[k];
	}
	e = cudaGetLastError();
	for (int f = 0; f < max_frames && e == cudaSuccess; f++) {
		rx_prep_kernel<<<(n + 3) / 4, 128, 0, cs>>>(d_iq, st, n, sps);
		rx_compact_kernel<<<1, 1024, 0, cs>>>(st, ls, n);
		launches += 2;
		for (int k = 0; k < 2 && e == cudaSuccess; k++) {
			DemodArgs a = {};
			a.iq = d_iq; a.ofs = ls.ofs[k]; a.n = n; a.sps = sps;
			a.win_len = BURST_SYMS * sps + (k == 0 ? 20 : 10) * sps;
			a.freq_shift = ls.fs[k];
			a.e_toa0 = -1.0f;
			a.ebits = eb[k]; a.ebits_stride = ebits[k];
			a.toa = toa[k]; a.freq_err = ferr[k];
			a.n_dev = ls.count + k;
			e = launch_demod(a, d_all + bt[k], &burst_tab(bt[k]), 1, 0, cs);
			if (e != cudaSuccess)
				break;
			DecodeArgs d = {};
			d.ebits = eb[k]; d.n = n; d.l2 = dl2[k]; d.conv = dconv[k]; d.crc = dcrc[k];
			d.n_dev = ls.count + k;
			d.dec_scratch = dscr[k];
			e = launch_decode(ch[k], d, cs);
			launches += 2;
		}
		if (e != cudaSuccess)
			break;
		rx_update_kernel<<<grid, tb, 0, cs>>>(st, bo, out, n, sps, f, max_frames);
		launches += 1;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess) {
		rx_final_kernel<<<grid, tb, 0, cs>>>(st, d_align_out, d_ferr_out, n);
		e = cudaGetLastError();
	}
	g_launches.fetch_add(launches);
	return s.finish(e, "rx_bcch_batch kernels");
}

extern "C" int gmr1b200_rx_bcch_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                                      const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                                      int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                                      int32_t *n_frames, int32_t *align_out, float *freq_err_out, void *stream)
{
	return rx_bcch_walk(iq, iq_len, rec_ofs, rec_len, align0, freq_err0, sps, n, max_frames, kind, fn, crc, conv, l2,
	                    n_frames, align_out, freq_err_out, nullptr, nullptr, stream);
}

extern "C" int gmr1b200_rx_bcch_ass_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs,
                                          const int32_t *rec_len, const int32_t *align0, const float *freq_err0, int sps,
                                          int n, int max_frames, int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv,
                                          uint8_t *l2, int32_t *n_frames, int32_t *align_out, float *freq_err_out,
                                          int32_t *tch3, float *tch3_energy, void *stream)
{
	if (!tch3)
		return set_err(-EINVAL, "rx_bcch_ass_batch: tch3 NULL");
	return rx_bcch_walk(iq, iq_len, rec_ofs, rec_len, align0, freq_err0, sps, n, max_frames, kind, fn, crc, conv, l2,
	                    n_frames, align_out, freq_err_out, tch3, tch3_energy, stream);
}

// ---- fused burst -> L2 for the common control channels ------------------------------------------------
extern "C" int gmr1b200_rx_xcch_batch(int chan, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                      int64_t win_stride, int win_len, int sps, const float *freq_shift,
                                      float freq_shift0, uint8_t *l2, int32_t *crc, int32_t *conv, float *toa,
                                      float *freq_err, int n, void *stream)
{
	if (chan < 0 || chan > 2 || !iq || !l2 || n < 0 || sps < 1 || sps > 16)
		return set_err(-EINVAL, "rx_xcch_batch: bad argument");
	const int bt = chan == 0 ? BT_BCCH : chan == 1 ? BT_DC6 : BT_DC12;
	const int ch = chan == 0 ? CH_BCCH : chan == 1 ? CH_CCCH : CH_DC12;
	const BurstTab &t = burst_tab(bt);
	if (win_len < t.len * sps)
		return set_err(-EINVAL, "rx_xcch_batch: window shorter than the burst");
	if (n == 0)
		return 0;
	if (!win_ofs && (win_stride < 0 || (int64_t)(n - 1) * win_stride + win_len > iq_len))
		return set_err(-EINVAL, "rx_xcch_batch: windows exceed iq_len");
	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");
	cudaStream_t cs = (cudaStream_t)stream;
	Stage s(stream);
	const size_t N = (size_t)n;
	DemodArgs a = {};
	a.iq = (const float2 *)s.in(iq, (size_t)iq_len * 2);
	a.ofs = s.in(win_ofs, N); a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = s.in(freq_shift, N); a.freq_shift0 = freq_shift0; a.e_toa0 = -1.0f;
	a.ebits = s.tmp<int8_t>(N * t.ebits); a.ebits_stride = t.ebits;      // soft bits never leave the device
	a.toa = s.out(toa, N); a.freq_err = s.out(freq_err, N);
	DecodeArgs d = {};
	d.ebits = a.ebits; d.n = n;
	d.l2 = s.out(l2, N * 24); d.crc = s.out(crc, N); d.conv = s.out(conv, N);
	if (const size_t sb = decode_scratch_bytes(ch, n))
		d.dec_scratch = s.tmp<uint8_t>(sb);
	if (!s.failed()) {
		e = launch_demod(a, d_all + bt, &t, 1, 0, cs);
		if (e == cudaSuccess) {
			g_launches.fetch_add(1);
			e = launch_decode(ch, d, cs);
			if (e == cudaSuccess)
				g_launches.fetch_add(1);
		}
	}
	return s.finish(e, "rx_xcch_batch kernels");
}

