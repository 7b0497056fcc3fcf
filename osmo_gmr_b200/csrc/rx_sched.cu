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
#include "rx_common.cuh"


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
	if (!recordings_in_range(rec_ofs, rec_len, n, iq_len))
		return set_err(-EINVAL, "rx_bcch_batch: a recording lies outside iq_len");
	WalkStream ws(stream);
	cudaStream_t cs = ws.get();
	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");

	Stage s((void *)cs);
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

	const int tb = 128, grid = (n + tb - 1) / tb;
	rx_init_kernel<<<grid, tb, 0, cs>>>(st, d_align0, d_ferr0, out.n_frames, out.tch3, out.tch3_energy, n);
	cudaMemsetAsync(out.kind, 0, NF * sizeof(int32_t), cs);
	cudaMemsetAsync(out.crc, 0xff, NF * sizeof(int32_t), cs);
	uint64_t launches = 1;
	RxBurstOut bo = {};
	for (int k = 0; k < 2; k++) {
		bo.toa[k] = toa[k]; bo.ferr[k] = ferr[k]; bo.crc[k] = dcrc[k]; bo.conv[k] = dconv[k]; bo.l2[k] = dl2[k];
	}
	int *d_frame = s.tmp<int>(1);
	if (s.failed())
		return s.finish(cudaSuccess, "rx_bcch_batch: staging");
	cudaMemsetAsync(d_frame, 0, sizeof(int), cs);
	e = cudaGetLastError();
	// the launches of one frame (identical for every frame: the frame index lives in d_frame)
	auto frame = [&]() -> cudaError_t {
		cudaError_t fe = cudaSuccess;
		rx_prep_kernel<<<(n + 3) / 4, 128, 0, cs>>>(d_iq, st, n, sps);
		rx_compact_kernel<<<1, 1024, 0, cs>>>(st, ls, n);
		for (int k = 0; k < 2 && fe == cudaSuccess; k++) {
			DemodArgs a = {};
			a.iq = d_iq; a.ofs = ls.ofs[k]; a.n = n; a.sps = sps;
			a.win_len = BURST_SYMS * sps + (k == 0 ? 20 : 10) * sps;
			a.freq_shift = ls.fs[k];
			a.e_toa0 = -1.0f;
			a.ebits = eb[k]; a.ebits_stride = ebits[k];
			a.toa = toa[k]; a.freq_err = ferr[k];
			a.n_dev = ls.count + k;
			fe = launch_demod(a, d_all + bt[k], &burst_tab(bt[k]), 1, 0, cs);
			if (fe != cudaSuccess)
				break;
			DecodeArgs d = {};
			d.ebits = eb[k]; d.n = n; d.l2 = dl2[k]; d.conv = dconv[k]; d.crc = dcrc[k];
			d.n_dev = ls.count + k;
			d.dec_scratch = dscr[k];
			fe = launch_decode(ch[k], d, cs);
		}
		if (fe != cudaSuccess)
			return fe;
		rx_update_kernel<<<grid, tb, 0, cs>>>(st, bo, out, n, sps, d_frame, max_frames, nullptr, nullptr, true);
		rx_tick_kernel<<<1, 1, 0, cs>>>(d_frame);
		return cudaGetLastError();
	};
	const int per_frame = 8;
	FrameGraph fg(cs);
	bool graph = false;
	for (int f = 0; f < max_frames && e == cudaSuccess; f++) {
		if (f == 1 && max_frames >= 4 && ws.capturable()) {
			if (fg.begin() == cudaSuccess) {
				const cudaError_t ce = frame();
				if (ce == cudaSuccess && fg.end() == cudaSuccess)
					graph = true;
				else
					fg.abort();
			} else
				cudaGetLastError();
		}
		e = graph ? fg.launch() : frame();
		launches += per_frame;
	}
	if (e == cudaSuccess) {
		rx_final_kernel<<<grid, tb, 0, cs>>>(st, d_align_out, d_ferr_out, n);
		e = cudaGetLastError();
	}
	g_launches.fetch_add(launches);
	const int rc = s.finish(e, "rx_bcch_batch kernels");
	ws.join();
	return rc;
}

// ---- the same walk, paced by each channel's own BCCH bursts --------------------------------------------------------
// The only frames of a channel that depend on each other are its BCCH frames: a good BCCH burst updates alignment,
// frequency error, frame numbering (SI1) and the energy gate; a CCCH frame reads that state and changes none of it
// (rx_ccch, src/gmr1_rx.c:805-851: its only side effect is the TCH3 hand-off record).  So instead of 64 dependent
// rounds of small kernels (one per frame) the walk alternates
//   B step   every channel whose next frame is a BCCH frame: window, demod, decode, tracking update, one frame on
//   C step   every channel: ALL following frames up to (not including) its next BCCH frame - at most 7 - as one batch
// i.e. two dependent rounds per eight frames.  Channels need not be frame-aligned with each other: every channel
// walks its own frame counter.  The results are those of the frame-by-frame walk, record for record.
namespace {

constexpr int RUN = 7;                        // frames of one C step per channel

struct Pace {                                 // device memory
	int32_t *frame;                           // [n] frames walked so far (index of the next record)
	int32_t *run_base, *run_len;              // [n] first record / number of frames of the current C step
	int32_t *vkey, *vslot;                    // [RUN n] list key / position of frame k of channel i at [i * RUN + k]
	int64_t *vofs;                            // [RUN n] absolute window start
	float   *vfe;                             // [RUN n] the channel's frequency error
	int32_t *bkey, *bslot;                    // [n] B step
	int64_t *bofs;
	int32_t *all_done;                        // [1] set by the B step when no channel has frames left
};

__device__ __forceinline__ bool pace_done(const RxState &st, const Pace &pc, int i, int max_frames)
{
	return st.done[i] || pc.frame[i] >= max_frames;
}

// C step, planning: one warp per channel walks its next frames up to the next BCCH frame (process_bcch :866-891)
__global__ void __launch_bounds__(128) pace_run_kernel(const float2 *__restrict__ iq, RxState st, Pace pc, RxOut out, int n,
                                                       int sps, int max_frames)
{
	const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (i >= n)
		return;
	const int frame_len = sps * SLOTS_PER_FRAME * SYM_PER_SLOT;
	int fn = st.fn[i], align = st.align[i], frame = pc.frame[i], done = st.done[i], k = 0;
	const int base = frame;
	for (; k < RUN && !done && frame < max_frames; k++) {
		const int m = (fn - st.delay[i]) & 7;
		if (m == 2)
			break;                                                            // the next frame is a BCCH frame
		int key = 0;
		int64_t wofs = 0;
		if (m != 0) {
			const int win = 10 * sps, etoa = win >> 1;                        // rx_ccch :815, burst_map :158-165
			const int begin = align + sps * st.stn[i] * SYM_PER_SLOT - etoa;
			const int len = BURST_SYMS * sps + win;
			if (begin >= 0 && begin + len <= st.rec_len[i]) {
				const float energy = window_energy(iq + st.rec_ofs[i] + begin, len, lane);
				// energy gate of the CCCH (:819-820); a NaN threshold (no BCCH seen yet) lets everything pass
				if (!(energy < st.bcch_energy[i] / 2.0f)) {
					key = 1;
					wofs = st.rec_ofs[i] + begin;
				}
			}
		}
		if (lane == 0) {
			pc.vkey[i * RUN + k] = key;
			pc.vofs[i * RUN + k] = wofs;
			pc.vfe[i * RUN + k] = st.freq_err[i];
			out.fn[(size_t)i * max_frames + frame] = fn;
		}
		fn++;
		align += frame_len;
		frame++;
		if (align + 2 * frame_len > st.rec_len[i])
			done = 1;
	}
	if (lane == 0) {
		for (int r = k; r < RUN; r++)
			pc.vkey[i * RUN + r] = 0;
		pc.run_base[i] = base;
		pc.run_len[i] = k;
		st.fn[i] = fn;
		st.align[i] = align;
		st.done[i] = done;
		pc.frame[i] = frame;
		out.n_frames[i] = frame;
	}
}

// C step, results: records of the run in frame order, IMM.ASS hand-off (rx_ccch :830-848)
__global__ void __launch_bounds__(128) pace_run_result_kernel(RxState st, Pace pc, RxBurstOut bo, RxOut out, int n,
                                                              int max_frames)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	for (int k = 0; k < pc.run_len[i]; k++) {
		if (!pc.vkey[i * RUN + k])
			continue;
		const int p = pc.vslot[i * RUN + k], frame = pc.run_base[i] + k;
		const size_t rec = (size_t)i * max_frames + frame;
		const uint8_t *l2 = bo.l2[1] + (size_t)p * 24;
		const int crc = bo.crc[1][p];
		out.kind[rec] = KIND_CCCH;
		out.crc[rec] = crc;
		out.conv[rec] = bo.conv[1][p];
		for (int b = 0; b < 24; b++)
			out.l2[rec * 24 + b] = l2[b];
		if (crc == 0 && out.tch3 && l2[1] == 0x06 && l2[2] == 0x3f) {          // ccch_is_imm_ass :236-239
			const float eb = st.bcch_energy[i] / 2.0f * 0.75f;                // rx_tch3_init :372-373
			out.tch3[4 * i + 0] = 1;
			out.tch3[4 * i + 1] = ((l2[8] & 0x03) << 3) | (l2[9] >> 5);
			out.tch3[4 * i + 2] = (l2[8] & 0xfc) >> 2;
			out.tch3[4 * i + 3] = frame;
			if (out.tch3_energy) {
				out.tch3_energy[2 * i + 0] = eb;
				out.tch3_energy[2 * i + 1] = eb / 8.0f;
			}
		}
	}
}

// B step, planning: the BCCH window of every channel that stands on a BCCH frame (rx_bcch :747-770)
__global__ void __launch_bounds__(128) pace_bcch_prep_kernel(const float2 *__restrict__ iq, RxState st, Pace pc, int n, int sps,
                                                             int max_frames)
{
	const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (i >= n)
		return;
	int key = 0;
	if (!pace_done(st, pc, i, max_frames) && ((st.fn[i] - st.delay[i]) & 7) == 2) {
		key = 2;                                                              // a BCCH frame without a mappable window
		const int win = 20 * sps, etoa = win >> 1;
		const int begin = st.align[i] + sps * st.stn[i] * SYM_PER_SLOT - etoa;
		const int len = BURST_SYMS * sps + win;
		if (begin >= 0 && begin + len <= st.rec_len[i]) {
			const float energy = window_energy(iq + st.rec_ofs[i] + begin, len, lane);
			if (lane == 0) {
				st.energy[i] = energy;
				pc.bofs[i] = st.rec_ofs[i] + begin;
			}
			key = 1;
		}
	}
	if (lane == 0)
		pc.bkey[i] = key;
}

// B step, results: rx_bcch :771-803 + bcch_tdma_align :194-236, then one frame on
__global__ void __launch_bounds__(128) pace_bcch_update_kernel(RxState st, Pace pc, RxBurstOut bo, RxOut out, int n, int sps,
                                                               int max_frames)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool left = false;
	if (i < n) {
		const int key = pc.bkey[i];
		if (key) {
			const int frame = pc.frame[i];
			const size_t rec = (size_t)i * max_frames + frame;
			out.fn[rec] = st.fn[i];
			if (key == 1) {
				const int p = pc.bslot[i];
				const int crc = bo.crc[0][p];
				const uint8_t *l2 = bo.l2[0] + (size_t)p * 24;
				out.kind[rec] = KIND_BCCH;
				out.crc[rec] = crc;
				out.conv[rec] = bo.conv[0][p];
				for (int b = 0; b < 24; b++)
					out.l2[rec * 24 + b] = l2[b];
				st.bcch_energy[i] = st.energy[i];                             // :773-774
				if (crc == 0) {
					const int etoa = (20 * sps) >> 1;
					int align = st.align[i] + ((int)roundf(bo.toa[0][p]) - etoa);     // :784
					st.freq_err[i] += bo.ferr[0][p];                                  // :785
					if ((l2[0] & 0xf8) == 0x08 && (l2[9] & 0xfc) == 0x80) {           // SI1 with segment 2Abis
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
			}
			const int frame_len = sps * SLOTS_PER_FRAME * SYM_PER_SLOT;       // process_bcch :884-891
			st.fn[i] += 1;
			st.align[i] += frame_len;
			pc.frame[i] = frame + 1;
			out.n_frames[i] = frame + 1;
			if (st.align[i] + 2 * frame_len > st.rec_len[i])
				st.done[i] = 1;
		}
		left = !pace_done(st, pc, i, max_frames);
	}
	if (__syncthreads_or(left) && threadIdx.x == 0)
		atomicExch(pc.all_done, 0);
}

__global__ void pace_init_kernel(Pace pc, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		pc.frame[i] = 0;
}

}  // namespace

static int rx_bcch_walk_paced(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
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
	if (!recordings_in_range(rec_ofs, rec_len, n, iq_len))
		return set_err(-EINVAL, "rx_bcch_batch: a recording lies outside iq_len");
	cudaStream_t cs = (cudaStream_t)stream;
	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");

	Stage s(stream);
	const size_t N = (size_t)n, NF = N * (size_t)max_frames, NV = N * RUN;
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
	Pace pc = {};
	pc.frame = s.tmp<int32_t>(N); pc.run_base = s.tmp<int32_t>(N); pc.run_len = s.tmp<int32_t>(N);
	pc.vkey = s.tmp<int32_t>(NV); pc.vslot = s.tmp<int32_t>(NV); pc.vofs = s.tmp<int64_t>(NV); pc.vfe = s.tmp<float>(NV);
	pc.bkey = s.tmp<int32_t>(N); pc.bslot = s.tmp<int32_t>(N); pc.bofs = s.tmp<int64_t>(N);
	pc.all_done = s.tmp<int32_t>(1);
	Lists<1> lb = {}, lc = {};                // B step: BCCH windows; C step: CCCH windows of up to RUN frames per channel
	lb.count = s.tmp<int32_t>(1); lb.idx = s.tmp<int32_t>(N); lb.ofs = s.tmp<int64_t>(N); lb.fs = s.tmp<float>(N);
	lc.count = s.tmp<int32_t>(1); lc.idx = s.tmp<int32_t>(NV); lc.ofs = s.tmp<int64_t>(NV); lc.fs = s.tmp<float>(NV);
	const int ebits[2] = {424, 432}, bt[2] = {BT_BCCH, BT_DC6}, ch[2] = {CH_BCCH, CH_CCCH};
	const size_t cap[2] = {N, NV};
	int8_t *eb[2]; float *toa[2], *ferr[2]; int32_t *dcrc[2], *dconv[2]; uint8_t *dl2[2], *dscr[2];
	for (int k = 0; k < 2; k++) {
		eb[k] = s.tmp<int8_t>(cap[k] * ebits[k]); toa[k] = s.tmp<float>(cap[k]); ferr[k] = s.tmp<float>(cap[k]);
		dcrc[k] = s.tmp<int32_t>(cap[k]); dconv[k] = s.tmp<int32_t>(cap[k]); dl2[k] = s.tmp<uint8_t>(cap[k] * 24);
		dscr[k] = s.tmp<uint8_t>(decode_scratch_bytes(ch[k], (int)cap[k]));
	}
	static thread_local int32_t *h_done = nullptr;    // page-locked word for the "every channel is through" check
	if (!h_done && cudaMallocHost(&h_done, sizeof(int32_t)) != cudaSuccess) {
		h_done = nullptr;
		return s.finish(cudaErrorMemoryAllocation, "rx_bcch_batch: pinned word");
	}
	if (s.failed())
		return s.finish(cudaSuccess, "rx_bcch_batch: staging");

	const int tb = 128, grid = (n + tb - 1) / tb, wgrid = (n + 3) / 4;
	rx_init_kernel<<<grid, tb, 0, cs>>>(st, d_align0, d_ferr0, out.n_frames, out.tch3, out.tch3_energy, n);
	pace_init_kernel<<<grid, tb, 0, cs>>>(pc, n);
	cudaMemsetAsync(out.kind, 0, NF * sizeof(int32_t), cs);
	cudaMemsetAsync(out.crc, 0xff, NF * sizeof(int32_t), cs);
	uint64_t launches = 2;
	RxBurstOut bo = {};
	for (int k = 0; k < 2; k++) {
		bo.toa[k] = toa[k]; bo.ferr[k] = ferr[k]; bo.crc[k] = dcrc[k]; bo.conv[k] = dconv[k]; bo.l2[k] = dl2[k];
	}
	auto chain = [&](int k, const Lists<1> &ls) -> cudaError_t {              // demod + decode of one list
		DemodArgs a = {};
		a.iq = d_iq; a.ofs = ls.ofs; a.n = (int)cap[k]; a.sps = sps;
		a.win_len = BURST_SYMS * sps + (k == 0 ? 20 : 10) * sps;
		a.freq_shift = ls.fs; a.e_toa0 = -1.0f;
		a.ebits = eb[k]; a.ebits_stride = ebits[k]; a.toa = toa[k]; a.freq_err = ferr[k];
		a.n_dev = ls.count;
		cudaError_t ce = launch_demod(a, d_all + bt[k], &burst_tab(bt[k]), 1, 0, cs);
		if (ce != cudaSuccess)
			return ce;
		DecodeArgs d = {};
		d.ebits = eb[k]; d.n = (int)cap[k]; d.l2 = dl2[k]; d.conv = dconv[k]; d.crc = dcrc[k];
		d.n_dev = ls.count; d.dec_scratch = dscr[k];
		launches += 2;
		return launch_decode(ch[k], d, cs);
	};
	auto c_step = [&]() -> cudaError_t {
		pace_run_kernel<<<wgrid, 128, 0, cs>>>(d_iq, st, pc, out, n, sps, max_frames);
		compact_kernel<1><<<1, 1024, 0, cs>>>(pc.vkey, pc.vofs, pc.vfe, nullptr, (int)NV, pc.vslot, lc);
		cudaError_t ce = chain(1, lc);
		pace_run_result_kernel<<<grid, tb, 0, cs>>>(st, pc, bo, out, n, max_frames);
		launches += 3;
		return ce != cudaSuccess ? ce : cudaGetLastError();
	};
	auto b_step = [&]() -> cudaError_t {
		pace_bcch_prep_kernel<<<wgrid, 128, 0, cs>>>(d_iq, st, pc, n, sps, max_frames);
		compact_kernel<1><<<1, 1024, 0, cs>>>(pc.bkey, pc.bofs, st.freq_err, nullptr, n, pc.bslot, lb);
		cudaError_t ce = chain(0, lb);
		cudaMemsetAsync(pc.all_done, 0xff, sizeof(int32_t), cs);
		pace_bcch_update_kernel<<<grid, tb, 0, cs>>>(st, pc, bo, out, n, sps, max_frames);
		launches += 3;
		return ce != cudaSuccess ? ce : cudaGetLastError();
	};
	// the B step's list holds key 1 only; key 2 (BCCH frame without a window) must not enter it
	e = cudaGetLastError();
	if (e == cudaSuccess)
		e = c_step();                         // frames in front of the first BCCH frame
	// a channel advances by 8 frames per (B, C) pair unless an SI1 re-times it; the planned pairs are followed by
	// checked ones until every channel is through
	int planned = (max_frames + 7) / 8 + 1, pairs = 0;
	while (e == cudaSuccess) {
		if ((e = b_step()) != cudaSuccess || (e = c_step()) != cudaSuccess)
			break;
		if (++pairs < planned)
			continue;
		cudaMemcpyAsync(h_done, pc.all_done, sizeof(int32_t), cudaMemcpyDeviceToHost, cs);
		if ((e = cudaStreamSynchronize(cs)) != cudaSuccess || *h_done != 0 || pairs > max_frames + 2)
			break;
	}
	if (e == cudaSuccess) {
		rx_final_kernel<<<grid, tb, 0, cs>>>(st, d_align_out, d_ferr_out, n);
		e = cudaGetLastError();
		launches++;
	}
	g_launches.fetch_add(launches);
	return s.finish(e, "rx_bcch_batch kernels");
}

static std::atomic<int> g_rx_lockstep{0};
static bool rx_lockstep() { return g_rx_lockstep.load(std::memory_order_relaxed) != 0; }

extern "C" int gmr1b200_set_rx_lockstep(int on) { return g_rx_lockstep.exchange(on ? 1 : 0); }

extern "C" int gmr1b200_rx_bcch_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                                      const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                                      int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                                      int32_t *n_frames, int32_t *align_out, float *freq_err_out, void *stream)
{
	return (rx_lockstep() ? rx_bcch_walk : rx_bcch_walk_paced)(iq, iq_len, rec_ofs, rec_len, align0, freq_err0, sps, n,
	                                                           max_frames, kind, fn, crc, conv, l2, n_frames, align_out,
	                                                           freq_err_out, nullptr, nullptr, stream);
}

extern "C" int gmr1b200_rx_bcch_ass_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs,
                                          const int32_t *rec_len, const int32_t *align0, const float *freq_err0, int sps,
                                          int n, int max_frames, int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv,
                                          uint8_t *l2, int32_t *n_frames, int32_t *align_out, float *freq_err_out,
                                          int32_t *tch3, float *tch3_energy, void *stream)
{
	if (!tch3)
		return set_err(-EINVAL, "rx_bcch_ass_batch: tch3 NULL");
	return (rx_lockstep() ? rx_bcch_walk : rx_bcch_walk_paced)(iq, iq_len, rec_ofs, rec_len, align0, freq_err0, sps, n,
	                                                           max_frames, kind, fn, crc, conv, l2, n_frames, align_out,
	                                                           freq_err_out, tch3, tch3_energy, stream);
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

