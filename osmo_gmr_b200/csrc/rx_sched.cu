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

