// rx_call.cu - the reference receiver's WHOLE frame loop for N channels in lock step: control channels, the TCH3
// burst loop behind an IMMEDIATE ASSIGNMENT and the TCH9 loop behind an ASSIGNMENT COMMAND 1.
// Replaces, for a batch of channels per call,
//   process_bcch                 src/gmr1_rx.c:853-895   frame walk: rx_bcch, rx_ccch, rx_tch3, rx_tch9 per frame
//   rx_bcch / rx_ccch / bcch_tdma_align   :194-236, 747-851   (kernels shared with rx_sched.cu, rx_common.cuh)
//   rx_tch3_init, rx_tch3        :362-381, 538-600       energy gate, DKAB vs burst, detect, running averages, release
//   _rx_tch3_dkab                :383-398
//   _rx_tch3_facch, _rx_tch3_facch_flush   :400-494      4-burst FACCH3 assembly, plain attempt, ciphered retry,
//                                                         cipher discovery, ASS.CMD -> TCH9
//   _rx_tch3_speech              :496-536                A5 mask of the frame, TCH3 decode
//   rx_tch9_init, rx_tch9        :264-355                NT9 demod, FACCH9 / TCH9 by sync id, depth-3 interleaver history
// The per-channel DECISIONS are the __host__ __device__ functions of tch3_state.cuh (verified on the CPU against the
// reference application by tests/test_tch3_state_emu.py); the signal processing between them is the batched kernels
// of this library run on compacted lists with device-side counts.  Channels are independent, frames of one channel
// are not: the loop runs over frames on the host (everything enqueued on one stream, no host round trip) and over
// channels on the device.  Per frame, after the control-channel part:
//   t3_prep      (warp / channel)   window on the traffic recording, burst_energy, energy gate       -> DKAB | burst
//   compact<2>, dkab_kernel, demod_kernel<detect>   on the two lists
//   t3_route     (thread / channel) DKAB result (averages, release) | detected type                  -> FACCH3 | speech
//   compact<2>, demod x 2, A5 masks + TCH3 decode for the speech list
//   t3_facch     (warp / entry)     flush-before / store / flush-after of the FACCH3 soft-bit store   -> flush snapshots
//   A5 masks x 2, FACCH3 decode x 2 (plain-or-known-cipher attempt, ciphered retry)
//   t3_result    (thread / channel) retry decision, cipher discovery, ASS.CMD -> TCH9 state, per-frame record
//   t9_prep, compact<1>, NT9 demod, A5 masks, t9_route (prev indices from the history ring), FACCH9 + TCH9 decode,
//   t9_result    (warp / entry)     record, history push
//   rx_advance
#include "rx_common.cuh"

namespace {

constexpr int NT3_SYMS = 117, NT9_SYMS = 351;
constexpr int TREC = 12, CREC = 6;                    // int32 fields per frame record (include/gmr1_b200.h)

struct CallState {
	const int64_t *tch_ofs, *csd_ofs;                 // [n] first sample of the traffic / CSD recording in iq, -1 = none
	const uint8_t *kc;                                // [n][8]
	Tch3State *t3;
	int8_t    *t3_store;                              // [n][416]
	Tch9State *t9;
	int32_t   *t9_head, *t9_cnt;                      // history ring: most recent slot, valid entries (0..2)
	int32_t   *key, *key2, *key9, *slot, *slot2, *slot9, *pay;    // [n] list key / position per stage, compaction payload
	float     *be;                                    // [n] energy of the traffic window
	int64_t   *wofs, *wofs9;                          // [n] absolute window start of the traffic / CSD window
};

// rx_tch3 up to the energy gate (:538-585)
__global__ void __launch_bounds__(128) t3_prep_kernel(const float2 *__restrict__ iq, RxState st, CallState cs, int n, int sps)
{
	const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (i >= n)
		return;
	int key = 0;
	if (!st.done[i] && cs.t3[i].active && cs.tch_ofs[i] >= 0) {
		const int win = sps + (sps >> 1), etoa = win >> 1;                    // :549-550
		const int begin = st.align[i] + sps * cs.t3[i].tn * SYM_PER_SLOT - etoa;
		const int len = NT3_SYMS * sps + win;
		if (begin >= 0 && begin + len <= st.rec_len[i]) {
			const float be = window_energy(iq + cs.tch_ofs[i] + begin, len, lane);
			if (lane == 0) {
				key = tch3_gate(cs.t3[i], be);                                // 1 = DKAB, 2 = burst
				cs.be[i] = be;
				cs.wofs[i] = cs.tch_ofs[i] + begin;
				cs.pay[i] = cs.t3[i].p;
			}
		}
	}
	if (lane == 0)
		cs.key[i] = key;
}

struct T3In {                                         // results of the first-stage kernels, per list entry
	const int32_t *dkab_rv;                           // DKAB list
	const int32_t *bt_id;                             // burst list: 0 = NT3 FACCH, else speech
};

// after gmr1_dkab_demod / gmr1_pi4cxpsk_detect (:558-598)
__global__ void __launch_bounds__(128) t3_route_kernel(RxState st, CallState cs, T3In in, int32_t *tch_rec, int n,
                                                       const int *frame_p, int max_frames)
{
	const int frame = *frame_p;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int32_t *rec = tch_rec + ((size_t)i * max_frames + frame) * TREC;
	const int key = cs.key[i];
	int key2 = 0;
	if (key == TCH3_GATE_DKAB) {
		const int rv = in.dkab_rv[cs.slot[i]];
		rec[0] = rv == 0 ? GMR1B200_TCH_DKAB : GMR1B200_TCH_DKAB_MISS;
		rec[1] = tch3_dkab_result(cs.t3[i], cs.be[i], rv);
	} else if (key == TCH3_GATE_BURST) {
		key2 = in.bt_id[cs.slot[i]] == 0 ? 1 : 2;
		rec[0] = key2 == 1 ? GMR1B200_TCH_FACCH3 : GMR1B200_TCH_SPEECH;
		cs.pay[i] = cs.t3[i].ciph;
	}
	cs.key2[i] = key2;
}

// unit t of a list with `mul` cipher streams per entry: Kc of the entry's channel, frame number of the channel
// (fn_unit == NULL) or per unit, algorithm per entry (alg_entry) or alg0
__global__ void __launch_bounds__(128) a5_gather_kernel(const int32_t *idx, const int32_t *count, int mul, const uint8_t *kc,
                                                        const int32_t *fn_ch, const uint32_t *fn_unit,
                                                        const int32_t *alg_entry, int alg0, uint8_t *key_out,
                                                        uint32_t *fn_out, int32_t *alg_out, int n)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= min(n, *count) * mul)
		return;
	const int p = t / mul, ch = idx[p];
	for (int b = 0; b < 8; b++)
		key_out[(size_t)t * 8 + b] = kc[(size_t)ch * 8 + b];
	fn_out[t] = fn_unit ? fn_unit[t] : (uint32_t)fn_ch[ch];
	alg_out[t] = alg_entry ? alg_entry[p] : alg0;
}

struct Flush {                                        // FACCH3 codewords to decode in this frame, per FACCH-list entry
	int8_t   *eb;                                     // [n][416] snapshot of the channel's store (zeros: no flush)
	uint32_t *fn;                                     // [n][4]   bi_fn at the time of the flush
	int32_t  *flag;                                   // [n] 0 = no flush, 1 = before storing this burst, 2 = after
	int32_t  *ciph;                                   // [n] the channel's cipher flag at the time of the flush
};

// _rx_tch3_facch (:455-494) for the entries of the FACCH list: one warp per entry
__global__ void __launch_bounds__(128) t3_facch_kernel(RxState st, CallState cs, const int32_t *idx, const int32_t *count,
                                                       const int8_t *ebits, const int32_t *sync_id, Flush fl, int n)
{
	const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (p >= min(n, *count))
		return;
	const int ch = idx[p];
	Tch3State &s = cs.t3[ch];
	int8_t *store = cs.t3_store + (size_t)ch * 416;
	int8_t *snap = fl.eb + (size_t)p * 416;
	const int sid = sync_id[p];
	const uint32_t fn = (uint32_t)st.fn[ch];
	const bool before = tch3_facch_flush_before(s, sid);
	__syncwarp();
	if (before) {                                     // the old group is closed first (:479-481)
		for (int k = lane; k < 416; k += 32) {
			snap[k] = store[k];
			store[k] = 0;
		}
		if (lane < 4)
			fl.fn[4 * p + lane] = s.bi_fn[lane];
		__syncwarp();
	}
	const int bi = (int)(fn & 3u);
	for (int k = lane; k < 104; k += 32)              // store this burst (:483-488)
		store[104 * bi + k] = ebits[(size_t)p * 104 + k];
	__syncwarp();
	int flag = before ? 1 : 0, ciph = s.ciph, cnt = 0;
	if (lane == 0) {
		if (before) {                                 // state part of the flush (tch3_flush_done without its result part)
			s.burst_cnt = 0;
			for (int k = 0; k < 4; k++)
				s.bi_fn[k] = 0xffffffffu;
		}
		s.sync_id = sid;
		s.bi_fn[bi] = fn;
		s.burst_cnt += 1;
		cnt = s.burst_cnt;
	}
	cnt = __shfl_sync(0xffffffffu, cnt, 0);
	if (cnt == 4) {                                   // the codeword is complete (:490-491); never after a flush-before
		__syncwarp();
		for (int k = lane; k < 416; k += 32) {
			snap[k] = store[k];
			store[k] = 0;
		}
		if (lane < 4)
			fl.fn[4 * p + lane] = s.bi_fn[lane];
		__syncwarp();
		if (lane == 0) {
			s.sync_id ^= 1;
			s.burst_cnt = 0;
			for (int k = 0; k < 4; k++)
				s.bi_fn[k] = 0xffffffffu;
		}
		flag = 2;
	} else if (!before) {
		for (int k = lane; k < 416; k += 32)
			snap[k] = 0;
		if (lane < 4)
			fl.fn[4 * p + lane] = 0xffffffffu;
	}
	if (lane == 0) {
		fl.flag[p] = flag;
		fl.ciph[p] = ciph;
	}
}

struct T3Res {                                        // per list entry
	const int32_t *f_sync;                            // FACCH list: sync id of the burst
	const int32_t *x_crc[2], *x_conv[2];              // FACCH list: the two FACCH3 decode attempts
	const uint8_t *x_l2[2];                           // [n][10]
	const uint8_t *s_f0, *s_f1;                       // speech list: frames [n][10]
	const int32_t *s_c0, *s_c1;
};

// the end of _rx_tch3_facch_flush (:417-449), speech results, the frame's record
__global__ void __launch_bounds__(128) t3_result_kernel(RxState st, CallState cs, Flush fl, T3Res r, const int32_t *tch3_ass,
                                                        int32_t *tch_rec, uint8_t *tch_data, int n, const int *frame_p,
                                                        int max_frames)
{
	const int frame = *frame_p;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const size_t fr = (size_t)i * max_frames + frame;
	int32_t *rec = tch_rec + fr * TREC;
	uint8_t *dat = tch_data + fr * 20;
	rec[10] = (tch3_ass[4 * i + 3] == frame && tch3_ass[4 * i + 0]) ? tch3_ass[4 * i + 1] : -1;
	const int key2 = cs.key2[i];
	if (key2 == 1) {
		const int p = cs.slot2[i];
		Tch3State &s = cs.t3[i];
		rec[2] = r.f_sync[p];
		const int flag = fl.flag[p];
		if (flag) {
			int crc = r.x_crc[0][p], att = 0;
			rec[3] = 1;
			rec[4] = crc;
			rec[5] = r.x_conv[0][p];
			const bool retried = !fl.ciph[p] && crc != 0;                      // :417-430
			if (retried) {
				att = 1;
				crc = r.x_crc[1][p];
				rec[3] = 2;
				rec[6] = crc;
				rec[7] = r.x_conv[1][p];
				if (!crc)
					s.ciph = 1;
			}
			if (!crc) {
				const uint8_t *l2 = r.x_l2[att] + (size_t)p * 10;
				for (int b = 0; b < 10; b++)
					dat[b] = l2[b];
				rec[11] = 1;                                                   // a good FACCH3 message in dat[0..9]
				if (cs.csd_ofs[i] >= 0 && tch9_init_from_facch3(cs.t9[i], l2, true)) {    // :436-441
					cs.t9_head[i] = 0;                                         // gmr1_interleaver_init :273
					cs.t9_cnt[i] = 0;
				}
			}
			// flush-after: sync_id was toggled in t3_facch_kernel; flush-before: it now is the new burst's
		}
	} else if (key2 == 2) {
		const int p = cs.slot2[i];
		for (int b = 0; b < 10; b++) {
			dat[b] = r.s_f0[(size_t)p * 10 + b];
			dat[10 + b] = r.s_f1[(size_t)p * 10 + b];
		}
		rec[8] = r.s_c0[p];
		rec[9] = r.s_c1[p];
	}
}

// rx_tch9 up to the window (:281-297)
__global__ void __launch_bounds__(128) t9_prep_kernel(RxState st, CallState cs, int n, int sps)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int key = 0;
	if (!st.done[i] && cs.t9[i].active && cs.csd_ofs[i] >= 0) {
		const int win = sps + (sps >> 1), etoa = win >> 1;
		const int begin = st.align[i] + sps * cs.t9[i].tn * SYM_PER_SLOT - etoa;
		const int len = NT9_SYMS * sps + win;
		if (begin >= 0 && begin + len <= st.rec_len[i]) {
			key = 1;
			cs.wofs9[i] = cs.csd_ofs[i] + begin;
		}
	}
	cs.key9[i] = key;
}

// predecessors of every NT9 burst in the channel's history ring (rows n + 2 ch + {0, 1} of the soft-bit / mask buffers)
__global__ void __launch_bounds__(128) t9_route_kernel(CallState cs, const int32_t *idx, const int32_t *count,
                                                       int32_t *prev1, int32_t *prev2, int n)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= min(n, *count))
		return;
	const int ch = idx[p], head = cs.t9_head[ch], cnt = cs.t9_cnt[ch];
	prev1[p] = cnt >= 1 ? n + 2 * ch + head : -1;
	prev2[p] = cnt >= 2 ? n + 2 * ch + (head ^ 1) : -1;
}

struct T9Res {
	const int32_t *sync;
	const int32_t *f_crc, *f_conv, *t_conv;
	const uint8_t *f_l2, *t_l2;                       // [n][38], [n][60]
};

// rx_tch9 after the demodulation (:305-352): one warp per entry
__global__ void __launch_bounds__(128) t9_result_kernel(CallState cs, const int32_t *idx, const int32_t *count, int8_t *eb,
                                                        uint8_t *ciph, T9Res r, int32_t *csd_rec, uint8_t *csd_data, int n,
                                                        const int *frame_p, int max_frames)
{
	const int frame = *frame_p;
	const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (p >= min(n, *count))
		return;
	const int ch = idx[p];
	const size_t fr = (size_t)ch * max_frames + frame;
	int32_t *rec = csd_rec + fr * CREC;
	uint8_t *dat = csd_data + fr * 60;
	const int sid = r.sync[p];
	const int8_t *cur = eb + (size_t)p * 662;
	if (tch9_is_facch9(sid)) {
		for (int b = lane; b < 38; b += 32)
			dat[b] = r.f_l2[(size_t)p * 38 + b];
		if (lane == 0) {
			rec[0] = GMR1B200_CSD_FACCH9;
			rec[1] = sid;
			rec[2] = r.f_crc[p];
			rec[3] = r.f_conv[p];
		}
		return;
	}
	int s = 0;                                        // mean soft-bit magnitude the reference prints (:325-327)
	for (int k = lane; k < 662; k += 32)
		s += cur[k] < 0 ? -(int)cur[k] : (int)cur[k];
#pragma unroll
	for (int o = 16; o; o >>= 1)
		s += __shfl_xor_sync(0xffffffffu, s, o);
	for (int b = lane; b < 60; b += 32)
		dat[b] = r.t_l2[(size_t)p * 60 + b];
	// this burst becomes the most recent entry of the channel's interleaver history
	const int slot = cs.t9_head[ch] ^ 1;
	int8_t *he = eb + (size_t)(n + 2 * ch + slot) * 662;
	uint8_t *hc = ciph + (size_t)(n + 2 * ch + slot) * 658;
	const uint8_t *cc = ciph + (size_t)p * 658;
	for (int k = lane; k < 662; k += 32)
		he[k] = cur[k];
	for (int k = lane; k < 658; k += 32)
		hc[k] = cc[k];
	if (lane == 0) {
		rec[0] = GMR1B200_CSD_TCH9;
		rec[1] = sid;
		rec[3] = r.t_conv[p];
		rec[4] = s / 662;
		cs.t9_head[ch] = slot;
		cs.t9_cnt[ch] = min(cs.t9_cnt[ch] + 1, 2);
	}
}

}  // namespace

extern "C" int gmr1b200_rx_call_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                                      const int64_t *tch_ofs, const int64_t *csd_ofs, const uint8_t *kc,
                                      const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                                      int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                                      int32_t *n_frames, int32_t *tch_rec, uint8_t *tch_data, int32_t *csd_rec,
                                      uint8_t *csd_data, void *stream)
{
	if (!iq || !rec_ofs || !rec_len || !tch_ofs || !align0 || n < 0 || max_frames < 1 || sps < 1 || sps > 16 || !kind ||
	    !fn || !crc || !conv || !l2 || !n_frames || !tch_rec || !tch_data || (csd_ofs && (!csd_rec || !csd_data)))
		return set_err(-EINVAL, "rx_call_batch: bad argument");
	if (n == 0)
		return 0;
	if (!recordings_in_range(rec_ofs, rec_len, n, iq_len))
		return set_err(-EINVAL, "rx_call_batch: a recording lies outside iq_len");
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
	out.tch3 = s.tmp<int32_t>(N * 4); out.tch3_energy = nullptr;
	int32_t *d_trec = s.out(tch_rec, NF * TREC);
	uint8_t *d_tdat = s.out(tch_data, NF * 20);
	int32_t *d_crec = csd_ofs ? s.out(csd_rec, NF * CREC) : nullptr;
	uint8_t *d_cdat = csd_ofs ? s.out(csd_data, NF * 60) : nullptr;

	st.align = s.tmp<int32_t>(N); st.freq_err = s.tmp<float>(N);
	st.fn = s.tmp<int32_t>(N); st.delay = s.tmp<int32_t>(N); st.stn = s.tmp<int32_t>(N);
	st.bcch_energy = s.tmp<float>(N); st.done = s.tmp<int32_t>(N);
	st.kind = s.tmp<int32_t>(N); st.begin = s.tmp<int32_t>(N); st.energy = s.tmp<float>(N); st.slot = s.tmp<int32_t>(N);
	RxLists ls = {};
	ls.count = s.tmp<int32_t>(2);
	int8_t *eb[2]; float *toa[2], *ferr[2]; int32_t *dcrc[2], *dconv[2]; uint8_t *dl2[2], *dscr[2];
	const int ebits[2] = {424, 432}, bt[2] = {BT_BCCH, BT_DC6}, ch[2] = {CH_BCCH, CH_CCCH};
	for (int k = 0; k < 2; k++) {
		ls.ofs[k] = s.tmp<int64_t>(N); ls.fs[k] = s.tmp<float>(N);
		eb[k] = s.tmp<int8_t>(N * ebits[k]); toa[k] = s.tmp<float>(N); ferr[k] = s.tmp<float>(N);
		dcrc[k] = s.tmp<int32_t>(N); dconv[k] = s.tmp<int32_t>(N); dl2[k] = s.tmp<uint8_t>(N * 24);
		dscr[k] = s.tmp<uint8_t>(decode_scratch_bytes(ch[k], n));
	}

	// ---- traffic-channel state and scratch
	CallState c = {};
	c.tch_ofs = s.in(tch_ofs, N);
	int64_t *no_csd = nullptr;
	if (csd_ofs)
		c.csd_ofs = s.in(csd_ofs, N);
	else
		c.csd_ofs = no_csd = s.tmp<int64_t>(N);
	uint8_t *zero_kc = nullptr;
	if (kc)
		c.kc = s.in(kc, N * 8);
	else
		c.kc = zero_kc = s.tmp<uint8_t>(N * 8);
	c.t3 = s.tmp<Tch3State>(N); c.t3_store = s.tmp<int8_t>(N * 416);
	c.t9 = s.tmp<Tch9State>(N); c.t9_head = s.tmp<int32_t>(N); c.t9_cnt = s.tmp<int32_t>(N);
	c.key = s.tmp<int32_t>(N); c.key2 = s.tmp<int32_t>(N); c.key9 = s.tmp<int32_t>(N);
	c.slot = s.tmp<int32_t>(N); c.slot2 = s.tmp<int32_t>(N); c.slot9 = s.tmp<int32_t>(N); c.pay = s.tmp<int32_t>(N);
	c.be = s.tmp<float>(N); c.wofs = s.tmp<int64_t>(N); c.wofs9 = s.tmp<int64_t>(N);
	auto mk2 = [&](Lists<2> &l) {
		l.count = s.tmp<int32_t>(2); l.idx = s.tmp<int32_t>(2 * N); l.ofs = s.tmp<int64_t>(2 * N);
		l.fs = s.tmp<float>(2 * N); l.pay = s.tmp<int32_t>(2 * N);
	};
	Lists<2> la = {}, lb = {};                       // stage 1: DKAB | burst;  stage 2: FACCH3 | speech
	mk2(la);
	mk2(lb);
	Lists<1> l9 = {};
	l9.count = s.tmp<int32_t>(1); l9.idx = s.tmp<int32_t>(N); l9.ofs = s.tmp<int64_t>(N); l9.fs = s.tmp<float>(N);
	l9.pay = nullptr;
	// stage 1 results
	int32_t *dk_rv = s.tmp<int32_t>(N), *det_bt = s.tmp<int32_t>(N);
	float *dk_toa = s.tmp<float>(N);
	BurstTab *d_det = s.tmp<BurstTab>(2);             // detect candidates in the reference's order: FACCH, speech (:540-544)
	// stage 2
	int8_t *f_eb = s.tmp<int8_t>(N * 104), *s_eb = s.tmp<int8_t>(N * 212);
	int32_t *f_sync = s.tmp<int32_t>(N);
	uint8_t *a_key = s.tmp<uint8_t>(4 * N * 8);
	uint32_t *a_fn = s.tmp<uint32_t>(4 * N);
	int32_t *a_alg = s.tmp<int32_t>(4 * N);
	uint8_t *s_ciph = s.tmp<uint8_t>(N * 208), *s_f0 = s.tmp<uint8_t>(N * 10), *s_f1 = s.tmp<uint8_t>(N * 10);
	int32_t *s_c0 = s.tmp<int32_t>(N), *s_c1 = s.tmp<int32_t>(N);
	uint8_t *scr_t3 = s.tmp<uint8_t>(decode_scratch_bytes(CH_TCH3, n));
	Flush fl = {};
	fl.eb = s.tmp<int8_t>(N * 416); fl.fn = s.tmp<uint32_t>(N * 4); fl.flag = s.tmp<int32_t>(N); fl.ciph = s.tmp<int32_t>(N);
	uint8_t *x_mask[2] = {s.tmp<uint8_t>(N * 384), s.tmp<uint8_t>(N * 384)};
	uint8_t *x_l2[2] = {s.tmp<uint8_t>(N * 10), s.tmp<uint8_t>(N * 10)};
	int32_t *x_crc[2] = {s.tmp<int32_t>(N), s.tmp<int32_t>(N)}, *x_conv[2] = {s.tmp<int32_t>(N), s.tmp<int32_t>(N)};
	uint8_t *scr_f3 = s.tmp<uint8_t>(decode_scratch_bytes(CH_FACCH3, n));
	// TCH9: rows 0..n-1 = this frame's bursts, rows n + 2 ch + {0, 1} = the channel's history
	int8_t *t_eb = nullptr;
	uint8_t *t_ciph = nullptr, *t_l2f = nullptr, *t_l2t = nullptr, *scr_f9 = nullptr, *scr_t9 = nullptr;
	int32_t *t_sync = nullptr, *t_prev1 = nullptr, *t_prev2 = nullptr, *t_fcrc = nullptr, *t_fconv = nullptr, *t_tconv = nullptr;
	if (csd_ofs) {
		t_eb = s.tmp<int8_t>(3 * N * 662); t_ciph = s.tmp<uint8_t>(3 * N * 658);
		t_sync = s.tmp<int32_t>(N); t_prev1 = s.tmp<int32_t>(N); t_prev2 = s.tmp<int32_t>(N);
		t_l2f = s.tmp<uint8_t>(N * 38); t_l2t = s.tmp<uint8_t>(N * 60);
		t_fcrc = s.tmp<int32_t>(N); t_fconv = s.tmp<int32_t>(N); t_tconv = s.tmp<int32_t>(N);
		scr_f9 = s.tmp<uint8_t>(decode_scratch_bytes(CH_FACCH9, n));
		scr_t9 = s.tmp<uint8_t>(decode_scratch_bytes(CH_TCH9_9K6, n));
	}
	int *d_frame = s.tmp<int>(1);
	if (s.failed())
		return s.finish(cudaSuccess, "rx_call_batch: staging");
	cudaMemsetAsync(d_frame, 0, sizeof(int), cs);

	const int tb = 128, grid = (n + tb - 1) / tb, wgrid = (n + 3) / 4;
	rx_init_kernel<<<grid, tb, 0, cs>>>(st, d_align0, d_ferr0, out.n_frames, out.tch3, nullptr, n);
	cudaMemsetAsync(out.kind, 0, NF * sizeof(int32_t), cs);
	cudaMemsetAsync(out.crc, 0xff, NF * sizeof(int32_t), cs);
	cudaMemsetAsync(d_trec, 0, NF * TREC * sizeof(int32_t), cs);
	cudaMemsetAsync(d_tdat, 0, NF * 20, cs);
	cudaMemsetAsync(c.t3, 0, N * sizeof(Tch3State), cs);        // chan_desc is zeroed in main(), gmr1_rx.c:906
	cudaMemsetAsync(c.t3_store, 0, N * 416, cs);
	cudaMemsetAsync(c.t9, 0, N * sizeof(Tch9State), cs);
	cudaMemsetAsync(c.t9_head, 0, N * sizeof(int32_t), cs);
	cudaMemsetAsync(c.t9_cnt, 0, N * sizeof(int32_t), cs);
	if (no_csd)
		cudaMemsetAsync(no_csd, 0xff, N * sizeof(int64_t), cs);
	if (zero_kc)
		cudaMemsetAsync(zero_kc, 0, N * 8, cs);
	if (csd_ofs) {
		cudaMemsetAsync(d_crec, 0, NF * CREC * sizeof(int32_t), cs);
		cudaMemsetAsync(d_cdat, 0, NF * 60, cs);
		cudaMemsetAsync(t_eb, 0, 3 * N * 662, cs);
		cudaMemsetAsync(t_ciph, 0, 3 * N * 658, cs);
	}
	cudaMemcpyAsync(d_det, d_all + BT_NT3_FACCH, sizeof(BurstTab), cudaMemcpyDeviceToDevice, cs);
	cudaMemcpyAsync(d_det + 1, d_all + BT_NT3_SPEECH, sizeof(BurstTab), cudaMemcpyDeviceToDevice, cs);
	const BurstTab h_det[2] = {burst_tab(BT_NT3_FACCH), burst_tab(BT_NT3_SPEECH)};
	uint64_t launches = 1;
	RxBurstOut bo = {};
	for (int k = 0; k < 2; k++) {
		bo.toa[k] = toa[k]; bo.ferr[k] = ferr[k]; bo.crc[k] = dcrc[k]; bo.conv[k] = dconv[k]; bo.l2[k] = dl2[k];
	}
	const int win3 = sps + (sps >> 1);
	const int wl3 = NT3_SYMS * sps + win3, wl9 = NT9_SYMS * sps + win3;
	e = cudaGetLastError();
	auto demod = [&](int btid, const int64_t *ofs, const float *fs, const int32_t *cnt, int wl, int8_t *ebits_out, int stride,
	                 int32_t *sync_out) -> cudaError_t {
		DemodArgs a = {};
		a.iq = d_iq; a.ofs = ofs; a.n = n; a.sps = sps; a.win_len = wl; a.freq_shift = fs; a.e_toa0 = -1.0f;
		a.ebits = ebits_out; a.ebits_stride = stride; a.sync_id = sync_out; a.n_dev = cnt;
		launches++;
		return launch_demod(a, d_all + btid, &burst_tab(btid), 1, 0, cs);
	};
	auto a5 = [&](const int32_t *idx, const int32_t *cnt, int mul, const uint32_t *fn_unit, const int32_t *alg_entry, int alg0,
	              int nbits, uint8_t *dl) -> cudaError_t {
		a5_gather_kernel<<<(mul * n + tb - 1) / tb, tb, 0, cs>>>(idx, cnt, mul, c.kc, st.fn, fn_unit, alg_entry, alg0, a_key,
		                                                         a_fn, a_alg, n);
		A5Args a = {};
		a.alg = a_alg; a.key = a_key; a.fn = a_fn; a.n = mul * n; a.nbits = nbits; a.stride = nbits; a.dl = dl;
		a.n_dev = cnt; a.n_dev_mul = mul;
		launches += 2;
		return launch_a5(a, cs);
	};
	// the launches of one frame (identical for every frame: the frame index lives in d_frame)
	auto frame = [&]() -> cudaError_t {
		cudaError_t e = cudaSuccess;
		// ---- control channels (as rx_bcch_walk)
		rx_prep_kernel<<<wgrid, 128, 0, cs>>>(d_iq, st, n, sps);
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
				return e;
			DecodeArgs d = {};
			d.ebits = eb[k]; d.n = n; d.l2 = dl2[k]; d.conv = dconv[k]; d.crc = dcrc[k];
			d.n_dev = ls.count + k;
			d.dec_scratch = dscr[k];
			e = launch_decode(ch[k], d, cs);
			launches += 2;
		}
		if (e != cudaSuccess)
			return e;
		rx_update_kernel<<<grid, tb, 0, cs>>>(st, bo, out, n, sps, d_frame, max_frames, c.t3, c.t3_store, false);
		// ---- rx_tch3
		t3_prep_kernel<<<wgrid, 128, 0, cs>>>(d_iq, st, c, n, sps);
		compact_kernel<2><<<1, 1024, 0, cs>>>(c.key, c.wofs, st.freq_err, c.pay, n, c.slot, la);
		launches += 3;
		{
			MiscArgs m = {};
			m.iq = d_iq; m.ofs = la.ofs; m.n = n; m.win_len = wl3; m.sps = sps; m.freq_shift = la.fs; m.dkab_p = la.pay;
			m.toa = dk_toa; m.rv = dk_rv; m.n_dev = la.count;
			if ((e = launch_dkab(m, cs)) != cudaSuccess)
				return e;
			DemodArgs a = {};
			a.iq = d_iq; a.ofs = la.ofs + N; a.n = n; a.sps = sps; a.win_len = wl3; a.freq_shift = la.fs + N;
			a.e_toa0 = (float)(win3 >> 1);                                         // :587-591
			a.bt_id = det_bt; a.n_dev = la.count + 1;
			if ((e = launch_demod(a, d_det, h_det, 2, 1, cs)) != cudaSuccess)
				return e;
			launches += 2;
		}
		T3In in = {dk_rv, det_bt};
		t3_route_kernel<<<grid, tb, 0, cs>>>(st, c, in, d_trec, n, d_frame, max_frames);
		compact_kernel<2><<<1, 1024, 0, cs>>>(c.key2, c.wofs, st.freq_err, c.pay, n, c.slot2, lb);
		launches += 2;
		if ((e = demod(BT_NT3_FACCH, lb.ofs, lb.fs, lb.count, wl3, f_eb, 104, f_sync)) != cudaSuccess)
			return e;
		if ((e = demod(BT_NT3_SPEECH, lb.ofs + N, lb.fs + N, lb.count + 1, wl3, s_eb, 212, nullptr)) != cudaSuccess)
			return e;
		// speech: A5 mask of (Kc, fn) where the channel is known to be ciphered, TCH3 decode (:518-524)
		if ((e = a5(lb.idx + N, lb.count + 1, 1, nullptr, lb.pay + N, 0, 208, s_ciph)) != cudaSuccess)
			return e;
		{
			DecodeArgs d = {};
			d.ebits = s_eb; d.ciph = s_ciph; d.n = n; d.l2 = s_f0; d.l2b = s_f1; d.conv = s_c0; d.conv1 = s_c1;
			d.tch3_m = 0; d.n_dev = lb.count + 1; d.dec_scratch = scr_t3;
			if ((e = launch_decode(CH_TCH3, d, cs)) != cudaSuccess)
				return e;
			launches++;
		}
		// FACCH3: store / flush, then both decode attempts of every flushed codeword
		t3_facch_kernel<<<wgrid, 128, 0, cs>>>(st, c, lb.idx, lb.count, f_eb, f_sync, fl, n);
		launches++;
		if ((e = a5(lb.idx, lb.count, 4, fl.fn, fl.ciph, 0, 96, x_mask[0])) != cudaSuccess)
			return e;
		if ((e = a5(lb.idx, lb.count, 4, fl.fn, nullptr, 1, 96, x_mask[1])) != cudaSuccess)
			return e;
		for (int k = 0; k < 2 && e == cudaSuccess; k++) {
			DecodeArgs d = {};
			d.ebits = fl.eb; d.ciph = x_mask[k]; d.n = n; d.l2 = x_l2[k]; d.conv = x_conv[k]; d.crc = x_crc[k];
			d.n_dev = lb.count; d.dec_scratch = scr_f3;
			e = launch_decode(CH_FACCH3, d, cs);
			launches++;
		}
		if (e != cudaSuccess)
			return e;
		T3Res r = {};
		r.f_sync = f_sync;
		for (int k = 0; k < 2; k++) {
			r.x_crc[k] = x_crc[k]; r.x_conv[k] = x_conv[k]; r.x_l2[k] = x_l2[k];
		}
		r.s_f0 = s_f0; r.s_f1 = s_f1; r.s_c0 = s_c0; r.s_c1 = s_c1;
		t3_result_kernel<<<grid, tb, 0, cs>>>(st, c, fl, r, out.tch3, d_trec, d_tdat, n, d_frame, max_frames);
		launches++;
		// ---- rx_tch9
		if (csd_ofs) {
			t9_prep_kernel<<<grid, tb, 0, cs>>>(st, c, n, sps);
			compact_kernel<1><<<1, 1024, 0, cs>>>(c.key9, c.wofs9, st.freq_err, nullptr, n, c.slot9, l9);
			launches += 2;
			if ((e = demod(BT_NT9, l9.ofs, l9.fs, l9.count, wl9, t_eb, 662, t_sync)) != cudaSuccess)
				return e;
			if ((e = a5(l9.idx, l9.count, 1, nullptr, nullptr, 1, 658, t_ciph)) != cudaSuccess)      // :308, :322
				return e;
			t9_route_kernel<<<grid, tb, 0, cs>>>(c, l9.idx, l9.count, t_prev1, t_prev2, n);
			DecodeArgs d = {};
			d.ebits = t_eb; d.ciph = t_ciph; d.n = n; d.l2 = t_l2f; d.conv = t_fconv; d.crc = t_fcrc;
			d.n_dev = l9.count; d.dec_scratch = scr_f9;
			if ((e = launch_decode(CH_FACCH9, d, cs)) != cudaSuccess)
				return e;
			DecodeArgs d9 = {};
			d9.ebits = t_eb; d9.ciph = t_ciph; d9.n = n; d9.l2 = t_l2t; d9.conv = t_tconv; d9.prev1 = t_prev1; d9.prev2 = t_prev2;
			d9.n_dev = l9.count; d9.dec_scratch = scr_t9;
			if ((e = launch_decode(CH_TCH9_9K6, d9, cs)) != cudaSuccess)
				return e;
			T9Res r9 = {t_sync, t_fcrc, t_fconv, t_tconv, t_l2f, t_l2t};
			t9_result_kernel<<<wgrid, 128, 0, cs>>>(c, l9.idx, l9.count, t_eb, t_ciph, r9, d_crec, d_cdat, n, d_frame, max_frames);
			launches += 4;
		}
		rx_advance_kernel<<<grid, tb, 0, cs>>>(st, out.n_frames, n, sps, d_frame);
		rx_tick_kernel<<<1, 1, 0, cs>>>(d_frame);
		launches += 3;
		return cudaGetLastError();
	};
	FrameGraph fg(cs);
	bool graph = false;
	uint64_t per_frame = 0;
	for (int f = 0; f < max_frames && e == cudaSuccess; f++) {
		if (f == 1 && max_frames >= 4 && ws.capturable()) {
			if (fg.begin() == cudaSuccess) {
				const uint64_t keep = launches;
				const cudaError_t ce = frame();          // captured, not executed
				launches = keep;
				if (ce == cudaSuccess && fg.end() == cudaSuccess)
					graph = true;
				else
					fg.abort();
			} else
				cudaGetLastError();
		}
		const uint64_t l0 = launches;
		e = graph ? fg.launch() : frame();
		if (!graph)
			per_frame = launches - l0;
		else
			launches += per_frame;
	}
	g_launches.fetch_add(launches);
	const int rc = s.finish(e, "rx_call_batch kernels");
	ws.join();
	return rc;
}
