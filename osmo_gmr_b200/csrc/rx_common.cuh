// rx_common.cuh - state, kernels and helpers of the receiver frame loop shared by the BCCH/CCCH walk (rx_sched.cu)
// and the whole-call walk (rx_call.cu): window geometry and energy (burst_map / burst_energy, src/gmr1_rx.c:149-182),
// which burst a frame carries (process_bcch :853-895), result handling and tracking feedback of rx_bcch / rx_ccch
// (:747-851), bcch_tdma_align (:194-236).  Included by exactly those two translation units.
#pragma once
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "decode_unit.cuh"
#include "launch.h"
#include "tch3_state.cuh"

#include <stdlib.h>
#include <mutex>
#include <vector>

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

// t3 / t3_store (optional): the channel's TCH3 state and FACCH3 soft-bit store, initialised by an IMM.ASS exactly as
// rx_tch3_init does (the whole-call walk); advance = false leaves the step to the next frame to rx_advance_kernel
// (the frame index comes from a device counter, so that the launches of one frame are the same for every frame
// and can be replayed as a CUDA graph: rx_tick_kernel closes a frame)
__global__ void __launch_bounds__(128) rx_update_kernel(RxState st, RxBurstOut bo, RxOut out, int n, int sps,
                                                        const int *frame_p, int max_frames, Tch3State *t3,
                                                        int8_t *t3_store, bool advance)
{
	const int frame = *frame_p;
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
			if (t3)
				tch3_init(t3[i], t3_store + (size_t)i * 416, l2, ref);
		}
	}
	out.crc[rec] = crc;
	out.conv[rec] = conv;
	if (!advance)
		return;
	// next frame (process_bcch :884-891)
	const int frame_len = sps * SLOTS_PER_FRAME * SYM_PER_SLOT;
	st.fn[i] += 1;
	st.align[i] += frame_len;
	out.n_frames[i] = frame + 1;
	if (st.align[i] + 2 * frame_len > st.rec_len[i])
		st.done[i] = 1;
}

// next frame (process_bcch :884-891), for walks that do more per frame after rx_update_kernel
__global__ void __launch_bounds__(128) rx_advance_kernel(RxState st, int32_t *n_frames, int n, int sps, const int *frame_p)
{
	const int frame = *frame_p;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || st.done[i])
		return;
	const int frame_len = sps * SLOTS_PER_FRAME * SYM_PER_SLOT;
	st.fn[i] += 1;
	st.align[i] += frame_len;
	n_frames[i] = frame + 1;
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

template <int NL>
struct Lists {                                        // NL compacted lists over the channels, [NL][n] each
	int32_t *count;                                   // [NL]
	int32_t *idx;                                     // channel of an entry
	int64_t *ofs;                                     // absolute window start
	float   *fs;                                      // freq_shift = -freq_err
	int32_t *pay;                                     // payload (DKAB position / cipher flag)
};

// ordered compaction of the channels by key (0 = in no list, k = list k-1); one CTA
template <int NL>
__global__ void __launch_bounds__(1024) compact_kernel(const int32_t *key, const int64_t *wofs, const float *freq_err,
                                                       const int32_t *pay, int n, int32_t *slot, Lists<NL> ls)
{
	__shared__ int wtot[NL][32], wbase[NL][32], base[NL];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < NL)
		base[tid] = 0;
	__syncthreads();
	for (int i0 = 0; i0 < n; i0 += 1024) {
		const int i = i0 + tid;
		const int k0 = i < n ? key[i] : 0;
		int pos[NL];
#pragma unroll
		for (int k = 0; k < NL; k++) {
			const unsigned m = __ballot_sync(0xffffffffu, k0 == k + 1);
			pos[k] = __popc(m & ((1u << lane) - 1u));
			if (lane == 0)
				wtot[k][warp] = __popc(m);
		}
		__syncthreads();
		if (warp == 0) {
#pragma unroll
			for (int k = 0; k < NL; k++) {
				const int v = wtot[k][lane];
				int s = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int t = __shfl_up_sync(0xffffffffu, s, o);
					if (lane >= o)
						s += t;
				}
				wbase[k][lane] = base[k] + s - v;
			}
		}
		__syncthreads();
		if (k0 > 0 && k0 <= NL) {
			int p = 0;
#pragma unroll
			for (int k = 0; k < NL; k++)
				if (k0 == k + 1)
					p = wbase[k][warp] + pos[k];
			slot[i] = p;
			const size_t e = (size_t)(k0 - 1) * n + p;
			ls.idx[e] = i;
			ls.ofs[e] = wofs[i];
			ls.fs[e] = -freq_err[i];
			if (pay)
				ls.pay[e] = pay[i];
		}
		__syncthreads();
		if (tid < NL)
			base[tid] = wbase[tid][31] + wtot[tid][31];
		__syncthreads();
	}
	if (tid < NL)
		ls.count[tid] = base[tid];
}

__global__ void rx_tick_kernel(int *frame_p) { *frame_p += 1; }

// ---- one frame's launches as a CUDA graph ----------------------------------------------------------------------------
// The walks enqueue the same sequence of small dependent kernels for every frame (449 launches per BCCH / CCCH walk,
// ~40 per frame with the traffic channels): frame 0 is launched directly (it also runs the launchers' lazy
// initialisation, which must not happen inside a capture), frame 1 is captured into a graph, frames 1 .. F-1 replay it.
// Executable graphs are destroyed later, when an event recorded behind their last launch has completed.
class FrameGraph {
public:
	explicit FrameGraph(cudaStream_t cs) : cs_(cs) { sweep(false); }
	static bool enabled()
	{
		static const bool off = [] { const char *e = getenv("GMR1B200_RX_NOGRAPH"); return e && atoi(e) != 0; }();
		return !off;
	}
	cudaError_t begin() { return cudaStreamBeginCapture(cs_, cudaStreamCaptureModeThreadLocal); }
	cudaError_t end()
	{
		cudaError_t e = cudaStreamEndCapture(cs_, &g_);
		if (e == cudaSuccess)
			e = cudaGraphInstantiate(&ex_, g_, 0);
		return e;
	}
	void abort()
	{
		cudaGraph_t g = nullptr;
		cudaStreamEndCapture(cs_, &g);
		if (g)
			cudaGraphDestroy(g);
		cudaGetLastError();
	}
	cudaError_t launch() { return cudaGraphLaunch(ex_, cs_); }
	~FrameGraph()
	{
		if (!ex_ && !g_)
			return;
		Dead d = {ex_, g_, nullptr};
		if (cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming) == cudaSuccess)
			cudaEventRecord(d.done, cs_);
		std::lock_guard<std::mutex> lk(mu());
		dead().push_back(d);
	}

private:
	struct Dead { cudaGraphExec_t ex; cudaGraph_t g; cudaEvent_t done; };
	static std::mutex &mu() { static std::mutex m; return m; }
	static std::vector<Dead> &dead() { static std::vector<Dead> v; return v; }
	static void sweep(bool all)
	{
		std::lock_guard<std::mutex> lk(mu());
		auto &v = dead();
		for (size_t i = 0; i < v.size();) {
			if (all || !v[i].done || cudaEventQuery(v[i].done) == cudaSuccess) {
				if (v[i].ex) cudaGraphExecDestroy(v[i].ex);
				if (v[i].g) cudaGraphDestroy(v[i].g);
				if (v[i].done) cudaEventDestroy(v[i].done);
				v[i] = v.back();
				v.pop_back();
			} else
				i++;
		}
		cudaGetLastError();
	}
	cudaStream_t cs_;
	cudaGraph_t g_ = nullptr;
	cudaGraphExec_t ex_ = nullptr;
};

// The legacy default stream cannot be captured: a walk called with stream = NULL runs on an internal stream that is
// ordered behind the default stream at entry and in front of it at exit (same semantics for the caller).
class WalkStream {
public:
	explicit WalkStream(void *user) : user_((cudaStream_t)user), cs_((cudaStream_t)user)
	{
		if (user_ || !FrameGraph::enabled())
			return;
		static thread_local cudaStream_t own[64] = {nullptr};
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64)
			return;
		if (!own[dev] && cudaStreamCreateWithFlags(&own[dev], cudaStreamNonBlocking) != cudaSuccess) {
			own[dev] = nullptr;
			cudaGetLastError();
			return;
		}
		cudaEvent_t ev;
		if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
			return;
		cudaEventRecord(ev, user_);
		cudaStreamWaitEvent(own[dev], ev, 0);
		cudaEventDestroy(ev);
		cs_ = own[dev];
		forked_ = true;
	}
	cudaStream_t get() const { return cs_; }
	bool capturable() const { return cs_ != nullptr && FrameGraph::enabled(); }
	void join()
	{
		if (!forked_)
			return;
		cudaEvent_t ev;
		if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
			cudaEventRecord(ev, cs_);
			cudaStreamWaitEvent(user_, ev, 0);
			cudaEventDestroy(ev);
		}
		forked_ = false;
	}
	~WalkStream() { join(); }

private:
	cudaStream_t user_, cs_;
	bool forked_ = false;
};

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
