// api_chan.cu - C ABI of the wideband channeliser (include/gmr1_b200.h, "wideband channeliser")
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "chan.h"

#include <math.h>
#include <mutex>
#include <new>

using namespace gmr1;

namespace {

constexpr int N_EV = 32;

struct Plan {
	ChanPlan p;
	std::mutex mu;                         // the phase walk and the device copies grow under it
	// host recordings travel in chunks on a copy stream of the plan while the kernels of the previous chunks run
	struct Feed {
		cudaStream_t st = nullptr;
		cudaEvent_t  ev[N_EV] = {};
		int          next = 0;
	} feed[64];
};

cudaError_t plan_feed(Plan &pl, Plan::Feed **out)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev < 0 || dev >= 64)
		return cudaErrorInvalidDevice;
	Plan::Feed &f = pl.feed[dev];
	if (!f.st) {
		if ((e = cudaStreamCreateWithFlags(&f.st, cudaStreamNonBlocking)) != cudaSuccess)
			return e;
		for (int i = 0; i < N_EV; i++)
			if ((e = cudaEventCreateWithFlags(&f.ev[i], cudaEventDisableTiming)) != cudaSuccess)
				return e;
	}
	*out = &f;
	return cudaSuccess;
}

// device copies of the plan's tables on the current device; the phase walk covers n_out outputs
cudaError_t plan_device(Plan &pl, size_t n_out, ChanPlan::Dev **out)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev < 0 || dev >= 64)
		return cudaErrorInvalidDevice;
	ChanPlan &p = pl.p;
	ChanPlan::Dev &d = p.dev[dev];
	auto up = [&](auto **dst, const auto &v) -> cudaError_t {
		if (*dst)
			return cudaSuccess;
		cudaError_t e2 = cudaMalloc((void **)dst, v.size() * sizeof(v[0]));
		if (e2 != cudaSuccess)
			return e2;
		return cudaMemcpy(*dst, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice);
	};
	if (!d.taps) {
		std::vector<float> padded((size_t)p.taps_per_branch * p.n_chans, 0.0f);
		std::copy(p.taps.begin(), p.taps.end(), padded.begin());
		if ((e = up(&d.taps, padded)) != cudaSuccess)
			return e;
	}
	if ((e = up(&d.filt, p.filt)) != cudaSuccess || (e = up(&d.dfilt, p.dfilt)) != cudaSuccess ||
	    (e = up(&d.twiddle, p.twiddle)) != cudaSuccess)
		return e;
	if (d.sched_n < n_out) {               // (re)upload the walk, with headroom so that a stream of equal calls uploads once
		const size_t n = p.sched_i.size();
		cudaFree(d.sched_i);
		cudaFree(d.sched_j);
		cudaFree(d.sched_acc);
		d.sched_i = nullptr; d.sched_j = nullptr; d.sched_acc = nullptr; d.sched_n = 0;
		if ((e = up(&d.sched_i, p.sched_i)) != cudaSuccess || (e = up(&d.sched_j, p.sched_j)) != cudaSuccess ||
		    (e = up(&d.sched_acc, p.sched_acc)) != cudaSuccess)
			return e;
		d.sched_n = n;
	}
	*out = &d;
	return cudaSuccess;
}

}  // namespace

extern "C" {

int gmr1b200_chan_create(int n_chans, int sps, void **plan)
{
	if (!plan)
		return set_err(-EINVAL, "chan_create: plan NULL");
	Plan *pl = new (std::nothrow) Plan;
	if (!pl)
		return set_err(-ENOMEM, "chan_create: out of memory");
	if (chan_plan_init(pl->p, n_chans, sps)) {
		delete pl;
		return set_err(-EINVAL, "chan_create: n_chans must be even, 2..4096, with prime factors <= 31; sps 1..16");
	}
	*plan = pl;
	return 0;
}

void gmr1b200_chan_destroy(void *plan)
{
	Plan *pl = (Plan *)plan;
	if (!pl)
		return;
	int cur = 0;
	cudaGetDevice(&cur);
	for (int dev = 0; dev < 64; dev++) {
		ChanPlan::Dev &d = pl->p.dev[dev];
		if (!d.taps && !d.filt && !d.sched_i && !pl->feed[dev].st)
			continue;
		cudaSetDevice(dev);
		if (pl->feed[dev].st) {
			cudaStreamSynchronize(pl->feed[dev].st);
			for (int i = 0; i < N_EV; i++)
				cudaEventDestroy(pl->feed[dev].ev[i]);
			cudaStreamDestroy(pl->feed[dev].st);
		}
		cudaFree(d.taps); cudaFree(d.filt); cudaFree(d.dfilt); cudaFree(d.twiddle);
		cudaFree(d.sched_i); cudaFree(d.sched_j); cudaFree(d.sched_acc);
	}
	cudaSetDevice(cur);
	delete pl;
}

int gmr1b200_chan_info(void *plan, struct gmr1b200_chan_info *info)
{
	Plan *pl = (Plan *)plan;
	if (!pl || !info)
		return set_err(-EINVAL, "chan_info: NULL argument");
	const ChanPlan &p = pl->p;
	info->n_chans = p.n_chans; info->sps = p.sps; info->n_taps = (int)p.taps.size(); info->taps_per_branch = p.taps_per_branch;
	info->n_taps_resamp = (int)p.taps_resamp.size(); info->fft_stages = (int)p.radix.size();
	info->samp_rate = p.samp_rate; info->mid_rate = p.mid_rate; info->resamp = p.resamp; info->delay_out = p.delay_out;
	return 0;
}

int gmr1b200_chan_taps(void *plan, float *taps, int max_taps, float *taps_resamp, int max_resamp)
{
	Plan *pl = (Plan *)plan;
	if (!pl)
		return set_err(-EINVAL, "chan_taps: plan NULL");
	const ChanPlan &p = pl->p;
	if ((taps && max_taps < (int)p.taps.size()) || (taps_resamp && max_resamp < (int)p.taps_resamp.size()))
		return set_err(-EINVAL, "chan_taps: buffer too small");
	if (taps)
		memcpy(taps, p.taps.data(), p.taps.size() * sizeof(float));
	if (taps_resamp)
		memcpy(taps_resamp, p.taps_resamp.data(), p.taps_resamp.size() * sizeof(float));
	return 0;
}

int64_t gmr1b200_chan_out_len(void *plan, int64_t n_wide)
{
	Plan *pl = (Plan *)plan;
	if (!pl || n_wide < 0)
		return set_err(-EINVAL, "chan_out_len: bad argument");
	std::lock_guard<std::mutex> lk(pl->mu);
	return chan_plan_out_len(pl->p, n_wide);
}

int gmr1b200_channelize(void *plan, const void *wide, int iq_format, int64_t n_wide, const int32_t *chan_idx, int n_wanted,
                        float *out, int64_t out_stride, void *stream)
{
	Plan *pl = (Plan *)plan;
	if (!pl || !wide || !out || n_wide < 0 || n_wanted < 0 || iq_format < 0 || iq_format > 1)
		return set_err(-EINVAL, "channelize: bad argument");
	ChanPlan &p = pl->p;
	if (n_wanted == 0 || n_wide == 0)
		return 0;
	if (chan_idx && host_pointer(chan_idx))
		for (int i = 0; i < n_wanted; i++)
			if (chan_idx[i] < 0 || chan_idx[i] >= p.n_chans)
				return set_err(-EINVAL, "channelize: channel index outside the bank");
	if (!chan_idx && n_wanted > p.n_chans)
		return set_err(-EINVAL, "channelize: more streams than channels");
	int64_t n_out, n_steps;
	ChanPlan::Dev *d = nullptr;
	Plan::Feed *feed = nullptr;
	int rows_max = 0, span_max = 0, to = 64;
	const bool wide_on_host = host_pointer(wide);
	const size_t samp_bytes = iq_format == 0 ? sizeof(float2) : 2 * sizeof(int16_t);
	// chunks: a host recording travels in up to 16 pieces (>= 2 MB each) so that the bank and the resampler of piece c
	// run under the copy of piece c + 1; a device-resident recording is one piece
	std::vector<int64_t> cut_m, cut_n;     // piece c makes steps [cut_m[c], cut_m[c+1]) and outputs [cut_n[c], cut_n[c+1])
	{
		std::lock_guard<std::mutex> lk(pl->mu);
		n_out = chan_plan_out_len(p, n_wide);
		n_steps = n_wide / (p.n_chans / 2);
		if (n_out > out_stride)
			return set_err(-EINVAL, "channelize: out_stride smaller than gmr1b200_chan_out_len(n_wide)");
		if (n_out == 0)
			return 0;
		cudaError_t e = plan_device(*pl, (size_t)n_out, &d);
		if (e == cudaSuccess && wide_on_host)
			e = plan_feed(*pl, &feed);
		if (e != cudaSuccess)
			return cuda_rc(e, "channelize: table upload");
		const int go = resamp_group_outputs();
		auto span_of = [&](int step) {
			int m = 0;
			for (int64_t n0 = 0; n0 < n_out; n0 += step) {
				const int64_t n1 = (n0 + step < n_out ? n0 + step : n_out) - 1;
				const int rows = p.sched_i[n1] - p.sched_i[n0] + p.tpf;
				m = rows > m ? rows : m;
			}
			return m;
		};
		span_max = span_of(go);
		to = resamp_tile_outputs(span_of(64), span_max, p.tpf);
		rows_max = span_of(to);
		int pieces = 1;
		if (wide_on_host) {
			pieces = (int)((size_t)n_wide * samp_bytes / (2u << 20));
			pieces = pieces < 1 ? 1 : pieces > 16 ? 16 : pieces;
		}
		cut_m.push_back(0);
		cut_n.push_back(0);
		for (int c = 1; c <= pieces; c++) {
			int64_t m1 = c == pieces ? n_steps : (n_steps * c / pieces) & ~(int64_t)63;
			if (m1 <= cut_m.back() && c < pieces)
				continue;
			// outputs whose newest input step is below m1, down to a whole tile (the rest waits for the next piece)
			int64_t lo = cut_n.back(), hi = n_out;
			while (lo < hi) {
				const int64_t mid = (lo + hi) / 2;
				if (p.sched_i[mid] < m1)
					lo = mid + 1;
				else
					hi = mid;
			}
			const int64_t n1 = c == pieces ? n_out : lo - lo % to;
			cut_m.push_back(m1);
			cut_n.push_back(n1 < cut_n.back() ? cut_n.back() : n1);
		}
	}
	cudaStream_t st = (cudaStream_t)stream;
	Stage s(stream);
	const void *d_wide;
	if (wide_on_host)
		d_wide = s.tmp<char>((size_t)n_wide * samp_bytes);
	else
		d_wide = wide;
	const int32_t *d_idx = s.in(chan_idx, (size_t)n_wanted);
	float2 *d_out = (float2 *)s.out(out, (size_t)n_wanted * (size_t)out_stride * 2);
	float2 *mid = s.tmp<float2>((size_t)n_steps * p.n_chans);
	if (s.failed())
		return s.finish(cudaSuccess, "channelize: staging");
	if ((void *)d_out != (void *)out && out_stride > n_out)      // staged host output: the row tails travel back too
		cudaMemsetAsync(d_out, 0, (size_t)n_wanted * (size_t)out_stride * sizeof(float2), st);
	PfbArgs pa = {};
	pa.wide = d_wide; pa.n_wide = n_wide; pa.n_chans = p.n_chans; pa.taps_per_branch = p.taps_per_branch;
	pa.taps = d->taps; pa.twiddle = d->twiddle; pa.n_stage = (int)p.radix.size();
	for (int i = 0; i < pa.n_stage; i++)
		pa.radix[i] = p.radix[i];
	pa.mid = mid; pa.n_steps = n_steps;
	ResampArgs ra = {};
	ra.mid = mid; ra.n_steps = n_steps; ra.n_chans = p.n_chans; ra.chan_idx = d_idx; ra.n_wanted = n_wanted;
	ra.sched_i = d->sched_i; ra.sched_j = d->sched_j; ra.sched_acc = d->sched_acc; ra.filt = d->filt; ra.dfilt = d->dfilt;
	ra.tpf = p.tpf; ra.rows_max = rows_max; ra.span_max = span_max; ra.tile_out = to; ra.out = d_out; ra.out_stride = out_stride; ra.n_out = n_out;
	cudaError_t e = cudaSuccess;
	// the plan's copy stream and event ring serve one call at a time: host-fed calls on one plan queue up here (they
	// only enqueue; device-resident calls do not take the lock)
	std::unique_lock<std::mutex> feed_lock(pl->mu, std::defer_lock);
	if (wide_on_host)
		feed_lock.lock();
	if (wide_on_host) {                    // the copy stream starts behind the allocation of its target
		cudaEvent_t ev = feed->ev[feed->next++ % N_EV];
		if ((e = cudaEventRecord(ev, st)) == cudaSuccess)
			e = cudaStreamWaitEvent(feed->st, ev, 0);
	}
	int64_t copied = 0;                    // wideband samples on the device so far
	for (size_t c = 0; c + 1 < cut_m.size() && e == cudaSuccess; c++) {
		const bool last = c + 2 == cut_m.size();
		if (wide_on_host) {
			// steps below cut_m[c+1] read samples up to (cut_m[c+1] - 1) n_chans / 2; the last piece takes the rest
			int64_t upto = last ? n_wide : (cut_m[c + 1] - 1) * (p.n_chans / 2) + 1;
			upto = upto > n_wide ? n_wide : upto;
			if (upto > copied) {
				e = cudaMemcpyAsync((char *)d_wide + (size_t)copied * samp_bytes, (const char *)wide + (size_t)copied * samp_bytes,
				                    (size_t)(upto - copied) * samp_bytes, cudaMemcpyHostToDevice, feed->st);
				copied = upto;
				cudaEvent_t ev = feed->ev[feed->next++ % N_EV];
				if (e == cudaSuccess)
					e = cudaEventRecord(ev, feed->st);
				if (e == cudaSuccess)
					e = cudaStreamWaitEvent(st, ev, 0);
				if (e != cudaSuccess)
					break;
			}
		}
		pa.m_begin = cut_m[c]; pa.m_end = cut_m[c + 1];
		if (pa.m_end > pa.m_begin) {
			if ((e = launch_pfb(pa, iq_format, st)) != cudaSuccess)
				break;
			g_launches.fetch_add(1);
		}
		ra.n_begin = cut_n[c]; ra.n_end = cut_n[c + 1];
		if (ra.n_end > ra.n_begin) {
			if ((e = launch_resamp(ra, st)) != cudaSuccess)
				break;
			g_launches.fetch_add(1);
		}
	}
	return s.finish(e, "channelize kernels");
}

int gmr1b200_set_chan_generic(int on)
{
	static std::atomic<int> cur{0};
	const int prev = cur.exchange(on ? 1 : 0);
	pfb_force_generic(on ? 1 : 0);
	return prev;
}

int gmr1b200_synth_wideband(void *plan, const float *streams, int64_t stream_stride, int64_t stream_len,
                            const int32_t *chan_idx, int n_streams, float esn0_db, float gain, uint64_t seed,
                            void *wide, int iq_format, int64_t n_wide, void *stream)
{
	Plan *pl = (Plan *)plan;
	if (!pl || !streams || !wide || n_streams < 1 || stream_len < 4 || stream_stride < stream_len || n_wide < 0 ||
	    iq_format < 0 || iq_format > 1)
		return set_err(-EINVAL, "synth_wideband: bad argument");
	ChanPlan &p = pl->p;
	ChanPlan::Dev *d = nullptr;
	{
		std::lock_guard<std::mutex> lk(pl->mu);
		cudaError_t e = plan_device(*pl, 0, &d);
		if (e != cudaSuccess)
			return cuda_rc(e, "synth_wideband: table upload");
	}
	Stage s(stream);
	WideSynthArgs a = {};
	a.streams = (const float2 *)s.in(streams, (size_t)n_streams * (size_t)stream_stride * 2);
	a.stream_stride = stream_stride; a.stream_len = stream_len;
	a.chan_idx = s.in(chan_idx, (size_t)n_streams);
	a.n_streams = n_streams; a.n_chans = p.n_chans;
	a.num = 468 * (int64_t)p.sps;          // stream samples per wideband sample: (23 400 sps) / (31 250 N) = 468 sps / (625 N)
	a.den = 625 * (int64_t)p.n_chans;
	a.twiddle = d->twiddle;
	a.sigma = esn0_db >= 100.0f ? 0.0f : sqrtf(exp10f(-esn0_db / 10.0f) * (float)(p.samp_rate / 23400.0) * 0.5f);
	a.gain = gain; a.seed = seed;
	a.wide = iq_format == 0 ? (void *)s.out((float *)wide, (size_t)n_wide * 2) : (void *)s.out((int16_t *)wide, (size_t)n_wide * 2);
	a.n_wide = n_wide;
	if (s.failed())
		return s.finish(cudaSuccess, "synth_wideband: staging");
	cudaError_t e = launch_wide_synth(a, iq_format, (cudaStream_t)stream);
	if (e == cudaSuccess)
		g_launches.fetch_add(1);
	return s.finish(e, "synth_wideband kernel");
}

}  // extern "C"
