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
	if (!pl || !wide || !out || n_wide < 0 || n_wanted < 0 || iq_format < 0 || iq_format > 2)
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
	const size_t samp_bytes = iq_format == 0 ? sizeof(float2) : iq_format == 2 ? 2 * sizeof(int8_t) : 2 * sizeof(int16_t);
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

// ---- streaming: consecutive blocks of one endless recording ------------------------------------------------------------
namespace {

struct ChanStream {
	Plan *pl = nullptr;
	int dev = 0, n_wanted = 0, fmt = -1;
	int32_t *d_idx = nullptr;              // [n_wanted] or NULL (identity)
	ChanWalk walk;
	int64_t s_total = 0, s_base = 0;       // samples received / first sample kept in `wide`
	int64_t m_done = 0, r_base = 0;        // bank steps made / first bank row kept in `mid`
	int64_t n_done = 0;                    // outputs delivered per channel
	char *wide[2] = {nullptr, nullptr};    // [cap_wide] bytes, ping-pong: history + the new block
	float2 *mid[2] = {nullptr, nullptr};   // [cap_rows][n_chans]
	size_t cap_wide = 0, cap_rows = 0;
	int cur_w = 0, cur_m = 0;
	int32_t *d_si = nullptr; uint8_t *d_sj = nullptr; float *d_sa = nullptr; size_t cap_sched = 0;
	int32_t *h_si = nullptr; uint8_t *h_sj = nullptr; float *h_sa = nullptr; size_t cap_hsched = 0;   // page-locked
	cudaEvent_t h_free = nullptr;          // the page-locked walk of the previous block has been read
	std::vector<int32_t> vi; std::vector<uint8_t> vj; std::vector<float> va;
};

void stream_free(ChanStream *cs)
{
	if (!cs)
		return;
	int cur = 0;
	cudaGetDevice(&cur);
	cudaSetDevice(cs->dev);
	cudaDeviceSynchronize();
	for (int k = 0; k < 2; k++) {
		cudaFree(cs->wide[k]);
		cudaFree(cs->mid[k]);
	}
	cudaFree(cs->d_idx); cudaFree(cs->d_si); cudaFree(cs->d_sj); cudaFree(cs->d_sa);
	cudaFreeHost(cs->h_si); cudaFreeHost(cs->h_sj); cudaFreeHost(cs->h_sa);
	if (cs->h_free)
		cudaEventDestroy(cs->h_free);
	cudaSetDevice(cur);
	delete cs;
}

}  // namespace

int gmr1b200_chan_stream_create(void *plan, const int32_t *chan_idx, int n_wanted, void **state)
{
	Plan *pl = (Plan *)plan;
	if (!pl || !state || n_wanted < 1 || (!chan_idx && n_wanted > pl->p.n_chans))
		return set_err(-EINVAL, "chan_stream_create: bad argument");
	std::vector<int32_t> idx;
	if (chan_idx) {
		idx.resize(n_wanted);
		cudaError_t e = cudaMemcpy(idx.data(), chan_idx, sizeof(int32_t) * n_wanted, cudaMemcpyDefault);   // host or device list
		if (e != cudaSuccess)
			return cuda_rc(e, "chan_stream_create: channel list");
		for (int i = 0; i < n_wanted; i++)
			if (idx[i] < 0 || idx[i] >= pl->p.n_chans)
				return set_err(-EINVAL, "chan_stream_create: channel index outside the bank");
	}
	ChanStream *cs = new (std::nothrow) ChanStream;
	if (!cs)
		return set_err(-ENOMEM, "chan_stream_create: out of memory");
	cs->pl = pl;
	cs->n_wanted = n_wanted;
	cs->walk = chan_walk_start(pl->p);
	cudaError_t e = cudaGetDevice(&cs->dev);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&cs->h_free, cudaEventDisableTiming);
	if (e == cudaSuccess && chan_idx) {
		e = cudaMalloc((void **)&cs->d_idx, sizeof(int32_t) * n_wanted);
		if (e == cudaSuccess)
			e = cudaMemcpy(cs->d_idx, idx.data(), sizeof(int32_t) * n_wanted, cudaMemcpyHostToDevice);
	}
	if (e != cudaSuccess) {
		stream_free(cs);
		return cuda_rc(e, "chan_stream_create");
	}
	*state = cs;
	return 0;
}

void gmr1b200_chan_stream_destroy(void *state) { stream_free((ChanStream *)state); }

int64_t gmr1b200_chan_stream_max_out(void *state, int64_t n_wide)
{
	ChanStream *cs = (ChanStream *)state;
	if (!cs || n_wide < 0)
		return set_err(-EINVAL, "chan_stream_max_out: bad argument");
	const ChanPlan &p = cs->pl->p;
	const int64_t steps = n_wide / (p.n_chans / 2) + 2;     // the new block plus what was left over from the blocks before
	return (int64_t)ceil((double)steps * p.resamp) + 2;
}

int gmr1b200_chan_stream_push(void *state, const void *wide, int iq_format, int64_t n_wide, float *out, int64_t out_stride,
                              int64_t *n_out, void *stream)
{
	ChanStream *cs = (ChanStream *)state;
	if (!cs || !n_out || n_wide < 0 || (n_wide && !wide) || iq_format < 0 || iq_format > 2)
		return set_err(-EINVAL, "chan_stream_push: bad argument");
	if (cs->fmt >= 0 && cs->fmt != iq_format)
		return set_err(-EINVAL, "chan_stream_push: the sample format of a stream cannot change");
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev != cs->dev)
		return set_err(-EINVAL, "chan_stream_push: the stream lives on another device");
	*n_out = 0;
	if (n_wide == 0)
		return 0;
	cs->fmt = iq_format;
	ChanPlan &p = cs->pl->p;
	const int N = p.n_chans, D = N / 2, P = p.taps_per_branch;
	const size_t sb = iq_format == 0 ? sizeof(float2) : iq_format == 2 ? 2 * sizeof(int8_t) : 2 * sizeof(int16_t);
	cudaStream_t st = (cudaStream_t)stream;
	ChanPlan::Dev *d = nullptr;
	{
		std::lock_guard<std::mutex> lk(cs->pl->mu);
		cudaError_t e = plan_device(*cs->pl, 0, &d);
		if (e != cudaSuccess)
			return cuda_rc(e, "chan_stream_push: table upload");
	}
	// what the next bank step (m_done) and the next output (the walk's input position) still need of the past
	const int64_t s_keep0 = (cs->m_done - 2 * (int64_t)(P - 1)) * D - (N - 1);
	const int64_t s_keep = s_keep0 < cs->s_base ? cs->s_base : s_keep0;
	const int64_t r_keep0 = cs->walk.i_in - (p.tpf - 1);
	const int64_t r_keep = r_keep0 < cs->r_base ? cs->r_base : r_keep0;
	const int64_t s_new = cs->s_total + n_wide;
	const int64_t m_avail = s_new / D;
	// the outputs of this block must fit: checked before anything changes (the walk has not passed step m_done)
	if (m_avail > cs->m_done) {
		const int64_t bound = (int64_t)ceil((double)(m_avail - cs->m_done + 1) * p.resamp) + 2;
		if (!out || out_stride < bound)
			return set_err(-EINVAL, "chan_stream_push: out NULL or out_stride smaller than gmr1b200_chan_stream_max_out(n_wide)");
	}
	cudaError_t e = cudaSuccess;
	// ---- samples: history + new block into the other buffer
	{
		const size_t need = (size_t)(s_new - s_keep) * sb;
		const int nxt = cs->cur_w ^ 1;
		if (need > cs->cap_wide) {                              // grow both (the old one is still read below)
			const size_t cap = need + need / 2 + 4096;
			char *nw[2] = {nullptr, nullptr};
			if ((e = cudaMalloc((void **)&nw[0], cap)) != cudaSuccess || (e = cudaMalloc((void **)&nw[1], cap)) != cudaSuccess) {
				cudaFree(nw[0]);
				return cuda_rc(e, "chan_stream_push: sample buffer");
			}
			if (cs->s_total > s_keep)
				e = cudaMemcpyAsync(nw[nxt], cs->wide[cs->cur_w] + (size_t)(s_keep - cs->s_base) * sb,
				                    (size_t)(cs->s_total - s_keep) * sb, cudaMemcpyDeviceToDevice, st);
			cudaStreamSynchronize(st);
			cudaFree(cs->wide[0]);
			cudaFree(cs->wide[1]);
			cs->wide[0] = nw[0]; cs->wide[1] = nw[1]; cs->cap_wide = cap;
		} else if (cs->s_total > s_keep) {
			e = cudaMemcpyAsync(cs->wide[nxt], cs->wide[cs->cur_w] + (size_t)(s_keep - cs->s_base) * sb,
			                    (size_t)(cs->s_total - s_keep) * sb, cudaMemcpyDeviceToDevice, st);
		}
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(cs->wide[nxt] + (size_t)(cs->s_total - s_keep) * sb, wide, (size_t)n_wide * sb, cudaMemcpyDefault, st);
		if (e != cudaSuccess)
			return cuda_rc(e, "chan_stream_push: sample copy");
		cs->cur_w = nxt;
		cs->s_base = s_keep;
		cs->s_total = s_new;
	}
	if (m_avail <= cs->m_done)
		return 0;                                               // not a whole bank step yet
	// ---- bank rows: history + new rows into the other buffer
	{
		const size_t need = (size_t)(m_avail - r_keep);
		const int nxt = cs->cur_m ^ 1;
		const size_t row_b = (size_t)N * sizeof(float2);
		if (need > cs->cap_rows) {
			const size_t cap = need + need / 2 + 64;
			float2 *nm[2] = {nullptr, nullptr};
			if ((e = cudaMalloc((void **)&nm[0], cap * row_b)) != cudaSuccess || (e = cudaMalloc((void **)&nm[1], cap * row_b)) != cudaSuccess) {
				cudaFree(nm[0]);
				return cuda_rc(e, "chan_stream_push: bank buffer");
			}
			if (cs->m_done > r_keep)
				e = cudaMemcpyAsync(nm[nxt], cs->mid[cs->cur_m] + (size_t)(r_keep - cs->r_base) * N, (size_t)(cs->m_done - r_keep) * row_b,
				                    cudaMemcpyDeviceToDevice, st);
			cudaStreamSynchronize(st);
			cudaFree(cs->mid[0]);
			cudaFree(cs->mid[1]);
			cs->mid[0] = nm[0]; cs->mid[1] = nm[1]; cs->cap_rows = cap;
		} else if (cs->m_done > r_keep) {
			e = cudaMemcpyAsync(cs->mid[nxt], cs->mid[cs->cur_m] + (size_t)(r_keep - cs->r_base) * N, (size_t)(cs->m_done - r_keep) * row_b,
			                    cudaMemcpyDeviceToDevice, st);
		}
		if (e != cudaSuccess)
			return cuda_rc(e, "chan_stream_push: bank history");
		cs->cur_m = nxt;
		cs->r_base = r_keep;
	}
	PfbArgs pa = {};
	pa.wide = cs->wide[cs->cur_w] - (ptrdiff_t)(cs->s_base * (int64_t)sb);      // absolute sample index -> buffer
	pa.n_wide = cs->s_total; pa.n_chans = N; pa.taps_per_branch = P;
	pa.taps = d->taps; pa.twiddle = d->twiddle; pa.n_stage = (int)p.radix.size();
	for (int i = 0; i < pa.n_stage; i++)
		pa.radix[i] = p.radix[i];
	pa.mid = cs->mid[cs->cur_m] - (ptrdiff_t)(cs->r_base * N);                  // absolute step -> buffer row
	pa.n_steps = m_avail; pa.m_begin = cs->m_done; pa.m_end = m_avail;
	if ((e = launch_pfb(pa, iq_format, st)) != cudaSuccess)
		return cuda_rc(e, "chan_stream_push: bank kernel");
	g_launches.fetch_add(1);
	cs->m_done = m_avail;
	// ---- outputs whose newest input step exists now
	cs->vi.clear(); cs->vj.clear(); cs->va.clear();
	chan_walk(p, cs->walk, m_avail, cs->vi, cs->vj, cs->va, cs->r_base);
	const size_t nn = cs->vi.size();
	if (nn == 0)
		return 0;
	if (nn > cs->cap_sched) {
		cudaStreamSynchronize(st);
		cudaFree(cs->d_si); cudaFree(cs->d_sj); cudaFree(cs->d_sa);
		cs->d_si = nullptr; cs->d_sj = nullptr; cs->d_sa = nullptr;
		const size_t cap = nn + nn / 2 + 256;
		if ((e = cudaMalloc((void **)&cs->d_si, cap * 4)) != cudaSuccess || (e = cudaMalloc((void **)&cs->d_sj, cap)) != cudaSuccess ||
		    (e = cudaMalloc((void **)&cs->d_sa, cap * 4)) != cudaSuccess)
			return cuda_rc(e, "chan_stream_push: walk buffers");
		cs->cap_sched = cap;
	}
	cudaEventSynchronize(cs->h_free);                           // the previous block's walk has left the page-locked arrays
	if (nn > cs->cap_hsched) {
		cudaFreeHost(cs->h_si); cudaFreeHost(cs->h_sj); cudaFreeHost(cs->h_sa);
		cs->h_si = nullptr; cs->h_sj = nullptr; cs->h_sa = nullptr;
		const size_t cap = nn + nn / 2 + 256;
		if ((e = cudaMallocHost((void **)&cs->h_si, cap * 4)) != cudaSuccess || (e = cudaMallocHost((void **)&cs->h_sj, cap)) != cudaSuccess ||
		    (e = cudaMallocHost((void **)&cs->h_sa, cap * 4)) != cudaSuccess)
			return cuda_rc(e, "chan_stream_push: walk staging");
		cs->cap_hsched = cap;
	}
	memcpy(cs->h_si, cs->vi.data(), nn * 4);
	memcpy(cs->h_sj, cs->vj.data(), nn);
	memcpy(cs->h_sa, cs->va.data(), nn * 4);
	cudaMemcpyAsync(cs->d_si, cs->h_si, nn * 4, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(cs->d_sj, cs->h_sj, nn, cudaMemcpyHostToDevice, st);
	cudaMemcpyAsync(cs->d_sa, cs->h_sa, nn * 4, cudaMemcpyHostToDevice, st);
	cudaEventRecord(cs->h_free, st);
	const int go = resamp_group_outputs();
	auto span_of = [&](int step) {
		int m = 0;
		for (size_t n0 = 0; n0 < nn; n0 += step) {
			const size_t n1 = (n0 + step < nn ? n0 + step : nn) - 1;
			const int rows = cs->vi[n1] - cs->vi[n0] + p.tpf;
			m = rows > m ? rows : m;
		}
		return m;
	};
	Stage s(stream);
	ResampArgs ra = {};
	ra.mid = cs->mid[cs->cur_m]; ra.n_steps = m_avail - cs->r_base; ra.n_chans = N; ra.chan_idx = cs->d_idx; ra.n_wanted = cs->n_wanted;
	ra.sched_i = cs->d_si; ra.sched_j = cs->d_sj; ra.sched_acc = cs->d_sa; ra.filt = d->filt; ra.dfilt = d->dfilt;
	ra.tpf = p.tpf; ra.span_max = span_of(go);
	ra.tile_out = resamp_tile_outputs(span_of(64), ra.span_max, p.tpf);
	ra.rows_max = span_of(ra.tile_out);
	ra.out = (float2 *)s.out(out, (size_t)cs->n_wanted * (size_t)out_stride * 2);
	ra.out_stride = out_stride; ra.n_out = (int64_t)nn; ra.n_begin = 0; ra.n_end = (int64_t)nn;
	if (s.failed())
		return s.finish(cudaSuccess, "chan_stream_push: staging");
	if ((void *)ra.out != (void *)out)                          // staged host output: the row tails travel back too
		cudaMemsetAsync(ra.out, 0, (size_t)cs->n_wanted * (size_t)out_stride * sizeof(float2), st);
	e = launch_resamp(ra, st);
	if (e == cudaSuccess)
		g_launches.fetch_add(1);
	cs->n_done += (int64_t)nn;
	*n_out = (int64_t)nn;
	return s.finish(e, "chan_stream_push kernels");
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
	    iq_format < 0 || iq_format > 2)
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
	a.wide = iq_format == 0 ? (void *)s.out((float *)wide, (size_t)n_wide * 2) :
	         iq_format == 2 ? (void *)s.out((int8_t *)wide, (size_t)n_wide * 2) : (void *)s.out((int16_t *)wide, (size_t)n_wide * 2);
	a.n_wide = n_wide;
	if (s.failed())
		return s.finish(cudaSuccess, "synth_wideband: staging");
	cudaError_t e = launch_wide_synth(a, iq_format, (cudaStream_t)stream);
	if (e == cudaSuccess)
		g_launches.fetch_add(1);
	return s.finish(e, "synth_wideband kernel");
}

}  // extern "C"
