// chan_plan.cpp - host side of the wideband channeliser: the filter designs and the resampler's phase walk.
//
// utils/gmr1_rx_sdr.py asks GNU Radio's firdes for its filters: firdes.low_pass(1.0, mid_samp_rate, chan_width * 0.50,
// chan_width * 0.25) for the bank (:433-438, default Hamming window) and firdes.root_raised_cosine(32.0,
// 32.0 * chan_rate * 2, sym_rate, 0.35, int(11.0 * 32 * chan_rate * 2 / sym_rate)) for the per-ARFCN resampler
// (:523-529); the resampler is pfb.arb_resampler_ccf(..., flt_size = 32) (:591-596).  GNU Radio is not part of the
// reference tree: the formulas below are the published ones of gr-filter (firdes.cc, pfb_arb_resampler.cc), written
// out here for the product; tests compare them with the independent numpy restatement in oracle/chan_port.py.
#include <errno.h>
#include <math.h>

#include "chan.h"

namespace gmr1 {

static const double CHAN_WIDTH = 31250.0, SYM_RATE = 23400.0;
static const int FLT = 32;

static std::vector<float> low_pass(double gain, double fs, double fc, double tw)
{
	int ntaps = (int)(53.0 * fs / (22.0 * tw));             // Hamming: 53 dB
	if ((ntaps & 1) == 0)
		ntaps++;
	const int M = (ntaps - 1) / 2;
	const double fwT0 = 2.0 * M_PI * fc / fs;
	std::vector<float> w(ntaps), taps(ntaps);
	for (int n = 0; n < ntaps; n++)
		w[n] = (float)(0.54 - 0.46 * cos((2.0 * M_PI * n) / (ntaps - 1)));
	for (int n = -M; n <= M; n++)
		taps[n + M] = n == 0 ? (float)(fwT0 / M_PI * w[n + M]) : (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
	double fmax = taps[M];
	for (int n = 1; n <= M; n++)
		fmax += 2.0 * taps[n + M];
	const double g = gain / fmax;
	for (auto &t : taps)
		t = (float)(t * g);
	return taps;
}

static std::vector<float> root_raised_cosine(double gain, double fs, double sym_rate, double alpha, int ntaps)
{
	ntaps |= 1;
	const double spb = fs / sym_rate;
	std::vector<double> t(ntaps);
	double scale = 0.0;
	for (int i = 0; i < ntaps; i++) {
		const double xindx = i - ntaps / 2;
		const double x1 = M_PI * xindx / spb;
		double x2 = 4.0 * alpha * xindx / spb;
		double x3 = x2 * x2 - 1.0;
		double num, den;
		if (fabs(x3) >= 0.000001) {
			if (i != ntaps / 2)
				num = cos((1.0 + alpha) * x1) + sin((1.0 - alpha) * x1) / (4.0 * alpha * xindx / spb);
			else
				num = cos((1.0 + alpha) * x1) + (1.0 - alpha) * M_PI / (4.0 * alpha);
			den = x3 * M_PI;
		} else {
			x3 = (1.0 - alpha) * x1;
			x2 = (1.0 + alpha) * x1;
			num = sin(x2) * (1.0 + alpha) * M_PI - cos(x3) * ((1.0 - alpha) * M_PI * spb) / (4.0 * alpha * xindx) +
			      sin(x3) * spb * spb / (4.0 * alpha * xindx * xindx);
			den = -32.0 * M_PI * alpha * alpha * xindx / spb;
		}
		t[i] = 4.0 * alpha * num / den;
		scale += t[i];
	}
	std::vector<float> out(ntaps);
	for (int i = 0; i < ntaps; i++)
		out[i] = (float)(t[i] * gain / scale);
	return out;
}

int chan_plan_init(ChanPlan &p, int n_chans, int sps)
{
	if (n_chans < 2 || (n_chans & 1) || n_chans > 4096 || sps < 1 || sps > 16)
		return -EINVAL;
	// radix plan of the bank's FFT: 4s, then a 2, then odd primes
	int n = n_chans;
	while (n % 4 == 0) {
		p.radix.push_back(4);
		n /= 4;
	}
	if (n % 2 == 0) {
		p.radix.push_back(2);
		n /= 2;
	}
	for (int f = 3; f <= CHAN_MAX_RADIX && n > 1; f += 2)
		while (n % f == 0) {
			p.radix.push_back(f);
			n /= f;
		}
	if (n != 1 || (int)p.radix.size() > CHAN_MAX_STAGE)
		return -EINVAL;                                      // a prime factor above CHAN_MAX_RADIX
	p.n_chans = n_chans;
	p.sps = sps;
	p.samp_rate = n_chans * CHAN_WIDTH;
	p.taps = low_pass(1.0, p.samp_rate, CHAN_WIDTH * 0.50, CHAN_WIDTH * 0.25);
	p.taps_per_branch = ((int)p.taps.size() + n_chans - 1) / n_chans;
	p.mid_rate = CHAN_WIDTH * 2.0;
	p.resamp = (SYM_RATE * sps) / p.mid_rate;
	p.taps_resamp = root_raised_cosine(32.0, 32.0 * p.mid_rate, SYM_RATE, 0.35, (int)(11.0 * 32 * p.mid_rate / SYM_RATE));
	const int nt = (int)p.taps_resamp.size();
	p.tpf = (nt + FLT - 1) / FLT;
	p.filt.assign((size_t)FLT * p.tpf, 0.0f);
	p.dfilt.assign((size_t)FLT * p.tpf, 0.0f);
	for (int i = 0; i < nt; i++) {
		const float d = i + 1 < nt ? p.taps_resamp[i + 1] - p.taps_resamp[i] : 0.0f;   // create_diff_taps: [-1, 1]
		p.filt[(size_t)(i % FLT) * p.tpf + i / FLT] = p.taps_resamp[i];
		p.dfilt[(size_t)(i % FLT) * p.tpf + i / FLT] = d;
	}
	p.twiddle.resize(n_chans);
	for (int t = 0; t < n_chans; t++)
		p.twiddle[t] = make_float2((float)cos(2.0 * M_PI * t / n_chans), (float)sin(2.0 * M_PI * t / n_chans));
	p.walk_j = (nt / 2) % FLT;                               // pfb_arb_resampler: d_last_filter = (ntaps / 2) % nfilts
	// group delay of the two filters; the walk starts walk_j / 32 of a bank step into the stream
	p.delay_out = (((p.taps.size() - 1) / 2.0) / p.samp_rate + ((nt - 1) / 2.0 - p.walk_j) / (FLT * p.mid_rate)) * SYM_RATE * sps;
	p.walk_acc = 0.0f;
	p.sched_in = 0;
	return 0;
}

// pfb_arb_resampler::filter's walk: every output takes filter j at input i with weight acc; then
// acc += flt_rate, j += dec_rate + floor(acc), acc = fmodf(acc, 1), and whole multiples of 32 in j move the input on.
// The walk is resumable: `w` carries (input position, filter, weight) from one call to the next, so a streaming caller
// gets exactly the outputs a one-shot caller gets.
void chan_walk(const ChanPlan &p, ChanWalk &w, int64_t n_steps, std::vector<int32_t> &out_i, std::vector<uint8_t> &out_j,
               std::vector<float> &out_acc, int64_t i_base)
{
	const int dec_rate = (int)floor(FLT / p.resamp);
	const float flt_rate = (float)(FLT / p.resamp - dec_rate);
	int64_t i_in = w.i_in;
	int j = w.j;
	float acc = w.acc;
	while (i_in < n_steps) {
		while (j < FLT) {
			out_i.push_back((int32_t)(i_in - i_base));
			out_j.push_back((uint8_t)j);
			out_acc.push_back(acc);
			acc += flt_rate;
			j += dec_rate + (int)floorf(acc);
			acc = fmodf(acc, 1.0f);
		}
		i_in += j / FLT;
		j = j % FLT;
	}
	w.i_in = i_in;
	w.j = j;
	w.acc = acc;
}

ChanWalk chan_walk_start(const ChanPlan &p)
{
	ChanWalk w;
	w.i_in = 0;
	w.j = ((int)p.taps_resamp.size() / 2) % FLT;           // pfb_arb_resampler: d_last_filter = (ntaps / 2) % nfilts
	w.acc = 0.0f;
	return w;
}

void chan_plan_walk(ChanPlan &p, int64_t n_steps)
{
	ChanWalk w;
	w.i_in = p.sched_in; w.j = p.walk_j; w.acc = p.walk_acc;
	chan_walk(p, w, n_steps, p.sched_i, p.sched_j, p.sched_acc, 0);
	p.sched_in = w.i_in; p.walk_j = w.j; p.walk_acc = w.acc;
}

int64_t chan_plan_out_len(ChanPlan &p, int64_t n_wide)
{
	const int64_t n_steps = n_wide / (p.n_chans / 2);
	if (n_steps <= 0)
		return 0;
	chan_plan_walk(p, n_steps);
	// outputs whose newest input lies inside the recording
	int64_t lo = 0, hi = (int64_t)p.sched_i.size();
	while (lo < hi) {
		const int64_t mid = (lo + hi) / 2;
		if (p.sched_i[mid] < n_steps)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

}  // namespace gmr1
