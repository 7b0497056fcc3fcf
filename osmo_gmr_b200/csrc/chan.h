// chan.h - wideband channeliser: kernel arguments and the host-side plan (chan_plan.cpp, chan_kernels.cu, api_chan.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "launch.h"

namespace gmr1 {

constexpr int CHAN_MAX_STAGE = 16;
constexpr int CHAN_MAX_RADIX = 31;         // largest odd prime factor of n_chans the bank's FFT takes

struct PfbArgs {
	const void   *wide;                    // wideband recording: cf32 (fmt 0) or interleaved int16 I/Q (fmt 1)
	int64_t       n_wide;
	int32_t       n_chans, taps_per_branch, groups;
	const float  *taps;                    // [taps_per_branch * n_chans] prototype low-pass, zero padded
	const float2 *twiddle;                 // [n_chans] e^{+j 2 pi t / N}
	int32_t       n_stage, radix[CHAN_MAX_STAGE];
	float2       *mid;                     // [n_steps][n_chans] bank output, 2 x 31.25 kS/s per channel
	int64_t       n_steps;
	int64_t       m_begin, m_end;          // the steps this launch makes (a chunk of the recording); samples up to
	                                       // (m_end - 1) n_chans / 2 are read
};

struct ResampArgs {
	const float2  *mid;
	int64_t        n_steps;
	int32_t        n_chans;
	const int32_t *chan_idx;               // [n_wanted] bank channel of each output stream, or NULL (identity)
	int32_t        n_wanted;
	const int32_t *sched_i;                // [n_out] newest input step of output n
	const uint8_t *sched_j;                // [n_out] filter phase
	const float   *sched_acc;              // [n_out] weight of the derivative filter
	const float   *filt, *dfilt;           // [32][tpf]
	int32_t        tpf, rows_max, span_max; // taps per phase; input rows of a tile / of an 8-output group, at most
	int32_t        tile_out;               // outputs per tile: resamp_tile_outputs()
	float2        *out;                    // [n_wanted][out_stride]
	int64_t        out_stride, n_out;
	int64_t        n_begin, n_end;         // the outputs this launch makes (n_begin a multiple of tile_out)
};

struct WideSynthArgs {
	const float2  *streams;                // [n_streams][stream_stride] per-ARFCN streams at sps x 23.4 kS/s
	int64_t        stream_stride, stream_len;
	const int32_t *chan_idx;               // [n_streams] or NULL (identity)
	int32_t        n_streams, n_chans;
	int64_t        num, den;               // stream samples per wideband sample = num / den
	const float2  *twiddle;
	float          sigma, gain;
	uint64_t       seed;
	void          *wide;
	int64_t        n_wide;
};

int pfb_groups(int n_chans);
cudaError_t launch_pfb(const PfbArgs &a, int fmt, cudaStream_t st);
// outputs per resampler tile (64, or 32 when 64 outputs span too many input rows for two CTAs per SM)
int resamp_tile_outputs(int rows64, int span_max, int tpf);
int resamp_group_outputs();
int pfb_is_fast(int n_chans, int taps_per_branch);
void pfb_force_generic(int on);          // tests: run the generic kernel where the fast one would
cudaError_t launch_resamp(const ResampArgs &a, cudaStream_t st);
cudaError_t launch_wide_synth(const WideSynthArgs &a, int fmt, cudaStream_t st);

// ---- host-side plan: what PFBBase / PFBOutputParameters of utils/gmr1_rx_sdr.py (:393-447, :501-529) compute
struct ChanPlan {
	int    n_chans = 0, sps = 0;
	double samp_rate = 0, mid_rate = 0, resamp = 0, delay_out = 0;
	std::vector<float> taps;               // firdes.low_pass(1, samp_rate, 15 625, 7 812.5), Hamming
	std::vector<float> taps_resamp;        // firdes.root_raised_cosine(32, 32 * 62 500, 23 400, 0.35, 11 symbols)
	int    taps_per_branch = 0, tpf = 0;
	std::vector<float> filt, dfilt;        // [32][tpf]
	std::vector<int>   radix;
	std::vector<float2> twiddle;
	// phase walk of the resampler for the first n outputs (grown on demand)
	std::vector<int32_t> sched_i;
	std::vector<uint8_t> sched_j;
	std::vector<float>   sched_acc;
	int64_t sched_in = 0;                  // input steps the walk has consumed
	int     walk_j = 0;
	float   walk_acc = 0.0f;
	// device copies, one set per device ordinal (created on first use under the init lock)
	struct Dev {
		float *taps = nullptr, *filt = nullptr, *dfilt = nullptr, *sched_acc = nullptr;
		float2 *twiddle = nullptr;
		int32_t *sched_i = nullptr;
		uint8_t *sched_j = nullptr;
		size_t sched_n = 0;
	} dev[64];
};

struct ChanWalk { int64_t i_in; int j; float acc; };            // the resampler's phase walk between two outputs
ChanWalk chan_walk_start(const ChanPlan &p);
// outputs whose newest input step lies below n_steps, appended to the three vectors (input steps relative to i_base)
void chan_walk(const ChanPlan &p, ChanWalk &w, int64_t n_steps, std::vector<int32_t> &out_i, std::vector<uint8_t> &out_j,
               std::vector<float> &out_acc, int64_t i_base);

int  chan_plan_init(ChanPlan &p, int n_chans, int sps);         // 0 / -EINVAL
void chan_plan_walk(ChanPlan &p, int64_t n_steps);              // extend the phase walk to cover n_steps input steps
int64_t chan_plan_out_len(ChanPlan &p, int64_t n_wide);         // outputs per channel for n_wide wideband samples

}  // namespace gmr1
