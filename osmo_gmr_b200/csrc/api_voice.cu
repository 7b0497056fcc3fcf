// api_voice.cu - SURVEY 8f N4, the AMBE hand-off: the speech frames of a call as the stream the reference's vocoder
// front end reads.  gmr1_tch3_decode (src/l1/tch3.c:116-184) returns two 10-byte frames per TCH3 burst;
// src/gmr1_ambe_decode.c:127-150 reads such frames 10 bytes at a time and calls gmr1_codec_decode_frame(codec, audio,
// 160, frame, bad) (include/osmocom/gmr1/codec/codec.h:40-45), gmr1_codec_decode_dtx for a 20 ms slot without one.
// The vocoder itself is out of scope; this is the format it is fed, built on the device from the
// per-frame records of gmr1b200_rx_call_batch so that a batch of calls leaves the GPU as one buffer.
#include "../../include/gmr1_b200.h"
#include "api_common.h"

using namespace gmr1;

namespace {

struct VoiceArgs {
	const int32_t *tch_rec;                // [n][max_frames][12]
	const uint8_t *tch_data;               // [n][max_frames][20]
	const int32_t *fn, *n_frames;          // [n][max_frames], [n]
	int32_t        n, max_frames;
	uint8_t       *voice, *flag;           // [n][2 max_frames][10], [n][2 max_frames]
	int32_t       *n_voice, *first_fn;     // [n]
};

// one CTA per channel: slot 0 is the first frame with traffic-channel activity, the last slot the release frame (or
// the last frame walked); a TDMA frame is two 20 ms slots
__global__ void __launch_bounds__(128) voice_stream_kernel(const VoiceArgs a)
{
	__shared__ int s_first, s_last;
	const int ch = blockIdx.x, tid = threadIdx.x;
	const int F = a.max_frames;
	const int nf = min(a.n_frames[ch], F);
	const int32_t *rec = a.tch_rec + (size_t)ch * F * 12;
	if (tid == 0) {
		s_first = INT_MAX;
		s_last = -1;
	}
	__syncthreads();
	int first = INT_MAX, last = -1;
	for (int f = tid; f < nf; f += blockDim.x) {
		const int kind = rec[f * 12 + 0], end = rec[f * 12 + 1];
		if (kind != GMR1B200_TCH_NONE || end) {
			first = min(first, f);
			last = max(last, f);
		}
	}
	if (first != INT_MAX) {
		atomicMin(&s_first, first);
		atomicMax(&s_last, last);
	}
	__syncthreads();
	first = s_first;
	last = s_last;
	uint8_t *voice = a.voice + (size_t)ch * 2 * F * 10;
	uint8_t *flag = a.flag + (size_t)ch * 2 * F;
	const int slots = last >= 0 ? 2 * (last - first + 1) : 0;
	if (tid == 0) {
		a.n_voice[ch] = slots;
		if (a.first_fn)
			a.first_fn[ch] = last >= 0 ? a.fn[(size_t)ch * F + first] : -1;
	}
	for (int i = tid; i < 2 * F * 10; i += blockDim.x) {
		const int slot = i / 10, b = i - slot * 10;
		const int f = first + (slot >> 1);
		const bool speech = slot < slots && rec[f * 12 + 0] == GMR1B200_TCH_SPEECH;
		voice[i] = speech ? a.tch_data[((size_t)ch * F + f) * 20 + (slot & 1) * 10 + b] : 0;
		if (b == 0)
			flag[slot] = slot < slots ? (speech ? 0 : 1) : 2;
	}
}

}  // namespace

extern "C" int gmr1b200_tch3_voice_stream_batch(const int32_t *tch_rec, const uint8_t *tch_data, const int32_t *fn,
                                                const int32_t *n_frames, int n, int max_frames, uint8_t *voice,
                                                uint8_t *flag, int32_t *n_voice, int32_t *first_fn, void *stream)
{
	if (n < 0 || max_frames < 1 || !tch_rec || !tch_data || !fn || !n_frames || !voice || !flag || !n_voice)
		return set_err(-EINVAL, "tch3_voice_stream_batch: bad argument");
	if (n == 0)
		return 0;
	const size_t nf = (size_t)n * max_frames;
	Stage s(stream);
	VoiceArgs a = {};
	a.tch_rec = s.in(tch_rec, nf * 12);
	a.tch_data = s.in(tch_data, nf * 20);
	a.fn = s.in(fn, nf);
	a.n_frames = s.in(n_frames, (size_t)n);
	a.n = n; a.max_frames = max_frames;
	a.voice = s.out(voice, nf * 20);
	a.flag = s.out(flag, nf * 2);
	a.n_voice = s.out(n_voice, (size_t)n);
	a.first_fn = s.out(first_fn, (size_t)n);
	if (s.failed())
		return s.finish(cudaSuccess, "tch3_voice_stream_batch: staging");
	voice_stream_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(a);
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess)
		g_launches.fetch_add(1);
	return s.finish(e, "voice_stream_kernel");
}
