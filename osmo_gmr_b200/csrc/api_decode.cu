// api_decode.cu - C ABI: library state + batched channel-decode entry points (include/gmr1_b200.h)
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "build_id.h"
#include "launch.h"

namespace gmr1 {
thread_local char g_err[512] = "";
thread_local int g_host_hint = 0;

// this thread's staging arena for the current device: 4 MB to start with (a 650 ms FCCH window is 0.5 MB), grown
// between calls to twice what the largest call so far would have needed
HostArena &host_arena()
{
	static thread_local HostArena ar[16];
	int dev = 0;
	cudaGetDevice(&dev);
	HostArena &a = ar[dev & 15];
	if (a.busy)
		return a;
	const size_t want = a.want > a.cap ? 2 * a.want : ((size_t)4 << 20);
	if (!a.h || a.want > a.cap) {
		if (a.h) {
			cudaDeviceSynchronize();
			cudaFreeHost(a.h);
			cudaFree(a.d);
			a.h = a.d = nullptr;
			a.cap = 0;
		}
		char *h = nullptr, *d = nullptr;
		if (cudaHostAlloc((void **)&h, want, cudaHostAllocPortable) == cudaSuccess && cudaMalloc((void **)&d, want) == cudaSuccess) {
			a.h = h;
			a.d = d;
			a.cap = want;
		} else {
			if (h) cudaFreeHost(h);
			cudaGetLastError();
		}
	}
	return a;
}
std::atomic<uint64_t> g_launches{0};
}

using namespace gmr1;

extern "C" {

int gmr1b200_init(int device)
{
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess)
		return cuda_rc(e, "cudaSetDevice");
	e = cudaFree(0);
	if (e != cudaSuccess)
		return cuda_rc(e, "CUDA context creation");
	return 0;
}

const char *gmr1b200_last_error(void) { return g_err; }
int gmr1b200_host_hint(int on)
{
	const int prev = g_host_hint;
	g_host_hint = on ? 1 : 0;
	return prev;
}
const char *gmr1b200_version(void) { return "gmr1_b200 0.2 (sm_100a) build " GMR1B200_BUILD_ID; }
uint64_t gmr1b200_kernel_launches(void) { return g_launches.load(); }

}  // extern "C"

// common tail: stage pointers, launch, copy back
static int run_decode(int ch, DecodeArgs a, int n_in, int n_ciph, int l2_bytes, void *stream,
                      size_t bits_s_per = 0, bool tch3 = false)
{
	if (a.n < 0 || !a.ebits || !a.l2 || (tch3 && !a.l2b))
		return set_err(-EINVAL, "decode_batch: NULL required pointer or negative n");
	if (a.n == 0)
		return 0;
	const size_t n = (size_t)a.n;
	Stage s(stream);
	a.ebits  = s.in(a.ebits, n * n_in);
	a.ciph   = s.in(a.ciph, n * n_ciph);
	a.prev1  = s.in(a.prev1, n);
	a.prev2  = s.in(a.prev2, n);
	a.sb_mask = s.in(a.sb_mask, n);
	a.l2     = s.out(a.l2, n * l2_bytes);
	a.l2b    = s.out(a.l2b, n * 10);
	a.conv   = s.out(a.conv, n);
	a.conv1  = s.out(a.conv1, n);
	a.crc    = s.out(a.crc, n);
	a.crc2   = s.out(a.crc2, n * 2);
	a.bits_s = s.out(a.bits_s, n * bits_s_per);
	a.sacch  = s.out(a.sacch, n * 10);
	a.status = s.out(a.status, n * 4);
	if (const size_t sb = decode_scratch_bytes(ch, a.n))
		a.dec_scratch = s.tmp<uint8_t>(sb);
	cudaError_t e = cudaSuccess;
	if (!s.failed()) {
		e = launch_decode(ch, a, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "decode kernel");
}

extern "C" {

int gmr1b200_bcch_decode_batch(uint8_t *l2, const int8_t *bits_e, int32_t *conv_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.n = n; a.l2 = l2; a.conv = conv_rv; a.crc = crc;
	return run_decode(CH_BCCH, a, 424, 0, 24, stream);
}

int gmr1b200_ccch_decode_batch(uint8_t *l2, const int8_t *bits_e, int32_t *conv_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.n = n; a.l2 = l2; a.conv = conv_rv; a.crc = crc;
	return run_decode(CH_CCCH, a, 432, 0, 24, stream);
}

int gmr1b200_facch3_decode_batch(uint8_t *l2, uint8_t *bits_s, const int8_t *bits_e, const uint8_t *ciph,
                                 int32_t *conv_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.ciph = ciph; a.n = n; a.l2 = l2; a.conv = conv_rv; a.crc = crc; a.bits_s = bits_s;
	return run_decode(CH_FACCH3, a, 416, 384, 10, stream, 32);
}

int gmr1b200_facch9_decode_batch(uint8_t *l2, int8_t *bits_sacch, int8_t *bits_status, const int8_t *bits_e,
                                 const uint8_t *ciph, int32_t *conv_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.ciph = ciph; a.n = n; a.l2 = l2; a.conv = conv_rv; a.crc = crc;
	a.sacch = bits_sacch; a.status = bits_status;
	return run_decode(CH_FACCH9, a, 662, 658, 38, stream);
}

int gmr1b200_tch3_decode_batch(uint8_t *frame0, uint8_t *frame1, uint8_t *bits_s, const int8_t *bits_e,
                               const uint8_t *ciph, int m, int32_t *conv0_rv, int32_t *conv1_rv, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.ciph = ciph; a.n = n; a.l2 = frame0; a.l2b = frame1; a.conv = conv0_rv; a.conv1 = conv1_rv;
	a.bits_s = bits_s; a.tch3_m = m ? 1 : 0;
	return run_decode(CH_TCH3, a, 212, 208, 10, stream, 4, true);
}

int gmr1b200_tch9_decode_batch(uint8_t *l2, int8_t *bits_sacch, int8_t *bits_status, const int8_t *bits_e, int mode,
                               const uint8_t *ciph, const int32_t *prev1, const int32_t *prev2,
                               int32_t *conv_rv, int n, void *stream)
{
	if (mode < 0 || mode > 2)
		return set_err(-EINVAL, "tch9_decode_batch: mode must be 0 (2k4), 1 (4k8) or 2 (9k6)");
	static const int bytes[3] = {18, 30, 60};
	DecodeArgs a = {};
	a.ebits = bits_e; a.ciph = ciph; a.n = n; a.l2 = l2; a.conv = conv_rv;
	a.sacch = bits_sacch; a.status = bits_status; a.prev1 = prev1; a.prev2 = prev2;
	return run_decode(CH_TCH9_2K4 + mode, a, 662, 658, bytes[mode], stream);
}

int gmr1b200_tch9_decode_rows_batch(uint8_t *l2, const int8_t *rows, int mode, int32_t *conv_rv, int n, void *stream)
{
	if (mode < 0 || mode > 2)
		return set_err(-EINVAL, "tch9_decode_rows_batch: mode must be 0 (2k4), 1 (4k8) or 2 (9k6)");
	static const int bytes[3] = {18, 30, 60};
	DecodeArgs a = {};
	a.ebits = rows; a.n = n; a.l2 = l2; a.conv = conv_rv; a.t9_rows = 1;
	return run_decode(CH_TCH9_2K4 + mode, a, 648, 0, bytes[mode], stream);
}

int gmr1b200_rach_decode_batch(uint8_t *rach, const int8_t *bits_e, const uint8_t *sb_mask, int sb_mask0,
                               int32_t *conv_rv, int32_t *crc_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.n = n; a.l2 = rach; a.conv = conv_rv; a.crc = crc; a.crc2 = crc_rv;
	a.sb_mask = sb_mask; a.sb_mask0 = sb_mask0 & 0xff;
	return run_decode(CH_RACH, a, 494, 0, 18, stream);
}

int gmr1b200_xch_dc12_decode_batch(uint8_t *l2, const int8_t *bits_e, int32_t *conv_rv, int32_t *crc, int n, void *stream)
{
	DecodeArgs a = {};
	a.ebits = bits_e; a.n = n; a.l2 = l2; a.conv = conv_rv; a.crc = crc;
	return run_decode(CH_DC12, a, 432, 0, 24, stream);
}

}  // extern "C"
