// gsmtap_kernels.cu - decode results as GSMTAP records on the device (second half).
//
// One record per decoded unit: the 16-byte gsmtap_hdr the reference's gmr1_gsmtap_makemsg builds
// (src/gsmtap.c:44-71: version 2, hdr_len 4 words, type GSMTAP_TYPE_GMR1_UM, timeslot, frame number in
// network byte order, sub_type = channel type, everything else 0) followed by the L2 bytes.  The L2 of a
// batch is on the device after the decode kernels; serialising it there leaves one contiguous D2H copy
// (or a send straight from a pinned buffer) instead of n small host-side message builds.
// Pure byte movement: one thread per output byte, consecutive threads write consecutive bytes.
#include <cuda_runtime.h>

#include "launch.h"

namespace gmr1 {

static constexpr int GSMTAP_HDR = 16;

__global__ void __launch_bounds__(256) gsmtap_kernel(const GsmtapArgs a)
{
	const int rec_bytes = GSMTAP_HDR + a.len;
	const int64_t total = (int64_t)a.n * rec_bytes;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
		const int r = (int)(i / rec_bytes), o = (int)(i - (int64_t)r * rec_bytes);
		uint8_t v;
		if (o >= GSMTAP_HDR) {
			v = a.l2[(size_t)r * a.l2_stride + (o - GSMTAP_HDR)];
		} else {
			const uint32_t fn = a.fn ? a.fn[r] : a.fn0 + (uint32_t)r;
			switch (o) {
			case 0:  v = 0x02; break;                                   // GSMTAP_VERSION
			case 1:  v = GSMTAP_HDR / 4; break;                         // header length in 32-bit words
			case 2:  v = 0x0a; break;                                   // GSMTAP_TYPE_GMR1_UM
			case 3:  v = a.tn ? a.tn[r] : a.tn0; break;                 // timeslot
			case 8:  v = (uint8_t)(fn >> 24); break;                    // frame number, big endian
			case 9:  v = (uint8_t)(fn >> 16); break;
			case 10: v = (uint8_t)(fn >> 8); break;
			case 11: v = (uint8_t)fn; break;
			case 12: v = a.chan_type ? a.chan_type[r] : a.chan_type0; break;   // sub_type
			default: v = 0; break;                                      // arfcn, signal, snr, antenna, sub-slot, reserved
			}
		}
		a.out[(size_t)r * a.out_stride + o] = v;
	}
}

cudaError_t launch_gsmtap(const GsmtapArgs &a, cudaStream_t st)
{
	if (a.n <= 0)
		return cudaSuccess;
	const int64_t total = (int64_t)a.n * (GSMTAP_HDR + a.len);
	int64_t grid = (total + 255) / 256;
	if (grid > 148 * 16)
		grid = 148 * 16;
	gsmtap_kernel<<<(int)grid, 256, 0, st>>>(a);
	return cudaGetLastError();
}

}  // namespace gmr1
