// api_demod.cu - C ABI: batched pi/4-CxPSK demodulation / burst-type detection (include/gmr1_b200.h)
#include "../../include/gmr1_b200.h"
#include "api_common.h"
#include "launch.h"

using namespace gmr1;

// device copy of the ten standard burst descriptors, one per device
static BurstTab *g_d_bursts[64] = {nullptr};

cudaError_t gmr1::device_bursts(const BurstTab **out)
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (dev >= 64)
		return cudaErrorInvalidDevice;
	GMR1_INIT_LOCK();
	if (!g_d_bursts[dev]) {
		BurstTab *d = nullptr;
		if ((e = cudaMalloc(&d, sizeof(BurstTab) * BT_COUNT)) != cudaSuccess)
			return e;
		for (int i = 0; i < BT_COUNT; i++)
			if ((e = cudaMemcpy(d + i, &burst_tab(i), sizeof(BurstTab), cudaMemcpyHostToDevice)) != cudaSuccess)
				return e;
		g_d_bursts[dev] = d;
	}
	*out = g_d_bursts[dev];
	return cudaSuccess;
}


static_assert(sizeof(gmr1b200_burst_desc) == sizeof(BurstTab), "public descriptor must mirror gmr1::BurstTab");

static int run_demod(const int *types, int n_types, int mode, DemodArgs a, int64_t iq_len, void *stream,
                     const BurstTab *custom = nullptr)
{
	if (a.n < 0 || !a.iq || n_types < 1 || n_types > 8 || a.win_len < 1)
		return set_err(-EINVAL, "pi4cxpsk batch: bad argument");
	for (int i = 0; i < n_types && !custom; i++)
		if (types[i] < 0 || types[i] >= BT_COUNT)
			return set_err(-EINVAL, "pi4cxpsk batch: unknown burst type");
	for (int i = 0; i < n_types && custom; i++) {
		const BurstTab &t = custom[i];
		bool ok = t.nbits >= 1 && t.nbits <= 2 && t.len > 0 && t.len <= 468 && t.n_sync >= 1 && t.n_sync <= MAX_SYNC &&
		          t.n_data >= 0 && t.n_data <= MAX_DATA_CHUNK;
		for (int s = 0; ok && s < t.n_sync; s++) {
			ok = t.n_chunk[s] >= 1 && t.n_chunk[s] <= MAX_SYNC_CHUNK;
			for (int c = 0; ok && c < t.n_chunk[s]; c++)
				ok = t.s_len[s][c] >= 1 && t.s_len[s][c] <= MAX_SYNC_SYMS && t.s_pos[s][c] >= 0 &&
				     t.s_pos[s][c] + t.s_len[s][c] <= t.len;
		}
		for (int s = 0; ok && s < t.n_sync; s++)
			for (int c = 0; ok && c < t.n_chunk[s]; c++)
				for (int k = 0; ok && k < t.s_len[s][c]; k++)
					ok = t.s_sym[s][c][k] < 4;
		int n_dsym = 0;
		for (int c = 0; ok && c < t.n_data; c++) {
			ok = t.d_pos[c] >= 0 && t.d_len[c] >= 0 && t.d_pos[c] + t.d_len[c] <= t.len;
			for (int c2 = 0; ok && c2 < c; c2++)          // data chunks must not overlap (each symbol sliced once)
				ok = t.d_pos[c] >= t.d_pos[c2] + t.d_len[c2] || t.d_pos[c2] >= t.d_pos[c] + t.d_len[c];
			n_dsym += t.d_len[c];
		}
		// the kernel writes nbits soft bits per data symbol whatever `ebits` says: they have to agree
		ok = ok && t.ebits == n_dsym * t.nbits;
		if (!ok)
			return set_err(-EINVAL, "pi4cxpsk batch: malformed burst descriptor");
	}
	if (a.sps < 1 || a.sps > 16)
		return set_err(-EINVAL, "pi4cxpsk batch: sps must be 1..16");
	const BurstTab &t0 = custom ? custom[0] : burst_tab(types[0]);
	if (a.win_len < t0.len * a.sps)
		return set_err(-EINVAL, "pi4cxpsk batch: window shorter than the burst");
	if (mode == 0 && (!a.ebits || a.ebits_stride < t0.ebits))
		return set_err(-EINVAL, "pi4cxpsk_demod_batch: ebits NULL or ebits_stride too small");
	if (a.n == 0)
		return 0;
	if (!a.ofs && (a.stride < 0 || (int64_t)(a.n - 1) * a.stride + a.win_len > iq_len))
		return set_err(-EINVAL, "pi4cxpsk batch: windows exceed iq_len");
	if (a.ofs && host_pointer(a.ofs))                     // offsets the host can read are range-checked here
		for (int i = 0; i < a.n; i++)
			if (a.ofs[i] < 0 || a.ofs[i] + a.win_len > iq_len)
				return set_err(-EINVAL, "pi4cxpsk batch: a window offset is outside iq_len");

	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");

	const size_t n = (size_t)a.n;
	Stage s(stream);
	a.iq = (const float2 *)s.in((const float *)a.iq, (size_t)iq_len * 2);
	a.ofs = s.in(a.ofs, n);
	a.freq_shift = s.in(a.freq_shift, n);
	a.e_toa = s.in(a.e_toa, n);
	a.ebits = s.out(a.ebits, n * (size_t)a.ebits_stride);
	a.sync_id = s.out(a.sync_id, n);
	a.bt_id = s.out(a.bt_id, n);
	a.toa = s.out(a.toa, n);
	a.freq_err = s.out(a.freq_err, n);
	a.pwr = s.out(a.pwr, n);

	// the kernel takes a dense array of the selected descriptors
	BurstTab h_sel[8];
	for (int i = 0; i < n_types; i++)
		h_sel[i] = custom ? custom[i] : burst_tab(types[i]);
	const BurstTab *d_sel = nullptr;
	if (custom) {
		d_sel = s.in(custom, (size_t)n_types);      // host descriptors: staged like any other input
	} else if (n_types == 1) {
		d_sel = d_all + types[0];
	} else {
		BurstTab *tmp = nullptr;
		if (!s.failed()) {
			cudaError_t e2 = cudaMallocAsync(&tmp, sizeof(BurstTab) * n_types, (cudaStream_t)stream);
			if (e2 == cudaSuccess) {
				for (int i = 0; i < n_types; i++)
					cudaMemcpyAsync(tmp + i, d_all + types[i], sizeof(BurstTab), cudaMemcpyDeviceToDevice,
					                (cudaStream_t)stream);
				d_sel = tmp;
			} else {
				return s.finish(e2, "descriptor scratch");
			}
		}
	}
	if (!s.failed()) {
		e = launch_demod(a, d_sel, h_sel, n_types, mode, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	if (!custom && n_types > 1 && d_sel)
		cudaFreeAsync((void *)d_sel, (cudaStream_t)stream);
	return s.finish(e, "pi4cxpsk kernel");
}

static int run_synth(int burst_type, SynthArgs a, int64_t iq_len, void *stream)
{
	if (burst_type < 0 || burst_type >= BT_COUNT || a.n < 0 || !a.ebits || !a.iq)
		return set_err(-EINVAL, "synth_bursts: bad argument");
	const BurstTab &t = burst_tab(burst_type);
	if (a.sps < 1 || a.sps > 16 || a.win_len < 1 || a.ebits_stride < t.ebits)
		return set_err(-EINVAL, "synth_bursts: bad sps / win_len / ebits_stride");
	if (a.n == 0)
		return 0;
	if (!a.ofs && (a.stride < 0 || (int64_t)(a.n - 1) * a.stride + a.win_len > iq_len))
		return set_err(-EINVAL, "synth_bursts: windows exceed iq_len");
	const BurstTab *d_all = nullptr;
	cudaError_t e = device_bursts(&d_all);
	if (e != cudaSuccess)
		return cuda_rc(e, "burst table upload");
	const size_t n = (size_t)a.n;
	Stage s(stream);
	a.ebits = s.in(a.ebits, n * (size_t)a.ebits_stride);
	a.sync_id = s.in(a.sync_id, n);
	a.toa = s.in(a.toa, n);
	a.cfo = s.in(a.cfo, n);
	a.phase = s.in(a.phase, n);
	a.esn0_db = s.in(a.esn0_db, n);
	a.amp = s.in(a.amp, n);
	a.ofs = s.in(a.ofs, n);
	a.iq = (float2 *)s.out((float *)a.iq, (size_t)iq_len * 2);
	if (!s.failed()) {
		e = launch_synth(a, d_all + burst_type, (cudaStream_t)stream);
		if (e == cudaSuccess)
			g_launches.fetch_add(1);
	}
	return s.finish(e, "synth kernel");
}

extern "C" {

int gmr1b200_synth_bursts(int burst_type, const uint8_t *ebits, int ebits_stride, const int32_t *sync_id,
                          int sps, int win_len, const float *toa, float toa0, const float *cfo, float cfo0,
                          const float *phase, float phase0, const float *esn0_db, float esn0_db0,
                          const float *amp, float amp0, uint64_t seed,
                          float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride, int n, void *stream)
{
	SynthArgs a = {};
	a.ebits = ebits; a.ebits_stride = ebits_stride; a.sync_id = sync_id; a.n = n; a.sps = sps; a.win_len = win_len;
	a.toa = toa; a.toa0 = toa0; a.cfo = cfo; a.cfo0 = cfo0; a.phase = phase; a.phase0 = phase0;
	a.esn0_db = esn0_db; a.esn0_db0 = esn0_db0; a.amp = amp; a.amp0 = amp0; a.seed = seed;
	a.iq = (float2 *)iq; a.ofs = win_ofs; a.stride = win_stride;
	return run_synth(burst_type, a, iq_len, stream);
}

int gmr1b200_synth_bursts_tx(int burst_type, const uint8_t *ebits, int ebits_stride, const int32_t *sync_id,
                             int sps, int win_len, const float *toa, float toa0, const float *cfo, float cfo0,
                             const float *phase, float phase0, const float *esn0_db, float esn0_db0,
                             const float *amp, float amp0, uint64_t seed,
                             float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride, int n, void *stream)
{
	SynthArgs a = {};
	a.ebits = ebits; a.ebits_stride = ebits_stride; a.sync_id = sync_id; a.n = n; a.sps = sps; a.win_len = win_len;
	a.toa = toa; a.toa0 = toa0; a.cfo = cfo; a.cfo0 = cfo0; a.phase = phase; a.phase0 = phase0;
	a.esn0_db = esn0_db; a.esn0_db0 = esn0_db0; a.amp = amp; a.amp0 = amp0; a.seed = seed;
	a.iq = (float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.tx_pulse = 1;
	return run_synth(burst_type, a, iq_len, stream);
}

int gmr1b200_set_sync_accumulator_reset(int on)
{
	return gmr1::g_sync_reset.exchange(on ? 1 : 0);
}

int gmr1b200_set_demod_generic(int on)
{
	return gmr1::g_demod_generic.exchange(on ? 1 : 0);
}

int gmr1b200_burst_desc_get(int bt, struct gmr1b200_burst_desc *out)
{
	if (bt < 0 || bt >= BT_COUNT || !out)
		return -EINVAL;
	memcpy(out, &burst_tab(bt), sizeof(*out));
	return 0;
}

int gmr1b200_pi4cxpsk_demod_desc_batch(const struct gmr1b200_burst_desc *desc, const float *iq, int64_t iq_len,
                                       const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                       const float *freq_shift, float freq_shift0, int8_t *ebits, int ebits_stride,
                                       int32_t *sync_id, float *toa, float *freq_err, float *pwr, int n, void *stream)
{
	if (!desc)
		return set_err(-EINVAL, "pi4cxpsk_demod_desc_batch: desc NULL");
	DemodArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.e_toa0 = -1.0f;
	a.ebits = ebits; a.ebits_stride = ebits_stride; a.sync_id = sync_id; a.toa = toa; a.freq_err = freq_err; a.pwr = pwr;
	const int dummy = 0;
	return run_demod(&dummy, 1, 0, a, iq_len, stream, reinterpret_cast<const BurstTab *>(desc));
}

int gmr1b200_pi4cxpsk_detect_desc_batch(const struct gmr1b200_burst_desc *descs, int n_types, const float *e_toa,
                                        float e_toa0, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                        int64_t win_stride, int win_len, int sps, const float *freq_shift,
                                        float freq_shift0, int32_t *bt_id, int32_t *sync_id, float *toa, int n,
                                        void *stream)
{
	if (!descs)
		return set_err(-EINVAL, "pi4cxpsk_detect_desc_batch: descs NULL");
	DemodArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.e_toa = e_toa; a.e_toa0 = e_toa0;
	a.bt_id = bt_id; a.sync_id = sync_id; a.toa = toa;
	const int dummy[8] = {0};
	return run_demod(dummy, n_types, 1, a, iq_len, stream, reinterpret_cast<const BurstTab *>(descs));
}

int gmr1b200_burst_len(int bt)
{
	return (bt < 0 || bt >= BT_COUNT) ? -EINVAL : burst_tab(bt).len;
}

int gmr1b200_burst_ebits(int bt)
{
	return (bt < 0 || bt >= BT_COUNT) ? -EINVAL : burst_tab(bt).ebits;
}

int gmr1b200_pi4cxpsk_demod_batch(int burst_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                  int64_t win_stride, int win_len, int sps, const float *freq_shift, float freq_shift0,
                                  int8_t *ebits, int ebits_stride, int32_t *sync_id, float *toa, float *freq_err,
                                  float *pwr, int n, void *stream)
{
	DemodArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.e_toa0 = -1.0f;
	a.ebits = ebits; a.ebits_stride = ebits_stride; a.sync_id = sync_id; a.toa = toa; a.freq_err = freq_err; a.pwr = pwr;
	return run_demod(&burst_type, 1, 0, a, iq_len, stream);
}

int gmr1b200_pi4cxpsk_detect_batch(const int *burst_types, int n_types, const float *e_toa, float e_toa0,
                                   const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                                   int win_len, int sps, const float *freq_shift, float freq_shift0,
                                   int32_t *bt_id, int32_t *sync_id, float *toa, int n, void *stream)
{
	if (!burst_types)
		return set_err(-EINVAL, "pi4cxpsk_detect_batch: burst_types NULL");
	DemodArgs a = {};
	a.iq = (const float2 *)iq; a.ofs = win_ofs; a.stride = win_stride; a.n = n; a.win_len = win_len; a.sps = sps;
	a.freq_shift = freq_shift; a.freq_shift0 = freq_shift0; a.e_toa = e_toa; a.e_toa0 = e_toa0;
	a.bt_id = bt_id; a.sync_id = sync_id; a.toa = toa;
	return run_demod(burst_types, n_types, 1, a, iq_len, stream);
}

}  // extern "C"
