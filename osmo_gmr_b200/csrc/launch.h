// launch.h - internal launcher entry points shared between the kernel translation units and
// the C-ABI layer (api_*.cu).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include "decode_unit.cuh"

namespace gmr1 {

// The lazy per-device initialisation inside the launchers (table uploads to __constant__ / __device__ memory,
// cudaFuncSetAttribute, cached device properties) can be reached from several host threads at once (the device pool
// of api_multi.cu runs one thread per GPU, callers may run one per stream): every launcher holds this lock from its
// first look at that state until its kernel is enqueued.  Uncontended cost ~20 ns per launch.
std::mutex &init_mutex();
#define GMR1_INIT_LOCK() std::lock_guard<std::mutex> gmr1_init_lock_(gmr1::init_mutex())

// ---- stage 3
cudaError_t launch_decode(int ch, const DecodeArgs &a, cudaStream_t st);
// bytes of DecodeArgs::dec_scratch for n units of channel ch (0: the channel keeps its decisions on chip)
size_t decode_scratch_bytes(int ch, int n);

// ---- stage 2
struct DemodArgs {
	const float2  *iq;          // complex float samples (device)
	const int64_t *ofs;         // [n] first sample of each burst window within iq, or NULL
	int64_t        stride;      // used when ofs == NULL: window b starts at b*stride
	int32_t        n, win_len, sps;
	const float   *freq_shift;  // [n] rad/symbol or NULL (then freq_shift0)
	float          freq_shift0;
	const float   *e_toa;       // detect: [n] expected TOA or NULL (then e_toa0; < 0 = none)
	float          e_toa0;
	int8_t        *ebits;       // [n][ebits_stride]
	int32_t        ebits_stride;
	int32_t       *sync_id;     // [n] or NULL
	int32_t       *bt_id;       // detect: [n] or NULL
	float         *toa;         // [n] or NULL
	float         *freq_err;    // [n] or NULL
	float         *pwr;         // [n] or NULL (sync power; detect: after the e_toa weighting)
	int32_t        sync_reset;  // filled in by launch_demod from g_sync_reset (callers leave it alone)
	const int32_t *n_dev;       // optional device-side burst count (<= n): bursts beyond it are skipped (rx scheduler)
};

// gmr1b200_set_sync_accumulator_reset: 0 = reference behaviour (accumulator never cleared between candidate
// sequences).  launch_demod applies it, so every caller (batch entry points, fused entries, frame loop) agrees.
extern std::atomic<int> g_sync_reset;
// d_bts: n_bt burst descriptors in device memory, h_bts: the same on the host (for geometry)
cudaError_t launch_demod(const DemodArgs &a, const BurstTab *d_bts, const BurstTab *h_bts, int n_bt, int mode,
                         cudaStream_t st);

// per-format kernels (demod_fast.cu): takes the batch when (standard burst type bt, sps 4, search width) is one of the
// compiled combinations and returns true (*err = launch status); false: the generic kernel has to run
extern std::atomic<int> g_demod_generic;      // gmr1b200_set_demod_generic
bool launch_demod_fast(const DemodArgs &a, int bt, cudaStream_t st, cudaError_t *err);

// ---- stage 1: FCCH
struct FcchArgs {
	const float2  *iq;
	const int64_t *ofs;
	int64_t        stride;
	const int32_t *rel;                    // fine: optional [n] extra offset (the rough TOA), clamped to [0, rel_max]
	int32_t        rel_max, rel_add;       //       rel_add: write toa = rel + fine TOA (the channel alignment)
	int32_t        n, win_len, sps, len;   // len = FCCH burst length in symbols
	float          freq;                   // chirp sweep (0.32 / 0.16)
	const float   *freq_shift;
	float          freq_shift0;
	int32_t       *toa;                    // [n] samples
	float         *freq_error;             // fine: [n] rad/symbol
	float         *snr;                    // snr: [n]
	float         *peak;                   // rough: [n] energy of the winning window, or NULL
	float         *en_out;                 // rough: if set, [n][win_len/sps - len + 1] |corr|^2 is written and
	                                       //        the peak search is left to the caller (rough_multi)
	const int32_t *skip;                   // optional [n]: entry b is left alone when skip[b] != 0 (device-side lists)
};
cudaError_t launch_fcch_rough(const FcchArgs &a, cudaStream_t st);
// second-generation coarse search (fcch_grid.cu): shifts == NULL: one search per window with the window's own shift;
// else n_shifts searches per window sharing one pass over the samples, toa / peak [n_shifts][n].
// cudaErrorNotSupported: geometry outside the kernel's range, use launch_fcch_rough
cudaError_t launch_fcch_grid(const FcchArgs &a, const float *shifts, int n_shifts, int32_t *toa, float *peak, cudaStream_t st);
// third generation (fcch_fft.cu): the same searches with the correlation in the frequency domain; same arguments and
// results as launch_fcch_grid; cudaErrorNotSupported: per-window shifts, sps != 4, windows beyond 8192 symbols, switched off
cudaError_t launch_fcch_fft(const FcchArgs &a, const float *shifts, int n_shifts, int32_t *toa, float *peak, cudaStream_t st);
void fcch_fft_enable(int on);
cudaError_t launch_fcch_fine(const FcchArgs &a, int mode, cudaStream_t st);

// ---- DKAB / modulation order
struct MiscArgs {
	const float2  *iq;
	const int64_t *ofs;
	int64_t        stride;
	int32_t        n, win_len, sps;
	const float   *freq_shift;
	float          freq_shift0;
	const int32_t *dkab_p;                 // [n] DKAB position or NULL (then dkab_p0)
	int32_t        dkab_p0;
	int8_t        *ebits;                  // dkab: [n][8]
	float         *toa;                    // dkab: [n]
	int32_t       *rv;                     // dkab: 0 found / 1 not a DKAB / -EINVAL;  mod_order: 2 / 4
	const int32_t *n_dev;                  // optional device-side unit count (<= n): units beyond it are skipped (rx scheduler)
};
cudaError_t launch_dkab(const MiscArgs &a, cudaStream_t st);
cudaError_t launch_mod_order(const MiscArgs &a, cudaStream_t st);

// ---- A5 keystreams
struct A5Args {
	const int32_t *alg;          // [n] 0 / 1 or NULL (then alg0)
	int32_t        alg0;
	const uint8_t *key;          // [n][8]
	const uint32_t *fn;          // [n]
	int32_t        n, nbits, stride;
	uint8_t       *dl, *ul;      // [n][stride] ubits, either may be NULL
	const int32_t *n_dev;        // optional device-side unit count (<= n): units beyond it are skipped (rx scheduler)
	int32_t        n_dev_mul;    // ... times this factor (0 = 1): e.g. four A5 streams per FACCH3 codeword
};
cudaError_t launch_a5(const A5Args &a, cudaStream_t st);
void a5_force_mode(int mode);           // tests / A-B: -1 by batch size, 0 one unit per thread, 1 bitsliced (32 units per thread)

// ---- GSMTAP records of decoded units
struct GsmtapArgs {
	const uint8_t  *chan_type;   // [n] GSMTAP_GMR1_* sub-type or NULL (then chan_type0)
	const uint32_t *fn;          // [n] frame numbers or NULL (then fn0 + i)
	const uint8_t  *tn;          // [n] timeslots or NULL (then tn0)
	uint8_t         chan_type0, tn0;
	uint32_t        fn0;
	const uint8_t  *l2;          // [n][l2_stride]
	int32_t         l2_stride, len, n, out_stride;
	uint8_t        *out;         // [n][out_stride], out_stride >= 16 + len
};
cudaError_t launch_gsmtap(const GsmtapArgs &a, cudaStream_t st);

// ---- workload synthesis
struct SynthArgs {
	const uint8_t *ebits;       // [n][ebits_stride] hard bits
	int32_t        ebits_stride;
	const int32_t *sync_id;     // [n] or NULL (0)
	int32_t        n, sps, win_len;
	const float   *toa, *cfo, *phase, *esn0_db, *amp;   // [n] each or NULL (then the *0 scalar)
	float          toa0, cfo0, phase0, esn0_db0, amp0;
	uint64_t       seed;
	float2        *iq;
	const int64_t *ofs;
	int64_t        stride;
	int32_t        tx_pulse;    // 0: raised cosine (after the matched filter), 1: root raised cosine (transmit side)
};
cudaError_t launch_synth(const SynthArgs &a, const BurstTab *d_bt, cudaStream_t st);
// device copy of the standard burst descriptors (api_demod.cu)
cudaError_t device_bursts(const BurstTab **out);

}  // namespace gmr1
