/* gmr1_b200.h - C ABI of the B200-native batched GMR-1 receiver PHY (libgmr1_b200.so).
 *
 * The library is a drop-in for the per-burst receive path of osmocom/osmo-gmr
 * (libgmr1-sdr + libgmr1-l1, consumed by src/gmr1_rx.c).  It exports
 *
 *   (1) batched entry points  gmr1b200_*_batch : n independent units per call, one or a few
 *       CUDA kernel launches, plain pointers + sizes.  Every pointer argument may be either
 *       HOST memory (pageable or pinned) or DEVICE memory; the library detects which
 *       (cudaPointerGetAttributes).  With only device pointers a call is asynchronous on
 *       `stream`; as soon as one pointer is host memory the call stages that buffer through
 *       device scratch on `stream` and returns after the results are back (the reference's
 *       synchronous semantics).  `stream` is a cudaStream_t passed as void* (NULL = default).
 *
 *   (2) the reference's own symbols (gmr1_bcch_decode, gmr1_pi4cxpsk_demod, ...), same
 *       signatures, as n = 1 wrappers over (1), declared in gmr1_b200_compat.h.
 *
 * Each declaration cites the reference interface it replaces (path:line under the reference
 * root).  Array layouts are "unit-major": n consecutive copies of the reference's per-call
 * array.  sbit_t = int8_t (+127 strong 0 ... -127 strong 1, 0 erased), ubit_t = uint8_t
 * holding one bit, L2 bytes LSB-first exactly as the reference packs them.
 *
 * Error convention (reference: 0 ok, -errno hard error, >0 soft condition):
 *   0 success, -EINVAL bad argument, -ENOMEM allocation failure, -EIO CUDA failure
 *   (text from gmr1b200_last_error()), -ENODEV no usable GPU.  There is NO CPU fallback.
 */
#ifndef GMR1_B200_H
#define GMR1_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int8_t  gmr1b200_sbit_t;
typedef uint8_t gmr1b200_ubit_t;

/* ---- library state --------------------------------------------------------------------- */

/* Select the CUDA device for the calling thread and upload the constant tables.
 * Optional: every entry point initialises the current device lazily. */
int gmr1b200_init(int device);
/* Text of the last CUDA / argument error on the calling thread ("" if none). */
const char *gmr1b200_last_error(void);
/* Library version string. */
const char *gmr1b200_version(void);
/* Calling-thread promise: until switched off again, every pointer this thread passes to a batched entry point is HOST
 * memory.  The library then skips the per-pointer classification and stages through a per-thread page-locked arena
 * (one H2D copy per input, one D2H copy for all outputs of a call).  Set by the n = 1 wrappers of the reference's own
 * symbols (csrc/compat.c); returns the previous setting. */
int gmr1b200_host_hint(int on);
/* Number of CUDA kernels this library has launched in this process (all threads). */
uint64_t gmr1b200_kernel_launches(void);

/* ---- stage 3: channel decode (soft bits -> L2) -------------------------------------------
 * crc[i]  = what the reference function returns for unit i (0 = CRC ok)
 * conv_rv[i] = osmo_conv_decode() return value (Viterbi path metric); may be NULL
 * All are bit-exact with the reference C path (integer arithmetic only). */

/* replaces gmr1_bcch_decode, src/l1/bcch.c:84 (include/osmocom/gmr1/l1/bcch.h:38)
 * l2 [n][24], bits_e [n][424] */
int gmr1b200_bcch_decode_batch(uint8_t *l2, const gmr1b200_sbit_t *bits_e,
                               int32_t *conv_rv, int32_t *crc, int n, void *stream);

/* replaces gmr1_ccch_decode, src/l1/ccch.c:88 (l1/ccch.h:38).  l2 [n][24], bits_e [n][432] */
int gmr1b200_ccch_decode_batch(uint8_t *l2, const gmr1b200_sbit_t *bits_e,
                               int32_t *conv_rv, int32_t *crc, int n, void *stream);

/* replaces gmr1_facch3_decode, src/l1/facch3.c:122 (l1/facch3.h:39-40)
 * l2 [n][10], bits_s [n][32] (may be NULL), bits_e [n][416] (4 bursts x 104),
 * ciph [n][384] ubit or NULL */
int gmr1b200_facch3_decode_batch(uint8_t *l2, gmr1b200_ubit_t *bits_s,
                                 const gmr1b200_sbit_t *bits_e, const gmr1b200_ubit_t *ciph,
                                 int32_t *conv_rv, int32_t *crc, int n, void *stream);

/* replaces gmr1_facch9_decode, src/l1/facch9.c:107 (l1/facch9.h:40-41)
 * l2 [n][38], bits_sacch [n][10] sbit, bits_status [n][4] sbit (either may be NULL),
 * bits_e [n][662], ciph [n][658] or NULL */
int gmr1b200_facch9_decode_batch(uint8_t *l2, gmr1b200_sbit_t *bits_sacch, gmr1b200_sbit_t *bits_status,
                                 const gmr1b200_sbit_t *bits_e, const gmr1b200_ubit_t *ciph,
                                 int32_t *conv_rv, int32_t *crc, int n, void *stream);

/* replaces gmr1_tch3_decode, src/l1/tch3.c:124 (l1/tch3.h:40-42)
 * frame0/frame1 [n][10] (MSB first), bits_s [n][4] ubit or NULL, bits_e [n][212],
 * ciph [n][208] or NULL, m = multiplexing mode, conv0_rv / conv1_rv [n] or NULL */
int gmr1b200_tch3_decode_batch(uint8_t *frame0, uint8_t *frame1, gmr1b200_ubit_t *bits_s,
                               const gmr1b200_sbit_t *bits_e, const gmr1b200_ubit_t *ciph, int m,
                               int32_t *conv0_rv, int32_t *conv1_rv, int n, void *stream);

/* replaces gmr1_tch9_decode + struct gmr1_interleaver, src/l1/tch9.c:140, interleave.c:168
 * (l1/tch9.h:50-53).  mode: 0 = 2k4 (l2 [n][18]), 1 = 4k8 ([n][30]), 2 = 9k6 ([n][60])
 * (enum gmr1_tch9_mode).  The stateful depth-3 inter-burst de-interleaver is expressed as a
 * gather: prev1[i] / prev2[i] = index (within this batch) of the burst received one / two
 * bursts before burst i on the same channel, or -1 when there is none (interleaver memory is
 * zero, as after gmr1_interleaver_init).  prev1 / prev2 NULL = -1 everywhere. */
int gmr1b200_tch9_decode_batch(uint8_t *l2, gmr1b200_sbit_t *bits_sacch, gmr1b200_sbit_t *bits_status,
                               const gmr1b200_sbit_t *bits_e, int mode, const gmr1b200_ubit_t *ciph,
                               const int32_t *prev1, const int32_t *prev2,
                               int32_t *conv_rv, int n, void *stream);

/* Second half of gmr1_tch9_decode for callers that keep the reference's stateful interleaver:
 * rows [n][648] are the soft bits AFTER decipher / descramble / gmr1_deinterleave_inter (tch9.c:150-165);
 * this does the intra de-interleave, de-puncturing and Viterbi (tch9.c:166-174).  l2 as above. */
int gmr1b200_tch9_decode_rows_batch(uint8_t *l2, const gmr1b200_sbit_t *rows, int mode,
                                    int32_t *conv_rv, int n, void *stream);

/* replaces gmr1_rach_decode, src/l1/rach.c:137 (l1/rach.h:38-39)
 * rach [n][18], bits_e [n][494], sb_mask [n] or NULL (then sb_mask0 for every unit),
 * crc_rv [n][2] or NULL, crc [n] = return value (crc_rv[0] || crc_rv[1]) */
int gmr1b200_rach_decode_batch(uint8_t *rach, const gmr1b200_sbit_t *bits_e,
                               const uint8_t *sb_mask, int sb_mask0,
                               int32_t *conv_rv, int32_t *crc_rv, int32_t *crc, int n, void *stream);

/* replaces gmr1_xch_dc12_decode, src/l1/xch_dc12.c:87 (l1/xch_dc12.h:38)
 * l2 [n][24], bits_e [n][432] */
int gmr1b200_xch_dc12_decode_batch(uint8_t *l2, const gmr1b200_sbit_t *bits_e,
                                   int32_t *conv_rv, int32_t *crc, int n, void *stream);

/* ---- stage 2: pi/4-CxPSK burst demodulation (IQ window -> soft bits) ------------------------
 * Burst formats: the ten descriptors of the reference's src/sdr/nb.c (include/osmocom/gmr1/sdr/
 * nb.h:37-46), selected by id.  A burst "window" is burst_len*sps + search_window complex
 * float samples, exactly what gmr1_rx.c:burst_map() cuts (src/gmr1_rx.c:149-170); the extra
 * samples are the TOA search range.  Windows are addressed inside one IQ buffer (a recording
 * or a batch of recordings) by sample offset, so nothing is copied to form them. */
enum gmr1b200_burst_type {
	GMR1B200_BT_BCCH = 0,       /* gmr1_bcch_burst        nb.c:54  */
	GMR1B200_BT_DC2,            /* gmr1_dc2_burst         nb.c:81  */
	GMR1B200_BT_DC6,            /* gmr1_dc6_burst         nb.c:112 */
	GMR1B200_BT_DC12,           /* gmr1_dc12_burst        nb.c:143 */
	GMR1B200_BT_NT3_SPEECH,     /* gmr1_nt3_speech_burst  nb.c:170 */
	GMR1B200_BT_NT3_FACCH,      /* gmr1_nt3_facch_burst   nb.c:202 */
	GMR1B200_BT_NT6,            /* gmr1_nt6_burst         nb.c:240 */
	GMR1B200_BT_NT9,            /* gmr1_nt9_burst         nb.c:281 */
	GMR1B200_BT_RACH,           /* gmr1_rach_burst        nb.c:317 */
	GMR1B200_BT_SDCCH,          /* gmr1_sdcch_burst       nb.c:369 */
	GMR1B200_BT_COUNT
};

/* Reference quirk switch.  _gmr1_pi4cxpsk_sync_find clears its correlation accumulator once per call,
 * not once per candidate training sequence (src/sdr/pi4cxpsk.c:207 vs :232-233), so for formats with
 * several sequences (NT3-FACCH, NT6, NT9, SDCCH) candidate i is scored on the sum of candidates 0..i and
 * the last one always wins.  Default (0) reproduces that bit for bit; 1 scores every candidate on its own
 * correlation (what the code evidently intends; needed to tell FACCH9 from TCH9).  Process-wide; returns
 * the previous setting. */
int gmr1b200_set_sync_accumulator_reset(int on);

/* Kernel selection switch (testing / A-B measurements).  The ten standard burst formats at sps 4 and their standard
 * search widths (what gmr1_rx cuts: BCCH 20*sps, DC6 10*sps, NT3 / NT9 sps + sps/2, src/gmr1_rx.c:290,549,759,809) run
 * on kernels specialised per format at compile time (RACH with sps + sps/2 included); everything else (other sps /
 * widths, custom descriptors, detect) runs on the generic kernel.  1 forces the generic kernel for everything; both compute the same function
 * and are held to the same parity tests.  Process-wide; returns the previous setting. */
int gmr1b200_set_demod_generic(int on);

/* symbols (incl. guard) and soft bits of a burst type; -EINVAL for a bad id */
int gmr1b200_burst_len(int burst_type);
int gmr1b200_burst_ebits(int burst_type);

/* replaces gmr1_pi4cxpsk_demod, src/sdr/pi4cxpsk.c:520 (sdr/pi4cxpsk.h:101-105), n bursts/call.
 *   iq          interleaved (re, im) float32, iq_len complex samples in total
 *   win_ofs     [n] first complex sample of each window, or NULL: window b starts at b*win_stride
 *   win_len     complex samples per window (same for the whole batch) = burst_len*sps + search
 *   sps         samples per symbol, 1..16 (4 = compile-time fast path; sps < 4 takes the reference's
 *               sinc-interpolating symbol alignment, pi4cxpsk.c:298-343)
 *   freq_shift  [n] rad/symbol pre-rotation (the reference's freq_shift argument) or NULL,
 *               then freq_shift0 applies to every burst
 *   ebits       [n][ebits_stride] soft bits out (ebits_stride >= burst ebits)
 *   sync_id, toa (samples, fractional), freq_err (rad/symbol), pwr: [n] each, any may be NULL
 * sync_id[i] = -1 marks a window in which no training sequence correlated (the reference
 * returns an error there); its ebits are zeroed. */
int gmr1b200_pi4cxpsk_demod_batch(int burst_type, const float *iq, int64_t iq_len,
                                  const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                  const float *freq_shift, float freq_shift0,
                                  gmr1b200_sbit_t *ebits, int ebits_stride,
                                  int32_t *sync_id, float *toa, float *freq_err, float *pwr,
                                  int n, void *stream);

/* replaces gmr1_pi4cxpsk_detect, src/sdr/pi4cxpsk.c:617 (sdr/pi4cxpsk.h:107-110).
 * burst_types[n_types] must have equal length and modulation rotation (as the reference
 * requires); e_toa [n] expected TOA or NULL (then e_toa0; negative = no weighting).
 * Outputs bt_id (index into burst_types), sync_id, toa: [n] each, any may be NULL. */
int gmr1b200_pi4cxpsk_detect_batch(const int *burst_types, int n_types,
                                   const float *e_toa, float e_toa0,
                                   const float *iq, int64_t iq_len,
                                   const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                   const float *freq_shift, float freq_shift0,
                                   int32_t *bt_id, int32_t *sync_id, float *toa,
                                   int n, void *stream);

/* ---- transmit side: channel encoders (host code; used by the signal generator and tools) ----
 * Bit-exact mirrors of the reference encoders; they run on the CPU on purpose (not the hot path). */

/* replaces gmr1_bcch_encode (src/l1/bcch.c:59), gmr1_ccch_encode (ccch.c:59), gmr1_xch_dc12_encode
 * (xch_dc12.c:63).  chan: 0 = BCCH (bits_e [n][424]), 1 = CCCH ([n][432]), 2 = DC12 ([n][432]);
 * l2 [n][24]; bits_e are hard ubits. */
int gmr1b200_xcch_encode_batch(int chan, gmr1b200_ubit_t *bits_e, const uint8_t *l2, int n);
/* replaces gmr1_facch3_encode, src/l1/facch3.c:64.  bits_e [416], l2 [10], bits_s [32], ciph [384]/NULL */
int gmr1b200_facch3_encode(gmr1b200_ubit_t *bits_e, const uint8_t *l2, const gmr1b200_ubit_t *bits_s,
                           const gmr1b200_ubit_t *ciph);
/* replaces gmr1_facch9_encode, src/l1/facch9.c:57.  bits_e [662], l2 [38], sacch [10], status [4], ciph [658]/NULL */
int gmr1b200_facch9_encode(gmr1b200_ubit_t *bits_e, const uint8_t *l2, const gmr1b200_ubit_t *bits_sacch,
                           const gmr1b200_ubit_t *bits_status, const gmr1b200_ubit_t *ciph);
/* replaces gmr1_tch9_encode + gmr1_interleaver_init/fini, src/l1/tch9.c:93, interleave.c:95-130 */
void *gmr1b200_tch9_interleaver_new(void);
void gmr1b200_tch9_interleaver_free(void *interleaver);
int gmr1b200_tch9_encode(gmr1b200_ubit_t *bits_e, const uint8_t *l2, int mode, const gmr1b200_ubit_t *bits_sacch,
                         const gmr1b200_ubit_t *bits_status, const gmr1b200_ubit_t *ciph, void *interleaver);
/* first half of gmr1_tch9_encode (tch9.c:105-107): conv-encode + puncture + intra interleave -> ep [648],
 * for callers that keep the reference's stateful inter-burst interleaver object */
int gmr1b200_tch9_encode_ep(gmr1b200_ubit_t *ep, const uint8_t *l2, int mode);
/* replaces gmr1_rach_encode, src/l1/rach.c:76.  bits_e [494], rach [18] */
int gmr1b200_rach_encode(gmr1b200_ubit_t *bits_e, const uint8_t *rach, int sb_mask);
/* replaces gmr1_tch3_encode, src/l1/tch3.c:60 - the reference passes the arguments of
 * osmo_conv_encode in the wrong order (tch3.c:81), this one encodes what gmr1_tch3_decode decodes.
 * bits_e [212], frame0/frame1 [10] MSB first, bits_s [4], ciph [208]/NULL, m = mux mode */
int gmr1b200_tch3_encode(gmr1b200_ubit_t *bits_e, const uint8_t *frame0, const uint8_t *frame1,
                         const gmr1b200_ubit_t *bits_s, const gmr1b200_ubit_t *ciph, int m);

/* Flattened burst-format descriptor for callers that bring their own format (the compat layer
 * converts the reference's struct gmr1_pi4cxpsk_burst, sdr/pi4cxpsk.h:77-98, into this). */
#define GMR1B200_MAX_SYNC        4     /* GMR1_MAX_SYNC      (sdr/pi4cxpsk.h:39) */
#define GMR1B200_MAX_SYNC_CHUNK  6
#define GMR1B200_MAX_SYNC_SYMS   32    /* GMR1_MAX_SYNC_SYMS (sdr/pi4cxpsk.h:40) */
#define GMR1B200_MAX_DATA_CHUNK  6
struct gmr1b200_burst_desc {
	float   rotation;                 /* per-symbol rotation, rad (pi/4 or pi/2) */
	int32_t nbits;                    /* bits per symbol, 1 or 2 */
	int32_t len;                      /* symbols incl. guard */
	int32_t ebits;                    /* soft bits produced */
	int32_t n_sync;                   /* alternative training sequences */
	int32_t n_chunk[GMR1B200_MAX_SYNC];
	int16_t s_pos[GMR1B200_MAX_SYNC][GMR1B200_MAX_SYNC_CHUNK];
	int16_t s_len[GMR1B200_MAX_SYNC][GMR1B200_MAX_SYNC_CHUNK];
	uint8_t s_sym[GMR1B200_MAX_SYNC][GMR1B200_MAX_SYNC_CHUNK][GMR1B200_MAX_SYNC_SYMS]; /* phase index 0..3 (k*pi/2) */
	int32_t n_data;
	int16_t d_pos[GMR1B200_MAX_DATA_CHUNK];
	int16_t d_len[GMR1B200_MAX_DATA_CHUNK];
};
/* copy of a built-in descriptor */
int gmr1b200_burst_desc_get(int burst_type, struct gmr1b200_burst_desc *out);
/* as gmr1b200_pi4cxpsk_demod_batch / _detect_batch with caller-supplied descriptors (host memory) */
int gmr1b200_pi4cxpsk_demod_desc_batch(const struct gmr1b200_burst_desc *desc, const float *iq, int64_t iq_len,
                                       const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                       const float *freq_shift, float freq_shift0,
                                       gmr1b200_sbit_t *ebits, int ebits_stride,
                                       int32_t *sync_id, float *toa, float *freq_err, float *pwr,
                                       int n, void *stream);
int gmr1b200_pi4cxpsk_detect_desc_batch(const struct gmr1b200_burst_desc *descs, int n_types,
                                        const float *e_toa, float e_toa0,
                                        const float *iq, int64_t iq_len,
                                        const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                        const float *freq_shift, float freq_shift0,
                                        int32_t *bt_id, int32_t *sync_id, float *toa,
                                        int n, void *stream);

/* replaces gmr1_fcch_rough_multi, src/sdr/fcch.c:341 (sdr/fcch.h:51-53): all overlapping FCCHs in one
 * >= 650 ms window.  The 117-tap correlation and the peak bookkeeping (two-cycle mixing, avg + 3 sigma
 * threshold, sorted de-duplicated insert, fcch.c:373-483, in the reference's summation order) both run on the
 * GPU.  toa [N] out (host memory; the 16 strongest at most, as gmr1_rx asks for); returns the number of FCCHs
 * found (>= 0) or -errno (-EINVAL: window shorter than 650 ms or the two cycles do not line up). */
int gmr1b200_fcch_rough_multi(int fcch_type, const float *iq, int64_t win_len, int sps, float freq_shift,
                              int32_t *toa, int N, void *stream);

/* ---- burst window -> L2 in one call ----------------------------------------------------------------
 * gmr1_pi4cxpsk_demod + gmr1_{bcch,ccch,xch_dc12}_decode (rx_bcch / rx_ccch, src/gmr1_rx.c:747-851) for n
 * windows: chan 0 = BCCH burst + BCCH decode, 1 = DC6 burst + CCCH decode, 2 = DC12 burst + DC12 decode.
 * One call, TWO kernel launches (demod, then decode): the soft bits stay in a device buffer between them (L2-resident
 * for batches up to ~250 k bursts) and never reach the host.  A single fused kernel was weighed and not
 * built: DESIGN.md 4.9 has the measurements.  l2 [n][24]; crc / conv / toa / freq_err [n], each may be NULL.
 * Results are identical to calling gmr1b200_pi4cxpsk_demod_batch and the *_decode_batch one after the other. */
int gmr1b200_rx_xcch_batch(int chan, const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                           int win_len, int sps, const float *freq_shift, float freq_shift0,
                           uint8_t *l2, int32_t *crc, int32_t *conv, float *toa, float *freq_err,
                           int n, void *stream);

/* ---- receiver frame loop for n channels in lock step --------------------------------
 * replaces process_bcch (src/gmr1_rx.c:853-895) with rx_bcch (:747-803), rx_ccch (:805-851; the TCH3
 * hand-off is reported by gmr1b200_rx_bcch_ass_batch below), bcch_tdma_align (:194-236), burst_map / burst_energy (:149-182) for n channels at once.
 * Channel i is the recording iq[rec_ofs[i] .. + rec_len[i]) (complex samples) and starts as the reference's
 * chan_desc after FCCH acquisition: align0[i] (samples from the start of the recording; what
 * gmr1b200_fcch_acquire_batch returns plus the offset of its search window), freq_err0[i] (rad/symbol, NULL = 0),
 * fn = sa_sirfn_delay = sa_bcch_stn = 0.  Frame by frame, every channel demodulates + decodes the BCCH burst
 * (SI-relative frame 2 of 8) or the CCCH burst (frames 1, 3..7, only when the window energy reaches half of the
 * last BCCH window's), feeds TOA / frequency error / SI1 timing of every good BCCH burst back into its state,
 * and stops when fewer than two frames of samples remain; all of it on the device, no host round trip.
 * Outputs, [n][max_frames] each (l2: [n][max_frames][24]): kind (0 nothing, 1 BCCH, 2 CCCH), fn (frame number
 * the channel believed in), crc (0 ok, -1 where kind = 0), conv (Viterbi metric), l2; n_frames [n] frames
 * walked; align_out / freq_err_out [n] final tracking state (may be NULL).  Frames beyond max_frames are not
 * walked.  kind and crc are set for every frame slot; fn is written for the n_frames[i] walked frames of channel i,
 * conv and l2 where kind != 0 - the other slots of fn / conv / l2 are unspecified.  Every pointer host or device memory. */
int gmr1b200_rx_bcch_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                           const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                           int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                           int32_t *n_frames, int32_t *align_out, float *freq_err_out, void *stream);

/* Scheduling switch of the two walks above (testing / A-B measurements).  Default (0): every channel is paced by its own
 * BCCH bursts - the only frames that change its tracking state - so a call alternates "BCCH frame of every channel"
 * and "all frames up to the next BCCH frame of every channel, as one batch": two dependent rounds of kernels per eight
 * frames.  1: all channels frame by frame in lock step (eight rounds per eight frames, each frame's launches replayed
 * as a CUDA graph).  Same results, record for record.  Process-wide; returns the previous setting. */
int gmr1b200_set_rx_lockstep(int on);

/* The same walk, and the TCH3 hand-off of rx_ccch (src/gmr1_rx.c:836-841): a CCCH burst with a good CRC that is an
 * IMMEDIATE ASSIGNMENT (ccch_is_imm_ass :236-239) initialises the channel's TCH3 state as rx_tch3_init does (:362-381).
 * tch3 [n][4]: active (0 / 1), tn (receive timeslot) and p (DKAB position) from ccch_imm_ass_parse (:240-245), frame
 * index (into the [max_frames] outputs) of the assignment, -1 without one; tch3_energy [n][2] (may be NULL):
 * energy_burst = 0.75 x the CCCH energy gate at that frame (half the energy of the last BCCH window) and
 * energy_dkab = energy_burst / 8.  A later IMM.ASS overwrites an earlier one, as in the reference.  The TCH3 burst
 * loop that follows (rx_tch3 :538-600, on the traffic-channel recording) stays with the caller, who has everything it
 * starts from: this record, fn of that frame, and the alignment / frequency tracking of the channel. */
int gmr1b200_rx_bcch_ass_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                               const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                               int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2,
                               int32_t *n_frames, int32_t *align_out, float *freq_err_out,
                               int32_t *tch3, float *tch3_energy, void *stream);

/* ---- the whole receiver frame loop, traffic channels included ----------------------------------
 * replaces process_bcch (src/gmr1_rx.c:853-895) with everything it calls per frame: rx_bcch / rx_ccch (as
 * gmr1b200_rx_bcch_batch above), rx_tch3 (:538-600) once an IMMEDIATE ASSIGNMENT has been seen on the CCCH - energy
 * gate, gmr1_dkab_demod or gmr1_pi4cxpsk_detect, running energy averages, release after ten silent frames, FACCH3
 * assembly over four frames with the plain attempt / ciphered retry / cipher discovery of _rx_tch3_facch_flush
 * (:394-452), speech bursts through gmr1_a5 + gmr1_tch3_decode - and rx_tch9 (:276-355) once a good FACCH3 message is
 * an ASSIGNMENT COMMAND 1: NT9 demodulation, FACCH9 (sync sequence 0) or TCH9-9k6 through the channel's depth-3
 * interleaver history.  n channels in lock step, all state on the device, no host round trip.
 *   tch_ofs [n]   first sample of the channel's traffic-carrier recording in iq (what gmr1_rx takes as its second
 *                 file), -1 = none; csd_ofs [n] or NULL the same for the TCH9 carrier (fourth argument of gmr1_rx);
 *                 both recordings are as long as the BCCH one (rec_len), as the reference assumes (:163)
 *   kc [n][8]     cipher keys, NULL = all zero
 * Control-channel outputs as gmr1b200_rx_bcch_batch.  Per channel and frame, [n][max_frames]:
 *   tch_rec [12]  int32: [0] GMR1B200_TCH_* of the frame, [1] 1 = the channel was released in this frame ("END"),
 *                 [2] sync id of a FACCH3 burst, [3] FACCH3 decode attempts made in this frame (0, 1, 2),
 *                 [4] [5] crc / Viterbi metric of the first attempt, [6] [7] of the ciphered retry,
 *                 [8] [9] Viterbi metrics of the two speech frames, [10] timeslot of an IMM.ASS seen in this frame or -1,
 *                 [11] 1 = tch_data holds a good FACCH3 message
 *   tch_data [20] speech: frame0 | frame1 (10 bytes each, as gmr1_tch3_decode); FACCH3: the 10-byte L2 message
 *   csd_rec [6]   int32: [0] GMR1B200_CSD_*, [1] sync id, [2] crc (FACCH9), [3] Viterbi metric, [4] mean soft-bit
 *                 magnitude (TCH9, what the reference prints as avg), [5] reserved
 *   csd_data [60] FACCH9: 38-byte L2; TCH9: the 60-byte block
 * Every pointer host or device memory. */
enum { GMR1B200_TCH_NONE = 0, GMR1B200_TCH_DKAB = 1, GMR1B200_TCH_DKAB_MISS = 2, GMR1B200_TCH_FACCH3 = 3,
       GMR1B200_TCH_SPEECH = 4 };
enum { GMR1B200_CSD_NONE = 0, GMR1B200_CSD_FACCH9 = 1, GMR1B200_CSD_TCH9 = 2 };
int gmr1b200_rx_call_batch(const float *iq, int64_t iq_len, const int64_t *rec_ofs, const int32_t *rec_len,
                           const int64_t *tch_ofs, const int64_t *csd_ofs, const uint8_t *kc,
                           const int32_t *align0, const float *freq_err0, int sps, int n, int max_frames,
                           int32_t *kind, int32_t *fn, int32_t *crc, int32_t *conv, uint8_t *l2, int32_t *n_frames,
                           int32_t *tch_rec, uint8_t *tch_data, int32_t *csd_rec, uint8_t *csd_data, void *stream);

/* ---- TCH3 speech frames -> vocoder input (the AMBE hand-off) ---------------------------------------
 * The reference hands speech to its vocoder as a stream of 10-byte frames, one per 20 ms, two per TCH3 burst:
 * gmr1_tch3_decode returns frame0 / frame1 (src/l1/tch3.c:116-184), src/gmr1_ambe_decode.c:127-150 reads such a stream
 * 10 bytes at a time into gmr1_codec_decode_frame(codec, audio, 160, frame, bad) (codec/codec.h:40-45) and
 * gmr1_codec_decode_dtx covers a slot without a frame.  The vocoder is out of scope; this builds what it is fed,
 * per channel and with the timing kept, from the per-frame records of gmr1b200_rx_call_batch (tch_rec, tch_data, fn,
 * n_frames as that call wrote them).  Slot 0 is the first 20 ms of the first frame with traffic-channel activity, the
 * last slot the second half of the release frame (or of the last frame walked); every TDMA frame is two slots.
 *   voice [n][2 max_frames][10]  the frames, zeros where flag != 0
 *   flag  [n][2 max_frames]      0 = speech frame (decode_frame, bad = 0); 1 = no speech in this slot - FACCH3-stolen,
 *                                DKAB / silence, missed burst (decode_dtx); 2 = behind the end of the stream
 *   n_voice [n]                  slots of the stream (0: the channel never had a traffic channel)
 *   first_fn [n] or NULL         frame number of slot 0, -1 without a stream
 * Concatenating the flag-0 frames of a channel gives the file gmr1_ambe_decode takes.  Host or device pointers. */
int gmr1b200_tch3_voice_stream_batch(const int32_t *tch_rec, const uint8_t *tch_data, const int32_t *fn,
                                     const int32_t *n_frames, int n, int max_frames, uint8_t *voice, uint8_t *flag,
                                     int32_t *n_voice, int32_t *first_fn, void *stream);

/* ---- A5 cipher stream (host; input to the ciphered decoders) ------------------------------------
 * replaces gmr1_a5 / gmr1_a5_1, src/l1/a5.c:57,226 (l1/a5.h:37-41): n = 0 (all zero) or 1 (A5/1-GMR);
 * key [8], dl / ul [nbits] ubits, either may be NULL */
void gmr1b200_a5(int n, const uint8_t *key, uint32_t fn, int nbits, gmr1b200_ubit_t *dl, gmr1b200_ubit_t *ul);

/* The same for n (Kc, frame number) pairs in one launch on the device: unit i gets
 * gmr1_a5(alg[i] or alg0, key + 8 i, fn[i], nbits, dl + i*stride, ul + i*stride), src/l1/a5.c:57.  The rows
 * are what the ciphered *_decode_batch entry points take as ciph, so the masks need never exist on the host.
 * dl or ul may be NULL; stride >= nbits (the stride - nbits bytes behind each row are unspecified afterwards);
 * every pointer host or device memory. */
int gmr1b200_a5_batch(const int32_t *alg, int alg0, const uint8_t *key, const uint32_t *fn, int nbits, int stride,
                      gmr1b200_ubit_t *dl, gmr1b200_ubit_t *ul, int n, void *stream);
/* Kernel selection switch (testing / A-B measurements): batches of 163 840 units and more run bitsliced - 32 keystreams
 * per thread, bit l of every register word belongs to unit l (csrc/a5_bitslice.cuh) - smaller ones one unit per
 * thread.  -1: by batch size (default), 0: one unit per thread always, 1: bitsliced always.  Same streams, bit for
 * bit.  Process-wide; returns the previous setting. */
int gmr1b200_set_a5_bitslice(int mode);

/* ---- decode results as GSMTAP records -------------------------------------------
 * replaces, for a whole batch on the device, gmr1_gsmtap_makemsg (src/gsmtap.c:44-71, gmr1/gsmtap.h): record i is
 * the 16-byte gsmtap_hdr {version 2, hdr_len 4, type GSMTAP_TYPE_GMR1_UM, timeslot tn[i], arfcn 0, signal 0, snr 0,
 * frame_number htonl(fn[i]), sub_type chan_type[i], antenna 0, sub_slot 0, res 0} followed by len bytes of
 * l2 + i*l2_stride, written to out + i*out_stride (out_stride >= 16 + len; the out_stride - 16 - len bytes behind a record are
 * unspecified afterwards).
 * chan_type / fn / tn may be NULL: then chan_type0 / fn0 + i / tn0 are used.  Every pointer host or device memory;
 * with device pointers the records of a batch leave the GPU in one copy. */
int gmr1b200_gsmtap_batch(const uint8_t *chan_type, int chan_type0, const uint32_t *fn, uint32_t fn0,
                          const uint8_t *tn, int tn0, const uint8_t *l2, int l2_stride, int len,
                          uint8_t *out, int out_stride, int n, void *stream);

/* ---- stage 1: FCCH chirp acquisition ----------------------------------------------------------
 * fcch_type: 0 = gmr1_fcch_burst (sweep 0.32, 117 symbols), 1 = gmr1_fcch3_lband_burst (0.32, 468),
 * 2 = gmr1_fcch3_sband_burst (0.16, 468)  (reference src/sdr/fcch.c:50-70, sdr/fcch.h:42-44).
 * Windows are addressed as for the demodulator (win_ofs / win_stride into iq). */

/* replaces gmr1_fcch_rough, src/sdr/fcch.c:211 (sdr/fcch.h:47-49): coarse TOA of the strongest FCCH
 * in each search window (gmr1_rx uses 330 ms = 30 888 samples at sps 4, src/gmr1_rx.c:612).
 * toa [n] in samples; peak [n] (energy of the winning 5-symbol window, our extra) may be NULL. */
int gmr1b200_fcch_rough_batch(int fcch_type, const float *iq, int64_t iq_len,
                              const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                              const float *freq_shift, float freq_shift0,
                              int32_t *toa, float *peak, int n, void *stream);

/* Kernel selection switch (testing / A-B measurements): the coarse FCCH searches (gmr1b200_fcch_rough_batch with one
 * shift for all windows, _rough_grid_batch, _acquire_batch; sps 4, windows up to 8192 symbols) correlate in the
 * frequency domain - one forward transform of the decimated window, one product + reverse transform per shift
 * (csrc/fcch_fft.cu); windows of 4097 .. 7808 symbols (the standard 330 ms window has 7722) as two overlapping
 * 4096-point blocks, others as one 8192-point block.  0 sends them to the direct-correlation kernels
 * (csrc/fcch_grid.cu), which also serve every other geometry; 2 = frequency domain, always one 8192-point block.
 * Same results up to float rounding.  Process-wide; returns the previous setting (default 1). */
int gmr1b200_set_fcch_fft(int on);
/* gmr1_fcch_rough for a GRID of frequency shifts per window (the "+-frequency-offset FCCH search" of a receiver that
 * does not know its carrier offset yet: n_shifts calls of gmr1_fcch_rough, src/sdr/fcch.c:211, with freq_shift =
 * shifts[k], on the same window).  The window is read, averaged, decimated and normalised once; the shift moves
 * onto the 117 reference taps, and a pair of shifts +-f shares its two real-tap correlations.  shifts [n_shifts]
 * host memory (1..16, rad/symbol); toa [n_shifts][n], peak [n_shifts][n] or NULL (energy of the winning window). */
int gmr1b200_fcch_rough_grid_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *win_ofs,
                                   int64_t win_stride, int win_len, int sps, const float *shifts, int n_shifts,
                                   int32_t *toa, float *peak, int n, void *stream);

/* replaces gmr1_fcch_fine, src/sdr/fcch.c:512 (sdr/fcch.h:55-57): each window is exactly
 * burst_len*sps samples (else the reference returns -EINVAL); toa [n] samples, freq_error [n] rad/symbol */
int gmr1b200_fcch_fine_batch(int fcch_type, const float *iq, int64_t iq_len,
                             const int64_t *win_ofs, int64_t win_stride, int sps,
                             const float *freq_shift, float freq_shift0,
                             int32_t *toa, float *freq_error, int n, void *stream);

/* replaces fcch_single_init, src/gmr1_rx.c:606-639, for n channels in two launches with no host round
 * trip: gmr1_fcch_rough (freq_shift 0) over each search window, then gmr1_fcch_fine (freq_shift 0) on the
 * burst_len*sps samples at the rough TOA.  align [n] = rough + fine TOA in samples from the window start
 * (what the reference accumulates into chan_desc.align), freq_error [n] rad/symbol; rough_toa [n] may be
 * NULL.  win_ofs / win_stride / win_len describe the SEARCH windows. */
int gmr1b200_fcch_acquire_batch(int fcch_type, const float *iq, int64_t iq_len,
                                const int64_t *win_ofs, int64_t win_stride, int win_len, int sps,
                                int32_t *rough_toa, int32_t *align, float *freq_error, int n, void *stream);

/* replaces gmr1_fcch_snr, src/sdr/fcch.c:643 (sdr/fcch.h:59-61): snr [n] */
int gmr1b200_fcch_snr_batch(int fcch_type, const float *iq, int64_t iq_len,
                            const int64_t *win_ofs, int64_t win_stride, int sps,
                            const float *freq_shift, float freq_shift0,
                            float *snr, int n, void *stream);

/* replaces fcch_multi_process, src/gmr1_rx.c:643-741, up to its callback, for n recordings: every
 * FCCH a receiver should follow.  Recording i is iq[rec_ofs[i] .. + rec_len[i]) (complex samples); align[i] /
 * freq_err[i] are its primary acquisition (gmr1b200_fcch_acquire_batch; freq_err NULL = 0).  Per recording:
 * gmr1_fcch_rough_multi over the 650 ms that start one FCCH burst before align[i] (freq_shift -freq_err[i], up to 16
 * peaks), gmr1_fcch_fine and gmr1_fcch_snr on each peak, then the reference's filter: the strongest peak always
 * survives, another one only with snr >= 2, snr >= (strongest snr) / 6 and a frequency error within 500 Hz of the
 * strongest's.  Correlation, fine and SNR stages are one launch each per 256 recordings; the scalar bookkeeping
 * between them runs on the host as in gmr1b200_fcch_rough_multi.
 * n_fcch [n]: survivors (0 .. max_cand, max_cand <= 16) or -EINVAL (fewer than 650 ms of samples behind the window
 * start - the reference prints "Not enough samples" - or the two FCCH cycles do not line up, fcch.c:426).
 * cand_align [n][max_cand]: alignment in samples from the start of the recording, strongest first: what
 * process_bcch starts from (chan_desc.align, gmr1_rx.c:731), i.e. align0 of gmr1b200_rx_bcch_batch with the
 * primary freq_err[i] as freq_err0.  cand_snr / cand_freq_err [n][max_cand] (linear SNR, fine frequency error in
 * rad/symbol relative to freq_err[i]) may be NULL; entries behind the n_fcch[i] survivors are unspecified.
 * iq may be host or device memory; the per-recording arrays and
 * the outputs are HOST memory (the call synchronises the stream three times per 256 recordings). */
int gmr1b200_fcch_multi_batch(int fcch_type, const float *iq, int64_t iq_len, const int64_t *rec_ofs,
                              const int32_t *rec_len, const int32_t *align, const float *freq_err, int sps, int n,
                              int max_cand, int32_t *n_fcch, int32_t *cand_align, float *cand_snr,
                              float *cand_freq_err, void *stream);

/* ---- DKAB and modulation order ------------------------------------------------------------------ */

/* replaces gmr1_dkab_demod, src/sdr/dkab.c:187 (sdr/dkab.h:40-42).  p [n] DKAB position or NULL (p0);
 * ebits [n][8] (written when rv == 0), toa [n], rv [n]: 0 = DKAB found, 1 = not a DKAB, -EINVAL window
 * shorter than a DKAB burst */
int gmr1b200_dkab_demod_batch(const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                              int win_len, int sps, const float *freq_shift, float freq_shift0,
                              const int32_t *p, int p0, gmr1b200_sbit_t *ebits, float *toa, int32_t *rv,
                              int n, void *stream);

/* replaces gmr1_pi4cxpsk_mod_order, src/sdr/pi4cxpsk.c:693 (sdr/pi4cxpsk.h:112-113): order [n] = 2 or 4 */
int gmr1b200_pi4cxpsk_mod_order_batch(const float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                                      int win_len, int sps, const float *freq_shift, float freq_shift0,
                                      int32_t *order, int n, void *stream);

/* ---- workload synthesis: pi/4-CxPSK burst generator (GPU) -------------------------------------
 * Not in the reference (its gmr1_pi4cxpsk_mod, src/sdr/pi4cxpsk.c:741, is 1 sample/symbol with no
 * pulse or channel).  Writes n burst windows of win_len complex samples into iq: hard bits ->
 * symbols per the burst format, pi/4 rotation, raised-cosine pulse (RRC 0.35 x RRC 0.35) at `sps`,
 * symbol 0 at fractional sample position toa inside the window, carrier offset cfo (rad/symbol),
 * phase (rad), amplitude amp, AWGN at Es/N0 = esn0_db (Philox4x32-10 keyed by seed, burst, sample).
 * Each per-burst parameter is an array [n] or NULL (then the scalar that follows it applies). */
int gmr1b200_synth_bursts(int burst_type, const gmr1b200_ubit_t *ebits, int ebits_stride, const int32_t *sync_id,
                          int sps, int win_len, const float *toa, float toa0, const float *cfo, float cfo0,
                          const float *phase, float phase0, const float *esn0_db, float esn0_db0,
                          const float *amp, float amp0, uint64_t seed,
                          float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                          int n, void *stream);
/* the same with the TRANSMIT pulse only (RRC 0.35, unit energy per symbol) and no noise when esn0_db >= 100: the
 * waveform in front of the receiver's matched filter, i.e. what the wideband channeliser below has to be fed */
int gmr1b200_synth_bursts_tx(int burst_type, const gmr1b200_ubit_t *ebits, int ebits_stride, const int32_t *sync_id,
                             int sps, int win_len, const float *toa, float toa0, const float *cfo, float cfo0,
                             const float *phase, float phase0, const float *esn0_db, float esn0_db0,
                             const float *amp, float amp0, uint64_t seed,
                             float *iq, int64_t iq_len, const int64_t *win_ofs, int64_t win_stride,
                             int n, void *stream);

/* ---- several GPUs of one box from one process: the device pool -------------------------------------
 * The reference handles its channels one after the other in one thread (process loop over chan_desc,
 * src/gmr1_rx.c:732-741).  ARFCNs never interact, so a pool shards them: ARFCN a is processed on device
 * a mod G.  Every device has one feeder thread (bound to the CPUs next to the GPU when
 * /sys/bus/pci/devices/<id>/local_cpulist names them), streams_per_dev streams and as many device buffers of
 * chunk_bytes; host IQ travels in chunks (strided H2D copy of every G-th ARFCN -> the batched entry points above
 * on device pointers -> strided D2H copy into the caller's result arrays).  No collective, no peer traffic.
 * Host buffers should be page-locked (gmr1b200_host_alloc, or the caller's own cudaHostAlloc / torch pinned
 * memory): pageable memory works but its copies are staged and synchronous.  Calls on one pool must not overlap. */
struct gmr1b200_pool;
int gmr1b200_pool_create(const int *devices /* [n_dev] CUDA ordinals, NULL: 0 .. n_dev-1 */, int n_dev,
                         int streams_per_dev, int64_t chunk_bytes, struct gmr1b200_pool **out);
void gmr1b200_pool_destroy(struct gmr1b200_pool *pool);
int gmr1b200_pool_size(const struct gmr1b200_pool *pool);
void *gmr1b200_host_alloc(size_t bytes);       /* page-locked, usable from every device; NULL on failure */
void gmr1b200_host_free(void *p);

/* gmr1b200_rx_xcch_batch over the pool: host_iq is [n_arfcn][per_arfcn][win_len] complex float (interleaved),
 * results [n_arfcn][per_arfcn] (l2: x 24 bytes), crc / conv / toa may be NULL. */
int gmr1b200_pool_rx_xcch(struct gmr1b200_pool *pool, int chan, const float *host_iq, int n_arfcn, int per_arfcn,
                          int win_len, int sps, float freq_shift0,
                          uint8_t *l2, int32_t *crc, int32_t *conv, float *toa);

/* gmr1b200_fcch_acquire_batch over the pool: one search window of win_len complex samples per ARFCN,
 * host_iq [n_arfcn][win_len]; align / freq_error [n_arfcn], rough [n_arfcn] or NULL. */
int gmr1b200_pool_fcch_acquire(struct gmr1b200_pool *pool, int fcch_type, const float *host_iq, int n_arfcn,
                               int win_len, int sps, int32_t *rough, int32_t *align, float *freq_error);

/* ---- wideband channeliser -----------------------------------------------------------------------------
 * Replaces the "PFB Channelizer mode" of utils/gmr1_rx_sdr.py (:391-604): one wideband recording in, one stream of
 * sps x 23.4 kS/s per ARFCN out - the per-ARFCN cfile contents gmr1_rx reads (gmr1_rx.c:924-988), here left in device
 * memory (or copied to the host) in exactly the layout the rx_* / *_batch entry points take as `iq` + offsets.
 * The reference wires GNU Radio blocks; the plan restates what PFBBase / PFBOutputParameters compute:
 *   bank       pfb.channelizer_ccf(n_chans, firdes.low_pass(1, samp_rate, 15 625, 7 812.5), oversample 2)   (:433-470)
 *   per ARFCN  pfb.arb_resampler_ccf(sps 23 400 / 62 500, firdes.root_raised_cosine(32, 32 x 62 500, 23 400, 0.35,
 *              11 symbols), flt_size 32)                                                               (:520-529, :591-596)
 * The recording's rate must sit on the carrier grid: samp_rate = n_chans x 31.25 kHz, n_chans even (the script's
 * pre-rotator / pre-resampler, :398-401 / :455-462, bring a capture there and are not part of this entry).
 * Bank channel k is centred k x 31.25 kHz above the recording's centre, k >= n_chans / 2 meaning k - n_chans
 * (PFBBase.freq2index, :489-496).  GNU Radio is not in the reference tree: outputs are checked against the CPU
 * restatement in oracle/chan_port.py (float tolerance) and by decoding them with the reference's C receive path. */
struct gmr1b200_chan_info {
	int32_t n_chans, sps;
	int32_t n_taps, taps_per_branch;      /* bank prototype filter */
	int32_t n_taps_resamp, fft_stages;
	double  samp_rate, mid_rate, resamp;  /* wideband rate, bank output rate per channel (62.5 kS/s), 1.4976 at sps 4 */
	double  delay_out;                    /* group delay of bank + RRC, in output samples */
};
int gmr1b200_chan_create(int n_chans, int sps, void **plan);
void gmr1b200_chan_destroy(void *plan);
int gmr1b200_chan_info(void *plan, struct gmr1b200_chan_info *info);
/* the two filter designs (host memory): taps [>= n_taps], taps_resamp [>= n_taps_resamp]; either may be NULL */
int gmr1b200_chan_taps(void *plan, float *taps, int max_taps, float *taps_resamp, int max_resamp);
/* output samples per channel for a recording of n_wide samples */
int64_t gmr1b200_chan_out_len(void *plan, int64_t n_wide);
/* wide: n_wide complex samples from the start of the recording (zeros are assumed in front of it), iq_format 0 =
 * complex float32, 1 = interleaved int16 I/Q scaled by 1 / 32768 (what SDR front ends deliver; half the bytes), 2 =
 * interleaved int8 I/Q scaled by 1 / 128 (8-bit front ends; a quarter of the bytes);
 * chan_idx [n_wanted] bank channels wanted, NULL = channels 0 .. n_wanted-1;
 * out [n_wanted][out_stride] complex float (interleaved), out_stride >= gmr1b200_chan_out_len(n_wide) samples.
 * Host or device pointers.  A HOST recording travels in up to 16 pieces on a copy stream of the plan while the bank
 * and the resampler of the pieces before run on `stream` (page-locked memory - gmr1b200_host_alloc - makes the copies
 * asynchronous); the results are the same as for one piece. */
int gmr1b200_channelize(void *plan, const void *wide, int iq_format, int64_t n_wide, const int32_t *chan_idx, int n_wanted,
                        float *out, int64_t out_stride, void *stream);
/* Streaming form: consecutive blocks of ONE endless recording (what the reference's flowgraph does with its SDR source).
 * The state keeps what the filters still need of the past - the last 18 bank steps of samples, the last rows of the
 * bank's output the 30-tap resampler phases reach back to, the resampler's phase walk - on the device, so the blocks
 * may have any length (a fraction of a bank step up to seconds) and the concatenated outputs are, bit for bit, what
 * gmr1b200_channelize makes of the whole recording at once.  One state per recording and device; calls on one state
 * must not overlap and should use one CUDA stream.
 *   create   chan_idx [n_wanted] bank channels (host or device memory, copied), NULL = channels 0 .. n_wanted-1
 *   push     wide: n_wide NEW samples (host or device memory; iq_format as above, fixed for the life of the state);
 *            out [n_wanted][out_stride] complex float receives this block's outputs from column 0, *n_out (host) their
 *            number per channel - 0 while less than a bank step has arrived; out_stride >= chan_stream_max_out(n_wide)
 *   max_out  upper bound of the outputs one push of n_wide samples can produce */
int gmr1b200_chan_stream_create(void *plan, const int32_t *chan_idx, int n_wanted, void **state);
void gmr1b200_chan_stream_destroy(void *state);
int64_t gmr1b200_chan_stream_max_out(void *state, int64_t n_wide);
int gmr1b200_chan_stream_push(void *state, const void *wide, int iq_format, int64_t n_wide, float *out, int64_t out_stride,
                              int64_t *n_out, void *stream);
/* Kernel selection switch (testing / A-B measurements): banks of 64 .. 2048 channels, a power of two, run on a kernel
 * with register radix-16 butterflies and compile-time geometry; every other channel count (mixed radix, odd primes up
 * to 31) on the generic shared-memory Stockham kernel.  1 forces the generic kernel for every bank; both compute the
 * same function.  Process-wide; returns the previous setting. */
int gmr1b200_set_chan_generic(int on);
/* workload construction (tests, bench): the inverse direction.  streams [n_streams][stream_stride] complex float at
 * sps x 23.4 kS/s (e.g. from gmr1b200_synth_bursts_tx) are interpolated to the wideband rate, mixed to their
 * channels and summed; white noise for a per-channel Es/N0 of esn0_db (unit-power streams; >= 100: none); the sum is
 * scaled by gain and written as iq_format. */
int gmr1b200_synth_wideband(void *plan, const float *streams, int64_t stream_stride, int64_t stream_len,
                            const int32_t *chan_idx, int n_streams, float esn0_db, float gain, uint64_t seed,
                            void *wide, int iq_format, int64_t n_wide, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GMR1_B200_H */
