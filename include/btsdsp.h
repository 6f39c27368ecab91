/*
 * btsdsp.h -- C ABI of the B200-native OpenBTS transceiver burst-DSP library (libbtsdsp.so).
 *
 * This is the drop-in boundary for the hot path of the reference's software transceiver: the free
 * functions of Transceiver/sigProcLib.h:101-384 (called from Transceiver.cpp:62-424 and
 * radioInterface.cpp:137-245).  Three layers, all plain pointers and sizes (no CUDA or torch types):
 *
 *   1. single-vector calls with HOST pointers, synchronous -- one per sigProcLib.h function, same
 *      argument meaning and error behaviour; openbts_ttsou_b200/host/sigProcLib.h re-wraps them in the
 *      reference's C++ signatures (signalVector/BitVector/SoftVector) so Transceiver.cpp /
 *      radioInterface.cpp link unchanged;
 *   2. batched calls with DEVICE pointers and a CUDA stream (passed as void*), thousands of
 *      (ARFCN, timeslot) bursts or resampler chunks per launch -- the throughput path;
 *   3. batched calls with HOST pointers that run layer 2 behind pinned staging buffers and overlapped
 *      copies -- what a radioInterface/Transceiver replacement (or bench.py's e2e leg) calls.
 *
 * Data: complex samples are interleaved float pairs {re, im} (== the reference's Complex<float>,
 * Complex.h:39-44, == CUDA float2); bits are one byte per bit with the value in bit 0 (BitVector);
 * soft bits are floats in [0,1] (SoftVector), hard bit = soft > 0.5F (BitVector.h:415-420).
 * All functions return BTSDSP_OK (0) or a negative error; btsdsp_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device btsdsp_create() fails.
 * Contexts are per device.  Layer-1 and layer-3 calls use the context's own staging buffers and streams: one such call
 * at a time per context (host/sigProcLib.cpp serialises them with a mutex).  Layer-2 calls may be issued from several
 * host threads on DISTINCT streams and then overlap on the device: every caller stream gets its own scratch set
 * (calls on one stream are ordered by the stream itself).
 */
#ifndef BTSDSP_H
#define BTSDSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden; these are its exports */
#endif

typedef struct btsdsp_ctx btsdsp_ctx;
typedef struct { float re, im; } btsdsp_cf32;

enum {
  BTSDSP_OK = 0,
  BTSDSP_EINVAL = -1,       /* bad argument (null pointer, size out of range, sps not supported here) */
  BTSDSP_ECUDA = -2,        /* CUDA runtime error; text in btsdsp_last_error() */
  BTSDSP_ENOMEM = -3,
  BTSDSP_EUNSUPPORTED = -4  /* e.g. the DFE path at sps != 1 (undefined in the reference too) */
};

/* ConvType, reference sigProcLib.h:40-47 */
enum { BTSDSP_FULL_SPAN = 0, BTSDSP_OVERLAP_ONLY = 1, BTSDSP_START_ONLY = 2, BTSDSP_WITH_TAIL = 3, BTSDSP_NO_DELAY = 4 };

/* table ids for btsdsp_get_table (parity inspection of the init-time globals, sigProcLib.cpp:39-59) */
enum {
  BTSDSP_T_COS = 0, BTSDSP_T_SIN = 1,            /* 1025 floats */
  BTSDSP_T_ROT = 2, BTSDSP_T_REVROT = 3,         /* 157*sps complex: GMSKRotation / GMSKReverseRotation */
  BTSDSP_T_PULSE = 4,                            /* 2*sps+1 complex: generateGSMPulse(2, sps) */
  BTSDSP_T_MID_SEQ = 5, BTSDSP_T_MID_META = 6,   /* idx = TSC: 16*sps complex ; {TOA, gain.re, gain.im} */
  BTSDSP_T_RACH_SEQ = 7, BTSDSP_T_RACH_META = 8, /* 41*sps complex ; {TOA, gain.re, gain.im} */
  BTSDSP_T_LPF_RX = 9, BTSDSP_T_LPF_TX = 10      /* 961 / 651 floats: createLPF(.,961,65) / createLPF(.,651,96) */
};

/* ---- lifetime.  Replaces sigProcLibSetup + generateGSMPulse + generateMidamble x8 +
 * generateRACHSequence + the two createLPF calls (sigProcLib.cpp:227,411,779,830,1102;
 * Transceiver.cpp:62-64,424,553; radioInterface.cpp:134-138,230-234).  sps in {1,2,4}. */
int btsdsp_create(btsdsp_ctx **ctx, int device, int sps);
int btsdsp_destroy(btsdsp_ctx *ctx);                       /* sigProcLibDestroy, sigProcLib.cpp:61 */
const char *btsdsp_last_error(const btsdsp_ctx *ctx);      /* ctx may be NULL: error of the last failed create */
int btsdsp_version(void);
int btsdsp_device(const btsdsp_ctx *ctx);
int btsdsp_sps(const btsdsp_ctx *ctx);
/* copies table `id` (entry idx for the per-TSC ones) to dst; returns the number of floats, or <0 */
int btsdsp_get_table(btsdsp_ctx *ctx, int id, int idx, float *dst, int cap);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long btsdsp_launch_count(const btsdsp_ctx *ctx);
int btsdsp_synchronize(btsdsp_ctx *ctx);
/* measurement aid: when enabled, btsdsp_demod_normal_dev brackets its two kernels (detect + DFE design, streaming
 * equaliser) with CUDA events on the caller's stream; btsdsp_get_timing returns the last call's durations in ms */
int btsdsp_set_timing(btsdsp_ctx *ctx, int enable);
int btsdsp_get_timing(btsdsp_ctx *ctx, float *detect_ms, float *equalize_ms);
/* measurement aid: when enabled, the host-buffer pipelines (btsdsp_rx_stream_host / _wire_host) issue exactly their
 * copies -- same staging buffers, segments, streams and events -- and launch no kernel: the copy roofline of the call */
int btsdsp_set_copy_only(btsdsp_ctx *ctx, int enable);

/* CUDA graphs for callers with small batches (one frame of 1024 ARFCN x 8 TS is two short kernels: launch overhead and
 * the host's call path are a visible part of it).  Between begin and end, layer-2 calls on `stream` (not the default
 * stream) are RECORDED instead of run; graph_launch replays the whole sequence -- same pointers, same sizes -- with one
 * launch.  The context's scratch buffers are pinned while a graph lives: issue the largest call once before capturing,
 * or a later, larger call fails with BTSDSP_EINVAL instead of moving a buffer the graph refers to.  The timing aid
 * (btsdsp_set_timing) must be off during capture. */
typedef struct btsdsp_graph btsdsp_graph;
int btsdsp_graph_begin(btsdsp_ctx *ctx, void *stream, btsdsp_graph **graph);
int btsdsp_graph_end(btsdsp_ctx *ctx, void *stream, btsdsp_graph *graph);
int btsdsp_graph_launch(btsdsp_ctx *ctx, btsdsp_graph *graph, void *stream);
int btsdsp_graph_destroy(btsdsp_ctx *ctx, btsdsp_graph *graph);

/* ---- layer 1: single vectors, HOST pointers, synchronous ------------------------------------- */
/* convolve / correlate, sigProcLib.cpp:267 / :474.  Returns the output length (c needs cap >= it). */
int btsdsp_convolve(btsdsp_ctx *ctx, const btsdsp_cf32 *a, int la, int a_real, const btsdsp_cf32 *b, int lb,
                    int b_real, btsdsp_cf32 *c, int cap, int span_type);
int btsdsp_correlate(btsdsp_ctx *ctx, const btsdsp_cf32 *a, int la, int a_real, const btsdsp_cf32 *b, int lb,
                     int b_real, btsdsp_cf32 *c, int cap, int span_type);
/* scaleVector :713 (in place) */
int btsdsp_scale_vector(btsdsp_ctx *ctx, btsdsp_cf32 *v, int n, int real_only, btsdsp_cf32 scale);
/* delayVector :573 (in place) */
int btsdsp_delay_vector(btsdsp_ctx *ctx, btsdsp_cf32 *v, int n, float delay);
/* peakDetect :663 ; interpolatePoint :639 ; energyDetect :916 */
int btsdsp_peak_detect(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, btsdsp_cf32 *peak, float *peak_index,
                       float *avg_power);
int btsdsp_interpolate_point(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, float ix, btsdsp_cf32 *out);
int btsdsp_energy_detect(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, unsigned window, float threshold,
                         float *avg_power, int *detected);
/* modulateBurst :521 with the library's GSM pulse; out gets sps*(nbits+guard) samples; returns that length */
int btsdsp_modulate_burst(btsdsp_ctx *ctx, const uint8_t *bits, int nbits, int guard, btsdsp_cf32 *out, int cap);
/* analyzeTrafficBurst :935.  chan (6*sps) / chan_offset are written only when request_channel && *detected. */
int btsdsp_analyze_traffic_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, unsigned tsc, float threshold,
                                 btsdsp_cf32 *amplitude, float *toa, int request_channel, btsdsp_cf32 *chan,
                                 float *chan_offset, int *detected);
/* detectRACHBurst :860 */
int btsdsp_detect_rach_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, float threshold,
                             btsdsp_cf32 *amplitude, float *toa, int *detected);
/* designDFE :1246 (w: nf taps, b: nchan-1 taps) */
int btsdsp_design_dfe(btsdsp_ctx *ctx, const btsdsp_cf32 *chan, int nchan, float snr, int nf, btsdsp_cf32 *w,
                      btsdsp_cf32 *b);
/* equalizeBurst :1343; like the reference it leaves the delayed burst in `burst`; soft gets n values. sps 1, nw 7, nb 5 */
int btsdsp_equalize_burst(btsdsp_ctx *ctx, btsdsp_cf32 *burst, int n, float toa, const btsdsp_cf32 *w, int nw,
                          const btsdsp_cf32 *b, int nb, float *soft);
/* demodulateBurst :1056; soft gets n/sps values; returns that count */
int btsdsp_demodulate_burst(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, btsdsp_cf32 channel, float toa,
                            float *soft);
/* polyphaseResampleVector :1157 with the library's filters: lpf 0 = RX (961 taps), 1 = TX (651 taps).
 * Returns the output length ceil(n*P/Q). */
int btsdsp_polyphase_resample(btsdsp_ctx *ctx, const btsdsp_cf32 *x, int n, int P, int Q, int lpf, btsdsp_cf32 *out,
                              int cap);

/* The same with the caller's own filter (any length): ntaps complex taps; lpf_real != 0 uses only their real parts
 * (the reference's realOnly branch, :1194-1200), 0 the complex branch (:1187-1193). */
int btsdsp_polyphase_resample_taps(btsdsp_ctx *ctx, const btsdsp_cf32 *x, int n, int P, int Q, const btsdsp_cf32 *lpf,
                                   int ntaps, int lpf_real, btsdsp_cf32 *out, int cap);
/* createLPF :1102-1150 (its cutoff argument is ignored there): filter_len == 651 -> the 651-tap prototype; any other
 * filter_len <= 961 -> a 961-entry vector whose first filter_len entries are the 961-tap prototype's; all scaled by
 * float(gain_dc / sum of the copied taps).  Returns the vector length (651 or 961).  filter_len > 961 overruns the
 * reference's vector and table: BTSDSP_EINVAL. */
int btsdsp_create_lpf(btsdsp_ctx *ctx, int filter_len, float gain_dc, float *taps, int cap);
/* the element-wise helpers: addVector :746 (x += y over the shorter length), offsetVector :760, conjugateVector :733,
 * vectorSlicer :507, GMSKRotate / GMSKReverseRotate :232-264 (n <= 157*sps) -- in place on x -- and vectorNorm2 :146
 * (*result = sum |x|^2 in the reference's order) */
enum { BTSDSP_VOP_ADD = 0, BTSDSP_VOP_OFFSET = 1, BTSDSP_VOP_CONJ = 2, BTSDSP_VOP_SLICE = 3, BTSDSP_VOP_NORM2 = 4,
       BTSDSP_VOP_ROTATE = 5, BTSDSP_VOP_REVROTATE = 6 };
int btsdsp_vector_op(btsdsp_ctx *ctx, int op, btsdsp_cf32 *x, int n, int real_only, const btsdsp_cf32 *y, int ny,
                     btsdsp_cf32 scalar, float *result);

/* ---- layer 2: batched, DEVICE pointers, asynchronous on `stream` (a cudaStream_t; NULL = default) ---- */
/* Burst addressing shared by the batched receive calls:
 *   pitch > 0 : burst i starts at bursts + i*pitch samples; its length is lens[i], or when lens == NULL
 *               (157 if (first+i)%4==0 else 156)*sps -- the reference's slot rule, radioInterface.cpp:375-378
 *   pitch == 0: `bursts` is a continuous slot stream beginning on a 157-sample slot; burst g = first+i
 *               starts at ((g/4)*625 + {0,157,313,469}[g%4])*sps.                                     */

/* GMSK modulate n bursts of nbits bits (n x nbits bytes).  guard < 0: 8 + ((first+i)%4==0) per burst
 * (Transceiver.cpp:105-106).  pitch as above (pitch == 0 writes a continuous slot stream). */
int btsdsp_modulate_dev(btsdsp_ctx *ctx, const uint8_t *bits, int nbits, long long n, int guard, long long first,
                        btsdsp_cf32 *out, long long pitch, void *stream);
/* RX resampler (radioInterface.cpp:238-259): nchunks chunks of 864 raw samples -> 585 each.
 * has_history != 0: raw[-192..-1] hold the previous samples; 0: the stream starts here (zeros). */
int btsdsp_resample_rx_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, int has_history, long long nchunks,
                           btsdsp_cf32 *out, void *stream);
/* The same from the radio's own sample format: interleaved int16 {I,Q} (4 B per sample), converted as
 * unUSRPifyVector does (radioInterface.cpp:91-116); swap_iq != 0 for the Q-first order of real USRP hardware,
 * 0 for the SWLOOPBACK order.  iq must be 16-byte aligned; has_history: iq[-384..-1] (192 samples) are valid. */
int btsdsp_resample_rx_i16_dev(btsdsp_ctx *ctx, const int16_t *iq, int swap_iq, int has_history, long long nchunks,
                               btsdsp_cf32 *out, void *stream);
/* The same for nstreams radios in ONE launch: stream a at iq + 2*a*iq_pitch int16 (iq_pitch samples, a multiple of 4),
 * its 585*nchunks outputs at out + a*out_pitch (a multiple of 2).  Every stream has the same nchunks / history flag. */
int btsdsp_resample_rx_i16_streams_dev(btsdsp_ctx *ctx, const int16_t *iq, long long iq_pitch, int nstreams, int swap_iq,
                                       int has_history, long long nchunks, btsdsp_cf32 *out, long long out_pitch,
                                       void *stream);
/* TX resampler (radioInterface.cpp:123-168 + USRPifyVector :74-89): nchunks chunks of 585 samples ->
 * 864 int16 {I,Q} pairs each, scaled by 13500 and truncated like the reference's (short) cast.  out must be 4-byte aligned. */
int btsdsp_resample_tx_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *in, int has_history, long long nchunks, int16_t *out,
                           void *stream);
/* Normal bursts, fused energy gate -> analyzeTrafficBurst(request channel) -> designDFE(Nf 7) ->
 * equalizeBurst, the TSC branch of Transceiver::pullRadioVector (Transceiver.cpp:298-396) with a channel
 * estimate per burst.  gate_thr < 0 disables the energy gate; snr_thr is the energy threshold in the
 * SNR estimate (:340).  Outputs per burst (any may be NULL): flag, amp, toa, soft[soft_pitch] (zeros when
 * not detected) and for inspection chan[6] (after 1/amp), chan_off, w[7], b[5].  sps must be 1. */
int btsdsp_demod_normal_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                            long long first, const uint8_t *tsc, long long n, float detect_thr, float gate_thr,
                            float snr_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch,
                            btsdsp_cf32 *chan, float *chan_off, btsdsp_cf32 *w, btsdsp_cf32 *b, void *stream);
/* The same with the soft bits in the RX datagram's wire format: (char) round(soft*255.0), 148 bytes per burst
 * (Transceiver::driveReceiveFIFO, Transceiver.cpp:668-670); rows soft_pitch_bytes apart (multiple of 4, >= 148). */
int btsdsp_demod_normal_u8_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                               long long first, const uint8_t *tsc, long long n, float detect_thr, float gate_thr,
                               float snr_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, uint8_t *soft_u8,
                               int soft_pitch_bytes, void *stream);
/* analyzeTrafficBurst alone (any sps). chan: 6*sps per burst. */
int btsdsp_analyze_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                       long long first, const uint8_t *tsc, long long n, float detect_thr, int request_channel,
                       int32_t *flag, btsdsp_cf32 *amp, float *toa, btsdsp_cf32 *chan, float *chan_off, void *stream);
/* Access bursts: detectRACHBurst, then demodulateBurst when soft != NULL (Transceiver.cpp:360-389). */
int btsdsp_rach_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long first,
                    long long n, float detect_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft,
                    int soft_pitch, void *stream);
/* designDFE for n channel estimates (chan n x 6, snr n) -> w n x 7, b n x 5 */
int btsdsp_design_dfe_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *chan, const float *snr, long long n, btsdsp_cf32 *w,
                          btsdsp_cf32 *b, void *stream);
/* equalizeBurst with supplied taps (the cached-DFE mode, Transceiver.cpp:391-396); burst_out may be NULL */
int btsdsp_equalize_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                        long long first, long long n, const float *toa, const btsdsp_cf32 *w, const btsdsp_cf32 *b,
                        float *soft, int soft_pitch, btsdsp_cf32 *burst_out, long long out_pitch, void *stream);
/* demodulateBurst alone with supplied amp/toa (any sps) */
int btsdsp_demodulate_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                          long long first, long long n, const btsdsp_cf32 *amp, const float *toa, float *soft,
                          int soft_pitch, void *stream);

/* ---- layer 3: batched, HOST pointers, synchronous (copies overlapped with kernels internally) ---- */
/* The north-star receive path end to end: raw 400 kS/s complex stream (nchunks x 864 samples, starting at
 * stream time 0) -> RX resample -> 157/156/156/156 slot cutting -> fused normal-burst demod of the first
 * nbursts slots (needs 625*nbursts/4 <= 585*nchunks).  tsc: one byte per burst.  Outputs as in
 * btsdsp_demod_normal_dev.  Host buffers may be pageable or pinned (pinned is faster).
 * One-shot: every call starts a stream (zero resampler history, as RadioInterface does at start-up); a receiver that
 * feeds a running radio call after call uses btsdsp_trx_radio_host, which carries the 192-sample history. */
int btsdsp_rx_stream_host(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, long long nchunks, const uint8_t *tsc,
                          long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                          btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch);
/* The same pipeline at the reference's wire formats on both sides: int16 {I,Q} samples in (what USRPDevice::readSamples
 * delivers, radioInterface.cpp:213-227), 148 soft bytes per burst out (what goes into the UDP datagram,
 * Transceiver.cpp:659-674).  Half the H2D bytes and a quarter of the D2H bytes of the float call. */
int btsdsp_rx_stream_wire_host(btsdsp_ctx *ctx, const int16_t *iq, int swap_iq, long long nchunks, const uint8_t *tsc,
                               long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                               btsdsp_cf32 *amp, float *toa, uint8_t *soft_u8);
/* The same on DEVICE-resident input and outputs (no copies): resample into an internal buffer, then demod. */
int btsdsp_rx_stream_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, long long nchunks, const uint8_t *tsc,
                         long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                         btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch, void *stream);
/* The same for a piece of a RUNNING stream: has_history != 0 says raw[-192..-1] hold the stream's previous 192 samples, the
 * state RadioInterface::pullBuffer carries from call to call (rcvHistory, radioInterface.cpp:238-259).  A piece must begin
 * on a 117-frame boundary of the stream (a multiple of 250 chunks = 936 bursts from its start), so that slot 0 of the piece
 * is a 157-sample slot.  Pieces processed call after call give exactly the one-call result. */
int btsdsp_rx_stream_cont_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *raw, int has_history, long long nchunks, const uint8_t *tsc,
                              long long nbursts, float detect_thr, float gate_thr, float snr_thr, int32_t *flag,
                              btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch, void *stream);
/* The transmit path: n x 148 bits -> modulateBurst (guard 8/9 by slot) -> slot stream -> TX resample ->
 * int16 {I,Q}.  n must be a multiple of 4 and 625*n/4 a multiple of 585 (n % 468 == 0), giving
 * 864*(625*n/4/585) output pairs.  Host pointers. */
int btsdsp_tx_stream_host(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int16_t *out);
int btsdsp_tx_stream_dev(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int16_t *out, void *stream);
/* The same for nstreams radios (one per ARFCN, Transceiver.cpp:412-426 runs one TX chain per Transceiver object) in ONE
 * launch: stream a's n bursts at bits148 + a*n*148, its output at out + a*2*864*(625*n/4/585) int16.  Every stream
 * starts from zero resampler history, as each RadioInterface does. */
int btsdsp_tx_streams_dev(btsdsp_ctx *ctx, const uint8_t *bits148, long long n, int nstreams, int16_t *out, void *stream);
/* Batched normal-burst demod / RACH detect+demod over HOST buffers (pitched bursts). */
int btsdsp_demod_normal_host(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens,
                             const uint8_t *tsc, long long n, float detect_thr, float gate_thr, float snr_thr,
                             int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch,
                             btsdsp_cf32 *chan, float *chan_off, btsdsp_cf32 *w, btsdsp_cf32 *b);
int btsdsp_rach_host(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long n,
                     float detect_thr, int32_t *flag, btsdsp_cf32 *amp, float *toa, float *soft, int soft_pitch);

/* pinned host memory helpers for layer-3 callers */
/* ---- caller policy: Transceiver::pullRadioVector + driveReceiveFIFO over batches ----
 * A btsdsp_trx holds, on the device, the receive-side state of `narfcn` Transceiver objects (one per ARFCN):
 * adaptive mEnergyThreshold, prevFalseDetectionTime, and per timeslot the cached channel estimate's DFE, offset,
 * SNR and timestamp (Transceiver.h:127-142; initial values Transceiver.cpp:72-89).  tsc: mTSC per ARFCN;
 * chan_type: mChanType per (ARFCN, TN), 0 NONE, 1..7 = I..VII, 8 LOOPBACK (Transceiver.h ChannelCombination).
 *
 * pull: one batch of nframes x narfcn x 8 received slots, laid out [frame][arfcn][tn] at `pitch` samples
 * (slot length 157 if tn%4==0 else 156), FN = fn0 + frame, processed with exactly the reference's burst-by-burst
 * semantics per ARFCN (Transceiver.cpp:271-410): slot map (expectedCorrType :207-269), energy gate against the
 * adaptive threshold and its updates (:298-304, :338-339, :353-357, :369-376), channel re-estimation every 50 frames
 * or after a miss (:315-350), equalisation with the cached DFE (:391-396) or RACH demodulation (:383-389).
 * Per burst: valid[i] (a datagram would be sent) and the 158-byte RX datagram of driveReceiveFIFO (:659-673) at
 * dgram + i*dgram_pitch: TN, FN (4 bytes, big endian), RSSI, timing offset (2 bytes, 1/256 symbol), 148 soft bytes,
 * 2 zero bytes.  Rows of invalid bursts are zero.  State carries over to the next pull. */
typedef struct btsdsp_trx btsdsp_trx;
int btsdsp_trx_create(btsdsp_ctx *ctx, int narfcn, const uint8_t *tsc, const uint8_t *chan_type, int start_fn,
                      btsdsp_trx **out);
int btsdsp_trx_destroy(btsdsp_ctx *ctx, btsdsp_trx *trx);
int btsdsp_trx_set_slot(btsdsp_ctx *ctx, btsdsp_trx *trx, int arfcn, int tn, int chan_type);   /* SETSLOT, Transceiver.cpp:549 */
/* state of one ARFCN for inspection: { double thr; int32 prev_false_fn, tsc, chan_type[8], est_fn[8], have[8];
 * float snr[8], chan_off[8]; cf32 w[8][7], b[8][5]; }  (btsdsp_trx_state_bytes() bytes) */
int btsdsp_trx_state_bytes(void);
int btsdsp_trx_get_state(btsdsp_ctx *ctx, btsdsp_trx *trx, int arfcn, void *dst, int cap);
/* device pointers, asynchronous on `stream`; dgram 8-byte aligned, dgram_pitch >= 160 and a multiple of 4 */
int btsdsp_trx_pull_dev(btsdsp_ctx *ctx, btsdsp_trx *trx, const btsdsp_cf32 *bursts, long long pitch, int nframes,
                        int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch, void *stream);
/* the same over narfcn continuous slot streams (what the RX resampler writes), stream_pitch samples apart: stream a holds
 * the slots of frames fn0.. back to back (157/156/156/156 samples), nframes*1250 samples each */
int btsdsp_trx_pull_streams_dev(btsdsp_ctx *ctx, btsdsp_trx *trx, const btsdsp_cf32 *streams, long long stream_pitch,
                                int nframes, int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch, void *stream);
/* Radio samples in, datagrams out (host pointers, synchronous): for each of the narfcn radios, nchunks chunks of 864
 * int16 {I,Q} samples (stream a at iq + 2*a*iq_pitch int16) -> RadioInterface::pullBuffer (unUSRPifyVector + 65/96
 * resample with the running 192-sample history, radioInterface.cpp:197-273) -> slot cutting (:370-394) -> the pull above.
 * nchunks % 250 == 0 (250 chunks = 117 frames exactly); the first call starts the stream (zero history) at FN fn0, TN 0,
 * later calls continue it.  Outputs laid out [frame][arfcn][tn], nchunks/250*117 frames. */
int btsdsp_trx_radio_host(btsdsp_ctx *ctx, btsdsp_trx *trx, const int16_t *iq, long long iq_pitch, long long nchunks,
                          int swap_iq, int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch);
/* The datagram's RSSI for n amplitude magnitudes (host pointers): (int) floor(20.0*log10(9450.0/|amp|)),
 * Transceiver.cpp:400, exactly as the HOST's libm evaluates it for every float (a threshold table built at create;
 * the device's own log10 is not bit-identical to glibc's at the floor boundaries). */
int btsdsp_trx_rssi(btsdsp_ctx *ctx, const float *abs_amp, int n, int32_t *rssi);
/* host pointers, synchronous; dgram_pitch >= 158 */
int btsdsp_trx_pull_host(btsdsp_ctx *ctx, btsdsp_trx *trx, const btsdsp_cf32 *bursts, long long pitch, int nframes,
                         int fn0, int32_t *valid, uint8_t *dgram, int dgram_pitch);

/* TX datagrams (GSM core -> transceiver, 154 bytes: TN, FN as 4 bytes big endian, power attenuation in dB, 148 bits;
 * Transceiver::driveTransmitPriorityQueue, Transceiver.cpp:582-632), batched: every burst is modulated with guard
 * 8 + (TN%4==0) and scaled by pow(10, -RSSI/10) (addRadioVector, :100-114), placed at slot (FN - fn0)*8 + TN of a
 * slot stream of nframes frames (nframes % 117 == 0: a whole number of 585-sample chunks), slots without a datagram
 * carry `filler` (148 bits, unscaled, e.g. the dummy burst; NULL = silence), then RadioInterface::pushBuffer's
 * 96/65 resample, x13500 and int16 conversion (radioInterface.cpp:123-168).  out: nframes*1250/585 chunks x 864
 * int16 {I,Q} pairs.  *placed (optional) = datagrams inside the window (TN 0..7, fn0 <= FN < fn0+nframes); a later
 * datagram for the same slot replaces an earlier one. */
int btsdsp_tx_datagrams_host(btsdsp_ctx *ctx, const uint8_t *dgram, long long n, int dgram_pitch, int fn0, int nframes,
                             const uint8_t *filler, int16_t *out, long long *placed);

/* ---- L1 FEC after the path: the XCCH block decoder (SACCH / SDCCH / CCCH-type channels,
 * GSM 05.03 4.1) as XCCHL1Decoder runs it (GSML1FEC.cpp:616-660): nframes L2 frames, each from the soft bytes of four
 * consecutive bursts (148 bytes each, burst_pitch apart, as in the RX datagram: probability = byte / 256.0F,
 * TRXManager.cpp:230) -> deinterleave -> soft-input Viterbi (SoftVector::decode / ViterbiR2O4, BitVector.cpp:290-540)
 * -> u[228] = d[184] : p[40] : tail[4] hard bits per frame, ok[f] = Fire-code syndrome of d : ~p is zero. ---- */
int btsdsp_xcch_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nframes, uint8_t *u,
                           int32_t *ok, void *stream);
int btsdsp_xcch_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nframes, uint8_t *u,
                            int32_t *ok);

/* TCH/FACCH block decoder (GSM 05.03 3.1 / 4.2; TCHFACCHL1Decoder::processBurst + deinterleave + decodeTCH + decode,
 * GSML1FEC.cpp:1031-1210) over the soft bytes of ONE traffic channel's consecutive traffic bursts (the 26-multiframe's
 * SACCH / idle frames already taken out by the caller's demultiplexer): 4*nblocks + 4 bursts, burst_pitch apart; block q
 * is diagonally interleaved over bursts 4q .. 4q+7 and completes with burst 4q+7, whose stealing flag Hl (bit 60) says
 * what it is.  stolen[q] == 0: a speech frame -- d[q][260] = the class-1 (Viterbi-decoded, reordered) and class-2 (sliced)
 * bits of GSM 05.03 3.1.2, good[q] = the 3-bit class-1A parity matches and the tail bits are zero.  stolen[q] != 0: a
 * FACCH frame -- facch_u[q][228] and facch_ok[q] exactly as btsdsp_xcch_decode gives them; good[q] = 0 (the reference feeds
 * bad-frame substitution then).  Rows of the other kind are zero.  Any output may be NULL.  GSM 06.10 frame formatting
 * and bad-frame substitution (:1161-1190) are the speech layer's. */
int btsdsp_tch_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nblocks, uint8_t *d, int32_t *good,
                          int32_t *stolen, uint8_t *facch_u, int32_t *facch_ok, void *stream);
int btsdsp_tch_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long nblocks, uint8_t *d, int32_t *good,
                           int32_t *stolen, uint8_t *facch_u, int32_t *facch_ok);

/* ---- L1 encoders on the TRANSMIT side: the producers of the 148-bit bursts modulateBurst is called with (the callers on the
 * other side of the path).  Bits are one byte each, value in bit 0; bursts are 148 bytes, burst_pitch (>= 148) apart; tsc = 0..7
 * writes that training sequence at bits 61..86 (gTrainingSequence[mTSC].copyToSegment(mBurst, 61), GSML1FEC.cpp:740), -1 leaves
 * them zero; tails are zero.  lsb8msb != 0 applies BitVector::LSB8MSB to the 184-bit frame first, as sendFrame does (:781).
 *
 * btsdsp_xcch_encode: XCCHL1Encoder::sendFrame = encode + interleave + transmit (GSML1FEC.cpp:763-850; GSM 05.03 4.1): frame
 * d[184] -> inverted 40-bit Fire-code word (Parity 0x10004820009) -> four tail zeros -> rate-1/2 K = 5 convolutional code
 * (BitVector::encode, BitVector.cpp:217-238) -> 456 bits block-interleaved over FOUR bursts, both stealing flags set
 * (:735-736).  frames: nframes x 184; bursts: 4 * nframes. */
int btsdsp_xcch_encode_dev(btsdsp_ctx *ctx, const uint8_t *frames, long long nframes, int lsb8msb, int tsc, uint8_t *bursts,
                           int burst_pitch, void *stream);
int btsdsp_xcch_encode_host(btsdsp_ctx *ctx, const uint8_t *frames, long long nframes, int lsb8msb, int tsc, uint8_t *bursts,
                            int burst_pitch);
/* btsdsp_tch_encode: TCHFACCHL1Encoder::encodeTCH / dispatch / interleave (GSML1FEC.cpp:1248-1392; GSM 05.03 3.1, 4.2) over ONE
 * traffic channel's block stream.  Block q is a speech frame d260[q] (260 bits in class order, i.e. after the g610BitOrder map of
 * :1255: 3-bit class-1A parity, reordering, four tail zeros, class 1 coded, class 2 appended) or, when steal[q] != 0, the FACCH
 * frame f184[q] coded like an XCCH block; its 456 bits are diagonally interleaved: the even e-bits and Hu (bit 87) of bursts
 * 4q .. 4q+3, the odd e-bits and Hl (bit 60) of bursts 4q+4 .. 4q+7.  bursts: 4 * nblocks + 4 -- the last four are half filled
 * (what the reference keeps in mI[] / mPreviousFACCH for its next dispatch); pass them as `carry` to the next call, which
 * completes them in its first four bursts.  carry == NULL: the channel starts here (zero-filled interleaver, :1219-1224).
 * d260, f184 and steal must all be valid (rows of the kind a block is not are never read).  nblocks == 0 (device entry point) just
 * completes the carry: four bursts out.  The reference's idle filler pattern (:1347-1350) is a constant of the caller. */
int btsdsp_tch_encode_dev(btsdsp_ctx *ctx, const uint8_t *d260, const uint8_t *f184, const uint8_t *steal, long long nblocks, int lsb8msb,
                          int tsc, const uint8_t *carry, uint8_t *bursts, int burst_pitch, void *stream);
int btsdsp_tch_encode_host(btsdsp_ctx *ctx, const uint8_t *d260, const uint8_t *f184, const uint8_t *steal, long long nblocks, int lsb8msb,
                           int tsc, const uint8_t *carry, uint8_t *bursts, int burst_pitch);

/* RACH block decoder (GSM 05.03 4.6; RACHL1Decoder::writeLowSide, GSML1FEC.cpp:474-515): per access burst, the 36
 * coded soft bytes at burst bits 49..84 -> Viterbi -> u[18] = d[8] : p[6] : tail[4] (u may be NULL), and
 * fields[i] = tail | bsic << 8 | ra << 16: tail = the 4 tail bits (a valid burst has 0), bsic = (~sent parity ^ computed
 * parity) & 0x3f (a valid burst matches the cell's BSIC), ra = the 8-bit RA value handed to the control layer. */
int btsdsp_rach_decode_dev(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long n, uint8_t *u,
                           int32_t *fields, void *stream);
int btsdsp_rach_decode_host(btsdsp_ctx *ctx, const uint8_t *soft_u8, int burst_pitch, long long n, uint8_t *u,
                            int32_t *fields);

/* ---- the reference's second transceiver variant (Transceiver52M/sigProcLib.cpp): the functions whose
 * arithmetic differs from the main variant.  analyzeTrafficBurst there searches only +-max_toa symbols around the
 * expected midamble position (convolve's CUSTOM span, :966-1077; max_toa < 3*sps is raised to 3*sps, max_toa <= 60) and
 * counts TOA from the window centre; energyDetect strides its window by four samples (:944-963, needs 4*(window-1) < n).
 * Everything else of that variant (modulate, delay, RACH, demodulate, DFE, resample) computes what the main entry points
 * compute.  chan: 6*sps per burst. ---- */
int btsdsp_analyze_52m_dev(btsdsp_ctx *ctx, const btsdsp_cf32 *bursts, long long pitch, const int32_t *lens, long long first,
                           const uint8_t *tsc, long long n, float detect_thr, unsigned max_toa, int request_channel,
                           int32_t *flag, btsdsp_cf32 *amp, float *toa, btsdsp_cf32 *chan, float *chan_off, void *stream);
int btsdsp_analyze_traffic_burst_52m(btsdsp_ctx *ctx, const btsdsp_cf32 *burst, int n, unsigned tsc, float threshold,
                                     unsigned max_toa, int request_channel, int *detected, btsdsp_cf32 *amp, float *toa,
                                     btsdsp_cf32 *chan, float *chan_off);
int btsdsp_energy_detect_52m(btsdsp_ctx *ctx, const btsdsp_cf32 *v, int n, unsigned window, float threshold, float *avg_pwr,
                             int *above);

/* The receive policy of that variant (Transceiver52M/Transceiver.cpp:268-404) on a btsdsp_trx: enable != 0 makes every later
 * pull use the stride-4 energy window and the windowed midamble search over +-max_expected_delay symbols (its
 * mMaxExpectedDelay, 0..60); with max_expected_delay <= 1 the channel is never estimated and detected normal bursts are
 * demodulated by demodulateBurst instead of the equaliser (needDFE == false, :272, :322, :382).  enable == 0 restores the
 * main variant.  State layout, slot map, threshold adaptation and datagrams are the same. */
int btsdsp_trx_set_variant_52m(btsdsp_ctx *ctx, btsdsp_trx *trx, int enable, int max_expected_delay);
/* Its transmit side: bursts are scaled at modulate time -- datagram bursts by 13500 * pow(10, -RSSI/10)
 * (Transceiver52M/Transceiver.cpp:111), `filler` slots by 13500 (:74; the filler table's initial content -- replaying the
 * last burst sent on a slot, :145-175, is queue logic left to the caller) -- and the symbol-rate radio only casts to short
 * (radioInterface.cpp:100-118): out gets nframes*1250 int16 {I,Q} pairs, any nframes > 0. */
int btsdsp_tx_datagrams_52m_host(btsdsp_ctx *ctx, const uint8_t *dgram, long long n, int dgram_pitch, int fn0, int nframes,
                                 const uint8_t *filler, int16_t *out, long long *placed);

void *btsdsp_host_alloc(size_t bytes);
void btsdsp_host_free(void *p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BTSDSP_H */
