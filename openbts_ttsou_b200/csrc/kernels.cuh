// kernels.cuh -- launchers of the sm_100a kernels (definitions in kernels.cu / resample.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tables.h"
#include "trx_policy.cuh"
#include "fec.cuh"

namespace btsdsp {

// Where the bursts of a batch live.  pitch > 0: burst i starts at base + i*pitch (complex samples),
// its length is lens[i] or, when lens is null, the GSM 157/156/156/156 rule on (first + i).
// pitch == 0: base is a continuous slot stream that begins on a 157-sample slot; burst g = first + i
// starts at (g/4)*625*sps + {0,157,313,469}[g%4]*sps  (reference radioInterface.cpp:370-394).
struct BurstSrc {
  const cf *base;
  long long pitch;
  const int *lens;
  long long first;
  int sps;
  const int *gather = nullptr;   // optional: burst i of the call is burst gather[i] of the array (outputs stay compact)
  // optional (pitch == 0 only): narfcn parallel slot streams, arfcn_pitch samples apart, with the bursts of the call
  // numbered [frame][arfcn][tn]: burst i lives in stream (i/8) % narfcn at slot (i / (8 narfcn)) * 8 + i % 8
  int narfcn = 0;
  long long arfcn_pitch = 0;
};

// which of the reference's two transceivers a caller-policy pull follows: the main one (Transceiver/), or the second
// (Transceiver52M/): windowed midamble search over +-max_toa symbols, stride-4 energy window, and -- when need_dfe is false
// (mMaxExpectedDelay <= 1, Transceiver52M/Transceiver.cpp:272) -- no channel estimate and no equaliser
struct TrxVariant {
  bool v52m = false, need_dfe = true;
  unsigned max_toa = 0;
};

struct NormalOut {      // per burst; null pointers are skipped
  int *flag;            // detection flag                                   (analyzeTrafficBurst return)
  cf *amp;              // amplitude estimate
  float *toa;           // TOA estimate, symbols
  cf *chan;             // [6*sps] channel response after the caller's 1/amp scaling (demod) or raw (analyze)
  float *off;           // channelResponseOffset
  cf *w;                // [7] DFE feed-forward taps
  cf *b;                // [5] DFE feedback taps
  float *soft;          // [soft_pitch] soft bits, zeros when not detected
  int soft_pitch;       // floats (soft) or bytes (soft_u8)
  unsigned char *soft_u8 = nullptr;   // alternative output: the RX datagram's bytes, round(soft*255), 148 per burst
};

constexpr int kTileStride = 33;      // complex samples between rows of a transposed tile (32 lanes + 1 pad)
constexpr int kBurstRows = 160;      // >= 157

// -- per-device kernel attributes (dynamic shared memory sizes) and constant-memory tables; call once per device
int configure_kernels();
int configure_resamplers();
void upload_resampler_taps(const DevTables *hostT);

// -- table construction (init only)
void launch_init_tables(DevTables *T, cudaStream_t st);
void launch_modulate_impulse(const DevTables *T, const uint8_t *bits, int nbits, cf *out, cudaStream_t st);

// -- single vectors in global memory (sigProcLib.h surface); one thread or one thread per output
void launch_convolve(const cf *a, int la, int a_real, const cf *b, int lb, int b_real, cf *c, int start, int outsz,
                     int corr, cudaStream_t st);
void launch_peak_detect(const DevTables *T, const cf *v, int n, cf *peak, float *idx, float *avg, cudaStream_t st);
void launch_interp_point(const DevTables *T, const cf *v, int n, float ix, cf *out, cudaStream_t st);
void launch_delay_vector(const DevTables *T, cf *v, int n, float delay, cf *tmp, cudaStream_t st);
void launch_scale_vector(cf *v, int n, int real_only, cf s, cudaStream_t st);
void launch_energy_detect(const cf *v, int n, unsigned win, float thr, float *avg, int *flag, cudaStream_t st);
void launch_rssi(const DevTables *T, const float *a, int n, int *rssi, cudaStream_t st);
enum { VOP_ADD = 0, VOP_OFFSET = 1, VOP_CONJ = 2, VOP_SLICE = 3, VOP_NORM2 = 4, VOP_ROTATE = 5, VOP_REVROTATE = 6 };
void launch_vector_op(const DevTables *T, int op, cf *x, int n, int real_only, const cf *y, int ny, cf s, float *res, cudaStream_t st);
void launch_resample_generic(const cf *x, int n, int P, int Q, const float *lpf, int L, cf *out, int outn,
                             cudaStream_t st);
// the same with the caller's own filter: ctaps = L complex taps, real_taps != 0 uses only their real parts (:1187-1200)
void launch_resample_taps(const cf *x, int n, int P, int Q, const cf *ctaps, int L, int real_taps, cf *out, int outn,
                          cudaStream_t st);
void launch_equalize_generic(const DevTables *T, cf *burst, int n, float toa, const cf *w, int nw, const cf *b, int nb,
                             cf *tmp, float *soft, cudaStream_t st);
void launch_design_dfe_generic(const cf *chan, int nchan, float snr, int nf, cf *w, cf *b, cudaStream_t st);

// -- batched
void launch_modulate(const DevTables *T, const uint8_t *bits, int nbits, long long nbursts, int guard_rule,
                     const uint8_t *guards, long long first, cf *out, long long pitch, cudaStream_t st,
                     const float *scale = nullptr);
void launch_resample_rx(const DevTables *T, const cf *in, int has_history, long long nchunks, cf *out, cudaStream_t st,
                        int max_ctas = 0, int light = 0);
void launch_tx_fused(const DevTables *T, const uint8_t *bits, const float *scale, long long nslots, int16_t *out, cudaStream_t st,
                     int nstreams = 1);
int launch_resample_rx_i16_multi(const int16_t *in, long long in_pitch, int nstreams, int swap_iq, int has_history,
                                 long long nchunks, cf *out, long long out_pitch, cudaStream_t st);
int launch_resample_rx_i16(const int16_t *in, int swap_iq, int has_history, long long nchunks, cf *out, cudaStream_t st);
void launch_resample_tx(const DevTables *T, const cf *in, int has_history, long long nchunks, int16_t *out,
                        cudaStream_t st);
size_t demod_scratch_bytes(long long n);   // per-launch scratch of launch_demod_normal (EqParams records)
size_t rach_scratch_bytes(long long n);    // per-launch scratch of launch_rach's tuned path (records + correlation rows)
int launch_demod_normal(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr,
                        float gate_thr, float snr_thr, NormalOut out, void *scratch, cudaStream_t st,
                        cudaEvent_t between = nullptr);
int launch_analyze(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr, int request,
                   NormalOut out, cf *scratch, int force_generic, cudaStream_t st);
int launch_rach(const DevTables *T, BurstSrc src, long long n, float detect_thr, int demod, NormalOut out, cf *scratch,
                int force_generic, cudaStream_t st, void *eq_scratch = nullptr);
void upload_rach_taps(const DevTables *hostT);
int launch_equalize(const DevTables *T, BurstSrc src, long long n, const float *toa, const cf *w, const cf *b,
                    float *soft, int soft_pitch, cf *burst_out, long long out_pitch, cudaStream_t st);
int launch_demodulate(const DevTables *T, BurstSrc src, long long n, const cf *amp, const float *toa, float *soft,
                      int soft_pitch, cf *scratch, cudaStream_t st);
int launch_design_dfe(const cf *chan, const float *snr, long long n, cf *w, cf *b, cudaStream_t st);
// the second transceiver variant (Transceiver52M): windowed analyzeTrafficBurst, stride-4 energyDetect
int analyze_52m_scratch_stride(unsigned max_toa, int sps);
int launch_analyze_52m(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr, unsigned max_toa,
                       int request, NormalOut out, cf *scratch, cudaStream_t st);
void launch_energy_detect_52m(const cf *v, int n, unsigned win, float thr, float *avg, int *flag, cudaStream_t st);
// L1 FEC after the path (fec.cuh / fec_kernels.cuh)
int launch_xcch_decode(const unsigned char *soft, int burst_pitch, long long nframes, unsigned char *u, int *ok, cudaStream_t st);
int launch_rach_decode(const unsigned char *soft, int burst_pitch, long long n, unsigned char *u, int *fields, cudaStream_t st);
int launch_tch_decode(const unsigned char *soft, int burst_pitch, long long nblocks, unsigned char *d, int *good, int *stolen,
                      unsigned char *fu, int *fok, cudaStream_t st);
// L1 encoders on the transmit side (the producers of modulateBurst's bits)
int launch_xcch_encode(const unsigned char *frames, long long nframes, int lsb8msb, unsigned tsc_word, int have_tsc, unsigned char *bursts,
                       int burst_pitch, cudaStream_t st);
int launch_tch_encode(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks, int lsb8msb,
                      unsigned tsc_word, int have_tsc, const unsigned char *carry, unsigned char *bursts, int burst_pitch, cudaStream_t st);
// the caller-policy pipeline (trx_policy.cuh / trx_kernels.cuh)
size_t trx_scratch_bytes(long long n, long long nr, int narfcn, bool slice_all = false);
void launch_usrpify(const cf *x, long long n, int16_t *out, cudaStream_t st);
int launch_trx_pull(const DevTables *T, TrxState *st, int narfcn, int nframes, int fn0, const cf *bursts, long long pitch,
                    long long stream_pitch, const uint8_t *kind, const uint8_t *tsc, const int *rach_idx, const int *rach_slot,
                    long long nr, void *scratch, int *valid, unsigned char *dgram, int dgram_pitch, cudaStream_t stream,
                    cudaStream_t side = nullptr, cudaEvent_t *ev = nullptr, TrxVariant var = TrxVariant());

// scratch (complex samples) the generic-sps global-memory variants need per burst
__host__ __device__ inline size_t scratch_per_burst(int sps) { return (size_t)(2 * 157 + 36) * sps + 64; }

}  // namespace btsdsp
