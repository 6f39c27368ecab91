// trx_policy.cuh -- the caller policy around the receive path, Transceiver::pullRadioVector and driveReceiveFIFO
// (Transceiver.cpp:207-269, 271-410, 641-676), as data-parallel passes plus ONE sequential scalar pass:
//
//   pass 1 (parallel, every burst)    energy of the first 20 samples; for TSC slots analyzeTrafficBurst with the channel
//                                     response, for RACH slots detectRACHBurst.  None of it depends on the adaptive
//                                     threshold, so it runs for the whole batch at once.
//   pass 2 (sequential per ARFCN)     this file: the reference's scalar state machine -- energy gate against the adaptive
//                                     mEnergyThreshold, its +-updates, the 50-frame channel/DFE cache, the SNR estimate --
//                                     walked in FIFO order by one thread per ARFCN over pass 1's results.  It decides,
//                                     per burst, what pass 3 does: nothing / demodulate a RACH burst / equalise with the
//                                     DFE designed from burst j (j == i: this burst re-estimates) or with the DFE carried
//                                     over from the previous batch.
//   pass 3 (parallel)                 designDFE for the re-estimating bursts, equalizeBurst / demodulateBurst for the
//                                     accepted ones, RSSI / timing integerisation and the RX datagram.
//
// One TrxState = the receive-side members of one Transceiver object (one ARFCN); it lives on the device between calls.
// Everything here is __host__ __device__ so tests/hostemu replays it on the CPU.
#pragma once
#include <math.h>

#include "cplx.cuh"

namespace btsdsp {

constexpr int kHyperframe = 2048 * 26 * 51;                     // GSMCommon.h:306
enum ChanType { CT_NONE = 0, CT_I, CT_II, CT_III, CT_IV, CT_V, CT_VI, CT_VII, CT_LOOPBACK };   // Transceiver.h ChannelCombination
enum CorrType { CORR_OFF = 0, CORR_TSC = 1, CORR_RACH = 2, CORR_IDLE = 3 };                      // Transceiver.h CorrType
enum TrxAct { ACT_SLICE = -4, ACT_NONE = -3, ACT_CARRIED = -2, ACT_RACH = -1 };   // >= 0: equalise with the DFE designed from that burst
// ACT_SLICE: a normal burst demodulated without the equaliser (the second variant's short-range mode, see need_dfe below)
constexpr int kExpTable = 1024;                                   // exp(-n), n = 0..1023, from the host's libm

struct TrxState {                       // layout is part of the ABI (btsdsp_trx_get_state)
  double thr;                           // mEnergyThreshold (Transceiver.h:132)
  int prev_false_fn;                    // prevFalseDetectionTime
  int tsc;                              // mTSC
  int chan_type[8];                     // mChanType
  int est_fn[8];                        // channelEstimateTime
  int have[8];                          // channelResponse[tn] != NULL
  float snr[8];                         // SNRestimate
  float chan_off[8];                    // chanRespOffset
  cf w[8][7];                           // DFEForward
  cf b[8][5];                           // DFEFeedback
};

struct __align__(16) DetRec {           // pass 1's result for one burst, 80 B
  float energy;                         // energyDetect's avgPwr
  float flag;                           // analyzeTrafficBurst / detectRACHBurst return (0/1)
  float amp_x, amp_y, toa, off;
  float pad0, pad1;
  cf chan[6];                           // channel response / gain, BEFORE the caller's 1/amplitude scaling
};

// GSM::FNDelta, GSMCommon.cpp:161-168
BTS_HD int fn_delta(int v1, int v2) {
  const int half = kHyperframe / 2;
  int d = v1 - v2;
  if (d >= half) d -= kHyperframe;
  else if (d < -half) d += kHyperframe;
  return d;
}

// Transceiver::expectedCorrType, Transceiver.cpp:207-269
BTS_HD int expected_corr_type(int chan_type, int fn) {
  const int m = fn % 51;
  switch (chan_type) {
    case CT_NONE: return CORR_OFF;
    case CT_I: return CORR_TSC;
    case CT_II: return (fn % 2 == 1) ? CORR_IDLE : CORR_TSC;
    case CT_III: return CORR_TSC;
    case CT_IV:
    case CT_VI: return (m % 10 < 2) ? CORR_RACH : CORR_OFF;
    case CT_V:
      if ((m <= 36 && m >= 14) || m == 4 || m == 5 || m == 45 || m == 46) return CORR_RACH;
      return CORR_TSC;
    case CT_VII: return (m == 12 || m == 13 || m == 14) ? CORR_IDLE : CORR_TSC;
    case CT_LOOPBACK: return (m <= 50 && m >= 48) ? CORR_IDLE : CORR_TSC;
    default: return CORR_OFF;
  }
}

// exp(-framesElapsed) as the reference's libm computes it (:355, :375): the table holds the host's exp(-n); beyond it
// the value is below the smallest double.  A negative argument cannot occur for bursts presented in time order.
BTS_HD double exp_neg_frames(const double *__restrict__ table, int frames) {
  if (frames < 0) return exp(-(double)frames);
  return frames < kExpTable ? table[frames] : 0.0;
}

// The scalar part of a TrxState: what pass 2 reads and writes.  Kept apart so that the walk holds it in registers (the
// timeslot loop is unrolled, every index is static) instead of dragging the 600-byte state through local memory.
struct TrxScalars {
  double thr;
  int prev_false_fn;
  int chan_type[8], est_fn[8], have[8];
  float snr[8], chan_off[8];
};
BTS_HD void trx_load_scalars(const TrxState &st, TrxScalars &s) {
  s.thr = st.thr; s.prev_false_fn = st.prev_false_fn;
#pragma unroll
  for (int tn = 0; tn < 8; tn++) {
    s.chan_type[tn] = st.chan_type[tn]; s.est_fn[tn] = st.est_fn[tn]; s.have[tn] = st.have[tn];
    s.snr[tn] = st.snr[tn]; s.chan_off[tn] = st.chan_off[tn];
  }
}
// chan_off is NOT stored back here: it belongs to the cached DFE and is committed with it (k_trx_commit)
BTS_HD void trx_store_scalars(TrxState &st, const TrxScalars &s) {
  st.thr = s.thr; st.prev_false_fn = s.prev_false_fn;
#pragma unroll
  for (int tn = 0; tn < 8; tn++) { st.est_fn[tn] = s.est_fn[tn]; st.have[tn] = s.have[tn]; st.snr[tn] = s.snr[tn]; }
}

// Pass 2 for one ARFCN `a` of a batch laid out [frame][arfcn][tn] (burst i = (f*narfcn + a)*8 + tn, FN = fn0 + f).
// rach_slot[i] = index of burst i among the batch's RACH-slot bursts (their detection results are compact), or -1.
// Outputs: act[i], thr_at[i] (the threshold right after this burst's update, valid where act[i] == i: pass 3 turns it
// into the SNR estimate of :340 -- the double division is kept OUT of this serial walk, where its latency would be paid
// 256 times per thread); commit[tn] = the burst whose DFE is the cache entry after the batch (>= 0), ACT_CARRIED
// (unchanged) -- the caller copies it into st.w/b/chan_off once pass 3 has designed it.
// A frame's eight records are fetched up front (independent loads), then walked in TN order.
BTS_HD float trx_snr_estimate(cf amp, double thr) { return (float)((double)cnorm2(amp) / (thr * thr + 1.0)); }   // :340
// need_dfe = false is the second transceiver variant with mMaxExpectedDelay <= 1 (Transceiver52M/Transceiver.cpp:272): the
// channel is never estimated (:322) and detected normal bursts go through demodulateBurst like access bursts (:382).
BTS_HD void trx_policy_arfcn(TrxScalars &st, int nframes, int fn0, int narfcn, int a, const DetRec *__restrict__ det,
                             const int *__restrict__ rach_slot, const int *__restrict__ rach_flag,
                             const double *__restrict__ exp_table, int *__restrict__ act, double *__restrict__ thr_at,
                             int *__restrict__ commit, bool need_dfe = true) {
  int src[8];
  float lax[8], lay[8];                     // amplitude and threshold of the last detected TSC burst per timeslot:
  double lthr[8];                           // SNRestimate[tn] of the state is computed from them once, at the end
  bool lhave[8];
#pragma unroll
  for (int tn = 0; tn < 8; tn++) { src[tn] = ACT_CARRIED; lhave[tn] = false; lax[tn] = lay[tn] = 0.0F; lthr[tn] = 0.0; }
  for (int f = 0; f < nframes; f++) {
    const int fn = (fn0 + f) % kHyperframe;
    const long long i0 = ((long long)f * narfcn + a) * 8;
    float en[8], fl[8], ax[8], ay[8], of[8];
    int rs[8], ac[8];
#pragma unroll
    for (int tn = 0; tn < 8; tn++) {
      const DetRec &d = det[i0 + tn];
      en[tn] = d.energy; fl[tn] = d.flag; ax[tn] = d.amp_x; ay[tn] = d.amp_y; of[tn] = d.off;
      rs[tn] = rach_slot[i0 + tn];
    }
#pragma unroll
    for (int tn = 0; tn < 8; tn++) {
      const long long i = i0 + tn;
      ac[tn] = ACT_NONE;
      const int corr = expected_corr_type(st.chan_type[tn], fn);
      if (corr == CORR_OFF || corr == CORR_IDLE) continue;                              // :290-293
      const float thrf = (float)st.thr;                                                 // energyDetect takes a float
      if (!(en[tn] > BTS_MUL(thrf, thrf))) {                                            // :298, sigProcLib.cpp:931
        if ((double)fn_delta(fn, st.prev_false_fn) > 50) { st.thr -= 10.0; st.prev_false_fn = fn; }   // :300-304
        continue;
      }
      bool success;
      if (corr == CORR_TSC) {
        bool estimate = false;
        if ((double)fn_delta(fn, st.est_fn[tn]) > 50 || !st.have[tn]) { st.have[tn] = 0; estimate = true; }   // :315-326
        if (!need_dfe) estimate = false;                                                 // Transceiver52M/Transceiver.cpp:322
        success = fl[tn] != 0.0F;
        if (success) {
          st.thr -= 1.0F;                                                               // :338-339
          if (st.thr < 0.0) st.thr = 0.0;
          lax[tn] = ax[tn]; lay[tn] = ay[tn]; lthr[tn] = st.thr; lhave[tn] = true;      // :340, evaluated later
          if (estimate) {                                                               // :341-350
            st.have[tn] = 1;
            st.chan_off[tn] = of[tn];
            st.est_fn[tn] = fn;
            src[tn] = (int)i;
            thr_at[i] = st.thr;
          }
          ac[tn] = need_dfe ? src[tn] : (int)ACT_SLICE;
        } else {
          st.thr += 10.0F * exp_neg_frames(exp_table, fn_delta(fn, st.prev_false_fn));  // :353-355
          st.prev_false_fn = fn;
          st.have[tn] = 0;                                                              // :357
        }
      } else {
        success = rach_flag[rs[tn]] != 0;
        if (success) {
          st.thr -= 1.0F;                                                               // :369-371
          if (st.thr < 0.0) st.thr = 0.0;
          st.have[tn] = 0;
          ac[tn] = ACT_RACH;
        } else {
          st.thr += 10.0F * exp_neg_frames(exp_table, fn_delta(fn, st.prev_false_fn));  // :374-376
          st.prev_false_fn = fn;
        }
      }
    }
#pragma unroll
    for (int tn = 0; tn < 8; tn++) act[i0 + tn] = ac[tn];
  }
#pragma unroll
  for (int tn = 0; tn < 8; tn++) {
    commit[tn] = src[tn];
    if (lhave[tn]) st.snr[tn] = trx_snr_estimate(mk(lax[tn], lay[tn]), lthr[tn]);
  }
}

// RX datagram header, Transceiver.cpp:400-402 and :659-666.  dg[0..7]
// RSSI = (int) floor(20.0*log10(9450.0/amplitude.abs())) (:400) through the host-built threshold table (tables.h): a
// binary search for the number of thresholds >= |amp|.  |amp| = 0, Inf or NaN falls outside the table; those keep the
// formula (the reference's own result there is an out-of-range int conversion).
BTS_HD int trx_rssi(const DevTables *__restrict__ T, float a) {
  if (!(a > 0.0F) || !(a <= 3.402823466e38F)) return (int)floor(20.0 * log10(9450.0 / (double)a));
  int lo = 0, hi = kRssiCount;                    // invariant: a <= thr[lo] (thr[0] = FLT_MAX), a > thr[hi] (virtual end)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a <= T->rssi_thr[mid]) lo = mid; else hi = mid;
  }
  return kRssiMin + lo;
}
BTS_HD void trx_datagram_header(const DevTables *__restrict__ T, unsigned char *dg, int tn, int fn, cf amp, float toa, int sps) {
  const int rssi = trx_rssi(T, cabs_(amp));                                             // :400
  const int timing = (int)round((double)toa * 256.0 / sps);                             // :402
  dg[0] = (unsigned char)tn;
  for (int k = 0; k < 4; k++) dg[1 + k] = (unsigned char)((fn >> ((3 - k) * 8)) & 0xff);
  dg[5] = (unsigned char)rssi;
  dg[6] = (unsigned char)((timing >> 8) & 0xff);
  dg[7] = (unsigned char)(timing & 0xff);
}
// soft bit -> datagram byte, Transceiver.cpp:669
BTS_HD unsigned char trx_soft_byte(float s) { return (unsigned char)((int)round((double)s * 255.0) & 0xff); }

}  // namespace btsdsp
