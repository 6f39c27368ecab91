// tables_host.h -- the host-side part of table construction: what the reference computes with libm in
// double (trig tables, GSM pulse) and the resampler filter scaling.  Everything that is float
// arithmetic over these tables (rotation tables, sinc grid, midamble/RACH peaks) is built on the
// device by capi.cu::build_tables through the product's own kernels.
#pragma once
#include <math.h>
#include <string.h>
#include "tables.h"
#include "lpf_taps.inc"

namespace btsdsp {

static const char *const kTSC[8] = {  // GSM 05.02 5.2.3 (the reference's gTrainingSequence, GSM/GSMCommon.cpp:44-53)
    "00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
    "01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
    "10100111110110001010011111", "11101111000100101110111100"};
static const char *const kRACH = "01001011011111111001100110101010001111000";  // GSM 05.02 5.2.7 (gRACHSynchSequence, :57)

// createLPF, sigProcLib.cpp:1102-1150: table copy, double sum, one float scale factor
inline void create_lpf(const unsigned int *bits, int len, float gainDC, float *out) {
  double sum = 0.0;
  for (int i = 0; i < len; i++) { float t; memcpy(&t, &bits[i], 4); out[i] = t; sum += t; }
  const float norm = (float)(gainDC / sum);
  for (int i = 0; i < len; i++) out[i] = out[i] * norm;
}

inline void host_fill_tables(DevTables *h, int sps) {
  memset(h, 0, sizeof(DevTables));
  h->sps = sps;
  for (int i = 0; i < kTrig + 1; i++) {                       // initTrigTables :207-212
    h->cosT[i] = (float)cos(2.0 * M_PI * i / kTrig);
    h->sinT[i] = (float)sin(2.0 * M_PI * i / kTrig);
  }
  h->cosT[kTrig + 1] = 0.0F;   // what the reference multiplies by delta == 0 when arg == 1 (finite, so harmless)
  h->sinT[kTrig + 1] = 0.0F;
  h->pulse_len = 2 * sps + 1;                                 // generateGSMPulse(2, sps) :411-430
  const int center = (h->pulse_len - 1) / 2;
  float e = 0.0F;
  for (int i = 0; i < h->pulse_len; i++) {
    const float arg = (float)(i - center) / (float)sps;
    h->pulse[i] = mk((float)(0.96 * exp(-1.1380 * arg * arg - 0.527 * arg * arg * arg * arg)), 0.0F);
  }
  for (int i = 0; i < h->pulse_len; i++) e += h->pulse[i].y * h->pulse[i].y + h->pulse[i].x * h->pulse[i].x;
  const float avg = sqrtf(e / sps);
  for (int i = 0; i < h->pulse_len; i++) { h->pulse[i].x /= avg; h->pulse[i].y /= avg; }
  create_lpf(LPF961_BITS, kRxTaps, (float)kRxP, h->lpf_rx);
  create_lpf(LPF651_BITS, kTxTaps, (float)kTxP, h->lpf_tx);
  for (int br = 0; br < kRxP; br++)
    for (int k = 0; k < 16; k++) h->rx_poly[br][k] = (br + kRxP * k < kRxTaps) ? h->lpf_rx[br + kRxP * k] : 0.0F;
  for (int br = 0; br < kTxP; br++)
    for (int k = 0; k < 8; k++) h->tx_poly[br][k] = (br + kTxP * k < kTxTaps) ? h->lpf_tx[br + kTxP * k] : 0.0F;
  for (int n = 0; n < 1024; n++) h->exp_neg[n] = exp(-(double)n);   // Transceiver.cpp:355 exp(-framesElapsed), host libm
  // rssi_thr[k]: largest positive finite float a with floor(20*log10(9450/a)) >= kRssiMin + k (non-increasing in a)
  for (int k = 0; k < kRssiCount; k++) {
    const int r = kRssiMin + k;
    unsigned lo = 0u, hi = 0x7f7fffffu;                              // bit patterns: 0 (RSSI = +inf) .. FLT_MAX
    auto rssi_of = [](unsigned bits) { float a; memcpy(&a, &bits, 4); return floor(20.0 * log10(9450.0 / (double)a)); };
    if (!(rssi_of(hi) < (double)r)) lo = hi;                         // every float qualifies
    while (hi - lo > 1u) {                                           // invariant: rssi_of(lo) >= r > rssi_of(hi)
      const unsigned mid = lo + (hi - lo) / 2u;
      if (rssi_of(mid) >= (double)r) lo = mid; else hi = mid;
    }
    memcpy(&h->rssi_thr[k], &lo, 4);
  }
}

}  // namespace btsdsp
