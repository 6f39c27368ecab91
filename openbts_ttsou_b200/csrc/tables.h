// tables.h -- the write-once constant tables of the burst-DSP library, resident in HBM.
//
// These are the reference's init-time globals (reference Transceiver/sigProcLib.cpp:39-59:
// cosTable/sinTable, GMSKRotation/GMSKReverseRotation, gMidambles[8], gRACHSequence) plus the GSM
// pulse (Transceiver.cpp:62) and the two resampler filters (radioInterface.cpp:134-138, :230-234).
// They are DATA computed once at btsdsp_create() with the reference's exact arithmetic (SURVEY F6:
// never "more accurately") and read-only afterwards, so any number of streams may share them.
#pragma once
#include "cplx.cuh"

namespace btsdsp {

constexpr int kMaxSps = 4;
constexpr int kTrig = 1024;                 // TABLESIZE, sigProcLib.cpp:36
constexpr int kRxTaps = 961, kRxP = 65, kRxQ = 96, kRxPoly = 15;   // RX: x65/96 through the 961-tap table
constexpr int kTxTaps = 651, kTxP = 96, kTxQ = 65, kTxPoly = 7;    // TX: x96/65 through the 651-tap table
constexpr int kDfeMax = 16;                // largest DFE filter the generic designDFE / equalizeBurst accept
constexpr int kSincGrid = 512;              // peakDetect resolves TOA to 1/512 symbol (9 halvings)

constexpr int kRssiMin = -700, kRssiCount = 1544;   // floats span RSSI -691 .. +838 (|amp| from 3.4e38 down to 1.4e-45)
struct DevTables {
  int sps;
  int pulse_len;
  float sinT[kTrig + 4];                    // [1025] = 0: the element the reference reads times delta == 0
  float cosT[kTrig + 4];
  cf rot[157 * kMaxSps];                    // GMSKRotation
  cf revrot[157 * kMaxSps];                 // GMSKReverseRotation
  cf pulse[2 * kMaxSps + 1];                // generateGSMPulse(2, sps), real-only
  cf mid_seq[8][16 * kMaxSps];              // gMidambles[t]->sequence
  float mid_toa[8];
  cf mid_gain[8];
  cf rach_seq[41 * kMaxSps];                // gRACHSequence->sequence
  float rach_toa;
  cf rach_gain;
  float lpf_rx[kRxTaps + 3];                // createLPF(.,961,65)
  float lpf_tx[kTxTaps + 1];                // createLPF(.,651,96)
  float rx_poly[kRxP][16];                  // rx_poly[br][k] = lpf_rx[br + 65 k], zero past the end
  float tx_poly[kTxP][8];                   // tx_poly[br][k] = lpf_tx[br + 96 k]
  // sinc(pi*(m - j/512)) for m = -10..10 (index m+10), j = 0..511: every fractional delay the path
  // itself produces lies on the 1/512 grid (peakDetect's early-late search), so these 21-tap rows
  // are the only interpolators the batched kernels ever need.  Row pitch 24 floats (96 B).
  alignas(16) float sinc_grid[kSincGrid][24];       // 16-byte aligned rows: the kernels fetch a row as six 16-byte loads
  // exp(-n), n = 0..1023, as the host's libm rounds it: the adaptive energy threshold of the caller policy
  // (Transceiver.cpp:355) is a double that must evolve exactly as on the CPU
  double exp_neg[1024];
  // RSSI = (int) floor(20.0*log10(9450.0/|amp|)) (Transceiver.cpp:400) as the HOST's libm evaluates it, for every float
  // |amp|: rssi_thr[k] = the largest float a whose RSSI is still >= kRssiMin + k (found by bisection over float bit
  // patterns with the host's own log10), so RSSI(a) = kRssiMin + #{k >= 1 : a <= rssi_thr[k]}.  The device's log10 is
  // not bit-identical to glibc's; a one-ulp disagreement at a floor boundary would flip a datagram byte.
  float rssi_thr[kRssiCount];
};

}  // namespace btsdsp
