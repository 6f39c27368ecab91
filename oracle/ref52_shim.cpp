/* ref52_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" face over the reference's SECOND transceiver variant,
 * Transceiver52M/sigProcLib.cpp, compiled in place like oracle/ref_shim.cpp (oracle/Makefile target
 * `ref52` -> oracle/_ref/libref52_oracle.so; it is a separate library because both variants define the same symbols).
 * Only what differs from the main variant is exposed: the windowed (CUSTOM-span) analyzeTrafficBurst with maxTOA
 * (Transceiver52M/sigProcLib.cpp:966-1077) and energyDetect's stride-4 window (:944-963); plus that variant's caller
 * policy (Transceiver52M/Transceiver.cpp:268-404: needDFE / mMaxExpectedDelay) and its modulate-time TX scaling (:74,111)
 * with the symbol-rate radio's short casts (Transceiver52M/radioInterface.cpp:100-118).  Transceiver.cpp cannot be compiled
 * here (libusrp headers), so its glue is restated; every DSP call is the real function of Transceiver52M/sigProcLib.cpp. */
#include "sigProcLib.h"
#include "GSMCommon.h"
#include <math.h>
#include <string.h>

extern "C" {

static signalVector *gPulse52 = NULL;
static int gSps52 = 0;

int ref52_setup(int sps) {
  if (gSps52 == sps) return 0;
  if (gSps52) { sigProcLibDestroy(); delete gPulse52; }
  sigProcLibSetup(sps);
  gPulse52 = generateGSMPulse(2, sps);
  for (int t = 0; t < 8; t++) generateMidamble(*gPulse52, sps, t);
  generateRACHSequence(*gPulse52, sps);
  gSps52 = sps;
  return 0;
}

/* burst: n complex64; chan: 6*sps complex64 (written when the return is 1 and request != 0) */
int ref52_analyze(const float *burst, int n, int tsc, float thr, int sps, unsigned maxTOA, int request,
                  float *amp, float *toa, float *chan, float *off) {
  signalVector b(n);
  memcpy(b.begin(), burst, n * sizeof(complex));
  complex a = 0.0;
  float t = 0.0F, o = 0.0F;
  signalVector *ch = NULL;
  bool ok = analyzeTrafficBurst(b, tsc, thr, sps, &a, &t, maxTOA, request != 0, &ch, &o);
  amp[0] = a.real(); amp[1] = a.imag(); *toa = t;
  if (ch) {
    for (unsigned i = 0; i < ch->size(); i++) { chan[2 * i] = (*ch)[i].real(); chan[2 * i + 1] = (*ch)[i].imag(); }
    *off = o;
    delete ch;
  }
  return ok;
}

int ref52_energy_detect(const float *v, int n, unsigned win, float thr, float *avg) {
  signalVector b(n);
  memcpy(b.begin(), v, n * sizeof(complex));
  return energyDetect(b, win, thr, avg);
}

/* ---- Transceiver::pullRadioVector + driveReceiveFIFO of the second variant (Transceiver52M/Transceiver.cpp:268-404,
 * 655-690).  Same state layout as oracle/ref_shim.cpp's ref_trx_state; bursts of ONE ARFCN in FIFO order. ---- */
struct ref52_trx_state {
  double thr;
  int prev_false_fn;
  int tsc;
  int chan_type[8];
  int est_fn[8];
  int have[8];
  float snr[8];
  float chan_off[8];
  float w[8][14];
  float b[8][10];
};
enum { CT_NONE = 0, CT_I, CT_II, CT_III, CT_IV, CT_V, CT_VI, CT_VII, CT_LOOPBACK };
enum { CORR_OFF = 0, CORR_TSC, CORR_RACH, CORR_IDLE };
static const int kHyper52 = 2048 * 26 * 51;
static int fn_delta52(int v1, int v2) {                   /* GSM::FNDelta, GSMCommon.cpp:161-168 */
  const int half = kHyper52 / 2;
  int d = v1 - v2;
  if (d >= half) d -= kHyper52; else if (d < -half) d += kHyper52;
  return d;
}
static int corr_type52(int chan_type, int fn) {           /* Transceiver52M/Transceiver.cpp:204-265 (same table as the main variant) */
  const int m = fn % 51;
  switch (chan_type) {
    case CT_I: case CT_III: return CORR_TSC;
    case CT_II: return (fn % 2 == 1) ? CORR_IDLE : CORR_TSC;
    case CT_IV: case CT_VI: return (m % 10 < 2) ? CORR_RACH : CORR_OFF;
    case CT_V: return ((m <= 36 && m >= 14) || m == 4 || m == 5 || m == 45 || m == 46) ? CORR_RACH : CORR_TSC;
    case CT_VII: return (m == 12 || m == 13 || m == 14) ? CORR_IDLE : CORR_TSC;
    case CT_LOOPBACK: return (m <= 50 && m >= 48) ? CORR_IDLE : CORR_TSC;
    default: return CORR_OFF;
  }
}
int ref52_trx_state_bytes(void) { return (int)sizeof(ref52_trx_state); }
void ref52_trx_init(ref52_trx_state *st, int tsc, const int *chan_type, int start_fn) {
  memset(st, 0, sizeof *st);
  st->thr = 250.0;
  st->prev_false_fn = start_fn;
  st->tsc = tsc;
  for (int i = 0; i < 8; i++) { st->chan_type[i] = chan_type[i]; st->est_fn[i] = start_fn; }
}
void ref52_trx_pull(ref52_trx_state *st, const float *bursts, int pitch, int nframes, int fn0, int max_expected_delay,
                    int *valid, unsigned char *dgram, int dgram_pitch) {
  const bool needDFE = (max_expected_delay > 1);                                          /* :272 */
  for (int f = 0; f < nframes; f++) {
    const int fn = (fn0 + f) % kHyper52;
    for (int tn = 0; tn < 8; tn++) {
      const long i = (long)f * 8 + tn;
      valid[i] = 0;
      unsigned char *dg = dgram + (size_t)dgram_pitch * i;
      memset(dg, 0, 158);
      const int corr = corr_type52(st->chan_type[tn], fn);
      if (corr == CORR_OFF || corr == CORR_IDLE) continue;
      const int len = (tn % 4 == 0) ? 157 : 156;
      signalVector burst(len);
      memcpy(burst.begin(), bursts + 2 * (size_t)pitch * i, len * sizeof(complex));
      complex amplitude = 0.0;
      float TOA = 0.0F, avgPwr = 0.0F;
      if (!energyDetect(burst, 20, st->thr, &avgPwr)) {                                   /* :293 */
        double framesElapsed = fn_delta52(fn, st->prev_false_fn);
        if (framesElapsed > 50) { st->thr -= 10.0; st->prev_false_fn = fn; }
        continue;
      }
      bool success = false;
      if (corr == CORR_TSC) {
        double framesElapsed = fn_delta52(fn, st->est_fn[tn]);
        bool estimateChannel = false;
        if (framesElapsed > 50 || !st->have[tn]) { st->have[tn] = 0; estimateChannel = true; }   /* :311-320 */
        if (!needDFE) estimateChannel = false;                                            /* :322 */
        signalVector *channelResp = NULL;
        float chanOffset = 0.0F;
        success = analyzeTrafficBurst(burst, st->tsc, 3.0, 1, &amplitude, &TOA, max_expected_delay, estimateChannel,
                                      &channelResp, &chanOffset);                         /* :324-333 */
        if (success) {
          st->thr -= 1.0F;
          if (st->thr < 0.0) st->thr = 0.0;
          st->snr[tn] = amplitude.norm2() / (st->thr * st->thr + 1.0);
          if (estimateChannel) {
            st->have[tn] = 1;
            st->chan_off[tn] = chanOffset;
            scaleVector(*channelResp, complex(1.0, 0.0) / amplitude);
            signalVector *W = NULL, *B = NULL;
            designDFE(*channelResp, st->snr[tn], 7, &W, &B);
            memcpy(st->w[tn], W->begin(), 7 * sizeof(complex));
            memcpy(st->b[tn], B->begin(), 5 * sizeof(complex));
            delete W; delete B;
            st->est_fn[tn] = fn;
          }
        } else {
          double fe = fn_delta52(fn, st->prev_false_fn);
          st->thr += 10.0F * exp(-fe);
          st->prev_false_fn = fn;
          st->have[tn] = 0;
        }
        if (channelResp) delete channelResp;
      } else {
        success = detectRACHBurst(burst, 5.0, 1, &amplitude, &TOA);
        if (success) {
          st->thr -= 1.0F;
          if (st->thr < 0.0) st->thr = 0.0;
          st->have[tn] = 0;
        } else {
          double fe = fn_delta52(fn, st->prev_false_fn);
          st->thr += 10.0F * exp(-fe);
          st->prev_false_fn = fn;
        }
      }
      if (!success) continue;
      SoftVector *soft;
      if (corr == CORR_RACH || !needDFE) {                                                /* :382 */
        soft = demodulateBurst(burst, *gPulse52, 1, amplitude, TOA);
      } else {
        scaleVector(burst, complex(1.0, 0.0) / amplitude);
        signalVector W(7), B(5);
        memcpy(W.begin(), st->w[tn], 7 * sizeof(complex));
        memcpy(B.begin(), st->b[tn], 5 * sizeof(complex));
        soft = equalizeBurst(burst, TOA - st->chan_off[tn], 1, W, B);
      }
      const int RSSI = (int)floor(20.0 * log10(9450.0 / amplitude.abs()));                /* :396 */
      const int timingOffset = (int)round(TOA * 256.0 / 1);                               /* :398 */
      valid[i] = 1;
      dg[0] = tn;
      for (int k = 0; k < 4; k++) dg[1 + k] = (fn >> ((3 - k) * 8)) & 0x0ff;
      dg[5] = RSSI;
      dg[6] = (timingOffset >> 8) & 0x0ff;
      dg[7] = timingOffset & 0x0ff;
      SoftVector::iterator it = soft->begin();
      for (int k = 0; k < 148; k++) dg[8 + k] = (char)round((*it++) * 255.0);
      delete soft;
    }
  }
}

/* The transmit side of that variant, batched over a window of nframes frames: addRadioVector's modulateBurst +
 * scaleVector(13500.0 * pow(10,-RSSI/10)) (:105-112), filler slots = the constructor's dummy-burst table entries,
 * modulateBurst + scaleVector(13500.0) (:70-78), and the radio's USRPifyVector short casts with powerScaling 1.0
 * (Transceiver52M/radioInterface.cpp:100-118).  out: nframes*1250 {I,Q} pairs. */
long ref52_tx_datagrams(const unsigned char *dgram, long n, int dgram_pitch, int fn0, int nframes,
                        const unsigned char *filler, short *out) {
  const long nslots = (long)nframes * 8, nsamp = nslots / 4 * 625;
  static const int slot_off[4] = {0, 157, 313, 469};
  long *src = new long[nslots];
  for (long s = 0; s < nslots; s++) src[s] = -1;
  long placed = 0;
  for (long i = 0; i < n; i++) {
    const char *buffer = (const char *)(dgram + (size_t)dgram_pitch * i);
    int timeSlot = (int)buffer[0];
    unsigned long frameNum = 0;
    for (int k = 0; k < 4; k++) frameNum = (frameNum << 8) | (0x0ff & buffer[k + 1]);
    if (timeSlot < 0 || timeSlot > 7) continue;
    int f = fn_delta52((int)(frameNum % kHyper52), fn0 % kHyper52);
    if (f < 0 || f >= nframes) continue;
    src[(long)f * 8 + timeSlot] = i;
    placed++;
  }
  memset(out, 0, (size_t)nsamp * 2 * sizeof(short));
  BitVector bv(148);
  for (long sl = 0; sl < nslots; sl++) {
    const int tn = (int)(sl % 8);
    signalVector *m = NULL;
    if (src[sl] >= 0) {
      const char *buffer = (const char *)(dgram + (size_t)dgram_pitch * src[sl]);
      int RSSI = (int)buffer[5];
      memcpy(bv.begin(), buffer + 6, 148);
      m = modulateBurst(bv, *gPulse52, 8 + (tn % 4 == 0), 1);
      scaleVector(*m, 13500.0 * pow(10, -RSSI / 10));
    } else if (filler) {
      memcpy(bv.begin(), filler, 148);
      m = modulateBurst(bv, *gPulse52, 8 + (tn % 4 == 0), 1);
      scaleVector(*m, 13500.0);
    }
    if (m) {
      short *o = out + 2 * ((sl / 4) * 625 + slot_off[sl % 4]);
      for (signalVector::iterator itr = m->begin(); itr < m->end(); itr++) {
        *o++ = (short)itr->real();
        *o++ = (short)itr->imag();
      }
      delete m;
    }
  }
  delete[] src;
  return placed;
}

}  // extern "C"
