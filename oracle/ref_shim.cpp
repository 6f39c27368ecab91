/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" face over the UNMODIFIED reference sources, compiled where they lie under
 * /root/reference (see oracle/Makefile; nothing from the reference is copied into this repo).
 * The result, oracle/_ref/libref_oracle.so, is "the reference itself run here": it pins the plain-C
 * restatement (oracle/sigproc_port.c), generates tests/golden/*, checks the CUDA path in
 * tests/ and is timed as bench.py's CPU baseline (cpu_baseline.kind == "reference").
 *
 * Every entry point has the same name/signature as its twin in sigproc_port.c, with prefix
 * ref_ instead of port_.  Complex vectors are interleaved float pairs (re, im) == Complex<float>
 * (reference Transceiver/Complex.h:39-44), bits are one char per bit, soft bits one float per bit.
 *
 * Functions wrapped (reference file:line):
 *   sigProcLibSetup sigProcLib.cpp:227, generateGSMPulse :411, generateMidamble :779,
 *   generateRACHSequence :830, modulateBurst :521, delayVector :573, convolve :267, correlate :474,
 *   peakDetect :663, interpolatePoint :639, energyDetect :916, analyzeTrafficBurst :935,
 *   detectRACHBurst :860, designDFE :1246, equalizeBurst :1343, demodulateBurst :1056,
 *   createLPF :1102, polyphaseResampleVector :1157, sinc :567, sinLookup/cosLookup :163-188.
 * Caller glue restated here because the callers need libusrp headers to compile:
 *   Transceiver::pullRadioVector TSC/RACH branches  Transceiver.cpp:327-396   (rx_normal / rx_rach)
 *   RadioInterface::pullBuffer                      radioInterface.cpp:238-259 (rx_resample_stream)
 *   RadioInterface::pushBuffer + USRPifyVector      radioInterface.cpp:74-89,123-194 (tx_resample_stream)
 *   RadioInterface::driveReceiveRadio slot cutting  radioInterface.cpp:370-394
 */
#include "sigProcLib.h"
#include "GSMCommon.h"
#include <string.h>
#include <math.h>

typedef struct {           // mirrors the file-local typedef at sigProcLib.cpp:52-56
  signalVector *sequence;
  float TOA;
  complex gain;
} CorrelationSequence;

extern CorrelationSequence *gMidambles[];
extern CorrelationSequence *gRACHSequence;
extern signalVector *GMSKRotation, *GMSKReverseRotation;
extern float cosTable[], sinTable[];
extern float sendLPF_961[], rcvLPF_651[];
float cosLookup(const float x);
float sinLookup(const float x);

static signalVector *gPulse = NULL;
static signalVector *gLpfRx = NULL;   // createLPF(.,961,65)  radioInterface.cpp:230-234
static signalVector *gLpfTx = NULL;   // createLPF(.,651,96)  radioInterface.cpp:134-138
static int gSps = 0;

static void put(const signalVector &v, float *dst) { memcpy(dst, v.begin(), v.size() * sizeof(complex)); }

extern "C" {

int ref_setup(int sps) {
  if (gSps) {
    sigProcLibDestroy();
    delete gPulse; delete gLpfRx; delete gLpfTx;
  }
  gSps = sps;
  gPulse = generateGSMPulse(2, sps);                 // Transceiver.cpp:62
  sigProcLibSetup(sps);                              // Transceiver.cpp:64
  for (int t = 0; t < 8; t++) generateMidamble(*gPulse, sps, t);   // Transceiver.cpp:553
  generateRACHSequence(*gPulse, sps);                // Transceiver.cpp:424
  gLpfRx = createLPF(1.0F / 96.0F, 961, 65);
  gLpfTx = createLPF(1.0F / 96.0F, 651, 96);
  return 0;
}

/* table ids shared with sigproc_port.c and the product's btsdsp_get_table */
int ref_get_table(int id, int idx, float *dst, int cap) {
  int n = 0;
  switch (id) {
    case 0: n = 1025; if (cap >= n) memcpy(dst, cosTable, n * 4); break;
    case 1: n = 1025; if (cap >= n) memcpy(dst, sinTable, n * 4); break;
    case 2: n = 2 * GMSKRotation->size(); if (cap >= n) put(*GMSKRotation, dst); break;
    case 3: n = 2 * GMSKReverseRotation->size(); if (cap >= n) put(*GMSKReverseRotation, dst); break;
    case 4: n = 2 * gPulse->size(); if (cap >= n) put(*gPulse, dst); break;
    case 5: n = 2 * gMidambles[idx]->sequence->size(); if (cap >= n) put(*gMidambles[idx]->sequence, dst); break;
    case 6: n = 3; if (cap >= n) { dst[0] = gMidambles[idx]->TOA; dst[1] = gMidambles[idx]->gain.real(); dst[2] = gMidambles[idx]->gain.imag(); } break;
    case 7: n = 2 * gRACHSequence->sequence->size(); if (cap >= n) put(*gRACHSequence->sequence, dst); break;
    case 8: n = 3; if (cap >= n) { dst[0] = gRACHSequence->TOA; dst[1] = gRACHSequence->gain.real(); dst[2] = gRACHSequence->gain.imag(); } break;
    case 9: n = 961; if (cap >= n) for (int i = 0; i < n; i++) dst[i] = (*gLpfRx)[i].real(); break;
    case 10: n = 651; if (cap >= n) for (int i = 0; i < n; i++) dst[i] = (*gLpfTx)[i].real(); break;
    case 11: {  // the 961-entry read of a 960-entry table; report what this build saw
      n = 1; if (cap >= n) { volatile float *p = sendLPF_961; dst[0] = p[960]; } break; }
    case 12: n = 960; if (cap >= n) memcpy(dst, sendLPF_961, n * 4); break;   // raw prototype tables
    case 13: n = 651; if (cap >= n) memcpy(dst, rcvLPF_651, n * 4); break;
    default: return -1;
  }
  return n;
}

float ref_sinc(float x) { return sinc(x); }
float ref_sin_lookup(float x) { return sinLookup(x); }
float ref_cos_lookup(float x) { return cosLookup(x); }

int ref_modulate(const char *bits, int nbits, int guard, int sps, float *out, int cap) {
  BitVector bv(nbits);
  memcpy(bv.begin(), bits, nbits);
  signalVector *m = modulateBurst(bv, *gPulse, guard, sps);
  int n = m->size();
  if (cap >= n) put(*m, out);
  delete m;
  return n;
}

void ref_delay_vector(float *v, int n, float delay) {
  signalVector sv((complex *)v, 0, n);
  delayVector(sv, delay);
}

void ref_scale_vector(float *v, int n, int real_only, const float *scale) {
  signalVector sv((complex *)v, 0, n);
  sv.isRealOnly(real_only);
  scaleVector(sv, complex(scale[0], scale[1]));
}

static int conv_or_corr(bool corr, const float *a, int la, int a_real, const float *b, int lb, int b_real,
                        float *c, int cap, int span) {
  signalVector A((complex *)a, 0, la), B((complex *)b, 0, lb);
  A.isRealOnly(a_real); B.isRealOnly(b_real);
  signalVector *r = corr ? correlate(&A, &B, NULL, (ConvType)span) : convolve(&A, &B, NULL, (ConvType)span);
  if (!r) return -1;
  int n = r->size();
  if (cap >= n) put(*r, c);
  delete r;
  return n;
}
int ref_convolve(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span) {
  return conv_or_corr(false, a, la, a_real, b, lb, b_real, c, cap, span);
}
int ref_correlate(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span) {
  return conv_or_corr(true, a, la, a_real, b, lb, b_real, c, cap, span);
}

void ref_peak_detect(const float *v, int n, float *peak, float *idx, float *avg) {
  signalVector sv((complex *)v, 0, n);
  complex p = peakDetect(sv, idx, avg);
  peak[0] = p.real(); peak[1] = p.imag();
}

void ref_interpolate_point(const float *v, int n, float ix, float *out) {
  signalVector sv((complex *)v, 0, n);
  complex p = interpolatePoint(sv, ix);
  out[0] = p.real(); out[1] = p.imag();
}

int ref_energy_detect(const float *v, int n, unsigned win, float thr, float *avg) {
  signalVector sv((complex *)v, 0, n);
  return energyDetect(sv, win, thr, avg) ? 1 : 0;
}

int ref_analyze(const float *burst, int n, int tsc, float thr, int sps, float *amp, float *toa,
                int request, float *chan, float *off) {
  signalVector sv((complex *)burst, 0, n);
  complex a = 0.0; float t = 0.0F, o = 0.0F;
  signalVector *ch = NULL;
  bool ok = analyzeTrafficBurst(sv, tsc, thr, sps, &a, &t, request != 0, &ch, &o);
  amp[0] = a.real(); amp[1] = a.imag(); *toa = t;
  if (ch) { if (chan) put(*ch, chan); if (off) *off = o; delete ch; }
  return ok ? 1 : 0;
}

int ref_detect_rach(const float *burst, int n, float thr, int sps, float *amp, float *toa) {
  signalVector sv((complex *)burst, 0, n);
  complex a = 0.0; float t = 0.0F;
  bool ok = detectRACHBurst(sv, thr, sps, &a, &t);
  amp[0] = a.real(); amp[1] = a.imag(); *toa = t;
  return ok ? 1 : 0;
}

int ref_design_dfe(const float *chan, int nchan, float snr, int Nf, float *w, float *b) {
  signalVector ch((complex *)chan, 0, nchan);
  signalVector *W = NULL, *B = NULL;
  bool ok = designDFE(ch, snr, Nf, &W, &B);
  put(*W, w); put(*B, b);
  delete W; delete B;
  return ok ? 1 : 0;
}

/* burst is modified in place exactly as the reference does (delayVector, sigProcLib.cpp:1350) */
int ref_equalize(float *burst, int n, float toa, int sps, const float *w, int nw, const float *b, int nb, float *soft) {
  signalVector sv((complex *)burst, 0, n);
  signalVector W((complex *)w, 0, nw), B((complex *)b, 0, nb);
  SoftVector *s = equalizeBurst(sv, toa, sps, W, B);
  int m = s->size();
  memcpy(soft, s->begin(), m * sizeof(float));
  delete s;
  return m;
}

int ref_demodulate(const float *burst, int n, int sps, const float *amp, float toa, float *soft) {
  signalVector sv((complex *)burst, 0, n);
  SoftVector *s = demodulateBurst(sv, *gPulse, sps, complex(amp[0], amp[1]), toa);
  int m = s->size();
  memcpy(soft, s->begin(), m * sizeof(float));
  delete s;
  return m;
}

/* lpf: 0 = the RX filter (961 taps, gain 65), 1 = the TX filter (651 taps, gain 96) */
int ref_resample(const float *x, int n, int P, int Q, int lpf, float *out, int cap) {
  signalVector sv((complex *)x, 0, n);
  signalVector *r = polyphaseResampleVector(sv, P, Q, lpf ? gLpfTx : gLpfRx);
  int m = r->size();
  if (cap >= m) put(*r, out);
  delete r;
  return m;
}

/* ---- caller glue, stateless per burst --------------------------------------------------------- */

/* Transceiver.cpp:327-396 with estimateChannel==true for every burst and a caller-supplied energy
 * threshold (the reference adapts mEnergyThreshold burst to burst; that policy is SURVEY next-1).
 * Per burst outputs: flag, amp[2], toa, chan[12] (after the 1/amp scaling, i.e. what designDFE saw),
 * off, w[14], b[10], soft[soft_pitch] (first len entries written, zeros when not detected). */
void ref_rx_normal_batch(const float *bursts, int pitch, const int *lens, const unsigned char *tsc, long n,
                         float detect_thr, float energy_thr,
                         int *flags, float *amp, float *toa, float *chan, float *off, float *w, float *b,
                         float *soft, int soft_pitch) {
  for (long i = 0; i < n; i++) {
    int len = lens[i];
    signalVector burst(len);
    memcpy(burst.begin(), bursts + 2 * (size_t)pitch * i, len * sizeof(complex));
    complex a = 0.0; float t = 0.0F, o = 0.0F;
    signalVector *ch = NULL;
    bool ok = analyzeTrafficBurst(burst, tsc[i], detect_thr, 1, &a, &t, true, &ch, &o);
    flags[i] = ok;
    amp[2 * i] = a.real(); amp[2 * i + 1] = a.imag(); toa[i] = t;
    float *sp = soft + (size_t)soft_pitch * i;
    memset(sp, 0, soft_pitch * sizeof(float));
    if (off) off[i] = 0.0F;
    if (chan) memset(chan + 12 * i, 0, 12 * sizeof(float));
    if (w) memset(w + 14 * i, 0, 14 * sizeof(float));
    if (b) memset(b + 10 * i, 0, 10 * sizeof(float));
    if (!ok) { if (ch) delete ch; continue; }
    float SNR = a.norm2() / (energy_thr * energy_thr + 1.0);          // Transceiver.cpp:340
    scaleVector(*ch, complex(1.0, 0.0) / a);                           // :346
    signalVector *W = NULL, *B = NULL;
    designDFE(*ch, SNR, 7, &W, &B);                                    // :347
    scaleVector(burst, complex(1.0, 0.0) / a);                         // :391
    SoftVector *s = equalizeBurst(burst, t - o, 1, *W, *B);            // :392-396
    memcpy(sp, s->begin(), len * sizeof(float));
    if (off) off[i] = o;
    if (chan) put(*ch, chan + 12 * i);
    if (w) put(*W, w + 14 * i);
    if (b) put(*B, b + 10 * i);
    delete s; delete W; delete B; delete ch;
  }
}

/* Transceiver.cpp:360-389: detectRACHBurst then demodulateBurst */
void ref_rx_rach_batch(const float *bursts, int pitch, const int *lens, long n, float detect_thr, int sps,
                       int *flags, float *amp, float *toa, float *soft, int soft_pitch) {
  for (long i = 0; i < n; i++) {
    int len = lens[i];
    signalVector burst(len);
    memcpy(burst.begin(), bursts + 2 * (size_t)pitch * i, len * sizeof(complex));
    complex a = 0.0; float t = 0.0F;
    bool ok = detectRACHBurst(burst, detect_thr, sps, &a, &t);
    flags[i] = ok; amp[2 * i] = a.real(); amp[2 * i + 1] = a.imag(); toa[i] = t;
    float *sp = soft + (size_t)soft_pitch * i;
    memset(sp, 0, soft_pitch * sizeof(float));
    if (!ok) continue;
    SoftVector *s = demodulateBurst(burst, *gPulse, sps, a, t);
    memcpy(sp, s->begin(), s->size() * sizeof(float));
    delete s;
  }
}

/* radioInterface.cpp:238-259: chunk c of 864 raw samples, prefixed by the previous 192 raw samples
 * (zeros before the first chunk when first_chunk==0), resampled 65/96 through the 961-tap filter;
 * outputs 130..714 kept (585 per chunk).  raw points at chunk `first_chunk`; when first_chunk>0 the
 * 192 samples before raw[0] must be readable. */
void ref_rx_resample_stream(const float *raw, long first_chunk, long nchunks, float *out) {
  signalVector input(192 + 864);
  for (long c = 0; c < nchunks; c++) {
    const float *src = raw + 2 * 864 * c;
    if (first_chunk + c == 0) {
      memset(input.begin(), 0, 192 * sizeof(complex));
      memcpy(input.begin() + 192, src, 864 * sizeof(complex));
    } else {
      memcpy(input.begin(), src - 2 * 192, (192 + 864) * sizeof(complex));
    }
    signalVector *r = polyphaseResampleVector(input, 65, 96, gLpfRx);
    memcpy(out + 2 * 585 * c, r->begin() + 130, 585 * sizeof(complex));
    delete r;
  }
}

/* The same from the radio's own samples: RadioInterface::pullBuffer reads interleaved shorts from the device and
 * widens them with unUSRPifyVector (radioInterface.cpp:91-116: Complex<float>(s[FLIP_IQ], s[1-FLIP_IQ]); FLIP_IQ = 1 on
 * real hardware, 0 on the SWLOOPBACK build; usrp_to_host_short is the identity on a little-endian host) before the
 * resample (:238-259).  iq: 2 shorts per sample; iq[-384..-1] are the previous chunk's last 192 samples. */
void ref_rx_resample_stream_i16(const short *iq, int flip_iq, long first_chunk, long nchunks, float *out) {
  signalVector input(192 + 864);
  for (long c = 0; c < nchunks; c++) {
    const short *src = iq + 2 * 864 * c;
    const int hist = (first_chunk + c == 0) ? 0 : 192;
    if (!hist) memset(input.begin(), 0, 192 * sizeof(complex));
    signalVector::iterator itr = input.begin() + (192 - hist);
    const short *sp = src - 2 * hist;
    while (itr < input.end()) {
      *itr++ = Complex<float>(*(sp + flip_iq), *(sp + 1 - flip_iq));
      sp += 2;
    }
    signalVector *r = polyphaseResampleVector(input, 65, 96, gLpfRx);
    memcpy(out + 2 * 585 * c, r->begin() + 130, 585 * sizeof(complex));
    delete r;
  }
}

/* the RX datagram's soft bytes, Transceiver.cpp:667-669: burstString[8+i] = (char) round(soft[i]*255.0), gSlotLen = 148 */
void ref_soft_to_wire(const float *soft, int soft_pitch, long n, unsigned char *out) {
  for (long i = 0; i < n; i++)
    for (int k = 0; k < 148; k++) out[148 * i + k] = (unsigned char)(char)round(soft[(size_t)soft_pitch * i + k] * 255.0);
}

/* radioInterface.cpp:123-168 + USRPifyVector :74-89 (SWLOOPBACK byte order: host_to_usrp_short is
 * the identity on the loopback build; element 0 = real, 1 = imag).  Chunk c of 585 samples prefixed
 * by the previous 130, resampled 96/65 through the 651-tap filter, x13500, (short) truncation,
 * outputs 192..1055 kept (864 per chunk). */
void ref_tx_resample_stream(const float *in, long first_chunk, long nchunks, short *out) {
  signalVector input(130 + 585);
  for (long c = 0; c < nchunks; c++) {
    const float *src = in + 2 * 585 * c;
    if (first_chunk + c == 0) {
      memset(input.begin(), 0, 130 * sizeof(complex));
      memcpy(input.begin() + 130, src, 585 * sizeof(complex));
    } else {
      memcpy(input.begin(), src - 2 * 130, (130 + 585) * sizeof(complex));
    }
    signalVector *r = polyphaseResampleVector(input, 96, 65, gLpfTx);
    scaleVector(*r, 13500.0);
    short *o = out + 2 * 864 * c;
    for (int i = 0; i < 864; i++) {
      o[2 * i] = (short)(*r)[192 + i].real();
      o[2 * i + 1] = (short)(*r)[192 + i].imag();
    }
    delete r;
  }
}

/* bits -> modulateBurst(guard 8 + (tn%4==0)) for slots tn0.. ; bursts are written back to back
 * (157/156/156/156), the layout driveTransmitRadio feeds pushBuffer (radioInterface.cpp:337-357). */
long ref_modulate_stream(const char *bits148, long nbursts, int tn0, float *out) {
  long pos = 0;
  BitVector bv(148);
  for (long i = 0; i < nbursts; i++) {
    memcpy(bv.begin(), bits148 + 148 * i, 148);
    signalVector *m = modulateBurst(bv, *gPulse, 8 + (((tn0 + i) % 8) % 4 == 0), 1);
    put(*m, out + 2 * pos);
    pos += m->size();
    delete m;
  }
  return pos;
}

/* The north-star path end to end on the CPU, stateless: raw 400 kS/s stream -> RX resample ->
 * 157/156/156/156 slot cutting -> rx_normal per burst.  nchunks raw chunks starting at chunk 0
 * give floor(585*nchunks / 625) whole frames... the caller passes nbursts <= what is available. */
void ref_rx_stream_demod(const float *resampled, long first_burst, long nbursts, const unsigned char *tsc,
                         float detect_thr, float energy_thr, int *flags, float *amp, float *toa,
                         float *soft, int soft_pitch) {
  static const int slot_off[4] = {0, 157, 313, 469};
  for (long i = 0; i < nbursts; i++) {
    long g = first_burst + i;
    long start = (g / 4) * 625 + slot_off[g % 4];
    int len = (g % 4 == 0) ? 157 : 156;
    ref_rx_normal_batch(resampled + 2 * start, 0, &len, tsc + i, 1, detect_thr, energy_thr,
                        flags + i, amp + 2 * i, toa + i, NULL, NULL, NULL, NULL,
                        soft + (size_t)soft_pitch * i, soft_pitch);
  }
}


/* ------------------------------------------------------------------------------------------------
 * Transceiver::pullRadioVector + driveReceiveFIFO (Transceiver.cpp:207-269, 271-410, 641-676), the caller policy
 * around the path: slot map, adaptive energy threshold, 50-frame channel/DFE cache, RSSI / timing integerisation
 * and the RX datagram.  Transceiver.cpp itself cannot be compiled here (needs libusrp headers), so its glue is
 * restated below line by line; every DSP call is the real reference function from sigProcLib.cpp.
 *
 * One `trx` = one Transceiver object (one ARFCN).  State layout (doubles/ints, mirrored by the port and the CUDA
 * side):  thr (mEnergyThreshold, double), prev_false_fn, per TN: est_fn, have (channelResponse != NULL),
 * snr (float), chan_off (float), chan_amp[2] (unused downstream), w[14], b[10].
 * Bursts arrive in FIFO order: frame f = 0..nframes-1 (FN = fn0 + f), TN = 0..7, at
 * bursts + 2*pitch*((f*8)+tn) floats, length 157 if tn%4==0 else 156 (sps 1).
 * Output per burst: valid[i], and a 158-byte datagram at dgram + dgram_pitch*i (byte 156 is never written by the
 * reference; it is set to 0 here).
 * ------------------------------------------------------------------------------------------------ */
struct ref_trx_state {
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
enum { CT_NONE = 0, CT_I, CT_II, CT_III, CT_IV, CT_V, CT_VI, CT_VII, CT_LOOPBACK };   /* Transceiver.h ChannelCombination */
enum { CORR_OFF = 0, CORR_TSC, CORR_RACH, CORR_IDLE };
static const int kHyper = 2048 * 26 * 51;
static int fn_delta(int v1, int v2) {                     /* GSM::FNDelta, GSMCommon.cpp:161-168 */
  const int half = kHyper / 2;
  int d = v1 - v2;
  if (d >= half) d -= kHyper; else if (d < -half) d += kHyper;
  return d;
}
int ref_expected_corr_type(int chan_type, int fn) {       /* Transceiver.cpp:207-269 */
  switch (chan_type) {
    case CT_NONE: return CORR_OFF;
    case CT_I: return CORR_TSC;
    case CT_II: return (fn % 2 == 1) ? CORR_IDLE : CORR_TSC;
    case CT_III: return CORR_TSC;
    case CT_IV: case CT_VI: return ((fn % 51) % 10 < 2) ? CORR_RACH : CORR_OFF;
    case CT_V: {
      int m = fn % 51;
      if (m <= 36 && m >= 14) return CORR_RACH;
      if (m == 4 || m == 5) return CORR_RACH;
      if (m == 45 || m == 46) return CORR_RACH;
      return CORR_TSC;
    }
    case CT_VII: {
      int m = fn % 51;
      if (m == 12 || m == 13 || m == 14) return CORR_IDLE;
      return CORR_TSC;
    }
    case CT_LOOPBACK: {
      int m = fn % 51;
      return (m <= 50 && m >= 48) ? CORR_IDLE : CORR_TSC;
    }
    default: return CORR_OFF;
  }
}
int ref_trx_state_bytes(void) { return (int)sizeof(ref_trx_state); }
void ref_trx_init(ref_trx_state *st, int tsc, const int *chan_type, int start_fn) {   /* Transceiver.cpp:40-90 */
  memset(st, 0, sizeof *st);
  st->thr = 250.0;
  st->prev_false_fn = start_fn;
  st->tsc = tsc;
  for (int i = 0; i < 8; i++) { st->chan_type[i] = chan_type[i]; st->est_fn[i] = start_fn; st->have[i] = 0; }
}
void ref_trx_pull(ref_trx_state *st, const float *bursts, int pitch, int nframes, int fn0,
                  int *valid, unsigned char *dgram, int dgram_pitch) {
  for (int f = 0; f < nframes; f++) {
    const int fn = (fn0 + f) % kHyper;
    for (int tn = 0; tn < 8; tn++) {
      const long i = (long)f * 8 + tn;
      valid[i] = 0;
      unsigned char *dg = dgram + (size_t)dgram_pitch * i;
      memset(dg, 0, 158);
      const int corr = ref_expected_corr_type(st->chan_type[tn], fn);
      if (corr == CORR_OFF || corr == CORR_IDLE) continue;                               /* :290-293 */
      const int len = (tn % 4 == 0) ? 157 : 156;
      signalVector burst(len);
      memcpy(burst.begin(), bursts + 2 * (size_t)pitch * i, len * sizeof(complex));
      complex amplitude = 0.0;
      float TOA = 0.0F, avgPwr = 0.0F;
      if (!energyDetect(burst, 20, st->thr, &avgPwr)) {                                  /* :298 (double -> float thr) */
        double framesElapsed = fn_delta(fn, st->prev_false_fn);
        if (framesElapsed > 50) { st->thr -= 10.0; st->prev_false_fn = fn; }           /* :300-304 */
        continue;
      }
      bool success = false;
      if (corr == CORR_TSC) {
        double framesElapsed = fn_delta(fn, st->est_fn[tn]);
        bool estimateChannel = false;
        if (framesElapsed > 50 || !st->have[tn]) { st->have[tn] = 0; estimateChannel = true; }   /* :317-326 */
        signalVector *channelResp = NULL;
        float chanOffset = 0.0F;
        success = analyzeTrafficBurst(burst, st->tsc, 3.0, 1, &amplitude, &TOA, estimateChannel, &channelResp, &chanOffset);
        if (success) {
          st->thr -= 1.0F;
          if (st->thr < 0.0) st->thr = 0.0;
          st->snr[tn] = amplitude.norm2() / (st->thr * st->thr + 1.0);                   /* :340 */
          if (estimateChannel) {
            st->have[tn] = 1;
            st->chan_off[tn] = chanOffset;
            scaleVector(*channelResp, complex(1.0, 0.0) / amplitude);
            signalVector *W = NULL, *B = NULL;
            designDFE(*channelResp, st->snr[tn], 7, &W, &B);
            put(*W, st->w[tn]); put(*B, st->b[tn]);
            delete W; delete B;
            st->est_fn[tn] = fn;
          }
        } else {
          double fe = fn_delta(fn, st->prev_false_fn);
          st->thr += 10.0F * exp(-fe);                                                   /* :355 */
          st->prev_false_fn = fn;
          st->have[tn] = 0;                                                              /* :357 channelResponse = NULL */
        }
        if (channelResp) delete channelResp;
      } else {
        success = detectRACHBurst(burst, 5.0, 1, &amplitude, &TOA);
        if (success) {
          st->thr -= 1.0F;
          if (st->thr < 0.0) st->thr = 0.0;
          st->have[tn] = 0;                                                              /* :371 */
        } else {
          double fe = fn_delta(fn, st->prev_false_fn);
          st->thr += 10.0F * exp(-fe);
          st->prev_false_fn = fn;
        }
      }
      if (!success) continue;
      SoftVector *soft;
      if (corr == CORR_RACH) {
        soft = demodulateBurst(burst, *gPulse, 1, amplitude, TOA);
      } else {
        scaleVector(burst, complex(1.0, 0.0) / amplitude);
        signalVector W(7), B(5);
        memcpy(W.begin(), st->w[tn], 7 * sizeof(complex));
        memcpy(B.begin(), st->b[tn], 5 * sizeof(complex));
        soft = equalizeBurst(burst, TOA - st->chan_off[tn], 1, W, B);
      }
      const int RSSI = (int)floor(20.0 * log10(9450.0 / amplitude.abs()));               /* :400 */
      const int timingOffset = (int)round(TOA * 256.0 / 1);                              /* :402 */
      valid[i] = 1;
      dg[0] = tn;                                                                        /* :659-673 */
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


/* Transceiver::driveTransmitPriorityQueue (:582-632) + addRadioVector (:100-114) + RadioInterface::pushBuffer
 * (radioInterface.cpp:123-168), batched over a window of nframes frames starting at fn0 (nframes % 117 == 0):
 * datagram = TN, FN[4] big endian, RSSI, 148 bits.  Slots without a datagram carry `filler` (unscaled) or silence.
 * The placement glue is restated; modulateBurst / scaleVector / polyphaseResampleVector are the reference's. */
long ref_tx_datagrams(const unsigned char *dgram, long n, int dgram_pitch, int fn0, int nframes,
                      const unsigned char *filler, short *out) {
  const long nslots = (long)nframes * 8, nsamp = nslots / 4 * 625;
  static const int slot_off[4] = {0, 157, 313, 469};
  signalVector stream(nsamp);                    /* zero-initialised */
  long *src = new long[nslots];
  for (long s = 0; s < nslots; s++) src[s] = -1;
  long placed = 0;
  for (long i = 0; i < n; i++) {
    const char *buffer = (const char *)(dgram + (size_t)dgram_pitch * i);
    int timeSlot = (int)buffer[0];
    unsigned long frameNum = 0;
    for (int k = 0; k < 4; k++) frameNum = (frameNum << 8) | (0x0ff & buffer[k + 1]);
    if (timeSlot < 0 || timeSlot > 7) continue;
    int f = fn_delta((int)(frameNum % kHyper), fn0 % kHyper);
    if (f < 0 || f >= nframes) continue;
    src[(long)f * 8 + timeSlot] = i;
    placed++;
  }
  BitVector bv(148);
  for (long sl = 0; sl < nslots; sl++) {
    const int tn = (int)(sl % 8);
    signalVector *m = NULL;
    if (src[sl] >= 0) {
      const char *buffer = (const char *)(dgram + (size_t)dgram_pitch * src[sl]);
      int RSSI = (int)buffer[5];
      memcpy(bv.begin(), buffer + 6, 148);
      m = modulateBurst(bv, *gPulse, 8 + (tn % 4 == 0), 1);
      scaleVector(*m, pow(10, -RSSI / 10));                    /* Transceiver.cpp:108 */
    } else if (filler) {
      memcpy(bv.begin(), filler, 148);
      m = modulateBurst(bv, *gPulse, 8 + (tn % 4 == 0), 1);
    }
    if (m) {
      memcpy(stream.begin() + (sl / 4) * 625 + slot_off[sl % 4], m->begin(), m->size() * sizeof(complex));
      delete m;
    }
  }
  delete[] src;
  ref_tx_resample_stream((const float *)stream.begin(), 0, nsamp / 585, out);
  return placed;
}


/* ------------------------------------------------------------------------------------------------
 * L1 FEC after the path: XCCHL1Decoder::deinterleave + decode (GSML1FEC.cpp:616-660) on frames of
 * four received bursts, given as the RX datagrams' soft bytes (ARFCNManager::driveRx: byte / 256.0F,
 * TRXManager.cpp:230); SoftVector::decode / ViterbiR2O4 / Parity are the reference's own classes
 * (CommonLibs/BitVector.cpp:290-540).  GSML1FEC.cpp itself needs the whole GSM stack, so the ten lines of glue are
 * restated.  u: 228 decoded bits per frame (d[184] : p[40] : tail[4], before the decoder's in-place parity
 * inversion); ok: syndrome == 0.
 * ------------------------------------------------------------------------------------------------ */
void ref_xcch_decode(const unsigned char *soft, int burst_pitch, long nframes, unsigned char *u, int *ok) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  for (long f = 0; f < nframes; f++) {
    SoftVector mI[4];
    SoftVector mC(456);
    BitVector mU(228);
    for (int B = 0; B < 4; B++) {
      const unsigned char *rp = soft + (size_t)burst_pitch * (4 * f + B);
      float data[148];
      for (int i = 0; i < 148; i++) data[i] = rp[i] / 256.0F;                 /* TRXManager.cpp:230 */
      mI[B] = SoftVector(114);
      for (int i = 0; i < 57; i++) { mI[B][i] = data[3 + i]; mI[B][57 + i] = data[88 + i]; }   /* :603-604 */
    }
    for (int k = 0; k < 456; k++) {                                           /* :616-630 */
      int B = k % 4;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      mC[k] = mI[B][j];
    }
    mC.decode(vcoder, mU);                                                    /* :641 */
    for (int i = 0; i < 228; i++) u[228 * f + i] = mU.bit(i);
    BitVector mP(mU.segment(184, 40)), mDP(mU.head(224));
    mP.invert();                                                              /* :649 */
    unsigned syndrome = blockCoder.syndrome(mDP);                             /* :652: an `unsigned` -- the low 32 of the 40 bits */
    ok[f] = syndrome == 0;                                                    /* :654 */
  }
}
/* XCCHL1Encoder::encode + interleave (GSML1FEC.cpp:795-819): d[184] -> the e-bits of four bursts (114 each).
 * Test-input generator for the decoder. */
void ref_xcch_encode(const unsigned char *d, long nframes, unsigned char *e) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  for (long f = 0; f < nframes; f++) {
    BitVector mU(228), mC(456);
    mU.fill(0);
    BitVector mD(mU.head(184)), mP(mU.segment(184, 40));
    for (int i = 0; i < 184; i++) mD[i] = d[184 * f + i] & 1;
    blockCoder.writeParityWord(mD, mP);
    mU.encode(vcoder, mC);
    for (int k = 0; k < 456; k++) {
      int B = k % 4;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      e[(4 * f + B) * 114 + j] = mC.bit(k);
    }
  }
}


/* the same with the parity word XORed by pflip[f] before it is written: frames whose Fire-code syndrome is a chosen
 * non-zero value (the decoder keeps the syndrome in an `unsigned`, GSML1FEC.cpp:652: bits 32..39 go unseen) */
void ref_xcch_encode_pflip(const unsigned char *d, const unsigned long long *pflip, long nframes, unsigned char *e) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  for (long f = 0; f < nframes; f++) {
    BitVector mU(228), mC(456);
    mU.fill(0);
    BitVector mD(mU.head(184)), mP(mU.segment(184, 40));
    for (int i = 0; i < 184; i++) mD[i] = d[184 * f + i] & 1;
    uint64_t pWord = ~mD.parity(blockCoder);                        /* Parity::writeParityWord, BitVector.cpp:411-416 */
    mP.fillField(0, pWord ^ pflip[f], 40);
    mU.encode(vcoder, mC);
    for (int k = 0; k < 456; k++) {
      int B = k % 4;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      e[(4 * f + B) * 114 + j] = mC.bit(k);
    }
  }
}

/* RACHL1Decoder::writeLowSide (GSML1FEC.cpp:474-515) up to the BSIC comparison: per burst u[18], tail = peekField(14,4),
 * bsic = (~sentParity ^ checkParity) & 0x3f, ra = RA after LSB8MSB.  Reference classes, restated glue. */
void ref_rach_decode(const unsigned char *soft, int burst_pitch, long n, unsigned char *u, int *tail, int *bsic, int *ra) {
  ViterbiR2O4 vcoder;
  Parity parity(0x06f, 6, 8);
  for (long i = 0; i < n; i++) {
    const unsigned char *rp = soft + (size_t)burst_pitch * i;
    SoftVector e(36);
    for (int k = 0; k < 36; k++) e[k] = rp[49 + k] / 256.0F;
    BitVector mU(18);
    BitVector mD(mU.head(8));
    e.decode(vcoder, mU);
    for (int k = 0; k < 18; k++) u[18 * i + k] = mU.bit(k);
    tail[i] = (int)mU.peekField(14, 4);
    unsigned sentParity = ~mU.peekField(8, 6);
    unsigned checkParity = mD.parity(parity);
    bsic[i] = (int)((sentParity ^ checkParity) & 0x03f);
    mD.LSB8MSB();
    ra[i] = (int)mD.peekField(0, 8);
  }
}
/* RACHL1Encoder (GSML1FEC.cpp, GSM 05.03 4.6): RA -> the 36 coded bits.  Test-input generator. */
void ref_rach_encode(const unsigned char *ra, const unsigned char *bsic, long n, unsigned char *e) {
  ViterbiR2O4 vcoder;
  Parity parity(0x06f, 6, 8);
  for (long i = 0; i < n; i++) {
    BitVector mU(18), mC(36);
    mU.fill(0);
    BitVector mD(mU.head(8)), mP(mU.segment(8, 6));
    mD.fillField(0, ra[i], 8);
    mD.LSB8MSB();
    parity.writeParityWord(mD, mP);                    /* inverted parity */
    unsigned p = mP.peekField(0, 6) ^ (bsic[i] & 0x3f);
    mP.fillField(0, p, 6);
    mU.encode(vcoder, mC);
    for (int k = 0; k < 36; k++) e[36 * i + k] = mC.bit(k);
  }
}

/* ------------------------------------------------------------------------------------------------
 * TCH/FACCH (GSML1FEC.cpp:1031-1210): TCHFACCHL1Decoder::processBurst + deinterleave + decodeTCH, and XCCHL1Decoder::decode
 * for a stolen (FACCH) block, over one traffic channel's consecutive traffic bursts given as RX-datagram soft bytes.
 * A decoder object is fed 4*nblocks + 4 bursts with B = 0, 1, 2, ... (mod 8); the first completion (B == 3, which would mix
 * in the object's initial zeros) is skipped, so block q is the one completed by burst 4q + 7 -- bursts 4q .. 4q+7.
 * The reference's classes do the work; the glue around them is restated line by line.
 * ------------------------------------------------------------------------------------------------ */
void ref_tch_decode(const unsigned char *soft, int burst_pitch, long nblocks, unsigned char *d, int *good, int *stolen_o,
                    unsigned char *fu, int *fok) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  Parity tchParity(0x0b, 3, 50);
  SoftVector mI[8];
  for (int i = 0; i < 8; i++) { mI[i] = SoftVector(114); mI[i].fill(.0); }       /* :1008-1012 */
  SoftVector mC(456);
  for (long n = 0; n < 4 * nblocks + 4; n++) {
    const unsigned char *rp = soft + (size_t)burst_pitch * n;
    SoftVector burst(148);
    for (int i = 0; i < 148; i++) burst[i] = rp[i] / 256.0F;                     /* TRXManager.cpp:230 */
    const int B = (int)(n % 8);
    burst.segment(3, 57).copyToSegment(mI[B], 0);                                /* data1(), :1062 */
    burst.segment(88, 57).copyToSegment(mI[B], 57);                              /* data2(), :1063 */
    if (B % 4 != 3) continue;                                                    /* :1067 */
    const int blockOffset = (B == 3) ? 4 : 0;                                    /* :1071-1072 */
    for (int k = 0; k < 456; k++) {                                              /* deinterleave, :1102-1110 */
      int Bk = (k + blockOffset) % 8;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      mC[k] = mI[Bk][j];
      mI[Bk][j] = 0.5F;
    }
    const long q = n / 4 - 1;
    if (q < 0) continue;                                                         /* the start-up block that reads the initial zeros */
    const bool stolen = burst.bit(60);                                           /* inBurst.Hl(), :1075 */
    stolen_o[q] = stolen;
    memset(d + 260 * q, 0, 260);
    memset(fu + 228 * q, 0, 228);
    good[q] = 0; fok[q] = 0;
    if (stolen) {                                                                /* XCCHL1Decoder::decode, :636-660 */
      BitVector mU(228);
      mC.decode(vcoder, mU);
      for (int i = 0; i < 228; i++) fu[228 * q + i] = mU.bit(i);
      BitVector mP(mU.segment(184, 40)), mDP(mU.head(224));
      mP.invert();
      unsigned syndrome = blockCoder.syndrome(mDP);
      fok[q] = syndrome == 0;
    } else {                                                                     /* decodeTCH(false), :1129-1160 */
      BitVector mTCHU(189), mTCHD(260);
      SoftVector mClass1_c(mC.head(378)), mClass2_c(mC.segment(378, 78));
      BitVector mClass1A_d(mTCHD.head(50));
      mClass1_c.decode(vcoder, mTCHU);
      mClass2_c.sliced().copyToSegment(mTCHD, 182);
      for (unsigned k = 0; k <= 90; k++) {
        mTCHD[2 * k] = mTCHU[k];
        mTCHD[2 * k + 1] = mTCHU[184 - k];
      }
      unsigned sentParity = (~mTCHU.peekField(91, 3)) & 0x07;
      unsigned calcParity = mClass1A_d.parity(tchParity) & 0x07;
      unsigned tail = mTCHU.peekField(185, 4);
      good[q] = (sentParity == calcParity) && (tail == 0);
      for (int i = 0; i < 260; i++) d[260 * q + i] = mTCHD.bit(i);
    }
  }
}
/* TCHFACCHL1Encoder::encodeTCH (:1248-1279) / XCCHL1Encoder::encode for a FACCH frame (:795-808) + the diagonal interleaver
 * (:1384-1392) + the stealing flags: nblocks blocks -> the 148 burst bits of 4*nblocks + 4 bursts (tails and midamble left
 * zero).  Test-input generator for the decoder: block q (speech d260[q], or, when steal[q], the FACCH payload f184[q]) goes to
 * bursts 4q .. 4q+7; Hl (bit 60) of bursts 4q+4 .. 4q+7 and Hu (bit 87) of bursts 4q .. 4q+3 carry steal[q]. */
void ref_tch_encode(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long nblocks,
                    unsigned char *bursts) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  Parity tchParity(0x0b, 3, 50);
  memset(bursts, 0, (size_t)(4 * nblocks + 4) * 148);
  for (long q = 0; q < nblocks; q++) {
    BitVector mC(456);
    if (steal[q]) {
      BitVector mU(228);
      mU.fill(0);
      BitVector mD(mU.head(184)), mP(mU.segment(184, 40));
      for (int i = 0; i < 184; i++) mD[i] = f184[184 * q + i] & 1;
      blockCoder.writeParityWord(mD, mP);
      mU.encode(vcoder, mC);
    } else {
      BitVector mTCHU(189), mTCHD(260);
      mTCHU.fill(0);
      for (int i = 0; i < 260; i++) mTCHD[i] = d260[260 * q + i] & 1;
      BitVector mClass1A_d(mTCHD.head(50)), mClass2_d(mTCHD.segment(182, 78));
      BitVector p = mTCHU.segment(91, 3);
      tchParity.writeParityWord(mClass1A_d, p);
      for (unsigned k = 0; k <= 90; k++) {
        mTCHU[k] = mTCHD[2 * k];
        mTCHU[184 - k] = mTCHD[2 * k + 1];
      }
      for (unsigned k = 185; k <= 188; k++) mTCHU[k] = 0;
      BitVector mClass1_c(mC.head(378));
      mTCHU.encode(vcoder, mClass1_c);
      mClass2_d.copyToSegment(mC, 378);
    }
    for (int k = 0; k < 456; k++) {
      const int r = k % 8;
      const int j = 2 * ((49 * k) % 57) + (r / 4);
      bursts[(size_t)(4 * q + r) * 148 + (j < 57 ? 3 + j : 88 + (j - 57))] = mC.bit(k);
    }
    for (int b = 0; b < 4; b++) {
      bursts[(size_t)(4 * q + 4 + b) * 148 + 60] = steal[q] ? 1 : 0;
      bursts[(size_t)(4 * q + b) * 148 + 87] = steal[q] ? 1 : 0;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * L1 encoders on the transmit side, in the reference's own flow and with its own classes (BitVector::LSB8MSB, Parity,
 * BitVector::encode, gTrainingSequence); the glue of GSML1FEC.cpp (which needs the whole GSM stack to compile) is restated.
 * XCCHL1Encoder: constructor :716-745 (stealing bits, midamble), sendFrame :763-790, encode :795-808, interleave :811-819,
 * transmit :823-850.  bursts: 4 * nframes x 148 bits.  tsc < 0: no midamble written.
 * ------------------------------------------------------------------------------------------------ */
void ref_xcch_send_frames(const unsigned char *frames, long nframes, int lsb8msb, int tsc, unsigned char *bursts) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  BitVector mBurst(148);
  mBurst.fill(0);
  mBurst[60] = 1;                                                      /* mBurst.Hl(1), GSMTransfer.h:47,120 */
  mBurst[87] = 1;                                                      /* mBurst.Hu(1), :48,117 */
  if (tsc >= 0) GSM::gTrainingSequence[tsc].copyToSegment(mBurst, 61);
  BitVector mC(456), mU(228);
  BitVector mD(mU.head(184)), mP(mU.segment(184, 40));
  mU.zero();
  BitVector mI[4];
  for (int k = 0; k < 4; k++) { mI[k] = BitVector(114); mI[k].fill(0); }
  for (long f = 0; f < nframes; f++) {
    BitVector frame(184);
    for (int i = 0; i < 184; i++) frame[i] = frames[184 * f + i] & 1;
    frame.copyToSegment(mU, 0);                                        /* :779, headerOffset() == 0 */
    if (lsb8msb) mD.LSB8MSB();                                         /* :781 */
    blockCoder.writeParityWord(mD, mP);                                /* :801 */
    mU.encode(vcoder, mC);                                             /* :805 */
    for (int k = 0; k < 456; k++) {                                    /* :814-818 */
      int B = k % 4;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      mI[B][j] = mC[k];
    }
    for (int B = 0; B < 4; B++) {                                      /* :837-848 */
      mI[B].segment(0, 57).copyToSegment(mBurst, 3);
      mI[B].segment(57, 57).copyToSegment(mBurst, 88);
      for (int i = 0; i < 148; i++) bursts[(size_t)(4 * f + B) * 148 + i] = mBurst.bit(i);
    }
  }
}

/* TCHFACCHL1Encoder: encodeTCH :1248-1279 (from the class-ordered d[260]; the g610BitOrder map of :1255 is the caller's),
 * dispatch :1299-1381 (FACCH branch :1323-1333, stealing flags :1365-1366, mOffset toggling :1373-1374), interleave :1384-1392.
 * One dispatch() per block sends four bursts; `state` (4 x 114 interleaver rows still to be sent, then mPreviousFACCH, then
 * mOffset: 458 bytes) carries the encoder's members across calls -- all zeros for a channel that starts here.
 * bursts: 4 * nblocks x 148 bits. */
void ref_tch_dispatch(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long nblocks, int lsb8msb,
                      int tsc, unsigned char *state, unsigned char *bursts) {
  ViterbiR2O4 vcoder;
  Parity blockCoder(0x10004820009ULL, 40, 224);
  Parity tchParity(0x0b, 3, 50);
  BitVector mBurst(148);
  mBurst.fill(0);
  if (tsc >= 0) GSM::gTrainingSequence[tsc].copyToSegment(mBurst, 61);
  BitVector mC(456), mU(228), mTCHU(189), mTCHD(260);
  BitVector mD(mU.head(184)), mP(mU.segment(184, 40));
  BitVector mClass1_c(mC.head(378)), mClass1A_d(mTCHD.head(50)), mClass2_d(mTCHD.segment(182, 78));
  mU.zero();
  mTCHU.fill(0);
  int mOffset = state[457] ? 4 : 0;
  bool mPreviousFACCH = state[456] != 0;
  BitVector mI[8];
  for (int k = 0; k < 8; k++) { mI[k] = BitVector(114); mI[k].fill(0); }
  for (int B = 0; B < 4; B++)
    for (int j = 0; j < 114; j++) mI[B + mOffset][j] = state[114 * B + j] & 1;
  for (long q = 0; q < nblocks; q++) {
    bool currentFACCH = false;
    if (steal[q]) {
      currentFACCH = true;
      BitVector fFrame(184);
      for (int i = 0; i < 184; i++) fFrame[i] = f184[184 * q + i] & 1;
      if (lsb8msb) fFrame.LSB8MSB();                                   /* :1327 */
      fFrame.copyTo(mU);                                               /* :1328 */
      blockCoder.writeParityWord(mD, mP);                              /* encode(), :801-805 */
      mU.encode(vcoder, mC);
    } else {
      for (int i = 0; i < 260; i++) mTCHD[i] = d260[260 * q + i] & 1;
      BitVector p = mTCHU.segment(91, 3);
      tchParity.writeParityWord(mClass1A_d, p);
      for (unsigned k = 0; k <= 90; k++) {
        mTCHU[k] = mTCHD[2 * k];
        mTCHU[184 - k] = mTCHD[2 * k + 1];
      }
      for (unsigned k = 185; k <= 188; k++) mTCHU[k] = 0;
      mTCHU.encode(vcoder, mClass1_c);
      mClass2_d.copyToSegment(mC, 378);
    }
    for (int k = 0; k < 456; k++) {                                    /* interleave(mOffset) :1384-1392 */
      int B = (k + mOffset) % 8;
      int j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      mI[B][j] = mC[k];
    }
    for (int B = 0; B < 4; B++) {                                      /* :1358-1370 */
      mI[B + mOffset].segment(0, 57).copyToSegment(mBurst, 3);
      mI[B + mOffset].segment(57, 57).copyToSegment(mBurst, 88);
      mBurst[87] = currentFACCH;
      mBurst[60] = mPreviousFACCH;
      for (int i = 0; i < 148; i++) bursts[(size_t)(4 * q + B) * 148 + i] = mBurst.bit(i);
    }
    if (mOffset == 0) mOffset = 4; else mOffset = 0;                   /* :1373-1374 */
    mPreviousFACCH = currentFACCH;
  }
  for (int B = 0; B < 4; B++)
    for (int j = 0; j < 114; j++) state[114 * B + j] = mI[B + mOffset].bit(j);
  state[456] = mPreviousFACCH ? 1 : 0;
  state[457] = mOffset ? 1 : 0;
}

}  // extern "C"
