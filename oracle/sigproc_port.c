/*
 * oracle/sigproc_port.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C, scalar, op-for-op restatement of the reference's burst-DSP path
 * (reference Transceiver/sigProcLib.cpp + Complex.h + the resample/slot glue of radioInterface.cpp and the
 * TSC/RACH branches of Transceiver.cpp::pullRadioVector).  It is NOT a copy: it is written from the
 * algorithm, with one rule -- every IEEE binary32/binary64 operation happens in the same order and
 * precision as in the reference, so results are bit-identical (build with -ffp-contract=off).
 *
 * Pinning: the reference ships no golden vectors for this path, so this port is pinned by
 * (1) oracle/_ref/libref_oracle.so = the unmodified reference compiled in place, compared call by call in
 * tests/test_oracle_port.py when that library is present, and (2) tests/golden/ (npz files), outputs of that same
 * compiled reference committed with their generator (oracle/gen_golden.py), compared everywhere.
 *
 * Each function cites the reference lines it follows.
 */
#define _GNU_SOURCE
#include "sigproc_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "lpf_taps.inc"

typedef struct { float r, i; } cpx;

/* ---- Complex<float> arithmetic, reference Transceiver/Complex.h ---------------------------------- */
static inline cpx C(float r, float i) { cpx z = {r, i}; return z; }
static inline cpx cadd(cpx a, cpx b) { return C(a.r + b.r, a.i + b.i); }                    /* :79 */
static inline cpx cmul(cpx a, cpx b) { return C(a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r); } /* :83 */
static inline cpx cmulr(cpx a, float s) { return C(a.r * s, a.i * s); }                     /* :84 */
static inline cpx cdivr(cpx a, float s) { return C(a.r / s, a.i / s); }                     /* :86 */
static inline cpx cconj(cpx a) { return C(a.r, -a.i); }                                     /* :121 */
static inline float cnorm2(cpx a) { return a.i * a.i + a.r * a.r; }                         /* :122 */
static inline cpx cinv(cpx a) { float n = cnorm2(a); return C(a.r / n, -a.i / n); }         /* :154-160 */
static inline cpx cdiv(cpx a, cpx b) { return cmul(a, cinv(b)); }                           /* :85 */
static inline float cabsf_(cpx a) { return sqrtf(cnorm2(a)); }                              /* :131 */

/* ---- library state (reference sigProcLib.cpp:39-59) ---------------------------------------------- */
#define TABLESIZE 1024
static float cosTable[TABLESIZE + 1], sinTable[TABLESIZE + 1];
static const float M_PI_F = (float)M_PI;
static const float M_2PI_F = (float)(2.0 * M_PI);
static float M_1_2PI_F;
static int gSps = 0;
static cpx *gRot = NULL, *gRevRot = NULL;    /* 157*sps each */
static cpx gPulse[64]; static int gPulseLen = 0;
typedef struct { cpx *seq; int len; float TOA; cpx gain; } corrseq;
static corrseq gMid[8], gRach;
static float gLpfRx[961], gLpfTx[651];

static const char *kTSC[8] = {  /* GSM 05.02 5.2.3 training sequences == reference GSM/GSMCommon.cpp:44-53 */
  "00100101110000100010010111", "00101101110111100010110111", "01000011101110100100001110",
  "01000111101101000100011110", "00011010111001000001101011", "01001110101100000100111010",
  "10100111110110001010011111", "11101111000100101110111100"};
static const char *kRACH = "01001011011111111001100110101010001111000";  /* GSM 05.02 5.2.7, GSMCommon.cpp:57 */

/* ---- table trig, sigProcLib.cpp:163-204 ---------------------------------------------------------- */
float port_cos_lookup(float x) {
  float arg = x * M_1_2PI_F;
  while (arg > 1.0F) arg -= 1.0F;
  while (arg < 0.0F) arg += 1.0F;
  const float argT = arg * ((float)TABLESIZE);
  const int argI = (int)argT;
  const float delta = argT - argI;
  const float iDelta = 1.0F - delta;
  /* argI can be 1024 (arg==1): the reference then reads cosTable[1025], one past the array, times
   * delta==0.  That element is sinTable[0] in the reference's layout; 0*finite = 0 either way. */
  float next = (argI + 1 <= TABLESIZE) ? cosTable[argI + 1] : sinTable[0];
  return iDelta * cosTable[argI] + delta * next;
}
float port_sin_lookup(float x) {
  float arg = x * M_1_2PI_F;
  while (arg > 1.0F) arg -= 1.0F;
  while (arg < 0.0F) arg += 1.0F;
  const float argT = arg * ((float)TABLESIZE);
  const int argI = (int)argT;
  const float delta = argT - argI;
  const float iDelta = 1.0F - delta;
  float next = (argI + 1 <= TABLESIZE) ? sinTable[argI + 1] : 0.0F;   /* delta==0 there */
  return iDelta * sinTable[argI] + delta * next;
}
static cpx expj_lookup(float x) { return C(port_cos_lookup(x), port_sin_lookup(x)); }   /* :192-204 */

float port_sinc(float x) {                                                               /* :567-571 */
  if ((x >= 0.01F) || (x <= -0.01F)) return port_sin_lookup(x) / x;
  return 1.0F;
}

/* ---- convolve / correlate, sigProcLib.cpp:267-408, :474-503 (symmetry NONE only; ABSSYM is never set
 * on the path) -------------------------------------------------------------------------------------- */
static int conv_sizes(int la, int lb, int span, int *start, int *outsz) {
  switch (span) {
    case PORT_FULL_SPAN:    *start = 0;  *outsz = la + lb - 1; break;
    case PORT_OVERLAP_ONLY: *start = la; *outsz = abs(la - lb) + 1; break;
    case PORT_START_ONLY:   *start = 0;  *outsz = la; break;
    case PORT_WITH_TAIL:    *start = lb; *outsz = la; break;
    case PORT_NO_DELAY:     *start = (lb % 2) ? lb / 2 : lb / 2 - 1; *outsz = la; break;
    default: return -1;
  }
  return 0;
}
static int convolve_(const cpx *a, int la, int a_real, const cpx *b, int lb, int b_real, cpx *c, int span) {
  int start, outsz;
  if (conv_sizes(la, lb, span, &start, &outsz)) return -1;
  for (int n = 0; n < outsz; n++) {
    int t = start + n;
    if (a_real && b_real) {
      float sum = 0.0F;
      for (int k = 0, ai = t; k < lb; k++, ai--) {
        if (ai < 0) break;
        if (ai < la) sum += a[ai].r * b[k].r;
      }
      c[n] = C(sum, 0.0F);
    } else if (a_real) {
      cpx sum = C(0, 0);
      for (int k = 0, ai = t; k < lb; k++, ai--) {
        if (ai < 0) break;
        if (ai < la) sum = cadd(sum, cmulr(b[k], a[ai].r));
      }
      c[n] = sum;
    } else if (b_real) {
      cpx sum = C(0, 0);
      for (int k = 0, ai = t; k < lb; k++, ai--) {
        if (ai < 0) break;
        if (ai < la) sum = cadd(sum, cmulr(a[ai], b[k].r));
      }
      c[n] = sum;
    } else {
      cpx sum = C(0, 0);
      for (int k = 0, ai = t; k < lb; k++, ai--) {
        if (ai < 0) break;
        if (ai < la) sum = cadd(sum, cmul(a[ai], b[k]));
      }
      c[n] = sum;
    }
  }
  return outsz;
}
static int correlate_(const cpx *a, int la, int a_real, const cpx *b, int lb, int b_real, cpx *c, int span) {
  cpx *tmp = (cpx *)calloc(lb, sizeof(cpx));
  for (int k = 0; k < lb; k++) tmp[lb - 1 - k] = b_real ? C(b[k].r, 0.0F) : cconj(b[k]);
  int r = convolve_(a, la, a_real, tmp, lb, b_real, c, span);
  free(tmp);
  return r;
}
int port_convolve(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span) {
  int start, outsz;
  if (conv_sizes(la, lb, span, &start, &outsz)) return -1;
  if (cap < outsz) return outsz;
  return convolve_((const cpx *)a, la, a_real, (const cpx *)b, lb, b_real, (cpx *)c, span);
}
int port_correlate(const float *a, int la, int a_real, const float *b, int lb, int b_real, float *c, int cap, int span) {
  int start, outsz;
  if (conv_sizes(la, lb, span, &start, &outsz)) return -1;
  if (cap < outsz) return outsz;
  return correlate_((const cpx *)a, la, a_real, (const cpx *)b, lb, b_real, (cpx *)c, span);
}

/* ---- scaleVector, sigProcLib.cpp:713-730 ---------------------------------------------------------- */
static void scale_(cpx *x, int n, int real_only, cpx s) {
  for (int i = 0; i < n; i++) x[i] = real_only ? cmulr(s, x[i].r) : cmul(x[i], s);
}
void port_scale_vector(float *v, int n, int real_only, const float *scale) {
  scale_((cpx *)v, n, real_only, C(scale[0], scale[1]));
}

/* ---- modulateBurst, sigProcLib.cpp:521-565 with GMSKRotate :232-247 ----------------------------- */
static int modulate_(const char *bits, int nbits, const cpx *pulse, int plen, int pulse_real, int guard, int sps, cpx *out) {
  int n = sps * (nbits + guard);
  cpx *x = (cpx *)calloc(n, sizeof(cpx));
  for (int i = 0; i < nbits; i++) x[i * sps] = C((float)(2.0 * (bits[i] & 0x01) - 1.0), 0.0F);
  for (int i = 0; i < n; i++) x[i] = cmulr(gRot[i], x[i].r);          /* realOnly branch of GMSKRotate */
  convolve_(x, n, 0, pulse, plen, pulse_real, out, PORT_NO_DELAY);
  free(x);
  return n;
}
int port_modulate(const char *bits, int nbits, int guard, int sps, float *out, int cap) {
  int n = sps * (nbits + guard);
  if (cap < n) return n;
  return modulate_(bits, nbits, gPulse, gPulseLen, 1, guard, sps, (cpx *)out);
}

/* ---- delayVector, sigProcLib.cpp:573-616 ---------------------------------------------------------- */
static void delay_(cpx *v, int n, int v_real, float delay) {
  int intOffset = (int)floor(delay);
  float fracOffset = delay - intOffset;
  cpx *shifted = v;
  if (fabs(fracOffset) > 1e-2) {
    cpx sincv[21];
    for (int i = 0; i < 21; i++) sincv[i] = C(port_sinc(M_PI_F * (i - 10 - fracOffset)), 0.0F);
    shifted = (cpx *)calloc(n, sizeof(cpx));
    convolve_(v, n, v_real, sincv, 21, 1, shifted, PORT_NO_DELAY);
  }
  if (intOffset < 0) {
    intOffset = -intOffset;
    int w = 0;
    for (int s = intOffset; s < n; s++) v[w++] = shifted[s];
    while (w < n) v[w++] = C(0, 0);
  } else {
    int w = n - 1;
    for (int s = n - 1 - intOffset; s >= 0; s--) v[w--] = shifted[s];
    while (w >= 0) v[w--] = C(0, 0);
  }
  if (shifted != v) free(shifted);
}
void port_delay_vector(float *v, int n, float delay) { delay_((cpx *)v, n, 0, delay); }

/* ---- interpolatePoint :639-659, peakDetect :663-711 ---------------------------------------------- */
static cpx interp_(const cpx *sig, int n, float ix) {
  int start = (int)(floor(ix) - 10);
  if (start < 0) start = 0;
  int end = (int)(floor(ix) + 11);
  if ((unsigned long)(unsigned)end > (unsigned long)n - 1) end = n - 1;
  cpx p = C(0, 0);
  for (int i = start; i < end; i++) p = cadd(p, cmulr(sig[i], port_sinc(M_PI_F * (i - ix))));
  return p;
}
void port_interpolate_point(const float *v, int n, float ix, float *out) {
  cpx p = interp_((const cpx *)v, n, ix); out[0] = p.r; out[1] = p.i;
}
static cpx peak_detect_(const cpx *v, int n, float *peakIndex, float *avgPwr) {
  float maxVal = 0.0F, maxIndex = -1, sumPower = 0.0F;
  for (int i = 0; i < n; i++) {
    float p = cnorm2(v[i]);
    if (p > maxVal) { maxVal = p; maxIndex = i; }
    sumPower += p;
  }
  float early = maxIndex - 1, late = maxIndex + 1, incr = 0.5F;
  while (incr > 1.0 / 1024.0) {
    float e = cnorm2(interp_(v, n, early)), l = cnorm2(interp_(v, n, late));
    if (e < l) early += incr;
    else if (e > l) early -= incr;
    else break;
    incr /= 2.0;
    late = early + 2.0;
  }
  maxIndex = early + 1.0;
  cpx pk = interp_(v, n, maxIndex);
  if (peakIndex) *peakIndex = maxIndex;
  if (avgPwr) *avgPwr = (sumPower - cnorm2(pk)) / (float)(unsigned long)(n - 1);
  return pk;
}
void port_peak_detect(const float *v, int n, float *peak, float *idx, float *avg) {
  cpx p = peak_detect_((const cpx *)v, n, idx, avg); peak[0] = p.r; peak[1] = p.i;
}

/* ---- energyDetect, sigProcLib.cpp:916-932 --------------------------------------------------------- */
int port_energy_detect(const float *v, int n, unsigned win, float thr, float *avg) {
  const cpx *x = (const cpx *)v;
  float energy = 0.0F;
  if (win > (unsigned)n) win = n;
  for (unsigned i = 0; i < win; i++) energy += cnorm2(x[i]);
  if (avg) *avg = energy / win;
  return (energy / win > thr * thr) ? 1 : 0;
}

/* ---- analyzeTrafficBurst, sigProcLib.cpp:935-1037 ------------------------------------------------- */
static int analyze_(const cpx *burst, int n, int tsc, float thr, int sps, cpx *amplitude, float *TOA,
                    int request, cpx *chan, float *chanOff) {
  (void)n;
  int L = 36 * sps;
  cpx *corr = (cpx *)calloc(L, sizeof(cpx));
  correlate_(burst + 56 * sps, L, 0, gMid[tsc].seq, gMid[tsc].len, 0, corr, PORT_NO_DELAY);
  float meanPower;
  *amplitude = peak_detect_(corr, L, TOA, &meanPower);
  float valley = 0.0F;
  if ((*TOA < 0.0) || (*TOA > (float)L)) { free(corr); *amplitude = C(0, 0); return 0; }
  int p = (int)rint(*TOA);
  int numRms = 0;
  for (int i = 2 * sps; i <= 5 * sps; i++) {
    if (p - i >= 0) { valley += cnorm2(corr[p - i]); numRms++; }
    if (p + i < L)  { valley += cnorm2(corr[p + i]); numRms++; }
  }
  if (numRms < 2) { free(corr); *amplitude = C(0, 0); return 0; }
  float RMS = sqrtf(valley / (float)numRms) + 0.00001;
  float peakToMean = cabsf_(*amplitude) / RMS;
  *amplitude = cdiv(*amplitude, gMid[tsc].gain);
  *TOA = (*TOA) - gMid[tsc].TOA;
  *TOA = (*TOA) - (66 - 56) * sps;
  if (request && (peakToMean > thr)) {
    float TOAoffset = gMid[tsc].TOA + (66 - 56) * sps;
    delay_(corr, L, 0, -(*TOA));
    int clen = 6 * sps;
    float maxEnergy = -1.0F;
    int maxI = -1;
    for (int i = 0; i < 7; i++) {
      if (TOAoffset + (i - 5) * sps + (float)clen > (float)L) continue;
      if (TOAoffset + (i - 5) * sps < 0) continue;
      const cpx *seg = corr + (int)floor(TOAoffset + (i - 5) * sps);
      float energy = 0.0F;
      for (int j = 0; j < clen; j++) energy += cnorm2(seg[j]);
      if (energy > 0.95 * maxEnergy) { maxI = i; maxEnergy = energy; }
    }
    const cpx *seg = corr + (int)floor(TOAoffset + (maxI - 5) * sps);
    cpx g = cdiv(C(1.0F, 0.0F), gMid[tsc].gain);
    for (int j = 0; j < clen; j++) chan[j] = cmul(seg[j], g);
    if (chanOff) *chanOff = 5 * sps - maxI;
  }
  free(corr);
  return peakToMean > thr;
}
int port_analyze(const float *burst, int n, int tsc, float thr, int sps, float *amp, float *toa,
                 int request, float *chan, float *off) {
  cpx a = C(0, 0); float t = 0.0F, o = 0.0F;
  cpx ch[64]; memset(ch, 0, sizeof(ch));
  int ok = analyze_((const cpx *)burst, n, tsc, thr, sps, &a, &t, request, ch, &o);
  amp[0] = a.r; amp[1] = a.i; *toa = t;
  if (request && ok) { if (chan) memcpy(chan, ch, 6 * sps * sizeof(cpx)); if (off) *off = o; }
  return ok;
}

/* ---- detectRACHBurst, sigProcLib.cpp:860-914 ------------------------------------------------------ */
static int detect_rach_(const cpx *burst, int n, float thr, int sps, cpx *amplitude, float *TOA) {
  cpx *corr = (cpx *)calloc(n, sizeof(cpx));
  correlate_(burst, n, 0, gRach.seq, gRach.len, 0, corr, PORT_NO_DELAY);
  float meanPower;
  cpx peakAmpl = peak_detect_(corr, n, TOA, &meanPower);
  float valley = 0.0F;
  if ((*TOA < 0.0) || (*TOA > (float)n)) { free(corr); *amplitude = C(0, 0); return 0; }
  int p = (int)rint(*TOA);
  float numSamples = 0.0F;
  for (int i = 57 * sps; i <= 107 * sps; i++) {
    if (p + i >= n) break;
    valley += cnorm2(corr[p + i]);
    numSamples++;
  }
  free(corr);
  if (numSamples < 2) { *amplitude = C(0, 0); return 0; }
  float RMS = sqrtf(valley / (float)numSamples) + 0.00001;
  float peakToMean = cabsf_(peakAmpl) / RMS;
  *amplitude = cdiv(peakAmpl, gRach.gain);
  *TOA = (*TOA) - gRach.TOA - 8 * sps;
  return peakToMean > thr;
}
int port_detect_rach(const float *burst, int n, float thr, int sps, float *amp, float *toa) {
  cpx a = C(0, 0); float t = 0.0F;
  int ok = detect_rach_((const cpx *)burst, n, thr, sps, &a, &t);
  amp[0] = a.r; amp[1] = a.i; *toa = t;
  return ok;
}

/* ---- designDFE, sigProcLib.cpp:1246-1340 (Al-Dhahir/Cioffi fast Cholesky recursion) -------------- */
static int design_dfe_(const cpx *chan, int nchan, float SNR, int Nf, cpx *w, cpx *b) {
  int nu = nchan - 1;
  int Ll = Nf + nu;
  cpx *G0 = (cpx *)calloc(Nf, sizeof(cpx)), *G1 = (cpx *)calloc(Nf, sizeof(cpx));
  cpx *G0n = (cpx *)calloc(Nf, sizeof(cpx)), *G1n = (cpx *)calloc(Nf, sizeof(cpx));
  cpx *L = (cpx *)calloc((size_t)Nf * Ll, sizeof(cpx));
  cpx *v = (cpx *)calloc(Nf, sizeof(cpx));
  G0[0] = C((float)(1.0 / sqrtf(SNR)), 0.0F);
  for (int j = 0; j <= nu && j < Nf; j++) G1[j] = cconj(chan[j]);
  float d = 0.0F;
  for (int i = 0; i < Nf; i++) {
    d = cnorm2(G0[0]) + cnorm2(G1[0]);
    cpx *Li = L + (size_t)i * Ll;
    for (int j = 0; j < Nf && i + j < Ll; j++)
      Li[i + j] = cdivr(cadd(cmul(G0[j], cconj(G0[0])), cmul(G1[j], cconj(G1[0]))), d);
    cpx k = cdiv(G1[0], G0[0]);
    if (i != Nf - 1) {
      cpx kc = cconj(k), nk = cmulr(k, (float)(-1.0));
      for (int j = 0; j < Nf; j++) G0n[j] = cadd(cmul(G1[j], kc), G0[j]);
      for (int j = 0; j < Nf; j++) G1n[j] = cadd(cmul(G0[j], nk), G1[j]);
      for (int j = 0; j + 1 < Nf; j++) G1n[j] = G1n[j + 1];       /* delayVector(G1new,-1.0) */
      G1n[Nf - 1] = C(0, 0);
      cpx s = C((float)(1.0 / sqrtf((float)(1.0 + cnorm2(k)))), 0.0F);            /* :1294-1295 */
      for (int j = 0; j < Nf; j++) { G0[j] = cmul(G0n[j], s); G1[j] = cmul(G1n[j], s); }
    }
  }
  const cpx *Llast = L + (size_t)(Nf - 1) * Ll;
  for (int j = 0; j < nu; j++) b[j] = cconj(cmul(Llast[Nf + j], C(-1.0F, 0.0F)));
  v[Nf - 1] = C(1.0F, 0.0F);
  for (int k = Nf - 2; k >= 0; k--) {
    const cpx *Lk = L + (size_t)k * Ll;
    cpx vk = C(0, 0);
    for (int j = k + 1; j < Nf; j++) { cpx pr = cmul(v[j], Lk[j]); vk.r -= pr.r; vk.i -= pr.i; }
    v[k] = vk;
  }
  for (int i = 0; i < Nf; i++) {
    cpx wi = C(0, 0);
    int endPt = (nu < (Nf - 1 - i)) ? nu : (Nf - 1 - i);
    for (int k = 0; k < endPt + 1; k++) wi = cadd(wi, cmul(v[i + k], cconj(chan[k])));
    w[i] = cdivr(wi, d);
  }
  free(G0); free(G1); free(G0n); free(G1n); free(L); free(v);
  return 1;
}
int port_design_dfe(const float *chan, int nchan, float snr, int Nf, float *w, float *b) {
  return design_dfe_((const cpx *)chan, nchan, snr, Nf, (cpx *)w, (cpx *)b);
}

/* ---- vectorSlicer :507-519 ------------------------------------------------------------------------ */
static float slice_(float x) {
  float s = (float)(0.5 * (x + 1.0F));
  if (s > 1.0) s = 1.0F;
  if (s < 0.0) s = 0.0F;
  return s;
}

/* ---- equalizeBurst, sigProcLib.cpp:1343-1399 ------------------------------------------------------ */
static int equalize_(cpx *burst, int n, float TOA, int sps, const cpx *w, int nw, const cpx *b, int nb, float *soft) {
  (void)sps;
  delay_(burst, n, 0, -TOA);
  cpx *full = (cpx *)calloc(n + nw - 1, sizeof(cpx));
  convolve_(burst, n, 0, w, nw, 0, full, PORT_FULL_SPAN);
  cpx *y = full + (nw - 1);
  for (int m = 0; m < n; m++) {
    for (int k = 0; k < nb && m - 1 - k >= 0; k++) y[m] = cadd(y[m], cmul(b[k], y[m - 1 - k]));
    y[m] = cmul(y[m], gRevRot[m]);
    float out = y[m].r;
    y[m] = cmul(C((y[m].r > 0.0) ? 1.0F : -1.0F, 0.0F), gRot[m]);
    soft[m] = slice_(out);
  }
  free(full);
  return n;
}
int port_equalize(float *burst, int n, float toa, int sps, const float *w, int nw, const float *b, int nb, float *soft) {
  return equalize_((cpx *)burst, n, toa, sps, (const cpx *)w, nw, (const cpx *)b, nb, soft);
}

/* ---- demodulateBurst :1056-1097 with decimateVector :1039-1053 ------------------------------------ */
static int demodulate_(const cpx *burst, int n, int sps, cpx channel, float TOA, float *soft) {
  cpx *x = (cpx *)malloc(n * sizeof(cpx));
  memcpy(x, burst, n * sizeof(cpx));
  scale_(x, n, 0, cdiv(C(1.0F, 0.0F), channel));
  delay_(x, n, 0, -TOA);
  for (int i = 0; i < n; i++) x[i] = cmul(gRevRot[i], x[i]);
  int m = (sps > 1) ? n / sps : n;
  for (int i = 0; i < m; i++) soft[i] = slice_(x[i * sps].r);
  free(x);
  return m;
}
int port_demodulate(const float *burst, int n, int sps, const float *amp, float toa, float *soft) {
  return demodulate_((const cpx *)burst, n, sps, C(amp[0], amp[1]), toa, soft);
}

/* ---- createLPF :1102-1150, polyphaseResampleVector :1157-1210 ------------------------------------ */
static void create_lpf_(const unsigned int *bits, int len, float gainDC, float *out) {
  double sum = 0.0;
  for (int i = 0; i < len; i++) { float t; memcpy(&t, &bits[i], 4); out[i] = t; sum += t; }
  float norm = gainDC / sum;
  for (int i = 0; i < len; i++) out[i] = out[i] * norm;
}
static int resample_(const cpx *x, int n, int P, int Q, const float *lpf, int L, cpx *out) {
  int outn = (int)ceil(n * (float)P / (float)Q);
  int outputIx = (L - 1) / 2 / Q;
  for (int o = 0; o < outn; o++, outputIx++) {
    int branch = (outputIx * Q) % P;
    int in = (outputIx * Q - branch) / P;
    int f = branch;
    while (in >= n) { in--; f += P; }
    cpx sum = C(0, 0);
    while (in >= 0 && f < L) { sum = cadd(sum, cmulr(x[in], lpf[f])); in--; f += P; }
    out[o] = sum;
  }
  return outn;
}
int port_resample(const float *x, int n, int P, int Q, int lpf, float *out, int cap) {
  int outn = (int)ceil(n * (float)P / (float)Q);
  if (cap < outn) return outn;
  return resample_((const cpx *)x, n, P, Q, lpf ? gLpfTx : gLpfRx, lpf ? 651 : 961, (cpx *)out);
}

/* ---- setup: sigProcLibSetup :207-230, generateGSMPulse :411-430, generateMidamble :779-828,
 *      generateRACHSequence :830-857 --------------------------------------------------------------- */
static void gen_corrseq_midamble(int t, int sps) {
  char bits[26];
  for (int i = 0; i < 26; i++) bits[i] = kTSC[t][i] == '1';
  cpx one = C(1.0F, 0.0F);
  int nmid = 16 * sps, nfull = 26 * sps;
  cpx *middle = (cpx *)calloc(nmid, sizeof(cpx)), *mid = (cpx *)calloc(nfull, sizeof(cpx));
  modulate_(bits + 5, 16, &one, 1, 0, 0, sps, middle);
  modulate_(bits, 26, gPulse, gPulseLen, 1, 0, sps, mid);
  scale_(middle, nmid, 0, C(-1.0F, 0.0F));
  scale_(mid, nfull, 0, C(0.0F, 1.0F));
  cpx *ac = (cpx *)calloc(nfull, sizeof(cpx));
  correlate_(mid, nfull, 0, middle, nmid, 0, ac, PORT_NO_DELAY);
  free(gMid[t].seq);
  gMid[t].seq = middle; gMid[t].len = nmid;
  gMid[t].gain = peak_detect_(ac, nfull, &gMid[t].TOA, NULL);
  gMid[t].TOA -= 5 * sps;
  free(ac); free(mid);
}
int port_setup(int sps) {
  M_1_2PI_F = 1 / M_2PI_F;
  gSps = sps;
  for (int i = 0; i < TABLESIZE + 1; i++) {
    cosTable[i] = cos(2.0 * M_PI * i / TABLESIZE);
    sinTable[i] = sin(2.0 * M_PI * i / TABLESIZE);
  }
  free(gRot); free(gRevRot);
  gRot = (cpx *)calloc(157 * sps, sizeof(cpx)); gRevRot = (cpx *)calloc(157 * sps, sizeof(cpx));
  float phase = 0.0F;
  for (int i = 0; i < 157 * sps; i++) {
    gRot[i] = expj_lookup(phase);
    gRevRot[i] = expj_lookup(-phase);
    phase += M_PI_F / 2.0F / (float)sps;
  }
  /* generateGSMPulse(symbolLength=2, sps) */
  gPulseLen = sps * 2 + 1;
  int center = (gPulseLen - 1) / 2;
  float e = 0.0F;
  for (int i = 0; i < gPulseLen; i++) {
    float arg = (float)(i - center) / (float)sps;
    gPulse[i] = C((float)(0.96 * exp(-1.1380 * arg * arg - 0.527 * arg * arg * arg * arg)), 0.0F);
  }
  for (int i = 0; i < gPulseLen; i++) e += cnorm2(gPulse[i]);
  float avgAbs = sqrtf(e / sps);
  for (int i = 0; i < gPulseLen; i++) gPulse[i] = cdivr(gPulse[i], avgAbs);
  for (int t = 0; t < 8; t++) gen_corrseq_midamble(t, sps);
  {
    char bits[41];
    for (int i = 0; i < 41; i++) bits[i] = kRACH[i] == '1';
    int n = 41 * sps;
    cpx *seq = (cpx *)calloc(n, sizeof(cpx)), *ac = (cpx *)calloc(n, sizeof(cpx));
    modulate_(bits, 41, gPulse, gPulseLen, 1, 0, sps, seq);
    correlate_(seq, n, 0, seq, n, 0, ac, PORT_NO_DELAY);
    free(gRach.seq);
    gRach.seq = seq; gRach.len = n;
    gRach.gain = peak_detect_(ac, n, &gRach.TOA, NULL);
    free(ac);
  }
  create_lpf_(LPF961_BITS, 961, 65, gLpfRx);
  create_lpf_(LPF651_BITS, 651, 96, gLpfTx);
  return 0;
}

int port_get_table(int id, int idx, float *dst, int cap) {
  int n = 0;
  switch (id) {
    case 0: n = 1025; if (cap >= n) memcpy(dst, cosTable, n * 4); break;
    case 1: n = 1025; if (cap >= n) memcpy(dst, sinTable, n * 4); break;
    case 2: n = 2 * 157 * gSps; if (cap >= n) memcpy(dst, gRot, n * 4); break;
    case 3: n = 2 * 157 * gSps; if (cap >= n) memcpy(dst, gRevRot, n * 4); break;
    case 4: n = 2 * gPulseLen; if (cap >= n) memcpy(dst, gPulse, n * 4); break;
    case 5: n = 2 * gMid[idx].len; if (cap >= n) memcpy(dst, gMid[idx].seq, n * 4); break;
    case 6: n = 3; if (cap >= n) { dst[0] = gMid[idx].TOA; dst[1] = gMid[idx].gain.r; dst[2] = gMid[idx].gain.i; } break;
    case 7: n = 2 * gRach.len; if (cap >= n) memcpy(dst, gRach.seq, n * 4); break;
    case 8: n = 3; if (cap >= n) { dst[0] = gRach.TOA; dst[1] = gRach.gain.r; dst[2] = gRach.gain.i; } break;
    case 9: n = 961; if (cap >= n) memcpy(dst, gLpfRx, n * 4); break;
    case 10: n = 651; if (cap >= n) memcpy(dst, gLpfTx, n * 4); break;
    case 11: n = 1; if (cap >= n) dst[0] = 0.0F; break;
    default: return -1;
  }
  return n;
}

/* ---- caller glue (see oracle/ref_shim.cpp for the reference lines) ------------------------------- */
void port_rx_normal_batch(const float *bursts, int pitch, const int *lens, const unsigned char *tsc, long n,
                          float detect_thr, float energy_thr,
                          int *flags, float *amp, float *toa, float *chan, float *off, float *w, float *b,
                          float *soft, int soft_pitch) {
  for (long i = 0; i < n; i++) {
    int len = lens[i];
    cpx burst[160];
    memcpy(burst, bursts + 2 * (size_t)pitch * i, len * sizeof(cpx));
    cpx a = C(0, 0), ch[6], W[7], B[5]; float t = 0.0F, o = 0.0F;
    memset(ch, 0, sizeof(ch));
    int ok = analyze_(burst, len, tsc[i], detect_thr, 1, &a, &t, 1, ch, &o);
    flags[i] = ok; amp[2 * i] = a.r; amp[2 * i + 1] = a.i; toa[i] = t;
    float *sp = soft + (size_t)soft_pitch * i;
    memset(sp, 0, soft_pitch * sizeof(float));
    if (off) off[i] = 0.0F;
    if (chan) memset(chan + 12 * i, 0, 12 * sizeof(float));
    if (w) memset(w + 14 * i, 0, 14 * sizeof(float));
    if (b) memset(b + 10 * i, 0, 10 * sizeof(float));
    if (!ok) continue;
    float SNR = cnorm2(a) / (energy_thr * energy_thr + 1.0);            /* Transceiver.cpp:340 */
    cpx ia = cdiv(C(1.0F, 0.0F), a);
    scale_(ch, 6, 0, ia);                                               /* :346 */
    design_dfe_(ch, 6, SNR, 7, W, B);                                   /* :347 */
    scale_(burst, len, 0, ia);                                          /* :391 */
    equalize_(burst, len, t - o, 1, W, 7, B, 5, sp);                    /* :392-396 */
    if (off) off[i] = o;
    if (chan) memcpy(chan + 12 * i, ch, sizeof(ch));
    if (w) memcpy(w + 14 * i, W, sizeof(W));
    if (b) memcpy(b + 10 * i, B, sizeof(B));
  }
}

void port_rx_rach_batch(const float *bursts, int pitch, const int *lens, long n, float detect_thr, int sps,
                        int *flags, float *amp, float *toa, float *soft, int soft_pitch) {
  for (long i = 0; i < n; i++) {
    int len = lens[i];
    const cpx *burst = (const cpx *)(bursts + 2 * (size_t)pitch * i);
    cpx a = C(0, 0); float t = 0.0F;
    int ok = detect_rach_(burst, len, detect_thr, sps, &a, &t);
    flags[i] = ok; amp[2 * i] = a.r; amp[2 * i + 1] = a.i; toa[i] = t;
    float *sp = soft + (size_t)soft_pitch * i;
    memset(sp, 0, soft_pitch * sizeof(float));
    if (ok) demodulate_(burst, len, sps, a, t, sp);
  }
}

void port_rx_resample_stream(const float *raw, long first_chunk, long nchunks, float *out) {
  cpx input[192 + 864], res[720];
  for (long c = 0; c < nchunks; c++) {
    const float *src = raw + 2 * 864 * c;
    if (first_chunk + c == 0) {
      memset(input, 0, 192 * sizeof(cpx));
      memcpy(input + 192, src, 864 * sizeof(cpx));
    } else {
      memcpy(input, src - 2 * 192, (192 + 864) * sizeof(cpx));
    }
    resample_(input, 192 + 864, 65, 96, gLpfRx, 961, res);
    memcpy(out + 2 * 585 * c, res + 130, 585 * sizeof(cpx));
  }
}

/* radioInterface.cpp:91-116 (unUSRPifyVector) in front of the same resample; see oracle/ref_shim.cpp */
void port_rx_resample_stream_i16(const short *iq, int flip_iq, long first_chunk, long nchunks, float *out) {
  cpx input[192 + 864], res[720];
  for (long c = 0; c < nchunks; c++) {
    const short *src = iq + 2 * 864 * c;
    const int hist = (first_chunk + c == 0) ? 0 : 192;
    if (!hist) memset(input, 0, 192 * sizeof(cpx));
    const short *sp = src - 2 * hist;
    for (int i = 192 - hist; i < 192 + 864; i++, sp += 2) input[i] = C((float)sp[flip_iq], (float)sp[1 - flip_iq]);
    resample_(input, 192 + 864, 65, 96, gLpfRx, 961, res);
    memcpy(out + 2 * 585 * c, res + 130, 585 * sizeof(cpx));
  }
}

/* Transceiver.cpp:667-669: (char) round(soft*255.0), 148 per burst */
void port_soft_to_wire(const float *soft, int soft_pitch, long n, unsigned char *out) {
  for (long i = 0; i < n; i++)
    for (int k = 0; k < 148; k++) out[148 * i + k] = (unsigned char)(char)round(soft[(size_t)soft_pitch * i + k] * 255.0);
}

void port_tx_resample_stream(const float *in, long first_chunk, long nchunks, short *out) {
  cpx input[130 + 585], res[1060];
  for (long c = 0; c < nchunks; c++) {
    const float *src = in + 2 * 585 * c;
    if (first_chunk + c == 0) {
      memset(input, 0, 130 * sizeof(cpx));
      memcpy(input + 130, src, 585 * sizeof(cpx));
    } else {
      memcpy(input, src - 2 * 130, (130 + 585) * sizeof(cpx));
    }
    resample_(input, 130 + 585, 96, 65, gLpfTx, 651, res);
    short *o = out + 2 * 864 * c;
    for (int i = 0; i < 864; i++) {
      cpx s = cmul(res[192 + i], C(13500.0F, 0.0F));                   /* scaleVector(.,13500.0) */
      o[2 * i] = (short)s.r;
      o[2 * i + 1] = (short)s.i;
    }
  }
}

long port_modulate_stream(const char *bits148, long nbursts, int tn0, float *out) {
  long pos = 0;
  for (long i = 0; i < nbursts; i++) {
    int guard = 8 + (((tn0 + i) % 8) % 4 == 0);
    pos += modulate_(bits148 + 148 * i, 148, gPulse, gPulseLen, 1, guard, 1, (cpx *)out + pos);
  }
  return pos;
}

void port_rx_stream_demod(const float *resampled, long first_burst, long nbursts, const unsigned char *tsc,
                          float detect_thr, float energy_thr, int *flags, float *amp, float *toa,
                          float *soft, int soft_pitch) {
  static const int slot_off[4] = {0, 157, 313, 469};
  for (long i = 0; i < nbursts; i++) {
    long g = first_burst + i;
    long start = (g / 4) * 625 + slot_off[g % 4];
    int len = (g % 4 == 0) ? 157 : 156;
    port_rx_normal_batch(resampled + 2 * start, 0, &len, tsc + i, 1, detect_thr, energy_thr,
                         flags + i, amp + 2 * i, toa + i, NULL, NULL, NULL, NULL,
                         soft + (size_t)soft_pitch * i, soft_pitch);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Transceiver::pullRadioVector + driveReceiveFIFO restated (Transceiver.cpp:207-269, 271-410, 641-676):
 * slot map, adaptive energy threshold, 50-frame channel/DFE cache, RSSI / timing integerisation, RX datagram.
 * Same state layout and entry points as oracle/ref_shim.cpp's ref_trx_*.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
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
} port_trx_state;
enum { CT_NONE = 0, CT_I, CT_II, CT_III, CT_IV, CT_V, CT_VI, CT_VII, CT_LOOPBACK };
enum { CORR_OFF = 0, CORR_TSC, CORR_RACH, CORR_IDLE };
#define HYPERFRAME (2048 * 26 * 51)
static int fn_delta(int v1, int v2) {                     /* GSM::FNDelta, GSMCommon.cpp:161-168 */
  const int half = HYPERFRAME / 2;
  int d = v1 - v2;
  if (d >= half) d -= HYPERFRAME; else if (d < -half) d += HYPERFRAME;
  return d;
}
int port_expected_corr_type(int chan_type, int fn) {      /* Transceiver.cpp:207-269 */
  int m = fn % 51;
  switch (chan_type) {
    case CT_NONE: return CORR_OFF;
    case CT_I: return CORR_TSC;
    case CT_II: return (fn % 2 == 1) ? CORR_IDLE : CORR_TSC;
    case CT_III: return CORR_TSC;
    case CT_IV: case CT_VI: return (m % 10 < 2) ? CORR_RACH : CORR_OFF;
    case CT_V:
      if ((m <= 36 && m >= 14) || m == 4 || m == 5 || m == 45 || m == 46) return CORR_RACH;
      return CORR_TSC;
    case CT_VII: return (m == 12 || m == 13 || m == 14) ? CORR_IDLE : CORR_TSC;
    case CT_LOOPBACK: return (m <= 50 && m >= 48) ? CORR_IDLE : CORR_TSC;
    default: return CORR_OFF;
  }
}
int port_trx_state_bytes(void) { return (int)sizeof(port_trx_state); }
void port_trx_init(void *state, int tsc, const int *chan_type, int start_fn) {   /* Transceiver.cpp:40-90 */
  port_trx_state *st = (port_trx_state *)state;
  memset(st, 0, sizeof *st);
  st->thr = 250.0;
  st->prev_false_fn = start_fn;
  st->tsc = tsc;
  for (int i = 0; i < 8; i++) { st->chan_type[i] = chan_type[i]; st->est_fn[i] = start_fn; st->have[i] = 0; }
}
void port_trx_pull(void *state, const float *bursts, int pitch, int nframes, int fn0,
                   int *valid, unsigned char *dgram, int dgram_pitch) {
  port_trx_state *st = (port_trx_state *)state;
  for (int f = 0; f < nframes; f++) {
    const int fn = (fn0 + f) % HYPERFRAME;
    for (int tn = 0; tn < 8; tn++) {
      const long i = (long)f * 8 + tn;
      unsigned char *dg = dgram + (size_t)dgram_pitch * i;
      valid[i] = 0;
      memset(dg, 0, 158);
      const int corr = port_expected_corr_type(st->chan_type[tn], fn);
      if (corr == CORR_OFF || corr == CORR_IDLE) continue;                               /* :290-293 */
      const int len = (tn % 4 == 0) ? 157 : 156;
      cpx burst[160];
      memcpy(burst, bursts + 2 * (size_t)pitch * i, len * sizeof(cpx));
      cpx amplitude = C(0, 0);
      float TOA = 0.0F, avgPwr = 0.0F;
      if (!port_energy_detect((const float *)burst, len, 20, (float)st->thr, &avgPwr)) { /* :298 */
        double framesElapsed = fn_delta(fn, st->prev_false_fn);
        if (framesElapsed > 50) { st->thr -= 10.0; st->prev_false_fn = fn; }           /* :300-304 */
        continue;
      }
      int success = 0;
      if (corr == CORR_TSC) {
        double framesElapsed = fn_delta(fn, st->est_fn[tn]);
        int estimateChannel = 0;
        if (framesElapsed > 50 || !st->have[tn]) { st->have[tn] = 0; estimateChannel = 1; }   /* :317-326 */
        cpx ch[6];
        float chanOffset = 0.0F;
        memset(ch, 0, sizeof ch);
        success = analyze_(burst, len, st->tsc, 3.0F, 1, &amplitude, &TOA, estimateChannel, ch, &chanOffset);
        if (success) {
          st->thr -= 1.0F;
          if (st->thr < 0.0) st->thr = 0.0;
          st->snr[tn] = (float)(cnorm2(amplitude) / (st->thr * st->thr + 1.0));          /* :340 */
          if (estimateChannel) {
            st->have[tn] = 1;
            st->chan_off[tn] = chanOffset;
            scale_(ch, 6, 0, cdiv(C(1.0F, 0.0F), amplitude));                            /* :346 */
            design_dfe_(ch, 6, st->snr[tn], 7, (cpx *)st->w[tn], (cpx *)st->b[tn]);      /* :347 */
            st->est_fn[tn] = fn;
          }
        } else {
          double fe = fn_delta(fn, st->prev_false_fn);
          st->thr += 10.0F * exp(-fe);                                                   /* :355 */
          st->prev_false_fn = fn;
          st->have[tn] = 0;                                                              /* :357 */
        }
      } else {
        success = detect_rach_(burst, len, 5.0F, 1, &amplitude, &TOA);
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
      float soft[160];
      memset(soft, 0, sizeof soft);
      if (corr == CORR_RACH) {
        port_demodulate((const float *)burst, len, 1, (const float *)&amplitude, TOA, soft);
      } else {
        scale_(burst, len, 0, cdiv(C(1.0F, 0.0F), amplitude));                           /* :391 */
        equalize_(burst, len, TOA - st->chan_off[tn], 1, (const cpx *)st->w[tn], 7, (const cpx *)st->b[tn], 5, soft);
      }
      const int RSSI = (int)floor(20.0 * log10(9450.0 / cabsf_(amplitude)));              /* :400 */
      const int timingOffset = (int)round(TOA * 256.0 / 1);                              /* :402 */
      valid[i] = 1;
      dg[0] = (unsigned char)tn;                                                         /* :659-673 */
      for (int k = 0; k < 4; k++) dg[1 + k] = (fn >> ((3 - k) * 8)) & 0x0ff;
      dg[5] = (unsigned char)RSSI;
      dg[6] = (timingOffset >> 8) & 0x0ff;
      dg[7] = timingOffset & 0x0ff;
      for (int k = 0; k < 148; k++) dg[8 + k] = (unsigned char)((int)round(soft[k] * 255.0) & 0xff);   /* what the x86 (char) cast yields */
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * L1 FEC after the path: XCCHL1Decoder::deinterleave + decode (GSML1FEC.cpp:616-660) restated, with
 * SoftVector::decode + ViterbiR2O4 (CommonLibs/BitVector.cpp:290-540) and the Parity/Generator syndrome
 * (BitVector.h:39-112).  Same entry point as oracle/ref_shim.cpp's ref_xcch_decode.
 * ------------------------------------------------------------------------------------------------ */
static unsigned apply_poly(unsigned val, unsigned poly, unsigned order) {      /* BitVector.cpp:41-47 */
  unsigned prod = val & poly, sum = prod;
  for (unsigned i = 1; i < order; i++) sum ^= prod >> i;
  return sum & 1u;
}
/* SoftVector::decode (BitVector.cpp:451-540) for nc coded probabilities c[] -> nc/2 decoded bits u[] */
static void viterbi_decode(const float *c, int nc, unsigned char *up) {
  enum { ORDER = 4, STATES = 16, CANDS = 32, DEFER = 24, MAXC = 456, CT = 456 + 2 * 24 };
  static unsigned gen[CANDS];
  static int have_gen = 0;
  if (!have_gen) {
    for (unsigned idx = 0; idx < CANDS; idx++) gen[idx] = (apply_poly(idx, 0x019, ORDER + 1) << 1) | apply_poly(idx, 0x01b, ORDER + 1);
    have_gen = 1;
  }
  const int nu = nc / 2, ct = nc + 2 * DEFER;
  float match[CT], mismatch[CT];
  unsigned history[CT];
  unsigned accum = 0;
  for (int i = 0; i < nc; i++) { accum = (accum << 1) | (c[i] > 0.5F ? 1u : 0u); history[i] = accum; }   /* :452-461 */
  for (int i = nc; i < ct; i++) { accum = (accum << 1) | (accum & 1u); history[i] = accum; }
  for (int i = 0; i < nc; i++) {                                               /* :476-490 */
    float pVal = c[i];
    if (pVal > 0.5F) pVal = 1.0F - pVal;
    float ipVal = 1.0F - pVal;
    if (pVal < 0.01F) pVal = 0.01;
    if (ipVal < 0.01F) ipVal = 0.01;
    match[i] = 0.25F / ipVal;
    mismatch[i] = 0.25F / pVal;
  }
  for (int i = nc; i < ct; i++) { match[i] = 0.5F; mismatch[i] = 0.5F; }
  float scost[STATES] = {0}, ccost[CANDS];
  unsigned sin_[STATES] = {0}, sout[STATES] = {0}, cin[CANDS], cout[CANDS];
  for (int s = 0; s < nu + DEFER; s++) {
    const unsigned in = history[2 * s + 1];
    const float *m0 = match + 2 * s, *m1 = mismatch + 2 * s;
    for (int i = 0; i < CANDS; i += 2) {                                       /* branchCandidates */
      const int sp = i / 2;
      const unsigned i0 = sin_[sp] << 1, i1 = i0 | 1u, osh = sout[sp] << 2;
      ccost[i] = scost[sp]; cout[i] = osh | gen[i0 & 0x1f]; cin[i] = i0;
      ccost[i + 1] = scost[sp]; cout[i + 1] = osh | gen[i1 & 0x1f]; cin[i + 1] = i1;
    }
    for (int i = 0; i < CANDS; i++) {                                          /* getSoftCostMetrics */
      const unsigned mm = in ^ cout[i];
      ccost[i] += ((mm & 1u) ? m1 : m0)[1] + (((mm >> 1) & 1u) ? m1 : m0)[0];
    }
    for (int i = 0; i < STATES; i++) {                                         /* pruneCandidates */
      const int w = ccost[i] < ccost[i + STATES] ? i : i + STATES;
      scost[i] = ccost[w]; sin_[i] = cin[w]; sout[i] = cout[w];
    }
    int best = 0;                                                              /* minCost */
    float bc = scost[0];
    for (int i = 1; i < STATES; i++) { if (scost[i] >= bc) continue; bc = scost[i]; best = i; }
    if (s >= DEFER) up[s - DEFER] = (sin_[best] >> DEFER) & 1u;
  }
}
/* Fire-code syndrome of d[184] : ~p[40], kept in an `unsigned` as GSML1FEC.cpp:652 does */
static int xcch_syndrome_ok(const unsigned char *up) {
  unsigned long long state = 0;
  for (int i = 0; i < 224; i++) {
    const unsigned bit = (i < 184 ? up[i] : ~up[i]) & 1u;
    const unsigned fb = (unsigned)(state >> 39) & 1u;
    state = (state << 1) ^ bit;
    if (fb) state ^= 0x10004820009ULL;
  }
  return (unsigned)(state & ((1ULL << 40) - 1)) == 0u;
}
void port_xcch_decode(const unsigned char *soft, int burst_pitch, long nframes, unsigned char *u, int *ok) {
  for (long f = 0; f < nframes; f++) {
    float c[456];
    for (int k = 0; k < 456; k++) {                                            /* GSML1FEC.cpp:620-624, :603-604 */
      int B = k % 4, j = 2 * ((49 * k) % 57) + ((k % 8) / 4);
      const unsigned char *rp = soft + (size_t)burst_pitch * (4 * f + B);
      c[k] = rp[j < 57 ? 3 + j : 88 + (j - 57)] / 256.0F;                       /* TRXManager.cpp:230 */
    }
    viterbi_decode(c, 456, u + 228 * f);
    ok[f] = xcch_syndrome_ok(u + 228 * f);
  }
}
/* TCHFACCHL1Decoder::processBurst / deinterleave / decodeTCH + XCCHL1Decoder::decode for stolen blocks (GSML1FEC.cpp:1031-1210),
 * same entry point and block numbering as oracle/ref_shim.cpp's ref_tch_decode (block q = bursts 4q .. 4q+7). */
void port_tch_decode(const unsigned char *soft, int burst_pitch, long nblocks, unsigned char *d, int *good, int *stolen_o,
                     unsigned char *fu, int *fok) {
  for (long q = 0; q < nblocks; q++) {
    float c[456];
    for (int k = 0; k < 456; k++) {                                            /* :1102-1110 */
      const int r = k % 8, j = 2 * ((49 * k) % 57) + (r / 4);
      const unsigned char *rp = soft + (size_t)burst_pitch * (4 * q + r);
      c[k] = rp[j < 57 ? 3 + j : 88 + (j - 57)] / 256.0F;
    }
    const int stolen = (soft[(size_t)burst_pitch * (4 * q + 7) + 60] / 256.0F) > 0.5F;   /* inBurst.Hl(), :1075 */
    stolen_o[q] = stolen;
    memset(d + 260 * q, 0, 260);
    memset(fu + 228 * q, 0, 228);
    good[q] = 0; fok[q] = 0;
    if (stolen) {
      viterbi_decode(c, 456, fu + 228 * q);
      fok[q] = xcch_syndrome_ok(fu + 228 * q);
    } else {                                                                   /* decodeTCH(false), :1129-1160 */
      unsigned char u[189], *dd = d + 260 * q;
      viterbi_decode(c, 378, u);
      for (int i = 0; i < 78; i++) dd[182 + i] = c[378 + i] > 0.5F;
      for (int k = 0; k <= 90; k++) { dd[2 * k] = u[k]; dd[2 * k + 1] = u[184 - k]; }
      const unsigned sent = (~((u[91] << 2) | (u[92] << 1) | u[93])) & 7u;
      unsigned state = 0;
      for (int i = 0; i < 50; i++) {
        const unsigned fb = ((state >> 2) ^ dd[i]) & 1u;
        state <<= 1;
        if (fb) state ^= 0x0bu;
      }
      const unsigned tail = (u[185] << 3) | (u[186] << 2) | (u[187] << 1) | u[188];
      good[q] = sent == (state & 7u) && tail == 0;
    }
  }
}
