// host/sigProcLib.cpp -- the reference's sigProcLib.h functions as thin calls into libbtsdsp.so.
//
// A maintainer replaces Transceiver/sigProcLib.cpp by this file (+ -lbtsdsp) and keeps Transceiver.cpp and
// radioInterface.cpp as they are: every function below has the reference's signature, ownership rules and
// return conventions (reference Transceiver/sigProcLib.cpp, lines cited per function).  All arithmetic runs
// in the CUDA kernels; this file only moves vectors across the C ABI.  Element-wise helpers that the
// reference applies to whole vectors on the CPU (scaleVector, addVector, ...) also go through the library so
// results stay bit-identical to the batched path.
#include "sigProcLib.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "btsdsp.h"

static btsdsp_ctx *gCtx = NULL;
static int gDevice = -1;
static int gSps = 0;

static btsdsp_ctx *ctx() {
  if (!gCtx) {
    fprintf(stderr, "sigProcLib(btsdsp): sigProcLibSetup() has not been called\n");
    abort();
  }
  return gCtx;
}
static void check(int rc, const char *what) {
  if (rc < 0) {
    fprintf(stderr, "sigProcLib(btsdsp): %s failed (%d): %s\n", what, rc, btsdsp_last_error(gCtx));
    abort();     // the reference has no error channel for these; a CUDA failure must not pass silently
  }
}
static const btsdsp_cf32 *cp(const signalVector &v) { return (const btsdsp_cf32 *)v.begin(); }
static btsdsp_cf32 *mp(signalVector &v) { return (btsdsp_cf32 *)v.begin(); }

/** which GPU sigProcLibSetup binds to: this setter, else $BTSDSP_DEVICE, else 0 (not part of the reference surface) */
extern "C" void sigProcLibSetDevice(int device) { gDevice = device; }

/* sigProcLib.cpp:227 -- also builds what generateMidamble / generateRACHSequence / createLPF produce */
void sigProcLibSetup(int samplesPerSymbol) {
  if (gCtx && gSps == samplesPerSymbol) return;
  sigProcLibDestroy();
  int dev = gDevice;
  if (dev < 0) { const char *e = getenv("BTSDSP_DEVICE"); dev = e ? atoi(e) : 0; }
  int rc = btsdsp_create(&gCtx, dev, samplesPerSymbol);
  if (rc != BTSDSP_OK) {
    fprintf(stderr, "sigProcLib(btsdsp): btsdsp_create failed (%d): %s\n", rc, btsdsp_last_error(NULL));
    abort();
  }
  gSps = samplesPerSymbol;
}

void sigProcLibDestroy(void) {               /* :61 */
  if (gCtx) btsdsp_destroy(gCtx);
  gCtx = NULL;
  gSps = 0;
}

float vectorNorm2(const signalVector &x) {   /* :146 -- caller-side scalar helper, same summation order */
  float e = 0.0F;
  for (const complex *p = x.begin(); p != x.end(); p++) e += p->norm2();
  return e;
}
float vectorPower(const signalVector &x) { return vectorNorm2(x) / x.size(); }

static signalVector *conv(bool corr, const signalVector *a, const signalVector *b, signalVector *c, ConvType span) {
  if (!a || !b) return NULL;
  int n = (corr ? btsdsp_correlate : btsdsp_convolve)(ctx(), cp(*a), a->size(), a->isRealOnly(), cp(*b), b->size(),
                                                       b->isRealOnly(), NULL, 0, span);
  if (n < 0) return NULL;                    /* unknown span type: the reference returns NULL (:302) */
  if (!c) c = new signalVector(n);
  else if ((int)c->size() != n) return NULL; /* :309 */
  check((corr ? btsdsp_correlate : btsdsp_convolve)(ctx(), cp(*a), a->size(), a->isRealOnly(), cp(*b), b->size(),
                                                    b->isRealOnly(), mp(*c), n, span), "convolve");
  return c;
}
signalVector *convolve(const signalVector *a, const signalVector *b, signalVector *c, ConvType s) { return conv(false, a, b, c, s); }  /* :267 */
signalVector *correlate(signalVector *a, signalVector *b, signalVector *c, ConvType s) { return conv(true, a, b, c, s); }            /* :474 */

signalVector *generateGSMPulse(int symbolLength, int samplesPerSymbol) {    /* :411 (callers pass (2, sps)) */
  if (symbolLength != 2 || samplesPerSymbol != gSps) return NULL;
  signalVector *p = new signalVector(2 * samplesPerSymbol + 1);
  check(btsdsp_get_table(ctx(), BTSDSP_T_PULSE, 0, (float *)p->begin(), 2 * p->size()), "get pulse");
  p->isRealOnly(true);
  return p;
}

bool vectorSlicer(signalVector *x) {          /* :507 */
  for (complex *p = x->begin(); p < x->end(); p++) {
    *p = (complex)(0.5 * (p->real() + 1.0F));
    if (p->real() > 1.0) *p = 1.0;
    if (p->real() < 0.0) *p = 0.0;
  }
  return true;
}

signalVector *modulateBurst(const BitVector &wBurst, const signalVector &gsmPulse, int guard, int sps) {   /* :521 */
  if (sps != gSps || (int)gsmPulse.size() != 2 * sps + 1) return NULL;   /* only the library's own GSM pulse */
  signalVector *out = new signalVector(sps * (wBurst.size() + guard));
  check(btsdsp_modulate_burst(ctx(), (const uint8_t *)wBurst.begin(), wBurst.size(), guard, mp(*out), out->size()),
        "modulateBurst");
  return out;
}

void delayVector(signalVector &wBurst, float delay) {   /* :573 */
  check(btsdsp_delay_vector(ctx(), mp(wBurst), wBurst.size(), delay), "delayVector");
}

bool addVector(signalVector &x, signalVector &y) {      /* :746 */
  complex *xp = x.begin(), *yp = y.begin();
  while (xp < x.end() && yp < y.end()) { *xp = *xp + *yp; xp++; yp++; }
  return true;
}

complex interpolatePoint(const signalVector &inSig, float ix) {   /* :639 */
  btsdsp_cf32 r;
  check(btsdsp_interpolate_point(ctx(), cp(inSig), inSig.size(), ix, &r), "interpolatePoint");
  return complex(r.re, r.im);
}

complex peakDetect(const signalVector &rxBurst, float *peakIndex, float *avgPwr) {   /* :663 */
  btsdsp_cf32 r;
  check(btsdsp_peak_detect(ctx(), cp(rxBurst), rxBurst.size(), &r, peakIndex, avgPwr), "peakDetect");
  return complex(r.re, r.im);
}

void scaleVector(signalVector &x, complex scale) {      /* :713 */
  btsdsp_cf32 s = {scale.real(), scale.imag()};
  check(btsdsp_scale_vector(ctx(), mp(x), x.size(), x.isRealOnly(), s), "scaleVector");
}

void offsetVector(signalVector &x, complex offset) {    /* :760 */
  for (complex *p = x.begin(); p < x.end(); p++) *p = x.isRealOnly() ? complex(p->real()) + offset : *p + offset;
}
void conjugateVector(signalVector &x) {                 /* :733 */
  if (x.isRealOnly()) return;
  for (complex *p = x.begin(); p < x.end(); p++) *p = p->conj();
}

/* :779 / :830 -- the correlation sequences are part of the context built by sigProcLibSetup */
bool generateMidamble(signalVector &, int sps, int TSC) { return sps == gSps && TSC >= 0 && TSC <= 7; }
bool generateRACHSequence(signalVector &, int sps) { return sps == gSps; }

bool energyDetect(signalVector &rxBurst, unsigned windowLength, float detectThreshold, float *avgPwr) {   /* :916 */
  int det = 0;
  check(btsdsp_energy_detect(ctx(), cp(rxBurst), rxBurst.size(), windowLength, detectThreshold, avgPwr, &det), "energyDetect");
  return det != 0;
}

bool detectRACHBurst(signalVector &rxBurst, float thr, int sps, complex *amplitude, float *TOA) {   /* :860 */
  (void)sps;
  btsdsp_cf32 a;
  int det = 0;
  check(btsdsp_detect_rach_burst(ctx(), cp(rxBurst), rxBurst.size(), thr, &a, TOA, &det), "detectRACHBurst");
  *amplitude = complex(a.re, a.im);
  return det != 0;
}

bool analyzeTrafficBurst(signalVector &rxBurst, unsigned TSC, float thr, int sps, complex *amplitude, float *TOA,
                         bool requestChannel, signalVector **channelResponse, float *channelResponseOffset) {   /* :935 */
  btsdsp_cf32 a, chan[6 * 4];
  float off = 0.0F;
  int det = 0;
  check(btsdsp_analyze_traffic_burst(ctx(), cp(rxBurst), rxBurst.size(), TSC, thr, &a, TOA, requestChannel, chan, &off, &det),
        "analyzeTrafficBurst");
  *amplitude = complex(a.re, a.im);
  if (requestChannel && det) {
    if (channelResponse) {
      *channelResponse = new signalVector(6 * sps);
      memcpy((*channelResponse)->begin(), chan, 6 * sps * sizeof(complex));
    }
    if (channelResponseOffset) *channelResponseOffset = off;
  }
  return det != 0;
}

/* The second transceiver variant's signature (Transceiver52M/sigProcLib.h:290-305: an extra maxTOA); C++ overload, so a
 * Transceiver52M caller links against it unchanged. */
bool analyzeTrafficBurst(signalVector &rxBurst, unsigned TSC, float thr, int sps, complex *amplitude, float *TOA,
                         unsigned maxTOA, bool requestChannel, signalVector **channelResponse,
                         float *channelResponseOffset) {                                  /* Transceiver52M/sigProcLib.cpp:966 */
  btsdsp_cf32 a, chan[6 * 4];
  float off = 0.0F;
  int det = 0;
  check(btsdsp_analyze_traffic_burst_52m(ctx(), cp(rxBurst), rxBurst.size(), TSC, thr, maxTOA, requestChannel, &det, &a, TOA,
                                         chan, &off),
        "analyzeTrafficBurst (52M)");
  *amplitude = complex(a.re, a.im);
  if (requestChannel && det) {
    if (channelResponse) {
      *channelResponse = new signalVector(6 * sps);
      memcpy((*channelResponse)->begin(), chan, 6 * sps * sizeof(complex));
    }
    if (channelResponseOffset) *channelResponseOffset = off;
  }
  return det != 0;
}

signalVector *decimateVector(signalVector &wVector, int decimationFactor) {   /* :1039 */
  if (decimationFactor <= 1) return NULL;
  signalVector *d = new signalVector(wVector.size() / decimationFactor);
  d->isRealOnly(wVector.isRealOnly());
  for (size_t k = 0; k < d->size(); k++) (*d)[k] = wVector[k * decimationFactor];
  return d;
}

SoftVector *demodulateBurst(const signalVector &rxBurst, const signalVector &, int sps, complex channel, float TOA) {   /* :1056 */
  SoftVector *s = new SoftVector(sps > 1 ? rxBurst.size() / sps : rxBurst.size());
  btsdsp_cf32 c = {channel.real(), channel.imag()};
  check(btsdsp_demodulate_burst(ctx(), cp(rxBurst), rxBurst.size(), c, TOA, s->begin()), "demodulateBurst");
  return s;
}

signalVector *createLPF(float, int filterLen, float) {   /* :1102 -- the cutoff argument is ignored there too */
  const int id = (filterLen == 651) ? BTSDSP_T_LPF_TX : BTSDSP_T_LPF_RX;
  const int n = (filterLen == 651) ? 651 : 961;
  float taps[961];
  check(btsdsp_get_table(ctx(), id, 0, taps, 961), "get LPF");
  signalVector *v = new signalVector(n);
  for (int k = 0; k < n; k++) (*v)[k] = complex(taps[k], 0.0F);
  v->isRealOnly(true);
  return v;
}

signalVector *polyphaseResampleVector(signalVector &wVector, int P, int Q, signalVector *LPF) {   /* :1157 */
  /* the two filters radioInterface.cpp creates are the two the library holds; pick by length */
  const int lpf = (LPF && LPF->size() == 651) ? 1 : 0;
  int n = btsdsp_polyphase_resample(ctx(), cp(wVector), wVector.size(), P, Q, lpf, NULL, 0);
  check(n, "polyphaseResampleVector");
  signalVector *out = new signalVector(n);
  check(btsdsp_polyphase_resample(ctx(), cp(wVector), wVector.size(), P, Q, lpf, mp(*out), n), "polyphaseResampleVector");
  return out;
}

bool designDFE(signalVector &channelResponse, float SNRestimate, int Nf, signalVector **w, signalVector **b) {   /* :1246 */
  const int nu = channelResponse.size() - 1;
  *w = new signalVector(Nf);
  *b = new signalVector(nu);
  int rc = btsdsp_design_dfe(ctx(), cp(channelResponse), channelResponse.size(), SNRestimate, Nf, mp(**w), mp(**b));
  if (rc < 0) { delete *w; delete *b; *w = *b = NULL; return false; }
  return true;
}

SoftVector *equalizeBurst(signalVector &rxBurst, float TOA, int, signalVector &w, signalVector &b) {   /* :1343 */
  SoftVector *s = new SoftVector(rxBurst.size());
  int rc = btsdsp_equalize_burst(ctx(), mp(rxBurst), rxBurst.size(), TOA, cp(w), w.size(), cp(b), b.size(), s->begin());
  if (rc < 0) { delete s; return NULL; }
  return s;
}
