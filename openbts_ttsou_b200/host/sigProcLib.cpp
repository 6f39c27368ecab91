// host/sigProcLib.cpp -- the reference's sigProcLib.h functions as thin calls into libbtsdsp.so.
//
// A maintainer replaces Transceiver/sigProcLib.cpp by this file (+ -lbtsdsp) and keeps Transceiver.cpp and
// radioInterface.cpp as they are: every function below has the reference's signature, ownership rules and
// return conventions (reference Transceiver/sigProcLib.cpp, lines cited per function).  All vector arithmetic runs
// in the CUDA kernels; this file only moves vectors across the C ABI (the scalar helpers dB / dBinv / sinc and the
// two never-called utilities gaussianNoise / resampleVector are evaluated here from the library's own tables).
//
// Threading: the reference's functions are re-entrant and are called concurrently from the RX-FIFO thread, the
// TX-queue thread and the two radio threads (Transceiver.cpp:412-426, radioInterface.cpp:305-330).  The library's
// single-vector calls share one context's staging buffers and stream, so every call below takes gLock: callers may
// come from any number of threads, calls are serialised inside (test: tests/test_gpu_shim.py::test_threads).
//
// Arguments the reference takes but this library holds as state (the GSM pulse, the two resampler filters) are
// CHECKED, not ignored: a pulse that is not generateGSMPulse(2, sps) makes modulateBurst / generateMidamble /
// generateRACHSequence return NULL / false; a filter that is not one of the two createLPF tables is uploaded and
// run through the generic resampler kernel.
#include "sigProcLib.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "btsdsp.h"

static btsdsp_ctx *gCtx = NULL;
static int gDevice = -1;
static int gSps = 0;
static std::recursive_mutex gLock;
static float gCos[1026], gSin[1026];          /* the library's trig tables (cosTable / sinTable, :207-212) */
static float gPulseTaps[2 * 9];               /* generateGSMPulse(2, sps) as the library holds it */
static float gLpfRx[961], gLpfTx[651];        /* createLPF(., 961, 65) / createLPF(., 651, 96) as the library holds them */
typedef std::lock_guard<std::recursive_mutex> Guard;

static btsdsp_ctx *ctx() {
  if (!gCtx) {
    fprintf(stderr, "sigProcLib(btsdsp): sigProcLibSetup() has not been called\n");
    abort();
  }
  return gCtx;
}
/* BTSDSP_EINVAL / EUNSUPPORTED = an input the kernels do not take (the reference would read out of bounds or is
 * undefined there): reported to the caller as the reference reports failure (false / NULL).  A CUDA failure has no
 * error channel in the reference's signatures and must not pass silently: abort. */
static bool check(int rc, const char *what) {
  if (rc >= 0) return true;
  fprintf(stderr, "sigProcLib(btsdsp): %s failed (%d): %s\n", what, rc, btsdsp_last_error(gCtx));
  if (rc == BTSDSP_EINVAL || rc == BTSDSP_EUNSUPPORTED) return false;
  abort();
}
static const btsdsp_cf32 *cp(const signalVector &v) { return (const btsdsp_cf32 *)v.begin(); }
static btsdsp_cf32 *mp(signalVector &v) { return (btsdsp_cf32 *)v.begin(); }

/** which GPU sigProcLibSetup binds to: this setter, else $BTSDSP_DEVICE, else 0 (not part of the reference surface) */
extern "C" void sigProcLibSetDevice(int device) { gDevice = device; }

/* sigProcLib.cpp:227 -- also builds what generateMidamble / generateRACHSequence / createLPF produce */
void sigProcLibSetup(int samplesPerSymbol) {
  Guard g(gLock);
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
  check(btsdsp_get_table(gCtx, BTSDSP_T_COS, 0, gCos, 1026), "get cos table");
  check(btsdsp_get_table(gCtx, BTSDSP_T_SIN, 0, gSin, 1026), "get sin table");
  check(btsdsp_get_table(gCtx, BTSDSP_T_PULSE, 0, gPulseTaps, 18), "get pulse");
  check(btsdsp_get_table(gCtx, BTSDSP_T_LPF_RX, 0, gLpfRx, 961), "get RX LPF");
  check(btsdsp_get_table(gCtx, BTSDSP_T_LPF_TX, 0, gLpfTx, 651), "get TX LPF");
}

void sigProcLibDestroy(void) {               /* :61 */
  Guard g(gLock);
  if (gCtx) btsdsp_destroy(gCtx);
  gCtx = NULL;
  gSps = 0;
}

/* is `p` the pulse the library was built with (generateGSMPulse(2, sps))?  Bit comparison of the real parts. */
static bool is_library_pulse(const signalVector &p, int sps) {
  if (sps != gSps || (int)p.size() != 2 * sps + 1) return false;
  for (int k = 0; k < 2 * sps + 1; k++)
    if (memcmp(&p.begin()[k].r, &gPulseTaps[2 * k], sizeof(float)) != 0) return false;
  return true;
}

/* dB :88-115 / dBinv :117-143: scalar float, no tables.  Both walk (arg, dB) down from (1, 0) in steps of
 * (/16, -12 dB), (/8, -9), (/4, -6), (/2, -3), each step size until the stop quantity passes x, then interpolate. */
static void dbDescend(float x, bool stopOnArg, float *argOut, float *dbOut) {
  float arg = 1.0F, level = 0.0F, div = 16.0F, drop = 12.0F;
  for (; div > 1.0F; div *= 0.5F, drop -= 3.0F) {
    float keepArg, keepLevel;
    do {
      keepArg = arg; keepLevel = level;
      arg /= div; level -= drop;
    } while (stopOnArg ? (arg > x) : (level > x));
    arg = keepArg; level = keepLevel;
  }
  *argOut = arg; *dbOut = level;
}
float dB(float x) {
  if (x >= 1.0F) return 0.0F;
  if (x <= 0.0F) return -200.0F;
  float arg, level;
  dbDescend(x, true, &arg, &level);
  return ((arg - x) * (level - 3.0F) + (x - arg * 0.5F) * level) / (arg - arg * 0.5F);
}
float dBinv(float x) {
  if (x >= 0.0F) return 1.0F;
  if (x <= -200.0F) return 0.0F;
  float arg, level;
  dbDescend(x, false, &arg, &level);
  return ((level - x) * (arg * 0.5F) + (x - (level - 3.0F)) * (arg)) / 3.0F;
}

/* cosLookup / sinLookup :163-188 over the library's tables, and sinc :567-571 */
static const float k2PiF = (float)(2.0 * M_PI);
static const float k1_2PiF = 1 / k2PiF;                       /* M_1_2PI_F, :45 */
static float trigLookup(const float *table, float x) {
  float arg = x * k1_2PiF;
  while (arg > 1.0F) arg -= 1.0F;
  while (arg < 0.0F) arg += 1.0F;
  const float argT = arg * 1024.0F;
  const int argI = (int)argT;
  const float delta = argT - argI;
  const float iDelta = 1.0F - delta;
  return iDelta * table[argI] + delta * table[argI + 1];
}
static complex expjLookupHost(float x) { return complex(trigLookup(gCos, x), trigLookup(gSin, x)); }   /* :190-204 */
float cosLookup(const float x) { return trigLookup(gCos, x); }
float sinLookup(const float x) { return trigLookup(gSin, x); }
complex expjLookup(float x) { return expjLookupHost(x); }
void initTrigTables() {}                      /* :207 / :214: the tables are built by sigProcLibSetup -> btsdsp_create */
void initGMSKRotationTables(int) {}
float sinc(float x) {
  if ((x >= 0.01F) || (x <= -0.01F)) return (trigLookup(gSin, x) / x);
  return 1.0F;
}

float vectorNorm2(const signalVector &x) {   /* :146 -- summed on the device in the reference's order */
  Guard g(gLock);
  float e = 0.0F;
  if (x.size() == 0) return e;
  btsdsp_cf32 z = {0.0F, 0.0F};
  check(btsdsp_vector_op(ctx(), BTSDSP_VOP_NORM2, (btsdsp_cf32 *)x.begin(), x.size(), 0, NULL, 0, z, &e), "vectorNorm2");
  return e;
}
float vectorPower(const signalVector &x) { return vectorNorm2(x) / x.size(); }   /* :157 */

static signalVector *conv(bool corr, const signalVector *a, const signalVector *b, signalVector *c, ConvType span) {
  if (!a || !b) return NULL;
  Guard g(gLock);
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
  Guard g(gLock);
  if (symbolLength != 2 || samplesPerSymbol != gSps) return NULL;
  signalVector *p = new signalVector(2 * samplesPerSymbol + 1);
  memcpy(p->begin(), gPulseTaps, p->size() * sizeof(complex));
  p->isRealOnly(true);
  return p;
}

/* frequencyShift :432-471: y[n] = x[n] * expjLookup(startPhase + n*freq), the phase accumulated in float */
signalVector *frequencyShift(signalVector *y, signalVector *x, float freq, float startPhase, float *finalPhase) {
  if (!x) return NULL;
  Guard g(gLock);
  ctx();
  if (!y) {
    y = new signalVector(x->size());
    y->isRealOnly(x->isRealOnly());
  }
  if (y->size() < x->size()) return NULL;
  float phase = startPhase;
  complex *out = y->begin();
  for (const complex *in = x->begin(); in < x->end(); in++, out++) {
    const complex e = expjLookupHost(phase);
    *out = x->isRealOnly() ? e * in->real() : (*in) * e;
    phase += freq;
  }
  if (finalPhase) *finalPhase = phase;
  return y;
}

static bool vop(int op, signalVector &x, const signalVector *y, complex s, const char *what) {
  if (x.size() == 0) return true;
  Guard g(gLock);
  btsdsp_cf32 sc = {s.real(), s.imag()};
  return check(btsdsp_vector_op(ctx(), op, mp(x), x.size(), x.isRealOnly(), y ? cp(*y) : NULL, y ? (int)y->size() : 0, sc, NULL),
               what);
}

bool vectorSlicer(signalVector *x) {          /* :507 */
  return vop(BTSDSP_VOP_SLICE, *x, NULL, complex(0.0F, 0.0F), "vectorSlicer");
}

signalVector *modulateBurst(const BitVector &wBurst, const signalVector &gsmPulse, int guard, int sps) {   /* :521 */
  Guard g(gLock);
  if (!is_library_pulse(gsmPulse, sps)) {
    fprintf(stderr, "sigProcLib(btsdsp): modulateBurst: the pulse is not generateGSMPulse(2, %d)\n", gSps);
    return NULL;
  }
  const int n = sps * (wBurst.size() + guard);
  signalVector *out = new signalVector(n);
  if (!check(btsdsp_modulate_burst(ctx(), (const uint8_t *)wBurst.begin(), wBurst.size(), guard, mp(*out), n), "modulateBurst")) {
    delete out;
    return NULL;
  }
  return out;
}

void delayVector(signalVector &wBurst, float delay) {   /* :573 */
  Guard g(gLock);
  check(btsdsp_delay_vector(ctx(), mp(wBurst), wBurst.size(), delay), "delayVector");
}

/* gaussianNoise :618-637 -- Box-Muller over libc rand(), a test utility of the reference; host side by nature */
signalVector *gaussianNoise(int length, float variance, complex mean) {
  signalVector *noise = new signalVector(length);
  const float stddev = sqrtf(variance);
  for (complex *p = noise->begin(); p < noise->end(); p++) {
    float u1 = (float)rand() / (float)RAND_MAX;
    while (u1 == 0.0) u1 = (float)rand() / (float)RAND_MAX;
    const float u2 = (float)rand() / (float)RAND_MAX;
    const float arg = 2.0 * M_PI * u2;
    *p = mean + stddev * complex(cos(arg), sin(arg)) * sqrtf(-2.0 * log(u1));
  }
  return noise;
}

bool addVector(signalVector &x, signalVector &y) {      /* :746 */
  if (y.size() == 0) return true;
  return vop(BTSDSP_VOP_ADD, x, &y, complex(0.0F, 0.0F), "addVector");
}

complex interpolatePoint(const signalVector &inSig, float ix) {   /* :639 */
  Guard g(gLock);
  btsdsp_cf32 r = {0.0F, 0.0F};
  check(btsdsp_interpolate_point(ctx(), cp(inSig), inSig.size(), ix, &r), "interpolatePoint");
  return complex(r.re, r.im);
}

complex peakDetect(const signalVector &rxBurst, float *peakIndex, float *avgPwr) {   /* :663 */
  Guard g(gLock);
  btsdsp_cf32 r = {0.0F, 0.0F};
  check(btsdsp_peak_detect(ctx(), cp(rxBurst), rxBurst.size(), &r, peakIndex, avgPwr), "peakDetect");
  return complex(r.re, r.im);
}

void scaleVector(signalVector &x, complex scale) {      /* :713 */
  Guard g(gLock);
  btsdsp_cf32 s = {scale.real(), scale.imag()};
  check(btsdsp_scale_vector(ctx(), mp(x), x.size(), x.isRealOnly(), s), "scaleVector");
}

void GMSKRotate(signalVector &x) { vop(BTSDSP_VOP_ROTATE, x, NULL, complex(0.0F, 0.0F), "GMSKRotate"); }                  /* :232 */
void GMSKReverseRotate(signalVector &x) { vop(BTSDSP_VOP_REVROTATE, x, NULL, complex(0.0F, 0.0F), "GMSKReverseRotate"); }   /* :249 */
void offsetVector(signalVector &x, complex offset) { vop(BTSDSP_VOP_OFFSET, x, NULL, offset, "offsetVector"); }   /* :760 */
void conjugateVector(signalVector &x) { vop(BTSDSP_VOP_CONJ, x, NULL, complex(0.0F, 0.0F), "conjugateVector"); }  /* :733 */

/* :779 / :830 -- the correlation sequences are part of the context built by sigProcLibSetup, from the library's own
 * pulse; a different pulse cannot be honoured and is refused */
bool generateMidamble(signalVector &gsmPulse, int sps, int TSC) {
  Guard g(gLock);
  return TSC >= 0 && TSC <= 7 && is_library_pulse(gsmPulse, sps);
}
bool generateRACHSequence(signalVector &gsmPulse, int sps) {
  Guard g(gLock);
  return is_library_pulse(gsmPulse, sps);
}

bool energyDetect(signalVector &rxBurst, unsigned windowLength, float detectThreshold, float *avgPwr) {   /* :916 */
  Guard g(gLock);
  int det = 0;
  if (!check(btsdsp_energy_detect(ctx(), cp(rxBurst), rxBurst.size(), windowLength, detectThreshold, avgPwr, &det), "energyDetect"))
    return false;
  return det != 0;
}

bool detectRACHBurst(signalVector &rxBurst, float thr, int sps, complex *amplitude, float *TOA) {   /* :860 */
  (void)sps;
  Guard g(gLock);
  btsdsp_cf32 a = {0.0F, 0.0F};
  int det = 0;
  if (!check(btsdsp_detect_rach_burst(ctx(), cp(rxBurst), rxBurst.size(), thr, &a, TOA, &det), "detectRACHBurst")) {
    *amplitude = 0.0F;
    return false;
  }
  *amplitude = complex(a.re, a.im);
  return det != 0;
}

bool analyzeTrafficBurst(signalVector &rxBurst, unsigned TSC, float thr, int sps, complex *amplitude, float *TOA,
                         bool requestChannel, signalVector **channelResponse, float *channelResponseOffset) {   /* :935 */
  Guard g(gLock);
  btsdsp_cf32 a = {0.0F, 0.0F}, chan[6 * 4];
  float off = 0.0F;
  int det = 0;
  /* a burst shorter than the correlation window (92*sps) or a TSC > 7 reads out of bounds in the reference; here: not detected */
  if (!check(btsdsp_analyze_traffic_burst(ctx(), cp(rxBurst), rxBurst.size(), TSC, thr, &a, TOA, requestChannel, chan, &off, &det),
             "analyzeTrafficBurst")) {
    *amplitude = 0.0F;
    return false;
  }
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
  Guard g(gLock);
  btsdsp_cf32 a = {0.0F, 0.0F}, chan[6 * 4];
  float off = 0.0F;
  int det = 0;
  if (!check(btsdsp_analyze_traffic_burst_52m(ctx(), cp(rxBurst), rxBurst.size(), TSC, thr, maxTOA, requestChannel, &det, &a, TOA,
                                              chan, &off),
             "analyzeTrafficBurst (52M)")) {
    *amplitude = 0.0F;
    return false;
  }
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

signalVector *decimateVector(signalVector &wVector, int decimationFactor) {   /* :1039 -- a strided copy, no arithmetic */
  if (decimationFactor <= 1) return NULL;
  signalVector *d = new signalVector(wVector.size() / decimationFactor);
  d->isRealOnly(wVector.isRealOnly());
  for (size_t k = 0; k < d->size(); k++) (*d)[k] = wVector[k * decimationFactor];
  return d;
}

SoftVector *demodulateBurst(const signalVector &rxBurst, const signalVector &, int sps, complex channel, float TOA) {   /* :1056 */
  /* (the reference does not use its gsmPulse argument either) */
  Guard g(gLock);
  SoftVector *s = new SoftVector(sps > 1 ? rxBurst.size() / sps : rxBurst.size());
  btsdsp_cf32 c = {channel.real(), channel.imag()};
  if (!check(btsdsp_demodulate_burst(ctx(), cp(rxBurst), rxBurst.size(), c, TOA, s->begin()), "demodulateBurst")) {
    delete s;
    return NULL;
  }
  return s;
}

signalVector *createLPF(float, int filterLen, float gainDC) {   /* :1102 -- the cutoff argument is ignored there too */
  Guard g(gLock);
  float taps[961];
  const int n = btsdsp_create_lpf(ctx(), filterLen, gainDC, taps, 961);
  if (!check(n, "createLPF")) return NULL;      /* filterLen > 961: the reference overruns its vector and its table */
  signalVector *v = new signalVector(n);
  for (int k = 0; k < n; k++) (*v)[k] = complex(taps[k], 0.0F);
  v->isRealOnly(true);
  return v;
}

/* is the caller's filter one of the two the library holds (createLPF(., 961, 65) / createLPF(., 651, 96))? */
static int held_lpf(const signalVector &f) {
  if (!f.isRealOnly()) return -1;
  const float *held = NULL;
  int id = -1;
  if (f.size() == 961) { held = gLpfRx; id = 0; }
  else if (f.size() == 651) { held = gLpfTx; id = 1; }
  else return -1;
  for (size_t k = 0; k < f.size(); k++)
    if (memcmp(&f.begin()[k].r, &held[k], sizeof(float)) != 0) return -1;
  return id;
}

signalVector *polyphaseResampleVector(signalVector &wVector, int P, int Q, signalVector *LPF) {   /* :1157 */
  Guard g(gLock);
  if (!LPF) {
    /* the reference builds createLPF(cutoff/3, 100*POLYPHASESPAN+1 = 1001, Q) here (:1165-1169), which writes 1001 taps
     * into a 961-entry vector from a 961-entry table: undefined behaviour, no caller does it.  Refused loudly. */
    fprintf(stderr, "sigProcLib(btsdsp): polyphaseResampleVector(LPF == NULL) is undefined in the reference; pass createLPF()'s filter\n");
    return NULL;
  }
  const int lpf = held_lpf(*LPF);
  int n;
  if (lpf >= 0) n = btsdsp_polyphase_resample(ctx(), cp(wVector), wVector.size(), P, Q, lpf, NULL, 0);
  else n = btsdsp_polyphase_resample_taps(ctx(), cp(wVector), wVector.size(), P, Q, cp(*LPF), LPF->size(), LPF->isRealOnly(), NULL, 0);
  if (!check(n, "polyphaseResampleVector")) return NULL;
  signalVector *out = new signalVector(n);
  out->isRealOnly(wVector.isRealOnly());
  if (lpf >= 0) n = btsdsp_polyphase_resample(ctx(), cp(wVector), wVector.size(), P, Q, lpf, mp(*out), n);
  else n = btsdsp_polyphase_resample_taps(ctx(), cp(wVector), wVector.size(), P, Q, cp(*LPF), LPF->size(), LPF->isRealOnly(), mp(*out), n);
  if (!check(n, "polyphaseResampleVector")) { delete out; return NULL; }
  return out;
}

/* resampleVector :1213-1241, never called in the reference.  Kept with its behaviour as written: the output iterator
 * is never advanced, so every interpolated point lands in element 0 and the rest of the vector stays zero. */
signalVector *resampleVector(signalVector &wVector, float expFactor, complex endPoint) {
  if (expFactor < 1.0) return NULL;
  signalVector *ret = new signalVector((int)ceil(wVector.size() * expFactor));
  float t = 0.0;
  complex *slot = ret->begin();
  while (slot < ret->end()) {
    const unsigned lo = (unsigned int)floor(t), hi = lo + 1;
    if (lo > wVector.size() - 1 || hi > wVector.size()) break;
    const complex p0 = wVector[lo], p1 = (hi == wVector.size()) ? endPoint : wVector[hi];
    const complex wa = (hi - t), wb = (t - lo);
    *slot = (wa * p0 + wb * p1);
    t += 1.0 / expFactor;
  }
  return ret;
}

bool designDFE(signalVector &channelResponse, float SNRestimate, int Nf, signalVector **w, signalVector **b) {   /* :1246 */
  Guard g(gLock);
  const int nu = channelResponse.size() - 1;
  *w = new signalVector(Nf);
  *b = new signalVector(nu);
  int rc = btsdsp_design_dfe(ctx(), cp(channelResponse), channelResponse.size(), SNRestimate, Nf, mp(**w), mp(**b));
  if (!check(rc, "designDFE")) { delete *w; delete *b; *w = *b = NULL; return false; }
  return true;
}

SoftVector *equalizeBurst(signalVector &rxBurst, float TOA, int, signalVector &w, signalVector &b) {   /* :1343 */
  Guard g(gLock);
  SoftVector *s = new SoftVector(rxBurst.size());
  int rc = btsdsp_equalize_burst(ctx(), mp(rxBurst), rxBurst.size(), TOA, cp(w), w.size(), cp(b), b.size(), s->begin());
  if (!check(rc, "equalizeBurst")) { delete s; return NULL; }
  return s;
}
