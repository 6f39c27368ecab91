// host/sigProcLib.h -- the reference's burst-DSP function surface (reference Transceiver/sigProcLib.h:101-384),
// implemented by host/sigProcLib.cpp on top of libbtsdsp.so (include/btsdsp.h): same names, argument meaning,
// ownership (returned vectors are new-allocated, caller deletes; delayVector / equalizeBurst mutate their input)
// and error behaviour (bool / NULL returns, amplitude = 0 on a bogus peak).  Callers written against the
// reference -- Transceiver.cpp, radioInterface.cpp -- compile against this header unchanged; see INTEGRATION.md.
// When the reference tree is on the include path, its own sigProcLib.h can be used instead: the definitions in
// host/sigProcLib.cpp match both.
#ifndef BTSDSP_HOST_SIGPROCLIB_H
#define BTSDSP_HOST_SIGPROCLIB_H
#include "BitVector.h"
#include "Complex.h"
#include "Vector.h"

enum Symmetry { NONE = 0, ABSSYM = 1 };
enum ConvType { FULL_SPAN = 0, OVERLAP_ONLY = 1, START_ONLY = 2, WITH_TAIL = 3, NO_DELAY = 4, UNDEFINED = 255 };

class signalVector : public Vector<complex> {
  Symmetry symmetry;
  bool realOnly;

 public:
  signalVector(int n = 0, Symmetry s = NONE) : Vector<complex>(n), symmetry(s), realOnly(false) {}
  signalVector(complex *data, size_t start, size_t span, Symmetry s = NONE)
      : Vector<complex>(NULL, data + start, data + start + span), symmetry(s), realOnly(false) {}
  signalVector(const signalVector &a, const signalVector &b) : Vector<complex>(a, b), symmetry(a.symmetry), realOnly(false) {}
  signalVector(const signalVector &v) : Vector<complex>(v), symmetry(v.symmetry), realOnly(false) {}   // drops realOnly, as the reference does
  Symmetry getSymmetry() const { return symmetry; }
  void setSymmetry(Symmetry s) { symmetry = s; }
  bool isRealOnly() const { return realOnly; }
  void isRealOnly(bool v) { realOnly = v; }
};

void sigProcLibSetup(int samplesPerSymbol);
void sigProcLibDestroy(void);

/* not declared by the reference's header but exported by its object; defined here too so every symbol resolves */
float cosLookup(const float x);
float sinLookup(const float x);
complex expjLookup(float x);
void GMSKRotate(signalVector &x);
void GMSKReverseRotate(signalVector &x);
float dB(float x);
float dBinv(float x);
float vectorNorm2(const signalVector &x);
float vectorPower(const signalVector &x);
signalVector *convolve(const signalVector *a, const signalVector *b, signalVector *c, ConvType spanType);
signalVector *correlate(signalVector *a, signalVector *b, signalVector *c, ConvType spanType);
signalVector *generateGSMPulse(int symbolLength, int samplesPerSymbol);
signalVector *frequencyShift(signalVector *y, signalVector *x, float freq = 0.0, float startPhase = 0.0,
                             float *finalPhase = NULL);
bool vectorSlicer(signalVector *x);
signalVector *modulateBurst(const BitVector &wBurst, const signalVector &gsmPulse, int guardPeriodLength,
                            int samplesPerSymbol);
float sinc(float x);
void delayVector(signalVector &wBurst, float delay);
signalVector *gaussianNoise(int length, float variance = 1.0, complex mean = complex(0.0));
bool addVector(signalVector &x, signalVector &y);
complex interpolatePoint(const signalVector &inSig, float ix);
complex peakDetect(const signalVector &rxBurst, float *peakIndex, float *avgPwr);
void scaleVector(signalVector &x, complex scale);
void offsetVector(signalVector &x, complex offset);
void conjugateVector(signalVector &x);
bool generateMidamble(signalVector &gsmPulse, int samplesPerSymbol, int TSC);
bool generateRACHSequence(signalVector &gsmPulse, int samplesPerSymbol);
bool energyDetect(signalVector &rxBurst, unsigned windowLength, float detectThreshold, float *avgPwr = NULL);
bool detectRACHBurst(signalVector &rxBurst, float detectThreshold, int samplesPerSymbol, complex *amplitude, float *TOA);
bool analyzeTrafficBurst(signalVector &rxBurst, unsigned TSC, float detectThreshold, int samplesPerSymbol,
                         complex *amplitude, float *TOA, bool requestChannel = false,
                         signalVector **channelResponse = NULL, float *channelResponseOffset = NULL);
/* Transceiver52M's form (extra maxTOA, windowed search) */
bool analyzeTrafficBurst(signalVector &rxBurst, unsigned TSC, float detectThreshold, int samplesPerSymbol,
                         complex *amplitude, float *TOA, unsigned maxTOA, bool requestChannel,
                         signalVector **channelResponse, float *channelResponseOffset);
signalVector *decimateVector(signalVector &wVector, int decimationFactor);
SoftVector *demodulateBurst(const signalVector &rxBurst, const signalVector &gsmPulse, int samplesPerSymbol,
                            complex channel, float TOA);
signalVector *createLPF(float cutoffFreq, int filterLen, float gainDC = 1.0);
signalVector *polyphaseResampleVector(signalVector &wVector, int P, int Q, signalVector *LPF);
signalVector *resampleVector(signalVector &wVector, float expFactor, complex endPoint);
bool designDFE(signalVector &channelResponse, float SNRestimate, int Nf, signalVector **feedForwardFilter,
               signalVector **feedbackFilter);
SoftVector *equalizeBurst(signalVector &rxBurst, float TOA, int samplesPerSymbol, signalVector &w, signalVector &b);
#endif
