/* ref52_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" face over the reference's SECOND transceiver variant,
 * Transceiver52M/sigProcLib.cpp, compiled in place like oracle/ref_shim.cpp (oracle/Makefile target
 * `ref52` -> oracle/_ref/libref52_oracle.so; it is a separate library because both variants define the same symbols).
 * Only what differs from the main variant is exposed: the windowed (CUSTOM-span) analyzeTrafficBurst with maxTOA
 * (Transceiver52M/sigProcLib.cpp:966-1077) and energyDetect's stride-4 window (:944-963). */
#include "sigProcLib.h"
#include "GSMCommon.h"
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

}  // extern "C"
