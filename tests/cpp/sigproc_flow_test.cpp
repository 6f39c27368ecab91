// tests/cpp/sigproc_flow_test.cpp -- the reference's own test flow (Transceiver/sigProcLibTest.cpp:29-181: build a
// RACH burst and a TSC-0 normal burst, modulate, resample 96/65 then 65/96, delay, 2-tap channel, noise, analyze,
// slicer-demodulate, design the DFE, equalize), written against the sigProcLib.h surface and linked with the
// btsdsp shim instead of the reference's sigProcLib.cpp.  The reference program only prints; this one writes
// every stage to a binary file so tests/test_gpu_shim.py can replay the same inputs through the oracle.
#include <stdio.h>
#include <stdlib.h>
#include "sigProcLib.h"

static unsigned lcg = 12345u;
static float unif() { lcg = lcg * 1664525u + 1013904223u; return ((lcg >> 8) + 0.5F) / 16777216.0F; }

static void put(FILE *f, const char *tag, const void *p, size_t bytes) {
  unsigned n = (unsigned)bytes;
  char t[8] = {0};
  for (int i = 0; i < 7 && tag[i]; i++) t[i] = tag[i];
  fwrite(t, 1, 8, f); fwrite(&n, 4, 1, f); fwrite(p, 1, bytes, f);
}

int main(int argc, char **argv) {
  const char *path = argc > 1 ? argv[1] : "flow.bin";
  FILE *f = fopen(path, "wb");
  if (!f) return 2;
  const int sps = 1;
  sigProcLibSetup(sps);
  signalVector *gsmPulse = generateGSMPulse(2, sps);
  generateRACHSequence(*gsmPulse, sps);
  generateMidamble(*gsmPulse, sps, 0);

  BitVector normalBits("0000101010100111110010101010010110101110011000111001101010000"
                       "00100101110000100010010111"
                       "0000101010100111110010101010010110101110011000111001101010000");
  BitVector rachBits("0011101001001011011111111001100110101010001111000110111101111110000111001001010110011000");
  signalVector *modBurst = modulateBurst(normalBits, *gsmPulse, 8, sps);
  signalVector *rachBurst = modulateBurst(rachBits, *gsmPulse, 156 - 88, sps);
  put(f, "tx", modBurst->begin(), modBurst->bytes());

  // resample up to the radio rate and back (sigProcLibTest.cpp:103-111)
  signalVector *lpfTx = createLPF(0.0F, 651, 96), *lpfRx = createLPF(0.0F, 961, 65);
  signalVector *up = polyphaseResampleVector(*modBurst, 96, 65, lpfTx);
  signalVector *down = polyphaseResampleVector(*up, 65, 96, lpfRx);
  put(f, "up", up->begin(), up->bytes());
  put(f, "down", down->begin(), down->bytes());

  signalVector rx(*modBurst);
  delayVector(rx, 6.932F);
  signalVector channel(4);
  channel[0] = complex(9000.0F, 0.0F); channel[1] = complex(3600.0F, 0.0F);
  signalVector *faded = convolve(&rx, &channel, NULL, NO_DELAY);
  for (size_t k = 0; k < faded->size(); k++) (*faded)[k] = (*faded)[k] + complex(40.0F * (unif() - 0.5F), 40.0F * (unif() - 0.5F));
  put(f, "rx", faded->begin(), faded->bytes());

  complex amp; float TOA = 0, off = 0;
  signalVector *chanResp = NULL;
  bool ok = analyzeTrafficBurst(*faded, 0, 8.0F, sps, &amp, &TOA, true, &chanResp, &off);
  float meta[5] = {(float)ok, amp.real(), amp.imag(), TOA, off};
  put(f, "meta", meta, sizeof meta);
  if (!ok) { fprintf(stderr, "normal burst not detected\n"); return 1; }
  put(f, "chan", chanResp->begin(), chanResp->bytes());

  SoftVector *slicer = demodulateBurst(*faded, *gsmPulse, sps, amp, TOA);
  put(f, "slicer", slicer->begin(), slicer->bytes());

  // the caller's glue, Transceiver.cpp:340-347, :391-396
  float SNR = amp.norm2() / (250.0F * 250.0F + 1.0);
  scaleVector(*chanResp, complex(1.0, 0.0) / amp);
  signalVector *w = NULL, *b = NULL;
  designDFE(*chanResp, SNR, 7, &w, &b);
  put(f, "w", w->begin(), w->bytes());
  put(f, "b", b->begin(), b->bytes());
  signalVector scaled(*faded);
  scaleVector(scaled, complex(1.0, 0.0) / amp);
  SoftVector *soft = equalizeBurst(scaled, TOA - off, sps, *w, *b);
  put(f, "soft", soft->begin(), soft->bytes());
  put(f, "after", scaled.begin(), scaled.bytes());

  int errs = 0, errsSlicer = 0;
  for (size_t k = 0; k < normalBits.size(); k++) { errs += soft->bit(k) != normalBits.bit(k); errsSlicer += slicer->bit(k) != normalBits.bit(k); }

  signalVector rrx(*rachBurst);
  delayVector(rrx, 3.4F);
  scaleVector(rrx, complex(2000.0F, 500.0F));
  put(f, "rachrx", rrx.begin(), rrx.bytes());
  complex ramp; float rtoa = 0;
  bool rok = detectRACHBurst(rrx, 5.0F, sps, &ramp, &rtoa);
  float rmeta[4] = {(float)rok, ramp.real(), ramp.imag(), rtoa};
  put(f, "rachm", rmeta, sizeof rmeta);
  float avg = 0;
  bool e = energyDetect(*faded, 20 * sps, 250.0F, &avg);
  float emeta[2] = {(float)e, avg};
  put(f, "energy", emeta, sizeof emeta);
  fclose(f);
  printf("normal: detected=%d amp=(%g,%g) TOA=%g off=%g DFE bit errors=%d slicer bit errors=%d | RACH: detected=%d TOA=%g\n",
         ok, amp.real(), amp.imag(), TOA, off, errs, errsSlicer, rok, rtoa);
  delete modBurst; delete rachBurst; delete lpfTx; delete lpfRx; delete up; delete down; delete faded; delete chanResp;
  delete slicer; delete w; delete b; delete soft; delete gsmPulse;
  sigProcLibDestroy();
  return (errs == 0 && rok) ? 0 : 1;
}
