// tests/cpp/surface_test.cpp -- the parts of the sigProcLib.h surface the reference's own test flow does not reach:
// the element-wise helpers, the scalar utilities, createLPF with every argument, polyphaseResampleVector with a filter
// that is not one of the two tables, and concurrent calls from four threads.  The SAME source is built twice:
//   * against the reference's own sigProcLib.cpp (oracle/gen_surface_golden.py, only where /root/reference exists);
//     its dump is committed as tests/golden/surface_ref.bin;
//   * against the btsdsp shim on the GPU box (tests/test_gpu_shim.py), whose dump must equal the golden byte for byte.
// Inputs come from a private LCG (never rand(), except gaussianNoise's own use of it behind srand(7)).
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sigProcLib.h"

// defined by sigProcLib.cpp but not declared in the reference's header
void conjugateVector(signalVector &x);
void GMSKRotate(signalVector &x);
void GMSKReverseRotate(signalVector &x);
float cosLookup(const float x);
float sinLookup(const float x);

static unsigned lcg = 2024u;
static float unif() { lcg = lcg * 1664525u + 1013904223u; return ((lcg >> 8) + 0.5F) / 16777216.0F; }
static FILE *gOut = NULL;

static void put(const char *tag, const void *p, size_t bytes) {
  unsigned n = (unsigned)bytes;
  char t[8] = {0};
  for (int i = 0; i < 7 && tag[i]; i++) t[i] = tag[i];
  fwrite(t, 1, 8, gOut); fwrite(&n, 4, 1, gOut); fwrite(p, 1, bytes, gOut);
}
static void putv(const char *tag, const signalVector &v) { put(tag, v.begin(), v.size() * sizeof(complex)); }
static void randomise(signalVector &v, float scale) {
  for (size_t k = 0; k < v.size(); k++) v[k] = complex(scale * (unif() - 0.5F), scale * (unif() - 0.5F));
}

// ---- the four-thread hammer: each thread runs the per-burst chain on its own bursts, concurrently with the others;
//      the results are written after the join, in thread order, so the dump does not depend on scheduling
struct Job {
  int id;
  signalVector *pulse, *lpfRx, *lpfTx;
  float meta[16][4];
  complex firstUp[16], firstDown[16], firstMod[16];
};
static void *hammer(void *arg) {
  Job *j = (Job *)arg;
  unsigned s = 77u + 1000u * j->id;
  for (int it = 0; it < 16; it++) {
    BitVector bits(148);
    for (int k = 0; k < 148; k++) { s = s * 1664525u + 1013904223u; bits.begin()[k] = (s >> 16) & 1; }
    const char *tsc = "00100101110000100010010111";
    for (int k = 0; k < 26; k++) bits.begin()[61 + k] = tsc[k] - '0';
    signalVector *m = modulateBurst(bits, *j->pulse, 8, 1);
    j->firstMod[it] = (*m)[5 + it];
    signalVector *up = polyphaseResampleVector(*m, 96, 65, j->lpfTx);
    signalVector *down = polyphaseResampleVector(*up, 65, 96, j->lpfRx);
    j->firstUp[it] = (*up)[40 + it];
    j->firstDown[it] = (*down)[30 + it];
    signalVector rx(*m);
    scaleVector(rx, complex(800.0F + 100.0F * j->id, 50.0F * it));
    delayVector(rx, 0.37F * it + 0.11F * j->id);
    complex amp; float toa = 0;
    bool ok = analyzeTrafficBurst(rx, 0, 3.0F, 1, &amp, &toa);
    j->meta[it][0] = ok; j->meta[it][1] = amp.real(); j->meta[it][2] = amp.imag(); j->meta[it][3] = toa;
    delete m; delete up; delete down;
  }
  return NULL;
}

int main(int argc, char **argv) {
  gOut = fopen(argc > 1 ? argv[1] : "surface.bin", "wb");
  if (!gOut) return 2;
  const int sps = 1;
  sigProcLibSetup(sps);
  signalVector *pulse = generateGSMPulse(2, sps);
  generateRACHSequence(*pulse, sps);
  for (int t = 0; t < 8; t++) generateMidamble(*pulse, sps, t);

  // scalar utilities
  {
    float v[64];
    for (int k = 0; k < 16; k++) v[k] = dB(0.9F / (1 + 37 * k * k));
    for (int k = 0; k < 16; k++) v[16 + k] = dBinv(-0.7F * k * k - 0.3F);
    for (int k = 0; k < 32; k++) v[32 + k] = sinc(-9.0F + 0.61F * k);
    v[32] = sinc(0.005F);
    put("scalars", v, sizeof v);
  }
  // element-wise helpers, complex and real-only operands
  {
    signalVector x(157), y(120), r(64);
    randomise(x, 2000.0F); randomise(y, 700.0F); randomise(r, 3.0F);
    r.isRealOnly(true);
    float nrm[4] = {vectorNorm2(x), vectorPower(x), vectorNorm2(y), vectorPower(r)};
    put("norms", nrm, sizeof nrm);
    addVector(x, y);                 putv("add", x);
    addVector(y, x);                 putv("add2", y);
    offsetVector(x, complex(1.5F, -2.25F));   putv("offc", x);
    offsetVector(r, complex(0.125F, 7.0F));   putv("offr", r);
    conjugateVector(x);              putv("conj", x);
    conjugateVector(r);              putv("conjr", r);
    signalVector s(40);
    for (int k = 0; k < 40; k++) s[k] = complex(-2.0F + 0.1F * k + 0.01F * unif(), unif());
    vectorSlicer(&s);                putv("slice", s);
    signalVector *d = decimateVector(y, 4);   putv("decim", *d);   delete d;   /* 120 = 4*30: the reference overruns its output otherwise */
    signalVector sc(r);
    sc.isRealOnly(true);
    scaleVector(sc, complex(0.5F, -3.0F));    putv("scaler", sc);
    signalVector g1(157), g2(100);
    randomise(g1, 2.0F); randomise(g2, 2.0F);
    g2.isRealOnly(true);
    GMSKRotate(g1);                  putv("rot", g1);
    GMSKReverseRotate(g1);           putv("rotrev", g1);
    GMSKRotate(g2);                  putv("rotr", g2);
    float tl[16];
    for (int k = 0; k < 8; k++) { tl[k] = cosLookup(-20.0F + 5.3F * k); tl[8 + k] = sinLookup(-20.0F + 5.3F * k); }
    put("trig", tl, sizeof tl);
  }
  // frequencyShift, both operand kinds, with and without a destination, phases beyond one turn
  {
    signalVector x(100), xr(50);
    randomise(x, 10.0F); randomise(xr, 10.0F);
    xr.isRealOnly(true);
    float fin = 0;
    signalVector *a = frequencyShift(NULL, &x, 0.3F, -1.0F, &fin);   putv("fshift", *a);  put("fshiftp", &fin, 4);
    signalVector dst(60);
    signalVector *b = frequencyShift(&dst, &xr, -0.05F, 7.5F, &fin); putv("fshiftr", *b); put("fshiftq", &fin, 4);
    delete a;
  }
  // gaussianNoise (libc rand() behind a fixed seed) and resampleVector as the reference wrote it
  {
    srand(7);
    signalVector *n = gaussianNoise(32, 0.25F, complex(1.0F, -1.0F));   putv("noise", *n);   delete n;
    signalVector x(20);
    randomise(x, 5.0F);
    signalVector *r = resampleVector(x, 1.7F, complex(3.0F, 4.0F));     putv("resamp", *r);  delete r;
    unsigned char isnull = resampleVector(x, 0.5F, complex(0.0F)) == NULL;
    put("resampn", &isnull, 1);
  }
  // createLPF: both tables, other gains, a shorter length
  signalVector *lpfTx = createLPF(0.0F, 651, 96), *lpfRx = createLPF(0.0F, 961, 65);
  {
    putv("lpf651", *lpfTx); putv("lpf961", *lpfRx);
    signalVector *a = createLPF(0.3F, 651, 1.0F);   putv("lpf651b", *a);  delete a;
    signalVector *b = createLPF(0.1F, 961, 2.5F);   putv("lpf961b", *b);  delete b;
    signalVector *c = createLPF(0.1F, 500, 96.0F);  putv("lpf500", *c);   delete c;
  }
  // polyphaseResampleVector with filters that are NOT the two tables: rescaled table, short real filter, complex filter
  {
    signalVector x(300);
    randomise(x, 1000.0F);
    signalVector *g = createLPF(0.1F, 961, 2.5F);
    signalVector *a = polyphaseResampleVector(x, 65, 96, g);    putv("pr961b", *a);  delete a;  delete g;
    signalVector h(101);
    for (int k = 0; k < 101; k++) h[k] = complex(0.02F * (50 - abs(k - 50)) * (0.5F + unif()), 0.0F);
    h.isRealOnly(true);
    signalVector *b = polyphaseResampleVector(x, 3, 2, &h);     putv("pr32", *b);    delete b;
    signalVector hc(64);
    randomise(hc, 0.2F);
    signalVector *c = polyphaseResampleVector(x, 2, 5, &hc);    putv("pr25c", *c);   delete c;
    signalVector *d = polyphaseResampleVector(x, 96, 65, lpfTx); putv("prtx", *d);   delete d;
  }
  // convolve / correlate spans the flow test does not use
  {
    signalVector a(50), b(9), br(7);
    randomise(a, 4.0F); randomise(b, 1.0F); randomise(br, 1.0F);
    br.isRealOnly(true);
    signalVector *c1 = convolve(&a, &b, NULL, FULL_SPAN);      putv("cvfull", *c1);  delete c1;
    signalVector *c2 = convolve(&a, &br, NULL, START_ONLY);    putv("cvstart", *c2); delete c2;
    signalVector *c3 = convolve(&a, &b, NULL, WITH_TAIL);      putv("cvtail", *c3);  delete c3;
    signalVector *c4 = correlate(&a, &b, NULL, NO_DELAY);      putv("crnd", *c4);    delete c4;
    float idx = 0, avg = 0;
    complex pk = peakDetect(a, &idx, &avg);
    float pm[4] = {pk.real(), pk.imag(), idx, avg};
    put("peak", pm, sizeof pm);
    complex ip = interpolatePoint(a, 17.3F);
    put("interp", &ip, sizeof ip);
  }
  // four threads at once
  {
    Job jobs[4];
    pthread_t th[4];
    for (int t = 0; t < 4; t++) { jobs[t].id = t; jobs[t].pulse = pulse; jobs[t].lpfRx = lpfRx; jobs[t].lpfTx = lpfTx; }
    for (int t = 0; t < 4; t++) pthread_create(&th[t], NULL, hammer, &jobs[t]);
    for (int t = 0; t < 4; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < 4; t++) {
      put("thmeta", jobs[t].meta, sizeof jobs[t].meta);
      put("thup", jobs[t].firstUp, sizeof jobs[t].firstUp);
      put("thdown", jobs[t].firstDown, sizeof jobs[t].firstDown);
      put("thmod", jobs[t].firstMod, sizeof jobs[t].firstMod);
    }
  }
  fclose(gOut);
  delete lpfTx; delete lpfRx; delete pulse;
  sigProcLibDestroy();
  printf("surface ok\n");
  return 0;
}
