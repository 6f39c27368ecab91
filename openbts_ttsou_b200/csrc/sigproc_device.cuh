// sigproc_device.cuh -- the burst-DSP algorithms as thread-serial functions over strided views.
//
// One CUDA thread owns one burst (or one vector) and walks it in exactly the order the reference's
// scalar loops do, so every comparison the reference makes (peak search, early/late balancing,
// detection thresholds, symbol decisions) sees bit-identical floats.  Parallelism comes from the
// batch: 32 bursts per warp, each lane reading its own column of a transposed shared-memory tile
// (View<33>, conflict-free), thousands of warps per launch.  The same functions run over plain
// global vectors (View<1>) for the single-vector sigProcLib.h surface.
//
// All functions are __host__ __device__ so tests/hostemu can replay the kernels' control flow on a
// CPU without a GPU; the product never calls them on the host.
//
// Reference lines are cited per function (Transceiver/sigProcLib.cpp unless another file is named).
#pragma once
#include <type_traits>
#include "cplx.cuh"
#include "tables.h"

namespace btsdsp {

constexpr float kPiF = (float)3.14159265358979323846;                 // M_PI_F   :43
constexpr float k2PiF = (float)(2.0 * 3.14159265358979323846);        // M_2PI_F  :44
constexpr float k1_2PiF = 1 / k2PiF;                                   // M_1_2PI_F :45

// cosLookup / sinLookup :163-188.  The reference range-reduces with `while (arg > 1) arg -= 1;
// while (arg < 0) arg += 1;`.  Every step of those loops but the last +1 is exact (the magnitude
// shrinks on the same ulp grid), so one subtraction of ceil(arg)-1, or one rounded addition of
// ceil(-arg), yields the same float -- and cannot spin on Inf/NaN/huge inputs as the loops would.
BTS_HD float trig_lookup(const float *__restrict__ T, float x) {
  float arg = BTS_MUL(x, k1_2PiF);
  if (arg > 1.0F) arg = BTS_SUB(arg, BTS_SUB(ceilf(arg), 1.0F));
  else if (arg < 0.0F) arg = BTS_ADD(arg, ceilf(-arg));
  const float argT = BTS_MUL(arg, (float)kTrig);
  int argI = (int)argT;
  argI = argI < 0 ? 0 : (argI > kTrig ? kTrig : argI);                 // only NaN/Inf get here out of range
  const float delta = BTS_SUB(argT, (float)argI);
  const float iDelta = BTS_SUB(1.0F, delta);
  return BTS_ADD(BTS_MUL(iDelta, T[argI]), BTS_MUL(delta, T[argI + 1]));
}

// sinc :567-571
BTS_HD float sinc_exact(const DevTables *__restrict__ T, float x) {
  if ((x >= 0.01F) || (x <= -0.01F)) return BTS_DIV(trig_lookup(T->sinT, x), x);
  return 1.0F;
}

// expjLookup :192-204
BTS_HD cf expj_lookup(const DevTables *__restrict__ T, float x) {
  return mk(trig_lookup(T->cosT, x), trig_lookup(T->sinT, x));
}

// One output sample of convolve() :267-408 (symmetry NONE), complex data x complex taps.
// c[t] = sum_{k=0..lb-1} a[t-k]*b[k], taps with t-k >= la skipped, loop ends at t-k < 0.
// `corr` applies correlate()'s tap transform on the fly (:474-503): tap k = conj(b[lb-1-k]).
template <int S>
BTS_HD cf conv_cc_at(View<S> a, int la, const cf *__restrict__ b, int lb, int t, bool corr) {
  cf sum = mk(0.0F, 0.0F);
  for (int k = 0; k < lb; k++) {
    int ai = t - k;
    if (ai < 0) break;
    if (ai < la) {
      cf tap = corr ? cconj(b[lb - 1 - k]) : b[k];
      sum = cadd(sum, cmul(a.ld(ai), tap));
    }
  }
  return sum;
}

// Same, complex data x real taps (the `b->isRealOnly()` branch :345-354): sum += a * b.real().
template <int S>
BTS_HD cf conv_cr_at(View<S> a, int la, const float *__restrict__ b, int lb, int t) {
  cf sum = mk(0.0F, 0.0F);
  for (int k = 0; k < lb; k++) {
    int ai = t - k;
    if (ai < 0) break;
    if (ai < la) sum = cadd(sum, cmulr(a.ld(ai), b[k]));
  }
  return sum;
}

BTS_HD int no_delay_start(int lb) { return (lb % 2) ? lb / 2 : lb / 2 - 1; }   // :295-301

// interpolatePoint :639-659 (complex input).
template <int S>
BTS_HD cf interp_point(const DevTables *__restrict__ T, View<S> sig, int n, float ix) {
  int start = (int)(floorf(ix) - 10);
  if (start < 0) start = 0;
  int end = (int)(floorf(ix) + 11);
  if ((unsigned long long)(unsigned)end > (unsigned long long)n - 1) end = n - 1;
  cf p = mk(0.0F, 0.0F);
  for (int i = start; i < end; i++)
    p = cadd(p, cmulr(sig.ld(i), sinc_exact(T, BTS_MUL(kPiF, BTS_SUB((float)i, ix)))));
  return p;
}

// interpolatePoint for ix on the 1/512 grid (all peakDetect ever asks for): the 21 sinc values are a
// row of DevTables::sinc_grid; identical floats, no trig/divide in the loop.
template <int S>
BTS_HD cf interp_point_grid(const float *__restrict__ row, View<S> sig, int n, int I) {
  int start = I - 10;
  if (start < 0) start = 0;
  int end = I + 11;
  if ((unsigned long long)(unsigned)end > (unsigned long long)n - 1) end = n - 1;
  cf p = mk(0.0F, 0.0F);
  for (int i = start; i < end; i++) p = cadd(p, cmulr(sig.ld(i), row[i - I + 10]));
  return p;
}

// peakDetect :663-711.  Returns the interpolated peak; *peakIndex in samples (integer + k/512).
// GRID selects the table-driven interpolator (exactly equal; see DevTables::sinc_grid).
template <int S, bool GRID>
BTS_HD cf peak_detect(const DevTables *__restrict__ T, View<S> v, int n, float *peakIndex, float *avgPwr) {
  float maxVal = 0.0F, maxIndex = -1.0F, sumPower = 0.0F;
  for (int i = 0; i < n; i++) {
    float p = cnorm2(v.ld(i));
    if (p > maxVal) { maxVal = p; maxIndex = (float)i; }
    sumPower = BTS_ADD(sumPower, p);
  }
  float early = BTS_SUB(maxIndex, 1.0F), late = BTS_ADD(maxIndex, 1.0F), incr = 0.5F;
  cf pk;
  if (GRID) {
    // early = I + j/512 exactly; late = early + 2 shares the row; so does the final early + 1.
    int e512 = ((int)maxIndex - 1) * kSincGrid;                        // early * 512, exact integer
    int step = kSincGrid / 2;
    while (step >= 1) {
      int I = e512 >> 9, j = e512 & (kSincGrid - 1);                   // floor and fraction (two's complement ok)
      const float *row = T->sinc_grid[j];
      float e = cnorm2(interp_point_grid<S>(row, v, n, I)), l = cnorm2(interp_point_grid<S>(row, v, n, I + 2));
      if (e < l) e512 += step;
      else if (e > l) e512 -= step;
      else break;
      step >>= 1;
    }
    int I = (e512 >> 9) + 1, j = e512 & (kSincGrid - 1);
    maxIndex = BTS_ADD((float)e512 * (1.0F / kSincGrid), 1.0F);        // exact: |e512| < 2^24
    pk = interp_point_grid<S>(T->sinc_grid[j], v, n, I);
  } else {
    while (incr > 1.0F / 1024.0F) {
      float e = cnorm2(interp_point<S>(T, v, n, early)), l = cnorm2(interp_point<S>(T, v, n, late));
      if (e < l) early = BTS_ADD(early, incr);
      else if (e > l) early = BTS_SUB(early, incr);
      else break;
      incr = incr * 0.5F;
      late = BTS_ADD(early, 2.0F);
    }
    maxIndex = BTS_ADD(early, 1.0F);
    pk = interp_point<S>(T, v, n, maxIndex);
  }
  if (peakIndex) *peakIndex = maxIndex;
  if (avgPwr) *avgPwr = BTS_DIV(BTS_SUB(sumPower, cnorm2(pk)), (float)(n - 1));
  return pk;
}

// The 21 sinc taps of delayVector :583-588 for fractional offset `frac`.
BTS_HD void delay_taps(const DevTables *__restrict__ T, float frac, float *taps) {
  float f512 = frac * (float)kSincGrid;
  int j = (int)f512;
  if ((float)j == f512 && j >= 0 && j < kSincGrid) {                   // on the 1/512 grid: table row, reversed role:
    // tap i multiplies sinc(pi*(i-10-frac)) = sinc(pi*(m - j/512)) with m = i-10
    const float *row = T->sinc_grid[j];
    for (int i = 0; i < 21; i++) taps[i] = row[i];
  } else {
    for (int i = 0; i < 21; i++) taps[i] = sinc_exact(T, BTS_MUL(kPiF, BTS_SUB((float)(i - 10), frac)));
  }
}

// delayVector :573-616, in place on v with scratch tmp (same length).  v_real is never set on the path.
template <int S>
BTS_HD void delay_vector(const DevTables *__restrict__ T, View<S> v, int n, float delay, View<S> tmp) {
  int intOffset = (int)floorf(delay);
  float frac = BTS_SUB(delay, (float)intOffset);
  bool shifted = false;
  if ((double)fabsf(frac) > 1e-2) {
    float taps[21];
    delay_taps(T, frac, taps);
    for (int i = 0; i < n; i++) tmp.st(i, conv_cr_at<S>(v, n, taps, 21, i + 10));   // NO_DELAY, Lb = 21 -> start 10
    shifted = true;
  }
  // integer part :597-613 (reads `tmp` when the fractional filter ran, else v itself, in a safe order)
  if (intOffset < 0) {
    int io = -intOffset, w = 0;
    for (int s = io; s < n; s++) v.st(w++, shifted ? tmp.ld(s) : v.ld(s));
    while (w < n) v.st(w++, mk(0.0F, 0.0F));
  } else {
    int w = n - 1;
    for (int s = n - 1 - intOffset; s >= 0; s--) v.st(w--, shifted ? tmp.ld(s) : v.ld(s));
    while (w >= 0) v.st(w--, mk(0.0F, 0.0F));
  }
}


// vectorSlicer :507-519 on one value
BTS_HD float soft_slice(float x) {
  float s = BTS_MUL(0.5F, BTS_ADD(x, 1.0F));        // 0.5*(x+1.0F): the double product is an exact scaling
  if (s > 1.0F) s = 1.0F;
  if (s < 0.0F) s = 0.0F;
  return s;
}

// energyDetect :916-932
template <int S>
BTS_HD bool energy_detect(View<S> v, int n, unsigned win, float thr, float *avg) {
  float energy = 0.0F;
  if (win > (unsigned)n) win = n;
  for (unsigned i = 0; i < win; i++) energy = BTS_ADD(energy, cnorm2(v.ld(i)));
  float a = BTS_DIV(energy, (float)win);
  if (avg) *avg = a;
  return a > BTS_MUL(thr, thr);
}

// analyzeTrafficBurst :935-1037.  burst = the whole received slot; corr/tmp = scratch of 36*sps samples.
// chan (6*sps, thread-local) and *chanOff are written only when request && detected.
template <int S, bool GRID>
BTS_HD bool analyze_traffic(const DevTables *__restrict__ T, View<S> burst, int tsc, float thr, int sps,
                            View<S> corr, View<S> tmp, cf *amplitude, float *TOA, bool request, cf *chan,
                            float *chanOff) {
  const int L = 36 * sps, lb = 16 * sps;
  View<S> seg = burst.at(56 * sps);
  const cf *seq = T->mid_seq[tsc];
  const int start = no_delay_start(lb);
  for (int i = 0; i < L; i++) corr.st(i, conv_cc_at<S>(seg, L, seq, lb, start + i, true));
  float toa;
  cf amp = peak_detect<S, GRID>(T, corr, L, &toa, nullptr);
  if ((toa < 0.0F) || (toa > (float)L)) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const int p = (int)rintf(toa);
  float valley = 0.0F;
  int numRms = 0;
  for (int i = 2 * sps; i <= 5 * sps; i++) {
    if (p - i >= 0) { valley = BTS_ADD(valley, cnorm2(corr.ld(p - i))); numRms++; }
    if (p + i < L)  { valley = BTS_ADD(valley, cnorm2(corr.ld(p + i))); numRms++; }
  }
  if (numRms < 2) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const float RMS = (float)((double)BTS_SQRT(BTS_DIV(valley, (float)numRms)) + 0.00001);
  const float peakToMean = BTS_DIV(cabs_(amp), RMS);
  amp = cdiv(amp, T->mid_gain[tsc]);
  toa = BTS_SUB(toa, T->mid_toa[tsc]);
  toa = BTS_SUB(toa, (float)((66 - 56) * sps));
  *amplitude = amp;
  *TOA = toa;
  const bool detected = peakToMean > thr;
  if (request && detected) {
    const float TOAoffset = BTS_ADD(T->mid_toa[tsc], (float)((66 - 56) * sps));
    delay_vector<S>(T, corr, L, -toa, tmp);
    const int clen = 6 * sps;
    float maxEnergy = -1.0F;
    int maxI = -1;
    for (int i = 0; i < 7; i++) {
      const float pos = BTS_ADD(TOAoffset, (float)((i - 5) * sps));
      if (BTS_ADD(pos, (float)clen) > (float)L) continue;
      if (pos < 0.0F) continue;
      const int s0 = (int)floorf(pos);
      float energy = 0.0F;
      for (int j = 0; j < clen; j++) energy = BTS_ADD(energy, cnorm2(corr.ld(s0 + j)));
      if ((double)energy > 0.95 * (double)maxEnergy) { maxI = i; maxEnergy = energy; }
    }
    const int s0 = (int)floorf(BTS_ADD(TOAoffset, (float)((maxI - 5) * sps)));
    const cf g = cdiv(mk(1.0F, 0.0F), T->mid_gain[tsc]);
    for (int j = 0; j < clen; j++) chan[j] = cmul(corr.ld(s0 + j), g);
    *chanOff = (float)(5 * sps - maxI);
  }
  return detected;
}

// ---- the reference's second transceiver variant (Transceiver52M/sigProcLib.cpp) -----------------
// energyDetect there strides its window by four samples (:944-963)
template <int S>
BTS_HD bool energy_detect_52m(View<S> v, int n, unsigned win, float thr, float *avg) {
  float energy = 0.0F;
  if (win > (unsigned)n) win = n;
  for (unsigned i = 0; i < win; i++) energy = BTS_ADD(energy, cnorm2(v.ld(4 * i)));
  float a = BTS_DIV(energy, (float)win);
  if (avg) *avg = a;
  return a > BTS_MUL(thr, thr);
}
// analyzeTrafficBurst with a search window of +-maxTOA symbols (:966-1077): the midamble correlation is evaluated only
// at the 2*maxTOA+1 lags around the expected peak (convolve's CUSTOM span), TOA is counted from the window centre.
// The midamble table is the main variant's (its stored TOA is 5*sps less than this variant's, :779-828 vs 52M).
// corr / tmp: scratch of 2*maxTOA+1 samples.  maxTOA <= 60.
template <int S>
BTS_HD bool analyze_traffic_52m(const DevTables *__restrict__ T, View<S> burst, int tsc, float thr, int sps, unsigned maxTOA,
                                View<S> corr, View<S> tmp, cf *amplitude, float *TOA, bool request, cf *chan, float *chanOff) {
  if (maxTOA < 3u * sps) maxTOA = 3 * sps;
  unsigned spanTOA = maxTOA;
  if (spanTOA < 5u * sps) spanTOA = 5 * sps;
  const int startIx = (66 - (int)spanTOA) * sps, endIx = (66 + 16 + (int)spanTOA) * sps;
  const int windowLen = endIx - startIx, corrLen = 2 * (int)maxTOA + 1, lb = 16 * sps;
  const float midTOA = BTS_ADD(T->mid_toa[tsc], (float)(5 * sps));       // exact: undoes the main variant's "- 5*sps"
  const unsigned expectedPeak = (unsigned)round((double)BTS_ADD(midTOA, (float)((lb - 1) / 2)));
  View<S> seg = burst.at(startIx);
  const cf *seq = T->mid_seq[tsc];
  const int start = (int)expectedPeak - (int)maxTOA;
  for (int i = 0; i < corrLen; i++) corr.st(i, conv_cc_at<S>(seg, windowLen, seq, lb, start + i, true));
  float toa;
  cf amp = peak_detect<S, false>(T, corr, corrLen, &toa, nullptr);
  if ((toa < 0.0F) || (toa > (float)corrLen)) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const int p = (int)rintf(toa);
  float valley = 0.0F;
  int numRms = 0;
  for (int i = 2 * sps; i <= 5 * sps; i++) {
    if (p - i >= 0)      { valley = BTS_ADD(valley, cnorm2(corr.ld(p - i))); numRms++; }
    if (p + i < corrLen) { valley = BTS_ADD(valley, cnorm2(corr.ld(p + i))); numRms++; }
  }
  if (numRms < 2) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const float RMS = (float)((double)BTS_SQRT(BTS_DIV(valley, (float)numRms)) + 0.00001);
  const float peakToMean = BTS_DIV(cabs_(amp), RMS);
  amp = cdiv(amp, T->mid_gain[tsc]);
  toa = BTS_SUB(toa, (float)maxTOA);
  *amplitude = amp;
  *TOA = toa;
  const bool detected = peakToMean > thr;
  if (request && detected) {
    const float TOAoffset = (float)maxTOA;
    delay_vector<S>(T, corr, corrLen, -toa, tmp);
    const int clen = 6 * sps;
    float maxEnergy = -1.0F;
    int maxI = -1;
    for (int i = 0; i < 7; i++) {
      const float pos = BTS_ADD(TOAoffset, (float)((i - 5) * sps));
      if (BTS_ADD(pos, (float)clen) > (float)corrLen) continue;
      if (pos < 0.0F) continue;
      const int s0 = (int)floorf(pos);
      float energy = 0.0F;
      for (int j = 0; j < clen; j++) energy = BTS_ADD(energy, cnorm2(corr.ld(s0 + j)));
      if ((double)energy > 0.95 * (double)maxEnergy) { maxI = i; maxEnergy = energy; }
    }
    // with a window narrower than 6*sps + 5*sps no position qualifies and the reference indexes with maxI = -1; keep that
    const int s0 = (int)floorf(BTS_ADD(TOAoffset, (float)((maxI - 5) * sps)));
    const cf g = cdiv(mk(1.0F, 0.0F), T->mid_gain[tsc]);
    for (int j = 0; j < clen; j++) chan[j] = (s0 + j >= 0 && s0 + j < corrLen) ? cmul(corr.ld(s0 + j), g) : mk(0.0F, 0.0F);
    *chanOff = (float)(5 * sps - maxI);
  }
  return detected;
}

// what analyze_traffic_52m touches at sps == 1: burst rows [startIx, startIx + windowLen) and corrLen lags (the tile-staged
// kernel k_detect_52m stages exactly these; tests/hostemu poisons everything else)
struct Geo52 { int startIx, windowLen, corrLen; };
BTS_HD Geo52 geo_52m(unsigned max_toa) {
  if (max_toa < 3u) max_toa = 3;
  unsigned span = max_toa < 5u ? 5u : max_toa;
  return Geo52{66 - (int)span, 16 + 2 * (int)span, 2 * (int)max_toa + 1};
}

// detectRACHBurst :860-914.  corr = scratch of n samples.
template <int S, bool GRID>
BTS_HD bool detect_rach(const DevTables *__restrict__ T, View<S> burst, int n, float thr, int sps, View<S> corr,
                        cf *amplitude, float *TOA) {
  const int lb = 41 * sps;
  const int start = no_delay_start(lb);
  for (int i = 0; i < n; i++) corr.st(i, conv_cc_at<S>(burst, n, T->rach_seq, lb, start + i, true));
  float toa;
  cf pk = peak_detect<S, GRID>(T, corr, n, &toa, nullptr);
  if ((toa < 0.0F) || (toa > (float)n)) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const int p = (int)rintf(toa);
  float valley = 0.0F, numSamples = 0.0F;
  for (int i = 57 * sps; i <= 107 * sps; i++) {
    if (p + i >= n) break;
    valley = BTS_ADD(valley, cnorm2(corr.ld(p + i)));
    numSamples = numSamples + 1.0F;
  }
  if (numSamples < 2.0F) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const float RMS = (float)((double)BTS_SQRT(BTS_DIV(valley, numSamples)) + 0.00001);
  const float peakToMean = BTS_DIV(cabs_(pk), RMS);
  *amplitude = cdiv(pk, T->rach_gain);
  *TOA = BTS_SUB(BTS_SUB(toa, T->rach_toa), (float)(8 * sps));
  return peakToMean > thr;
}

// designDFE :1246-1340 (Al-Dhahir & Cioffi fast Cholesky recursion).  NF/NU > 0 fix the sizes at compile
// time (everything unrolls into registers); NF == 0 takes them from nf/nu at run time (<= kDfeMax).
template <int NF, int NU>
BTS_HD void design_dfe(const cf *chan, int nu_rt, float SNR, int nf_rt, cf *w, cf *b) {
  const int Nf = NF ? NF : nf_rt, nu = NF ? NU : nu_rt;
  constexpr int MF = NF ? NF : kDfeMax, ML = NF ? NF + NU : 2 * kDfeMax;
  cf G0[MF], G1[MF], G0n[MF], G1n[MF], v[MF];
  cf L[MF][ML];
#pragma unroll
  for (int j = 0; j < MF; j++) { G0[j] = mk(0.0F, 0.0F); G1[j] = mk(0.0F, 0.0F); }
#pragma unroll
  for (int i = 0; i < MF; i++)
#pragma unroll
    for (int j = 0; j < ML; j++) L[i][j] = mk(0.0F, 0.0F);
  G0[0] = mk(BTS_DIV(1.0F, BTS_SQRT(SNR)), 0.0F);                      // 1.0/sqrtf(SNR) :1261 (double div is innocuous)
#pragma unroll
  for (int j = 0; j < MF; j++) if (j <= nu && j < Nf) G1[j] = cconj(chan[j]);
  float d = 0.0F;
#pragma unroll
  for (int i = 0; i < MF; i++) {
    if (i < Nf) {
      d = BTS_ADD(cnorm2(G0[0]), cnorm2(G1[0]));
      const cf g0c = cconj(G0[0]), g1c = cconj(G1[0]);
#pragma unroll
      for (int j = 0; j < MF; j++)
        if (j < Nf && i + j < Nf + nu) L[i][i + j] = cdivr(cadd(cmul(G0[j], g0c), cmul(G1[j], g1c)), d);
      const cf k = cdiv(G1[0], G0[0]);
      if (i != Nf - 1) {
        const cf kc = cconj(k), nk = cmulr(k, -1.0F);
#pragma unroll
        for (int j = 0; j < MF; j++) if (j < Nf) {
          G0n[j] = cadd(cmul(G1[j], kc), G0[j]);
          G1n[j] = cadd(cmul(G0[j], nk), G1[j]);
        }
        const cf s = mk(BTS_DIV(1.0F, BTS_SQRT(BTS_ADD(1.0F, cnorm2(k)))), 0.0F);   // :1294-1295
#pragma unroll
        for (int j = 0; j < MF; j++) if (j < Nf) {
          G0[j] = cmul(G0n[j], s);
          G1[j] = cmul((j + 1 < Nf) ? G1n[j + 1] : mk(0.0F, 0.0F), s);             // delayVector(G1new,-1.0)
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < (NF ? NU : kDfeMax); j++)
    if (j < nu) b[j] = cconj(cmul(L[Nf - 1][Nf + j], mk(-1.0F, 0.0F)));
  v[Nf - 1] = mk(1.0F, 0.0F);
#pragma unroll
  for (int kk = MF - 2; kk >= 0; kk--) {
    if (kk <= Nf - 2) {
      cf vk = mk(0.0F, 0.0F);
#pragma unroll
      for (int j = 1; j < MF; j++) if (j >= kk + 1 && j < Nf) vk = csub(vk, cmul(v[j], L[kk][j]));
      v[kk] = vk;
    }
  }
#pragma unroll
  for (int i = 0; i < MF; i++) {
    if (i < Nf) {
      cf wi = mk(0.0F, 0.0F);
      const int endPt = (nu < (Nf - 1 - i)) ? nu : (Nf - 1 - i);
#pragma unroll
      for (int k = 0; k < MF; k++) if (k < endPt + 1 && i + k < MF) wi = cadd(wi, cmul(v[i + k], cconj(chan[k])));
      w[i] = cdivr(wi, d);
    }
  }
}

// (designDFE with its outer iteration kept as a loop -- L rows parked in the lane's tile column, ~500-instruction body run seven
// times instead of ~4000 straight-line instructions -- was measured: 0.484 ms against 0.457 ms per 800 280 bursts for the
// detect kernel, bit-identical; the serial G0/G1 chain no longer overlaps across iterations.  profiles/README.md r3d.)

// equalizeBurst :1343-1399.  burst is delayed in place (as the reference does), tmp = scratch of n
// samples that ends up holding the post-feedback symbols; soft[m*SS] receives the n soft bits.
template <int S, int SS>
BTS_HD void equalize_burst(const DevTables *__restrict__ T, View<S> burst, int n, float TOA, const cf *w, int nw,
                           const cf *b, int nb, View<S> tmp, float *soft) {
  delay_vector<S>(T, burst, n, -TOA, tmp);
  // feed-forward: FULL_SPAN convolution, samples nw-1 .. nw-1+n kept  :1352-1356
  for (int m = 0; m < n; m++) tmp.st(m, conv_cc_at<S>(burst, n, w, nw, m + nw - 1, false));
  // decision feedback :1367-1384.  The reference overwrites postForward[m] with the decided symbol and
  // reads the previous nb of them back; only those nb are ever needed, so they live in registers and
  // tmp[m] is free to receive the soft bit (soft may alias tmp: it is written after tmp[m] is read).
  cf hist[kDfeMax];
#pragma unroll
  for (int k = 0; k < kDfeMax; k++) hist[k] = mk(0.0F, 0.0F);
  for (int m = 0; m < n; m++) {
    cf y = tmp.ld(m);
#pragma unroll
    for (int k = 0; k < kDfeMax; k++)
      if (k < nb && m - 1 - k >= 0) y = cadd(y, cmul(b[k], hist[k]));
    y = cmul(y, T->revrot[m]);
    const float out = y.x;
#pragma unroll
    for (int k = kDfeMax - 1; k >= 1; k--) if (k < nb) hist[k] = hist[k - 1];
    hist[0] = cmul(mk((out > 0.0F) ? 1.0F : -1.0F, 0.0F), T->rot[m]);
    soft[m * SS] = soft_slice(out);
  }
}

// demodulateBurst :1056-1097 (+ decimateVector :1039-1053).  x = private copy of the burst (modified),
// tmp = scratch; returns the number of soft bits written.
template <int S, int SS>
BTS_HD int demodulate_burst(const DevTables *__restrict__ T, View<S> x, int n, int sps, cf channel, float TOA,
                            View<S> tmp, float *soft) {
  const cf s = cdiv(mk(1.0F, 0.0F), channel);
  for (int i = 0; i < n; i++) x.st(i, cmul(x.ld(i), s));
  delay_vector<S>(T, x, n, -TOA, tmp);
  const int m = (sps > 1) ? n / sps : n;
  for (int i = 0; i < m; i++) {
    cf r = cmul(T->revrot[i * sps], x.ld(i * sps));
    soft[i * SS] = soft_slice(r.x);
  }
  return m;
}

// One output sample of modulateBurst :521-565 = GMSKRotate (:232-247, real-only branch) followed by
// convolve(.., pulse, NO_DELAY) with the real-only pulse.  bits: one per byte, value in bit 0.
BTS_HD cf modulate_at(const DevTables *__restrict__ T, const unsigned char *__restrict__ bits, int nbits, int n,
                      int sps, const cf *__restrict__ pulse, int plen, bool pulse_real, int t) {
  cf sum = mk(0.0F, 0.0F);
  const int start = no_delay_start(plen);
  for (int k = 0; k < plen; k++) {
    int ai = t + start - k;
    if (ai < 0) break;
    if (ai < n) {
      float sym = 0.0F;
      if (ai % sps == 0 && ai / sps < nbits) sym = (bits[ai / sps] & 0x01) ? 1.0F : -1.0F;   // 2.0*(bit&1)-1.0 :548
      cf x = cmulr(T->rot[ai], sym);
      sum = cadd(sum, pulse_real ? cmulr(x, pulse[k].x) : cmul(x, pulse[k]));
    }
  }
  return sum;
}


// One output of the per-chunk polyphase resampler (polyphaseResampleVector :1157-1210 as called per
// chunk by radioInterface.cpp:142 / :245).  x = the chunk's input incl. history (nin samples),
// hp[k*HP + branch] = lpf[branch + P*k].  m = index among the KEPT outputs (drop already removed).
// Outputs whose first taps would read past the chunk (in >= nin) skip those taps (:1183-1186).
template <int P, int Q, int NTAPS, int NPOLY, int HP>
BTS_HD cf resample_at(const cf *__restrict__ x, int nin, const float *__restrict__ hp, int drop, int m) {
  const int t = Q * (m + drop + (NTAPS - 1) / 2 / Q);                   // outputIx*Q, outputIx0 = (L-1)/2/Q :1177
  const int br = t % P;
  int ix = t / P;
  int k = 0;
  if (ix >= nin) { k = ix - (nin - 1); ix = nin - 1; }
  cf sum = mk(0.0F, 0.0F);
  for (; k < NPOLY && br + P * k < NTAPS && ix >= 0; k++, ix--) sum = cadd(sum, cmulr(x[ix], hp[k * HP + br]));
  return sum;
}

// TX tail: scaleVector(.,13500.0) (radioInterface.cpp:148) then USRPifyVector's (short) casts (:74-89);
// the x86 cast converts to int32 toward zero and keeps the low 16 bits.
BTS_HD short2 tx_quantise(cf sum) {
  const cf s = cmul(sum, mk(13500.0F, 0.0F));
  short2 o;
  o.x = (short)(int)s.x;
  o.y = (short)(int)s.y;
  return o;
}

// ---- tuned RX resampler (see resample.cu for the derivation and the kernel) --------------------------------------
constexpr int kRxDropC = 130;   // INHISTORY outputs dropped per chunk (radioInterface.cpp:249-252)
__host__ __device__ constexpr int rx_ix(int r) { return (kRxQ * (r + kRxDropC + 5)) / kRxP; }
__host__ __device__ constexpr int rx_br(int r) { return (kRxQ * (r + kRxDropC + 5)) % kRxP; }
__host__ __device__ constexpr int rx_ntaps(int r) { return (kRxTaps - 1 - rx_br(r)) / kRxP + 1; }
__host__ __device__ constexpr int rx_trunc(int r) { return rx_ix(r) > 287 ? rx_ix(r) - 287 : 0; }   // 1055 - 96*8 = 287
// sample offset of (phase r, tap k) relative to the tile origin of the lane's period, then its padded address
__host__ __device__ constexpr int rx_c(int r, int k) { return rx_ix(r) - k - 96; }
__host__ __device__ constexpr int rx_pad(int c) { return c + 2 * (c / 96); }

constexpr int kRxTileRows = 34, kRxRowPitch = 98;                      // samples
constexpr int kRxTileIn = kRxTileRows * kRxRowPitch;                   // 3332 samples
constexpr int kRxTileOut = 32 * kRxP;                                  // 2080 samples

template <int R0, int NR = 5>
BTS_HD void rx_group(const float *__restrict__ taps, const cf *__restrict__ xl, cf *__restrict__ ol, bool q8) {
  constexpr int CLO = (rx_c(R0, 14)) & ~1;                             // even-aligned lowest sample offset
  constexpr int CHI = rx_c(R0 + NR - 1, 0);
  constexpr int NP = (CHI - CLO) / 2 + 1;                              // 16-byte pairs
  cf win[2 * NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const float4 v = *reinterpret_cast<const float4 *>(xl + rx_pad(CLO + 2 * p));
    win[2 * p] = mk(v.x, v.y);
    win[2 * p + 1] = mk(v.z, v.w);
  }
#pragma unroll
  for (int d = 0; d < NR; d++) {
    const int r = R0 + d;
    cf sum = mk(0.0F, 0.0F);
#pragma unroll
    for (int k = 0; k < 15; k++) {
      if (k < rx_ntaps(r)) {
        const cf x = win[rx_c(r, k) - CLO];
        cf p = pmul0(x, taps[r * 16 + k]);
        if (k < rx_trunc(r)) {                                         // only phases 60..64: dropped in period q == 8
          p.x = q8 ? 0.0F : p.x;
          p.y = q8 ? 0.0F : p.y;
        }
        sum = padd(sum, p);
      }
    }
    ol[r] = sum;
  }
}

// all 65 phases of one period: xl = the lane's row of the padded input tile, ol = its 65 outputs
BTS_HD void rx_period(const float *__restrict__ taps, const cf *__restrict__ xl, cf *__restrict__ ol, bool q8) {
  rx_group<0>(taps, xl, ol, q8);  rx_group<5>(taps, xl, ol, q8);  rx_group<10>(taps, xl, ol, q8); rx_group<15>(taps, xl, ol, q8);
  rx_group<20>(taps, xl, ol, q8); rx_group<25>(taps, xl, ol, q8); rx_group<30>(taps, xl, ol, q8); rx_group<35>(taps, xl, ol, q8);
  rx_group<40>(taps, xl, ol, q8); rx_group<45>(taps, xl, ol, q8); rx_group<50>(taps, xl, ol, q8); rx_group<55>(taps, xl, ol, q8);
  rx_group<60>(taps, xl, ol, q8);
}
// the same split in two for a pair of warps sharing one tile: phases 0..34 and 35..64
template <int HALF>
BTS_HD void rx_half(const float *__restrict__ taps, const cf *__restrict__ xl, cf *__restrict__ ol, bool q8) {
  if (HALF == 0) {
    rx_group<0>(taps, xl, ol, q8);  rx_group<5>(taps, xl, ol, q8);  rx_group<10>(taps, xl, ol, q8); rx_group<15>(taps, xl, ol, q8);
    rx_group<20>(taps, xl, ol, q8); rx_group<25>(taps, xl, ol, q8); rx_group<30>(taps, xl, ol, q8);
  } else {
    rx_group<35>(taps, xl, ol, q8); rx_group<40>(taps, xl, ol, q8); rx_group<45>(taps, xl, ol, q8); rx_group<50>(taps, xl, ol, q8);
    rx_group<55>(taps, xl, ol, q8); rx_group<60>(taps, xl, ol, q8);
  }
}
// ... or across NW warps: warp w does phases [65 w / NW, 65 (w+1) / NW), cut into register-blocked groups of <= 6
// The same sums kept in registers instead of stored: the gated form of rx_range holds a whole part's outputs back until its
// GATE has been called -- the kernel uses that to wait, as late as possible, until the output block may be overwritten
// (resample.cu); the default gate does nothing and stores each sum as soon as it is complete.
struct RxNoGate { BTS_HD void operator()() const {} };
template <int R0, int NR>
BTS_HD void rx_group_acc(const float *__restrict__ taps, const cf *__restrict__ xl, bool q8, cf *sums) {
  constexpr int CLO = (rx_c(R0, 14)) & ~1;
  constexpr int CHI = rx_c(R0 + NR - 1, 0);
  constexpr int NP = (CHI - CLO) / 2 + 1;
  cf win[2 * NP];
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const float4 v = *reinterpret_cast<const float4 *>(xl + rx_pad(CLO + 2 * p));
    win[2 * p] = mk(v.x, v.y);
    win[2 * p + 1] = mk(v.z, v.w);
  }
#pragma unroll
  for (int d = 0; d < NR; d++) {
    const int r = R0 + d;
    cf sum = mk(0.0F, 0.0F);
#pragma unroll
    for (int k = 0; k < 15; k++) {
      if (k < rx_ntaps(r)) {
        const cf x = win[rx_c(r, k) - CLO];
        cf p = pmul0(x, taps[r * 16 + k]);
        if (k < rx_trunc(r)) {
          p.x = q8 ? 0.0F : p.x;
          p.y = q8 ? 0.0F : p.y;
        }
        sum = padd(sum, p);
      }
    }
    sums[d] = sum;
  }
}
template <int A, int B>
BTS_HD void rx_range_acc(const float *__restrict__ taps, const cf *__restrict__ xl, bool q8, cf *sums) {
  constexpr int n = B - A, ng = (n + 5) / 6, first = (n + ng - 1) / ng;
  rx_group_acc<A, first>(taps, xl, q8, sums);
  if constexpr (n > first) rx_range_acc<A + first, B>(taps, xl, q8, sums + first);
}
template <int A, int B, class GATE = RxNoGate>
BTS_HD void rx_range(const float *__restrict__ taps, const cf *__restrict__ xl, cf *__restrict__ ol, bool q8, GATE gate = GATE()) {
  constexpr int n = B - A, ng = (n + 5) / 6, first = (n + ng - 1) / ng;
  if constexpr (!std::is_same<GATE, RxNoGate>::value) {                // all of the part's sums, then the gate, then the stores
    cf sums[n];
    rx_range_acc<A, B>(taps, xl, q8, sums);
    gate();
#pragma unroll
    for (int d = 0; d < n; d++) ol[A + d] = sums[d];
  } else {
    rx_group<A, first>(taps, xl, ol, q8);
    if constexpr (n > first) rx_range<A + first, B>(taps, xl, ol, q8);
  }
}
template <int NW, int W = 0, class GATE = RxNoGate>
BTS_HD void rx_part(int warp, const float *__restrict__ taps, const cf *__restrict__ xl, cf *__restrict__ ol, bool q8, GATE gate = GATE()) {
  if (warp == W) rx_range<kRxP * W / NW, kRxP * (W + 1) / NW, GATE>(taps, xl, ol, q8, gate);
  else if constexpr (W + 1 < NW) rx_part<NW, W + 1, GATE>(warp, taps, xl, ol, q8, gate);
}
// ---- tuned TX chain (see resample.cu: k_tx_fused) ------------------------------------------------------------------
// Per chunk the reference resamples 130 history + 585 new samples to 1056 outputs and sends outputs 192..1055
// (radioInterface.cpp:123-168): 864 = 9 x 96 outputs per 585 = 9 x 65 inputs, so the loop is periodic in
// (96 outputs, 65 inputs).  For global period G = 9*chunk + q and phase r = 0..95:
//     output 96 G + r = sum_k x[65 G - 130 + ix_r - k] * h[br_r + 96 k],  ix_r = (65 (r+197)) / 96, br_r = (65 (r+197)) % 96
// (o = r + 192 + 5, the filter's group delay in output samples, sigProcLib.cpp:1177), k < ntaps_r (6 or 7), with the two
// chunk effects kept: samples before the stream start are zero, and in the last period of a chunk (q == 8) phases with
// ix_r > 194 lose their first ix_r - 194 taps (the reference cannot see the next chunk, :1183-1186).
constexpr int kTxDropC = 192;
__host__ __device__ constexpr int tx_ix(int r) { return (kTxQ * (r + kTxDropC + 5)) / kTxP; }
__host__ __device__ constexpr int tx_br(int r) { return (kTxQ * (r + kTxDropC + 5)) % kTxP; }
__host__ __device__ constexpr int tx_ntaps(int r) { return (kTxTaps - 1 - tx_br(r)) / kTxP + 1; }
__host__ __device__ constexpr int tx_trunc(int r) { return tx_ix(r) > 194 ? tx_ix(r) - 194 : 0; }   // 714 - 65*8 = 194
// sample offset of (phase r, tap k) relative to the first sample of the lane's period (stream sample 65 G): -3 .. 67
__host__ __device__ constexpr int tx_c(int r, int k) { return tx_ix(r) - k - 130; }
constexpr int kTxHalo = 4;                                             // tile starts 4 samples before its first period

// NR adjacent phases of one period: xl = the lane's period start inside the tile, taps[r*8 + k] = lpf_tx[br_r + 96 k],
// ol[r] receives the quantised output (x13500, (short) casts: tx_quantise)
template <int R0, int NR>
BTS_HD void tx_group(const float *__restrict__ taps, const cf *__restrict__ xl, short2 *__restrict__ ol, bool q8) {
  constexpr int CLO = tx_c(R0, 6);                                     // lowest sample any phase of the group can read
  constexpr int CHI = tx_c(R0 + NR - 1, 0);
  constexpr int NW = CHI - CLO + 1;
  cf win[NW];
#pragma unroll
  for (int p = 0; p < NW; p++) win[p] = xl[CLO + p];
#pragma unroll
  for (int d = 0; d < NR; d++) {
    const int r = R0 + d;
    cf sum = mk(0.0F, 0.0F);
#pragma unroll
    for (int k = 0; k < 7; k++) {
      if (k < tx_ntaps(r)) {
        cf p = pmul0(win[tx_c(r, k) - CLO], taps[r * 8 + k]);
        if (k < tx_trunc(r)) {                                         // only phases 91..95, dropped in period q == 8
          p.x = q8 ? 0.0F : p.x;
          p.y = q8 ? 0.0F : p.y;
        }
        sum = padd(sum, p);
      }
    }
    ol[r] = tx_quantise(sum);
  }
}
template <int A, int B>
BTS_HD void tx_range(const float *__restrict__ taps, const cf *__restrict__ xl, short2 *__restrict__ ol, bool q8) {
  constexpr int n = B - A, ng = (n + 5) / 6, first = (n + ng - 1) / ng;
  tx_group<A, first>(taps, xl, ol, q8);
  if constexpr (n > first) tx_range<A + first, B>(taps, xl, ol, q8);
}
// warp `part` of NW does phases [96 part / NW, 96 (part+1) / NW)
template <int NW, int W = 0>
BTS_HD void tx_part(int part, const float *__restrict__ taps, const cf *__restrict__ xl, short2 *__restrict__ ol, bool q8) {
  if (part == W) tx_range<kTxP * W / NW, kTxP * (W + 1) / NW>(taps, xl, ol, q8);
  else if constexpr (W + 1 < NW) tx_part<NW, W + 1>(part, taps, xl, ol, q8);
}
// Modulated sample t of one burst (modulateBurst :521-565 at sps 1: x[n] = (2 bit - 1) rot[n], 3-tap real pulse,
// NO_DELAY) from a table q[ai*3 + k] = rot[ai] * pulse[k] (rounded once, as the reference rounds (rot*sym)*pulse:
// multiplying by sym = +-1 first only flips signs).  Terms with a zero symbol (guard, ai >= 148) add +-0 in the
// reference and are skipped here; the sum starts at +0 like the reference's.
BTS_HD cf tx_burst_sample(const cf *__restrict__ q, const unsigned char *__restrict__ bits, int t) {
  cf sum = mk(0.0F, 0.0F);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int ai = t + 1 - k;                                         // no_delay_start(3) = 1
    if (ai >= 0 && ai < 148) {
      const cf v = q[ai * 3 + k];
      sum = cadd(sum, (bits[ai] & 0x01) ? v : mk(-v.x, -v.y));
    }
  }
  return sum;
}
BTS_HD void tx_fill_q(const DevTables *__restrict__ T, cf *__restrict__ q, int i) {   // i < 148*3
  q[i] = cmulr(T->rot[i / 3], T->pulse[i % 3].x);
}
// slot and in-slot offset of sample w (0..624) of a 4-slot group (157/156/156/156)
BTS_HD void tx_slot_of(int w, int *j, int *t) {
  *j = w < 157 ? 0 : (w < 313 ? 1 : (w < 469 ? 2 : 3));
  *t = w - (*j == 0 ? 0 : (*j == 1 ? 157 : (*j == 2 ? 313 : 469)));
}
// taps[r*16 + k] = lpf_rx[br_r + 65 k] from the [branch][k] table
inline void rx_fill_taps(const DevTables *hostT, float *taps) {
  for (int r = 0; r < kRxP; r++)
    for (int k = 0; k < 16; k++) taps[r * 16 + k] = hostT->rx_poly[rx_br(r)][k];
}

}  // namespace btsdsp
