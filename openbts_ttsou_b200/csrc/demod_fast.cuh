// demod_fast.cuh -- the tuned one-burst-per-thread receive path for sps == 1 (k_demod_normal).
//
// Same arithmetic, operand order and accumulation order as sigproc_device.cuh (which restates the reference
// loop by loop) -- only the *schedule* differs:
//   * FIRs are register-blocked: a thread produces 4 (or 12) consecutive outputs per sweep over its column of
//     the transposed tile, so a sample is loaded once per 4 outputs instead of once per tap, the tap loops are
//     fully unrolled (no loop-carried index arithmetic, no bounds branches on interior blocks) and the 8
//     independent accumulation chains give a single warp enough ILP to keep its scheduler busy;
//   * the equaliser is streamed: delayVector's 21-tap fractional FIR, the 7-tap feed-forward filter and the
//     decision-feedback recursion run as one pipeline over the burst with a 10-sample register window, so the
//     delayed burst and the feed-forward output never exist in memory (one tile per warp instead of three);
//   * the channel-window search needs only 12 samples of the delayed correlation, so only those are computed;
//   * all sinc interpolators are rows of the 1/512-grid table (see tables.h) held in shared memory.
// Every function is __host__ __device__: tests/hostemu replays them on the CPU against the reference outputs.
#pragma once
#include "sigproc_device.cuh"

namespace btsdsp {

constexpr int kGridPitch = 21;      // floats per sinc-grid row in the compact (shared-memory) copy

// a sinc-grid table and its row pitch: the compact shared-memory copy (pitch 21) or DevTables::sinc_grid (pitch 24)
struct Grid {
  const float *p;
  int pitch;
  bool vec = false;     // rows are 16-byte aligned at pitch 24 and may be fetched as six 16-byte loads (the search kernels, which
                        // fetch ten rows per burst; the equaliser's single row per burst measured better as scalar loads)
};

BTS_HD void load_grid_row(Grid grid, int j, float s[21]) {
#ifdef __CUDA_ARCH__
  if (grid.vec) {                               // DevTables::sinc_grid: 96-byte rows, 16-byte aligned -> six vector loads
    const float4 *p = reinterpret_cast<const float4 *>(grid.p + j * 24);
    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
    s[8] = c.x; s[9] = c.y; s[10] = c.z; s[11] = c.w; s[12] = d.x; s[13] = d.y; s[14] = d.z; s[15] = d.w;
    s[16] = e.x; s[17] = e.y; s[18] = e.z; s[19] = e.w; s[20] = f.x;
    return;
  }
#endif
#pragma unroll
  for (int t = 0; t < 21; t++) s[t] = grid.p[j * grid.pitch + t];
}

// taps of correlate(window, midamble) (:474-503): tap[k] = conj(seq[15-k])
BTS_HD void load_corr_taps(const DevTables *__restrict__ T, int tsc, cf tap[16]) {
#pragma unroll
  for (int k = 0; k < 16; k++) tap[k] = cconj(T->mid_seq[tsc][15 - k]);
}

// correlate(burst[56..92), midamble, NO_DELAY): 36 outputs x 16 complex taps, start index 7 (:295-301).
// c[n] = sum_k win[n+7-k]*tap[k], window indices outside [0,36) are not part of the vector.
template <int S>
BTS_HD void corr36(View<S> win, View<S> out, const cf tap[16]) {
  cf taps[16];                                           // (-im, re) copies for cmac_tap
#pragma unroll
  for (int k = 0; k < 16; k++) taps[k] = cswapneg(tap[k]);
#pragma unroll 3                                         // 9 blocks: rolled 0.444 ms, three at a time 0.436, fully unrolled 0.487 (code size)
  for (int n0 = 0; n0 < 36; n0 += 4) {
    cf acc[4];
#pragma unroll
    for (int r = 0; r < 4; r++) acc[r] = mk(0.0F, 0.0F);
#pragma unroll
    for (int j = 0; j < 19; j++) {
      const int idx = n0 + 10 - j;                       // descending window index == ascending tap index
      if ((unsigned)idx < 36u) {
        const cf v = win.ld(idx);
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const int k = j + r - 3;                       // = (n0 + r + 7) - idx
          if (k >= 0 && k < 16) acc[r] = cmac_tap(acc[r], tap[k], taps[k], v);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) out.st(n0 + r, acc[r]);
  }
}

// interpolatePoint (:639-659) at ix = I + j/512 with the grid row s[] (s[t] = sinc(pi*((t-10) - j/512)))
template <int S>
BTS_HD cf interp21(const float s[21], View<S> c, int n, int I) {
  int start = I - 10;
  if (start < 0) start = 0;
  int end = I + 11;
  if ((unsigned long long)(unsigned)end > (unsigned long long)n - 1) end = n - 1;
  cf p = mk(0.0F, 0.0F);
#pragma unroll
  for (int t = 0; t < 21; t++) {
    const int i = I - 10 + t;
    if (i >= start && i < end) p = padd(p, pmul0(c.ld(i), s[t]));
  }
  return p;
}

// interpolatePoint when none of its 21 taps is clipped (0 <= I - 10 and I + 11 <= n - 1): no bounds logic
template <int S>
BTS_HD cf interp21_interior(const float s[21], View<S> c, int I) {
  const View<S> t0 = c.at(I - 10);
  cf p = mk(0.0F, 0.0F);
#pragma unroll
  for (int t = 0; t < 21; t++) p = padd(p, pmul0(t0.ld(t), s[t]));
  return p;
}

// peakDetect (:663-711) on the 1/512 grid; avgPwr is not needed by analyzeTrafficBurst.
// The early/late search interpolates at I, I + 2 and finally I + 1 with I in [imax - 2, imax - 1]: when 12 <= imax <= n - 13 no
// tap of any of them is clipped, and when that holds for every active lane of the warp the search runs without bounds logic
// (the same products and sums, so the same bits).
template <int S>
BTS_HD cf peak_detect_fast(Grid grid, View<S> c, int n, float *peakIndex) {
  float maxVal = 0.0F;
  int imax = -1;
  for (int i = 0; i < n; i++) {
    const float p = cnorm2(c.ld(i));
    if (p > maxVal) { maxVal = p; imax = i; }
  }
  const bool interior = imax >= 12 && imax <= n - 13;
#if defined(__CUDA_ARCH__) && !defined(BTS_NO_INTERIOR_SEARCH)
  const bool fast = __all_sync(__activemask(), interior);
#else
  const bool fast = false && interior;
#endif
  int e512 = (imax - 1) * kSincGrid;                     // earlyIndex * 512
  float s[21];
  if (fast) {
    for (int step = kSincGrid / 2; step >= 1; step >>= 1) {
      const int I = e512 >> 9, j = e512 & (kSincGrid - 1);
      load_grid_row(grid, j, s);
      const float e = cnorm2(interp21_interior<S>(s, c, I)), l = cnorm2(interp21_interior<S>(s, c, I + 2));
      if (e < l) e512 += step;
      else if (e > l) e512 -= step;
      else break;
    }
    load_grid_row(grid, e512 & (kSincGrid - 1), s);
    *peakIndex = BTS_ADD((float)e512 * (1.0F / kSincGrid), 1.0F);
    return interp21_interior<S>(s, c, (e512 >> 9) + 1);
  }
  for (int step = kSincGrid / 2; step >= 1; step >>= 1) {
    const int I = e512 >> 9, j = e512 & (kSincGrid - 1);
    load_grid_row(grid, j, s);
    const float e = cnorm2(interp21<S>(s, c, n, I)), l = cnorm2(interp21<S>(s, c, n, I + 2));
    if (e < l) e512 += step;
    else if (e > l) e512 -= step;
    else break;
  }
  load_grid_row(grid, e512 & (kSincGrid - 1), s);
  *peakIndex = BTS_ADD((float)e512 * (1.0F / kSincGrid), 1.0F);
  return interp21<S>(s, c, n, (e512 >> 9) + 1);
}

// the 21 taps of delayVector's fractional filter (:583-588); on the grid they are a table row
BTS_HD void load_delay_taps(Grid grid, const DevTables *__restrict__ T, float frac, float s[21]) {
  const float f512 = frac * (float)kSincGrid;
  const int j = (int)f512;
  if ((float)j == f512 && j >= 0 && j < kSincGrid) load_grid_row(grid, j, s);
  else {
#pragma unroll
    for (int i = 0; i < 21; i++) s[i] = sinc_exact(T, BTS_MUL(kPiF, BTS_SUB((float)(i - 10), frac)));
  }
}

// delayVector(corr, delay) (:573-616) evaluated only at the 12 positions [smin, smin+12) the channel-window
// search of analyzeTrafficBurst reads: out[n] = F[n - io] with F = the 21-tap filtered vector (or corr itself
// when the fraction is <= 0.01), zero where n - io falls outside the vector.
// PAD = rows L .. L+8 of c are scratch the caller no longer needs: when every active lane's 32 input rows
// [x0 - 10, x0 + 21] lie inside [0, L + 8] those rows are zeroed and the sweep runs without bounds logic.  A zero sample
// contributes a +-0 product, and an accumulator that started at +0 is never -0, so the sums are the same bits as with
// the out-of-range taps skipped.
template <int S, bool PAD = false>
BTS_HD void delayed12(Grid grid, const DevTables *__restrict__ T, View<S> c, int L, float delay,
                      int smin, cf dl[12]) {
  const int io = (int)floorf(delay);
  const float frac = BTS_SUB(delay, (float)io);
  const int x0 = smin - io;
  if ((double)fabsf(frac) > 1e-2) {
    float s[21];
    load_delay_taps(grid, T, frac, s);
    cf acc[12];
#pragma unroll
    for (int q = 0; q < 12; q++) acc[q] = mk(0.0F, 0.0F);
#if defined(__CUDA_ARCH__) && !defined(BTS_NO_PAD_DELAY)
    if (PAD && __all_sync(__activemask(), x0 >= 10 && x0 + 21 <= L + 8)) {
#pragma unroll
      for (int r = 0; r < 9; r++) c.st(L + r, mk(0.0F, 0.0F));
      const View<S> t0 = c.at(x0 - 10);
#pragma unroll
      for (int jj = 0; jj < 32; jj++) {
        const cf v = t0.ld(31 - jj);
#pragma unroll
        for (int q = 0; q < 12; q++) {
          const int k = q - 11 + jj;
          if (k >= 0 && k <= 20) acc[q] = padd(acc[q], pmul0(v, s[k]));
        }
      }
#pragma unroll
      for (int q = 0; q < 12; q++) dl[q] = (x0 + q < L) ? acc[q] : mk(0.0F, 0.0F);
      return;
    }
#endif
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
      const int row = x0 + 21 - jj;
      if ((unsigned)row < (unsigned)L) {
        const cf v = c.ld(row);
#pragma unroll
        for (int q = 0; q < 12; q++) {
          const int k = q - 11 + jj;                     // = (x0 + q + 10) - row
          if (k >= 0 && k <= 20) acc[q] = padd(acc[q], pmul0(v, s[k]));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 12; q++) dl[q] = ((unsigned)(x0 + q) < (unsigned)L) ? acc[q] : mk(0.0F, 0.0F);
  } else {
#pragma unroll
    for (int q = 0; q < 12; q++) dl[q] = ((unsigned)(x0 + q) < (unsigned)L) ? c.ld(x0 + q) : mk(0.0F, 0.0F);
  }
}

// analyzeTrafficBurst (:935-1037) at sps == 1 with requestChannel == true.  win = burst rows 56..91,
// corr = 36 scratch rows.  Same outputs as analyze_traffic<S, true>.
template <int S>
BTS_HD void analyze_corr(const DevTables *__restrict__ T, View<S> win, View<S> corr, int tsc) {
  cf tap[16];
  load_corr_taps(T, tsc, tap);
  corr36<S>(win, corr, tap);
}

// everything of analyzeTrafficBurst behind the peak search: valley RMS, threshold, channel estimate
template <int S, bool PAD = false>
BTS_HD bool analyze_tail(Grid grid, const DevTables *__restrict__ T, View<S> corr, int tsc, float thr, cf amp, float toa,
                         cf *amplitude, float *TOA, cf chan[6], float *chanOff) {
  constexpr int L = 36;
  if ((toa < 0.0F) || (toa > (float)L)) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const int p = (int)rintf(toa);
  float valley = 0.0F;
  int numRms = 0;
#pragma unroll
  for (int i = 2; i <= 5; i++) {
    if (p - i >= 0) { valley = BTS_ADD(valley, cnorm2(corr.ld(p - i))); numRms++; }
    if (p + i < L)  { valley = BTS_ADD(valley, cnorm2(corr.ld(p + i))); numRms++; }
  }
  if (numRms < 2) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const float RMS = (float)((double)BTS_SQRT(BTS_DIV(valley, (float)numRms)) + 0.00001);
  const float peakToMean = BTS_DIV(cabs_(amp), RMS);
  const float mtoa = T->mid_toa[tsc];
  const cf gain = T->mid_gain[tsc];
  amp = cdiv(amp, gain);
  toa = BTS_SUB(BTS_SUB(toa, mtoa), 10.0F);
  *amplitude = amp;
  *TOA = toa;
  if (!(peakToMean > thr)) return false;
  // channel estimate :1005-1031
  const float TOAoffset = BTS_ADD(mtoa, 10.0F);
  const int smin = (int)floorf(BTS_ADD(TOAoffset, -5.0F));          // window i starts at smin + i (mtoa = 8 +- k/512)
  cf dl[12];
  delayed12<S, PAD>(grid, T, corr, L, -toa, smin, dl);
  float maxEnergy = -1.0F;
  int maxI = -1;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const float pos = BTS_ADD(TOAoffset, (float)(i - 5));
    if (BTS_ADD(pos, 6.0F) > (float)L) continue;
    if (pos < 0.0F) continue;
    float energy = 0.0F;
#pragma unroll
    for (int j = 0; j < 6; j++) energy = BTS_ADD(energy, cnorm2(dl[i + j]));
    if ((double)energy > 0.95 * (double)maxEnergy) { maxI = i; maxEnergy = energy; }
  }
  const cf g = cdiv(mk(1.0F, 0.0F), gain);
#pragma unroll
  for (int j = 0; j < 6; j++) {
    cf v = dl[j];
#pragma unroll
    for (int i = 1; i < 7; i++) if (maxI == i) v = dl[i + j];
    chan[j] = cmul(v, g);
  }
  *chanOff = (float)(5 - maxI);
  return true;
}

template <int S>
BTS_HD bool analyze_fast(Grid grid, const DevTables *__restrict__ T, View<S> win, View<S> corr,
                         int tsc, float thr, cf *amplitude, float *TOA, cf chan[6], float *chanOff) {
  analyze_corr<S>(T, win, corr, tsc);
  float toa;
  const cf amp = peak_detect_fast<S>(grid, corr, 36, &toa);
  return analyze_tail<S>(grid, T, corr, tsc, thr, amp, toa, amplitude, TOA, chan, chanOff);
}

// equalizeBurst (:1343-1399) as a streaming pipeline over one lane's column of a ROLLING tile (the burst,
// already scaled by 1/amplitude; rows outside the burst hold zeros).  F = delayVector's fractional FIR output,
// D[n] = F[n - io] the delayed burst, y = feed-forward output, then decision feedback.
// Window Fw[i] = F[q0 + i], q0 = m0 - io for the block of outputs m0..m0+3.  `base` = burst row held in tile row 0.
#ifndef BTS_EQ_ROWS
#define BTS_EQ_ROWS 56
#endif
constexpr int kEqRows = BTS_EQ_ROWS;  // rows of the rolling tile (56 rows = 14.8 KB per warp: 14-15 one-warp CTAs per SM at 128 registers)
#ifndef BTS_EQ_CROT
#define BTS_EQ_CROT 1                // the ring kernels read rot / revrot from a __constant__ copy (the index is warp-uniform: one constant-
                                     // cache broadcast instead of 32 lanes' global loads): equaliser 0.961 -> 0.931 ms; 0 = DevTables in global memory
#endif
#if BTS_EQ_CROT && defined(__CUDACC__)
__constant__ float4 c_eq_rr[160];    // {rot.x, rot.y, revrot.x, revrot.y} per symbol, sps == 1 (uploaded by upload_rach_taps)
#endif
constexpr int kEqRing = 32;          // rows of the ring tile (k_equalize_ring): a step reads 24, four more arrive, four are free
constexpr int kEqStart = -12;        // first pipeline step: three priming steps fill the window
constexpr int kEqLook = 23;          // step(m0) reads burst rows m0 - io .. m0 - io + 23

template <int S>
struct EqLane {
  View<S> a;
  int N, io;
  bool nofrac;
  float s[21];
  cf w[7], b[5], hist[5], Fw[10];
  cf ws[7], bs[5];                                                // (-im, re) copies of the taps for cmac_tap

  BTS_HD void init(Grid grid, const DevTables *__restrict__ T, View<S> tile, int n, float TOA, const cf *w_, const cf *b_) {
    a = tile;
    N = n;
    const float delay = -TOA;                                     // delayVector(rxBurst, -TOA) :1350
    io = (int)floorf(delay);
    const float frac = BTS_SUB(delay, (float)io);
    nofrac = !((double)fabsf(frac) > 1e-2);
    if (nofrac) {
#pragma unroll
      for (int t = 0; t < 21; t++) s[t] = 0.0F;
    } else {
      load_delay_taps(grid, T, frac, s);
    }
#pragma unroll
    for (int k = 0; k < 7; k++) { w[k] = w_[k]; ws[k] = cswapneg(w_[k]); }
#pragma unroll
    for (int k = 0; k < 5; k++) { b[k] = b_[k]; bs[k] = cswapneg(b_[k]); hist[k] = mk(0.0F, 0.0F); }
#pragma unroll
    for (int i = 0; i < 10; i++) Fw[i] = mk(0.0F, 0.0F);
    // the window is primed by running the pipeline from m0 = kEqStart: compute_y(-8) and compute_y(-4) leave
    // F[-io .. -io+5] in Fw[0..5]; their feed-forward outputs (m < 0) are discarded by feedback4
  }

  // F[x0..x0+3].  Rows outside the burst read as zero from the tile (a zero sample adds +-0 where the reference
  // skips the tap).  CHECKED additionally zeroes F where x or x + io falls outside the burst (delayVector's zero
  // fill and the vector ends).
  template <bool CHECKED>
  BTS_HD void newF4(int base, int x0, cf out[4]) const {
    cf acc[4], center[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { acc[r] = mk(0.0F, 0.0F); center[r] = mk(0.0F, 0.0F); }
    const View<S> t = a.at(x0 + 13 - base);
#pragma unroll
    for (int jj = 0; jj < 24; jj++) {
      const cf v = t.ld(-jj);                                     // burst row x0 + 13 - jj
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int k = r - 3 + jj;                                 // = (x0 + r + 10) - row
        if (k >= 0 && k <= 20) acc[r] = padd(acc[r], pmul0(v, s[k]));
        if (k == 10) center[r] = v;                               // row == x0 + r
      }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      cf f = nofrac ? center[r] : acc[r];
      if (CHECKED) {
        const int x = x0 + r;
        if (!((unsigned)x < (unsigned)N && (unsigned)(x + io) < (unsigned)N)) f = mk(0.0F, 0.0F);
      }
      out[r] = f;
    }
  }

  // The same over a RING of kEqRing tile rows indexed by mu = burst row + io -- the lane's own output timeline: block mb
  // (F[x0..x0+3], x0 = mb - io + 6) reads burst rows x0-10 .. x0+13, i.e. mu = mb-4 .. mb+19, which sit in the same ring
  // slots for every lane of the warp whatever its integer delay.  mb is a multiple of 4, so the 24 rows are six groups of
  // four consecutive slots: six wrapped base addresses, compile-time offsets inside a group.
  template <bool CHECKED>
  BTS_HD void newF4_ring(int mb, cf out[4]) const {
    cf acc[4], center[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { acc[r] = mk(0.0F, 0.0F); center[r] = mk(0.0F, 0.0F); }
    View<S> grp[6];
#pragma unroll
    for (int g = 0; g < 6; g++) grp[g] = a.at((mb - 4 + 4 * g) & (kEqRing - 1));
#pragma unroll
    for (int jj = 0; jj < 24; jj++) {
      const int c = 23 - jj;                                       // row mu = mb - 4 + c = burst row x0 + 13 - jj
      const cf v = grp[c >> 2].ld(c & 3);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int k = r - 3 + jj;
        if (k >= 0 && k <= 20) acc[r] = padd(acc[r], pmul0(v, s[k]));
        if (k == 10) center[r] = v;
      }
    }
    const int x0 = mb - io + 6;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      cf f = nofrac ? center[r] : acc[r];
      if (CHECKED) {
        const int x = x0 + r;
        if (!((unsigned)x < (unsigned)N && (unsigned)(x + io) < (unsigned)N)) f = mk(0.0F, 0.0F);
      }
      out[r] = f;
    }
  }

  // step(m0 - 4) may take the unchecked path: every F of block m0 is a sample of the delayed burst, and the block
  // fed back in that step (m0 - 4) has five past decisions (m0 - 4 >= 8 keeps the EDGE feedback for the first blocks)
  BTS_HD bool interior(int m0) const {
    const int x0 = m0 - io + 6;
    return x0 >= 0 && x0 + 3 <= N - 1 && m0 + 9 <= N - 1 && m0 - 4 >= 8;
  }

  // feed-forward outputs y[m0..m0+3] (consumes the window, then slides it by 4)
  template <bool CHECKED, bool RING = false>
  BTS_HD void compute_y(int base, int m0, cf y[4]) {
    if (RING) newF4_ring<CHECKED>(m0, &Fw[6]);
    else newF4<CHECKED>(base, m0 - io + 6, &Fw[6]);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      cf sum = mk(0.0F, 0.0F);
#pragma unroll
      for (int k = 0; k < 7; k++) sum = cmac_tap(sum, w[k], ws[k], Fw[r + 6 - k]);
      y[r] = sum;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) Fw[i] = Fw[i + 4];
  }

  // decision feedback + slicer for m = m0..m0+3 (:1367-1386); rot/revrot = the table entries for those m.
  // EDGE = the block may contain m < 0 (pipeline priming) or m < 5 (fewer than five past decisions exist: the
  // reference's `dBackPtr >= begin` test); every later block runs the branch-free form, which lets the compiler
  // schedule this serial chain in the same basic block as the next block's feed-forward MACs.
  template <bool EDGE>
  BTS_HD void feedback4(int m0, const cf y[4], const cf rot[4], const cf revrot[4], float soft[4]) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int m = m0 + r;
      if (EDGE && m < 0) { soft[r] = 0.0F; continue; }
      cf v = y[r];
#pragma unroll
      for (int k = 0; k < 5; k++)
        if (!EDGE || m - 1 - k >= 0) v = cmac_tap(v, b[k], bs[k], hist[k]);
      const float out = BTS_SUB(BTS_MUL(v.x, revrot[r].x), BTS_MUL(v.y, revrot[r].y));   // real part of v * revrot[m]
#pragma unroll
      for (int k = 4; k >= 1; k--) hist[k] = hist[k - 1];
      hist[0] = cmul(mk((out > 0.0F) ? 1.0F : -1.0F, 0.0F), rot[r]);
      soft[r] = soft_slice(out);
    }
  }

  // one pipeline step: feed-forward for the NEXT block (independent work) + feedback for the current one
  template <bool CHECKED, bool RING = false>
  BTS_HD void step(const DevTables *__restrict__ T, int base, int m0, cf ycur[4], float soft[4]) {
    cf rot[4], revrot[4];                                         // fetched first: off the feedback chain's critical path
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int mi = m0 + r < 0 ? 0 : (m0 + r > 156 ? 156 : m0 + r);
#if BTS_EQ_CROT && defined(__CUDA_ARCH__)
      const float4 q = c_eq_rr[mi];
      rot[r] = mk(q.x, q.y);
      revrot[r] = mk(q.z, q.w);
#else
      rot[r] = T->rot[mi];
      revrot[r] = T->revrot[mi];
#endif
    }
    cf ynext[4];
    compute_y<CHECKED, RING>(base, m0 + 4, ynext);
    if (CHECKED && m0 < 8) feedback4<true>(m0, ycur, rot, revrot, soft);
    else feedback4<false>(m0, ycur, rot, revrot, soft);
#pragma unroll
    for (int r = 0; r < 4; r++) ycur[r] = ynext[r];
  }
};

// warp-uniform rolling-tile policy shared by the kernel and the host emulation: before step(m0) the tile must hold
// burst rows [m0 - io_max, m0 - io_min + kEqLook]; when it does not, re-stage with row (m0 - io_max) at tile row 0.
BTS_HD bool eq_needs_restage(int base, int m0, int io_min, int io_max) {
  return (m0 - io_max < base) || (m0 - io_min + kEqLook >= base + kEqRows);
}

// ---- access bursts (sps == 1) -----------------------------------------------------------------------------------
// correlate(burst, RACH sequence, NO_DELAY) (:867-869): c[n] = sum_{k=0..40} r[n+20-k] * tap[k], tap[k] =
// conj(seq[40-k]), burst indices outside [0, N) not part of the vector; then detectRACHBurst (:860-914).
// (Round 1 wrote the correlation in place over a full-burst tile: 47 KB per warp, 4 warps per SM.)
// The warp walks the correlation in blocks of four lags over a tile of kRachRollRows burst rows (zero outside the burst, so every
// block runs unchecked: a zero sample adds +-0 where the reference skips the tap), keeps the running first-maximum
// of |c|^2, and parks the correlation in a global scratch row per burst (one full 32-byte sector per lane and block).
// Afterwards each lane brings the 26 lags around its own maximum back into shared memory for the early/late search and
// reads the 51 valley lags straight from its scratch row.  21 KB per warp: 10 warps per SM.
constexpr int kRachRollRows = 80;                         // block n0 reads burst rows n0-20 .. n0+23
constexpr int kRachWin = 26;                              // lags imax-12 .. imax+13 cover every read of the peak search
BTS_HD bool rach_needs_restage(int base, int n0) { return n0 + 23 >= base + kRachRollRows; }

template <int S>
BTS_HD void rach_corr4_roll(View<S> tile, int base, const cf *__restrict__ tap, int n0, cf acc[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) acc[r] = mk(0.0F, 0.0F);
  const View<S> t = tile.at(n0 + 23 - base);
#pragma unroll
  for (int j = 0; j < 44; j++) {
    const cf v = t.ld(-j);                                // burst row n0 + 23 - j
    const cf vs = cswapneg(v);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int k = j + r - 3;                            // = (n0 + r + 20) - row
      if (k >= 0 && k <= 40) acc[r] = cmac_tap(acc[r], v, vs, tap[k]);
    }
  }
}
// peakDetect's early/late search (:684-700) around a maximum already found at lag imax; c must serve lags imax-12..imax+13
template <int S>
BTS_HD cf peak_refine_fast(Grid grid, View<S> c, int n, int imax, float *peakIndex) {
  const bool interior = imax >= 12 && imax <= n - 13;     // no tap of any interpolation is clipped (see peak_detect_fast)
#if defined(__CUDA_ARCH__) && !defined(BTS_NO_INTERIOR_SEARCH)
  const bool fast = __all_sync(__activemask(), interior);
#else
  const bool fast = false && interior;
#endif
  int e512 = (imax - 1) * kSincGrid;
  float s[21];
  if (fast) {
    for (int step = kSincGrid / 2; step >= 1; step >>= 1) {
      const int I = e512 >> 9, j = e512 & (kSincGrid - 1);
      load_grid_row(grid, j, s);
      const float e = cnorm2(interp21_interior<S>(s, c, I)), l = cnorm2(interp21_interior<S>(s, c, I + 2));
      if (e < l) e512 += step;
      else if (e > l) e512 -= step;
      else break;
    }
    load_grid_row(grid, e512 & (kSincGrid - 1), s);
    *peakIndex = BTS_ADD((float)e512 * (1.0F / kSincGrid), 1.0F);
    return interp21_interior<S>(s, c, (e512 >> 9) + 1);
  }
  for (int step = kSincGrid / 2; step >= 1; step >>= 1) {
    const int I = e512 >> 9, j = e512 & (kSincGrid - 1);
    load_grid_row(grid, j, s);
    const float e = cnorm2(interp21<S>(s, c, n, I)), l = cnorm2(interp21<S>(s, c, n, I + 2));
    if (e < l) e512 += step;
    else if (e > l) e512 -= step;
    else break;
  }
  load_grid_row(grid, e512 & (kSincGrid - 1), s);
  *peakIndex = BTS_ADD((float)e512 * (1.0F / kSincGrid), 1.0F);
  return interp21<S>(s, c, n, (e512 >> 9) + 1);
}
// everything after the correlation sweep: win = the lane's window column (row k = lag imax-12+k), cs = its scratch row
template <int S>
BTS_HD bool rach_finish(Grid grid, const DevTables *__restrict__ T, View<S> win, const cf *__restrict__ cs, int N, int imax,
                        float thr, cf *amplitude, float *TOA) {
  float toa;
  const cf pk = peak_refine_fast<S>(grid, win.at(-(imax - 12)), N, imax, &toa);
  if ((toa < 0.0F) || (toa > (float)N)) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const int p = (int)rintf(toa);
  float valley = 0.0F, numSamples = 0.0F;
  for (int i = 57; i <= 107; i++) {
    if (p + i >= N) break;
    valley = BTS_ADD(valley, cnorm2(cs[p + i]));
    numSamples = numSamples + 1.0F;
  }
  if (numSamples < 2.0F) { *amplitude = mk(0.0F, 0.0F); *TOA = toa; return false; }
  const float RMS = (float)((double)BTS_SQRT(BTS_DIV(valley, numSamples)) + 0.00001);
  const float peakToMean = BTS_DIV(cabs_(pk), RMS);
  *amplitude = cdiv(pk, T->rach_gain);
  *TOA = BTS_SUB(BTS_SUB(toa, T->rach_toa), 8.0F);
  return peakToMean > thr;
}

// demodulateBurst (:1056-1097) at sps == 1 as a stream over the rolling tile (the burst already scaled by
// 1/channel).  The warp walks the FILTER index x in lockstep (rows x-10..x+13 are then the same for every lane,
// whatever its TOA -- access-burst TOAs spread over the whole slot), and each lane places F[x] at its own output
// m = x + io: soft[m] = slice(Re(revrot[m] * F[m - io])).  Outputs no x reaches are delayVector's zero fill:
// slice(Re(revrot * 0)) = 0.5 exactly, which the caller pre-fills.
template <int S>
struct SlicerLane {
  EqLane<S> f;          // reuses the fractional-delay block filter (newF4) and its tap/offset set-up
  BTS_HD void init(Grid grid, const DevTables *__restrict__ T, View<S> tile, int n, float TOA) {
    cf zw[7], zb[5];
#pragma unroll
    for (int k = 0; k < 7; k++) zw[k] = mk(0.0F, 0.0F);
#pragma unroll
    for (int k = 0; k < 5; k++) zb[k] = mk(0.0F, 0.0F);
    f.init(grid, T, tile, n, TOA, zw, zb);
  }
  // F[x0..x0+3] -> soft[4] for outputs m = x0 + io + r; valid[r] tells which of them exist
  BTS_HD void step(const DevTables *__restrict__ T, int base, int x0, float soft[4], bool valid[4]) {
    cf d[4];
    f.template newF4<false>(base, x0, d);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int x = x0 + r, m = x + f.io;
      valid[r] = (unsigned)x < (unsigned)f.N && (unsigned)m < (unsigned)f.N;
      const cf rr = T->revrot[valid[r] ? m : 0];
      soft[r] = soft_slice(BTS_SUB(BTS_MUL(rr.x, d[r].x), BTS_MUL(rr.y, d[r].y)));     // Re(revrot[m] * D[m]) :232-264
    }
  }
};
// rows step(x0) reads: x0 - 10 .. x0 + 13 (the same for every lane)
BTS_HD bool slicer_needs_restage(int base, int x0) { return x0 + 13 >= base + kEqRows; }

}  // namespace btsdsp
