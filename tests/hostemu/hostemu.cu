// tests/hostemu/hostemu.cu -- TEST INFRASTRUCTURE ONLY.
//
// Replays the product's kernels on a CPU: the same __host__ __device__ functions of
// openbts_ttsou_b200/csrc/sigproc_device.cuh, driven with the same views (stride-33 transposed tiles,
// 32 "lanes" per warp processed one after another) and the same table-construction order as
// capi.cu::build_tables.  It lets the `-m "not gpu"` suite check the kernels' control flow, indexing
// and arithmetic order against the oracle without a GPU.  It is never loaded by the product, is not a
// fallback (libbtsdsp.so has none) and says nothing about device-only behaviour (intrinsic rounding,
// warp staging), which the `-m gpu` tests cover through the C ABI.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../openbts_ttsou_b200/csrc/kernels.cuh"
#include "../../openbts_ttsou_b200/csrc/sigproc_device.cuh"
#include "../../openbts_ttsou_b200/csrc/demod_fast.cuh"
#include "../../openbts_ttsou_b200/csrc/tables_host.h"
#include "../../openbts_ttsou_b200/csrc/fec_lane.cuh"

using namespace btsdsp;

static DevTables *T = nullptr;

static void burst_loc_h(const cf *base, long long pitch, const int *lens, long long first, int sps, long long i,
                        long long *start, int *len) {
  const long long g = first + i;
  const int q = (int)(g & 3);
  const int rule_len = (q == 0 ? 157 : 156) * sps;
  if (pitch > 0) { *start = i * pitch; *len = lens ? lens[i] : rule_len; }
  else {
    const int off = q == 0 ? 0 : (q == 1 ? 157 : (q == 2 ? 313 : 469));
    *start = ((g >> 2) * 625 + off) * sps; *len = rule_len;
  }
}

extern "C" {

int emu_rssi(float a) { return trx_rssi(T, a); }
int emu_rssi_table(float *dst) { memcpy(dst, T->rssi_thr, sizeof T->rssi_thr); return kRssiMin; }

int emu_setup(int sps) {
  if (!T) T = (DevTables *)malloc(sizeof(DevTables));
  host_fill_tables(T, sps);
  // k_init_tables
  float phase = 0.0F;
  const float inc = BTS_DIV(BTS_DIV(kPiF, 2.0F), (float)sps);
  for (int i = 0; i < 157 * sps; i++) {
    T->rot[i] = expj_lookup(T, phase);
    T->revrot[i] = expj_lookup(T, -phase);
    phase = BTS_ADD(phase, inc);
  }
  // k_init_sinc_grid
  for (int j = 0; j < kSincGrid; j++)
    for (int i = 0; i < 24; i++) {
      float v = 0.0F;
      if (i < 21) {
        const float d = BTS_SUB((float)(i - 10), (float)j * (1.0F / (float)kSincGrid));
        v = sinc_exact(T, BTS_MUL(kPiF, d));
      }
      T->sinc_grid[j][i] = v;
    }
  // midambles / RACH as in capi.cu::build_tables
  const cf one = mk(1.0F, 0.0F);
  for (int t = 0; t < 8; t++) {
    uint8_t bits[26];
    for (int i = 0; i < 26; i++) bits[i] = kTSC[t][i] == '1';
    const int nmid = 16 * sps, nfull = 26 * sps;
    std::vector<cf> mid(nmid), full(nfull), ac(nfull);
    for (int i = 0; i < nmid; i++) mid[i] = modulate_at(T, bits + 5, 16, nmid, sps, &one, 1, false, i);
    for (int i = 0; i < nfull; i++) full[i] = modulate_at(T, bits, 26, nfull, sps, T->pulse, T->pulse_len, true, i);
    for (int i = 0; i < nmid; i++) mid[i] = cmul(mid[i], mk(-1.0F, 0.0F));
    for (int i = 0; i < nfull; i++) full[i] = cmul(full[i], mk(0.0F, 1.0F));
    const int start = no_delay_start(nmid);
    for (int i = 0; i < nfull; i++) ac[i] = conv_cc_at<1>(View<1>{full.data()}, nfull, mid.data(), nmid, start + i, true);
    float toa;
    T->mid_gain[t] = peak_detect<1, false>(T, View<1>{ac.data()}, nfull, &toa, nullptr);
    T->mid_toa[t] = toa - (float)(5 * sps);
    memcpy(T->mid_seq[t], mid.data(), nmid * sizeof(cf));
  }
  {
    uint8_t bits[41];
    for (int i = 0; i < 41; i++) bits[i] = kRACH[i] == '1';
    const int n = 41 * sps;
    std::vector<cf> seq(n), ac(n);
    for (int i = 0; i < n; i++) seq[i] = modulate_at(T, bits, 41, n, sps, T->pulse, T->pulse_len, true, i);
    const int start = no_delay_start(n);
    for (int i = 0; i < n; i++) ac[i] = conv_cc_at<1>(View<1>{seq.data()}, n, seq.data(), n, start + i, true);
    float toa;
    T->rach_gain = peak_detect<1, false>(T, View<1>{ac.data()}, n, &toa, nullptr);
    T->rach_toa = toa;
    memcpy(T->rach_seq, seq.data(), n * sizeof(cf));
  }
  return 0;
}

int emu_get_table(int id, int idx, float *dst, int cap) {
  const int sps = T->sps;
  const void *src = nullptr;
  int n = 0;
  float meta[3];
  switch (id) {
    case 0: src = T->cosT; n = kTrig + 1; break;
    case 1: src = T->sinT; n = kTrig + 1; break;
    case 2: src = T->rot; n = 2 * 157 * sps; break;
    case 3: src = T->revrot; n = 2 * 157 * sps; break;
    case 4: src = T->pulse; n = 2 * T->pulse_len; break;
    case 5: src = T->mid_seq[idx]; n = 2 * 16 * sps; break;
    case 6: meta[0] = T->mid_toa[idx]; meta[1] = T->mid_gain[idx].x; meta[2] = T->mid_gain[idx].y; src = meta; n = 3; break;
    case 7: src = T->rach_seq; n = 2 * 41 * sps; break;
    case 8: meta[0] = T->rach_toa; meta[1] = T->rach_gain.x; meta[2] = T->rach_gain.y; src = meta; n = 3; break;
    case 9: src = T->lpf_rx; n = kRxTaps; break;
    case 10: src = T->lpf_tx; n = kTxTaps; break;
    default: return -1;
  }
  if (cap >= n) memcpy(dst, src, n * sizeof(float));
  return n;
}

float emu_sinc(float x) { return sinc_exact(T, x); }
float emu_sin_lookup(float x) { return trig_lookup(T->sinT, x); }
float emu_cos_lookup(float x) { return trig_lookup(T->cosT, x); }

// the sinc grid must equal sinc_exact at the arguments the reference would form
int emu_check_sinc_grid(void) {
  int bad = 0;
  for (int j = 0; j < kSincGrid; j++)
    for (int m = -10; m <= 10; m++) {
      const float frac = (float)j / 512.0F;
      const float a = sinc_exact(T, BTS_MUL(kPiF, BTS_SUB((float)m, frac)));
      uint32_t x, y;
      memcpy(&x, &a, 4); memcpy(&y, &T->sinc_grid[j][m + 10], 4);
      bad += x != y;
    }
  return bad;
}

// k_detect_design + k_equalize_fast (kernels.cu / demod_fast.cuh), warp by warp, lane by lane:
// phase 1 on a 45-row tile (burst samples 56..91 in rows 9..44, the correlation written in place to rows 0..35,
// the energy-gate window staged into rows 0..19 and evaluated first), phase 2 on a 160-row tile holding the detected bursts scaled by 1/amp.
static int g_eq_ring = 1;          // which equaliser kernel's tile policy emu_demod_normal replays: ring (default) or rolling
void emu_set_eq_ring(int on) { g_eq_ring = on; }
void emu_demod_normal(const float *bursts, long long pitch, const int *lens, long long first, const uint8_t *tsc,
                      long long n, float detect_thr, float gate_thr, float snr_thr, int *flag, float *amp, float *toa,
                      float *soft, int soft_pitch, float *chan_o, float *off_o, float *w_o, float *b_o) {
  std::vector<float> gridv(kSincGrid * kGridPitch);
  for (int i = 0; i < kSincGrid * kGridPitch; i++) gridv[i] = T->sinc_grid[i / kGridPitch][i % kGridPitch];
  const Grid gsm{gridv.data(), kGridPitch}, ggl{&T->sinc_grid[0][0], 24};
  std::vector<cf> tileA(45 * kTileStride), tileB(kEqRows * kTileStride);
  cf *A = tileA.data(), *B = tileB.data();
  const bool gated = gate_thr >= 0.0F;
  for (long long w0 = 0; w0 < n; w0 += 32) {
    const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
    for (size_t k = 0; k < tileA.size(); k++) A[k] = mk(1e30F, -1e30F);       // poison: catch reads of unstaged rows
    bool passv[32];
    for (int j = 0; j < 32; j++) passv[j] = true;
    if (gated) {
      for (int j = 0; j < nv; j++) {
        long long start; int len;
        burst_loc_h((const cf *)bursts, pitch, lens, first, 1, w0 + j, &start, &len);
        const cf *g = (const cf *)bursts + start;
        for (int i = 0; i < 20; i++) A[i * kTileStride + j] = g[i];
      }
      for (int j = 0; j < nv; j++) {
        long long start; int len;
        burst_loc_h((const cf *)bursts, pitch, lens, first, 1, w0 + j, &start, &len);
        passv[j] = energy_detect<kTileStride>(View<kTileStride>{A + j}, len, 20, gate_thr, nullptr);
      }
      for (size_t k = 0; k < tileA.size(); k++) A[k] = mk(1e30F, -1e30F);
    }
    for (int j = 0; j < nv; j++) {
      long long start; int len;
      burst_loc_h((const cf *)bursts, pitch, lens, first, 1, w0 + j, &start, &len);
      const cf *g = (const cf *)bursts + start;
      for (int i = 0; i < 36; i++) A[(9 + i) * kTileStride + j] = g[56 + i];
    }
    bool okv[32]; int lenv[32]; long long startv[32]; cf iav[32], wv[32][7], fbv[32][5]; float teq[32];
    for (int lane = 0; lane < nv; lane++) {
      const long long i = w0 + lane;
      long long start; int len;
      burst_loc_h((const cf *)bursts, pitch, lens, first, 1, i, &start, &len);
      const View<kTileStride> a{A + lane};
      cf ampv = mk(0.0F, 0.0F), ia = mk(0.0F, 0.0F), chan[6], w[7], fb[5];
      float tv = 0.0F, off = 0.0F;
      bool pass = passv[lane], ok = false;
      if (pass) ok = analyze_fast<kTileStride>(gsm, T, a.at(9), a, tsc[i], detect_thr, &ampv, &tv, chan, &off);
      if (ok) {
        const float SNR = (float)((double)cnorm2(ampv) / ((double)BTS_MUL(snr_thr, snr_thr) + 1.0));
        ia = cdiv(mk(1.0F, 0.0F), ampv);
        for (int j = 0; j < 6; j++) chan[j] = cmul(chan[j], ia);
        design_dfe<7, 5>(chan, 5, SNR, 7, w, fb);
      }
      flag[i] = ok; amp[2 * i] = ampv.x; amp[2 * i + 1] = ampv.y; toa[i] = tv;
      if (off_o) off_o[i] = ok ? off : 0.0F;
      for (int j = 0; j < 6 && chan_o; j++) ((cf *)chan_o)[i * 6 + j] = ok ? chan[j] : mk(0.0F, 0.0F);
      for (int j = 0; j < 7 && w_o; j++) ((cf *)w_o)[i * 7 + j] = ok ? w[j] : mk(0.0F, 0.0F);
      for (int j = 0; j < 5 && b_o; j++) ((cf *)b_o)[i * 5 + j] = ok ? fb[j] : mk(0.0F, 0.0F);
      if (len > kBurstRows - 3) len = kBurstRows - 3;
      okv[lane] = ok; lenv[lane] = len; startv[lane] = start; iav[lane] = ia; teq[lane] = BTS_SUB(tv, off);
      for (int j = 0; j < 7; j++) wv[lane][j] = w[j];
      for (int j = 0; j < 5; j++) fbv[lane][j] = fb[j];
    }
    // ---- k_equalize_fast: the 32 lanes advance together through the pipeline; the tile rolls when any lane's
    //      look-ahead leaves it (warp-uniform decision from io_min / io_max)
    for (int lane = 0; lane < nv; lane++) {
      float *row = soft + (w0 + lane) * soft_pitch;
      for (int m = 0; m < soft_pitch; m++) row[m] = 0.0F;
    }
    EqLane<kTileStride> eq[32];
    cf ycur[32][4];
    int io_min = 0x7fffffff, io_max = (int)0x80000000, nmax = 0;
    bool any = false;
    for (int lane = 0; lane < nv; lane++) {
      if (!okv[lane]) continue;
      any = true;
      eq[lane].init(ggl, T, View<kTileStride>{B + lane}, lenv[lane], teq[lane], wv[lane], fbv[lane]);
      io_min = eq[lane].io < io_min ? eq[lane].io : io_min;
      io_max = eq[lane].io > io_max ? eq[lane].io : io_max;
      nmax = lenv[lane] > nmax ? lenv[lane] : nmax;
      for (int r = 0; r < 4; r++) ycur[lane][r] = mk(0.0F, 0.0F);
    }
    if (!any) continue;
    if (g_eq_ring) {
      // ---- k_equalize_ring: 32-row ring indexed on each lane's output timeline (slot = (burst row + io) & 31); a group of
      //      four rows is stored at the top of the step that first needs it; slots never stored hold poison
      std::vector<cf> R(kEqRing * kTileStride);
      for (auto &x : R) x = mk(1e30F, -1e30F);
      auto store_group = [&](int mu0) {
        for (int j = 0; j < 32; j++)
          for (int qq = 0; qq < 4; qq++) {
            const int mu = mu0 + qq;
            cf v = mk(0.0F, 0.0F), ia = mk(0.0F, 0.0F);
            if (j < nv && okv[j]) {
              const int r = mu - eq[j].io;
              ia = iav[j];
              if (r >= 0 && r < lenv[j]) v = ((const cf *)bursts + startv[j])[r];
            }
            R[(mu & (kEqRing - 1)) * kTileStride + j] = cmul(v, ia);
          }
      };
      for (int lane = 0; lane < nv; lane++) if (okv[lane]) eq[lane].a = View<kTileStride>{R.data() + lane};
      for (int g = 0; g < 5; g++) store_group(kEqStart + 4 * g);
      for (int m0 = kEqStart; m0 < nmax; m0 += 4) {
        store_group(m0 + 20);
        // the slots of rows below this step's window are the ones the NEXT store overwrites: poison them now, so a read of a
        // row the kernel no longer holds shows up
        for (int j = 0; j < 32; j++) for (int qq = 0; qq < 4; qq++) R[((m0 - 4 + qq) & (kEqRing - 1)) * kTileStride + j] = mk(1e30F, -1e30F);
        bool all_int = true;
        for (int lane = 0; lane < nv; lane++) if (okv[lane]) all_int = all_int && eq[lane].interior(m0 + 4);
        for (int lane = 0; lane < nv; lane++) {
          if (!okv[lane]) continue;
          float s4[4];
          if (all_int) eq[lane].template step<false, true>(T, 0, m0, ycur[lane], s4);
          else eq[lane].template step<true, true>(T, 0, m0, ycur[lane], s4);
          float *row = soft + (w0 + lane) * soft_pitch;
          for (int r = 0; r < 4; r++) if (m0 + r >= 0 && m0 + r < lenv[lane] && m0 + r < soft_pitch) row[m0 + r] = s4[r];
        }
      }
      continue;
    }
    int base = 0;
    bool staged = false;
    for (int m0 = kEqStart; m0 < nmax; m0 += 4) {
      if (!staged || eq_needs_restage(base, m0, io_min, io_max)) {
        base = m0 - io_max;
        staged = true;
        for (size_t k = 0; k < (size_t)kEqRows * kTileStride; k++) B[k] = mk(1e30F, -1e30F);
        for (int j = 0; j < nv; j++) {
          if (!okv[j]) continue;
          const cf *g = (const cf *)bursts + startv[j];
          for (int tr = 0; tr < kEqRows; tr++) {
            const int r = base + tr;
            B[tr * kTileStride + j] = ((unsigned)r < (unsigned)lenv[j]) ? cmul(g[r], iav[j]) : mk(0.0F, 0.0F);
          }
        }
      }
      bool all_int = true;
      for (int lane = 0; lane < nv; lane++) if (okv[lane]) all_int = all_int && eq[lane].interior(m0 + 4);
      for (int lane = 0; lane < nv; lane++) {
        if (!okv[lane]) continue;
        float s4[4];
        if (all_int) eq[lane].template step<false>(T, base, m0, ycur[lane], s4);
        else eq[lane].template step<true>(T, base, m0, ycur[lane], s4);
        float *row = soft + (w0 + lane) * soft_pitch;
        for (int r = 0; r < 4; r++) if (m0 + r >= 0 && m0 + r < lenv[lane] && m0 + r < soft_pitch) row[m0 + r] = s4[r];
      }
    }
  }
}

// tiles != 0: k_rach_detect + k_slicer_fast (sps 1, in-place correlation, rolling-tile slicer), warp by warp;
// tiles == 0: k_rach<false> (any sps, global scratch)

void emu_rach(const float *bursts, long long pitch, const int *lens, long long first, long long n, float detect_thr,
              int sps, int tiles, int *flag, float *amp, float *toa, float *soft, int soft_pitch) {
  if (!tiles) {
    std::vector<cf> scratch(scratch_per_burst(sps));
    for (long long i = 0; i < n; i++) {
      long long start; int len;
      burst_loc_h((const cf *)bursts, pitch, lens, first, sps, i, &start, &len);
      const cf *src = (const cf *)bursts + start;
      cf ampv = mk(0.0F, 0.0F);
      float toav = 0.0F;
      float *sp = soft + i * soft_pitch;
      for (int m = 0; m < soft_pitch; m++) sp[m] = 0.0F;
      cf *s = scratch.data();
      const View<1> corr{s}, x{s + 157 * sps};
      const bool ok = detect_rach<1, true>(T, View<1>{(cf *)src}, len, detect_thr, sps, corr, &ampv, &toav);
      if (ok) {
        for (int m = 0; m < len; m++) x.st(m, src[m]);
        demodulate_burst<1, 1>(T, x, len, sps, ampv, toav, corr, sp);
      }
      flag[i] = ok; amp[2 * i] = ampv.x; amp[2 * i + 1] = ampv.y; toa[i] = toav;
    }
    return;
  }
  const Grid ggl{&T->sinc_grid[0][0], 24};
  cf taps[41];
  for (int k = 0; k < 41; k++) taps[k] = mk(T->rach_seq[40 - k].x, -T->rach_seq[40 - k].y);
  std::vector<cf> tileB(kEqRows * kTileStride), tileR(kRachRollRows * kTileStride), cs(32 * 160);
  cf *B = tileB.data(), *R = tileR.data();
  for (long long w0 = 0; w0 < n; w0 += 32) {
    const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
    long long startv[32]; int lenv[32]; bool okv[32]; cf iav[32]; float toav[32];
    for (int j = 0; j < nv; j++) {
      burst_loc_h((const cf *)bursts, pitch, lens, first, 1, w0 + j, &startv[j], &lenv[j]);
      if (lenv[j] > 157) lenv[j] = 157;
    }
    int nmax = 0;
    for (int j = 0; j < nv; j++) nmax = lenv[j] > nmax ? lenv[j] : nmax;
    // the rolling-tile detector (k_rach_detect): correlation sweep with the running maximum, the correlation parked
    // in a scratch row per burst, then the peak search on a 26-lag window
    int imaxv[32]; float maxv[32];
    {
      for (int j = 0; j < 32; j++) { imaxv[j] = -1; maxv[j] = 0.0F; }
      for (size_t k = 0; k < cs.size(); k++) cs[k] = mk(1e30F, -1e30F);
      int base = 0;
      bool staged = false;
      for (int n0 = 0; n0 < nmax; n0 += 4) {
        if (!staged || rach_needs_restage(base, n0)) {
          base = n0 - 20;
          staged = true;
          for (size_t k = 0; k < tileR.size(); k++) R[k] = mk(1e30F, -1e30F);
          for (int j = 0; j < nv; j++) {
            const cf *g = (const cf *)bursts + startv[j];
            for (int tr = 0; tr < kRachRollRows; tr++) {
              const int r = base + tr;
              R[tr * kTileStride + j] = (r >= 0 && r < lenv[j]) ? g[r] : mk(0.0F, 0.0F);
            }
          }
        }
        for (int lane = 0; lane < nv; lane++) {
          cf acc[4];
          rach_corr4_roll<kTileStride>(View<kTileStride>{R + lane}, base, taps, n0, acc);
          for (int r = 0; r < 4; r++) {
            if (n0 + r >= lenv[lane]) continue;
            cs[lane * 160 + n0 + r] = acc[r];
            const float p = cnorm2(acc[r]);
            if (p > maxv[lane]) { maxv[lane] = p; imaxv[lane] = n0 + r; }
          }
        }
      }
      for (size_t k = 0; k < tileR.size(); k++) R[k] = mk(1e30F, -1e30F);
      for (int lane = 0; lane < nv; lane++)
        for (int k = 0; k < kRachWin; k++) {
          const int idx = imaxv[lane] - 12 + k;
          R[k * kTileStride + lane] = (idx >= 0 && idx < lenv[lane]) ? cs[lane * 160 + idx] : mk(0.0F, 0.0F);
        }
    }
    for (int lane = 0; lane < nv; lane++) {
      const long long i = w0 + lane;
      cf ampv = mk(0.0F, 0.0F);
      float tv = 0.0F;
      okv[lane] = rach_finish<kTileStride>(ggl, T, View<kTileStride>{R + lane}, cs.data() + lane * 160, lenv[lane], imaxv[lane],
                                           detect_thr, &ampv, &tv);
      flag[i] = okv[lane]; amp[2 * i] = ampv.x; amp[2 * i + 1] = ampv.y; toa[i] = tv;
      iav[lane] = okv[lane] ? cdiv(mk(1.0F, 0.0F), ampv) : mk(0.0F, 0.0F);
      toav[lane] = tv;
      float *row = soft + i * soft_pitch;
      for (int m = 0; m < soft_pitch; m++) row[m] = 0.0F;
    }
    // slicer: the warp walks the filter index x0; each lane writes at its own m = x + io
    SlicerLane<kTileStride> sl[32];
    int nm = 0;
    bool any = false;
    for (int lane = 0; lane < nv; lane++) {
      if (!okv[lane]) continue;
      any = true;
      sl[lane].init(ggl, T, View<kTileStride>{B + lane}, lenv[lane], toav[lane]);
      nm = lenv[lane] > nm ? lenv[lane] : nm;
      float *row = soft + (w0 + lane) * soft_pitch;
      for (int m = 0; m < soft_pitch; m++) row[m] = m < lenv[lane] ? 0.5F : 0.0F;
    }
    if (!any) continue;
    if (g_eq_ring) {
      // ---- k_slicer_ring: the ring tile on each lane's output timeline, blocks m0 = -2, 2, 6, ...; dead rows poisoned
      std::vector<cf> Rg(kEqRing * kTileStride);
      for (auto &x : Rg) x = mk(1e30F, -1e30F);
      auto store_group = [&](int mu0) {
        for (int j = 0; j < 32; j++)
          for (int qq = 0; qq < 4; qq++) {
            const int mu = mu0 + qq;
            cf v = mk(0.0F, 0.0F), ia = mk(0.0F, 0.0F);
            if (j < nv && okv[j]) {
              const int r = mu - sl[j].f.io;
              ia = iav[j];
              if (r >= 0 && r < lenv[j]) v = ((const cf *)bursts + startv[j])[r];
            }
            Rg[(mu & (kEqRing - 1)) * kTileStride + j] = cmul(v, ia);
          }
      };
      for (int lane = 0; lane < nv; lane++) if (okv[lane]) sl[lane].f.a = View<kTileStride>{Rg.data() + lane};
      const int kM0 = -2;
      for (int g = 0; g < 5; g++) store_group(kM0 - 10 + 4 * g);
      for (int m0 = kM0; m0 < nm; m0 += 4) {
        store_group(m0 + 10);
        for (int j = 0; j < 32; j++) for (int qq = 0; qq < 4; qq++) Rg[((m0 - 14 + qq) & (kEqRing - 1)) * kTileStride + j] = mk(1e30F, -1e30F);
        for (int lane = 0; lane < nv; lane++) {
          if (!okv[lane]) continue;
          cf d[4];
          sl[lane].f.template newF4_ring<true>(m0 - 6, d);
          float *row = soft + (w0 + lane) * soft_pitch;
          for (int r = 0; r < 4; r++) {
            const int m = m0 + r;
            if (m < 0 || m >= lenv[lane] || m >= soft_pitch) continue;
            const cf rr = T->revrot[m > 156 ? 156 : m];
            row[m] = soft_slice(BTS_SUB(BTS_MUL(rr.x, d[r].x), BTS_MUL(rr.y, d[r].y)));
          }
        }
      }
      continue;
    }
    int base = 0;
    bool staged = false;
    for (int x0 = 0; x0 < nm; x0 += 4) {
      if (!staged || slicer_needs_restage(base, x0)) {
        base = x0 - 10;
        staged = true;
        for (size_t k = 0; k < tileB.size(); k++) B[k] = mk(1e30F, -1e30F);
        for (int j = 0; j < nv; j++) {
          if (!okv[j]) continue;
          const cf *g = (const cf *)bursts + startv[j];
          for (int tr = 0; tr < kEqRows; tr++) {
            const int r = base + tr;
            B[tr * kTileStride + j] = cmul(((unsigned)r < (unsigned)lenv[j]) ? g[r] : mk(0.0F, 0.0F), iav[j]);
          }
        }
      }
      for (int lane = 0; lane < nv; lane++) {
        if (!okv[lane]) continue;
        float s4[4];
        bool valid[4];
        sl[lane].step(T, base, x0, s4, valid);
        float *row = soft + (w0 + lane) * soft_pitch;
        for (int r = 0; r < 4; r++) {
          const int m = x0 + r + sl[lane].f.io;
          if (valid[r] && m < soft_pitch) row[m] = s4[r];
        }
      }
    }
  }
}

// k_analyze<false>: any sps
void emu_analyze(const float *bursts, long long pitch, const int *lens, long long first, const uint8_t *tsc, long long n,
                 float detect_thr, int request, int sps, int *flag, float *amp, float *toa, float *chan_o, float *off_o) {
  std::vector<cf> scratch(scratch_per_burst(sps));
  for (long long i = 0; i < n; i++) {
    long long start; int len;
    burst_loc_h((const cf *)bursts, pitch, lens, first, sps, i, &start, &len);
    cf ampv = mk(0.0F, 0.0F), chan[6 * kMaxSps];
    float toav = 0.0F, off = 0.0F;
    cf *s = scratch.data();
    bool ok = analyze_traffic<1, true>(T, View<1>{(cf *)bursts + start}, tsc[i], detect_thr, sps, View<1>{s},
                                       View<1>{s + 36 * sps}, &ampv, &toav, request != 0, chan, &off);
    flag[i] = ok; amp[2 * i] = ampv.x; amp[2 * i + 1] = ampv.y; toa[i] = toav;
    const bool have = ok && request;
    if (off_o) off_o[i] = have ? off : 0.0F;
    for (int j = 0; j < 6 * sps && chan_o; j++) ((cf *)chan_o)[i * 6 * sps + j] = have ? chan[j] : mk(0.0F, 0.0F);
  }
}

// generic single-vector pieces
void emu_peak_detect(const float *v, int n, float *peak, float *idx, float *avg) {
  cf p = peak_detect<1, false>(T, View<1>{(cf *)v}, n, idx, avg);
  peak[0] = p.x; peak[1] = p.y;
}
void emu_peak_detect_grid(const float *v, int n, float *peak, float *idx, float *avg) {
  cf p = peak_detect<1, true>(T, View<1>{(cf *)v}, n, idx, avg);
  peak[0] = p.x; peak[1] = p.y;
}
void emu_delay_vector(float *v, int n, float delay) {
  std::vector<cf> tmp(n);
  delay_vector<1>(T, View<1>{(cf *)v}, n, delay, View<1>{tmp.data()});
}
void emu_design_dfe(const float *chan, int nchan, float snr, int nf, float *w, float *b, int fixed) {
  cf ch[kDfeMax], W[kDfeMax], F[kDfeMax];
  for (int i = 0; i < nchan; i++) ch[i] = ((const cf *)chan)[i];
  if (fixed) design_dfe<7, 5>(ch, 5, snr, 7, W, F);
  else design_dfe<0, 0>(ch, nchan - 1, snr, nf, W, F);
  for (int i = 0; i < nf; i++) ((cf *)w)[i] = W[i];
  for (int i = 0; i < nchan - 1; i++) ((cf *)b)[i] = F[i];
}
void emu_equalize(float *burst, int n, float toa, const float *w, int nw, const float *b, int nb, float *soft) {
  std::vector<cf> tmp(n + kDfeMax);
  equalize_burst<1, 1>(T, View<1>{(cf *)burst}, n, toa, (const cf *)w, nw, (const cf *)b, nb, View<1>{tmp.data()}, soft);
}
int emu_modulate(const uint8_t *bits, int nbits, int guard, float *out) {
  const int sps = T->sps, n = sps * (nbits + guard);
  for (int t = 0; t < n; t++) ((cf *)out)[t] = modulate_at(T, bits, nbits, n, sps, T->pulse, T->pulse_len, true, t);
  return n;
}

// k_resample_rx / k_resample_tx chunk loops
void emu_resample_rx(const float *in, int has_history, long long nchunks, float *out) {
  std::vector<cf> x(192 + 864);
  std::vector<float> hp(kRxPoly * (kRxP + 1));
  for (int i = 0; i < kRxPoly * kRxP; i++) hp[(i / kRxP) * (kRxP + 1) + i % kRxP] = T->rx_poly[i % kRxP][i / kRxP];
  for (long long c = 0; c < nchunks; c++) {
    const cf *src = (const cf *)in + c * 864;
    for (int i = 0; i < 192 + 864; i++) x[i] = (i >= 192 || has_history || c > 0) ? src[i - 192] : mk(0.0F, 0.0F);
    for (int m = 0; m < 585; m++)
      ((cf *)out)[c * 585 + m] = resample_at<kRxP, kRxQ, kRxTaps, kRxPoly, kRxP + 1>(x.data(), 192 + 864, hp.data(), 130, m);
  }
}
// k_resample_rx_v2: 32-period tiles, padded rows, all 65 phases per "lane"
void emu_resample_rx_v2(const float *in_, int has_history, long long nchunks, float *out_) {
  const cf *in = (const cf *)in_;
  cf *out = (cf *)out_;
  std::vector<float> taps(kRxP * 16);
  rx_fill_taps(T, taps.data());
  const long long nperiods = nchunks * 9, nsamples = nchunks * 864;
  std::vector<cf> xt(kRxTileIn), ot(kRxTileOut);
  for (long long G0 = 0; G0 < nperiods; G0 += 32) {
    const long long raw0 = 96 * G0 - 96, lo = has_history ? -192 : 0;
    for (size_t k = 0; k < xt.size(); k++) xt[k] = mk(1e30F, 1e30F);
    for (int row = 0; row < kRxTileRows; row++)
      for (int c = 0; c < 96; c++) {
        const long long s = raw0 + (long long)row * 96 + c;
        xt[row * kRxRowPitch + c] = (s >= lo && s < nsamples) ? in[s] : mk(0.0F, 0.0F);
      }
    for (int lane = 0; lane < 32; lane++) {
      const long long G = G0 + lane;
      for (int part = 0; part < 8; part++)      // the kernel's split of the 65 phases over eight warps
        rx_part<8>(part, taps.data(), xt.data() + lane * kRxRowPitch, ot.data() + lane * kRxP, (G % 9) == 8);
    }
    const long long nvalid = (nperiods - G0 < 32 ? nperiods - G0 : 32) * kRxP;
    for (long long i = 0; i < nvalid; i++) out[G0 * kRxP + i] = ot[i];
  }
}
void emu_resample_tx(const float *in, int has_history, long long nchunks, short *out) {
  std::vector<cf> x(130 + 585);
  std::vector<float> hp(kTxPoly * (kTxP + 1));
  for (int i = 0; i < kTxPoly * kTxP; i++) hp[(i / kTxP) * (kTxP + 1) + i % kTxP] = T->tx_poly[i % kTxP][i / kTxP];
  for (long long c = 0; c < nchunks; c++) {
    const cf *src = (const cf *)in + c * 585;
    for (int i = 0; i < 130 + 585; i++) x[i] = (i >= 130 || has_history || c > 0) ? src[i - 130] : mk(0.0F, 0.0F);
    for (int m = 0; m < 864; m++) {
      short2 o = tx_quantise(resample_at<kTxP, kTxQ, kTxTaps, kTxPoly, kTxP + 1>(x.data(), 130 + 585, hp.data(), 192, m));
      out[(c * 864 + m) * 2] = o.x; out[(c * 864 + m) * 2 + 1] = o.y;
    }
  }
}

// The caller-policy pipeline (trx_policy.cuh / trx_kernels.cuh) replayed on the CPU: pass 1 and pass 3 with the generic
// single-burst device functions, pass 2 with the SAME trx_policy_arfcn the policy kernel runs.
// state: narfcn TrxState records (in/out).  bursts laid out [frame][arfcn][tn] at `pitch`.
int emu_trx_state_bytes(void) { return (int)sizeof(TrxState); }
void emu_trx_init(void *state, int narfcn, const uint8_t *tsc, const uint8_t *chan_type, int start_fn) {
  TrxState *st = (TrxState *)state;
  memset(st, 0, sizeof(TrxState) * narfcn);
  for (int a = 0; a < narfcn; a++) {
    st[a].thr = 250.0; st[a].prev_false_fn = start_fn; st[a].tsc = tsc[a];
    for (int tn = 0; tn < 8; tn++) { st[a].chan_type[tn] = chan_type[a * 8 + tn]; st[a].est_fn[tn] = start_fn; }
  }
}
static void emu_trx_pull_impl(void *state, int narfcn, const float *bursts, long long pitch, int nframes, int fn0, int *valid,
                              unsigned char *dgram, int dgram_pitch, bool v52m, unsigned max_delay) {
  TrxState *st = (TrxState *)state;
  const bool need_dfe = !v52m || max_delay > 1;
  const long long n = (long long)nframes * narfcn * 8;
  std::vector<DetRec> det(n);
  std::vector<int> act(n), slot(n, -1), rflag, ridx;
  std::vector<double> thr_at(n, 0.0);
  std::vector<float> rtoa;
  std::vector<cf> ramp;
  std::vector<cf> scratch(scratch_per_burst(1));
  // pass 1
  for (long long i = 0; i < n; i++) {
    const int f = (int)(i / (8LL * narfcn)), a = (int)((i >> 3) % narfcn), tn = (int)(i & 7);
    const int fn = (fn0 + f) % kHyperframe, len = (tn % 4 == 0) ? 157 : 156;
    const int corr = expected_corr_type(st[a].chan_type[tn], fn);
    cf *src = (cf *)bursts + i * pitch;
    DetRec d;
    memset(&d, 0, sizeof d);
    if (v52m) energy_detect_52m<1>(View<1>{src}, len, 20, 0.0F, &d.energy);
    else energy_detect<1>(View<1>{src}, len, 20, 0.0F, &d.energy);
    cf amp = mk(0.0F, 0.0F);
    float toa = 0.0F, off = 0.0F;
    if (corr == CORR_TSC) {
      cf chan[6];
      for (int j = 0; j < 6; j++) chan[j] = mk(0.0F, 0.0F);
      cf *s = scratch.data();
      bool ok;
      if (v52m) {
        std::vector<cf> s52(2 * 130);
        ok = analyze_traffic_52m<1>(T, View<1>{src}, st[a].tsc, 3.0F, 1, max_delay, View<1>{s52.data()}, View<1>{s52.data() + 130},
                                    &amp, &toa, need_dfe, chan, &off);
        if (!need_dfe) off = 0.0F;
      } else {
        ok = analyze_traffic<1, true>(T, View<1>{src}, st[a].tsc, 3.0F, 1, View<1>{s}, View<1>{s + 36}, &amp, &toa, true, chan, &off);
      }
      d.flag = ok ? 1.0F : 0.0F; d.amp_x = amp.x; d.amp_y = amp.y; d.toa = toa; d.off = ok ? off : 0.0F;
      for (int j = 0; j < 6; j++) d.chan[j] = (ok && need_dfe) ? chan[j] : mk(0.0F, 0.0F);
    } else if (corr == CORR_RACH) {
      cf *s = scratch.data();
      const bool ok = detect_rach<1, true>(T, View<1>{src}, len, 5.0F, 1, View<1>{s}, &amp, &toa);
      slot[i] = (int)ridx.size();
      ridx.push_back((int)i); rflag.push_back(ok); ramp.push_back(amp); rtoa.push_back(toa);
    }
    det[i] = d;
  }
  if (rflag.empty()) { rflag.push_back(0); }
  // pass 2
  std::vector<int> commit(narfcn * 8);
  for (int a = 0; a < narfcn; a++) {
    TrxScalars sc;
    trx_load_scalars(st[a], sc);
    trx_policy_arfcn(sc, nframes, fn0, narfcn, a, det.data(), slot.data(), rflag.data(), T->exp_neg, act.data(), thr_at.data(),
                     commit.data() + a * 8, need_dfe);
    trx_store_scalars(st[a], sc);
  }
  // pass 3
  struct Dfe { cf w[7], b[5]; float off; };
  std::vector<Dfe> dfe(n);
  for (long long i = 0; i < n; i++) {
    if (act[i] != (int)i) continue;
    const cf ia = cdiv(mk(1.0F, 0.0F), mk(det[i].amp_x, det[i].amp_y));
    cf ch[6];
    for (int j = 0; j < 6; j++) ch[j] = cmul(det[i].chan[j], ia);
    design_dfe<7, 5>(ch, 5, trx_snr_estimate(mk(det[i].amp_x, det[i].amp_y), thr_at[i]), 7, dfe[i].w, dfe[i].b);
    dfe[i].off = det[i].off;
  }
  for (long long i = 0; i < n; i++) {
    const int f = (int)(i / (8LL * narfcn)), a = (int)((i >> 3) % narfcn), tn = (int)(i & 7);
    const int fn = (fn0 + f) % kHyperframe, len = (tn % 4 == 0) ? 157 : 156;
    unsigned char *dg = dgram + i * (long long)dgram_pitch;
    memset(dg, 0, 158);
    valid[i] = 0;
    if (act[i] == ACT_NONE) continue;
    cf *src = (cf *)bursts + i * pitch;
    std::vector<cf> x(160), tmp(200);
    float soft[160];
    memset(soft, 0, sizeof soft);
    cf amp;
    float toa;
    if (act[i] == ACT_RACH || act[i] == ACT_SLICE) {
      if (act[i] == ACT_RACH) { amp = ramp[slot[i]]; toa = rtoa[slot[i]]; }
      else { amp = mk(det[i].amp_x, det[i].amp_y); toa = det[i].toa; }
      for (int m = 0; m < len; m++) x[m] = src[m];
      demodulate_burst<1, 1>(T, View<1>{x.data()}, len, 1, amp, toa, View<1>{tmp.data()}, soft);
    } else {
      amp = mk(det[i].amp_x, det[i].amp_y); toa = det[i].toa;
      const cf ia = cdiv(mk(1.0F, 0.0F), amp);
      for (int m = 0; m < len; m++) x[m] = cmul(src[m], ia);
      const cf *w, *b;
      float off;
      if (act[i] >= 0) { w = dfe[act[i]].w; b = dfe[act[i]].b; off = dfe[act[i]].off; }
      else { w = st[a].w[tn]; b = st[a].b[tn]; off = st[a].chan_off[tn]; }   // carried over: still the pre-commit values
      cf W[7], B[5];
      for (int j = 0; j < 7; j++) W[j] = w[j];
      for (int j = 0; j < 5; j++) B[j] = b[j];
      equalize_burst<1, 1>(T, View<1>{x.data()}, len, BTS_SUB(toa, off), W, 7, B, 5, View<1>{tmp.data()}, soft);
    }
    trx_datagram_header(T, dg, tn, fn, amp, toa, 1);
    for (int m = 0; m < 148; m++) dg[8 + m] = trx_soft_byte(soft[m]);
    valid[i] = 1;
  }
  for (int k = 0; k < narfcn * 8; k++) {
    if (commit[k] < 0) continue;
    const int a = k >> 3, tn = k & 7;
    for (int j = 0; j < 7; j++) st[a].w[tn][j] = dfe[commit[k]].w[j];
    for (int j = 0; j < 5; j++) st[a].b[tn][j] = dfe[commit[k]].b[j];
    st[a].chan_off[tn] = dfe[commit[k]].off;
  }
}

// k_tch_decode's arithmetic, block by block (fec.cuh: tch_decode_block_seq)
void emu_tch_decode(const unsigned char *soft, int burst_pitch, long long nblocks, unsigned char *d, int *good, int *stolen,
                    unsigned char *fu, int *fok) {
  for (long long q = 0; q < nblocks; q++)
    stolen[q] = tch_decode_block_seq(soft + q * 4 * (long long)burst_pitch, burst_pitch, d + 260 * q, good + q, fu + 228 * q, fok + q) ? 1 : 0;
}

void emu_trx_pull(void *state, int narfcn, const float *bursts, long long pitch, int nframes, int fn0, int *valid,
                  unsigned char *dgram, int dgram_pitch) {
  emu_trx_pull_impl(state, narfcn, bursts, pitch, nframes, fn0, valid, dgram, dgram_pitch, false, 0);
}
// the second transceiver variant's policy (Transceiver52M/Transceiver.cpp:268-404)
void emu_trx_pull_52m(void *state, int narfcn, const float *bursts, long long pitch, int nframes, int fn0, int max_delay,
                      int *valid, unsigned char *dgram, int dgram_pitch) {
  emu_trx_pull_impl(state, narfcn, bursts, pitch, nframes, fn0, valid, dgram, dgram_pitch, true, (unsigned)max_delay);
}

// k_tx_fused replayed on the CPU: modulate into the step's tile, 96 phases per period through tx_part, quantise
void emu_tx_fused(const uint8_t *bits, const float *scale, long long nslots, short *out) {
  constexpr int PER = 96, IN = kTxQ * PER + 2 * kTxHalo, PITCH = kTxP + 1;
  const long long nsamples = nslots / 4 * 625, nchunks = nsamples / 585, nperiods = nchunks * 9;
  std::vector<float> taps(kTxP * 8);
  for (int r = 0; r < kTxP; r++)
    for (int k = 0; k < 8; k++) taps[r * 8 + k] = T->tx_poly[tx_br(r)][k];
  std::vector<cf> xs(IN), q(148 * 3);
  std::vector<short2> os(PER * PITCH);
  for (int i = 0; i < 148 * 3; i++) tx_fill_q(T, q.data(), i);
  for (long long G0 = 0; G0 < nperiods; G0 += PER) {
    const long long s0 = (long long)kTxQ * G0 - kTxHalo;
    const long long sa = s0 < 0 ? 0 : s0, ga4 = sa / 625;
    const int w0 = (int)(s0 - ga4 * 625);
    for (int j = 0; j < IN; j++) {
      const int w = w0 + j;
      cf x = mk(0.0F, 0.0F);
      if (w >= 0 && s0 + j < nsamples) {
        const int q4 = w / 625;
        int sl, t;
        tx_slot_of(w - q4 * 625, &sl, &t);
        const long long g = (ga4 + q4) * 4 + sl;
        x = tx_burst_sample(q.data(), bits + g * 148, t);
        if (scale) x = cmul(x, mk(scale[g], 0.0F));
      }
      xs[j] = x;
    }
    for (int row = 0; row < PER; row++)
      for (int part = 0; part < 8; part++)
        tx_part<8>(part, taps.data(), xs.data() + kTxQ * row + kTxHalo, os.data() + row * PITCH, ((G0 + row) % 9) == 8);
    const int nper = (int)(nperiods - G0 < PER ? nperiods - G0 : PER);
    for (int p = 0; p < nper; p++)
      for (int w = 0; w < kTxP; w++) {
        out[((G0 + p) * kTxP + w) * 2] = os[p * PITCH + w].x;
        out[((G0 + p) * kTxP + w) * 2 + 1] = os[p * PITCH + w].y;
      }
  }
}

// XCCH block decoder (fec.cuh), sequential form
void emu_xcch_decode(const unsigned char *soft, int burst_pitch, long long nframes, unsigned char *u, int *ok) {
  for (long long f = 0; f < nframes; f++)
    ok[f] = xcch_decode_frame_seq(soft + f * 4 * (long long)burst_pitch, burst_pitch, u + f * kXcchU) ? 1 : 0;
}

// L1 encoders on the transmit side (fec.cuh, sequential forms); tsc < 0: no midamble
static unsigned emu_tsc_word(int tsc) {
  unsigned w = 0;
  if (tsc >= 0 && tsc < 8) for (int i = 0; i < 26; i++) w |= (unsigned)(kTSC[tsc][i] == '1') << (25 - i);
  return w;
}
void emu_xcch_encode(const unsigned char *frames, long long nframes, int lsb8msb, int tsc, unsigned char *bursts, int burst_pitch) {
  for (long long f = 0; f < nframes; f++)
    xcch_encode_frame_seq(frames + f * 184, lsb8msb, emu_tsc_word(tsc), tsc >= 0, bursts + f * 4 * (long long)burst_pitch, burst_pitch);
}
// the lane form of the encoder kernels (fec_lane.cuh): bit-packed words, compile-time positions; pitch 148 only
void emu_xcch_encode_lanes(const unsigned char *frames, long long nframes, int lsb8msb, int tsc, unsigned char *bursts) {
  static const CrcTable crc = make_fire_crc_table();
  const unsigned sp = enc_sp_base(emu_tsc_word(tsc), tsc >= 0);
  for (long long f = 0; f < nframes; f++) xcch_encode_frame_lane(frames + f * 184, lsb8msb, crc.t, sp, bursts + f * 592);
}
void emu_xcch_encode_lanes_popc(const unsigned char *frames, long long nframes, int lsb8msb, int tsc, unsigned char *bursts) {
  const unsigned sp = enc_sp_base(emu_tsc_word(tsc), tsc >= 0);
  for (long long f = 0; f < nframes; f++) {
    unsigned W[8], pl[16];
    xcch_u_words_popc(frames + f * 184, lsb8msb, W);
    conv_planes(W, pl, pl + 8);
    enc_out_group<XcchTab>(pl, sp | (1u << 2) | (1u << 3), bursts + f * 592);
  }
}
// traffic channel, lane form: groups 1 .. nblocks (group 0 needs the carry: the warp-per-block kernel's job); bursts = 4*nblocks+4 rows of 148
void emu_tch_encode_lanes(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks, int lsb8msb,
                          int tsc, unsigned char *bursts) {
  static const CrcTable crc = make_fire_crc_table();
  const unsigned sp = enc_sp_base(emu_tsc_word(tsc), tsc >= 0);
  for (long long g = 1; g <= nblocks; g++) {
    unsigned pl[32];
    for (int i = 0; i < 32; i++) pl[i] = 0;
    tch_block_planes(steal[g - 1], d260 + (g - 1) * 260, f184 + (g - 1) * 184, lsb8msb, crc.t, pl, pl + 8);
    if (g < nblocks) tch_block_planes(steal[g], d260 + g * 260, f184 + g * 184, lsb8msb, crc.t, pl + 16, pl + 24);
    tch_encode_group_lane(pl, steal[g - 1], g < nblocks ? steal[g] : 0, sp, bursts + g * 592);
  }
}
void emu_tch_encode(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks, int lsb8msb,
                    int tsc, const unsigned char *carry, unsigned char *bursts, int burst_pitch) {
  tch_encode_stream_seq(d260, f184, steal, nblocks, lsb8msb, emu_tsc_word(tsc), tsc >= 0, carry, bursts, burst_pitch);
}

void emu_rach_decode(const unsigned char *soft, int burst_pitch, long long n, unsigned char *u, int *tail, int *bsic, int *ra) {
  for (long long i = 0; i < n; i++) rach_decode_burst_seq(soft + i * (long long)burst_pitch, u + i * kRachU, tail + i, bsic + i, ra + i);
}

// Transceiver52M's analyzeTrafficBurst / energyDetect (the functions that differ from the main variant)
int emu_analyze_52m(const float *burst, int n, int tsc, float thr, unsigned max_toa, int request, float *amp, float *toa,
                    float *chan, float *off) {
  std::vector<cf> scratch(512);
  cf a = mk(0.0F, 0.0F), ch[6 * kMaxSps];
  for (int j = 0; j < 6 * kMaxSps; j++) ch[j] = mk(0.0F, 0.0F);
  float t = 0.0F, o = 0.0F;
  (void)n;
  const bool ok = analyze_traffic_52m<1>(T, View<1>{(cf *)burst}, tsc, thr, T->sps, max_toa, View<1>{scratch.data()},
                                         View<1>{scratch.data() + 256}, &a, &t, request != 0, ch, &o);
  amp[0] = a.x; amp[1] = a.y; *toa = t;
  if (ok && request) { memcpy(chan, ch, sizeof(cf) * 6 * T->sps); *off = o; }
  return ok ? 1 : 0;
}
// k_detect_52m's tile: only the search window is staged (tile rows 0 .. windowLen-1 of the lane's column), the correlation and
// delayVector's temporary live in the rows behind it, the burst view is offset so that row startIx is tile row 0; every other
// row of the tile holds poison
int emu_analyze_52m_tiled(const float *burst, int n, int tsc, float thr, unsigned max_toa, int request, float *amp, float *toa,
                          float *chan, float *off) {
  const Geo52 g = geo_52m(max_toa);
  const int rows = g.windowLen + 2 * g.corrLen + 1, lane = 5;
  std::vector<cf> tile((size_t)(rows + 8) * kTileStride, mk(1e30F, -1e30F));
  cf *A = tile.data() + 4 * kTileStride;                                 // poisoned margin rows before and after
  for (int r = 0; r < g.windowLen; r++) A[r * kTileStride + lane] = (g.startIx + r < n) ? ((const cf *)burst)[g.startIx + r] : mk(0.0F, 0.0F);
  const View<kTileStride> win{A + lane};
  cf a = mk(0.0F, 0.0F), ch[6];
  for (int j = 0; j < 6; j++) ch[j] = mk(0.0F, 0.0F);
  float t = 0.0F, o = 0.0F;
  const bool ok = analyze_traffic_52m<kTileStride>(T, win.at(-g.startIx), tsc, thr, 1, max_toa, win.at(g.windowLen),
                                                   win.at(g.windowLen + g.corrLen), &a, &t, request != 0, ch, &o);
  amp[0] = a.x; amp[1] = a.y; *toa = t;
  if (ok && request) { memcpy(chan, ch, sizeof(cf) * 6); *off = o; }
  return ok ? 1 : 0;
}
int emu_energy_detect_52m(const float *v, int n, unsigned win, float thr, float *avg) {
  return energy_detect_52m<1>(View<1>{(cf *)v}, n, win, thr, avg) ? 1 : 0;
}

}  // extern "C"
