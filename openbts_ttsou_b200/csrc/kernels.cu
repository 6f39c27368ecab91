// kernels.cu -- sm_100a kernels of the burst-DSP path (everything but the resamplers).
//
// Layout in HBM: bursts are interleaved complex-float I/Q (float2), either pitched (one burst every
// `pitch` samples) or cut on the fly from a continuous 157/156/156/156 slot stream (BurstSrc).
// The batched receive kernels run ONE BURST PER THREAD: a warp stages its 32 bursts with coalesced
// loads into a transposed shared-memory tile (row = sample index, column = lane, row stride 33 so
// both the transposing store and the per-lane column walk are bank-conflict free), then every lane
// runs the reference's scalar algorithm over its own column (sigproc_device.cuh).  That keeps each
// comparison bit-identical to the reference while all 32 lanes of every instruction do useful work;
// throughput comes from bursts in flight, not from splitting one burst across lanes.
#include <stdlib.h>
#include "kernels.cuh"
#include "sigproc_device.cuh"
#include "demod_fast.cuh"
#include "fec_lane.cuh"

namespace btsdsp {

// ------------------------------------------------------------------------------------------------
// burst addressing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void burst_loc(const BurstSrc &s, long long i, long long *start, int *len) {
  if (s.gather) i = s.gather[i];
  long long g = s.first + i, stream0 = 0;
  if (s.narfcn > 0) {                               // [frame][arfcn][tn] over per-ARFCN slot streams
    const long long grp = i >> 3;
    g = (grp / s.narfcn) * 8 + (i & 7);
    stream0 = (grp % s.narfcn) * s.arfcn_pitch;
  }
  const int q = (int)(g & 3);
  const int rule_len = (q == 0 ? 157 : 156) * s.sps;
  if (s.pitch > 0) {
    *start = i * s.pitch;
    *len = s.lens ? s.lens[i] : rule_len;
  } else {
    const int off = q == 0 ? 0 : (q == 1 ? 157 : (q == 2 ? 313 : 469));
    *start = stream0 + ((g >> 2) * 625 + off) * s.sps;
    *len = rule_len;
  }
}

// warp-cooperative: copy the warp's nv bursts into tile rows (transposed), coalesced 8-byte loads
__device__ __forceinline__ void stage_in(cf *tile, const BurstSrc &src, long long w0, int nv, int lane) {
  for (int j = 0; j < nv; j++) {
    long long start; int len;
    burst_loc(src, w0 + j, &start, &len);
    if (len > kBurstRows) len = kBurstRows;
    const cf *g = src.base + start;
    for (int i = lane; i < len; i += 32) tile[i * kTileStride + j] = __ldg(g + i);
  }
  __syncwarp();
}

// warp-cooperative: write the soft bits held in column j of `tile` (.x of each row) to global
__device__ __forceinline__ void stage_out_soft(const cf *tile, float *soft, int soft_pitch, long long w0, int nv,
                                               int lane, bool ok, int len) {
  __syncwarp();
  const float *tf = (const float *)tile;
  for (int j = 0; j < nv; j++) {
    const bool okj = __shfl_sync(0xffffffffu, (int)ok, j) != 0;
    const int lenj = __shfl_sync(0xffffffffu, len, j);
    float *dst = soft + (w0 + j) * (long long)soft_pitch;
    for (int m = lane; m < soft_pitch; m += 32)
      dst[m] = (okj && m < lenj) ? tf[(m * kTileStride + j) * 2] : 0.0F;
  }
}

// ------------------------------------------------------------------------------------------------
// table construction (init only; single thread, same serial order as the reference's init code)
// ------------------------------------------------------------------------------------------------
__global__ void k_init_tables(DevTables *T) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // initGMSKRotationTables, sigProcLib.cpp:214-225
  float phase = 0.0F;
  const float inc = BTS_DIV(BTS_DIV(kPiF, 2.0F), (float)T->sps);
  for (int i = 0; i < 157 * T->sps; i++) {
    T->rot[i] = expj_lookup(T, phase);
    T->revrot[i] = expj_lookup(T, -phase);
    phase = BTS_ADD(phase, inc);
  }
}
__global__ void k_init_sinc_grid(DevTables *T) {
  const int j = blockIdx.x, i = threadIdx.x;          // 512 blocks x 24 threads
  float v = 0.0F;
  if (i < 21) {
    // sinc(M_PI_F*(i-10-frac)) (:588) == sinc(M_PI_F*(i'-ix)) (:651) for frac = j/512: the float
    // difference (i-10) - j/512 is exact either way.
    const float d = BTS_SUB((float)(i - 10), (float)j * (1.0F / (float)kSincGrid));
    v = sinc_exact(T, BTS_MUL(kPiF, d));
  }
  T->sinc_grid[j][i] = v;
}
void launch_init_tables(DevTables *T, cudaStream_t st) {
  k_init_tables<<<1, 32, 0, st>>>(T);
  k_init_sinc_grid<<<kSincGrid, 24, 0, st>>>(T);
}

// ------------------------------------------------------------------------------------------------
// single vectors in global memory
// ------------------------------------------------------------------------------------------------
// convolve / correlate, all four realOnly branches of sigProcLib.cpp:319-367; one thread per output.
__global__ void k_convolve(const cf *a, int la, int a_real, const cf *b, int lb, int b_real, cf *c, int start,
                           int outsz, int corr) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= outsz) return;
  const int t = start + n;
  cf sum = mk(0.0F, 0.0F);
  float rsum = 0.0F;
  for (int k = 0; k < lb; k++) {
    const int ai = t - k;
    if (ai < 0) break;
    if (ai >= la) continue;
    cf tap = corr ? (b_real ? mk(b[lb - 1 - k].x, 0.0F) : cconj(b[lb - 1 - k])) : b[k];
    const cf av = a[ai];
    if (a_real && b_real) rsum = BTS_ADD(rsum, BTS_MUL(av.x, tap.x));
    else if (a_real) sum = cadd(sum, cmulr(tap, av.x));
    else if (b_real) sum = cadd(sum, cmulr(av, tap.x));
    else sum = cadd(sum, cmul(av, tap));
  }
  c[n] = (a_real && b_real) ? mk(rsum, 0.0F) : sum;
}
void launch_convolve(const cf *a, int la, int a_real, const cf *b, int lb, int b_real, cf *c, int start, int outsz,
                     int corr, cudaStream_t st) {
  k_convolve<<<(outsz + 127) / 128, 128, 0, st>>>(a, la, a_real, b, lb, b_real, c, start, outsz, corr);
}

__global__ void k_peak_detect(const DevTables *T, const cf *v, int n, cf *peak, float *idx, float *avg) {
  if (threadIdx.x != 0) return;
  *peak = peak_detect<1, false>(T, View<1>{(cf *)v}, n, idx, avg);
}
void launch_peak_detect(const DevTables *T, const cf *v, int n, cf *peak, float *idx, float *avg, cudaStream_t st) {
  k_peak_detect<<<1, 32, 0, st>>>(T, v, n, peak, idx, avg);
}

__global__ void k_interp_point(const DevTables *T, const cf *v, int n, float ix, cf *out) {
  if (threadIdx.x != 0) return;
  *out = interp_point<1>(T, View<1>{(cf *)v}, n, ix);
}
void launch_interp_point(const DevTables *T, const cf *v, int n, float ix, cf *out, cudaStream_t st) {
  k_interp_point<<<1, 32, 0, st>>>(T, v, n, ix, out);
}

// delayVector on one long vector: fractional FIR one thread per sample into tmp, then the integer shift.
__global__ void k_delay_frac(const DevTables *T, const cf *v, int n, float frac, cf *tmp) {
  __shared__ float taps[21];
  if (threadIdx.x == 0) delay_taps(T, frac, taps);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tmp[i] = conv_cr_at<1>(View<1>{(cf *)v}, n, taps, 21, i + 10);
}
__global__ void k_delay_shift(const cf *src, cf *dst, int n, int intOffset) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = i - intOffset;                          // dst[i] = src[i - intOffset], zero outside
  dst[i] = (s >= 0 && s < n) ? src[s] : mk(0.0F, 0.0F);
}
void launch_delay_vector(const DevTables *T, cf *v, int n, float delay, cf *tmp, cudaStream_t st) {
  const int intOffset = (int)floorf(delay);
  const float frac = delay - (float)intOffset;
  const int nb = (n + 127) / 128;
  if ((double)fabsf(frac) > 1e-2) {
    k_delay_frac<<<nb, 128, 0, st>>>(T, v, n, frac, tmp);
    k_delay_shift<<<nb, 128, 0, st>>>(tmp, v, n, intOffset);
  } else if (intOffset != 0) {
    cudaMemcpyAsync(tmp, v, (size_t)n * sizeof(cf), cudaMemcpyDeviceToDevice, st);
    k_delay_shift<<<nb, 128, 0, st>>>(tmp, v, n, intOffset);
  }
}

__global__ void k_scale_vector(cf *v, int n, int real_only, cf s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = real_only ? cmulr(s, v[i].x) : cmul(v[i], s);      // scaleVector :713-730
}
void launch_scale_vector(cf *v, int n, int real_only, cf s, cudaStream_t st) {
  k_scale_vector<<<(n + 127) / 128, 128, 0, st>>>(v, n, real_only, s);
}

// the element-wise helpers of sigProcLib.cpp: addVector :746, offsetVector :760, conjugateVector :733, vectorSlicer :507
// GMSKRotate / GMSKReverseRotate :232-264 (one thread per element) and vectorNorm2 :146 (one thread, the reference's summation order; res[0] = sum |x|^2)
__global__ void k_vector_op(const DevTables *__restrict__ T, int op, cf *x, int n, int real_only, const cf *y, int ny, cf s,
                            float *res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (op == VOP_NORM2) {
    if (i != 0) return;
    float e = 0.0F;
    for (int k = 0; k < n; k++) e = BTS_ADD(e, cnorm2(x[k]));
    *res = e;
    return;
  }
  if (i >= n) return;
  const cf v = x[i];
  switch (op) {
    case VOP_ADD: if (i < ny) x[i] = cadd(v, y[i]); break;
    case VOP_OFFSET: x[i] = real_only ? mk(BTS_ADD(s.x, v.x), s.y) : cadd(v, s); break;   // `real + offset` = (offset.r + real, offset.i)
    case VOP_CONJ: if (!real_only) x[i] = cconj(v); break;
    case VOP_SLICE: x[i] = mk(soft_slice(v.x), 0.0F); break;
    case VOP_ROTATE: x[i] = real_only ? cmulr(T->rot[i], v.x) : cmul(T->rot[i], v); break;          // GMSKRotate :232-247
    case VOP_REVROTATE: x[i] = real_only ? cmulr(T->revrot[i], v.x) : cmul(T->revrot[i], v); break; // GMSKReverseRotate :249-264
  }
}
void launch_vector_op(const DevTables *T, int op, cf *x, int n, int real_only, const cf *y, int ny, cf s, float *res,
                      cudaStream_t st) {
  if (n <= 0) return;
  k_vector_op<<<op == VOP_NORM2 ? 1 : (n + 127) / 128, op == VOP_NORM2 ? 32 : 128, 0, st>>>(T, op, x, n, real_only, y, ny, s, res);
}

// RSSI of the RX datagram (Transceiver.cpp:400) for n amplitude magnitudes, through the host-built threshold table
__global__ void k_rssi(const DevTables *__restrict__ T, const float *__restrict__ a, int n, int *__restrict__ rssi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rssi[i] = trx_rssi(T, a[i]);
}
void launch_rssi(const DevTables *T, const float *a, int n, int *rssi, cudaStream_t st) {
  if (n > 0) k_rssi<<<(n + 127) / 128, 128, 0, st>>>(T, a, n, rssi);
}

__global__ void k_energy_detect(const cf *v, int n, unsigned win, float thr, float *avg, int *flag) {
  if (threadIdx.x != 0) return;
  *flag = energy_detect<1>(View<1>{(cf *)v}, n, win, thr, avg) ? 1 : 0;
}
void launch_energy_detect(const cf *v, int n, unsigned win, float thr, float *avg, int *flag, cudaStream_t st) {
  k_energy_detect<<<1, 32, 0, st>>>(v, n, win, thr, avg, flag);
}

__global__ void k_design_dfe_generic(const cf *chan, int nchan, float snr, int nf, cf *w, cf *b) {
  if (threadIdx.x != 0) return;
  cf ch[kDfeMax], W[kDfeMax], B[kDfeMax];
  for (int i = 0; i < nchan; i++) ch[i] = chan[i];
  design_dfe<0, 0>(ch, nchan - 1, snr, nf, W, B);
  for (int i = 0; i < nf; i++) w[i] = W[i];
  for (int i = 0; i < nchan - 1; i++) b[i] = B[i];
}
void launch_design_dfe_generic(const cf *chan, int nchan, float snr, int nf, cf *w, cf *b, cudaStream_t st) {
  k_design_dfe_generic<<<1, 32, 0, st>>>(chan, nchan, snr, nf, w, b);
}

// ------------------------------------------------------------------------------------------------
// GMSK modulation, batched: one thread per output sample (modulateBurst :521-565)
// guard = guards[i] when given, else 8 + ((first+i) % 4 == 0)  (Transceiver.cpp:105-106)
// ------------------------------------------------------------------------------------------------
__global__ void k_modulate(const DevTables *__restrict__ T, const uint8_t *__restrict__ bits, int nbits,
                           long long nbursts, int guard_rule, const uint8_t *__restrict__ guards, long long first,
                           cf *__restrict__ out, long long pitch, const float *__restrict__ scale) {
  const int sps = T->sps;
  // normal bursts at symbol rate: every sample is a signed sum of up to three entries of rot[n]*pulse[k]
  // (sigproc_device.cuh: tx_burst_sample), the table and the burst's bits held in shared memory
  const bool fast = sps == 1 && nbits == 148 && T->pulse_len == 3;
  __shared__ cf q[148 * 3];
  __shared__ unsigned char sb[148];
  if (fast) {
    for (int i = threadIdx.x; i < 148 * 3; i += blockDim.x) tx_fill_q(T, q, i);
    __syncthreads();
  }
  const long long i = blockIdx.x;
  for (long long bi = i; bi < nbursts; bi += gridDim.x) {
    const long long g = first + bi;
    const int guard = guards ? guards[bi] : (guard_rule >= 0 ? guard_rule : 8 + ((g & 3) == 0));
    const int n = sps * (nbits + guard);
    long long start;
    if (pitch > 0) start = bi * pitch;
    else {
      const int qd = (int)(g & 3);
      start = ((g >> 2) * 625 + (qd == 0 ? 0 : (qd == 1 ? 157 : (qd == 2 ? 313 : 469)))) * sps;
    }
    const uint8_t *bb = bits + bi * nbits;
    if (fast) {
      __syncthreads();
      for (int t = threadIdx.x; t < 148; t += blockDim.x) sb[t] = bb[t];
      __syncthreads();
    }
    const cf sc = mk(scale ? scale[bi] : 1.0F, 0.0F);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      cf x = fast ? tx_burst_sample(q, sb, t) : modulate_at(T, bb, nbits, n, sps, T->pulse, T->pulse_len, true, t);
      if (scale) x = cmul(x, sc);                  // addRadioVector's scaleVector(*modBurst, pow(10,-RSSI/10)), Transceiver.cpp:108
      out[start + t] = x;
    }
  }
}
void launch_modulate(const DevTables *T, const uint8_t *bits, int nbits, long long nbursts, int guard_rule,
                     const uint8_t *guards, long long first, cf *out, long long pitch, cudaStream_t st, const float *scale) {
  if (nbursts <= 0) return;
  const int grid = (int)(nbursts < 148 * 64 ? nbursts : 148 * 64);
  k_modulate<<<grid, 160, 0, st>>>(T, bits, nbits, nbursts, guard_rule, guards, first, out, pitch, scale);
}
// USRPifyVector of the second variant (Transceiver52M/radioInterface.cpp:100-118, powerScaling == 1.0): (short) casts of the
// already-scaled samples, no resampling -- that radio runs at the symbol rate
__global__ void k_usrpify(const cf *__restrict__ x, long long n, short2 *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cf v = x[i];
  short2 o;
  o.x = (short)(int)v.x;
  o.y = (short)(int)v.y;
  out[i] = o;
}
void launch_usrpify(const cf *x, long long n, int16_t *out, cudaStream_t st) {
  if (n > 0) k_usrpify<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, reinterpret_cast<short2 *>(out));
}
// the impulse-pulse variant generateMidamble needs (:794-797): pulse = {1.0} complex, guard 0
__global__ void k_modulate_impulse(const DevTables *T, const uint8_t *bits, int nbits, cf *out) {
  const int sps = T->sps, n = sps * nbits;
  const cf one = mk(1.0F, 0.0F);
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = modulate_at(T, bits, nbits, n, sps, &one, 1, false, t);
}
void launch_modulate_impulse(const DevTables *T, const uint8_t *bits, int nbits, cf *out, cudaStream_t st) {
  k_modulate_impulse<<<1, 128, 0, st>>>(T, bits, nbits, out);
}

// ------------------------------------------------------------------------------------------------
// normal-burst receive: energy gate -> analyzeTrafficBurst -> designDFE -> equalizeBurst
// (reference Transceiver.cpp:298-396 with estimateChannel == true for every burst); sps == 1.
// One burst per lane.  Two kernels so that each runs at the occupancy its working set allows:
//   k_detect_design : stages only the 36-sample midamble window (the 20-sample energy-gate window first, when
//                     gated), writes the correlation in place (45 tile rows = 12 KB per one-warp CTA, 16 per SM; the
//                     sinc grid is read from global memory), and leaves {1/amp, TOA - offset, w[7], b[5]} per burst
//                     in an EqParams record;
//   k_equalize_fast : streams the detected bursts, scaled by 1/amp on the way in, through a ROLLING 56-row tile
//                     (14.8 KB per warp, re-staged every ~31 rows; 14-15 warps per SM) under the pipelined equaliser
//                     of demod_fast.cuh.
// Staging loads are issued in batches (36 / 20 independent loads per lane) before their shared-memory stores so a
// warp has many requests in flight instead of one.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8z(void *smem_dst, const void *gsrc, bool valid) {   // zero-fills when !valid
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int nbytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct __align__(16) EqParams {       // 28 floats = 112 B per burst
  float ia_x, ia_y, toa_eq, ok;
  cf w[7];
  cf b[5];
};
constexpr size_t kGridBytes = (size_t)kSincGrid * kGridPitch * sizeof(float);
constexpr int kDetRows = 45;          // burst samples 56..91 in rows 9..44; the correlation is written IN PLACE to rows 0..35
constexpr int kDetWin = 9;           // (output n only needs window samples >= n - 8, so row n is dead when c[n] is stored)
constexpr size_t kDetTileBytes = (size_t)kDetRows * kTileStride * sizeof(cf);
constexpr size_t kEqTileBytes = (size_t)kEqRows * kTileStride * sizeof(cf);
// Up to 16 warps per CTA read the sinc grid from global memory (43 KB, L1/L2-resident): 12 KB of shared memory per warp
// lets 16 warps share an SM (registers then bind), 0.47 ms per 800 280 bursts as one-warp CTAs.  WARPS > 16 keeps a
// shared-memory copy of the grid per CTA (measured at 15 warps, 223 KB: 0.58 ms, and its 43 KB copy dominated small launches).
#ifndef BTS_DET_WARPS
#define BTS_DET_WARPS 5
#endif
#ifndef BTS_DET_SYNC
#define BTS_DET_SYNC 1
#endif
// Large batches run 5-warp CTAs (three resident per SM = 15 warps, what 128-130 registers allow) whose warps re-join at
// the phase boundaries: the kernel is ~140 KB of straight-line code that every warp walks once, its top stall is
// instruction fetch, and warps that stay together touch 3 code regions per SM instead of 15.  Per 800 280 bursts:
// one-warp CTAs 0.442 ms, 4 warps 0.408 (with barriers 0.413), 5 warps 0.487 without / 0.400 with barriers,
// 2/3/6/7 with barriers 0.411/0.445/0.443/0.448, 12/15 0.52/0.51 (profiles/README.md r3h).
// Small batches keep one-warp CTAs so they spread over all SMs.
constexpr int kDetWarps = BTS_DET_WARPS;
#ifndef BTS_EQ_WARPS
#define BTS_EQ_WARPS 1
#endif
constexpr int kEqWarps = BTS_EQ_WARPS;
constexpr long long kDetWideMin = 4096;     // warps
template <int WARPS> __host__ __device__ constexpr bool detect_grid_shared() { return WARPS > 16; }
template <int WARPS> __host__ __device__ constexpr size_t detect_grid_bytes() { return detect_grid_shared<WARPS>() ? kGridBytes : 0; }
template <int WARPS> constexpr size_t detect_smem() { return detect_grid_bytes<WARPS>() + WARPS * kDetTileBytes; }
template <int WARPS> constexpr size_t equalize_smem() { return WARPS * kEqTileBytes; }

// POLICY = pass 1 of the caller-policy pipeline (trx_policy.cuh): the energy is measured but not judged, the analysis
// runs on the slots `kind` marks as TSC, and the results go to a DetRec instead of into a DFE design.
// SPLIT = the stateless path with designDFE left to a second launch (k_design_eqp): this kernel stops after the analysis
// and parks {flag, amp, TOA, offset, channel} in a DetRec.
// (Asking ptxas for 18-19 resident warps per SM -- 113 / 107 registers -- measured 0.50 ms against 0.457: profiles/README.md r3e.
// No min-blocks argument here: even "1" changes the register allocation and costs 40 %.)
template <int WARPS, bool POLICY = false, bool SPLIT = false>
__global__ void __launch_bounds__(WARPS * 32) k_detect_design(const DevTables *__restrict__ T, BurstSrc src,
                                                              const uint8_t *__restrict__ tsc, long long n,
                                                              float detect_thr, float gate_thr, float snr_thr,
                                                              NormalOut out, EqParams *__restrict__ eqp,
                                                              const uint8_t *__restrict__ kind = nullptr,
                                                              DetRec *__restrict__ det = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *grid = reinterpret_cast<float *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf *A = reinterpret_cast<cf *>(smem_raw + detect_grid_bytes<WARPS>()) + (size_t)warp * kDetRows * kTileStride;
  if (detect_grid_shared<WARPS>()) {
    for (int i = threadIdx.x; i < kSincGrid * kGridPitch; i += WARPS * 32) grid[i] = T->sinc_grid[i / kGridPitch][i % kGridPitch];
    __syncthreads();
  }
  const long long w0 = ((long long)blockIdx.x * WARPS + warp) * 32;
  if (w0 >= n) return;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const bool gated = POLICY || gate_thr >= 0.0F;
  const long long i = w0 + lane;
  long long start = 0;
  int len = 0;
  if (lane < nv) burst_loc(src, i, &start, &len);

  // ---- energy gate first (its 20-sample window shares the tile): stage, evaluate, then overwrite
  bool pass = true;
  float avg_pwr = 0.0F;
  if (gated) {
    cf v[20];
#pragma unroll
    for (int it = 0; it < 20; it++) {
      const int e = it * 32 + lane, j = e / 20, r = e - j * 20;
      const long long sj = __shfl_sync(0xffffffffu, start, j);
      v[it] = (j < nv) ? __ldg(src.base + sj + r) : mk(0.0F, 0.0F);
    }
#pragma unroll
    for (int it = 0; it < 20; it++) {
      const int e = it * 32 + lane, j = e / 20, r = e - j * 20;
      A[r * kTileStride + j] = v[it];
    }
    __syncwarp();
    if (lane < nv) pass = energy_detect<kTileStride>(View<kTileStride>{A + lane}, len, 20, gate_thr, POLICY ? &avg_pwr : nullptr);   // Transceiver.cpp:298
    if (POLICY && lane < nv) pass = kind[i] == CORR_TSC;
    __syncwarp();
  }
  // ---- staging: element e = it*32 + lane of the warp's nv x 36 window samples (burst e/36, sample 56 + e%36)
  {
    cf v[36];
#pragma unroll
    for (int it = 0; it < 36; it++) {
      const int e = it * 32 + lane, j = e / 36, r = e - j * 36;
      const long long sj = __shfl_sync(0xffffffffu, start, j);
      v[it] = (j < nv) ? __ldg(src.base + sj + 56 + r) : mk(0.0F, 0.0F);
    }
#pragma unroll
    for (int it = 0; it < 36; it++) {
      const int e = it * 32 + lane, j = e / 36, r = e - j * 36;
      A[(kDetWin + r) * kTileStride + j] = v[it];
    }
  }
  __syncwarp();
  if (lane >= nv) return;

  const View<kTileStride> a{A + lane};
  const Grid g = detect_grid_shared<WARPS>() ? Grid{grid, kGridPitch} : Grid{&T->sinc_grid[0][0], 24, true};
  bool ok = false;
  cf amp = mk(0.0F, 0.0F), ia = mk(0.0F, 0.0F), chan[6], w[7], fb[5];
  float toa = 0.0F, off = 0.0F;
#if BTS_DET_SYNC
  // The CTA's warps re-join at the phase boundaries, so they stay in the same region of the straight-line code and share
  // its instruction-cache lines (exited warps and lanes do not count for the barrier).
  {
    if (pass) analyze_corr<kTileStride>(T, a.at(kDetWin), a, tsc[i]);
    __syncthreads();
    cf pk = mk(0.0F, 0.0F);
    float pt = 0.0F;
    if (pass) pk = peak_detect_fast<kTileStride>(g, a, 36, &pt);
    __syncthreads();
    if (pass) ok = analyze_tail<kTileStride, true>(g, T, a, tsc[i], detect_thr, pk, pt, &amp, &toa, chan, &off);
    __syncthreads();
  }
#else
  if (pass) ok = analyze_fast<kTileStride>(g, T, a.at(kDetWin), a, tsc[i], detect_thr, &amp, &toa, chan, &off);
#endif
  if (POLICY || SPLIT) {
    float4 *q = reinterpret_cast<float4 *>(det + i);
    q[0] = make_float4(avg_pwr, ok ? 1.0F : 0.0F, amp.x, amp.y);
    q[1] = make_float4(toa, ok ? off : 0.0F, 0.0F, 0.0F);
    q[2] = ok ? make_float4(chan[0].x, chan[0].y, chan[1].x, chan[1].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    q[3] = ok ? make_float4(chan[2].x, chan[2].y, chan[3].x, chan[3].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    q[4] = ok ? make_float4(chan[4].x, chan[4].y, chan[5].x, chan[5].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    return;
  }
  if (ok) {
    // Transceiver.cpp:340  SNRestimate = amplitude.norm2()/(thr*thr + 1.0)  (double division)
    const float SNR = (float)((double)cnorm2(amp) / ((double)BTS_MUL(snr_thr, snr_thr) + 1.0));
    ia = cdiv(mk(1.0F, 0.0F), amp);
#pragma unroll
    for (int j = 0; j < 6; j++) chan[j] = cmul(chan[j], ia);                                    // :346
    design_dfe<7, 5>(chan, 5, SNR, 7, w, fb);                                                   // :347
  } else {
#pragma unroll
    for (int j = 0; j < 7; j++) w[j] = mk(0.0F, 0.0F);
#pragma unroll
    for (int j = 0; j < 5; j++) fb[j] = mk(0.0F, 0.0F);
  }
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  if (out.off) out.off[i] = ok ? off : 0.0F;
  if (out.chan) for (int j = 0; j < 6; j++) out.chan[i * 6 + j] = ok ? chan[j] : mk(0.0F, 0.0F);
  if (out.w) for (int j = 0; j < 7; j++) out.w[i * 7 + j] = w[j];
  if (out.b) for (int j = 0; j < 5; j++) out.b[i * 5 + j] = fb[j];
  if (eqp) {
    float4 *q = reinterpret_cast<float4 *>(eqp + i);
    q[0] = make_float4(ia.x, ia.y, BTS_SUB(toa, off), ok ? 1.0F : 0.0F);                        // TOA - chanRespOffset :393
    q[1] = make_float4(w[0].x, w[0].y, w[1].x, w[1].y);
    q[2] = make_float4(w[2].x, w[2].y, w[3].x, w[3].y);
    q[3] = make_float4(w[4].x, w[4].y, w[5].x, w[5].y);
    q[4] = make_float4(w[6].x, w[6].y, fb[0].x, fb[0].y);
    q[5] = make_float4(fb[1].x, fb[1].y, fb[2].x, fb[2].y);
    q[6] = make_float4(fb[3].x, fb[3].y, fb[4].x, fb[4].y);
  }
}

__device__ __forceinline__ void store_soft4(float *row, bool vec, int pitch, int m0, const float s4[4], int len) {
  if (vec && m0 >= 0 && m0 + 3 < len && m0 + 3 < pitch) {
    *reinterpret_cast<float4 *>(row + m0) = make_float4(s4[0], s4[1], s4[2], s4[3]);
  } else {
#pragma unroll
    for (int r = 0; r < 4; r++) if (m0 + r >= 0 && m0 + r < len && m0 + r < pitch) row[m0 + r] = s4[r];
  }
}

// the wire format of the reference's RX datagram: (char) round(soft * 255.0) (Transceiver.cpp:669), 148 per burst
__device__ __forceinline__ unsigned soft_u8(float s) { return (unsigned)(int)round((double)s * 255.0) & 0xffu; }

// U8 = false: soft bits as float, row pitch soft_pitch floats.  U8 = true: soft bits as the datagram's bytes,
// row pitch soft_pitch BYTES (a multiple of 4), 148 per burst.
template <int WARPS, bool U8>
__global__ void __launch_bounds__(WARPS * 32) k_equalize_fast(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                              const EqParams *__restrict__ eqp, void *__restrict__ soft_,
                                                              int soft_pitch, int row_bytes = 0) {
  const int row_words = (row_bytes > 0 ? row_bytes : soft_pitch) / 4;   // U8: bytes of each row this kernel owns
  float *soft = reinterpret_cast<float *>(soft_);
  unsigned char *soft8 = reinterpret_cast<unsigned char *>(soft_);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  cf *A = reinterpret_cast<cf *>(smem_raw) + (size_t)warp * kEqRows * kTileStride;
  const long long w0 = ((long long)blockIdx.x * WARPS + warp) * 32;
  if (w0 >= n) return;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  long long start = 0;
  int len = 0;
  bool ok = false;
  cf ia = mk(0.0F, 0.0F), w[7], fb[5];
  float toa_eq = 0.0F;
  if (lane < nv) {
    burst_loc(src, i, &start, &len);
    if (len > kBurstRows - 3) len = kBurstRows - 3;
    const float4 *q = reinterpret_cast<const float4 *>(eqp + i);
    const float4 q0 = __ldg(q);
    ok = q0.w != 0.0F;
    if (ok) {
      const float4 q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5), q6 = __ldg(q + 6);
      ia = mk(q0.x, q0.y); toa_eq = q0.z;
      w[0] = mk(q1.x, q1.y); w[1] = mk(q1.z, q1.w); w[2] = mk(q2.x, q2.y); w[3] = mk(q2.z, q2.w);
      w[4] = mk(q3.x, q3.y); w[5] = mk(q3.z, q3.w); w[6] = mk(q4.x, q4.y);
      fb[0] = mk(q4.z, q4.w); fb[1] = mk(q5.x, q5.y); fb[2] = mk(q5.z, q5.w); fb[3] = mk(q6.x, q6.y); fb[4] = mk(q6.z, q6.w);
    }
  }
  float *row = soft + i * (long long)soft_pitch;
  unsigned *row8 = reinterpret_cast<unsigned *>(soft8 + i * (long long)soft_pitch);
  const bool vec = ((reinterpret_cast<uintptr_t>(row) | (uintptr_t)(soft_pitch * 4)) & 15) == 0;
  if (!ok && lane < nv) {                       // undetected: the row is all zeros (written while the others equalise)
    if (U8) for (int m = 0; m < row_words; m++) row8[m] = 0u;
    else if (vec) for (int m = 0; m < soft_pitch; m += 4) *reinterpret_cast<float4 *>(row + m) = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
    else for (int m = 0; m < soft_pitch; m++) row[m] = 0.0F;
  }
  const unsigned okmask = __ballot_sync(0xffffffffu, ok);
  if (okmask == 0) return;

  EqLane<kTileStride> eq;
  if (ok) eq.init(Grid{&T->sinc_grid[0][0], 24}, T, View<kTileStride>{A + lane}, len, toa_eq, w, fb);      // :392-396
  else eq.io = 0;
  const int io_min = __reduce_min_sync(0xffffffffu, ok ? eq.io : 0x7fffffff);
  const int io_max = __reduce_max_sync(0xffffffffu, ok ? eq.io : (int)0x80000000);
  const int nmax = __reduce_max_sync(0xffffffffu, ok ? len : 0);
  int base = 0;
  bool staged = false;
  cf ycur[4];
#pragma unroll
  for (int r = 0; r < 4; r++) ycur[r] = mk(0.0F, 0.0F);

  // Only the outputs the caller keeps are computed: the recursion is causal (soft bit m depends on decisions < m only),
  // so stopping after the last stored symbol leaves every stored value unchanged.  The reference's caller keeps the
  // first gSlotLen = 148 of the 156/157 (Transceiver.cpp:668); a wider soft_pitch gets them all.
  const int mend = U8 ? (nmax < 148 ? nmax : 148) : (nmax < soft_pitch ? nmax : soft_pitch);
  for (int m0 = kEqStart; m0 < mend; m0 += 4) {
    if (!staged || eq_needs_restage(base, m0, io_min, io_max)) {
      // ---- roll the tile: rows [base, base + kEqRows) of every detected burst (zeros outside the burst) arrive
      //      by cp.async, then each lane scales its column by 1/amplitude (scaleVector, Transceiver.cpp:391)
      base = m0 - io_max;
      staged = true;
      __syncwarp();
      for (unsigned rem = okmask; rem; rem &= rem - 1) {            // raw samples, global -> shared, all in flight
        const int j = __ffs(rem) - 1;
        const long long sj = __shfl_sync(0xffffffffu, start, j);
        const int lj = __shfl_sync(0xffffffffu, len, j);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int tr = lane + 32 * k, r = base + tr;
          if (tr < kEqRows) {
            const bool valid = (unsigned)r < (unsigned)lj;
            cp_async8z(A + tr * kTileStride + j, src.base + sj + (valid ? r : 0), valid);
          }
        }
      }
      cp_async_wait_all();
      __syncwarp();
      if (ok) {                                                     // scale the lane's own column in place
        const View<kTileStride> a{A + lane};
#pragma unroll 8
        for (int tr = 0; tr < kEqRows; tr++) a.st(tr, cmul(a.ld(tr), ia));
      }
    }
    if (ok) {
      float s4[4];
      if (__all_sync(okmask, eq.interior(m0 + 4))) eq.template step<false>(T, base, m0, ycur, s4);
      else eq.template step<true>(T, base, m0, ycur, s4);
      if (U8) {
        if (m0 >= 0 && m0 + 3 < 148)
          row8[m0 >> 2] = soft_u8(s4[0]) | (soft_u8(s4[1]) << 8) | (soft_u8(s4[2]) << 16) | (soft_u8(s4[3]) << 24);
      } else {
        store_soft4(row, vec, soft_pitch, m0, s4, len);
      }
    }
  }
  if (ok) {
    if (U8) for (int m = 37; m < row_words; m++) row8[m] = 0u;
    else for (int m = len; m < soft_pitch; m++) row[m] = 0.0F;
  }
}


// ------------------------------------------------------------------------------------------------
// k_equalize_ring: the same equaliser over a RING tile fed by a software pipeline instead of a rolling tile that is
// re-staged every eight steps.
//   * tile = kEqRing (32) rows per lane, row slot = (burst row + io) & 31: indexed on each lane's own output timeline,
//     so a step's 24 rows are in the same slots for every lane whatever its integer delay (no io_min/io_max policy,
//     no spread limit), 8.4 KB per warp instead of 14.8 KB (registers, not shared memory, now bound residency: 16/SM);
//   * every step the warp brings in the four rows the NEXT step newly needs, for all 32 bursts: lane L loads row
//     (L & 3) of bursts (L >> 2) + 8t, t = 0..3 -- 32-byte runs of eight bursts per instruction -- into registers at the
//     top of step s, and at the top of step s + 1 scales them by that burst's 1/amplitude (scaleVector,
//     Transceiver.cpp:391) and stores them.  The loads have a whole step (~700 instructions) to land; nothing waits;
//   * each burst row is fetched and scaled exactly once (the rolling tile re-fetched 24 of every 56 rows).
// Per-burst source records (pointer already offset by -io, valid mu range, 1/amp) sit in a 1 KB shared-memory table.
// ------------------------------------------------------------------------------------------------
struct __align__(16) EqSrc {
  const cf *p;          // burst base - io: sample of timeline index mu is p[mu]
  int lo, hi;           // mu in [lo, hi) exists in the burst; elsewhere the tile holds zeros
  float iax, iay;       // 1/amplitude
  int pad0, pad1;
};
#ifndef BTS_SLICER_RING_DEFAULT
#define BTS_SLICER_RING_DEFAULT true    // GPU parity run: identical on all tests and on 1 000 064 access bursts; 1.37 ms against 1.47 ms per 10^6 bursts
#endif
#ifndef BTS_EQ_PROLOGUE
#define BTS_EQ_PROLOGUE 1
#endif
#ifndef BTS_EQ_MINCTAS
#define BTS_EQ_MINCTAS 16
#endif
constexpr size_t kEqRingBytes = (size_t)kEqRing * kTileStride * sizeof(cf) + 32 * sizeof(EqSrc);

__device__ __forceinline__ void eq_ring_fetch(const EqSrc *__restrict__ tab, int lane, int mu0, cf v[4]) {
  const int mu = mu0 + (lane & 3);
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int4 e = *reinterpret_cast<const int4 *>(&tab[(lane >> 2) + 8 * t]);       // p, lo, hi
    const cf *p = reinterpret_cast<const cf *>((unsigned long long)(unsigned)e.x | ((unsigned long long)(unsigned)e.y << 32));
    v[t] = (mu >= e.z && mu < e.w) ? __ldg(p + mu) : mk(0.0F, 0.0F);
  }
}
__device__ __forceinline__ void eq_ring_store(cf *__restrict__ A, const EqSrc *__restrict__ tab, int lane, int mu0, const cf v[4]) {
  cf *dst = A + ((mu0 + (lane & 3)) & (kEqRing - 1)) * kTileStride + (lane >> 2);
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const float2 ia = *reinterpret_cast<const float2 *>(&tab[(lane >> 2) + 8 * t].iax);
    dst[8 * t] = cmul(v[t], mk(ia.x, ia.y));
  }
}

template <bool U8>
__global__ void __launch_bounds__(32, BTS_EQ_MINCTAS) k_equalize_ring(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                      const EqParams *__restrict__ eqp, void *__restrict__ soft_,
                                                      int soft_pitch, int row_bytes = 0) {
  const int row_words = (row_bytes > 0 ? row_bytes : soft_pitch) / 4;
  float *soft = reinterpret_cast<float *>(soft_);
  unsigned char *soft8 = reinterpret_cast<unsigned char *>(soft_);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  cf *A = reinterpret_cast<cf *>(smem_raw);
  EqSrc *tab = reinterpret_cast<EqSrc *>(A + kEqRing * kTileStride);
  const long long w0 = (long long)blockIdx.x * 32;
  if (w0 >= n) return;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  long long start = 0;
  int len = 0;
  bool ok = false;
  cf ia = mk(0.0F, 0.0F), w[7], fb[5];
  float toa_eq = 0.0F;
  if (lane < nv) {
    burst_loc(src, i, &start, &len);
    if (len > kBurstRows - 3) len = kBurstRows - 3;
    const float4 *q = reinterpret_cast<const float4 *>(eqp + i);
    const float4 q0 = __ldg(q);
    ok = q0.w != 0.0F;
    if (ok) {
      const float4 q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4), q5 = __ldg(q + 5), q6 = __ldg(q + 6);
      ia = mk(q0.x, q0.y); toa_eq = q0.z;
      w[0] = mk(q1.x, q1.y); w[1] = mk(q1.z, q1.w); w[2] = mk(q2.x, q2.y); w[3] = mk(q2.z, q2.w);
      w[4] = mk(q3.x, q3.y); w[5] = mk(q3.z, q3.w); w[6] = mk(q4.x, q4.y);
      fb[0] = mk(q4.z, q4.w); fb[1] = mk(q5.x, q5.y); fb[2] = mk(q5.z, q5.w); fb[3] = mk(q6.x, q6.y); fb[4] = mk(q6.z, q6.w);
    }
  }
  float *row = soft + i * (long long)soft_pitch;
  unsigned *row8 = reinterpret_cast<unsigned *>(soft8 + i * (long long)soft_pitch);
  const bool vec = ((reinterpret_cast<uintptr_t>(row) | (uintptr_t)(soft_pitch * 4)) & 15) == 0;
  if (!ok && lane < nv) {                       // undetected: the row is all zeros
    if (U8) for (int m = 0; m < row_words; m++) row8[m] = 0u;
    else if (vec) for (int m = 0; m < soft_pitch; m += 4) *reinterpret_cast<float4 *>(row + m) = make_float4(0.0F, 0.0F, 0.0F, 0.0F);
    else for (int m = 0; m < soft_pitch; m++) row[m] = 0.0F;
  }
  const unsigned okmask = __ballot_sync(0xffffffffu, ok);
  if (okmask == 0) return;

  const int io = ok ? (int)floorf(-toa_eq) : 0;              // EqLane::init's integer delay, needed before it for the table
  {
    EqSrc e;
    e.p = src.base + start - io;                // never dereferenced outside [lo, hi)
    e.lo = ok ? io : 0;
    e.hi = ok ? len + io : 0;
    e.iax = ia.x; e.iay = ia.y; e.pad0 = e.pad1 = 0;
    tab[lane] = e;
  }
  __syncwarp();
  // prologue: step(kEqStart) reads mu = kEqStart .. kEqStart+23.  All six groups are requested at once (their latency
  // overlaps the rotation-table copy and the lane's own set-up); five are stored before the loop, the sixth stays
  // pending in registers exactly as every later step finds it
  cf pend[4];
#if BTS_EQ_PROLOGUE
  cf pro[5][4];
#pragma unroll
  for (int g = 0; g < 5; g++) eq_ring_fetch(tab, lane, kEqStart + 4 * g, pro[g]);
#else
#pragma unroll 1
  for (int g = 0; g < 5; g++) {
    eq_ring_fetch(tab, lane, kEqStart + 4 * g, pend);
    eq_ring_store(A, tab, lane, kEqStart + 4 * g, pend);
  }
#endif
  eq_ring_fetch(tab, lane, kEqStart + 20, pend);
  // (A shared-memory copy of the rotation tables in place of the per-step global loads measured 1.40 ms against
  // 0.96 ms: profiles/README.md r3b.)
  EqLane<kTileStride> eq;
  if (ok) eq.init(Grid{&T->sinc_grid[0][0], 24}, T, View<kTileStride>{A + lane}, len, toa_eq, w, fb);      // :392-396
  else eq.io = 0;
  const int nmax = __reduce_max_sync(0xffffffffu, ok ? len : 0);
  const int mend = U8 ? (nmax < 148 ? nmax : 148) : (nmax < soft_pitch ? nmax : soft_pitch);
  cf ycur[4];
#pragma unroll
  for (int r = 0; r < 4; r++) ycur[r] = mk(0.0F, 0.0F);
#if BTS_EQ_PROLOGUE
#pragma unroll
  for (int g = 0; g < 5; g++) eq_ring_store(A, tab, lane, kEqStart + 4 * g, pro[g]);
#endif
#pragma unroll 1
  for (int m0 = kEqStart; m0 < mend; m0 += 4) {
    eq_ring_store(A, tab, lane, m0 + 20, pend);            // the rows this step newly needs (fetched during the previous step)
    __syncwarp();
    eq_ring_fetch(tab, lane, m0 + 24, pend);               // ... and the next step's, in flight under this step's arithmetic
    if (ok) {
      float s4[4];
      if (__all_sync(okmask, eq.interior(m0 + 4))) eq.template step<false, true>(T, 0, m0, ycur, s4);
      else eq.template step<true, true>(T, 0, m0, ycur, s4);
      if (U8) {
        if (m0 >= 0 && m0 + 3 < 148)
          row8[m0 >> 2] = soft_u8(s4[0]) | (soft_u8(s4[1]) << 8) | (soft_u8(s4[2]) << 16) | (soft_u8(s4[3]) << 24);
      } else {
        store_soft4(row, vec, soft_pitch, m0, s4, len);
      }
    }
    // (no second barrier: the next step's store goes to the slots of rows mu < m0 - 4, which no lane still reads)
  }
  if (ok) {
    if (U8) for (int m = 37; m < row_words; m++) row8[m] = 0u;
    else for (int m = len; m < soft_pitch; m++) row[m] = 0.0F;
  }
}

// second half of the split stateless path: designDFE from the parked records, then the same outputs as the fused kernel
__global__ void __launch_bounds__(64) k_design_eqp(long long n, const DetRec *__restrict__ det, float snr_thr, NormalOut out,
                                                   EqParams *__restrict__ eqp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 *dq = reinterpret_cast<const float4 *>(det + i);
  const float4 d0 = __ldg(dq), d1 = __ldg(dq + 1);
  const bool ok = d0.y != 0.0F;
  const cf amp = mk(d0.z, d0.w);
  const float toa = d1.x, off = d1.y;
  cf ia = mk(0.0F, 0.0F), chan[6], w[7], fb[5];
  if (ok) {
    const float4 c0 = __ldg(dq + 2), c1 = __ldg(dq + 3), c2 = __ldg(dq + 4);
    chan[0] = mk(c0.x, c0.y); chan[1] = mk(c0.z, c0.w); chan[2] = mk(c1.x, c1.y);
    chan[3] = mk(c1.z, c1.w); chan[4] = mk(c2.x, c2.y); chan[5] = mk(c2.z, c2.w);
    const float SNR = (float)((double)cnorm2(amp) / ((double)BTS_MUL(snr_thr, snr_thr) + 1.0));   // Transceiver.cpp:340
    ia = cdiv(mk(1.0F, 0.0F), amp);
#pragma unroll
    for (int j = 0; j < 6; j++) chan[j] = cmul(chan[j], ia);                                    // :346
    design_dfe<7, 5>(chan, 5, SNR, 7, w, fb);                                                   // :347
  } else {
#pragma unroll
    for (int j = 0; j < 6; j++) chan[j] = mk(0.0F, 0.0F);
#pragma unroll
    for (int j = 0; j < 7; j++) w[j] = mk(0.0F, 0.0F);
#pragma unroll
    for (int j = 0; j < 5; j++) fb[j] = mk(0.0F, 0.0F);
  }
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  if (out.off) out.off[i] = ok ? off : 0.0F;
  if (out.chan) for (int j = 0; j < 6; j++) out.chan[i * 6 + j] = chan[j];
  if (out.w) for (int j = 0; j < 7; j++) out.w[i * 7 + j] = w[j];
  if (out.b) for (int j = 0; j < 5; j++) out.b[i * 5 + j] = fb[j];
  if (eqp) {
    float4 *q = reinterpret_cast<float4 *>(eqp + i);
    q[0] = make_float4(ia.x, ia.y, BTS_SUB(toa, off), ok ? 1.0F : 0.0F);
    q[1] = make_float4(w[0].x, w[0].y, w[1].x, w[1].y);
    q[2] = make_float4(w[2].x, w[2].y, w[3].x, w[3].y);
    q[3] = make_float4(w[4].x, w[4].y, w[5].x, w[5].y);
    q[4] = make_float4(w[6].x, w[6].y, fb[0].x, fb[0].y);
    q[5] = make_float4(fb[1].x, fb[1].y, fb[2].x, fb[2].y);
    q[6] = make_float4(fb[3].x, fb[3].y, fb[4].x, fb[4].y);
  }
}
static bool g_eq_ring = true;
static bool g_det_split = false;
// EqParams records, then (split path) DetRec records behind them
size_t demod_scratch_bytes(long long n) { return (size_t)n * (sizeof(EqParams) + sizeof(DetRec)); }
// access bursts: the records plus a 160-sample correlation scratch row per burst (128-byte aligned behind the records)
size_t rach_scratch_bytes(long long n) { return (((size_t)n * sizeof(EqParams) + 127) & ~(size_t)127) + (size_t)n * 160 * sizeof(cf); }

// The two launches of the normal-burst receive path; `between` (optional) is recorded between them so a caller
// can time the kernels separately.
int launch_demod_normal(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr,
                        float gate_thr, float snr_thr, NormalOut out, void *scratch, cudaStream_t st,
                        cudaEvent_t between) {
  if (n <= 0) return 0;
  const long long nwarps = (n + 31) / 32;
  EqParams *eqp = (out.soft || out.soft_u8) ? reinterpret_cast<EqParams *>(scratch) : nullptr;
  if (g_det_split) {    // analysis and designDFE as two launches (BTSDSP_DET_SPLIT=1): measurement variant
    DetRec *det = reinterpret_cast<DetRec *>(reinterpret_cast<EqParams *>(scratch) + n);
    NormalOut none{};
    if (nwarps >= kDetWideMin)
      k_detect_design<kDetWarps, false, true><<<(unsigned)((nwarps + kDetWarps - 1) / kDetWarps), 32 * kDetWarps, detect_smem<kDetWarps>(), st>>>(T, src, tsc, n, detect_thr, gate_thr, snr_thr, none, nullptr, nullptr, det);
    else
      k_detect_design<1, false, true><<<(unsigned)nwarps, 32, detect_smem<1>(), st>>>(T, src, tsc, n, detect_thr, gate_thr, snr_thr, none, nullptr, nullptr, det);
    k_design_eqp<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(n, det, snr_thr, out, eqp);
  } else
  // one-warp CTAs, 12 KB of shared memory and 124 registers each: 16 resident per SM
  if (nwarps >= kDetWideMin)
    k_detect_design<kDetWarps><<<(unsigned)((nwarps + kDetWarps - 1) / kDetWarps), 32 * kDetWarps, detect_smem<kDetWarps>(), st>>>(T, src, tsc, n, detect_thr, gate_thr, snr_thr, out, eqp);
  else
    k_detect_design<1><<<(unsigned)nwarps, 32, detect_smem<1>(), st>>>(T, src, tsc, n, detect_thr, gate_thr, snr_thr, out, eqp);
  if (between) cudaEventRecord(between, st);
  if (!out.soft && !out.soft_u8) return 1;
  if (g_eq_ring) {     // ring tile + software-pipelined loads (default); BTSDSP_EQ_RING=0 selects the rolling-tile kernel
    if (out.soft_u8) k_equalize_ring<true><<<(unsigned)nwarps, 32, kEqRingBytes, st>>>(T, src, n, eqp, out.soft_u8, out.soft_pitch);
    else k_equalize_ring<false><<<(unsigned)nwarps, 32, kEqRingBytes, st>>>(T, src, n, eqp, out.soft, out.soft_pitch);
    return 2;
  }
  // one-warp CTAs: 14.8 KB of shared memory and 128 registers per thread each, 14-15 resident per SM
  if (out.soft_u8)
    k_equalize_fast<kEqWarps, true><<<(unsigned)((nwarps + kEqWarps - 1) / kEqWarps), 32 * kEqWarps, equalize_smem<kEqWarps>(), st>>>(T, src, n, eqp, out.soft_u8, out.soft_pitch);
  else
    k_equalize_fast<kEqWarps, false><<<(unsigned)((nwarps + kEqWarps - 1) / kEqWarps), 32 * kEqWarps, equalize_smem<kEqWarps>(), st>>>(T, src, n, eqp, out.soft, out.soft_pitch);
  return 2;
}

constexpr int kCorrRowsA = 36;
// ------------------------------------------------------------------------------------------------
// analyzeTrafficBurst alone, batched.  SM = true: sps == 1, shared-memory tiles.  SM = false: any
// sps, one thread per burst over global scratch (the functional path for sps = 4).
// ------------------------------------------------------------------------------------------------
template <bool SM>
__global__ void __launch_bounds__(32) k_analyze(const DevTables *__restrict__ T, BurstSrc src,
                                                const uint8_t *__restrict__ tsc, long long n, float detect_thr,
                                                int request, NormalOut out, cf *scratch) {
  extern __shared__ cf tile[];
  const int lane = threadIdx.x, sps = src.sps;
  const long long w0 = (long long)blockIdx.x * 32;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  if (SM) stage_in(tile, src, w0, nv, lane);
  if (lane >= nv) return;
  const long long i = w0 + lane;
  long long start; int len;
  burst_loc(src, i, &start, &len);
  cf amp = mk(0.0F, 0.0F), chan[6 * kMaxSps];
  float toa = 0.0F, off = 0.0F;
  bool ok;
  if (SM) {
    cf *B = tile + kBurstRows * kTileStride, *C = B + kCorrRowsA * kTileStride;
    ok = analyze_traffic<kTileStride, true>(T, View<kTileStride>{tile + lane}, tsc[i], detect_thr, 1,
                                            View<kTileStride>{B + lane}, View<kTileStride>{C + lane}, &amp, &toa,
                                            request != 0, chan, &off);
  } else {
    cf *s = scratch + (size_t)i * scratch_per_burst(sps);
    ok = analyze_traffic<1, true>(T, View<1>{(cf *)src.base + start}, tsc[i], detect_thr, sps, View<1>{s},
                                  View<1>{s + 36 * sps}, &amp, &toa, request != 0, chan, &off);
  }
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  const bool have = ok && request;
  if (out.off) out.off[i] = have ? off : 0.0F;
  if (out.chan) for (int j = 0; j < 6 * sps; j++) out.chan[i * 6 * sps + j] = have ? chan[j] : mk(0.0F, 0.0F);
}
constexpr int kCorrRows = 36;
constexpr size_t kAnalyzeSmem = (size_t)(kBurstRows + 2 * kCorrRows) * kTileStride * sizeof(cf);
int launch_analyze(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr, int request,
                   NormalOut out, cf *scratch, int force_generic, cudaStream_t st) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (src.sps == 1 && !force_generic) {
    k_analyze<true><<<grid, 32, kAnalyzeSmem, st>>>(T, src, tsc, n, detect_thr, request, out, scratch);
  } else {
    k_analyze<false><<<grid, 32, 0, st>>>(T, src, tsc, n, detect_thr, request, out, scratch);
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------
// access bursts, sps == 1, tuned: k_rach_detect (detectRACHBurst, one burst per lane: the 157 x 41 correlation
// register-blocked in blocks of four lags over a ROLLING 80-row tile -- 21 KB per warp, 10 warps per SM -- with the
// correlation parked in a global scratch row per burst; taps warp-uniform from __constant__ memory) then
// k_slicer_fast (demodulateBurst as a stream over the rolling tile).
// ------------------------------------------------------------------------------------------------
__constant__ cf c_rach_taps[41];                  // conj(rach_seq[40-k]) (:474-503)


constexpr size_t kRachRollBytes = (size_t)kRachRollRows * kTileStride * sizeof(cf);
constexpr int kRachScratchPitch = 160;                  // complex samples of correlation scratch per burst

// detectRACHBurst over a rolling 80-row tile (demod_fast.cuh: rach_corr4_roll / rach_finish); cs = n x 160 complex scratch
__global__ void __launch_bounds__(32) k_rach_detect(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                    float detect_thr, NormalOut out, EqParams *__restrict__ eqp,
                                                    cf *__restrict__ cs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *A = reinterpret_cast<cf *>(smem_raw);
  const int lane = threadIdx.x;
  const long long w0 = (long long)blockIdx.x * 32;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  const bool live = lane < nv;
  long long start = 0;
  int len = 0;
  if (live) { burst_loc(src, i, &start, &len); if (len > 157) len = 157; }
  const int nmax = __reduce_max_sync(0xffffffffu, len);
  cf *row = cs + (live ? i : w0) * kRachScratchPitch;
  int imax = -1, base = 0;
  float maxv = 0.0F;
  bool staged = false;
  for (int n0 = 0; n0 < nmax; n0 += 4) {
    if (!staged || rach_needs_restage(base, n0)) {
      // ---- roll the tile: burst rows [n0 - 20, +80) of every burst, zeros outside the burst; cp.async, all in flight
      base = n0 - 20;
      staged = true;
      __syncwarp();
      for (int j = 0; j < nv; j++) {
        const long long sj = __shfl_sync(0xffffffffu, start, j);
        const int lj = __shfl_sync(0xffffffffu, len, j);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int tr = lane + 32 * k, r = base + tr;
          if (tr < kRachRollRows) {
            const bool valid = (unsigned)r < (unsigned)lj;
            cp_async8z(A + tr * kTileStride + j, src.base + sj + (valid ? r : 0), valid);
          }
        }
      }
      cp_async_wait_all();
      __syncwarp();
    }
    if (live) {
      cf acc[4];
      rach_corr4_roll<kTileStride>(View<kTileStride>{A + lane}, base, c_rach_taps, n0, acc);
      if (n0 + 3 < len) {                                   // one full 32-byte sector per lane
        float4 *q = reinterpret_cast<float4 *>(row + n0);
        q[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
        q[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++) if (n0 + r < len) row[n0 + r] = acc[r];
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (n0 + r < len) {                                 // peakDetect's first strict maximum (:673-681)
          const float p = cnorm2(acc[r]);
          if (p > maxv) { maxv = p; imax = n0 + r; }
        }
      }
    }
  }
  __syncwarp();                                             // every lane is done with the burst rows
  if (!live) return;
  for (int k = 0; k < kRachWin; k++) {                      // the 26 lags around the lane's own maximum, back from scratch
    const int idx = imax - 12 + k;
    A[k * kTileStride + lane] = ((unsigned)idx < (unsigned)len) ? row[idx] : mk(0.0F, 0.0F);
  }
  cf amp = mk(0.0F, 0.0F);
  float toa = 0.0F;
  const bool ok = rach_finish<kTileStride>(Grid{&T->sinc_grid[0][0], 24, true}, T, View<kTileStride>{A + lane}, row, len, imax,
                                           detect_thr, &amp, &toa);
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  if (eqp) {
    const cf ia = ok ? cdiv(mk(1.0F, 0.0F), amp) : mk(0.0F, 0.0F);          // ((complex) 1.0)/channel :1066
    reinterpret_cast<float4 *>(eqp + i)[0] = make_float4(ia.x, ia.y, toa, ok ? 1.0F : 0.0F);
  }
}

__global__ void __launch_bounds__(32) k_slicer_fast(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                    const EqParams *__restrict__ eqp, float *__restrict__ soft, int soft_pitch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *A = reinterpret_cast<cf *>(smem_raw);
  const int lane = threadIdx.x;
  const long long w0 = (long long)blockIdx.x * 32;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  long long start = 0;
  int len = 0;
  bool ok = false;
  cf ia = mk(0.0F, 0.0F);
  float toa = 0.0F;
  if (lane < nv) {
    burst_loc(src, i, &start, &len);
    if (len > 157) len = 157;
    const float4 q0 = __ldg(reinterpret_cast<const float4 *>(eqp + i));
    ok = q0.w != 0.0F; ia = mk(q0.x, q0.y); toa = q0.z;
  }
  const unsigned okmask = __ballot_sync(0xffffffffu, ok);
  // ---- prefill, one coalesced row at a time: not detected -> zeros; detected -> delayVector's zero fill slices to 0.5
  for (int j = 0; j < nv; j++) {
    const bool okj = (okmask >> j) & 1u;
    const int lj = __shfl_sync(0xffffffffu, len, j);
    float *rj = soft + (w0 + j) * (long long)soft_pitch;
    for (int m = lane; m < soft_pitch; m += 32) rj[m] = (okj && m < lj) ? 0.5F : 0.0F;
  }
  if (okmask == 0) return;
  SlicerLane<kTileStride> sl;
  if (ok) sl.init(Grid{&T->sinc_grid[0][0], 24}, T, View<kTileStride>{A + lane}, len, toa);
  const int io = ok ? sl.f.io : 0;
  const int nmax = __reduce_max_sync(0xffffffffu, ok ? len : 0);
  int base = 0, xr = 0, ks = 0;                 // tile origin, filter index at the last re-stage, steps since then
  bool staged = false;
  // A lane's soft bits go to global memory as scattered 4-byte stores if written where they are produced (each store
  // instruction touches 32 rows).  Instead step k of a tile interval parks its four outputs in tile rows 4k..4k+3 of
  // the lane's own column -- burst rows the walk has just left behind -- and before the tile is re-staged the warp
  // writes every lane's run of outputs as coalesced 128-byte stores.
  auto flush = [&]() {
    __syncwarp();
    const int cnt = 4 * ks;
    for (unsigned rem = okmask; rem; rem &= rem - 1) {
      const int j = __ffs(rem) - 1;
      const int mj = xr + __shfl_sync(0xffffffffu, io, j);
      float *rj = soft + (w0 + j) * (long long)soft_pitch;
      for (int t = lane; t < cnt; t += 32) {
        const float v = reinterpret_cast<const float *>(A + t * kTileStride + j)[0];
        const int m = mj + t;
        if (v >= 0.0F && (unsigned)m < (unsigned)soft_pitch) rj[m] = v;
      }
    }
    __syncwarp();
  };
  for (int x0 = 0; x0 < nmax; x0 += 4) {
    if (!staged || slicer_needs_restage(base, x0)) {
      if (staged) flush();
      base = x0 - 10;
      xr = x0;
      ks = 0;
      staged = true;
      __syncwarp();
      for (unsigned rem = okmask; rem; rem &= rem - 1) {
        const int j = __ffs(rem) - 1;
        const long long sj = __shfl_sync(0xffffffffu, start, j);
        const int lj = __shfl_sync(0xffffffffu, len, j);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int tr = lane + 32 * k, r = base + tr;
          if (tr < kEqRows) {
            const bool valid = (unsigned)r < (unsigned)lj;
            cp_async8z(A + tr * kTileStride + j, src.base + sj + (valid ? r : 0), valid);
          }
        }
      }
      cp_async_wait_all();
      __syncwarp();
      if (ok) {
        const View<kTileStride> a{A + lane};
#pragma unroll 8
        for (int tr = 0; tr < kEqRows; tr++) a.st(tr, cmul(a.ld(tr), ia));           // scaleVector :1066
      }
    }
    if (ok) {
      float s4[4];
      bool valid[4];
      sl.step(T, base, x0, s4, valid);                    // reads tile rows >= 4*ks only
#pragma unroll
      for (int r = 0; r < 4; r++)                          // soft bits are in [0, 1]: -1 marks "no output here"
        reinterpret_cast<float *>(A + (4 * ks + r) * kTileStride + lane)[0] = valid[r] ? s4[r] : -1.0F;
    }
    ks++;
  }
  flush();
}

// demodulateBurst (:1056-1097) over the same ring tile and software pipeline as k_equalize_ring: each lane walks its OWN
// output index m (the tile is indexed on the lane's timeline, so lanes with TOAs anywhere in the slot read the same slots),
// soft[m] = slice(Re(revrot[m] * F[m - io])), F zeroed outside the burst (delayVector's zero fill slices to 0.5).  Blocks
// start at m = -2 (mod 4 == 2) so that the filter block index m - 6 is a multiple of four, as the ring's groups require;
// a lane's four outputs leave as two 8-byte stores.  Replaces k_slicer_fast's rolling tile + parked-output flush.
__global__ void __launch_bounds__(32) k_slicer_ring(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                    const EqParams *__restrict__ eqp, float *__restrict__ soft, int soft_pitch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  cf *A = reinterpret_cast<cf *>(smem_raw);
  EqSrc *tab = reinterpret_cast<EqSrc *>(A + kEqRing * kTileStride);
  const long long w0 = (long long)blockIdx.x * 32;
  if (w0 >= n) return;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  long long start = 0;
  int len = 0;
  bool ok = false;
  cf ia = mk(0.0F, 0.0F);
  float toa = 0.0F;
  if (lane < nv) {
    burst_loc(src, i, &start, &len);
    if (len > 157) len = 157;
    const float4 q0 = __ldg(reinterpret_cast<const float4 *>(eqp + i));
    ok = q0.w != 0.0F; ia = mk(q0.x, q0.y); toa = q0.z;
  }
  float *row = soft + i * (long long)soft_pitch;
  if (!ok && lane < nv) for (int m = 0; m < soft_pitch; m++) row[m] = 0.0F;
  const unsigned okmask = __ballot_sync(0xffffffffu, ok);
  if (okmask == 0) return;
  const int io = ok ? (int)floorf(-toa) : 0;
  {
    EqSrc e;
    e.p = src.base + start - io;
    e.lo = ok ? io : 0;
    e.hi = ok ? len + io : 0;
    e.iax = ia.x; e.iay = ia.y; e.pad0 = e.pad1 = 0;
    tab[lane] = e;
  }
  __syncwarp();
  constexpr int kM0 = -2;                                   // first block; its filter block kM0 - 6 reads mu = kM0-10 .. kM0+13
  cf pend[4], pro[5][4];
#pragma unroll
  for (int g = 0; g < 5; g++) eq_ring_fetch(tab, lane, kM0 - 10 + 4 * g, pro[g]);
  eq_ring_fetch(tab, lane, kM0 + 10, pend);
  EqLane<kTileStride> f;
  {
    cf zw[7], zb[5];
#pragma unroll
    for (int k = 0; k < 7; k++) zw[k] = mk(0.0F, 0.0F);
#pragma unroll
    for (int k = 0; k < 5; k++) zb[k] = mk(0.0F, 0.0F);
    if (ok) f.init(Grid{&T->sinc_grid[0][0], 24}, T, View<kTileStride>{A + lane}, len, toa, zw, zb);
    else f.io = 0;
  }
  const int nmax = __reduce_max_sync(0xffffffffu, ok ? len : 0);
  const int mend = nmax < soft_pitch ? nmax : soft_pitch;
#pragma unroll
  for (int g = 0; g < 5; g++) eq_ring_store(A, tab, lane, kM0 - 10 + 4 * g, pro[g]);
#pragma unroll 1
  for (int m0 = kM0; m0 < mend; m0 += 4) {
    eq_ring_store(A, tab, lane, m0 + 10, pend);
    __syncwarp();
    eq_ring_fetch(tab, lane, m0 + 14, pend);
    if (ok) {
      cf d[4];
      f.template newF4_ring<true>(m0 - 6, d);
      float s4[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int m = m0 + r;
#if BTS_EQ_CROT
        const float4 q = c_eq_rr[m < 0 ? 0 : (m > 156 ? 156 : m)];
        const cf rr = mk(q.z, q.w);
#else
        const cf rr = T->revrot[m < 0 ? 0 : (m > 156 ? 156 : m)];
#endif
        s4[r] = soft_slice(BTS_SUB(BTS_MUL(rr.x, d[r].x), BTS_MUL(rr.y, d[r].y)));     // Re(revrot[m] * D[m]) :232-264
      }
      if (m0 >= 0 && m0 + 3 < len && m0 + 3 < soft_pitch && ((reinterpret_cast<uintptr_t>(row + m0) & 7) == 0)) {
        *reinterpret_cast<float2 *>(row + m0) = make_float2(s4[0], s4[1]);
        *reinterpret_cast<float2 *>(row + m0 + 2) = make_float2(s4[2], s4[3]);
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++) if (m0 + r >= 0 && m0 + r < len && m0 + r < soft_pitch) row[m0 + r] = s4[r];
      }
    }
  }
  if (ok) for (int m = len; m < soft_pitch; m++) row[m] = 0.0F;
}
static bool g_slicer_ring = BTS_SLICER_RING_DEFAULT;
void launch_slicer(const DevTables *T, BurstSrc src, long long n, const EqParams *eqp, float *soft, int soft_pitch, cudaStream_t st) {
  if (n <= 0) return;
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (g_slicer_ring) k_slicer_ring<<<grid, 32, kEqRingBytes, st>>>(T, src, n, eqp, soft, soft_pitch);
  else k_slicer_fast<<<grid, 32, kEqTileBytes, st>>>(T, src, n, eqp, soft, soft_pitch);
}

void upload_rach_taps(const DevTables *hostT) {
#if BTS_EQ_CROT
  {
    float4 rr[160];
    for (int m = 0; m < 160; m++) {
      const int k = m < 157 ? m : 156;
      rr[m] = make_float4(hostT->rot[k].x, hostT->rot[k].y, hostT->revrot[k].x, hostT->revrot[k].y);
    }
    cudaMemcpyToSymbol(c_eq_rr, rr, sizeof rr);
  }
#endif
  cf h[41];
  for (int k = 0; k < 41; k++) h[k] = mk(hostT->rach_seq[40 - k].x, -hostT->rach_seq[40 - k].y);
  cudaMemcpyToSymbol(c_rach_taps, h, sizeof h);
}

// ------------------------------------------------------------------------------------------------
// access bursts: detectRACHBurst (+ demodulateBurst when demod != 0), Transceiver.cpp:360-389
// ------------------------------------------------------------------------------------------------
template <bool SM>
__global__ void __launch_bounds__(32) k_rach(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                             float detect_thr, int demod, NormalOut out, cf *scratch) {
  extern __shared__ cf tile[];
  const int lane = threadIdx.x, sps = src.sps;
  const long long w0 = (long long)blockIdx.x * 32;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  if (SM) stage_in(tile, src, w0, nv, lane);
  const long long i = w0 + lane;
  bool ok = false;
  int len = 0, nsoft = 0;
  cf *B = tile + kBurstRows * kTileStride;
  if (lane < nv) {
    long long start;
    burst_loc(src, i, &start, &len);
    cf amp = mk(0.0F, 0.0F);
    float toa = 0.0F;
    if (SM) {
      const View<kTileStride> a{tile + lane}, b{B + lane};
      ok = detect_rach<kTileStride, true>(T, a, len, detect_thr, 1, b, &amp, &toa);
      if (ok && demod) nsoft = demodulate_burst<kTileStride, 2 * kTileStride>(T, a, len, 1, amp, toa, b, (float *)(B + lane));
    } else {
      cf *s = scratch + (size_t)i * scratch_per_burst(sps);
      const View<1> corr{s}, x{s + 157 * sps};
      ok = detect_rach<1, true>(T, View<1>{(cf *)src.base + start}, len, detect_thr, sps, corr, &amp, &toa);
      if (ok && demod) {
        for (int m = 0; m < len; m++) x.st(m, src.base[start + m]);
        float *sp = out.soft + i * (long long)out.soft_pitch;
        nsoft = demodulate_burst<1, 1>(T, x, len, sps, amp, toa, corr, sp);
        for (int m = nsoft; m < out.soft_pitch; m++) sp[m] = 0.0F;
      } else if (demod && out.soft) {
        float *sp = out.soft + i * (long long)out.soft_pitch;
        for (int m = 0; m < out.soft_pitch; m++) sp[m] = 0.0F;
      }
    }
    if (out.flag) out.flag[i] = ok ? 1 : 0;
    if (out.amp) out.amp[i] = amp;
    if (out.toa) out.toa[i] = toa;
  }
  if (SM && demod && out.soft) stage_out_soft(B, out.soft, out.soft_pitch, w0, nv, lane, ok, nsoft);
}
constexpr size_t kRachSmem = (size_t)(2 * kBurstRows) * kTileStride * sizeof(cf);
int launch_rach(const DevTables *T, BurstSrc src, long long n, float detect_thr, int demod, NormalOut out, cf *scratch,
                int force_generic, cudaStream_t st, void *eq_scratch) {
  if (n <= 0) return 0;
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (src.sps == 1 && !force_generic && eq_scratch) {
    EqParams *eqp = reinterpret_cast<EqParams *>(eq_scratch);
    k_rach_detect<<<grid, 32, kRachRollBytes, st>>>(T, src, n, detect_thr, out, (demod && out.soft) ? eqp : nullptr,
                                                    reinterpret_cast<cf *>(reinterpret_cast<char *>(eq_scratch) + (((size_t)n * sizeof(EqParams) + 127) & ~(size_t)127)));
    if (!(demod && out.soft)) return 1;
    launch_slicer(T, src, n, eqp, out.soft, out.soft_pitch, st);
    return 2;
  }
  k_rach<false><<<grid, 32, 0, st>>>(T, src, n, detect_thr, demod, out, scratch);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// equalizeBurst with caller-supplied DFE taps (the cached-filter mode of Transceiver.cpp:391-396);
// sps == 1.  burst_out (optional) receives the delayed burst: the reference mutates its input.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_equalize(const DevTables *__restrict__ T, BurstSrc src, long long n,
                                                 const float *__restrict__ toa, const cf *__restrict__ w,
                                                 const cf *__restrict__ b, float *soft, int soft_pitch,
                                                 cf *burst_out, long long out_pitch) {
  extern __shared__ cf tile[];
  cf *A = tile, *B = tile + kBurstRows * kTileStride;
  const int lane = threadIdx.x;
  const long long w0 = (long long)blockIdx.x * 32;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  stage_in(A, src, w0, nv, lane);
  const long long i = w0 + lane;
  int len = 0;
  if (lane < nv) {
    long long start;
    burst_loc(src, i, &start, &len);
    cf W[7], F[5];
    for (int j = 0; j < 7; j++) W[j] = w[i * 7 + j];
    for (int j = 0; j < 5; j++) F[j] = b[i * 5 + j];
    equalize_burst<kTileStride, 2 * kTileStride>(T, View<kTileStride>{A + lane}, len, toa[i], W, 7, F, 5,
                                                 View<kTileStride>{B + lane}, (float *)(B + lane));
  }
  stage_out_soft(B, soft, soft_pitch, w0, nv, lane, lane < nv, len);
  if (burst_out) {
    for (int j = 0; j < nv; j++) {
      const int lenj = __shfl_sync(0xffffffffu, len, j);
      for (int m = lane; m < lenj; m += 32) burst_out[(w0 + j) * out_pitch + m] = A[m * kTileStride + j];
    }
  }
}
int launch_equalize(const DevTables *T, BurstSrc src, long long n, const float *toa, const cf *w, const cf *b,
                    float *soft, int soft_pitch, cf *burst_out, long long out_pitch, cudaStream_t st) {
  if (n <= 0) return 0;
  k_equalize<<<(unsigned)((n + 31) / 32), 32, kRachSmem, st>>>(T, src, n, toa, w, b, soft, soft_pitch, burst_out, out_pitch);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// demodulateBurst alone (slicer path), any sps: one thread per burst over global scratch
// ------------------------------------------------------------------------------------------------
__global__ void k_demodulate(const DevTables *__restrict__ T, BurstSrc src, long long n, const cf *__restrict__ amp,
                             const float *__restrict__ toa, float *soft, int soft_pitch, cf *scratch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long start; int len;
  burst_loc(src, i, &start, &len);
  cf *s = scratch + (size_t)i * scratch_per_burst(src.sps);
  const View<1> x{s}, tmp{s + 157 * src.sps};
  for (int m = 0; m < len; m++) x.st(m, src.base[start + m]);
  float *sp = soft + i * (long long)soft_pitch;
  const int ns = demodulate_burst<1, 1>(T, x, len, src.sps, amp[i], toa[i], tmp, sp);
  for (int m = ns; m < soft_pitch; m++) sp[m] = 0.0F;
}
int launch_demodulate(const DevTables *T, BurstSrc src, long long n, const cf *amp, const float *toa, float *soft,
                      int soft_pitch, cf *scratch, cudaStream_t st) {
  if (n <= 0) return 0;
  k_demodulate<<<(unsigned)((n + 31) / 32), 32, 0, st>>>(T, src, n, amp, toa, soft, soft_pitch, scratch);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// designDFE batched (Nf = 7, nu = 5): one thread per channel estimate, registers only
// ------------------------------------------------------------------------------------------------
__global__ void k_design_dfe(const cf *__restrict__ chan, const float *__restrict__ snr, long long n, cf *w, cf *b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cf ch[6], W[7], F[5];
  for (int j = 0; j < 6; j++) ch[j] = chan[i * 6 + j];
  design_dfe<7, 5>(ch, 5, snr[i], 7, W, F);
  for (int j = 0; j < 7; j++) w[i * 7 + j] = W[j];
  for (int j = 0; j < 5; j++) b[i * 5 + j] = F[j];
}
int launch_design_dfe(const cf *chan, const float *snr, long long n, cf *w, cf *b, cudaStream_t st) {
  if (n <= 0) return 0;
  k_design_dfe<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(chan, snr, n, w, b);
  return 1;
}

// equalizeBurst on one vector with arbitrary tap counts (the sigProcLib.h entry point); one thread.
__global__ void k_equalize_generic(const DevTables *T, cf *burst, int n, float toa, const cf *w, int nw, const cf *b,
                                   int nb, cf *tmp, float *soft) {
  if (threadIdx.x != 0) return;
  cf W[kDfeMax], F[kDfeMax];
  for (int i = 0; i < nw; i++) W[i] = w[i];
  for (int i = 0; i < nb; i++) F[i] = b[i];
  equalize_burst<1, 1>(T, View<1>{burst}, n, toa, W, nw, F, nb, View<1>{tmp}, soft);
}
void launch_equalize_generic(const DevTables *T, cf *burst, int n, float toa, const cf *w, int nw, const cf *b, int nb,
                             cf *tmp, float *soft, cudaStream_t st) {
  k_equalize_generic<<<1, 32, 0, st>>>(T, burst, n, toa, w, nw, b, nb, tmp, soft);
}

// ------------------------------------------------------------------------------------------------
// The second transceiver variant's analyzeTrafficBurst (windowed search, Transceiver52M/sigProcLib.cpp:966-1077):
// one thread per burst over global scratch (2 x (2*maxTOA+1) samples each); functional, any sps.
// ------------------------------------------------------------------------------------------------
__global__ void k_analyze_52m(const DevTables *__restrict__ T, BurstSrc src, const uint8_t *__restrict__ tsc, long long n,
                              float detect_thr, unsigned max_toa, int request, NormalOut out, cf *scratch, int scratch_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long start; int len;
  burst_loc(src, i, &start, &len);
  cf amp = mk(0.0F, 0.0F), chan[6 * kMaxSps];
  for (int j = 0; j < 6 * kMaxSps; j++) chan[j] = mk(0.0F, 0.0F);
  float toa = 0.0F, off = 0.0F;
  cf *s = scratch + (size_t)i * scratch_stride;
  const bool ok = analyze_traffic_52m<1>(T, View<1>{(cf *)src.base + start}, tsc[i], detect_thr, src.sps, max_toa, View<1>{s},
                                         View<1>{s + scratch_stride / 2}, &amp, &toa, request != 0, chan, &off);
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  const bool have = ok && request;
  if (out.off) out.off[i] = have ? off : 0.0F;
  if (out.chan) for (int j = 0; j < 6 * src.sps; j++) out.chan[i * 6 * src.sps + j] = have ? chan[j] : mk(0.0F, 0.0F);
}
int analyze_52m_scratch_stride(unsigned max_toa, int sps) {
  if (max_toa < 3u * sps) max_toa = 3 * sps;
  return 2 * (2 * (int)max_toa + 2);
}
// The same one burst per lane over a shared-memory tile (sps == 1): the warp stages the search window of its 32 bursts
// (burst rows 66-span .. 82+span, 16 + 2*span rows) transposed, and the correlation (2*maxTOA+1 lags) and delayVector's
// temporary live in the rows behind it -- (windowLen + 2*corrLen) x 33 x 8 B per warp: 10.6 KB at maxTOA <= 3, 24 KB at 12,
// 100 KB at the limit of 60.  POLICY = pass 1 of the caller-policy pipeline for this variant: the stride-4 energy of the
// slot is measured (Transceiver52M/sigProcLib.cpp:944-963), the analysis runs on the TSC slots, results go to a DetRec.
size_t detect_52m_smem(unsigned max_toa) {
  const Geo52 g = geo_52m(max_toa);
  return (size_t)(g.windowLen + 2 * g.corrLen + 1) * kTileStride * sizeof(cf);
}
template <bool POLICY>
__global__ void __launch_bounds__(32) k_detect_52m(const DevTables *__restrict__ T, BurstSrc src, const uint8_t *__restrict__ tsc,
                                                   long long n, float detect_thr, unsigned max_toa, int request, NormalOut out,
                                                   const uint8_t *__restrict__ kind, DetRec *__restrict__ det) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *A = reinterpret_cast<cf *>(smem_raw);
  const int lane = threadIdx.x;
  const long long w0 = (long long)blockIdx.x * 32;
  if (w0 >= n) return;
  const int nv = (int)((n - w0) < 32 ? (n - w0) : 32);
  const long long i = w0 + lane;
  const Geo52 g = geo_52m(max_toa);
  long long start = 0;
  int len = 0;
  if (lane < nv) burst_loc(src, i, &start, &len);
  for (int e = lane; e < 32 * g.windowLen; e += 32) {          // element e: burst e / windowLen, window row e % windowLen
    const int j = e / g.windowLen, r = e - j * g.windowLen;
    const long long sj = __shfl_sync(0xffffffffu, start, j);
    const int lj = __shfl_sync(0xffffffffu, len, j);
    A[r * kTileStride + j] = (j < nv && g.startIx + r < lj) ? __ldg(src.base + sj + g.startIx + r) : mk(0.0F, 0.0F);
  }
  __syncwarp();
  if (lane >= nv) return;
  const View<kTileStride> win{A + lane};
  bool ok = false, run = true;
  float avg_pwr = 0.0F;
  if (POLICY) {
    energy_detect_52m<1>(View<1>{(cf *)src.base + start}, len, 20, 0.0F, &avg_pwr);
    run = kind[i] == CORR_TSC;
  }
  cf amp = mk(0.0F, 0.0F), chan[6];
#pragma unroll
  for (int j = 0; j < 6; j++) chan[j] = mk(0.0F, 0.0F);
  float toa = 0.0F, off = 0.0F;
  // analyze_traffic_52m indexes the burst from its start: hand it a view whose row startIx is tile row 0
  if (run) ok = analyze_traffic_52m<kTileStride>(T, win.at(-g.startIx), tsc[i], detect_thr, 1, max_toa, win.at(g.windowLen),
                                                 win.at(g.windowLen + g.corrLen), &amp, &toa, request != 0, chan, &off);
  const bool have = ok && request;
  if (POLICY) {
    float4 *q = reinterpret_cast<float4 *>(det + i);
    q[0] = make_float4(avg_pwr, ok ? 1.0F : 0.0F, amp.x, amp.y);
    q[1] = make_float4(toa, have ? off : 0.0F, 0.0F, 0.0F);
    q[2] = have ? make_float4(chan[0].x, chan[0].y, chan[1].x, chan[1].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    q[3] = have ? make_float4(chan[2].x, chan[2].y, chan[3].x, chan[3].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    q[4] = have ? make_float4(chan[4].x, chan[4].y, chan[5].x, chan[5].y) : make_float4(0.F, 0.F, 0.F, 0.F);
    return;
  }
  if (out.flag) out.flag[i] = ok ? 1 : 0;
  if (out.amp) out.amp[i] = amp;
  if (out.toa) out.toa[i] = toa;
  if (out.off) out.off[i] = have ? off : 0.0F;
  if (out.chan) for (int j = 0; j < 6; j++) out.chan[i * 6 + j] = have ? chan[j] : mk(0.0F, 0.0F);
}
int configure_detect_52m() {
  const int bytes = (int)detect_52m_smem(60);
  cudaError_t e = cudaFuncSetAttribute(k_detect_52m<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFuncSetAttribute(k_detect_52m<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
int launch_analyze_52m(const DevTables *T, BurstSrc src, const uint8_t *tsc, long long n, float detect_thr, unsigned max_toa,
                       int request, NormalOut out, cf *scratch, cudaStream_t st) {
  if (n <= 0) return 0;
  if (src.sps == 1 && max_toa <= 60) {                  // tile-staged: the global scratch is not touched
    k_detect_52m<false><<<(unsigned)((n + 31) / 32), 32, detect_52m_smem(max_toa), st>>>(T, src, tsc, n, detect_thr, max_toa, request, out,
                                                                                        nullptr, nullptr);
    return 1;
  }
  k_analyze_52m<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(T, src, tsc, n, detect_thr, max_toa, request, out, scratch,
                                                         analyze_52m_scratch_stride(max_toa, src.sps));
  return 1;
}
__global__ void k_energy_detect_52m(const cf *v, int n, unsigned win, float thr, float *avg, int *flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *flag = energy_detect_52m<1>(View<1>{(cf *)v}, n, win, thr, avg) ? 1 : 0;
}
void launch_energy_detect_52m(const cf *v, int n, unsigned win, float thr, float *avg, int *flag, cudaStream_t st) {
  k_energy_detect_52m<<<1, 32, 0, st>>>(v, n, win, thr, avg, flag);
}

#include "trx_kernels.cuh"
#include "fec_kernels.cuh"

int configure_kernels() {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_detect_design<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<1>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_detect_design<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<1>());
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_detect_design<kDetWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<kDetWarps>());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_detect_design<kDetWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<kDetWarps>());
  if (e != cudaSuccess) return (int)e;
  if (kEqWarps != 1) {
    e = cudaFuncSetAttribute(k_equalize_fast<kEqWarps, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)equalize_smem<kEqWarps>());
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_equalize_fast<kEqWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)equalize_smem<kEqWarps>());
    if (e != cudaSuccess) return (int)e;
  }
  e = cudaFuncSetAttribute(k_equalize_fast<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)equalize_smem<1>());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_equalize_fast<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)equalize_smem<1>());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_analyze<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAnalyzeSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_rach_detect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRachRollBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_slicer_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEqTileBytes);
  if (e != cudaSuccess) return (int)e;
  if (configure_detect_52m() != 0) return -52;
  if (const char *e = getenv("BTSDSP_EQ_RING")) g_eq_ring = atoi(e) != 0;
  if (const char *e = getenv("BTSDSP_DET_SPLIT")) g_det_split = atoi(e) != 0;
  if (const char *e = getenv("BTSDSP_SLICER_RING")) g_slicer_ring = atoi(e) != 0;
  if (const char *e = getenv("BTSDSP_ENC_LANES")) g_enc_lanes = atoi(e);
  e = cudaFuncSetAttribute(k_xcch_encode_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncTileSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_tch_encode_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTchTileSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_slicer_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEqRingBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_detect_design<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<1>());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_detect_design<kDetWarps, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)detect_smem<kDetWarps>());
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_equalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRachSmem);
  return (int)e;
}

}  // namespace btsdsp
