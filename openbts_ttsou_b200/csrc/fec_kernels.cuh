// fec_kernels.cuh -- k_xcch_decode / k_rach_decode: one warp per code block, one lane per trellis candidate (see
// fec.cuh); included by kernels.cu.  Shared memory per warp: the block's match / mismatch costs, received hard-bit
// pairs and the decoded bits.
template <int NC, int NU>
struct __align__(8) VitSmem {
  static constexpr int STEPS = NU + kVitDeferral, TABLE = 2 * STEPS;
  float match[TABLE], mismatch[TABLE];
  unsigned char hard[TABLE];
  unsigned char in2[STEPS + 2];         // the two received hard bits of each step
  unsigned char u[NU + 6];
};

// tables already hold entries 0..NC-1 (match, mismatch, hard); pads them, runs the trellis, leaves u[NU] in S.u
template <int NC, int NU, typename ST = VitSmem<NC, NU>>       // ST: storage with arrays at least as large as VitSmem<NC, NU>'s
__device__ __forceinline__ void viterbi_warp(ST &S, int lane) {
  constexpr int STEPS = NU + kVitDeferral, TABLE = 2 * STEPS;
  __syncwarp();
  for (int k = NC + lane; k < TABLE; k += 32) { S.match[k] = 0.5F; S.mismatch[k] = 0.5F; S.hard[k] = S.hard[NC - 1]; }
  __syncwarp();
  for (int k = lane; k < STEPS; k += 32) S.in2[k] = (unsigned char)((S.hard[2 * k] << 1) | S.hard[2 * k + 1]);
  __syncwarp();
  constexpr unsigned long long GEN = vit_generator_lut();
  // lane c = candidate c; lanes 0..15 also hold survivor `lane` between steps
  float cost = 0.0F;
  unsigned ist = 0, ost = 0;
  const unsigned FULL = 0xffffffffu;
  for (int s = 0; s < STEPS; s++) {
    const int sp = lane >> 1;
    const float pc = __shfl_sync(FULL, cost, sp);                        // branchCandidates :338-358
    const unsigned pi = __shfl_sync(FULL, ist, sp), po = __shfl_sync(FULL, ost, sp);
    const unsigned ci = (pi << 1) | (unsigned)(lane & 1);
    const unsigned co = (po << 2) | (unsigned)((GEN >> (2 * (ci & 0x1fu))) & 3u);
    const unsigned mm = (unsigned)S.in2[s] ^ co;                         // getSoftCostMetrics :361-371
    const float2 ma = *reinterpret_cast<const float2 *>(&S.match[2 * s]), mi = *reinterpret_cast<const float2 *>(&S.mismatch[2 * s]);
    const float t = __fadd_rn((mm & 1u) ? mi.y : ma.y, ((mm >> 1) & 1u) ? mi.x : ma.x);
    const float cc = __fadd_rn(pc, t);
    const float hc = __shfl_down_sync(FULL, cc, 16);                     // pruneCandidates :374-382
    const unsigned hi = __shfl_down_sync(FULL, ci, 16), ho = __shfl_down_sync(FULL, co, 16);
    const bool low = cc < hc;
    cost = low ? cc : hc; ist = low ? ci : hi; ost = low ? co : ho;      // meaningful in lanes 0..15
    // minCost :385-397: first strict minimum over survivors 0..15.  Path costs are sums of positive floats, so their
    // bit patterns order like the values: one warp-wide integer min, then the lowest lane that holds it.
    const unsigned key = lane < 16 ? __float_as_uint(cost) : 0xffffffffu;
    const unsigned mn = __reduce_min_sync(FULL, key);
    const int bi = __ffs(__ballot_sync(FULL, key == mn)) - 1;
    if (s >= kVitDeferral) {
      const unsigned wi = __shfl_sync(FULL, ist, bi);
      if (lane == 0) S.u[s - kVitDeferral] = (unsigned char)((wi >> kVitDeferral) & 1u);
    }
  }
  __syncwarp();
}

constexpr int kXcchWarps = 8;
__global__ void __launch_bounds__(kXcchWarps * 32) k_xcch_decode(const unsigned char *__restrict__ soft, int burst_pitch, long long nframes,
                                                                unsigned char *__restrict__ u, int *__restrict__ ok) {
  __shared__ VitSmem<kXcchC, kXcchU> sm[kXcchWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * kXcchWarps + warp;
  if (f >= nframes) return;
  VitSmem<kXcchC, kXcchU> &S = sm[warp];
  const unsigned char *fs = soft + f * 4 * (long long)burst_pitch;
  // ---- deinterleave + cost tables (:616-630; BitVector.cpp:440-497), 456 entries over the lanes
  for (int k = lane; k < kXcchC; k += 32) {
    int B;
    const int bit = xcch_source_bit(k, &B);
    unsigned h;
    vit_costs((float)fs[B * burst_pitch + bit] / 256.0F, &S.match[k], &S.mismatch[k], &h);
    S.hard[k] = (unsigned char)h;
  }
  viterbi_warp<kXcchC, kXcchU>(S, lane);
  for (int i = lane; i < kXcchU; i += 32) u[f * kXcchU + i] = S.u[i];
  if (lane == 0) ok[f] = xcch_parity_ok(S.u) ? 1 : 0;
}

int launch_xcch_decode(const unsigned char *soft, int burst_pitch, long long nframes, unsigned char *u, int *ok, cudaStream_t st) {
  if (nframes <= 0) return 0;
  k_xcch_decode<<<(unsigned)((nframes + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(soft, burst_pitch, nframes, u, ok);
  return 1;
}

// TCH/FACCH: one warp per block of a traffic channel's burst stream (block q = bursts 4q .. 4q+7, fec.cuh); a stolen block
// runs the 456-bit FACCH (XCCH) decode, the others the 378-bit class-1 decode + class-2 slice + parity / tail checks
__global__ void __launch_bounds__(kXcchWarps * 32) k_tch_decode(const unsigned char *__restrict__ soft, int burst_pitch, long long nblocks,
                                                               unsigned char *__restrict__ d, int *__restrict__ good, int *__restrict__ stolen_o,
                                                               unsigned char *__restrict__ fu, int *__restrict__ fok) {
  __shared__ VitSmem<kXcchC, kXcchU> sm[kXcchWarps];
  __shared__ unsigned char c2[kXcchWarps][kTchC2 + 2];
  __shared__ unsigned char dd[kXcchWarps][kTchD + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kXcchWarps + warp;
  if (q >= nblocks) return;
  VitSmem<kXcchC, kXcchU> &S = sm[warp];
  const unsigned char *bs = soft + q * 4 * (long long)burst_pitch;
  for (int k = lane; k < kXcchC; k += 32) {
    int B;
    const int bit = tch_source_bit(k, &B);
    unsigned h;
    vit_costs((float)bs[B * burst_pitch + bit] / 256.0F, &S.match[k], &S.mismatch[k], &h);
    S.hard[k] = (unsigned char)h;
    if (k >= kTchC1) c2[warp][k - kTchC1] = (unsigned char)h;                  // class 2, before the trellis pads over it
  }
  const bool stolen = (float)bs[7 * burst_pitch + 60] / 256.0F > 0.5F;         // warp-uniform
  if (lane == 0 && stolen_o) stolen_o[q] = stolen ? 1 : 0;
  if (stolen) {
    viterbi_warp<kXcchC, kXcchU>(S, lane);
    if (fu) for (int i = lane; i < kXcchU; i += 32) fu[q * kXcchU + i] = S.u[i];
    if (lane == 0 && fok) fok[q] = xcch_parity_ok(S.u) ? 1 : 0;
    if (d) for (int i = lane; i < kTchD; i += 32) d[q * kTchD + i] = 0;
    if (lane == 0 && good) good[q] = 0;
  } else {
    viterbi_warp<kTchC1, kTchU, VitSmem<kXcchC, kXcchU>>(S, lane);
    int g = 0;
    if (lane == 0) g = tch_fields(S.u, c2[warp], dd[warp]) ? 1 : 0;
    __syncwarp();
    if (d) for (int i = lane; i < kTchD; i += 32) d[q * kTchD + i] = dd[warp][i];
    if (lane == 0 && good) good[q] = g;
    if (fu) for (int i = lane; i < kXcchU; i += 32) fu[q * kXcchU + i] = 0;
    if (lane == 0 && fok) fok[q] = 0;
  }
}
int launch_tch_decode(const unsigned char *soft, int burst_pitch, long long nblocks, unsigned char *d, int *good, int *stolen,
                      unsigned char *fu, int *fok, cudaStream_t st) {
  if (nblocks <= 0) return 0;
  k_tch_decode<<<(unsigned)((nblocks + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(soft, burst_pitch, nblocks, d, good, stolen, fu, fok);
  return 1;
}

// RACH: one warp per access burst; out[i] = {tail, bsic, ra, 0} packed in one int32: tail | bsic << 8 | ra << 16
__global__ void __launch_bounds__(kXcchWarps * 32) k_rach_decode(const unsigned char *__restrict__ soft, int burst_pitch, long long n,
                                                                unsigned char *__restrict__ u, int *__restrict__ fields) {
  __shared__ VitSmem<kRachC, kRachU> sm[kXcchWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * kXcchWarps + warp;
  if (i >= n) return;
  VitSmem<kRachC, kRachU> &S = sm[warp];
  const unsigned char *bs = soft + i * (long long)burst_pitch;
  for (int k = lane; k < kRachC; k += 32) {
    unsigned h;
    vit_costs((float)bs[49 + k] / 256.0F, &S.match[k], &S.mismatch[k], &h);   // burst.segment(49,36) :478
    S.hard[k] = (unsigned char)h;
  }
  viterbi_warp<kRachC, kRachU>(S, lane);
  if (u && lane < kRachU) u[i * kRachU + lane] = S.u[lane];
  if (lane == 0) {
    int tail, bsic, ra;
    rach_fields(S.u, &tail, &bsic, &ra);
    fields[i] = tail | (bsic << 8) | (ra << 16);
  }
}
int launch_rach_decode(const unsigned char *soft, int burst_pitch, long long n, unsigned char *u, int *fields, cudaStream_t st) {
  if (n <= 0) return 0;
  k_rach_decode<<<(unsigned)((n + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(soft, burst_pitch, n, u, fields);
  return 1;
}
