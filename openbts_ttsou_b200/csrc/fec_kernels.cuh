// fec_kernels.cuh -- k_xcch_decode: one warp per L2 frame, one lane per trellis candidate (see fec.cuh); included by
// kernels.cu.  Shared memory per warp: the frame's 504 match / mismatch costs, hard bits and the 228 decoded bits.
constexpr int kXcchWarps = 8;
struct __align__(8) XcchSmem {
  float match[kXcchTable], mismatch[kXcchTable];
  unsigned char hard[kXcchTable];
  unsigned char in2[kXcchSteps];        // the two received hard bits of each step
  unsigned char u[kXcchU + 4];
};

__global__ void __launch_bounds__(kXcchWarps * 32) k_xcch_decode(const unsigned char *__restrict__ soft, int burst_pitch, long long nframes,
                                                                unsigned char *__restrict__ u, int *__restrict__ ok) {
  __shared__ XcchSmem sm[kXcchWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * kXcchWarps + warp;
  if (f >= nframes) return;
  XcchSmem &S = sm[warp];
  const unsigned char *fs = soft + f * 4 * (long long)burst_pitch;
  // ---- deinterleave + cost tables (:616-630; BitVector.cpp:440-497), 456 entries over the lanes
  for (int k = lane; k < kXcchTable; k += 32) {
    if (k < kXcchC) {
      int B;
      const int bit = xcch_source_bit(k, &B);
      unsigned h;
      vit_costs((float)fs[B * burst_pitch + bit] / 256.0F, &S.match[k], &S.mismatch[k], &h);
      S.hard[k] = (unsigned char)h;
    } else {
      S.match[k] = 0.5F; S.mismatch[k] = 0.5F;
    }
  }
  __syncwarp();
  for (int k = kXcchC + lane; k < kXcchTable; k += 32) S.hard[k] = S.hard[kXcchC - 1];
  __syncwarp();
  for (int k = lane; k < kXcchSteps; k += 32) S.in2[k] = (unsigned char)((S.hard[2 * k] << 1) | S.hard[2 * k + 1]);
  __syncwarp();
  constexpr unsigned long long GEN = vit_generator_lut();
  // ---- Viterbi: lane c = candidate c; lanes 0..15 also hold survivor `lane` between steps
  float cost = 0.0F;
  unsigned ist = 0, ost = 0;
  const unsigned FULL = 0xffffffffu;
  for (int s = 0; s < kXcchSteps; s++) {
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
  for (int i = lane; i < kXcchU; i += 32) u[f * kXcchU + i] = S.u[i];
  if (lane == 0) ok[f] = xcch_parity_ok(S.u) ? 1 : 0;
}

int launch_xcch_decode(const unsigned char *soft, int burst_pitch, long long nframes, unsigned char *u, int *ok, cudaStream_t st) {
  if (nframes <= 0) return 0;
  k_xcch_decode<<<(unsigned)((nframes + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(soft, burst_pitch, nframes, u, ok);
  return 1;
}
