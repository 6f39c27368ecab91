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

// ---- L1 encoders on the transmit side (fec.cuh): one warp per XCCH frame / per traffic-channel block ----------------------
// Parity words are remainders of a linear code: the word of a frame is the XOR of the words of its set bits.  r[i] = the
// encoder state a single 1 at position i of an n-bit message leaves behind (Generator::encoderShift run over the rest as zeros),
// built at compile time; a lane XORs the entries of its bits and the warp combines them with shuffles.
template <int N>
struct ParityTable { unsigned long long r[N]; };
template <int N>
__host__ __device__ constexpr ParityTable<N> make_parity_table(unsigned long long coeff, int len) {
  ParityTable<N> t{};
  const unsigned long long mask = (1ULL << len) - 1;
  for (int i = 0; i < N; i++) {
    unsigned long long state = coeff & mask;                             // the 1 at position i meets a zero state: fb = 1
    for (int s = i + 1; s < N; s++) {
      const unsigned long long fb = (state >> (len - 1)) & 1ULL;
      state = (state << 1) & mask;
      if (fb) state ^= coeff & mask;
    }
    t.r[i] = state;
  }
  return t;
}
__constant__ ParityTable<184> c_fire_par = make_parity_table<184>(0x10004820009ULL, 40);
__constant__ ParityTable<50> c_tch_par = make_parity_table<50>(0x0bULL, 3);
__device__ __forceinline__ unsigned long long warp_xor64(unsigned long long v) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __reduce_xor_sync(0xffffffffu, lo);
  hi = __reduce_xor_sync(0xffffffffu, hi);
  return ((unsigned long long)hi << 32) | lo;
}
// u[0..184) = the frame (optionally LSB8MSB), then the inverted Fire-code word and four tail zeros; c = its 456 coded bits
__device__ __forceinline__ void xcch_encode_warp(const unsigned char *__restrict__ frame, int lsb8msb, unsigned char *u, unsigned char *c,
                                                 int lane) {
  unsigned long long acc = 0;
  for (int i = lane; i < 184; i += 32) {
    const unsigned char b = frame[lsb8msb ? lsb8msb_src(i) : i] & 1;
    u[i] = b;
    if (b) acc ^= c_fire_par.r[i];
  }
  const unsigned long long p = ~warp_xor64(acc);
  for (int j = lane; j < 44; j += 32) u[184 + j] = j < 40 ? (unsigned char)((p >> (39 - j)) & 1ULL) : 0;
  __syncwarp();
  constexpr unsigned long long GEN = vit_generator_lut();
  for (int k = lane; k < kXcchU; k += 32) {
    unsigned h = 0;
#pragma unroll
    for (int t = 0; t < 5; t++) if (k - t >= 0) h |= (unsigned)u[k - t] << t;
    const unsigned g = (unsigned)((GEN >> (2 * h)) & 3u);
    c[2 * k] = (unsigned char)(g >> 1);
    c[2 * k + 1] = (unsigned char)(g & 1u);
  }
  __syncwarp();
}
struct EncSmem { unsigned char u[232], c[kXcchC]; };

__global__ void __launch_bounds__(kXcchWarps * 32) k_xcch_encode(const unsigned char *__restrict__ frames, long long nframes, int lsb8msb,
                                                                unsigned tsc_word, int have_tsc, unsigned char *__restrict__ bursts,
                                                                int burst_pitch) {
  __shared__ EncSmem sm[kXcchWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * kXcchWarps + warp;
  if (f >= nframes) return;
  EncSmem &S = sm[warp];
  xcch_encode_warp(frames + f * 184, lsb8msb, S.u, S.c, lane);
  unsigned char *out = bursts + f * 4 * (long long)burst_pitch;
  for (int idx = lane; idx < 4 * 148; idx += 32) {                       // fixed fields: tails, stealing flags (both set), midamble
    const int B = idx / 148, pos = idx - B * 148;
    if (pos < 3 || pos >= 145) out[B * burst_pitch + pos] = 0;
    else if (pos == 60 || pos == 87) out[B * burst_pitch + pos] = 1;
    else if (pos >= 61 && pos < 87) out[B * burst_pitch + pos] = burst_tsc_bit(tsc_word, have_tsc, pos);
  }
  for (int k = lane; k < kXcchC; k += 32) {                              // interleave :811-819 + mapping on the bursts :842-843
    int B;
    const int pos = xcch_source_bit(k, &B);
    out[B * burst_pitch + pos] = S.c[k];
  }
}
int launch_xcch_encode(const unsigned char *frames, long long nframes, int lsb8msb, unsigned tsc_word, int have_tsc, unsigned char *bursts,
                       int burst_pitch, cudaStream_t st) {
  if (nframes <= 0) return 0;
  k_xcch_encode<<<(unsigned)((nframes + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(frames, nframes, lsb8msb, tsc_word, have_tsc,
                                                                                              bursts, burst_pitch);
  return 1;
}

// Warp q in [-1, nblocks]: block q writes its own bytes of bursts 4q..4q+7 -- the first four bursts' tails, midamble, Hu and even
// e-bits, the last four bursts' Hl and odd e-bits -- so no two warps write the same byte.  q = -1 stands for the previous call's
// last block (`carry`, or a channel that starts here: zeros), q = nblocks for the block that is not there yet (zeros).
__global__ void __launch_bounds__(kXcchWarps * 32) k_tch_encode(const unsigned char *__restrict__ d260, const unsigned char *__restrict__ f184,
                                                               const unsigned char *__restrict__ steal, long long nblocks, int lsb8msb,
                                                               unsigned tsc_word, int have_tsc, const unsigned char *__restrict__ carry,
                                                               unsigned char *__restrict__ bursts, int burst_pitch) {
  __shared__ EncSmem sm[kXcchWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kXcchWarps + warp - 1;
  if (q > nblocks) return;
  EncSmem &S = sm[warp];
  const bool real = q >= 0 && q < nblocks;
  int st = 0;
  if (real) {
    st = steal[q] ? 1 : 0;
    if (st) {
      xcch_encode_warp(f184 + q * 184, lsb8msb, S.u, S.c, lane);
    } else {
      const unsigned char *d = d260 + q * kTchD;
      unsigned long long acc = 0;
      for (int i = lane; i < 50; i += 32) if (d[i] & 1) acc ^= c_tch_par.r[i];
      const unsigned p = ~(unsigned)warp_xor64(acc);
      for (int k = lane; k <= 90; k += 32) { S.u[k] = d[2 * k] & 1; S.u[184 - k] = d[2 * k + 1] & 1; }   // :1262-1265
      if (lane < 3) S.u[91 + lane] = (unsigned char)((p >> (2 - lane)) & 1u);                              // :1258-1259
      if (lane >= 4 && lane < 8) S.u[185 + lane - 4] = 0;                                                  // :1269
      __syncwarp();
      constexpr unsigned long long GEN = vit_generator_lut();
      for (int k = lane; k < kTchU; k += 32) {
        unsigned h = 0;
#pragma unroll
        for (int t = 0; t < 5; t++) if (k - t >= 0) h |= (unsigned)S.u[k - t] << t;
        const unsigned g = (unsigned)((GEN >> (2 * h)) & 3u);
        S.c[2 * k] = (unsigned char)(g >> 1);
        S.c[2 * k + 1] = (unsigned char)(g & 1u);
      }
      for (int i = lane; i < kTchC2; i += 32) S.c[kTchC1 + i] = d[182 + i] & 1;                           // :1275
      __syncwarp();
    }
  }
  unsigned char *lo = bursts + 4 * q * (long long)burst_pitch;           // bursts 4q .. 4q+3 (q >= 0)
  unsigned char *hi = lo + 4 * (long long)burst_pitch;                   // bursts 4q+4 .. 4q+7 (q < nblocks)
  if (q >= 0) {
    for (int idx = lane; idx < 4 * 148; idx += 32) {
      const int B = idx / 148, pos = idx - B * 148;
      if (pos < 3 || pos >= 145) lo[B * burst_pitch + pos] = 0;
      else if (pos == 87) lo[B * burst_pitch + pos] = (unsigned char)st;
      else if (pos >= 61 && pos < 87) lo[B * burst_pitch + pos] = burst_tsc_bit(tsc_word, have_tsc, pos);
    }
  }
  if (q < nblocks && lane < 4)
    hi[lane * burst_pitch + 60] = q >= 0 ? (unsigned char)st : (carry ? (unsigned char)(carry[lane * burst_pitch + 60] & 1) : 0);
  for (int k = lane; k < kXcchC; k += 32) {
    int r;
    const int pos = tch_source_bit(k, &r);
    if (r < 4) {
      if (q >= 0) lo[r * burst_pitch + pos] = real ? S.c[k] : 0;
    } else if (q < nblocks) {
      hi[(r - 4) * burst_pitch + pos] = real ? S.c[k] : (carry ? (unsigned char)(carry[(r - 4) * burst_pitch + pos] & 1) : 0);
    }
  }
}
int launch_tch_encode(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks, int lsb8msb,
                      unsigned tsc_word, int have_tsc, const unsigned char *carry, unsigned char *bursts, int burst_pitch, cudaStream_t st) {
  if (nblocks < 0) return 0;
  const long long warps = nblocks + 2;
  k_tch_encode<<<(unsigned)((warps + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(d260, f184, steal, nblocks, lsb8msb, tsc_word,
                                                                                           have_tsc, carry, bursts, burst_pitch);
  return 1;
}
