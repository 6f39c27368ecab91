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

// ---- L1 encoders on the transmit side (fec.cuh) --------------------------------------------------------------------------
// HBM-bound byte work (184 bytes in, 4 x 148 out per frame), so the shape is: coded bits of a frame in shared memory, then every
// output byte fetched through a compile-time permutation table (interleaver + mapping on the burst + fixed fields folded into one
// index per output byte) and the bursts written as coalesced 32-bit words.
//
// Parity words are remainders of a linear code: the word of a frame is the XOR of the words of its set bits.  r[i] = the
// encoder state a single 1 at position i of an n-bit message leaves behind (Generator::encoderShift run over the rest as zeros),
// built at compile time; a lane XORs the entries of its bits and the warp combines them.
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
__device__ const ParityTable<184> d_fire_par = make_parity_table<184>(0x10004820009ULL, 40);   // lane-indexed reads: global, not __constant__
__device__ const ParityTable<50> d_tch_par = make_parity_table<50>(0x0bULL, 3);

__device__ const EncTable d_xcch_table = make_xcch_table();
__device__ const EncTable d_tch_table = make_tch_table();

__device__ __forceinline__ unsigned long long warp_xor64(unsigned long long v) {
  unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
  lo = __reduce_xor_sync(0xffffffffu, lo);
  hi = __reduce_xor_sync(0xffffffffu, hi);
  return ((unsigned long long)hi << 32) | lo;
}
// the rate-1/2 K = 5 code over u[0..n): c[2k], c[2k+1] (BitVector::encode)
__device__ __forceinline__ void conv_encode_warp(const unsigned char *u, int n, unsigned char *c, int lane) {
  constexpr unsigned long long GEN = vit_generator_lut();
  for (int k = lane; k < n; k += 32) {
    unsigned h = 0;
#pragma unroll
    for (int t = 0; t < 5; t++) if (k - t >= 0) h |= (unsigned)u[k - t] << t;
    const unsigned g = (unsigned)((GEN >> (2 * h)) & 3u);
    *reinterpret_cast<unsigned short *>(c + 2 * k) = (unsigned short)((g >> 1) | ((g & 1u) << 8));
  }
}
// u[0..184) = the frame (optionally LSB8MSB), then the inverted Fire-code word and four tail zeros; c = its 456 coded bits
__device__ __forceinline__ void xcch_encode_warp(const unsigned char *__restrict__ frame, int lsb8msb, unsigned char *u, unsigned char *c,
                                                 int lane) {
  unsigned long long acc = 0;
  for (int i = lane; i < 184; i += 32) {
    const unsigned char b = frame[lsb8msb ? lsb8msb_src(i) : i] & 1;
    u[i] = b;
    if (b) acc ^= d_fire_par.r[i];
  }
  const unsigned long long p = ~warp_xor64(acc);
  for (int j = lane; j < 44; j += 32) u[184 + j] = j < 40 ? (unsigned char)((p >> (39 - j)) & 1ULL) : 0;
  __syncwarp();
  conv_encode_warp(u, kXcchU, c, lane);
  __syncwarp();
}
// a traffic channel's speech frame d[260] -> c[456] (encodeTCH :1248-1279)
__device__ __forceinline__ void tch_speech_encode_warp(const unsigned char *__restrict__ d, unsigned char *u, unsigned char *c, int lane) {
  unsigned long long acc = 0;
  for (int i = lane; i < 50; i += 32) if (d[i] & 1) acc ^= d_tch_par.r[i];
  const unsigned p = ~(unsigned)warp_xor64(acc);
  for (int k = lane; k <= 90; k += 32) { u[k] = d[2 * k] & 1; u[184 - k] = d[2 * k + 1] & 1; }       // :1262-1265
  if (lane < 3) u[91 + lane] = (unsigned char)((p >> (2 - lane)) & 1u);                               // :1258-1259
  if (lane >= 4 && lane < 8) u[185 + lane - 4] = 0;                                                   // :1269
  __syncwarp();
  conv_encode_warp(u, kTchU, c, lane);
  for (int i = lane; i < kTchC2; i += 32) c[kTchC1 + i] = d[182 + i] & 1;                             // :1275
  __syncwarp();
}
// four bursts out: every byte through the table, from the coded bits at `src` or the group's fixed-field word `sp`
__device__ __forceinline__ unsigned enc_byte(const unsigned char *src, unsigned sp, unsigned idx) {
  return (idx & kEncSpecial) ? (sp >> (idx & 31u)) & 1u : (unsigned)src[idx];
}
__device__ __forceinline__ void enc_write_group(const EncTable &tab, const unsigned char *src, unsigned sp, unsigned char *out, int burst_pitch,
                                                int lane) {
  if (((reinterpret_cast<uintptr_t>(out) | (uintptr_t)burst_pitch) & 3) == 0) {
    for (int w = lane; w < 148; w += 32) {
      const int B = w / 37, wi = w - 37 * B;
      const uint2 t = __ldg(reinterpret_cast<const uint2 *>(&tab.idx[4 * w]));
      const unsigned v = enc_byte(src, sp, t.x & 0xffffu) | (enc_byte(src, sp, t.x >> 16) << 8) | (enc_byte(src, sp, t.y & 0xffffu) << 16) |
                         (enc_byte(src, sp, t.y >> 16) << 24);
      *reinterpret_cast<unsigned *>(out + (long long)B * burst_pitch + 4 * wi) = v;
    }
  } else {
    for (int o = lane; o < 4 * 148; o += 32) {
      const int B = o / 148, pos = o - 148 * B;
      out[(long long)B * burst_pitch + pos] = (unsigned char)enc_byte(src, sp, tab.idx[o]);
    }
  }
}
struct alignas(8) EncSmem { unsigned char u[232]; };

// one warp per frame
__global__ void __launch_bounds__(kXcchWarps * 32) k_xcch_encode(const unsigned char *__restrict__ frames, long long nframes, int lsb8msb,
                                                                unsigned sp_base, unsigned char *__restrict__ bursts, int burst_pitch) {
  __shared__ EncSmem sm[kXcchWarps];
  __shared__ __align__(8) unsigned char cs[kXcchWarps][kXcchC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * kXcchWarps + warp;
  if (f >= nframes) return;
  xcch_encode_warp(frames + f * 184, lsb8msb, sm[warp].u, cs[warp], lane);
  enc_write_group(d_xcch_table, cs[warp], sp_base | (1u << 2) | (1u << 3), bursts + f * 4 * (long long)burst_pitch, burst_pitch, lane);   // Hl = Hu = 1 :735-736
}
// lane form (fec_lane.cuh): one lane per frame, the block in registers as bit-packed words
__device__ const CrcTable d_fire_crc = make_fire_crc_table();
constexpr int kEncLaneThreads = 128;
__global__ void __launch_bounds__(kEncLaneThreads) k_xcch_encode_lanes(const unsigned char *__restrict__ frames, long long nframes, int lsb8msb,
                                                                     unsigned sp_base, unsigned char *__restrict__ bursts) {
  __shared__ unsigned long long crc[256];
  for (int i = threadIdx.x; i < 256; i += kEncLaneThreads) crc[i] = d_fire_crc.t[i];
  __syncthreads();
  const long long f = (long long)blockIdx.x * kEncLaneThreads + threadIdx.x;
  if (f >= nframes) return;
  xcch_encode_frame_lane(frames + f * 184, lsb8msb, crc, sp_base, bursts + f * 592);
}
// The same with the OUTPUT rows staged through shared memory: every lane assembles its four bursts in its own 592-byte row and
// sends the row to global memory as ONE bulk async copy (cp.async.bulk, 592 contiguous bytes) instead of 37 stores of 16 bytes at a
// 592-byte lane stride (0.61 -> 0.22 ms per 2^20 frames).  Row pitch 592 B = 148 words keeps a quarter-warp's 16-byte stores on
// distinct banks.  The frames are read straight from global memory: staging them too (coalesced 16-byte loads into a 5.9 KB tile
// per warp) measured 0.31 ms -- the extra shared memory costs more residency than the lane-strided loads cost bandwidth.
constexpr int kEncTileWarps = 4;
constexpr size_t kEncTileSmem = kEncTileWarps * 32 * 592 + 2048, kTchTileSmem = kEncTileWarps * 32 * 304 + 2048;
__device__ __forceinline__ void enc_bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\ncp.async.bulk.commit_group;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
// HALVES: the group leaves in two pieces (quads 0..18 and 19..36: 304 + 288 bytes) through one 304-byte row per lane -- half the
// shared memory, twice the resident warps, one more wait on the copy engine.  Measured per 2^20 blocks: traffic channel 0.278 ->
// 0.234 ms (kept), XCCH 0.212 -> 0.293 ms (its lanes sit on the serial CRC chain already; it keeps the whole-group row).
template <class TAB, bool HALVES>
__device__ __forceinline__ void enc_store_group(const unsigned *pl, unsigned sp, unsigned char *row, unsigned char *gdst) {
  if (HALVES) {
    enc_out_range<TAB, 0, 19>(pl, sp, row);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the row is visible to the copy engine
    enc_bulk_store(gdst, row, 304u);
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // ... and has been read: it can be written again
    enc_out_range<TAB, 19, 18>(pl, sp, row);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    enc_bulk_store(gdst + 304, row, 288u);
  } else {
    enc_out_range<TAB, 0, 37>(pl, sp, row);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    enc_bulk_store(gdst, row, 592u);
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // before the CTA's memory goes away
}
__global__ void __launch_bounds__(kEncTileWarps * 32) k_xcch_encode_tiles(const unsigned char *__restrict__ frames, long long nframes, int lsb8msb,
                                                                        unsigned sp_base, unsigned char *__restrict__ bursts) {
  extern __shared__ __align__(16) unsigned char enc_smem[];
  unsigned long long *crc = reinterpret_cast<unsigned long long *>(enc_smem);
  for (int i = threadIdx.x; i < 256; i += kEncTileWarps * 32) crc[i] = d_fire_crc.t[i];
  __syncthreads();
  const long long f = (long long)blockIdx.x * (kEncTileWarps * 32) + threadIdx.x;
  if (f >= nframes) return;
  unsigned W[8], pl[16];
  xcch_u_words(frames + f * 184, lsb8msb, crc, W);
  conv_planes(W, pl, pl + 8);
  enc_store_group<XcchTab, false>(pl, sp_base | (1u << 2) | (1u << 3), enc_smem + 2048 + threadIdx.x * 592, bursts + f * 592);   // Hl = Hu = 1 :735-736
}
static int g_enc_lanes = 2;              // BTSDSP_ENC_LANES: 0 = the warp-per-block kernels everywhere, 1 = lane form storing straight to global memory, 2 = rows out through shared memory
int launch_xcch_encode(const unsigned char *frames, long long nframes, int lsb8msb, unsigned tsc_word, int have_tsc, unsigned char *bursts,
                       int burst_pitch, cudaStream_t st) {
  if (nframes <= 0) return 0;
  if (g_enc_lanes && burst_pitch == 148 && (reinterpret_cast<uintptr_t>(frames) & 3) == 0 && (reinterpret_cast<uintptr_t>(bursts) & 15) == 0) {
    if (g_enc_lanes == 2) {
      k_xcch_encode_tiles<<<(unsigned)((nframes + kEncTileWarps * 32 - 1) / (kEncTileWarps * 32)), kEncTileWarps * 32, kEncTileSmem, st>>>(
          frames, nframes, lsb8msb, enc_sp_base(tsc_word, have_tsc), bursts);
      return 1;
    }
    k_xcch_encode_lanes<<<(unsigned)((nframes + kEncLaneThreads - 1) / kEncLaneThreads), kEncLaneThreads, 0, st>>>(
        frames, nframes, lsb8msb, enc_sp_base(tsc_word, have_tsc), bursts);
    return 1;
  }
  k_xcch_encode<<<(unsigned)((nframes + kXcchWarps - 1) / kXcchWarps), kXcchWarps * 32, 0, st>>>(frames, nframes, lsb8msb,
                                                                                              enc_sp_base(tsc_word, have_tsc), bursts, burst_pitch);
  return 1;
}

// A CTA of kTchGroups + 1 warps makes kTchGroups consecutive groups of four bursts: warp w codes block G0 - 1 + w into row w of
// the shared array; after the CTA barrier warp w >= 1 assembles group g = G0 + w - 1 from rows w - 1 (block g - 1: odd e-bits, Hl)
// and w (block g: even e-bits, Hu) -- adjacent rows, so one table index reaches both.  Block -1 is the previous call's last block
// (`carry`: its odd e-bits and Hl are read back from the four half-filled bursts) or nothing (zeros); block nblocks is not
// there yet (zeros): group nblocks is the half-filled carry of the next call.  One block in kTchGroups is coded twice.
constexpr int kTchGroups = 8;
__global__ void __launch_bounds__((kTchGroups + 1) * 32) k_tch_encode(const unsigned char *__restrict__ d260, const unsigned char *__restrict__ f184,
                                                                      const unsigned char *__restrict__ steal, long long nblocks, int lsb8msb,
                                                                      unsigned sp_base, const unsigned char *__restrict__ carry,
                                                                      unsigned char *__restrict__ bursts, int burst_pitch) {
  __shared__ EncSmem sm[kTchGroups + 1];
  __shared__ __align__(8) unsigned char cs[kTchGroups + 1][kXcchC];
  __shared__ unsigned char stolen[kTchGroups + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kTchGroups - 1 + warp;                  // the block this warp codes
  unsigned char *c = cs[warp];
  int st = 0;
  if (q >= 0 && q < nblocks) {
    st = steal[q] ? 1 : 0;
    if (st) xcch_encode_warp(f184 + q * 184, lsb8msb, sm[warp].u, c, lane);           // dispatch :1323-1333
    else tch_speech_encode_warp(d260 + q * kTchD, sm[warp].u, c, lane);
  } else if (q == -1 && carry) {
    for (int k = lane; k < kXcchC; k += 32) {
      int r;
      const int pos = tch_source_bit(k, &r);
      c[k] = r >= 4 ? (unsigned char)(carry[(long long)(r - 4) * burst_pitch + pos] & 1) : 0;
    }
    st = carry[60] & 1;                                                               // mPreviousFACCH
  } else {
    for (int k = lane; k < kXcchC; k += 32) c[k] = 0;
  }
  if (lane == 0) stolen[warp] = (unsigned char)st;
  __syncthreads();
  const long long g = q;                                                              // warp w >= 1 assembles group g = its own block's index
  if (warp == 0 || g > nblocks) return;
  const unsigned sp = sp_base | ((unsigned)stolen[warp - 1] << 2) | ((unsigned)stolen[warp] << 3);   // Hl = previous, Hu = current :1365-1366
  enc_write_group(d_tch_table, cs[warp - 1], sp, bursts + g * 4 * (long long)burst_pitch, burst_pitch, lane);
}
// Lane form for the traffic channel (fec_lane.cuh): lane L of a warp codes block b0 + L into bit-packed code planes, takes the
// planes of block b0 + L - 1 from its neighbour by shuffle, and lanes 1..31 assemble groups b0 + 1 .. b0 + 31 (31 groups per warp:
// one block in 32 is coded twice), each in its own 592-byte shared-memory row that leaves as one bulk async copy.  Groups
// 0 .. kTchLaneFirst - 1 (group 0 needs the previous call's carry) stay with the first CTA of k_tch_encode.
constexpr int kTchLaneFirst = kTchGroups, kTchLaneGroups = 31;
__global__ void __launch_bounds__(kEncTileWarps * 32) k_tch_encode_tiles(const unsigned char *__restrict__ d260, const unsigned char *__restrict__ f184,
                                                                       const unsigned char *__restrict__ steal, long long nblocks, int lsb8msb,
                                                                       unsigned sp_base, unsigned char *__restrict__ bursts) {
  extern __shared__ __align__(16) unsigned char enc_smem[];
  unsigned long long *crc = reinterpret_cast<unsigned long long *>(enc_smem);
  for (int i = threadIdx.x; i < 256; i += kEncTileWarps * 32) crc[i] = d_fire_crc.t[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long G = kTchLaneFirst + ((long long)blockIdx.x * kEncTileWarps + warp) * kTchLaneGroups;   // the warp's first group
  if (G > nblocks) return;
  const long long b = G - 1 + lane;                                      // the block this lane codes = the group it assembles
  unsigned pl[32];
  int st = 0;
  if (b < nblocks) {
    st = steal[b] ? 1 : 0;
    tch_block_planes(st, d260 + b * kTchD, f184 + b * 184, lsb8msb, crc, pl + 16, pl + 24);
  } else {
#pragma unroll
    for (int i = 16; i < 32; i++) pl[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) pl[i] = __shfl_up_sync(0xffffffffu, pl[16 + i], 1);
  const int pst = __shfl_up_sync(0xffffffffu, st, 1);
  if (lane == 0 || b > nblocks) return;
  enc_store_group<TchTab, true>(pl, sp_base | ((unsigned)pst << 2) | ((unsigned)st << 3), enc_smem + 2048 + (warp * 32 + lane) * 304,
                          bursts + b * 592);                             // Hl = the previous block's flag, Hu = this block's :1365-1366
}
int launch_tch_encode(const unsigned char *d260, const unsigned char *f184, const unsigned char *steal, long long nblocks, int lsb8msb,
                      unsigned tsc_word, int have_tsc, const unsigned char *carry, unsigned char *bursts, int burst_pitch, cudaStream_t st) {
  if (nblocks < 0) return 0;
  if (g_enc_lanes && nblocks >= kTchLaneFirst && burst_pitch == 148 && (reinterpret_cast<uintptr_t>(bursts) & 15) == 0 &&
      ((reinterpret_cast<uintptr_t>(d260) | reinterpret_cast<uintptr_t>(f184)) & 3) == 0) {
    // groups 0 .. 7 (blocks -1 .. 7): the first CTA of the warp-per-block kernel; the rest: lanes
    k_tch_encode<<<1, (kTchGroups + 1) * 32, 0, st>>>(d260, f184, steal, nblocks, lsb8msb, enc_sp_base(tsc_word, have_tsc), carry, bursts, burst_pitch);
    const long long warps = (nblocks - kTchLaneFirst + kTchLaneGroups) / kTchLaneGroups;      // groups kTchLaneFirst .. nblocks
    k_tch_encode_tiles<<<(unsigned)((warps + kEncTileWarps - 1) / kEncTileWarps), kEncTileWarps * 32, kTchTileSmem, st>>>(
        d260, f184, steal, nblocks, lsb8msb, enc_sp_base(tsc_word, have_tsc), bursts);
    return 2;
  }
  const long long groups = nblocks + 1;
  k_tch_encode<<<(unsigned)((groups + kTchGroups - 1) / kTchGroups), (kTchGroups + 1) * 32, 0, st>>>(d260, f184, steal, nblocks, lsb8msb,
                                                                                                  enc_sp_base(tsc_word, have_tsc), carry, bursts,
                                                                                                  burst_pitch);
  return 1;
}
