// resample.cu -- the 65<->96 polyphase resamplers between the 400 kS/s radio stream and the
// 270.833 kS/s (1 sample/symbol) burst stream.
//
// Reference: polyphaseResampleVector, Transceiver/sigProcLib.cpp:1157-1210, as called per chunk by
// RadioInterface::pullBuffer (RX, radioInterface.cpp:238-259: 192 history + 864 new samples -> 715
// outputs, first 130 dropped) and RadioInterface::pushBuffer (TX, :123-168: 130 history + 585 new ->
// 1056 outputs, x13500, (short), first 192 dropped).  The chunk structure is part of the contract:
// the last outputs of every chunk see a filter truncated at the chunk's right edge (the `while
// (inputItr >= wVector.end())` skip, :1183-1186), and the stream's first chunk sees zero history.
//
// Output n of a chunk (o = n + 5, the filter's group delay in output samples, :1177):
//   RX: branch = (96 o) mod 65, in = (96 o - branch)/65, y = sum_k x[in-k] * h[branch + 65 k]
//   TX: branch = (65 o) mod 96, in = (65 o - branch)/96, y = sum_k x[in-k] * h[branch + 96 k]
// accumulated in k order, complex x real, every product and sum rounded separately.
//
// Kernel shape: one CTA per chunk; the 1056 (715) input samples are staged in shared memory with
// coalesced 16-byte loads, the polyphase taps sit in shared memory transposed [k][branch] so that a
// warp's tap reads are spread over the banks; each thread produces outputs n, n+256, ...
#include "kernels.cuh"
#include "sigproc_device.cuh"

namespace btsdsp {

constexpr int kRxIn = 192 + 864, kRxOut = 585, kRxDrop = 130;
constexpr int kTxIn = 130 + 585, kTxOut = 864, kTxDrop = 192;

// cooperative load of `count` samples starting `hist` samples before `in` (zeros when !has_history)
__device__ __forceinline__ void load_chunk(cf *x, const cf *__restrict__ in, int hist, int body, bool has_history) {
  for (int i = threadIdx.x; i < hist + body; i += blockDim.x) {
    cf v = mk(0.0F, 0.0F);
    if (i >= hist || has_history) v = __ldg(in + (i - hist));
    x[i] = v;
  }
}

__global__ void __launch_bounds__(256) k_resample_rx(const DevTables *__restrict__ T, const cf *__restrict__ in,
                                                     int has_history, long long nchunks, cf *__restrict__ out) {
  __shared__ cf x[kRxIn];
  __shared__ float hp[kRxPoly][kRxP + 1];
  for (int i = threadIdx.x; i < kRxPoly * kRxP; i += blockDim.x) hp[i / kRxP][i % kRxP] = T->rx_poly[i % kRxP][i / kRxP];
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    __syncthreads();
    load_chunk(x, in + c * 864, 192, 864, has_history || c > 0);
    __syncthreads();
    for (int m = threadIdx.x; m < kRxOut; m += blockDim.x)
      out[c * kRxOut + m] = resample_at<kRxP, kRxQ, kRxTaps, kRxPoly, kRxP + 1>(x, kRxIn, &hp[0][0], kRxDrop, m);
  }
}
// ------------------------------------------------------------------------------------------------
// RX resampler, tuned kernel.
//
// 585 = 9 x 65 outputs and 864 = 9 x 96 inputs per chunk: the chunked reference loop is periodic with period
// (65 outputs, 96 inputs).  For global period G (= 9*chunk + q) and phase r = 0..64:
//     output 65 G + r  =  sum_{k} raw[96 G - 192 + ix_r - k] * h[br_r + 65 k],   ix_r = (96 (r+135)) / 65,
//                                                                               br_r = (96 (r+135)) % 65,
// with two chunk effects kept exactly: raw indices before the stream start read as zero (no history), and in
// the last period of every chunk (q == 8) phases r >= 60 lose their first ix_r - 287 taps (the reference cannot
// see samples of the next chunk; sigProcLib.cpp:1183-1186).
//
// Mapping: one lane = one period, one warp = 32 consecutive periods, ALL 65 phases per lane.  Because the phase
// is the same across a warp, every tap is a warp-uniform constant: the taps live in __constant__ memory and are
// read as immediate c[bank][offset] operands of the multiplies (no tap loads at all), and the whole 65-phase
// body is unrolled into straight-line code with compile-time sample offsets.  Phases are processed in groups of
// five adjacent outputs whose input windows overlap (15 + ~6 samples), so a lane loads ~22 samples per five
// outputs as aligned 16-byte pairs instead of 75.
// Shared memory per warp: the 32-period input tile, stored in rows of 96 samples padded to 98 (lane stride
// 98 samples = 49 x 16 B, odd in 16-byte units -> conflict-free LDS.128), and a 32 x 65 output tile (lane
// stride 65, odd -> conflict-free) that is written back as one contiguous, fully coalesced 16.6 KB block.
// ------------------------------------------------------------------------------------------------
__constant__ float c_rx_poly[kRxP * 16];          // [r][k] = lpf_rx[br_r + 65 k], zero past the end

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc, bool valid) {   // zero-fills when !valid
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16z(void *smem_dst, const void *gsrc, bool valid) {  // zero-fills when !valid
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

#ifndef BTS_RXV2_WARPS
#define BTS_RXV2_WARPS 4
#endif
#ifndef BTS_RXV2_TILES
#define BTS_RXV2_TILES 2
#endif
// A CTA works on kRxV2Tiles consecutive 32-period tiles at once (one super-tile of 32*T + 2 input rows); warp w does
// phase part w % kRxV2Warps of tile w / kRxV2Warps.  With kRxV2Warps = 4 the warps that share an SM sub-partition
// (w % 4) all run the SAME quarter of the unrolled phase code, released together by the barrier, so they walk the
// instruction stream in near lock-step and share its fetches.
constexpr int kRxV2Warps = BTS_RXV2_WARPS, kRxV2Tiles = BTS_RXV2_TILES, kRxV2Threads = 32 * kRxV2Warps * kRxV2Tiles;
constexpr int kRxV2Periods = 32 * kRxV2Tiles, kRxV2Rows = kRxV2Periods + 2;
constexpr int kRxV2In = kRxV2Rows * kRxRowPitch, kRxV2Out = kRxV2Periods * kRxP;
constexpr size_t kRxV2Smem = (size_t)(kRxV2In + kRxV2Out) * sizeof(cf);
constexpr int kRxV2CtasPerSm = (int)(227 * 1024 / (kRxV2Smem + 1024)) < 2048 / kRxV2Threads ? (int)(227 * 1024 / (kRxV2Smem + 1024)) : 2048 / kRxV2Threads;

// I16 = the radio's own sample format: interleaved int16 {I,Q} pairs (4 B per sample, `swap_iq` for the Q-first
// order of real USRP hardware), converted exactly as unUSRPifyVector does (radioInterface.cpp:91-116) while the
// tile is built -- the raw samples land in the (still unused) output tile via cp.async and are expanded to float.
template <bool I16>
__global__ void __launch_bounds__(kRxV2Threads) k_resample_rx_v2(const void *__restrict__ in_, int has_history, int swap_iq,
                                                       long long nperiods, long long nsamples, cf *__restrict__ out) {
  const cf *in = reinterpret_cast<const cf *>(in_);
  const short2 *in16 = reinterpret_cast<const short2 *>(in_);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *xt = reinterpret_cast<cf *>(smem_raw);
  cf *ot = xt + kRxV2In;
  const int warp = threadIdx.x >> 5, part = warp % kRxV2Warps;
  const int row = (warp / kRxV2Warps) * 32 + (threadIdx.x & 31);      // this lane's period within the super-tile
  for (long long tile = blockIdx.x; tile * kRxV2Periods < nperiods; tile += gridDim.x) {
    const long long G0 = tile * kRxV2Periods;
    const long long raw0 = 96 * G0 - 96;                               // tile origin: sample (G, r, k) sits at 96*l + ix_r - k - 96
    // ---- load 34 rows x 96 samples as 16-byte cp.async copies (global -> shared, no registers, all 51 per lane
    //      in flight at once); samples outside [lo, nsamples) are zero-filled by the copy's src-size operand
    const long long lo = has_history ? -192 : 0;
    __syncthreads();                                                    // previous tile fully written back
    if (!I16) {
      for (int i4 = threadIdx.x; i4 < kRxV2Rows * 48; i4 += kRxV2Threads) {
        const int rw = i4 / 48, c4 = i4 - rw * 48;
        const long long s = raw0 + (long long)rw * 96 + 2 * c4;
        cf *dst = xt + rw * kRxRowPitch + 2 * c4;
        if (s >= lo && s + 1 < nsamples) cp_async16(dst, in + s);
        else {
          cp_async8(dst, in + (s >= lo && s < nsamples ? s : 0), s >= lo && s < nsamples);
          cp_async8(dst + 1, in + (s + 1 >= lo && s + 1 < nsamples ? s + 1 : 0), s + 1 >= lo && s + 1 < nsamples);
        }
      }
      cp_async_wait_all();
      __syncthreads();
    } else {
      // raw0, lo and nsamples are multiples of 4 samples, so every 16-byte group is wholly inside or outside
      short2 *stage = reinterpret_cast<short2 *>(ot);
      for (int i = threadIdx.x; i < kRxV2Rows * 24; i += kRxV2Threads) {
        const long long s = raw0 + 4LL * i;
        const bool valid = s >= lo && s < nsamples;
        cp_async16z(stage + 4 * i, in16 + (valid ? s : 0), valid);
      }
      cp_async_wait_all();
      __syncthreads();
      for (int i = threadIdx.x; i < kRxV2Rows * 24; i += kRxV2Threads) {
        const int rw = i / 24, c = 4 * (i - rw * 24);
        const int4 v = *reinterpret_cast<const int4 *>(stage + 4 * i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        cf *dst = xt + rw * kRxRowPitch + c;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float a = (float)(short)(w[k] & 0xffff), b = (float)(short)(w[k] >> 16);     // first, second int16
          dst[k] = swap_iq ? mk(b, a) : mk(a, b);
        }
      }
      __syncthreads();
    }
    // ---- 65 phases for this lane's period
    const long long G = G0 + row;
    const bool q8 = (G % 9) == 8;
    const cf *xl = xt + row * kRxRowPitch;
    cf *ol = ot + row * kRxP;
    rx_part<kRxV2Warps>(part, c_rx_poly, xl, ol, q8);
    __syncthreads();
    // ---- write the 32 x 65 outputs back: contiguous in both shared and global memory
    const long long nvalid = (nperiods - G0 < kRxV2Periods ? nperiods - G0 : kRxV2Periods) * kRxP;
    cf *og = out + G0 * kRxP;
    for (int i4 = threadIdx.x; i4 < kRxV2Out / 2; i4 += kRxV2Threads) {
      if (2 * i4 + 1 < nvalid) *reinterpret_cast<float4 *>(og + 2 * i4) = *reinterpret_cast<const float4 *>(ot + 2 * i4);
      else if (2 * i4 < nvalid) og[2 * i4] = ot[2 * i4];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// v3: one persistent CTA per SM, T tiles per step, software-pipelined.
// v2 loses a third of its issue slots waiting for instruction fetches: the unrolled 65-phase body is ~60 KB and
// five independent CTAs per SM walk it at unrelated places.  Here a CTA owns T consecutive tiles; warp w runs
// phase quarter w % 4 of tile w / 4, so the warps of one SM sub-partition all run the SAME quarter and are released
// together by the barrier -- they walk the code in near lock-step and share its fetches.  What the single CTA
// loses in load/compute overlap is put back explicitly:
//   float input : two input buffers; the cp.async loads of super-tile s+1 are in flight while s is computed;
//   int16 input : the raw int16 rows of s+1 land in a staging buffer while s is computed, and are widened to
//                 float (unUSRPifyVector, radioInterface.cpp:94-110) into the single float tile at the next step;
//   output      : the finished 32T x 65 block is contiguous in shared AND global memory and leaves as ONE bulk
//                 async copy (cp.async.bulk, issued by one thread) that overlaps the next step.
// ------------------------------------------------------------------------------------------------
template <bool I16, int T>
struct RxV3 {
  static constexpr int kPeriods = 32 * T, kRows = kPeriods + 2;
  static constexpr int kIn = kRows * kRxRowPitch, kOut = kPeriods * kRxP;         // samples
  static constexpr int kThreads = 128 * T + 32;                                // 4T compute warps + the producer warp
  static constexpr size_t kSmem = I16 ? (size_t)(kIn + kOut) * sizeof(cf) + (size_t)2 * kRows * 96 * 4
                                      : (size_t)(2 * kIn + kOut) * sizeof(cf);
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\ncp.async.bulk.commit_group;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarriers between the producer warp (cp.async loads) and the compute warps
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrives (without raising the expected count) once all of this thread's earlier cp.async copies have landed
__device__ __forceinline__ void mbar_arrive_after_cp_async(unsigned long long *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
template <int NTHREADS>
__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

// Producer warp: issue (not wait for) the loads of the super-tile whose first raw sample is raw0.
// Samples outside [lo, nsamples) are zero-filled by the copies' src-size operand.
template <int ROWS>
__device__ __forceinline__ void rxv3_load_f32(cf *xt, const cf *__restrict__ in, long long raw0, long long lo, long long nsamples) {
  const int lane = threadIdx.x & 31;
  if (raw0 >= lo && raw0 + (long long)ROWS * 96 <= nsamples) {                      // interior: no bounds tests
    const cf *src = in + raw0 + 2 * lane;
    cf *dst = xt + 2 * lane;
#pragma unroll 2
    for (int rw = 0; rw < ROWS; rw++) {
      cp_async16(dst + rw * kRxRowPitch, src + rw * 96);
      if (lane < 16) cp_async16(dst + rw * kRxRowPitch + 64, src + rw * 96 + 64);
    }
  } else {
    for (int i4 = lane; i4 < ROWS * 48; i4 += 32) {
      const int rw = i4 / 48, c4 = i4 - rw * 48;
      const long long s = raw0 + (long long)rw * 96 + 2 * c4;
      const bool ok = s >= lo && s < nsamples;                                     // bounds are even: both samples or none
      cp_async16z(xt + rw * kRxRowPitch + 2 * c4, in + (ok ? s : 0), ok);
    }
  }
}
template <int ROWS>
__device__ __forceinline__ void rxv3_load_i16(short2 *stage, const short2 *__restrict__ in16, long long raw0, long long lo,
                                              long long nsamples) {
  const int lane = threadIdx.x & 31;
  if (raw0 >= lo && raw0 + (long long)ROWS * 96 <= nsamples) {
    const short2 *src = in16 + raw0 + 4 * lane;
    short2 *dst = stage + 4 * lane;
#pragma unroll 4
    for (int i = 0; i < ROWS * 24 / 32; i++) cp_async16(dst + 128 * i, src + 128 * i);
    if (lane < ROWS * 24 % 32) cp_async16(dst + 128 * (ROWS * 24 / 32), src + 128 * (ROWS * 24 / 32));
  } else {
    for (int i = lane; i < ROWS * 24; i += 32) {
      const long long s = raw0 + 4LL * i;
      const bool ok = s >= lo && s < nsamples;
      cp_async16z(stage + 4 * i, in16 + (ok ? s : 0), ok);
    }
  }
}
template <int ROWS, int NTHREADS>
__device__ __forceinline__ void rxv3_widen(cf *xt, const short2 *stage, int swap_iq) {
  for (int i = threadIdx.x; i < ROWS * 24; i += NTHREADS) {
    const int rw = i / 24, c = 4 * (i - rw * 24);
    const int4 v = *reinterpret_cast<const int4 *>(stage + 4 * i);
    const int w[4] = {v.x, v.y, v.z, v.w};
    cf *dst = xt + rw * kRxRowPitch + c;
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const float a0 = (float)(short)(w[k] & 0xffff), b0 = (float)(short)(w[k] >> 16);       // first, second int16
      const float a1 = (float)(short)(w[k + 1] & 0xffff), b1 = (float)(short)(w[k + 1] >> 16);
      *reinterpret_cast<float4 *>(dst + k) = swap_iq ? make_float4(b0, a0, b1, a1) : make_float4(a0, b0, a1, b1);
    }
  }
}

template <bool I16, int T>
__global__ void __launch_bounds__(128 * T + 32, 1) k_resample_rx_v3(const void *__restrict__ in_, int has_history, int swap_iq,
                                                                   long long nperiods, long long nsamples, cf *__restrict__ out) {
  using C = RxV3<I16, T>;
  constexpr int NC = 128 * T;                                            // compute threads; warp 4T is the producer
  const cf *in = reinterpret_cast<const cf *>(in_);
  const short2 *in16 = reinterpret_cast<const short2 *>(in_);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *ot = reinterpret_cast<cf *>(smem_raw);                             // output block first: 16-byte aligned for the bulk copy
  cf *xbuf = ot + C::kOut;                                               // float: two input buffers; int16: one + 2 staging
  short2 *stage = reinterpret_cast<short2 *>(xbuf + C::kIn);
  __shared__ __align__(8) unsigned long long full[2], empty[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long lo = has_history ? -192 : 0;
  const long long ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 32);
    mbar_init(&full[1], 32);
    mbar_init(&empty[0], 4 * T);
    mbar_init(&empty[1], 4 * T);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 4 * T) {
    // ---- producer: keeps up to two super-tiles in flight ahead of the compute warps
    int s = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, s++) {
      const int b = s & 1;
      if (s >= 2) mbar_wait(&empty[b], ((s >> 1) - 1) & 1);              // the compute warps are done with this buffer
      const long long raw0 = 96 * tile * C::kPeriods - 96;               // sample (G, r, k) sits at 96*l + ix_r - k - 96
      if (I16) rxv3_load_i16<C::kRows>(stage + b * (C::kRows * 96), in16, raw0, lo, nsamples);
      else rxv3_load_f32<C::kRows>(xbuf + b * C::kIn, in, raw0, lo, nsamples);
      mbar_arrive_after_cp_async(&full[b]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }
  // ---- compute warps: warp w runs phase quarter w % 4 of tile w / 4
  const int part = warp & 3;
  const int row = (warp >> 2) * 32 + lane;                               // this lane's period within the super-tile
  bool store_pending = false;
  int s = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, s++) {
    const long long G0 = tile * C::kPeriods;
    const int b = s & 1;
    cf *xt = xbuf;
    if (I16) {
      mbar_wait(&full[b], (s >> 1) & 1);                                 // raw int16 rows have landed
      rxv3_widen<C::kRows, NC>(xt, stage + b * (C::kRows * 96), swap_iq);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
      if (threadIdx.x == 0 && store_pending) bulk_store_wait_read();
      compute_warps_sync<NC>();                                          // xt ready, ot free
    } else {
      xt = xbuf + b * C::kIn;
      if (threadIdx.x == 0 && store_pending) bulk_store_wait_read();
      compute_warps_sync<NC>();                                          // ot free
      mbar_wait(&full[b], (s >> 1) & 1);                                 // this step's tile has landed
    }
    const long long G = G0 + row;
    const bool q8 = (G % 9) == 8;
    rx_part<4>(part, c_rx_poly, xt + row * kRxRowPitch, ot + row * kRxP, q8);
    if (!I16) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }
    // ---- the 32T x 65 outputs are contiguous in shared and global memory
    const long long nvalid = (nperiods - G0 < C::kPeriods ? nperiods - G0 : C::kPeriods) * kRxP;
    cf *og = out + G0 * kRxP;
    if (nvalid == C::kOut) {
      fence_async_smem();
      compute_warps_sync<NC>();                                          // all outputs written; all reads of xt done
      if (threadIdx.x == 0) bulk_store(og, ot, (unsigned)(C::kOut * sizeof(cf)));
      store_pending = true;
    } else {
      compute_warps_sync<NC>();
      for (int i = threadIdx.x; i < nvalid; i += NC) og[i] = ot[i];
      store_pending = false;
      compute_warps_sync<NC>();
    }
  }
  if (threadIdx.x == 0 && store_pending) bulk_store_wait_all();
}

#ifndef BTS_RXV3_TILES_F32
#define BTS_RXV3_TILES_F32 3
#endif
#ifndef BTS_RXV3_TILES_I16
#define BTS_RXV3_TILES_I16 3
#endif
constexpr int kRxV3TilesF32 = BTS_RXV3_TILES_F32, kRxV3TilesI16 = BTS_RXV3_TILES_I16;
static int g_num_sms = 148;

void upload_resampler_taps(const DevTables *hostT) {
  float h[kRxP * 16];
  rx_fill_taps(hostT, h);
  cudaMemcpyToSymbol(c_rx_poly, h, sizeof h);
}

void launch_resample_rx(const DevTables *T, const cf *in, int has_history, long long nchunks, cf *out, cudaStream_t st) {
  if (nchunks <= 0) return;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (aligned) {
    using C = RxV3<false, kRxV3TilesF32>;
    const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
    const unsigned grid = (unsigned)(ntiles < g_num_sms ? ntiles : g_num_sms);
    k_resample_rx_v3<false, kRxV3TilesF32><<<grid, C::kThreads, C::kSmem, st>>>(in, has_history, 0, nperiods, nchunks * 864, out);
  } else {
    const unsigned grid = (unsigned)(nchunks < 148 * 32 ? nchunks : 148 * 32);
    k_resample_rx<<<grid, 256, 0, st>>>(T, in, has_history, nchunks, out);
  }
}
// int16 {I,Q} ingest (the radio's format); `in` must be 16-byte aligned
int launch_resample_rx_i16(const int16_t *in, int swap_iq, int has_history, long long nchunks, cf *out, cudaStream_t st) {
  if (nchunks <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
  using C = RxV3<true, kRxV3TilesI16>;
  const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
  const unsigned grid = (unsigned)(ntiles < g_num_sms ? ntiles : g_num_sms);
  k_resample_rx_v3<true, kRxV3TilesI16><<<grid, C::kThreads, C::kSmem, st>>>(in, has_history, swap_iq, nperiods, nchunks * 864, out);
  return 1;
}
int configure_resamplers() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_num_sms = sms;
  cudaError_t e = cudaFuncSetAttribute(k_resample_rx_v3<false, kRxV3TilesF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)RxV3<false, kRxV3TilesF32>::kSmem);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFuncSetAttribute(k_resample_rx_v3<true, kRxV3TilesI16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)RxV3<true, kRxV3TilesI16>::kSmem);
}

// TX: also applies the x13500 scaling and int16 truncation (tx_quantise).
__global__ void __launch_bounds__(256) k_resample_tx(const DevTables *__restrict__ T, const cf *__restrict__ in,
                                                     int has_history, long long nchunks, int16_t *__restrict__ out) {
  __shared__ cf x[kTxIn];
  __shared__ float hp[kTxPoly][kTxP + 1];
  for (int i = threadIdx.x; i < kTxPoly * kTxP; i += blockDim.x) hp[i / kTxP][i % kTxP] = T->tx_poly[i % kTxP][i / kTxP];
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    __syncthreads();
    load_chunk(x, in + c * 585, 130, 585, has_history || c > 0);
    __syncthreads();
    for (int m = threadIdx.x; m < kTxOut; m += blockDim.x)
      reinterpret_cast<short2 *>(out)[c * kTxOut + m] =
          tx_quantise(resample_at<kTxP, kTxQ, kTxTaps, kTxPoly, kTxP + 1>(x, kTxIn, &hp[0][0], kTxDrop, m));
  }
}
void launch_resample_tx(const DevTables *T, const cf *in, int has_history, long long nchunks, int16_t *out,
                        cudaStream_t st) {
  if (nchunks <= 0) return;
  const unsigned grid = (unsigned)(nchunks < 148 * 32 ? nchunks : 148 * 32);
  k_resample_tx<<<grid, 256, 0, st>>>(T, in, has_history, nchunks, out);
}

// polyphaseResampleVector on one arbitrary vector (the sigProcLib.h entry point), one thread per output.
__global__ void k_resample_generic(const cf *__restrict__ x, int n, int P, int Q, const float *__restrict__ lpf, int L,
                                   cf *__restrict__ out, int outn) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= outn) return;
  const int outputIx = (L - 1) / 2 / Q + o;
  const int br = (outputIx * Q) % P;
  int ix = (outputIx * Q - br) / P;
  int f = br;
  while (ix >= n) { ix--; f += P; }
  cf sum = mk(0.0F, 0.0F);
  while (ix >= 0 && f < L) { sum = cadd(sum, cmulr(x[ix], lpf[f])); ix--; f += P; }
  out[o] = sum;
}
void launch_resample_generic(const cf *x, int n, int P, int Q, const float *lpf, int L, cf *out, int outn,
                             cudaStream_t st) {
  if (outn <= 0) return;
  k_resample_generic<<<(outn + 127) / 128, 128, 0, st>>>(x, n, P, Q, lpf, L, out, outn);
}

}  // namespace btsdsp
