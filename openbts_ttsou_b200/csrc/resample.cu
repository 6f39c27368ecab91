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
// Kernels: k_resample_rx_v3 (RX, the hot one: persistent warp-specialised CTA per SM, TMA tensor copies in, bulk copies
// out, lane = 65-output period, straight-line 65-phase body -- described where it is defined) and k_tx_fused (the whole
// TX chain, same shape, its input tile computed from the bits).  k_resample_rx / k_resample_tx are the plain
// one-CTA-per-chunk forms (chunk staged in shared memory, taps transposed [k][branch], a thread per output): the
// fallback for unaligned pointers and the stand-alone TX resampler entry point.
#include <stdlib.h>
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through the runtime, no -lcuda)

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
// Mapping: one lane = one period, one warp = 32 consecutive periods (a "tile").  Because the phase is the same
// across a warp, every tap is warp-uniform and the 65-phase body is unrolled into straight-line code with
// compile-time sample offsets.  Phases are processed in groups of 4-6 adjacent outputs whose input windows overlap
// (15 + ~6 samples), so a lane loads ~22 samples per five outputs as aligned 16-byte pairs instead of 75.  Each tap
// costs TWO instructions: FFMA2 (x * h + 0, both components) and FADD2 (see cplx.cuh pmul0/padd) -- bit-identical
// to the reference's separately rounded multiply and add.
// Shared memory: the input tile in rows of 96 samples at a 98-sample pitch (lane stride 49 x 16 B, odd ->
// conflict-free LDS.128), and a 32 x 65 output tile (lane stride 65, odd -> conflict-free) that leaves as one
// contiguous, fully coalesced block.
// ------------------------------------------------------------------------------------------------
__constant__ float c_rx_poly[kRxP * 16];          // [r][k] = lpf_rx[br_r + 65 k], zero past the end

// ------------------------------------------------------------------------------------------------
// k_resample_rx_v3: one persistent, warp-specialised CTA per SM.
//   * T consecutive tiles per step (a "super-tile" of 32T periods, 32T + 2 input rows).
//   * mover: ONE thread drives the copy engine in both directions.  In: ONE 2-D tensor (TMA) copy per super-tile into a
//     ring of input buffers, completion counted on a `full` mbarrier.  The tensor map describes the stream as rows of 96
//     samples and the box is 98 samples wide, so the copy engine itself writes the rows at the conflict-free pitch; rows
//     outside the stream arrive as zeros (= the reference's zero history and nothing past the end).  Out: the finished
//     32T x 65 block is contiguous in shared AND global memory and leaves as one bulk async copy (cp.async.bulk).
//   * compute warps: warp w runs phase part w % P of tile w / P.  Warps of one SM sub-partition (w % 4) run the same part
//     of the unrolled code, so they share its instruction fetches.  The taps come from shared memory (uniform-address
//     LDS.128): as constant-bank operands the 4 KB table thrashes the per-SM constant cache and every miss stalls a warp
//     for ~100 cycles.
//   * float input: no CTA barrier in the loop.  A warp that has delivered its part arrives on `outfull`; when all have, the
//     mover issues the store, THEN the refill of the input slot that step read (the copy engine serves its queue in order:
//     a 77 KB refill queued ahead of the store kept the output block busy ~1 us longer), waits until the block has been
//     read out of shared memory and reports it on `drained`.  A compute warp starts the next step as soon as its input has
//     landed and checks `drained` only before its stores (rx_range's gate: a part keeps its sums in registers until
//     then), so the block's drain time is covered by arithmetic.  Per-step tile bookkeeping is 32-bit and division-free (RxTileIter): three 64-bit
//     divisions per step used to sit on every warp's critical path.
//   * int16 input (the radio's format): the raw rows land in a staging ring and are widened to float
//     (unUSRPifyVector, radioInterface.cpp:94-110) into the single float tile at the start of each step; this path keeps
//     the barrier-per-step flow (producer thread + full/empty mbarriers, the compute warps meet on a named barrier).
// History (profiles/README.md, r2): five independent 2-warp CTAs per SM 0.714 ms -> lock-step quarters 0.65 ->
// 2-instruction taps 0.615 -> producer warp 0.52 -> shared-memory taps + tensor copy 0.553 -> barrier-free flow, store
// ahead of the refill, division-free bookkeeping 0.490 ms (213 750 chunks; the copy pipeline alone, arithmetic removed,
// runs 0.399 ms = 0.94 of the measured HBM peak: -DBTS_RXV3_NOMATH).
// tools/rxv3_trace.cu prints per-step timestamps of one CTA (build with -DBTS_RXV3_TRACE).
// ------------------------------------------------------------------------------------------------
#ifndef BTS_RXV3_PARTS
#define BTS_RXV3_PARTS 8
#endif
constexpr int kRxV3Parts = BTS_RXV3_PARTS;         // warps per 32-period tile (each does 65/kRxV3Parts phases)
#ifndef BTS_RXV3_RING
#define BTS_RXV3_RING 2
#endif
constexpr int kRxV3Ring = BTS_RXV3_RING;           // input buffers in flight (float tiles, or int16 staging rows)
#ifndef BTS_RXV3_OBUFS
#define BTS_RXV3_OBUFS 1
#endif
constexpr int kRxV3Obufs = BTS_RXV3_OBUFS;         // output blocks: with one, every step waits for the previous block to drain
// (An experiment that split the 65 phases over pairs of CTAs, so each SM runs half the unrolled code out of its
// instruction cache, measured 0.92 ms against 0.50 ms: every input byte then crosses the L2 -> SM fabric twice and the
// per-GPC fabric becomes the limit -- profiles/README.md r2.  It is not kept in the source.)
#ifndef BTS_RXV3_DECOUPLED
#define BTS_RXV3_DECOUPLED 1
#endif
// float input: the compute warps never meet at a CTA barrier -- the mover thread issues the output block's bulk copy when the
// last warp has delivered its part, refills the input slot and reports when the block has been read out; each compute
// warp checks that report only before the stores of its next step (see the kernel)
constexpr bool kRxV3Decoupled = BTS_RXV3_DECOUPLED != 0;
template <bool I16, int T>
struct RxV3 {
  static constexpr int kPeriods = 32 * T, kRows = kPeriods + 2;
  static constexpr int kIn = kRows * kRxRowPitch, kOut = kPeriods * kRxP;        // samples
  static constexpr int kInSlot = (kIn + 15) & ~15;                               // ring slots start on 128-byte lines (TMA)
  static constexpr bool kDec = !I16 && kRxV3Decoupled;
  static constexpr int kThreads = 32 * kRxV3Parts * T + 32;                      // compute warps + the warp that drives the copy engine
  static constexpr int kStage = kRows * 96;                                      // int16 pairs per staging buffer
  static constexpr unsigned kTileBytes = I16 ? kStage * 4u : kIn * 8u;           // what one tensor copy delivers
  // output block; then float: kRxV3Ring input tiles / int16: one float tile + kRxV3Ring staging buffers
  static constexpr size_t kSmem = 128 + (I16 ? (size_t)(kRxV3Obufs * kOut + kInSlot) * sizeof(cf) + (size_t)kRxV3Ring * kStage * 4
                                             : (size_t)(kRxV3Obufs * kOut + kRxV3Ring * kInSlot) * sizeof(cf));
  static_assert((kOut * sizeof(cf)) % 128 == 0 && (kStage * 4) % 128 == 0, "TMA destinations must be 128-byte aligned");
  static_assert(kSmem + 6 * 1024 <= 227 * 1024, "resampler tile does not fit in shared memory");
};

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\ncp.async.bulk.commit_group;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
template <int PENDING>           // returns when at most PENDING of the newest bulk stores are still reading shared memory
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
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
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// A persistent CTA's walk over the super-tiles (tile t = super-tile t % tps of stream t / tps, t = first, first + stride, ...)
// without a 64-bit division per step: those sat on every warp's critical path, several hundred cycles per step.
struct RxTileIter {
  unsigned tl, tps, stride;      // super-tile within the stream; super-tiles per stream; grid size
  int st;                        // stream
  __device__ __forceinline__ RxTileIter(unsigned first, unsigned tps_, unsigned stride_)
      : tl(first % tps_), tps(tps_), stride(stride_), st((int)(first / tps_)) {}
  __device__ __forceinline__ void next() {
    tl += stride;
    if (tl >= tps) { const unsigned q = tl / tps; tl -= q * tps; st += (int)q; }
  }
  // is period (tl * PERIODS + row) the last of its 9-period chunk?
  template <int PERIODS>
  __device__ __forceinline__ bool q8(int row) const { return ((tl % 9u) * (unsigned)(PERIODS % 9) + (unsigned)row) % 9u == 8u; }
};
// rx_range's gate: wait (once per step) until the previous step's output block has been read out of shared memory
struct RxDrainGate {
  unsigned long long *bar;
  unsigned parity;
  bool on;
  __device__ __forceinline__ void operator()() const { if (on) mbar_wait(bar, parity); }
};
template <int NTHREADS>
__device__ __forceinline__ void compute_warps_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// One tensor (TMA) copy brings in a whole super-tile.  The tensor map describes the raw streams as [stream][row][96 samples];
// the box is ROWS x 98 samples for float input -- two samples wider than the tensor, so the copy itself lays the rows
// down at the 98-sample pitch that keeps the compute warps' LDS.128 conflict-free (the two extra columns are
// out-of-bounds and arrive as zeros) -- and ROWS x 96 int16 pairs for int16 input.  Rows before the start of a
// history-less stream and past its end are out of bounds too and arrive as zeros, which is exactly the reference's
// zero history / truncated right edge.  One thread issues it; completion is counted in bytes on the mbarrier.
__device__ __forceinline__ void tma_load_tile(void *sdst, const CUtensorMap *tmap, int row0, int stream, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(sdst)),
               "l"(tmap), "r"(0), "r"(row0), "r"(stream), "r"(smem_u32(bar))
               : "memory");
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

#ifdef BTS_RXV3_TRACE            // tools/rxv3_trace.cu: per-iteration timestamps of CTA 0 (debug builds only)
__device__ long long g_rx_trace[2][64][8];
__device__ long long g_rx_cta[256][4];
#ifndef BTS_RXV3_TRACE_CTA
#define BTS_RXV3_TRACE_CTA 0
#endif
#define RX_TRACE(who, it, slot) do { if (blockIdx.x == BTS_RXV3_TRACE_CTA && (who == 1 || threadIdx.x == 0) && (it) < 64) g_rx_trace[who][it][slot] = clock64(); } while (0)
#else
#define RX_TRACE(who, it, slot) do { } while (0)
#endif
template <bool I16, int T>
__global__ void __launch_bounds__(32 * kRxV3Parts * T + 32, 1) k_resample_rx_v3(const __grid_constant__ CUtensorMap tmap, int row_bias, int swap_iq,
                                                                             long long nperiods, cf *__restrict__ out, int nstreams,
                                                                             long long out_pitch) {
  using C = RxV3<I16, T>;
  constexpr int NW = kRxV3Parts * T, NC = 32 * NW, R = kRxV3Ring;        // compute warps / threads; warp NW is the producer
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *obuf = reinterpret_cast<cf *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);   // 128-byte lines
  cf *xbuf = obuf + kRxV3Obufs * C::kOut;
  short2 *stage = reinterpret_cast<short2 *>(xbuf + C::kInSlot);
  constexpr bool DEC = C::kDec;
  __shared__ __align__(8) unsigned long long full[R], empty[R], outfull, drained;
  // the taps are read from shared memory (uniform-address LDS.128, ~30 cycles) rather than as constant-bank operands:
  // 4 KB of taps cycled through by 12+ warps misses the small per-SM constant cache and each miss stalls a warp ~100s
  // of cycles on the LDCU that feeds its next four multiplies
  __shared__ __align__(16) float s_taps[kRxP * 16];
  for (int i = threadIdx.x; i < kRxP * 16; i += blockDim.x) s_taps[i] = c_rx_poly[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // nstreams independent streams of nperiods periods each; tile t = super-tile t % tps of stream t / tps
  const long long tps = (nperiods + C::kPeriods - 1) / C::kPeriods, ntiles = tps * nstreams;
  const long long first = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < R; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NW);
    }
    mbar_init(&outfull, NW);
    mbar_init(&drained, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NW) {
    if (lane != 0) return;
    if constexpr (DEC) {
      // ---- mover (float input): ONE thread runs both directions of the copy engine.  R super-tiles are requested up front;
      //      then, per step, when every compute warp has delivered its part: the output block's bulk copy, the refill of the
      //      input slot that step read (after the store, so the copy engine sees the store first), and -- once the block has
      //      been read out of shared memory -- the `drained` report the compute warps check before their next first store.
      RxTileIter tl_((unsigned)first, (unsigned)tps, (unsigned)stride);      // the tile whose input is requested next
      long long ltile = first;
      for (int i = 0; i < R && ltile < ntiles; i++, ltile += stride, tl_.next()) {
        mbar_expect_tx(&full[i], C::kTileBytes);
        tma_load_tile(xbuf + i * C::kInSlot, &tmap, (int)(tl_.tl * C::kPeriods) - 1 + row_bias, tl_.st, &full[i]);
      }
      int b = 0, it = 0;
      RxTileIter ti((unsigned)first, (unsigned)tps, (unsigned)stride);
      for (long long tile = first; tile < ntiles; tile += stride, it++, ti.next()) {
        const long long G0 = (long long)ti.tl * C::kPeriods;
        RX_TRACE(1, it, 0);
        mbar_wait(&outfull, it & 1);
        RX_TRACE(1, it, 1);
        const bool whole = nperiods - G0 >= C::kPeriods;
        if (whole) bulk_store(out + ti.st * out_pitch + G0 * kRxP, obuf, (unsigned)(C::kOut * sizeof(cf)));
        if (ltile < ntiles) {                                              // every warp has left slot b: refill it
          mbar_expect_tx(&full[b], C::kTileBytes);
          tma_load_tile(xbuf + b * C::kInSlot, &tmap, (int)(tl_.tl * C::kPeriods) - 1 + row_bias, tl_.st, &full[b]);
          ltile += stride;
          tl_.next();
        }
        RX_TRACE(1, it, 2);
        if (++b == R) b = 0;
        if (whole) bulk_store_wait_read<0>();                              // (a ragged block is copied out by the compute warps)
        mbar_arrive(&drained);                                             // one phase per step
      }
      bulk_store_wait_all();
    } else {
      // ---- producer: one thread keeps up to R super-tiles in flight ahead of the compute warps
      int b = 0, use = 0, it = 0;                                        // ring slot and how many times it has been used
      RxTileIter ti((unsigned)first, (unsigned)tps, (unsigned)stride);
      for (long long tile = first; tile < ntiles; tile += stride, it++, ti.next()) {
        RX_TRACE(1, it, 0);
        if (use > 0) mbar_wait(&empty[b], (use - 1) & 1);                // the compute warps are done with this slot
        RX_TRACE(1, it, 1);
        // sample (G, r, k) sits at 96*l + ix_r - k - 96 from the tile origin: the tile starts one row before period G0
        mbar_expect_tx(&full[b], C::kTileBytes);
        tma_load_tile(I16 ? (void *)(stage + b * C::kStage) : (void *)(xbuf + b * C::kInSlot), &tmap,
                      (int)(ti.tl * C::kPeriods) - 1 + row_bias, ti.st, &full[b]);
        RX_TRACE(1, it, 2);
        if (++b == R) { b = 0; use++; }
      }
    }
    return;
  }
  // ---- compute warps: warp w runs phase part w % kRxV3Parts of tile w / kRxV3Parts
  const int part = warp % kRxV3Parts;
  const int tl = warp / kRxV3Parts;
  const int row = tl * 32 + lane;                                        // this lane's period within the super-tile
  int b = 0, use = 0;
  bool store_pending = false;
  int it = 0;
#ifdef BTS_RXV3_TRACE
  if (threadIdx.x == 0) g_rx_cta[blockIdx.x][0] = clock64();
#endif
  if constexpr (DEC) {
    RxTileIter ti((unsigned)first, (unsigned)tps, (unsigned)stride);
    for (long long tile = first; tile < ntiles; tile += stride, it++, ti.next()) {
      const long long G0 = (long long)ti.tl * C::kPeriods;
      RX_TRACE(0, it, 0);
      mbar_wait(&full[b], use & 1);                                      // this step's input has landed
      RX_TRACE(0, it, 1);
      const cf *xt = xbuf + b * C::kInSlot;
      const bool q8 = ti.template q8<C::kPeriods>(row);
      RX_TRACE(0, it, 2);
      rx_part<kRxV3Parts, 0, RxDrainGate>(part, s_taps, xt + row * kRxRowPitch, obuf + row * kRxP, q8,
                                           RxDrainGate{&drained, (unsigned)(it - 1) & 1u, it > 0});
      RX_TRACE(0, it, 3);
      fence_async_smem();                                                // outputs visible to the copy engine
      __syncwarp();
      if (lane == 0) mbar_arrive(&outfull);                              // this warp's part is in the block; its reads of xt are done
      if (++b == R) { b = 0; use++; }
      const long long nper = nperiods - G0 < C::kPeriods ? nperiods - G0 : C::kPeriods;
      RX_TRACE(0, it, 4);
      if (nper != C::kPeriods) {                                         // ragged end of a stream: plain stores by all compute warps
        mbar_wait(&outfull, it & 1);
        cf *og = out + ti.st * out_pitch + G0 * kRxP;
        for (int i = threadIdx.x; i < nper * kRxP; i += NC) og[i] = obuf[i];
        compute_warps_sync<NC>();                                        // nobody writes the block before everybody has read it
      }
      RX_TRACE(0, it, 5);
    }
#ifdef BTS_RXV3_TRACE
    if (threadIdx.x == 0) { g_rx_cta[blockIdx.x][1] = clock64(); g_rx_cta[blockIdx.x][2] = it; unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); g_rx_cta[blockIdx.x][3] = sm; }
#endif
    return;
  } else {
  RxTileIter ti((unsigned)first, (unsigned)tps, (unsigned)stride);
  for (long long tile = first; tile < ntiles; tile += stride, it++, ti.next()) {
    const long long G0 = (long long)ti.tl * C::kPeriods;
    cf *xt = xbuf, *ot = obuf + (it % kRxV3Obufs) * C::kOut;
    RX_TRACE(0, it, 0);
    if (threadIdx.x == 0 && store_pending) bulk_store_wait_read<kRxV3Obufs - 1>();   // the block this step writes has drained
    compute_warps_sync<NC>();
    mbar_wait(&full[b], use & 1);                                        // this step's input has landed
    RX_TRACE(0, it, 1);
    if (I16) {
      rxv3_widen<C::kRows, NC>(xt, stage + b * C::kStage, swap_iq);      // previous compute left xt at the closing barrier
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
      compute_warps_sync<NC>();                                          // xt ready
    } else {
      xt = xbuf + b * C::kInSlot;
    }
    const bool q8 = ti.template q8<C::kPeriods>(row);
    RX_TRACE(0, it, 2);
#ifndef BTS_RXV3_NOMATH          // (defined: the memory pipeline alone -- tensor copies in, bulk copies out -- for profiles/README.md r3j)
    rx_part<kRxV3Parts>(part, s_taps, xt + row * kRxRowPitch, ot + row * kRxP, q8);
#endif
    RX_TRACE(0, it, 3);
#ifndef BTS_RXV3_LATE_RELEASE
    if (!I16) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[b]);
    }
#endif
#ifdef BTS_RXV3_LATE_RELEASE
    const int bdone = b;
#endif
    if (++b == R) { b = 0; use++; }
    const long long nper = nperiods - G0 < C::kPeriods ? nperiods - G0 : C::kPeriods;
    // the 32T x 65 outputs are contiguous in shared and global memory: one bulk copy, issued by one thread,
    // that drains while the next step waits for its input
    cf *og = out + ti.st * out_pitch + G0 * kRxP;
    fence_async_smem();
    compute_warps_sync<NC>();                                            // all outputs written; all reads of xt done
    RX_TRACE(0, it, 4);
    if (nper == C::kPeriods) {
      if (threadIdx.x == 0) bulk_store(og, ot, (unsigned)(C::kOut * sizeof(cf)));
      store_pending = true;
    } else {
      for (int i = threadIdx.x; i < nper * kRxP; i += NC) og[i] = ot[i];
      store_pending = false;
    }
    RX_TRACE(0, it, 5);
#ifdef BTS_RXV3_LATE_RELEASE
    // (measurement variant of this barrier-per-step flow: hand the input slot back only after the store has been issued)
    if (!I16) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[bdone]);
    }
#endif
  }
  if (threadIdx.x == 0) bulk_store_wait_all();                            // nothing may still read this CTA's shared memory
#ifdef BTS_RXV3_TRACE
  if (threadIdx.x == 0) { g_rx_cta[blockIdx.x][1] = clock64(); g_rx_cta[blockIdx.x][2] = it; unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); g_rx_cta[blockIdx.x][3] = sm; }
#endif
  }
}

#ifndef BTS_RXV3_TILES_F32
#define BTS_RXV3_TILES_F32 3
#endif
#ifndef BTS_RXV3_TILES_I16
#define BTS_RXV3_TILES_I16 3
#endif
constexpr int kRxV3TilesF32 = BTS_RXV3_TILES_F32, kRxV3TilesI16 = BTS_RXV3_TILES_I16;
static int g_num_sms = 148;
static unsigned rxv3_grid(long long ntiles, int max_ctas = 0) {        // at most one CTA per SM
  const long long cap = (max_ctas > 0 && max_ctas < g_num_sms) ? max_ctas : g_num_sms;
  return (unsigned)(ntiles < cap ? ntiles : cap);
}

void upload_resampler_taps(const DevTables *hostT) {
  float h[kRxP * 16];
  rx_fill_taps(hostT, h);
  cudaMemcpyToSymbol(c_rx_poly, h, sizeof h);
}

// Tensor map over the raw stream as [rows][96 samples]; with history the map starts two rows (192 samples) before `in`.
// Returns the row bias to add to a period index (tile origin = period - 1).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
template <bool I16, int T>
static int rxv3_make_map(CUtensorMap *map, const void *in, int has_history, long long nchunks, int nstreams = 1,
                         long long stream_pitch = 0) {
  using C = RxV3<I16, T>;
  if (!g_encode_tiled) return -1;
  const size_t sample = I16 ? 4 : 8;
  const char *base = reinterpret_cast<const char *>(in) - (has_history ? 192 * sample : 0);
  const cuuint64_t rows = (cuuint64_t)nchunks * 9 + (has_history ? 2 : 0);
  // float: 32-bit elements, 192 per row, box 196 wide (= 98 samples); int16: one 32-bit element per I/Q pair
  // third dimension: the streams, stream_pitch samples apart (one stream: any pitch that keeps the stride legal)
  const cuuint64_t spitch = nstreams > 1 ? (cuuint64_t)stream_pitch * sample : rows * 96 * sample;
  if (spitch % 16) return -1;
  const cuuint64_t gdim[3] = {I16 ? 96u : 192u, rows, (cuuint64_t)nstreams};
  const cuuint64_t gstride[2] = {(cuuint64_t)(96 * sample), spitch};
  const cuuint32_t box[3] = {I16 ? 96u : 2u * kRxRowPitch, (cuuint32_t)C::kRows, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUresult r = g_encode_tiled(map, I16 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                                    const_cast<char *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? (has_history ? 2 : 0) : -1;
}

// max_ctas > 0 caps the persistent grid (the caller overlaps the launch with other kernels and leaves them the other SMs)
// light != 0 runs the one-tile configuration (9 warps, 21 K registers, 70 KB per CTA instead of 26 warps / 60 K / 204 KB): it
// leaves most of every SM to kernels of another stream (the segmented btsdsp_rx_stream_dev overlaps it with the demod kernels)
void launch_resample_rx(const DevTables *T, const cf *in, int has_history, long long nchunks, cf *out, cudaStream_t st,
                        int max_ctas, int light) {
  if (nchunks <= 0) return;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  alignas(64) CUtensorMap map;
  if (light && aligned) {
    using C = RxV3<false, 1>;
    const int bias = rxv3_make_map<false, 1>(&map, in, has_history, nchunks);
    if (bias >= 0) {
      const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
      k_resample_rx_v3<false, 1><<<rxv3_grid(ntiles, max_ctas), C::kThreads, C::kSmem, st>>>(map, bias, 0, nperiods, out, 1, 0);
      return;
    }
  }
  const int bias = aligned ? rxv3_make_map<false, kRxV3TilesF32>(&map, in, has_history, nchunks) : -1;
  if (bias >= 0) {
    using C = RxV3<false, kRxV3TilesF32>;
    const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
    k_resample_rx_v3<false, kRxV3TilesF32><<<rxv3_grid(ntiles, max_ctas), C::kThreads, C::kSmem, st>>>(map, bias, 0, nperiods, out, 1, 0);
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
  alignas(64) CUtensorMap map;
  const int bias = rxv3_make_map<true, kRxV3TilesI16>(&map, in, has_history, nchunks);
  if (bias < 0) return -1;
  const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods;
  k_resample_rx_v3<true, kRxV3TilesI16><<<rxv3_grid(ntiles), C::kThreads, C::kSmem, st>>>(map, bias, swap_iq, nperiods, out, 1, 0);
  return 1;
}
// nstreams radios at once: stream a = in + 2*a*in_pitch int16, its outputs at out + a*out_pitch samples
int launch_resample_rx_i16_multi(const int16_t *in, long long in_pitch, int nstreams, int swap_iq, int has_history,
                                 long long nchunks, cf *out, long long out_pitch, cudaStream_t st) {
  if (nchunks <= 0 || nstreams <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
  if ((in_pitch * 4) % 16 || (out_pitch * 8) % 16) return -1;
  using C = RxV3<true, kRxV3TilesI16>;
  alignas(64) CUtensorMap map;
  const int bias = rxv3_make_map<true, kRxV3TilesI16>(&map, in, has_history, nchunks, nstreams, in_pitch);
  if (bias < 0) return -1;
  const long long nperiods = nchunks * 9, ntiles = (nperiods + C::kPeriods - 1) / C::kPeriods * nstreams;
  k_resample_rx_v3<true, kRxV3TilesI16><<<rxv3_grid(ntiles), C::kThreads, C::kSmem, st>>>(map, bias, swap_iq, nperiods, out, nstreams,
                                                                                        out_pitch);
  return 1;
}
// ------------------------------------------------------------------------------------------------
// k_tx_fused: the whole transmit chain in one kernel -- bits -> modulateBurst -> (power scaling) -> slot stream ->
// pushBuffer's 96/65 polyphase resample -> x13500 -> int16 {I,Q}.  The modulated stream (1250 B per burst) never
// exists in global memory: a persistent CTA per SM computes the 96 periods (6240 samples + halo) of its step straight
// into shared memory from the 148-byte bursts (L2-resident), resamples them exactly like the RX kernel does (lane =
// period, phase parts per warp, two-instruction taps from shared memory), and writes the 96 x 96 int16 pairs back as
// coalesced 128-byte rows.  HBM traffic: 148 B in and 923 B out per burst instead of 1398 + 2173.
// ------------------------------------------------------------------------------------------------
constexpr int kTxFusedTiles = 3, kTxFusedParts = 8;
constexpr int kTxFusedPeriods = 32 * kTxFusedTiles;                              // periods per step
constexpr int kTxFusedIn = kTxQ * kTxFusedPeriods + 2 * kTxHalo;               // samples in the input tile
constexpr int kTxFusedOutPitch = kTxP + 1;                                      // 97 words: conflict-free STS.32
constexpr int kTxFusedThreads = 32 * kTxFusedTiles * kTxFusedParts;
constexpr int kTxFusedBursts = ((kTxFusedIn + 624) / 625 + 1) * 4;             // slots a step's tile can touch (from a group start)
constexpr size_t kTxFusedSmem = (size_t)kTxFusedIn * sizeof(cf) + (size_t)kTxFusedPeriods * kTxFusedOutPitch * 4 + kTxP * 8 * 4 +
                                148 * 3 * sizeof(cf) + (size_t)kTxFusedBursts * 148;

// BITS = false: the same resampler over an already modulated stream `samples` (the stand-alone pushBuffer entry point,
// btsdsp_resample_tx_dev): the tile is loaded instead of computed; has_history says samples[-4..-1] are real stream samples
template <bool BITS>
__global__ void __launch_bounds__(kTxFusedThreads, 1) k_tx_fused(const DevTables *__restrict__ T, const uint8_t *__restrict__ bits,
                                                                const float *__restrict__ scale, long long nsamples,
                                                                long long nperiods, int nstreams, short2 *__restrict__ out,
                                                                const cf *__restrict__ samples = nullptr, int has_history = 0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *xs = reinterpret_cast<cf *>(smem_raw);                             // xs[j] = stream sample 65*G0 - 4 + j
  short2 *os = reinterpret_cast<short2 *>(xs + kTxFusedIn);
  float *taps = reinterpret_cast<float *>(os + kTxFusedPeriods * kTxFusedOutPitch);
  cf *q = reinterpret_cast<cf *>(taps + kTxP * 8);                       // rot[ai] * pulse[k]
  unsigned char *sb = reinterpret_cast<unsigned char *>(q + 148 * 3);    // this step's bursts, 148 bytes each
  if (BITS) for (int i = threadIdx.x; i < 148 * 3; i += kTxFusedThreads) tx_fill_q(T, q, i);
  for (int i = threadIdx.x; i < kTxP * 8; i += kTxFusedThreads) {        // taps[r*8 + k] = lpf_tx[br_r + 96 k]
    const int r = i >> 3, k = i & 7;
    int br = (kTxQ * (r + kTxDropC + 5)) % kTxP;
    taps[i] = T->tx_poly[br][k];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int part = warp % kTxFusedParts, row = (warp / kTxFusedParts) * 32 + lane;
  // nstreams independent slot streams of nsamples samples each (bits, scale and out laid out stream after stream); a
  // global step gs is step gs % nsteps of stream gs / nsteps, and every stream starts from zero filter history
  const long long nsteps = (nperiods + kTxFusedPeriods - 1) / kTxFusedPeriods, ngsteps = nsteps * nstreams;
  // the bits a step needs (<= 44 bursts = 6.5 KB) are fetched one step ahead into a register per thread and parked in
  // shared memory at the top of the step, so their latency hides under the previous step's arithmetic
  const long long nslots = nsamples / 625 * 4;
  const bool vec = (reinterpret_cast<uintptr_t>(bits) & 15) == 0;        // 4-slot groups are 592 B = 37 x 16 B
  const unsigned nst32 = (unsigned)nsteps;                               // 32-bit division per step (2^31 steps would be 79 TB of output)
  auto group_of = [](long long step_) {
    const long long s0_ = (long long)kTxQ * step_ * kTxFusedPeriods - kTxHalo;
    return (s0_ < 0 ? 0 : s0_) / 625;
  };
  auto fetch = [&](long long gs_) {
    int4 v = make_int4(0, 0, 0, 0);
    const long long a_ = (unsigned)gs_ / nst32, gfirst = group_of(gs_ - a_ * nsteps) * 4;
    long long nb = nslots - gfirst;
    if (nb > kTxFusedBursts) nb = kTxFusedBursts;
    if (vec && (int)threadIdx.x * 16 < (int)nb * 148)
      v = __ldg(reinterpret_cast<const int4 *>(bits + (a_ * nslots + gfirst) * 148) + threadIdx.x);
    return v;
  };
  static_assert(kTxFusedBursts * 148 <= kTxFusedThreads * 16, "one 16-byte fetch per thread must cover a step's bits");
  int4 pre = (BITS && blockIdx.x < ngsteps) ? fetch(blockIdx.x) : make_int4(0, 0, 0, 0);
  for (long long gs = blockIdx.x; gs < ngsteps; gs += gridDim.x) {
    const long long a = (unsigned)gs / nst32, step = gs - a * nsteps;
    const long long G0 = step * kTxFusedPeriods;
    __syncthreads();                                                     // previous step's tiles are free
    // ---- park the bits of the bursts this step touches, then modulate its samples into the tile
    const long long s0 = (long long)kTxQ * G0 - kTxHalo;
    if (!BITS) {
      // ---- the stand-alone resampler: the step's 6248 samples come from the modulated stream (coalesced 8-byte loads)
      const cf *sp = samples + a * nsamples;
      for (int j = threadIdx.x; j < kTxFusedIn; j += kTxFusedThreads) {
        const long long t = s0 + j;
        xs[j] = ((t >= 0 || (has_history && a == 0)) && t < nsamples) ? __ldg(sp + t) : mk(0.0F, 0.0F);
      }
    } else {
      const long long ga4 = group_of(step);                              // staging starts at slot 4*ga4
      if (vec) {
        if ((int)threadIdx.x * 16 < kTxFusedBursts * 148) reinterpret_cast<int4 *>(sb)[threadIdx.x] = pre;
      } else {
        const long long gfirst = ga4 * 4;
        long long nb = nslots - gfirst;
        if (nb > kTxFusedBursts) nb = kTxFusedBursts;
        const unsigned char *src = bits + (a * nslots + gfirst) * 148;
        for (int i = threadIdx.x; i < (int)nb * 148; i += kTxFusedThreads) sb[i] = src[i];
      }
      __syncthreads();
      if (gs + gridDim.x < ngsteps) pre = fetch(gs + gridDim.x);           // in flight during this step
      const int w0 = (int)(s0 - ga4 * 625);                                // in [-4, 624]
      for (int j = threadIdx.x; j < kTxFusedIn; j += kTxFusedThreads) {
        const int w = w0 + j;
        cf x = mk(0.0F, 0.0F);
        if (w >= 0 && s0 + j < nsamples) {
          const int q4 = w / 625;
          int sl, t;
          tx_slot_of(w - q4 * 625, &sl, &t);
          const int lb = q4 * 4 + sl;                                      // burst index within the staged bits
          x = tx_burst_sample(q, sb + lb * 148, t);
          if (scale) x = cmul(x, mk(scale[a * nslots + ga4 * 4 + lb], 0.0F));   // addRadioVector's scaleVector, Transceiver.cpp:108
        }
        xs[j] = x;
      }
    }
    __syncthreads();
    // ---- this warp's part of the 96 phases of this lane's period
    const bool q8 = (((unsigned)(step % 9) * (unsigned)(kTxFusedPeriods % 9) + (unsigned)row) % 9u) == 8u;   // (G0 + row) % 9 == 8
    tx_part<kTxFusedParts>(part, taps, xs + kTxQ * row + kTxHalo, os + row * kTxFusedOutPitch, q8);
    __syncthreads();
    // ---- rows of 96 int16 pairs (384 B) back to global, coalesced
    const int nper = (int)(nperiods - G0 < kTxFusedPeriods ? nperiods - G0 : kTxFusedPeriods);
    for (int p = warp; p < nper; p += kTxFusedThreads / 32) {
      short2 *og = out + (a * nperiods + G0 + p) * kTxP;
      const short2 *src = os + p * kTxFusedOutPitch;
      og[lane] = src[lane];
      og[lane + 32] = src[lane + 32];
      og[lane + 64] = src[lane + 64];
    }
  }
}
// bits: nstreams x nslots x 148 bytes (slot g of stream a), nslots % 4 == 0 and a whole number of 585-sample chunks
void launch_tx_fused(const DevTables *T, const uint8_t *bits, const float *scale, long long nslots, int16_t *out, cudaStream_t st,
                     int nstreams) {
  const long long nsamples = nslots / 4 * 625, nchunks = nsamples / 585, nperiods = nchunks * 9;
  if (nperiods <= 0 || nstreams <= 0) return;
  const long long nsteps = (nperiods + kTxFusedPeriods - 1) / kTxFusedPeriods * nstreams;
  const unsigned grid = (unsigned)(nsteps < g_num_sms ? nsteps : g_num_sms);
  k_tx_fused<true><<<grid, kTxFusedThreads, kTxFusedSmem, st>>>(T, bits, scale, nsamples, nperiods, nstreams,
                                                                reinterpret_cast<short2 *>(out));
}

int configure_resamplers() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_num_sms = sms;
  if (!g_encode_tiled) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    if (!g_encode_tiled) return -1;
  }
  if (cudaFuncSetAttribute(k_tx_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTxFusedSmem) != cudaSuccess) return -2;
  if (cudaFuncSetAttribute(k_tx_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTxFusedSmem) != cudaSuccess) return -2;
  cudaError_t e = cudaFuncSetAttribute(k_resample_rx_v3<false, kRxV3TilesF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)RxV3<false, kRxV3TilesF32>::kSmem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_resample_rx_v3<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RxV3<false, 1>::kSmem);
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
#ifndef BTS_TX_TUNED_DEFAULT
#define BTS_TX_TUNED_DEFAULT true       // GPU parity run: all tests pass with it, 0.229 ms against 0.306 ms per 64 000 chunks (profiles/README.md r3e)
#endif
void launch_resample_tx(const DevTables *T, const cf *in, int has_history, long long nchunks, int16_t *out,
                        cudaStream_t st) {
  if (nchunks <= 0) return;
  static const bool tuned = [] { const char *e = getenv("BTSDSP_TX_TUNED"); return e ? atoi(e) != 0 : BTS_TX_TUNED_DEFAULT; }();
  if (tuned && ((reinterpret_cast<uintptr_t>(in) & 7) | (reinterpret_cast<uintptr_t>(out) & 3)) == 0) {
    // the tuned shape (persistent CTA per SM, lane = period, two-instruction taps): k_tx_fused with a loaded input tile
    const long long nsamples = nchunks * 585, nperiods = nchunks * 9;
    const long long nsteps = (nperiods + kTxFusedPeriods - 1) / kTxFusedPeriods;
    const unsigned grid = (unsigned)(nsteps < g_num_sms ? nsteps : g_num_sms);
    k_tx_fused<false><<<grid, kTxFusedThreads, kTxFusedSmem, st>>>(T, nullptr, nullptr, nsamples, nperiods, 1, reinterpret_cast<short2 *>(out),
                                                                 in, has_history);
    return;
  }
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

// the same with a filter supplied by the caller (any length; complex taps when the LPF is not real-only, :1187-1193)
__global__ void k_resample_taps(const cf *__restrict__ x, int n, int P, int Q, const cf *__restrict__ lpf, int L, int real_taps,
                                cf *__restrict__ out, int outn) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= outn) return;
  const int outputIx = (L - 1) / 2 / Q + o;
  const int br = (outputIx * Q) % P;
  int ix = (outputIx * Q - br) / P;
  int f = br;
  while (ix >= n) { ix--; f += P; }
  cf sum = mk(0.0F, 0.0F);
  if (real_taps) while (ix >= 0 && f < L) { sum = cadd(sum, cmulr(x[ix], lpf[f].x)); ix--; f += P; }
  else while (ix >= 0 && f < L) { sum = cadd(sum, cmul(x[ix], lpf[f])); ix--; f += P; }
  out[o] = sum;
}
void launch_resample_taps(const cf *x, int n, int P, int Q, const cf *ctaps, int L, int real_taps, cf *out, int outn,
                          cudaStream_t st) {
  if (outn <= 0) return;
  k_resample_taps<<<(outn + 127) / 128, 128, 0, st>>>(x, n, P, Q, ctaps, L, real_taps, out, outn);
}

}  // namespace btsdsp
