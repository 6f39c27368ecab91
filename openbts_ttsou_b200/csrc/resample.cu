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

constexpr size_t kRxV2Smem = (size_t)(kRxTileIn + kRxTileOut) * sizeof(cf);

// I16 = the radio's own sample format: interleaved int16 {I,Q} pairs (4 B per sample, `swap_iq` for the Q-first
// order of real USRP hardware), converted exactly as unUSRPifyVector does (radioInterface.cpp:91-116) while the
// tile is built -- the raw samples land in the (still unused) output tile via cp.async and are expanded to float.
template <bool I16>
__global__ void __launch_bounds__(64) k_resample_rx_v2(const void *__restrict__ in_, int has_history, int swap_iq,
                                                       long long nperiods, long long nsamples, cf *__restrict__ out) {
  const cf *in = reinterpret_cast<const cf *>(in_);
  const short2 *in16 = reinterpret_cast<const short2 *>(in_);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf *xt = reinterpret_cast<cf *>(smem_raw);
  cf *ot = xt + kRxTileIn;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      // two warps share a tile: each does half the phases
  for (long long tile = blockIdx.x; tile * 32 < nperiods; tile += gridDim.x) {
    const long long G0 = tile * 32;
    const long long raw0 = 96 * G0 - 96;                               // tile origin: sample (G, r, k) sits at 96*l + ix_r - k - 96
    // ---- load 34 rows x 96 samples as 16-byte cp.async copies (global -> shared, no registers, all 51 per lane
    //      in flight at once); samples outside [lo, nsamples) are zero-filled by the copy's src-size operand
    const long long lo = has_history ? -192 : 0;
    __syncthreads();                                                    // previous tile fully written back
    if (!I16) {
      for (int i4 = threadIdx.x; i4 < kRxTileRows * 48; i4 += 64) {
        const int row = i4 / 48, c4 = i4 - row * 48;
        const long long s = raw0 + (long long)row * 96 + 2 * c4;
        cf *dst = xt + row * kRxRowPitch + 2 * c4;
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
      for (int i = threadIdx.x; i < kRxTileRows * 24; i += 64) {
        const long long s = raw0 + 4LL * i;
        const bool valid = s >= lo && s < nsamples;
        cp_async16z(stage + 4 * i, in16 + (valid ? s : 0), valid);
      }
      cp_async_wait_all();
      __syncthreads();
      for (int i = threadIdx.x; i < kRxTileRows * 24; i += 64) {
        const int row = i / 24, c = 4 * (i - row * 24);
        const int4 v = *reinterpret_cast<const int4 *>(stage + 4 * i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        cf *dst = xt + row * kRxRowPitch + c;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float a = (float)(short)(w[k] & 0xffff), b = (float)(short)(w[k] >> 16);     // first, second int16
          dst[k] = swap_iq ? mk(b, a) : mk(a, b);
        }
      }
      __syncthreads();
    }
    // ---- 65 phases for this lane's period
    const long long G = G0 + lane;
    const bool q8 = (G % 9) == 8;
    const cf *xl = xt + lane * kRxRowPitch;
    cf *ol = ot + lane * kRxP;
    if (warp == 0) rx_half<0>(c_rx_poly, xl, ol, q8);
    else rx_half<1>(c_rx_poly, xl, ol, q8);
    __syncthreads();
    // ---- write the 32 x 65 outputs back: contiguous in both shared and global memory
    const long long nvalid = (nperiods - G0 < 32 ? nperiods - G0 : 32) * kRxP;
    cf *og = out + G0 * kRxP;
    for (int i4 = threadIdx.x; i4 < kRxTileOut / 2; i4 += 64) {
      if (2 * i4 + 1 < nvalid) *reinterpret_cast<float4 *>(og + 2 * i4) = *reinterpret_cast<const float4 *>(ot + 2 * i4);
      else if (2 * i4 < nvalid) og[2 * i4] = ot[2 * i4];
    }
  }
}

void upload_resampler_taps(const DevTables *hostT) {
  float h[kRxP * 16];
  rx_fill_taps(hostT, h);
  cudaMemcpyToSymbol(c_rx_poly, h, sizeof h);
}

void launch_resample_rx(const DevTables *T, const cf *in, int has_history, long long nchunks, cf *out, cudaStream_t st) {
  if (nchunks <= 0) return;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (aligned) {
    const long long nperiods = nchunks * 9, ntiles = (nperiods + 31) / 32;
    const unsigned grid = (unsigned)(ntiles < 148 * 5 * 8 ? ntiles : 148 * 5 * 8);
    k_resample_rx_v2<false><<<grid, 64, kRxV2Smem, st>>>(in, has_history, 0, nperiods, nchunks * 864, out);
  } else {
    const unsigned grid = (unsigned)(nchunks < 148 * 32 ? nchunks : 148 * 32);
    k_resample_rx<<<grid, 256, 0, st>>>(T, in, has_history, nchunks, out);
  }
}
// int16 {I,Q} ingest (the radio's format); `in` must be 16-byte aligned
int launch_resample_rx_i16(const int16_t *in, int swap_iq, int has_history, long long nchunks, cf *out, cudaStream_t st) {
  if (nchunks <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) return -1;
  const long long nperiods = nchunks * 9, ntiles = (nperiods + 31) / 32;
  const unsigned grid = (unsigned)(ntiles < 148 * 5 * 8 ? ntiles : 148 * 5 * 8);
  k_resample_rx_v2<true><<<grid, 64, kRxV2Smem, st>>>(in, has_history, swap_iq, nperiods, nchunks * 864, out);
  return 1;
}
int configure_resamplers() {
  cudaError_t e = cudaFuncSetAttribute(k_resample_rx_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRxV2Smem);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaFuncSetAttribute(k_resample_rx_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRxV2Smem);
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
