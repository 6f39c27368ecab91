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
void launch_resample_rx(const DevTables *T, const cf *in, int has_history, long long nchunks, cf *out, cudaStream_t st) {
  if (nchunks <= 0) return;
  const unsigned grid = (unsigned)(nchunks < 148 * 32 ? nchunks : 148 * 32);
  k_resample_rx<<<grid, 256, 0, st>>>(T, in, has_history, nchunks, out);
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
