// Debug harness: per-iteration clock64 timestamps of CTA 0 of the RX resampler (compute warp 0 and the producer warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I openbts_ttsou_b200/csrc -I include \
//        [-DBTS_RXV3_...=n] -o tools/rxv3_trace tools/rxv3_trace.cu && tools/rxv3_trace [nchunks]
#define BTS_RXV3_TRACE 1
#include "../openbts_ttsou_b200/csrc/resample.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace btsdsp;
int main(int argc, char **argv) {
  const long long nchunks = argc > 1 ? atoll(argv[1]) : 213750;
  cf *in, *out;
  cudaMalloc(&in, (nchunks * 864 + 192) * sizeof(cf));
  cudaMalloc(&out, nchunks * 585 * sizeof(cf));
  cudaMemset(in, 0, (nchunks * 864 + 192) * sizeof(cf));
  std::vector<float> taps(kRxP * 16, 0.01f);
  cudaMemcpyToSymbol(c_rx_poly, taps.data(), taps.size() * 4);
  if (configure_resamplers()) { printf("configure failed\n"); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    launch_resample_rx(nullptr, in + 192, 1, nchunks, out, 0);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: %.4f ms\n", rep, ms);
  }
  static long long tr[2][64][8];
  cudaMemcpyFromSymbol(tr, g_rx_trace, sizeof tr);
  const long long t0 = tr[1][0][0];
  printf("it | compute warp0: step_top input_landed compute_start compute_end(incl. drain gate) arrived step_end | mover: step_top all_parts_in store+refill_issued\n");
  for (int it = 0; it < 24; it++) {
    printf("%2d |", it);
    for (int k = 0; k < 6; k++) printf(" %7lld", tr[0][it][k] - t0);
    printf(" |");
    for (int k = 0; k < 3; k++) printf(" %7lld", tr[1][it][k] - t0);
    printf("\n");
  }
  static long long ct[256][4];
  cudaMemcpyFromSymbol(ct, g_rx_cta, sizeof ct);
  printf("per-CTA cycles (cta:sm:iters:cycles):");
  for (int c = 0; c < 148; c++) { if (c % 8 == 0) printf("\n"); printf(" %3d:%3lld:%3lld:%7lld", c, ct[c][3], ct[c][2], ct[c][1] - ct[c][0]); }
  printf("\n");
  return 0;
}
