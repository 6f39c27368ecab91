// tools/microbench_fp32x2.cu -- measurement tool (not product code).
// Issue rates on sm_100a of the instruction mixes a bit-exact (unfused) complex x real MAC can be built from:
//   mode 0  2 FMUL + 2 FADD            (scalar, what -fmad=false gives)
//   mode 1  FMUL2 + 2 FADD             (packed multiply, scalar adds)
//   mode 2  2 FMUL + FADD2             (scalar multiplies, packed add)
//   mode 3  FFMA2                      (fused; NOT bit-exact -- ptxas 12.9 turns mul.rn.f32x2 + add.rn.f32x2 into this
//                                       even with --fmad=false, which is why modes 1/2 exist)
//   mode 4  2 FFMA                     (fused scalar, reference point)
//   mode 5  LDS.64 conflict-free + FADD2
//   mode 6  dependent FADD chain       (latency)
// Reports warp-level MACs per cycle per SM at 1..16 warps/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/mb tools/microbench_fp32x2.cu && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float x, float y) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;
}
__device__ __forceinline__ float2 upk(unsigned long long v) {
  float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r;
}

template <int MODE>
__global__ void k(float2 *out, int iters, long long *cycles, float hval) {
  __shared__ float2 sm[33 * 64];
  for (int i = threadIdx.x; i < 33 * 64; i += blockDim.x) sm[i] = make_float2(i * 1e-3f, 1.0f);
  __syncthreads();
  float2 a[8];
  for (int j = 0; j < 8; j++) a[j] = make_float2(threadIdx.x * 1e-3f + j, 1.0f + j);
  const float h = hval, c = 1e-3f;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (MODE == 0) {
        a[j].x = __fadd_rn(__fmul_rn(a[j].x, h), c);
        a[j].y = __fadd_rn(__fmul_rn(a[j].y, h), c);
      } else if (MODE == 1) {
        unsigned long long m;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(pk(a[j].x, a[j].y)), "l"(pk(h, h)));
        float2 p = upk(m);
        a[j].x = __fadd_rn(p.x, c); a[j].y = __fadd_rn(p.y, c);
      } else if (MODE == 2) {
        float px = __fmul_rn(a[j].x, h), py = __fmul_rn(a[j].y, h);
        unsigned long long s;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(pk(px, py)), "l"(pk(c, c)));
        a[j] = upk(s);
      } else if (MODE == 3) {
        unsigned long long s;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(s) : "l"(pk(a[j].x, a[j].y)), "l"(pk(h, h)), "l"(pk(c, c)));
        a[j] = upk(s);
      } else if (MODE == 4) {
        a[j].x = __fmaf_rn(a[j].x, h, c); a[j].y = __fmaf_rn(a[j].y, h, c);
      } else if (MODE == 5) {
        unsigned long long s;
        float2 v = sm[((i + j) & 63) * 33 + (threadIdx.x & 31)];
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(pk(a[j].x, a[j].y)), "l"(pk(v.x, v.y)));
        a[j] = upk(s);
      } else {
        a[0].x = __fadd_rn(a[0].x, c);
      }
    }
  }
  long long t1 = clock64();
  float2 s = make_float2(0, 0);
  for (int j = 0; j < 8; j++) { s.x += a[j].x; s.y += a[j].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE> void run(const char *name, float2 *out, long long *cyc) {
  const int iters = 4096;
  for (int warps = 1; warps <= 16; warps *= 2) {
    long long h = 0;
    for (int rep = 0; rep < 2; rep++) { k<MODE><<<148, warps * 32>>>(out, iters, cyc, 0.999f); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM %2d  cycles %9lld  warp-MACs/cycle/SM %.3f\n", name, warps, h, (double)iters * 8 * warps / h);
  }
}

int main() {
  float2 *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float2));
  cudaMalloc(&cyc, 8);
  run<0>("2 FMUL + 2 FADD", out, cyc);
  run<1>("FMUL2 + 2 FADD", out, cyc);
  run<2>("2 FMUL + FADD2", out, cyc);
  run<3>("FFMA2 (fused)", out, cyc);
  run<4>("2 FFMA (fused)", out, cyc);
  run<5>("LDS.64 + FADD2", out, cyc);
  run<6>("dependent FADD (x8)", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
