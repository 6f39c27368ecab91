// cplx.cuh -- exact (never contracted) complex<float> arithmetic for the burst-DSP kernels.
//
// Bit-parity with the reference's scalar IEEE binary32 arithmetic (reference Transceiver/Complex.h)
// is the contract: every product and sum is rounded separately, in the reference's operand order.
// The __f*_rn intrinsics are never fused into FMAs by nvcc, independent of -fmad.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace btsdsp {

#define BTS_HD __host__ __device__ __forceinline__

typedef float2 cf;   // .x = real, .y = imag ; same memory layout as the reference's Complex<float>

BTS_HD cf mk(float r, float i) { return make_float2(r, i); }

#ifdef __CUDA_ARCH__
#define BTS_MUL(a, b) __fmul_rn((a), (b))
#define BTS_ADD(a, b) __fadd_rn((a), (b))
#define BTS_SUB(a, b) __fsub_rn((a), (b))
#define BTS_DIV(a, b) __fdiv_rn((a), (b))
#define BTS_SQRT(a) __fsqrt_rn((a))
#else
#define BTS_MUL(a, b) ((a) * (b))
#define BTS_ADD(a, b) ((a) + (b))
#define BTS_SUB(a, b) ((a) - (b))
#define BTS_DIV(a, b) ((a) / (b))
#define BTS_SQRT(a) sqrtf((a))
#endif

// Complex.h:79  operator+
BTS_HD cf cadd(cf a, cf b) { return mk(BTS_ADD(a.x, b.x), BTS_ADD(a.y, b.y)); }
BTS_HD cf csub(cf a, cf b) { return mk(BTS_SUB(a.x, b.x), BTS_SUB(a.y, b.y)); }
// Packed products.  sm_100 has FMUL2 (mul.rn.f32x2): two independent, separately rounded binary32 products in one
// instruction = bit-identical to two FMULs.  It issues at half rate, so it saves instruction BYTES (the unrolled
// kernels are instruction-cache bound), not issue slots.  Its result must only feed SCALAR adds: ptxas 12.9 fuses
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (tools/microbench_fp32x2.cu); the build checks
// that no FFMA2 with a live addend was emitted (see pmul0 below for the one allowed form).
BTS_HD cf pmul(cf a, float s) {                 // (a.x*s, a.y*s)
#ifdef __CUDA_ARCH__
  unsigned long long pa, ps, pr;
  cf r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(ps));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
#else
  return mk(a.x * s, a.y * s);
#endif
}
// Packed multiply-then-add for accumulating real-tap FIRs (sum += x * h) in TWO instructions per tap, still
// with two roundings: the product is taken as fma.rn.f32x2(x, h, +0) -- an FFMA2 whose addend is the zero register,
// which ptxas cannot merge with the add that follows (an FMA is not contractable) -- and accumulated with
// add.rn.f32x2 (FADD2).  RN(x*h + 0) == RN(x*h) except that a -0 product becomes +0, and adding either zero to an
// accumulator that started at +0 gives the same bits (such an accumulator is never -0), so sums are bit-identical
// to the scalar mul-then-add.  The build accepts FFMA2 only in this zero-addend form.
BTS_HD cf pmul0(cf a, float s) {                // (a.x*s, a.y*s); a -0 product may come back as +0
#ifdef __CUDA_ARCH__
  unsigned long long pa, ps, pz, pr;
  cf r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ps) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pz) : "f"(0.0F));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pr) : "l"(pa), "l"(ps), "l"(pz));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
#else
  return mk(a.x * s, a.y * s);
#endif
}
BTS_HD cf padd(cf a, cf b) {                    // (a.x+b.x, a.y+b.y): only for sums fed by pmul0 or by loads
#ifdef __CUDA_ARCH__
  unsigned long long pa, pb, pr;
  cf r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pr) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(pr));
  return r;
#else
  return mk(a.x + b.x, a.y + b.y);
#endif
}
// acc + a*b for complex a (a loop-invariant tap, `as` = (-a.y, a.x) precomputed) and complex b, in FOUR instructions:
// a*b = (a.x*b.x - a.y*b.y, a.y*b.x + a.x*b.y) = pmul0(a, b.x) + pmul0(as, b.y), each product rounded once, the
// difference formed as x + (-y) (negating a factor negates the rounded product exactly), then one packed add into acc.
// Values are bit-identical to cadd(acc, cmul(a, b)) / cadd(acc, cmul(b, a)) (products and sums commute); only the
// sign of an exactly-zero product term can differ, which an accumulator that is not -0 absorbs (see pmul0).
BTS_HD cf cswapneg(cf a) { return mk(-a.y, a.x); }
BTS_HD cf cmac_tap(cf acc, cf a, cf as, cf b) {
#ifdef __CUDA_ARCH__
  return padd(acc, padd(pmul0(a, b.x), pmul0(as, b.y)));
#else
  (void)as;
  return mk(acc.x + (a.x * b.x - a.y * b.y), acc.y + (a.x * b.y + a.y * b.x));
#endif
}
// Complex.h:83  operator*(Complex): (r*a.r - i*a.i, r*a.i + i*a.r)
BTS_HD cf cmul(cf a, cf b) {
  const cf t1 = pmul(a, b.x), t2 = pmul(a, b.y);        // (a.x*b.x, a.y*b.x), (a.x*b.y, a.y*b.y)
  return mk(BTS_SUB(t1.x, t2.y), BTS_ADD(t2.x, t1.y));
}
// Complex.h:84  operator*(Real)
BTS_HD cf cmulr(cf a, float s) { return pmul(a, s); }
// Complex.h:86  operator/(Real)
BTS_HD cf cdivr(cf a, float s) { return mk(BTS_DIV(a.x, s), BTS_DIV(a.y, s)); }
BTS_HD cf cconj(cf a) { return mk(a.x, -a.y); }
// Complex.h:122 norm2 = i*i + r*r  (imaginary product first)
BTS_HD float cnorm2(cf a) { return BTS_ADD(BTS_MUL(a.y, a.y), BTS_MUL(a.x, a.x)); }
// Complex.h:154-160 inv = (r/n, -i/n)
BTS_HD cf cinv(cf a) {
  float n = cnorm2(a);
  return mk(BTS_DIV(a.x, n), BTS_DIV(-a.y, n));
}
// Complex.h:85  operator/(Complex) = *this * a.inv()
BTS_HD cf cdiv(cf a, cf b) { return cmul(a, cinv(b)); }
// Complex.h:131 abs = sqrt(norm2)
BTS_HD float cabs_(cf a) { return BTS_SQRT(cnorm2(a)); }

// Strided view of a complex vector.  S = 1 for ordinary (global) vectors; S = 33 for the per-lane
// columns of the transposed shared-memory tiles used by the one-burst-per-thread kernels.
template <int S>
struct View {
  cf *p;
  BTS_HD cf ld(int i) const { return p[i * S]; }
  BTS_HD void st(int i, cf v) const { p[i * S] = v; }
  BTS_HD View<S> at(int k) const { return View<S>{p + k * S}; }
};

}  // namespace btsdsp
