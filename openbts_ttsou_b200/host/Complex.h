// host/Complex.h -- minimal complex value type with the reference's arithmetic conventions, for builds of the
// sigProcLib.h shim that do not have the reference's own Transceiver/Complex.h on the include path.
// (With the reference tree available, put ITS headers first on the include path; the shim compiles against both.)
// Conventions that matter to callers (reference Transceiver/Complex.h): operator/ multiplies by the reciprocal
// (r/n, -i/n) with n = i*i + r*r; ordering operators compare squared magnitudes.
#ifndef BTSDSP_HOST_COMPLEX_H
#define BTSDSP_HOST_COMPLEX_H
#include <math.h>
#include <ostream>

template <class Real> class Complex {
 public:
  Real r, i;
  Complex() : r(0), i(0) {}
  Complex(Real re) : r(re), i(0) {}
  Complex(Real re, Real im) : r(re), i(im) {}
  template <class U> Complex(const Complex<U> &z) : r((Real)z.r), i((Real)z.i) {}

  Real real() const { return r; }
  Real imag() const { return i; }
  Real norm2() const { return i * i + r * r; }
  Real abs() const { return (Real)::sqrt(norm2()); }
  Real arg() const { return (Real)::atan2(i, r); }
  Complex conj() const { return Complex(r, -i); }
  Complex inv() const { Real n = norm2(); return Complex(r / n, -i / n); }
  bool isZero() const { return r == (Real)0 && i == (Real)0; }

  Complex operator+(const Complex &a) const { return Complex(r + a.r, i + a.i); }
  Complex operator-(const Complex &a) const { return Complex(r - a.r, i - a.i); }
  Complex operator*(const Complex &a) const { return Complex(r * a.r - i * a.i, r * a.i + i * a.r); }
  Complex operator/(const Complex &a) const { return (*this) * a.inv(); }
  Complex operator+(Real a) const { return Complex(r + a, i); }
  Complex operator-(Real a) const { return Complex(r - a, i); }
  Complex operator*(Real a) const { return Complex(r * a, i * a); }
  Complex operator/(Real a) const { return Complex(r / a, i / a); }
  Complex &operator+=(const Complex &a) { r += a.r; i += a.i; return *this; }
  Complex &operator-=(const Complex &a) { r -= a.r; i -= a.i; return *this; }
  Complex &operator*=(Real a) { r *= a; i *= a; return *this; }
  Complex &operator/=(Real a) { r /= a; i /= a; return *this; }
  bool operator==(const Complex &a) const { return r == a.r && i == a.i; }
  bool operator!=(const Complex &a) const { return !(*this == a); }
  bool operator<(const Complex &a) const { return norm2() < a.norm2(); }
  bool operator>(const Complex &a) const { return norm2() > a.norm2(); }
};

template <class Real> Complex<Real> operator*(Real a, const Complex<Real> &z) { return Complex<Real>(z.r * a, z.i * a); }
template <class Real> std::ostream &operator<<(std::ostream &os, const Complex<Real> &z) { return os << z.r << ' ' << z.i << "j"; }

typedef Complex<float> complex;
#endif
