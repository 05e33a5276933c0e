// nmp_math.h — portable, bit-reproducible transcendental functions ("M1" math of DESIGN.md).
//
// The Noah-MP column physics (reference: phys/module_sf_noahmplsm.F90) uses the Fortran
// intrinsics EXP, LOG, LOG10, ATAN, TAN, ACOS, COS, TANH, SQRT and real-exponent `**`.
// The reference takes them from the host libm; CUDA's libdevice versions differ from any
// libm in the last ulp, and the physics has hard thresholds (Newton exit tests, snow-layer
// splits) that can flip on one ulp.  To make GPU-vs-CPU parity an *exact* compare we evaluate
// every transcendental with the code in this header, which uses only IEEE-754 basic
// operations (+,-,*,/,sqrt on fp64, explicit conversions, integer bit moves).  The same
// source is compiled by nvcc for sm_100a (with -fmad=false in parity builds) and by g++
// (with -ffp-contract=off), and yields identical bits on both.
//
// Accuracy: the fp64 cores are accurate to ~1e-15 relative, so the fp32 results are the
// correctly rounded value in all but ~1e-7 of calls — i.e. as close to glibc's (≤1 ulp)
// functions as two good libms are to each other.
//
// This header is product code (it is what the parity-mode kernels execute); the CPU oracle
// under oracle/ includes it for its M1 mode and uses glibc for its M0 mode.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define NMP_HD __host__ __device__ __forceinline__
#else
#define NMP_HD inline
#endif

namespace nmpm {

NMP_HD double bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}
NMP_HD uint64_t d2bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}

// fp64 square root, IEEE correctly rounded on both targets.
NMP_HD double sqrt_d(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}

// ---------------------------------------------------------------------------------------
// exp(x), fp64 core.  x = k*ln2 + r, |r| <= ln2/2, Taylor to r^13, scale by 2^k.
// ---------------------------------------------------------------------------------------
NMP_HD double exp_d(double x) {
  if (!(x == x)) return x;                         // NaN
  if (x > 709.78) return bits2d(0x7FF0000000000000ull);   // +inf
  if (x < -745.2) return 0.0;
  const double LOG2E  = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;       // 33 significant bits
  const double LN2_LO = 1.90821492927058770002e-10;
  double t = x * LOG2E;
  // round to nearest integer without relying on the rounding-mode intrinsics
  double kf = (t >= 0.0) ? (double)(long long)(t + 0.5) : (double)(long long)(t - 0.5);
  long long k = (long long)kf;
  double r = (x - kf * LN2_HI) - kf * LN2_LO;
  // Horner, Taylor coefficients 1/n!
  double p = 1.0 / 6227020800.0;                   // 1/13!
  p = p * r + 1.0 / 479001600.0;
  p = p * r + 1.0 / 39916800.0;
  p = p * r + 1.0 / 3628800.0;
  p = p * r + 1.0 / 362880.0;
  p = p * r + 1.0 / 40320.0;
  p = p * r + 1.0 / 5040.0;
  p = p * r + 1.0 / 720.0;
  p = p * r + 1.0 / 120.0;
  p = p * r + 1.0 / 24.0;
  p = p * r + 1.0 / 6.0;
  p = p * r + 0.5;
  p = p * r + 1.0;
  p = p * r + 1.0;
  // scale by 2^k in two steps so that subnormal results round once, acceptably
  if (k > 1000) { p *= bits2d((uint64_t)(1023 + 1000) << 52); k -= 1000; }
  if (k < -1000) { p *= bits2d((uint64_t)(1023 - 1000) << 52); k += 1000; }
  return p * bits2d((uint64_t)(1023 + k) << 52);
}

// ---------------------------------------------------------------------------------------
// log(x), fp64 core.  x = 2^e * m, m in [sqrt(1/2), sqrt(2)), s=(m-1)/(m+1),
// log(m) = 2s(1 + s^2/3 + s^4/5 + ...), series to s^26.
// ---------------------------------------------------------------------------------------
NMP_HD double log_d(double x) {
  if (!(x == x)) return x;
  if (x < 0.0) return bits2d(0x7FF8000000000000ull);       // NaN
  if (x == 0.0) return bits2d(0xFFF0000000000000ull);      // -inf
  uint64_t u = d2bits(x);
  if (u == 0x7FF0000000000000ull) return x;                // +inf
  int e = 0;
  if ((u >> 52) == 0) {                                    // subnormal: renormalise
    x *= 18014398509481984.0;                              // 2^54
    u = d2bits(x);
    e = -54;
  }
  e += (int)(u >> 52) - 1023;
  u = (u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double m = bits2d(u);                                    // [1,2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  double p = 1.0 / 27.0;
  p = p * z + 1.0 / 25.0;
  p = p * z + 1.0 / 23.0;
  p = p * z + 1.0 / 21.0;
  p = p * z + 1.0 / 19.0;
  p = p * z + 1.0 / 17.0;
  p = p * z + 1.0 / 15.0;
  p = p * z + 1.0 / 13.0;
  p = p * z + 1.0 / 11.0;
  p = p * z + 1.0 / 9.0;
  p = p * z + 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z + 1.0 / 3.0;
  p = p * z;                                               // s^2/3 + s^4/5 + ...
  double lm = 2.0 * s + 2.0 * s * p;
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  double ef = (double)e;
  return (ef * LN2_LO + lm) + ef * LN2_HI;
}

// pow in fp64 (used by GROUNDWATER's REAL(KIND=8) S_NODE**(-BEXP), noahmplsm.F90:8504-8506).
NMP_HD double pow_d(double x, double y) {
  if (y == 0.0) return 1.0;
  if (x == 1.0) return 1.0;
  if (x == 0.0) return (y > 0.0) ? 0.0 : bits2d(0x7FF0000000000000ull);
  if (x < 0.0) {
    // integral y only (the physics never raises a negative base to a fractional power
    // without having produced NaN in the reference as well)
    double yi = (double)(long long)y;
    if (yi != y) return bits2d(0x7FF8000000000000ull);
    double r = exp_d(y * log_d(-x));
    return (((long long)y) & 1) ? -r : r;
  }
  return exp_d(y * log_d(x));
}

// atan in fp64: fold to [0,1], two half-angle steps, then Taylor to x^25.
NMP_HD double atan_d(double x) {
  if (!(x == x)) return x;
  const double PIO2 = 1.57079632679489661923;
  double ax = x < 0.0 ? -x : x;
  bool inv = ax > 1.0;
  if (inv) ax = 1.0 / ax;
  ax = ax / (1.0 + sqrt_d(1.0 + ax * ax));
  ax = ax / (1.0 + sqrt_d(1.0 + ax * ax));           // |ax| <= tan(pi/16)
  double z = ax * ax;
  double p = 1.0 / 25.0;
  p = 1.0 / 23.0 - p * z;
  p = 1.0 / 21.0 - p * z;
  p = 1.0 / 19.0 - p * z;
  p = 1.0 / 17.0 - p * z;
  p = 1.0 / 15.0 - p * z;
  p = 1.0 / 13.0 - p * z;
  p = 1.0 / 11.0 - p * z;
  p = 1.0 / 9.0 - p * z;
  p = 1.0 / 7.0 - p * z;
  p = 1.0 / 5.0 - p * z;
  p = 1.0 / 3.0 - p * z;
  p = 1.0 - p * z;
  double r = 4.0 * (ax * p);
  if (inv) r = PIO2 - r;
  return x < 0.0 ? -r : r;
}

// sin/cos in fp64 for moderate arguments (|x| < ~1e5; the physics only needs [0, pi/2]).
NMP_HD void sincos_d(double x, double* sn, double* cs) {
  const double TWO_OVER_PI = 0.63661977236758134308;
  const double PIO2_HI = 1.57079632673412561417e+00;     // 33 bits
  const double PIO2_LO = 6.07710050650619224932e-11;
  double t = x * TWO_OVER_PI;
  double kf = (t >= 0.0) ? (double)(long long)(t + 0.5) : (double)(long long)(t - 0.5);
  long long k = (long long)kf;
  double r = (x - kf * PIO2_HI) - kf * PIO2_LO;          // |r| <= pi/4
  double z = r * r;
  // sin: r(1 - z/3! + z^2/5! ...) to r^17 ; cos: 1 - z/2! + ... to r^18
  double ps = -1.0 / 355687428096000.0;                  // -1/17!
  ps = ps * z + 1.0 / 1307674368000.0;
  ps = ps * z - 1.0 / 6227020800.0;
  ps = ps * z + 1.0 / 39916800.0;
  ps = ps * z - 1.0 / 362880.0;
  ps = ps * z + 1.0 / 5040.0;
  ps = ps * z - 1.0 / 120.0;
  ps = ps * z + 1.0 / 6.0;
  double s = r - r * (z * ps);
  double pc = 1.0 / 6402373705728000.0;                  // 1/18!
  pc = pc * z - 1.0 / 20922789888000.0;
  pc = pc * z + 1.0 / 87178291200.0;
  pc = pc * z - 1.0 / 479001600.0;
  pc = pc * z + 1.0 / 3628800.0;
  pc = pc * z - 1.0 / 40320.0;
  pc = pc * z + 1.0 / 720.0;
  pc = pc * z - 1.0 / 24.0;
  pc = pc * z + 0.5;
  double c = 1.0 - z * pc;
  switch ((int)(k & 3)) {
    case 0: *sn = s;  *cs = c;  break;
    case 1: *sn = c;  *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}

// ---------------------------------------------------------------------------------------
// fp32 front ends (what the physics calls)
// ---------------------------------------------------------------------------------------
NMP_HD float expf_(float x)   { return (float)exp_d((double)x); }
NMP_HD float logf_(float x)   { return (float)log_d((double)x); }
NMP_HD float log10f_(float x) { return (float)(log_d((double)x) * 0.43429448190325182765); }
NMP_HD float powf_(float x, float y) { return (float)pow_d((double)x, (double)y); }
NMP_HD float atanf_(float x)  { return (float)atan_d((double)x); }
NMP_HD float cosf_(float x)   { double s, c; sincos_d((double)x, &s, &c); return (float)c; }
NMP_HD float sinf_(float x)   { double s, c; sincos_d((double)x, &s, &c); return (float)s; }
NMP_HD float tanf_(float x)   { double s, c; sincos_d((double)x, &s, &c); return (float)(s / c); }
NMP_HD float acosf_(float x) {
  // acos(x) = 2 atan( sqrt((1-x)/(1+x)) ),  x in (-1,1]
  double xd = (double)x;
  if (xd >= 1.0) return 0.0f;
  if (xd <= -1.0) return 3.14159274f;
  return (float)(2.0 * atan_d(sqrt_d((1.0 - xd) / (1.0 + xd))));
}
NMP_HD float asinf_(float x) {
  // asin(x) = atan( x / sqrt(1 - x^2) ),  |x| < 1
  double xd = (double)x;
  if (xd >= 1.0) return 1.57079637f;
  if (xd <= -1.0) return -1.57079637f;
  return (float)atan_d(xd / sqrt_d(1.0 - xd * xd));
}
NMP_HD float tanhf_(float x) {
  double xd = (double)x;
  if (xd > 20.0) return 1.0f;
  if (xd < -20.0) return -1.0f;
  double ax = xd < 0.0 ? -xd : xd;
  if (ax < 1.0e-4) return x;                          // tanh(x) = x - x^3/3: below fp32 resolution
  double e = exp_d(2.0 * xd);
  return (float)((e - 1.0) / (e + 1.0));
}

}  // namespace nmpm
