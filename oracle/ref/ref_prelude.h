// ref_prelude.h — run-time support of the C++ that oracle/ref/f90cxx.py generates from the reference's Fortran
// (TEST INFRASTRUCTURE).  Fortran array views with declared lower bounds, the intrinsics with gfortran's semantics,
// by-reference temporaries for expression arguments.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include "../../noahmp_b200/csrc/nmp_math.h"

// 0 = host libm (what a gfortran build of the reference links), 1 = the portable nmp_math.h routines the CUDA PARITY
// build and the oracle's portable mode use
extern int ref_math_mode;

#ifdef REF_BOUNDS
#define REF_CHECK(i, lo, n, what) \
  do { if ((i) < (lo) || (long)(i) - (lo) >= (long)(n)) throw std::out_of_range(what); } while (0)
#else
#define REF_CHECK(i, lo, n, what) ((void)0)
#endif

// statement marker (source line of the Fortran statement): only the value-tracing build records it
#if defined(NMO_OPCOUNT)
#define REF_LINE(n) nmo_count::mark(n)
#elif defined(REF_COVERAGE)
extern unsigned char ref_cover[];  // statement coverage of the reference (tools/ref_coverage.py)
#define REF_LINE(n) (ref_cover[n] = 1)
#else
#define REF_LINE(n) ((void)0)
#endif

// column-major views; the last extent is not needed for addressing
template <class T>
struct FV1 {
  T* p; int l1;
  T& operator()(int i) const { return p[i - l1]; }
};
template <class T>
struct FV2 {
  T* p; int l1, n1, l2;
  T& operator()(int i, int j) const { return p[(long)(i - l1) + (long)n1 * (j - l2)]; }
};
template <class T>
struct FV3 {
  T* p; int l1, n1, l2, n2, l3;
  T& operator()(int i, int j, int k) const { return p[(long)(i - l1) + (long)n1 * ((j - l2) + (long)n2 * (k - l3))]; }
};

template <class T>
inline void F_ZERO(T* p, long n) { std::memset((void*)p, 0, sizeof(T) * (size_t)n); }

// an expression passed to a (by-reference) dummy argument
template <class T>
struct F_TMP {
  T v;
  template <class U> F_TMP(U x) : v((T)x) {}
  operator T&() { return v; }
};
template <class T>
inline T& F_ABSENT() { static T dummy{}; return dummy; }
template <class T>
inline bool F_IS_ABSENT(T& x) { return &x == &F_ABSENT<T>(); }

struct ref_fatal : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] inline void F_FATAL(const char* msg) { throw ref_fatal(msg); }

// ---- intrinsics ---------------------------------------------------------------------------------------------------
inline int F_ABS(int a) { return a < 0 ? -a : a; }
inline float F_ABS(float a) { return __builtin_fabsf(a); }
inline double F_ABS(double a) { return __builtin_fabs(a); }
inline float F_SQRT(float a);
inline double F_SQRT(double a) { return __builtin_sqrt(a); }
#ifdef NMO_OPCOUNT
#define REF_TICK(cls) nmo_count::tick(nmo_count::cls)  // the op-counting instantiation tallies the transcendentals too
#else
#define REF_TICK(cls) ((void)0)
#endif
#define REF_MATH1(NAME, port, lm, cls)                                                                 \
  inline float NAME(float x) { REF_TICK(cls); return ref_math_mode ? nmpm::port(x) : lm##f(x); }      \
  inline double NAME(double x) { return lm(x); }
REF_MATH1(F_EXP, expf_, __builtin_exp, EXP)
REF_MATH1(F_LOG, logf_, __builtin_log, LOG)
REF_MATH1(F_LOG10, log10f_, __builtin_log10, LOG10)
REF_MATH1(F_SIN, sinf_, __builtin_sin, SIN)
REF_MATH1(F_COS, cosf_, __builtin_cos, COS)
REF_MATH1(F_TAN, tanf_, __builtin_tan, TAN)
REF_MATH1(F_ATAN, atanf_, __builtin_atan, ATAN)
REF_MATH1(F_ASIN, asinf_, __builtin_asin, ASIN)
REF_MATH1(F_ACOS, acosf_, __builtin_acos, ACOS)
REF_MATH1(F_TANH, tanhf_, __builtin_tanh, TANH)
// The reference's own gfortran configuration (arch/makefile.in.linux.*.gcc: F90FLAGS without -O) compiles without
// optimisation, so x**y stays a libm call even for a constant y (no pow(x,2.0) -> x*x folding) and x**n is libgcc's
// __powisf2 (square-and-multiply) whatever n is -- not the power-tree expansion -O2 would substitute for a constant
// n (x**5 = x2*(x2*x) there, x*(x2*x2) here).  Both are spelled out so that the result does not depend on the
// optimisation level this file is compiled with.
float ref_powf(float x, float y);    // out of line (ref_shim.cpp): never folded
double ref_pow(double x, double y);
inline float F_POW(float x, float y) { REF_TICK(POW); return ref_math_mode ? float(nmpm::powf_(x, y)) : float(ref_powf(x, y)); }
inline double F_POW(double x, double y) { return ref_math_mode ? nmpm::pow_d(x, y) : ref_pow(x, y); }
template <class T>
inline T ref_powi(T x, int m) {
  unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
  T y = (n % 2) ? x : T(1.0f);
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return m < 0 ? T(1.0f) / y : y;
}
inline float F_POWI(float x, int n) { return ref_powi<float>(x, n); }
inline double F_POWI(double x, int n) { return ref_powi<double>(x, n); }
inline int F_IPOW(int x, int n) {
  if (n < 0) return x == 1 ? 1 : (x == -1 ? ((n & 1) ? -1 : 1) : 0);
  int r = 1;
  while (n) { if (n & 1) r *= x; x *= x; n >>= 1; }
  return r;
}
// MIN / MAX as gfortran expands them without -ffinite-math-only: m = a; if (b .op. m || isnan(m)) m = b
template <class T> inline T F_MAX(T a, T b) { return (b > a || a != a) ? b : a; }
template <class T> inline T F_MIN(T a, T b) { return (b < a || a != a) ? b : a; }
inline float F_SIGN(float a, float b) { return __builtin_copysignf(a, b); }
inline double F_SIGN(double a, double b) { return __builtin_copysign(a, b); }
inline int F_SIGN(int a, int b) { return b >= 0 ? F_ABS(a) : -F_ABS(a); }
inline int F_MOD(int a, int b) { return a % b; }
inline float F_MOD(float a, float b) { return __builtin_fmodf(a, b); }
inline double F_MOD(double a, double b) { return __builtin_fmod(a, b); }
inline int F_NINT(float a) { return (int)__builtin_lroundf(a); }
inline int F_NINT(double a) { return (int)__builtin_lround(a); }
inline float F_AINT(float a) { return __builtin_truncf(a); }
inline float F_ANINT(float a) { return __builtin_roundf(a); }
inline int F_CEILING(float a) { return (int)__builtin_ceilf(a); }
inline int F_FLOOR(float a) { return (int)__builtin_floorf(a); }
inline bool F_ISNAN(float a) { return a != a; }
inline bool F_ISNAN(double a) { return a != a; }
inline float F_SQRT(float a) { REF_TICK(SQRT); return __builtin_sqrtf(a); }
