// nmo_count.h — OP-COUNTING INSTANTIATION of the oracle (test / measurement infrastructure).
//
// Force-included (-include nmo_count.h) in front of the oracle's translation units: after the standard headers and
// the C-ABI header have been seen with the real `float`, the keyword is redefined to a wrapper type that performs the
// same fp32 arithmetic and counts every operation.  The physics sources stay untouched; the results are bit-identical
// to the ordinary oracle (tests/test_opcount.py checks that), and the counters give the ALGORITHMIC work per
// column-step — additions, multiplications, divisions, comparisons / min / max, and each transcendental by class —
// that bench.py's compute roofline uses as its numerator (profiles/r02_opcount.json, tools/opcount.py).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>
#include "../include/noahmp_b200.h"
#include "../noahmp_b200/csrc/nmp_math.h"

#define NMO_OPCOUNT 1

namespace nmo_count {

typedef float true_float;  // the parameter tables and the caller's arrays stay plain fp32 words

enum Op { ADD = 0, MUL, DIV, CMP, EXP, LOG, LOG10, POW, DPOW, SQRT, ATAN, TAN, COS, SIN, ASIN, ACOS, TANH, NOPS };
struct Counters { unsigned long long n[NOPS]; };
extern thread_local Counters tl;
inline void tick(Op o) { ++tl.n[o]; }
// optional value trace (debugging aid of the reference pin, tools/ref_trace_diff.py): every arithmetic result with its
// operands, and statement markers (op = -1, a = source line) where the translated reference emits them
struct TraceRec { int op; unsigned a, b, r; };
extern thread_local std::vector<TraceRec>* trace;
inline unsigned fbits(float x) { unsigned u; std::memcpy(&u, &x, 4); return u; }
inline void rec(Op o, float a, float b, float r) { if (trace) trace->push_back(TraceRec{(int)o, fbits(a), fbits(b), fbits(r)}); }
inline void mark(int line) { if (trace) trace->push_back(TraceRec{-1, (unsigned)line, 0u, 0u}); }

struct Real {
  float v;
  Real() = default;
  constexpr Real(float x) : v(x) {}
  constexpr Real(double x) : v((float)x) {}
  constexpr Real(int x) : v((float)x) {}
  constexpr Real(unsigned x) : v((float)x) {}
  constexpr Real(long x) : v((float)x) {}
  constexpr Real(unsigned long x) : v((float)x) {}
  constexpr operator float() const { return v; }
  Real operator-() const { return Real(-v); }
  Real operator+() const { return *this; }
  Real& operator+=(Real o) { tick(ADD); float r = v + o.v; rec(ADD, v, o.v, r); v = r; return *this; }
  Real& operator-=(Real o) { tick(ADD); float r = v - o.v; rec(ADD, v, -o.v, r); v = r; return *this; }
  Real& operator*=(Real o) { tick(MUL); float r = v * o.v; rec(MUL, v, o.v, r); v = r; return *this; }
  Real& operator/=(Real o) { tick(DIV); float r = v / o.v; rec(DIV, v, o.v, r); v = r; return *this; }
};
static_assert(std::is_trivially_copyable<Real>::value && sizeof(Real) == 4, "Real is a float");

template <class T>
using arith = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;

#define NMO_BINOP(op, cls, sgn)                                                                   \
  inline Real operator op(Real a, Real b) { tick(cls); float r = a.v op b.v; rec(cls, a.v, sgn b.v, r); return Real(r); } \
  template <class T, arith<T> = 0> inline Real operator op(Real a, T b) { return a op Real((float)b); } \
  template <class T, arith<T> = 0> inline Real operator op(T a, Real b) { return Real((float)a) op b; }
NMO_BINOP(+, ADD, +)
NMO_BINOP(-, ADD, -)
NMO_BINOP(*, MUL, +)
NMO_BINOP(/, DIV, +)
#undef NMO_BINOP
#define NMO_CMPOP(op)                                                                             \
  inline bool operator op(Real a, Real b) { tick(CMP); return a.v op b.v; }                        \
  template <class T, arith<T> = 0> inline bool operator op(Real a, T b) { tick(CMP); return a.v op (float)b; } \
  template <class T, arith<T> = 0> inline bool operator op(T a, Real b) { tick(CMP); return (float)a op b.v; }
NMO_CMPOP(<)
NMO_CMPOP(<=)
NMO_CMPOP(>)
NMO_CMPOP(>=)
NMO_CMPOP(==)
NMO_CMPOP(!=)
#undef NMO_CMPOP

}  // namespace nmo_count

// transcendental hooks used by the math front end of nmo.h
#define NMO_TICK(cls) nmo_count::tick(nmo_count::cls)

#define float nmo_count::Real
