// nmp_common.cuh — shared device-side definitions for the Noah-MP column kernels (sm_100a).
//
// One thread owns one land (or glacier) column.  The per-column working set (7 snow+soil layers,
// ~80 prognostic words, ~22 soil/veg parameters) lives in registers / L1-resident local memory.
// Compile-time switches:
//   NMP_PARITY=1  : every transcendental goes through nmp_math.h (bit-identical to the CPU oracle's
//                   M1 mode); the translation unit is also compiled with -fmad=false.
//   NMP_PARITY=0  : CUDA libdevice single-precision functions, FMA contraction allowed ("fast").
// Physics options (noahmp.namelist opt_*) are template constants when the kernel is instantiated
// for a fixed option set, or read from kernel parameters when the template value is 0.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/noahmp_b200.h"
#include "nmp_math.h"

#ifndef NMP_PARITY
#define NMP_PARITY 0
#endif
#ifndef NMP_FASTMATH
#define NMP_FASTMATH 0
#endif
// Threads per block of the physics kernels and the phase barrier: with NMP_PHASE_SYNC the warps of a block
// walk through the (several hundred KB of) straight-line physics code in step, so an instruction-cache line
// fetched from L2 by the first warp is reused by the others (DESIGN.md "instruction fetch").
#ifndef NMP_BLOCK
#define NMP_BLOCK 128
#endif
#ifndef NMP_MINBLOCKS
#define NMP_MINBLOCKS 1
#endif
#ifndef NMP_PHASE_SYNC
#define NMP_PHASE_SYNC 0
#endif
// NMP_PHASE_SYNC: 0 = no barriers, 1 = barriers between all ~16 phases, 2 = only between the major phases
#if NMP_PHASE_SYNC == 1
#define NMP_PHASE() __syncthreads()
#define NMP_PHASE_MAJOR() __syncthreads()
#elif NMP_PHASE_SYNC == 2
#define NMP_PHASE() ((void)0)
// NMP_PHASE_GROUP = g > 0 (experiment): only groups of g threads of a block keep in step (named barriers 1..), so a
// slow warp holds back g/32 - 1 others instead of the whole block
#if defined(NMP_PHASE_GROUP) && NMP_PHASE_GROUP > 0
#define NMP_PHASE_MAJOR() asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / NMP_PHASE_GROUP)), "n"(NMP_PHASE_GROUP) : "memory")
#else
#define NMP_PHASE_MAJOR() __syncthreads()
#endif
#else
#define NMP_PHASE() ((void)0)
#define NMP_PHASE_MAJOR() ((void)0)
#endif

#define NMP_DEV __device__ __forceinline__
#define NMP_DEVN __device__ __noinline__

namespace nmp {

// ---- constants: noahmp_globals (noahmplsm.F90:12-28, 180-188), identical in glacier.F90:10-26 ----
constexpr float GRAV = 9.80616f, SB = 5.67E-08f, VKC = 0.40f, TFRZ = 273.16f, HSUB = 2.8440E06f,
                HVAP = 2.5104E06f, HFUS = 0.3336E06f, CWAT = 4.188E06f, CICE = 2.094E06f,
                CPAIR = 1004.64f, TKWAT = 0.6f, TKICE = 2.2f, TKAIR = 0.023f, RAIR = 287.04f,
                RW = 461.269f, DENH2O = 1000.f, DENICE = 917.f;
constexpr float TIMEAN = 10.5f, FSATMX = 0.38f, M_MELT = 2.50f, Z0SNO = 0.002f, SSI = 0.03f, SWEMX = 1.00f;
constexpr int NSOIL = NOAHMP_NSOIL, NSNOW = NOAHMP_NSNOW, NLAY = NSOIL + NSNOW;

// ---- math front end ----------------------------------------------------------------------------
#if NMP_PARITY
NMP_DEV float EXP(float x) { return nmpm::expf_(x); }
NMP_DEV float LOG(float x) { return nmpm::logf_(x); }
NMP_DEV float LOG10(float x) { return nmpm::log10f_(x); }
NMP_DEV float POW(float x, float y) { return nmpm::powf_(x, y); }
NMP_DEV double DPOW(double x, double y) { return nmpm::pow_d(x, y); }
NMP_DEV float ATAN(float x) { return nmpm::atanf_(x); }
NMP_DEV float TAN(float x) { return nmpm::tanf_(x); }
NMP_DEV float COS(float x) { return nmpm::cosf_(x); }
NMP_DEV float SIN(float x) { return nmpm::sinf_(x); }
NMP_DEV float ACOS(float x) { return nmpm::acosf_(x); }
NMP_DEV float TANH(float x) { return nmpm::tanhf_(x); }
NMP_DEV float SQRT(float x) { return __fsqrt_rn(x); }
NMP_DEV float DIV(float a, float b) { return __fdiv_rn(a, b); }
#elif NMP_FASTMATH
// Production arithmetic: SFU-based exp2/log2 (MUFU.EX2 / MUFU.LG2) instead of the ~100-instruction libdevice
// powf/expf/logf expansions; relative error <= ~1e-6 over the ranges the physics uses (|y*log2 x| < 60).
// The translation unit is compiled with -prec-div=false -prec-sqrt=false as well.  Rarely used functions
// (opt_rad=1 geometry, TANH of the snow-cover fraction, ATAN of the unstable profile) stay on libdevice.
NMP_DEV float EXP(float x) { return __expf(x); }
NMP_DEV float LOG(float x) { return __logf(x); }
NMP_DEV float LOG10(float x) { return __log10f(x); }
NMP_DEV float POW(float x, float y) { return __powf(x, y); }
NMP_DEV double DPOW(double x, double y) { return pow(x, y); }
NMP_DEV float ATAN(float x) { return atanf(x); }
NMP_DEV float TAN(float x) { return tanf(x); }
NMP_DEV float COS(float x) { return cosf(x); }
NMP_DEV float SIN(float x) { return sinf(x); }
NMP_DEV float ACOS(float x) { return acosf(x); }
NMP_DEV float TANH(float x) { return tanhf(x); }
NMP_DEV float SQRT(float x) { return sqrtf(x); }
NMP_DEV float DIV(float a, float b) { return a / b; }
#else
NMP_DEV float EXP(float x) { return expf(x); }
NMP_DEV float LOG(float x) { return logf(x); }
NMP_DEV float LOG10(float x) { return log10f(x); }
NMP_DEV float POW(float x, float y) { return powf(x, y); }
NMP_DEV double DPOW(double x, double y) { return pow(x, y); }
NMP_DEV float ATAN(float x) { return atanf(x); }
NMP_DEV float TAN(float x) { return tanf(x); }
NMP_DEV float COS(float x) { return cosf(x); }
NMP_DEV float SIN(float x) { return sinf(x); }
NMP_DEV float ACOS(float x) { return acosf(x); }
NMP_DEV float TANH(float x) { return tanhf(x); }
NMP_DEV float SQRT(float x) { return sqrtf(x); }
NMP_DEV float DIV(float a, float b) { return a / b; }
#endif

// x**2. (a REAL exponent): a libm pow call in the reference's own gfortran configuration (no -O, so no folding to x*x;
// the translated reference of oracle/ref pins this).  The production build multiplies: __powf has no negative bases.
#if NMP_FASTMATH
NMP_DEV float POWR2(float x) { return x * x; }
#else
NMP_DEV float POWR2(float x) { return POW(x, 2.0f); }
#endif

// x**n for integer n as libgcc's __powisf2 evaluates it (what gfortran emits for REAL**INTEGER)
NMP_DEV float POW2(float x) { return x * x; }
NMP_DEV float POW3(float x) { return x * (x * x); }
NMP_DEV float POW4(float x) { float t = x * x; return t * t; }
NMP_DEV float POW5(float x) { float t = x * x; return x * (t * t); }

// Fortran MIN / MAX / SIGN / ABS
#if NMP_FASTMATH
NMP_DEV float MIN(float a, float b) { return fminf(a, b); }  // one FMNMX; differs from the select only for NaN
NMP_DEV float MAX(float a, float b) { return fmaxf(a, b); }
#else
// gfortran's expansion (m = a; if (b .op. m || isnan(m)) m = b): a NaN first argument gives way to the second
NMP_DEV float MIN(float a, float b) { return (b < a || a != a) ? b : a; }
NMP_DEV float MAX(float a, float b) { return (b > a || a != a) ? b : a; }
#endif
NMP_DEV float ABS(float a) { return fabsf(a); }
NMP_DEV float SIGN(float a, float b) { return (b >= 0.0f) ? fabsf(a) : -fabsf(a); }

// ---- layer arrays with Fortran bounds -----------------------------------------------------------
// L7: (-2:4) snow+soil, S4: (1:4) soil, N3: (-2:0) snow.  operator() takes the Fortran index.
struct L7 {
  float v[NLAY];
  NMP_DEV float& operator()(int k) { return v[k + 2]; }
  NMP_DEV const float& operator()(int k) const { return v[k + 2]; }
};
struct S4 {
  float v[NSOIL];
  NMP_DEV float& operator()(int k) { return v[k - 1]; }
  NMP_DEV const float& operator()(int k) const { return v[k - 1]; }
};
struct N3 {
  float v[NSNOW];
  NMP_DEV float& operator()(int k) { return v[k + 2]; }
  NMP_DEV const float& operator()(int k) const { return v[k + 2]; }
};
struct I7 {
  int v[NLAY];
  NMP_DEV int& operator()(int k) { return v[k + 2]; }
  NMP_DEV const int& operator()(int k) const { return v[k + 2]; }
};
struct B2 {
  float v[2];
  NMP_DEV float& operator()(int k) { return v[k - 1]; }
  NMP_DEV const float& operator()(int k) const { return v[k - 1]; }
};

// "Is layer k the top active layer (k == ISNOW+1)?" for a compile-time k, asked through a bit mask whose value is
// hidden from the optimiser.  Written as `k == ISNOW + 1`, NVVM rewrites the guarded access A(k) into the
// run-time indexed A(ISNOW+1); one such dynamic index anywhere keeps the whole column structure (~200 words) in
// local memory instead of registers (profiles/r01_notes.md).
NMP_DEV unsigned opaque_u(unsigned x) {
  asm volatile("" : "+r"(x));
  return x;
}
struct TopLayer {
  unsigned bit;
  NMP_DEV explicit TopLayer(int isnow) : bit(opaque_u(1u << (isnow + 3))) {}
  NMP_DEV bool is(int k) const { return ((bit >> (k + 2)) & 1u) != 0u; }
};

// ---- physics options ----------------------------------------------------------------------------
struct Opts {  // runtime values (kernel parameter)
  int dveg, crs, btr, run, sfc, frz, inf, rad, alb, snf, tbot, stc;
};
// Compile-time option set; a 0 entry defers to the runtime value.
template <int DVEG, int CRS, int BTR, int RUN, int SFC, int FRZ, int INF, int RAD, int ALB, int SNF, int TBOT,
          int STC>
struct OptSet {
  static constexpr int dveg = DVEG, crs = CRS, btr = BTR, run = RUN, sfc = SFC, frz = FRZ, inf = INF, rad = RAD,
                       alb = ALB, snf = SNF, tbot = TBOT, stc = STC;
};
using OptRuntime = OptSet<0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0>;
#define NMP_OPT(name) ((O::name != 0) ? O::name : c.o.name)

// ---- per-column parameters written by REDPRM (noahmplsm.F90:9202-9349) ---------------------------
struct Prm {
  int NROOT;
  float RGL, RSMIN, HS, RSMAX, TOPT;
  float BEXP, SMCDRY, F1, SMCMAX, SMCREF, PSISAT, DKSAT, DWSAT, SMCWLT, QUARTZ;
  float SLOPE, CSOIL, ZBOT, CZIL, KDT, FRZX;
};

// execution context of one column: tables (shared memory), options, parameters, error latch
struct Ctx {
  const noahmp_tables* T;
  Opts o;
  Prm P;
  int err;
  float errv;
  NMP_DEV void fatal(int code, float v) {
    if (!err) { err = code; errv = v; }
  }
};

NMP_DEV float tv1(const float* a, int vegtyp) { return a[vegtyp - 1]; }

}  // namespace nmp
