// nmo.h — shared declarations of the CPU ORACLE (test infrastructure, NOT product code).
//
// The oracle is a scalar C++17 restatement of the reference's hot path
//   phys/module_sf_noahmpdrv.F90 :: noahmplsm          (dispatcher)
//   phys/module_sf_noahmplsm.F90 :: REDPRM, NOAHMP_SFLX (land columns)
//   phys/module_sf_noahmp_glacier.F90 :: NOAHMP_GLACIER (land-ice columns)
// routine by routine, in source order, fp32 with the one fp64 temporary the reference has.
//
// PARITY PINNED BY TRANSLATION: the reference ships no golden vectors for this path and the build image has no Fortran
// compiler, so the reference's own Fortran text is machine-translated to C++ (oracle/ref/f90cxx.py, a translator of the
// language), compiled into oracle/_ref/libnoahmp_ref.so and this restatement is held to it bit for bit
// (tests/test_reference_pin.py: noahmplsm on C1..C4 populations and every accepted option value, NOAHMP_INIT,
// WTABLE_mmf_noahmp, both math modes; vectors it produced are committed under tests/golden/).  It is not a Fortran
// compiler's output: DESIGN.md section 5 says what the translation assumes (gfortran's unoptimised configuration, which
// is the reference's own; its MIN/MAX expansion).  Also pinned by the model's conservation checks (ERROR /
// ERROR_GLACIER) and by physics-derived known answers (tests/test_oracle_*.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../include/noahmp_b200.h"
#include "../noahmp_b200/csrc/nmp_math.h"

namespace nmo {

// ---- math backend: 0 = host libm (what the reference links), 1 = portable nmp_math.h --------
extern int g_math_mode;
// op-counting instantiation (nmo_count.h): every transcendental call is tallied by class; a no-op otherwise
#ifndef NMO_TICK
#define NMO_TICK(cls) ((void)0)
#endif
inline float EXP(float x)   { NMO_TICK(EXP); return g_math_mode ? nmpm::expf_(x)   : std::exp(x); }
inline float LOG(float x)   { NMO_TICK(LOG); return g_math_mode ? nmpm::logf_(x)   : std::log(x); }
inline float LOG10(float x) { NMO_TICK(LOG10); return g_math_mode ? nmpm::log10f_(x) : std::log10(x); }
// x**y with a REAL exponent is a libm call in the reference's own gfortran configuration (arch/makefile.in.*.gcc sets
// no -O, so no pow(x, 2.0) -> x*x folding): the libm leg goes through out-of-line functions (nmo_driver.cpp) that the
// compiler cannot fold at a call site with a constant exponent either
float powf_libm(float x, float y);
double pow_libm(double x, double y);
inline float POW(float x, float y) { NMO_TICK(POW); return g_math_mode ? float(nmpm::powf_(x, y)) : float(powf_libm(x, y)); }
inline double DPOW(double x, double y) { NMO_TICK(DPOW); return g_math_mode ? nmpm::pow_d(x, y) : pow_libm(x, y); }
inline float ATAN(float x)  { NMO_TICK(ATAN); return g_math_mode ? nmpm::atanf_(x)  : std::atan(x); }
inline float TAN(float x)   { NMO_TICK(TAN); return g_math_mode ? nmpm::tanf_(x)   : std::tan(x); }
inline float COS(float x)   { NMO_TICK(COS); return g_math_mode ? nmpm::cosf_(x)   : std::cos(x); }
inline float SIN(float x)   { NMO_TICK(SIN); return g_math_mode ? nmpm::sinf_(x)   : std::sin(x); }
inline float ASIN(float x)  { NMO_TICK(ASIN); return g_math_mode ? nmpm::asinf_(x)  : std::asin(x); }
inline float ACOS(float x)  { NMO_TICK(ACOS); return g_math_mode ? nmpm::acosf_(x)  : std::acos(x); }
inline float TANH(float x)  { NMO_TICK(TANH); return g_math_mode ? nmpm::tanhf_(x)  : std::tanh(x); }
inline float SQRT(float x)  { NMO_TICK(SQRT); return std::sqrt(x); }
// integer power x**n exactly as libgcc's __powisf2 evaluates it (square-and-multiply)
inline float POWI(float x, int m) {
  unsigned n = m < 0 ? -(unsigned)m : (unsigned)m;
  float y = (n & 1) ? x : float(1.0f);
  while (n >>= 1) { x = x * x; if (n & 1) y *= x; }
  return m < 0 ? 1.0f / y : y;
}
// Fortran MIN/MAX/SIGN semantics.  gfortran expands MAX(a, b) as  m = a; if (b > m || isnan(m)) m = b  (trans-intrinsic.c,
// without -ffinite-math-only): a NaN first argument is replaced by the second -- GROUNDWATER_INIT relies on it when its
// Newton iteration for the deep soil moisture diverges: MAX(SMC, 1.E-4) is then 1.E-4, not NaN.
#ifdef NMO_OPCOUNT
inline bool ISNAN_(nmo_count::true_float a) { return a != a; }  // not an operation of the algorithm: not counted
#else
inline bool ISNAN_(float a) { return a != a; }
#endif
inline float MIN(float a, float b) { return (b < a || ISNAN_(a)) ? b : a; }
inline float MAX(float a, float b) { return (b > a || ISNAN_(a)) ? b : a; }
inline float MIN3(float a, float b, float c) { return MIN(MIN(a, b), c); }
inline int IMIN(int a, int b) { return b < a ? b : a; }
inline int IMAX(int a, int b) { return b > a ? b : a; }
inline float ABS(float a) { return std::fabs(a); }
inline float SIGN(float a, float b) { return (b >= 0.0f) ? std::fabs(a) : -std::fabs(a); }

// Fortran array with explicit lower bound
template <int LO, int HI>
struct FA {
  float v[HI - LO + 1];
  float& operator()(int i) { return v[i - LO]; }
  const float& operator()(int i) const { return v[i - LO]; }
  void fill(float x) { for (auto& e : v) e = x; }
};
template <int LO, int HI>
struct IA {
  int v[HI - LO + 1];
  int& operator()(int i) { return v[i - LO]; }
  const int& operator()(int i) const { return v[i - LO]; }
};
constexpr int NSOIL = NOAHMP_NSOIL;
constexpr int NSNOW = NOAHMP_NSNOW;
using ASnSo = FA<-NSNOW + 1, NSOIL>;  // (-2:4)
using ASoil = FA<1, NSOIL>;           // (1:4)
using ASnow = FA<-NSNOW + 1, 0>;      // (-2:0)
using ABand = FA<1, 2>;

// noahmp_globals constants (noahmplsm.F90:12-28, 180-188)
constexpr float GRAV = 9.80616f, SB = 5.67E-08f, VKC = 0.40f, TFRZ = 273.16f, HSUB = 2.8440E06f,
                HVAP = 2.5104E06f, HFUS = 0.3336E06f, CWAT = 4.188E06f, CICE = 2.094E06f,
                CPAIR = 1004.64f, TKWAT = 0.6f, TKICE = 2.2f, TKAIR = 0.023f, RAIR = 287.04f,
                RW = 461.269f, DENH2O = 1000.f, DENICE = 917.f;
constexpr float TIMEAN = 10.5f, FSATMX = 0.38f, M_MELT = 2.50f, Z0SNO = 0.002f, SSI = 0.03f,
                SWEMX = 1.00f;

// noahmp_options (noahmplsm.F90:9352-9388)
struct Options {
  int DVEG, OPT_CRS, OPT_BTR, OPT_RUN, OPT_SFC, OPT_FRZ, OPT_INF, OPT_RAD, OPT_ALB, OPT_SNF,
      OPT_TBOT, OPT_STC;
};

// per-column module globals written by REDPRM (noahmplsm.F90:33-38, 69-95)
struct Params {
  int NROOT;
  float RGL, RSMIN, HS, RSMAX, TOPT;
  float BEXP, SMCDRY, F1, SMCMAX, SMCREF, PSISAT, DKSAT, DWSAT, SMCWLT, QUARTZ;
  float SLOPE, CSOIL, ZBOT, CZIL, KDT, FRZX;
};

struct Ctx {
  const noahmp_tables* T;
  Options O;
  Params P;
  int err_code = 0;   // first fatal condition hit in this column (enum in noahmp_b200.h)
  float err_value = 0.f;
  void fatal(int code, float v) { if (!err_code) { err_code = code; err_value = v; } }
};

// 1-based accessors for the table struct (tblf: the tables' plain fp32 words, also in the op-counting instantiation)
#ifdef NMO_OPCOUNT
typedef nmo_count::true_float tblf;
#else
typedef float tblf;
#endif
inline float TV1(const tblf* a, int vegtyp) { return a[vegtyp - 1]; }
inline float TV2(const tblf (*a)[NOAHMP_MVT], int vegtyp, int k) { return a[k - 1][vegtyp - 1]; }

int REDPRM(Ctx& c, int VEGTYP, int SOILTYP, int SLOPETYP, const ASoil& ZSOIL, int ISURBAN);

// everything NOAHMP_SFLX exchanges with the dispatcher (noahmplsm.F90:518-548)
struct SflxIO {
  // IN
  int ILOC, JLOC; float LAT; int YEARLEN; float JULIAN, COSZ, DT, DX, DZ8W; ASoil ZSOIL;
  float SHDFAC, SHDMAX; int VEGTYP, ISURBAN, ICE, IST, ISC; ASoil SMCEQ; int IZ0TLND;
  float SFCTMP, SFCPRS, PSFC, UU, VV, Q2, QC, SOLDN, LWDN, PRCP, TBOT, CO2AIR, O2AIR, FOLN;
  ASnow FICEOLD; float PBLH, ZLVL;
  // INOUT
  float ALBOLD, SNEQVO; ASnSo STC; ASoil SH2O, SMC; float TAH, EAH, FWET, CANLIQ, CANICE, TV, TG,
      QSFC, QSNOW; int ISNOW; ASnSo ZSNSO; float SNOWH, SNEQV; ASnow SNICE, SNLIQ;
  float ZWT, WA, WT, WSLAKE, LFMASS, RTMASS, STMASS, WOOD, STBLCP, FASTCP, LAI, SAI, CM, CH, TAUSS,
      SMCWTD, DEEPRECH, RECH;
  // OUT
  float FSA, FSR, FIRA, FSH, SSOIL, FCEV, FGEV, FCTR, ECAN, ETRAN, EDIR, TRAD, TGB, TGV, T2MV, T2MB,
      Q2V, Q2B, RUNSRF, RUNSUB, APAR, PSN, SAV, SAG, FSNO, NEE, GPP, NPP, FVEG, ALBEDO, QSNBOT,
      PONDING, PONDING1, PONDING2, RSSUN, RSSHA, BGAP, WGAP, CHV, CHB, EMISSI, SHG, SHC, SHB, EVG,
      EVB, GHV, GHB, IRG, IRC, IRB, TR, EVC, CHLEAF, CHUC, CHV2, CHB2, FPICE;
  // diagnostics kept for tests (not part of the reference's list)
  float ERRWAT, ERRENG, ERRSW; IA<-NSNOW + 1, NSOIL> IMELT; int VEGE_ITERS;
};
void NOAHMP_SFLX(Ctx& c, SflxIO& s);

// NOAHMP_GLACIER argument list (glacier.F90:150-167)
struct GlacIO {
  int ILOC, JLOC; float COSZ, DT, SFCTMP, SFCPRS, UU, VV, Q2, SOLDN, PRCP, LWDN, TBOT, ZLVL;
  ASnow FICEOLD; ASoil ZSOIL;
  float QSNOW, SNEQVO, ALBOLD, CM, CH; int ISNOW; float SNEQV; ASoil SMC; ASnSo ZSNSO; float SNOWH;
  ASnow SNICE, SNLIQ; float TG; ASnSo STC; ASoil SH2O; float TAUSS, QSFC;
  float FSA, FSR, FIRA, FSH, FGEV, SSOIL, TRAD, EDIR, RUNSRF, RUNSUB, SAG, ALBEDO, QSNBOT, PONDING,
      PONDING1, PONDING2, T2M, Q2E, EMISSI, FPICE, CH2B;
  float ERRWAT, ERRENG, ERRSW;
};
void NOAHMP_GLACIER(Ctx& c, GlacIO& g);

}  // namespace nmo
