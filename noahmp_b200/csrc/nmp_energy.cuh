// nmp_energy.cuh — device code of the ENERGY subtree of NOAHMP_SFLX
// (phys/module_sf_noahmplsm.F90:1231-6377): THERMOPROP/CSNOW/TDFCND, RADIATION (ALBEDO, SNOW_AGE,
// SNOWALB_BATS/CLASS, GROUNDALB, TWOSTREAM, SURRAD), VEGE_FLUX / BARE_FLUX with SFCDIF1/SFCDIF2, RAGRB,
// ESAT, STOMATA / CANRES, TSNOSOI (HRT, HSTEP, ROSR12) and PHASECHANGE (FRH2O).
//
// GPU shape: one thread = one column; all layer loops run over the fixed range -2..4 with an
// "active layer" predicate (k > ISNOW) and are fully unrolled, so layer arrays are indexed by
// compile-time constants and stay in registers; accesses to the top active layer (ISNOW+1) go through
// the 4-way select top7().  Expression order follows the Fortran source so that the parity build
// (NMP_PARITY=1, -fmad=false) is bit-identical with the CPU oracle.
#pragma once
#include "nmp_common.cuh"

namespace nmp {

// value of a (-2:4) array at the top active layer ISNOW+1 (ISNOW in -3..0)
NMP_DEV float top7(const L7& a, int isnow) {
  return isnow == 0 ? a.v[3] : (isnow == -1 ? a.v[2] : (isnow == -2 ? a.v[1] : a.v[0]));
}

NMP_DEV float TDC(float T) { return MIN(50.f, MAX(-50.f, (T - TFRZ))); }

// noahmplsm.F90:5272-5321 (identical in glacier.F90:1150-1199): four degree-6 polynomials in T (deg C) — saturation
// vapour pressure over water / ice and their temperature derivatives.  Every caller keeps only the pair the sign of
// T selects (`IF (T .GT. 0.) THEN ESTV = ESATW ... ELSE ESTV = ESATI`), so the two entry points below evaluate only
// what is kept: the warp skips the water (ice) polynomials when no lane is above (below) freezing.  Each polynomial is
// the reference's Horner form, so the selected values have the bits the four-output routine gives.
namespace esat_c {
constexpr float EA0 = 6.107799961f, EA1 = 4.436518521E-01f, EA2 = 1.428945805E-02f, EA3 = 2.650648471E-04f,
                EA4 = 3.031240396E-06f, EA5 = 2.034080948E-08f, EA6 = 6.136820929E-11f;
constexpr float EB0 = 6.109177956f, EB1 = 5.034698970E-01f, EB2 = 1.886013408E-02f, EB3 = 4.176223716E-04f,
                EB4 = 5.824720280E-06f, EB5 = 4.838803174E-08f, EB6 = 1.838826904E-10f;
constexpr float EC0 = 4.438099984E-01f, EC1 = 2.857002636E-02f, EC2 = 7.938054040E-04f, EC3 = 1.215215065E-05f,
                EC4 = 1.036561403E-07f, EC5 = 3.532421810e-10f, EC6 = -7.090244804E-13f;
constexpr float ED0 = 5.030305237E-01f, ED1 = 3.773255020E-02f, ED2 = 1.267995369E-03f, ED3 = 2.477563108E-05f,
                ED4 = 3.005693132E-07f, ED5 = 2.158542548E-09f, ED6 = 7.131097725E-12f;
}  // namespace esat_c
NMP_DEV float ESAT_W(float T) {
  using namespace esat_c;
  return 100.f * (EA0 + T * (EA1 + T * (EA2 + T * (EA3 + T * (EA4 + T * (EA5 + T * EA6))))));
}
NMP_DEV float ESAT_I(float T) {
  using namespace esat_c;
  return 100.f * (EB0 + T * (EB1 + T * (EB2 + T * (EB3 + T * (EB4 + T * (EB5 + T * EB6))))));
}
NMP_DEV float DESAT_W(float T) {
  using namespace esat_c;
  return 100.f * (EC0 + T * (EC1 + T * (EC2 + T * (EC3 + T * (EC4 + T * (EC5 + T * EC6))))));
}
NMP_DEV float DESAT_I(float T) {
  using namespace esat_c;
  return 100.f * (ED0 + T * (ED1 + T * (ED2 + T * (ED3 + T * (ED4 + T * (ED5 + T * ED6))))));
}
NMP_DEV void ESAT(float T, float& ESW, float& ESI, float& DESW, float& DESI) {
  ESW = ESAT_W(T); ESI = ESAT_I(T); DESW = DESAT_W(T); DESI = DESAT_I(T);
}
// NMP_ESAT_SEL=1 skips the unused pair behind warp votes: measured SLOWER (2.333 vs 2.248 ms per 4.4 M columns over a
// diurnal cycle, profiles/r02_notes.md) — the kernel is latency-bound and the four independent Horner chains overlap
// for free, while the votes and branches add to the dependent path.  Default: evaluate and select.
#ifndef NMP_ESAT_SEL
#define NMP_ESAT_SEL 0
#endif
// ES = (T > 0) ? ESATW : ESATI and the matching derivative
NMP_DEV void ESAT_SEL(float T, float& ES, float& DES) {
#if NMP_ESAT_SEL
  const bool w = T > 0.f;
  const unsigned m = __activemask();
  float ew = 0.f, dw = 0.f, ei = 0.f, di = 0.f;
  if (__any_sync(m, w)) { ew = ESAT_W(T); dw = DESAT_W(T); }
  if (__any_sync(m, !w)) { ei = ESAT_I(T); di = DESAT_I(T); }
  ES = w ? ew : ei;
  DES = w ? dw : di;
#else
  float a, b, c, d;
  ESAT(T, a, b, c, d);
  if (T > 0.f) { ES = a; DES = c; } else { ES = b; DES = d; }
#endif
}
NMP_DEV float ESAT_SEL1(float T) {
#if NMP_ESAT_SEL
  const bool w = T > 0.f;
  const unsigned m = __activemask();
  float ew = 0.f, ei = 0.f;
  if (__any_sync(m, w)) ew = ESAT_W(T);
  if (__any_sync(m, !w)) ei = ESAT_I(T);
  return w ? ew : ei;
#else
  return T > 0.f ? ESAT_W(T) : ESAT_I(T);
#endif
}

// noahmplsm.F90:1957-2011
NMP_DEV void CSNOW(int ISNOW, const N3& SNICE, const N3& SNLIQ, const L7& DZSNSO, N3& TKSNO, N3& CVSNO,
                   N3& SNICEV, N3& SNLIQV, N3& EPORE) {
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ) {
    if (IZ > ISNOW) {
      SNICEV(IZ) = MIN(1.f, SNICE(IZ) / (DZSNSO(IZ) * DENICE));
      EPORE(IZ) = 1.f - SNICEV(IZ);
      SNLIQV(IZ) = MIN(EPORE(IZ), SNLIQ(IZ) / (DZSNSO(IZ) * DENH2O));
      float BDSNOI = (SNICE(IZ) + SNLIQ(IZ)) / DZSNSO(IZ);
      CVSNO(IZ) = CICE * SNICEV(IZ) + CWAT * SNLIQV(IZ);
      TKSNO(IZ) = 3.2217E-6f * POWR2(BDSNOI);
    }
  }
}

// noahmplsm.F90:2014-2118
NMP_DEV float TDFCND(const Prm& P, float SMC, float SH2O) {
  float SATRATIO = SMC / P.SMCMAX;
  const float THKW = 0.57f, THKO = 2.0f, THKQTZ = 7.7f;
  float THKS = POW(THKQTZ, P.QUARTZ) * POW(THKO, 1.f - P.QUARTZ);
  float XUNFROZ = SH2O / SMC;
  float XU = XUNFROZ * P.SMCMAX;
  float THKSAT = POW(THKS, 1.f - P.SMCMAX) * POW(TKICE, P.SMCMAX - XU) * POW(THKW, XU);
  float GAMMD = (1.f - P.SMCMAX) * 2700.f;
  float THKDRY = (0.135f * GAMMD + 64.7f) / (2700.f - 0.947f * GAMMD);
  float AKE;
  if ((SH2O + 0.0005f) < SMC) {
    AKE = SATRATIO;
  } else {
    if (SATRATIO > 0.1f) AKE = LOG10(SATRATIO) + 1.0f;
    else AKE = 0.0f;
  }
  return AKE * (THKSAT - THKDRY) + THKDRY;
}

// noahmplsm.F90:1845-1954
NMP_DEV void THERMOPROP(const Ctx& c, int ISNOW, int IST, const L7& DZSNSO, float DT, float SNOWH,
                        const N3& SNICE, const N3& SNLIQ, float CSOIL, const S4& SMC, const S4& SH2O,
                        const L7& STC, bool URBAN, L7& DF, L7& HCPCT, N3& SNICEV, N3& SNLIQV, N3& EPORE,
                        L7& FACT) {
  N3 CVSNO, TKSNO;
  CSNOW(ISNOW, SNICE, SNLIQ, DZSNSO, TKSNO, CVSNO, SNICEV, SNLIQV, EPORE);
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ)
    if (IZ > ISNOW) { DF(IZ) = TKSNO(IZ); HCPCT(IZ) = CVSNO(IZ); }
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    float SICE = SMC(IZ) - SH2O(IZ);
    HCPCT(IZ) = SH2O(IZ) * CWAT + (1.0f - c.P.SMCMAX) * CSOIL + (c.P.SMCMAX - SMC(IZ)) * CPAIR + SICE * CICE;
    DF(IZ) = TDFCND(c.P, SMC(IZ), SH2O(IZ));
  }
  if (URBAN) {
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) DF(IZ) = 3.24f;
  }
  if (IST == 2) {
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) {
      if (STC(IZ) > TFRZ) { HCPCT(IZ) = CWAT; DF(IZ) = TKWAT; }
      else { HCPCT(IZ) = CICE; DF(IZ) = TKICE; }
    }
  }
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ)
    if (IZ > ISNOW) FACT(IZ) = DT / (HCPCT(IZ) * DZSNSO(IZ));
  if (ISNOW == 0) DF(1) = (DF(1) * DZSNSO(1) + 0.35f * SNOWH) / (SNOWH + DZSNSO(1));
  else DF(1) = (DF(1) * DZSNSO(1) + DF(0) * DZSNSO(0)) / (DZSNSO(0) + DZSNSO(1));
}

// noahmplsm.F90:2547-2596 (identical in glacier.F90:794-845)
NMP_DEV void SNOW_AGE(float DT, float TG, float SNEQVO, float SNEQV, float& TAUSS, float& FAGE) {
  if (SNEQV <= 0.0f) {
    TAUSS = 0.f;
  } else if (SNEQV > 800.f) {
    TAUSS = 0.f;
  } else {
    float DELA0 = 1.E-6f * DT;
    float ARG = 5.E3f * (1.f / TFRZ - 1.f / TG);
    float AGE1 = EXP(ARG);
    float AGE2 = EXP(MIN(0.f, 10.f * ARG));
    float AGE3 = 0.3f;
    float TAGE = AGE1 + AGE2 + AGE3;
    float DELA = DELA0 * TAGE;
    float DELS = MAX(0.0f, SNEQV - SNEQVO) / SWEMX;
    float SGE = (TAUSS + DELA) * (1.0f - DELS);
    TAUSS = MAX(0.f, SGE);
  }
  FAGE = TAUSS / (TAUSS + 1.f);
}

// noahmplsm.F90:2599-2649
NMP_DEV void SNOWALB_BATS(float COSZ, float FAGE, B2& ALBSND, B2& ALBSNI) {
  const float C1 = 0.2f, C2 = 0.5f;
  float SL = 2.0f;
  float SL1 = 1.f / SL;
  float SL2 = 2.f * SL;
  float CF1 = ((1.f + SL1) / (1.f + SL2 * COSZ) - SL1);
  float FZEN = MAX(CF1, 0.f);
  ALBSNI(1) = 0.95f * (1.f - C1 * FAGE);
  ALBSNI(2) = 0.65f * (1.f - C2 * FAGE);
  ALBSND(1) = ALBSNI(1) + 0.4f * FZEN * (1.f - ALBSNI(1));
  ALBSND(2) = ALBSNI(2) + 0.4f * FZEN * (1.f - ALBSNI(2));
}

// noahmplsm.F90:2652-2700
NMP_DEV void SNOWALB_CLASS(float QSNOW, float DT, float& ALB, float ALBOLD, B2& ALBSND, B2& ALBSNI) {
  ALB = 0.55f + (ALBOLD - 0.55f) * EXP(-0.01f * DT / 3600.f);
  if (QSNOW > 0.f) ALB = ALB + MIN(QSNOW * DT, SWEMX) * (0.84f - ALB) / (SWEMX);
  ALBSNI(1) = ALB; ALBSNI(2) = ALB; ALBSND(1) = ALB; ALBSND(2) = ALB;
}

// NOAHMP_RAD_PARAMETERS (noahmplsm.F90:427-445); only soil colour ISC=4 is reachable (noahmpdrv.F90:527)
__device__ __constant__ float kALBSAT[2][9] = {{0.15f, 0.11f, 0.10f, 0.09f, 0.08f, 0.07f, 0.06f, 0.05f, 0.f},
                                               {0.30f, 0.22f, 0.20f, 0.18f, 0.16f, 0.14f, 0.12f, 0.10f, 0.f}};
__device__ __constant__ float kALBDRY[2][9] = {{0.27f, 0.22f, 0.20f, 0.18f, 0.16f, 0.14f, 0.12f, 0.10f, 0.f},
                                               {0.54f, 0.44f, 0.40f, 0.36f, 0.32f, 0.28f, 0.24f, 0.20f, 0.f}};

// noahmplsm.F90:2703-2765
NMP_DEV void GROUNDALB(int IST, int ISC, float FSNO, float SMC1, const B2& ALBSND, const B2& ALBSNI,
                       float COSZ, float TG, B2& ALBGRD, B2& ALBGRI) {
#pragma unroll
  for (int IB = 1; IB <= 2; ++IB) {
    float INC = MAX(0.11f - 0.40f * SMC1, 0.f);
    float ALBSOD, ALBSOI;
    if (IST == 1) {
      ALBSOD = MIN(kALBSAT[IB - 1][ISC - 1] + INC, kALBDRY[IB - 1][ISC - 1]);
      ALBSOI = ALBSOD;
    } else if (TG > TFRZ) {
      ALBSOD = 0.06f / (POW(MAX(0.01f, COSZ), 1.7f) + 0.15f);
      ALBSOI = 0.06f;
    } else {
      ALBSOD = (IB == 1) ? 0.60f : 0.40f;  // ALBLAK
      ALBSOI = ALBSOD;
    }
    if (IST == 1 && ISC == 9) {
      ALBSOD = ALBSOD + 0.10f;
      ALBSOI = ALBSOI + 0.10f;
    }
    ALBGRD(IB) = ALBSOD * (1.f - FSNO) + ALBSND(IB) * FSNO;
    ALBGRI(IB) = ALBSOI * (1.f - FSNO) + ALBSNI(IB) * FSNO;
  }
}

// Geometry part of TWOSTREAM that does not depend on band / beam type: gap fractions (:2860-2889)
template <class O>
NMP_DEV void TWOSTREAM_GAPS(const Ctx& c, int VEGTYP, float COSZ, float VAI, float FVEG, float& GAP,
                            float& KOPEN, float& BGAP, float& WGAP) {
  const noahmp_tables& T = *c.T;
  const float PAI = 3.14159265f;
  GAP = 0.f; KOPEN = 0.f;
  if (VAI == 0.0f) {
    GAP = 1.0f;
    KOPEN = 1.0f;
  } else {
    const int rad = NMP_OPT(rad);
    if (rad == 1) {
      float rc = tv1(T.rc, VEGTYP);
      float DENFVEG = -LOG(MAX(1.0f - FVEG, 0.01f)) / (PAI * (rc * rc));
      float HD = tv1(T.hvt, VEGTYP) - tv1(T.hvb, VEGTYP);
      float BB = 0.5f * HD;
      float THETAP = ATAN(BB / rc * TAN(ACOS(MAX(0.01f, COSZ))));
      BGAP = EXP(-DENFVEG * PAI * (rc * rc) / COS(THETAP));
      float FA = VAI / (1.33f * PAI * POW(rc, 3.0f) * (BB / rc) * DENFVEG);
      float NEWVAI = HD * FA;
      WGAP = (1.0f - BGAP) * EXP(-0.5f * NEWVAI / COSZ);
      GAP = MIN(1.0f - FVEG, BGAP + WGAP);
      KOPEN = 0.05f;
    }
    if (rad == 2) { GAP = 0.0f; KOPEN = 0.0f; }
    if (rad == 3) { GAP = 1.0f - FVEG; KOPEN = 1.0f - FVEG; }
  }
}

// noahmplsm.F90:2768-3016.  The reference calls this 4x per column (2 bands x {direct,diffuse}); the
// gap-fraction block is evaluated each time there with identical results, here once (TWOSTREAM_GAPS).
NMP_DEV void TWOSTREAM(const Ctx& c, int IB, int IC, int VEGTYP, float COSZ, float VAI, float FWET, float Tv,
                       const B2& ALBGRD, const B2& ALBGRI, const B2& RHO, const B2& TAU, float GAP,
                       float KOPEN, float& FAB, float& FRE, float& FTD, float& FTI, float& GDIR,
                       float& FREV, float& FREG) {
  const noahmp_tables& T = *c.T;
  const float OMEGAS = (IB == 1) ? 0.8f : 0.4f, BETADS = 0.5f, BETAIS = 0.5f;
  float COSZI = MAX(0.001f, COSZ);
  float CHIL = MIN(MAX(tv1(T.xl, VEGTYP), -0.4f), 0.6f);
  if (ABS(CHIL) <= 0.01f) CHIL = 0.01f;
  float PHI1 = 0.5f - 0.633f * CHIL - 0.330f * CHIL * CHIL;
  float PHI2 = 0.877f * (1.f - 2.f * PHI1);
  GDIR = PHI1 + PHI2 * COSZI;
  float EXT = GDIR / COSZI;
  float AVMU = (1.f - PHI1 / PHI2 * LOG((PHI1 + PHI2) / PHI1)) / PHI2;
  float OMEGAL = RHO(IB) + TAU(IB);
  float TMP0 = GDIR + PHI2 * COSZI;
  float TMP1 = PHI1 * COSZI;
  float ASU = 0.5f * OMEGAL * GDIR / TMP0 * (1.f - TMP1 / TMP0 * LOG((TMP1 + TMP0) / TMP1));
  float BETADL = (1.f + AVMU * EXT) / (OMEGAL * AVMU * EXT) * ASU;
  float hc = (1.f + CHIL) / 2.f;
  float BETAIL = 0.5f * (RHO(IB) + TAU(IB) + (RHO(IB) - TAU(IB)) * (hc * hc)) / OMEGAL;
  float TMP2;
  if (Tv > TFRZ) {
    TMP0 = OMEGAL; TMP1 = BETADL; TMP2 = BETAIL;
  } else {
    TMP0 = (1.f - FWET) * OMEGAL + FWET * OMEGAS;
    TMP1 = ((1.f - FWET) * OMEGAL * BETADL + FWET * OMEGAS * BETADS) / TMP0;
    TMP2 = ((1.f - FWET) * OMEGAL * BETAIL + FWET * OMEGAS * BETAIS) / TMP0;
  }
  float OMEGA = TMP0, BETAD = TMP1, BETAI = TMP2;
  float B = 1.f - OMEGA + OMEGA * BETAI;
  float C = OMEGA * BETAI;
  TMP0 = AVMU * EXT;
  float D = TMP0 * OMEGA * BETAD;
  float F = TMP0 * OMEGA * (1.f - BETAD);
  TMP1 = B * B - C * C;
  float H = SQRT(TMP1) / AVMU;
  float SIGMA = TMP0 * TMP0 - TMP1;
  if (ABS(SIGMA) < 1.e-6f) SIGMA = SIGN(1.e-6f, SIGMA);
  float P1 = B + AVMU * H;
  float P2 = B - AVMU * H;
  float P3 = B + TMP0;
  float P4 = B - TMP0;
  float S1 = EXP(-H * VAI);
  float S2 = EXP(-EXT * VAI);
  const float ALBG = (IC == 0) ? ALBGRD(IB) : ALBGRI(IB);
  float U1 = B - C / ALBG;
  float U2 = B - C * ALBG;
  float U3 = F + C * ALBG;
  TMP2 = U1 - AVMU * H;
  float TMP3 = U1 + AVMU * H;
  float D1 = P1 * TMP2 / S1 - P2 * TMP3 * S1;
  float TMP4 = U2 + AVMU * H;
  float TMP5 = U2 - AVMU * H;
  float D2 = TMP4 / S1 - TMP5 * S1;
  float H1 = -D * P4 - C * F;
  float TMP6 = D - H1 * P3 / SIGMA;
  float TMP7 = (D - C - H1 / SIGMA * (U1 + TMP0)) * S2;
  float H2 = (TMP6 * TMP2 / S1 - P2 * TMP7) / D1;
  float H3 = -(TMP6 * TMP3 * S1 - P1 * TMP7) / D1;
  float H4 = -F * P3 - C * D;
  float TMP8 = H4 / SIGMA;
  float TMP9 = (U3 - TMP8 * (U2 - TMP0)) * S2;
  float H5 = -(TMP8 * TMP4 / S1 + TMP9) / D2;
  float H6 = (TMP8 * TMP5 * S1 + TMP9) / D2;
  float H7 = (C * TMP2) / (D1 * S1);
  float H8 = (-C * TMP3 * S1) / D1;
  float H9 = TMP4 / (D2 * S1);
  float H10 = (-TMP5 * S1) / D2;
  float FTDS, FTIS, FRES, FREVEG, FREBAR;
  if (IC == 0) {
    FTDS = S2 * (1.0f - GAP) + GAP;
    FTIS = (H4 * S2 / SIGMA + H5 * S1 + H6 / S1) * (1.0f - GAP);
    FRES = (H1 / SIGMA + H2 + H3) * (1.0f - GAP) + ALBGRD(IB) * GAP;
    FREVEG = (H1 / SIGMA + H2 + H3) * (1.0f - GAP);
    FREBAR = ALBGRD(IB) * GAP;
  } else {
    FTDS = 0.f;
    FTIS = (H9 * S1 + H10 / S1) * (1.0f - KOPEN) + KOPEN;
    FRES = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREVEG = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREBAR = 0.f;
  }
  FTD = FTDS;
  FTI = FTIS;
  FRE = FRES;
  FREV = FREVEG;
  FREG = FREBAR;
  FAB = 1.f - FRE - (1.f - ALBGRD(IB)) * FTD - (1.f - ALBGRI(IB)) * FTI;
}

// RADIATION = ALBEDO + SURRAD (noahmplsm.F90:2120-2544)
struct RadOut {
  float FSUN, LAISUN, LAISHA, PARSUN, PARSHA, SAV, SAG, FSR, FSA, FSRV, FSRG, BGAP, WGAP;
};
template <class O>
NMP_DEV void RADIATION(const Ctx& c, int VEGTYP, int IST, int ISC, float SNEQVO, float SNEQV, float DT,
                       float COSZ, float TG, float TV, float FSNO, float QSNOW, float FWET, float ELAI,
                       float ESAI, float SMC1, const B2& SOLAD, const B2& SOLAI, float FVEG, float& ALBOLD,
                       float& TAUSS, RadOut& r) {
  const noahmp_tables& T = *c.T;
  const float MPE = 1.E-06f;
  B2 ALBGRD, ALBGRI, ALBD, ALBI, FABD, FABI, FTDD, FTID, FTII, FREVI, FREVD, FREGD, FREGI;
#pragma unroll
  for (int IB = 1; IB <= 2; ++IB) {
    ALBD(IB) = 0.f; ALBI(IB) = 0.f; ALBGRD(IB) = 0.f; ALBGRI(IB) = 0.f; FABD(IB) = 0.f; FABI(IB) = 0.f;
    FTDD(IB) = 0.f; FTID(IB) = 0.f; FTII(IB) = 0.f;
    FREVI(IB) = 0.f; FREVD(IB) = 0.f; FREGD(IB) = 0.f; FREGI(IB) = 0.f;  // undefined at night in the reference
  }
  r.BGAP = 0.f; r.WGAP = 0.f; r.FSUN = 0.f;
  // ---- ALBEDO (:2243-2423): everything below is skipped at night ----
  if (COSZ > 0.f) {
    B2 RHO, TAU, ALBSND, ALBSNI;
    float VAI = ELAI + ESAI;
    float WL = ELAI / MAX(VAI, MPE);
    float WS = ESAI / MAX(VAI, MPE);
#pragma unroll
    for (int IB = 1; IB <= 2; ++IB) {
      RHO(IB) = MAX(T.rhol[IB - 1][VEGTYP - 1] * WL + T.rhos[IB - 1][VEGTYP - 1] * WS, MPE);
      TAU(IB) = MAX(T.taul[IB - 1][VEGTYP - 1] * WL + T.taus[IB - 1][VEGTYP - 1] * WS, MPE);
    }
    float FAGE;
    SNOW_AGE(DT, TG, SNEQVO, SNEQV, TAUSS, FAGE);
    ALBSND(1) = 0.f; ALBSND(2) = 0.f; ALBSNI(1) = 0.f; ALBSNI(2) = 0.f;
    const int alb = NMP_OPT(alb);
    if (alb == 1) SNOWALB_BATS(COSZ, FAGE, ALBSND, ALBSNI);
    if (alb == 2) {
      float ALB;
      SNOWALB_CLASS(QSNOW, DT, ALB, ALBOLD, ALBSND, ALBSNI);
      ALBOLD = ALB;
    }
    GROUNDALB(IST, ISC, FSNO, SMC1, ALBSND, ALBSNI, COSZ, TG, ALBGRD, ALBGRI);
    float GAP, KOPEN, GDIR = 0.f, FTDI;
    TWOSTREAM_GAPS<O>(c, VEGTYP, COSZ, VAI, FVEG, GAP, KOPEN, r.BGAP, r.WGAP);
#pragma unroll
    for (int IB = 1; IB <= 2; ++IB) {
      TWOSTREAM(c, IB, 0, VEGTYP, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, GAP, KOPEN, FABD(IB), ALBD(IB),
                FTDD(IB), FTID(IB), GDIR, FREVD(IB), FREGD(IB));
      TWOSTREAM(c, IB, 1, VEGTYP, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, GAP, KOPEN, FABI(IB), ALBI(IB),
                FTDI, FTII(IB), GDIR, FREVI(IB), FREGI(IB));
    }
    float EXT = GDIR / COSZ * SQRT(1.f - RHO(1) - TAU(1));
    float FSUN = (1.f - EXP(-EXT * VAI)) / MAX(EXT * VAI, MPE);
    EXT = FSUN;
    if (EXT < 0.01f) WL = 0.f; else WL = EXT;
    r.FSUN = WL;
  }
  // ---- RADIATION tail (:2221-2229) and SURRAD (:2426-2544) ----
  float FSHA = 1.f - r.FSUN;
  r.LAISUN = ELAI * r.FSUN;
  r.LAISHA = ELAI * FSHA;
  float VAI = ELAI + ESAI;
  B2 CAD, CAI;
  r.SAG = 0.f; r.SAV = 0.f; r.FSA = 0.f;
#pragma unroll
  for (int IB = 1; IB <= 2; ++IB) {
    CAD(IB) = SOLAD(IB) * FABD(IB);
    CAI(IB) = SOLAI(IB) * FABI(IB);
    r.SAV = r.SAV + CAD(IB) + CAI(IB);
    r.FSA = r.FSA + CAD(IB) + CAI(IB);
    float TRD = SOLAD(IB) * FTDD(IB);
    float TRI = SOLAD(IB) * FTID(IB) + SOLAI(IB) * FTII(IB);
    float ABSG = TRD * (1.f - ALBGRD(IB)) + TRI * (1.f - ALBGRI(IB));
    r.SAG = r.SAG + ABSG;
    r.FSA = r.FSA + ABSG;
  }
  float LAIFRA = ELAI / MAX(VAI, MPE);
  if (r.FSUN > 0.f) {
    r.PARSUN = (CAD(1) + r.FSUN * CAI(1)) * LAIFRA / MAX(r.LAISUN, MPE);
    r.PARSHA = (FSHA * CAI(1)) * LAIFRA / MAX(r.LAISHA, MPE);
  } else {
    r.PARSUN = 0.f;
    r.PARSHA = (CAD(1) + CAI(1)) * LAIFRA / MAX(r.LAISHA, MPE);
  }
  float RVIS = ALBD(1) * SOLAD(1) + ALBI(1) * SOLAI(1);
  float RNIR = ALBD(2) * SOLAD(2) + ALBI(2) * SOLAI(2);
  r.FSR = RVIS + RNIR;
  r.FSRV = FREVD(1) * SOLAD(1) + FREVI(1) * SOLAI(1) + FREVD(2) * SOLAD(2) + FREVI(2) * SOLAI(2);
  r.FSRG = FREGD(1) * SOLAD(1) + FREGI(1) * SOLAI(1) + FREGD(2) * SOLAD(2) + FREGI(2) * SOLAI(2);
}

// ---- surface exchange --------------------------------------------------------------------------
struct SfcState {  // variables VEGE_FLUX / BARE_FLUX keep alive across SFCDIF calls
  float MOZ, FM, FH, FM2, FH2, FV, WSTAR;
  int MOZSGN;
};

// The four logarithms of SFCDIF1 (:4127-4130) depend only on the geometry, which is fixed during the canopy / ground
// Newton iterations: the callers evaluate them once per column instead of once per pass (same expressions, same bits).
struct SfcLogs {
  float TMPCM, TMPCH, TMPCM2, TMPCH2;
};
NMP_DEV SfcLogs sfcdif1_logs(float ZLVL, float ZPD, float Z0M, float Z0H) {
  SfcLogs g;
  g.TMPCM = LOG((ZLVL - ZPD) / Z0M);
  g.TMPCH = LOG((ZLVL - ZPD) / Z0H);
  g.TMPCM2 = LOG((2.0f + Z0M) / Z0M);
  g.TMPCH2 = LOG((2.0f + Z0H) / Z0H);
  return g;
}

// noahmplsm.F90:4061-4220 (maths identical in glacier.F90:1202-1358)
NMP_DEV void SFCDIF1(Ctx& c, int ITER, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL, float ZPD,
                     float Z0M, float Z0H, float UR, float MPE, const SfcLogs& G, SfcState& s, float& CM, float& CH) {
  float MOZOLD = s.MOZ;
  if (ZLVL <= ZPD) c.fatal(NOAHMP_ERR_ZLVL, ZLVL - ZPD);
  const float TMPCM = G.TMPCM, TMPCH = G.TMPCH, TMPCM2 = G.TMPCM2, TMPCH2 = G.TMPCH2;
  float MOZ2;
  if (ITER == 1) {
    s.FV = 0.0f; s.MOZ = 0.0f; MOZ2 = 0.0f;
  } else {
    float TVIR = (1.f + 0.61f * QAIR) * SFCTMP;
    float TMP1 = VKC * (GRAV / TVIR) * H / (RHOAIR * CPAIR);
    if (ABS(TMP1) <= MPE) TMP1 = MPE;
    float MOL = -1.f * POW3(s.FV) / TMP1;
    s.MOZ = MIN((ZLVL - ZPD) / MOL, 1.f);
    MOZ2 = MIN((2.0f + Z0H) / MOL, 1.f);
  }
  if (MOZOLD * s.MOZ < 0.f) s.MOZSGN = s.MOZSGN + 1;
  if (s.MOZSGN >= 2) {
    s.MOZ = 0.f; s.FM = 0.f; s.FH = 0.f; MOZ2 = 0.f; s.FM2 = 0.f; s.FH2 = 0.f;
  }
  float FMNEW, FHNEW, FM2NEW, FH2NEW;
  if (s.MOZ < 0.f) {
    float TMP1 = POW(1.f - 16.f * s.MOZ, 0.25f);
    float TMP2 = LOG((1.f + TMP1 * TMP1) / 2.f);
    float TMP3 = LOG((1.f + TMP1) / 2.f);
    FMNEW = 2.f * TMP3 + TMP2 - 2.f * ATAN(TMP1) + 1.5707963f;
    FHNEW = 2.f * TMP2;
    float TMP12 = POW(1.f - 16.f * MOZ2, 0.25f);
    float TMP22 = LOG((1.f + TMP12 * TMP12) / 2.f);
    float TMP32 = LOG((1.f + TMP12) / 2.f);
    FM2NEW = 2.f * TMP32 + TMP22 - 2.f * ATAN(TMP12) + 1.5707963f;
    FH2NEW = 2.f * TMP22;
  } else {
    FMNEW = -5.f * s.MOZ;
    FHNEW = FMNEW;
    FM2NEW = -5.f * MOZ2;
    FH2NEW = FM2NEW;
  }
  if (ITER == 1) {
    s.FM = FMNEW; s.FH = FHNEW; s.FM2 = FM2NEW; s.FH2 = FH2NEW;
  } else {
    s.FM = 0.5f * (s.FM + FMNEW);
    s.FH = 0.5f * (s.FH + FHNEW);
    s.FM2 = 0.5f * (s.FM2 + FM2NEW);
    s.FH2 = 0.5f * (s.FH2 + FH2NEW);
  }
  s.FH = MIN(s.FH, 0.9f * TMPCH);
  s.FM = MIN(s.FM, 0.9f * TMPCM);
  s.FH2 = MIN(s.FH2, 0.9f * TMPCH2);
  s.FM2 = MIN(s.FM2, 0.9f * TMPCM2);
  float CMFM = TMPCM - s.FM;
  float CHFH = TMPCH - s.FH;
  if (ABS(CMFM) <= MPE) CMFM = MPE;
  if (ABS(CHFH) <= MPE) CHFH = MPE;
  CM = VKC * VKC / (CMFM * CMFM);
  CH = VKC * VKC / (CMFM * CHFH);
  s.FV = UR * SQRT(CM);
  // CH2 (2-m exchange coefficient, :4213-4218) is computed by the reference but never used by its callers
}

// noahmplsm.F90:4224-4422
NMP_DEV void SFCDIF2(int ITER, float Z0, float THZ0, float THLM, float SFCSPD, float CZIL, float ZLM,
                     float& AKMS, float& AKHS, float& RLMO, float& WSTAR2, float& USTAR) {
  const float WWST = 1.2f, WWST2 = WWST * WWST, VKRM = 0.40f, EXCM = 0.001f, BETA = 1.0f / 270.0f,
              BTG = BETA * GRAV, ELFC = VKRM * BTG, WOLD = 0.15f, WNEW = 1.0f - WOLD, PIHF = 3.14159265f / 2.f,
              EPSU2 = 1.E-4f, EPSUST = 0.07f, ZTMIN = -5.0f, ZTMAX = 1.0f, HPBL = 1000.0f, SQVISC = 258.2f;
  // Paulson (ILECH = 0) stability functions (:4292-4297)
  auto PSPMU = [&](float XX) {
    return -2.f * LOG((XX + 1.f) * 0.5f) - LOG((XX * XX + 1.f) * 0.5f) + 2.f * ATAN(XX) - PIHF;
  };
  auto PSPHU = [&](float XX) { return -2.f * LOG((XX * XX + 1.f) * 0.5f); };
  float ZILFC = -CZIL * VKRM * SQVISC;
  float ZU = Z0;
  float RDZ = 1.f / ZLM;
  float CXCH = EXCM * RDZ;
  float DTHV = THLM - THZ0;
  float DU2 = MAX(SFCSPD * SFCSPD, EPSU2);
  float BTGH = BTG * HPBL;
  if (ITER == 1) {
    if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * POW(ABS(BTGH * AKHS * DTHV), 2.f / 3.f);
    else WSTAR2 = 0.0f;
    USTAR = MAX(SQRT(AKMS * SQRT(DU2 + WSTAR2)), EPSUST);
    RLMO = ELFC * AKHS * DTHV / POW3(USTAR);
  }
  float ZT = MAX(1.E-6f, EXP(ZILFC * SQRT(USTAR * Z0)) * Z0);
  float ZSLU = ZLM + ZU;
  float ZSLT = ZLM + ZT;
  float RLOGU = LOG(ZSLU / ZU);
  float RLOGT = LOG(ZSLT / ZT);
  float ZETALT = MAX(ZSLT * RLMO, ZTMIN);
  RLMO = ZETALT / ZSLT;
  float ZETALU = ZSLU * RLMO;
  float ZETAU = ZU * RLMO;
  float ZETAT = ZT * RLMO;
  float PSMZ, SIMM, PSHZ, SIMH;
  if (RLMO < 0.f) {
    float XLU4 = 1.f - 16.f * ZETALU;
    float XLT4 = 1.f - 16.f * ZETALT;
    float XU4 = 1.f - 16.f * ZETAU;
    float XT4 = 1.f - 16.f * ZETAT;
    float XLU = SQRT(SQRT(XLU4));
    float XLT = SQRT(SQRT(XLT4));
    float XU = SQRT(SQRT(XU4));
    float XT = SQRT(SQRT(XT4));
    PSMZ = PSPMU(XU);
    SIMM = PSPMU(XLU) - PSMZ + RLOGU;
    PSHZ = PSPHU(XT);
    SIMH = PSPHU(XLT) - PSHZ + RLOGT;
  } else {
    ZETALU = MIN(ZETALU, ZTMAX);
    ZETALT = MIN(ZETALT, ZTMAX);
    PSMZ = 5.f * ZETAU;
    SIMM = 5.f * ZETALU - PSMZ + RLOGU;
    PSHZ = 5.f * ZETAT;
    SIMH = 5.f * ZETALT - PSHZ + RLOGT;
  }
  USTAR = MAX(SQRT(AKMS * SQRT(DU2 + WSTAR2)), EPSUST);
  float USTARK = USTAR * VKRM;
  AKMS = MAX(USTARK / SIMM, CXCH);
  AKHS = MAX(USTARK / SIMH, CXCH);
  if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * POW(ABS(BTGH * AKHS * DTHV), 2.f / 3.f);
  else WSTAR2 = 0.0f;
  float RLMN = ELFC * AKHS * DTHV / POW3(USTAR);
  float RLMA = RLMO * WOLD + RLMN * WNEW;
  RLMO = RLMA;
}

// noahmplsm.F90:3960-4057
NMP_DEV void RAGRB(const Ctx& c, int ITER, float VAI, float RHOAIR, float HG, float TAH, float ZPD,
                   float Z0MG, float Z0HG, float HCAN, float UC, float Z0H, float FV, float CWP, int VEGTYP,
                   float MPE, float& FHG, float& RAHG, float& RAWG, float& RB) {
  float MOZG = 0.f;
  if (ITER > 1) {
    float TMP1 = VKC * (GRAV / TAH) * HG / (RHOAIR * CPAIR);
    if (ABS(TMP1) <= MPE) TMP1 = MPE;
    float MOLG = -1.f * POW3(FV) / TMP1;
    MOZG = MIN((ZPD - Z0MG) / MOLG, 1.f);
  }
  float FHGNEW;
  if (MOZG < 0.f) FHGNEW = POW(1.f - 15.f * MOZG, -0.25f);
  else FHGNEW = 1.f + 4.7f * MOZG;
  if (ITER == 1) FHG = FHGNEW;
  else FHG = 0.5f * (FHG + FHGNEW);
  float CWPC = POW(CWP * VAI * HCAN * FHG, 0.5f);
  float TMP1 = EXP(-CWPC * Z0HG / HCAN);
  float TMP2 = EXP(-CWPC * (Z0H + ZPD) / HCAN);
  float TMPRAH2 = HCAN * EXP(CWPC) / CWPC * (TMP1 - TMP2);
  float KH = MAX(VKC * FV * (HCAN - ZPD), MPE);
  RAHG = TMPRAH2 / KH;
  RAWG = RAHG;
  float TMPRB = CWPC * 50.f / (1.f - EXP(-CWPC / 2.f));
  RB = TMPRB * SQRT(tv1(c.T->dleaf, VEGTYP) / UC);
}

// noahmplsm.F90:5323-5464 (with the internal CI2CI)
NMP_DEV void STOMATA(const Ctx& c, int VEGTYP, float MPE, float APAR, float FOLN, float TV, float EI, float EA,
                     float SFCTMP, float SFCPRS, float O2, float CO2, float IGS, float BTRAN, float RB,
                     float& RS, float& PSN) {
  const noahmp_tables& T = *c.T;
  const float CIERR = 5e-2f;
  const float bp = tv1(T.bp, VEGTYP);
  float CF = SFCPRS / (8.314f * SFCTMP) * 1.0e06f;
  RS = 1.0f / bp * CF;
  PSN = 0.0f;
  if (APAR <= 0.0f) return;
  const float c3 = tv1(T.c3psn, VEGTYP), mp = tv1(T.mp, VEGTYP);
  float FNF = MIN(FOLN / MAX(MPE, tv1(T.folnmx, VEGTYP)), 1.0f);
  float TC = TV - TFRZ;
  float PPF = 4.6f * APAR;
  float J = PPF * tv1(T.qe25, VEGTYP);
  float KC = tv1(T.kc25, VEGTYP) * POW(tv1(T.akc, VEGTYP), (TC - 25.0f) / 10.0f);
  float KO = tv1(T.ko25, VEGTYP) * POW(tv1(T.ako, VEGTYP), (TC - 25.0f) / 10.0f);
  float AWC = KC * (1.0f + O2 / KO);
  float CP = 0.5f * KC / KO * O2 * 0.21f;
  float VCMX = tv1(T.vcmx25, VEGTYP) / (1.0f + EXP((-2.2E05f + 710.0f * (TC + TFRZ)) / (8.314f * (TC + TFRZ)))) *
               FNF * BTRAN * POW(tv1(T.avcmx, VEGTYP), (TC - 25.0f) / 10.0f);
  float RLB = RB / CF;
  float CIHI = 1.5f * CO2;
  float CILOW = 0.0f;
#if NMP_FASTMATH
  // production build: the loop-invariant quotients of CI2CI are formed once (five reciprocals per bisection step
  // instead of seven; the parity build below keeps the reference's expression order)
  const float WE0 = 0.5f * VCMX * c3, WE1 = 4000.0f * VCMX / SFCPRS * (1.f - c3);
  const float J1 = J * (1.f - c3), V1 = VCMX * (1.f - c3), EAEI = EA / EI, MPP = mp * SFCPRS;
  const float K137 = 1.37f * RLB * SFCPRS, K165 = SFCPRS * 1.65f;
#endif
  for (int ITER = 1; ITER <= 20; ++ITER) {
    float CI = 0.5f * (CIHI + CILOW);
    // CI2CI (:5430-5463)
#if NMP_FASTMATH
    const float CICP = MAX(CI - CP, 0.0f);
    float WJ = CICP * J / (CI + 2.0f * CP) * c3 + J1;
    float WC = CICP * VCMX / (CI + AWC) * c3 + V1;
    float WE = WE0 + WE1 * CI;
    PSN = MIN(MIN(WJ, WC), WE) * IGS;
    float CS = MAX(CO2 - K137 * PSN, MPE);
    const float TQ = MPP * PSN / CS;
    float A = TQ * EAEI + bp;
    float B = (TQ + bp) * RLB - 1.f;
    float C = -RLB;
    float Q;
    if (B >= 0.0f) Q = -0.5f * (B + SQRT(B * B - 4.0f * A * C));
    else Q = -0.5f * (B - SQRT(B * B - 4.0f * A * C));
    float R1 = Q / A;
    float R2 = C / Q;
    RS = MAX(R1, R2);
    float FCI = MAX(CS - PSN * K165 * RS, 0.0f);
#else
    float WJ = MAX(CI - CP, 0.0f) * J / (CI + 2.0f * CP) * c3 + J * (1.f - c3);
    float WC = MAX(CI - CP, 0.0f) * VCMX / (CI + AWC) * c3 + VCMX * (1.f - c3);
    float WE = 0.5f * VCMX * c3 + 4000.0f * VCMX * CI / SFCPRS * (1.f - c3);
    PSN = MIN(MIN(WJ, WC), WE) * IGS;
    float CS = MAX(CO2 - 1.37f * RLB * SFCPRS * PSN, MPE);
    float A = mp * PSN * SFCPRS * EA / (CS * EI) + bp;
    float B = (mp * PSN * SFCPRS / CS + bp) * RLB - 1.f;
    float C = -RLB;
    float Q;
    if (B >= 0.0f) Q = -0.5f * (B + SQRT(B * B - 4.0f * A * C));
    else Q = -0.5f * (B - SQRT(B * B - 4.0f * A * C));
    float R1 = Q / A;
    float R2 = C / Q;
    RS = MAX(R1, R2);
    float FCI = MAX(CS - PSN * SFCPRS * 1.65f * RS, 0.0f);
#endif
    if (((CIHI - CILOW) <= CIERR) || ABS(FCI - CI) <= MPE) break;
    else if (FCI > CI) CILOW = CI;
    else CIHI = CI;
  }
  RS = RS * CF;
}

// noahmplsm.F90:5598-5705 (CANRES with CALHUM)
NMP_DEV void CANRES(const Ctx& c, float PAR, float SFCTMP, float RCSOIL, float EAH, float SFCPRS, float& RC,
                    float& PSN) {
  const Prm& P = c.P;
  const float A3 = 273.15f, ELWV = 2.501E6f, E0 = 0.611f, RV = 461.0f, EPSILON = 0.622f;
  float Q2 = 0.622f * EAH / (SFCPRS - 0.378f * EAH);
  Q2 = Q2 / (1.0f + Q2);
  // CALHUM: only Q2SAT is used by CANRES
  float ES = E0 * EXP(ELWV / RV * (1.f / A3 - 1.f / SFCTMP));
  float SFCPRSX = SFCPRS * 1.E-3f;
  float Q2SAT = EPSILON * ES / (SFCPRSX - ES);
  Q2SAT = Q2SAT * 1.E3f;
  Q2SAT = Q2SAT / 1.E3f;
  float FF = 2.0f * PAR / P.RGL;
  float RCS = (FF + P.RSMIN / P.RSMAX) / (1.0f + FF);
  RCS = MAX(RCS, 0.0001f);
  float RCT = 1.0f - 0.0016f * POWR2(P.TOPT - SFCTMP);
  RCT = MAX(RCT, 0.0001f);
  float RCQ = 1.0f / (1.0f + P.HS * MAX(0.f, Q2SAT - Q2));
  RCQ = MAX(RCQ, 0.01f);
  RC = P.RSMIN / (RCS * RCT * RCQ * RCSOIL);
  PSN = -999.99f;
}

// inputs shared by VEGE_FLUX and BARE_FLUX
struct FluxIn {
  int ISNOW, VEGTYP;
  float DT, SAV, SAG, LWDN, UR, UU, VV, SFCTMP, THAIR, QAIR, EAIR, RHOAIR, SNOWH, SFCPRS, PSFC;
  float RSURF, RHSUR, EMG, ZLVL;
  float DF_TOP, DZ_TOP, STC_TOP;  // DF/DZSNSO/STC at layer ISNOW+1
};

struct VegOut {
  float TAUXV, TAUYV, IRG, IRC, SHG, SHC, EVG, EVC, TR, GH, T2MV, PSNSUN, PSNSHA, Q2V, CAH2, CHLEAF, CHUC;
};

// noahmplsm.F90:3018-3589
template <class O>
NMP_DEV void VEGE_FLUX(Ctx& c, const FluxIn& in, float VAI, float GAMMAV, float GAMMAG, float FWET,
                       float LAISUN, float LAISHA, float CWP, float HTOP, float ZPD, float Z0M, float FVEG,
                       float Z0MG, float EMV, float CANLIQ, float CANICE, float& RSSUN, float& RSSHA,
                       float LATHEAV, float PARSUN, float PARSHA, float IGS, float FOLN, float CO2AIR,
                       float O2AIR, float BTRAN, float& EAH, float& TAH, float& TV, float& TG, float& CM,
                       float& CH, float& QSFC, VegOut& o, int& NITER_OUT) {
  const float MPE = 1E-6f;
  const int sfc = NMP_OPT(sfc), crs = NMP_OPT(crs);
  const float SFCTMP = in.SFCTMP, RHOAIR = in.RHOAIR, EAIR = in.EAIR, UR = in.UR, EMG = in.EMG, LWDN = in.LWDN;
  SfcState s;
  s.MOZ = 0.f; s.FM = 0.f; s.FH = 0.f; s.FM2 = 0.f; s.FH2 = 0.f; s.FV = 0.1f; s.WSTAR = 0.f; s.MOZSGN = 0;
  int LITER = 0;
  float DTV = 0.f, HG = 0.f, H = 0.f;
  float FHG = 0.f, RAHG = 0.f, RAWG = 0.f, RB = 0.f;
  float ESTV = 0.f, DESTV = 0.f, ESTG, DESTG = 0.f;
  float CAH = 0.f, CVH = 0.f, CGH, COND, ATA, BTA, CSH, CAW, CEW, CTW, CGW, AEA, BEA, CEV, CTR, A, B;
  float RAHC = 1.f, RAWC;
  o.PSNSUN = 0.f; o.PSNSHA = 0.f;

  float VAIE = MIN(6.f, VAI / FVEG);
  float LAISUNE = MIN(6.f, LAISUN / FVEG);
  float LAISHAE = MIN(6.f, LAISHA / FVEG);

  float T = TDC(TG);
  ESTG = ESAT_SEL1(T);

  QSFC = 0.622f * EAIR / (in.PSFC - 0.378f * EAIR);

  float HCAN = HTOP;
  float UC = UR * LOG(HCAN / Z0M) / LOG(in.ZLVL / Z0M);
  if ((HCAN - ZPD) <= 0.f) c.fatal(NOAHMP_ERR_HCAN, HCAN - ZPD);

  float AIR = -EMV * (1.f + (1.f - EMV) * (1.f - EMG)) * LWDN - EMV * EMG * SB * POW4(TG);
  float CIR = (2.f - EMV * (1.f - EMG)) * EMV * SB;

  const float Z0H = Z0M, Z0HG = Z0MG;
  SfcLogs G{};
  if (sfc == 1) G = sfcdif1_logs(in.ZLVL, ZPD, Z0M, Z0H);
  int ITER;
  for (ITER = 1; ITER <= 20; ++ITER) {
    if (sfc == 1) SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, in.QAIR, in.ZLVL, ZPD, Z0M, Z0H, UR, MPE, G, s, CM, CH);
    if (sfc == 2) {
      SFCDIF2(ITER, Z0M, TAH, in.THAIR, UR, c.P.CZIL, in.ZLVL, CM, CH, s.MOZ, s.WSTAR, s.FV);
      CH = CH / UR;
      CM = CM / UR;
    }
    RAHC = MAX(1.f, 1.f / (CH * UR));
    RAWC = RAHC;

    RAGRB(c, ITER, VAIE, RHOAIR, HG, TAH, ZPD, Z0MG, Z0HG, HCAN, UC, Z0H, s.FV, CWP, in.VEGTYP, MPE, FHG, RAHG,
          RAWG, RB);

    T = TDC(TV);
    ESAT_SEL(T, ESTV, DESTV);

    if (ITER == 1) {
      if (crs == 1) {
        STOMATA(c, in.VEGTYP, MPE, PARSUN, FOLN, TV, ESTV, EAH, SFCTMP, in.SFCPRS, O2AIR, CO2AIR, IGS, BTRAN, RB,
                RSSUN, o.PSNSUN);
        STOMATA(c, in.VEGTYP, MPE, PARSHA, FOLN, TV, ESTV, EAH, SFCTMP, in.SFCPRS, O2AIR, CO2AIR, IGS, BTRAN, RB,
                RSSHA, o.PSNSHA);
      }
      if (crs == 2) {
        CANRES(c, PARSUN, TV, BTRAN, EAH, in.SFCPRS, RSSUN, o.PSNSUN);
        CANRES(c, PARSHA, TV, BTRAN, EAH, in.SFCPRS, RSSHA, o.PSNSHA);
      }
    }

    CAH = 1.f / RAHC;
    CVH = 2.f * VAIE / RB;
    CGH = 1.f / RAHG;
    COND = CAH + CVH + CGH;
    ATA = (SFCTMP * CAH + TG * CGH) / COND;
    BTA = CVH / COND;
    CSH = (1.f - BTA) * RHOAIR * CPAIR * CVH;

    CAW = 1.f / RAWC;
    CEW = FWET * VAIE / RB;
    CTW = (1.f - FWET) * (LAISUNE / (RB + RSSUN) + LAISHAE / (RB + RSSHA));
    CGW = 1.f / (RAWG + in.RSURF);
    COND = CAW + CEW + CTW + CGW;
    AEA = (EAIR * CAW + ESTG * CGW) / COND;
    BEA = (CEW + CTW) / COND;
    CEV = (1.f - BEA) * CEW * RHOAIR * CPAIR / GAMMAV;
    CTR = (1.f - BEA) * CTW * RHOAIR * CPAIR / GAMMAV;

    TAH = ATA + BTA * TV;
    EAH = AEA + BEA * ESTV;

    o.IRC = FVEG * (AIR + CIR * POW4(TV));
    o.SHC = FVEG * RHOAIR * CPAIR * CVH * (TV - TAH);
    o.EVC = FVEG * RHOAIR * CPAIR * CEW * (ESTV - EAH) / GAMMAV;
    o.TR = FVEG * RHOAIR * CPAIR * CTW * (ESTV - EAH) / GAMMAV;
    if (TV > TFRZ) o.EVC = MIN(CANLIQ * LATHEAV / in.DT, o.EVC);
    else o.EVC = MIN(CANICE * LATHEAV / in.DT, o.EVC);

    B = in.SAV - o.IRC - o.SHC - o.EVC - o.TR;
    A = FVEG * (4.f * CIR * POW3(TV) + CSH + (CEV + CTR) * DESTV);
    DTV = B / A;

    o.IRC = o.IRC + FVEG * 4.f * CIR * POW3(TV) * DTV;
    o.SHC = o.SHC + FVEG * CSH * DTV;
    o.EVC = o.EVC + FVEG * CEV * DESTV * DTV;
    o.TR = o.TR + FVEG * CTR * DESTV * DTV;

    TV = TV + DTV;

    H = RHOAIR * CPAIR * (TAH - SFCTMP) / RAHC;
    HG = RHOAIR * CPAIR * (TG - TAH) / RAHG;

    QSFC = (0.622f * EAH) / (in.SFCPRS - 0.378f * EAH);

    NITER_OUT = ITER;
    if (LITER == 1) break;
    if (ITER >= 5 && ABS(DTV) <= 0.01f && LITER == 0) LITER = 1;
  }

  AIR = -EMG * (1.f - EMV) * LWDN - EMG * EMV * SB * POW4(TV);
  CIR = EMG * SB;
  CSH = RHOAIR * CPAIR / RAHG;
  CEV = RHOAIR * CPAIR / (GAMMAG * (RAWG + in.RSURF));
  CGH = 2.f * in.DF_TOP / in.DZ_TOP;

#pragma unroll 1
  for (ITER = 1; ITER <= 5; ++ITER) {
    T = TDC(TG);
    ESAT_SEL(T, ESTG, DESTG);

    o.IRG = CIR * POW4(TG) + AIR;
    o.SHG = CSH * (TG - TAH);
    o.EVG = CEV * (ESTG * in.RHSUR - EAH);
    o.GH = CGH * (TG - in.STC_TOP);

    B = in.SAG - o.IRG - o.SHG - o.EVG - o.GH;
    A = 4.f * CIR * POW3(TG) + CSH + CEV * DESTG + CGH;
    float DTG = B / A;

    o.IRG = o.IRG + 4.f * CIR * POW3(TG) * DTG;
    o.SHG = o.SHG + CSH * DTG;
    o.EVG = o.EVG + CEV * DESTG * DTG;
    o.GH = o.GH + CGH * DTG;
    TG = TG + DTG;
  }

  if (NMP_OPT(stc) == 1) {
    if (in.SNOWH > 0.05f && TG > TFRZ) {
      TG = TFRZ;
      o.IRG = CIR * POW4(TG) - EMG * (1.f - EMV) * LWDN - EMG * EMV * SB * POW4(TV);
      o.SHG = CSH * (TG - TAH);
      o.EVG = CEV * (ESTG * in.RHSUR - EAH);
      o.GH = in.SAG - (o.IRG + o.SHG + o.EVG);
    }
  }

  o.TAUXV = -RHOAIR * CM * UR * in.UU;
  o.TAUYV = -RHOAIR * CM * UR * in.VV;

  if (sfc == 1 || sfc == 2) {
    // FH2 is assigned only by SFCDIF1; with OPT_SFC=2 the reference reads it undefined: defined as 0 here.
    float CAH2 = s.FV * VKC / (LOG((2.f + Z0H) / Z0H) - s.FH2);
    float CQ2V = CAH2;
    if (CAH2 < 1.E-5f) {
      o.T2MV = TAH;
      o.Q2V = QSFC;
    } else {
      o.T2MV = TAH - (o.SHG + o.SHC / FVEG) / (RHOAIR * CPAIR) * 1.f / CAH2;
      o.Q2V = QSFC - ((o.EVC + o.TR) / FVEG + o.EVG) / (LATHEAV * RHOAIR) * 1.f / CQ2V;
    }
    o.CAH2 = CAH2;
  }
  CH = CAH;
  o.CHLEAF = CVH;
  o.CHUC = 1.f / RAHG;
}

struct BareOut {
  float TAUXB, TAUYB, IRB, SHB, EVB, GHB, T2MB, Q2B, EHB2;
};

// noahmplsm.F90:3591-3958
template <class O>
NMP_DEV void BARE_FLUX(Ctx& c, const FluxIn& in, float ZPD, float Z0M, float LATHEA, float GAMMA, float& TGB,
                       float& CM, float& CH, float& QSFC, bool URBAN, BareOut& o) {
  const float MPE = 1E-6f;
  const int sfc = NMP_OPT(sfc);
  const float SFCTMP = in.SFCTMP, RHOAIR = in.RHOAIR, EAIR = in.EAIR, UR = in.UR, EMG = in.EMG, LWDN = in.LWDN;
  SfcState s;
  s.MOZ = 0.f; s.FM = 0.f; s.FH = 0.f; s.FM2 = 0.f; s.FH2 = 0.f; s.FV = 0.1f; s.WSTAR = 0.f; s.MOZSGN = 0;
  float H = 0.f;
  float ESTG = 0.f, DESTG;
  float CSH = 0.f, CEV = 0.f, EHB = 0.f;
  const float Z0H = Z0M;
  float CIR = EMG * SB;
  float CGH = 2.f * in.DF_TOP / in.DZ_TOP;

  SfcLogs G{};
  if (sfc == 1) G = sfcdif1_logs(in.ZLVL, ZPD, Z0M, Z0H);
#pragma unroll 1
  for (int ITER = 1; ITER <= 5; ++ITER) {
    if (sfc == 1) SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, in.QAIR, in.ZLVL, ZPD, Z0M, Z0H, UR, MPE, G, s, CM, CH);
    if (sfc == 2) {
      SFCDIF2(ITER, Z0M, TGB, in.THAIR, UR, c.P.CZIL, in.ZLVL, CM, CH, s.MOZ, s.WSTAR, s.FV);
      CH = CH / UR;
      CM = CM / UR;
      if (in.SNOWH > 0.f) {
        CM = MIN(0.01f, CM);
        CH = MIN(0.01f, CH);
      }
    }
    float RAHB = MAX(1.f, 1.f / (CH * UR));
    float RAWB = RAHB;
    EHB = 1.f / RAHB;

    float T = TDC(TGB);
    ESAT_SEL(T, ESTG, DESTG);

    CSH = RHOAIR * CPAIR / RAHB;
    CEV = RHOAIR * CPAIR / GAMMA / (in.RSURF + RAWB);

    o.IRB = CIR * POW4(TGB) - EMG * LWDN;
    o.SHB = CSH * (TGB - SFCTMP);
    o.EVB = CEV * (ESTG * in.RHSUR - EAIR);
    o.GHB = CGH * (TGB - in.STC_TOP);

    float B = in.SAG - o.IRB - o.SHB - o.EVB - o.GHB;
    float A = 4.f * CIR * POW3(TGB) + CSH + CEV * DESTG + CGH;
    float DTG = B / A;

    o.IRB = o.IRB + 4.f * CIR * POW3(TGB) * DTG;
    o.SHB = o.SHB + CSH * DTG;
    o.EVB = o.EVB + CEV * DESTG * DTG;
    o.GHB = o.GHB + CGH * DTG;

    TGB = TGB + DTG;

    H = CSH * (TGB - SFCTMP);

    T = TDC(TGB);
    ESTG = ESAT_SEL1(T);
    QSFC = 0.622f * (ESTG * in.RHSUR) / (in.PSFC - 0.378f * (ESTG * in.RHSUR));
  }

  if (NMP_OPT(stc) == 1) {
    if (in.SNOWH > 0.05f && TGB > TFRZ) {
      TGB = TFRZ;
      o.IRB = CIR * POW4(TGB) - EMG * LWDN;
      o.SHB = CSH * (TGB - SFCTMP);
      o.EVB = CEV * (ESTG * in.RHSUR - EAIR);
      o.GHB = in.SAG - (o.IRB + o.SHB + o.EVB);
    }
  }

  o.TAUXB = -RHOAIR * CM * UR * in.UU;
  o.TAUYB = -RHOAIR * CM * UR * in.VV;

  if (sfc == 1 || sfc == 2) {
    float EHB2 = s.FV * VKC / (LOG((2.f + Z0H) / Z0H) - s.FH2);
    float CQ2B = EHB2;
    if (EHB2 < 1.E-5f) {
      o.T2MB = TGB;
      o.Q2B = QSFC;
    } else {
      o.T2MB = TGB - o.SHB / (RHOAIR * CPAIR) * 1.f / EHB2;
      o.Q2B = QSFC - o.EVB / (LATHEA * RHOAIR) * (1.f / CQ2B + in.RSURF);
    }
    if (URBAN) o.Q2B = QSFC;
    o.EHB2 = EHB2;
  }
  CH = EHB;
}

// ---- tridiagonal solve: ROSR12 (noahmplsm.F90:5979-6036) on layers NTOP..NSOIL of (-2:4) arrays.
// P enters as workspace and returns the solution; C(NSOIL) is taken as 0.
NMP_DEV void ROSR12(L7& P, const L7& A, const L7& B, const L7& C, const L7& D, L7& DELTA, int NTOP) {
  const TopLayer top(NTOP - 1);
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) {
    if (top.is(K)) {
      P(K) = -C(K) / B(K);
      DELTA(K) = D(K) / B(K);
    } else if (K > NTOP) {
      const float CK = (K == NSOIL) ? 0.0f : C(K);
      P(K) = -CK * (1.0f / (B(K) + A(K) * P(K - (K > -2 ? 1 : 0))));
      DELTA(K) = (D(K) - A(K) * DELTA(K - (K > -2 ? 1 : 0))) * (1.0f / (B(K) + A(K) * P(K - (K > -2 ? 1 : 0))));
    }
  }
  P(NSOIL) = DELTA(NSOIL);
#pragma unroll
  for (int KK = NSOIL - 1; KK >= -2; --KK) {
    if (KK >= NTOP) P(KK) = P(KK) * P(KK + 1) + DELTA(KK);
  }
}

// TSNOSOI = HRT + HSTEP (noahmplsm.F90:5707-5977); glacier.F90:1360-1573 is the same code with ZBOT=-8
template <class O>
NMP_DEV void TSNOSOI(const Ctx& c, int ISNOW, float TBOT, const L7& ZSNSO, float SSOIL, const L7& DF,
                     const L7& HCPCT, float ZBOT, float DT, float SNOWH, L7& STC) {
  const int tbot = NMP_OPT(tbot), stc = NMP_OPT(stc);
  L7 AI, BI, CI, RHSTS, DDZ, DTSDZ;
  float ZBOTSNO = ZBOT - SNOWH;
  const int NTOP = ISNOW + 1;
  const TopLayer top(ISNOW);
  // HRT (:5825-5922); PHI == 0
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) {
    AI(K) = 0.f; BI(K) = 0.f; CI(K) = 0.f; RHSTS(K) = 0.f; DDZ(K) = 0.f; DTSDZ(K) = 0.f;
  }
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) {
    if (K >= NTOP) {
      float DENOM, EFLUX;
      if (top.is(K)) {
        DENOM = -ZSNSO(K) * HCPCT(K);
        float TEMP1 = -ZSNSO(K + (K < NSOIL ? 1 : 0));
        DDZ(K) = 2.0f / TEMP1;
        DTSDZ(K) = 2.0f * (STC(K) - STC(K + (K < NSOIL ? 1 : 0))) / TEMP1;
        EFLUX = DF(K) * DTSDZ(K) - SSOIL - 0.f;
      } else if (K < NSOIL) {
        DENOM = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
        float TEMP1 = ZSNSO(K - 1) - ZSNSO(K + 1);
        DDZ(K) = 2.0f / TEMP1;
        DTSDZ(K) = 2.0f * (STC(K) - STC(K + 1)) / TEMP1;
        EFLUX = (DF(K) * DTSDZ(K) - DF(K - 1) * DTSDZ(K - 1)) - 0.f;
      } else {
        DENOM = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
        float BOTFLX = 0.f;
        if (tbot == 2) {
          DTSDZ(K) = (STC(K) - TBOT) / (0.5f * (ZSNSO(K - 1) + ZSNSO(K)) - ZBOTSNO);
          BOTFLX = -DF(K) * DTSDZ(K);
        }
        EFLUX = (-BOTFLX - DF(K - 1) * DTSDZ(K - 1)) - 0.f;
      }
      if (top.is(K)) {
        AI(K) = 0.0f;
        CI(K) = -DF(K) * DDZ(K) / DENOM;
        if (stc == 1) BI(K) = -CI(K);
        if (stc == 2) BI(K) = -CI(K) + DF(K) / (0.5f * ZSNSO(K) * ZSNSO(K) * HCPCT(K));
      } else if (K < NSOIL) {
        AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM;
        CI(K) = -DF(K) * DDZ(K) / DENOM;
        BI(K) = -(AI(K) + CI(K));
      } else {
        AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM;
        CI(K) = 0.0f;
        BI(K) = -(AI(K) + CI(K));
      }
      RHSTS(K) = EFLUX / (-DENOM);
    }
  }
  // HSTEP (:5925-5977)
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) {
    if (K >= NTOP) {
      RHSTS(K) = RHSTS(K) * DT;
      AI(K) = AI(K) * DT;
      BI(K) = 1.f + BI(K) * DT;
      CI(K) = CI(K) * DT;
    }
  }
  L7 SOL, DELTA;
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) { SOL(K) = 0.f; DELTA(K) = 0.f; }
  ROSR12(SOL, AI, BI, CI, RHSTS, DELTA, NTOP);
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K)
    if (K >= NTOP) STC(K) = STC(K) + SOL(K);
}

// noahmplsm.F90:6247-6377
NMP_DEV float FRH2O(const Prm& P, float TKELV, float SMC, float SH2O) {
  const float CK = 8.0f, BLIM = 5.5f, ERROR_ = 0.005f;
  float BX = P.BEXP;
  if (P.BEXP > BLIM) BX = BLIM;
  int NLOG = 0, KCOUNT = 0;
  float FREE;
  if (TKELV > (TFRZ - 1.E-3f)) {
    FREE = SMC;
  } else {
    float SWL = SMC - SH2O;
    if (SWL > (SMC - 0.02f)) SWL = SMC - 0.02f;
    if (SWL < 0.f) SWL = 0.f;
    while ((NLOG < 10) && (KCOUNT == 0)) {
      NLOG = NLOG + 1;
      float t1 = (1.f + CK * SWL);
      float DF = LOG((P.PSISAT * GRAV / HFUS) * POWR2(t1) * POW(P.SMCMAX / (SMC - SWL), BX)) -
                 LOG(-(TKELV - TFRZ) / TKELV);
      float DENOM = 2.f * CK / (1.f + CK * SWL) + BX / (SMC - SWL);
      float SWLK = SWL - DF / DENOM;
      if (SWLK > (SMC - 0.02f)) SWLK = SMC - 0.02f;
      if (SWLK < 0.f) SWLK = 0.f;
      float DSWL = ABS(SWLK - SWL);
      SWL = SWLK;
      if (DSWL <= ERROR_) KCOUNT = KCOUNT + 1;
    }
    FREE = SMC - SWL;
    if (KCOUNT == 0) {
      float FK = POW((HFUS / (GRAV * (-P.PSISAT))) * ((TKELV - TFRZ) / TKELV), -1.f / BX) * P.SMCMAX;
      if (FK < 0.02f) FK = 0.02f;
      FREE = MIN(FK, SMC);
    }
  }
  return FREE;
}

// noahmplsm.F90:6039-6245
template <class O>
NMP_DEV void PHASECHANGE(const Ctx& c, int ISNOW, float DT, const L7& FACT, const L7& DZSNSO, int IST, L7& STC,
                         N3& SNICE, N3& SNLIQ, float& SNEQV, float& SNOWH, S4& SMC, S4& SH2O, float& QMELT,
                         I7& IMELT, float& PONDING) {
  const Prm& P = c.P;
  const int frz = NMP_OPT(frz);
  L7 HM, XM, WMASS0, WICE0, MICE, MLIQ, SUPERCOOL;
  QMELT = 0.f;
  PONDING = 0.f;
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    SUPERCOOL(J) = 0.0f; IMELT(J) = 0; HM(J) = 0.f; XM(J) = 0.f; MICE(J) = 0.f; MLIQ(J) = 0.f;
    WICE0(J) = 0.f; WMASS0(J) = 0.f;
  }
#pragma unroll
  for (int J = -2; J <= 0; ++J)
    if (J > ISNOW) { MICE(J) = SNICE(J); MLIQ(J) = SNLIQ(J); }
#pragma unroll
  for (int J = 1; J <= NSOIL; ++J) {
    MLIQ(J) = SH2O(J) * DZSNSO(J) * 1000.f;
    MICE(J) = (SMC(J) - SH2O(J)) * DZSNSO(J) * 1000.f;
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J)
    if (J > ISNOW) { WICE0(J) = MICE(J); WMASS0(J) = MICE(J) + MLIQ(J); }
  if (IST == 1) {
#pragma unroll
    for (int J = 1; J <= NSOIL; ++J) {
      if (frz == 1) {
        if (STC(J) < TFRZ) {
          float SMP = HFUS * (TFRZ - STC(J)) / (GRAV * STC(J));
          SUPERCOOL(J) = P.SMCMAX * POW(SMP / P.PSISAT, -1.f / P.BEXP);
          SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.f;
        }
      }
      if (frz == 2) {
        SUPERCOOL(J) = FRH2O(P, STC(J), SMC(J), SH2O(J));
        SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.f;
      }
    }
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    if (J > ISNOW) {
      if (MICE(J) > 0.f && STC(J) >= TFRZ) IMELT(J) = 1;
      if (MLIQ(J) > SUPERCOOL(J) && STC(J) < TFRZ) IMELT(J) = 2;
      if (ISNOW == 0 && SNEQV > 0.f && J == 1) {
        if (STC(J) >= TFRZ) IMELT(J) = 1;
      }
    }
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    if (J > ISNOW) {
      if (IMELT(J) > 0) {
        HM(J) = (STC(J) - TFRZ) / FACT(J);
        STC(J) = TFRZ;
      }
      if (IMELT(J) == 1 && HM(J) < 0.f) { HM(J) = 0.f; IMELT(J) = 0; }
      if (IMELT(J) == 2 && HM(J) > 0.f) { HM(J) = 0.f; IMELT(J) = 0; }
      XM(J) = HM(J) * DT / HFUS;
    }
  }
  if (ISNOW == 0 && SNEQV > 0.f && XM(1) > 0.f) {
    float TEMP1 = SNEQV;
    SNEQV = MAX(0.f, TEMP1 - XM(1));
    float PROPOR = SNEQV / TEMP1;
    SNOWH = MAX(0.f, PROPOR * SNOWH);
    float HEATR = HM(1) - HFUS * (TEMP1 - SNEQV) / DT;
    if (HEATR > 0.f) {
      XM(1) = HEATR * DT / HFUS;
      HM(1) = HEATR;
    } else {
      XM(1) = 0.f;
      HM(1) = 0.f;
    }
    QMELT = MAX(0.f, (TEMP1 - SNEQV)) / DT;
    PONDING = TEMP1 - SNEQV;
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    if (J > ISNOW) {
      if (IMELT(J) > 0 && ABS(HM(J)) > 0.f) {
        float HEATR = 0.f;
        if (XM(J) > 0.f) {
          MICE(J) = MAX(0.f, WICE0(J) - XM(J));
          HEATR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
        } else if (XM(J) < 0.f) {
          if (J <= 0) {
            MICE(J) = MIN(WMASS0(J), WICE0(J) - XM(J));
          } else {
            if (WMASS0(J) < SUPERCOOL(J)) {
              MICE(J) = 0.f;
            } else {
              MICE(J) = MIN(WMASS0(J) - SUPERCOOL(J), WICE0(J) - XM(J));
              MICE(J) = MAX(MICE(J), 0.0f);
            }
          }
          HEATR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
        }
        MLIQ(J) = MAX(0.f, WMASS0(J) - MICE(J));
        if (ABS(HEATR) > 0.f) {
          STC(J) = STC(J) + FACT(J) * HEATR;
          if (J <= 0) {
            if (MLIQ(J) * MICE(J) > 0.f) STC(J) = TFRZ;
          }
        }
        if (J < 1) QMELT = QMELT + MAX(0.f, (WICE0(J) - MICE(J))) / DT;
      }
    }
  }
#pragma unroll
  for (int J = -2; J <= 0; ++J)
    if (J > ISNOW) { SNLIQ(J) = MLIQ(J); SNICE(J) = MICE(J); }
#pragma unroll
  for (int J = 1; J <= NSOIL; ++J) {
    SH2O(J) = MLIQ(J) / (1000.f * DZSNSO(J));
    SMC(J) = (MLIQ(J) + MICE(J)) / (1000.f * DZSNSO(J));
  }
}

}  // namespace nmp
