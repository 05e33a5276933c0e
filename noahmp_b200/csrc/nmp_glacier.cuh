// nmp_glacier.cuh — device code of NOAHMP_GLACIER (phys/module_sf_noahmp_glacier.F90:150-2972).
//
// Routines whose bodies are identical to the land versions (CSNOW, SNOW_AGE, SNOWALB_*, SFCDIF1, ESAT,
// HRT/HSTEP/ROSR12, COMBO, COMPACT — SURVEY.md §8a diff table) reuse the device functions of
// nmp_energy.cuh / nmp_water.cuh; the snow-pack routines are the <GLACIER=true> instantiations.
#pragma once
#include "nmp_sflx.cuh"

namespace nmp {

// glacier.F90:575-645
NMP_DEV void THERMOPROP_GLACIER(int ISNOW, const L7& DZSNSO, float DT, float SNOWH, const N3& SNICE,
                                const N3& SNLIQ, L7& DF, L7& HCPCT, L7& FACT) {
  N3 CVSNO, TKSNO, SNICEV, SNLIQV, EPORE;
  CSNOW(ISNOW, SNICE, SNLIQ, DZSNSO, TKSNO, CVSNO, SNICEV, SNLIQV, EPORE);
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ)
    if (IZ > ISNOW) { DF(IZ) = TKSNO(IZ); HCPCT(IZ) = CVSNO(IZ); }
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    float ZMID = 0.5f * (DZSNSO(IZ));
#pragma unroll
    for (int IZ2 = 1; IZ2 <= IZ - 1; ++IZ2) ZMID = ZMID + DZSNSO(IZ2);
    HCPCT(IZ) = 1.E6f * (0.8194f + 0.1309f * ZMID);
    DF(IZ) = 0.32333f + (0.10073f * ZMID);
  }
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ)
    if (IZ > ISNOW) FACT(IZ) = DT / (HCPCT(IZ) * DZSNSO(IZ));
  if (ISNOW == 0) DF(1) = (DF(1) * DZSNSO(1) + 0.35f * SNOWH) / (SNOWH + DZSNSO(1));
  else DF(1) = (DF(1) * DZSNSO(1) + DF(0) * DZSNSO(0)) / (DZSNSO(0) + DZSNSO(1));
}

// glacier.F90:704-792: no canopy; snow ages and ALBOLD is updated every step, day or night
template <class O>
NMP_DEV void RADIATION_GLACIER(const Ctx& c, float DT, float TG, float SNEQVO, float SNEQV, float COSZ, float QSNOW,
                               const B2& SOLAD, const B2& SOLAI, float& ALBOLD, float& TAUSS, float& SAG,
                               float& FSR, float& FSA) {
  B2 ALBSND, ALBSNI;
  ALBSND(1) = 0.f; ALBSND(2) = 0.f; ALBSNI(1) = 0.f; ALBSNI(2) = 0.f;
  float FAGE;
  SNOW_AGE(DT, TG, SNEQVO, SNEQV, TAUSS, FAGE);
  const int alb = NMP_OPT(alb);
  if (alb == 1) SNOWALB_BATS(COSZ, FAGE, ALBSND, ALBSNI);
  if (alb == 2) {
    float ALB;
    SNOWALB_CLASS(QSNOW, DT, ALB, ALBOLD, ALBSND, ALBSNI);
    ALBOLD = ALB;
  }
  SAG = 0.f; FSA = 0.f; FSR = 0.f;
  float FSNO = 0.0f;
  if (SNEQV > 0.0f) FSNO = 1.0f;
#pragma unroll
  for (int IB = 1; IB <= 2; ++IB) {
    const float ALBICE = (IB == 1) ? 0.80f : 0.55f;
    ALBSND(IB) = ALBICE * (1.f - FSNO) + ALBSND(IB) * FSNO;
    ALBSNI(IB) = ALBICE * (1.f - FSNO) + ALBSNI(IB) * FSNO;
    float ABS_ = SOLAD(IB) * (1.f - ALBSND(IB)) + SOLAI(IB) * (1.f - ALBSNI(IB));
    SAG = SAG + ABS_;
    FSA = FSA + ABS_;
    float REF = SOLAD(IB) * ALBSND(IB) + SOLAI(IB) * ALBSNI(IB);
    FSR = FSR + REF;
  }
}

// glacier.F90:942-1148
template <class O>
NMP_DEV void GLACIER_FLUX(Ctx& c, float EMG, float DF_TOP, float DZ_TOP, float STC_TOP, float Z0M, float ZLVL,
                          float ZPD, float QAIR, float SFCTMP, float RHOAIR, float SFCPRS, float UR, float GAMMA,
                          float RSURF, float LWDN, float RHSUR, const S4& SMC, float EAIR, float SAG, float SNOWH,
                          float LATHEA, const S4& SH2O, float& CM, float& CH, float& TGB, float& QSFC, float& IRB,
                          float& SHB, float& EVB, float& GHB, float& T2MB, float& Q2B, float& EHB2) {
  const float MPE = 1E-6f;
  SfcState s;
  s.MOZ = 0.f; s.FM = 0.f; s.FH = 0.f; s.FM2 = 0.f; s.FH2 = 0.f; s.FV = 0.1f; s.WSTAR = 0.f; s.MOZSGN = 0;
  float H = 0.f;
  float ESTG = 0.f, DESTG, CSH = 0.f, CEV = 0.f, RAHB = 1.f;
  const float Z0H = Z0M;
  float CIR = EMG * SB;
  float CGH = 2.f * DF_TOP / DZ_TOP;
  const SfcLogs G = sfcdif1_logs(ZLVL, ZPD, Z0M, Z0H);
#pragma unroll 1
  for (int ITER = 1; ITER <= 5; ++ITER) {
    SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE, G, s, CM, CH);
    RAHB = MAX(1.f, 1.f / (CH * UR));
    float RAWB = RAHB;
    float T = TDC(TGB);
    ESAT_SEL(T, ESTG, DESTG);
    CSH = RHOAIR * CPAIR / RAHB;
    CEV = RHOAIR * CPAIR / GAMMA / (RSURF + RAWB);
    IRB = CIR * POW4(TGB) - EMG * LWDN;
    SHB = CSH * (TGB - SFCTMP);
    EVB = CEV * (ESTG * RHSUR - EAIR);
    GHB = CGH * (TGB - STC_TOP);
    float B = SAG - IRB - SHB - EVB - GHB;
    float A = 4.f * CIR * POW3(TGB) + CSH + CEV * DESTG + CGH;
    float DTG = B / A;
    IRB = IRB + 4.f * CIR * POW3(TGB) * DTG;
    SHB = SHB + CSH * DTG;
    EVB = EVB + CEV * DESTG * DTG;
    GHB = GHB + CGH * DTG;
    TGB = TGB + DTG;
    H = CSH * (TGB - SFCTMP);
    T = TDC(TGB);
    ESTG = ESAT_SEL1(T);
    QSFC = 0.622f * (ESTG * RHSUR) / (SFCPRS - 0.378f * (ESTG * RHSUR));
  }
  float SICEMAXV = SMC(1) - SH2O(1);
#pragma unroll
  for (int K = 2; K <= NSOIL; ++K) SICEMAXV = MAX(SICEMAXV, SMC(K) - SH2O(K));
  if (NMP_OPT(stc) == 1) {
    if ((SICEMAXV > 0.0f || SNOWH > 0.0f) && TGB > TFRZ) {
      TGB = TFRZ;
      IRB = CIR * POW4(TGB) - EMG * LWDN;
      SHB = CSH * (TGB - SFCTMP);
      EVB = CEV * (ESTG * RHSUR - EAIR);
      GHB = SAG - (IRB + SHB + EVB);
    }
  }
  EHB2 = s.FV * VKC / (LOG((2.f + Z0H) / Z0H) - s.FH2);
  float CQ2B = EHB2;
  if (EHB2 < 1.E-5f) {
    T2MB = TGB;
    Q2B = QSFC;
  } else {
    T2MB = TGB - SHB / (RHOAIR * CPAIR) * 1.f / EHB2;
    Q2B = QSFC - EVB / (LATHEA * RHOAIR) * (1.f / CQ2B + RSURF);
  }
  CH = 1.f / RAHB;
}

// glacier.F90:1635-1922
NMP_DEV void PHASECHANGE_GLACIER(int ISNOW, float DT, const L7& FACT, const L7& DZSNSO, L7& STC, N3& SNICE,
                                 N3& SNLIQ, float& SNEQV, float& SNOWH, S4& SMC, S4& SH2O, float& QMELT, I7& IMELT,
                                 float& PONDING) {
  L7 HM, XM, WMASS0, WICE0, MICE, MLIQ, HEATR;
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    HM(J) = 0.f; XM(J) = 0.f; WMASS0(J) = 0.f; WICE0(J) = 0.f; MICE(J) = 0.f; MLIQ(J) = 0.f; HEATR(J) = 0.f;
    IMELT(J) = 0;
  }
  QMELT = 0.f; PONDING = 0.f;
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
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    if (J > ISNOW) {
      if (MICE(J) > 0.f && STC(J) >= TFRZ) IMELT(J) = 1;
      if (MLIQ(J) > 0.f && STC(J) < TFRZ) IMELT(J) = 2;
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
    float HR = HM(1) - HFUS * (TEMP1 - SNEQV) / DT;
    if (HR > 0.f) {
      XM(1) = HR * DT / HFUS;
      HM(1) = HR;
      IMELT(1) = 1;
    } else {
      XM(1) = 0.f;
      HM(1) = 0.f;
      IMELT(1) = 0;
    }
    QMELT = MAX(0.f, (TEMP1 - SNEQV)) / DT;
    PONDING = TEMP1 - SNEQV;
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) {
    if (J > ISNOW) {
      if (IMELT(J) > 0 && ABS(HM(J)) > 0.f) {
        float HR = 0.f;
        if (XM(J) > 0.f) {
          MICE(J) = MAX(0.f, WICE0(J) - XM(J));
          HR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
        } else if (XM(J) < 0.f) {
          MICE(J) = MIN(WMASS0(J), WICE0(J) - XM(J));
          HR = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
        }
        MLIQ(J) = MAX(0.f, WMASS0(J) - MICE(J));
        if (ABS(HR) > 0.f) {
          STC(J) = STC(J) + FACT(J) * HR;
          if (J <= 0) {
            if (MLIQ(J) * MICE(J) > 0.f) STC(J) = TFRZ;
          }
        }
        if (J < 1) QMELT = QMELT + MAX(0.f, (WICE0(J) - MICE(J))) / DT;
      }
    }
  }
#pragma unroll
  for (int J = -2; J <= NSOIL; ++J) { HEATR(J) = 0.f; XM(J) = 0.f; }

  // The four redistribution sweeps over STC(1:4) (:1804-1908; NSOIL=4 is hard-coded in the reference)
  auto any_stc_gt = [&]() { return STC(1) > TFRZ || STC(2) > TFRZ || STC(3) > TFRZ || STC(4) > TFRZ; };
  auto any_stc_lt = [&]() { return STC(1) < TFRZ || STC(2) < TFRZ || STC(3) < TFRZ || STC(4) < TFRZ; };
  // (1) warm layers give heat to cold layers
  if (any_stc_gt() && any_stc_lt()) {
#pragma unroll
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) > TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
#pragma unroll
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && STC(K) < TFRZ && HEATR(J) > 0.1f) {
            HEATR(K) = (STC(K) - TFRZ) / FACT(K);
            if (ABS(HEATR(K)) > HEATR(J)) {
              HEATR(K) = HEATR(K) + HEATR(J);
              STC(K) = TFRZ + HEATR(K) * FACT(K);
              HEATR(J) = 0.0f;
            } else {
              HEATR(J) = HEATR(J) + HEATR(K);
              HEATR(K) = 0.0f;
              STC(K) = TFRZ;
            }
          }
        }
        STC(J) = TFRZ + HEATR(J) * FACT(J);
      }
    }
  }
  // (2) cold layers take heat from warm layers
  if (any_stc_gt() && any_stc_lt()) {
#pragma unroll
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) < TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
#pragma unroll
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && STC(K) > TFRZ && HEATR(J) < -0.1f) {
            HEATR(K) = (STC(K) - TFRZ) / FACT(K);
            if (HEATR(K) > ABS(HEATR(J))) {
              HEATR(K) = HEATR(K) + HEATR(J);
              STC(K) = TFRZ + HEATR(K) * FACT(K);
              HEATR(J) = 0.0f;
            } else {
              HEATR(J) = HEATR(J) + HEATR(K);
              HEATR(K) = 0.0f;
              STC(K) = TFRZ;
            }
          }
        }
        STC(J) = TFRZ + HEATR(J) * FACT(J);
      }
    }
  }
  // (3) warm layers melt ice elsewhere
  if (any_stc_gt() && (MICE(1) > 0.f || MICE(2) > 0.f || MICE(3) > 0.f || MICE(4) > 0.f)) {
#pragma unroll
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) > TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
        XM(J) = HEATR(J) * DT / HFUS;
#pragma unroll
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && MICE(K) > 0.f && XM(J) > 0.1f) {
            if (MICE(K) > XM(J)) {
              MICE(K) = MICE(K) - XM(J);
              STC(K) = TFRZ;
              XM(J) = 0.0f;
            } else {
              XM(J) = XM(J) - MICE(K);
              MICE(K) = 0.0f;
              STC(K) = TFRZ;
            }
            MLIQ(K) = MAX(0.f, WMASS0(K) - MICE(K));
          }
        }
        HEATR(J) = XM(J) * HFUS / DT;
        STC(J) = TFRZ + HEATR(J) * FACT(J);
      }
    }
  }
  // (4) cold layers freeze liquid elsewhere
  if (any_stc_lt() && (MLIQ(1) > 0.f || MLIQ(2) > 0.f || MLIQ(3) > 0.f || MLIQ(4) > 0.f)) {
#pragma unroll
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) < TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
        XM(J) = HEATR(J) * DT / HFUS;
#pragma unroll
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && MLIQ(K) > 0.f && XM(J) < -0.1f) {
            if (MLIQ(K) > ABS(XM(J))) {
              MICE(K) = MICE(K) - XM(J);
              STC(K) = TFRZ;
              XM(J) = 0.0f;
            } else {
              XM(J) = XM(J) + MLIQ(K);
              MICE(K) = WMASS0(K);
              STC(K) = TFRZ;
            }
            MLIQ(K) = MAX(0.f, WMASS0(K) - MICE(K));
          }
        }
        HEATR(J) = XM(J) * HFUS / DT;
        STC(J) = TFRZ + HEATR(J) * FACT(J);
      }
    }
  }
#pragma unroll
  for (int J = -2; J <= 0; ++J)
    if (J > ISNOW) { SNLIQ(J) = MLIQ(J); SNICE(J) = MICE(J); }
#pragma unroll
  for (int J = 1; J <= NSOIL; ++J) {
    SH2O(J) = MLIQ(J) / (1000.f * DZSNSO(J));
    SH2O(J) = MAX(0.0f, MIN(1.0f, SH2O(J)));
    SMC(J) = 1.0f;
  }
}

// glacier.F90:2113-2237
NMP_DEV void SNOWWATER_GLACIER(const I7& IMELT, float DT, float SFCTMP, float SNOWHIN, float QSNOW, float QSNFRO,
                               float QSNSUB, float QRAIN, const N3& FICEOLD, const S4& ZSOIL, int& ISNOW,
                               float& SNOWH, float& SNEQV, N3& SNICE, N3& SNLIQ, float& SH2O1, float& SICE1, L7& STC,
                               L7& DZSNSO, L7& ZSNSO, float& QSNBOT, float& SNOFLOW, float& PONDING1,
                               float& PONDING2) {
  SNOFLOW = 0.0f; PONDING1 = 0.0f; PONDING2 = 0.0f;
  const float DZ1 = DZSNSO(1);
  SnowPack p;
#pragma unroll
  for (int J = -2; J <= 0; ++J) {
    p.dz[J + 2] = DZSNSO(J); p.ice[J + 2] = SNICE(J); p.liq[J + 2] = SNLIQ(J); p.t[J + 2] = STC(J);
  }
  SNOWFALL<true>(DT, QSNOW, SNOWHIN, SFCTMP, ISNOW, SNOWH, p, SNEQV);
  if (ISNOW < 0) {
    SnowPack pc = p;
    COMPACT(DT, pc, IMELT, FICEOLD, ISNOW, p);
    COMBINE<true>(ISNOW, SH2O1, SICE1, DZ1, p, SNOWH, SNEQV, PONDING1, PONDING2);
    DIVIDE<true>(ISNOW, p);
  }
  // empty layers are zeroed BEFORE SNOWH2O here (:2182-2188; the land routine does it after)
#pragma unroll
  for (int J = 0; J < NSNOW; ++J)
    if (J - 2 <= ISNOW) { p.ice[J] = 0.f; p.liq[J] = 0.f; p.t[J] = 0.f; p.dz[J] = 0.f; }
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ)
    if (IZ <= ISNOW) ZSNSO(IZ) = 0.f;
  SNOWH2O<true>(DT, QSNFRO, QSNSUB, QRAIN, ISNOW, p, DZ1, SNOWH, SNEQV, SH2O1, SICE1, QSNBOT, PONDING1, PONDING2);
#pragma unroll
  for (int J = -2; J <= 0; ++J) {
    DZSNSO(J) = p.dz[J + 2]; SNICE(J) = p.ice[J + 2]; SNLIQ(J) = p.liq[J + 2]; STC(J) = p.t[J + 2];
  }
  if (SNEQV > 2000.f) {
    float BDSNOW = SNICE(0) / DZSNSO(0);
    SNOFLOW = (SNEQV - 2000.f);
    SNICE(0) = SNICE(0) - SNOFLOW;
    DZSNSO(0) = DZSNSO(0) - SNOFLOW / BDSNOW;
    SNOFLOW = SNOFLOW / DT;
  }
  if (ISNOW != 0) {
    SNEQV = 0.f;
#pragma unroll
    for (int IZ = -2; IZ <= 0; ++IZ)
      if (IZ > ISNOW) SNEQV = SNEQV + SNICE(IZ) + SNLIQ(IZ);
  }
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ)
    if (IZ > ISNOW) DZSNSO(IZ) = -DZSNSO(IZ);
  DZSNSO(1) = ZSOIL(1);
#pragma unroll
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = (ZSOIL(IZ) - ZSOIL(IZ - 1));
  const TopLayer top_new(ISNOW);
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ) {
    if (top_new.is(IZ)) ZSNSO(IZ) = DZSNSO(IZ);
    else if (IZ > ISNOW + 1) ZSNSO(IZ) = ZSNSO(IZ - (IZ > -2 ? 1 : 0)) + DZSNSO(IZ);
  }
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ)
    if (IZ > ISNOW) DZSNSO(IZ) = -DZSNSO(IZ);
}

// glacier.F90:150-338 with ATM_GLACIER (:340-390), ENERGY_GLACIER (:393-573), WATER_GLACIER (:1924-2110),
// ERROR_GLACIER (:2898-2972).  Uses the glacier subset of Col; land-only members are filled by the caller.
template <class O>
NMP_DEV void NOAHMP_GLACIER(Ctx& c, Col& g) {
  const float ZBOT = -8.0f;  // glacier.F90:260
  float THAIR, QAIR, EAIR, RHOAIR, SWDOWN, QPRECC, QPRECL, QMELT = 0.f;
  B2 SOLAD, SOLAI;
  L7 DZSNSO;
  S4 SICE;
  I7 IMELT;
  ATM(g.SFCPRS, g.SFCTMP, g.Q2, g.PRCP, g.SOLDN, g.COSZ, THAIR, QAIR, EAIR, RHOAIR, QPRECC, QPRECL, SOLAD, SOLAI,
      SWDOWN);
  const float BEG_WB = g.SNEQV;
  const TopLayer top0(g.ISNOW);
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ) {
    DZSNSO(IZ) = 0.f;
    if (top0.is(IZ)) DZSNSO(IZ) = -g.ZSNSO(IZ);
    else if (IZ > g.ISNOW + 1) DZSNSO(IZ) = g.ZSNSO(IZ - (IZ > -2 ? 1 : 0)) - g.ZSNSO(IZ);
  }
  // ---- ENERGY_GLACIER ----
  const float LATHEA = HSUB;
  {
    L7 DF, HCPCT, FACT;
#pragma unroll
    for (int K = -2; K <= NSOIL; ++K) { DF(K) = 0.f; HCPCT(K) = 0.f; FACT(K) = 0.f; }
    float UR = MAX(SQRT(POWR2(g.UU) + POWR2(g.VV)), 1.f);
    float Z0MG = Z0SNO;
    float ZPD = g.SNOWH;
    float ZLVL = ZPD + g.ZLVL;
    THERMOPROP_GLACIER(g.ISNOW, DZSNSO, g.DT, g.SNOWH, g.SNICE, g.SNLIQ, DF, HCPCT, FACT);
    RADIATION_GLACIER<O>(c, g.DT, g.TG, g.SNEQVO, g.SNEQV, g.COSZ, g.QSNOW, SOLAD, SOLAI, g.ALBOLD, g.TAUSS, g.SAG,
                         g.FSR, g.FSA);
    const float EMG = 0.98f, RHSUR = 1.0f, RSURF = 1.0f;
    float GAMMA = CPAIR * g.SFCPRS / (0.622f * LATHEA);
    NMP_PHASE_MAJOR();
    GLACIER_FLUX<O>(c, EMG, top7(DF, g.ISNOW), top7(DZSNSO, g.ISNOW), top7(g.STC, g.ISNOW), Z0MG, ZLVL, ZPD, QAIR,
                    g.SFCTMP, RHOAIR, g.SFCPRS, UR, GAMMA, RSURF, g.LWDN, RHSUR, g.SMC, EAIR, g.SAG, g.SNOWH, LATHEA,
                    g.SH2O, g.CM, g.CH, g.TG, g.QSFC, g.FIRA, g.FSH, g.FGEV, g.SSOIL, g.T2MB, g.Q2B, g.CHB2);
    float FIRE = g.LWDN + g.FIRA;
    if (FIRE <= 0.f) c.fatal(NOAHMP_ERR_FIRE, FIRE);
    g.EMISSI = EMG;
    g.TRAD = POW((FIRE - (1.f - g.EMISSI) * g.LWDN) / (g.EMISSI * SB), 0.25f);
    NMP_PHASE_MAJOR();
    TSNOSOI<O>(c, g.ISNOW, g.TBOT, g.ZSNSO, g.SSOIL, DF, HCPCT, ZBOT, g.DT, g.SNOWH, g.STC);
    if (NMP_OPT(stc) == 2) {
      if (g.SNOWH > 0.05f && g.TG > TFRZ) g.TG = TFRZ;
    }
    PHASECHANGE_GLACIER(g.ISNOW, g.DT, FACT, DZSNSO, g.STC, g.SNICE, g.SNLIQ, g.SNEQV, g.SNOWH, g.SMC, g.SH2O, QMELT,
                        IMELT, g.PONDING);
  }
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) SICE(K) = MAX(0.0f, g.SMC(K) - g.SH2O(K));
  g.SNEQVO = g.SNEQV;
  float QVAP = MAX(g.FGEV / LATHEA, 0.f);
  float QDEW = ABS(MIN(g.FGEV / LATHEA, 0.f));
  g.EDIR = QVAP - QDEW;
  NMP_PHASE_MAJOR();
  // ---- WATER_GLACIER ----
  {
    float SNOFLOW = 0.f;
    g.RUNSUB = 0.f; g.RUNSRF = 0.f;
    S4 SICE_SAVE = SICE, SH2O_SAVE = g.SH2O;
    g.FPICE = 0.f;
    const int snf = NMP_OPT(snf);
    if (snf == 1) {
      if (g.SFCTMP > TFRZ + 2.5f) {
        g.FPICE = 0.f;
      } else {
        if (g.SFCTMP <= TFRZ + 0.5f) g.FPICE = 1.0f;
        else if (g.SFCTMP <= TFRZ + 2.f) g.FPICE = 1.f - (-54.632f + 0.2f * g.SFCTMP);
        else g.FPICE = 0.6f;
      }
    }
    if (snf == 2) {
      if (g.SFCTMP >= TFRZ + 2.2f) g.FPICE = 0.f; else g.FPICE = 1.0f;
    }
    if (snf == 3) {
      if (g.SFCTMP >= TFRZ) g.FPICE = 0.f; else g.FPICE = 1.0f;
    }
    float BDFALL = MIN(120.f, 67.92f + 51.25f * EXP((g.SFCTMP - TFRZ) / 2.59f));
    float QRAIN = g.PRCP * (1.f - g.FPICE);
    g.QSNOW = g.PRCP * g.FPICE;
    float SNOWHIN = g.QSNOW / BDFALL;
    float QSNSUB = QVAP;
    float QSNFRO = QDEW;
    SNOWWATER_GLACIER(IMELT, g.DT, g.SFCTMP, SNOWHIN, g.QSNOW, QSNFRO, QSNSUB, QRAIN, g.FICEOLD, g.ZSOIL, g.ISNOW,
                      g.SNOWH, g.SNEQV, g.SNICE, g.SNLIQ, g.SH2O(1), SICE(1), g.STC, DZSNSO, g.ZSNSO, g.QSNBOT,
                      SNOFLOW, g.PONDING1, g.PONDING2);
    g.RUNSRF = (g.PONDING + g.PONDING1 + g.PONDING2) / g.DT;
    if (g.ISNOW == 0) g.RUNSRF = g.RUNSRF + g.QSNBOT + QRAIN;
    else g.RUNSRF = g.RUNSRF + g.QSNBOT;
    float REPLACE = 0.0f;
#pragma unroll
    for (int ILEV = 1; ILEV <= NSOIL; ++ILEV)
      REPLACE = REPLACE + DZSNSO(ILEV) * (SICE(ILEV) - SICE_SAVE(ILEV) + g.SH2O(ILEV) - SH2O_SAVE(ILEV));
    REPLACE = REPLACE * 1000.0f / g.DT;
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) {
      SICE(K) = MIN(1.0f, SICE_SAVE(K));
      g.SH2O(K) = 1.0f - SICE(K);
    }
    g.RUNSUB = SNOFLOW + REPLACE;
  }
  // ---- ERROR_GLACIER: one-sided SW / energy tests ----
  g.ERRSW = SWDOWN - (g.FSA + g.FSR);
  if (g.ERRSW > 0.01f) c.fatal(NOAHMP_ERR_ERRSW, g.ERRSW);
  g.ERRENG = g.SAG - (g.FIRA + g.FSH + g.FGEV + g.SSOIL);
  if (g.ERRENG > 0.01f) c.fatal(NOAHMP_ERR_ERRENG, g.ERRENG);
  float END_WB = g.SNEQV;
  g.ERRWAT = END_WB - BEG_WB - (g.PRCP - g.EDIR - g.RUNSRF - g.RUNSUB) * g.DT;
  if (ABS(g.ERRWAT) > 0.1f) c.fatal(NOAHMP_ERR_ERRWAT, g.ERRWAT);
  if (g.SNOWH <= 1.E-6f || g.SNEQV <= 1.E-3f) {
    g.SNOWH = 0.0f;
    g.SNEQV = 0.0f;
  }
  if (SWDOWN != 0.f) g.ALBEDO = g.FSR / SWDOWN; else g.ALBEDO = -999.9f;
}

}  // namespace nmp
