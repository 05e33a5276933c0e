// nmo_glacier.cpp — ORACLE (test infrastructure): NOAHMP_GLACIER tree.
// Restates phys/module_sf_noahmp_glacier.F90:150-2972. Routines whose bodies are identical to their
// land counterparts (CSNOW, SNOW_AGE, SNOWALB_*, SFCDIF1, ESAT, HRT/HSTEP/ROSR12, COMBO, COMPACT; see the
// mechanical diff in SURVEY.md §8a) call the land restatement; the rest is restated here.
#include "nmo_land.h"

namespace nmo {

static inline float TDC(float T) { return MIN(50.f, MAX(-50.f, (T - TFRZ))); }

// glacier.F90:575-645
static void THERMOPROP_GLACIER(int ISNOW, const ASnSo& DZSNSO, float DT, float SNOWH, const ASnow& SNICE,
                               const ASnow& SNLIQ, ASnSo& DF, ASnSo& HCPCT, ASnow& SNICEV, ASnow& SNLIQV,
                               ASnow& EPORE, ASnSo& FACT) {
  ASnow CVSNO, TKSNO;
  CSNOW(ISNOW, SNICE, SNLIQ, DZSNSO, TKSNO, CVSNO, SNICEV, SNLIQV, EPORE);
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) { DF(IZ) = TKSNO(IZ); HCPCT(IZ) = CVSNO(IZ); }
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    float ZMID = 0.5f * (DZSNSO(IZ));
    for (int IZ2 = 1; IZ2 <= IZ - 1; ++IZ2) ZMID = ZMID + DZSNSO(IZ2);
    HCPCT(IZ) = 1.E6f * (0.8194f + 0.1309f * ZMID);
    DF(IZ) = 0.32333f + (0.10073f * ZMID);
  }
  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) FACT(IZ) = DT / (HCPCT(IZ) * DZSNSO(IZ));
  if (ISNOW == 0) DF(1) = (DF(1) * DZSNSO(1) + 0.35f * SNOWH) / (SNOWH + DZSNSO(1));
  else DF(1) = (DF(1) * DZSNSO(1) + DF(0) * DZSNSO(0)) / (DZSNSO(0) + DZSNSO(1));
}

// glacier.F90:704-792
static void RADIATION_GLACIER(Ctx& c, float DT, float TG, float SNEQVO, float SNEQV, float COSZ, float QSNOW,
                              const ABand& SOLAD, const ABand& SOLAI, float& ALBOLD, float& TAUSS,
                              float& SAG, float& FSR, float& FSA) {
  ABand ALBSND, ALBSNI, ALBICE;
  ALBSND.fill(0.f); ALBSNI.fill(0.f);
  ALBICE(1) = 0.80f; ALBICE(2) = 0.55f;
  float FAGE, ALB;
  SNOW_AGE(DT, TG, SNEQVO, SNEQV, TAUSS, FAGE);
  if (c.O.OPT_ALB == 1) SNOWALB_BATS(0.f, COSZ, FAGE, ALBSND, ALBSNI);
  if (c.O.OPT_ALB == 2) {
    SNOWALB_CLASS(QSNOW, DT, ALB, ALBOLD, ALBSND, ALBSNI);
    ALBOLD = ALB;
  }
  SAG = 0.f; FSA = 0.f; FSR = 0.f;
  float FSNO = 0.0f;
  if (SNEQV > 0.0f) FSNO = 1.0f;
  for (int IB = 1; IB <= 2; ++IB) {
    ALBSND(IB) = ALBICE(IB) * (1.f - FSNO) + ALBSND(IB) * FSNO;
    ALBSNI(IB) = ALBICE(IB) * (1.f - FSNO) + ALBSNI(IB) * FSNO;
    float ABS_ = SOLAD(IB) * (1.f - ALBSND(IB)) + SOLAI(IB) * (1.f - ALBSNI(IB));
    SAG = SAG + ABS_;
    FSA = FSA + ABS_;
    float REF = SOLAD(IB) * ALBSND(IB) + SOLAI(IB) * ALBSNI(IB);
    FSR = FSR + REF;
  }
}

// glacier.F90:942-1148
static void GLACIER_FLUX(Ctx& c, float EMG, int ISNOW, const ASnSo& DF, const ASnSo& DZSNSO, float Z0M,
                         float ZLVL, float ZPD, float QAIR, float SFCTMP, float RHOAIR, float SFCPRS,
                         float UR, float GAMMA, float RSURF, float LWDN, float RHSUR, const ASoil& SMC,
                         float EAIR, const ASnSo& STC, float SAG, float SNOWH, float LATHEA,
                         const ASoil& SH2O, float& CM, float& CH, float& TGB, float& QSFC, float& IRB,
                         float& SHB, float& EVB, float& GHB, float& T2MB, float& Q2B, float& EHB2) {
  const int NITERB = 5;
  const float MPE = 1E-6f;
  int MOZSGN = 0;
  float H = 0.f, FV = 0.1f, MOZ = 0.f, FM = 0.f, FH = 0.f, FM2 = 0.f, FH2 = 0.f, CH2 = 0.f;
  float ESATW, ESATI, DSATW, DSATI, ESTG = 0.f, DESTG, CSH = 0.f, CEV = 0.f, RAHB = 1.f, Z0H = Z0M;
  float CIR = EMG * SB;
  float CGH = 2.f * DF(ISNOW + 1) / DZSNSO(ISNOW + 1);
  for (int ITER = 1; ITER <= NITERB; ++ITER) {
    Z0H = Z0M;
    SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE, MOZ, MOZSGN, FM, FH, FM2, FH2, CM,
            CH, FV, CH2);
    float RAMB = MAX(1.f, 1.f / (CM * UR));
    RAHB = MAX(1.f, 1.f / (CH * UR));
    float RAWB = RAHB;
    (void)RAMB;
    float T = TDC(TGB);
    ESAT(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.f) { ESTG = ESATW; DESTG = DSATW; }
    else { ESTG = ESATI; DESTG = DSATI; }
    CSH = RHOAIR * CPAIR / RAHB;
    CEV = RHOAIR * CPAIR / GAMMA / (RSURF + RAWB);
    IRB = CIR * POWI(TGB, 4) - EMG * LWDN;
    SHB = CSH * (TGB - SFCTMP);
    EVB = CEV * (ESTG * RHSUR - EAIR);
    GHB = CGH * (TGB - STC(ISNOW + 1));
    float B = SAG - IRB - SHB - EVB - GHB;
    float A = 4.f * CIR * POWI(TGB, 3) + CSH + CEV * DESTG + CGH;
    float DTG = B / A;
    IRB = IRB + 4.f * CIR * POWI(TGB, 3) * DTG;
    SHB = SHB + CSH * DTG;
    EVB = EVB + CEV * DESTG * DTG;
    GHB = GHB + CGH * DTG;
    TGB = TGB + DTG;
    H = CSH * (TGB - SFCTMP);
    T = TDC(TGB);
    ESAT(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.f) ESTG = ESATW; else ESTG = ESATI;
    QSFC = 0.622f * (ESTG * RHSUR) / (SFCPRS - 0.378f * (ESTG * RHSUR));
  }
  float SICEMAXV = SMC(1) - SH2O(1);
  for (int K = 2; K <= NSOIL; ++K) SICEMAXV = MAX(SICEMAXV, SMC(K) - SH2O(K));
  if (c.O.OPT_STC == 1) {
    if ((SICEMAXV > 0.0f || SNOWH > 0.0f) && TGB > TFRZ) {
      TGB = TFRZ;
      IRB = CIR * POWI(TGB, 4) - EMG * LWDN;
      SHB = CSH * (TGB - SFCTMP);
      EVB = CEV * (ESTG * RHSUR - EAIR);
      GHB = SAG - (IRB + SHB + EVB);
    }
  }
  EHB2 = FV * VKC / (LOG((2.f + Z0H) / Z0H) - FH2);
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
static void PHASECHANGE_GLACIER(int ISNOW, float DT, const ASnSo& FACT, const ASnSo& DZSNSO, ASnSo& STC,
                                ASnow& SNICE, ASnow& SNLIQ, float& SNEQV, float& SNOWH, ASoil& SMC,
                                ASoil& SH2O, float& QMELT, IA<-NSNOW + 1, NSOIL>& IMELT, float& PONDING) {
  ASnSo HM, XM, WMASS0, WICE0, WLIQ0, MICE, MLIQ, HEATR;
  HM.fill(0.f); XM.fill(0.f); WMASS0.fill(0.f); WICE0.fill(0.f); WLIQ0.fill(0.f); MICE.fill(0.f);
  MLIQ.fill(0.f); HEATR.fill(0.f);
  for (int J = -NSNOW + 1; J <= NSOIL; ++J) IMELT(J) = 0;
  QMELT = 0.f; PONDING = 0.f;
  float XMF = 0.f;
  for (int J = ISNOW + 1; J <= 0; ++J) { MICE(J) = SNICE(J); MLIQ(J) = SNLIQ(J); }
  for (int J = 1; J <= NSOIL; ++J) {
    MLIQ(J) = SH2O(J) * DZSNSO(J) * 1000.f;
    MICE(J) = (SMC(J) - SH2O(J)) * DZSNSO(J) * 1000.f;
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    IMELT(J) = 0; HM(J) = 0.f; XM(J) = 0.f;
    WICE0(J) = MICE(J); WLIQ0(J) = MLIQ(J); WMASS0(J) = MICE(J) + MLIQ(J);
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    if (MICE(J) > 0.f && STC(J) >= TFRZ) IMELT(J) = 1;
    if (MLIQ(J) > 0.f && STC(J) < TFRZ) IMELT(J) = 2;
    if (ISNOW == 0 && SNEQV > 0.f && J == 1) {
      if (STC(J) >= TFRZ) IMELT(J) = 1;
    }
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    if (IMELT(J) > 0) {
      HM(J) = (STC(J) - TFRZ) / FACT(J);
      STC(J) = TFRZ;
    }
    if (IMELT(J) == 1 && HM(J) < 0.f) { HM(J) = 0.f; IMELT(J) = 0; }
    if (IMELT(J) == 2 && HM(J) > 0.f) { HM(J) = 0.f; IMELT(J) = 0; }
    XM(J) = HM(J) * DT / HFUS;
  }
  if (ISNOW == 0 && SNEQV > 0.f && XM(1) > 0.f) {
    float TEMP1 = SNEQV;
    SNEQV = MAX(0.f, TEMP1 - XM(1));
    float PROPOR = SNEQV / TEMP1;
    SNOWH = MAX(0.f, PROPOR * SNOWH);
    HEATR(1) = HM(1) - HFUS * (TEMP1 - SNEQV) / DT;
    if (HEATR(1) > 0.f) {
      XM(1) = HEATR(1) * DT / HFUS;
      HM(1) = HEATR(1);
      IMELT(1) = 1;
    } else {
      XM(1) = 0.f;
      HM(1) = 0.f;
      IMELT(1) = 0;
    }
    QMELT = MAX(0.f, (TEMP1 - SNEQV)) / DT;
    XMF = HFUS * QMELT;
    PONDING = TEMP1 - SNEQV;
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    if (IMELT(J) > 0 && ABS(HM(J)) > 0.f) {
      HEATR(J) = 0.f;
      if (XM(J) > 0.f) {
        MICE(J) = MAX(0.f, WICE0(J) - XM(J));
        HEATR(J) = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
      } else if (XM(J) < 0.f) {
        MICE(J) = MIN(WMASS0(J), WICE0(J) - XM(J));
        HEATR(J) = HM(J) - HFUS * (WICE0(J) - MICE(J)) / DT;
      }
      MLIQ(J) = MAX(0.f, WMASS0(J) - MICE(J));
      if (ABS(HEATR(J)) > 0.f) {
        STC(J) = STC(J) + FACT(J) * HEATR(J);
        if (J <= 0) {
          if (MLIQ(J) * MICE(J) > 0.f) STC(J) = TFRZ;
        }
      }
      if (J > 0) XMF = XMF + HFUS * (WICE0(J) - MICE(J)) / DT;
      if (J < 1) QMELT = QMELT + MAX(0.f, (WICE0(J) - MICE(J))) / DT;
    }
  }
  HEATR.fill(0.f);
  XM.fill(0.f);

  auto any_gt = [&](const ASnSo& a, float v) { return a(1) > v || a(2) > v || a(3) > v || a(4) > v; };
  auto any_lt = [&](const ASnSo& a, float v) { return a(1) < v || a(2) < v || a(3) < v || a(4) < v; };

  // (1) warm layers give heat to cold layers (:1804-1825)
  if (any_gt(STC, TFRZ) && any_lt(STC, TFRZ)) {
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) > TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
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
  // (2) cold layers take heat from warm layers (:1829-1850)
  if (any_gt(STC, TFRZ) && any_lt(STC, TFRZ)) {
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) < TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
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
  // (3) warm layers melt ice elsewhere (:1854-1879)
  if (any_gt(STC, TFRZ) && any_gt(MICE, 0.f)) {
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) > TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
        XM(J) = HEATR(J) * DT / HFUS;
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && MICE(K) > 0.f && XM(J) > 0.1f) {
            if (MICE(K) > XM(J)) {
              MICE(K) = MICE(K) - XM(J);
              XMF = XMF + HFUS * XM(J) / DT;
              STC(K) = TFRZ;
              XM(J) = 0.0f;
            } else {
              XM(J) = XM(J) - MICE(K);
              XMF = XMF + HFUS * MICE(K) / DT;
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
  // (4) cold layers freeze liquid elsewhere (:1883-1908)
  if (any_lt(STC, TFRZ) && any_gt(MLIQ, 0.f)) {
    for (int J = 1; J <= NSOIL; ++J) {
      if (STC(J) < TFRZ) {
        HEATR(J) = (STC(J) - TFRZ) / FACT(J);
        XM(J) = HEATR(J) * DT / HFUS;
        for (int K = 1; K <= NSOIL; ++K) {
          if (J != K && MLIQ(K) > 0.f && XM(J) < -0.1f) {
            if (MLIQ(K) > ABS(XM(J))) {
              MICE(K) = MICE(K) - XM(J);
              XMF = XMF + HFUS * XM(J) / DT;
              STC(K) = TFRZ;
              XM(J) = 0.0f;
            } else {
              XM(J) = XM(J) + MLIQ(K);
              XMF = XMF - HFUS * MLIQ(K) / DT;
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
  (void)XMF;
  for (int J = ISNOW + 1; J <= 0; ++J) { SNLIQ(J) = MLIQ(J); SNICE(J) = MICE(J); }
  for (int J = 1; J <= NSOIL; ++J) {
    SH2O(J) = MLIQ(J) / (1000.f * DZSNSO(J));
    SH2O(J) = MAX(0.0f, MIN(1.0f, SH2O(J)));
    SMC(J) = 1.0f;
  }
}

// glacier.F90:393-573
static void ENERGY_GLACIER(Ctx& c, GlacIO& g, int ISNOW, float RHOAIR, float EAIR, float QAIR,
                           const ABand& SOLAD, const ABand& SOLAI, float ZBOT, ASnSo& DZSNSO,
                           IA<-NSNOW + 1, NSOIL>& IMELT, float& QMELT, float& LATHEA) {
  ASnSo DF, HCPCT, FACT;
  DF.fill(0.f); HCPCT.fill(0.f); FACT.fill(0.f);
  ASnow SNICEV, SNLIQV, EPORE;
  float UR = MAX(SQRT(POW(g.UU, 2.f) + POW(g.VV, 2.f)), 1.f);  // UU**2. : a REAL exponent, libm pow (nmo.h)
  float Z0MG = Z0SNO;
  float ZPD = g.SNOWH;
  float ZLVL = ZPD + g.ZLVL;
  THERMOPROP_GLACIER(ISNOW, DZSNSO, g.DT, g.SNOWH, g.SNICE, g.SNLIQ, DF, HCPCT, SNICEV, SNLIQV, EPORE, FACT);
  RADIATION_GLACIER(c, g.DT, g.TG, g.SNEQVO, g.SNEQV, g.COSZ, g.QSNOW, SOLAD, SOLAI, g.ALBOLD, g.TAUSS,
                    g.SAG, g.FSR, g.FSA);
  float EMG = 0.98f, RHSUR = 1.0f, RSURF = 1.0f;
  LATHEA = HSUB;
  float GAMMA = CPAIR * g.SFCPRS / (0.622f * LATHEA);
  GLACIER_FLUX(c, EMG, ISNOW, DF, DZSNSO, Z0MG, ZLVL, ZPD, QAIR, g.SFCTMP, RHOAIR, g.SFCPRS, UR, GAMMA,
               RSURF, g.LWDN, RHSUR, g.SMC, EAIR, g.STC, g.SAG, g.SNOWH, LATHEA, g.SH2O, g.CM, g.CH, g.TG,
               g.QSFC, g.FIRA, g.FSH, g.FGEV, g.SSOIL, g.T2M, g.Q2E, g.CH2B);
  float FIRE = g.LWDN + g.FIRA;
  if (FIRE <= 0.f) c.fatal(NOAHMP_ERR_FIRE, FIRE);
  g.EMISSI = EMG;
  g.TRAD = POW((FIRE - (1.f - g.EMISSI) * g.LWDN) / (g.EMISSI * SB), 0.25f);
  // TSNOSOI_GLACIER = HRT/HSTEP/ROSR12 with identical bodies (glacier.F90:1360-1632)
  TSNOSOI(c, -1, ISNOW, 1, g.TBOT, g.ZSNSO, g.SSOIL, DF, HCPCT, ZBOT, g.SAG, g.DT, g.SNOWH, DZSNSO, g.TG,
          g.STC);
  if (c.O.OPT_STC == 2) {
    if (g.SNOWH > 0.05f && g.TG > TFRZ) g.TG = TFRZ;
  }
  PHASECHANGE_GLACIER(ISNOW, g.DT, FACT, DZSNSO, g.STC, g.SNICE, g.SNLIQ, g.SNEQV, g.SNOWH, g.SMC, g.SH2O,
                      QMELT, IMELT, g.PONDING);
}

// glacier.F90:2239-2301
static void SNOWFALL_GLACIER(float DT, float QSNOW, float SNOWHIN, float SFCTMP, int& ISNOW, float& SNOWH,
                             ASnSo& DZSNSO, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, float& SNEQV) {
  int NEWNODE = 0;
  if (ISNOW == 0 && QSNOW > 0.f) {
    SNOWH = SNOWH + SNOWHIN * DT;
    SNEQV = SNEQV + QSNOW * DT;
  }
  if (ISNOW == 0 && QSNOW > 0.f && SNOWH >= 0.05f) {
    ISNOW = -1;
    NEWNODE = 1;
    DZSNSO(0) = SNOWH;
    SNOWH = 0.f;
    STC(0) = MIN(273.16f, SFCTMP);
    SNICE(0) = SNEQV;
    SNLIQ(0) = 0.f;
  }
  if (ISNOW < 0 && NEWNODE == 0 && QSNOW > 0.f) {
    SNICE(ISNOW + 1) = SNICE(ISNOW + 1) + QSNOW * DT;
    DZSNSO(ISNOW + 1) = DZSNSO(ISNOW + 1) + SNOWHIN * DT;
  }
}

// glacier.F90:2403-2571
static void COMBINE_GLACIER(int& ISNOW, ASoil& SH2O, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, ASnSo& DZSNSO,
                            ASoil& SICE, float& SNOWH, float& SNEQV, float& PONDING1, float& PONDING2) {
  static const float DZMIN[3] = {0.045f, 0.05f, 0.2f};
  int ISNOW_OLD = ISNOW;
  for (int J = ISNOW_OLD + 1; J <= 0; ++J) {
    if (SNICE(J) <= .1f) {
      if (J != 0) {
        SNLIQ(J + 1) = SNLIQ(J + 1) + SNLIQ(J);
        SNICE(J + 1) = SNICE(J + 1) + SNICE(J);
      } else {
        if (ISNOW_OLD < -1) {
          SNLIQ(J - 1) = SNLIQ(J - 1) + SNLIQ(J);
          SNICE(J - 1) = SNICE(J - 1) + SNICE(J);
        } else {
          PONDING1 = PONDING1 + SNLIQ(J);
          SNEQV = SNICE(J);
          SNOWH = DZSNSO(J);
          SNLIQ(J) = 0.0f;
          SNICE(J) = 0.0f;
          DZSNSO(J) = 0.0f;
        }
      }
      if (J > ISNOW + 1 && ISNOW < -1) {
        for (int I = J; I >= ISNOW + 2; --I) {
          STC(I) = STC(I - 1);
          SNLIQ(I) = SNLIQ(I - 1);
          SNICE(I) = SNICE(I - 1);
          DZSNSO(I) = DZSNSO(I - 1);
        }
      }
      ISNOW = ISNOW + 1;
    }
  }
  if (SICE(1) < 0.f) {
    SH2O(1) = SH2O(1) + SICE(1);
    SICE(1) = 0.f;
  }
  if (ISNOW == 0) return;
  SNEQV = 0.f; SNOWH = 0.f;
  float ZWICE = 0.f, ZWLIQ = 0.f;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    SNEQV = SNEQV + SNICE(J) + SNLIQ(J);
    SNOWH = SNOWH + DZSNSO(J);
    ZWICE = ZWICE + SNICE(J);
    ZWLIQ = ZWLIQ + SNLIQ(J);
  }
  if (SNOWH < 0.05f && ISNOW < 0) {
    ISNOW = 0;
    SNEQV = ZWICE;
    PONDING2 = PONDING2 + ZWLIQ;
    if (SNEQV <= 0.f) SNOWH = 0.f;
  }
  if (ISNOW < -1) {
    ISNOW_OLD = ISNOW;
    int MSSI = 1;
    for (int I = ISNOW_OLD + 1; I <= 0; ++I) {
      if (DZSNSO(I) < DZMIN[MSSI - 1]) {
        int NEIBOR;
        if (I == ISNOW + 1) NEIBOR = I + 1;
        else if (I == 0) NEIBOR = I - 1;
        else {
          NEIBOR = I + 1;
          if ((DZSNSO(I - 1) + DZSNSO(I)) < (DZSNSO(I + 1) + DZSNSO(I))) NEIBOR = I - 1;
        }
        int J, L;
        if (NEIBOR > I) { J = NEIBOR; L = I; }
        else { J = I; L = NEIBOR; }
        COMBO(DZSNSO(J), SNLIQ(J), SNICE(J), STC(J), DZSNSO(L), SNLIQ(L), SNICE(L), STC(L));
        if (J - 1 > ISNOW + 1) {
          for (int K = J - 1; K >= ISNOW + 2; --K) {
            STC(K) = STC(K - 1);
            SNICE(K) = SNICE(K - 1);
            SNLIQ(K) = SNLIQ(K - 1);
            DZSNSO(K) = DZSNSO(K - 1);
          }
        }
        ISNOW = ISNOW + 1;
        if (ISNOW >= -1) break;
      } else {
        MSSI = MSSI + 1;
      }
    }
  }
}

// glacier.F90:2626-2749
static void DIVIDE_GLACIER(int& ISNOW, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, ASnSo& DZSNSO) {
  FA<1, NSNOW> DZ, SWICE, SWLIQ, TSNO;
  DZ.fill(0.f); SWICE.fill(0.f); SWLIQ.fill(0.f); TSNO.fill(0.f);
  for (int J = 1; J <= NSNOW; ++J) {
    if (J <= std::abs(ISNOW)) {
      DZ(J) = DZSNSO(J + ISNOW);
      SWICE(J) = SNICE(J + ISNOW);
      SWLIQ(J) = SNLIQ(J + ISNOW);
      TSNO(J) = STC(J + ISNOW);
    }
  }
  int MSNO = std::abs(ISNOW);
  if (MSNO == 1) {
    if (DZ(1) > 0.05f) {
      MSNO = 2;
      DZ(1) = DZ(1) / 2.f;
      SWICE(1) = SWICE(1) / 2.f;
      SWLIQ(1) = SWLIQ(1) / 2.f;
      DZ(2) = DZ(1);
      SWICE(2) = SWICE(1);
      SWLIQ(2) = SWLIQ(1);
      TSNO(2) = TSNO(1);
    }
  }
  if (MSNO > 1) {
    if (DZ(1) > 0.05f) {
      float DRR = DZ(1) - 0.05f;
      float PROPOR = DRR / DZ(1);
      float ZWICE = PROPOR * SWICE(1);
      float ZWLIQ = PROPOR * SWLIQ(1);
      PROPOR = 0.05f / DZ(1);
      SWICE(1) = PROPOR * SWICE(1);
      SWLIQ(1) = PROPOR * SWLIQ(1);
      DZ(1) = 0.05f;
      COMBO(DZ(2), SWLIQ(2), SWICE(2), TSNO(2), DRR, ZWLIQ, ZWICE, TSNO(1));
      if (MSNO <= 2 && DZ(2) > 0.10f) {
        MSNO = 3;
        float DTDZ = (TSNO(1) - TSNO(2)) / ((DZ(1) + DZ(2)) / 2.f);
        DZ(2) = DZ(2) / 2.f;
        SWICE(2) = SWICE(2) / 2.f;
        SWLIQ(2) = SWLIQ(2) / 2.f;
        DZ(3) = DZ(2);
        SWICE(3) = SWICE(2);
        SWLIQ(3) = SWLIQ(2);
        TSNO(3) = TSNO(2) - DTDZ * DZ(2) / 2.f;
        if (TSNO(3) >= TFRZ) TSNO(3) = TSNO(2);
        else TSNO(2) = TSNO(2) + DTDZ * DZ(2) / 2.f;
      }
    }
  }
  if (MSNO > 2) {
    if (DZ(2) > 0.2f) {
      float DRR = DZ(2) - 0.2f;
      float PROPOR = DRR / DZ(2);
      float ZWICE = PROPOR * SWICE(2);
      float ZWLIQ = PROPOR * SWLIQ(2);
      PROPOR = 0.2f / DZ(2);
      SWICE(2) = PROPOR * SWICE(2);
      SWLIQ(2) = PROPOR * SWLIQ(2);
      DZ(2) = 0.2f;
      COMBO(DZ(3), SWLIQ(3), SWICE(3), TSNO(3), DRR, ZWLIQ, ZWICE, TSNO(2));
    }
  }
  ISNOW = -MSNO;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    DZSNSO(J) = DZ(J - ISNOW);
    SNICE(J) = SWICE(J - ISNOW);
    SNLIQ(J) = SWLIQ(J - ISNOW);
    STC(J) = TSNO(J - ISNOW);
  }
}

// glacier.F90:2751-2895
static void SNOWH2O_GLACIER(float DT, float QSNFRO, float QSNSUB, float QRAIN, int& ISNOW, ASnSo& DZSNSO,
                            float& SNOWH, float& SNEQV, ASnow& SNICE, ASnow& SNLIQ, ASoil& SH2O, ASoil& SICE,
                            ASnSo& STC, float& PONDING1, float& PONDING2, float& QSNBOT) {
  ASnow VOL_LIQ, VOL_ICE, EPORE;
  VOL_LIQ.fill(0.f); VOL_ICE.fill(0.f); EPORE.fill(0.f);
  if (SNEQV == 0.f) SICE(1) = SICE(1) + (QSNFRO - QSNSUB) * DT / (DZSNSO(1) * 1000.f);
  if (ISNOW == 0 && SNEQV > 0.f) {
    float TEMP = SNEQV;
    SNEQV = SNEQV - QSNSUB * DT + QSNFRO * DT;
    float PROPOR = SNEQV / TEMP;
    SNOWH = MAX(0.f, PROPOR * SNOWH);
    if (SNEQV < 0.f) {
      SICE(1) = SICE(1) + SNEQV / (DZSNSO(1) * 1000.f);
      SNEQV = 0.f;
      SNOWH = 0.f;
    }
    if (SICE(1) < 0.f) {
      SH2O(1) = SH2O(1) + SICE(1);
      SICE(1) = 0.f;
    }
  }
  if (SNOWH <= 1.E-8f || SNEQV <= 1.E-6f) {
    SNOWH = 0.0f;
    SNEQV = 0.0f;
  }
  if (ISNOW < 0) {
    float WGDIF = SNICE(ISNOW + 1) - QSNSUB * DT + QSNFRO * DT;
    SNICE(ISNOW + 1) = WGDIF;
    if (WGDIF < 1.e-6f && ISNOW < 0)
      COMBINE_GLACIER(ISNOW, SH2O, STC, SNICE, SNLIQ, DZSNSO, SICE, SNOWH, SNEQV, PONDING1, PONDING2);
    if (ISNOW < 0) {
      SNLIQ(ISNOW + 1) = SNLIQ(ISNOW + 1) + QRAIN * DT;
      SNLIQ(ISNOW + 1) = MAX(0.f, SNLIQ(ISNOW + 1));
    }
  }
  for (int J = -NSNOW + 1; J <= 0; ++J) {
    if (J >= ISNOW + 1) {
      VOL_ICE(J) = MIN(1.f, SNICE(J) / (DZSNSO(J) * DENICE));
      EPORE(J) = 1.f - VOL_ICE(J);
      VOL_LIQ(J) = MIN(EPORE(J), SNLIQ(J) / (DZSNSO(J) * DENH2O));
    }
  }
  float QIN = 0.f, QOUT = 0.f;
  for (int J = -NSNOW + 1; J <= 0; ++J) {
    if (J >= ISNOW + 1) {
      SNLIQ(J) = SNLIQ(J) + QIN;
      if (J <= -1) {
        if (EPORE(J) < 0.05f || EPORE(J + 1) < 0.05f) {
          QOUT = 0.f;
        } else {
          QOUT = MAX(0.f, (VOL_LIQ(J) - SSI * EPORE(J)) * DZSNSO(J));
          QOUT = MIN(QOUT, (1.f - VOL_ICE(J + 1) - VOL_LIQ(J + 1)) * DZSNSO(J + 1));
        }
      } else {
        QOUT = MAX(0.f, (VOL_LIQ(J) - SSI * EPORE(J)) * DZSNSO(J));
      }
      QOUT = QOUT * 1000.f;
      SNLIQ(J) = SNLIQ(J) - QOUT;
      QIN = QOUT;
    }
  }
  QSNBOT = QOUT / DT;
}

// glacier.F90:2113-2237
static void SNOWWATER_GLACIER(const IA<-NSNOW + 1, NSOIL>& IMELT, float DT, float SFCTMP, float SNOWHIN,
                              float QSNOW, float QSNFRO, float QSNSUB, float QRAIN, const ASnow& FICEOLD,
                              const ASoil& ZSOIL, int& ISNOW, float& SNOWH, float& SNEQV, ASnow& SNICE,
                              ASnow& SNLIQ, ASoil& SH2O, ASoil& SICE, ASnSo& STC, ASnSo& DZSNSO,
                              ASnSo& ZSNSO, float& QSNBOT, float& SNOFLOW, float& PONDING1,
                              float& PONDING2) {
  SNOFLOW = 0.0f; PONDING1 = 0.0f; PONDING2 = 0.0f;
  SNOWFALL_GLACIER(DT, QSNOW, SNOWHIN, SFCTMP, ISNOW, SNOWH, DZSNSO, STC, SNICE, SNLIQ, SNEQV);
  if (ISNOW < 0) {
    COMPACT(DT, STC, SNICE, SNLIQ, IMELT, FICEOLD, ISNOW, DZSNSO);
    COMBINE_GLACIER(ISNOW, SH2O, STC, SNICE, SNLIQ, DZSNSO, SICE, SNOWH, SNEQV, PONDING1, PONDING2);
    DIVIDE_GLACIER(ISNOW, STC, SNICE, SNLIQ, DZSNSO);
  }
  for (int IZ = -NSNOW + 1; IZ <= ISNOW; ++IZ) {
    SNICE(IZ) = 0.f; SNLIQ(IZ) = 0.f; STC(IZ) = 0.f; DZSNSO(IZ) = 0.f; ZSNSO(IZ) = 0.f;
  }
  SNOWH2O_GLACIER(DT, QSNFRO, QSNSUB, QRAIN, ISNOW, DZSNSO, SNOWH, SNEQV, SNICE, SNLIQ, SH2O, SICE, STC,
                  PONDING1, PONDING2, QSNBOT);
  if (SNEQV > 2000.f) {
    float BDSNOW = SNICE(0) / DZSNSO(0);
    SNOFLOW = (SNEQV - 2000.f);
    SNICE(0) = SNICE(0) - SNOFLOW;
    DZSNSO(0) = DZSNSO(0) - SNOFLOW / BDSNOW;
    SNOFLOW = SNOFLOW / DT;
  }
  if (ISNOW != 0) {
    SNEQV = 0.f;
    for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) SNEQV = SNEQV + SNICE(IZ) + SNLIQ(IZ);
  }
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) DZSNSO(IZ) = -DZSNSO(IZ);
  DZSNSO(1) = ZSOIL(1);
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = (ZSOIL(IZ) - ZSOIL(IZ - 1));
  ZSNSO(ISNOW + 1) = DZSNSO(ISNOW + 1);
  for (int IZ = ISNOW + 2; IZ <= NSOIL; ++IZ) ZSNSO(IZ) = ZSNSO(IZ - 1) + DZSNSO(IZ);
  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = -DZSNSO(IZ);
}

// glacier.F90:1924-2110
static void WATER_GLACIER(Ctx& c, GlacIO& g, const IA<-NSNOW + 1, NSOIL>& IMELT, float QVAP, float QDEW,
                          ASnSo& DZSNSO, ASoil& SICE) {
  float SNOFLOW = 0.f;
  g.RUNSUB = 0.f; g.RUNSRF = 0.f;
  ASoil SICE_SAVE = SICE, SH2O_SAVE = g.SH2O;
  g.FPICE = 0.f;
  if (c.O.OPT_SNF == 1) {
    if (g.SFCTMP > TFRZ + 2.5f) {
      g.FPICE = 0.f;
    } else {
      if (g.SFCTMP <= TFRZ + 0.5f) g.FPICE = 1.0f;
      else if (g.SFCTMP <= TFRZ + 2.f) g.FPICE = 1.f - (-54.632f + 0.2f * g.SFCTMP);
      else g.FPICE = 0.6f;
    }
  }
  if (c.O.OPT_SNF == 2) {
    if (g.SFCTMP >= TFRZ + 2.2f) g.FPICE = 0.f; else g.FPICE = 1.0f;
  }
  if (c.O.OPT_SNF == 3) {
    if (g.SFCTMP >= TFRZ) g.FPICE = 0.f; else g.FPICE = 1.0f;
  }
  float BDFALL = MIN(120.f, 67.92f + 51.25f * EXP((g.SFCTMP - TFRZ) / 2.59f));
  float QRAIN = g.PRCP * (1.f - g.FPICE);
  g.QSNOW = g.PRCP * g.FPICE;
  float SNOWHIN = g.QSNOW / BDFALL;
  float QSNSUB = QVAP;
  float QSNFRO = QDEW;
  SNOWWATER_GLACIER(IMELT, g.DT, g.SFCTMP, SNOWHIN, g.QSNOW, QSNFRO, QSNSUB, QRAIN, g.FICEOLD, g.ZSOIL,
                    g.ISNOW, g.SNOWH, g.SNEQV, g.SNICE, g.SNLIQ, g.SH2O, SICE, g.STC, DZSNSO, g.ZSNSO,
                    g.QSNBOT, SNOFLOW, g.PONDING1, g.PONDING2);
  g.RUNSRF = (g.PONDING + g.PONDING1 + g.PONDING2) / g.DT;
  if (g.ISNOW == 0) g.RUNSRF = g.RUNSRF + g.QSNBOT + QRAIN;
  else g.RUNSRF = g.RUNSRF + g.QSNBOT;
  float REPLACE = 0.0f;
  for (int ILEV = 1; ILEV <= NSOIL; ++ILEV)
    REPLACE = REPLACE + DZSNSO(ILEV) * (SICE(ILEV) - SICE_SAVE(ILEV) + g.SH2O(ILEV) - SH2O_SAVE(ILEV));
  REPLACE = REPLACE * 1000.0f / g.DT;
  for (int K = 1; K <= NSOIL; ++K) {
    SICE(K) = MIN(1.0f, SICE_SAVE(K));
    g.SH2O(K) = 1.0f - SICE(K);
  }
  g.RUNSUB = SNOFLOW + REPLACE;
}

// glacier.F90:150-338
void NOAHMP_GLACIER(Ctx& c, GlacIO& g) {
  const float ZBOT = -8.0f;  // glacier.F90:260
  float THAIR, QAIR, EAIR, RHOAIR, SWDOWN, QMELT = 0.f, LATHEA;
  ABand SOLAD, SOLAI;
  ASnSo DZSNSO; DZSNSO.fill(0.f);
  ASoil SICE;
  IA<-NSNOW + 1, NSOIL> IMELT;
  // ATM_GLACIER (glacier.F90:340-390)
  {
    float PAIR = g.SFCPRS;
    THAIR = g.SFCTMP * POW(g.SFCPRS / PAIR, RAIR / CPAIR);
    QAIR = g.Q2;
    EAIR = QAIR * g.SFCPRS / (0.622f + 0.378f * QAIR);
    RHOAIR = (g.SFCPRS - 0.378f * EAIR) / (RAIR * g.SFCTMP);
    if (g.COSZ <= 0.f) SWDOWN = 0.f; else SWDOWN = g.SOLDN;
    SOLAD(1) = SWDOWN * 0.7f * 0.5f;
    SOLAD(2) = SWDOWN * 0.7f * 0.5f;
    SOLAI(1) = SWDOWN * 0.3f * 0.5f;
    SOLAI(2) = SWDOWN * 0.3f * 0.5f;
    (void)THAIR;
  }
  float BEG_WB = g.SNEQV;
  for (int IZ = g.ISNOW + 1; IZ <= NSOIL; ++IZ) {
    if (IZ == g.ISNOW + 1) DZSNSO(IZ) = -g.ZSNSO(IZ);
    else DZSNSO(IZ) = g.ZSNSO(IZ - 1) - g.ZSNSO(IZ);
  }
  ENERGY_GLACIER(c, g, g.ISNOW, RHOAIR, EAIR, QAIR, SOLAD, SOLAI, ZBOT, DZSNSO, IMELT, QMELT, LATHEA);
  for (int K = 1; K <= NSOIL; ++K) SICE(K) = MAX(0.0f, g.SMC(K) - g.SH2O(K));
  g.SNEQVO = g.SNEQV;
  float QVAP = MAX(g.FGEV / LATHEA, 0.f);
  float QDEW = ABS(MIN(g.FGEV / LATHEA, 0.f));
  g.EDIR = QVAP - QDEW;
  WATER_GLACIER(c, g, IMELT, QVAP, QDEW, DZSNSO, SICE);
  // ERROR_GLACIER (glacier.F90:2898-2972): one-sided SW / energy tests
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

}  // namespace nmo

extern "C" {
// PHASECHANGE_GLACIER probe (glacier.F90:1635-1922): arrays in the oracle's Fortran bounds packed from the lowest index
void nmo_phasechange_glacier(int ISNOW, float DT, const float* FACT7, const float* DZSNSO7, float* STC7, float* SNICE3,
                             float* SNLIQ3, float* SNEQV, float* SNOWH, float* SMC4, float* SH2O4, float* QMELT, int* IMELT7,
                             float* PONDING) {
  using namespace nmo;
  ASnSo fact, dz, stc; ASnow ice, liq; ASoil smc, sh; IA<-NSNOW + 1, NSOIL> im;
  for (int k = -2; k <= NSOIL; ++k) { fact(k) = FACT7[k + 2]; dz(k) = DZSNSO7[k + 2]; stc(k) = STC7[k + 2]; im(k) = 0; }
  for (int k = -2; k <= 0; ++k) { ice(k) = SNICE3[k + 2]; liq(k) = SNLIQ3[k + 2]; }
  for (int k = 1; k <= NSOIL; ++k) { smc(k) = SMC4[k - 1]; sh(k) = SH2O4[k - 1]; }
  PHASECHANGE_GLACIER(ISNOW, DT, fact, dz, stc, ice, liq, *SNEQV, *SNOWH, smc, sh, *QMELT, im, *PONDING);
  for (int k = -2; k <= NSOIL; ++k) { STC7[k + 2] = stc(k); IMELT7[k + 2] = im(k); }
  for (int k = -2; k <= 0; ++k) { SNICE3[k + 2] = ice(k); SNLIQ3[k + 2] = liq(k); }
  for (int k = 1; k <= NSOIL; ++k) { SMC4[k - 1] = smc(k); SH2O4[k - 1] = sh(k); }
}
}

