// nmo_land1.cpp — ORACLE (test infrastructure): NOAHMP_SFLX orchestration, ATM, PHENOLOGY, ERROR,
// ENERGY, THERMOPROP/CSNOW/TDFCND, RADIATION/ALBEDO/SNOW_AGE/SNOWALB_*/GROUNDALB/TWOSTREAM/SURRAD.
// Restates phys/module_sf_noahmplsm.F90:518-3016 routine by routine (line refs at each function).
#include "nmo_land.h"

namespace nmo {

// noahmplsm.F90:949-1007
static void ATM(float SFCPRS, float SFCTMP, float Q2, float PRCP, float SOLDN, float COSZ, float& THAIR,
                float& QAIR, float& EAIR, float& RHOAIR, float& QPRECC, float& QPRECL, ABand& SOLAD,
                ABand& SOLAI, float& SWDOWN) {
  float PAIR = SFCPRS;
  THAIR = SFCTMP * POW(SFCPRS / PAIR, RAIR / CPAIR);
  QAIR = Q2;
  EAIR = QAIR * SFCPRS / (0.622f + 0.378f * QAIR);
  RHOAIR = (SFCPRS - 0.378f * EAIR) / (RAIR * SFCTMP);
  QPRECC = 0.10f * PRCP;
  QPRECL = 0.90f * PRCP;
  if (COSZ <= 0.f) SWDOWN = 0.f; else SWDOWN = SOLDN;
  SOLAD(1) = SWDOWN * 0.7f * 0.5f;
  SOLAD(2) = SWDOWN * 0.7f * 0.5f;
  SOLAI(1) = SWDOWN * 0.3f * 0.5f;
  SOLAI(2) = SWDOWN * 0.3f * 0.5f;
}

// noahmplsm.F90:1010-1104
static void PHENOLOGY(Ctx& c, int VEGTYP, int ISURBAN, float SNOWH, float TV, float LAT, int YEARLEN,
                      float JULIAN, float& LAI, float& SAI, float TROOT, float& HTOP, float& ELAI,
                      float& ESAI, float& IGS) {
  const noahmp_tables& T = *c.T;
  (void)TROOT;
  if (c.O.DVEG == 1 || c.O.DVEG == 3 || c.O.DVEG == 4) {
    float DAY;
    if (LAT >= 0.f) {
      DAY = JULIAN;
    } else {
      DAY = std::fmod(JULIAN + (0.5f * (float)YEARLEN), (float)YEARLEN);
    }
    float Tm = 12.f * DAY / (float)YEARLEN;
    int IT1 = (int)(Tm + 0.5f);
    int IT2 = IT1 + 1;
    float WT1 = ((float)IT1 + 0.5f) - Tm;
    float WT2 = 1.f - WT1;
    if (IT1 < 1) IT1 = 12;
    if (IT2 > 12) IT2 = 1;
    LAI = WT1 * TV2(T.laim, VEGTYP, IT1) + WT2 * TV2(T.laim, VEGTYP, IT2);
    SAI = WT1 * TV2(T.saim, VEGTYP, IT1) + WT2 * TV2(T.saim, VEGTYP, IT2);
  }
  if (SAI < 0.01f) SAI = 0.0f;
  if (LAI < 0.05f || SAI == 0.0f) LAI = 0.0f;
  if (VEGTYP == T.iswater || VEGTYP == T.isbarren || VEGTYP == T.issnow || VEGTYP == ISURBAN) {
    LAI = 0.f;
    SAI = 0.f;
  }
  float hvt = TV1(T.hvt, VEGTYP), hvb = TV1(T.hvb, VEGTYP);
  float DB = MIN(MAX(SNOWH - hvb, 0.f), hvt - hvb);
  float FB = DB / MAX(1.E-06f, hvt - hvb);
  if (hvt > 0.f && hvt <= 1.0f) {
    float SNOWHC = hvt * EXP(-SNOWH / 0.2f);
    FB = MIN(SNOWH, SNOWHC) / SNOWHC;
  }
  ELAI = LAI * (1.f - FB);
  ESAI = SAI * (1.f - FB);
  if (ESAI < 0.01f) ESAI = 0.0f;
  if (ELAI < 0.05f || ESAI == 0.0f) ELAI = 0.0f;
  if (TV > TV1(T.tmin, VEGTYP)) IGS = 1.f; else IGS = 0.f;
  HTOP = hvt;
}

// noahmplsm.F90:1106-1228
static void ERROR(Ctx& c, SflxIO& s, const SflxLocal& L) {
  float ERRSW = L.SWDOWN - (s.FSA + s.FSR);
  s.ERRSW = ERRSW;
  if (ABS(ERRSW) > 0.01f) c.fatal(NOAHMP_ERR_ERRSW, ERRSW);
  float ERRENG = s.SAV + s.SAG - (s.FIRA + s.FSH + s.FCEV + s.FGEV + s.FCTR + s.SSOIL);
  s.ERRENG = ERRENG;
  if (ABS(ERRENG) > 0.01f) c.fatal(NOAHMP_ERR_ERRENG, ERRENG);
  if (s.IST == 1) {
    float END_WB = s.CANLIQ + s.CANICE + s.SNEQV + s.WA;
    for (int IZ = 1; IZ <= NSOIL; ++IZ) END_WB = END_WB + s.SMC(IZ) * L.DZSNSO(IZ) * 1000.f;
    s.ERRWAT = END_WB - L.BEG_WB - (s.PRCP - s.ECAN - s.ETRAN - s.EDIR - s.RUNSRF - s.RUNSUB) * s.DT;
    if (ABS(s.ERRWAT) > 0.1f) c.fatal(NOAHMP_ERR_ERRWAT, s.ERRWAT);
  } else {
    s.ERRWAT = 0.0f;
  }
}

// noahmplsm.F90:518-947
void NOAHMP_SFLX(Ctx& c, SflxIO& s) {
  const noahmp_tables& T = *c.T;
  SflxLocal L;
  std::memset(&L, 0, sizeof(L));
  s.NEE = 0.0f; s.NPP = 0.0f; s.GPP = 0.0f;

  ATM(s.SFCPRS, s.SFCTMP, s.Q2, s.PRCP, s.SOLDN, s.COSZ, L.THAIR, L.QAIR, L.EAIR, L.RHOAIR, L.QPRECC,
      L.QPRECL, L.SOLAD, L.SOLAI, L.SWDOWN);

  for (int IZ = s.ISNOW + 1; IZ <= NSOIL; ++IZ) {
    if (IZ == s.ISNOW + 1) L.DZSNSO(IZ) = -s.ZSNSO(IZ);
    else L.DZSNSO(IZ) = s.ZSNSO(IZ - 1) - s.ZSNSO(IZ);
  }

  L.TROOT = 0.f;
  for (int IZ = 1; IZ <= c.P.NROOT; ++IZ)
    L.TROOT = L.TROOT + s.STC(IZ) * L.DZSNSO(IZ) / (-s.ZSOIL(c.P.NROOT));

  if (s.IST == 1) {
    L.BEG_WB = s.CANLIQ + s.CANICE + s.SNEQV + s.WA;
    for (int IZ = 1; IZ <= NSOIL; ++IZ) L.BEG_WB = L.BEG_WB + s.SMC(IZ) * L.DZSNSO(IZ) * 1000.f;
  }

  PHENOLOGY(c, s.VEGTYP, s.ISURBAN, s.SNOWH, s.TV, s.LAT, s.YEARLEN, s.JULIAN, s.LAI, s.SAI, L.TROOT,
            L.HTOP, L.ELAI, L.ESAI, L.IGS);

  if (c.O.DVEG == 1) {
    s.FVEG = s.SHDFAC;
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else if (c.O.DVEG == 2 || c.O.DVEG == 3) {
    s.FVEG = 1.f - EXP(-0.52f * (s.LAI + s.SAI));
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else if (c.O.DVEG == 4 || c.O.DVEG == 5) {
    s.FVEG = s.SHDMAX;
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else {
    c.fatal(NOAHMP_ERR_OPTION, (float)c.O.DVEG);
    return;
  }
  if (s.VEGTYP == s.ISURBAN || s.VEGTYP == T.isbarren) s.FVEG = 0.0f;
  if (L.ELAI + L.ESAI == 0.0f) s.FVEG = 0.0f;

  ENERGY(c, s, L);

  for (int IZ = 1; IZ <= NSOIL; ++IZ) L.SICE(IZ) = MAX(0.0f, s.SMC(IZ) - s.SH2O(IZ));
  s.SNEQVO = s.SNEQV;

  L.QVAP = MAX(s.FGEV / L.LATHEAG, 0.f);
  L.QDEW = ABS(MIN(s.FGEV / L.LATHEAG, 0.f));
  s.EDIR = L.QVAP - L.QDEW;

  WATER(c, s, L);

  if (c.O.DVEG == 2 || c.O.DVEG == 5) CARBON(c, s, L);

  ERROR(c, s, L);

  float QFX = s.ETRAN + s.ECAN + s.EDIR;
  if (s.VEGTYP == s.ISURBAN) {
    s.QSFC = (QFX / L.RHOAIR * s.CH) + L.QAIR;
    s.Q2B = s.QSFC;
  }
  if (s.SNOWH <= 1.E-6f || s.SNEQV <= 1.E-3f) {
    s.SNOWH = 0.0f;
    s.SNEQV = 0.0f;
  }
  if (L.SWDOWN != 0.f) s.ALBEDO = s.FSR / L.SWDOWN; else s.ALBEDO = -999.9f;
  s.IMELT = L.IMELT;
}

// noahmplsm.F90:1957-2011
void CSNOW(int ISNOW, const ASnow& SNICE, const ASnow& SNLIQ, const ASnSo& DZSNSO, ASnow& TKSNO,
                  ASnow& CVSNO, ASnow& SNICEV, ASnow& SNLIQV, ASnow& EPORE) {
  ASnow BDSNOI;
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {
    SNICEV(IZ) = MIN(1.f, SNICE(IZ) / (DZSNSO(IZ) * DENICE));
    EPORE(IZ) = 1.f - SNICEV(IZ);
    SNLIQV(IZ) = MIN(EPORE(IZ), SNLIQ(IZ) / (DZSNSO(IZ) * DENH2O));
  }
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {
    BDSNOI(IZ) = (SNICE(IZ) + SNLIQ(IZ)) / DZSNSO(IZ);
    CVSNO(IZ) = CICE * SNICEV(IZ) + CWAT * SNLIQV(IZ);
  }
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) TKSNO(IZ) = 3.2217E-6f * POW(BDSNOI(IZ), 2.f);
}

// noahmplsm.F90:2014-2118
static void TDFCND(const Ctx& c, float& DF, float SMC, float SH2O) {
  const Params& P = c.P;
  float SATRATIO = SMC / P.SMCMAX;
  float THKW = 0.57f, THKO = 2.0f, THKQTZ = 7.7f;
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
  DF = AKE * (THKSAT - THKDRY) + THKDRY;
}

// noahmplsm.F90:1845-1954
static void THERMOPROP(Ctx& c, int ISNOW, int IST, const ASnSo& DZSNSO, float DT, float SNOWH,
                       const ASnow& SNICE, const ASnow& SNLIQ, float CSOIL, const ASoil& SMC,
                       const ASoil& SH2O, const ASnSo& STC, int VEGTYP, int ISURBAN, ASnSo& DF,
                       ASnSo& HCPCT, ASnow& SNICEV, ASnow& SNLIQV, ASnow& EPORE, ASnSo& FACT) {
  ASnow CVSNO, TKSNO;
  ASoil SICE;
  CSNOW(ISNOW, SNICE, SNLIQ, DZSNSO, TKSNO, CVSNO, SNICEV, SNLIQV, EPORE);
  for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {
    DF(IZ) = TKSNO(IZ);
    HCPCT(IZ) = CVSNO(IZ);
  }
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    SICE(IZ) = SMC(IZ) - SH2O(IZ);
    HCPCT(IZ) = SH2O(IZ) * CWAT + (1.0f - c.P.SMCMAX) * CSOIL + (c.P.SMCMAX - SMC(IZ)) * CPAIR +
                SICE(IZ) * CICE;
    TDFCND(c, DF(IZ), SMC(IZ), SH2O(IZ));
  }
  if (VEGTYP == ISURBAN) {
    for (int IZ = 1; IZ <= NSOIL; ++IZ) DF(IZ) = 3.24f;
  }
  if (IST == 2) {
    for (int IZ = 1; IZ <= NSOIL; ++IZ) {
      if (STC(IZ) > TFRZ) { HCPCT(IZ) = CWAT; DF(IZ) = TKWAT; }
      else { HCPCT(IZ) = CICE; DF(IZ) = TKICE; }
    }
  }
  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) FACT(IZ) = DT / (HCPCT(IZ) * DZSNSO(IZ));
  if (ISNOW == 0) {
    DF(1) = (DF(1) * DZSNSO(1) + 0.35f * SNOWH) / (SNOWH + DZSNSO(1));
  } else {
    DF(1) = (DF(1) * DZSNSO(1) + DF(0) * DZSNSO(0)) / (DZSNSO(0) + DZSNSO(1));
  }
}

// noahmplsm.F90:2547-2596
void SNOW_AGE(float DT, float TG, float SNEQVO, float SNEQV, float& TAUSS, float& FAGE) {
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
void SNOWALB_BATS(float FSNO, float COSZ, float FAGE, ABand& ALBSND, ABand& ALBSNI) {
  (void)FSNO;
  const float C1 = 0.2f, C2 = 0.5f;
  ALBSND.fill(0.f); ALBSNI.fill(0.f);
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
void SNOWALB_CLASS(float QSNOW, float DT, float& ALB, float ALBOLD, ABand& ALBSND, ABand& ALBSNI) {
  ALBSND.fill(0.f); ALBSNI.fill(0.f);
  ALB = 0.55f + (ALBOLD - 0.55f) * EXP(-0.01f * DT / 3600.f);
  if (QSNOW > 0.f) ALB = ALB + MIN(QSNOW * DT, SWEMX) * (0.84f - ALB) / (SWEMX);
  ALBSNI(1) = ALB; ALBSNI(2) = ALB; ALBSND(1) = ALB; ALBSND(2) = ALB;
}

// NOAHMP_RAD_PARAMETERS data (noahmplsm.F90:427-445). ALBSAT/ALBDRY(9,:) are never set in the
// reference (DATA covers I=1,8); ISC is hard-wired to 4 by the dispatcher so they are never read.
static const float ALBSAT[2][9] = {{0.15f, 0.11f, 0.10f, 0.09f, 0.08f, 0.07f, 0.06f, 0.05f, 0.f},
                                   {0.30f, 0.22f, 0.20f, 0.18f, 0.16f, 0.14f, 0.12f, 0.10f, 0.f}};
static const float ALBDRY[2][9] = {{0.27f, 0.22f, 0.20f, 0.18f, 0.16f, 0.14f, 0.12f, 0.10f, 0.f},
                                   {0.54f, 0.44f, 0.40f, 0.36f, 0.32f, 0.28f, 0.24f, 0.20f, 0.f}};
static const float ALBLAK[2] = {0.60f, 0.40f};
static const float OMEGAS[2] = {0.8f, 0.4f};
static const float BETADS = 0.5f, BETAIS = 0.5f;
static const float EG[2] = {0.97f, 0.98f};

// noahmplsm.F90:2703-2765
static void GROUNDALB(int IST, int ISC, float FSNO, const ASoil& SMC, const ABand& ALBSND,
                      const ABand& ALBSNI, float COSZ, float TG, ABand& ALBGRD, ABand& ALBGRI) {
  for (int IB = 1; IB <= 2; ++IB) {
    float INC = MAX(0.11f - 0.40f * SMC(1), 0.f);
    float ALBSOD, ALBSOI;
    if (IST == 1) {
      ALBSOD = MIN(ALBSAT[IB - 1][ISC - 1] + INC, ALBDRY[IB - 1][ISC - 1]);
      ALBSOI = ALBSOD;
    } else if (TG > TFRZ) {
      ALBSOD = 0.06f / (POW(MAX(0.01f, COSZ), 1.7f) + 0.15f);
      ALBSOI = 0.06f;
    } else {
      ALBSOD = ALBLAK[IB - 1];
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

// noahmplsm.F90:2768-3016
static void TWOSTREAM(Ctx& c, int IB, int IC, int VEGTYP, float COSZ, float VAI, float FWET, float Tv,
                      const ABand& ALBGRD, const ABand& ALBGRI, const ABand& RHO, const ABand& TAU,
                      float FVEG, ABand& FAB, ABand& FRE, ABand& FTD, ABand& FTI, float& GDIR,
                      ABand& FREV, ABand& FREG, float& BGAP, float& WGAP) {
  const noahmp_tables& T = *c.T;
  const float PAI = 3.14159265f;
  float GAP = 0.f, KOPEN = 0.f;
  if (VAI == 0.0f) {
    GAP = 1.0f;
    KOPEN = 1.0f;
  } else {
    if (c.O.OPT_RAD == 1) {
      float rc = TV1(T.rc, VEGTYP);
      float DENFVEG = -LOG(MAX(1.0f - FVEG, 0.01f)) / (PAI * (rc * rc));
      float HD = TV1(T.hvt, VEGTYP) - TV1(T.hvb, VEGTYP);
      float BB = 0.5f * HD;
      float THETAP = ATAN(BB / rc * TAN(ACOS(MAX(0.01f, COSZ))));
      BGAP = EXP(-DENFVEG * PAI * (rc * rc) / COS(THETAP));
      float FA_ = VAI / (1.33f * PAI * POW(rc, 3.0f) * (BB / rc) * DENFVEG);
      float NEWVAI = HD * FA_;
      WGAP = (1.0f - BGAP) * EXP(-0.5f * NEWVAI / COSZ);
      GAP = MIN(1.0f - FVEG, BGAP + WGAP);
      KOPEN = 0.05f;
    }
    if (c.O.OPT_RAD == 2) { GAP = 0.0f; KOPEN = 0.0f; }
    if (c.O.OPT_RAD == 3) { GAP = 1.0f - FVEG; KOPEN = 1.0f - FVEG; }
  }
  float COSZI = MAX(0.001f, COSZ);
  float CHIL = MIN(MAX(TV1(T.xl, VEGTYP), -0.4f), 0.6f);
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
    TMP0 = (1.f - FWET) * OMEGAL + FWET * OMEGAS[IB - 1];
    TMP1 = ((1.f - FWET) * OMEGAL * BETADL + FWET * OMEGAS[IB - 1] * BETADS) / TMP0;
    TMP2 = ((1.f - FWET) * OMEGAL * BETAIL + FWET * OMEGAS[IB - 1] * BETAIS) / TMP0;
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
  float U1, U2, U3;
  if (IC == 0) {
    U1 = B - C / ALBGRD(IB); U2 = B - C * ALBGRD(IB); U3 = F + C * ALBGRD(IB);
  } else {
    U1 = B - C / ALBGRI(IB); U2 = B - C * ALBGRI(IB); U3 = F + C * ALBGRI(IB);
  }
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
  float FTDS, FTIS;
  if (IC == 0) {
    FTDS = S2 * (1.0f - GAP) + GAP;
    FTIS = (H4 * S2 / SIGMA + H5 * S1 + H6 / S1) * (1.0f - GAP);
  } else {
    FTDS = 0.f;
    FTIS = (H9 * S1 + H10 / S1) * (1.0f - KOPEN) + KOPEN;
  }
  FTD(IB) = FTDS;
  FTI(IB) = FTIS;
  float FRES, FREVEG, FREBAR;
  if (IC == 0) {
    FRES = (H1 / SIGMA + H2 + H3) * (1.0f - GAP) + ALBGRD(IB) * GAP;
    FREVEG = (H1 / SIGMA + H2 + H3) * (1.0f - GAP);
    FREBAR = ALBGRD(IB) * GAP;
  } else {
    FRES = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREVEG = (H7 + H8) * (1.0f - KOPEN) + ALBGRI(IB) * KOPEN;
    FREBAR = 0.f;
  }
  FRE(IB) = FRES;
  FREV(IB) = FREVEG;
  FREG(IB) = FREBAR;
  FAB(IB) = 1.f - FRE(IB) - (1.f - ALBGRD(IB)) * FTD(IB) - (1.f - ALBGRI(IB)) * FTI(IB);
}

// noahmplsm.F90:2243-2423
static void ALBEDO(Ctx& c, int VEGTYP, int IST, int ISC, float DT, float COSZ, float& FAGE, float ELAI,
                   float ESAI, float TG, float TV, float FSNO, float FWET, const ASoil& SMC, float SNEQVO,
                   float SNEQV, float QSNOW, float FVEG, float& ALBOLD, float& TAUSS, ABand& ALBGRD,
                   ABand& ALBGRI, ABand& ALBD, ABand& ALBI, ABand& FABD, ABand& FABI, ABand& FTDD,
                   ABand& FTID, ABand& FTII, float& FSUN, ABand& FREVI, ABand& FREVD, ABand& FREGD,
                   ABand& FREGI, float& BGAP, float& WGAP) {
  const noahmp_tables& T = *c.T;
  const float MPE = 1.E-06f;
  ABand RHO, TAU, FTDI, ALBSND, ALBSNI;
  float ALB = 0.f, GDIR = 0.f, VAI = 0.f;
  BGAP = 0.f; WGAP = 0.f;
  for (int IB = 1; IB <= 2; ++IB) {
    ALBD(IB) = 0.f; ALBI(IB) = 0.f; ALBGRD(IB) = 0.f; ALBGRI(IB) = 0.f; FABD(IB) = 0.f;
    FABI(IB) = 0.f; FTDD(IB) = 0.f; FTID(IB) = 0.f; FTII(IB) = 0.f;
    if (IB == 1) FSUN = 0.f;
  }
  // FREVI/FREVD/FREGD/FREGI are left undefined by the reference at night; they are only ever
  // multiplied by SOLAD=SOLAI=0 there (SURRAD :2540-2541) and feed diagnostics FSRV/FSRG only.
  FREVI.fill(0.f); FREVD.fill(0.f); FREGD.fill(0.f); FREGI.fill(0.f);
  if (COSZ <= 0.f) return;

  for (int IB = 1; IB <= 2; ++IB) {
    VAI = ELAI + ESAI;
    float WL = ELAI / MAX(VAI, MPE);
    float WS = ESAI / MAX(VAI, MPE);
    RHO(IB) = MAX(TV2(T.rhol, VEGTYP, IB) * WL + TV2(T.rhos, VEGTYP, IB) * WS, MPE);
    TAU(IB) = MAX(TV2(T.taul, VEGTYP, IB) * WL + TV2(T.taus, VEGTYP, IB) * WS, MPE);
  }
  SNOW_AGE(DT, TG, SNEQVO, SNEQV, TAUSS, FAGE);
  if (c.O.OPT_ALB == 1) SNOWALB_BATS(FSNO, COSZ, FAGE, ALBSND, ALBSNI);
  if (c.O.OPT_ALB == 2) {
    SNOWALB_CLASS(QSNOW, DT, ALB, ALBOLD, ALBSND, ALBSNI);
    ALBOLD = ALB;
  }
  GROUNDALB(IST, ISC, FSNO, SMC, ALBSND, ALBSNI, COSZ, TG, ALBGRD, ALBGRI);
  for (int IB = 1; IB <= 2; ++IB) {
    TWOSTREAM(c, IB, 0, VEGTYP, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, FVEG, FABD, ALBD, FTDD,
              FTID, GDIR, FREVD, FREGD, BGAP, WGAP);
    TWOSTREAM(c, IB, 1, VEGTYP, COSZ, VAI, FWET, TV, ALBGRD, ALBGRI, RHO, TAU, FVEG, FABI, ALBI, FTDI,
              FTII, GDIR, FREVI, FREGI, BGAP, WGAP);
  }
  float EXT = GDIR / COSZ * SQRT(1.f - RHO(1) - TAU(1));
  FSUN = (1.f - EXP(-EXT * VAI)) / MAX(EXT * VAI, MPE);
  EXT = FSUN;
  float WL;
  if (EXT < 0.01f) WL = 0.f; else WL = EXT;
  FSUN = WL;
}

// noahmplsm.F90:2426-2544
static void SURRAD(float MPE, float FSUN, float FSHA, float ELAI, float VAI, float LAISUN, float LAISHA,
                   const ABand& SOLAD, const ABand& SOLAI, const ABand& FABD, const ABand& FABI,
                   const ABand& FTDD, const ABand& FTID, const ABand& FTII, const ABand& ALBGRD,
                   const ABand& ALBGRI, const ABand& ALBD, const ABand& ALBI, float& PARSUN,
                   float& PARSHA, float& SAV, float& SAG, float& FSA, float& FSR, const ABand& FREVI,
                   const ABand& FREVD, const ABand& FREGD, const ABand& FREGI, float& FSRV, float& FSRG) {
  ABand CAD, CAI;
  SAG = 0.f; SAV = 0.f; FSA = 0.f;
  for (int IB = 1; IB <= 2; ++IB) {
    CAD(IB) = SOLAD(IB) * FABD(IB);
    CAI(IB) = SOLAI(IB) * FABI(IB);
    SAV = SAV + CAD(IB) + CAI(IB);
    FSA = FSA + CAD(IB) + CAI(IB);
    float TRD = SOLAD(IB) * FTDD(IB);
    float TRI = SOLAD(IB) * FTID(IB) + SOLAI(IB) * FTII(IB);
    float ABS_ = TRD * (1.f - ALBGRD(IB)) + TRI * (1.f - ALBGRI(IB));
    SAG = SAG + ABS_;
    FSA = FSA + ABS_;
  }
  float LAIFRA = ELAI / MAX(VAI, MPE);
  if (FSUN > 0.f) {
    PARSUN = (CAD(1) + FSUN * CAI(1)) * LAIFRA / MAX(LAISUN, MPE);
    PARSHA = (FSHA * CAI(1)) * LAIFRA / MAX(LAISHA, MPE);
  } else {
    PARSUN = 0.f;
    PARSHA = (CAD(1) + CAI(1)) * LAIFRA / MAX(LAISHA, MPE);
  }
  float RVIS = ALBD(1) * SOLAD(1) + ALBI(1) * SOLAI(1);
  float RNIR = ALBD(2) * SOLAD(2) + ALBI(2) * SOLAI(2);
  FSR = RVIS + RNIR;
  FSRV = FREVD(1) * SOLAD(1) + FREVI(1) * SOLAI(1) + FREVD(2) * SOLAD(2) + FREVI(2) * SOLAI(2);
  FSRG = FREGD(1) * SOLAD(1) + FREGI(1) * SOLAI(1) + FREGD(2) * SOLAD(2) + FREGI(2) * SOLAI(2);
}

// noahmplsm.F90:2120-2240
static void RADIATION(Ctx& c, int VEGTYP, int IST, int ISC, float SNEQVO, float SNEQV, float DT, float COSZ,
                      float TG, float TV, float FSNO, float QSNOW, float FWET, float ELAI, float ESAI,
                      const ASoil& SMC, const ABand& SOLAD, const ABand& SOLAI, float FVEG, float& ALBOLD,
                      float& TAUSS, float& FSUN, float& LAISUN, float& LAISHA, float& PARSUN,
                      float& PARSHA, float& SAV, float& SAG, float& FSR, float& FSA, float& FSRV,
                      float& FSRG, float& BGAP, float& WGAP) {
  const float MPE = 1.E-6f;
  float FAGE = 0.f;
  ABand ALBGRD, ALBGRI, ALBD, ALBI, FABD, FABI, FTDD, FTID, FTII, FREVI, FREVD, FREGI, FREGD;
  ALBEDO(c, VEGTYP, IST, ISC, DT, COSZ, FAGE, ELAI, ESAI, TG, TV, FSNO, FWET, SMC, SNEQVO, SNEQV, QSNOW,
         FVEG, ALBOLD, TAUSS, ALBGRD, ALBGRI, ALBD, ALBI, FABD, FABI, FTDD, FTID, FTII, FSUN, FREVI, FREVD,
         FREGD, FREGI, BGAP, WGAP);
  float FSHA = 1.f - FSUN;
  LAISUN = ELAI * FSUN;
  LAISHA = ELAI * FSHA;
  float VAI = ELAI + ESAI;
  SURRAD(MPE, FSUN, FSHA, ELAI, VAI, LAISUN, LAISHA, SOLAD, SOLAI, FABD, FABI, FTDD, FTID, FTII, ALBGRD,
         ALBGRI, ALBD, ALBI, PARSUN, PARSHA, SAV, SAG, FSA, FSR, FREVI, FREVD, FREGD, FREGI, FSRV, FSRG);
}

// noahmplsm.F90:1231-1843
void ENERGY(Ctx& c, SflxIO& s, SflxLocal& L) {
  const noahmp_tables& T = *c.T;
  const Params& P = c.P;
  const float MPE = 1.E-6f, PSIWLT = -150.f, Z0 = 0.01f;
  ASnSo FACT, DF, HCPCT;
  FACT.fill(0.f); DF.fill(0.f); HCPCT.fill(0.f);
  float TAUXV = 0.f, TAUYV = 0.f, TAUXB = 0.f, TAUYB = 0.f;
  s.IRC = 0.f; s.SHC = 0.f; s.IRG = 0.f; s.SHG = 0.f; s.EVG = 0.f; s.EVC = 0.f; s.TR = 0.f; s.GHV = 0.f;
  float PSNSUN = 0.f, PSNSHA = 0.f;
  s.T2MV = 0.f; s.Q2V = 0.f; s.CHV = 0.f; s.CHLEAF = 0.f; s.CHUC = 0.f; s.CHV2 = 0.f;

  float UR = MAX(SQRT(POW(s.UU, 2.f) + POW(s.VV, 2.f)), 1.f);  // UU**2. : a REAL exponent, libm pow (nmo.h)
  float VAI = L.ELAI + L.ESAI;
  bool VEG = false;
  if (VAI > 0.f) VEG = true;

  s.FSNO = 0.f;
  if (s.SNOWH > 0.f) {
    float BDSNO = s.SNEQV / s.SNOWH;
    float FMELT = POW(BDSNO / 100.f, M_MELT);
    s.FSNO = TANH(s.SNOWH / (2.5f * Z0 * FMELT));
  }
  float Z0MG;
  if (s.IST == 2) {
    if (s.TG <= TFRZ) Z0MG = 0.01f * (1.0f - s.FSNO) + s.FSNO * Z0SNO;
    else Z0MG = 0.01f;
  } else {
    Z0MG = Z0 * (1.0f - s.FSNO) + s.FSNO * Z0SNO;
  }
  float ZPDG = s.SNOWH, Z0M, ZPD;
  if (VEG) {
    Z0M = TV1(T.z0mvt, s.VEGTYP);
    ZPD = 0.65f * L.HTOP;
    if (s.SNOWH > ZPD) ZPD = s.SNOWH;
  } else {
    Z0M = Z0MG;
    ZPD = ZPDG;
  }
  float ZLVL = MAX(ZPD, L.HTOP) + s.ZLVL;
  if (ZPDG >= ZLVL) ZLVL = ZPDG + s.ZLVL;

  float CWP = TV1(T.cwpvt, s.VEGTYP);

  THERMOPROP(c, s.ISNOW, s.IST, L.DZSNSO, s.DT, s.SNOWH, s.SNICE, s.SNLIQ, P.CSOIL, s.SMC, s.SH2O, s.STC,
             s.VEGTYP, s.ISURBAN, DF, HCPCT, L.SNICEV, L.SNLIQV, L.EPORE, FACT);

  float FSUN, LAISUN, LAISHA, PARSUN, PARSHA;
  RADIATION(c, s.VEGTYP, s.IST, s.ISC, s.SNEQVO, s.SNEQV, s.DT, s.COSZ, s.TG, s.TV, s.FSNO, s.QSNOW,
            s.FWET, L.ELAI, L.ESAI, s.SMC, L.SOLAD, L.SOLAI, s.FVEG, s.ALBOLD, s.TAUSS, FSUN, LAISUN,
            LAISHA, PARSUN, PARSHA, s.SAV, s.SAG, s.FSR, s.FSA, L.FSRV, L.FSRG, s.BGAP, s.WGAP);

  float EMV = 1.f - EXP(-(L.ELAI + L.ESAI) / 1.0f);
  float EMG;
  if (s.ICE == 1) EMG = 0.98f * (1.f - s.FSNO) + 1.0f * s.FSNO;
  else EMG = EG[s.IST - 1] * (1.f - s.FSNO) + 1.0f * s.FSNO;

  L.BTRAN = 0.f;
  float PSI, GX = 0.f;
  if (s.IST == 1) {
    for (int IZ = 1; IZ <= P.NROOT; ++IZ) {
      if (c.O.OPT_BTR == 1) GX = (s.SH2O(IZ) - P.SMCWLT) / (P.SMCREF - P.SMCWLT);
      if (c.O.OPT_BTR == 2) {
        PSI = MAX(PSIWLT, -P.PSISAT * POW(MAX(0.01f, s.SH2O(IZ)) / P.SMCMAX, -P.BEXP));
        GX = (1.f - PSI / PSIWLT) / (1.f + P.PSISAT / PSIWLT);
      }
      if (c.O.OPT_BTR == 3) {
        PSI = MAX(PSIWLT, -P.PSISAT * POW(MAX(0.01f, s.SH2O(IZ)) / P.SMCMAX, -P.BEXP));
        GX = 1.f - EXP(-5.8f * (LOG(PSIWLT / PSI)));
      }
      GX = MIN(1.f, MAX(0.f, GX));
      L.BTRANI(IZ) = MAX(MPE, L.DZSNSO(IZ) / (-s.ZSOIL(P.NROOT)) * GX);
      L.BTRAN = L.BTRAN + L.BTRANI(IZ);
    }
    L.BTRAN = MAX(MPE, L.BTRAN);
    for (int IZ = 1; IZ <= P.NROOT; ++IZ) L.BTRANI(IZ) = L.BTRANI(IZ) / L.BTRAN;
  }

  float RSURF, RHSUR;
  if (s.IST == 2) {
    RSURF = 1.f;
    RHSUR = 1.0f;
  } else {
    float L_RSURF = (-s.ZSOIL(1)) * (EXP(POWI(1.0f - MIN(1.0f, s.SH2O(1) / P.SMCMAX), 5)) - 1.0f) /
                    (2.71828f - 1.0f);
    float D_RSURF = 2.2E-5f * P.SMCMAX * P.SMCMAX * POW(1.0f - P.SMCWLT / P.SMCMAX, 2.0f + 3.0f / P.BEXP);
    RSURF = L_RSURF / D_RSURF;
    if (s.SH2O(1) < 0.01f && s.SNOWH == 0.f) RSURF = 1.E6f;
    PSI = -P.PSISAT * POW(MAX(0.01f, s.SH2O(1)) / P.SMCMAX, -P.BEXP);
    RHSUR = s.FSNO + (1.f - s.FSNO) * EXP(PSI * GRAV / (RW * s.TG));
  }
  if (s.VEGTYP == s.ISURBAN && s.SNOWH == 0.f) RSURF = 1.E6f;

  if (s.TV > TFRZ) { L.LATHEAV = HVAP; L.FROZEN_CANOPY = false; }
  else { L.LATHEAV = HSUB; L.FROZEN_CANOPY = true; }
  float GAMMAV = CPAIR * s.SFCPRS / (0.622f * L.LATHEAV);
  if (s.TG > TFRZ) { L.LATHEAG = HVAP; L.FROZEN_GROUND = false; }
  else { L.LATHEAG = HSUB; L.FROZEN_GROUND = true; }
  float GAMMAG = CPAIR * s.SFCPRS / (0.622f * L.LATHEAG);

  float CMV = 0.f, CMB = 0.f;
  s.VEGE_ITERS = 0;
  if (VEG && s.FVEG > 0.f) {
    s.TGV = s.TG;
    CMV = s.CM;
    s.CHV = s.CH;
    VEGE_FLUX(c, s, L, s.ISNOW, s.VEGTYP, s.DT, s.SAV, s.SAG, s.LWDN, UR, s.UU, s.VV, s.SFCTMP, L.THAIR,
              L.QAIR, L.EAIR, L.RHOAIR, s.SNOWH, VAI, GAMMAV, GAMMAG, s.FWET, LAISUN, LAISHA, CWP,
              L.DZSNSO, L.HTOP, ZLVL, ZPD, Z0M, s.FVEG, Z0MG, EMV, EMG, s.CANLIQ, s.CANICE, s.STC, DF,
              s.RSSUN, s.RSSHA, RSURF, L.LATHEAV, L.LATHEAG, PARSUN, PARSHA, L.IGS, s.FOLN, s.CO2AIR,
              s.O2AIR, L.BTRAN, s.SFCPRS, RHSUR, s.Q2, s.EAH, s.TAH, s.TV, s.TGV, CMV, s.CHV, s.DX,
              s.DZ8W, TAUXV, TAUYV, s.IRG, s.IRC, s.SHG, s.SHC, s.EVG, s.EVC, s.TR, s.GHV, s.T2MV, PSNSUN,
              PSNSHA, s.QSFC, s.PSFC, s.ISURBAN, s.IZ0TLND, s.Q2V, s.CHV2, s.CHLEAF, s.CHUC);
  }

  s.TGB = s.TG;
  CMB = s.CM;
  s.CHB = s.CH;
  BARE_FLUX(c, s, s.ISNOW, s.DT, s.SAG, s.LWDN, UR, s.UU, s.VV, s.SFCTMP, L.THAIR, L.QAIR, L.EAIR,
            L.RHOAIR, s.SNOWH, L.DZSNSO, ZLVL, ZPDG, Z0MG, EMG, s.STC, DF, RSURF, L.LATHEAG, GAMMAG, RHSUR,
            s.Q2, s.TGB, CMB, s.CHB, TAUXB, TAUYB, s.IRB, s.SHB, s.EVB, s.GHB, s.T2MB, s.DX, s.DZ8W,
            s.VEGTYP, s.QSFC, s.PSFC, s.ISURBAN, s.IZ0TLND, s.SFCPRS, s.Q2B, s.CHB2);

  if (VEG && s.FVEG > 0.f) {
    L.TAUX = s.FVEG * TAUXV + (1.0f - s.FVEG) * TAUXB;
    L.TAUY = s.FVEG * TAUYV + (1.0f - s.FVEG) * TAUYB;
    s.FIRA = s.FVEG * s.IRG + (1.0f - s.FVEG) * s.IRB + s.IRC;
    s.FSH = s.FVEG * s.SHG + (1.0f - s.FVEG) * s.SHB + s.SHC;
    s.FGEV = s.FVEG * s.EVG + (1.0f - s.FVEG) * s.EVB;
    s.SSOIL = s.FVEG * s.GHV + (1.0f - s.FVEG) * s.GHB;
    s.FCEV = s.EVC;
    s.FCTR = s.TR;
    s.TG = s.FVEG * s.TGV + (1.0f - s.FVEG) * s.TGB;
    L.T2M = s.FVEG * s.T2MV + (1.0f - s.FVEG) * s.T2MB;
    L.TS = s.FVEG * s.TV + (1.0f - s.FVEG) * s.TGB;
    s.CM = s.FVEG * CMV + (1.0f - s.FVEG) * CMB;
    s.CH = s.FVEG * s.CHV + (1.0f - s.FVEG) * s.CHB;
    L.Q1 = s.FVEG * (s.EAH * 0.622f / (s.SFCPRS - 0.378f * s.EAH)) + (1.0f - s.FVEG) * s.QSFC;
    L.Q2E = s.FVEG * s.Q2V + (1.0f - s.FVEG) * s.Q2B;
  } else {
    L.TAUX = TAUXB;
    L.TAUY = TAUYB;
    s.FIRA = s.IRB;
    s.FSH = s.SHB;
    s.FGEV = s.EVB;
    s.SSOIL = s.GHB;
    s.TG = s.TGB;
    L.T2M = s.T2MB;
    s.FCEV = 0.f;
    s.FCTR = 0.f;
    L.TS = s.TG;
    s.CM = CMB;
    s.CH = s.CHB;
    L.Q1 = s.QSFC;
    L.Q2E = s.Q2B;
    s.RSSUN = 0.0f;
    s.RSSHA = 0.0f;
    s.TGV = s.TGB;
    s.CHV = s.CHB;
  }

  float FIRE = s.LWDN + s.FIRA;
  if (FIRE <= 0.f) c.fatal(NOAHMP_ERR_FIRE, FIRE);

  s.EMISSI = s.FVEG * (EMG * (1.f - EMV) + EMV + EMV * (1.f - EMV) * (1.f - EMG)) + (1.f - s.FVEG) * EMG;
  s.TRAD = POW((FIRE - (1.f - s.EMISSI) * s.LWDN) / (s.EMISSI * SB), 0.25f);
  s.APAR = PARSUN * LAISUN + PARSHA * LAISHA;
  s.PSN = PSNSUN * LAISUN + PSNSHA * LAISHA;

  TSNOSOI(c, s.ICE, s.ISNOW, s.IST, s.TBOT, s.ZSNSO, s.SSOIL, DF, HCPCT, P.ZBOT, s.SAG, s.DT, s.SNOWH,
          L.DZSNSO, s.TG, s.STC);

  if (c.O.OPT_STC == 2) {
    if (s.SNOWH > 0.05f && s.TG > TFRZ) {
      s.TGV = TFRZ;
      s.TGB = TFRZ;
      if (VEG && s.FVEG > 0.f) {
        s.TG = s.FVEG * s.TGV + (1.0f - s.FVEG) * s.TGB;
        L.TS = s.FVEG * s.TV + (1.0f - s.FVEG) * s.TGB;
      } else {
        s.TG = s.TGB;
        L.TS = s.TGB;
      }
    }
  }

  PHASECHANGE(c, s.ISNOW, s.DT, FACT, L.DZSNSO, HCPCT, s.IST, s.STC, s.SNICE, s.SNLIQ, s.SNEQV, s.SNOWH,
              s.SMC, s.SH2O, L.QMELT, L.IMELT, s.PONDING);
}

}  // namespace nmo

// TDFCND probe: soil thermal conductivity for given total / liquid water, porosity and quartz content
extern "C" float nmo_tdfcnd(float SMC, float SH2O, float SMCMAX, float QUARTZ) {
  nmo::Ctx c{};
  c.P.SMCMAX = SMCMAX;
  c.P.QUARTZ = QUARTZ;
  float DF = 0.f;
  nmo::TDFCND(c, DF, SMC, SH2O);
  return DF;
}

// ---- probe for the known-answer test of the two-stream solution (tests/test_oracle.py) -------------------------
extern "C" void nmo_twostream(const noahmp_tables* tables, int opt_rad, int IB, int IC, int VEGTYP, float COSZ, float VAI,
                              float FWET, float Tv, float ALBGRD, float ALBGRI, float RHO, float TAU, float FVEG,
                              float* out /* FAB FRE FTD FTI GDIR */) {
  using namespace nmo;
  Ctx c{};
  c.T = tables;
  c.O.OPT_RAD = opt_rad;
  ABand agd, agi, rho, tau, FAB, FRE, FTD, FTI, FREV, FREG;
  for (int b = 1; b <= 2; ++b) { agd(b) = ALBGRD; agi(b) = ALBGRI; rho(b) = RHO; tau(b) = TAU; }
  float GDIR = 0.f, BGAP = 0.f, WGAP = 0.f;
  TWOSTREAM(c, IB, IC, VEGTYP, COSZ, VAI, FWET, Tv, agd, agi, rho, tau, FVEG, FAB, FRE, FTD, FTI, GDIR, FREV, FREG, BGAP,
            WGAP);
  out[0] = FAB(IB); out[1] = FRE(IB); out[2] = FTD(IB); out[3] = FTI(IB); out[4] = GDIR;
}

