// nmo_land2.cpp — ORACLE (test infrastructure): VEGE_FLUX, BARE_FLUX, RAGRB, SFCDIF1, SFCDIF2, ESAT,
// STOMATA, CANRES/CALHUM, TSNOSOI/HRT/HSTEP/ROSR12, PHASECHANGE/FRH2O.
// Restates phys/module_sf_noahmplsm.F90:3018-4422 and :5272-6377 (SFCDIF3/4 at :4425-5269 are
// non-functional offline and excluded).
#include "nmo_land.h"

namespace nmo {

static inline float TDC(float T) { return MIN(50.f, MAX(-50.f, (T - TFRZ))); }

// noahmplsm.F90:5272-5321
void ESAT(float T, float& ESW, float& ESI, float& DESW, float& DESI) {
  const float A0 = 6.107799961f, A1 = 4.436518521E-01f, A2 = 1.428945805E-02f, A3 = 2.650648471E-04f,
              A4 = 3.031240396E-06f, A5 = 2.034080948E-08f, A6 = 6.136820929E-11f;
  const float B0 = 6.109177956f, B1 = 5.034698970E-01f, B2 = 1.886013408E-02f, B3 = 4.176223716E-04f,
              B4 = 5.824720280E-06f, B5 = 4.838803174E-08f, B6 = 1.838826904E-10f;
  const float C0 = 4.438099984E-01f, C1 = 2.857002636E-02f, C2 = 7.938054040E-04f, C3 = 1.215215065E-05f,
              C4 = 1.036561403E-07f, C5 = 3.532421810e-10f, C6 = -7.090244804E-13f;
  const float D0 = 5.030305237E-01f, D1 = 3.773255020E-02f, D2 = 1.267995369E-03f, D3 = 2.477563108E-05f,
              D4 = 3.005693132E-07f, D5 = 2.158542548E-09f, D6 = 7.131097725E-12f;
  ESW = 100.f * (A0 + T * (A1 + T * (A2 + T * (A3 + T * (A4 + T * (A5 + T * A6))))));
  ESI = 100.f * (B0 + T * (B1 + T * (B2 + T * (B3 + T * (B4 + T * (B5 + T * B6))))));
  DESW = 100.f * (C0 + T * (C1 + T * (C2 + T * (C3 + T * (C4 + T * (C5 + T * C6))))));
  DESI = 100.f * (D0 + T * (D1 + T * (D2 + T * (D3 + T * (D4 + T * (D5 + T * D6))))));
}

// noahmplsm.F90:4061-4220
void SFCDIF1(Ctx& c, int ITER, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL, float ZPD,
                    float Z0M, float Z0H, float UR, float MPE, float& MOZ, int& MOZSGN, float& FM,
                    float& FH, float& FM2, float& FH2, float& CM, float& CH, float& FV, float& CH2) {
  float MOZOLD = MOZ;
  if (ZLVL <= ZPD) c.fatal(NOAHMP_ERR_ZLVL, ZLVL - ZPD);
  float TMPCM = LOG((ZLVL - ZPD) / Z0M);
  float TMPCH = LOG((ZLVL - ZPD) / Z0H);
  float TMPCM2 = LOG((2.0f + Z0M) / Z0M);
  float TMPCH2 = LOG((2.0f + Z0H) / Z0H);
  float MOL, MOZ2;
  if (ITER == 1) {
    FV = 0.0f; MOZ = 0.0f; MOL = 0.0f; MOZ2 = 0.0f;
  } else {
    float TVIR = (1.f + 0.61f * QAIR) * SFCTMP;
    float TMP1 = VKC * (GRAV / TVIR) * H / (RHOAIR * CPAIR);
    if (ABS(TMP1) <= MPE) TMP1 = MPE;
    MOL = -1.f * POWI(FV, 3) / TMP1;
    MOZ = MIN((ZLVL - ZPD) / MOL, 1.f);
    MOZ2 = MIN((2.0f + Z0H) / MOL, 1.f);
  }
  if (MOZOLD * MOZ < 0.f) MOZSGN = MOZSGN + 1;
  if (MOZSGN >= 2) {
    MOZ = 0.f; FM = 0.f; FH = 0.f; MOZ2 = 0.f; FM2 = 0.f; FH2 = 0.f;
  }
  float FMNEW, FHNEW, FM2NEW, FH2NEW;
  if (MOZ < 0.f) {
    float TMP1 = POW(1.f - 16.f * MOZ, 0.25f);
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
    FMNEW = -5.f * MOZ;
    FHNEW = FMNEW;
    FM2NEW = -5.f * MOZ2;
    FH2NEW = FM2NEW;
  }
  if (ITER == 1) {
    FM = FMNEW; FH = FHNEW; FM2 = FM2NEW; FH2 = FH2NEW;
  } else {
    FM = 0.5f * (FM + FMNEW);
    FH = 0.5f * (FH + FHNEW);
    FM2 = 0.5f * (FM2 + FM2NEW);
    FH2 = 0.5f * (FH2 + FH2NEW);
  }
  FH = MIN(FH, 0.9f * TMPCH);
  FM = MIN(FM, 0.9f * TMPCM);
  FH2 = MIN(FH2, 0.9f * TMPCH2);
  FM2 = MIN(FM2, 0.9f * TMPCM2);
  float CMFM = TMPCM - FM;
  float CHFH = TMPCH - FH;
  float CM2FM2 = TMPCM2 - FM2;
  float CH2FH2 = TMPCH2 - FH2;
  if (ABS(CMFM) <= MPE) CMFM = MPE;
  if (ABS(CHFH) <= MPE) CHFH = MPE;
  if (ABS(CM2FM2) <= MPE) CM2FM2 = MPE;
  if (ABS(CH2FH2) <= MPE) CH2FH2 = MPE;
  CM = VKC * VKC / (CMFM * CMFM);
  CH = VKC * VKC / (CMFM * CHFH);
  CH2 = VKC * VKC / (CM2FM2 * CH2FH2);
  FV = UR * SQRT(CM);
  CH2 = VKC * FV / CH2FH2;
}

// noahmplsm.F90:4224-4422
static void SFCDIF2(int ITER, float Z0, float THZ0, float THLM, float SFCSPD, float CZIL, float ZLM,
                    float& AKMS, float& AKHS, float& RLMO, float& WSTAR2, float& USTAR) {
  const float WWST = 1.2f, WWST2 = WWST * WWST, VKRM = 0.40f, EXCM = 0.001f, BETA = 1.0f / 270.0f,
              BTG = BETA * GRAV, ELFC = VKRM * BTG, WOLD = 0.15f, WNEW = 1.0f - WOLD,
              PIHF = 3.14159265f / 2.f, EPSU2 = 1.E-4f, EPSUST = 0.07f, ZTMIN = -5.0f, ZTMAX = 1.0f,
              HPBL = 1000.0f, SQVISC = 258.2f, RIC = 0.183f, RRIC = 1.0f / RIC, FHNEU = 0.8f, RFC = 0.191f,
              RFAC = RIC / (FHNEU * RFC * RFC);
  auto PSLMU = [&](float ZZ) { return -0.96f * LOG(1.0f - 4.5f * ZZ); };
  auto PSLMS = [&](float ZZ) { return ZZ * RRIC - 2.076f * (1.f - 1.f / (ZZ + 1.f)); };
  auto PSLHU = [&](float ZZ) { return -0.96f * LOG(1.0f - 4.5f * ZZ); };
  auto PSLHS = [&](float ZZ) { return ZZ * RFAC - 2.076f * (1.f - 1.f / (ZZ + 1.f)); };
  auto PSPMU = [&](float XX) {
    return -2.f * LOG((XX + 1.f) * 0.5f) - LOG((XX * XX + 1.f) * 0.5f) + 2.f * ATAN(XX) - PIHF;
  };
  auto PSPMS = [&](float YY) { return 5.f * YY; };
  auto PSPHU = [&](float XX) { return -2.f * LOG((XX * XX + 1.f) * 0.5f); };
  auto PSPHS = [&](float YY) { return 5.f * YY; };

  const int ILECH = 0;
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
    RLMO = ELFC * AKHS * DTHV / POWI(USTAR, 3);
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
  if (ILECH == 0) {
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
      PSMZ = PSPMS(ZETAU);
      SIMM = PSPMS(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSPHS(ZETAT);
      SIMH = PSPHS(ZETALT) - PSHZ + RLOGT;
    }
  } else {
    if (RLMO < 0.f) {
      PSMZ = PSLMU(ZETAU);
      SIMM = PSLMU(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSLHU(ZETAT);
      SIMH = PSLHU(ZETALT) - PSHZ + RLOGT;
    } else {
      ZETALU = MIN(ZETALU, ZTMAX);
      ZETALT = MIN(ZETALT, ZTMAX);
      PSMZ = PSLMS(ZETAU);
      SIMM = PSLMS(ZETALU) - PSMZ + RLOGU;
      PSHZ = PSLHS(ZETAT);
      SIMH = PSLHS(ZETALT) - PSHZ + RLOGT;
    }
  }
  USTAR = MAX(SQRT(AKMS * SQRT(DU2 + WSTAR2)), EPSUST);
  ZT = MAX(1.E-6f, EXP(ZILFC * SQRT(USTAR * Z0)) * Z0);
  ZSLT = ZLM + ZT;
  RLOGT = LOG(ZSLT / ZT);
  float USTARK = USTAR * VKRM;
  AKMS = MAX(USTARK / SIMM, CXCH);
  AKHS = MAX(USTARK / SIMH, CXCH);
  if (BTGH * AKHS * DTHV != 0.0f) WSTAR2 = WWST2 * POW(ABS(BTGH * AKHS * DTHV), 2.f / 3.f);
  else WSTAR2 = 0.0f;
  float RLMN = ELFC * AKHS * DTHV / POWI(USTAR, 3);
  float RLMA = RLMO * WOLD + RLMN * WNEW;
  RLMO = RLMA;
  (void)RLOGT;
}

// noahmplsm.F90:3960-4057
static void RAGRB(Ctx& c, int ITER, float VAI, float RHOAIR, float HG, float TAH, float ZPD, float Z0MG,
                  float Z0HG, float HCAN, float UC, float Z0H, float FV, float CWP, int VEGTYP, float MPE,
                  float& MOZG, float& FHG, float& RAMG, float& RAHG, float& RAWG, float& RB) {
  MOZG = 0.f;
  float MOLG = 0.f;
  if (ITER > 1) {
    float TMP1 = VKC * (GRAV / TAH) * HG / (RHOAIR * CPAIR);
    if (ABS(TMP1) <= MPE) TMP1 = MPE;
    MOLG = -1.f * POWI(FV, 3) / TMP1;
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
  RAMG = 0.f;
  RAHG = TMPRAH2 / KH;
  RAWG = RAHG;
  float TMPRB = CWPC * 50.f / (1.f - EXP(-CWPC / 2.f));
  RB = TMPRB * SQRT(TV1(c.T->dleaf, VEGTYP) / UC);
}

// noahmplsm.F90:5323-5464 (incl. the internal CI2CI)
static void STOMATA(Ctx& c, int VEGTYP, float MPE, float APAR, float FOLN, float TV, float EI, float EA,
                    float SFCTMP, float SFCPRS, float O2, float CO2, float IGS, float BTRAN, float RB,
                    float& RS, float& PSN) {
  const noahmp_tables& T = *c.T;
  const float CIERR = 5e-2f;
  const int NITER = 20;
  float CF = SFCPRS / (8.314f * SFCTMP) * 1.0e06f;
  RS = 1.0f / TV1(T.bp, VEGTYP) * CF;
  PSN = 0.0f;
  float CI = CO2;
  if (APAR <= 0.0f) return;
  float FNF = MIN(FOLN / MAX(MPE, TV1(T.folnmx, VEGTYP)), 1.0f);
  float TC = TV - TFRZ;
  float PPF = 4.6f * APAR;
  float J = PPF * TV1(T.qe25, VEGTYP);
  float KC = TV1(T.kc25, VEGTYP) * POW(TV1(T.akc, VEGTYP), (TC - 25.0f) / 10.0f);
  float KO = TV1(T.ko25, VEGTYP) * POW(TV1(T.ako, VEGTYP), (TC - 25.0f) / 10.0f);
  float AWC = KC * (1.0f + O2 / KO);
  float CP = 0.5f * KC / KO * O2 * 0.21f;
  float VCMX = TV1(T.vcmx25, VEGTYP) /
               (1.0f + EXP((-2.2E05f + 710.0f * (TC + TFRZ)) / (8.314f * (TC + TFRZ)))) * FNF * BTRAN *
               POW(TV1(T.avcmx, VEGTYP), (TC - 25.0f) / 10.0f);
  float RLB = RB / CF;
  const float c3 = TV1(T.c3psn, VEGTYP), mp = TV1(T.mp, VEGTYP), bp = TV1(T.bp, VEGTYP);
  auto CI2CI = [&](float CIx, float& FCI) {
    float WJ = MAX(CIx - CP, 0.0f) * J / (CIx + 2.0f * CP) * c3 + J * (1.f - c3);
    float WC = MAX(CIx - CP, 0.0f) * VCMX / (CIx + AWC) * c3 + VCMX * (1.f - c3);
    float WE = 0.5f * VCMX * c3 + 4000.0f * VCMX * CIx / SFCPRS * (1.f - c3);
    PSN = MIN3(WJ, WC, WE) * IGS;
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
    FCI = MAX(CS - PSN * SFCPRS * 1.65f * RS, 0.0f);
  };
  float CIHI = 1.5f * CO2;
  float CILOW = 0.0f;
  for (int ITER = 1; ITER <= NITER; ++ITER) {
    CI = 0.5f * (CIHI + CILOW);
    float FCI;
    CI2CI(CI, FCI);
    if (((CIHI - CILOW) <= CIERR) || ABS(FCI - CI) <= MPE) break;
    else if (FCI > CI) CILOW = CI;
    else CIHI = CI;
  }
  RS = RS * CF;
}

// noahmplsm.F90:5679-5705
static void CALHUM(float SFCTMP, float SFCPRS, float& Q2SAT, float& DQSDT2) {
  const float A2 = 17.67f, A3 = 273.15f, A4 = 29.65f, ELWV = 2.501E6f, A23M4 = A2 * (A3 - A4), E0 = 0.611f,
              RV = 461.0f, EPSILON = 0.622f;
  float ES = E0 * EXP(ELWV / RV * (1.f / A3 - 1.f / SFCTMP));
  float SFCPRSX = SFCPRS * 1.E-3f;
  Q2SAT = EPSILON * ES / (SFCPRSX - ES);
  Q2SAT = Q2SAT * 1.E3f;
  DQSDT2 = (Q2SAT / (1.f + Q2SAT)) * A23M4 / ((SFCTMP - A4) * (SFCTMP - A4));
  Q2SAT = Q2SAT / 1.E3f;
}

// noahmplsm.F90:5598-5677
static void CANRES(Ctx& c, float PAR, float SFCTMP, float RCSOIL, float EAH, float SFCPRS, float& RC,
                   float& PSN) {
  const Params& P = c.P;
  RC = 0.0f;
  float RCS = 0.0f, RCT = 0.0f, RCQ = 0.0f, Q2SAT, DQSDT2;
  float Q2 = 0.622f * EAH / (SFCPRS - 0.378f * EAH);
  Q2 = Q2 / (1.0f + Q2);
  CALHUM(SFCTMP, SFCPRS, Q2SAT, DQSDT2);
  float FF = 2.0f * PAR / P.RGL;
  RCS = (FF + P.RSMIN / P.RSMAX) / (1.0f + FF);
  RCS = MAX(RCS, 0.0001f);
  RCT = 1.0f - 0.0016f * POW(P.TOPT - SFCTMP, 2.0f);
  RCT = MAX(RCT, 0.0001f);
  RCQ = 1.0f / (1.0f + P.HS * MAX(0.f, Q2SAT - Q2));
  RCQ = MAX(RCQ, 0.01f);
  RC = P.RSMIN / (RCS * RCT * RCQ * RCSOIL);
  PSN = -999.99f;
}

// noahmplsm.F90:3018-3589
void VEGE_FLUX(Ctx& c, SflxIO& s, SflxLocal& L, int ISNOW, int VEGTYP, float DT, float SAV, float SAG,
               float LWDN, float UR, float UU, float VV, float SFCTMP, float THAIR, float QAIR, float EAIR,
               float RHOAIR, float SNOWH, float VAI, float GAMMAV, float GAMMAG, float FWET, float LAISUN,
               float LAISHA, float CWP, const ASnSo& DZSNSO, float HTOP, float ZLVL, float ZPD, float Z0M,
               float FVEG, float Z0MG, float EMV, float EMG, float CANLIQ, float CANICE, const ASnSo& STC,
               const ASnSo& DF, float& RSSUN, float& RSSHA, float RSURF, float LATHEAV, float LATHEAG,
               float PARSUN, float PARSHA, float IGS, float FOLN, float CO2AIR, float O2AIR, float BTRAN,
               float SFCPRS, float RHSUR, float Q2, float& EAH, float& TAH, float& TV, float& TG, float& CM,
               float& CH, float DX, float DZ8W, float& TAUXV, float& TAUYV, float& IRG, float& IRC,
               float& SHG, float& SHC, float& EVG, float& EVC, float& TR, float& GH, float& T2MV,
               float& PSNSUN, float& PSNSHA, float& QSFC, float PSFC, int ISURBAN, int IZ0TLND, float& Q2V,
               float& CAH2, float& CHLEAF, float& CHUC) {
  (void)L; (void)Q2; (void)DX; (void)DZ8W; (void)ISURBAN; (void)IZ0TLND; (void)LATHEAG; (void)GAMMAG;
  const int NITERC = 20, NITERG = 5;
  const float MPE = 1E-6f;
  int LITER = 0;
  float FV = 0.1f;
  float DTV = 0.f, DTG = 0.f;
  int MOZSGN = 0;
  float HG = 0.f, H = 0.f;
  float MOZ = 0.f, FM = 0.f, FH = 0.f, FM2 = 0.f, FH2 = 0.f, CH2 = 0.f, WSTAR = 0.f;
  float MOZG = 0.f, FHG = 0.f, RAMG = 0.f, RAHG = 0.f, RAWG = 0.f, RB = 0.f;
  float ESATW, ESATI, DSATW, DSATI, ESTV = 0.f, DESTV = 0.f, ESTG, DESTG = 0.f;
  float CAH = 0.f, CVH = 0.f, CGH, COND, ATA, BTA, CSH, CAW, CEW, CTW, CGW, AEA, BEA, CEV, CTR, A, B;
  float RAMC, RAHC = 1.f, RAWC;

  float VAIE = MIN(6.f, VAI / FVEG);
  float LAISUNE = MIN(6.f, LAISUN / FVEG);
  float LAISHAE = MIN(6.f, LAISHA / FVEG);

  float T = TDC(TG);
  ESAT(T, ESATW, ESATI, DSATW, DSATI);
  if (T > 0.f) ESTG = ESATW; else ESTG = ESATI;

  QSFC = 0.622f * EAIR / (PSFC - 0.378f * EAIR);

  float HCAN = HTOP;
  float UC = UR * LOG(HCAN / Z0M) / LOG(ZLVL / Z0M);
  if ((HCAN - ZPD) <= 0.f) c.fatal(NOAHMP_ERR_HCAN, HCAN - ZPD);

  float AIR = -EMV * (1.f + (1.f - EMV) * (1.f - EMG)) * LWDN - EMV * EMG * SB * POWI(TG, 4);
  float CIR = (2.f - EMV * (1.f - EMG)) * EMV * SB;

  int ITER;
  for (ITER = 1; ITER <= NITERC; ++ITER) {
    float Z0H = Z0M;
    float Z0HG = Z0MG;
    if (c.O.OPT_SFC == 1) {
      SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE, MOZ, MOZSGN, FM, FH, FM2, FH2,
              CM, CH, FV, CH2);
    }
    if (c.O.OPT_SFC == 2) {
      SFCDIF2(ITER, Z0M, TAH, THAIR, UR, c.P.CZIL, ZLVL, CM, CH, MOZ, WSTAR, FV);
      CH = CH / UR;
      CM = CM / UR;
    }
    RAMC = MAX(1.f, 1.f / (CM * UR));
    RAHC = MAX(1.f, 1.f / (CH * UR));
    RAWC = RAHC;
    (void)RAMC;

    RAGRB(c, ITER, VAIE, RHOAIR, HG, TAH, ZPD, Z0MG, Z0HG, HCAN, UC, Z0H, FV, CWP, VEGTYP, MPE, MOZG, FHG,
          RAMG, RAHG, RAWG, RB);

    T = TDC(TV);
    ESAT(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.f) { ESTV = ESATW; DESTV = DSATW; }
    else { ESTV = ESATI; DESTV = DSATI; }

    if (ITER == 1) {
      if (c.O.OPT_CRS == 1) {
        STOMATA(c, VEGTYP, MPE, PARSUN, FOLN, TV, ESTV, EAH, SFCTMP, SFCPRS, O2AIR, CO2AIR, IGS, BTRAN, RB,
                RSSUN, PSNSUN);
        STOMATA(c, VEGTYP, MPE, PARSHA, FOLN, TV, ESTV, EAH, SFCTMP, SFCPRS, O2AIR, CO2AIR, IGS, BTRAN, RB,
                RSSHA, PSNSHA);
      }
      if (c.O.OPT_CRS == 2) {
        CANRES(c, PARSUN, TV, BTRAN, EAH, SFCPRS, RSSUN, PSNSUN);
        CANRES(c, PARSHA, TV, BTRAN, EAH, SFCPRS, RSSHA, PSNSHA);
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
    CGW = 1.f / (RAWG + RSURF);
    COND = CAW + CEW + CTW + CGW;
    AEA = (EAIR * CAW + ESTG * CGW) / COND;
    BEA = (CEW + CTW) / COND;
    CEV = (1.f - BEA) * CEW * RHOAIR * CPAIR / GAMMAV;
    CTR = (1.f - BEA) * CTW * RHOAIR * CPAIR / GAMMAV;

    TAH = ATA + BTA * TV;
    EAH = AEA + BEA * ESTV;

    IRC = FVEG * (AIR + CIR * POWI(TV, 4));
    SHC = FVEG * RHOAIR * CPAIR * CVH * (TV - TAH);
    EVC = FVEG * RHOAIR * CPAIR * CEW * (ESTV - EAH) / GAMMAV;
    TR = FVEG * RHOAIR * CPAIR * CTW * (ESTV - EAH) / GAMMAV;
    if (TV > TFRZ) EVC = MIN(CANLIQ * LATHEAV / DT, EVC);
    else EVC = MIN(CANICE * LATHEAV / DT, EVC);

    B = SAV - IRC - SHC - EVC - TR;
    A = FVEG * (4.f * CIR * POWI(TV, 3) + CSH + (CEV + CTR) * DESTV);
    DTV = B / A;

    IRC = IRC + FVEG * 4.f * CIR * POWI(TV, 3) * DTV;
    SHC = SHC + FVEG * CSH * DTV;
    EVC = EVC + FVEG * CEV * DESTV * DTV;
    TR = TR + FVEG * CTR * DESTV * DTV;

    TV = TV + DTV;

    H = RHOAIR * CPAIR * (TAH - SFCTMP) / RAHC;
    HG = RHOAIR * CPAIR * (TG - TAH) / RAHG;

    QSFC = (0.622f * EAH) / (SFCPRS - 0.378f * EAH);

    s.VEGE_ITERS = ITER;
    if (LITER == 1) break;
    if (ITER >= 5 && ABS(DTV) <= 0.01f && LITER == 0) LITER = 1;
  }

  AIR = -EMG * (1.f - EMV) * LWDN - EMG * EMV * SB * POWI(TV, 4);
  CIR = EMG * SB;
  CSH = RHOAIR * CPAIR / RAHG;
  CEV = RHOAIR * CPAIR / (GAMMAG * (RAWG + RSURF));
  CGH = 2.f * DF(ISNOW + 1) / DZSNSO(ISNOW + 1);

  for (ITER = 1; ITER <= NITERG; ++ITER) {
    T = TDC(TG);
    ESAT(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.f) { ESTG = ESATW; DESTG = DSATW; }
    else { ESTG = ESATI; DESTG = DSATI; }

    IRG = CIR * POWI(TG, 4) + AIR;
    SHG = CSH * (TG - TAH);
    EVG = CEV * (ESTG * RHSUR - EAH);
    GH = CGH * (TG - STC(ISNOW + 1));

    B = SAG - IRG - SHG - EVG - GH;
    A = 4.f * CIR * POWI(TG, 3) + CSH + CEV * DESTG + CGH;
    DTG = B / A;

    IRG = IRG + 4.f * CIR * POWI(TG, 3) * DTG;
    SHG = SHG + CSH * DTG;
    EVG = EVG + CEV * DESTG * DTG;
    GH = GH + CGH * DTG;
    TG = TG + DTG;
  }

  if (c.O.OPT_STC == 1) {
    if (SNOWH > 0.05f && TG > TFRZ) {
      TG = TFRZ;
      IRG = CIR * POWI(TG, 4) - EMG * (1.f - EMV) * LWDN - EMG * EMV * SB * POWI(TV, 4);
      SHG = CSH * (TG - TAH);
      EVG = CEV * (ESTG * RHSUR - EAH);
      GH = SAG - (IRG + SHG + EVG);
    }
  }

  TAUXV = -RHOAIR * CM * UR * UU;
  TAUYV = -RHOAIR * CM * UR * VV;

  if (c.O.OPT_SFC == 1 || c.O.OPT_SFC == 2) {
    float Z0H = Z0M;
    // FH2 is only ever assigned by SFCDIF1; with OPT_SFC=2 the reference reads it undefined
    //  — the oracle defines it as 0 there.
    CAH2 = FV * VKC / LOG((2.f + Z0H) / Z0H);
    CAH2 = FV * VKC / (LOG((2.f + Z0H) / Z0H) - FH2);
    float CQ2V = CAH2;
    if (CAH2 < 1.E-5f) {
      T2MV = TAH;
      Q2V = QSFC;
    } else {
      T2MV = TAH - (SHG + SHC / FVEG) / (RHOAIR * CPAIR) * 1.f / CAH2;
      Q2V = QSFC - ((EVC + TR) / FVEG + EVG) / (LATHEAV * RHOAIR) * 1.f / CQ2V;
    }
  }
  CH = CAH;
  CHLEAF = CVH;
  CHUC = 1.f / RAHG;
}

// noahmplsm.F90:3591-3958
void BARE_FLUX(Ctx& c, SflxIO& s, int ISNOW, float DT, float SAG, float LWDN, float UR, float UU, float VV,
               float SFCTMP, float THAIR, float QAIR, float EAIR, float RHOAIR, float SNOWH,
               const ASnSo& DZSNSO, float ZLVL, float ZPD, float Z0M, float EMG, const ASnSo& STC,
               const ASnSo& DF, float RSURF, float LATHEA, float GAMMA, float RHSUR, float Q2, float& TGB,
               float& CM, float& CH, float& TAUXB, float& TAUYB, float& IRB, float& SHB, float& EVB,
               float& GHB, float& T2MB, float DX, float DZ8W, int IVGTYP, float& QSFC, float PSFC,
               int ISURBAN, int IZ0TLND, float SFCPRS, float& Q2B, float& EHB2) {
  (void)s; (void)DT; (void)Q2; (void)DX; (void)DZ8W; (void)IZ0TLND; (void)SFCPRS;
  const int NITERB = 5;
  const float MPE = 1E-6f;
  float DTG = 0.f;
  int MOZSGN = 0;
  float H = 0.f, QFX = 0.f, FV = 0.1f;
  float MOZ = 0.f, FM = 0.f, FH = 0.f, FM2 = 0.f, FH2 = 0.f, CH2 = 0.f, WSTAR = 0.f;
  float ESATW, ESATI, DSATW, DSATI, ESTG = 0.f, DESTG;
  float CSH = 0.f, CEV = 0.f, A, B, EHB = 0.f, Z0H = Z0M;
  (void)QFX; (void)DTG;

  float CIR = EMG * SB;
  float CGH = 2.f * DF(ISNOW + 1) / DZSNSO(ISNOW + 1);

  for (int ITER = 1; ITER <= NITERB; ++ITER) {
    Z0H = Z0M;
    if (c.O.OPT_SFC == 1) {
      SFCDIF1(c, ITER, SFCTMP, RHOAIR, H, QAIR, ZLVL, ZPD, Z0M, Z0H, UR, MPE, MOZ, MOZSGN, FM, FH, FM2, FH2,
              CM, CH, FV, CH2);
    }
    if (c.O.OPT_SFC == 2) {
      SFCDIF2(ITER, Z0M, TGB, THAIR, UR, c.P.CZIL, ZLVL, CM, CH, MOZ, WSTAR, FV);
      CH = CH / UR;
      CM = CM / UR;
      if (SNOWH > 0.f) {
        CM = MIN(0.01f, CM);
        CH = MIN(0.01f, CH);
      }
    }
    float RAMB = MAX(1.f, 1.f / (CM * UR));
    float RAHB = MAX(1.f, 1.f / (CH * UR));
    float RAWB = RAHB;
    float EMB = 1.f / RAMB;
    EHB = 1.f / RAHB;
    (void)EMB;

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

    B = SAG - IRB - SHB - EVB - GHB;
    A = 4.f * CIR * POWI(TGB, 3) + CSH + CEV * DESTG + CGH;
    DTG = B / A;

    IRB = IRB + 4.f * CIR * POWI(TGB, 3) * DTG;
    SHB = SHB + CSH * DTG;
    EVB = EVB + CEV * DESTG * DTG;
    GHB = GHB + CGH * DTG;

    TGB = TGB + DTG;

    H = CSH * (TGB - SFCTMP);

    T = TDC(TGB);
    ESAT(T, ESATW, ESATI, DSATW, DSATI);
    if (T > 0.f) ESTG = ESATW; else ESTG = ESATI;
    QSFC = 0.622f * (ESTG * RHSUR) / (PSFC - 0.378f * (ESTG * RHSUR));
    QFX = (QSFC - QAIR) * CEV * GAMMA / CPAIR;
  }

  if (c.O.OPT_STC == 1) {
    if (SNOWH > 0.05f && TGB > TFRZ) {
      TGB = TFRZ;
      IRB = CIR * POWI(TGB, 4) - EMG * LWDN;
      SHB = CSH * (TGB - SFCTMP);
      EVB = CEV * (ESTG * RHSUR - EAIR);
      GHB = SAG - (IRB + SHB + EVB);
    }
  }

  TAUXB = -RHOAIR * CM * UR * UU;
  TAUYB = -RHOAIR * CM * UR * VV;

  if (c.O.OPT_SFC == 1 || c.O.OPT_SFC == 2) {
    EHB2 = FV * VKC / LOG((2.f + Z0H) / Z0H);
    EHB2 = FV * VKC / (LOG((2.f + Z0H) / Z0H) - FH2);
    float CQ2B = EHB2;
    if (EHB2 < 1.E-5f) {
      T2MB = TGB;
      Q2B = QSFC;
    } else {
      T2MB = TGB - SHB / (RHOAIR * CPAIR) * 1.f / EHB2;
      Q2B = QSFC - EVB / (LATHEA * RHOAIR) * (1.f / CQ2B + RSURF);
    }
    if (IVGTYP == ISURBAN) Q2B = QSFC;
  }
  CH = EHB;
}

// noahmplsm.F90:5979-6036
void ROSR12(ASnSo& P, const ASnSo& A, const ASnSo& B, ASnSo& C, const ASnSo& D, ASnSo& DELTA, int NTOP,
            int NSOILX, int NSNOWX) {
  (void)NSNOWX;
  C(NSOILX) = 0.0f;
  P(NTOP) = -C(NTOP) / B(NTOP);
  DELTA(NTOP) = D(NTOP) / B(NTOP);
  for (int K = NTOP + 1; K <= NSOILX; ++K) {
    P(K) = -C(K) * (1.0f / (B(K) + A(K) * P(K - 1)));
    DELTA(K) = (D(K) - A(K) * DELTA(K - 1)) * (1.0f / (B(K) + A(K) * P(K - 1)));
  }
  P(NSOILX) = DELTA(NSOILX);
  for (int K = NTOP + 1; K <= NSOILX; ++K) {
    int KK = NSOILX - K + (NTOP - 1) + 1;
    P(KK) = P(KK) * P(KK + 1) + DELTA(KK);
  }
}

// noahmplsm.F90:5825-5922
static void HRT(Ctx& c, int ISNOW, const ASnSo& ZSNSO, const ASnSo& STC, float TBOT, float ZBOT,
                const ASnSo& DF, const ASnSo& HCPCT, float SSOIL, const ASnSo& PHI, ASnSo& AI, ASnSo& BI,
                ASnSo& CI, ASnSo& RHSTS, float& BOTFLX) {
  ASnSo DDZ, DENOM, DTSDZ, EFLUX;
  DDZ.fill(0.f); DENOM.fill(0.f); DTSDZ.fill(0.f); EFLUX.fill(0.f);
  BOTFLX = 0.f;
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {
    if (K == ISNOW + 1) {
      DENOM(K) = -ZSNSO(K) * HCPCT(K);
      float TEMP1 = -ZSNSO(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DTSDZ(K) = 2.0f * (STC(K) - STC(K + 1)) / TEMP1;
      EFLUX(K) = DF(K) * DTSDZ(K) - SSOIL - PHI(K);
    } else if (K < NSOIL) {
      DENOM(K) = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
      float TEMP1 = ZSNSO(K - 1) - ZSNSO(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DTSDZ(K) = 2.0f * (STC(K) - STC(K + 1)) / TEMP1;
      EFLUX(K) = (DF(K) * DTSDZ(K) - DF(K - 1) * DTSDZ(K - 1)) - PHI(K);
    } else if (K == NSOIL) {
      DENOM(K) = (ZSNSO(K - 1) - ZSNSO(K)) * HCPCT(K);
      if (c.O.OPT_TBOT == 1) BOTFLX = 0.f;
      if (c.O.OPT_TBOT == 2) {
        DTSDZ(K) = (STC(K) - TBOT) / (0.5f * (ZSNSO(K - 1) + ZSNSO(K)) - ZBOT);
        BOTFLX = -DF(K) * DTSDZ(K);
      }
      EFLUX(K) = (-BOTFLX - DF(K - 1) * DTSDZ(K - 1)) - PHI(K);
    }
  }
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {
    if (K == ISNOW + 1) {
      AI(K) = 0.0f;
      CI(K) = -DF(K) * DDZ(K) / DENOM(K);
      if (c.O.OPT_STC == 1) BI(K) = -CI(K);
      if (c.O.OPT_STC == 2) BI(K) = -CI(K) + DF(K) / (0.5f * ZSNSO(K) * ZSNSO(K) * HCPCT(K));
    } else if (K < NSOIL) {
      AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = -DF(K) * DDZ(K) / DENOM(K);
      BI(K) = -(AI(K) + CI(K));
    } else if (K == NSOIL) {
      AI(K) = -DF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = 0.0f;
      BI(K) = -(AI(K) + CI(K));
    }
    RHSTS(K) = EFLUX(K) / (-DENOM(K));
  }
}

// noahmplsm.F90:5925-5977
static void HSTEP(int ISNOW, float DT, ASnSo& AI, ASnSo& BI, ASnSo& CI, ASnSo& RHSTS, ASnSo& STC) {
  ASnSo RHSTSIN, CIIN;
  RHSTSIN.fill(0.f); CIIN.fill(0.f);
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {
    RHSTS(K) = RHSTS(K) * DT;
    AI(K) = AI(K) * DT;
    BI(K) = 1.f + BI(K) * DT;
    CI(K) = CI(K) * DT;
  }
  for (int K = ISNOW + 1; K <= NSOIL; ++K) {
    RHSTSIN(K) = RHSTS(K);
    CIIN(K) = CI(K);
  }
  ROSR12(CI, AI, BI, CIIN, RHSTSIN, RHSTS, ISNOW + 1, NSOIL, NSNOW);
  for (int K = ISNOW + 1; K <= NSOIL; ++K) STC(K) = STC(K) + CI(K);
}

// noahmplsm.F90:5707-5822 (everything after the RETURN at :5797 is dead code)
void TSNOSOI(Ctx& c, int ICE, int ISNOW, int IST, float TBOT, const ASnSo& ZSNSO, float SSOIL,
             const ASnSo& DF, const ASnSo& HCPCT, float ZBOT, float SAG, float DT, float SNOWH,
             const ASnSo& DZSNSO, float TG, ASnSo& STC) {
  (void)ICE; (void)IST; (void)SAG; (void)DZSNSO; (void)TG;
  ASnSo AI, BI, CI, RHSTS, PHI;
  AI.fill(0.f); BI.fill(0.f); CI.fill(0.f); RHSTS.fill(0.f); PHI.fill(0.f);
  float ZBOTSNO = ZBOT - SNOWH;
  float EFLXB;
  HRT(c, ISNOW, ZSNSO, STC, TBOT, ZBOTSNO, DF, HCPCT, SSOIL, PHI, AI, BI, CI, RHSTS, EFLXB);
  HSTEP(ISNOW, DT, AI, BI, CI, RHSTS, STC);
}

// noahmplsm.F90:6247-6377
static void FRH2O(Ctx& c, float& FREE, float TKELV, float SMC, float SH2O) {
  const Params& P = c.P;
  const float CK = 8.0f, BLIM = 5.5f, ERROR_ = 0.005f;
  float BX = P.BEXP;
  if (P.BEXP > BLIM) BX = BLIM;
  int NLOG = 0, KCOUNT = 0;
  if (TKELV > (TFRZ - 1.E-3f)) {
    FREE = SMC;
  } else {
    float SWL = SMC - SH2O;
    if (SWL > (SMC - 0.02f)) SWL = SMC - 0.02f;
    if (SWL < 0.f) SWL = 0.f;
    while ((NLOG < 10) && (KCOUNT == 0)) {
      NLOG = NLOG + 1;
      float t1 = (1.f + CK * SWL);
      float DF = LOG((P.PSISAT * GRAV / HFUS) * POW(t1, 2.f) * POW(P.SMCMAX / (SMC - SWL), BX)) -
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
}

// noahmplsm.F90:6039-6245
void PHASECHANGE(Ctx& c, int ISNOW, float DT, const ASnSo& FACT, const ASnSo& DZSNSO, const ASnSo& HCPCT,
                 int IST, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, float& SNEQV, float& SNOWH, ASoil& SMC,
                 ASoil& SH2O, float& QMELT, IA<-NSNOW + 1, NSOIL>& IMELT, float& PONDING) {
  (void)HCPCT;
  const Params& P = c.P;
  ASnSo HM, XM, WMASS0, WICE0, WLIQ0, MICE, MLIQ, SUPERCOOL;
  HM.fill(0.f); XM.fill(0.f); WMASS0.fill(0.f); WICE0.fill(0.f); WLIQ0.fill(0.f); MICE.fill(0.f);
  MLIQ.fill(0.f);
  QMELT = 0.f;
  PONDING = 0.f;
  float XMF = 0.f;
  for (int J = -NSNOW + 1; J <= NSOIL; ++J) { SUPERCOOL(J) = 0.0f; IMELT(J) = 0; }
  for (int J = ISNOW + 1; J <= 0; ++J) { MICE(J) = SNICE(J); MLIQ(J) = SNLIQ(J); }
  for (int J = 1; J <= NSOIL; ++J) {
    MLIQ(J) = SH2O(J) * DZSNSO(J) * 1000.f;
    MICE(J) = (SMC(J) - SH2O(J)) * DZSNSO(J) * 1000.f;
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    IMELT(J) = 0; HM(J) = 0.f; XM(J) = 0.f;
    WICE0(J) = MICE(J); WLIQ0(J) = MLIQ(J); WMASS0(J) = MICE(J) + MLIQ(J);
  }
  if (IST == 1) {
    for (int J = 1; J <= NSOIL; ++J) {
      if (c.O.OPT_FRZ == 1) {
        if (STC(J) < TFRZ) {
          float SMP = HFUS * (TFRZ - STC(J)) / (GRAV * STC(J));
          SUPERCOOL(J) = P.SMCMAX * POW(SMP / P.PSISAT, -1.f / P.BEXP);
          SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.f;
        }
      }
      if (c.O.OPT_FRZ == 2) {
        FRH2O(c, SUPERCOOL(J), STC(J), SMC(J), SH2O(J));
        SUPERCOOL(J) = SUPERCOOL(J) * DZSNSO(J) * 1000.f;
      }
    }
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
    if (MICE(J) > 0.f && STC(J) >= TFRZ) IMELT(J) = 1;
    if (MLIQ(J) > SUPERCOOL(J) && STC(J) < TFRZ) IMELT(J) = 2;
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
    float HEATR = HM(1) - HFUS * (TEMP1 - SNEQV) / DT;
    if (HEATR > 0.f) {
      XM(1) = HEATR * DT / HFUS;
      HM(1) = HEATR;
    } else {
      XM(1) = 0.f;
      HM(1) = 0.f;
    }
    QMELT = MAX(0.f, (TEMP1 - SNEQV)) / DT;
    XMF = HFUS * QMELT;
    PONDING = TEMP1 - SNEQV;
  }
  for (int J = ISNOW + 1; J <= NSOIL; ++J) {
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
      XMF = XMF + HFUS * (WICE0(J) - MICE(J)) / DT;
      if (J < 1) QMELT = QMELT + MAX(0.f, (WICE0(J) - MICE(J))) / DT;
    }
  }
  (void)XMF;
  for (int J = ISNOW + 1; J <= 0; ++J) { SNLIQ(J) = MLIQ(J); SNICE(J) = MICE(J); }
  for (int J = 1; J <= NSOIL; ++J) {
    SH2O(J) = MLIQ(J) / (1000.f * DZSNSO(J));
    SMC(J) = (MLIQ(J) + MICE(J)) / (1000.f * DZSNSO(J));
  }
}

}  // namespace nmo

// ---- scalar probes for known-answer tests (tests/test_oracle.py) ------------------------------------------------
extern "C" {
// First pass of SFCDIF1 (ITER = 1: neutral stratification, MOZ = 0): out = {CM, CH, FV, CH2}
void nmo_sfcdif1_neutral(float ZLVL, float ZPD, float Z0M, float Z0H, float UR, float* out) {
  nmo::Ctx c{};
  float MOZ = 0.f, FM = 0.f, FH = 0.f, FM2 = 0.f, FH2 = 0.f, CM = 0.f, CH = 0.f, FV = 0.f, CH2 = 0.f;
  int MOZSGN = 0;
  nmo::SFCDIF1(c, 1, 280.f, 1.2f, 0.f, 0.005f, ZLVL, ZPD, Z0M, Z0H, UR, 1.E-6f, MOZ, MOZSGN, FM, FH, FM2, FH2, CM, CH, FV,
               CH2);
  out[0] = CM; out[1] = CH; out[2] = FV; out[3] = CH2;
}
// TSNOSOI: one implicit heat-diffusion step of the snow/soil column; arrays are the (-2:4) layers, element IZ + 2
void nmo_tsnosoi(int opt_stc, int opt_tbot, int ISNOW, float TBOT, const float* ZSNSO, float SSOIL, const float* DF,
                 const float* HCPCT, float ZBOT, float DT, float SNOWH, float* STC) {
  using namespace nmo;
  Ctx c{};
  c.O.OPT_STC = opt_stc;
  c.O.OPT_TBOT = opt_tbot;
  ASnSo z, df, hc, dz, stc;
  for (int k = -2; k <= NSOIL; ++k) { z(k) = ZSNSO[k + 2]; df(k) = DF[k + 2]; hc(k) = HCPCT[k + 2]; stc(k) = STC[k + 2]; dz(k) = 0.f; }
  TSNOSOI(c, 0, ISNOW, 1, TBOT, z, SSOIL, df, hc, ZBOT, 0.f, DT, SNOWH, dz, 0.f, stc);
  for (int k = -2; k <= NSOIL; ++k) STC[k + 2] = stc(k);
}
// PHASECHANGE on a land column; (-2:4) arrays as element IZ + 2, snow arrays (-2:0) likewise, soil arrays (1:4) as K - 1
void nmo_phasechange(int opt_frz, int ISNOW, float DT, const float* FACT, const float* DZSNSO, float* STC, float* SNICE,
                     float* SNLIQ, float* SNEQV, float* SNOWH, float* SMC, float* SH2O, float BEXP, float PSISAT,
                     float SMCMAX, float* QMELT, int* IMELT, float* PONDING) {
  using namespace nmo;
  Ctx c{};
  c.O.OPT_FRZ = opt_frz;
  c.P.BEXP = BEXP; c.P.PSISAT = PSISAT; c.P.SMCMAX = SMCMAX;
  ASnSo fact, dz, hc, stc;
  ASnow ice, liq;
  ASoil smc, sh2o;
  IA<-NSNOW + 1, NSOIL> im;
  for (int k = -2; k <= NSOIL; ++k) { fact(k) = FACT[k + 2]; dz(k) = DZSNSO[k + 2]; hc(k) = 0.f; stc(k) = STC[k + 2]; im(k) = 0; }
  for (int k = -2; k <= 0; ++k) { ice(k) = SNICE[k + 2]; liq(k) = SNLIQ[k + 2]; }
  for (int k = 1; k <= NSOIL; ++k) { smc(k) = SMC[k - 1]; sh2o(k) = SH2O[k - 1]; }
  PHASECHANGE(c, ISNOW, DT, fact, dz, hc, 1, stc, ice, liq, *SNEQV, *SNOWH, smc, sh2o, *QMELT, im, *PONDING);
  for (int k = -2; k <= NSOIL; ++k) { STC[k + 2] = stc(k); IMELT[k + 2] = im(k); }
  for (int k = -2; k <= 0; ++k) { SNICE[k + 2] = ice(k); SNLIQ[k + 2] = liq(k); }
  for (int k = 1; k <= NSOIL; ++k) { SMC[k - 1] = smc(k); SH2O[k - 1] = sh2o(k); }
}
// FRH2O: liquid water a soil layer keeps below freezing
float nmo_frh2o(float TKELV, float SMC, float SH2O, float BEXP, float PSISAT, float SMCMAX) {
  nmo::Ctx c{};
  c.P.BEXP = BEXP; c.P.PSISAT = PSISAT; c.P.SMCMAX = SMCMAX;
  float FREE = 0.f;
  nmo::FRH2O(c, FREE, TKELV, SMC, SH2O);
  return FREE;
}
// STOMATA: in = {APAR FOLN TV EI EA SFCTMP SFCPRS O2 CO2 IGS BTRAN RB}; out = {RS PSN}
void nmo_stomata(const noahmp_tables* T, int VEGTYP, const float* in, float* out2) {
  nmo::Ctx c{};
  c.T = T;
  nmo::STOMATA(c, VEGTYP, 1.E-6f, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], in[9], in[10], in[11],
               out2[0], out2[1]);
}
}
