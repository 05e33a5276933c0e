// nmo_land3.cpp — ORACLE (test infrastructure): WATER, CANWATER, SNOWWATER (SNOWFALL, COMPACT, COMBINE,
// DIVIDE, COMBO, SNOWH2O), SOILWATER (ZWTEQ, INFIL, SRT, SSTEP, WDFCND1/2), GROUNDWATER,
// SHALLOWWATERTABLE, CARBON/CO2FLUX, REDPRM.
// Restates phys/module_sf_noahmplsm.F90:6382-9104 and :9202-9349.
#include "nmo_land.h"

namespace nmo {

// noahmplsm.F90:6615-6865
static void CANWATER(Ctx& c, int VEGTYP, float DT, float SFCTMP, float UU, float VV, float FCEV, float FCTR,
                     float QPRECC, float QPRECL, float ELAI, float ESAI, int IST, float TG, float FVEG,
                     bool FROZEN_CANOPY, float& CANLIQ, float& CANICE, float& TV, float& CMC, float& ECAN,
                     float& ETRAN, float& QRAIN, float& QSNOW, float& SNOWHIN, float& FWET, float& FPICE) {
  float FP = 0.0f, RAIN = 0.0f, SNOW = 0.0f, QINTR = 0.f, QDRIPR = 0.f, QTHROR = 0.f, QINTS = 0.f,
        QDRIPS = 0.0f, QTHROS = 0.f;
  float QEVAC, QDEWC, QSUBC, QFROC;
  QRAIN = 0.0f; QSNOW = 0.0f; SNOWHIN = 0.0f; ECAN = 0.0f;
  FPICE = 0.f;
  if (c.O.OPT_SNF == 1) {
    if (SFCTMP > TFRZ + 2.5f) {
      FPICE = 0.f;
    } else {
      if (SFCTMP <= TFRZ + 0.5f) FPICE = 1.0f;
      else if (SFCTMP <= TFRZ + 2.f) FPICE = 1.f - (-54.632f + 0.2f * SFCTMP);
      else FPICE = 0.6f;
    }
  }
  if (c.O.OPT_SNF == 2) {
    if (SFCTMP >= TFRZ + 2.2f) FPICE = 0.f; else FPICE = 1.0f;
  }
  if (c.O.OPT_SNF == 3) {
    if (SFCTMP >= TFRZ) FPICE = 0.f; else FPICE = 1.0f;
  }
  float BDFALL = MIN(120.f, 67.92f + 51.25f * EXP((SFCTMP - TFRZ) / 2.59f));
  RAIN = (QPRECC + QPRECL) * (1.f - FPICE);
  SNOW = (QPRECC + QPRECL) * FPICE;
  if (QPRECC + QPRECL > 0.f) FP = (QPRECC + QPRECL) / (10.f * QPRECC + QPRECL);
  float MAXLIQ = TV1(c.T->ch2op, VEGTYP) * (ELAI + ESAI);
  if ((ELAI + ESAI) > 0.f) {
    QINTR = FVEG * RAIN * FP;
    QINTR = MIN(QINTR, (MAXLIQ - CANLIQ) / DT * (1.f - EXP(-RAIN * DT / MAXLIQ)));
    QINTR = MAX(QINTR, 0.f);
    QDRIPR = FVEG * RAIN - QINTR;
    QTHROR = (1.f - FVEG) * RAIN;
  } else {
    QINTR = 0.f; QDRIPR = 0.f; QTHROR = RAIN;
  }
  if (!FROZEN_CANOPY) {
    ETRAN = MAX(FCTR / HVAP, 0.f);
    QEVAC = MAX(FCEV / HVAP, 0.f);
    QDEWC = ABS(MIN(FCEV / HVAP, 0.f));
    QSUBC = 0.f; QFROC = 0.f;
  } else {
    ETRAN = MAX(FCTR / HSUB, 0.f);
    QEVAC = 0.f; QDEWC = 0.f;
    QSUBC = MAX(FCEV / HSUB, 0.f);
    QFROC = ABS(MIN(FCEV / HSUB, 0.f));
  }
  QEVAC = MIN(CANLIQ / DT, QEVAC);
  CANLIQ = MAX(0.f, CANLIQ + (QINTR + QDEWC - QEVAC) * DT);
  if (CANLIQ <= 1.E-06f) CANLIQ = 0.0f;
  float MAXSNO = 6.6f * (0.27f + 46.f / BDFALL) * (ELAI + ESAI);
  if ((ELAI + ESAI) > 0.f) {
    QINTS = FVEG * SNOW * FP;
    QINTS = MIN(QINTS, (MAXSNO - CANICE) / DT * (1.f - EXP(-SNOW * DT / MAXSNO)));
    QINTS = MAX(QINTS, 0.f);
    float FT = MAX(0.0f, (TV - 270.15f) / 1.87E5f);
    float FV = SQRT(UU * UU + VV * VV) / 1.56E5f;
    QDRIPS = MAX(0.f, CANICE) * (FV + FT);
    QTHROS = (1.0f - FVEG) * SNOW + (FVEG * SNOW - QINTS);
  } else {
    QINTS = 0.f; QDRIPS = 0.f; QTHROS = SNOW;
  }
  QSUBC = MIN(CANICE / DT, QSUBC);
  CANICE = MAX(0.f, CANICE + (QINTS - QDRIPS) * DT + (QFROC - QSUBC) * DT);
  if (CANICE <= 1.E-6f) CANICE = 0.f;
  if (CANICE > 0.f) FWET = MAX(0.f, CANICE) / MAX(MAXSNO, 1.E-06f);
  else FWET = MAX(0.f, CANLIQ) / MAX(MAXLIQ, 1.E-06f);
  FWET = POW(MIN(FWET, 1.f), 0.667f);
  float QMELTC = 0.f, QFRZC = 0.f;
  if (CANICE > 1.E-6f && TV > TFRZ) {
    QMELTC = MIN(CANICE / DT, (TV - TFRZ) * CICE * CANICE / DENICE / (DT * HFUS));
    CANICE = MAX(0.f, CANICE - QMELTC * DT);
    CANLIQ = MAX(0.f, CANLIQ + QMELTC * DT);
    TV = FWET * TFRZ + (1.f - FWET) * TV;
  }
  if (CANLIQ > 1.E-6f && TV < TFRZ) {
    QFRZC = MIN(CANLIQ / DT, (TFRZ - TV) * CWAT * CANLIQ / DENH2O / (DT * HFUS));
    CANLIQ = MAX(0.f, CANLIQ - QFRZC * DT);
    CANICE = MAX(0.f, CANICE + QFRZC * DT);
    TV = FWET * TFRZ + (1.f - FWET) * TV;
  }
  CMC = CANLIQ + CANICE;
  ECAN = QEVAC + QSUBC - QDEWC - QFROC;
  QRAIN = QDRIPR + QTHROR;
  QSNOW = QDRIPS + QTHROS;
  SNOWHIN = QSNOW / BDFALL;
  if (IST == 2 && TG > TFRZ) { QSNOW = 0.f; SNOWHIN = 0.f; }
}

// noahmplsm.F90:6998-7063
static void SNOWFALL(float DT, float QSNOW, float SNOWHIN, float SFCTMP, int& ISNOW, float& SNOWH,
                     ASnSo& DZSNSO, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, float& SNEQV) {
  int NEWNODE = 0;
  if (ISNOW == 0 && QSNOW > 0.f) {
    SNOWH = SNOWH + SNOWHIN * DT;
    SNEQV = SNEQV + QSNOW * DT;
  }
  if (ISNOW == 0 && QSNOW > 0.f && SNOWH >= 0.025f) {
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

// noahmplsm.F90:7375-7424
void COMBO(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2, float WICE2,
                  float T2) {
  float DZC = DZ + DZ2;
  float WICEC = (WICE + WICE2);
  float WLIQC = (WLIQ + WLIQ2);
  float H = (CICE * WICE + CWAT * WLIQ) * (T - TFRZ) + HFUS * WLIQ;
  float H2 = (CICE * WICE2 + CWAT * WLIQ2) * (T2 - TFRZ) + HFUS * WLIQ2;
  float HC = H + H2;
  float TC;
  if (HC < 0.f) TC = TFRZ + HC / (CICE * WICEC + CWAT * WLIQC);
  else if (HC <= HFUS * WLIQC) TC = TFRZ;
  else TC = TFRZ + (HC - HFUS * WLIQC) / (CICE * WICEC + CWAT * WLIQC);
  DZ = DZC; WICE = WICEC; WLIQ = WLIQC; T = TC;
}

// noahmplsm.F90:7065-7246
static void COMBINE(int& ISNOW, ASoil& SH2O, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, ASnSo& DZSNSO,
                    ASoil& SICE, float& SNOWH, float& SNEQV, float& PONDING1, float& PONDING2) {
  static const float DZMIN[3] = {0.025f, 0.025f, 0.1f};
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
          if (SNICE(J) >= 0.f) {
            PONDING1 = SNLIQ(J);
            SNEQV = SNICE(J);
            SNOWH = DZSNSO(J);
          } else {
            PONDING1 = SNLIQ(J) + SNICE(J);
            if (PONDING1 < 0.f) {
              SICE(1) = MAX(0.0f, SICE(1) + PONDING1 / (DZSNSO(1) * 1000.f));
              PONDING1 = 0.0f;
            }
            SNEQV = 0.0f;
            SNOWH = 0.0f;
          }
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
  if (SNOWH < 0.025f && ISNOW < 0) {
    ISNOW = 0;
    SNEQV = ZWICE;
    PONDING2 = ZWLIQ;
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

// noahmplsm.F90:7248-7371
static void DIVIDE(int& ISNOW, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, ASnSo& DZSNSO) {
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
      if (MSNO <= 2 && DZ(2) > 0.20f) {
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

// noahmplsm.F90:7427-7528
void COMPACT(float DT, const ASnSo& STC, const ASnow& SNICE, const ASnow& SNLIQ,
                    const IA<-NSNOW + 1, NSOIL>& IMELT, const ASnow& FICEOLD, int ISNOW, ASnSo& DZSNSO) {
  const float C2 = 21.e-3f, C3 = 2.5e-6f, C4 = 0.04f, C5 = 2.0f, DM = 100.0f, ETA0 = 0.8e+6f;
  ASnow FICE;
  float BURDEN = 0.0f;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    float WX = SNICE(J) + SNLIQ(J);
    FICE(J) = SNICE(J) / WX;
    float VOID = 1.f - (SNICE(J) / DENICE + SNLIQ(J) / DENH2O) / DZSNSO(J);
    if (VOID > 0.001f && SNICE(J) > 0.1f) {
      float BI = SNICE(J) / DZSNSO(J);
      float TD = MAX(0.f, TFRZ - STC(J));
      float DEXPF = EXP(-C4 * TD);
      float DDZ1 = -C3 * DEXPF;
      if (BI > DM) DDZ1 = DDZ1 * EXP(-46.0E-3f * (BI - DM));
      if (SNLIQ(J) > 0.01f * DZSNSO(J)) DDZ1 = DDZ1 * C5;
      float DDZ2 = -(BURDEN + 0.5f * WX) * EXP(-0.08f * TD - C2 * BI) / ETA0;
      float DDZ3;
      if (IMELT(J) == 1) {
        DDZ3 = MAX(0.f, (FICEOLD(J) - FICE(J)) / MAX(1.E-6f, FICEOLD(J)));
        DDZ3 = -DDZ3 / DT;
      } else {
        DDZ3 = 0.f;
      }
      float PDZDTC = (DDZ1 + DDZ2 + DDZ3) * DT;
      PDZDTC = MAX(-0.5f, PDZDTC);
      DZSNSO(J) = DZSNSO(J) * (1.f + PDZDTC);
    }
    BURDEN = BURDEN + WX;
  }
}

// noahmplsm.F90:7530-7678
static void SNOWH2O(float DT, float QSNFRO, float QSNSUB, float QRAIN, int& ISNOW, ASnSo& DZSNSO,
                    float& SNOWH, float& SNEQV, ASnow& SNICE, ASnow& SNLIQ, ASoil& SH2O, ASoil& SICE,
                    ASnSo& STC, float& QSNBOT, float& PONDING1, float& PONDING2) {
  ASnow VOL_LIQ, VOL_ICE, EPORE;
  VOL_LIQ.fill(0.f); VOL_ICE.fill(0.f); EPORE.fill(0.f);
  if (SNEQV == 0.f) {
    SICE(1) = SICE(1) + (QSNFRO - QSNSUB) * DT / (DZSNSO(1) * 1000.f);
    if (SICE(1) < 0.f) {
      SH2O(1) = SH2O(1) + SICE(1);
      SICE(1) = 0.f;
    }
  }
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
      COMBINE(ISNOW, SH2O, STC, SNICE, SNLIQ, DZSNSO, SICE, SNOWH, SNEQV, PONDING1, PONDING2);
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

// noahmplsm.F90:6868-6996
static void SNOWWATER(const IA<-NSNOW + 1, NSOIL>& IMELT, float DT, const ASoil& ZSOIL, float SFCTMP,
                      float SNOWHIN, float QSNOW, float QSNFRO, float QSNSUB, float QRAIN,
                      const ASnow& FICEOLD, int& ISNOW, float& SNOWH, float& SNEQV, ASnow& SNICE,
                      ASnow& SNLIQ, ASoil& SH2O, ASoil& SICE, ASnSo& STC, ASnSo& ZSNSO, ASnSo& DZSNSO,
                      float& QSNBOT, float& SNOFLOW, float& PONDING1, float& PONDING2) {
  SNOFLOW = 0.0f; PONDING1 = 0.0f; PONDING2 = 0.0f;
  SNOWFALL(DT, QSNOW, SNOWHIN, SFCTMP, ISNOW, SNOWH, DZSNSO, STC, SNICE, SNLIQ, SNEQV);
  if (ISNOW < 0) COMPACT(DT, STC, SNICE, SNLIQ, IMELT, FICEOLD, ISNOW, DZSNSO);
  if (ISNOW < 0) COMBINE(ISNOW, SH2O, STC, SNICE, SNLIQ, DZSNSO, SICE, SNOWH, SNEQV, PONDING1, PONDING2);
  if (ISNOW < 0) DIVIDE(ISNOW, STC, SNICE, SNLIQ, DZSNSO);
  SNOWH2O(DT, QSNFRO, QSNSUB, QRAIN, ISNOW, DZSNSO, SNOWH, SNEQV, SNICE, SNLIQ, SH2O, SICE, STC, QSNBOT,
          PONDING1, PONDING2);
  for (int IZ = -NSNOW + 1; IZ <= ISNOW; ++IZ) {
    SNICE(IZ) = 0.f; SNLIQ(IZ) = 0.f; STC(IZ) = 0.f; DZSNSO(IZ) = 0.f; ZSNSO(IZ) = 0.f;
  }
  if (SNEQV > 2000.f) {
    float BDSNOW = SNICE(0) / DZSNSO(0);
    SNOFLOW = (SNEQV - 2000.f);
    SNICE(0) = SNICE(0) - SNOFLOW;
    DZSNSO(0) = DZSNSO(0) - SNOFLOW / BDSNOW;
    SNOFLOW = SNOFLOW / DT;
  }
  if (ISNOW < 0) {
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

// noahmplsm.F90:8329-8362
static void WDFCND1(const Ctx& c, float& WDF, float& WCND, float SMC, float FCR) {
  const Params& P = c.P;
  float FACTR = MAX(0.01f, SMC / P.SMCMAX);
  float EXPON = P.BEXP + 2.0f;
  WDF = P.DWSAT * POW(FACTR, EXPON);
  WDF = WDF * (1.0f - FCR);
  EXPON = 2.0f * P.BEXP + 3.0f;
  WCND = P.DKSAT * POW(FACTR, EXPON);
  WCND = WCND * (1.0f - FCR);
}

// noahmplsm.F90:8364-8400
static void WDFCND2(const Ctx& c, float& WDF, float& WCND, float SMC, float SICE) {
  const Params& P = c.P;
  float FACTR = MAX(0.01f, SMC / P.SMCMAX);
  float EXPON = P.BEXP + 2.0f;
  WDF = P.DWSAT * POW(FACTR, EXPON);
  if (SICE > 0.0f) {
    float VKWGT = 1.f / (1.f + POW(500.f * SICE, 3.f));
    WDF = VKWGT * WDF + (1.f - VKWGT) * P.DWSAT * POW(0.2f / P.SMCMAX, EXPON);
  }
  EXPON = 2.0f * P.BEXP + 3.0f;
  WCND = P.DKSAT * POW(FACTR, EXPON);
}

// noahmplsm.F90:7938-7989
static void ZWTEQ(const Ctx& c, const ASoil& ZSOIL, const ASnSo& DZSNSO, const ASoil& SH2O, float& ZWT) {
  const Params& P = c.P;
  const int NFINE = 100;
  float ZFINE[NFINE + 1];
  float WD1 = 0.f;
  for (int K = 1; K <= NSOIL; ++K) WD1 = WD1 + (P.SMCMAX - SH2O(K)) * DZSNSO(K);
  float DZFINE = 3.0f * (-ZSOIL(NSOIL)) / (float)NFINE;
  for (int K = 1; K <= NFINE; ++K) ZFINE[K] = (float)K * DZFINE;
  ZWT = -3.f * ZSOIL(NSOIL) - 0.001f;
  float WD2 = 0.f;
  for (int K = 1; K <= NFINE; ++K) {
    float TEMP = 1.f + (ZWT - ZFINE[K]) / P.PSISAT;
    WD2 = WD2 + P.SMCMAX * (1.f - POW(TEMP, -1.f / P.BEXP)) * DZFINE;
    if (ABS(WD2 - WD1) <= 0.01f) {
      ZWT = ZFINE[K];
      break;
    }
  }
}

// noahmplsm.F90:7992-8087
static void INFIL(const Ctx& c, float DT, const ASoil& ZSOIL, const ASoil& SH2O, const ASoil& SICE,
                  float SICEMAX, float QINSUR, float& PDDUM, float& RUNSRF) {
  const Params& P = c.P;
  const int CVFRZ = 3;
  ASoil DMAX;
  if (QINSUR > 0.0f) {
    float DT1 = DT / 86400.f;
    float SMCAV = P.SMCMAX - P.SMCWLT;
    DMAX(1) = -ZSOIL(1) * SMCAV;
    float DICE = -ZSOIL(1) * SICE(1);
    DMAX(1) = DMAX(1) * (1.0f - (SH2O(1) + SICE(1) - P.SMCWLT) / SMCAV);
    float DD = DMAX(1);
    for (int K = 2; K <= NSOIL; ++K) {
      DICE = DICE + (ZSOIL(K - 1) - ZSOIL(K)) * SICE(K);
      DMAX(K) = (ZSOIL(K - 1) - ZSOIL(K)) * SMCAV;
      DMAX(K) = DMAX(K) * (1.0f - (SH2O(K) + SICE(K) - P.SMCWLT) / SMCAV);
      DD = DD + DMAX(K);
    }
    float VAL = (1.f - EXP(-P.KDT * DT1));
    float DDT = DD * VAL;
    float PX = MAX(0.f, QINSUR * DT);
    float INFMAX = (PX * (DDT / (PX + DDT))) / DT;
    float FCR = 1.f;
    if (DICE > 1.E-2f) {
      float ACRT = (float)CVFRZ * P.FRZX / DICE;
      float SUM = 1.f;
      int IALP1 = CVFRZ - 1;
      for (int J = 1; J <= IALP1; ++J) {
        int K = 1;
        for (int JJ = J + 1; JJ <= IALP1; ++JJ) K = K * JJ;
        SUM = SUM + POWI(ACRT, CVFRZ - J) / (float)K;
      }
      FCR = 1.f - EXP(-ACRT) * SUM;
    }
    INFMAX = INFMAX * FCR;
    float WDF, WCND;
    WDFCND2(c, WDF, WCND, SH2O(1), SICEMAX);
    INFMAX = MAX(INFMAX, WCND);
    INFMAX = MIN(INFMAX, PX);
    RUNSRF = MAX(0.f, QINSUR - INFMAX);
    PDDUM = QINSUR - RUNSRF;
  }
}

// noahmplsm.F90:8089-8217
static void SRT(const Ctx& c, const ASoil& ZSOIL, float DT, float PDDUM, const ASoil& ETRANI, float QSEVA,
                const ASoil& SH2O, const ASoil& SMC, float ZWT, const ASoil& FCR, float SICEMAX,
                float FCRMAX, float SMCWTD, ASoil& RHSTT, ASoil& AI, ASoil& BI, ASoil& CI, float& QDRAIN,
                ASoil& WCND) {
  (void)DT;
  const Params& P = c.P;
  ASoil DDZ, DENOM, DSMDZ, WFLUX, WDF, SMX;
  DDZ.fill(0.f); DENOM.fill(0.f); DSMDZ.fill(0.f); WFLUX.fill(0.f); WDF.fill(0.f); SMX.fill(0.f);
  float SMXWTD = 0.f;
  if (c.O.OPT_INF == 1) {
    for (int K = 1; K <= NSOIL; ++K) {
      WDFCND1(c, WDF(K), WCND(K), SMC(K), FCR(K));
      SMX(K) = SMC(K);
    }
    if (c.O.OPT_RUN == 5) SMXWTD = SMCWTD;
  }
  if (c.O.OPT_INF == 2) {
    for (int K = 1; K <= NSOIL; ++K) {
      WDFCND2(c, WDF(K), WCND(K), SH2O(K), SICEMAX);
      SMX(K) = SH2O(K);
    }
    if (c.O.OPT_RUN == 5) SMXWTD = SMCWTD * SH2O(NSOIL) / SMC(NSOIL);
  }
  for (int K = 1; K <= NSOIL; ++K) {
    if (K == 1) {
      DENOM(K) = -ZSOIL(K);
      float TEMP1 = -ZSOIL(K + 1);
      DDZ(K) = 2.0f / TEMP1;
      DSMDZ(K) = 2.0f * (SMX(K) - SMX(K + 1)) / TEMP1;
      WFLUX(K) = WDF(K) * DSMDZ(K) + WCND(K) - PDDUM + ETRANI(K) + QSEVA;
    } else if (K < NSOIL) {
      DENOM(K) = (ZSOIL(K - 1) - ZSOIL(K));
      float TEMP1 = (ZSOIL(K - 1) - ZSOIL(K + 1));
      DDZ(K) = 2.0f / TEMP1;
      DSMDZ(K) = 2.0f * (SMX(K) - SMX(K + 1)) / TEMP1;
      WFLUX(K) = WDF(K) * DSMDZ(K) + WCND(K) - WDF(K - 1) * DSMDZ(K - 1) - WCND(K - 1) + ETRANI(K);
    } else {
      DENOM(K) = (ZSOIL(K - 1) - ZSOIL(K));
      if (c.O.OPT_RUN == 1 || c.O.OPT_RUN == 2) QDRAIN = 0.f;
      if (c.O.OPT_RUN == 3) QDRAIN = P.SLOPE * WCND(K);
      if (c.O.OPT_RUN == 4) QDRAIN = (1.0f - FCRMAX) * WCND(K);
      if (c.O.OPT_RUN == 5) {
        float TEMP1 = 2.0f * DENOM(K);
        float SMXBOT;
        if (ZWT < ZSOIL(NSOIL) - DENOM(NSOIL)) {
          SMXBOT = SMX(K) - (SMX(K) - SMXWTD) * DENOM(K) * 2.f / (DENOM(K) + ZSOIL(K) - ZWT);
        } else {
          SMXBOT = SMXWTD;
        }
        DSMDZ(K) = 2.0f * (SMX(K) - SMXBOT) / TEMP1;
        QDRAIN = WDF(K) * DSMDZ(K) + WCND(K);
      }
      WFLUX(K) = -(WDF(K - 1) * DSMDZ(K - 1)) - WCND(K - 1) + ETRANI(K) + QDRAIN;
    }
  }
  for (int K = 1; K <= NSOIL; ++K) {
    if (K == 1) {
      AI(K) = 0.0f;
      BI(K) = WDF(K) * DDZ(K) / DENOM(K);
      CI(K) = -BI(K);
    } else if (K < NSOIL) {
      AI(K) = -WDF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = -WDF(K) * DDZ(K) / DENOM(K);
      BI(K) = -(AI(K) + CI(K));
    } else {
      AI(K) = -WDF(K - 1) * DDZ(K - 1) / DENOM(K);
      CI(K) = 0.0f;
      BI(K) = -(AI(K) + CI(K));
    }
    RHSTT(K) = WFLUX(K) / (-DENOM(K));
  }
}

// noahmplsm.F90:8220-8327
static void SSTEP(const Ctx& c, float DT, const ASoil& ZSOIL, const ASnSo& DZSNSO, const ASoil& SICE,
                  float ZWT, ASoil& SH2O, ASoil& SMC, ASoil& AI, ASoil& BI, ASoil& CI, ASoil& RHSTT,
                  float& SMCWTD, float& QDRAIN, float& DEEPRECH, float& WPLUS) {
  const Params& P = c.P;
  WPLUS = 0.0f;
  for (int K = 1; K <= NSOIL; ++K) {
    RHSTT(K) = RHSTT(K) * DT;
    AI(K) = AI(K) * DT;
    BI(K) = 1.f + BI(K) * DT;
    CI(K) = CI(K) * DT;
  }
  // ROSR12 on (1:NSOIL) arrays: embed into (-2:4) work arrays
  ASnSo Pw, Aw, Bw, Cw, Dw, DELTAw;
  Pw.fill(0.f); Aw.fill(0.f); Bw.fill(0.f); Cw.fill(0.f); Dw.fill(0.f); DELTAw.fill(0.f);
  for (int K = 1; K <= NSOIL; ++K) { Aw(K) = AI(K); Bw(K) = BI(K); Cw(K) = CI(K); Dw(K) = RHSTT(K); }
  ROSR12(Pw, Aw, Bw, Cw, Dw, DELTAw, 1, NSOIL, 0);
  for (int K = 1; K <= NSOIL; ++K) { CI(K) = Pw(K); RHSTT(K) = DELTAw(K); }
  for (int K = 1; K <= NSOIL; ++K) SH2O(K) = SH2O(K) + CI(K);
  if (c.O.OPT_RUN == 5) {
    if (ZWT < ZSOIL(NSOIL) - DZSNSO(NSOIL)) {
      DEEPRECH = DEEPRECH + DT * QDRAIN;
    } else {
      SMCWTD = SMCWTD + DT * QDRAIN / DZSNSO(NSOIL);
      WPLUS = MAX((SMCWTD - P.SMCMAX), 0.0f) * DZSNSO(NSOIL);
      float WMINUS = MAX((1.E-4f - SMCWTD), 0.0f) * DZSNSO(NSOIL);
      SMCWTD = MAX(MIN(SMCWTD, P.SMCMAX), 1.E-4f);
      SH2O(NSOIL) = SH2O(NSOIL) + WPLUS / DZSNSO(NSOIL);
      QDRAIN = QDRAIN - WPLUS / DT;
      DEEPRECH = DEEPRECH - WMINUS;
    }
  }
  for (int K = NSOIL; K >= 2; --K) {
    float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(K)));
    WPLUS = MAX((SH2O(K) - EPORE), 0.0f) * DZSNSO(K);
    SH2O(K) = MIN(EPORE, SH2O(K));
    SH2O(K - 1) = SH2O(K - 1) + WPLUS / DZSNSO(K - 1);
  }
  float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(1)));
  WPLUS = MAX((SH2O(1) - EPORE), 0.0f) * DZSNSO(1);
  SH2O(1) = MIN(EPORE, SH2O(1));
  for (int K = 1; K <= NSOIL; ++K) SMC(K) = SH2O(K) + SICE(K);
}

// noahmplsm.F90:7680-7936
static void SOILWATER(Ctx& c, float DT, const ASoil& ZSOIL, const ASnSo& DZSNSO, float QINSUR, float QSEVA,
                      const ASoil& ETRANI, const ASoil& SICE, ASoil& SH2O, ASoil& SMC, float& ZWT,
                      int ISURBAN, int VEGTYP, float& SMCWTD, float& DEEPRECH, float& RUNSRF,
                      float& QDRAIN, float& RUNSUB, ASoil& WCND, float& FCRMAX) {
  const Params& P = c.P;
  const float A = 4.0f;
  ASoil RHSTT, AI, BI, CI, MLIQ, FCR;
  RHSTT.fill(0.f); AI.fill(0.f); BI.fill(0.f); CI.fill(0.f); MLIQ.fill(0.f);
  RUNSRF = 0.0f;
  float PDDUM = 0.0f, RSAT = 0.0f, FSAT, FFF, RSBMX, WPLUS;
  QDRAIN = 0.f;
  for (int K = 1; K <= NSOIL; ++K) {
    float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(K)));
    RSAT = RSAT + MAX(0.f, SH2O(K) - EPORE) * DZSNSO(K);
    SH2O(K) = MIN(EPORE, SH2O(K));
  }
  for (int K = 1; K <= NSOIL; ++K) {
    float FICE = MIN(1.0f, SICE(K) / P.SMCMAX);
    FCR(K) = MAX(0.0f, EXP(-A * (1.f - FICE)) - EXP(-A)) / (1.0f - EXP(-A));
  }
  float SICEMAX = 0.0f;
  FCRMAX = 0.0f;
  float SH2OMIN = P.SMCMAX;
  for (int K = 1; K <= NSOIL; ++K) {
    if (SICE(K) > SICEMAX) SICEMAX = SICE(K);
    if (FCR(K) > FCRMAX) FCRMAX = FCR(K);
    if (SH2O(K) < SH2OMIN) SH2OMIN = SH2O(K);
  }
  if (c.O.OPT_RUN == 2) {
    FFF = 2.0f;
    RSBMX = 4.0f;
    ZWTEQ(c, ZSOIL, DZSNSO, SH2O, ZWT);
    RUNSUB = (1.0f - FCRMAX) * RSBMX * EXP(-TIMEAN) * EXP(-FFF * ZWT);
  }
  if (VEGTYP == ISURBAN) FCR(1) = 0.95f;
  if (c.O.OPT_RUN == 1) {
    FFF = 6.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * (ZWT - 2.0f));
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (c.O.OPT_RUN == 5) {
    FFF = 6.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * MAX(-2.0f - ZWT, 0.f));
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (c.O.OPT_RUN == 2) {
    FFF = 2.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * ZWT);
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (c.O.OPT_RUN == 3) INFIL(c, DT, ZSOIL, SH2O, SICE, SICEMAX, QINSUR, PDDUM, RUNSRF);
  if (c.O.OPT_RUN == 4) {
    float SMCTOT = 0.f, DZTOT = 0.f;
    for (int K = 1; K <= NSOIL; ++K) {
      DZTOT = DZTOT + DZSNSO(K);
      SMCTOT = SMCTOT + SMC(K) * DZSNSO(K);
      if (DZTOT >= 2.0f) break;
    }
    SMCTOT = SMCTOT / DZTOT;
    FSAT = POW(MAX(0.01f, SMCTOT / P.SMCMAX), 4.f);
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  int NITER = 1;
  if (c.O.OPT_INF == 1) {
    NITER = 3;
    if (PDDUM * DT > DZSNSO(1) * P.SMCMAX) NITER = NITER * 2;
  }
  float DTFINE = DT / (float)NITER;
  float QDRAIN_SAVE = 0.0f;
  for (int ITER = 1; ITER <= NITER; ++ITER) {
    SRT(c, ZSOIL, DTFINE, PDDUM, ETRANI, QSEVA, SH2O, SMC, ZWT, FCR, SICEMAX, FCRMAX, SMCWTD, RHSTT, AI,
        BI, CI, QDRAIN, WCND);
    SSTEP(c, DTFINE, ZSOIL, DZSNSO, SICE, ZWT, SH2O, SMC, AI, BI, CI, RHSTT, SMCWTD, QDRAIN, DEEPRECH,
          WPLUS);
    RSAT = RSAT + WPLUS;
    QDRAIN_SAVE = QDRAIN_SAVE + QDRAIN;
  }
  QDRAIN = QDRAIN_SAVE / (float)NITER;
  RUNSRF = RUNSRF * 1000.f + RSAT * 1000.f / DT;
  QDRAIN = QDRAIN * 1000.f;
  if (c.O.OPT_RUN == 2) {
    float WTSUB = 0.f;
    for (int K = 1; K <= NSOIL; ++K) WTSUB = WTSUB + WCND(K) * DZSNSO(K);
    for (int K = 1; K <= NSOIL; ++K) {
      float MH2O = RUNSUB * DT * (WCND(K) * DZSNSO(K)) / WTSUB;
      SH2O(K) = SH2O(K) - MH2O / (DZSNSO(K) * 1000.f);
    }
  }
  if (c.O.OPT_RUN != 1) {
    for (int IZ = 1; IZ <= NSOIL; ++IZ) MLIQ(IZ) = SH2O(IZ) * DZSNSO(IZ) * 1000.f;
    float WATMIN = 0.01f, XS;
    for (int IZ = 1; IZ <= NSOIL - 1; ++IZ) {
      if (MLIQ(IZ) < 0.f) XS = WATMIN - MLIQ(IZ);
      else XS = 0.f;
      MLIQ(IZ) = MLIQ(IZ) + XS;
      MLIQ(IZ + 1) = MLIQ(IZ + 1) - XS;
    }
    int IZ = NSOIL;
    if (MLIQ(IZ) < WATMIN) XS = WATMIN - MLIQ(IZ);
    else XS = 0.f;
    MLIQ(IZ) = MLIQ(IZ) + XS;
    RUNSUB = RUNSUB - XS / DT;
    if (c.O.OPT_RUN == 5) DEEPRECH = DEEPRECH - XS * 1.E-3f;
    for (IZ = 1; IZ <= NSOIL; ++IZ) SH2O(IZ) = MLIQ(IZ) / (DZSNSO(IZ) * 1000.f);
  }
}

// noahmplsm.F90:8403-8585
static void GROUNDWATER(Ctx& c, float DT, const ASoil& SICE, const ASoil& ZSOIL, const ASoil& WCND,
                        float FCRMAX, ASoil& SH2O, float& ZWT, float& WA, float& WT, float& QIN,
                        float& QDIS) {
  const Params& P = c.P;
  const float ROUS = 0.2f, CMIC = 0.20f;
  ASoil DZMM, ZNODE, MLIQ, EPORE, HK, SMC;
  QDIS = 0.0f; QIN = 0.0f;
  DZMM(1) = -ZSOIL(1) * 1.E3f;
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZMM(IZ) = 1.E3f * (ZSOIL(IZ - 1) - ZSOIL(IZ));
  ZNODE(1) = -ZSOIL(1) / 2.f;
  for (int IZ = 2; IZ <= NSOIL; ++IZ) ZNODE(IZ) = -ZSOIL(IZ - 1) + 0.5f * (ZSOIL(IZ - 1) - ZSOIL(IZ));
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    SMC(IZ) = SH2O(IZ) + SICE(IZ);
    MLIQ(IZ) = SH2O(IZ) * DZMM(IZ);
    EPORE(IZ) = MAX(0.01f, P.SMCMAX - SICE(IZ));
    HK(IZ) = 1.E3f * WCND(IZ);
  }
  int IWT = NSOIL;
  for (int IZ = 2; IZ <= NSOIL; ++IZ) {
    if (ZWT <= -ZSOIL(IZ)) {
      IWT = IZ - 1;
      break;
    }
  }
  float FFF = 6.0f, RSBMX = 5.0f;
  QDIS = (1.0f - FCRMAX) * RSBMX * EXP(-TIMEAN) * EXP(-FFF * (ZWT - 2.0f));
  // S_NODE is REAL(KIND=8) in the reference (:8443): the pow runs in fp64
  double S_NODE = (double)MIN(1.0f, SMC(IWT) / P.SMCMAX);
  { double lo = (double)0.01f; if (lo > S_NODE) S_NODE = lo; }
  float SMPFZ = (float)(-((double)(P.PSISAT * 1000.f) * DPOW(S_NODE, (double)(-P.BEXP))));
  SMPFZ = MAX(-120000.0f, CMIC * SMPFZ);
  float KA = HK(IWT);
  float WH_ZWT = -ZWT * 1.E3f;
  float WH = SMPFZ - ZNODE(IWT) * 1.E3f;
  QIN = -KA * (WH_ZWT - WH) / ((ZWT - ZNODE(IWT)) * 1.E3f);
  QIN = MAX(-10.0f / DT, MIN(10.f / DT, QIN));
  WT = WT + (QIN - QDIS) * DT;
  if (IWT == NSOIL) {
    WA = WA + (QIN - QDIS) * DT;
    WT = WA;
    ZWT = (-ZSOIL(NSOIL) + 25.f) - WA / 1000.f / ROUS;
    MLIQ(NSOIL) = MLIQ(NSOIL) - QIN * DT;
    MLIQ(NSOIL) = MLIQ(NSOIL) + MAX(0.f, (WA - 5000.f));
    WA = MIN(WA, 5000.f);
  } else {
    if (IWT == NSOIL - 1) {
      ZWT = -ZSOIL(NSOIL) - (WT - ROUS * 1000.f * 25.f) / (EPORE(NSOIL)) / 1000.f;
    } else {
      float WS = 0.f;
      for (int IZ = IWT + 2; IZ <= NSOIL; ++IZ) WS = WS + EPORE(IZ) * DZMM(IZ);
      ZWT = -ZSOIL(IWT + 1) - (WT - ROUS * 1000.f * 25.f - WS) / (EPORE(IWT + 1)) / 1000.f;
    }
    float WTSUB = 0.f;
    for (int IZ = 1; IZ <= NSOIL; ++IZ) WTSUB = WTSUB + HK(IZ) * DZMM(IZ);
    for (int IZ = 1; IZ <= NSOIL; ++IZ) MLIQ(IZ) = MLIQ(IZ) - QDIS * DT * HK(IZ) * DZMM(IZ) / WTSUB;
  }
  ZWT = MAX(1.5f, ZWT);
  float WATMIN = 0.01f, XS;
  for (int IZ = 1; IZ <= NSOIL - 1; ++IZ) {
    if (MLIQ(IZ) < 0.f) XS = WATMIN - MLIQ(IZ);
    else XS = 0.f;
    MLIQ(IZ) = MLIQ(IZ) + XS;
    MLIQ(IZ + 1) = MLIQ(IZ + 1) - XS;
  }
  int IZ = NSOIL;
  if (MLIQ(IZ) < WATMIN) XS = WATMIN - MLIQ(IZ);
  else XS = 0.f;
  MLIQ(IZ) = MLIQ(IZ) + XS;
  WA = WA - XS;
  WT = WT - XS;
  for (IZ = 1; IZ <= NSOIL; ++IZ) SH2O(IZ) = MLIQ(IZ) / DZMM(IZ);
}

// noahmplsm.F90:8588-8718
static void SHALLOWWATERTABLE(Ctx& c, const ASoil& ZSOIL, float DT, const ASnSo& DZSNSO, const ASoil& SMCEQ,
                              const ASoil& SMC, float& WTD, float& SMCWTD, float& RECH, float& QDRAIN) {
  (void)DT; (void)QDRAIN;
  const Params& P = c.P;
  FA<0, NSOIL> ZSOIL0;
  for (int K = 1; K <= NSOIL; ++K) ZSOIL0(K) = ZSOIL(K);
  ZSOIL0(0) = 0.f;
  int IZ;
  for (IZ = NSOIL; IZ >= 1; --IZ) {
    if (WTD + 1.E-6f < ZSOIL0(IZ)) break;
  }
  int IWTD = IZ;
  int KWTD = IWTD + 1;
  float WTDOLD;
  if (KWTD <= NSOIL) {
    WTDOLD = WTD;
    if (SMC(KWTD) > SMCEQ(KWTD)) {
      if (SMC(KWTD) == P.SMCMAX) {
        WTD = ZSOIL0(IWTD);
        RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ(KWTD));
        IWTD = IWTD - 1;
        KWTD = KWTD - 1;
        if (KWTD >= 1) {
          if (SMC(KWTD) > SMCEQ(KWTD)) {
            WTDOLD = WTD;
            WTD = MIN((SMC(KWTD) * DZSNSO(KWTD) - SMCEQ(KWTD) * ZSOIL0(IWTD) + P.SMCMAX * ZSOIL0(KWTD)) /
                          (P.SMCMAX - SMCEQ(KWTD)),
                      ZSOIL0(IWTD));
            RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQ(KWTD));
          }
        }
      } else {
        WTD = MIN((SMC(KWTD) * DZSNSO(KWTD) - SMCEQ(KWTD) * ZSOIL0(IWTD) + P.SMCMAX * ZSOIL0(KWTD)) /
                      (P.SMCMAX - SMCEQ(KWTD)),
                  ZSOIL0(IWTD));
        RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ(KWTD));
      }
    } else {
      WTD = ZSOIL0(KWTD);
      RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ(KWTD));
      KWTD = KWTD + 1;
      IWTD = IWTD + 1;
      if (KWTD <= NSOIL) {
        WTDOLD = WTD;
        if (SMC(KWTD) > SMCEQ(KWTD)) {
          WTD = MIN((SMC(KWTD) * DZSNSO(KWTD) - SMCEQ(KWTD) * ZSOIL0(IWTD) + P.SMCMAX * ZSOIL0(KWTD)) /
                        (P.SMCMAX - SMCEQ(KWTD)),
                    ZSOIL0(IWTD));
        } else {
          WTD = ZSOIL0(KWTD);
        }
        RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQ(KWTD));
      } else {
        WTDOLD = WTD;
        float SMCEQDEEP = P.SMCMAX * POW(-P.PSISAT / (-P.PSISAT - DZSNSO(NSOIL)), 1.f / P.BEXP);
        WTD = MIN((SMCWTD * DZSNSO(NSOIL) - SMCEQDEEP * ZSOIL0(NSOIL) +
                   P.SMCMAX * (ZSOIL0(NSOIL) - DZSNSO(NSOIL))) /
                      (P.SMCMAX - SMCEQDEEP),
                  ZSOIL0(NSOIL));
        RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQDEEP);
      }
    }
  } else if (WTD >= ZSOIL0(NSOIL) - DZSNSO(NSOIL)) {
    WTDOLD = WTD;
    float SMCEQDEEP = P.SMCMAX * POW(-P.PSISAT / (-P.PSISAT - DZSNSO(NSOIL)), 1.f / P.BEXP);
    if (SMCWTD > SMCEQDEEP) {
      WTD = MIN((SMCWTD * DZSNSO(NSOIL) - SMCEQDEEP * ZSOIL0(NSOIL) +
                 P.SMCMAX * (ZSOIL0(NSOIL) - DZSNSO(NSOIL))) /
                    (P.SMCMAX - SMCEQDEEP),
                ZSOIL0(NSOIL));
      RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQDEEP);
    } else {
      RECH = -(WTDOLD - (ZSOIL0(NSOIL) - DZSNSO(NSOIL))) * (P.SMCMAX - SMCEQDEEP);
      WTDOLD = ZSOIL0(NSOIL) - DZSNSO(NSOIL);
      float DZUP = (SMCEQDEEP - SMCWTD) * DZSNSO(NSOIL) / (P.SMCMAX - SMCEQDEEP);
      WTD = WTDOLD - DZUP;
      RECH = RECH - (P.SMCMAX - SMCEQDEEP) * DZUP;
      SMCWTD = SMCEQDEEP;
    }
  }
  if (IWTD < NSOIL) SMCWTD = P.SMCMAX;
}

// noahmplsm.F90:6382-6613
void WATER(Ctx& c, SflxIO& s, SflxLocal& L) {
  const Params& P = c.P;
  const float WSLMAX = 5000.f;
  ASoil ETRANI, WCND;
  ETRANI.fill(0.f); WCND.fill(0.f);
  float SNOFLOW = 0.f, QINSUR = 0.f, QRAIN, SNOWHIN, QSNSUB, QSEVA, QSNFRO, QSDEW, QDRAIN = 0.f,
        FCRMAX = 0.f;
  s.RUNSUB = 0.f;

  CANWATER(c, s.VEGTYP, s.DT, s.SFCTMP, s.UU, s.VV, s.FCEV, s.FCTR, L.QPRECC, L.QPRECL, L.ELAI, L.ESAI,
           s.IST, s.TG, s.FVEG, L.FROZEN_CANOPY, s.CANLIQ, s.CANICE, s.TV, L.CMC, s.ECAN, s.ETRAN, QRAIN,
           s.QSNOW, SNOWHIN, s.FWET, s.FPICE);

  QSNSUB = 0.f;
  if (s.SNEQV > 0.f) QSNSUB = MIN(L.QVAP, s.SNEQV / s.DT);
  QSEVA = L.QVAP - QSNSUB;
  QSNFRO = 0.f;
  if (s.SNEQV > 0.f) QSNFRO = L.QDEW;
  QSDEW = L.QDEW - QSNFRO;

  SNOWWATER(L.IMELT, s.DT, s.ZSOIL, s.SFCTMP, SNOWHIN, s.QSNOW, QSNFRO, QSNSUB, QRAIN, s.FICEOLD, s.ISNOW,
            s.SNOWH, s.SNEQV, s.SNICE, s.SNLIQ, s.SH2O, L.SICE, s.STC, s.ZSNSO, L.DZSNSO, s.QSNBOT, SNOFLOW,
            s.PONDING1, s.PONDING2);

  if (L.FROZEN_GROUND) {
    L.SICE(1) = L.SICE(1) + (QSDEW - QSEVA) * s.DT / (L.DZSNSO(1) * 1000.f);
    QSDEW = 0.0f;
    QSEVA = 0.0f;
    if (L.SICE(1) < 0.f) {
      s.SH2O(1) = s.SH2O(1) + L.SICE(1);
      L.SICE(1) = 0.f;
    }
  }

  QINSUR = (s.PONDING + s.PONDING1 + s.PONDING2) / s.DT * 0.001f;
  if (s.ISNOW == 0) QINSUR = QINSUR + (s.QSNBOT + QSDEW + QRAIN) * 0.001f;
  else QINSUR = QINSUR + (s.QSNBOT + QSDEW) * 0.001f;
  QSEVA = QSEVA * 0.001f;
  for (int IZ = 1; IZ <= P.NROOT; ++IZ) ETRANI(IZ) = s.ETRAN * L.BTRANI(IZ) * 0.001f;

  if (s.IST == 2) {
    s.RUNSRF = 0.f;
    if (s.WSLAKE >= WSLMAX) s.RUNSRF = QINSUR * 1000.f;
    s.WSLAKE = s.WSLAKE + (QINSUR - QSEVA) * 1000.f * s.DT - s.RUNSRF * s.DT;
  } else {
    // the reference passes (VEGTYP, ISURBAN) into dummies named (ISURBAN, VEGTYP) (:6579 vs :7682);
    // the only use is the symmetric test VEGTYP == ISURBAN.
    SOILWATER(c, s.DT, s.ZSOIL, L.DZSNSO, QINSUR, QSEVA, ETRANI, L.SICE, s.SH2O, s.SMC, s.ZWT, s.VEGTYP,
              s.ISURBAN, s.SMCWTD, s.DEEPRECH, s.RUNSRF, QDRAIN, s.RUNSUB, WCND, FCRMAX);
    if (c.O.OPT_RUN == 1) {
      GROUNDWATER(c, s.DT, L.SICE, s.ZSOIL, WCND, FCRMAX, s.SH2O, s.ZWT, s.WA, s.WT, L.QIN, L.QDIS);
      s.RUNSUB = L.QDIS;
    }
    if (c.O.OPT_RUN == 3 || c.O.OPT_RUN == 4) s.RUNSUB = s.RUNSUB + QDRAIN;
    for (int IZ = 1; IZ <= NSOIL; ++IZ) s.SMC(IZ) = s.SH2O(IZ) + L.SICE(IZ);
    if (c.O.OPT_RUN == 5) {
      SHALLOWWATERTABLE(c, s.ZSOIL, s.DT, L.DZSNSO, s.SMCEQ, s.SMC, s.ZWT, s.SMCWTD, s.RECH, QDRAIN);
      s.SH2O(NSOIL) = s.SMC(NSOIL) - L.SICE(NSOIL);
      s.RUNSUB = s.RUNSUB + QDRAIN;
      s.WA = 0.f;
    }
  }
  s.RUNSUB = s.RUNSUB + SNOFLOW;
}

// noahmplsm.F90:8837-9104
static void CO2FLUX(Ctx& c, int VEGTYP, float IGS, float DT, const ASnSo& STC, float PSN, float TV,
                    float WROOT, float WSTRES, float FOLN, float LAPM, float& XLAI, float& XSAI,
                    float& LFMASS, float& RTMASS, float& STMASS, float& FASTCP, float& STBLCP, float& WOOD,
                    float& GPP, float& NPP, float& NEE, float& AUTORS, float& HETERS, float& TOTSC,
                    float& TOTLB) {
  const noahmp_tables& T = *c.T;
  auto R = [](float x) { return EXP(0.08f * (x - 298.16f)); };
  float RTOVRC = 2.0E-8f, RSWOODC = 3.0E-10f, BF = 0.90f, WSTRC = 100.0f, LAIMIN = 0.05f, XSAMIN = 0.01f;
  float SAPM = 3.f * 0.001f;
  float LFMSMN = LAIMIN / LAPM;
  float STMSMN = XSAMIN / SAPM;
  float RF;
  if (IGS == 0.f) RF = 0.5f; else RF = 1.0f;
  float FNF = MIN(FOLN / MAX(1.E-06f, TV1(T.folnmx, VEGTYP)), 1.0f);
  float TF = POW(TV1(T.arm, VEGTYP), (TV - 298.16f) / 10.f);
  float RESP = TV1(T.rmf25, VEGTYP) * TF * FNF * XLAI * RF * (1.f - WSTRES);
  float RSLEAF = MIN(LFMASS / DT, RESP * 12.e-6f);
  float RSROOT = TV1(T.rmr25, VEGTYP) * (RTMASS * 1E-3f) * TF * RF * 12.e-6f;
  float RSSTEM = TV1(T.rms25, VEGTYP) * (STMASS * 1E-3f) * TF * RF * 12.e-6f;
  float RSWOOD = RSWOODC * R(TV) * WOOD * TV1(T.wdpool, VEGTYP);
  float CARBFX = PSN * 12.e-6f;
  float LEAFPT = EXP(0.01f * (1.f - EXP(0.75f * XLAI)) * XLAI);
  if (VEGTYP == T.eblforest) LEAFPT = EXP(0.01f * (1.f - EXP(0.50f * XLAI)) * XLAI);
  float NONLEF = 1.0f - LEAFPT;
  float STEMPT = XLAI / 10.0f;
  LEAFPT = LEAFPT - STEMPT;
  float WOODF;
  if (WOOD > 0.f) WOODF = (1.f - EXP(-BF * (TV1(T.wrrat, VEGTYP) * RTMASS / WOOD)) / BF) * TV1(T.wdpool, VEGTYP);
  else WOODF = 0.f;
  float ROOTPT = NONLEF * (1.f - WOODF);
  float WOODPT = NONLEF * WOODF;
  float LFTOVR = TV1(T.ltovrc, VEGTYP) * 1.E-6f * LFMASS;
  float STTOVR = TV1(T.ltovrc, VEGTYP) * 1.E-6f * STMASS;
  float RTTOVR = RTOVRC * RTMASS;
  float WDTOVR = 9.5E-10f * WOOD;
  float SC = EXP(-0.3f * MAX(0.f, TV - TV1(T.tdlef, VEGTYP))) * (LFMASS / 120.f);
  float SD = EXP((WSTRES - 1.f) * WSTRC);
  float DIELF = LFMASS * 1.E-6f * (TV1(T.dilefw, VEGTYP) * SD + TV1(T.dilefc, VEGTYP) * SC);
  float DIEST = STMASS * 1.E-6f * (TV1(T.dilefw, VEGTYP) * SD + TV1(T.dilefc, VEGTYP) * SC);
  float fragr = TV1(T.fragr, VEGTYP);
  float GRLEAF = MAX(0.0f, fragr * (LEAFPT * CARBFX - RSLEAF));
  float GRSTEM = MAX(0.0f, fragr * (STEMPT * CARBFX - RSSTEM));
  float GRROOT = MAX(0.0f, fragr * (ROOTPT * CARBFX - RSROOT));
  float GRWOOD = MAX(0.0f, fragr * (WOODPT * CARBFX - RSWOOD));
  float ADDNPPLF = MAX(0.f, LEAFPT * CARBFX - GRLEAF - RSLEAF);
  float ADDNPPST = MAX(0.f, STEMPT * CARBFX - GRSTEM - RSSTEM);
  if (TV < TV1(T.tmin, VEGTYP)) ADDNPPLF = 0.f;
  if (TV < TV1(T.tmin, VEGTYP)) ADDNPPST = 0.f;
  float LFDEL = (LFMASS - LFMSMN) / DT;
  float STDEL = (STMASS - STMSMN) / DT;
  DIELF = MIN(DIELF, LFDEL + ADDNPPLF - LFTOVR);
  DIEST = MIN(DIEST, STDEL + ADDNPPST - STTOVR);
  float NPPL = MAX(ADDNPPLF, -LFDEL);
  float NPPS = MAX(ADDNPPST, -STDEL);
  float NPPR = ROOTPT * CARBFX - RSROOT - GRROOT;
  float NPPW = WOODPT * CARBFX - RSWOOD - GRWOOD;
  LFMASS = LFMASS + (NPPL - LFTOVR - DIELF) * DT;
  STMASS = STMASS + (NPPS - STTOVR - DIEST) * DT;
  RTMASS = RTMASS + (NPPR - RTTOVR) * DT;
  if (RTMASS < 0.0f) {
    RTTOVR = NPPR;
    RTMASS = 0.0f;
  }
  WOOD = (WOOD + (NPPW - WDTOVR) * DT) * TV1(T.wdpool, VEGTYP);
  FASTCP = FASTCP + (RTTOVR + LFTOVR + STTOVR + WDTOVR + DIELF) * DT;
  float FST = POW(2.0f, (STC(1) - 283.16f) / 10.f);
  float FSW = WROOT / (0.20f + WROOT) * 0.23f / (0.23f + WROOT);
  float RSSOIL = FSW * FST * TV1(T.mrp, VEGTYP) * MAX(0.f, FASTCP * 1.E-3f) * 12.E-6f;
  float STABLC = 0.1f * RSSOIL;
  FASTCP = FASTCP - (RSSOIL + STABLC) * DT;
  STBLCP = STBLCP + STABLC * DT;
  GPP = CARBFX;
  NPP = NPPL + NPPW + NPPR;
  AUTORS = RSROOT + RSWOOD + RSLEAF + GRLEAF + GRROOT + GRWOOD;
  HETERS = RSSOIL;
  NEE = (AUTORS + HETERS - GPP) * 44.f / 12.f;
  TOTSC = FASTCP + STBLCP;
  TOTLB = LFMASS + RTMASS + WOOD;
  XLAI = MAX(LFMASS * LAPM, LAIMIN);
  XSAI = MAX(STMASS * SAPM, XSAMIN);
  (void)GRSTEM; (void)NPPS;
}

// noahmplsm.F90:8723-8835
void CARBON(Ctx& c, SflxIO& s, SflxLocal& L) {
  const noahmp_tables& T = *c.T;
  const Params& P = c.P;
  if (s.VEGTYP == T.iswater || s.VEGTYP == T.isbarren || s.VEGTYP == T.issnow || s.VEGTYP == s.ISURBAN) {
    s.LAI = 0.f; s.SAI = 0.f; s.GPP = 0.f; s.NPP = 0.f; s.NEE = 0.f;
    L.AUTORS = 0.f; L.HETERS = 0.f; L.TOTSC = 0.f; L.TOTLB = 0.f;
    s.LFMASS = 0.f; s.RTMASS = 0.f; s.STMASS = 0.f; s.WOOD = 0.f; s.STBLCP = 0.f; s.FASTCP = 0.f;
    return;
  }
  float LAPM = TV1(T.sla, s.VEGTYP) / 1000.f;
  float WSTRES = 1.f - L.BTRAN;
  float WROOT = 0.f;
  for (int J = 1; J <= P.NROOT; ++J)
    WROOT = WROOT + s.SMC(J) / P.SMCMAX * L.DZSNSO(J) / (-s.ZSOIL(P.NROOT));
  CO2FLUX(c, s.VEGTYP, L.IGS, s.DT, s.STC, s.PSN, s.TV, WROOT, WSTRES, s.FOLN, LAPM, s.LAI, s.SAI,
          s.LFMASS, s.RTMASS, s.STMASS, s.FASTCP, s.STBLCP, s.WOOD, s.GPP, s.NPP, s.NEE, L.AUTORS,
          L.HETERS, L.TOTSC, L.TOTLB);
}

// noahmplsm.F90:9202-9349
int REDPRM(Ctx& c, int VEGTYP, int SOILTYP, int SLOPETYP, const ASoil& ZSOIL, int ISURBAN) {
  (void)ZSOIL;
  const noahmp_tables& T = *c.T;
  Params& P = c.P;
  if (SOILTYP > T.slcats || SOILTYP < 1) { c.fatal(NOAHMP_ERR_REDPRM, (float)SOILTYP); return 1; }
  if (VEGTYP > T.lucats || VEGTYP < 1) { c.fatal(NOAHMP_ERR_REDPRM, (float)VEGTYP); return 1; }
  P.CSOIL = T.csoil_data;
  P.BEXP = T.bb[SOILTYP - 1];
  P.DKSAT = T.satdk[SOILTYP - 1];
  P.DWSAT = T.satdw[SOILTYP - 1];
  P.F1 = T.f11[SOILTYP - 1];
  P.PSISAT = T.satpsi[SOILTYP - 1];
  P.QUARTZ = T.qtz[SOILTYP - 1];
  P.SMCDRY = T.drysmc[SOILTYP - 1];
  P.SMCMAX = T.maxsmc[SOILTYP - 1];
  P.SMCREF = T.refsmc[SOILTYP - 1];
  P.SMCWLT = T.wltsmc[SOILTYP - 1];
  if (VEGTYP == ISURBAN) {
    P.SMCMAX = 0.45f; P.SMCREF = 0.42f; P.SMCWLT = 0.40f; P.SMCDRY = 0.40f; P.CSOIL = 3.E6f;
  }
  P.ZBOT = T.zbot_data;
  P.CZIL = T.czil_data;
  float FRZK = T.frzk_data, REFDK = T.refdk_data, REFKDT = T.refkdt_data;
  P.KDT = REFKDT * P.DKSAT / REFDK;
  P.SLOPE = T.slope_data[SLOPETYP - 1];
  if (SOILTYP != 14) {
    float FRZFACT = (P.SMCMAX / P.SMCREF) * (0.412f / 0.468f);
    P.FRZX = FRZK * FRZFACT;
  } else {
    // The reference leaves FRZX at the previous column's value here (module global, :9316-9319).
    // Deliberate deviation: defined as 0 (only INFIL / opt_run=3 reads it).
    P.FRZX = 0.f;
  }
  P.TOPT = T.topt_data;
  P.RGL = T.rgltbl[VEGTYP - 1];
  P.RSMAX = T.rsmax_data;
  P.RSMIN = T.rstbl[VEGTYP - 1];
  P.HS = T.hstbl[VEGTYP - 1];
  P.NROOT = T.nrotbl[VEGTYP - 1];
  if (VEGTYP == ISURBAN) P.RSMIN = 400.0f;
  if (P.NROOT > NSOIL) { c.fatal(NOAHMP_ERR_REDPRM, (float)P.NROOT); return 1; }
  return 0;
}

}  // namespace nmo

// ---- probes for the known-answer / conservation tests of the WATER side (tests/test_oracle_water.py) ----------------
extern "C" {
// CANWATER: in = {DT SFCTMP UU VV FCEV FCTR QPRECC QPRECL ELAI ESAI TG FVEG frozen_canopy}, io = {CANLIQ CANICE TV},
// out = {CMC ECAN ETRAN QRAIN QSNOW SNOWHIN FWET FPICE}
void nmo_canwater(const noahmp_tables* T, int opt_snf, int VEGTYP, const float* in, float* io, float* out) {
  using namespace nmo;
  Ctx c{};
  c.T = T; c.O.OPT_SNF = opt_snf;
  CANWATER(c, VEGTYP, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], in[9], 1, in[10], in[11],
           in[12] != 0.f, io[0], io[1], io[2], out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
}
// SNOWWATER: sc = {DT SFCTMP SNOWHIN QSNOW QSNFRO QSNSUB QRAIN}; arrays in the oracle's Fortran bounds packed from the
// lowest index; io scalars = {SNOWH SNEQV}; out = {QSNBOT SNOFLOW PONDING1 PONDING2}
void nmo_snowwater(const int* IMELT7, const float* sc, const float* ZSOIL4, const float* FICEOLD3, int* ISNOW, float* io2,
                   float* SNICE3, float* SNLIQ3, float* SH2O4, float* SICE4, float* STC7, float* ZSNSO7, float* DZSNSO7,
                   float* out4) {
  using namespace nmo;
  IA<-NSNOW + 1, NSOIL> im;
  ASoil zs, sh, si; ASnow fo, ice, liq; ASnSo stc, zsn, dz;
  for (int k = -2; k <= NSOIL; ++k) { im(k) = IMELT7[k + 2]; stc(k) = STC7[k + 2]; zsn(k) = ZSNSO7[k + 2]; dz(k) = DZSNSO7[k + 2]; }
  for (int k = 1; k <= NSOIL; ++k) { zs(k) = ZSOIL4[k - 1]; sh(k) = SH2O4[k - 1]; si(k) = SICE4[k - 1]; }
  for (int k = -2; k <= 0; ++k) { fo(k) = FICEOLD3[k + 2]; ice(k) = SNICE3[k + 2]; liq(k) = SNLIQ3[k + 2]; }
  SNOWWATER(im, sc[0], zs, sc[1], sc[2], sc[3], sc[4], sc[5], sc[6], fo, *ISNOW, io2[0], io2[1], ice, liq, sh, si, stc,
            zsn, dz, out4[0], out4[1], out4[2], out4[3]);
  for (int k = -2; k <= NSOIL; ++k) { STC7[k + 2] = stc(k); ZSNSO7[k + 2] = zsn(k); DZSNSO7[k + 2] = dz(k); }
  for (int k = 1; k <= NSOIL; ++k) { SH2O4[k - 1] = sh(k); SICE4[k - 1] = si(k); }
  for (int k = -2; k <= 0; ++k) { SNICE3[k + 2] = ice(k); SNLIQ3[k + 2] = liq(k); }
}
// SOILWATER for one soil type: sc = {DT QINSUR QSEVA}; io = {ZWT SMCWTD DEEPRECH}; out = {RUNSRF QDRAIN RUNSUB FCRMAX}
int nmo_soilwater(const noahmp_tables* T, int opt_run, int opt_inf, int SOILTYP, const float* sc, const float* ZSOIL4,
                  const float* ETRANI4, const float* SICE4, float* SH2O4, float* SMC4, float* io3, float* out4) {
  using namespace nmo;
  Ctx c{};
  c.T = T; c.O.OPT_RUN = opt_run; c.O.OPT_INF = opt_inf;
  ASoil zs, et, si, sh, smc, wcnd;
  for (int k = 1; k <= NSOIL; ++k) { zs(k) = ZSOIL4[k - 1]; et(k) = ETRANI4[k - 1]; si(k) = SICE4[k - 1]; sh(k) = SH2O4[k - 1]; smc(k) = SMC4[k - 1]; }
  if (REDPRM(c, 7, SOILTYP, 1, zs, 1)) return 1;
  ASnSo dz; dz.fill(0.f);
  dz(1) = -zs(1);
  for (int k = 2; k <= NSOIL; ++k) dz(k) = zs(k - 1) - zs(k);
  out4[2] = 0.f;  // RUNSUB: the caller's zero survives for opt_run 3/4/5
  SOILWATER(c, sc[0], zs, dz, sc[1], sc[2], et, si, sh, smc, io3[0], 1, 7, io3[1], io3[2], out4[0], out4[1], out4[2], wcnd,
            out4[3]);
  for (int k = 1; k <= NSOIL; ++k) { SH2O4[k - 1] = sh(k); SMC4[k - 1] = smc(k); }
  return 0;
}
// CO2FLUX: sc = {IGS DT STC1 PSN TV WROOT WSTRES FOLN LAPM}; pools = {XLAI XSAI LFMASS RTMASS STMASS FASTCP STBLCP WOOD};
// out = {GPP NPP NEE AUTORS HETERS TOTSC TOTLB}
void nmo_co2flux(const noahmp_tables* T, int VEGTYP, const float* sc, float* pools, float* out7) {
  using namespace nmo;
  Ctx c{};
  c.T = T;
  ASnSo stc; stc.fill(0.f);
  stc(1) = sc[2];
  CO2FLUX(c, VEGTYP, sc[0], sc[1], stc, sc[3], sc[4], sc[5], sc[6], sc[7], sc[8], pools[0], pools[1], pools[2], pools[3],
          pools[4], pools[5], pools[6], pools[7], out7[0], out7[1], out7[2], out7[3], out7[4], out7[5], out7[6]);
}
}

