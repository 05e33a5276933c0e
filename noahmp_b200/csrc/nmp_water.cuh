// nmp_water.cuh — device code of the WATER and CARBON subtrees of NOAHMP_SFLX
// (phys/module_sf_noahmplsm.F90:6382-9104): CANWATER, SNOWWATER (SNOWFALL, COMPACT, COMBINE, DIVIDE,
// COMBO, SNOWH2O), SOILWATER (ZWTEQ, INFIL, SRT, SSTEP, WDFCND1/2), GROUNDWATER, SHALLOWWATERTABLE,
// CARBON/CO2FLUX.
//
// GPU shape: soil loops (1..4) are fully unrolled and index registers.  The snow-pack routines shift
// layers by data-dependent amounts (COMBINE/DIVIDE); they work on a private 3-layer copy (SnowPack)
// that is indexed dynamically (L1-resident local memory, 12 words) so the 7-layer column arrays of the
// caller stay in registers.
#pragma once
#include "nmp_common.cuh"

namespace nmp {

NMP_DEV float at4(const S4& a, int k) {  // a(k), k in 1..4, without dynamic register indexing
  return k == 1 ? a.v[0] : (k == 2 ? a.v[1] : (k == 3 ? a.v[2] : a.v[3]));
}

// noahmplsm.F90:6615-6865
struct CanOut {
  float CMC, ECAN, ETRAN, QRAIN, QSNOW, SNOWHIN, FPICE;
};
template <class O>
NMP_DEV void CANWATER(const Ctx& c, int VEGTYP, float DT, float SFCTMP, float UU, float VV, float FCEV,
                      float FCTR, float QPRECC, float QPRECL, float ELAI, float ESAI, int IST, float TG,
                      float FVEG, bool FROZEN_CANOPY, float& CANLIQ, float& CANICE, float& TV, float& FWET,
                      CanOut& o) {
  float FP = 0.0f, RAIN, SNOW, QINTR, QDRIPR, QTHROR, QINTS, QDRIPS, QTHROS;
  float QEVAC, QDEWC, QSUBC, QFROC, ETRAN;
  float FPICE = 0.f;
  const int snf = NMP_OPT(snf);
  if (snf == 1) {
    if (SFCTMP > TFRZ + 2.5f) {
      FPICE = 0.f;
    } else {
      if (SFCTMP <= TFRZ + 0.5f) FPICE = 1.0f;
      else if (SFCTMP <= TFRZ + 2.f) FPICE = 1.f - (-54.632f + 0.2f * SFCTMP);
      else FPICE = 0.6f;
    }
  }
  if (snf == 2) {
    if (SFCTMP >= TFRZ + 2.2f) FPICE = 0.f; else FPICE = 1.0f;
  }
  if (snf == 3) {
    if (SFCTMP >= TFRZ) FPICE = 0.f; else FPICE = 1.0f;
  }
  float BDFALL = MIN(120.f, 67.92f + 51.25f * EXP((SFCTMP - TFRZ) / 2.59f));
  RAIN = (QPRECC + QPRECL) * (1.f - FPICE);
  SNOW = (QPRECC + QPRECL) * FPICE;
  if (QPRECC + QPRECL > 0.f) FP = (QPRECC + QPRECL) / (10.f * QPRECC + QPRECL);
  float MAXLIQ = tv1(c.T->ch2op, VEGTYP) * (ELAI + ESAI);
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
  if (CANICE > 1.E-6f && TV > TFRZ) {
    float QMELTC = MIN(CANICE / DT, (TV - TFRZ) * CICE * CANICE / DENICE / (DT * HFUS));
    CANICE = MAX(0.f, CANICE - QMELTC * DT);
    CANLIQ = MAX(0.f, CANLIQ + QMELTC * DT);
    TV = FWET * TFRZ + (1.f - FWET) * TV;
  }
  if (CANLIQ > 1.E-6f && TV < TFRZ) {
    float QFRZC = MIN(CANLIQ / DT, (TFRZ - TV) * CWAT * CANLIQ / DENH2O / (DT * HFUS));
    CANLIQ = MAX(0.f, CANLIQ - QFRZC * DT);
    CANICE = MAX(0.f, CANICE + QFRZC * DT);
    TV = FWET * TFRZ + (1.f - FWET) * TV;
  }
  o.CMC = CANLIQ + CANICE;
  o.ECAN = QEVAC + QSUBC - QDEWC - QFROC;
  o.ETRAN = ETRAN;
  o.QRAIN = QDRIPR + QTHROR;
  o.QSNOW = QDRIPS + QTHROS;
  o.SNOWHIN = o.QSNOW / BDFALL;
  o.FPICE = FPICE;
  if (IST == 2 && TG > TFRZ) { o.QSNOW = 0.f; o.SNOWHIN = 0.f; }
}

// ---- snow pack: private dynamically indexed copy of the 3 snow layers -------------------------------
// Fortran index J in -2..0 maps to element J+2.  Shared by the land and glacier paths; the thresholds
// that differ between noahmplsm.F90 and glacier.F90 are template constants (SURVEY.md §8a diff table).
struct SnowPack {
  float dz[NSNOW], ice[NSNOW], liq[NSNOW], t[NSNOW];
};

// noahmplsm.F90:7375-7424 (identical in glacier.F90:2575-2624)
NMP_DEV void COMBO(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2, float WICE2, float T2) {
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

// noahmplsm.F90:6998-7063 / glacier.F90:2239-2302 (new layer at SNOWH >= 0.025 / 0.05)
template <bool GLACIER>
NMP_DEV void SNOWFALL(float DT, float QSNOW, float SNOWHIN, float SFCTMP, int& ISNOW, float& SNOWH, SnowPack& p,
                      float& SNEQV) {
  const float HNEW = GLACIER ? 0.05f : 0.025f;
  int NEWNODE = 0;
  if (ISNOW == 0 && QSNOW > 0.f) {
    SNOWH = SNOWH + SNOWHIN * DT;
    SNEQV = SNEQV + QSNOW * DT;
  }
  if (ISNOW == 0 && QSNOW > 0.f && SNOWH >= HNEW) {
    ISNOW = -1;
    NEWNODE = 1;
    p.dz[2] = SNOWH;
    SNOWH = 0.f;
    p.t[2] = MIN(273.16f, SFCTMP);
    p.ice[2] = SNEQV;
    p.liq[2] = 0.f;
  }
  if (ISNOW < 0 && NEWNODE == 0 && QSNOW > 0.f) {
    p.ice[ISNOW + 3] = p.ice[ISNOW + 3] + QSNOW * DT;
    p.dz[ISNOW + 3] = p.dz[ISNOW + 3] + SNOWHIN * DT;
  }
}

// noahmplsm.F90:7427-7528 (maths identical in glacier.F90:2304-2401)
NMP_DEV void COMPACT(float DT, const SnowPack& pc, const I7& IMELT, const N3& FICEOLD, int ISNOW, SnowPack& p) {
  const float C2 = 21.e-3f, C3 = 2.5e-6f, C4 = 0.04f, C5 = 2.0f, DM = 100.0f, ETA0 = 0.8e+6f;
  float BURDEN = 0.0f;
#pragma unroll
  for (int J = -2; J <= 0; ++J) {
    if (J > ISNOW) {
      const float ICE = pc.ice[J + 2], LIQ = pc.liq[J + 2], DZ = p.dz[J + 2];
      float WX = ICE + LIQ;
      float FICE = ICE / WX;
      float VOID = 1.f - (ICE / DENICE + LIQ / DENH2O) / DZ;
      if (VOID > 0.001f && ICE > 0.1f) {
        float BI = ICE / DZ;
        float TD = MAX(0.f, TFRZ - pc.t[J + 2]);
        float DEXPF = EXP(-C4 * TD);
        float DDZ1 = -C3 * DEXPF;
        if (BI > DM) DDZ1 = DDZ1 * EXP(-46.0E-3f * (BI - DM));
        if (LIQ > 0.01f * DZ) DDZ1 = DDZ1 * C5;
        float DDZ2 = -(BURDEN + 0.5f * WX) * EXP(-0.08f * TD - C2 * BI) / ETA0;
        float DDZ3;
        if (IMELT(J) == 1) {
          DDZ3 = MAX(0.f, (FICEOLD(J) - FICE) / MAX(1.E-6f, FICEOLD(J)));
          DDZ3 = -DDZ3 / DT;
        } else {
          DDZ3 = 0.f;
        }
        float PDZDTC = (DDZ1 + DDZ2 + DDZ3) * DT;
        PDZDTC = MAX(-0.5f, PDZDTC);
        p.dz[J + 2] = DZ * (1.f + PDZDTC);
      }
      BURDEN = BURDEN + WX;
    }
  }
}

// noahmplsm.F90:7065-7246 / glacier.F90:2403-2573.  SICE1/SH2O1/DZ1 = first soil layer (land only).
template <bool GLACIER>
NMP_DEV void COMBINE(int& ISNOW, float& SH2O1, float& SICE1, float DZ1, SnowPack& p, float& SNOWH, float& SNEQV,
                     float& PONDING1, float& PONDING2) {
  const float DZMIN0 = GLACIER ? 0.045f : 0.025f, DZMIN1 = GLACIER ? 0.05f : 0.025f,
              DZMIN2 = GLACIER ? 0.2f : 0.1f;
  const float HMIN = GLACIER ? 0.05f : 0.025f;
  int ISNOW_OLD = ISNOW;
  for (int J = ISNOW_OLD + 1; J <= 0; ++J) {
    if (p.ice[J + 2] <= .1f) {
      if (J != 0) {
        p.liq[J + 3] = p.liq[J + 3] + p.liq[J + 2];
        p.ice[J + 3] = p.ice[J + 3] + p.ice[J + 2];
      } else {
        if (ISNOW_OLD < -1) {
          p.liq[J + 1] = p.liq[J + 1] + p.liq[J + 2];
          p.ice[J + 1] = p.ice[J + 1] + p.ice[J + 2];
        } else {
          if (GLACIER) {
            PONDING1 = PONDING1 + p.liq[J + 2];
            SNEQV = p.ice[J + 2];
            SNOWH = p.dz[J + 2];
          } else {
            if (p.ice[J + 2] >= 0.f) {
              PONDING1 = p.liq[J + 2];
              SNEQV = p.ice[J + 2];
              SNOWH = p.dz[J + 2];
            } else {
              PONDING1 = p.liq[J + 2] + p.ice[J + 2];
              if (PONDING1 < 0.f) {
                SICE1 = MAX(0.0f, SICE1 + PONDING1 / (DZ1 * 1000.f));
                PONDING1 = 0.0f;
              }
              SNEQV = 0.0f;
              SNOWH = 0.0f;
            }
          }
          p.liq[J + 2] = 0.0f;
          p.ice[J + 2] = 0.0f;
          p.dz[J + 2] = 0.0f;
        }
      }
      if (J > ISNOW + 1 && ISNOW < -1) {
        for (int I = J; I >= ISNOW + 2; --I) {
          p.t[I + 2] = p.t[I + 1];
          p.liq[I + 2] = p.liq[I + 1];
          p.ice[I + 2] = p.ice[I + 1];
          p.dz[I + 2] = p.dz[I + 1];
        }
      }
      ISNOW = ISNOW + 1;
    }
  }
  if (SICE1 < 0.f) {
    SH2O1 = SH2O1 + SICE1;
    SICE1 = 0.f;
  }
  if (ISNOW == 0) return;
  SNEQV = 0.f; SNOWH = 0.f;
  float ZWICE = 0.f, ZWLIQ = 0.f;
  for (int J = ISNOW + 1; J <= 0; ++J) {
    SNEQV = SNEQV + p.ice[J + 2] + p.liq[J + 2];
    SNOWH = SNOWH + p.dz[J + 2];
    ZWICE = ZWICE + p.ice[J + 2];
    ZWLIQ = ZWLIQ + p.liq[J + 2];
  }
  if (SNOWH < HMIN && ISNOW < 0) {
    ISNOW = 0;
    SNEQV = ZWICE;
    if (GLACIER) PONDING2 = PONDING2 + ZWLIQ;
    else PONDING2 = ZWLIQ;
    if (SNEQV <= 0.f) SNOWH = 0.f;
  }
  if (ISNOW < -1) {
    ISNOW_OLD = ISNOW;
    int MSSI = 1;
    for (int I = ISNOW_OLD + 1; I <= 0; ++I) {
      const float DZMIN = MSSI == 1 ? DZMIN0 : (MSSI == 2 ? DZMIN1 : DZMIN2);
      if (p.dz[I + 2] < DZMIN) {
        int NEIBOR;
        if (I == ISNOW + 1) NEIBOR = I + 1;
        else if (I == 0) NEIBOR = I - 1;
        else {
          NEIBOR = I + 1;
          if ((p.dz[I + 1] + p.dz[I + 2]) < (p.dz[I + 3] + p.dz[I + 2])) NEIBOR = I - 1;
        }
        int J, L;
        if (NEIBOR > I) { J = NEIBOR; L = I; }
        else { J = I; L = NEIBOR; }
        COMBO(p.dz[J + 2], p.liq[J + 2], p.ice[J + 2], p.t[J + 2], p.dz[L + 2], p.liq[L + 2], p.ice[L + 2],
              p.t[L + 2]);
        if (J - 1 > ISNOW + 1) {
          for (int K = J - 1; K >= ISNOW + 2; --K) {
            p.t[K + 2] = p.t[K + 1];
            p.ice[K + 2] = p.ice[K + 1];
            p.liq[K + 2] = p.liq[K + 1];
            p.dz[K + 2] = p.dz[K + 1];
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

// noahmplsm.F90:7248-7371 / glacier.F90:2626-2749 (a 2-layer pack subdivides layer 2 above 0.20 / 0.10 m)
template <bool GLACIER>
NMP_DEV void DIVIDE(int& ISNOW, SnowPack& p) {
  const float DZ2MAX = GLACIER ? 0.10f : 0.20f;
  // top-down copies: registers (indices are static below)
  float DZ1 = 0.f, DZ2 = 0.f, DZ3 = 0.f, WI1 = 0.f, WI2 = 0.f, WI3 = 0.f, WL1 = 0.f, WL2 = 0.f, WL3 = 0.f,
        T1 = 0.f, T2 = 0.f, T3 = 0.f;
  int MSNO = -ISNOW;
  if (MSNO >= 1) { DZ1 = p.dz[ISNOW + 3]; WI1 = p.ice[ISNOW + 3]; WL1 = p.liq[ISNOW + 3]; T1 = p.t[ISNOW + 3]; }
  if (MSNO >= 2) { DZ2 = p.dz[ISNOW + 4]; WI2 = p.ice[ISNOW + 4]; WL2 = p.liq[ISNOW + 4]; T2 = p.t[ISNOW + 4]; }
  if (MSNO >= 3) { DZ3 = p.dz[ISNOW + 5]; WI3 = p.ice[ISNOW + 5]; WL3 = p.liq[ISNOW + 5]; T3 = p.t[ISNOW + 5]; }
  if (MSNO == 1) {
    if (DZ1 > 0.05f) {
      MSNO = 2;
      DZ1 = DZ1 / 2.f;
      WI1 = WI1 / 2.f;
      WL1 = WL1 / 2.f;
      DZ2 = DZ1; WI2 = WI1; WL2 = WL1; T2 = T1;
    }
  }
  if (MSNO > 1) {
    if (DZ1 > 0.05f) {
      float DRR = DZ1 - 0.05f;
      float PROPOR = DRR / DZ1;
      float ZWICE = PROPOR * WI1;
      float ZWLIQ = PROPOR * WL1;
      PROPOR = 0.05f / DZ1;
      WI1 = PROPOR * WI1;
      WL1 = PROPOR * WL1;
      DZ1 = 0.05f;
      COMBO(DZ2, WL2, WI2, T2, DRR, ZWLIQ, ZWICE, T1);
      if (MSNO <= 2 && DZ2 > DZ2MAX) {
        MSNO = 3;
        float DTDZ = (T1 - T2) / ((DZ1 + DZ2) / 2.f);
        DZ2 = DZ2 / 2.f;
        WI2 = WI2 / 2.f;
        WL2 = WL2 / 2.f;
        DZ3 = DZ2; WI3 = WI2; WL3 = WL2;
        T3 = T2 - DTDZ * DZ2 / 2.f;
        if (T3 >= TFRZ) T3 = T2;
        else T2 = T2 + DTDZ * DZ2 / 2.f;
      }
    }
  }
  if (MSNO > 2) {
    if (DZ2 > 0.2f) {  // 0.2 in both noahmplsm.F90:7351 and glacier.F90:2721
      float DRR = DZ2 - 0.2f;
      float PROPOR = DRR / DZ2;
      float ZWICE = PROPOR * WI2;
      float ZWLIQ = PROPOR * WL2;
      PROPOR = 0.2f / DZ2;
      WI2 = PROPOR * WI2;
      WL2 = PROPOR * WL2;
      DZ2 = 0.2f;
      COMBO(DZ3, WL3, WI3, T3, DRR, ZWLIQ, ZWICE, T2);
    }
  }
  ISNOW = -MSNO;
  if (MSNO >= 1) { p.dz[ISNOW + 3] = DZ1; p.ice[ISNOW + 3] = WI1; p.liq[ISNOW + 3] = WL1; p.t[ISNOW + 3] = T1; }
  if (MSNO >= 2) { p.dz[ISNOW + 4] = DZ2; p.ice[ISNOW + 4] = WI2; p.liq[ISNOW + 4] = WL2; p.t[ISNOW + 4] = T2; }
  if (MSNO >= 3) { p.dz[ISNOW + 5] = DZ3; p.ice[ISNOW + 5] = WI3; p.liq[ISNOW + 5] = WL3; p.t[ISNOW + 5] = T3; }
}

// noahmplsm.F90:7530-7678 / glacier.F90:2751-2896
template <bool GLACIER>
NMP_DEV void SNOWH2O(float DT, float QSNFRO, float QSNSUB, float QRAIN, int& ISNOW, SnowPack& p, float DZ1,
                     float& SNOWH, float& SNEQV, float& SH2O1, float& SICE1, float& QSNBOT, float& PONDING1,
                     float& PONDING2) {
  if (SNEQV == 0.f) {
    if (GLACIER) {
      SICE1 = SICE1 + (QSNFRO - QSNSUB) * DT / (DZ1 * 1000.f);
    } else {
      SICE1 = SICE1 + (QSNFRO - QSNSUB) * DT / (DZ1 * 1000.f);
      if (SICE1 < 0.f) {
        SH2O1 = SH2O1 + SICE1;
        SICE1 = 0.f;
      }
    }
  }
  if (ISNOW == 0 && SNEQV > 0.f) {
    float TEMP = SNEQV;
    SNEQV = SNEQV - QSNSUB * DT + QSNFRO * DT;
    float PROPOR = SNEQV / TEMP;
    SNOWH = MAX(0.f, PROPOR * SNOWH);
    if (SNEQV < 0.f) {
      SICE1 = SICE1 + SNEQV / (DZ1 * 1000.f);
      SNEQV = 0.f;
      SNOWH = 0.f;
    }
    if (SICE1 < 0.f) {
      SH2O1 = SH2O1 + SICE1;
      SICE1 = 0.f;
    }
  }
  if (SNOWH <= 1.E-8f || SNEQV <= 1.E-6f) {
    SNOWH = 0.0f;
    SNEQV = 0.0f;
  }
  if (ISNOW < 0) {
    float WGDIF = p.ice[ISNOW + 3] - QSNSUB * DT + QSNFRO * DT;
    p.ice[ISNOW + 3] = WGDIF;
    if (WGDIF < 1.e-6f && ISNOW < 0)
      COMBINE<GLACIER>(ISNOW, SH2O1, SICE1, DZ1, p, SNOWH, SNEQV, PONDING1, PONDING2);
    if (ISNOW < 0) {
      p.liq[ISNOW + 3] = p.liq[ISNOW + 3] + QRAIN * DT;
      p.liq[ISNOW + 3] = MAX(0.f, p.liq[ISNOW + 3]);
    }
  }
  float VOL_LIQ[NSNOW], VOL_ICE[NSNOW], EPORE[NSNOW];
#pragma unroll
  for (int J = 0; J < NSNOW; ++J) {
    VOL_LIQ[J] = 0.f; VOL_ICE[J] = 0.f; EPORE[J] = 0.f;
    if (J - 2 >= ISNOW + 1) {
      VOL_ICE[J] = MIN(1.f, p.ice[J] / (p.dz[J] * DENICE));
      EPORE[J] = 1.f - VOL_ICE[J];
      VOL_LIQ[J] = MIN(EPORE[J], p.liq[J] / (p.dz[J] * DENH2O));
    }
  }
  float QIN = 0.f, QOUT = 0.f;
#pragma unroll
  for (int J = 0; J < NSNOW; ++J) {
    if (J - 2 >= ISNOW + 1) {
      p.liq[J] = p.liq[J] + QIN;
      if (J <= 1) {
        if (EPORE[J] < 0.05f || EPORE[J + (J < 2 ? 1 : 0)] < 0.05f) {
          QOUT = 0.f;
        } else {
          QOUT = MAX(0.f, (VOL_LIQ[J] - SSI * EPORE[J]) * p.dz[J]);
          QOUT = MIN(QOUT, (1.f - VOL_ICE[J + (J < 2 ? 1 : 0)] - VOL_LIQ[J + (J < 2 ? 1 : 0)]) *
                               p.dz[J + (J < 2 ? 1 : 0)]);
        }
      } else {
        QOUT = MAX(0.f, (VOL_LIQ[J] - SSI * EPORE[J]) * p.dz[J]);
      }
      QOUT = QOUT * 1000.f;
      p.liq[J] = p.liq[J] - QOUT;
      QIN = QOUT;
    }
  }
  QSNBOT = QOUT / DT;
}

// noahmplsm.F90:6868-6996.  STC/SNICE/SNLIQ/DZSNSO/ZSNSO are the caller's register arrays.
NMP_DEV void SNOWWATER(const I7& IMELT, float DT, const S4& ZSOIL, float SFCTMP, float SNOWHIN, float QSNOW,
                       float QSNFRO, float QSNSUB, float QRAIN, const N3& FICEOLD, int& ISNOW, float& SNOWH,
                       float& SNEQV, N3& SNICE, N3& SNLIQ, float& SH2O1, float& SICE1, L7& STC, L7& ZSNSO,
                       L7& DZSNSO, float& QSNBOT, float& SNOFLOW, float& PONDING1, float& PONDING2) {
  SNOFLOW = 0.0f; PONDING1 = 0.0f; PONDING2 = 0.0f;
  QSNBOT = 0.f;
  const float DZ1 = DZSNSO(1);
  if (ISNOW == 0 && !(QSNOW > 0.f)) {
    // no snow layers and no snowfall: only the bulk-snow branch of SNOWH2O can act
    SnowPack p;  // never indexed on this path
#pragma unroll
    for (int J = 0; J < NSNOW; ++J) { p.dz[J] = 0.f; p.ice[J] = 0.f; p.liq[J] = 0.f; p.t[J] = 0.f; }
    SNOWH2O<false>(DT, QSNFRO, QSNSUB, QRAIN, ISNOW, p, DZ1, SNOWH, SNEQV, SH2O1, SICE1, QSNBOT, PONDING1,
                   PONDING2);
  } else {
    SnowPack p;
#pragma unroll
    for (int J = -2; J <= 0; ++J) {
      p.dz[J + 2] = DZSNSO(J); p.ice[J + 2] = SNICE(J); p.liq[J + 2] = SNLIQ(J); p.t[J + 2] = STC(J);
    }
    SNOWFALL<false>(DT, QSNOW, SNOWHIN, SFCTMP, ISNOW, SNOWH, p, SNEQV);
    if (ISNOW < 0) {
      SnowPack pc = p;
      COMPACT(DT, pc, IMELT, FICEOLD, ISNOW, p);
    }
    if (ISNOW < 0) COMBINE<false>(ISNOW, SH2O1, SICE1, DZ1, p, SNOWH, SNEQV, PONDING1, PONDING2);
    if (ISNOW < 0) DIVIDE<false>(ISNOW, p);
    SNOWH2O<false>(DT, QSNFRO, QSNSUB, QRAIN, ISNOW, p, DZ1, SNOWH, SNEQV, SH2O1, SICE1, QSNBOT, PONDING1,
                   PONDING2);
#pragma unroll
    for (int J = -2; J <= 0; ++J) {
      DZSNSO(J) = p.dz[J + 2]; SNICE(J) = p.ice[J + 2]; SNLIQ(J) = p.liq[J + 2]; STC(J) = p.t[J + 2];
    }
  }
#pragma unroll
  for (int IZ = -2; IZ <= 0; ++IZ) {
    if (IZ <= ISNOW) { SNICE(IZ) = 0.f; SNLIQ(IZ) = 0.f; STC(IZ) = 0.f; DZSNSO(IZ) = 0.f; ZSNSO(IZ) = 0.f; }
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

// noahmplsm.F90:8329-8362
NMP_DEV void WDFCND1(const Prm& P, float& WDF, float& WCND, float SMC, float FCR) {
  float FACTR = MAX(0.01f, SMC / P.SMCMAX);
  float EXPON = P.BEXP + 2.0f;
  WDF = P.DWSAT * POW(FACTR, EXPON);
  WDF = WDF * (1.0f - FCR);
  EXPON = 2.0f * P.BEXP + 3.0f;
  WCND = P.DKSAT * POW(FACTR, EXPON);
  WCND = WCND * (1.0f - FCR);
}

// noahmplsm.F90:8364-8400
NMP_DEV void WDFCND2(const Prm& P, float& WDF, float& WCND, float SMC, float SICE) {
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
NMP_DEV float ZWTEQ(const Prm& P, const S4& ZSOIL, const L7& DZSNSO, const S4& SH2O) {
  const int NFINE = 100;
  float WD1 = 0.f;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) WD1 = WD1 + (P.SMCMAX - SH2O(K)) * DZSNSO(K);
  float DZFINE = 3.0f * (-ZSOIL(NSOIL)) / (float)NFINE;
  float ZWT = -3.f * ZSOIL(NSOIL) - 0.001f;
  float WD2 = 0.f;
#pragma unroll 1
  for (int K = 1; K <= NFINE; ++K) {
    float ZFINE = (float)K * DZFINE;
    float TEMP = 1.f + (ZWT - ZFINE) / P.PSISAT;
    WD2 = WD2 + P.SMCMAX * (1.f - POW(TEMP, -1.f / P.BEXP)) * DZFINE;
    if (ABS(WD2 - WD1) <= 0.01f) {
      ZWT = ZFINE;
      break;
    }
  }
  return ZWT;
}

// noahmplsm.F90:7992-8087
NMP_DEV void INFIL(const Prm& P, float DT, const S4& ZSOIL, const S4& SH2O, const S4& SICE, float SICEMAX,
                   float QINSUR, float& PDDUM, float& RUNSRF) {
  const int CVFRZ = 3;
  if (QINSUR > 0.0f) {
    float DT1 = DT / 86400.f;
    float SMCAV = P.SMCMAX - P.SMCWLT;
    float DMAX = -ZSOIL(1) * SMCAV;
    float DICE = -ZSOIL(1) * SICE(1);
    DMAX = DMAX * (1.0f - (SH2O(1) + SICE(1) - P.SMCWLT) / SMCAV);
    float DD = DMAX;
#pragma unroll
    for (int K = 2; K <= NSOIL; ++K) {
      DICE = DICE + (ZSOIL(K - 1) - ZSOIL(K)) * SICE(K);
      DMAX = (ZSOIL(K - 1) - ZSOIL(K)) * SMCAV;
      DMAX = DMAX * (1.0f - (SH2O(K) + SICE(K) - P.SMCWLT) / SMCAV);
      DD = DD + DMAX;
    }
    float VAL = (1.f - EXP(-P.KDT * DT1));
    float DDT = DD * VAL;
    float PX = MAX(0.f, QINSUR * DT);
    float INFMAX = (PX * (DDT / (PX + DDT))) / DT;
    float FCR = 1.f;
    if (DICE > 1.E-2f) {
      float ACRT = (float)CVFRZ * P.FRZX / DICE;
      // SUM = 1 + ACRT**2/2 + ACRT**1/1  (J=1: K=2; J=2: K=1)
      float SUM = 1.f;
      SUM = SUM + POW2(ACRT) / 2.f;
      SUM = SUM + ACRT / 1.f;
      FCR = 1.f - EXP(-ACRT) * SUM;
    }
    INFMAX = INFMAX * FCR;
    float WDF, WCND;
    WDFCND2(P, WDF, WCND, SH2O(1), SICEMAX);
    INFMAX = MAX(INFMAX, WCND);
    INFMAX = MIN(INFMAX, PX);
    RUNSRF = MAX(0.f, QINSUR - INFMAX);
    PDDUM = QINSUR - RUNSRF;
  }
}

// noahmplsm.F90:8089-8217
template <class O>
NMP_DEV void SRT(const Ctx& c, const S4& ZSOIL, float PDDUM, const S4& ETRANI, float QSEVA, const S4& SH2O,
                 const S4& SMC, float ZWT, const S4& FCR, float SICEMAX, float FCRMAX, float SMCWTD, S4& RHSTT,
                 S4& AI, S4& BI, S4& CI, float& QDRAIN, S4& WCND) {
  const Prm& P = c.P;
  const int inf = NMP_OPT(inf), run = NMP_OPT(run);
  S4 DDZ, DENOM, DSMDZ, WFLUX, WDF, SMX;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) { DDZ(K) = 0.f; DENOM(K) = 0.f; DSMDZ(K) = 0.f; WFLUX(K) = 0.f; WDF(K) = 0.f; SMX(K) = 0.f; }
  float SMXWTD = 0.f;
  if (inf == 1) {
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) {
      WDFCND1(P, WDF(K), WCND(K), SMC(K), FCR(K));
      SMX(K) = SMC(K);
    }
    if (run == 5) SMXWTD = SMCWTD;
  }
  if (inf == 2) {
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) {
      WDFCND2(P, WDF(K), WCND(K), SH2O(K), SICEMAX);
      SMX(K) = SH2O(K);
    }
    if (run == 5) SMXWTD = SMCWTD * SH2O(NSOIL) / SMC(NSOIL);
  }
#pragma unroll
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
      if (run == 1 || run == 2) QDRAIN = 0.f;
      if (run == 3) QDRAIN = P.SLOPE * WCND(K);
      if (run == 4) QDRAIN = (1.0f - FCRMAX) * WCND(K);
      if (run == 5) {
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
#pragma unroll
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

// ROSR12 (noahmplsm.F90:5979-6036) on a 4-layer system, layers 1..NSOIL.  Returns the solution in CI
// and the forward-sweep DELTA in RHSTT, as SSTEP's call does (:8270).
NMP_DEV void ROSR12_SOIL(S4& P, const S4& A, const S4& B, const S4& C, const S4& D, S4& DELTA) {
  P(1) = -C(1) / B(1);
  DELTA(1) = D(1) / B(1);
#pragma unroll
  for (int K = 2; K <= NSOIL; ++K) {
    const float CK = (K == NSOIL) ? 0.0f : C(K);
    P(K) = -CK * (1.0f / (B(K) + A(K) * P(K - 1)));
    DELTA(K) = (D(K) - A(K) * DELTA(K - 1)) * (1.0f / (B(K) + A(K) * P(K - 1)));
  }
  P(NSOIL) = DELTA(NSOIL);
#pragma unroll
  for (int KK = NSOIL - 1; KK >= 1; --KK) P(KK) = P(KK) * P(KK + 1) + DELTA(KK);
}

// noahmplsm.F90:8220-8327
template <class O>
NMP_DEV void SSTEP(const Ctx& c, float DT, const S4& ZSOIL, const L7& DZSNSO, const S4& SICE, float ZWT, S4& SH2O,
                   S4& SMC, S4& AI, S4& BI, S4& CI, S4& RHSTT, float& SMCWTD, float& QDRAIN, float& DEEPRECH,
                   float& WPLUS) {
  const Prm& P = c.P;
  WPLUS = 0.0f;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    RHSTT(K) = RHSTT(K) * DT;
    AI(K) = AI(K) * DT;
    BI(K) = 1.f + BI(K) * DT;
    CI(K) = CI(K) * DT;
  }
  S4 SOL, DELTA;
  ROSR12_SOIL(SOL, AI, BI, CI, RHSTT, DELTA);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) SH2O(K) = SH2O(K) + SOL(K);
  if (NMP_OPT(run) == 5) {
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
#pragma unroll
  for (int K = NSOIL; K >= 2; --K) {
    float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(K)));
    WPLUS = MAX((SH2O(K) - EPORE), 0.0f) * DZSNSO(K);
    SH2O(K) = MIN(EPORE, SH2O(K));
    SH2O(K - 1) = SH2O(K - 1) + WPLUS / DZSNSO(K - 1);
  }
  float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(1)));
  WPLUS = MAX((SH2O(1) - EPORE), 0.0f) * DZSNSO(1);
  SH2O(1) = MIN(EPORE, SH2O(1));
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) SMC(K) = SH2O(K) + SICE(K);
}

// noahmplsm.F90:7680-7936
template <class O>
NMP_DEV void SOILWATER(const Ctx& c, float DT, const S4& ZSOIL, const L7& DZSNSO, float QINSUR, float QSEVA,
                       const S4& ETRANI, const S4& SICE, S4& SH2O, S4& SMC, float& ZWT, bool URBAN, float& SMCWTD,
                       float& DEEPRECH, float& RUNSRF, float& QDRAIN, float& RUNSUB, S4& WCND, float& FCRMAX) {
  const Prm& P = c.P;
  const int run = NMP_OPT(run), inf = NMP_OPT(inf);
  const float A = 4.0f;
  S4 RHSTT, AI, BI, CI, FCR;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) { RHSTT(K) = 0.f; AI(K) = 0.f; BI(K) = 0.f; CI(K) = 0.f; }
  RUNSRF = 0.0f;
  float PDDUM = 0.0f, RSAT = 0.0f, FSAT, FFF, RSBMX, WPLUS = 0.f;
  QDRAIN = 0.f;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    float EPORE = MAX(1.E-4f, (P.SMCMAX - SICE(K)));
    RSAT = RSAT + MAX(0.f, SH2O(K) - EPORE) * DZSNSO(K);
    SH2O(K) = MIN(EPORE, SH2O(K));
  }
  const float EXPA = EXP(-A);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    float FICE = MIN(1.0f, SICE(K) / P.SMCMAX);
    FCR(K) = MAX(0.0f, EXP(-A * (1.f - FICE)) - EXPA) / (1.0f - EXPA);
  }
  float SICEMAX = 0.0f;
  FCRMAX = 0.0f;
  float SH2OMIN = P.SMCMAX;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    if (SICE(K) > SICEMAX) SICEMAX = SICE(K);
    if (FCR(K) > FCRMAX) FCRMAX = FCR(K);
    if (SH2O(K) < SH2OMIN) SH2OMIN = SH2O(K);
  }
  if (run == 2) {
    FFF = 2.0f;
    RSBMX = 4.0f;
    ZWT = ZWTEQ(P, ZSOIL, DZSNSO, SH2O);
    RUNSUB = (1.0f - FCRMAX) * RSBMX * EXP(-TIMEAN) * EXP(-FFF * ZWT);
  }
  if (URBAN) FCR(1) = 0.95f;
  if (run == 1) {
    FFF = 6.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * (ZWT - 2.0f));
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (run == 5) {
    FFF = 6.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * MAX(-2.0f - ZWT, 0.f));
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (run == 2) {
    FFF = 2.0f;
    FSAT = FSATMX * EXP(-0.5f * FFF * ZWT);
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  if (run == 3) INFIL(P, DT, ZSOIL, SH2O, SICE, SICEMAX, QINSUR, PDDUM, RUNSRF);
  if (run == 4) {
    float SMCTOT = 0.f, DZTOT = 0.f;
    bool done = false;
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) {
      if (!done) {
        DZTOT = DZTOT + DZSNSO(K);
        SMCTOT = SMCTOT + SMC(K) * DZSNSO(K);
        if (DZTOT >= 2.0f) done = true;
      }
    }
    SMCTOT = SMCTOT / DZTOT;
    FSAT = POW(MAX(0.01f, SMCTOT / P.SMCMAX), 4.f);
    if (QINSUR > 0.f) {
      RUNSRF = QINSUR * ((1.0f - FCR(1)) * FSAT + FCR(1));
      PDDUM = QINSUR - RUNSRF;
    }
  }
  int NITER = 1;
  if (inf == 1) {
    NITER = 3;
    if (PDDUM * DT > DZSNSO(1) * P.SMCMAX) NITER = NITER * 2;
  }
  float DTFINE = DT / (float)NITER;
  float QDRAIN_SAVE = 0.0f;
#pragma unroll 1
  for (int ITER = 1; ITER <= NITER; ++ITER) {
    SRT<O>(c, ZSOIL, PDDUM, ETRANI, QSEVA, SH2O, SMC, ZWT, FCR, SICEMAX, FCRMAX, SMCWTD, RHSTT, AI, BI, CI, QDRAIN,
           WCND);
    SSTEP<O>(c, DTFINE, ZSOIL, DZSNSO, SICE, ZWT, SH2O, SMC, AI, BI, CI, RHSTT, SMCWTD, QDRAIN, DEEPRECH, WPLUS);
    RSAT = RSAT + WPLUS;
    QDRAIN_SAVE = QDRAIN_SAVE + QDRAIN;
  }
  QDRAIN = QDRAIN_SAVE / (float)NITER;
  RUNSRF = RUNSRF * 1000.f + RSAT * 1000.f / DT;
  QDRAIN = QDRAIN * 1000.f;
  if (run == 2) {
    float WTSUB = 0.f;
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) WTSUB = WTSUB + WCND(K) * DZSNSO(K);
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) {
      float MH2O = RUNSUB * DT * (WCND(K) * DZSNSO(K)) / WTSUB;
      SH2O(K) = SH2O(K) - MH2O / (DZSNSO(K) * 1000.f);
    }
  }
  if (run != 1) {
    S4 MLIQ;
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) MLIQ(IZ) = SH2O(IZ) * DZSNSO(IZ) * 1000.f;
    float WATMIN = 0.01f, XS;
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL - 1; ++IZ) {
      if (MLIQ(IZ) < 0.f) XS = WATMIN - MLIQ(IZ);
      else XS = 0.f;
      MLIQ(IZ) = MLIQ(IZ) + XS;
      MLIQ(IZ + 1) = MLIQ(IZ + 1) - XS;
    }
    if (MLIQ(NSOIL) < WATMIN) XS = WATMIN - MLIQ(NSOIL);
    else XS = 0.f;
    MLIQ(NSOIL) = MLIQ(NSOIL) + XS;
    RUNSUB = RUNSUB - XS / DT;
    if (run == 5) DEEPRECH = DEEPRECH - XS * 1.E-3f;
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) SH2O(IZ) = MLIQ(IZ) / (DZSNSO(IZ) * 1000.f);
  }
}

// noahmplsm.F90:8403-8585
NMP_DEV void GROUNDWATER(const Prm& P, float DT, const S4& SICE, const S4& ZSOIL, const S4& WCND, float FCRMAX,
                         S4& SH2O, float& ZWT, float& WA, float& WT, float& QIN, float& QDIS) {
  const float ROUS = 0.2f, CMIC = 0.20f;
  S4 DZMM, ZNODE, MLIQ, EPORE, HK, SMC;
  DZMM(1) = -ZSOIL(1) * 1.E3f;
#pragma unroll
  for (int IZ = 2; IZ <= NSOIL; ++IZ) DZMM(IZ) = 1.E3f * (ZSOIL(IZ - 1) - ZSOIL(IZ));
  ZNODE(1) = -ZSOIL(1) / 2.f;
#pragma unroll
  for (int IZ = 2; IZ <= NSOIL; ++IZ) ZNODE(IZ) = -ZSOIL(IZ - 1) + 0.5f * (ZSOIL(IZ - 1) - ZSOIL(IZ));
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) {
    SMC(IZ) = SH2O(IZ) + SICE(IZ);
    MLIQ(IZ) = SH2O(IZ) * DZMM(IZ);
    EPORE(IZ) = MAX(0.01f, P.SMCMAX - SICE(IZ));
    HK(IZ) = 1.E3f * WCND(IZ);
  }
  int IWT = NSOIL;
  {
    bool found = false;
#pragma unroll
    for (int IZ = 2; IZ <= NSOIL; ++IZ) {
      if (!found && ZWT <= -ZSOIL(IZ)) { IWT = IZ - 1; found = true; }
    }
  }
  const float FFF = 6.0f, RSBMX = 5.0f;
  QDIS = (1.0f - FCRMAX) * RSBMX * EXP(-TIMEAN) * EXP(-FFF * (ZWT - 2.0f));
  // S_NODE is REAL(KIND=8) in the reference (:8443): the pow runs in fp64
  double S_NODE = (double)MIN(1.0f, at4(SMC, IWT) / P.SMCMAX);
  { double lo = (double)0.01f; if (lo > S_NODE) S_NODE = lo; }
  float SMPFZ = (float)(-((double)(P.PSISAT * 1000.f) * DPOW(S_NODE, (double)(-P.BEXP))));
  SMPFZ = MAX(-120000.0f, CMIC * SMPFZ);
  float KA = at4(HK, IWT);
  float WH_ZWT = -ZWT * 1.E3f;
  float WH = SMPFZ - at4(ZNODE, IWT) * 1.E3f;
  QIN = -KA * (WH_ZWT - WH) / ((ZWT - at4(ZNODE, IWT)) * 1.E3f);
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
#pragma unroll
      for (int IZ = 3; IZ <= NSOIL; ++IZ)
        if (IZ >= IWT + 2) WS = WS + EPORE(IZ) * DZMM(IZ);
      ZWT = -at4(ZSOIL, IWT + 1) - (WT - ROUS * 1000.f * 25.f - WS) / (at4(EPORE, IWT + 1)) / 1000.f;
    }
    float WTSUB = 0.f;
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) WTSUB = WTSUB + HK(IZ) * DZMM(IZ);
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) MLIQ(IZ) = MLIQ(IZ) - QDIS * DT * HK(IZ) * DZMM(IZ) / WTSUB;
  }
  ZWT = MAX(1.5f, ZWT);
  float WATMIN = 0.01f, XS;
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL - 1; ++IZ) {
    if (MLIQ(IZ) < 0.f) XS = WATMIN - MLIQ(IZ);
    else XS = 0.f;
    MLIQ(IZ) = MLIQ(IZ) + XS;
    MLIQ(IZ + 1) = MLIQ(IZ + 1) - XS;
  }
  if (MLIQ(NSOIL) < WATMIN) XS = WATMIN - MLIQ(NSOIL);
  else XS = 0.f;
  MLIQ(NSOIL) = MLIQ(NSOIL) + XS;
  WA = WA - XS;
  WT = WT - XS;
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) SH2O(IZ) = MLIQ(IZ) / DZMM(IZ);
}

// noahmplsm.F90:8588-8718.  Layer walk with data-dependent indices: done on a 5-entry local copy.
NMP_DEV void SHALLOWWATERTABLE(const Prm& P, const S4& ZSOIL, const L7& DZSNSO, const S4& SMCEQ_, const S4& SMC_,
                               float& WTD, float& SMCWTD, float& RECH) {
  float ZSOIL0[NSOIL + 1], SMC[NSOIL + 1], SMCEQ[NSOIL + 1], DZ[NSOIL + 1];
  ZSOIL0[0] = 0.f; SMC[0] = 0.f; SMCEQ[0] = 0.f; DZ[0] = 0.f;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) { ZSOIL0[K] = ZSOIL(K); SMC[K] = SMC_(K); SMCEQ[K] = SMCEQ_(K); DZ[K] = DZSNSO(K); }
  const float DZN = DZ[NSOIL];
  int IZ;
  for (IZ = NSOIL; IZ >= 1; --IZ) {
    if (WTD + 1.E-6f < ZSOIL0[IZ]) break;
  }
  int IWTD = IZ;
  int KWTD = IWTD + 1;
  float WTDOLD;
  if (KWTD <= NSOIL) {
    WTDOLD = WTD;
    if (SMC[KWTD] > SMCEQ[KWTD]) {
      if (SMC[KWTD] == P.SMCMAX) {
        WTD = ZSOIL0[IWTD];
        RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ[KWTD]);
        IWTD = IWTD - 1;
        KWTD = KWTD - 1;
        if (KWTD >= 1) {
          if (SMC[KWTD] > SMCEQ[KWTD]) {
            WTDOLD = WTD;
            WTD = MIN((SMC[KWTD] * DZ[KWTD] - SMCEQ[KWTD] * ZSOIL0[IWTD] + P.SMCMAX * ZSOIL0[KWTD]) /
                          (P.SMCMAX - SMCEQ[KWTD]),
                      ZSOIL0[IWTD]);
            RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQ[KWTD]);
          }
        }
      } else {
        WTD = MIN((SMC[KWTD] * DZ[KWTD] - SMCEQ[KWTD] * ZSOIL0[IWTD] + P.SMCMAX * ZSOIL0[KWTD]) /
                      (P.SMCMAX - SMCEQ[KWTD]),
                  ZSOIL0[IWTD]);
        RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ[KWTD]);
      }
    } else {
      WTD = ZSOIL0[KWTD];
      RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQ[KWTD]);
      KWTD = KWTD + 1;
      IWTD = IWTD + 1;
      if (KWTD <= NSOIL) {
        WTDOLD = WTD;
        if (SMC[KWTD] > SMCEQ[KWTD]) {
          WTD = MIN((SMC[KWTD] * DZ[KWTD] - SMCEQ[KWTD] * ZSOIL0[IWTD] + P.SMCMAX * ZSOIL0[KWTD]) /
                        (P.SMCMAX - SMCEQ[KWTD]),
                    ZSOIL0[IWTD]);
        } else {
          WTD = ZSOIL0[KWTD];
        }
        RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQ[KWTD]);
      } else {
        WTDOLD = WTD;
        float SMCEQDEEP = P.SMCMAX * POW(-P.PSISAT / (-P.PSISAT - DZN), 1.f / P.BEXP);
        WTD = MIN((SMCWTD * DZN - SMCEQDEEP * ZSOIL0[NSOIL] + P.SMCMAX * (ZSOIL0[NSOIL] - DZN)) /
                      (P.SMCMAX - SMCEQDEEP),
                  ZSOIL0[NSOIL]);
        RECH = RECH - (WTDOLD - WTD) * (P.SMCMAX - SMCEQDEEP);
      }
    }
  } else if (WTD >= ZSOIL0[NSOIL] - DZN) {
    WTDOLD = WTD;
    float SMCEQDEEP = P.SMCMAX * POW(-P.PSISAT / (-P.PSISAT - DZN), 1.f / P.BEXP);
    if (SMCWTD > SMCEQDEEP) {
      WTD = MIN((SMCWTD * DZN - SMCEQDEEP * ZSOIL0[NSOIL] + P.SMCMAX * (ZSOIL0[NSOIL] - DZN)) /
                    (P.SMCMAX - SMCEQDEEP),
                ZSOIL0[NSOIL]);
      RECH = -(WTDOLD - WTD) * (P.SMCMAX - SMCEQDEEP);
    } else {
      RECH = -(WTDOLD - (ZSOIL0[NSOIL] - DZN)) * (P.SMCMAX - SMCEQDEEP);
      WTDOLD = ZSOIL0[NSOIL] - DZN;
      float DZUP = (SMCEQDEEP - SMCWTD) * DZN / (P.SMCMAX - SMCEQDEEP);
      WTD = WTDOLD - DZUP;
      RECH = RECH - (P.SMCMAX - SMCEQDEEP) * DZUP;
      SMCWTD = SMCEQDEEP;
    }
  }
  if (IWTD < NSOIL) SMCWTD = P.SMCMAX;
}

// noahmplsm.F90:8837-9104
struct CarbonState {
  float LFMASS, RTMASS, STMASS, WOOD, STBLCP, FASTCP, LAI, SAI, GPP, NPP, NEE;
};
NMP_DEV void CO2FLUX(const Ctx& c, int VEGTYP, float IGS, float DT, float STC1, float PSN, float TV, float WROOT,
                     float WSTRES, float FOLN, float LAPM, CarbonState& s) {
  const noahmp_tables& T = *c.T;
  const float RTOVRC = 2.0E-8f, RSWOODC = 3.0E-10f, BF = 0.90f, WSTRC = 100.0f, LAIMIN = 0.05f, XSAMIN = 0.01f;
  const float SAPM = 3.f * 0.001f;
  float XLAI = s.LAI, LFMASS = s.LFMASS, RTMASS = s.RTMASS, STMASS = s.STMASS, WOOD = s.WOOD, FASTCP = s.FASTCP,
        STBLCP = s.STBLCP;
  float LFMSMN = LAIMIN / LAPM;
  float STMSMN = XSAMIN / SAPM;
  float RF;
  if (IGS == 0.f) RF = 0.5f; else RF = 1.0f;
  float FNF = MIN(FOLN / MAX(1.E-06f, tv1(T.folnmx, VEGTYP)), 1.0f);
  float TF = POW(tv1(T.arm, VEGTYP), (TV - 298.16f) / 10.f);
  float RESP = tv1(T.rmf25, VEGTYP) * TF * FNF * XLAI * RF * (1.f - WSTRES);
  float RSLEAF = MIN(LFMASS / DT, RESP * 12.e-6f);
  float RSROOT = tv1(T.rmr25, VEGTYP) * (RTMASS * 1E-3f) * TF * RF * 12.e-6f;
  float RSSTEM = tv1(T.rms25, VEGTYP) * (STMASS * 1E-3f) * TF * RF * 12.e-6f;
  float RSWOOD = RSWOODC * EXP(0.08f * (TV - 298.16f)) * WOOD * tv1(T.wdpool, VEGTYP);
  float CARBFX = PSN * 12.e-6f;
  float LEAFPT = EXP(0.01f * (1.f - EXP(0.75f * XLAI)) * XLAI);
  if (VEGTYP == T.eblforest) LEAFPT = EXP(0.01f * (1.f - EXP(0.50f * XLAI)) * XLAI);
  float NONLEF = 1.0f - LEAFPT;
  float STEMPT = XLAI / 10.0f;
  LEAFPT = LEAFPT - STEMPT;
  float WOODF;
  if (WOOD > 0.f) WOODF = (1.f - EXP(-BF * (tv1(T.wrrat, VEGTYP) * RTMASS / WOOD)) / BF) * tv1(T.wdpool, VEGTYP);
  else WOODF = 0.f;
  float ROOTPT = NONLEF * (1.f - WOODF);
  float WOODPT = NONLEF * WOODF;
  float LFTOVR = tv1(T.ltovrc, VEGTYP) * 1.E-6f * LFMASS;
  float STTOVR = tv1(T.ltovrc, VEGTYP) * 1.E-6f * STMASS;
  float RTTOVR = RTOVRC * RTMASS;
  float WDTOVR = 9.5E-10f * WOOD;
  float SC = EXP(-0.3f * MAX(0.f, TV - tv1(T.tdlef, VEGTYP))) * (LFMASS / 120.f);
  float SD = EXP((WSTRES - 1.f) * WSTRC);
  float DIELF = LFMASS * 1.E-6f * (tv1(T.dilefw, VEGTYP) * SD + tv1(T.dilefc, VEGTYP) * SC);
  float DIEST = STMASS * 1.E-6f * (tv1(T.dilefw, VEGTYP) * SD + tv1(T.dilefc, VEGTYP) * SC);
  float fragr = tv1(T.fragr, VEGTYP);
  float GRLEAF = MAX(0.0f, fragr * (LEAFPT * CARBFX - RSLEAF));
  float GRSTEM = MAX(0.0f, fragr * (STEMPT * CARBFX - RSSTEM));
  float GRROOT = MAX(0.0f, fragr * (ROOTPT * CARBFX - RSROOT));
  float GRWOOD = MAX(0.0f, fragr * (WOODPT * CARBFX - RSWOOD));
  float ADDNPPLF = MAX(0.f, LEAFPT * CARBFX - GRLEAF - RSLEAF);
  float ADDNPPST = MAX(0.f, STEMPT * CARBFX - GRSTEM - RSSTEM);
  if (TV < tv1(T.tmin, VEGTYP)) ADDNPPLF = 0.f;
  if (TV < tv1(T.tmin, VEGTYP)) ADDNPPST = 0.f;
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
  WOOD = (WOOD + (NPPW - WDTOVR) * DT) * tv1(T.wdpool, VEGTYP);
  FASTCP = FASTCP + (RTTOVR + LFTOVR + STTOVR + WDTOVR + DIELF) * DT;
  float FST = POW(2.0f, (STC1 - 283.16f) / 10.f);
  float FSW = WROOT / (0.20f + WROOT) * 0.23f / (0.23f + WROOT);
  float RSSOIL = FSW * FST * tv1(T.mrp, VEGTYP) * MAX(0.f, FASTCP * 1.E-3f) * 12.E-6f;
  float STABLC = 0.1f * RSSOIL;
  FASTCP = FASTCP - (RSSOIL + STABLC) * DT;
  STBLCP = STBLCP + STABLC * DT;
  s.GPP = CARBFX;
  s.NPP = NPPL + NPPW + NPPR;
  float AUTORS = RSROOT + RSWOOD + RSLEAF + GRLEAF + GRROOT + GRWOOD;
  float HETERS = RSSOIL;
  s.NEE = (AUTORS + HETERS - s.GPP) * 44.f / 12.f;
  s.LAI = MAX(LFMASS * LAPM, LAIMIN);
  s.SAI = MAX(STMASS * SAPM, XSAMIN);
  s.LFMASS = LFMASS; s.RTMASS = RTMASS; s.STMASS = STMASS; s.WOOD = WOOD; s.FASTCP = FASTCP; s.STBLCP = STBLCP;
  (void)GRSTEM; (void)NPPS;
}

// noahmplsm.F90:8723-8835
NMP_DEV void CARBON(const Ctx& c, int VEGTYP, bool URBAN, float IGS, float DT, float STC1, float PSN, float TV,
                    float FOLN, float BTRAN, const S4& SMC, const L7& DZSNSO, const S4& ZSOIL, CarbonState& s) {
  const noahmp_tables& T = *c.T;
  const Prm& P = c.P;
  if (VEGTYP == T.iswater || VEGTYP == T.isbarren || VEGTYP == T.issnow || URBAN) {
    s.LAI = 0.f; s.SAI = 0.f; s.GPP = 0.f; s.NPP = 0.f; s.NEE = 0.f;
    s.LFMASS = 0.f; s.RTMASS = 0.f; s.STMASS = 0.f; s.WOOD = 0.f; s.STBLCP = 0.f; s.FASTCP = 0.f;
    return;
  }
  float LAPM = tv1(T.sla, VEGTYP) / 1000.f;
  float WSTRES = 1.f - BTRAN;
  float WROOT = 0.f;
  const float ZR = -at4(ZSOIL, P.NROOT);
#pragma unroll
  for (int J = 1; J <= NSOIL; ++J)
    if (J <= P.NROOT) WROOT = WROOT + SMC(J) / P.SMCMAX * DZSNSO(J) / ZR;
  CO2FLUX(c, VEGTYP, IGS, DT, STC1, PSN, TV, WROOT, WSTRES, FOLN, LAPM, s);
}

}  // namespace nmp
