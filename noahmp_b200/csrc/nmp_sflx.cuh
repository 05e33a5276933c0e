// nmp_sflx.cuh — device code of the per-column drivers: REDPRM, NOAHMP_SFLX (ATM, PHENOLOGY, ENERGY,
// WATER, CARBON, ERROR) — phys/module_sf_noahmplsm.F90:518-1843, :6382-6613, :9202-9349.
#pragma once
#include "nmp_energy.cuh"
#include "nmp_io.cuh"
#include "nmp_water.cuh"

namespace nmp {

// Everything one column exchanges with the dispatcher (the NOAHMP_SFLX / NOAHMP_GLACIER dummy lists,
// noahmplsm.F90:518-548, glacier.F90:150-167).  Lives in registers of the owning thread.
struct Col {
  // IN
  float LAT, JULIAN, COSZ, DT, SHDFAC, SHDMAX, SFCTMP, SFCPRS, PSFC, UU, VV, Q2, SOLDN, LWDN, PRCP, TBOT, CO2AIR,
      O2AIR, FOLN, ZLVL;
  int YEARLEN, VEGTYP, ICE;
  bool URBAN;  // VEGTYP == ISURBAN
  S4 ZSOIL, SMCEQ;
  N3 FICEOLD;
  // INOUT
  float ALBOLD, SNEQVO, TAH, EAH, FWET, CANLIQ, CANICE, TV, TG, QSFC, QSNOW, SNOWH, SNEQV, ZWT, WA, WT, WSLAKE,
      LFMASS, RTMASS, STMASS, WOOD, STBLCP, FASTCP, LAI, SAI, CM, CH, TAUSS, SMCWTD, DEEPRECH, RECH;
  int ISNOW;
  L7 STC, ZSNSO;
  S4 SH2O, SMC;
  N3 SNICE, SNLIQ;
  // OUT
  float FSA, FSR, FIRA, FSH, SSOIL, FCEV, FGEV, FCTR, ECAN, ETRAN, EDIR, TRAD, TGB, TGV, T2MV, T2MB, Q2V, Q2B,
      RUNSRF, RUNSUB, APAR, PSN, SAV, SAG, FSNO, NEE, GPP, NPP, FVEG, ALBEDO, QSNBOT, PONDING, PONDING1, PONDING2,
      RSSUN, RSSHA, BGAP, WGAP, CHV, CHB, EMISSI, SHG, SHC, SHB, EVG, EVB, GHV, GHB, IRG, IRC, IRB, TR, EVC,
      CHLEAF, CHUC, CHV2, CHB2, FPICE;
  // diagnostics (not in the reference's list)
  float ERRWAT, ERRENG, ERRSW;
  int VEGE_ITERS;
};

// noahmplsm.F90:9202-9349.  Returns nonzero on a range error.
NMP_DEV int REDPRM(Ctx& c, int VEGTYP, int SOILTYP, int SLOPETYP, bool URBAN) {
  const noahmp_tables& T = *c.T;
  Prm& P = c.P;
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
  if (URBAN) {
    P.SMCMAX = 0.45f; P.SMCREF = 0.42f; P.SMCWLT = 0.40f; P.SMCDRY = 0.40f; P.CSOIL = 3.E6f;
  }
  P.ZBOT = T.zbot_data;
  P.CZIL = T.czil_data;
  const float FRZK = T.frzk_data, REFDK = T.refdk_data, REFKDT = T.refkdt_data;
  P.KDT = REFKDT * P.DKSAT / REFDK;
  P.SLOPE = T.slope_data[SLOPETYP - 1];
  if (SOILTYP != 14) {
    float FRZFACT = (P.SMCMAX / P.SMCREF) * (0.412f / 0.468f);
    P.FRZX = FRZK * FRZFACT;
  } else {
    // The reference leaves FRZX at the previously processed column's value (module global, :9316-9319);
    // columns are independent threads here, so it is defined as 0 (only INFIL / opt_run=3 reads it).
    P.FRZX = 0.f;
  }
  P.TOPT = T.topt_data;
  P.RGL = T.rgltbl[VEGTYP - 1];
  P.RSMAX = T.rsmax_data;
  P.RSMIN = T.rstbl[VEGTYP - 1];
  P.HS = T.hstbl[VEGTYP - 1];
  P.NROOT = T.nrotbl[VEGTYP - 1];
  if (URBAN) P.RSMIN = 400.0f;
  if (P.NROOT > NSOIL) { c.fatal(NOAHMP_ERR_REDPRM, (float)P.NROOT); return 1; }
  return 0;
}

// locals of NOAHMP_SFLX that travel between ENERGY, WATER, CARBON and ERROR (noahmplsm.F90:647-771)
struct SflxLocal {
  I7 IMELT;
  L7 DZSNSO;
  float THAIR, QAIR, EAIR, RHOAIR, QPRECC, QPRECL, SWDOWN;
  B2 SOLAD, SOLAI;
  float IGS, ELAI, ESAI, HTOP;
  S4 BTRANI;
  float BTRAN;
  S4 SICE;
  float QDEW, QVAP, QMELT, BEG_WB, LATHEAV, LATHEAG;
  bool FROZEN_GROUND, FROZEN_CANOPY;
};

// noahmplsm.F90:949-1007 (same maths in glacier.F90:340-390)
NMP_DEV void ATM(float SFCPRS, float SFCTMP, float Q2, float PRCP, float SOLDN, float COSZ, float& THAIR,
                 float& QAIR, float& EAIR, float& RHOAIR, float& QPRECC, float& QPRECL, B2& SOLAD, B2& SOLAI,
                 float& SWDOWN) {
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
template <class O>
NMP_DEV void PHENOLOGY(const Ctx& c, int VEGTYP, bool URBAN, float SNOWH, float TV, float LAT, int YEARLEN,
                       float JULIAN, float& LAI, float& SAI, float& HTOP, float& ELAI, float& ESAI, float& IGS) {
  const noahmp_tables& T = *c.T;
  const int dveg = NMP_OPT(dveg);
  if (dveg == 1 || dveg == 3 || dveg == 4) {
    float DAY;
    if (LAT >= 0.f) {
      DAY = JULIAN;
    } else {
      DAY = fmodf(JULIAN + (0.5f * (float)YEARLEN), (float)YEARLEN);
    }
    float Tm = 12.f * DAY / (float)YEARLEN;
    int IT1 = (int)(Tm + 0.5f);
    int IT2 = IT1 + 1;
    float WT1 = ((float)IT1 + 0.5f) - Tm;
    float WT2 = 1.f - WT1;
    if (IT1 < 1) IT1 = 12;
    if (IT2 > 12) IT2 = 1;
    LAI = WT1 * T.laim[IT1 - 1][VEGTYP - 1] + WT2 * T.laim[IT2 - 1][VEGTYP - 1];
    SAI = WT1 * T.saim[IT1 - 1][VEGTYP - 1] + WT2 * T.saim[IT2 - 1][VEGTYP - 1];
  }
  if (SAI < 0.01f) SAI = 0.0f;
  if (LAI < 0.05f || SAI == 0.0f) LAI = 0.0f;
  if (VEGTYP == T.iswater || VEGTYP == T.isbarren || VEGTYP == T.issnow || URBAN) {
    LAI = 0.f;
    SAI = 0.f;
  }
  float hvt = tv1(T.hvt, VEGTYP), hvb = tv1(T.hvb, VEGTYP);
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
  if (TV > tv1(T.tmin, VEGTYP)) IGS = 1.f; else IGS = 0.f;
  HTOP = hvt;
}

// noahmplsm.F90:1231-1843 (IST = 1, soil; the lake branches of the reference are unreachable: IST is
// hard-coded to 1 by the dispatcher, noahmpdrv.F90:526)
template <class O>
NMP_DEV void ENERGY(Ctx& c, Col& s, SflxLocal& L) {
  const noahmp_tables& T = *c.T;
  const Prm& P = c.P;
  const float MPE = 1.E-6f, PSIWLT = -150.f, Z0 = 0.01f;
  const int IST = 1, ISC = 4;
  L7 FACT, DF, HCPCT;
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) { FACT(K) = 0.f; DF(K) = 0.f; HCPCT(K) = 0.f; }
  s.IRC = 0.f; s.SHC = 0.f; s.IRG = 0.f; s.SHG = 0.f; s.EVG = 0.f; s.EVC = 0.f; s.TR = 0.f; s.GHV = 0.f;
  float PSNSUN = 0.f, PSNSHA = 0.f;
  s.T2MV = 0.f; s.Q2V = 0.f; s.CHV = 0.f; s.CHLEAF = 0.f; s.CHUC = 0.f; s.CHV2 = 0.f;

  float UR = MAX(SQRT(POWR2(s.UU) + POWR2(s.VV)), 1.f);
  float VAI = L.ELAI + L.ESAI;
  const bool VEG = VAI > 0.f;

  s.FSNO = 0.f;
  if (s.SNOWH > 0.f) {
    float BDSNO = s.SNEQV / s.SNOWH;
    float FMELT = POW(BDSNO / 100.f, M_MELT);
    s.FSNO = TANH(s.SNOWH / (2.5f * Z0 * FMELT));
  }
  float Z0MG = Z0 * (1.0f - s.FSNO) + s.FSNO * Z0SNO;
  float ZPDG = s.SNOWH, Z0M, ZPD;
  if (VEG) {
    Z0M = tv1(T.z0mvt, s.VEGTYP);
    ZPD = 0.65f * L.HTOP;
    if (s.SNOWH > ZPD) ZPD = s.SNOWH;
  } else {
    Z0M = Z0MG;
    ZPD = ZPDG;
  }
  float ZLVL = MAX(ZPD, L.HTOP) + s.ZLVL;
  if (ZPDG >= ZLVL) ZLVL = ZPDG + s.ZLVL;

  float CWP = tv1(T.cwpvt, s.VEGTYP);

  NMP_PHASE();
  N3 SNICEV, SNLIQV, EPORE;
  THERMOPROP(c, s.ISNOW, IST, L.DZSNSO, s.DT, s.SNOWH, s.SNICE, s.SNLIQ, P.CSOIL, s.SMC, s.SH2O, s.STC, s.URBAN, DF,
             HCPCT, SNICEV, SNLIQV, EPORE, FACT);

  NMP_PHASE();
  RadOut r;
  RADIATION<O>(c, s.VEGTYP, IST, ISC, s.SNEQVO, s.SNEQV, s.DT, s.COSZ, s.TG, s.TV, s.FSNO, s.QSNOW, s.FWET, L.ELAI,
               L.ESAI, s.SMC(1), L.SOLAD, L.SOLAI, s.FVEG, s.ALBOLD, s.TAUSS, r);
  s.SAV = r.SAV; s.SAG = r.SAG; s.FSR = r.FSR; s.FSA = r.FSA; s.BGAP = r.BGAP; s.WGAP = r.WGAP;

  NMP_PHASE();
  float EMV = 1.f - EXP(-(L.ELAI + L.ESAI) / 1.0f);
  float EMG;
  if (s.ICE == 1) EMG = 0.98f * (1.f - s.FSNO) + 1.0f * s.FSNO;
  else EMG = 0.97f * (1.f - s.FSNO) + 1.0f * s.FSNO;  // EG(IST=1)

  L.BTRAN = 0.f;
  float PSI, GX = 0.f;
  {
    const int btr = NMP_OPT(btr);
    const float ZR = -at4(s.ZSOIL, P.NROOT);
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) {
      L.BTRANI(IZ) = 0.f;
      if (IZ <= P.NROOT) {
        if (btr == 1) GX = (s.SH2O(IZ) - P.SMCWLT) / (P.SMCREF - P.SMCWLT);
        if (btr == 2) {
          PSI = MAX(PSIWLT, -P.PSISAT * POW(MAX(0.01f, s.SH2O(IZ)) / P.SMCMAX, -P.BEXP));
          GX = (1.f - PSI / PSIWLT) / (1.f + P.PSISAT / PSIWLT);
        }
        if (btr == 3) {
          PSI = MAX(PSIWLT, -P.PSISAT * POW(MAX(0.01f, s.SH2O(IZ)) / P.SMCMAX, -P.BEXP));
          GX = 1.f - EXP(-5.8f * (LOG(PSIWLT / PSI)));
        }
        GX = MIN(1.f, MAX(0.f, GX));
        L.BTRANI(IZ) = MAX(MPE, L.DZSNSO(IZ) / ZR * GX);
        L.BTRAN = L.BTRAN + L.BTRANI(IZ);
      }
    }
    L.BTRAN = MAX(MPE, L.BTRAN);
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ)
      if (IZ <= P.NROOT) L.BTRANI(IZ) = L.BTRANI(IZ) / L.BTRAN;
  }

  float RSURF, RHSUR;
  {
    float L_RSURF = (-s.ZSOIL(1)) * (EXP(POW5(1.0f - MIN(1.0f, s.SH2O(1) / P.SMCMAX))) - 1.0f) / (2.71828f - 1.0f);
    float D_RSURF = 2.2E-5f * P.SMCMAX * P.SMCMAX * POW(1.0f - P.SMCWLT / P.SMCMAX, 2.0f + 3.0f / P.BEXP);
    RSURF = L_RSURF / D_RSURF;
    if (s.SH2O(1) < 0.01f && s.SNOWH == 0.f) RSURF = 1.E6f;
    PSI = -P.PSISAT * POW(MAX(0.01f, s.SH2O(1)) / P.SMCMAX, -P.BEXP);
    RHSUR = s.FSNO + (1.f - s.FSNO) * EXP(PSI * GRAV / (RW * s.TG));
  }
  if (s.URBAN && s.SNOWH == 0.f) RSURF = 1.E6f;

  if (s.TV > TFRZ) { L.LATHEAV = HVAP; L.FROZEN_CANOPY = false; }
  else { L.LATHEAV = HSUB; L.FROZEN_CANOPY = true; }
  float GAMMAV = CPAIR * s.SFCPRS / (0.622f * L.LATHEAV);
  if (s.TG > TFRZ) { L.LATHEAG = HVAP; L.FROZEN_GROUND = false; }
  else { L.LATHEAG = HSUB; L.FROZEN_GROUND = true; }
  float GAMMAG = CPAIR * s.SFCPRS / (0.622f * L.LATHEAG);

  NMP_PHASE();
  FluxIn in;
  in.ISNOW = s.ISNOW; in.VEGTYP = s.VEGTYP; in.DT = s.DT; in.SAV = s.SAV; in.SAG = s.SAG; in.LWDN = s.LWDN;
  in.UR = UR; in.UU = s.UU; in.VV = s.VV; in.SFCTMP = s.SFCTMP; in.THAIR = L.THAIR; in.QAIR = L.QAIR;
  in.EAIR = L.EAIR; in.RHOAIR = L.RHOAIR; in.SNOWH = s.SNOWH; in.SFCPRS = s.SFCPRS; in.PSFC = s.PSFC;
  in.RSURF = RSURF; in.RHSUR = RHSUR; in.EMG = EMG; in.ZLVL = ZLVL;
  in.DF_TOP = top7(DF, s.ISNOW); in.DZ_TOP = top7(L.DZSNSO, s.ISNOW); in.STC_TOP = top7(s.STC, s.ISNOW);

  float CMV = 0.f, CMB = 0.f, TAUXV = 0.f, TAUYV = 0.f;
  s.VEGE_ITERS = 0;
  const bool VEGTILE = VEG && s.FVEG > 0.f;
  if (VEGTILE) {
    s.TGV = s.TG;
    CMV = s.CM;
    s.CHV = s.CH;
    VegOut vo;
    VEGE_FLUX<O>(c, in, VAI, GAMMAV, GAMMAG, s.FWET, r.LAISUN, r.LAISHA, CWP, L.HTOP, ZPD, Z0M, s.FVEG, Z0MG, EMV,
                 s.CANLIQ, s.CANICE, s.RSSUN, s.RSSHA, L.LATHEAV, r.PARSUN, r.PARSHA, L.IGS, s.FOLN, s.CO2AIR,
                 s.O2AIR, L.BTRAN, s.EAH, s.TAH, s.TV, s.TGV, CMV, s.CHV, s.QSFC, vo, s.VEGE_ITERS);
    TAUXV = vo.TAUXV; TAUYV = vo.TAUYV; s.IRG = vo.IRG; s.IRC = vo.IRC; s.SHG = vo.SHG; s.SHC = vo.SHC;
    s.EVG = vo.EVG; s.EVC = vo.EVC; s.TR = vo.TR; s.GHV = vo.GH; s.T2MV = vo.T2MV; PSNSUN = vo.PSNSUN;
    PSNSHA = vo.PSNSHA; s.Q2V = vo.Q2V; s.CHV2 = vo.CAH2; s.CHLEAF = vo.CHLEAF; s.CHUC = vo.CHUC;
  }

  NMP_PHASE_MAJOR();
  s.TGB = s.TG;
  CMB = s.CM;
  s.CHB = s.CH;
  BareOut bo;
  BARE_FLUX<O>(c, in, ZPDG, Z0MG, L.LATHEAG, GAMMAG, s.TGB, CMB, s.CHB, s.QSFC, s.URBAN, bo);
  s.IRB = bo.IRB; s.SHB = bo.SHB; s.EVB = bo.EVB; s.GHB = bo.GHB; s.T2MB = bo.T2MB; s.Q2B = bo.Q2B;
  s.CHB2 = bo.EHB2;
  (void)TAUXV; (void)TAUYV;

  NMP_PHASE_MAJOR();
  if (VEGTILE) {
    s.FIRA = s.FVEG * s.IRG + (1.0f - s.FVEG) * s.IRB + s.IRC;
    s.FSH = s.FVEG * s.SHG + (1.0f - s.FVEG) * s.SHB + s.SHC;
    s.FGEV = s.FVEG * s.EVG + (1.0f - s.FVEG) * s.EVB;
    s.SSOIL = s.FVEG * s.GHV + (1.0f - s.FVEG) * s.GHB;
    s.FCEV = s.EVC;
    s.FCTR = s.TR;
    s.TG = s.FVEG * s.TGV + (1.0f - s.FVEG) * s.TGB;
    s.CM = s.FVEG * CMV + (1.0f - s.FVEG) * CMB;
    s.CH = s.FVEG * s.CHV + (1.0f - s.FVEG) * s.CHB;
  } else {
    s.FIRA = s.IRB;
    s.FSH = s.SHB;
    s.FGEV = s.EVB;
    s.SSOIL = s.GHB;
    s.TG = s.TGB;
    s.FCEV = 0.f;
    s.FCTR = 0.f;
    s.CM = CMB;
    s.CH = s.CHB;
    s.RSSUN = 0.0f;
    s.RSSHA = 0.0f;
    s.TGV = s.TGB;
    s.CHV = s.CHB;
  }

  float FIRE = s.LWDN + s.FIRA;
  if (FIRE <= 0.f) c.fatal(NOAHMP_ERR_FIRE, FIRE);

  s.EMISSI = s.FVEG * (EMG * (1.f - EMV) + EMV + EMV * (1.f - EMV) * (1.f - EMG)) + (1.f - s.FVEG) * EMG;
  s.TRAD = POW((FIRE - (1.f - s.EMISSI) * s.LWDN) / (s.EMISSI * SB), 0.25f);
  s.APAR = r.PARSUN * r.LAISUN + r.PARSHA * r.LAISHA;
  s.PSN = PSNSUN * r.LAISUN + PSNSHA * r.LAISHA;

  NMP_PHASE();
  TSNOSOI<O>(c, s.ISNOW, s.TBOT, s.ZSNSO, s.SSOIL, DF, HCPCT, P.ZBOT, s.DT, s.SNOWH, s.STC);

  if (NMP_OPT(stc) == 2) {
    if (s.SNOWH > 0.05f && s.TG > TFRZ) {
      s.TGV = TFRZ;
      s.TGB = TFRZ;
      if (VEGTILE) s.TG = s.FVEG * s.TGV + (1.0f - s.FVEG) * s.TGB;
      else s.TG = s.TGB;
    }
  }

  NMP_PHASE();
  PHASECHANGE<O>(c, s.ISNOW, s.DT, FACT, L.DZSNSO, IST, s.STC, s.SNICE, s.SNLIQ, s.SNEQV, s.SNOWH, s.SMC, s.SH2O,
                 L.QMELT, L.IMELT, s.PONDING);
}

// noahmplsm.F90:6382-6613 (IST = 1)
template <class O>
NMP_DEV void WATER(Ctx& c, Col& s, SflxLocal& L) {
  const Prm& P = c.P;
  const int run = NMP_OPT(run);
  S4 ETRANI, WCND;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) { ETRANI(K) = 0.f; WCND(K) = 0.f; }
  float SNOFLOW = 0.f, QINSUR = 0.f, QSNSUB, QSEVA, QSNFRO, QSDEW, QDRAIN = 0.f, FCRMAX = 0.f;
  s.RUNSUB = 0.f;

  CanOut co;
  CANWATER<O>(c, s.VEGTYP, s.DT, s.SFCTMP, s.UU, s.VV, s.FCEV, s.FCTR, L.QPRECC, L.QPRECL, L.ELAI, L.ESAI, 1, s.TG,
              s.FVEG, L.FROZEN_CANOPY, s.CANLIQ, s.CANICE, s.TV, s.FWET, co);
  s.ECAN = co.ECAN; s.ETRAN = co.ETRAN; s.QSNOW = co.QSNOW; s.FPICE = co.FPICE;
  const float QRAIN = co.QRAIN, SNOWHIN = co.SNOWHIN;

  NMP_PHASE();
  QSNSUB = 0.f;
  if (s.SNEQV > 0.f) QSNSUB = MIN(L.QVAP, s.SNEQV / s.DT);
  QSEVA = L.QVAP - QSNSUB;
  QSNFRO = 0.f;
  if (s.SNEQV > 0.f) QSNFRO = L.QDEW;
  QSDEW = L.QDEW - QSNFRO;

  SNOWWATER(L.IMELT, s.DT, s.ZSOIL, s.SFCTMP, SNOWHIN, s.QSNOW, QSNFRO, QSNSUB, QRAIN, s.FICEOLD, s.ISNOW, s.SNOWH,
            s.SNEQV, s.SNICE, s.SNLIQ, s.SH2O(1), L.SICE(1), s.STC, s.ZSNSO, L.DZSNSO, s.QSNBOT, SNOFLOW, s.PONDING1,
            s.PONDING2);

  NMP_PHASE();
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
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ)
    if (IZ <= P.NROOT) ETRANI(IZ) = s.ETRAN * L.BTRANI(IZ) * 0.001f;

  NMP_PHASE_MAJOR();
  SOILWATER<O>(c, s.DT, s.ZSOIL, L.DZSNSO, QINSUR, QSEVA, ETRANI, L.SICE, s.SH2O, s.SMC, s.ZWT, s.URBAN, s.SMCWTD,
               s.DEEPRECH, s.RUNSRF, QDRAIN, s.RUNSUB, WCND, FCRMAX);
  NMP_PHASE();
  if (run == 1) {
    float QIN, QDIS;
    GROUNDWATER(P, s.DT, L.SICE, s.ZSOIL, WCND, FCRMAX, s.SH2O, s.ZWT, s.WA, s.WT, QIN, QDIS);
    s.RUNSUB = QDIS;
  }
  if (run == 3 || run == 4) s.RUNSUB = s.RUNSUB + QDRAIN;
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) s.SMC(IZ) = s.SH2O(IZ) + L.SICE(IZ);
  if (run == 5) {
    SHALLOWWATERTABLE(P, s.ZSOIL, L.DZSNSO, s.SMCEQ, s.SMC, s.ZWT, s.SMCWTD, s.RECH);
    s.SH2O(NSOIL) = s.SMC(NSOIL) - L.SICE(NSOIL);
    s.RUNSUB = s.RUNSUB + QDRAIN;
    s.WA = 0.f;
  }
  s.RUNSUB = s.RUNSUB + SNOFLOW;
}

// Scatter of the quantities that are final once ENERGY has run (noahmpdrv.F90:728-835 for these fields).
// Storing them here, not at the end of the column program, ends their live ranges before WATER / CARBON.
NMP_DEV void store_energy_outputs(const ColumnIO& io, const Col& s) {
  io.st(NMP_SLOT(tsk), s.TRAD);
  io.st(NMP_SLOT(tradxy), s.TRAD);
  io.st(NMP_SLOT(hfx), s.FSH);
  io.st(NMP_SLOT(lh), s.FCEV + s.FGEV + s.FCTR);
  io.st(NMP_SLOT(grdflx), s.SSOIL);
  io.st(NMP_SLOT(snowc), s.FSNO);
  io.st(NMP_SLOT(emiss), s.EMISSI);
  io.st(NMP_SLOT(tgxy), s.TG);
  io.st(NMP_SLOT(eahxy), s.EAH);
  io.st(NMP_SLOT(tahxy), s.TAH);
  io.st(NMP_SLOT(cmxy), s.CM);
  io.st(NMP_SLOT(chxy), s.CH);
  io.st(NMP_SLOT(alboldxy), s.ALBOLD);
  io.st(NMP_SLOT(taussxy), s.TAUSS);
  io.st(NMP_SLOT(t2mvxy), s.T2MV);
  io.st(NMP_SLOT(t2mbxy), s.T2MB);
  io.st(NMP_SLOT(q2mvxy), s.Q2V / (1.0f - s.Q2V));
  io.st(NMP_SLOT(fvegxy), s.FVEG);
  io.st(NMP_SLOT(fsaxy), s.FSA);
  io.st(NMP_SLOT(firaxy), s.FIRA);
  io.st(NMP_SLOT(aparxy), s.APAR);
  io.st(NMP_SLOT(psnxy), s.PSN);
  io.st(NMP_SLOT(savxy), s.SAV);
  io.st(NMP_SLOT(sagxy), s.SAG);
  io.st(NMP_SLOT(rssunxy), s.RSSUN);
  io.st(NMP_SLOT(rsshaxy), s.RSSHA);
  io.st(NMP_SLOT(bgapxy), s.BGAP);
  io.st(NMP_SLOT(wgapxy), s.WGAP);
  io.st(NMP_SLOT(tgvxy), s.TGV);
  io.st(NMP_SLOT(tgbxy), s.TGB);
  io.st(NMP_SLOT(chvxy), s.CHV);
  io.st(NMP_SLOT(chbxy), s.CHB);
  io.st(NMP_SLOT(ircxy), s.IRC);
  io.st(NMP_SLOT(irgxy), s.IRG);
  io.st(NMP_SLOT(shcxy), s.SHC);
  io.st(NMP_SLOT(shgxy), s.SHG);
  io.st(NMP_SLOT(evgxy), s.EVG);
  io.st(NMP_SLOT(ghvxy), s.GHV);
  io.st(NMP_SLOT(irbxy), s.IRB);
  io.st(NMP_SLOT(shbxy), s.SHB);
  io.st(NMP_SLOT(evbxy), s.EVB);
  io.st(NMP_SLOT(ghbxy), s.GHB);
  io.st(NMP_SLOT(trxy), s.TR);
  io.st(NMP_SLOT(evcxy), s.EVC);
  io.st(NMP_SLOT(chleafxy), s.CHLEAF);
  io.st(NMP_SLOT(chucxy), s.CHUC);
  io.st(NMP_SLOT(chv2xy), s.CHV2);
  io.st(NMP_SLOT(chb2xy), s.CHB2);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) io.st(NMP_SLOT(tslb) + K - 1, s.STC(K));  // soil temperatures: final after ENERGY
}

// ---- ENERGY | WATER as two kernels (NMP_SPLIT build): what crosses the cut -----------------------------------------
// PART 0 = the fused column program; PART 1 = ATM ... ENERGY and the scatter of everything ENERGY finalised, then the
// hand-off; PART 2 = hand-off, WATER, CARBON, ERROR's water balance and the rest of the scatter.  The arithmetic of a
// column is the same in both forms (the PARITY build of either is bit-identical to the oracle).
enum { HO_FCEV = 0, HO_FCTR, HO_FGEV, HO_FVEG, HO_ELAI, HO_ESAI, HO_IGS, HO_BTRAN, HO_BTRANI0, HO_FLAGS = HO_BTRANI0 + 4,
       HO_FICEOLD0, HO_PONDING = HO_FICEOLD0 + 3, HO_BEG_WB, HO_PSN, HO_Q2B, HO_COUNT };
static_assert(HO_COUNT <= 20, "hand-off planes");
NMP_DEV void store_handoff(const ColumnIO& io, const Col& s, const SflxLocal& L, int failed) {
  const int H = nmpf::PLANE_HANDOFF0;
  io.st(H + HO_FCEV, s.FCEV); io.st(H + HO_FCTR, s.FCTR); io.st(H + HO_FGEV, s.FGEV); io.st(H + HO_FVEG, s.FVEG);
  io.st(H + HO_ELAI, L.ELAI); io.st(H + HO_ESAI, L.ESAI); io.st(H + HO_IGS, L.IGS); io.st(H + HO_BTRAN, L.BTRAN);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) io.st(H + HO_BTRANI0 + K - 1, L.BTRANI(K));
  int flags = (L.FROZEN_CANOPY ? 1 : 0) | (L.FROZEN_GROUND ? 2 : 0) | (failed ? 4 : 0);
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) flags |= (L.IMELT(K) & 3) << (4 + 2 * (K + 2));
  io.sti(H + HO_FLAGS, flags);
#pragma unroll
  for (int K = -2; K <= 0; ++K) io.st(H + HO_FICEOLD0 + K + 2, s.FICEOLD(K));
  io.st(H + HO_PONDING, s.PONDING); io.st(H + HO_BEG_WB, L.BEG_WB); io.st(H + HO_PSN, s.PSN); io.st(H + HO_Q2B, s.Q2B);
  // state ENERGY changed and WATER reads back through the ordinary planes
#pragma unroll
  for (int K = -2; K <= 0; ++K) {
    io.st(NMP_SLOT(tsnoxy) + K + 2, s.STC(K));
    io.st(NMP_SLOT(snicexy) + K + 2, s.SNICE(K));
    io.st(NMP_SLOT(snliqxy) + K + 2, s.SNLIQ(K));
  }
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    io.st(NMP_SLOT(smois) + K - 1, s.SMC(K));
    io.st(NMP_SLOT(sh2o) + K - 1, s.SH2O(K));
  }
  io.st(NMP_SLOT(snow), s.SNEQV); io.st(NMP_SLOT(snowh), s.SNOWH); io.st(NMP_SLOT(tvxy), s.TV);
  io.st(NMP_SLOT(qsfc), s.QSFC); io.st(NMP_SLOT(xlaixy), s.LAI); io.st(NMP_SLOT(xsaixy), s.SAI);
}
NMP_DEV int load_handoff(const ColumnIO& io, Col& s, SflxLocal& L) {
  const int H = nmpf::PLANE_HANDOFF0;
  s.FCEV = io.ld(H + HO_FCEV); s.FCTR = io.ld(H + HO_FCTR); s.FGEV = io.ld(H + HO_FGEV); s.FVEG = io.ld(H + HO_FVEG);
  L.ELAI = io.ld(H + HO_ELAI); L.ESAI = io.ld(H + HO_ESAI); L.IGS = io.ld(H + HO_IGS); L.BTRAN = io.ld(H + HO_BTRAN);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) L.BTRANI(K) = io.ld(H + HO_BTRANI0 + K - 1);
  const int flags = io.ldi(H + HO_FLAGS);
  L.FROZEN_CANOPY = (flags & 1) != 0;
  L.FROZEN_GROUND = (flags & 2) != 0;
  L.LATHEAG = L.FROZEN_GROUND ? HSUB : HVAP;
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) L.IMELT(K) = (flags >> (4 + 2 * (K + 2))) & 3;
#pragma unroll
  for (int K = -2; K <= 0; ++K) s.FICEOLD(K) = io.ld(H + HO_FICEOLD0 + K + 2);
  s.PONDING = io.ld(H + HO_PONDING); L.BEG_WB = io.ld(H + HO_BEG_WB); s.PSN = io.ld(H + HO_PSN); s.Q2B = io.ld(H + HO_Q2B);
  return (flags & 4) != 0;
}

// noahmplsm.F90:518-947, with the dispatcher's scatter of the column (noahmpdrv.F90:713-835) folded in
template <class O, int PART = 0>
NMP_DEV void NOAHMP_SFLX(Ctx& c, Col& s, const ColumnIO& io) {
  const float DTX = io.p.dt;
  const noahmp_tables& T = *c.T;
  const int dveg = NMP_OPT(dveg);
  SflxLocal L;
  s.NEE = 0.0f; s.NPP = 0.0f; s.GPP = 0.0f;
  L.QMELT = 0.f;

  ATM(s.SFCPRS, s.SFCTMP, s.Q2, s.PRCP, s.SOLDN, s.COSZ, L.THAIR, L.QAIR, L.EAIR, L.RHOAIR, L.QPRECC, L.QPRECL,
      L.SOLAD, L.SOLAI, L.SWDOWN);

  const TopLayer top0(s.ISNOW);
#pragma unroll
  for (int IZ = -2; IZ <= NSOIL; ++IZ) {
    L.DZSNSO(IZ) = 0.f;
    if (top0.is(IZ)) L.DZSNSO(IZ) = -s.ZSNSO(IZ);
    else if (IZ > s.ISNOW + 1) L.DZSNSO(IZ) = s.ZSNSO(IZ - (IZ > -2 ? 1 : 0)) - s.ZSNSO(IZ);
  }

  // TROOT (:798-801) is computed by the reference but never used by PHENOLOGY's body.

  if constexpr (PART != 2) {
  L.BEG_WB = s.CANLIQ + s.CANICE + s.SNEQV + s.WA;
#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) L.BEG_WB = L.BEG_WB + s.SMC(IZ) * L.DZSNSO(IZ) * 1000.f;

  PHENOLOGY<O>(c, s.VEGTYP, s.URBAN, s.SNOWH, s.TV, s.LAT, s.YEARLEN, s.JULIAN, s.LAI, s.SAI, L.HTOP, L.ELAI, L.ESAI,
               L.IGS);

  if (dveg == 1) {
    s.FVEG = s.SHDFAC;
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else if (dveg == 2 || dveg == 3) {
    s.FVEG = 1.f - EXP(-0.52f * (s.LAI + s.SAI));
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else if (dveg == 4 || dveg == 5) {
    s.FVEG = s.SHDMAX;
    if (s.FVEG <= 0.01f) s.FVEG = 0.01f;
  } else {
    c.fatal(NOAHMP_ERR_OPTION, (float)dveg);
    return;
  }
  if (s.URBAN || s.VEGTYP == T.isbarren) s.FVEG = 0.0f;
  if (L.ELAI + L.ESAI == 0.0f) s.FVEG = 0.0f;

  NMP_PHASE_MAJOR();
  ENERGY<O>(c, s, L);

  // the SW and energy-balance checks of ERROR (:1164-1199) depend on ENERGY outputs only; they are evaluated
  // here (same order of the fatal latch: ERRSW, ERRENG, then ERRWAT after WATER) so those fluxes can be retired
  s.ERRSW = L.SWDOWN - (s.FSA + s.FSR);
  if (ABS(s.ERRSW) > 0.01f) c.fatal(NOAHMP_ERR_ERRSW, s.ERRSW);
  s.ERRENG = s.SAV + s.SAG - (s.FIRA + s.FSH + s.FCEV + s.FGEV + s.FCTR + s.SSOIL);
  if (ABS(s.ERRENG) > 0.01f) c.fatal(NOAHMP_ERR_ERRENG, s.ERRENG);
  if (L.SWDOWN != 0.f) s.ALBEDO = s.FSR / L.SWDOWN; else s.ALBEDO = -999.9f;
  if (s.ALBEDO > -999.f) io.st(NMP_SLOT(albedo), s.ALBEDO);
  store_energy_outputs(io, s);
  s.SNEQVO = s.SNEQV;
  io.st(NMP_SLOT(sneqvoxy), s.SNEQVO);
  }  // PART != 2
  if constexpr (PART == 1) {
    store_handoff(io, s, L, c.err != 0);
    return;
  }
  int energy_failed = 0;
  if constexpr (PART == 2) energy_failed = load_handoff(io, s, L);

#pragma unroll
  for (int IZ = 1; IZ <= NSOIL; ++IZ) L.SICE(IZ) = MAX(0.0f, s.SMC(IZ) - s.SH2O(IZ));

  // water-table state is first needed here (late load)
  s.ZWT = io.ld(NMP_SLOT(zwtxy));
  s.WA = io.ld(NMP_SLOT(waxy));
  s.WT = io.ld(NMP_SLOT(wtxy));
  s.SMCWTD = io.ld(NMP_SLOT(smcwtdxy));
  if (NMP_OPT(run) == 5) {
#pragma unroll
    for (int K = 1; K <= NSOIL; ++K) s.SMCEQ(K) = io.ld(NMP_SLOT(smoiseq) + K - 1);
  }

  L.QVAP = MAX(s.FGEV / L.LATHEAG, 0.f);
  L.QDEW = ABS(MIN(s.FGEV / L.LATHEAG, 0.f));
  s.EDIR = L.QVAP - L.QDEW;

  NMP_PHASE_MAJOR();
  WATER<O>(c, s, L);

  NMP_PHASE_MAJOR();
  if (dveg == 2 || dveg == 5) {
    // carbon pools: loaded here, stored right after (with dveg 1/3/4 they pass through untouched in HBM)
    CarbonState cs;
    cs.LFMASS = io.ld(NMP_SLOT(lfmassxy)); cs.RTMASS = io.ld(NMP_SLOT(rtmassxy));
    cs.STMASS = io.ld(NMP_SLOT(stmassxy)); cs.WOOD = io.ld(NMP_SLOT(woodxy));
    cs.STBLCP = io.ld(NMP_SLOT(stblcpxy)); cs.FASTCP = io.ld(NMP_SLOT(fastcpxy));
    cs.LAI = s.LAI; cs.SAI = s.SAI; cs.GPP = s.GPP; cs.NPP = s.NPP; cs.NEE = s.NEE;
    CARBON(c, s.VEGTYP, s.URBAN, L.IGS, s.DT, s.STC(1), s.PSN, s.TV, s.FOLN, L.BTRAN, s.SMC, L.DZSNSO, s.ZSOIL, cs);
    io.st(NMP_SLOT(lfmassxy), cs.LFMASS); io.st(NMP_SLOT(rtmassxy), cs.RTMASS);
    io.st(NMP_SLOT(stmassxy), cs.STMASS); io.st(NMP_SLOT(woodxy), cs.WOOD);
    io.st(NMP_SLOT(stblcpxy), cs.STBLCP); io.st(NMP_SLOT(fastcpxy), cs.FASTCP);
    s.LAI = cs.LAI; s.SAI = cs.SAI; s.GPP = cs.GPP; s.NPP = cs.NPP; s.NEE = cs.NEE;
  }

  NMP_PHASE_MAJOR();
  // ERROR (:1106-1228): water balance (the SW / energy checks ran after ENERGY)
  {
    float END_WB = s.CANLIQ + s.CANICE + s.SNEQV + s.WA;
#pragma unroll
    for (int IZ = 1; IZ <= NSOIL; ++IZ) END_WB = END_WB + s.SMC(IZ) * L.DZSNSO(IZ) * 1000.f;
    s.ERRWAT = END_WB - L.BEG_WB - (s.PRCP - s.ECAN - s.ETRAN - s.EDIR - s.RUNSRF - s.RUNSUB) * s.DT;
    if (ABS(s.ERRWAT) > 0.1f && !energy_failed) c.fatal(NOAHMP_ERR_ERRWAT, s.ERRWAT);
  }

  float QFX = s.ETRAN + s.ECAN + s.EDIR;
  if (s.URBAN) {
    s.QSFC = (QFX / L.RHOAIR * s.CH) + L.QAIR;
    s.Q2B = s.QSFC;
  }
  if (s.SNOWH <= 1.E-6f || s.SNEQV <= 1.E-3f) {
    s.SNOWH = 0.0f;
    s.SNEQV = 0.0f;
  }

  // ---- scatter of everything WATER / CARBON finalised (noahmpdrv.F90:713-835) ----
  io.st(NMP_SLOT(qfx), s.ECAN + s.EDIR + s.ETRAN);
  io.st(NMP_SLOT(smstav), 0.0f);
  io.st(NMP_SLOT(smstot), 0.0f);
  io.acc(NMP_SLOT(sfcrunoff), s.RUNSRF * DTX);
  io.acc(NMP_SLOT(udrunoff), s.RUNSUB * DTX);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    io.st(NMP_SLOT(smois) + K - 1, s.SMC(K));
    io.st(NMP_SLOT(sh2o) + K - 1, s.SH2O(K));
  }
  io.st(NMP_SLOT(snow), s.SNEQV);
  io.st(NMP_SLOT(snowh), s.SNOWH);
  io.st(NMP_SLOT(canwat), s.CANLIQ + s.CANICE);
  io.acc(NMP_SLOT(acsnow), s.PRCP * s.FPICE);
  io.acc(NMP_SLOT(acsnom), s.QSNBOT * DTX + s.PONDING + s.PONDING1 + s.PONDING2);
  io.st(NMP_SLOT(qsfc), s.QSFC);
  io.sti(NMP_SLOT(isnowxy), s.ISNOW);
  io.st(NMP_SLOT(tvxy), s.TV);
  io.st(NMP_SLOT(canliqxy), s.CANLIQ);
  io.st(NMP_SLOT(canicexy), s.CANICE);
  io.st(NMP_SLOT(fwetxy), s.FWET);
  io.st(NMP_SLOT(qsnowxy), s.QSNOW);
  io.st(NMP_SLOT(zwtxy), s.ZWT);
  io.st(NMP_SLOT(waxy), s.WA);
  io.st(NMP_SLOT(wtxy), s.WT);
#pragma unroll
  for (int K = -2; K <= 0; ++K) {
    io.st(NMP_SLOT(tsnoxy) + K + 2, s.STC(K));
    io.st(NMP_SLOT(snicexy) + K + 2, s.SNICE(K));
    io.st(NMP_SLOT(snliqxy) + K + 2, s.SNLIQ(K));
  }
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) io.st(NMP_SLOT(zsnsoxy) + K + 2, s.ZSNSO(K));
  io.st(NMP_SLOT(xlaixy), s.LAI);
  io.st(NMP_SLOT(xsaixy), s.SAI);
  io.st(NMP_SLOT(q2mbxy), s.Q2B / (1.0f - s.Q2B));
  io.st(NMP_SLOT(neexy), s.NEE);
  io.st(NMP_SLOT(gppxy), s.GPP);
  io.st(NMP_SLOT(nppxy), s.NPP);
  io.st(NMP_SLOT(runsfxy), s.RUNSRF);
  io.st(NMP_SLOT(runsbxy), s.RUNSUB);
  io.st(NMP_SLOT(ecanxy), s.ECAN);
  io.st(NMP_SLOT(edirxy), s.EDIR);
  io.st(NMP_SLOT(etranxy), s.ETRAN);
  io.acc(NMP_SLOT(rechxy), s.RECH * 1.E3f);
  io.acc(NMP_SLOT(deeprechxy), s.DEEPRECH);
  io.st(NMP_SLOT(smcwtdxy), s.SMCWTD);
}

}  // namespace nmp
