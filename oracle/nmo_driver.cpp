// nmo_driver.cpp — ORACLE (test infrastructure): the `noahmplsm` grid->column dispatcher
// (phys/module_sf_noahmpdrv.F90:11-844) and the oracle's C entry points for ctypes.
#include <algorithm>
#include <thread>
#include <vector>
#include "nmo_land.h"

namespace nmo {

int g_math_mode = 0;

struct Idx {
  int ims, ime, jms, jme, kms, kme, ni, nj, nk;
  Idx(const noahmp_lsm_args& a)
      : ims(a.ims), ime(a.ime), jms(a.jms), jme(a.jme), kms(a.kms), kme(a.kme), ni(a.ime - a.ims + 1),
        nj(a.jme - a.jms + 1), nk(a.kme - a.kms + 1) {}
  size_t d2(int i, int j) const { return (size_t)(i - ims) + (size_t)(j - jms) * ni; }
  // 3-D A(i,k,j) with k in [k0, k0+nkk-1]
  size_t d3(int i, int k, int j, int k0, int nkk) const {
    return (size_t)(i - ims) + (size_t)(k - k0) * ni + (size_t)(j - jms) * ni * nkk;
  }
};

// One column of the ILOOP body (noahmpdrv.F90:426-837). Returns the column's error code.
static int column(const noahmp_lsm_args& a, const noahmp_tables& T, const Idx& X, int I, int J, int YEARLEN,
                  const ASoil& ZSOIL, float* err_value, int* vege_iters) {
  const int NS = NSOIL;
  const float undefined_value = -1.E36f, undefined_value2 = 0.0f;
  size_t p = X.d2(I, J);
  int ICE;
  if (a.xice[p] >= a.xice_thres) ICE = 1;
  else if (a.ivgtyp[p] == a.isice) ICE = -1;
  else ICE = 0;
  if ((a.xland[p] - 1.5f) >= 0.f) return 0;  // open water
  if (ICE == 1) {
    for (int K = 1; K <= NS; ++K) a.sh2o[X.d3(I, K, J, 1, NS)] = 1.0f;
    a.xlaixy[p] = 0.01f;
    return 0;
  }
  Ctx c;
  c.T = &T;
  c.O = Options{a.idveg, a.iopt_crs, a.iopt_btr, a.iopt_run, a.iopt_sfc, a.iopt_frz, a.iopt_inf,
                a.iopt_rad, a.iopt_alb, a.iopt_snf, a.iopt_tbot, a.iopt_stc};
  std::memset(&c.P, 0, sizeof(c.P));

  float COSZ = a.coszin[p];
  float LAT = a.xlatin[p];
  float Z_ML = 0.5f * a.dz8w[X.d3(I, X.kms, J, X.kms, X.nk)];
  int VEGTYP = a.ivgtyp[p];
  int SOILTYP = a.isltyp[p];
  float FVEG = a.vegfra[p] / 100.f;
  float FVGMAX = a.vegmax[p] / 100.f;
  float TBOT = a.tmn[p];
  float T_ML = a.t3d[X.d3(I, 1, J, X.kms, X.nk)];
  float qv = a.qv3d[X.d3(I, 1, J, X.kms, X.nk)];
  float Q_ML = qv / (1.0f + qv);
  float U_ML = a.u_phy[X.d3(I, 1, J, X.kms, X.nk)];
  float V_ML = a.v_phy[X.d3(I, 1, J, X.kms, X.nk)];
  float SWDN = a.swdown[p];
  float LWDN = a.glw[p];
  float P_ML = (a.p8w3d[X.d3(I, a.kts + 1, J, X.kms, X.nk)] + a.p8w3d[X.d3(I, a.kts, J, X.kms, X.nk)]) * 0.5f;
  float PSFC = a.p8w3d[X.d3(I, 1, J, X.kms, X.nk)];
  float PRCP = a.rainbl[p] / a.dt;

  int ISNOW = a.isnowxy[p];
  ASoil SMC, SMH2O, SMCEQ;
  ASnSo STC, ZSNSO;
  ASnow SNICE, SNLIQ, FICEOLD;
  for (int K = 1; K <= NS; ++K) {
    SMC(K) = a.smois[X.d3(I, K, J, 1, NS)];
    SMH2O(K) = a.sh2o[X.d3(I, K, J, 1, NS)];
    STC(K) = a.tslb[X.d3(I, K, J, 1, NS)];
    SMCEQ(K) = a.smoiseq[X.d3(I, K, J, 1, NS)];
  }
  for (int K = -NSNOW + 1; K <= 0; ++K) {
    STC(K) = a.tsnoxy[X.d3(I, K, J, -NSNOW + 1, NSNOW)];
    SNICE(K) = a.snicexy[X.d3(I, K, J, -NSNOW + 1, NSNOW)];
    SNLIQ(K) = a.snliqxy[X.d3(I, K, J, -NSNOW + 1, NSNOW)];
  }
  for (int K = -NSNOW + 1; K <= NS; ++K) ZSNSO(K) = a.zsnsoxy[X.d3(I, K, J, -NSNOW + 1, NSNOW + NS)];
  float SWE = a.snow[p], SNDPTH = a.snowh[p], QSFC1D = a.qsfc[p];
  float TV = a.tvxy[p], TG = a.tgxy[p], CANLIQ = a.canliqxy[p], CANICE = a.canicexy[p], EAH = a.eahxy[p],
        TAH = a.tahxy[p], CM = a.cmxy[p], CH = a.chxy[p], FWET = a.fwetxy[p], SNEQVO = a.sneqvoxy[p],
        ALBOLD = a.alboldxy[p], QSNOW = a.qsnowxy[p], WSLAKE = a.wslakexy[p], ZWT = a.zwtxy[p],
        WA = a.waxy[p], WT = a.wtxy[p], LFMASS = a.lfmassxy[p], RTMASS = a.rtmassxy[p],
        STMASS = a.stmassxy[p], WOOD = a.woodxy[p], STBLCP = a.stblcpxy[p], FASTCP = a.fastcpxy[p],
        PLAI = a.xlaixy[p], PSAI = a.xsaixy[p], TAUSS = a.taussxy[p], SMCWTD = a.smcwtdxy[p];
  float RECH = 0.f, DEEPRECH = 0.f;

  FICEOLD.fill(0.0f);
  for (int K = ISNOW + 1; K <= 0; ++K) FICEOLD(K) = SNICE(K) / (SNICE(K) + SNLIQ(K));
  const float CO2 = 395.e-06f, O2 = 0.209f;
  float CO2PP = CO2 * P_ML;
  float O2PP = O2 * P_ML;
  float FOLN = 1.0f;
  float QC = undefined_value, PBLH = undefined_value;
  float DZ8W1D = a.dz8w[X.d3(I, 1, J, X.kms, X.nk)];
  int SLOPETYP = 1, IST = 1, ISC = 4;

  if (SOILTYP == 14 && a.xice[p] == 0.f) SOILTYP = 7;
  if (a.ivgtyp[p] == a.isurban || a.ivgtyp[p] == 31 || a.ivgtyp[p] == 32 || a.ivgtyp[p] == 33)
    VEGTYP = a.isurban;
  if (VEGTYP == 25) FVEG = 0.0f;
  if (VEGTYP == 25) PLAI = 0.0f;
  if (VEGTYP == 26) FVEG = 0.0f;
  if (VEGTYP == 26) PLAI = 0.0f;
  if (VEGTYP == 27) FVEG = 0.0f;
  if (VEGTYP == 27) PLAI = 0.0f;

  if (REDPRM(c, VEGTYP, SOILTYP, SLOPETYP, ZSOIL, a.isurban)) {
    *err_value = c.err_value;
    return c.err_code;
  }

  // outputs
  float FSA, FSR, FIRA, FSH, SSOIL, FCEV, FGEV, FCTR, ECAN, ETRAN, ESOIL, TRAD, TGB, TGV, T2MV, T2MB, Q2MV,
      Q2MB, RUNSF, RUNSB, APAR, PSN, SAV, SAG, FSNO, NEE, GPP, NPP, FVEGMP, SALB, QSNBOT, PONDING, PONDING1,
      PONDING2, RSSUN, RSSHA, BGAP, WGAP, CHV, CHB, EMISSI, SHG, SHC, SHB, EVG, EVB, GHV, GHB, IRG, IRC,
      IRB, TR, EVC, CHLEAF, CHUC, CHV2, CHB2, FPICE;

  if (ICE == -1) {
    GlacIO g;
    std::memset(&g, 0, sizeof(g));
    TBOT = MIN(TBOT, 263.15f);
    g.ILOC = I; g.JLOC = J; g.COSZ = COSZ; g.DT = a.dt; g.SFCTMP = T_ML; g.SFCPRS = P_ML; g.UU = U_ML;
    g.VV = V_ML; g.Q2 = Q_ML; g.SOLDN = SWDN; g.PRCP = PRCP; g.LWDN = LWDN; g.TBOT = TBOT; g.ZLVL = Z_ML;
    g.FICEOLD = FICEOLD; g.ZSOIL = ZSOIL;
    g.QSNOW = QSNOW; g.SNEQVO = SNEQVO; g.ALBOLD = ALBOLD; g.CM = CM; g.CH = CH; g.ISNOW = ISNOW;
    g.SNEQV = SWE; g.SMC = SMC; g.ZSNSO = ZSNSO; g.SNOWH = SNDPTH; g.SNICE = SNICE; g.SNLIQ = SNLIQ;
    g.TG = TG; g.STC = STC; g.SH2O = SMH2O; g.TAUSS = TAUSS; g.QSFC = QSFC1D;
    NOAHMP_GLACIER(c, g);
    QSNOW = g.QSNOW; SNEQVO = g.SNEQVO; ALBOLD = g.ALBOLD; CM = g.CM; CH = g.CH; ISNOW = g.ISNOW;
    SWE = g.SNEQV; SMC = g.SMC; ZSNSO = g.ZSNSO; SNDPTH = g.SNOWH; SNICE = g.SNICE; SNLIQ = g.SNLIQ;
    TG = g.TG; STC = g.STC; SMH2O = g.SH2O; TAUSS = g.TAUSS; QSFC1D = g.QSFC;
    FSA = g.FSA; FSR = g.FSR; FIRA = g.FIRA; FSH = g.FSH; FGEV = g.FGEV; SSOIL = g.SSOIL; TRAD = g.TRAD;
    ESOIL = g.EDIR; RUNSF = g.RUNSRF; RUNSB = g.RUNSUB; SAG = g.SAG; SALB = g.ALBEDO; QSNBOT = g.QSNBOT;
    PONDING = g.PONDING; PONDING1 = g.PONDING1; PONDING2 = g.PONDING2; T2MB = g.T2M; Q2MB = g.Q2E;
    EMISSI = g.EMISSI; FPICE = g.FPICE; CHB2 = g.CH2B;
    (void)FSR;

    FSNO = 1.0f;
    TV = undefined_value; TGB = TG; CANICE = undefined_value2; CANLIQ = undefined_value2;
    EAH = undefined_value; TAH = undefined_value; FWET = undefined_value2; WSLAKE = undefined_value2;
    ZWT = undefined_value; WA = undefined_value; WT = undefined_value; LFMASS = undefined_value2;
    RTMASS = undefined_value2; STMASS = undefined_value2; WOOD = undefined_value2;
    STBLCP = undefined_value; FASTCP = undefined_value; PLAI = undefined_value2; PSAI = undefined_value2;
    T2MV = undefined_value; Q2MV = undefined_value; NEE = undefined_value2; GPP = undefined_value2;
    NPP = undefined_value2; FVEGMP = 0.0f; ECAN = undefined_value2; ETRAN = undefined_value2;
    APAR = undefined_value2; PSN = undefined_value2; SAV = undefined_value2; RSSUN = undefined_value;
    RSSHA = undefined_value; BGAP = undefined_value; WGAP = undefined_value; TGV = undefined_value;
    CHV = undefined_value; CHB = CH; IRC = undefined_value; IRG = undefined_value; SHC = undefined_value;
    SHG = undefined_value; EVG = undefined_value; GHV = undefined_value; IRB = FIRA; SHB = FSH; EVB = FGEV;
    GHB = SSOIL; TR = undefined_value2; EVC = undefined_value2; CHLEAF = undefined_value;
    CHUC = undefined_value; CHV2 = undefined_value; FCEV = undefined_value2; FCTR = undefined_value2;
    a.qfx[p] = ESOIL;
    a.lh[p] = FGEV;
    *vege_iters = 0;
  } else {
    SflxIO s;
    std::memset(&s, 0, sizeof(s));
    s.ILOC = I; s.JLOC = J; s.LAT = LAT; s.YEARLEN = YEARLEN; s.JULIAN = a.julian; s.COSZ = COSZ;
    s.DT = a.dt; s.DX = a.dx; s.DZ8W = DZ8W1D; s.ZSOIL = ZSOIL; s.SHDFAC = FVEG; s.SHDMAX = FVGMAX;
    s.VEGTYP = VEGTYP; s.ISURBAN = a.isurban; s.ICE = ICE; s.IST = IST; s.ISC = ISC; s.SMCEQ = SMCEQ;
    s.IZ0TLND = a.iz0tlnd; s.SFCTMP = T_ML; s.SFCPRS = P_ML; s.PSFC = PSFC; s.UU = U_ML; s.VV = V_ML;
    s.Q2 = Q_ML; s.QC = QC; s.SOLDN = SWDN; s.LWDN = LWDN; s.PRCP = PRCP; s.TBOT = TBOT; s.CO2AIR = CO2PP;
    s.O2AIR = O2PP; s.FOLN = FOLN; s.FICEOLD = FICEOLD; s.PBLH = PBLH; s.ZLVL = Z_ML;
    s.ALBOLD = ALBOLD; s.SNEQVO = SNEQVO; s.STC = STC; s.SH2O = SMH2O; s.SMC = SMC; s.TAH = TAH; s.EAH = EAH;
    s.FWET = FWET; s.CANLIQ = CANLIQ; s.CANICE = CANICE; s.TV = TV; s.TG = TG; s.QSFC = QSFC1D;
    s.QSNOW = QSNOW; s.ISNOW = ISNOW; s.ZSNSO = ZSNSO; s.SNOWH = SNDPTH; s.SNEQV = SWE; s.SNICE = SNICE;
    s.SNLIQ = SNLIQ; s.ZWT = ZWT; s.WA = WA; s.WT = WT; s.WSLAKE = WSLAKE; s.LFMASS = LFMASS;
    s.RTMASS = RTMASS; s.STMASS = STMASS; s.WOOD = WOOD; s.STBLCP = STBLCP; s.FASTCP = FASTCP; s.LAI = PLAI;
    s.SAI = PSAI; s.CM = CM; s.CH = CH; s.TAUSS = TAUSS; s.SMCWTD = SMCWTD; s.DEEPRECH = DEEPRECH;
    s.RECH = RECH;
    NOAHMP_SFLX(c, s);
    ALBOLD = s.ALBOLD; SNEQVO = s.SNEQVO; STC = s.STC; SMH2O = s.SH2O; SMC = s.SMC; TAH = s.TAH; EAH = s.EAH;
    FWET = s.FWET; CANLIQ = s.CANLIQ; CANICE = s.CANICE; TV = s.TV; TG = s.TG; QSFC1D = s.QSFC;
    QSNOW = s.QSNOW; ISNOW = s.ISNOW; ZSNSO = s.ZSNSO; SNDPTH = s.SNOWH; SWE = s.SNEQV; SNICE = s.SNICE;
    SNLIQ = s.SNLIQ; ZWT = s.ZWT; WA = s.WA; WT = s.WT; WSLAKE = s.WSLAKE; LFMASS = s.LFMASS;
    RTMASS = s.RTMASS; STMASS = s.STMASS; WOOD = s.WOOD; STBLCP = s.STBLCP; FASTCP = s.FASTCP; PLAI = s.LAI;
    PSAI = s.SAI; CM = s.CM; CH = s.CH; TAUSS = s.TAUSS; SMCWTD = s.SMCWTD; DEEPRECH = s.DEEPRECH;
    RECH = s.RECH;
    FSA = s.FSA; FSR = s.FSR; FIRA = s.FIRA; FSH = s.FSH; SSOIL = s.SSOIL; FCEV = s.FCEV; FGEV = s.FGEV;
    FCTR = s.FCTR; ECAN = s.ECAN; ETRAN = s.ETRAN; ESOIL = s.EDIR; TRAD = s.TRAD; TGB = s.TGB; TGV = s.TGV;
    T2MV = s.T2MV; T2MB = s.T2MB; Q2MV = s.Q2V; Q2MB = s.Q2B; RUNSF = s.RUNSRF; RUNSB = s.RUNSUB;
    APAR = s.APAR; PSN = s.PSN; SAV = s.SAV; SAG = s.SAG; FSNO = s.FSNO; NEE = s.NEE; GPP = s.GPP;
    NPP = s.NPP; FVEGMP = s.FVEG; SALB = s.ALBEDO; QSNBOT = s.QSNBOT; PONDING = s.PONDING;
    PONDING1 = s.PONDING1; PONDING2 = s.PONDING2; RSSUN = s.RSSUN; RSSHA = s.RSSHA; BGAP = s.BGAP;
    WGAP = s.WGAP; CHV = s.CHV; CHB = s.CHB; EMISSI = s.EMISSI; SHG = s.SHG; SHC = s.SHC; SHB = s.SHB;
    EVG = s.EVG; EVB = s.EVB; GHV = s.GHV; GHB = s.GHB; IRG = s.IRG; IRC = s.IRC; IRB = s.IRB; TR = s.TR;
    EVC = s.EVC; CHLEAF = s.CHLEAF; CHUC = s.CHUC; CHV2 = s.CHV2; CHB2 = s.CHB2; FPICE = s.FPICE;
    (void)FSR;
    a.qfx[p] = ECAN + ESOIL + ETRAN;
    a.lh[p] = FCEV + FGEV + FCTR;
    *vege_iters = s.VEGE_ITERS;
  }

  a.tsk[p] = TRAD;
  a.hfx[p] = FSH;
  a.grdflx[p] = SSOIL;
  a.smstav[p] = 0.0f;
  a.smstot[p] = 0.0f;
  a.sfcrunoff[p] = a.sfcrunoff[p] + RUNSF * a.dt;
  a.udrunoff[p] = a.udrunoff[p] + RUNSB * a.dt;
  if (SALB > -999.f) a.albedo[p] = SALB;
  a.snowc[p] = FSNO;
  for (int K = 1; K <= NS; ++K) {
    a.smois[X.d3(I, K, J, 1, NS)] = SMC(K);
    a.sh2o[X.d3(I, K, J, 1, NS)] = SMH2O(K);
    a.tslb[X.d3(I, K, J, 1, NS)] = STC(K);
  }
  a.snow[p] = SWE;
  a.snowh[p] = SNDPTH;
  a.canwat[p] = CANLIQ + CANICE;
  a.acsnow[p] = a.acsnow[p] + PRCP * FPICE;
  a.acsnom[p] = a.acsnom[p] + QSNBOT * a.dt + PONDING + PONDING1 + PONDING2;
  a.emiss[p] = EMISSI;
  a.qsfc[p] = QSFC1D;
  a.isnowxy[p] = ISNOW;
  a.tvxy[p] = TV; a.tgxy[p] = TG; a.canliqxy[p] = CANLIQ; a.canicexy[p] = CANICE; a.eahxy[p] = EAH;
  a.tahxy[p] = TAH; a.cmxy[p] = CM; a.chxy[p] = CH; a.fwetxy[p] = FWET; a.sneqvoxy[p] = SNEQVO;
  a.alboldxy[p] = ALBOLD; a.qsnowxy[p] = QSNOW; a.wslakexy[p] = WSLAKE; a.zwtxy[p] = ZWT; a.waxy[p] = WA;
  a.wtxy[p] = WT;
  for (int K = -NSNOW + 1; K <= 0; ++K) {
    a.tsnoxy[X.d3(I, K, J, -NSNOW + 1, NSNOW)] = STC(K);
    a.snicexy[X.d3(I, K, J, -NSNOW + 1, NSNOW)] = SNICE(K);
    a.snliqxy[X.d3(I, K, J, -NSNOW + 1, NSNOW)] = SNLIQ(K);
  }
  for (int K = -NSNOW + 1; K <= NS; ++K) a.zsnsoxy[X.d3(I, K, J, -NSNOW + 1, NSNOW + NS)] = ZSNSO(K);
  a.lfmassxy[p] = LFMASS; a.rtmassxy[p] = RTMASS; a.stmassxy[p] = STMASS; a.woodxy[p] = WOOD;
  a.stblcpxy[p] = STBLCP; a.fastcpxy[p] = FASTCP; a.xlaixy[p] = PLAI; a.xsaixy[p] = PSAI;
  a.taussxy[p] = TAUSS;
  a.t2mvxy[p] = T2MV; a.t2mbxy[p] = T2MB;
  a.q2mvxy[p] = Q2MV / (1.0f - Q2MV);
  a.q2mbxy[p] = Q2MB / (1.0f - Q2MB);
  a.tradxy[p] = TRAD; a.neexy[p] = NEE; a.gppxy[p] = GPP; a.nppxy[p] = NPP; a.fvegxy[p] = FVEGMP;
  a.runsfxy[p] = RUNSF; a.runsbxy[p] = RUNSB; a.ecanxy[p] = ECAN; a.edirxy[p] = ESOIL; a.etranxy[p] = ETRAN;
  a.fsaxy[p] = FSA; a.firaxy[p] = FIRA; a.aparxy[p] = APAR; a.psnxy[p] = PSN; a.savxy[p] = SAV;
  a.sagxy[p] = SAG; a.rssunxy[p] = RSSUN; a.rsshaxy[p] = RSSHA; a.bgapxy[p] = BGAP; a.wgapxy[p] = WGAP;
  a.tgvxy[p] = TGV; a.tgbxy[p] = TGB; a.chvxy[p] = CHV; a.chbxy[p] = CHB; a.ircxy[p] = IRC; a.irgxy[p] = IRG;
  a.shcxy[p] = SHC; a.shgxy[p] = SHG; a.evgxy[p] = EVG; a.ghvxy[p] = GHV; a.irbxy[p] = IRB; a.shbxy[p] = SHB;
  a.evbxy[p] = EVB; a.ghbxy[p] = GHB; a.trxy[p] = TR; a.evcxy[p] = EVC; a.chleafxy[p] = CHLEAF;
  a.chucxy[p] = CHUC; a.chv2xy[p] = CHV2; a.chb2xy[p] = CHB2;
  a.rechxy[p] = a.rechxy[p] + RECH * 1.E3f;
  a.deeprechxy[p] = a.deeprechxy[p] + DEEPRECH;
  a.smcwtdxy[p] = SMCWTD;

  *err_value = c.err_value;
  return c.err_code;
}

// noahmpdrv.F90:376-840.  nthreads > 1 splits the J rows over std::threads (columns are independent).
static int noahmplsm(const noahmp_lsm_args& a, const noahmp_tables& T, noahmp_status* st, int nthreads,
                     int32_t* vege_iters_out) {
  Idx X(a);
  if (a.nsoil != NSOIL) { if (st) { st->code = NOAHMP_ERR_ARG; } return NOAHMP_ERR_ARG; }
  int YEARLEN = 365;
  if (a.yr % 4 == 0) {
    YEARLEN = 366;
    if (a.yr % 100 == 0) {
      YEARLEN = 365;
      if (a.yr % 400 == 0) YEARLEN = 366;
    }
  }
  ASoil ZSOIL;
  ZSOIL(1) = -a.dzs[0];
  for (int K = 2; K <= NSOIL; ++K) ZSOIL(K) = -a.dzs[K - 1] + ZSOIL(K - 1);

  struct Res { int code = 0, i = 0, j = 0, count = 0; float value = 0.f; };
  nthreads = std::max(1, nthreads);
  std::vector<Res> res(nthreads);
  auto work = [&](int t) {
    Res& r = res[t];
    for (int J = a.jts + t; J <= a.jte; J += nthreads) {
      if (a.itimestep == 1) {
        for (int I = a.its; I <= a.ite; ++I) {
          size_t p = X.d2(I, J);
          if ((a.xland[p] - 1.5f) >= 0.f) {
            a.smstav[p] = 1.0f;
            a.smstot[p] = 1.0f;
            for (int K = 1; K <= NSOIL; ++K) {
              a.smois[X.d3(I, K, J, 1, NSOIL)] = 1.0f;
              a.tslb[X.d3(I, K, J, 1, NSOIL)] = 273.16f;
            }
          } else if (a.xice[p] == 1.f) {
            a.smstav[p] = 1.0f;
            a.smstot[p] = 1.0f;
            for (int K = 1; K <= NSOIL; ++K) a.smois[X.d3(I, K, J, 1, NSOIL)] = 1.0f;
          }
        }
      }
      for (int I = a.its; I <= a.ite; ++I) {
        float ev = 0.f;
        int vi = 0;
        int code = column(a, T, X, I, J, YEARLEN, ZSOIL, &ev, &vi);
        if (vege_iters_out) vege_iters_out[X.d2(I, J)] = vi;
        if (code) {
          if (!r.count || J < r.j || (J == r.j && I < r.i)) { r.code = code; r.i = I; r.j = J; r.value = ev; }
          r.count++;
        }
      }
    }
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  Res tot;
  for (auto& r : res) {
    if (r.count) {
      if (!tot.count || r.j < tot.j || (r.j == tot.j && r.i < tot.i)) {
        tot.code = r.code; tot.i = r.i; tot.j = r.j; tot.value = r.value;
      }
      tot.count += r.count;
    }
  }
  if (st) { st->code = tot.code; st->i = tot.i; st->j = tot.j; st->count = tot.count; st->value = tot.value; }
  return tot.code;
}

}  // namespace nmo

namespace nmo {
__attribute__((noinline)) float powf_libm(float x, float y) { return std::pow(x, y); }
__attribute__((noinline)) double pow_libm(double x, double y) { return std::pow(x, y); }
}  // namespace nmo

extern "C" {

// math_mode: 0 = host libm (reference-like), 1 = portable nmp_math.h (bit-comparable with the GPU
// parity build).
void nmo_set_math_mode(int mode) { nmo::g_math_mode = mode; }
int nmo_get_math_mode(void) { return nmo::g_math_mode; }

int nmo_noahmplsm(const noahmp_lsm_args* args, const noahmp_tables* tables, noahmp_status* status,
                  int nthreads, int32_t* vege_iters /* may be NULL; ni*nj */) {
  return nmo::noahmplsm(*args, *tables, status, nthreads, vege_iters);
}

// scalar probes for micro known-answer tests
void nmo_esat(float T, float* out4) { nmo::ESAT(T, out4[0], out4[1], out4[2], out4[3]); }
float nmo_math1(int fn, float x, float y) {
  using namespace nmo;
  switch (fn) {
    case 0: return EXP(x);
    case 1: return LOG(x);
    case 2: return LOG10(x);
    case 3: return POW(x, y);
    case 4: return ATAN(x);
    case 5: return TAN(x);
    case 6: return COS(x);
    case 7: return ACOS(x);
    case 8: return TANH(x);
    case 9: return POWI(x, (int)y);
    case 10: return (float)DPOW((double)x, (double)y);
    default: return 0.f;
  }
}
// ROSR12 probe: n unknowns (n <= 7) placed on layers NSOIL-n+1..NSOIL of the (-2:4) work arrays
void nmo_rosr12(int n, const float* a, const float* b, const float* c, const float* d, float* x) {
  using namespace nmo;
  ASnSo P, A, B, C, D, DELTA;
  P.fill(0.f); A.fill(0.f); B.fill(0.f); C.fill(0.f); D.fill(0.f); DELTA.fill(0.f);
  const int top = NSOIL - n + 1;
  for (int k = 0; k < n; ++k) { A(top + k) = a[k]; B(top + k) = b[k]; C(top + k) = c[k]; D(top + k) = d[k]; }
  ROSR12(P, A, B, C, D, DELTA, top, NSOIL, NSNOW);
  for (int k = 0; k < n; ++k) x[k] = P(top + k);
}
// COMBO probe: v = {DZ, WLIQ, WICE, T} of layer 1 (updated), w = the same of layer 2
void nmo_combo(float* v, const float* w) { nmo::COMBO(v[0], v[1], v[2], v[3], w[0], w[1], w[2], w[3]); }

void nmo_math_array(int fn, const float* x, const float* y, float* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = nmo_math1(fn, x[i], y ? y[i] : float(0.f));
}
}
