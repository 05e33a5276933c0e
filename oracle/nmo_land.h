// nmo_land.h — internal declarations shared by the oracle's land translation units.
// ORACLE = test infrastructure (see nmo.h).
#pragma once
#include "nmo.h"

namespace nmo {

// locals of NOAHMP_SFLX that travel between ENERGY, WATER, CARBON and ERROR (noahmplsm.F90:647-771)
struct SflxLocal {
  IA<-NSNOW + 1, NSOIL> IMELT;
  ASnSo DZSNSO;
  float THAIR, QAIR, EAIR, RHOAIR, QPRECC, QPRECL, SWDOWN;
  ABand SOLAD, SOLAI;
  float IGS, ELAI, ESAI, HTOP, TROOT;
  ASoil BTRANI; float BTRAN;
  ASoil SICE; ASnow SNICEV, SNLIQV, EPORE;
  float T2M, QDEW, QVAP, QMELT, BEG_WB, TS, TAUX, TAUY, FSRV, FSRG, Q1, Q2E, CMC, QIN, QDIS;
  float LATHEAV, LATHEAG; bool FROZEN_GROUND, FROZEN_CANOPY;
  float AUTORS, HETERS, TOTSC, TOTLB;
};

// land1
void ENERGY(Ctx& c, SflxIO& s, SflxLocal& L);
// land2
void ESAT(float T, float& ESW, float& ESI, float& DESW, float& DESI);
void VEGE_FLUX(Ctx& c, SflxIO& s, SflxLocal& L, int ISNOW, int VEGTYP, float DT, float SAV, float SAG,
               float LWDN, float UR, float UU, float VV, float SFCTMP, float THAIR, float QAIR,
               float EAIR, float RHOAIR, float SNOWH, float VAI, float GAMMAV, float GAMMAG, float FWET,
               float LAISUN, float LAISHA, float CWP, const ASnSo& DZSNSO, float HTOP, float ZLVL,
               float ZPD, float Z0M, float FVEG, float Z0MG, float EMV, float EMG, float CANLIQ,
               float CANICE, const ASnSo& STC, const ASnSo& DF, float& RSSUN, float& RSSHA, float RSURF,
               float LATHEAV, float LATHEAG, float PARSUN, float PARSHA, float IGS, float FOLN,
               float CO2AIR, float O2AIR, float BTRAN, float SFCPRS, float RHSUR, float Q2, float& EAH,
               float& TAH, float& TV, float& TG, float& CM, float& CH, float DX, float DZ8W,
               float& TAUXV, float& TAUYV, float& IRG, float& IRC, float& SHG, float& SHC, float& EVG,
               float& EVC, float& TR, float& GH, float& T2MV, float& PSNSUN, float& PSNSHA, float& QSFC,
               float PSFC, int ISURBAN, int IZ0TLND, float& Q2V, float& CAH2, float& CHLEAF, float& CHUC);
void BARE_FLUX(Ctx& c, SflxIO& s, int ISNOW, float DT, float SAG, float LWDN, float UR, float UU, float VV,
               float SFCTMP, float THAIR, float QAIR, float EAIR, float RHOAIR, float SNOWH,
               const ASnSo& DZSNSO, float ZLVL, float ZPD, float Z0M, float EMG, const ASnSo& STC,
               const ASnSo& DF, float RSURF, float LATHEA, float GAMMA, float RHSUR, float Q2, float& TGB,
               float& CM, float& CH, float& TAUXB, float& TAUYB, float& IRB, float& SHB, float& EVB,
               float& GHB, float& T2MB, float DX, float DZ8W, int IVGTYP, float& QSFC, float PSFC,
               int ISURBAN, int IZ0TLND, float SFCPRS, float& Q2B, float& EHB2);
void TSNOSOI(Ctx& c, int ICE, int ISNOW, int IST, float TBOT, const ASnSo& ZSNSO, float SSOIL,
             const ASnSo& DF, const ASnSo& HCPCT, float ZBOT, float SAG, float DT, float SNOWH,
             const ASnSo& DZSNSO, float TG, ASnSo& STC);
void PHASECHANGE(Ctx& c, int ISNOW, float DT, const ASnSo& FACT, const ASnSo& DZSNSO,
                 const ASnSo& HCPCT, int IST, ASnSo& STC, ASnow& SNICE, ASnow& SNLIQ, float& SNEQV,
                 float& SNOWH, ASoil& SMC, ASoil& SH2O, float& QMELT, IA<-NSNOW + 1, NSOIL>& IMELT,
                 float& PONDING);
// solves the tridiagonal system (noahmplsm.F90:5979-6036); arrays are (-2:4), NTOP..NSOIL active
void ROSR12(ASnSo& P, const ASnSo& A, const ASnSo& B, ASnSo& C, const ASnSo& D, ASnSo& DELTA, int NTOP,
            int NSOILX, int NSNOWX);
// shared with the glacier path (bodies identical in glacier.F90)
void CSNOW(int ISNOW, const ASnow& SNICE, const ASnow& SNLIQ, const ASnSo& DZSNSO, ASnow& TKSNO,
           ASnow& CVSNO, ASnow& SNICEV, ASnow& SNLIQV, ASnow& EPORE);
void SNOW_AGE(float DT, float TG, float SNEQVO, float SNEQV, float& TAUSS, float& FAGE);
void SNOWALB_BATS(float FSNO, float COSZ, float FAGE, ABand& ALBSND, ABand& ALBSNI);
void SNOWALB_CLASS(float QSNOW, float DT, float& ALB, float ALBOLD, ABand& ALBSND, ABand& ALBSNI);
void SFCDIF1(Ctx& c, int ITER, float SFCTMP, float RHOAIR, float H, float QAIR, float ZLVL, float ZPD,
             float Z0M, float Z0H, float UR, float MPE, float& MOZ, int& MOZSGN, float& FM, float& FH,
             float& FM2, float& FH2, float& CM, float& CH, float& FV, float& CH2);
void COMBO(float& DZ, float& WLIQ, float& WICE, float& T, float DZ2, float WLIQ2, float WICE2, float T2);
void COMPACT(float DT, const ASnSo& STC, const ASnow& SNICE, const ASnow& SNLIQ,
             const IA<-NSNOW + 1, NSOIL>& IMELT, const ASnow& FICEOLD, int ISNOW, ASnSo& DZSNSO);
// land3
void WATER(Ctx& c, SflxIO& s, SflxLocal& L);
void CARBON(Ctx& c, SflxIO& s, SflxLocal& L);

}  // namespace nmo
