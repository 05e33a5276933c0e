// nmo_init.cpp — ORACLE (test infrastructure): the cold start NOAHMP_INIT with SNOW_INIT, GROUNDWATER_INIT and
// EQSMOISTURE.  Restates phys/module_sf_noahmpdrv.F90:847-1522 in source order, fp32.
// Pinned bit for bit against the machine-translated reference (see nmo.h; tests/test_reference_pin.py::test_noahmp_init);
// also cross-checked in tests/ against the
// independent numpy restatement of the same routines in noahmp_b200/synthetic.py.
#include <cmath>
#include <vector>
#include "nmo.h"

namespace nmo {

static const float KLATFACTOR_I[19] = {2.f, 3.f, 4.f, 10.f, 10.f, 12.f, 14.f, 20.f, 24.f, 28.f,
                                       40.f, 48.f, 2.f, 0.f, 10.f, 0.f, 20.f, 2.f, 2.f};

// noahmpdrv.F90:1473-1522
static void EQSMOISTURE(const ASoil& ZSOIL, float SMCMAX, float SMCWLT, float DWSAT, float DKSAT, float BEXP, ASoil& SMCEQ) {
  (void)SMCWLT;
  for (int K = 1; K <= NSOIL; ++K) {
    float DDZ;
    if (K == 1) DDZ = -ZSOIL(K + 1) * 0.5f;
    else if (K < NSOIL) DDZ = (ZSOIL(K - 1) - ZSOIL(K + 1)) * 0.5f;
    else DDZ = ZSOIL(K - 1) - ZSOIL(K);
    const float EXPON = BEXP + 1.f;
    const float AA = DWSAT / DDZ;
    const float BB = DKSAT / POW(SMCMAX, EXPON);
    float SMC = 0.5f * SMCMAX;
    for (int ITER = 1; ITER <= 100; ++ITER) {
      const float FUNC = (SMC - SMCMAX) * AA + BB * POW(SMC, EXPON);
      const float DFUNC = AA + BB * EXPON * POW(SMC, BEXP);
      const float DX = FUNC / DFUNC;
      SMC = SMC - DX;
      if (ABS(DX) < 1.E-6f) break;
    }
    SMCEQ(K) = MIN(MAX(SMC, 1.E-4f), SMCMAX * 0.99f);
  }
}

static int INIT(const noahmp_init_args& a, const noahmp_tables& T) {
  if (a.nsoil != NSOIL) return NOAHMP_ERR_ARG;
  if (a.restart) return 0;
  const int ni = a.ime - a.ims + 1;
  auto d2 = [&](int I, int J) { return (size_t)(I - a.ims) + (size_t)(J - a.jms) * ni; };
  auto dL = [&](int I, int K, int K0, int NK, int J) {
    return (size_t)(I - a.ims) + (size_t)(K - K0) * ni + (size_t)(J - a.jms) * ni * NK;
  };
  const int itf = IMIN(a.ite, a.ide - 1), jtf = IMIN(a.jte, a.jde - 1);
  const float BLIM = 5.5f, HLICE = 3.335E5f, GRAV_I = 9.81f, T0 = 273.15f;
  (void)BLIM;
  // :996-1005
  if (!a.fndsnowh)
    for (int J = a.jts; J <= jtf; ++J)
      for (int I = a.its; I <= itf; ++I) a.snowh[d2(I, J)] = a.snow[d2(I, J)] * 0.005f;
  // :1007-1021
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I)
      if (a.isltyp[d2(I, J)] < 1) return NOAHMP_ERR_ISLTYP;
  // :1033-1069
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I) {
      const size_t p = d2(I, J);
      if (a.ivgtyp[p] == a.isice && a.xice[p] <= 0.0f) {
        for (int NS = 1; NS <= NSOIL; ++NS) {
          a.smois[dL(I, NS, 1, NSOIL, J)] = 1.0f;
          a.sh2o[dL(I, NS, 1, NSOIL, J)] = 0.0f;
          a.tslb[dL(I, NS, 1, NSOIL, J)] = MIN(a.tslb[dL(I, NS, 1, NSOIL, J)], 263.15f);
        }
        a.snow[p] = MAX(a.snow[p], 10.0f);
        a.snowh[p] = a.snow[p] * 0.01f;
      } else {
        const float BX = T.bb[a.isltyp[p] - 1];
        const float SMCMAX = T.maxsmc[a.isltyp[p] - 1];
        for (int NS = 1; NS <= NSOIL; ++NS)
          if (a.smois[dL(I, NS, 1, NSOIL, J)] > SMCMAX) a.smois[dL(I, NS, 1, NSOIL, J)] = SMCMAX;
        const float PSISAT = T.satpsi[a.isltyp[p] - 1];
        if (BX > 0.0f && SMCMAX > 0.0f && PSISAT > 0.0f) {
          for (int NS = 1; NS <= NSOIL; ++NS) {
            const size_t q = dL(I, NS, 1, NSOIL, J);
            if (a.tslb[q] < 273.149f) {
              float FK = POW((HLICE / (GRAV_I * (-PSISAT))) * ((a.tslb[q] - T0) / a.tslb[q]), -1.f / BX) * SMCMAX;
              FK = MAX(FK, 0.02f);
              a.sh2o[q] = MIN(FK, a.smois[q]);
            } else {
              a.sh2o[q] = a.smois[q];
            }
          }
        } else {
          for (int NS = 1; NS <= NSOIL; ++NS) a.sh2o[dL(I, NS, 1, NSOIL, J)] = a.smois[dL(I, NS, 1, NSOIL, J)];
        }
      }
    }
  // :1073-1120
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I) {
      const size_t p = d2(I, J);
      const bool melt = a.snow[p] > 0.0f && a.tsk[p] > 273.15f;
      a.tvxy[p] = melt ? 273.15f : a.tsk[p];
      a.tgxy[p] = melt ? 273.15f : a.tsk[p];
      a.canwat[p] = 0.0f;
      a.canliqxy[p] = a.canwat[p];
      a.canicexy[p] = 0.f;
      a.eahxy[p] = 2000.f;
      a.tahxy[p] = melt ? 273.15f : a.tsk[p];
      a.t2mvxy[p] = melt ? 273.15f : a.tsk[p];
      a.t2mbxy[p] = melt ? 273.15f : a.tsk[p];
      a.chstarxy[p] = 0.1f;
      a.cmxy[p] = 0.0f;
      a.chxy[p] = 0.0f;
      a.fwetxy[p] = 0.0f;
      a.sneqvoxy[p] = 0.0f;
      a.alboldxy[p] = 0.65f;
      a.qsnowxy[p] = 0.0f;
      a.wslakexy[p] = 0.0f;
      if (a.iopt_run != 5) {
        a.waxy[p] = 4900.f;
        a.wtxy[p] = a.waxy[p];
        a.zwtxy[p] = (25.f + 2.0f) - a.waxy[p] / 1000.f / 0.2f;
      } else {
        a.waxy[p] = 0.f;
        a.wtxy[p] = 0.f;
        if (!a.areaxy || !a.msftx || !a.msfty) return NOAHMP_ERR_ARG;
        a.areaxy[p] = (a.dx * a.dy) / (a.msftx[p] * a.msfty[p]);
      }
      a.lfmassxy[p] = 50.f;
      a.stmassxy[p] = 50.0f;
      a.rtmassxy[p] = 500.0f;
      a.woodxy[p] = 500.0f;
      a.stblcpxy[p] = 1000.0f;
      a.fastcpxy[p] = 1000.0f;
      a.xsaixy[p] = 0.1f;
    }
  ASoil ZSOIL;
  ZSOIL(1) = -a.dzs[0];
  for (int NS = 2; NS <= NSOIL; ++NS) ZSOIL(NS) = ZSOIL(NS - 1) - a.dzs[NS - 1];
  // ---- SNOW_INIT :1182-1283 (SWE = SNOW, SNODEP = SNOWH) ----
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I) {
      const size_t p = d2(I, J);
      const float SNODEP = a.snowh[p], SWE = a.snow[p];
      ASnow DZSNO;
      ASnSo DZSNSO;
      DZSNO.fill(0.f);
      DZSNSO.fill(0.f);
      int ISNOW;
      if (SNODEP < 0.025f) {
        ISNOW = 0;
      } else if (SNODEP >= 0.025f && SNODEP <= 0.05f) {
        ISNOW = -1;
        DZSNO(0) = SNODEP;
      } else if (SNODEP > 0.05f && SNODEP <= 0.10f) {
        ISNOW = -2;
        DZSNO(-1) = SNODEP / 2.f;
        DZSNO(0) = SNODEP / 2.f;
      } else if (SNODEP > 0.10f && SNODEP <= 0.25f) {
        ISNOW = -2;
        DZSNO(-1) = 0.05f;
        DZSNO(0) = SNODEP - DZSNO(-1);
      } else if (SNODEP > 0.25f && SNODEP <= 0.45f) {
        ISNOW = -3;
        DZSNO(-2) = 0.05f;
        DZSNO(-1) = 0.5f * (SNODEP - DZSNO(-2));
        DZSNO(0) = 0.5f * (SNODEP - DZSNO(-2));
      } else if (SNODEP > 0.45f) {
        ISNOW = -3;
        DZSNO(-2) = 0.05f;
        DZSNO(-1) = 0.20f;
        DZSNO(0) = SNODEP - DZSNO(-1) - DZSNO(-2);
      } else {
        return NOAHMP_ERR_ARG;  // NaN snow depth: "Problem with the logic assigning snow layers."
      }
      a.isnowxy[p] = ISNOW;
      for (int IZ = -NSNOW + 1; IZ <= 0; ++IZ) {
        a.tsnoxy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = 0.f;
        a.snicexy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = 0.f;
        a.snliqxy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = 0.f;
      }
      for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) {
        a.tsnoxy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = a.tgxy[p];
        a.snliqxy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = 0.00f;
        a.snicexy[dL(I, IZ, -NSNOW + 1, NSNOW, J)] = 1.00f * DZSNO(IZ) * (SWE / SNODEP);
      }
      for (int IZ = ISNOW + 1; IZ <= 0; ++IZ) DZSNSO(IZ) = -DZSNO(IZ);
      DZSNSO(1) = ZSOIL(1);
      for (int IZ = 2; IZ <= NSOIL; ++IZ) DZSNSO(IZ) = ZSOIL(IZ) - ZSOIL(IZ - 1);
      const int NZ = NSNOW + NSOIL;
      a.zsnsoxy[dL(I, ISNOW + 1, -NSNOW + 1, NZ, J)] = DZSNSO(ISNOW + 1);
      for (int IZ = ISNOW + 2; IZ <= NSOIL; ++IZ)
        a.zsnsoxy[dL(I, IZ, -NSNOW + 1, NZ, J)] = a.zsnsoxy[dL(I, IZ - 1, -NSNOW + 1, NZ, J)] + DZSNSO(IZ);
    }
  if (a.iopt_run != 5) return 0;
  // :1133-1172
  if (!(a.smoiseq && a.smcwtdxy && a.rechxy && a.deeprechxy && a.areaxy && a.msftx && a.msfty && a.stepwtd && a.qrfsxy &&
        a.qspringsxy && a.qslatxy && a.fdepthxy && a.ht && a.riverbedxy && a.eqzwt && a.rivercondxy && a.pexpxy))
    return NOAHMP_ERR_ARG;
  {
    // nint(): round half away from zero
    const float x = a.wtddt * 60.f / a.dt;
    int s = (int)(x >= 0.f ? std::floor(x + 0.5f) : -std::floor(-x + 0.5f));
    *a.stepwtd = IMAX(s, 1);
  }
  // ---- GROUNDWATER_INIT :1286-1470 ----
  const int nj = a.jme - a.jms + 1;
  FA<1, NSOIL> DZS;
  for (int K = 1; K <= NSOIL; ++K) DZS(K) = a.dzs[K - 1];
  const float DELTAT = a.wtddt * 60.f;
  std::vector<int> LANDMASK((size_t)ni * nj);
  std::vector<float> QLAT((size_t)ni * nj, 0.f), QRF((size_t)ni * nj, 0.f), KCELL((size_t)ni * nj, 0.f),
      HEAD((size_t)ni * nj, 0.f);
  for (size_t p = 0; p < (size_t)ni * nj; ++p)
    LANDMASK[p] = (a.ivgtyp[p] != a.iswater && a.ivgtyp[p] != a.isice) ? 1 : -1;
  float* WTD = a.zwtxy;
  {  // LATERALFLOW (groundwater.F90:201-295) with the caller's ids..kte
    const float FANGLE = 0.45508986056f;
    int itsh = IMAX(a.its - 1, a.ids), iteh = IMIN(a.ite + 1, a.ide - 1);
    int jtsh = IMAX(a.jts - 1, a.jds), jteh = IMIN(a.jte + 1, a.jde - 1);
    for (int J = jtsh; J <= jteh; ++J)
      for (int I = itsh; I <= iteh; ++I) {
        const size_t p = d2(I, J);
        if (a.fdepthxy[p] > 0.f) {
          const float KLAT = T.satdk[a.isltyp[p] - 1] * KLATFACTOR_I[a.isltyp[p] - 1];
          if (WTD[p] < -1.5f) KCELL[p] = a.fdepthxy[p] * KLAT * EXP((WTD[p] + 1.5f) / a.fdepthxy[p]);
          else KCELL[p] = KLAT * (WTD[p] + 1.5f + a.fdepthxy[p]);
        } else {
          KCELL[p] = 0.f;
        }
        HEAD[p] = a.ht[p] + WTD[p];
      }
    itsh = IMAX(a.its, a.ids + 1); iteh = IMIN(a.ite, a.ide - 2);
    jtsh = IMAX(a.jts, a.jds + 1); jteh = IMIN(a.jte, a.jde - 2);
    const float SQRT2 = SQRT(2.f);
    for (int J = jtsh; J <= jteh; ++J)
      for (int I = itsh; I <= iteh; ++I) {
        const size_t p = d2(I, J);
        if (LANDMASK[p] > 0) {
          float Q = 0.f;
          auto KC = [&](int i, int j) { return KCELL[d2(i, j)]; };
          auto HD = [&](int i, int j) { return HEAD[d2(i, j)]; };
          Q = Q + (KC(I - 1, J + 1) + KC(I, J)) * (HD(I - 1, J + 1) - HD(I, J)) / SQRT2;
          Q = Q + (KC(I - 1, J) + KC(I, J)) * (HD(I - 1, J) - HD(I, J));
          Q = Q + (KC(I - 1, J - 1) + KC(I, J)) * (HD(I - 1, J - 1) - HD(I, J)) / SQRT2;
          Q = Q + (KC(I, J + 1) + KC(I, J)) * (HD(I, J + 1) - HD(I, J));
          Q = Q + (KC(I, J - 1) + KC(I, J)) * (HD(I, J - 1) - HD(I, J));
          Q = Q + (KC(I + 1, J + 1) + KC(I, J)) * (HD(I + 1, J + 1) - HD(I, J)) / SQRT2;
          Q = Q + (KC(I + 1, J) + KC(I, J)) * (HD(I + 1, J) - HD(I, J));
          Q = Q + (KC(I + 1, J - 1) + KC(I, J)) * (HD(I + 1, J - 1) - HD(I, J)) / SQRT2;
          QLAT[p] = FANGLE * Q * DELTAT / a.areaxy[p];
        }
      }
  }
  // :1352-1368
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I) {
      const size_t p = d2(I, J);
      if (LANDMASK[p] > 0) {
        float RCOND;
        if (WTD[p] > a.riverbedxy[p] && a.eqzwt[p] > a.riverbedxy[p])
          RCOND = a.rivercondxy[p] * EXP(a.pexpxy[p] * (WTD[p] - a.eqzwt[p]));
        else
          RCOND = a.rivercondxy[p];
        QRF[p] = RCOND * (WTD[p] - a.riverbedxy[p]) * DELTAT / a.areaxy[p];
        QRF[p] = MAX(QRF[p], 0.f);
      } else {
        QRF[p] = 0.f;
      }
    }
  // :1372-1462
  for (int J = a.jts; J <= jtf; ++J)
    for (int I = a.its; I <= itf; ++I) {
      const size_t p = d2(I, J);
      const float BX = T.bb[a.isltyp[p] - 1];
      float SMCMAX = T.maxsmc[a.isltyp[p] - 1];
      float SMCWLT = T.wltsmc[a.isltyp[p] - 1];
      if (a.ivgtyp[p] == a.isurban) { SMCMAX = 0.45f; SMCWLT = 0.40f; }
      const float DWSAT = T.satdw[a.isltyp[p] - 1];
      const float DKSAT = T.satdk[a.isltyp[p] - 1];
      const float PSISAT = -T.satpsi[a.isltyp[p] - 1];
      if (BX > 0.0f && SMCMAX > 0.0f && -PSISAT > 0.0f) {
        ASoil SMCEQ;
        EQSMOISTURE(ZSOIL, SMCMAX, SMCWLT, DWSAT, DKSAT, BX, SMCEQ);
        for (int K = 1; K <= NSOIL; ++K) a.smoiseq[dL(I, K, 1, NSOIL, J)] = SMCEQ(K);
        if (WTD[p] < ZSOIL(NSOIL) - DZS(NSOIL)) {
          const float EXPON = 2.f * BX + 3.f;
          const float DDZ = ZSOIL(NSOIL) - WTD[p];
          const float CC = PSISAT / DDZ;
          const float FLUX = (QLAT[p] - QRF[p]) / DELTAT;
          float SMC = 0.5f * SMCMAX;
          for (int ITER = 1; ITER <= 100; ++ITER) {
            const float DD = (SMC + SMCMAX) / (2.f * SMCMAX);
            const float AA = -DKSAT * POW(DD, EXPON);
            const float BBB = CC * (POW(SMCMAX / SMC, BX) - 1.f) + 1.f;
            const float FUNC = AA * BBB - FLUX;
            const float DFUNC = -DKSAT * (EXPON / (2.f * SMCMAX)) * POW(DD, EXPON - 1.f) * BBB +
                                AA * CC * (-BX) * POW(SMCMAX, BX) * POW(SMC, -BX - 1.f);
            const float DX = FUNC / DFUNC;
            SMC = SMC - DX;
            if (ABS(DX) < 1.E-6f) break;
          }
          a.smcwtdxy[p] = MAX(SMC, 1.E-4f);
        } else if (WTD[p] < ZSOIL(NSOIL)) {
          float SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS(NSOIL)), 1.f / BX);
          SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
          a.smcwtdxy[p] = SMCMAX * (WTD[p] - (ZSOIL(NSOIL) - DZS(NSOIL))) + SMCEQDEEP * (ZSOIL(NSOIL) - WTD[p]);
        } else {
          a.smcwtdxy[p] = SMCMAX;
          for (int K = NSOIL; K >= 2; --K) {
            const size_t q = dL(I, K, 1, NSOIL, J);
            if (WTD[p] >= ZSOIL(K - 1)) {
              const float FRLIQ = a.sh2o[q] / a.smois[q];
              a.smois[q] = SMCMAX;
              a.sh2o[q] = SMCMAX * FRLIQ;
            } else {
              if (a.smois[q] < SMCEQ(K)) WTD[p] = ZSOIL(K);
              else WTD[p] = (a.smois[q] * DZS(K) - SMCEQ(K) * ZSOIL(K - 1) + SMCMAX * ZSOIL(K)) / (SMCMAX - SMCEQ(K));
              break;
            }
          }
        }
      } else {
        for (int K = 1; K <= NSOIL; ++K) a.smoiseq[dL(I, K, 1, NSOIL, J)] = SMCMAX;
        a.smcwtdxy[p] = SMCMAX;
        WTD[p] = 0.f;
      }
      a.deeprechxy[p] = 0.f;
      a.rechxy[p] = 0.f;
      a.qslatxy[p] = 0.f;
      a.qrfsxy[p] = 0.f;
      a.qspringsxy[p] = 0.f;
    }
  return 0;
}

}  // namespace nmo

extern "C" int nmo_init(const noahmp_init_args* args, const noahmp_tables* tables) { return nmo::INIT(*args, *tables); }
