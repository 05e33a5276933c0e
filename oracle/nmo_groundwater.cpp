// nmo_groundwater.cpp — ORACLE (test infrastructure): WTABLE_mmf_noahmp, LATERALFLOW, UPDATEWTD.
// Restates phys/module_sf_noahmp_groundwater.F90:14-606 in source order, fp32.
// Pinned bit for bit against the machine-translated reference (see nmo.h; tests/test_reference_pin.py::test_wtable_*).
#include <vector>
#include "nmo.h"

namespace nmo {

// groundwater.F90:298-606.  ZSOIL is (0:NSOIL), the other arrays (1:NSOIL).
static void UPDATEWTD(const FA<1, NSOIL>& DZS, const FA<0, NSOIL>& ZSOIL, const ASoil& SMCEQ, float SMCMAX, float SMCWLT,
                      float PSISAT, float BEXP, float& TOTWATER, float& WTD, ASoil& SMC, ASoil& SH2O, float& SMCWTD,
                      float& QSPRING) {
  (void)SMCWLT;
  int K, K1, IWTD, KWTD;
  float MAXWATUP, MAXWATDW, WTDOLD, WGPMID, SYIELDDW, DZUP, SMCEQDEEP;
  ASoil SICE;
  QSPRING = 0.f;
  for (K = 1; K <= NSOIL; ++K) SICE(K) = SMC(K) - SH2O(K);
  IWTD = 1;
  if (TOTWATER > 0.f) {
    if (WTD >= ZSOIL(NSOIL)) {
      for (K = NSOIL - 1; K >= 1; --K)
        if (WTD < ZSOIL(K)) break;
      IWTD = K;
      KWTD = IWTD + 1;
      MAXWATUP = DZS(KWTD) * (SMCMAX - SMC(KWTD));
      if (TOTWATER <= MAXWATUP) {
        SMC(KWTD) = SMC(KWTD) + TOTWATER / DZS(KWTD);
        SMC(KWTD) = MIN(SMC(KWTD), SMCMAX);
        if (SMC(KWTD) > SMCEQ(KWTD))
          WTD = MIN((SMC(KWTD) * DZS(KWTD) - SMCEQ(KWTD) * ZSOIL(IWTD) + SMCMAX * ZSOIL(KWTD)) / (SMCMAX - SMCEQ(KWTD)),
                    ZSOIL(IWTD));
        TOTWATER = 0.f;
      } else {
        SMC(KWTD) = SMCMAX;
        TOTWATER = TOTWATER - MAXWATUP;
        K1 = IWTD;
        for (K = K1; K >= 0; --K) {
          WTD = ZSOIL(K);
          IWTD = K - 1;
          if (K == 0) break;
          MAXWATUP = DZS(K) * (SMCMAX - SMC(K));
          if (TOTWATER <= MAXWATUP) {
            SMC(K) = SMC(K) + TOTWATER / DZS(K);
            SMC(K) = MIN(SMC(K), SMCMAX);
            if (SMC(K) > SMCEQ(K))
              WTD = MIN((SMC(K) * DZS(K) - SMCEQ(K) * ZSOIL(IWTD) + SMCMAX * ZSOIL(K)) / (SMCMAX - SMCEQ(K)), ZSOIL(IWTD));
            TOTWATER = 0.f;
            break;
          } else {
            SMC(K) = SMCMAX;
            TOTWATER = TOTWATER - MAXWATUP;
          }
        }
      }
    } else if (WTD >= ZSOIL(NSOIL) - DZS(NSOIL)) {
      SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS(NSOIL)), 1.f / BEXP);
      SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
      MAXWATUP = (SMCMAX - SMCWTD) * DZS(NSOIL);
      if (TOTWATER <= MAXWATUP) {
        SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
        SMCWTD = MIN(SMCWTD, SMCMAX);
        if (SMCWTD > SMCEQDEEP)
          WTD = MIN((SMCWTD * DZS(NSOIL) - SMCEQDEEP * ZSOIL(NSOIL) + SMCMAX * (ZSOIL(NSOIL) - DZS(NSOIL))) /
                        (SMCMAX - SMCEQDEEP),
                    ZSOIL(NSOIL));
        TOTWATER = 0.f;
      } else {
        SMCWTD = SMCMAX;
        TOTWATER = TOTWATER - MAXWATUP;
        for (K = NSOIL; K >= 0; --K) {
          WTD = ZSOIL(K);
          IWTD = K - 1;
          if (K == 0) break;
          MAXWATUP = DZS(K) * (SMCMAX - SMC(K));
          if (TOTWATER <= MAXWATUP) {
            SMC(K) = MIN(SMC(K) + TOTWATER / DZS(K), SMCMAX);
            if (SMC(K) > SMCEQ(K))
              WTD = MIN((SMC(K) * DZS(K) - SMCEQ(K) * ZSOIL(IWTD) + SMCMAX * ZSOIL(K)) / (SMCMAX - SMCEQ(K)), ZSOIL(IWTD));
            TOTWATER = 0.f;
            break;
          } else {
            SMC(K) = SMCMAX;
            TOTWATER = TOTWATER - MAXWATUP;
          }
        }
      }
    } else {
      MAXWATUP = (SMCMAX - SMCWTD) * (ZSOIL(NSOIL) - DZS(NSOIL) - WTD);
      if (TOTWATER <= MAXWATUP) {
        WTD = WTD + TOTWATER / (SMCMAX - SMCWTD);
        TOTWATER = 0.f;
      } else {
        TOTWATER = TOTWATER - MAXWATUP;
        WTD = ZSOIL(NSOIL) - DZS(NSOIL);
        MAXWATUP = (SMCMAX - SMCWTD) * DZS(NSOIL);
        if (TOTWATER <= MAXWATUP) {
          SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS(NSOIL)), 1.f / BEXP);
          SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
          SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
          SMCWTD = MIN(SMCWTD, SMCMAX);
          WTD = (SMCWTD * DZS(NSOIL) - SMCEQDEEP * ZSOIL(NSOIL) + SMCMAX * (ZSOIL(NSOIL) - DZS(NSOIL))) /
                (SMCMAX - SMCEQDEEP);
          TOTWATER = 0.f;
        } else {
          SMCWTD = SMCMAX;
          TOTWATER = TOTWATER - MAXWATUP;
          for (K = NSOIL; K >= 0; --K) {
            WTD = ZSOIL(K);
            IWTD = K - 1;
            if (K == 0) break;
            MAXWATUP = DZS(K) * (SMCMAX - SMC(K));
            if (TOTWATER <= MAXWATUP) {
              SMC(K) = SMC(K) + TOTWATER / DZS(K);
              SMC(K) = MIN(SMC(K), SMCMAX);
              if (SMC(K) > SMCEQ(K))
                WTD = (SMC(K) * DZS(K) - SMCEQ(K) * ZSOIL(IWTD) + SMCMAX * ZSOIL(K)) / (SMCMAX - SMCEQ(K));
              TOTWATER = 0.f;
              break;
            } else {
              SMC(K) = SMCMAX;
              TOTWATER = TOTWATER - MAXWATUP;
            }
          }
        }
      }
    }
    QSPRING = TOTWATER;
  } else if (TOTWATER < 0.f) {
    if (WTD >= ZSOIL(NSOIL)) {
      for (K = NSOIL - 1; K >= 1; --K)
        if (WTD < ZSOIL(K)) break;
      IWTD = K;
      K1 = IWTD + 1;
      for (KWTD = K1; KWTD <= NSOIL; ++KWTD) {
        MAXWATDW = DZS(KWTD) * (SMC(KWTD) - MAX(SMCEQ(KWTD), SICE(KWTD)));
        if (-TOTWATER <= MAXWATDW) {
          SMC(KWTD) = SMC(KWTD) + TOTWATER / DZS(KWTD);
          if (SMC(KWTD) > SMCEQ(KWTD)) {
            WTD = (SMC(KWTD) * DZS(KWTD) - SMCEQ(KWTD) * ZSOIL(IWTD) + SMCMAX * ZSOIL(KWTD)) / (SMCMAX - SMCEQ(KWTD));
          } else {
            WTD = ZSOIL(KWTD);
            IWTD = IWTD + 1;
          }
          TOTWATER = 0.f;
          break;
        } else {
          WTD = ZSOIL(KWTD);
          IWTD = IWTD + 1;
          if (MAXWATDW >= 0.f) {
            SMC(KWTD) = SMC(KWTD) + MAXWATDW / DZS(KWTD);
            TOTWATER = TOTWATER + MAXWATDW;
          }
        }
      }
      if (IWTD == NSOIL && TOTWATER < 0.f) {
        SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS(NSOIL)), 1.f / BEXP);
        SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
        MAXWATDW = DZS(NSOIL) * (SMCWTD - SMCEQDEEP);
        if (-TOTWATER <= MAXWATDW) {
          SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
          WTD = MAX((SMCWTD * DZS(NSOIL) - SMCEQDEEP * ZSOIL(NSOIL) + SMCMAX * (ZSOIL(NSOIL) - DZS(NSOIL))) /
                        (SMCMAX - SMCEQDEEP),
                    ZSOIL(NSOIL) - DZS(NSOIL));
        } else {
          WTD = ZSOIL(NSOIL) - DZS(NSOIL);
          SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
          DZUP = (SMCEQDEEP - SMCWTD) * DZS(NSOIL) / (SMCMAX - SMCEQDEEP);
          WTD = WTD - DZUP;
          SMCWTD = SMCEQDEEP;
        }
      }
    } else if (WTD >= ZSOIL(NSOIL) - DZS(NSOIL)) {
      SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS(NSOIL)), 1.f / BEXP);
      SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
      MAXWATDW = DZS(NSOIL) * (SMCWTD - SMCEQDEEP);
      if (-TOTWATER <= MAXWATDW) {
        SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
        WTD = MAX((SMCWTD * DZS(NSOIL) - SMCEQDEEP * ZSOIL(NSOIL) + SMCMAX * (ZSOIL(NSOIL) - DZS(NSOIL))) /
                      (SMCMAX - SMCEQDEEP),
                  ZSOIL(NSOIL) - DZS(NSOIL));
      } else {
        WTD = ZSOIL(NSOIL) - DZS(NSOIL);
        SMCWTD = SMCWTD + TOTWATER / DZS(NSOIL);
        DZUP = (SMCEQDEEP - SMCWTD) * DZS(NSOIL) / (SMCMAX - SMCEQDEEP);
        WTD = WTD - DZUP;
        SMCWTD = SMCEQDEEP;
      }
    } else {
      WGPMID = SMCMAX * POW(PSISAT / (PSISAT - (ZSOIL(NSOIL) - WTD)), 1.f / BEXP);
      WGPMID = MAX(WGPMID, 1.E-4f);
      SYIELDDW = SMCMAX - WGPMID;
      WTDOLD = WTD;
      WTD = WTDOLD + TOTWATER / SYIELDDW;
      SMCWTD = (SMCWTD * (ZSOIL(NSOIL) - WTDOLD) + WGPMID * (WTDOLD - WTD)) / (ZSOIL(NSOIL) - WTD);
    }
    QSPRING = 0.f;
  }
  for (K = 1; K <= NSOIL; ++K) SH2O(K) = SMC(K) - SICE(K);
}

static const float KLATFACTOR[19] = {2.f, 3.f, 4.f, 10.f, 10.f, 12.f, 14.f, 20.f, 24.f, 28.f,
                                     40.f, 48.f, 2.f, 0.f, 10.f, 0.f, 20.f, 2.f, 2.f};

// groundwater.F90:14-198 with LATERALFLOW (:201-295) inlined at its call site
static int WTABLE(const noahmp_wtable_args& a, const noahmp_tables& T) {
  if (a.nsoil != NSOIL) return NOAHMP_ERR_ARG;
  const int ni = a.ime - a.ims + 1, nj = a.jme - a.jms + 1;
  auto d2 = [&](int I, int J) { return (size_t)(I - a.ims) + (size_t)(J - a.jms) * ni; };
  auto d3 = [&](int I, int K, int J) { return (size_t)(I - a.ims) + (size_t)(K - 1) * ni + (size_t)(J - a.jms) * ni * NSOIL; };
  const float DELTAT = a.wtddt * 60.f;
  FA<0, NSOIL> ZSOIL;
  FA<1, NSOIL> DZS;
  for (int K = 1; K <= NSOIL; ++K) DZS(K) = a.dzs[K - 1];
  ZSOIL(0) = 0.f;
  ZSOIL(1) = -DZS(1);
  for (int K = 2; K <= NSOIL; ++K) ZSOIL(K) = -DZS(K) + ZSOIL(K - 1);
  std::vector<int> LANDMASK((size_t)ni * nj);
  std::vector<float> QLAT((size_t)ni * nj, 0.f), KCELL((size_t)ni * nj, 0.f), HEAD((size_t)ni * nj, 0.f);
  for (int J = a.jms; J <= a.jme; ++J)
    for (int I = a.ims; I <= a.ime; ++I) {
      size_t p = d2(I, J);
      LANDMASK[p] = ((a.xland[p] - 1.5f) < 0.f && a.xice[p] < a.xice_threshold && a.ivgtyp[p] != a.isice) ? 1 : -1;
    }
  // ---- LATERALFLOW ----
  {
    const float FANGLE = 0.45508986056f;
    int itsh = IMAX(a.its - 1, a.ids), iteh = IMIN(a.ite + 1, a.ide - 1);
    int jtsh = IMAX(a.jts - 1, a.jds), jteh = IMIN(a.jte + 1, a.jde - 1);
    for (int J = jtsh; J <= jteh; ++J)
      for (int I = itsh; I <= iteh; ++I) {
        size_t p = d2(I, J);
        if (a.fdepth[p] > 0.f) {
          float KLAT = T.satdk[a.isltyp[p] - 1] * KLATFACTOR[a.isltyp[p] - 1];
          if (a.wtd[p] < -1.5f) KCELL[p] = a.fdepth[p] * KLAT * EXP((a.wtd[p] + 1.5f) / a.fdepth[p]);
          else KCELL[p] = KLAT * (a.wtd[p] + 1.5f + a.fdepth[p]);
        } else {
          KCELL[p] = 0.f;
        }
        HEAD[p] = a.topo[p] + a.wtd[p];
      }
    itsh = IMAX(a.its, a.ids + 1); iteh = IMIN(a.ite, a.ide - 2);
    jtsh = IMAX(a.jts, a.jds + 1); jteh = IMIN(a.jte, a.jde - 2);
    const float SQRT2 = SQRT(2.f);
    for (int J = jtsh; J <= jteh; ++J)
      for (int I = itsh; I <= iteh; ++I) {
        size_t p = d2(I, J);
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
          QLAT[p] = FANGLE * Q * DELTAT / a.area[p];
        }
      }
  }
  // ---- river flux ----
  for (int J = a.jts; J <= a.jte; ++J)
    for (int I = a.its; I <= a.ite; ++I) {
      size_t p = d2(I, J);
      if (LANDMASK[p] > 0) {
        float RCOND;
        if (a.wtd[p] > a.riverbed[p] && a.eqwtd[p] > a.riverbed[p])
          RCOND = a.rivercond[p] * EXP(a.pexp[p] * (a.wtd[p] - a.eqwtd[p]));
        else
          RCOND = a.rivercond[p];
        a.qrf[p] = RCOND * (a.wtd[p] - a.riverbed[p]) * DELTAT / a.area[p];
        a.qrf[p] = MAX(a.qrf[p], 0.f);
      } else {
        a.qrf[p] = 0.f;
      }
    }
  // ---- deep recharge + UPDATEWTD ----
  for (int J = a.jts; J <= a.jte; ++J)
    for (int I = a.its; I <= a.ite; ++I) {
      size_t p = d2(I, J);
      // QSPRING is INTENT(OUT) and only assigned at land cells in the reference; defined as 0 elsewhere here
      a.qspring[p] = 0.f;
      if (LANDMASK[p] > 0) {
        const int st = a.isltyp[p];
        float BEXP = T.bb[st - 1], DKSAT = T.satdk[st - 1], SMCMAX = T.maxsmc[st - 1], PSISAT = -T.satpsi[st - 1],
              SMCWLT = T.wltsmc[st - 1];
        if (a.ivgtyp[p] == a.isurban) { SMCMAX = 0.45f; SMCWLT = 0.40f; }
        if (a.wtd[p] < ZSOIL(NSOIL) - DZS(NSOIL)) {
          float DDZ = ZSOIL(NSOIL) - a.wtd[p];
          float SMCWTDMID = 0.5f * (a.smcwtd[p] + SMCMAX);
          float PSI = PSISAT * POW(SMCMAX / a.smcwtd[p], BEXP);
          float WCNDDEEP = DKSAT * POW(SMCWTDMID / SMCMAX, 2.0f * BEXP + 3.0f);
          float WFLUXDEEP = -DELTAT * WCNDDEEP * ((PSISAT - PSI) / DDZ - 1.f);
          a.smcwtd[p] = a.smcwtd[p] + (a.deeprech[p] - WFLUXDEEP) / DDZ;
          float WPLUS = MAX((a.smcwtd[p] - SMCMAX), 0.0f) * DDZ;
          float WMINUS = MAX((1.E-4f - a.smcwtd[p]), 0.0f) * DDZ;
          a.smcwtd[p] = MAX(MIN(a.smcwtd[p], SMCMAX), 1.E-4f);
          WFLUXDEEP = WFLUXDEEP + WPLUS - WMINUS;
          a.deeprech[p] = WFLUXDEEP;
        }
        float TOTWATER = QLAT[p] - a.qrf[p] + a.deeprech[p];
        ASoil SMC, SH2O, SMCEQ;
        for (int K = 1; K <= NSOIL; ++K) {
          SMC(K) = a.smois[d3(I, K, J)];
          SH2O(K) = a.sh2oxy[d3(I, K, J)];
          SMCEQ(K) = a.smoiseq[d3(I, K, J)];
        }
        UPDATEWTD(DZS, ZSOIL, SMCEQ, SMCMAX, SMCWLT, PSISAT, BEXP, TOTWATER, a.wtd[p], SMC, SH2O, a.smcwtd[p],
                  a.qspring[p]);
        for (int K = 1; K <= NSOIL; ++K) {
          a.smois[d3(I, K, J)] = SMC(K);
          a.sh2oxy[d3(I, K, J)] = SH2O(K);
        }
      }
    }
  // ---- accumulate ----
  for (int J = a.jts; J <= a.jte; ++J)
    for (int I = a.its; I <= a.ite; ++I) {
      size_t p = d2(I, J);
      a.qslat[p] = a.qslat[p] + QLAT[p] * 1.E3f;
      a.qrfs[p] = a.qrfs[p] + a.qrf[p] * 1.E3f;
      a.qsprings[p] = a.qsprings[p] + a.qspring[p] * 1.E3f;
      a.rech[p] = a.rech[p] + a.deeprech[p] * 1.E3f;
      a.deeprech[p] = 0.f;
    }
  return 0;
}

}  // namespace nmo

extern "C" int nmo_wtable(const noahmp_wtable_args* args, const noahmp_tables* tables) {
  return nmo::WTABLE(*args, *tables);
}
