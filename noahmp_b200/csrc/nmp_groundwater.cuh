// nmp_groundwater.cuh — opt_run = 5 (Miguez-Macho & Fan) groundwater step on the device:
// WTABLE_mmf_noahmp, LATERALFLOW, UPDATEWTD (phys/module_sf_noahmp_groundwater.F90:14-606).
//
// LATERALFLOW is the only non-column-local piece of the whole path: pass 1 (KCELL, HEAD per cell) writes two
// grid-order planes with a one-cell halo ring, pass 2 (8-neighbour flux) is evaluated by the thread that owns the
// land column, which then runs the river flux, the deep-recharge update and UPDATEWTD on the column's planes of
// the compact state.  Included by both physics builds (fast / parity).
#pragma once
#include "nmp_common.cuh"
#include "nmp_fields.h"

namespace {

using namespace nmp;

__device__ __constant__ float kKLATFACTOR[19] = {2.f, 3.f, 4.f, 10.f, 10.f, 12.f, 14.f, 20.f, 24.f, 28.f,
                                                 40.f, 48.f, 2.f, 0.f, 10.f, 0.f, 20.f, 2.f, 2.f};

// LATERALFLOW pass 1 (:236-252) on the cells of this tile that lie inside max(its-1,ids)..min(ite+1,ide-1) (the
// ring cells themselves belong to the neighbouring tiles and arrive by halo exchange)
__global__ void wt_head_kernel(const nmpf::WtParams w) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (long long)w.ni * w.nj) return;
  const int il = (int)(c % w.ni), jl = (int)(c / w.ni);
  const int I = w.its + il, J = w.jts + jl;
  float kc = 0.f, hd = 0.f;
  if (I >= w.ids && I <= w.ide - 1 && J >= w.jds && J <= w.jde - 1) {
    const float fd = w.fdepth[c], wtd = w.wtd_grid[c];
    if (fd > 0.f) {
      const int st = __float_as_int(w.isltyp[c]);
      const float KLAT = w.tables->satdk[st - 1] * kKLATFACTOR[st - 1];
      if (wtd < -1.5f) kc = fd * KLAT * EXP((wtd + 1.5f) / fd);
      else kc = KLAT * (wtd + 1.5f + fd);
    }
    hd = w.topo[c] + wtd;
  }
  const long long h = (long long)(il + 1) + (long long)(jl + 1) * (w.ni + 2);
  w.kcell[h] = kc;
  w.head[h] = hd;
}

// UPDATEWTD (:298-606).  ZSOIL(0:4), others (1:4); small dynamically indexed local arrays.
__device__ void UPDATEWTD(const float* DZS, const float* ZSOIL, const float* SMCEQ, float SMCMAX, float PSISAT,
                          float BEXP, float& TOTWATER, float& WTD, float* SMC, float* SH2O, float& SMCWTD,
                          float& QSPRING) {
  const int NS = NSOIL;
  int K, K1, IWTD, KWTD;
  float MAXWATUP, MAXWATDW, WTDOLD, WGPMID, SYIELDDW, DZUP, SMCEQDEEP;
  float SICE[NSOIL + 1];
  QSPRING = 0.f;
  for (K = 1; K <= NS; ++K) SICE[K] = SMC[K] - SH2O[K];
  IWTD = 1;
  if (TOTWATER > 0.f) {
    if (WTD >= ZSOIL[NS]) {
      for (K = NS - 1; K >= 1; --K)
        if (WTD < ZSOIL[K]) break;
      IWTD = K;
      KWTD = IWTD + 1;
      MAXWATUP = DZS[KWTD] * (SMCMAX - SMC[KWTD]);
      if (TOTWATER <= MAXWATUP) {
        SMC[KWTD] = SMC[KWTD] + TOTWATER / DZS[KWTD];
        SMC[KWTD] = MIN(SMC[KWTD], SMCMAX);
        if (SMC[KWTD] > SMCEQ[KWTD])
          WTD = MIN((SMC[KWTD] * DZS[KWTD] - SMCEQ[KWTD] * ZSOIL[IWTD] + SMCMAX * ZSOIL[KWTD]) / (SMCMAX - SMCEQ[KWTD]),
                    ZSOIL[IWTD]);
        TOTWATER = 0.f;
      } else {
        SMC[KWTD] = SMCMAX;
        TOTWATER = TOTWATER - MAXWATUP;
        K1 = IWTD;
        for (K = K1; K >= 0; --K) {
          WTD = ZSOIL[K];
          IWTD = K - 1;
          if (K == 0) break;
          MAXWATUP = DZS[K] * (SMCMAX - SMC[K]);
          if (TOTWATER <= MAXWATUP) {
            SMC[K] = SMC[K] + TOTWATER / DZS[K];
            SMC[K] = MIN(SMC[K], SMCMAX);
            if (SMC[K] > SMCEQ[K])
              WTD = MIN((SMC[K] * DZS[K] - SMCEQ[K] * ZSOIL[IWTD] + SMCMAX * ZSOIL[K]) / (SMCMAX - SMCEQ[K]), ZSOIL[IWTD]);
            TOTWATER = 0.f;
            break;
          } else {
            SMC[K] = SMCMAX;
            TOTWATER = TOTWATER - MAXWATUP;
          }
        }
      }
    } else if (WTD >= ZSOIL[NS] - DZS[NS]) {
      SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS[NS]), 1.f / BEXP);
      SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
      MAXWATUP = (SMCMAX - SMCWTD) * DZS[NS];
      if (TOTWATER <= MAXWATUP) {
        SMCWTD = SMCWTD + TOTWATER / DZS[NS];
        SMCWTD = MIN(SMCWTD, SMCMAX);
        if (SMCWTD > SMCEQDEEP)
          WTD = MIN((SMCWTD * DZS[NS] - SMCEQDEEP * ZSOIL[NS] + SMCMAX * (ZSOIL[NS] - DZS[NS])) / (SMCMAX - SMCEQDEEP),
                    ZSOIL[NS]);
        TOTWATER = 0.f;
      } else {
        SMCWTD = SMCMAX;
        TOTWATER = TOTWATER - MAXWATUP;
        for (K = NS; K >= 0; --K) {
          WTD = ZSOIL[K];
          IWTD = K - 1;
          if (K == 0) break;
          MAXWATUP = DZS[K] * (SMCMAX - SMC[K]);
          if (TOTWATER <= MAXWATUP) {
            SMC[K] = MIN(SMC[K] + TOTWATER / DZS[K], SMCMAX);
            if (SMC[K] > SMCEQ[K])
              WTD = MIN((SMC[K] * DZS[K] - SMCEQ[K] * ZSOIL[IWTD] + SMCMAX * ZSOIL[K]) / (SMCMAX - SMCEQ[K]), ZSOIL[IWTD]);
            TOTWATER = 0.f;
            break;
          } else {
            SMC[K] = SMCMAX;
            TOTWATER = TOTWATER - MAXWATUP;
          }
        }
      }
    } else {
      MAXWATUP = (SMCMAX - SMCWTD) * (ZSOIL[NS] - DZS[NS] - WTD);
      if (TOTWATER <= MAXWATUP) {
        WTD = WTD + TOTWATER / (SMCMAX - SMCWTD);
        TOTWATER = 0.f;
      } else {
        TOTWATER = TOTWATER - MAXWATUP;
        WTD = ZSOIL[NS] - DZS[NS];
        MAXWATUP = (SMCMAX - SMCWTD) * DZS[NS];
        if (TOTWATER <= MAXWATUP) {
          SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS[NS]), 1.f / BEXP);
          SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
          SMCWTD = SMCWTD + TOTWATER / DZS[NS];
          SMCWTD = MIN(SMCWTD, SMCMAX);
          WTD = (SMCWTD * DZS[NS] - SMCEQDEEP * ZSOIL[NS] + SMCMAX * (ZSOIL[NS] - DZS[NS])) / (SMCMAX - SMCEQDEEP);
          TOTWATER = 0.f;
        } else {
          SMCWTD = SMCMAX;
          TOTWATER = TOTWATER - MAXWATUP;
          for (K = NS; K >= 0; --K) {
            WTD = ZSOIL[K];
            IWTD = K - 1;
            if (K == 0) break;
            MAXWATUP = DZS[K] * (SMCMAX - SMC[K]);
            if (TOTWATER <= MAXWATUP) {
              SMC[K] = SMC[K] + TOTWATER / DZS[K];
              SMC[K] = MIN(SMC[K], SMCMAX);
              if (SMC[K] > SMCEQ[K])
                WTD = (SMC[K] * DZS[K] - SMCEQ[K] * ZSOIL[IWTD] + SMCMAX * ZSOIL[K]) / (SMCMAX - SMCEQ[K]);
              TOTWATER = 0.f;
              break;
            } else {
              SMC[K] = SMCMAX;
              TOTWATER = TOTWATER - MAXWATUP;
            }
          }
        }
      }
    }
    QSPRING = TOTWATER;
  } else if (TOTWATER < 0.f) {
    if (WTD >= ZSOIL[NS]) {
      for (K = NS - 1; K >= 1; --K)
        if (WTD < ZSOIL[K]) break;
      IWTD = K;
      K1 = IWTD + 1;
      for (KWTD = K1; KWTD <= NS; ++KWTD) {
        MAXWATDW = DZS[KWTD] * (SMC[KWTD] - MAX(SMCEQ[KWTD], SICE[KWTD]));
        if (-TOTWATER <= MAXWATDW) {
          SMC[KWTD] = SMC[KWTD] + TOTWATER / DZS[KWTD];
          if (SMC[KWTD] > SMCEQ[KWTD]) {
            WTD = (SMC[KWTD] * DZS[KWTD] - SMCEQ[KWTD] * ZSOIL[IWTD] + SMCMAX * ZSOIL[KWTD]) / (SMCMAX - SMCEQ[KWTD]);
          } else {
            WTD = ZSOIL[KWTD];
            IWTD = IWTD + 1;
          }
          TOTWATER = 0.f;
          break;
        } else {
          WTD = ZSOIL[KWTD];
          IWTD = IWTD + 1;
          if (MAXWATDW >= 0.f) {
            SMC[KWTD] = SMC[KWTD] + MAXWATDW / DZS[KWTD];
            TOTWATER = TOTWATER + MAXWATDW;
          }
        }
      }
      if (IWTD == NS && TOTWATER < 0.f) {
        SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS[NS]), 1.f / BEXP);
        SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
        MAXWATDW = DZS[NS] * (SMCWTD - SMCEQDEEP);
        if (-TOTWATER <= MAXWATDW) {
          SMCWTD = SMCWTD + TOTWATER / DZS[NS];
          WTD = MAX((SMCWTD * DZS[NS] - SMCEQDEEP * ZSOIL[NS] + SMCMAX * (ZSOIL[NS] - DZS[NS])) / (SMCMAX - SMCEQDEEP),
                    ZSOIL[NS] - DZS[NS]);
        } else {
          WTD = ZSOIL[NS] - DZS[NS];
          SMCWTD = SMCWTD + TOTWATER / DZS[NS];
          DZUP = (SMCEQDEEP - SMCWTD) * DZS[NS] / (SMCMAX - SMCEQDEEP);
          WTD = WTD - DZUP;
          SMCWTD = SMCEQDEEP;
        }
      }
    } else if (WTD >= ZSOIL[NS] - DZS[NS]) {
      SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS[NS]), 1.f / BEXP);
      SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
      MAXWATDW = DZS[NS] * (SMCWTD - SMCEQDEEP);
      if (-TOTWATER <= MAXWATDW) {
        SMCWTD = SMCWTD + TOTWATER / DZS[NS];
        WTD = MAX((SMCWTD * DZS[NS] - SMCEQDEEP * ZSOIL[NS] + SMCMAX * (ZSOIL[NS] - DZS[NS])) / (SMCMAX - SMCEQDEEP),
                  ZSOIL[NS] - DZS[NS]);
      } else {
        WTD = ZSOIL[NS] - DZS[NS];
        SMCWTD = SMCWTD + TOTWATER / DZS[NS];
        DZUP = (SMCEQDEEP - SMCWTD) * DZS[NS] / (SMCMAX - SMCEQDEEP);
        WTD = WTD - DZUP;
        SMCWTD = SMCEQDEEP;
      }
    } else {
      WGPMID = SMCMAX * POW(PSISAT / (PSISAT - (ZSOIL[NS] - WTD)), 1.f / BEXP);
      WGPMID = MAX(WGPMID, 1.E-4f);
      SYIELDDW = SMCMAX - WGPMID;
      WTDOLD = WTD;
      WTD = WTDOLD + TOTWATER / SYIELDDW;
      SMCWTD = (SMCWTD * (ZSOIL[NS] - WTDOLD) + WGPMID * (WTDOLD - WTD)) / (ZSOIL[NS] - WTD);
    }
    QSPRING = 0.f;
  }
  for (K = 1; K <= NS; ++K) SH2O[K] = SMC[K] - SICE[K];
}

// land columns: LATERALFLOW pass 2 (:259-292), river flux (:112-129), deep recharge (:147-161), UPDATEWTD,
// accumulation (:186-195)
__global__ void wt_column_kernel(const nmpf::WtParams w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= w.nland) return;
  const long long n = t;
  const int c = w.cell[n];
  const int il = c % w.ni, jl = c / w.ni;
  const int I = w.its + il, J = w.jts + jl;
  const long long np = w.np;
  float* S = w.state;
  float QLAT = 0.f;
  if (I >= max(w.its, w.ids + 1) && I <= min(w.ite, w.ide - 2) && J >= max(w.jts, w.jds + 1) &&
      J <= min(w.jte, w.jde - 2)) {
    const int P = w.ni + 2;
    const long long h = (long long)(il + 1) + (long long)(jl + 1) * P;
    const float* KC = w.kcell;
    const float* HD = w.head;
    const float k0 = KC[h], h0 = HD[h];
    const float SQRT2 = SQRT(2.f);
    float Q = 0.f;
    Q = Q + (KC[h - 1 + P] + k0) * (HD[h - 1 + P] - h0) / SQRT2;
    Q = Q + (KC[h - 1] + k0) * (HD[h - 1] - h0);
    Q = Q + (KC[h - 1 - P] + k0) * (HD[h - 1 - P] - h0) / SQRT2;
    Q = Q + (KC[h + P] + k0) * (HD[h + P] - h0);
    Q = Q + (KC[h - P] + k0) * (HD[h - P] - h0);
    Q = Q + (KC[h + 1 + P] + k0) * (HD[h + 1 + P] - h0) / SQRT2;
    Q = Q + (KC[h + 1] + k0) * (HD[h + 1] - h0);
    Q = Q + (KC[h + 1 - P] + k0) * (HD[h + 1 - P] - h0) / SQRT2;
    QLAT = 0.45508986056f * Q * w.deltat / w.area[c];
  }
  float WTD = S[(long long)NMP_SLOT(zwtxy) * np + n];
  float QRF;
  {
    float RCOND;
    const float rb = w.riverbed[c], eq = w.eqwtd[c];
    if (WTD > rb && eq > rb) RCOND = w.rivercond[c] * EXP(w.pexp[c] * (WTD - eq));
    else RCOND = w.rivercond[c];
    QRF = RCOND * (WTD - rb) * w.deltat / w.area[c];
    QRF = MAX(QRF, 0.f);
  }
  const noahmp_tables& T = *w.tables;
  const int st = __float_as_int(w.isltyp[c]);
  float BEXP = T.bb[st - 1], DKSAT = T.satdk[st - 1], SMCMAX = T.maxsmc[st - 1], PSISAT = -T.satpsi[st - 1];
  if (__float_as_int(w.ivgtyp[c]) == w.isurban) SMCMAX = 0.45f;
  float DZS[NSOIL + 1], ZSOIL[NSOIL + 1], SMC[NSOIL + 1], SH2O[NSOIL + 1], SMCEQ[NSOIL + 1];
  ZSOIL[0] = 0.f; DZS[0] = 0.f; SMC[0] = 0.f; SH2O[0] = 0.f; SMCEQ[0] = 0.f;
  for (int K = 1; K <= NSOIL; ++K) {
    DZS[K] = w.dzs[K - 1];
    ZSOIL[K] = (K == 1) ? -DZS[1] : -DZS[K] + ZSOIL[K - 1];
  }
  float SMCWTD = S[(long long)NMP_SLOT(smcwtdxy) * np + n];
  float DEEPRECH = S[(long long)NMP_SLOT(deeprechxy) * np + n];
  if (WTD < ZSOIL[NSOIL] - DZS[NSOIL]) {
    float DDZ = ZSOIL[NSOIL] - WTD;
    float SMCWTDMID = 0.5f * (SMCWTD + SMCMAX);
    float PSI = PSISAT * POW(SMCMAX / SMCWTD, BEXP);
    float WCNDDEEP = DKSAT * POW(SMCWTDMID / SMCMAX, 2.0f * BEXP + 3.0f);
    float WFLUXDEEP = -w.deltat * WCNDDEEP * ((PSISAT - PSI) / DDZ - 1.f);
    SMCWTD = SMCWTD + (DEEPRECH - WFLUXDEEP) / DDZ;
    float WPLUS = MAX((SMCWTD - SMCMAX), 0.0f) * DDZ;
    float WMINUS = MAX((1.E-4f - SMCWTD), 0.0f) * DDZ;
    SMCWTD = MAX(MIN(SMCWTD, SMCMAX), 1.E-4f);
    WFLUXDEEP = WFLUXDEEP + WPLUS - WMINUS;
    DEEPRECH = WFLUXDEEP;
  }
  float TOTWATER = QLAT - QRF + DEEPRECH;
  for (int K = 1; K <= NSOIL; ++K) {
    SMC[K] = S[(long long)(NMP_SLOT(smois) + K - 1) * np + n];
    SH2O[K] = S[(long long)(NMP_SLOT(sh2o) + K - 1) * np + n];
    SMCEQ[K] = S[(long long)(NMP_SLOT(smoiseq) + K - 1) * np + n];
  }
  float QSPRING;
  UPDATEWTD(DZS, ZSOIL, SMCEQ, SMCMAX, PSISAT, BEXP, TOTWATER, WTD, SMC, SH2O, SMCWTD, QSPRING);
  for (int K = 1; K <= NSOIL; ++K) {
    S[(long long)(NMP_SLOT(smois) + K - 1) * np + n] = SMC[K];
    S[(long long)(NMP_SLOT(sh2o) + K - 1) * np + n] = SH2O[K];
  }
  S[(long long)NMP_SLOT(zwtxy) * np + n] = WTD;
  S[(long long)NMP_SLOT(smcwtdxy) * np + n] = SMCWTD;
  w.qrf[c] = QRF;
  w.qspring[c] = QSPRING;
  w.qslat[c] = w.qslat[c] + QLAT * 1.E3f;
  w.qrfs[c] = w.qrfs[c] + QRF * 1.E3f;
  w.qsprings[c] = w.qsprings[c] + QSPRING * 1.E3f;
  S[(long long)NMP_SLOT(rechxy) * np + n] = S[(long long)NMP_SLOT(rechxy) * np + n] + DEEPRECH * 1.E3f;
  S[(long long)NMP_SLOT(deeprechxy) * np + n] = 0.f;
}

void launch_wtable(const nmpf::WtParams& w, cudaStream_t s, long long* launches, int phase) {
  const int T = 256;
  if (phase == 0) {
    const long long nc = (long long)w.ni * w.nj;
    wt_head_kernel<<<(unsigned)((nc + T - 1) / T), T, 0, s>>>(w);
    ++*launches;
  } else if (w.nland > 0) {
    wt_column_kernel<<<(w.nland + 127) / 128, 128, 0, s>>>(w);
    ++*launches;
  }
}

}  // namespace
