// nmp_init.cuh — cold start on the device: NOAHMP_INIT, SNOW_INIT, GROUNDWATER_INIT and
// EQSMOISTURE (phys/module_sf_noahmpdrv.F90:847-1522), one thread per grid cell of the tile, arrays in the caller's
// Fortran layout.  Built in the parity translation unit only: the cold start runs once, so it always uses the
// portable transcendentals and rounds exactly as the CPU oracle does.
#pragma once
#include "nmp_common.cuh"
#include "nmp_fields.h"
#include "nmp_groundwater.cuh"

namespace {

using namespace nmp;

// element (il, K, jl) of a layered array whose first layer has Fortran index K0 and which has NK layers
__device__ __forceinline__ long long lay(const nmpf::InitParams& P, int il, int K, int K0, int NK, int jl) {
  return (long long)il + (long long)(K - K0) * P.ni + (long long)jl * P.ni * NK;
}

// noahmpdrv.F90:996-1120 and SNOW_INIT :1182-1283; with iopt_run = 5 also AREAXY and LATERALFLOW pass 1
// (groundwater.F90:236-252) on the water-table depth the caller supplied
__global__ void init_cell_kernel(const nmpf::InitParams P) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (long long)P.ni * P.nj) return;
  const int il = (int)(c % P.ni), jl = (int)(c / P.ni);
  auto F = [&](int f) -> float& { return P.f[f][c]; };
  if (P.iopt_run == 5) {
    const int I = P.its + il, J = P.jts + jl;
    float kc = 0.f, hd = 0.f;
    if (I >= max(P.its - 1, P.ids) && I <= min(P.ite + 1, P.ide - 1) && J >= max(P.jts - 1, P.jds) &&
        J <= min(P.jte + 1, P.jde - 1)) {
      const float fd = P.f[nmpf::IF_fdepthxy][c], wtd = P.f[nmpf::IF_zwtxy][c];
      const int st = __float_as_int(P.f[nmpf::IF_isltyp][c]);
      if (fd > 0.f && st >= 1) {
        const float KLAT = P.tables->satdk[st - 1] * kKLATFACTOR[st - 1];
        if (wtd < -1.5f) kc = fd * KLAT * EXP((wtd + 1.5f) / fd);
        else kc = KLAT * (wtd + 1.5f + fd);
      }
      hd = P.f[nmpf::IF_ht][c] + wtd;
    }
    P.kcell[c] = kc;
    P.head[c] = hd;
  }
  if (il >= P.itf_n || jl >= P.jtf_n) return;
  if (!P.fndsnowh) F(nmpf::IF_snowh) = F(nmpf::IF_snow) * 0.005f;
  const int ISLTYP = __float_as_int(F(nmpf::IF_isltyp)), IVGTYP = __float_as_int(F(nmpf::IF_ivgtyp));
  if (ISLTYP < 1) { atomicMax(P.err, 1); return; }
  const noahmp_tables& T = *P.tables;
  const float HLICE = 3.335E5f, GRAV_I = 9.81f, T0 = 273.15f;
  if (IVGTYP == P.isice && F(nmpf::IF_xice) <= 0.0f) {
    for (int NS = 1; NS <= NSOIL; ++NS) {
      const long long q = lay(P, il, NS, 1, NSOIL, jl);
      P.f[nmpf::IF_smois][q] = 1.0f;
      P.f[nmpf::IF_sh2o][q] = 0.0f;
      P.f[nmpf::IF_tslb][q] = MIN(P.f[nmpf::IF_tslb][q], 263.15f);
    }
    F(nmpf::IF_snow) = MAX(F(nmpf::IF_snow), 10.0f);
    F(nmpf::IF_snowh) = F(nmpf::IF_snow) * 0.01f;
  } else {
    const float BX = T.bb[ISLTYP - 1], SMCMAX = T.maxsmc[ISLTYP - 1], PSISAT = T.satpsi[ISLTYP - 1];
    const bool soil_ok = BX > 0.0f && SMCMAX > 0.0f && PSISAT > 0.0f;
    for (int NS = 1; NS <= NSOIL; ++NS) {
      const long long q = lay(P, il, NS, 1, NSOIL, jl);
      float smois = P.f[nmpf::IF_smois][q];
      if (smois > SMCMAX) { smois = SMCMAX; P.f[nmpf::IF_smois][q] = smois; }
      const float tslb = P.f[nmpf::IF_tslb][q];
      float sh2o = smois;
      if (soil_ok && tslb < 273.149f) {
        float FK = POW((HLICE / (GRAV_I * (-PSISAT))) * ((tslb - T0) / tslb), -1.f / BX) * SMCMAX;
        FK = MAX(FK, 0.02f);
        sh2o = MIN(FK, smois);
      }
      P.f[nmpf::IF_sh2o][q] = sh2o;
    }
  }
  const float snow = F(nmpf::IF_snow), tsk = F(nmpf::IF_tsk);
  const float tstart = (snow > 0.0f && tsk > 273.15f) ? 273.15f : tsk;
  F(nmpf::IF_tvxy) = tstart;
  F(nmpf::IF_tgxy) = tstart;
  F(nmpf::IF_canwat) = 0.0f;
  F(nmpf::IF_canliqxy) = 0.0f;
  F(nmpf::IF_canicexy) = 0.f;
  F(nmpf::IF_eahxy) = 2000.f;
  F(nmpf::IF_tahxy) = tstart;
  F(nmpf::IF_t2mvxy) = tstart;
  F(nmpf::IF_t2mbxy) = tstart;
  F(nmpf::IF_chstarxy) = 0.1f;
  F(nmpf::IF_cmxy) = 0.0f;
  F(nmpf::IF_chxy) = 0.0f;
  F(nmpf::IF_fwetxy) = 0.0f;
  F(nmpf::IF_sneqvoxy) = 0.0f;
  F(nmpf::IF_alboldxy) = 0.65f;
  F(nmpf::IF_qsnowxy) = 0.0f;
  F(nmpf::IF_wslakexy) = 0.0f;
  if (P.iopt_run != 5) {
    const float wa = 4900.f;
    F(nmpf::IF_waxy) = wa;
    F(nmpf::IF_wtxy) = wa;
    F(nmpf::IF_zwtxy) = (25.f + 2.0f) - wa / 1000.f / 0.2f;
  } else {
    F(nmpf::IF_waxy) = 0.f;
    F(nmpf::IF_wtxy) = 0.f;
    F(nmpf::IF_areaxy) = (P.dx * P.dy) / (F(nmpf::IF_msftx) * F(nmpf::IF_msfty));
  }
  F(nmpf::IF_lfmassxy) = 50.f;
  F(nmpf::IF_stmassxy) = 50.0f;
  F(nmpf::IF_rtmassxy) = 500.0f;
  F(nmpf::IF_woodxy) = 500.0f;
  F(nmpf::IF_stblcpxy) = 1000.0f;
  F(nmpf::IF_fastcpxy) = 1000.0f;
  F(nmpf::IF_xsaixy) = 0.1f;
  // ---- SNOW_INIT (SWE = SNOW, SNODEP = SNOWH, TGXY as just set) ----
  float ZSOIL[NSOIL + 1];
  ZSOIL[0] = 0.f;
  ZSOIL[1] = -P.dzs[0];
  for (int NS = 2; NS <= NSOIL; ++NS) ZSOIL[NS] = ZSOIL[NS - 1] - P.dzs[NS - 1];
  const float SNODEP = F(nmpf::IF_snowh), SWE = snow;
  float DZSNO[NSNOW] = {0.f, 0.f, 0.f};  // index IZ + 2
  int ISNOW;
  if (SNODEP < 0.025f) {
    ISNOW = 0;
  } else if (SNODEP >= 0.025f && SNODEP <= 0.05f) {
    ISNOW = -1;
    DZSNO[2] = SNODEP;
  } else if (SNODEP > 0.05f && SNODEP <= 0.10f) {
    ISNOW = -2;
    DZSNO[1] = SNODEP / 2.f;
    DZSNO[2] = SNODEP / 2.f;
  } else if (SNODEP > 0.10f && SNODEP <= 0.25f) {
    ISNOW = -2;
    DZSNO[1] = 0.05f;
    DZSNO[2] = SNODEP - DZSNO[1];
  } else if (SNODEP > 0.25f && SNODEP <= 0.45f) {
    ISNOW = -3;
    DZSNO[0] = 0.05f;
    DZSNO[1] = 0.5f * (SNODEP - DZSNO[0]);
    DZSNO[2] = 0.5f * (SNODEP - DZSNO[0]);
  } else if (SNODEP > 0.45f) {
    ISNOW = -3;
    DZSNO[0] = 0.05f;
    DZSNO[1] = 0.20f;
    DZSNO[2] = SNODEP - DZSNO[1] - DZSNO[0];
  } else {
    atomicMax(P.err, 2);
    return;
  }
  P.f[nmpf::IF_isnowxy][c] = __int_as_float(ISNOW);
  for (int IZ = -NSNOW + 1; IZ <= 0; ++IZ) {
    const long long q = lay(P, il, IZ, -NSNOW + 1, NSNOW, jl);
    const bool in = IZ >= ISNOW + 1;
    P.f[nmpf::IF_tsnoxy][q] = in ? tstart : 0.f;
    P.f[nmpf::IF_snliqxy][q] = 0.f;
    P.f[nmpf::IF_snicexy][q] = in ? 1.00f * DZSNO[IZ + 2] * (SWE / SNODEP) : 0.f;
  }
  float z = 0.f;
  for (int IZ = ISNOW + 1; IZ <= NSOIL; ++IZ) {
    const float dz = IZ <= 0 ? -DZSNO[IZ + 2] : (IZ == 1 ? ZSOIL[1] : ZSOIL[IZ] - ZSOIL[IZ - 1]);
    z = (IZ == ISNOW + 1) ? dz : z + dz;
    P.f[nmpf::IF_zsnsoxy][lay(P, il, IZ, -NSNOW + 1, NSNOW + NSOIL, jl)] = z;
  }
}

// EQSMOISTURE (noahmpdrv.F90:1473-1522)
__device__ void EQSMOISTURE(const float* ZSOIL /*0:NSOIL*/, float SMCMAX, float DWSAT, float DKSAT, float BEXP,
                            float* SMCEQ /*1:NSOIL*/) {
  for (int K = 1; K <= NSOIL; ++K) {
    float DDZ;
    if (K == 1) DDZ = -ZSOIL[K + 1] * 0.5f;
    else if (K < NSOIL) DDZ = (ZSOIL[K - 1] - ZSOIL[K + 1]) * 0.5f;
    else DDZ = ZSOIL[K - 1] - ZSOIL[K];
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
    SMCEQ[K] = MIN(MAX(SMC, 1.E-4f), SMCMAX * 0.99f);
  }
}

// GROUNDWATER_INIT (noahmpdrv.F90:1286-1470): LATERALFLOW pass 2, river flux, equilibrium and deep soil moisture
__global__ void init_groundwater_kernel(const nmpf::InitParams P) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= (long long)P.ni * P.nj) return;
  const int il = (int)(c % P.ni), jl = (int)(c / P.ni);
  if (il >= P.itf_n || jl >= P.jtf_n) return;
  const int I = P.its + il, J = P.jts + jl;
  const int ISLTYP = __float_as_int(P.f[nmpf::IF_isltyp][c]), IVGTYP = __float_as_int(P.f[nmpf::IF_ivgtyp][c]);
  // a missing-field fill (e.g. -9999) was flagged by init_cell_kernel (err = 1, "lsminit: out of range value of
  // ISLTYP"); it must not index the soil tables here
  if (ISLTYP < 1 || ISLTYP > NOAHMP_NSLTYPE) return;
  const bool land = IVGTYP != P.iswater && IVGTYP != P.isice;
  const float area = P.f[nmpf::IF_areaxy][c];
  float WTD = P.f[nmpf::IF_zwtxy][c];
  float QLAT = 0.f;
  if (land && I >= max(P.its, P.ids + 1) && I <= min(P.ite, P.ide - 2) && J >= max(P.jts, P.jds + 1) &&
      J <= min(P.jte, P.jde - 2)) {
    const int W = P.ni;
    const float* KC = P.kcell;
    const float* HD = P.head;
    const float k0 = KC[c], h0 = HD[c];
    const float SQRT2 = SQRT(2.f);
    float Q = 0.f;
    Q = Q + (KC[c - 1 + W] + k0) * (HD[c - 1 + W] - h0) / SQRT2;
    Q = Q + (KC[c - 1] + k0) * (HD[c - 1] - h0);
    Q = Q + (KC[c - 1 - W] + k0) * (HD[c - 1 - W] - h0) / SQRT2;
    Q = Q + (KC[c + W] + k0) * (HD[c + W] - h0);
    Q = Q + (KC[c - W] + k0) * (HD[c - W] - h0);
    Q = Q + (KC[c + 1 + W] + k0) * (HD[c + 1 + W] - h0) / SQRT2;
    Q = Q + (KC[c + 1] + k0) * (HD[c + 1] - h0);
    Q = Q + (KC[c + 1 - W] + k0) * (HD[c + 1 - W] - h0) / SQRT2;
    QLAT = 0.45508986056f * Q * P.deltat / area;
  }
  float QRF = 0.f;
  if (land) {
    const float rb = P.f[nmpf::IF_riverbedxy][c], eq = P.f[nmpf::IF_eqzwt][c], rc = P.f[nmpf::IF_rivercondxy][c];
    float RCOND;
    if (WTD > rb && eq > rb) RCOND = rc * EXP(P.f[nmpf::IF_pexpxy][c] * (WTD - eq));
    else RCOND = rc;
    QRF = RCOND * (WTD - rb) * P.deltat / area;
    QRF = MAX(QRF, 0.f);
  }
  const noahmp_tables& T = *P.tables;
  const float BX = T.bb[ISLTYP - 1];
  float SMCMAX = T.maxsmc[ISLTYP - 1];
  if (IVGTYP == P.isurban) SMCMAX = 0.45f;
  const float DWSAT = T.satdw[ISLTYP - 1], DKSAT = T.satdk[ISLTYP - 1], PSISAT = -T.satpsi[ISLTYP - 1];
  float ZSOIL[NSOIL + 1], DZS[NSOIL + 1], SMCEQ[NSOIL + 1];
  ZSOIL[0] = 0.f; DZS[0] = 0.f; SMCEQ[0] = 0.f;
  for (int K = 1; K <= NSOIL; ++K) {
    DZS[K] = P.dzs[K - 1];
    ZSOIL[K] = K == 1 ? -DZS[1] : ZSOIL[K - 1] - DZS[K];
  }
  float SMCWTD;
  if (BX > 0.0f && SMCMAX > 0.0f && -PSISAT > 0.0f) {
    EQSMOISTURE(ZSOIL, SMCMAX, DWSAT, DKSAT, BX, SMCEQ);
    for (int K = 1; K <= NSOIL; ++K) P.f[nmpf::IF_smoiseq][lay(P, il, K, 1, NSOIL, jl)] = SMCEQ[K];
    if (WTD < ZSOIL[NSOIL] - DZS[NSOIL]) {
      const float EXPON = 2.f * BX + 3.f;
      const float DDZ = ZSOIL[NSOIL] - WTD;
      const float CC = PSISAT / DDZ;
      const float FLUX = (QLAT - QRF) / P.deltat;
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
      SMCWTD = MAX(SMC, 1.E-4f);
    } else if (WTD < ZSOIL[NSOIL]) {
      float SMCEQDEEP = SMCMAX * POW(PSISAT / (PSISAT - DZS[NSOIL]), 1.f / BX);
      SMCEQDEEP = MAX(SMCEQDEEP, 1.E-4f);
      SMCWTD = SMCMAX * (WTD - (ZSOIL[NSOIL] - DZS[NSOIL])) + SMCEQDEEP * (ZSOIL[NSOIL] - WTD);
    } else {
      SMCWTD = SMCMAX;
      for (int K = NSOIL; K >= 2; --K) {
        const long long q = lay(P, il, K, 1, NSOIL, jl);
        const float smois = P.f[nmpf::IF_smois][q];
        if (WTD >= ZSOIL[K - 1]) {
          const float FRLIQ = P.f[nmpf::IF_sh2o][q] / smois;
          P.f[nmpf::IF_smois][q] = SMCMAX;
          P.f[nmpf::IF_sh2o][q] = SMCMAX * FRLIQ;
        } else {
          if (smois < SMCEQ[K]) WTD = ZSOIL[K];
          else WTD = (smois * DZS[K] - SMCEQ[K] * ZSOIL[K - 1] + SMCMAX * ZSOIL[K]) / (SMCMAX - SMCEQ[K]);
          break;
        }
      }
    }
  } else {
    for (int K = 1; K <= NSOIL; ++K) P.f[nmpf::IF_smoiseq][lay(P, il, K, 1, NSOIL, jl)] = SMCMAX;
    SMCWTD = SMCMAX;
    WTD = 0.f;
  }
  P.f[nmpf::IF_smcwtdxy][c] = SMCWTD;
  P.f[nmpf::IF_zwtxy][c] = WTD;
  P.f[nmpf::IF_deeprechxy][c] = 0.f;
  P.f[nmpf::IF_rechxy][c] = 0.f;
  P.f[nmpf::IF_qslatxy][c] = 0.f;
  P.f[nmpf::IF_qrfsxy][c] = 0.f;
  P.f[nmpf::IF_qspringsxy][c] = 0.f;
}

void launch_init(const nmpf::InitParams& P, cudaStream_t s, long long* launches) {
  const int T = 256;
  const long long nc = (long long)P.ni * P.nj;
  init_cell_kernel<<<(unsigned)((nc + T - 1) / T), T, 0, s>>>(P);
  ++*launches;
  if (P.iopt_run == 5) {
    init_groundwater_kernel<<<(unsigned)((nc + T - 1) / T), T, 0, s>>>(P);
    ++*launches;
  }
}

}  // namespace
