// nmo_forcing.cpp — ORACLE (test infrastructure): the driver-side forcing preparation, restated from
// driver/module_hrldas_netcdf_io.F90:1369-1404 (hrldas_input_interpolate), driver/module_hrldas_noahmp_driver.F90:
// 336-354 (VEGFRA*100, level-2 copies, RAINBL, DZ8W, the CALC_DECLIN loop) and :813-863 (CALC_DECLIN).
#include "nmo.h"

extern "C" float nmo_forcing(const noahmp_forcing_fields* A, const noahmp_forcing_fields* B, const float* lat2d,
                             const float* lon2d, long ncell, float fraction, int iday, int ihour, int iminute,
                             int isecond, float model_timestep, float zlvl, float** out /* [NOAHMP_NFORCING] */) {
  using namespace nmo;
  const float DEGRAD = 3.14159265f / 180.f, DPD = 360.f / 365.f;
  const float JULIAN = (float)iday + (float)ihour / 24.f;
  float DECLIN = 0.f;
  const float OBECL = 23.5f * DEGRAD;
  const float SINOB = SIN(OBECL);
  float SXLONG = 0.f;
  if (JULIAN >= 80.f) SXLONG = DPD * (JULIAN - 80.f) * DEGRAD;
  if (JULIAN < 80.f) SXLONG = DPD * (JULIAN + 285.f) * DEGRAD;
  const float ARG = SINOB * SIN(SXLONG);
  DECLIN = ASIN(ARG);
  for (long c = 0; c < ncell; ++c) {
    const float om = 1.0f - fraction;
    out[1][c] = (A->t[c] * fraction) + (B->t[c] * om);
    out[2][c] = (A->q[c] * fraction) + (B->q[c] * om);
    out[3][c] = (A->u[c] * fraction) + (B->u[c] * om);
    out[4][c] = (A->v[c] * fraction) + (B->v[c] * om);
    const float p = (A->p[c] * fraction) + (B->p[c] * om);
    out[7][c] = p;
    out[8][c] = p;
    out[6][c] = (A->lw[c] * fraction) + (B->lw[c] * om);
    out[5][c] = (A->sw[c] * fraction) + (B->sw[c] * om);
    out[9][c] = A->pcp[c] * model_timestep;
    out[10][c] = A->fpar[c] * 100.0f;
    out[11][c] = 2.0f * zlvl;
    float TLOCTIM = (float)ihour + (float)iminute / 60.0f + (float)isecond / 3600.0f + lon2d[c] / 15.0f;
    TLOCTIM = std::fmod(TLOCTIM + 24.0f, 24.0f);
    const float HRANG = 15.f * (TLOCTIM - 12.f) * DEGRAD;
    out[0][c] = SIN(lat2d[c] * DEGRAD) * SIN(DECLIN) + COS(lat2d[c] * DEGRAD) * COS(DECLIN) * COS(HRANG);
  }
  return JULIAN;
}
