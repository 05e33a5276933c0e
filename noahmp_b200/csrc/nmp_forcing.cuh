// nmp_forcing.cuh — the driver-side forcing preparation on the device:
// hrldas_input_interpolate (driver/module_hrldas_netcdf_io.F90:1369-1404), the fills of
// driver/module_hrldas_noahmp_driver.F90:336-344 and the per-cell part of CALC_DECLIN (:845-861).
#pragma once
#include "nmp_common.cuh"
#include "nmp_fields.h"

namespace {

using namespace nmp;

__global__ void forcing_kernel(const nmpf::ForcingParams f) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= f.ncell) return;
  const float fr = f.fraction, om = 1.0f - fr;
  // (instructA%x * fraction) + (instructB%x * (1.0-fraction))
  const float t = (f.A[0][c] * fr) + (f.B[0][c] * om);
  const float q = (f.A[1][c] * fr) + (f.B[1][c] * om);
  const float u = (f.A[2][c] * fr) + (f.B[2][c] * om);
  const float v = (f.A[3][c] * fr) + (f.B[3][c] * om);
  const float p = (f.A[4][c] * fr) + (f.B[4][c] * om);
  const float lw = (f.A[5][c] * fr) + (f.B[5][c] * om);
  const float sw = (f.A[6][c] * fr) + (f.B[6][c] * om);
  const float pcp = f.A[7][c];
  const float fpar = f.A[8][c];
  // CALC_DECLIN, per-cell part
  const float DEGRAD = 3.14159265f / 180.f;
  float TLOCTIM = f.hour_frac + f.lon[c] / 15.0f;
  TLOCTIM = fmodf(TLOCTIM + 24.0f, 24.0f);
  const float HRANG = 15.f * (TLOCTIM - 12.f) * DEGRAD;
  const float lat = f.lat[c];
  const float COSZ = SIN(lat * DEGRAD) * f.sin_declin + COS(lat * DEGRAD) * f.cos_declin * COS(HRANG);
  f.out[nmpf::FC_COSZIN][c] = COSZ;
  f.out[nmpf::FC_T][c] = t;
  f.out[nmpf::FC_QV][c] = q;
  f.out[nmpf::FC_U][c] = u;
  f.out[nmpf::FC_V][c] = v;
  f.out[nmpf::FC_SWDOWN][c] = sw;
  f.out[nmpf::FC_GLW][c] = lw;
  f.out[nmpf::FC_P1][c] = p;
  f.out[nmpf::FC_P2][c] = p;  // P8W(:,2,:) = P8W(:,1,:)
  f.out[nmpf::FC_RAINBL][c] = pcp * f.dt;
  f.out[nmpf::FC_VEGFRA][c] = fpar * 100.0f;
  f.out[nmpf::FC_DZ8W][c] = 2.0f * f.zlvl;
}

void launch_forcing(const nmpf::ForcingParams& f, cudaStream_t s, long long* launches) {
  const int T = 256;
  forcing_kernel<<<(unsigned)((f.ncell + T - 1) / T), T, 0, s>>>(f);
  ++*launches;
}

}  // namespace
