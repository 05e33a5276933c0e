// nmp_kernels.cuh — the column-physics kernels: the ILOOP body of `noahmplsm`
// (phys/module_sf_noahmpdrv.F90:424-837) for land and glacier columns.
//
// Included by nmp_kernels_fast.cu (NMP_PARITY=0: libdevice math, FMA contraction) and by
// nmp_kernels_parity.cu (NMP_PARITY=1, compiled with -fmad=false: bit-identical to the CPU oracle's
// portable-math mode).  One thread = one column; columns of one class are contiguous in the compact
// state, so the land and the glacier kernels never diverge on column class.
#pragma once
#include <cstdlib>
#include "nmp_fields.h"
#include "nmp_glacier.cuh"
#include "nmp_io.cuh"

namespace {

using namespace nmp;
using nmpf::StepParams;

__device__ __forceinline__ void report_error(const StepParams& p, int cell, int code, float value) {
  atomicAdd(p.err_count, 1);
  unsigned long long key = ((unsigned long long)(unsigned)cell << 39) | ((unsigned long long)(code & 0x7f) << 32) |
                           (unsigned long long)__float_as_uint(value);
  atomicMin(p.err_key, key);
}

// gather of the pieces shared by land and glacier columns (noahmpdrv.F90:449-520)
__device__ __forceinline__ void load_common(const ColumnIO& io, const StepParams& p, Col& s) {
  s.COSZ = io.forc(nmpf::FC_COSZIN);
  s.LAT = io.stat(nmpf::ST_XLATIN);
  s.ZLVL = 0.5f * io.forc(nmpf::FC_DZ8W);
  s.TBOT = io.stat(nmpf::ST_TMN);
  s.SFCTMP = io.forc(nmpf::FC_T);
  const float qv = io.forc(nmpf::FC_QV);
  s.Q2 = qv / (1.0f + qv);
  s.UU = io.forc(nmpf::FC_U);
  s.VV = io.forc(nmpf::FC_V);
  s.SOLDN = io.forc(nmpf::FC_SWDOWN);
  s.LWDN = io.forc(nmpf::FC_GLW);
  const float p1 = io.forc(nmpf::FC_P1);
  s.SFCPRS = (io.forc(nmpf::FC_P2) + p1) * 0.5f;
  s.PSFC = p1;
  s.PRCP = io.forc(nmpf::FC_RAINBL) / p.dt;
  s.DT = p.dt;
  s.JULIAN = p.julian;
  s.YEARLEN = p.yearlen;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) s.ZSOIL(K) = p.zsoil[K - 1];

  s.ISNOW = io.ldi(NMP_SLOT(isnowxy));
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    s.SMC(K) = io.ld(NMP_SLOT(smois) + K - 1);
    s.SH2O(K) = io.ld(NMP_SLOT(sh2o) + K - 1);
    s.STC(K) = io.ld(NMP_SLOT(tslb) + K - 1);
  }
#pragma unroll
  for (int K = -2; K <= 0; ++K) {
    s.STC(K) = io.ld(NMP_SLOT(tsnoxy) + K + 2);
    s.SNICE(K) = io.ld(NMP_SLOT(snicexy) + K + 2);
    s.SNLIQ(K) = io.ld(NMP_SLOT(snliqxy) + K + 2);
  }
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) s.ZSNSO(K) = io.ld(NMP_SLOT(zsnsoxy) + K + 2);
  s.SNEQV = io.ld(NMP_SLOT(snow));
  s.SNOWH = io.ld(NMP_SLOT(snowh));
  s.QSFC = io.ld(NMP_SLOT(qsfc));
  s.TG = io.ld(NMP_SLOT(tgxy));
  s.CM = io.ld(NMP_SLOT(cmxy));
  s.CH = io.ld(NMP_SLOT(chxy));
  s.SNEQVO = io.ld(NMP_SLOT(sneqvoxy));
  s.ALBOLD = io.ld(NMP_SLOT(alboldxy));
  s.QSNOW = io.ld(NMP_SLOT(qsnowxy));
  s.TAUSS = io.ld(NMP_SLOT(taussxy));
#pragma unroll
  for (int K = -2; K <= 0; ++K) {
    s.FICEOLD(K) = 0.0f;
    if (K > s.ISNOW) s.FICEOLD(K) = s.SNICE(K) / (s.SNICE(K) + s.SNLIQ(K));
  }
}

// scatter of one column (noahmpdrv.F90:728-835); quantities are already set for the column class
struct ColumnOut {
  float QFX, LH, TV, CANICE, CANLIQ, EAH, TAH, FWET, WSLAKE, ZWT, WA, WT, LFMASS, RTMASS, STMASS, WOOD, STBLCP,
      FASTCP, PLAI, PSAI, T2MV, T2MB, Q2MV, Q2MB, NEE, GPP, NPP, FVEGMP, ECAN, ETRAN, ESOIL, APAR, PSN, SAV, RSSUN,
      RSSHA, BGAP, WGAP, TGV, TGB, CHV, CHB, IRC, IRG, SHC, SHG, EVG, GHV, IRB, SHB, EVB, GHB, TR, EVC, CHLEAF, CHUC,
      CHV2, CHB2, FSNO, RECH, DEEPRECH, SMCWTD;
};

__device__ __forceinline__ void store_column(const ColumnIO& io, const StepParams& p, const Col& s,
                                             const ColumnOut& o) {
  io.st(NMP_SLOT(tsk), s.TRAD);
  io.st(NMP_SLOT(hfx), s.FSH);
  io.st(NMP_SLOT(qfx), o.QFX);
  io.st(NMP_SLOT(lh), o.LH);
  io.st(NMP_SLOT(grdflx), s.SSOIL);
  io.st(NMP_SLOT(smstav), 0.0f);
  io.st(NMP_SLOT(smstot), 0.0f);
  io.st(NMP_SLOT(sfcrunoff), io.ld(NMP_SLOT(sfcrunoff)) + s.RUNSRF * p.dt);
  io.st(NMP_SLOT(udrunoff), io.ld(NMP_SLOT(udrunoff)) + s.RUNSUB * p.dt);
  if (s.ALBEDO > -999.f) io.st(NMP_SLOT(albedo), s.ALBEDO);
  io.st(NMP_SLOT(snowc), o.FSNO);
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) {
    io.st(NMP_SLOT(smois) + K - 1, s.SMC(K));
    io.st(NMP_SLOT(sh2o) + K - 1, s.SH2O(K));
    io.st(NMP_SLOT(tslb) + K - 1, s.STC(K));
  }
  io.st(NMP_SLOT(snow), s.SNEQV);
  io.st(NMP_SLOT(snowh), s.SNOWH);
  io.st(NMP_SLOT(canwat), o.CANLIQ + o.CANICE);
  io.st(NMP_SLOT(acsnow), io.ld(NMP_SLOT(acsnow)) + s.PRCP * s.FPICE);
  io.st(NMP_SLOT(acsnom), io.ld(NMP_SLOT(acsnom)) + s.QSNBOT * p.dt + s.PONDING + s.PONDING1 + s.PONDING2);
  io.st(NMP_SLOT(emiss), s.EMISSI);
  io.st(NMP_SLOT(qsfc), s.QSFC);
  io.sti(NMP_SLOT(isnowxy), s.ISNOW);
  io.st(NMP_SLOT(tvxy), o.TV);
  io.st(NMP_SLOT(tgxy), s.TG);
  io.st(NMP_SLOT(canliqxy), o.CANLIQ);
  io.st(NMP_SLOT(canicexy), o.CANICE);
  io.st(NMP_SLOT(eahxy), o.EAH);
  io.st(NMP_SLOT(tahxy), o.TAH);
  io.st(NMP_SLOT(cmxy), s.CM);
  io.st(NMP_SLOT(chxy), s.CH);
  io.st(NMP_SLOT(fwetxy), o.FWET);
  io.st(NMP_SLOT(sneqvoxy), s.SNEQVO);
  io.st(NMP_SLOT(alboldxy), s.ALBOLD);
  io.st(NMP_SLOT(qsnowxy), s.QSNOW);
  io.st(NMP_SLOT(wslakexy), o.WSLAKE);
  io.st(NMP_SLOT(zwtxy), o.ZWT);
  io.st(NMP_SLOT(waxy), o.WA);
  io.st(NMP_SLOT(wtxy), o.WT);
#pragma unroll
  for (int K = -2; K <= 0; ++K) {
    io.st(NMP_SLOT(tsnoxy) + K + 2, s.STC(K));
    io.st(NMP_SLOT(snicexy) + K + 2, s.SNICE(K));
    io.st(NMP_SLOT(snliqxy) + K + 2, s.SNLIQ(K));
  }
#pragma unroll
  for (int K = -2; K <= NSOIL; ++K) io.st(NMP_SLOT(zsnsoxy) + K + 2, s.ZSNSO(K));
  io.st(NMP_SLOT(lfmassxy), o.LFMASS);
  io.st(NMP_SLOT(rtmassxy), o.RTMASS);
  io.st(NMP_SLOT(stmassxy), o.STMASS);
  io.st(NMP_SLOT(woodxy), o.WOOD);
  io.st(NMP_SLOT(stblcpxy), o.STBLCP);
  io.st(NMP_SLOT(fastcpxy), o.FASTCP);
  io.st(NMP_SLOT(xlaixy), o.PLAI);
  io.st(NMP_SLOT(xsaixy), o.PSAI);
  io.st(NMP_SLOT(taussxy), s.TAUSS);
  io.st(NMP_SLOT(t2mvxy), o.T2MV);
  io.st(NMP_SLOT(t2mbxy), o.T2MB);
  io.st(NMP_SLOT(q2mvxy), o.Q2MV / (1.0f - o.Q2MV));
  io.st(NMP_SLOT(q2mbxy), o.Q2MB / (1.0f - o.Q2MB));
  io.st(NMP_SLOT(tradxy), s.TRAD);
  io.st(NMP_SLOT(neexy), o.NEE);
  io.st(NMP_SLOT(gppxy), o.GPP);
  io.st(NMP_SLOT(nppxy), o.NPP);
  io.st(NMP_SLOT(fvegxy), o.FVEGMP);
  io.st(NMP_SLOT(runsfxy), s.RUNSRF);
  io.st(NMP_SLOT(runsbxy), s.RUNSUB);
  io.st(NMP_SLOT(ecanxy), o.ECAN);
  io.st(NMP_SLOT(edirxy), o.ESOIL);
  io.st(NMP_SLOT(etranxy), o.ETRAN);
  io.st(NMP_SLOT(fsaxy), s.FSA);
  io.st(NMP_SLOT(firaxy), s.FIRA);
  io.st(NMP_SLOT(aparxy), o.APAR);
  io.st(NMP_SLOT(psnxy), o.PSN);
  io.st(NMP_SLOT(savxy), o.SAV);
  io.st(NMP_SLOT(sagxy), s.SAG);
  io.st(NMP_SLOT(rssunxy), o.RSSUN);
  io.st(NMP_SLOT(rsshaxy), o.RSSHA);
  io.st(NMP_SLOT(bgapxy), o.BGAP);
  io.st(NMP_SLOT(wgapxy), o.WGAP);
  io.st(NMP_SLOT(tgvxy), o.TGV);
  io.st(NMP_SLOT(tgbxy), o.TGB);
  io.st(NMP_SLOT(chvxy), o.CHV);
  io.st(NMP_SLOT(chbxy), o.CHB);
  io.st(NMP_SLOT(ircxy), o.IRC);
  io.st(NMP_SLOT(irgxy), o.IRG);
  io.st(NMP_SLOT(shcxy), o.SHC);
  io.st(NMP_SLOT(shgxy), o.SHG);
  io.st(NMP_SLOT(evgxy), o.EVG);
  io.st(NMP_SLOT(ghvxy), o.GHV);
  io.st(NMP_SLOT(irbxy), o.IRB);
  io.st(NMP_SLOT(shbxy), o.SHB);
  io.st(NMP_SLOT(evbxy), o.EVB);
  io.st(NMP_SLOT(ghbxy), o.GHB);
  io.st(NMP_SLOT(trxy), o.TR);
  io.st(NMP_SLOT(evcxy), o.EVC);
  io.st(NMP_SLOT(chleafxy), o.CHLEAF);
  io.st(NMP_SLOT(chucxy), o.CHUC);
  io.st(NMP_SLOT(chv2xy), o.CHV2);
  io.st(NMP_SLOT(chb2xy), o.CHB2);
  io.st(NMP_SLOT(rechxy), io.ld(NMP_SLOT(rechxy)) + o.RECH * 1.E3f);
  io.st(NMP_SLOT(deeprechxy), io.ld(NMP_SLOT(deeprechxy)) + o.DEEPRECH);
  io.st(NMP_SLOT(smcwtdxy), o.SMCWTD);
}

// NMP_SMEM_TABLES: the parameter tables (noahmp_tables, 12.7 KB) staged per block in shared memory and indexed by
// VEGTYP / SOILTYP from there (north_star item 2) instead of through L1 from global memory.
#ifndef NMP_SMEM_TABLES
#define NMP_SMEM_TABLES 0
#endif
#if NMP_SMEM_TABLES
__device__ __forceinline__ const noahmp_tables* stage_tables(const StepParams& p) {
  __shared__ noahmp_tables sT;
  static_assert(sizeof(noahmp_tables) % 4 == 0, "tables are copied word by word");
  const unsigned* src = reinterpret_cast<const unsigned*>(p.tables);
  unsigned* dst = reinterpret_cast<unsigned*>(&sT);
  for (unsigned k = threadIdx.x; k < sizeof(noahmp_tables) / 4; k += blockDim.x) dst[k] = __ldg(src + k);
  __syncthreads();
  return &sT;
}
#endif

__device__ __forceinline__ void init_ctx(Ctx& c, const StepParams& p) {
  c.T = p.tables;
  c.o.dveg = p.opt[0]; c.o.crs = p.opt[1]; c.o.btr = p.opt[2]; c.o.run = p.opt[3]; c.o.sfc = p.opt[4];
  c.o.frz = p.opt[5]; c.o.inf = p.opt[6]; c.o.rad = p.opt[7]; c.o.alb = p.opt[8]; c.o.snf = p.opt[9];
  c.o.tbot = p.opt[10]; c.o.stc = p.opt[11];
  c.err = 0; c.errv = 0.f;
}

// ---- land columns: REDPRM + NOAHMP_SFLX (noahmpdrv.F90:449-547, :681-714) ---------------------------
#ifndef NMP_WATER_MINBLOCKS
#define NMP_WATER_MINBLOCKS 3
#endif
// PART 0: the fused column program.  NMP_SPLIT build: PART 1 = ENERGY half (same launch bounds), PART 2 = WATER half
// (3 blocks of 256 threads per SM: <= 85 registers).
// NMP_TILE_PREFETCH = d > 0: a block starts the HBM -> L2 transfer of the INOUT / static planes of the columns that
// block (blockIdx + d) will load (d = the number of resident blocks: that block starts about when this one ends).
#ifndef NMP_TILE_PREFETCH
#define NMP_TILE_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_tile(const StepParams& p, int block) {
  constexpr int NIN = NMP_SLOT(t2mvxy);            // planes [0, NIN) are the INOUT arrays
  constexpr int NPF = NIN + 1 + nmpf::NSTATIC;     // + PREV_ITERS + compact static planes
  constexpr int LINES = NMP_BLOCK / 32;            // 128-byte lines per plane and block
  const long long c0 = (long long)p.first - (p.first & 31) + (long long)block * NMP_BLOCK;
  if (c0 >= (long long)p.first + p.count) return;
  for (int i = threadIdx.x; i < NPF * LINES; i += NMP_BLOCK) {
    int pl = i / LINES, ln = i % LINES;
    if (pl >= NIN) pl += nmpf::PLANE_PREV_ITERS - NIN;
    long long col = c0 + ln * 32;
    if (col < p.first) col = p.first;
    if (col >= (long long)p.first + p.count) continue;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.state + (long long)pl * p.np + col));
  }
  if (threadIdx.x < LINES) {
    long long col = c0 + threadIdx.x * 32;
    if (col >= p.first && col < (long long)p.first + p.count) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.cell + col));
  }
}

template <class O, int PART = 0>
__global__ void __launch_bounds__(NMP_BLOCK, PART == 2 ? NMP_WATER_MINBLOCKS : NMP_MINBLOCKS)
land_kernel(const __grid_constant__ StepParams p) {
  // Warps are aligned to 32-column (128-byte) boundaries of the planes whatever the first column of the launch is
  // (row chunks of the RESIDENT pipeline start anywhere): the launch is shifted down by first % 32 lanes.
  int t = blockIdx.x * blockDim.x + threadIdx.x - (p.first & 31);
  // threads outside the range redo a column of it and skip the stores, so that every thread of the block reaches
  // the phase barriers inside NOAHMP_SFLX
  bool live = t >= 0 && t < p.count;
  if (!live) t = t < 0 ? 0 : p.count - 1;
  ColumnIO io(p, (long long)p.first + t, live);
#if NMP_TILE_PREFETCH
  prefetch_tile(p, blockIdx.x + NMP_TILE_PREFETCH);
#endif
  Ctx c;
  init_ctx(c, p);
#if NMP_SMEM_TABLES
  c.T = stage_tables(p);
#endif
  Col s;
  load_common(io, p, s);
  s.ICE = 0;
  int VEGTYP = io.stati(nmpf::ST_IVGTYP);
  int SOILTYP = io.stati(nmpf::ST_ISLTYP);
  const int ivg = VEGTYP;
  s.SHDFAC = io.forc(nmpf::FC_VEGFRA) / 100.f;
  s.SHDMAX = io.stat(nmpf::ST_VEGMAX) / 100.f;
  s.TV = io.ld(NMP_SLOT(tvxy));
  s.CANLIQ = io.ld(NMP_SLOT(canliqxy));
  s.CANICE = io.ld(NMP_SLOT(canicexy));
  if (PART != 2) {
    s.EAH = io.ld(NMP_SLOT(eahxy));
    s.TAH = io.ld(NMP_SLOT(tahxy));
  }
  s.FWET = io.ld(NMP_SLOT(fwetxy));
  s.WA = io.ld(NMP_SLOT(waxy));  // for BEG_WB; reloaded with the rest of the water-table state before WATER
  s.LAI = io.ld(NMP_SLOT(xlaixy));
  s.SAI = io.ld(NMP_SLOT(xsaixy));
  // state that NOAHMP_SFLX loads late (water table, carbon pools): start the HBM -> L2 transfer now
  io.prefetch(NMP_SLOT(zwtxy)); io.prefetch(NMP_SLOT(wtxy)); io.prefetch(NMP_SLOT(smcwtdxy));
  if (p.opt[0] == 2 || p.opt[0] == 5) {
    io.prefetch(NMP_SLOT(lfmassxy)); io.prefetch(NMP_SLOT(rtmassxy)); io.prefetch(NMP_SLOT(stmassxy));
    io.prefetch(NMP_SLOT(woodxy)); io.prefetch(NMP_SLOT(stblcpxy)); io.prefetch(NMP_SLOT(fastcpxy));
  }
  // WSLAKE (IST = 1 always) and, unless dveg is 2 or 5, the carbon pools pass through NOAHMP_SFLX unchanged:
  // they stay where they are in HBM.  The remaining state is loaded inside NOAHMP_SFLX where first needed.
  s.ZWT = 0.f; s.WT = 0.f; s.SMCWTD = 0.f; s.WSLAKE = 0.f;
#pragma unroll
  for (int K = 1; K <= NSOIL; ++K) s.SMCEQ(K) = 0.f;
  s.RECH = 0.f;
  s.DEEPRECH = 0.f;
  const float CO2 = 395.e-06f, O2 = 0.209f;
  s.CO2AIR = CO2 * s.SFCPRS;
  s.O2AIR = O2 * s.SFCPRS;
  s.FOLN = 1.0f;

  if (SOILTYP == 14 && io.stat(nmpf::ST_XICE) == 0.f) SOILTYP = 7;
  if (ivg == p.isurban || ivg == 31 || ivg == 32 || ivg == 33) VEGTYP = p.isurban;
  if (VEGTYP == 25 || VEGTYP == 26 || VEGTYP == 27) { s.SHDFAC = 0.0f; s.LAI = 0.0f; }
  s.VEGTYP = VEGTYP;
  s.URBAN = (VEGTYP == p.isurban);

  const bool bad_index = REDPRM(c, VEGTYP, SOILTYP, 1, s.URBAN) != 0;
  if (bad_index) {
    // REDPRM range violation: the column is reported and left untouched; it still walks through the physics with
    // a valid parameter row so that the block's barriers stay matched
    Ctx c2;
    init_ctx(c2, p);
    REDPRM(c2, 19, 1, 1, false);
    c.P = c2.P;
    s.VEGTYP = 19;
    s.URBAN = false;
    io.on = false;
  }
  // every OUT member is assigned by NOAHMP_SFLX before use except on the dveg error path
  s.PONDING = 0.f; s.PONDING1 = 0.f; s.PONDING2 = 0.f; s.QSNBOT = 0.f; s.FPICE = 0.f;
  NOAHMP_SFLX<O, PART>(c, s, io);
  if (PART != 2) {
    if (io.on && p.vege_iters) p.vege_iters[io.cell] = s.VEGE_ITERS;
    io.sti(nmpf::PLANE_PREV_ITERS, s.VEGE_ITERS);  // key of the column re-binning (nmp_lib.cu: rebin)
  }
  // (the WATER half of the split build does not report again what the ENERGY half reported for a rejected column)
  if (live && c.err && !(PART == 2 && bad_index)) report_error(p, io.cell, c.err, c.errv);
}

// ---- glacier columns: NOAHMP_GLACIER + sentinel fills (noahmpdrv.F90:552-628) ------------------------
// Same launch shape as the land kernel (NMP_BLOCK threads, warps aligned to 128-byte plane segments, the block
// walking NOAHMP_GLACIER in barrier-separated phases); padding threads and rejected columns redo a valid column
// with the stores switched off so that every thread reaches the barriers.
template <class O>
__global__ void __launch_bounds__(NMP_BLOCK, NMP_MINBLOCKS) glacier_kernel(const __grid_constant__ StepParams p) {
  int t = blockIdx.x * blockDim.x + threadIdx.x - (p.first & 31);
  bool live = t >= 0 && t < p.count;
  if (!live) t = t < 0 ? 0 : p.count - 1;
  ColumnIO io(p, (long long)p.first + t, live);
  Ctx c;
  init_ctx(c, p);
#if NMP_SMEM_TABLES
  c.T = stage_tables(p);
#endif
  Col s;
  load_common(io, p, s);
  s.ICE = -1;
  s.TBOT = MIN(s.TBOT, 263.15f);
  s.PONDING = 0.f; s.PONDING1 = 0.f; s.PONDING2 = 0.f; s.QSNBOT = 0.f;
  // REDPRM runs for glacier columns too (noahmpdrv.F90:547); its range checks are the only effect
  {
    int VEGTYP = io.stati(nmpf::ST_IVGTYP), SOILTYP = io.stati(nmpf::ST_ISLTYP);
    if (SOILTYP == 14 && io.stat(nmpf::ST_XICE) == 0.f) SOILTYP = 7;
    const int ivg = VEGTYP;
    if (ivg == p.isurban || ivg == 31 || ivg == 32 || ivg == 33) VEGTYP = p.isurban;
    if (REDPRM(c, VEGTYP, SOILTYP, 1, VEGTYP == p.isurban)) {
      if (live) report_error(p, io.cell, c.err, c.errv);
      c.err = 0;
      io.on = false;
      live = false;
    }
  }
  NOAHMP_GLACIER<O>(c, s);

  const float U1 = -1.E36f, U2 = 0.0f;  // undefined_value / undefined_value2 (noahmpdrv.F90:368-369)
  ColumnOut o;
  o.QFX = s.EDIR;
  o.LH = s.FGEV;
  o.FSNO = 1.0f;
  o.TV = U1; o.TGB = s.TG; o.CANICE = U2; o.CANLIQ = U2; o.EAH = U1; o.TAH = U1; o.FWET = U2; o.WSLAKE = U2;
  o.ZWT = U1; o.WA = U1; o.WT = U1; o.LFMASS = U2; o.RTMASS = U2; o.STMASS = U2; o.WOOD = U2; o.STBLCP = U1;
  o.FASTCP = U1; o.PLAI = U2; o.PSAI = U2; o.T2MV = U1; o.Q2MV = U1; o.NEE = U2; o.GPP = U2; o.NPP = U2;
  o.FVEGMP = 0.0f; o.ECAN = U2; o.ETRAN = U2; o.APAR = U2; o.PSN = U2; o.SAV = U2; o.RSSUN = U1; o.RSSHA = U1;
  o.BGAP = U1; o.WGAP = U1; o.TGV = U1; o.CHV = U1; o.CHB = s.CH; o.IRC = U1; o.IRG = U1; o.SHC = U1; o.SHG = U1;
  o.EVG = U1; o.GHV = U1; o.IRB = s.FIRA; o.SHB = s.FSH; o.EVB = s.FGEV; o.GHB = s.SSOIL; o.TR = U2; o.EVC = U2;
  o.CHLEAF = U1; o.CHUC = U1; o.CHV2 = U1; o.CHB2 = s.CHB2;
  o.T2MB = s.T2MB; o.Q2MB = s.Q2B; o.ESOIL = s.EDIR;
  o.RECH = 0.f; o.DEEPRECH = 0.f;
  o.SMCWTD = io.ld(NMP_SLOT(smcwtdxy));
  store_column(io, p, s, o);
  if (io.on && p.vege_iters) p.vege_iters[io.cell] = 0;
  if (live && c.err) report_error(p, io.cell, c.err, c.errv);
}

template <class O>
void launch_pair(const StepParams& base, const nmpf::StepRange& r, cudaStream_t stream, long long* launches) {
  const int nland = r.land_count, nglac = r.glac_count;
  if (nland > 0) {
    StepParams p = base;
    p.first = r.land_first;
    p.count = nland;
#if NMP_SPLIT
    land_kernel<O, 1><<<(nland + (r.land_first & 31) + NMP_BLOCK - 1) / NMP_BLOCK, NMP_BLOCK, 0, stream>>>(p);
    land_kernel<O, 2><<<(nland + (r.land_first & 31) + NMP_BLOCK - 1) / NMP_BLOCK, NMP_BLOCK, 0, stream>>>(p);
    *launches += 2;
#else
    land_kernel<O><<<(nland + (r.land_first & 31) + NMP_BLOCK - 1) / NMP_BLOCK, NMP_BLOCK, 0, stream>>>(p);
    ++*launches;
#endif
  }
  if (nglac > 0) {
    StepParams p = base;
    p.first = r.glac_first;
    p.count = nglac;
    glacier_kernel<O><<<(nglac + (r.glac_first & 31) + NMP_BLOCK - 1) / NMP_BLOCK, NMP_BLOCK, 0, stream>>>(p);
    ++*launches;
  }
}

// compile-time option sets: dveg crs btr run sfc frz inf rad alb snf tbot stc
using OptDefault = OptSet<4, 1, 1, 1, 1, 1, 1, 3, 2, 1, 2, 1>;  // BASELINE configs C1/C2/C4
using OptDynVeg = OptSet<2, 1, 1, 1, 1, 1, 1, 3, 2, 1, 2, 1>;   // C3: dveg=2 dynamic vegetation
using OptDynVegMMF = OptSet<2, 1, 1, 5, 1, 1, 1, 3, 2, 1, 2, 1>;  // C5: C3 with the MMF groundwater scheme (opt_run=5)

template <class O>
bool matches(const int* opt) {
  const int want[12] = {O::dveg, O::crs, O::btr, O::run, O::sfc, O::frz, O::inf, O::rad, O::alb, O::snf, O::tbot,
                        O::stc};
  for (int k = 0; k < 12; ++k)
    if (want[k] != opt[k]) return false;
  return true;
}

// Picks the specialised instantiation when the namelist options match one, else the generic kernel that
// reads the options at run time.  Returns the name of the variant (for logs / tests).
const char* launch_step(const StepParams& base, const nmpf::StepRange& r, cudaStream_t stream, long long* launches) {
#ifndef NMP_NO_SPECIALISE
  if (getenv("NOAHMP_B200_FORCE_RUNTIME")) {
    launch_pair<OptRuntime>(base, r, stream, launches);
    return "runtime";
  }
  if (matches<OptDefault>(base.opt)) { launch_pair<OptDefault>(base, r, stream, launches); return "default"; }
  if (matches<OptDynVeg>(base.opt)) { launch_pair<OptDynVeg>(base, r, stream, launches); return "dynveg"; }
  if (matches<OptDynVegMMF>(base.opt)) { launch_pair<OptDynVegMMF>(base, r, stream, launches); return "dynveg_mmf"; }
#endif
  launch_pair<OptRuntime>(base, r, stream, launches);
  return "runtime";
}

}  // namespace
