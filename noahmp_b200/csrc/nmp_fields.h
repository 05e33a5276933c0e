// nmp_fields.h — the field registry shared by the host library and the kernels.
//
// One entry per array of the `noahmplsm` dummy list (phys/module_sf_noahmpdrv.F90:11-44, decls :51-211),
// in the order of noahmp_lsm_args: X(member, layers, kind).  `layers` is the extent of the middle
// dimension of the Fortran (i,k,j) layout (1 for 2-D arrays).
//
// HBM layout:
//   * forcing / static inputs stay in the driver's 2-D (i,j) order ("grid order"), one dense plane each;
//   * INOUT and OUT arrays live in a column-compact structure of arrays: plane (SLOT_x + k) of `state`
//     holds layer k of field x for the NP active (non-water) columns, columns ordered land | glacier |
//     sea-ice, each class in grid order.  One thread owns one column, so every load/store of a plane is
//     a unit-stride, fully coalesced access.
#pragma once
#include "../../include/noahmp_b200.h"

#define NMP_K_INOUT 0
#define NMP_K_OUT 1

// INOUT + OUT arrays that live in the compact state (args order)
#define NMP_STATE_FIELDS(X)                                                                                      \
  X(tsk, 1, 0) X(hfx, 1, 0) X(qfx, 1, 0) X(lh, 1, 0) X(grdflx, 1, 0) X(smstav, 1, 0) X(smstot, 1, 0)            \
  X(sfcrunoff, 1, 0) X(udrunoff, 1, 0) X(albedo, 1, 0) X(snowc, 1, 0) X(smois, 4, 0) X(sh2o, 4, 0)              \
  X(tslb, 4, 0) X(snow, 1, 0) X(snowh, 1, 0) X(canwat, 1, 0) X(acsnom, 1, 0) X(acsnow, 1, 0) X(emiss, 1, 0)     \
  X(qsfc, 1, 0) X(isnowxy, 1, 0) X(tvxy, 1, 0) X(tgxy, 1, 0) X(canicexy, 1, 0) X(canliqxy, 1, 0)                \
  X(eahxy, 1, 0) X(tahxy, 1, 0) X(cmxy, 1, 0) X(chxy, 1, 0) X(fwetxy, 1, 0) X(sneqvoxy, 1, 0)                   \
  X(alboldxy, 1, 0) X(qsnowxy, 1, 0) X(wslakexy, 1, 0) X(zwtxy, 1, 0) X(waxy, 1, 0) X(wtxy, 1, 0)               \
  X(tsnoxy, 3, 0) X(zsnsoxy, 7, 0) X(snicexy, 3, 0) X(snliqxy, 3, 0) X(lfmassxy, 1, 0) X(rtmassxy, 1, 0)        \
  X(stmassxy, 1, 0) X(woodxy, 1, 0) X(stblcpxy, 1, 0) X(fastcpxy, 1, 0) X(xlaixy, 1, 0) X(xsaixy, 1, 0)         \
  X(taussxy, 1, 0) X(smoiseq, 4, 0) X(smcwtdxy, 1, 0) X(deeprechxy, 1, 0) X(rechxy, 1, 0)                       \
  X(t2mvxy, 1, 1) X(t2mbxy, 1, 1) X(q2mvxy, 1, 1) X(q2mbxy, 1, 1) X(tradxy, 1, 1) X(neexy, 1, 1)                \
  X(gppxy, 1, 1) X(nppxy, 1, 1) X(fvegxy, 1, 1) X(runsfxy, 1, 1) X(runsbxy, 1, 1) X(ecanxy, 1, 1)               \
  X(edirxy, 1, 1) X(etranxy, 1, 1) X(fsaxy, 1, 1) X(firaxy, 1, 1) X(aparxy, 1, 1) X(psnxy, 1, 1)                \
  X(savxy, 1, 1) X(sagxy, 1, 1) X(rssunxy, 1, 1) X(rsshaxy, 1, 1) X(bgapxy, 1, 1) X(wgapxy, 1, 1)               \
  X(tgvxy, 1, 1) X(tgbxy, 1, 1) X(chvxy, 1, 1) X(chbxy, 1, 1) X(shgxy, 1, 1) X(shcxy, 1, 1) X(shbxy, 1, 1)      \
  X(evgxy, 1, 1) X(evbxy, 1, 1) X(ghvxy, 1, 1) X(ghbxy, 1, 1) X(irgxy, 1, 1) X(ircxy, 1, 1) X(irbxy, 1, 1)      \
  X(trxy, 1, 1) X(evcxy, 1, 1) X(chleafxy, 1, 1) X(chucxy, 1, 1) X(chv2xy, 1, 1) X(chb2xy, 1, 1)

namespace nmpf {

// field ids
enum FieldId {
#define X(name, nl, kind) F_##name,
  NMP_STATE_FIELDS(X)
#undef X
      NFIELDS
};

// first plane (slot) of every field in the compact state
struct SlotTable {
  int slot[NFIELDS + 1];
  int layers[NFIELDS];
  int kind[NFIELDS];
  constexpr SlotTable() : slot(), layers(), kind() {
    int i = 0, s = 0;
#define X(name, nl, k) slot[i] = s; layers[i] = nl; kind[i] = k; s += nl; ++i;
    NMP_STATE_FIELDS(X)
#undef X
    slot[i] = s;
  }
};
constexpr SlotTable kSlots{};
constexpr int NPLANES = kSlots.slot[NFIELDS];
// internal planes after the argument-list fields: [NPLANES] = VEGE_FLUX pass count of the previous step (a key of
// the column re-binning), then the NSTATIC static inputs in compact order (so that re-binned columns still read
// them with unit stride)
// NMP_SPLIT=1 (experimental build, tools/build_split.sh): ENERGY and WATER as two kernels; the values that cross the cut
// travel through NHANDOFF extra planes (written by the first kernel, read by the second, never re-binned).
#ifndef NMP_SPLIT
#define NMP_SPLIT 0
#endif
constexpr int PLANE_PREV_ITERS = NPLANES;
constexpr int PLANE_STATIC0 = NPLANES + 1;
constexpr int NHANDOFF = NMP_SPLIT ? 20 : 0;
constexpr int PLANE_HANDOFF0 = NPLANES + 1 + 7;
constexpr int NPLANES_ALLOC = NPLANES + 1 + 7 + NHANDOFF;
template <int F>
struct SlotOf {
  static constexpr int value = kSlots.slot[F];
};
#define NMP_SLOT(name) (nmpf::SlotOf<nmpf::F_##name>::value)

// forcing / static planes kept in grid order (dense ni*nj each); see noahmp_b200_device_forcing()
enum ForcingId {
  FC_COSZIN = 0, FC_T, FC_QV, FC_U, FC_V, FC_SWDOWN, FC_GLW, FC_P1, FC_P2, FC_RAINBL, FC_VEGFRA, FC_DZ8W, NFORC
};
static_assert(NFORC == NOAHMP_NFORCING, "forcing plane count");
enum StaticId { ST_IVGTYP = 0, ST_ISLTYP, ST_VEGMAX, ST_TMN, ST_XLATIN, ST_XLAND, ST_XICE, NSTATIC };
static_assert(NPLANES_ALLOC == NPLANES + 1 + NSTATIC + NHANDOFF, "internal plane count");

// column classes (noahmpdrv.F90:426-441)
enum ColClass { CL_WATER = 0, CL_LAND = 1, CL_GLACIER = 2, CL_SEAICE = 3 };

// kernel parameter block of one step
struct StepParams {
  const float* forc[NFORC];
  const float* stat[NSTATIC];  // ivgtyp / isltyp planes are int32 bit patterns
  float* state;                // NPLANES planes of `np` words
  const int* cell;             // compact column -> grid cell (i-1) + (j-1)*ni  (0-based, tile-local)
  const noahmp_tables* tables;
  unsigned long long* err_key; // min over failing columns of (cell<<39 | code<<32 | value bits)
  int* err_count;
  int* vege_iters;             // optional per-column VEGE_FLUX pass count (diagnostic), may be NULL
  long long np;                // plane stride (number of active columns)
  unsigned np4;                // the same in bytes (fits: a tile has < 2^25 cells)
  int first, count;            // compact range this launch covers
  int ni;
  int itimestep, yearlen;
  float julian, dt, dx, xice_thres;
  int isice, isurban, iz0tlnd;
  float zsoil[NOAHMP_NSOIL];
  int opt[12];  // dveg crs btr run sfc frz inf rad alb snf tbot stc
};

// kernel parameter block of the opt_run=5 groundwater step (nmp_groundwater.cuh)
struct WtParams {
  const float *fdepth, *area, *topo, *rivercond, *riverbed, *eqwtd, *pexp;  // grid-order planes
  const float *isltyp, *ivgtyp;                                           // int32 bit patterns
  const float* wtd_grid;                                                  // WTD of every cell (staging plane of ZWTXY)
  float *qrf, *qspring, *qslat, *qrfs, *qsprings;                         // grid-order planes
  float *kcell, *head;                                                    // (ni+2) x (nj+2), one-cell halo ring
  float* state;
  const int* cell;
  const noahmp_tables* tables;
  long long np;
  int nland, ni, nj;
  int ids, ide, jds, jde, its, ite, jts, jte;
  int isurban;
  float deltat;
  float dzs[NOAHMP_NSOIL];
};

// kernel parameter block of the on-device forcing pipeline (nmp_forcing.cuh)
struct ForcingParams {
  const float* A[9];  // t q u v p lw sw pcp fpar of the earlier bracket
  const float* B[9];  // ... of the later bracket (only t..sw are read)
  const float *lat, *lon;
  float* out[NFORC];
  long long ncell;
  float fraction, sin_declin, cos_declin, hour_frac /* IHOUR + IMINUTE/60 + ISECOND/3600 */, dt, zlvl;
};

// ---- cold start (nmp_init.cuh): arrays of NOAHMP_INIT in dummy-argument order -----------------------------
// X(name, layers, io)  io: 1 = input only (uploaded), 3 = updated (uploaded and downloaded)
#define NMP_INIT_FIELDS(X)                                                                                          \
  X(snow, 1, 3) X(snowh, 1, 3) X(canwat, 1, 3) X(isltyp, 1, 1) X(ivgtyp, 1, 1) X(tslb, 4, 3) X(smois, 4, 3)          \
  X(sh2o, 4, 3) X(tsk, 1, 1) X(isnowxy, 1, 3) X(tvxy, 1, 3) X(tgxy, 1, 3) X(canicexy, 1, 3) X(xice, 1, 1)            \
  X(canliqxy, 1, 3) X(eahxy, 1, 3) X(tahxy, 1, 3) X(cmxy, 1, 3) X(chxy, 1, 3) X(fwetxy, 1, 3) X(sneqvoxy, 1, 3)      \
  X(alboldxy, 1, 3) X(qsnowxy, 1, 3) X(wslakexy, 1, 3) X(zwtxy, 1, 3) X(waxy, 1, 3) X(wtxy, 1, 3) X(tsnoxy, 3, 3)    \
  X(zsnsoxy, 7, 3) X(snicexy, 3, 3) X(snliqxy, 3, 3) X(lfmassxy, 1, 3) X(rtmassxy, 1, 3) X(stmassxy, 1, 3)           \
  X(woodxy, 1, 3) X(stblcpxy, 1, 3) X(fastcpxy, 1, 3) X(xsaixy, 1, 3) X(t2mvxy, 1, 3) X(t2mbxy, 1, 3)                \
  X(chstarxy, 1, 3)
// the optional groundwater block (iopt_run = 5)
#define NMP_INIT_GW_FIELDS(X)                                                                                       \
  X(smoiseq, 4, 3) X(smcwtdxy, 1, 3) X(rechxy, 1, 3) X(deeprechxy, 1, 3) X(areaxy, 1, 3) X(msftx, 1, 1)              \
  X(msfty, 1, 1) X(qrfsxy, 1, 3) X(qspringsxy, 1, 3) X(qslatxy, 1, 3) X(fdepthxy, 1, 1) X(ht, 1, 1)                  \
  X(riverbedxy, 1, 1) X(eqzwt, 1, 1) X(rivercondxy, 1, 1) X(pexpxy, 1, 1)
enum InitField {
#define X(nm, nl, io) IF_##nm,
  NMP_INIT_FIELDS(X) IF_GW0,
  IF_GW_PREV = IF_GW0 - 1,
  NMP_INIT_GW_FIELDS(X)
#undef X
  NINITF
};
struct InitParams {
  float* f[NINITF];     // grid-order device arrays, Fortran (i,k,j) layout; integer arrays as int32 bit patterns
  float *kcell, *head;  // LATERALFLOW pass-1 planes (iopt_run = 5)
  int* err;             // 0, or 1 = ISLTYP < 1, 2 = snow depth is not a number
  const noahmp_tables* tables;
  int ni, nj, itf_n, jtf_n;  // tile size; cells with il < itf_n and jl < jtf_n are initialised
  int ids, ide, jds, jde, its, ite, jts, jte;
  int isurban, isice, iswater, iopt_run, fndsnowh;
  float dx, dy, deltat;
  float dzs[NOAHMP_NSOIL];
};

// compact-column ranges one launch of the physics covers (a whole tile, or one row chunk of it)
struct StepRange {
  int land_first, land_count, glac_first, glac_count;
};

}  // namespace nmpf
