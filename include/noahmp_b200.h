/* noahmp_b200.h — C-ABI of the B200-native Noah-MP column-physics step.
 *
 * This is the drop-in boundary for the reference's
 *     SUBROUTINE noahmplsm(...)            phys/module_sf_noahmpdrv.F90:11-44 (decls :51-211)
 * as called once per model step by
 *     land_driver_exe                      driver/module_hrldas_noahmp_driver.F90:386-415
 * plus the table loaders it depends on
 *     read_mp_veg_parameters(DATASET)      phys/module_sf_noahmplsm.F90:274-404   (MPTABLE.TBL)
 *     SOIL_VEG_GEN_PARM(MMINLU,MMINSL)     phys/module_sf_noahmpdrv.F90:1528-1821 (VEGPARM/SOILPARM/GENPARM.TBL)
 *
 * Plain C: pointers, ints, floats. No torch / CUDA types cross this boundary.  Host arrays use
 * the Fortran memory layout of the reference (2-D A(i,j): i fastest; 3-D A(i,k,j): i fastest,
 * layer k in the middle), REAL = float, INTEGER = int32.  INTEGRATION.md shows the
 * ISO_C_BINDING shim that forwards the unchanged Fortran `noahmplsm` signature here.
 */
#ifndef NOAHMP_B200_H
#define NOAHMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NOAHMP_MVT 27      /* NOAHMP_VEG_PARAMETERS::MVT, noahmplsm.F90:206 */
#define NOAHMP_MBAND 2     /* noahmplsm.F90:207 */
#define NOAHMP_NLUS 50     /* noahmp_globals::NLUS, noahmplsm.F90:42 */
#define NOAHMP_NSLTYPE 30  /* noahmplsm.F90:83 */
#define NOAHMP_NSLOPE 30   /* noahmplsm.F90:100 */
#define NOAHMP_NSOIL 4     /* the path is built for NSOIL=4 (glacier PHASECHANGE hard-codes it, glacier.F90:1804-1908) */
#define NOAHMP_NSNOW 3     /* noahmpdrv.F90:367 */

/* Parameter tables, exactly the module data the reference physics reads.
 * 2-D Fortran arrays X(MVT,n) are stored x[n][MVT] (same memory order). Index 0 = Fortran index 1. */
typedef struct noahmp_tables {
  /* ---- MPTABLE.TBL (noahmplsm.F90:209-259, namelist :291-301) ---- */
  int32_t nveg;
  int32_t isurban_mp; /* PRIVATE in the reference; kept for completeness, unused by physics */
  int32_t iswater, isbarren, issnow, eblforest;
  float ch2op[NOAHMP_MVT], dleaf[NOAHMP_MVT], z0mvt[NOAHMP_MVT], hvt[NOAHMP_MVT], hvb[NOAHMP_MVT];
  float den[NOAHMP_MVT], rc[NOAHMP_MVT];
  float rhol[NOAHMP_MBAND][NOAHMP_MVT], rhos[NOAHMP_MBAND][NOAHMP_MVT];
  float taul[NOAHMP_MBAND][NOAHMP_MVT], taus[NOAHMP_MBAND][NOAHMP_MVT];
  float xl[NOAHMP_MVT], cwpvt[NOAHMP_MVT], c3psn[NOAHMP_MVT], kc25[NOAHMP_MVT], akc[NOAHMP_MVT];
  float ko25[NOAHMP_MVT], ako[NOAHMP_MVT], avcmx[NOAHMP_MVT], aqe[NOAHMP_MVT], ltovrc[NOAHMP_MVT];
  float dilefc[NOAHMP_MVT], dilefw[NOAHMP_MVT], rmf25[NOAHMP_MVT], sla[NOAHMP_MVT], fragr[NOAHMP_MVT];
  float tmin[NOAHMP_MVT], vcmx25[NOAHMP_MVT], tdlef[NOAHMP_MVT], bp[NOAHMP_MVT], mp[NOAHMP_MVT];
  float qe25[NOAHMP_MVT], rms25[NOAHMP_MVT], rmr25[NOAHMP_MVT], arm[NOAHMP_MVT], folnmx[NOAHMP_MVT];
  float wdpool[NOAHMP_MVT], wrrat[NOAHMP_MVT], mrp[NOAHMP_MVT];
  float saim[12][NOAHMP_MVT], laim[12][NOAHMP_MVT];
  float slarea[NOAHMP_MVT], eps[5][NOAHMP_MVT];
  /* ---- VEGPARM.TBL (noahmpdrv.F90:1569-1646); only NROTBL,RSTBL,RGLTBL,HSTBL,TOPT,RSMAX reach physics ---- */
  int32_t lucats;
  int32_t nrotbl[NOAHMP_NLUS];
  float shdtbl[NOAHMP_NLUS], rstbl[NOAHMP_NLUS], rgltbl[NOAHMP_NLUS], hstbl[NOAHMP_NLUS];
  float snuptbl[NOAHMP_NLUS], maxalb[NOAHMP_NLUS], laimintbl[NOAHMP_NLUS], laimaxtbl[NOAHMP_NLUS];
  float emissmintbl[NOAHMP_NLUS], emissmaxtbl[NOAHMP_NLUS], albedomintbl[NOAHMP_NLUS], albedomaxtbl[NOAHMP_NLUS];
  float z0mintbl[NOAHMP_NLUS], z0maxtbl[NOAHMP_NLUS], ztopvtbl[NOAHMP_NLUS], zbotvtbl[NOAHMP_NLUS];
  float topt_data, cmcmax_data, cfactr_data, rsmax_data;
  int32_t bare, natural;
  /* ---- SOILPARM.TBL (noahmpdrv.F90:1681-1726) ---- */
  int32_t slcats;
  float bb[NOAHMP_NSLTYPE], drysmc[NOAHMP_NSLTYPE], f11[NOAHMP_NSLTYPE], maxsmc[NOAHMP_NSLTYPE];
  float refsmc[NOAHMP_NSLTYPE], satpsi[NOAHMP_NSLTYPE], satdk[NOAHMP_NSLTYPE], satdw[NOAHMP_NSLTYPE];
  float wltsmc[NOAHMP_NSLTYPE], qtz[NOAHMP_NSLTYPE];
  /* ---- GENPARM.TBL (noahmpdrv.F90:1755-1800) ---- */
  int32_t slpcats;
  float slope_data[NOAHMP_NSLOPE];
  float sbeta_data, fxexp_data, csoil_data, salp_data, refdk_data, refkdt_data, frzk_data, zbot_data;
  float czil_data, smlow_data, smhigh_data, lvcoef_data;
} noahmp_tables;

/* Argument list of `noahmplsm` (noahmpdrv.F90:11-44) as a struct of the same names, same order.
 * Arrays are the caller's (host) arrays in Fortran layout with memory bounds ims:ime, kms:kme,
 * jms:jme (soil arrays 1:nsoil, snow arrays -2:0, ZSNSOXY -2:nsoil in the middle dimension). */
typedef struct noahmp_lsm_args {
  /* IN: time/space */
  int32_t itimestep, yr;
  float julian;
  const float* coszin;
  const float* xlatin;
  const float* dz8w; /* (ims:ime,kms:kme,jms:jme) */
  float dt;
  const float* dzs; /* (1:nsoil) */
  int32_t nsoil;
  float dx;
  const int32_t* ivgtyp;
  const int32_t* isltyp;
  const float* vegfra;
  const float* vegmax;
  const float* tmn;
  const float* xland;
  const float* xice;
  float xice_thres;
  int32_t isice, isurban;
  /* IN: options */
  int32_t idveg, iopt_crs, iopt_btr, iopt_run, iopt_sfc, iopt_frz, iopt_inf, iopt_rad, iopt_alb;
  int32_t iopt_snf, iopt_tbot, iopt_stc, iz0tlnd;
  /* IN: forcing */
  const float *t3d, *qv3d, *u_phy, *v_phy; /* 3-D (ims:ime,kms:kme,jms:jme) */
  const float *swdown, *glw;
  const float* p8w3d; /* 3-D */
  const float* rainbl;
  /* INOUT: generic LSM (21) */
  float *tsk, *hfx, *qfx, *lh, *grdflx, *smstav, *smstot, *sfcrunoff, *udrunoff, *albedo, *snowc;
  float *smois, *sh2o, *tslb; /* (ims:ime,1:nsoil,jms:jme) */
  float *snow, *snowh, *canwat, *acsnom, *acsnow, *emiss, *qsfc;
  /* INOUT: Noah-MP (34) */
  int32_t* isnowxy;
  float *tvxy, *tgxy, *canicexy, *canliqxy, *eahxy, *tahxy, *cmxy, *chxy, *fwetxy, *sneqvoxy, *alboldxy;
  float *qsnowxy, *wslakexy, *zwtxy, *waxy, *wtxy;
  float* tsnoxy;  /* (ims:ime,-2:0,jms:jme) */
  float* zsnsoxy; /* (ims:ime,-2:nsoil,jms:jme) */
  float *snicexy, *snliqxy; /* (ims:ime,-2:0,jms:jme) */
  float *lfmassxy, *rtmassxy, *stmassxy, *woodxy, *stblcpxy, *fastcpxy, *xlaixy, *xsaixy, *taussxy;
  float* smoiseq; /* (ims:ime,1:nsoil,jms:jme) */
  float *smcwtdxy, *deeprechxy, *rechxy;
  /* OUT: Noah-MP (44) */
  float *t2mvxy, *t2mbxy, *q2mvxy, *q2mbxy, *tradxy, *neexy, *gppxy, *nppxy, *fvegxy, *runsfxy;
  float *runsbxy, *ecanxy, *edirxy, *etranxy, *fsaxy, *firaxy, *aparxy, *psnxy, *savxy, *sagxy;
  float *rssunxy, *rsshaxy, *bgapxy, *wgapxy, *tgvxy, *tgbxy, *chvxy, *chbxy, *shgxy, *shcxy;
  float *shbxy, *evgxy, *evbxy, *ghvxy, *ghbxy, *irgxy, *ircxy, *irbxy, *trxy, *evcxy;
  float *chleafxy, *chucxy, *chv2xy, *chb2xy;
  /* index bounds */
  int32_t ids, ide, jds, jde, kds, kde;
  int32_t ims, ime, jms, jme, kms, kme;
  int32_t its, ite, jts, jte, kts, kte;
} noahmp_lsm_args;

/* Error record: the reference aborts the process via wrf_error_fatal (util/module_wrf_utilities.F:12-24)
 * from ERROR (noahmplsm.F90:1164-1226), ERROR_GLACIER (glacier.F90:2932-2970), ENERGY (:1787) etc.
 * Here the step returns a status and the first failing column; the Fortran shim turns a nonzero
 * status into the reference's wrf_error_fatal message. */
enum {
  NOAHMP_OK = 0,
  NOAHMP_ERR_ERRSW = 1,    /* "Stop in Noah-MP" (noahmplsm.F90:1185)                      */
  NOAHMP_ERR_ERRENG = 2,   /* "Energy budget problem in NOAHMP LSM" (noahmplsm.F90:1196) */
  NOAHMP_ERR_ERRWAT = 3,   /* "Water budget problem in NOAHMP LSM" (noahmplsm.F90:1221)  */
  NOAHMP_ERR_FIRE = 4,     /* "STOP in Noah-MP" emitted longwave <0 (noahmplsm.F90:1792)  */
  NOAHMP_ERR_HCAN = 5,     /* "CRITICAL PROBLEM: HCAN <= ZPD" (noahmplsm.F90:3289)        */
  NOAHMP_ERR_ZLVL = 6,     /* "STOP in Noah-MP" ZLVL <= ZPD (noahmplsm.F90:4124)          */
  NOAHMP_ERR_REDPRM = 7,   /* REDPRM range checks (noahmplsm.F90:9266-9277, :9340)        */
  NOAHMP_ERR_OPTION = 8,   /* unsupported option value (opt_sfc 3/4: non-functional offline) */
  NOAHMP_ERR_ISLTYP = 9,   /* NOAHMP_INIT: ISLTYP < 1 (noahmpdrv.F90:1008-1021)            */
  NOAHMP_ERR_CUDA = 100,   /* CUDA runtime failure; see noahmp_b200_last_error()          */
  NOAHMP_ERR_ARG = 101
};

typedef struct noahmp_status {
  int32_t code;      /* one of the enum above */
  int32_t i, j;      /* Fortran (i,j) of the first failing column, 0 if none */
  int32_t count;     /* number of failing columns in this step */
  float value;       /* offending residual (ERRSW / ERRENG / ERRWAT ...) */
} noahmp_status;

typedef struct noahmp_b200_ctx noahmp_b200_ctx; /* opaque */

/* ---- table loading: replaces read_mp_veg_parameters + SOIL_VEG_GEN_PARM ------------------- */
/* Reads MPTABLE.TBL, VEGPARM.TBL, SOILPARM.TBL, GENPARM.TBL from `dir` (the reference reads the
 * CWD; pass "." for identical behaviour). dataset = "USGS" | "MODIFIED_IGBP_MODIS_NOAH",
 * soil = "STAS". Returns 0 on success. Host-only, no GPU needed. */
int noahmp_b200_read_tables(const char* dir, const char* dataset, const char* soil, noahmp_tables* out);
/* Message of the last table-reading failure (the text the reference would pass to wrf_error_fatal). */
const char* noahmp_b200_tables_error(void);
/* sizeof(noahmp_tables) / sizeof(noahmp_lsm_args) as compiled into the library (ABI checks of bindings). */
unsigned long long noahmp_b200_sizeof_tables(void);
unsigned long long noahmp_b200_sizeof_args(void);

/* ---- context ------------------------------------------------------------------------------ */
/* Creates a device context on CUDA device `device` for a tile of ni x nj columns. Fails (returns
 * NULL) when no CUDA device is usable: there is no CPU fallback.
 * Host memory: caller arrays of 4 MiB and more that the step-path calls copy are page-locked (cudaHostRegister) on
 * first use and remembered by address until noahmp_b200_destroy, as a Fortran driver's arrays live for the whole
 * run.  A caller that frees or reallocates such arrays while the context exists must create the context with the
 * environment variable NOAHMP_B200_PIN=0 (pageable copies).  noahmp_b200_init never page-locks. */
noahmp_b200_ctx* noahmp_b200_create(int device, const noahmp_tables* tables, int ni, int nj);
void noahmp_b200_destroy(noahmp_b200_ctx* ctx);
const char* noahmp_b200_last_error(void);

/* mode flags for noahmp_b200_set_mode */
#define NOAHMP_SYNC_FULL 0     /* every call uploads INOUT state and downloads INOUT+OUT: strict drop-in */
#define NOAHMP_SYNC_RESIDENT 1 /* state stays in HBM; forcing is uploaded each call; host INOUT/OUT arrays
                                  are refreshed only by noahmp_b200_sync_host (output/restart cadence) */
int noahmp_b200_set_mode(noahmp_b200_ctx* ctx, int sync_mode);

/* Arithmetic of the physics kernels.  FAST: CUDA libdevice single-precision functions with FMA contraction
 * (the production build).  PARITY: every transcendental through csrc/nmp_math.h and no FMA contraction, which
 * makes the results bit-identical with the CPU oracle's portable-math mode (used by the parity tests; also
 * selectable with the environment variable NOAHMP_B200_MATH=parity). */
#define NOAHMP_MATH_FAST 0
#define NOAHMP_MATH_PARITY 1
int noahmp_b200_set_math(noahmp_b200_ctx* ctx, int math_mode);
/* Name of the kernel instantiation the last step used: "default" / "dynveg" (opt_* compiled in as template
 * constants) or "runtime" (options read at run time). */
const char* noahmp_b200_kernel_variant(const noahmp_b200_ctx* ctx);

/* ---- the step: replaces `CALL noahmplsm(...)` ------------------------------------------------ */
int noahmp_b200_noahmplsm(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args, noahmp_status* status);

/* Refresh the caller's INOUT/OUT host arrays from HBM (RESIDENT mode). */
int noahmp_b200_sync_host(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args);

/* RESIDENT mode: fields (comma separated noahmp_lsm_args member names, e.g. "tsk,hfx,lh,grdflx"; "" = none) that
 * every noahmp_b200_noahmplsm call refreshes in the caller's arrays; the rest waits for noahmp_b200_sync_host.
 * The HRLDAS driver reads TSLB and LAI every step (module_hrldas_noahmp_driver.F90:567-572) and the output list
 * only every output_timestep (:440-565). */
int noahmp_b200_set_fetch(noahmp_b200_ctx* ctx, const char* fields);
/* RESIDENT mode: INOUT fields whose HOST content every noahmp_b200_noahmplsm call takes again before the step (same
 * syntax as set_fetch).  land_driver_exe overwrites LAI (passed as XLAIXY) from the forcing file before every call
 * (driver/module_hrldas_noahmp_driver.F90:335, :403; driver/module_hrldas_netcdf_io.F90:1365, :1402): a resident
 * drop-in lists "xlaixy" here. */
int noahmp_b200_set_push(noahmp_b200_ctx* ctx, const char* fields);
/* RESIDENT mode: forcing planes of the following calls that need no upload (a hint is a promise of the caller).
 *   DZ8W_CONSTANT      DZ8W = 2*zlvl is set once by the driver (module_hrldas_noahmp_driver.F90:344)
 *   VEGFRA_UNCHANGED   VEGFRA has not changed since the previous call (it changes when a forcing file is read,
 *                      module_hrldas_netcdf_io.F90:1238-1246, :1401); clear the bit for the call after a change
 *   P8W_LEVELS_EQUAL   levels kts and kts+1 of P8W3D hold the same values (the driver copies level 1 into level 2,
 *                      module_hrldas_noahmp_driver.F90:338): level 2 is read from the level-1 plane
 * A plane is sent at least once; withdrawing a hint makes the next call send it again. */
#define NOAHMP_HINT_DZ8W_CONSTANT 1u
#define NOAHMP_HINT_VEGFRA_UNCHANGED 2u
#define NOAHMP_HINT_P8W_LEVELS_EQUAL 4u
int noahmp_b200_set_forcing_hints(noahmp_b200_ctx* ctx, unsigned hints);
/* Release the page-lock the library holds on one caller array (before the caller frees it). */
int noahmp_b200_unpin(noahmp_b200_ctx* ctx, const void* host_array);
/* RESIDENT mode runs as a pipeline over `nchunks` row chunks (forcing upload | physics | result download overlap);
 * 0 = automatic: one chunk below 2^20 cells, 5 below 2^21, else 9; the first and the last chunk are half as tall as the
 * others from 4 chunks on.  After the first re-binning the chunking is fixed: other counts than the
 * binned one and 1 fall back to it. */
int noahmp_b200_set_chunks(noahmp_b200_ctx* ctx, int nchunks);
/* Divergence control (north_star item 4): in RESIDENT mode the land columns are physically re-ordered every
 * `interval` steps (default 20, 0 = never), inside eighths of their row chunk (bands of the grid-order planes that fit
 * the L2), into bins of equal snow-layer count and canopy tile computed / not computed in the previous step, so the
 * threads of a block take the same branches.  Results do
 * not depend on the order; noahmp_b200_column_map returns the current one.  The environment variable
 * NOAHMP_B200_REBIN sets the initial interval (a tuning aid: profiles/r02_rebin_interval.log).  When a re-binning is
 * due the library first counts how many columns left their bin (asynchronously: the answer is used by a later step)
 * and permutes only above NOAHMP_B200_REBIN_MIN_CHANGED (default 0.001 of the land columns; 0 = always).
 * noahmp_b200_rebin_count counts the permutations done. */
int noahmp_b200_set_rebin(noahmp_b200_ctx* ctx, int interval);
int noahmp_b200_rebin_count(const noahmp_b200_ctx* ctx);
/* Refresh ONE caller array (named like the noahmp_lsm_args member, e.g. "tsk") from HBM in RESIDENT mode. */
int noahmp_b200_fetch(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args, const char* field);

/* ---- device-resident stepping used by bench.py / drivers that keep forcing on the GPU -------- */
/* Upload static fields + initial state from host arrays (Fortran layout) once. */
int noahmp_b200_upload(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args);
/* Pointers to device forcing staging buffers (column-compact order is internal; these are in the
 * Fortran 2-D (i,j) layout, ni*nj floats each) so a driver can fill them on the device. Order:
 * 0 COSZIN 1 T3D(k=1) 2 QV3D(k=1) 3 U_PHY 4 V_PHY 5 SWDOWN 6 GLW 7 P8W3D(k=1) 8 P8W3D(k=2) 9 RAINBL
 * 10 VEGFRA 11 DZ8W(k=1) */
#define NOAHMP_NFORCING 12
int noahmp_b200_device_forcing(noahmp_b200_ctx* ctx, float** dev_ptrs /* [NOAHMP_NFORCING] */);
/* Use caller-owned device planes (same order/layout) as the forcing of the following steps; NULL = back to
 * the context's own staging planes. */
int noahmp_b200_bind_forcing(noahmp_b200_ctx* ctx, float* const* dev_ptrs /* [NOAHMP_NFORCING] or NULL */);
/* One step entirely on the device with the forcing currently in the staging buffers.
 * scalars as in noahmplsm; `stream` is a cudaStream_t passed as void* (NULL = the context's stream). */
int noahmp_b200_step_device(noahmp_b200_ctx* ctx, int itimestep, int yr, float julian, float dt,
                            void* stream);
/* Status since the previous get_status (or upload): the first failing column and the number of failing column-steps
 * of all noahmp_b200_step_device calls in between (synchronises the device, then resets the latch). */
int noahmp_b200_get_status(noahmp_b200_ctx* ctx, noahmp_status* status);
/* Number of kernels launched by this context since creation (for bench accounting). */
long long noahmp_b200_launch_count(const noahmp_b200_ctx* ctx);
/* Column census: [0] land columns, [1] glacier columns, [2] sea-ice columns, [3] water columns. */
int noahmp_b200_census(const noahmp_b200_ctx* ctx, int64_t counts[4]);

/* Compact column -> 0-based tile-local cell index ((i-1) + (j-1)*ni), ordered land | glacier | sea-ice, each
 * class in grid order: the permutation the kernels run on (bit-exact contract, tests/test_partition.py). */
int noahmp_b200_column_map(noahmp_b200_ctx* ctx, int32_t* cells, long long capacity);
/* Device pointer to layer `layer` (0-based) of a state field (named like the noahmp_lsm_args member) in the
 * column-compact structure of arrays; *np = number of active columns (the plane length). */
float* noahmp_b200_device_state(noahmp_b200_ctx* ctx, const char* field, int layer, long long* np);
/* Optional diagnostic: per-cell VEGE_FLUX pass count of the last step (ni*nj int32, grid order). */
int noahmp_b200_enable_iteration_counts(noahmp_b200_ctx* ctx, int enable);
int noahmp_b200_get_iteration_counts(noahmp_b200_ctx* ctx, int32_t* out);

/* ---- forcing pipeline on the device -----------------------------------------------------
 * Replaces, for a driver that hands the raw forcing-file fields to the library at INPUT cadence instead of
 * interpolated arrays at every step:
 *   hrldas_input_interpolate / hrldas_input_copy   driver/module_hrldas_netcdf_io.F90:1351-1404
 *   VEGFRA*100, level-2 copies, RAINBL = rate*dt, DZ8W = 2*zlvl        module_hrldas_noahmp_driver.F90:336-344
 *   the CALC_DECLIN loop (COSZEN, JULIAN)                              module_hrldas_noahmp_driver.F90:350-354, :813-863
 * The interpolated planes never exist on the host; per step nothing crosses PCIe but the fetch list. */
typedef struct noahmp_forcing_fields {
  /* one forcing file (xstart:xend, ystart:yend): T2D Q2D U2D V2D PSFC LWDOWN SWDOWN RAINRATE, VEGFRA as fraction */
  const float *t, *q, *u, *v, *p, *lw, *sw, *pcp, *fpar;
} noahmp_forcing_fields;
/* lat2d / lon2d in degrees (the CALC_DECLIN arguments), zlvl = config%zlvl (30 m). */
int noahmp_b200_forcing_static(noahmp_b200_ctx* ctx, const float* lat2d, const float* lon2d, float zlvl);
/* Asynchronous upload of one bracket: slot 0 = instructA (earlier file), 1 = instructB (later file). */
int noahmp_b200_forcing_upload(noahmp_b200_ctx* ctx, int slot, const noahmp_forcing_fields* fields);
/* instructB becomes instructA (the model time passed the later file). */
int noahmp_b200_forcing_swap(noahmp_b200_ctx* ctx);
/* Fill the device forcing planes for one model step: fraction = real(idts2-idts)/real(idts2) (1 = exactly at A,
 * which is hrldas_input_copy); date parts of the step's NOWDATE; returns JULIAN through *julian. */
int noahmp_b200_forcing_apply(noahmp_b200_ctx* ctx, float fraction, int iday, int ihour, int iminute, int isecond,
                              float model_timestep, float* julian);
/* noahmp_b200_noahmplsm for a context whose forcing planes were filled by noahmp_b200_forcing_apply: the forcing
 * pointers of `args` are ignored (may be NULL); RESIDENT mode only; same row-chunk pipeline and fetch list. */
int noahmp_b200_noahmplsm_device_forcing(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args, noahmp_status* status);

/* ---- opt_run = 5: Miguez-Macho & Fan groundwater, replaces `CALL WTABLE_mmf_noahmp(...)` ----------------------
 * phys/module_sf_noahmp_groundwater.F90:14-198 (LATERALFLOW :201-295, UPDATEWTD :298-606), called by
 * land_driver_exe every STEPWTD steps (driver/module_hrldas_noahmp_driver.F90:420-436).  One member per dummy
 * argument, same names, same order; 2-D arrays (ims:ime,jms:jme), 3-D (ims:ime,1:nsoil,jms:jme). */
typedef struct noahmp_wtable_args {
  int32_t nsoil;
  const float *xland, *xice;
  float xice_threshold;
  int32_t isice;
  const int32_t* isltyp;
  const float* smoiseq;
  const float* dzs;
  float wtddt; /* minutes */
  const float *fdepth, *area, *topo;
  int32_t isurban;
  const int32_t* ivgtyp;
  const float *rivercond, *riverbed, *eqwtd, *pexp;
  float *smois, *sh2oxy, *smcwtd, *wtd, *qrf, *deeprech, *qspring, *qslat, *qrfs, *qsprings, *rech;
  int32_t ids, ide, jds, jde, kds, kde;
  int32_t ims, ime, jms, jme, kms, kme;
  int32_t its, ite, jts, jte, kts, kte;
} noahmp_wtable_args;

/* Whole call on one tile.  SYNC_FULL: uploads the INOUT arrays, downloads INOUT+OUT.  SYNC_RESIDENT: SMOIS, SH2O,
 * SMCWTD, WTD (= ZWTXY), DEEPRECH, RECH are the planes the resident noahmplsm state already holds; the other
 * arrays are uploaded on the first call and kept on the device (sync with noahmp_b200_wtable_sync_host).
 * With ids..ide / jds..jde describing the GLOBAL domain and its..ite / jts..jte the tile, a tile whose neighbours
 * are other GPUs needs its halo filled between _begin and _end (below); noahmp_b200_wtable = _begin + _end. */
int noahmp_b200_wtable(noahmp_b200_ctx* ctx, const noahmp_wtable_args* args);
int noahmp_b200_wtable_begin(noahmp_b200_ctx* ctx, const noahmp_wtable_args* args);
int noahmp_b200_wtable_end(noahmp_b200_ctx* ctx, const noahmp_wtable_args* args);
/* Device planes of KCELL and HEAD (LATERALFLOW pass 1) with a one-cell halo ring: (nj+2) rows of (ni+2) floats,
 * element (i - its + 1) + (j - jts + 1) * (ni + 2).  The caller exchanges the ring with the neighbouring tiles
 * (NCCL send/recv through its own communicator) between _begin and _end. */
int noahmp_b200_wtable_halo(noahmp_b200_ctx* ctx, float** kcell, float** head);
/* ---- multi-GPU exchanges of the path, over NCCL inside the library ----------------------------------------------
 * One process per GPU, ranks laid out as mpp_land_partition does (rank = iprocx + iprocy*nprocx,
 * mpp/module_mpp_land.F90:83-84, :124-141).  Rank 0 obtains the 128-byte NCCL id with comm_unique_id, the host
 * program broadcasts it (MPI_Bcast in the Fortran driver, torch.distributed under torchrun), every rank calls
 * comm_init.  From then on noahmp_b200_wtable exchanges the KCELL / HEAD halo with the up-to-8 neighbouring tiles
 * itself (grouped ncclSend/ncclRecv: columns, then rows carrying the corners) and the tiles reproduce the sequential
 * single-domain result; the reference's MPI build exchanges nothing and clips LATERALFLOW at tile edges
 * (module_sf_noahmp_groundwater.F90:231-234, :254-257 with ids = its). */
int noahmp_b200_comm_unique_id(void* id128);
int noahmp_b200_comm_init(noahmp_b200_ctx* ctx, const void* id128, int rank, int nranks);
/* ranks of the left, right, lower (smaller j) and upper neighbour tile, -1 at the domain edge */
int noahmp_b200_comm_neighbours(const noahmp_b200_ctx* ctx, int neighbours[4]);
void noahmp_b200_tile_neighbours(int nranks, int rank, int neighbours[4]);
/* The halo exchange alone, between _begin and _end, on `stream` (cudaStream_t as void*, NULL = the context's). */
int noahmp_b200_wtable_exchange(noahmp_b200_ctx* ctx, void* stream);
/* noahmp_b200_wtable for a device-side stepping loop (RESIDENT mode): pass 1, halo, pass 2 and the column update are
 * enqueued on `stream` behind the preceding noahmp_b200_step_device; nothing is copied and the host does not wait. */
int noahmp_b200_wtable_device(noahmp_b200_ctx* ctx, const noahmp_wtable_args* args, void* stream);

/* Global water / energy budget (north_star: NCCL "for global water/energy-balance diagnostics").  When enabled, a
 * reduction kernel follows every step and keeps eight fp64 sums over the land + glacier columns of the tile:
 *   [0] storage now: CANWAT + SNOW + WA + sum_k SMOIS_k*DZS_k*1000 (mm)   [1] precipitation RAINBL (mm, accumulated)
 *   [2] (ECAN+EDIR+ETRAN)*DT (mm, accumulated)   [3] (RUNSF+RUNSB)*DT (mm, accumulated)
 *   [4] SAV+SAG-(FIRA+HFX+LH+GRDFLX), the ERRENG residual of ERROR (noahmplsm.F90:1188-1199) (W/m2, accumulated)
 *   [5] SNOW now (mm)   [6] columns   [7] steps accumulated
 * budget_read copies them out; global != 0 first sums them over the communicator (one ncclAllReduce of 8 doubles).
 * Over an interval, d[0] - ([1]-[2]-[3]) is the sum of the per-column ERRWAT (noahmplsm.F90:1204-1226). */
int noahmp_b200_budget_enable(noahmp_b200_ctx* ctx, int enable);
int noahmp_b200_budget_read(noahmp_b200_ctx* ctx, double* out8, int global, int reset);
int noahmp_b200_wtable_sync_host(noahmp_b200_ctx* ctx, const noahmp_wtable_args* args);

/* ---- output / restart staging --------------------------------------------------
 * For a context whose state lives in HBM (RESIDENT mode): snapshot the named state fields (comma separated
 * noahmp_lsm_args member names, or "*" for all INOUT/OUT arrays) as of the latest step, and copy them into the
 * caller's host arrays in the background while later steps run.  mask_water != 0 writes -1.E33 where
 * IVGTYP == ISWATER, which is what put_var_2d / put_var_3d do to history output
 * (driver/module_hrldas_netcdf_io.F90:1950-2052: `where (vegtyp == ISWATER .and. .not. restart_flag)`); pass 0 for
 * restart files (:612-672 of module_hrldas_noahmp_driver.F90).  The host arrays are valid after _output_wait.
 * One snapshot in flight at a time (_begin waits for the previous one). */
int noahmp_b200_output_begin(noahmp_b200_ctx* ctx, const noahmp_lsm_args* args, const char* fields, int mask_water);
int noahmp_b200_output_wait(noahmp_b200_ctx* ctx);

/* ---- cold start: replaces `CALL NOAHMP_INIT(...)` -------------------------------
 * phys/module_sf_noahmpdrv.F90:847-1179 with SNOW_INIT :1182-1283, GROUNDWATER_INIT :1286-1470 and EQSMOISTURE
 * :1473-1522; call site driver/module_hrldas_noahmp_driver.F90:281-297.  One member per dummy argument, same names,
 * same order; LOGICALs are int32 (0 = .FALSE.).  MMINLU is not needed (the context already holds the tables that
 * read_mp_veg_parameters / SOIL_VEG_GEN_PARM would read); allowed_to_read is accepted and ignored for the same
 * reason.  2-D arrays (ims:ime,jms:jme); TSLB/SMOIS/SH2O/SMOISEQ (ims:ime,1:nsoil,jms:jme); TSNOXY/SNICEXY/SNLIQXY
 * (ims:ime,-2:0,jms:jme); ZSNSOXY (ims:ime,-2:nsoil,jms:jme).  The groundwater members may be NULL unless
 * iopt_run == 5.  With restart != 0 nothing is touched, as in the reference. */
typedef struct noahmp_init_args {
  float *snow, *snowh, *canwat;
  const int32_t *isltyp, *ivgtyp;
  int32_t isurban;
  float *tslb, *smois, *sh2o;
  const float* dzs;
  int32_t fndsoilw, fndsnowh, isice, iswater;
  const float* tsk;
  int32_t* isnowxy;
  float *tvxy, *tgxy, *canicexy;
  float* tmn;
  const float* xice;
  float *canliqxy, *eahxy, *tahxy, *cmxy, *chxy;
  float *fwetxy, *sneqvoxy, *alboldxy, *qsnowxy, *wslakexy, *zwtxy, *waxy;
  float *wtxy, *tsnoxy, *zsnsoxy, *snicexy, *snliqxy, *lfmassxy, *rtmassxy;
  float *stmassxy, *woodxy, *stblcpxy, *fastcpxy, *xsaixy;
  float *t2mvxy, *t2mbxy, *chstarxy;
  int32_t nsoil, restart, allowed_to_read, iopt_run;
  int32_t ids, ide, jds, jde, kds, kde;
  int32_t ims, ime, jms, jme, kms, kme;
  int32_t its, ite, jts, jte, kts, kte;
  /* optional groundwater block */
  float *smoiseq, *smcwtdxy, *rechxy, *deeprechxy, *areaxy;
  float dx, dy;
  const float *msftx, *msfty;
  float wtddt;      /* minutes */
  int32_t* stepwtd; /* OUT: max(nint(wtddt*60/dt),1) */
  float dt;
  float *qrfsxy, *qspringsxy, *qslatxy;
  const float *fdepthxy, *ht, *riverbedxy, *eqzwt, *rivercondxy, *pexpxy;
} noahmp_init_args;

/* Uploads the arrays, runs the initialisation kernels, downloads the results into the caller's arrays.  Returns
 * NOAHMP_ERR_ISLTYP when a cell of the tile has ISLTYP < 1 ("lsminit: out of range value of ISLTYP", :1018-1020),
 * NOAHMP_ERR_ARG when iopt_run == 5 and a groundwater member is missing ("Not enough fields to use groundwater
 * option in Noah-MP", :1171). */
int noahmp_b200_init(noahmp_b200_ctx* ctx, const noahmp_init_args* args);
unsigned long long noahmp_b200_sizeof_init_args(void);

/* ---- one host process driving several GPUs ------------------------------------------------
 * Replaces the IO-rank scatter / gather of the reference's MPI build (decompose_data_real/_int, write_io_real/_int,
 * mpp/module_mpp_land.F90:645-857) by tile slicing: the process holds the whole-domain arrays, GPU r owns the tile
 * mpp_land_partition_calc gives rank r, and every per-tile context copies its rows directly out of / into the global
 * arrays.  The per-tile entry points accept this as well: memory bounds (ims:ime, jms:jme) may exceed the tile bounds. */
typedef struct noahmp_b200_domain noahmp_b200_domain; /* opaque */
noahmp_b200_domain* noahmp_b200_domain_create(const noahmp_tables* tables, int global_ni, int global_nj, int ntiles,
                                              const int* devices /* [ntiles] CUDA device of each tile */);
void noahmp_b200_domain_destroy(noahmp_b200_domain* d);
int noahmp_b200_domain_ntiles(const noahmp_b200_domain* d);
noahmp_b200_ctx* noahmp_b200_domain_tile(noahmp_b200_domain* d, int r);
int noahmp_b200_domain_tile_bounds(const noahmp_b200_domain* d, int r, int* xstart, int* xend, int* ystart, int* yend);
/* sync / math mode, fetch and push lists (NULL = leave), forcing hints of every tile */
int noahmp_b200_domain_configure(noahmp_b200_domain* d, int sync_mode, int math_mode, const char* fetch, const char* push,
                                 unsigned hints);
/* CALL noahmplsm(...) with the GLOBAL arrays (memory bounds = the domain; the tile bounds of `args` are ignored): one
 * worker thread per GPU runs its tile's call.  status = first failing column in the reference's loop order. */
int noahmp_b200_domain_noahmplsm(noahmp_b200_domain* d, const noahmp_lsm_args* args, noahmp_status* status);
int noahmp_b200_domain_sync_host(noahmp_b200_domain* d, const noahmp_lsm_args* args);

/* ---- domain decomposition: replaces mpp_land_partition arithmetic ----------------------------
 * mpp/module_mpp_land.F90:124-141 (process grid), :245-288 (tile extents). All 1-based inclusive. */
void noahmp_b200_proc_grid(int nproc, int* nprocx, int* nprocy);
void noahmp_b200_tile(int global_nx, int global_ny, int nproc, int rank, int* xstart, int* xend,
                      int* ystart, int* yend);

#ifdef __cplusplus
}
#endif
#endif /* NOAHMP_B200_H */
