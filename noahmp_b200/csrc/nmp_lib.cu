// nmp_lib.cu — host side of libnoahmp_b200.so: the C-ABI of include/noahmp_b200.h.
//
// Replaces the grid loop of `noahmplsm` (phys/module_sf_noahmpdrv.F90:376-840): classification of cells
// into water / land / glacier / sea-ice (:426-441), the ITIMESTEP==1 initialisation of water and sea-ice
// rows (:399-419), the per-column gather/scatter (:449-512, :728-835) and the status convention that
// stands in for wrf_error_fatal.  There is no CPU fallback: without a CUDA device create() fails.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "nmp_fields.h"
#include "nmp_math.h"

using namespace nmpf;

// launchers of the two physics builds (nmp_kernels_fast.cu / nmp_kernels_parity.cu)
const char* nmp_launch_step_fast(const StepParams& base, const StepRange& r, cudaStream_t stream, long long* launches);
const char* nmp_launch_step_parity(const StepParams& base, const StepRange& r, cudaStream_t stream, long long* launches);
void nmp_launch_wtable_fast(const WtParams& w, cudaStream_t stream, long long* launches, int phase);
void nmp_launch_wtable_parity(const WtParams& w, cudaStream_t stream, long long* launches, int phase);
void nmp_launch_forcing_fast(const ForcingParams& f, cudaStream_t stream, long long* launches);
void nmp_launch_forcing_parity(const ForcingParams& f, cudaStream_t stream, long long* launches);
void nmp_launch_init(const InitParams& p, cudaStream_t stream, long long* launches);

static thread_local std::string g_last_error;
static void set_error(const std::string& s) { g_last_error = s; }

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
      return NOAHMP_ERR_CUDA;                                                                        \
    }                                                                                                \
  } while (0)

// ---- field table: host pointer of every state field inside noahmp_lsm_args ---------------------------
struct FieldInfo {
  const char* name;
  size_t arg_offset;
  int layers, kind, slot;
};
static const FieldInfo kFields[NFIELDS] = {
#define X(nm, nl, k) {#nm, offsetof(noahmp_lsm_args, nm), nl, k, 0},
    NMP_STATE_FIELDS(X)
#undef X
};
static inline float* host_ptr(const noahmp_lsm_args* a, int f) {
  return *reinterpret_cast<float* const*>(reinterpret_cast<const char*>(a) + kFields[f].arg_offset);
}

// descriptor of one (field, layer) plane for the batched gather / scatter kernels
struct PlaneDesc {
  float* grid;     // staging array of the field, Fortran (i,k,j) layout
  int plane;       // plane index in the compact state
  int layer, layers;
};

#include "nmp_comm.cuh"

struct noahmp_b200_ctx {
  int device = 0, ni = 0, nj = 0;
  // Layout of the caller's host arrays in the current call: memory extent mi x mj (ims:ime, jms:jme) and the offset
  // (x0, y0) of this context's tile (its:ite, jts:jte) inside it.  HRLDAS allocates exactly the tile (mi = ni, x0 = 0);
  // a single host process that holds the whole domain and drives one context per GPU passes the same global arrays to
  // every context with different tile bounds (row f4: tile slicing instead of the IO-rank scatter / gather).
  int h_mi = 0, h_mj = 0, h_x0 = 0, h_y0 = 0, h_its = 1, h_jts = 1;
  NmpComm comm;                            // NCCL communicator of the tiles of one domain (optional)
  double* d_budget = nullptr;              // NBUDGET running sums (noahmp_b200_budget_*)
  bool budget_on = false;
  int budget_steps = 0;
  long long ncell = 0;
  cudaStream_t stream = nullptr;
  noahmp_tables* d_tables = nullptr;
  float* d_forc[NFORC] = {};
  float* d_stat[NSTATIC] = {};
  float* d_state = nullptr;
  long long np = 0, np_alloc = 0;
  float* d_grid[NFIELDS] = {};
  bool grid_init[NFIELDS] = {};
  PlaneDesc* d_planes = nullptr;
  int nplanes_desc = 0;
  int* d_cell = nullptr;
  unsigned char* d_class = nullptr;
  void* d_cub = nullptr;
  size_t cub_bytes = 0;
  int* d_nsel = nullptr;
  int nclass[4] = {0, 0, 0, 0};  // water, land, glacier, seaice
  unsigned long long* d_errkey = nullptr;
  int* d_errcount = nullptr;
  unsigned long long* h_errkey = nullptr;  // pinned
  int* h_errcount = nullptr;
  int* d_vege_iters = nullptr;
  int sync_mode = NOAHMP_SYNC_FULL;
  int math_mode = 0;
  bool uploaded = false, classified = false;
  bool pin_host = true;
  StepParams base{};
  long long launches = 0;
  std::string variant;
  std::unordered_map<const void*, size_t> registered;
  std::unordered_map<const void*, unsigned long long> reg_used;  // last use (pin_clock) of a registration
  unsigned long long pin_clock = 0;
  size_t pinned_bytes = 0, pin_budget = (size_t)64 << 30;         // LRU bound of the page-locked bytes
  // chunk pipeline of the RESIDENT-mode call
  std::vector<int> h_cell;                 // host copy of the column map (grid order, as classified)
  // column re-binning (divergence control): land columns are physically re-ordered inside each row chunk by
  // (canopy tile computed in the previous step yes/no, snow-layer count)
  int rebin_interval = 20, steps_since_rebin = 0, rebins = 0, bin_chunks = 0;
  // A due re-binning first asks whether the order still holds: the bin keys are recomputed, the places where they
  // decrease along the compact order counted (two per column whose bin changed), the count read back without waiting
  // (pinned word + event, looked at in a later step) -- and the 4.9 ms permutation is done only if more than
  // rebin_min_changed of the land columns are out of place (NOAHMP_B200_REBIN_MIN_CHANGED; 0 = always permute).
  float rebin_min_changed = 1e-3f;
  bool check_pending = false;
  cudaEvent_t ev_check = nullptr;
  int *d_viol = nullptr, *h_viol = nullptr;
  int rebin_checks = 0, rebin_skips = 0;
  int bin_sub = 8;  // row groups per chunk that get bins of their own (NOAHMP_B200_BIN_SUB; see bin_key_kernel)
  bool binned = false;
  float* d_state2 = nullptr;               // PERMUTE_GROUP scratch planes of the re-binning
  std::vector<int> moved_planes;           // planes the re-binning permutes (INOUT + internal)
  cudaEvent_t ev_rebin = nullptr;
  unsigned char* d_plane_kind = nullptr;
  std::vector<int> h_chunk;  // staging of bin_key_kernel's chunk table (kept alive for the asynchronous copy)
  int *d_cell2 = nullptr, *d_keys = nullptr, *d_keys2 = nullptr, *d_perm = nullptr, *d_iota = nullptr, *d_chunk = nullptr;
  std::vector<int> ch_land, ch_glac, ch_sea;  // compact range boundaries of the row chunks (size nchunks+1)
  std::vector<int> fetch;                  // fields refreshed on the host by every noahmplsm call
  std::vector<int> push;                   // INOUT fields re-read from the host by every RESIDENT-mode call
  PlaneDesc* d_fetch_planes = nullptr;     // plane descriptors of the fetch list (one scatter launch per range)
  int n_fetch_planes = 0;
  unsigned hints = 0;                      // NOAHMP_HINT_* (forcing planes that need no upload this call)
  bool forc_valid[NFORC] = {};             // plane holds an upload of the caller's array
  int nchunks = 0;                         // 0 = automatic
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_k, ev_out, ev_plane;
  cudaEvent_t ev_t0 = nullptr;
  // output staging (row f3): snapshot buffer and its events
  float* d_outstage = nullptr;
  size_t outstage_words = 0;
  cudaEvent_t ev_outready = nullptr, ev_outdone = nullptr;
  bool out_pending = false;
  cudaStream_t last_step_stream = nullptr;  // a caller stream the latest device-side step ran on
  int iswater = 16;
  bool trace = false;  // NOAHMP_B200_TRACE: print the per-chunk timeline of every RESIDENT-mode call
  // opt_run = 5 groundwater: grid-order planes (see WtPlane) and the haloed KCELL / HEAD planes
  float* d_wt[12] = {};
  float *d_kcell = nullptr, *d_head = nullptr;
  bool wt_init = false;
  // on-device forcing pipeline (row f2): two brackets of 9 planes, lat / lon
  float* d_fb[2][9] = {};
  float *d_lat = nullptr, *d_lon = nullptr;
  float zlvl = 30.f;
  cudaEvent_t ev_fb[2] = {};
  bool fb_static = false, fb_loaded[2] = {false, false};
};

// ---- small kernels ------------------------------------------------------------------------------------
// noahmpdrv.F90:426-441
__global__ void classify_kernel(const float* __restrict__ xland, const float* __restrict__ xice,
                                const float* __restrict__ ivgtyp_bits, float xice_thres, int isice, long long ncell,
                                unsigned char* __restrict__ cls) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  int ICE;
  if (xice[c] >= xice_thres) ICE = 1;
  else if (__float_as_int(ivgtyp_bits[c]) == isice) ICE = -1;
  else ICE = 0;
  unsigned char k;
  if ((xland[c] - 1.5f) >= 0.f) k = CL_WATER;
  else if (ICE == 1) k = CL_SEAICE;
  else if (ICE == -1) k = CL_GLACIER;
  else k = CL_LAND;
  cls[c] = k;
}

struct ClassIs {
  const unsigned char* cls;
  unsigned char want;
  __host__ __device__ bool operator()(int c) const { return cls[c] == want; }
};

// grid order (Fortran (i,k,j)) <-> compact planes; blockIdx.y = plane descriptor
__global__ void gather_kernel(const PlaneDesc* __restrict__ planes, const int* __restrict__ cell, float* state,
                              long long np, int ni) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= np) return;
  const PlaneDesc d = planes[blockIdx.y];
  const int c = cell[n];
  const int i = c % ni, j = c / ni;
  state[(long long)d.plane * np + n] = d.grid[(long long)i + (long long)d.layer * ni + (long long)j * ni * d.layers];
}
__global__ void scatter_kernel(const PlaneDesc* __restrict__ planes, const int* __restrict__ cell,
                               const float* __restrict__ state, long long np, int ni, long long first = 0,
                               long long count = -1) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (count >= 0) {
    if (n >= count) return;
    n += first;
  }
  if (n >= np) return;
  const PlaneDesc d = planes[blockIdx.y];
  const int c = cell[n];
  const int i = c % ni, j = c / ni;
  d.grid[(long long)i + (long long)d.layer * ni + (long long)j * ni * d.layers] = state[(long long)d.plane * np + n];
}

// static inputs (grid order) -> their compact planes; blockIdx.y = static plane
struct StatPtrs { const float* p[NSTATIC]; };
__global__ void gather_static_kernel(StatPtrs st, const int* __restrict__ cell, float* state, long long np) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= np) return;
  state[(long long)(PLANE_STATIC0 + blockIdx.y) * np + n] = st.p[blockIdx.y][cell[n]];
}

// ITIMESTEP == 1 initialisation of open-water cells (noahmpdrv.F90:399-411); works on the grid-order staging
// arrays because those cells have no compact column (the XICE == 1 branch lives in seaice_kernel).
__global__ void first_step_water_kernel(const float* __restrict__ xland, const float* __restrict__ xice, float* smstav,
                                        float* smstot, float* smois, float* tslb, int ni, long long ncell) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int i = (int)(c % ni);
  const long long j = c / ni;
  if ((xland[c] - 1.5f) >= 0.f) {
    smstav[c] = 1.0f;
    smstot[c] = 1.0f;
    for (int k = 0; k < NOAHMP_NSOIL; ++k) {
      smois[i + (long long)k * ni + j * ni * NOAHMP_NSOIL] = 1.0f;
      tslb[i + (long long)k * ni + j * ni * NOAHMP_NSOIL] = 273.16f;
    }
  }
}

// sea-ice columns: SH2O = 1, XLAI = 0.01 (noahmpdrv.F90:436-441); at ITIMESTEP == 1 the row initialisation
// (:412-418) also sets SMSTAV = SMSTOT = SMOIS = 1 where XICE == 1
__global__ void seaice_kernel(float* state, const int* __restrict__ cell, const float* __restrict__ xice, long long np,
                              int first, int count, int itimestep) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  long long n = (long long)first + t;
  if (itimestep == 1 && xice[cell[n]] == 1.f) {
    state[(long long)NMP_SLOT(smstav) * np + n] = 1.0f;
    state[(long long)NMP_SLOT(smstot) * np + n] = 1.0f;
    for (int k = 0; k < NOAHMP_NSOIL; ++k) state[(long long)(NMP_SLOT(smois) + k) * np + n] = 1.0f;
  }
  for (int k = 0; k < NOAHMP_NSOIL; ++k) state[(long long)(NMP_SLOT(sh2o) + k) * np + n] = 1.0f;
  state[(long long)NMP_SLOT(xlaixy) * np + n] = 0.01f;
}

// ---- helpers ------------------------------------------------------------------------------------------
static void unpin_one(noahmp_b200_ctx* ctx, const void* p) {
  auto it = ctx->registered.find(p);
  if (it == ctx->registered.end()) return;
  if (it->second) {
    if (cudaHostUnregister(const_cast<void*>(p)) != cudaSuccess) cudaGetLastError();
    ctx->pinned_bytes -= std::min(ctx->pinned_bytes, it->second);
  }
  ctx->registered.erase(it);
  ctx->reg_used.erase(p);
}
static void pin(noahmp_b200_ctx* ctx, const void* p, size_t bytes) {
  // Only large arrays are page-locked: they come from mmap'ed allocations of their own, whereas small heap
  // arrays can share a page with other data, and a partially registered range makes cudaMemcpyAsync fail.
  if (!ctx->pin_host || !p || bytes < (size_t)(4u << 20)) return;
  auto it = ctx->registered.find(p);
  if (it != ctx->registered.end() && (it->second >= bytes || it->second == 0)) {
    ctx->reg_used[p] = ctx->pin_clock;
    return;
  }
  if (it != ctx->registered.end()) unpin_one(ctx, p);
  // A driver that hands fresh arrays to every call (a Python loop allocating its forcing per step) must not grow the
  // page-locked set without bound: registrations not used by the current or the previous call are dropped, oldest
  // first, once the budget (NOAHMP_B200_PIN_BUDGET_GB, default 64) would be exceeded.
  if (ctx->pinned_bytes + bytes > ctx->pin_budget) {
    std::vector<std::pair<unsigned long long, const void*>> old;
    for (auto& kv : ctx->reg_used)
      if (kv.second + 1 < ctx->pin_clock) old.push_back({kv.second, kv.first});
    std::sort(old.begin(), old.end());
    if (!old.empty()) cudaDeviceSynchronize();  // no copy may still be reading what is about to be unlocked
    for (auto& e : old) {
      if (ctx->pinned_bytes + bytes <= ctx->pin_budget) break;
      unpin_one(ctx, e.second);
    }
  }
  cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterPortable);  // contexts on other GPUs may share the array
  if (e == cudaSuccess) { ctx->registered[p] = bytes; ctx->pinned_bytes += bytes; }
  else {
    cudaGetLastError();         // already pinned by the caller, or not pinnable: plain pageable copy
    ctx->registered[p] = 0;     // remember not to retry every call
  }
  ctx->reg_used[p] = ctx->pin_clock;
}
static void unpin_all(noahmp_b200_ctx* ctx) {
  for (auto& kv : ctx->registered)
    if (kv.second) { if (cudaHostUnregister(const_cast<void*>(kv.first)) != cudaSuccess) cudaGetLastError(); }
  ctx->registered.clear();
  ctx->reg_used.clear();
  ctx->pinned_bytes = 0;
}

static int check_bounds(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  // HRLDAS allocates exactly the tile: ims=its ... (driver/module_hrldas_noahmp_driver.F90:112-129); memory bounds that
  // CONTAIN the tile are accepted too (WRF-style halos; one host process holding the whole domain)
  if (a->ims > a->its || a->ime < a->ite || a->jms > a->jts || a->jme < a->jte) {
    set_error("memory bounds must contain the tile bounds (ims<=its, ime>=ite, jms<=jts, jme>=jte)");
    return NOAHMP_ERR_ARG;
  }
  if (a->ite - a->its + 1 != ctx->ni || a->jte - a->jts + 1 != ctx->nj) {
    set_error("tile extent differs from the context's ni x nj");
    return NOAHMP_ERR_ARG;
  }
  if (a->nsoil != NOAHMP_NSOIL) { set_error("nsoil must be 4"); return NOAHMP_ERR_ARG; }
  if (a->kme - a->kms + 1 < 2 || a->kms > 1 || a->kme < 2 || a->kts != 1) {
    set_error("vertical bounds must contain levels 1 and 2 with kts=1");
    return NOAHMP_ERR_ARG;
  }
  ctx->h_mi = a->ime - a->ims + 1;
  ctx->h_mj = a->jme - a->jms + 1;
  ctx->h_x0 = a->its - a->ims;
  ctx->h_y0 = a->jts - a->jms;
  ctx->h_its = a->its;
  ctx->h_jts = a->jts;
  return 0;
}

// Copy rows [j0, j1) of a tile-shaped device array (ni x layers x nj, Fortran (i,k,j) order) from / to the caller's
// array of the same field, whose memory extent may exceed the tile (h_mi, h_x0, h_y0).  One contiguous copy when the
// memory is the tile, else a pitched copy of (j1-j0)*layers rows of ni words.
static cudaError_t copy_field_rows(const noahmp_b200_ctx* ctx, float* dev, float* host, int layers, int j0, int j1,
                                   bool to_device, cudaStream_t st) {
  const size_t ni = ctx->ni, mi = ctx->h_mi, L = layers;
  float* h = host + (size_t)ctx->h_x0 + ((size_t)(ctx->h_y0 + j0) * L) * mi;
  float* d = dev + (size_t)j0 * L * ni;
  const size_t rows = (size_t)(j1 - j0) * L;
  if (rows == 0) return cudaSuccess;
  if (mi == ni)
    return to_device ? cudaMemcpyAsync(d, h, sizeof(float) * ni * rows, cudaMemcpyHostToDevice, st)
                     : cudaMemcpyAsync(h, d, sizeof(float) * ni * rows, cudaMemcpyDeviceToHost, st);
  return to_device ? cudaMemcpy2DAsync(d, sizeof(float) * ni, h, sizeof(float) * mi, sizeof(float) * ni, rows,
                                       cudaMemcpyHostToDevice, st)
                   : cudaMemcpy2DAsync(h, sizeof(float) * mi, d, sizeof(float) * ni, sizeof(float) * ni, rows,
                                       cudaMemcpyDeviceToHost, st);
}
// bytes of the caller's whole array of a field with `layers` layers (what gets page-locked)
static size_t host_bytes(const noahmp_b200_ctx* ctx, int layers) {
  return sizeof(float) * (size_t)ctx->h_mi * (size_t)ctx->h_mj * (size_t)layers;
}

static void fill_scalars(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  StepParams& b = ctx->base;
  b.dx = a->dx;
  b.xice_thres = a->xice_thres;
  b.isice = a->isice;
  b.isurban = a->isurban;
  b.iz0tlnd = a->iz0tlnd;
  // ZSOIL (noahmpdrv.F90:392-395)
  b.zsoil[0] = -a->dzs[0];
  for (int k = 1; k < NOAHMP_NSOIL; ++k) b.zsoil[k] = -a->dzs[k] + b.zsoil[k - 1];
  const int o[12] = {a->idveg,    a->iopt_crs, a->iopt_btr, a->iopt_run, a->iopt_sfc,  a->iopt_frz,
                     a->iopt_inf, a->iopt_rad, a->iopt_alb, a->iopt_snf, a->iopt_tbot, a->iopt_stc};
  for (int k = 0; k < 12; ++k) b.opt[k] = o[k];
}

static int year_length(int yr) {  // noahmpdrv.F90:381-390
  int YEARLEN = 365;
  if (yr % 4 == 0) {
    YEARLEN = 366;
    if (yr % 100 == 0) {
      YEARLEN = 365;
      if (yr % 400 == 0) YEARLEN = 366;
    }
  }
  return YEARLEN;
}

static int check_options(const noahmp_b200_ctx* ctx) {
  const int* o = ctx->base.opt;
  const int lo[12] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  const int hi[12] = {5, 2, 3, 5, 2, 2, 2, 3, 2, 3, 2, 2};
  static const char* nm[12] = {"dveg", "opt_crs", "opt_btr", "opt_run", "opt_sfc", "opt_frz",
                               "opt_inf", "opt_rad", "opt_alb", "opt_snf", "opt_tbot", "opt_stc"};
  for (int k = 0; k < 12; ++k) {
    if (o[k] < lo[k] || o[k] > hi[k]) {
      // opt_sfc 3/4 (MYJ / YSU) read lookup tables nothing initialises offline
      set_error(std::string("unsupported option value for ") + nm[k]);
      return NOAHMP_ERR_OPTION;
    }
  }
  return 0;
}

static int ensure_grid(noahmp_b200_ctx* ctx, int f) {
  if (!ctx->d_grid[f]) {
    CK(cudaMalloc(&ctx->d_grid[f], sizeof(float) * ctx->ncell * kFields[f].layers));
    CK(cudaMemsetAsync(ctx->d_grid[f], 0, sizeof(float) * ctx->ncell * kFields[f].layers, ctx->stream));
    // descriptors must be refreshed
    ctx->nplanes_desc = 0;
  }
  return 0;
}

static int build_plane_descs(noahmp_b200_ctx* ctx) {
  if (ctx->nplanes_desc == NPLANES) return 0;
  std::vector<PlaneDesc> h;
  for (int f = 0; f < NFIELDS; ++f) {
    int rc = ensure_grid(ctx, f);
    if (rc) return rc;
    for (int k = 0; k < kFields[f].layers; ++k) h.push_back({ctx->d_grid[f], kSlots.slot[f] + k, k, kFields[f].layers});
  }
  if (!ctx->d_planes) CK(cudaMalloc(&ctx->d_planes, sizeof(PlaneDesc) * NPLANES));
  CK(cudaMemcpyAsync(ctx->d_planes, h.data(), sizeof(PlaneDesc) * NPLANES, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->nplanes_desc = NPLANES;
  return 0;
}

// H2D of rows [j0, j1) of one 2-D plane; for 3-D atmospheric arrays (i,k,j) picks level `lev` (1-based)
static int h2d_rows(noahmp_b200_ctx* ctx, float* dst, const float* src, int nk, int kms, int lev, int j0, int j1,
                    cudaStream_t st) {
  if (nk == 1) {
    CK(copy_field_rows(ctx, dst, const_cast<float*>(src), 1, j0, j1, true, st));
  } else {
    const size_t ni = ctx->ni, mi = ctx->h_mi;
    const float* h = src + (size_t)ctx->h_x0 + ((size_t)(ctx->h_y0 + j0) * nk + (size_t)(lev - kms)) * mi;
    CK(cudaMemcpy2DAsync(dst + (size_t)j0 * ni, sizeof(float) * ni, h, sizeof(float) * mi * nk, sizeof(float) * ni, j1 - j0,
                         cudaMemcpyHostToDevice, st));
  }
  return 0;
}
static int h2d_plane(noahmp_b200_ctx* ctx, float* dst, const float* src, int nk, int kms, int lev) {
  pin(ctx, src, host_bytes(ctx, nk));
  return h2d_rows(ctx, dst, src, nk, kms, lev, 0, ctx->nj, ctx->stream);
}

static int upload_forcing(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  const int nk = a->kme - a->kms + 1, kms = a->kms;
  int rc = 0;
  rc |= h2d_plane(ctx, ctx->d_forc[FC_COSZIN], a->coszin, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_T], a->t3d, nk, kms, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_QV], a->qv3d, nk, kms, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_U], a->u_phy, nk, kms, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_V], a->v_phy, nk, kms, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_SWDOWN], a->swdown, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_GLW], a->glw, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_P1], a->p8w3d, nk, kms, a->kts);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_P2], a->p8w3d, nk, kms, a->kts + 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_RAINBL], a->rainbl, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_VEGFRA], a->vegfra, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_forc[FC_DZ8W], a->dz8w, nk, kms, 1);
  return rc ? NOAHMP_ERR_CUDA : 0;
}

static int upload_static(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  int rc = 0;
  rc |= h2d_plane(ctx, ctx->d_stat[ST_IVGTYP], reinterpret_cast<const float*>(a->ivgtyp), 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_ISLTYP], reinterpret_cast<const float*>(a->isltyp), 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_VEGMAX], a->vegmax, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_TMN], a->tmn, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_XLATIN], a->xlatin, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_XLAND], a->xland, 1, 1, 1);
  rc |= h2d_plane(ctx, ctx->d_stat[ST_XICE], a->xice, 1, 1, 1);
  return rc ? NOAHMP_ERR_CUDA : 0;
}

// cell classification + stable compaction: land | glacier | sea-ice, each in grid order
static int classify(noahmp_b200_ctx* ctx) {
  const long long nc = ctx->ncell;
  const int T = 256;
  classify_kernel<<<(unsigned)((nc + T - 1) / T), T, 0, ctx->stream>>>(
      ctx->d_stat[ST_XLAND], ctx->d_stat[ST_XICE], ctx->d_stat[ST_IVGTYP], ctx->base.xice_thres, ctx->base.isice, nc,
      ctx->d_class);
  ctx->launches++;
  thrust::counting_iterator<int> it(0);
  const unsigned char order[3] = {CL_LAND, CL_GLACIER, CL_SEAICE};
  int offset = 0;
  for (int k = 0; k < 3; ++k) {
    ClassIs pred{ctx->d_class, order[k]};
    size_t need = 0;
    CK(cub::DeviceSelect::If(nullptr, need, it, ctx->d_cell + offset, ctx->d_nsel, (int)nc, pred, ctx->stream));
    if (need > ctx->cub_bytes) {
      if (ctx->d_cub) CK(cudaFree(ctx->d_cub));
      CK(cudaMalloc(&ctx->d_cub, need));
      ctx->cub_bytes = need;
    }
    CK(cub::DeviceSelect::If(ctx->d_cub, need, it, ctx->d_cell + offset, ctx->d_nsel, (int)nc, pred, ctx->stream));
    int nsel = 0;
    CK(cudaMemcpyAsync(&nsel, ctx->d_nsel, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nclass[order[k]] = nsel;
    offset += nsel;
  }
  ctx->np = offset;
  ctx->nclass[CL_WATER] = (int)(nc - offset);
  if (ctx->np > ctx->np_alloc) {
    if (ctx->d_state) CK(cudaFree(ctx->d_state));
    if (ctx->d_state2) {
      CK(cudaFree(ctx->d_state2)); ctx->d_state2 = nullptr;
      for (int** q : {&ctx->d_cell2, &ctx->d_keys, &ctx->d_keys2, &ctx->d_perm, &ctx->d_iota, &ctx->d_chunk}) {
        CK(cudaFree(*q)); *q = nullptr;
      }
    }
    CK(cudaMalloc(&ctx->d_state, sizeof(float) * (size_t)NPLANES_ALLOC * (size_t)ctx->np));
    CK(cudaMemsetAsync(ctx->d_state + (size_t)PLANE_PREV_ITERS * (size_t)ctx->np, 0, sizeof(float) * (size_t)ctx->np,
                       ctx->stream));
    ctx->np_alloc = ctx->np;
  }
  ctx->h_cell.resize((size_t)ctx->np);
  if (ctx->np) CK(cudaMemcpy(ctx->h_cell.data(), ctx->d_cell, sizeof(int) * ctx->np, cudaMemcpyDeviceToHost));
  ctx->classified = true;
  ctx->binned = false;
  ctx->check_pending = false;
  ctx->bin_chunks = 0;
  ctx->steps_since_rebin = 0;
  return 0;
}

static int gather_fields(noahmp_b200_ctx* ctx) {
  if (ctx->np == 0) return 0;
  const int T = 256;
  {
    StatPtrs st;
    for (int f = 0; f < NSTATIC; ++f) st.p[f] = ctx->d_stat[f];
    dim3 g((unsigned)((ctx->np + T - 1) / T), NSTATIC);
    gather_static_kernel<<<g, T, 0, ctx->stream>>>(st, ctx->d_cell, ctx->d_state, ctx->np);
    ctx->launches++;
  }
  dim3 grid((unsigned)((ctx->np + T - 1) / T), NPLANES);
  gather_kernel<<<grid, T, 0, ctx->stream>>>(ctx->d_planes, ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}
static int scatter_fields(noahmp_b200_ctx* ctx) {
  if (ctx->np == 0) return 0;
  const int T = 256;
  dim3 grid((unsigned)((ctx->np + T - 1) / T), NPLANES);
  scatter_kernel<<<grid, T, 0, ctx->stream>>>(ctx->d_planes, ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// H2D of the state arrays into staging; `all` also brings the OUT arrays (needed once so that untouched
// water cells keep the caller's values on the way back)
static int upload_state(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, bool all) {
  for (int f = 0; f < NFIELDS; ++f) {
    const bool want = kFields[f].kind == NMP_K_INOUT || all || !ctx->grid_init[f];
    if (!want) continue;
    const float* src = host_ptr(a, f);
    if (!src) { set_error(std::string("null array: ") + kFields[f].name); return NOAHMP_ERR_ARG; }
    pin(ctx, src, host_bytes(ctx, kFields[f].layers));
    CK(copy_field_rows(ctx, ctx->d_grid[f], const_cast<float*>(src), kFields[f].layers, 0, ctx->nj, true, ctx->stream));
    ctx->grid_init[f] = true;
  }
  return 0;
}

static int download_state(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  for (int f = 0; f < NFIELDS; ++f) {
    if (f == F_smoiseq) continue;  // INTENT(IN) in effect: never written by noahmplsm
    float* dst = host_ptr(a, f);
    pin(ctx, dst, host_bytes(ctx, kFields[f].layers));
    CK(copy_field_rows(ctx, ctx->d_grid[f], dst, kFields[f].layers, 0, ctx->nj, false, ctx->stream));
  }
  return 0;
}

static void decode_status(noahmp_b200_ctx* ctx, noahmp_status* st) {
  if (!st) return;
  st->code = 0; st->i = 0; st->j = 0; st->count = *ctx->h_errcount; st->value = 0.f;
  if (*ctx->h_errcount > 0) {
    const unsigned long long key = *ctx->h_errkey;
    const unsigned cell = (unsigned)(key >> 39);
    st->code = (int)((key >> 32) & 0x7f);
    unsigned vb = (unsigned)(key & 0xffffffffu);
    memcpy(&st->value, &vb, 4);
    st->i = (int)(cell % (unsigned)ctx->ni) + ctx->h_its;  // Fortran (i,j) in the caller's index space
    st->j = (int)(cell / (unsigned)ctx->ni) + ctx->h_jts;
  }
}

// ---- C ABI ----------------------------------------------------------------------------------------------
extern "C" {

struct ChunkRanges;
static int budget_accumulate(noahmp_b200_ctx* ctx, float dt, cudaStream_t s);
static int chunk_ranges(noahmp_b200_ctx* ctx, int nchunks, ChunkRanges* out);
static int auto_chunks(const noahmp_b200_ctx* ctx);
static int rebin(noahmp_b200_ctx* ctx, cudaStream_t s);

const char* noahmp_b200_last_error(void) { return g_last_error.c_str(); }

noahmp_b200_ctx* noahmp_b200_create(int device, const noahmp_tables* tables, int ni, int nj) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (this library has no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= ndev || !tables || ni <= 0 || nj <= 0 || (long long)ni * nj >= (1LL << 25)) {
    set_error("bad arguments to noahmp_b200_create (need 0 < ni*nj < 2^25 per tile)");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major < 10) {
    set_error(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
              "; this library carries sm_100a code only");
    return nullptr;
  }
  auto* ctx = new noahmp_b200_ctx();
  ctx->device = device; ctx->ni = ni; ctx->nj = nj; ctx->ncell = (long long)ni * nj;
  ctx->iswater = tables->iswater;
  const char* env = getenv("NOAHMP_B200_MATH");
  if (env && (!strcmp(env, "parity") || !strcmp(env, "1"))) ctx->math_mode = 1;
  env = getenv("NOAHMP_B200_PIN");
  if (env && !strcmp(env, "0")) ctx->pin_host = false;
  env = getenv("NOAHMP_B200_REBIN_MIN_CHANGED");
  if (env && atof(env) >= 0.) ctx->rebin_min_changed = (float)atof(env);
  env = getenv("NOAHMP_B200_BIN_SUB");
  if (env && atoi(env) >= 1 && atoi(env) <= 8) ctx->bin_sub = atoi(env);
  env = getenv("NOAHMP_B200_REBIN");  // steps between two re-binnings of the land columns (0 = never); tuning aid
  if (env && atoi(env) >= 0) ctx->rebin_interval = atoi(env);
  env = getenv("NOAHMP_B200_PIN_BUDGET_GB");
  if (env && atof(env) > 0.) ctx->pin_budget = (size_t)(atof(env) * 1073741824.0);
  bool ok = true;
  auto A = [&](void** p, size_t bytes) { if (ok && cudaMalloc(p, bytes) != cudaSuccess) ok = false; };
  ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
  A((void**)&ctx->d_tables, sizeof(noahmp_tables));
  for (int f = 0; f < NFORC; ++f) A((void**)&ctx->d_forc[f], sizeof(float) * ctx->ncell);
  for (int f = 0; f < NSTATIC; ++f) A((void**)&ctx->d_stat[f], sizeof(float) * ctx->ncell);
  A((void**)&ctx->d_cell, sizeof(int) * ctx->ncell);
  A((void**)&ctx->d_class, ctx->ncell);
  A((void**)&ctx->d_nsel, sizeof(int));
  A((void**)&ctx->d_errkey, sizeof(unsigned long long));
  A((void**)&ctx->d_errcount, sizeof(int));
  if (ok) ok = cudaMallocHost((void**)&ctx->h_errkey, sizeof(unsigned long long)) == cudaSuccess;
  if (ok) ok = cudaMallocHost((void**)&ctx->h_errcount, sizeof(int)) == cudaSuccess;
  if (ok) ok = cudaMemcpy(ctx->d_tables, tables, sizeof(noahmp_tables), cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) {
    set_error(std::string("allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    noahmp_b200_destroy(ctx);
    return nullptr;
  }
  *ctx->h_errcount = 0;
  *ctx->h_errkey = 0;
  StepParams& b = ctx->base;
  for (int f = 0; f < NFORC; ++f) b.forc[f] = ctx->d_forc[f];
  for (int f = 0; f < NSTATIC; ++f) b.stat[f] = ctx->d_stat[f];
  b.tables = ctx->d_tables;
  b.err_key = ctx->d_errkey;
  b.err_count = ctx->d_errcount;
  b.vege_iters = nullptr;
  b.ni = ni;
  return ctx;
}

void noahmp_b200_destroy(noahmp_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  unpin_all(ctx);
  cudaFree(ctx->d_tables);
  for (auto p : ctx->d_forc) cudaFree(p);
  for (auto p : ctx->d_stat) cudaFree(p);
  for (auto p : ctx->d_grid) cudaFree(p);
  cudaFree(ctx->d_state); cudaFree(ctx->d_planes); cudaFree(ctx->d_cell); cudaFree(ctx->d_class);
  cudaFree(ctx->d_cub); cudaFree(ctx->d_nsel); cudaFree(ctx->d_errkey); cudaFree(ctx->d_errcount);
  cudaFree(ctx->d_vege_iters); cudaFree(ctx->d_fetch_planes);
  cudaFree(ctx->d_budget); cudaFree(ctx->comm.d_send); cudaFree(ctx->comm.d_recv); cudaFree(ctx->comm.d_budget_sum);
  if (ctx->comm.comm && nccl_api()) nccl_api()->CommDestroy(ctx->comm.comm);
  cudaFree(ctx->d_state2); cudaFree(ctx->d_cell2); cudaFree(ctx->d_keys); cudaFree(ctx->d_keys2);
  cudaFree(ctx->d_perm); cudaFree(ctx->d_iota); cudaFree(ctx->d_chunk); cudaFree(ctx->d_plane_kind);
  cudaFree(ctx->d_viol); if (ctx->h_viol) cudaFreeHost(ctx->h_viol); if (ctx->ev_check) cudaEventDestroy(ctx->ev_check);
  for (auto& b : ctx->d_fb) for (auto p : b) cudaFree(p);
  cudaFree(ctx->d_lat); cudaFree(ctx->d_lon);
  for (auto e : ctx->ev_fb) if (e) cudaEventDestroy(e);
  for (auto p : ctx->d_wt) cudaFree(p);
  cudaFree(ctx->d_kcell); cudaFree(ctx->d_head);
  if (ctx->h_errkey) cudaFreeHost(ctx->h_errkey);
  if (ctx->h_errcount) cudaFreeHost(ctx->h_errcount);
  for (auto e : ctx->ev_in) cudaEventDestroy(e);
  for (auto e : ctx->ev_k) cudaEventDestroy(e);
  for (auto e : ctx->ev_out) cudaEventDestroy(e);
  if (ctx->ev_outready) cudaEventDestroy(ctx->ev_outready);
  if (ctx->ev_outdone) cudaEventDestroy(ctx->ev_outdone);
  if (ctx->d_outstage) cudaFree(ctx->d_outstage);
  for (auto e : ctx->ev_plane) if (e) cudaEventDestroy(e);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_rebin) cudaEventDestroy(ctx->ev_rebin);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

int noahmp_b200_set_mode(noahmp_b200_ctx* ctx, int sync_mode) {
  if (!ctx || (sync_mode != NOAHMP_SYNC_FULL && sync_mode != NOAHMP_SYNC_RESIDENT)) return NOAHMP_ERR_ARG;
  ctx->sync_mode = sync_mode;
  return 0;
}

int noahmp_b200_set_math(noahmp_b200_ctx* ctx, int math_mode) {
  if (!ctx || (math_mode != NOAHMP_MATH_FAST && math_mode != NOAHMP_MATH_PARITY)) return NOAHMP_ERR_ARG;
  ctx->math_mode = math_mode;
  return 0;
}

const char* noahmp_b200_kernel_variant(const noahmp_b200_ctx* ctx) { return ctx ? ctx->variant.c_str() : ""; }

int noahmp_b200_upload(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  if (!ctx || !a) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = check_bounds(ctx, a);
  if (rc) return rc;
  fill_scalars(ctx, a);
  if ((rc = check_options(ctx))) return rc;
  if ((rc = build_plane_descs(ctx))) return rc;
  if ((rc = upload_static(ctx, a))) return rc;
  if ((rc = upload_forcing(ctx, a))) return rc;
  if ((rc = classify(ctx))) return rc;
  // OUT arrays travel up as well when the tile has open-water cells: those cells have no column, and what comes
  // back for them must be what the caller holds now, not what it held at the first call
  if ((rc = upload_state(ctx, a, /*all=*/!ctx->uploaded || ctx->nclass[CL_WATER] > 0))) return rc;
  if ((rc = gather_fields(ctx))) return rc;
  ctx->base.state = ctx->d_state;
  ctx->base.cell = ctx->d_cell;
  ctx->base.np = ctx->np;
  ctx->base.np4 = (unsigned)(ctx->np * 4);
  ctx->uploaded = true;
  for (int f = 0; f < NFORC; ++f) { ctx->forc_valid[f] = true; ctx->base.forc[f] = ctx->d_forc[f]; }
  CK(cudaMemsetAsync(ctx->d_errkey, 0xff, sizeof(unsigned long long), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_errcount, 0, sizeof(int), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int noahmp_b200_device_forcing(noahmp_b200_ctx* ctx, float** dev_ptrs) {
  if (!ctx || !dev_ptrs) return NOAHMP_ERR_ARG;
  for (int f = 0; f < NFORC; ++f) dev_ptrs[f] = ctx->d_forc[f];
  return 0;
}

int noahmp_b200_enable_iteration_counts(noahmp_b200_ctx* ctx, int enable) {
  if (!ctx) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (enable && !ctx->d_vege_iters) {
    CK(cudaMalloc(&ctx->d_vege_iters, sizeof(int) * ctx->ncell));
    CK(cudaMemset(ctx->d_vege_iters, 0, sizeof(int) * ctx->ncell));
  }
  ctx->base.vege_iters = enable ? ctx->d_vege_iters : nullptr;
  return 0;
}
int noahmp_b200_get_iteration_counts(noahmp_b200_ctx* ctx, int32_t* out) {
  if (!ctx || !out || !ctx->d_vege_iters) return NOAHMP_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, ctx->d_vege_iters, sizeof(int) * ctx->ncell, cudaMemcpyDeviceToHost));
  return 0;
}

// `clear_latch`: the per-call entry points report the status of their own step; a device-side loop of step_device
// calls keeps the first failure (smallest key) and the total count until noahmp_b200_get_status reads them.
static int step_device_impl(noahmp_b200_ctx* ctx, int itimestep, int yr, float julian, float dt, void* stream,
                            bool clear_latch) {
  if (!ctx || !ctx->uploaded) { set_error("step_device before upload"); return NOAHMP_ERR_ARG; }
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  // a snapshot taken by output_begin on the library's stream must be complete before a caller stream changes the state
  if (ctx->out_pending && s != ctx->stream) CK(cudaStreamWaitEvent(s, ctx->ev_outready, 0));
  if (ctx->last_step_stream && ctx->last_step_stream != s) {  // the stream changed between steps: keep them ordered
    if (!ctx->ev_rebin) CK(cudaEventCreateWithFlags(&ctx->ev_rebin, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ev_rebin, ctx->last_step_stream));
    CK(cudaStreamWaitEvent(s, ctx->ev_rebin, 0));
  }
  if (ctx->sync_mode == NOAHMP_SYNC_RESIDENT && ctx->rebin_interval > 0 && ctx->nclass[CL_LAND] > 0) {
    if (itimestep > 1 && (!ctx->binned ? ctx->steps_since_rebin >= 2 : ctx->steps_since_rebin >= ctx->rebin_interval)) {
      int rc = chunk_ranges(ctx, ctx->binned ? ctx->bin_chunks : auto_chunks(ctx), nullptr);
      if (rc) return rc;
      if ((rc = rebin(ctx, s))) return rc;
    }
    ctx->steps_since_rebin++;
  }
  ctx->last_step_stream = s;
  StepParams p = ctx->base;
  p.itimestep = itimestep;
  p.yearlen = year_length(yr);
  p.julian = julian;
  p.dt = dt;
  if (clear_latch) {
    CK(cudaMemsetAsync(ctx->d_errkey, 0xff, sizeof(unsigned long long), s));
    CK(cudaMemsetAsync(ctx->d_errcount, 0, sizeof(int), s));
  }
  if (itimestep == 1 && ctx->nclass[CL_WATER] > 0) {
    const int T = 256;
    first_step_water_kernel<<<(unsigned)((ctx->ncell + T - 1) / T), T, 0, s>>>(
        ctx->d_stat[ST_XLAND], ctx->d_stat[ST_XICE], ctx->d_grid[F_smstav], ctx->d_grid[F_smstot],
        ctx->d_grid[F_smois], ctx->d_grid[F_tslb], ctx->ni, ctx->ncell);
    ctx->launches++;
  }
  const int nland = ctx->nclass[CL_LAND], nglac = ctx->nclass[CL_GLACIER], nsea = ctx->nclass[CL_SEAICE];
  const StepRange r{0, nland, nland, nglac};
  const char* v = ctx->math_mode == NOAHMP_MATH_PARITY ? nmp_launch_step_parity(p, r, s, &ctx->launches)
                                                       : nmp_launch_step_fast(p, r, s, &ctx->launches);
  ctx->variant = v;
  if (nsea > 0) {
    seaice_kernel<<<(nsea + 255) / 256, 256, 0, s>>>(ctx->d_state, ctx->d_cell, ctx->d_stat[ST_XICE], ctx->np, nland + nglac,
                                                     nsea, itimestep);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  if (ctx->budget_on) return budget_accumulate(ctx, dt, s);
  return 0;
}

int noahmp_b200_step_device(noahmp_b200_ctx* ctx, int itimestep, int yr, float julian, float dt, void* stream) {
  return step_device_impl(ctx, itimestep, yr, julian, dt, stream, /*clear_latch=*/false);
}

int noahmp_b200_get_status(noahmp_b200_ctx* ctx, noahmp_status* status) {
  if (!ctx) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(ctx->h_errkey, ctx->d_errkey, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ctx->h_errcount, ctx->d_errcount, sizeof(int), cudaMemcpyDeviceToHost));
  decode_status(ctx, status);
  // the latch has been read: start a new accumulation period for the device-side steps that follow
  CK(cudaMemset(ctx->d_errkey, 0xff, sizeof(unsigned long long)));
  CK(cudaMemset(ctx->d_errcount, 0, sizeof(int)));
  return status ? status->code : 0;
}

int noahmp_b200_sync_host(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a) {
  if (!ctx || !a || !ctx->uploaded) return NOAHMP_ERR_ARG;
  ++ctx->pin_clock;
  CK(cudaSetDevice(ctx->device));
  int rc = check_bounds(ctx, a);
  if (rc) return rc;
  CK(cudaDeviceSynchronize());  // steps may have been issued on a caller stream
  if ((rc = scatter_fields(ctx))) return rc;
  if ((rc = download_state(ctx, a))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- column re-binning -------------------------------------------------------------------------------------------
// Row-chunk boundaries in the compact order.  Columns of a class are classified in grid order, so a row chunk is a
// contiguous compact range per class; re-binning permutes land columns only INSIDE their chunk, which keeps these
// ranges (and with them the upload | physics | download pipeline) valid.
// First row of chunk c.  The first and the last chunk are half as tall as the others: the physics starts after a
// short first upload and the step ends with a short last kernel + download (the pipeline's fill and drain).
static int chunk_row(const noahmp_b200_ctx* ctx, int c, int nchunks) {
  if (nchunks < 4) return (int)((long long)ctx->nj * c / nchunks);
  const int units = 2 * nchunks - 2;
  const int u = c <= 0 ? 0 : (c >= nchunks ? units : 2 * c - 1);
  return (int)((long long)ctx->nj * u / units);
}
// Automatic number of row chunks of the RESIDENT-mode pipeline: 9 (the first and the last half as tall as the others)
// from 2^21 cells, 5 from 2^20, one below.  What a chunk costs is ~20 asynchronous API calls (0.04 ms of host time,
// enqueued ahead of the GPU); what it buys is a shorter fill and drain of the upload | physics | download pipeline: the
// call ends one chunk's kernel + download after the last upload (profiles/r02_notes.md: 3 equal chunks on a 4.4 M-cell
// tile end 1.5 ms after the last byte went up, 9 chunks 0.3 ms).
static int auto_chunks(const noahmp_b200_ctx* ctx) {
  if (ctx->nchunks) return std::min(ctx->nchunks, ctx->nj);
  if (ctx->ncell < (1LL << 20)) return 1;
  return std::min(ctx->ncell >= (1LL << 21) ? 9 : 5, ctx->nj);
}
struct ChunkRanges {
  int n = 0;
  std::vector<int> row, land, glac, sea;  // size n + 1
};
// Compact ranges of `nchunks` row chunks.  Once the land columns have been re-binned they stay inside the row chunk
// they were binned in, so only that chunking (ctx->bin_chunks) or a single chunk are valid afterwards: any other
// request falls back to the binned chunking.
static int chunk_ranges(noahmp_b200_ctx* ctx, int nchunks, ChunkRanges* out) {
  const int nland = ctx->nclass[CL_LAND], nglac = ctx->nclass[CL_GLACIER], nsea = ctx->nclass[CL_SEAICE];
  if (ctx->binned && nchunks != ctx->bin_chunks && nchunks != 1) nchunks = ctx->bin_chunks;
  const bool stored = ctx->bin_chunks == nchunks && (int)ctx->ch_land.size() == nchunks + 1;
  if (!stored && !(ctx->binned && nchunks == 1)) {
    const int* cl = ctx->h_cell.data();
    auto lower = [&](int lo, int hi, int cell) { return (int)(std::lower_bound(cl + lo, cl + hi, cell) - cl); };
    ctx->ch_land.assign(nchunks + 1, 0); ctx->ch_glac.assign(nchunks + 1, 0); ctx->ch_sea.assign(nchunks + 1, 0);
    for (int c = 0; c <= nchunks; ++c) {
      const int j = chunk_row(ctx, c, nchunks);
      const int cell = j * ctx->ni;
      ctx->ch_land[c] = c == nchunks ? nland : lower(0, nland, cell);
      ctx->ch_glac[c] = c == nchunks ? nland + nglac : lower(nland, nland + nglac, cell);
      ctx->ch_sea[c] = c == nchunks ? nland + nglac + nsea : lower(nland + nglac, nland + nglac + nsea, cell);
    }
    ctx->bin_chunks = nchunks;
  }
  if (out) {
    out->n = nchunks;
    if (ctx->binned && nchunks == 1 && ctx->bin_chunks != 1) {
      out->row = {0, ctx->nj};
      out->land = {0, nland}; out->glac = {nland, nland + nglac}; out->sea = {nland + nglac, nland + nglac + nsea};
    } else {
      out->row.resize(nchunks + 1);
      for (int c = 0; c <= nchunks; ++c) out->row[c] = chunk_row(ctx, c, nchunks);
      out->land = ctx->ch_land; out->glac = ctx->ch_glac; out->sea = ctx->ch_sea;
    }
  }
  return 0;
}

// chunk_first[0..nchunks]: first compact column of every row chunk; chunk_first[80 + c]: its first row.  With nsub > 1
// every chunk is cut into nsub row groups that are binned separately: a bin then sweeps a band of the grid-order planes
// (forcing, groundwater inputs) small enough to stay in L2 until the next bin of the same band reads it again.
__global__ void bin_key_kernel(const float* __restrict__ state, long long np, int nland, const int* __restrict__ chunk_first,
                               int nchunks, int nsub, const int* __restrict__ cell, int ni, int* __restrict__ keys,
                               int* __restrict__ iota) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nland) return;
  int c = 0;
  while (c + 1 < nchunks && n >= chunk_first[c + 1]) ++c;
  if (nsub > 1) {
    const int r0 = chunk_first[80 + c], r1 = chunk_first[80 + c + 1];
    const int j = cell[n] / ni;
    c = c * nsub + min(nsub - 1, (int)((long long)(j - r0) * nsub / max(r1 - r0, 1)));
  }
  const int isnow = __float_as_int(state[(long long)NMP_SLOT(isnowxy) * np + n]);
  const int prev = __float_as_int(state[(long long)PLANE_PREV_ITERS * np + n]);
  // The number of canopy Newton passes itself is not a useful key: it is unpredictable from one step to the next
  // (tools/binning_study.py), and finer bins only scatter the column -> cell map that the forcing is read through.
  const int pb = prev == 0 ? 0 : 1;
  keys[n] = c * 32 + pb * 4 + min(max(-isnow, 0), 3);
  iota[n] = n;
}
// number of places where the bin key decreases along the compact order (0 = the columns are still sorted)
__global__ void key_order_kernel(const int* __restrict__ keys, int nland, int* __restrict__ viol) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  int v = (n > 0 && n < nland && keys[n] < keys[n - 1]) ? 1 : 0;
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(viol, v);
}
// scratch[g][i] = state[planes[g]][perm[i]] for the land columns i < nland of one group of planes.  One thread moves
// its column in PERMUTE_GROUP planes: the permutation index is read once per group and the group's gathers are in
// flight together.  The permuted group is then copied back over the planes it came from (a dense device-to-device
// copy), so the re-binning needs PERMUTE_GROUP scratch planes instead of a second copy of the whole state.
// OUT planes are rewritten for every land column by the step that follows and are not moved; the glacier / sea-ice
// tail of every plane stays where it is.
constexpr int PERMUTE_GROUP = 8;
struct PermuteGroup { int plane[PERMUTE_GROUP]; int n; };
__global__ void permute_state_kernel(const float* __restrict__ state, float* __restrict__ scratch,
                                     const int* __restrict__ perm, PermuteGroup grp, long long np, int nland) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nland) return;
  const long long j = (long long)perm[i];
  float v[PERMUTE_GROUP];
#pragma unroll
  for (int g = 0; g < PERMUTE_GROUP; ++g)
    if (g < grp.n) v[g] = state[(long long)grp.plane[g] * np + j];
#pragma unroll
  for (int g = 0; g < PERMUTE_GROUP; ++g)
    if (g < grp.n) scratch[(long long)g * nland + i] = v[g];
}
__global__ void permute_copyback_kernel(const float* __restrict__ scratch, float* __restrict__ state, PermuteGroup grp,
                                        long long np, int nland) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nland) return;
#pragma unroll
  for (int g = 0; g < PERMUTE_GROUP; ++g)
    if (g < grp.n) state[(long long)grp.plane[g] * np + i] = scratch[(long long)g * nland + i];
}
__global__ void permute_cell_kernel(const int* __restrict__ src, int* __restrict__ dst, const int* __restrict__ perm,
                                    long long np, int nland) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  dst[i] = src[i < nland ? (long long)perm[i] : i];
}

// Runs entirely in the order of stream `s` (the stream the steps run on): no device-wide synchronisation, the host
// only swaps the two column-map buffers.
static int rebin(noahmp_b200_ctx* ctx, cudaStream_t s) {
  const int nland = ctx->nclass[CL_LAND];
  const long long np = ctx->np;
  const int nch = ctx->bin_chunks;
  if (nland == 0 || nch == 0) return 0;
  if (ctx->last_step_stream && ctx->last_step_stream != s) {  // steps in flight on another stream: order after them
    if (!ctx->ev_rebin) CK(cudaEventCreateWithFlags(&ctx->ev_rebin, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ev_rebin, ctx->last_step_stream));
    CK(cudaStreamWaitEvent(s, ctx->ev_rebin, 0));
  }
  if (!ctx->d_state2) {
    CK(cudaMalloc(&ctx->d_state2, sizeof(float) * (size_t)PERMUTE_GROUP * (size_t)nland));
    CK(cudaMalloc(&ctx->d_cell2, sizeof(int) * ctx->ncell));
    CK(cudaMalloc(&ctx->d_keys, sizeof(int) * ctx->ncell));
    CK(cudaMalloc(&ctx->d_keys2, sizeof(int) * ctx->ncell));
    CK(cudaMalloc(&ctx->d_perm, sizeof(int) * ctx->ncell));
    CK(cudaMalloc(&ctx->d_iota, sizeof(int) * ctx->ncell));
    CK(cudaMalloc(&ctx->d_chunk, sizeof(int) * 160));
    ctx->moved_planes.clear();
    std::vector<unsigned char> kind(NPLANES_ALLOC, 0);
    for (int f = 0; f < NFIELDS; ++f)
      for (int k = 0; k < kFields[f].layers; ++k) kind[kSlots.slot[f] + k] = (unsigned char)kFields[f].kind;
    for (int pl = 0; pl < NPLANES_ALLOC; ++pl)
      if (kind[pl] == NMP_K_INOUT && pl < PLANE_HANDOFF0) ctx->moved_planes.push_back(pl);
    size_t need = 0;
    int bits = 5;
    while ((1 << (bits - 5)) < 64 * 8) ++bits;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->d_keys, ctx->d_keys2, ctx->d_iota, ctx->d_perm, nland, 0, bits, s));
    if (need > ctx->cub_bytes) {
      if (ctx->d_cub) CK(cudaFree(ctx->d_cub));
      CK(cudaMalloc(&ctx->d_cub, need));
      ctx->cub_bytes = need;
    }
  }
  const int T = 256;
  auto compute_keys = [&]() -> int {
    ctx->h_chunk.assign(160, 0);
    for (int c = 0; c <= nch && c < 80; ++c) {
      ctx->h_chunk[c] = ctx->ch_land[c];
      ctx->h_chunk[80 + c] = chunk_row(ctx, c, nch);
    }
    CK(cudaMemcpyAsync(ctx->d_chunk, ctx->h_chunk.data(), sizeof(int) * 160, cudaMemcpyHostToDevice, s));
    bin_key_kernel<<<(nland + T - 1) / T, T, 0, s>>>(ctx->d_state, np, nland, ctx->d_chunk, nch, ctx->bin_sub,
                                                     ctx->d_cell, ctx->ni, ctx->d_keys, ctx->d_iota);
    ctx->launches++;
    return 0;
  };
  if (ctx->binned && ctx->rebin_min_changed > 0.f) {
    if (!ctx->check_pending) {
      if (!ctx->ev_check) {
        CK(cudaEventCreateWithFlags(&ctx->ev_check, cudaEventDisableTiming));
        CK(cudaMalloc(&ctx->d_viol, sizeof(int)));
        CK(cudaMallocHost((void**)&ctx->h_viol, sizeof(int)));
      }
      int rc = compute_keys();
      if (rc) return rc;
      CK(cudaMemsetAsync(ctx->d_viol, 0, sizeof(int), s));
      key_order_kernel<<<(nland + T - 1) / T, T, 0, s>>>(ctx->d_keys, nland, ctx->d_viol);
      ctx->launches++;
      CK(cudaMemcpyAsync(ctx->h_viol, ctx->d_viol, sizeof(int), cudaMemcpyDeviceToHost, s));
      CK(cudaEventRecord(ctx->ev_check, s));
      ctx->check_pending = true;
      ctx->rebin_checks++;
      return 0;  // the answer is looked at when a later step comes by (no waiting here)
    }
    const cudaError_t q = cudaEventQuery(ctx->ev_check);
    if (q == cudaErrorNotReady) return 0;
    CK(q);
    ctx->check_pending = false;
    if ((double)*ctx->h_viol <= 2.0 * (double)ctx->rebin_min_changed * (double)nland) {
      ctx->steps_since_rebin = 0;
      ctx->rebin_skips++;
      if (ctx->trace) fprintf(stderr, "[noahmp_b200 trace] re-binning skipped: %d order breaks in %d land columns\n",
                              *ctx->h_viol, nland);
      return 0;
    }
  }
  {
    int rc = compute_keys();
    if (rc) return rc;
  }
  int bits = 5;
  while ((1 << (bits - 5)) < nch * ctx->bin_sub) ++bits;
  size_t need = ctx->cub_bytes;
  // stable LSD radix sort: columns of equal key keep their current relative order
  CK(cub::DeviceRadixSort::SortPairs(ctx->d_cub, need, ctx->d_keys, ctx->d_keys2, ctx->d_iota, ctx->d_perm, nland, 0, bits, s));
  const unsigned nb = (unsigned)((nland + T - 1) / T);
  for (size_t g0 = 0; g0 < ctx->moved_planes.size(); g0 += PERMUTE_GROUP) {
    PermuteGroup grp;
    grp.n = (int)std::min((size_t)PERMUTE_GROUP, ctx->moved_planes.size() - g0);
    for (int g = 0; g < PERMUTE_GROUP; ++g) grp.plane[g] = g < grp.n ? ctx->moved_planes[g0 + g] : 0;
    permute_state_kernel<<<nb, T, 0, s>>>(ctx->d_state, ctx->d_state2, ctx->d_perm, grp, np, nland);
    permute_copyback_kernel<<<nb, T, 0, s>>>(ctx->d_state2, ctx->d_state, grp, np, nland);
    ctx->launches += 2;
  }
  permute_cell_kernel<<<(unsigned)((np + T - 1) / T), T, 0, s>>>(ctx->d_cell, ctx->d_cell2, ctx->d_perm, np, nland);
  ctx->launches++;
  CK(cudaGetLastError());
  std::swap(ctx->d_cell, ctx->d_cell2);
  ctx->base.cell = ctx->d_cell;
  ctx->binned = true;
  ctx->steps_since_rebin = 0;
  ctx->rebins++;
  return 0;
}

// gather of compact range [first, first+count) of one field (all its layers) from the grid-order staging
__global__ void gather_range_kernel(const PlaneDesc* __restrict__ planes, const int* __restrict__ cell, float* state,
                                    long long np, int ni, long long first, long long count) {
  long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= count) return;
  n += first;
  const PlaneDesc d = planes[blockIdx.y];
  const int c = cell[n];
  const int i = c % ni, j = c / ni;
  state[(long long)d.plane * np + n] = d.grid[(long long)i + (long long)d.layer * ni + (long long)j * ni * d.layers];
}

// RESIDENT-mode step as a row-chunk pipeline: while chunk c is computed, the forcing rows of chunk c+1 are on
// their way up and the requested result fields of chunk c-1 on their way down (three streams, events in between).
// Columns of a class are stored in grid order, so a row chunk is one contiguous compact range per class.
static int step_resident_pipelined(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, int nchunks, bool upload = true) {
  const int nk = a->kme - a->kms + 1, kms = a->kms;
  const auto t_begin = std::chrono::steady_clock::now();
  if (!ctx->s_in) {
    CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  }
  if (!ctx->ev_t0) {
    ctx->trace = getenv("NOAHMP_B200_TRACE") != nullptr;
    CK(cudaEventCreate(&ctx->ev_t0));
  }
  const unsigned evflags = ctx->trace ? cudaEventDefault : cudaEventDisableTiming;
  while ((int)ctx->ev_in.size() < std::max(nchunks, 1)) {
    cudaEvent_t e1, e2, e3;
    CK(cudaEventCreateWithFlags(&e1, evflags));
    CK(cudaEventCreateWithFlags(&e2, evflags));
    CK(cudaEventCreateWithFlags(&e3, evflags));
    ctx->ev_in.push_back(e1);
    ctx->ev_k.push_back(e2);
    ctx->ev_out.push_back(e3);
  }
  if (ctx->trace) CK(cudaEventRecord(ctx->ev_t0, ctx->s_in));
  struct Plane { int id; const float* src; int nk, lev; };
  const Plane all_planes[NFORC] = {
      {FC_COSZIN, a->coszin, 1, 1}, {FC_T, a->t3d, nk, 1},       {FC_QV, a->qv3d, nk, 1},      {FC_U, a->u_phy, nk, 1},
      {FC_V, a->v_phy, nk, 1},      {FC_SWDOWN, a->swdown, 1, 1}, {FC_GLW, a->glw, 1, 1},      {FC_P1, a->p8w3d, nk, a->kts},
      {FC_P2, a->p8w3d, nk, a->kts + 1}, {FC_RAINBL, a->rainbl, 1, 1}, {FC_VEGFRA, a->vegfra, 1, 1}, {FC_DZ8W, a->dz8w, nk, 1}};
  // Forcing planes the caller declared unchanged (noahmp_b200_set_forcing_hints) are not sent again once the device
  // holds a copy; with P8W_LEVELS_EQUAL level 2 of P8W3D is read from the level-1 plane.
  Plane planes[NFORC];
  int nplanes = 0;
  if (upload) {
    for (const Plane& pl : all_planes) {
      bool skip = false;
      if (pl.id == FC_DZ8W && (ctx->hints & NOAHMP_HINT_DZ8W_CONSTANT) && ctx->forc_valid[FC_DZ8W]) skip = true;
      if (pl.id == FC_VEGFRA && (ctx->hints & NOAHMP_HINT_VEGFRA_UNCHANGED) && ctx->forc_valid[FC_VEGFRA]) skip = true;
      if (pl.id == FC_P2 && (ctx->hints & NOAHMP_HINT_P8W_LEVELS_EQUAL)) skip = true;
      if (!skip) planes[nplanes++] = pl;
    }
    for (int k = 0; k < nplanes; ++k) pin(ctx, planes[k].src, host_bytes(ctx, planes[k].nk));
  }
  for (int f : ctx->fetch) pin(ctx, host_ptr(a, f), host_bytes(ctx, kFields[f].layers));
  for (int f : ctx->push) {
    if (!host_ptr(a, f)) { set_error(std::string("null array in the push list: ") + kFields[f].name); return NOAHMP_ERR_ARG; }
    pin(ctx, host_ptr(a, f), host_bytes(ctx, kFields[f].layers));
  }

  // host forcing supersedes device pointers bound earlier with bind_forcing()
  if (upload) {
    for (int f = 0; f < NFORC; ++f) ctx->base.forc[f] = ctx->d_forc[f];
    if (ctx->hints & NOAHMP_HINT_P8W_LEVELS_EQUAL) ctx->base.forc[FC_P2] = ctx->d_forc[FC_P1];
  }
  StepParams p = ctx->base;
  p.itimestep = a->itimestep;
  p.yearlen = year_length(a->yr);
  p.julian = a->julian;
  p.dt = a->dt;
  cudaStream_t sk = ctx->stream;
  if (ctx->last_step_stream && ctx->last_step_stream != sk) {  // device-side steps issued earlier on a caller stream
    if (!ctx->ev_rebin) CK(cudaEventCreateWithFlags(&ctx->ev_rebin, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ev_rebin, ctx->last_step_stream));
    CK(cudaStreamWaitEvent(sk, ctx->ev_rebin, 0));
  }
  ctx->last_step_stream = sk;
  CK(cudaMemsetAsync(ctx->d_errkey, 0xff, sizeof(unsigned long long), sk));
  CK(cudaMemsetAsync(ctx->d_errcount, 0, sizeof(int), sk));
  if (a->itimestep == 1 && ctx->nclass[CL_WATER] > 0) {
    const int T = 256;
    first_step_water_kernel<<<(unsigned)((ctx->ncell + T - 1) / T), T, 0, sk>>>(
        ctx->d_stat[ST_XLAND], ctx->d_stat[ST_XICE], ctx->d_grid[F_smstav], ctx->d_grid[F_smstot], ctx->d_grid[F_smois],
        ctx->d_grid[F_tslb], ctx->ni, ctx->ncell);
    ctx->launches++;
  }
  const int nland = ctx->nclass[CL_LAND];
  int rc0;
  if (ctx->rebin_interval > 0 && nland > 0 && a->itimestep > 1 &&
      (!ctx->binned ? ctx->steps_since_rebin >= 2 : ctx->steps_since_rebin >= ctx->rebin_interval)) {
    if ((rc0 = chunk_ranges(ctx, ctx->binned ? ctx->bin_chunks : auto_chunks(ctx), nullptr))) return rc0;
    if ((rc0 = rebin(ctx, sk))) return rc0;
    p.state = ctx->d_state;
    p.cell = ctx->d_cell;
  }
  ChunkRanges R;
  if ((rc0 = chunk_ranges(ctx, nchunks, &R))) return rc0;
  nchunks = R.n;
  ctx->steps_since_rebin++;
  const int T = 256;
  for (int c = 0; c < nchunks; ++c) {
    const int j0 = R.row[c], j1 = R.row[c + 1];
    if (j1 <= j0) continue;
    StepRange r;
    r.land_first = R.land[c];
    r.land_count = R.land[c + 1] - r.land_first;
    r.glac_first = R.glac[c];
    r.glac_count = R.glac[c + 1] - r.glac_first;
    const int s0 = R.sea[c], s1 = R.sea[c + 1];
    const int rng[3][2] = {{r.land_first, r.land_count}, {r.glac_first, r.glac_count}, {s0, s1 - s0}};
    bool waited = false;
    if (upload && nplanes > 0) {
      for (int k = 0; k < nplanes; ++k) {
        const Plane& pl = planes[k];
        int rc = h2d_rows(ctx, ctx->d_forc[pl.id], pl.src, pl.nk, kms, pl.lev, j0, j1, ctx->s_in);
        if (rc) return rc;
        if (ctx->trace && c == 0) {
          if ((int)ctx->ev_plane.size() <= pl.id) ctx->ev_plane.resize(NFORC, nullptr);
          if (!ctx->ev_plane[pl.id]) CK(cudaEventCreate(&ctx->ev_plane[pl.id]));
          CK(cudaEventRecord(ctx->ev_plane[pl.id], ctx->s_in));
        }
      }
      waited = true;
    }
    // INOUT arrays the driver rewrites before every call (e.g. XLAIXY from the forcing file): rows up, then into
    // the compact planes of this chunk's columns
    for (int f : ctx->push) {
      CK(copy_field_rows(ctx, ctx->d_grid[f], host_ptr(a, f), kFields[f].layers, j0, j1, true, ctx->s_in));
      waited = true;
    }
    if (waited) {
      CK(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
      CK(cudaStreamWaitEvent(sk, ctx->ev_in[c], 0));
    }
    for (int f : ctx->push) {
      for (auto& q : rng) {
        if (q[1] <= 0) continue;
        dim3 grid((unsigned)((q[1] + T - 1) / T), kFields[f].layers);
        gather_range_kernel<<<grid, T, 0, sk>>>(ctx->d_planes + kSlots.slot[f], p.cell, ctx->d_state, ctx->np, ctx->ni,
                                                (long long)q[0], (long long)q[1]);
        ctx->launches++;
      }
    }
    const char* v = ctx->math_mode == NOAHMP_MATH_PARITY ? nmp_launch_step_parity(p, r, sk, &ctx->launches)
                                                         : nmp_launch_step_fast(p, r, sk, &ctx->launches);
    ctx->variant = v;
    if (s1 > s0) {
      seaice_kernel<<<(s1 - s0 + 255) / 256, 256, 0, sk>>>(ctx->d_state, p.cell, ctx->d_stat[ST_XICE], ctx->np, s0,
                                                         s1 - s0, a->itimestep);
      ctx->launches++;
    }
    if (!ctx->fetch.empty()) {
      // scatter the requested fields of this chunk's columns into the grid-order staging (one launch per class
      // range for the whole fetch list), then send the rows down
      for (auto& q : rng) {
        if (q[1] <= 0) continue;
        dim3 grid((unsigned)((q[1] + T - 1) / T), ctx->n_fetch_planes);
        scatter_kernel<<<grid, T, 0, sk>>>(ctx->d_fetch_planes, p.cell, ctx->d_state, ctx->np, ctx->ni, (long long)q[0],
                                            (long long)q[1]);
        ctx->launches++;
      }
      CK(cudaEventRecord(ctx->ev_k[c], sk));
      CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[c], 0));
      for (int f : ctx->fetch) {
        CK(copy_field_rows(ctx, ctx->d_grid[f], host_ptr(a, f), kFields[f].layers, j0, j1, false, ctx->s_out));
      }
      if (ctx->trace) CK(cudaEventRecord(ctx->ev_out[c], ctx->s_out));
    } else if (ctx->trace) {
      CK(cudaEventRecord(ctx->ev_k[c], sk));
    }
  }
  CK(cudaGetLastError());
  if (ctx->budget_on && (rc0 = budget_accumulate(ctx, a->dt, sk))) return rc0;
  if (upload)
    for (int k = 0; k < nplanes; ++k) ctx->forc_valid[planes[k].id] = true;
  CK(cudaMemcpyAsync(ctx->h_errkey, ctx->d_errkey, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sk));
  CK(cudaMemcpyAsync(ctx->h_errcount, ctx->d_errcount, sizeof(int), cudaMemcpyDeviceToHost, sk));
  const auto t_enq = std::chrono::steady_clock::now();
  CK(cudaStreamSynchronize(sk));
  if (!ctx->fetch.empty()) CK(cudaStreamSynchronize(ctx->s_out));
  if (ctx->trace) {
    fprintf(stderr, "[noahmp_b200 trace] host: enqueue %.2f ms, total %.2f ms (%d chunks, %d forcing planes up)\n",
            std::chrono::duration<double, std::milli>(t_enq - t_begin).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), nchunks,
            upload ? nplanes : 0);
    // completion times (ms after the call's first enqueue): forcing rows up | physics (+ scatter) | results down
    fprintf(stderr, "[noahmp_b200 trace] step %d:", a->itimestep);
    for (int c = 0; c < nchunks; ++c) {
      float t[3] = {-1.f, -1.f, -1.f};
      if (upload && nplanes > 0) cudaEventElapsedTime(&t[0], ctx->ev_t0, ctx->ev_in[c]);
      cudaEventElapsedTime(&t[1], ctx->ev_t0, ctx->ev_k[c]);
      if (!ctx->fetch.empty()) cudaEventElapsedTime(&t[2], ctx->ev_t0, ctx->ev_out[c]);
      fprintf(stderr, " [%d] %.2f|%.2f|%.2f", c, t[0], t[1], t[2]);
    }
    fprintf(stderr, "\n");
    cudaGetLastError();
  }
  return 0;
}

// Fields (comma separated noahmp_lsm_args member names, "" = none) that every RESIDENT-mode noahmplsm call
// refreshes in the caller's host arrays, e.g. "tsk,hfx,lh,grdflx"; everything else waits for sync_host().
static int parse_fields(const char* fields, std::vector<int>& list) {
  std::string s(fields), tok;
  size_t pos = 0;
  while (pos <= s.size()) {
    size_t e = s.find(',', pos);
    if (e == std::string::npos) e = s.size();
    tok = s.substr(pos, e - pos);
    pos = e + 1;
    while (!tok.empty() && tok.front() == ' ') tok.erase(0, 1);
    while (!tok.empty() && tok.back() == ' ') tok.pop_back();
    if (tok.empty()) continue;
    int f = 0;
    for (; f < NFIELDS; ++f)
      if (tok == kFields[f].name) break;
    if (f == NFIELDS) { set_error("unknown field " + tok); return NOAHMP_ERR_ARG; }
    list.push_back(f);
  }
  return 0;
}
int noahmp_b200_set_fetch(noahmp_b200_ctx* ctx, const char* fields) {
  if (!ctx || !fields) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<int> list;
  int rc = parse_fields(fields, list);
  if (rc) return rc;
  if ((rc = build_plane_descs(ctx))) return rc;
  std::vector<PlaneDesc> h;
  for (int f : list)
    for (int k = 0; k < kFields[f].layers; ++k) h.push_back({ctx->d_grid[f], kSlots.slot[f] + k, k, kFields[f].layers});
  CK(cudaStreamSynchronize(ctx->stream));  // a call in flight may still read the old descriptor list
  if (ctx->d_fetch_planes) { CK(cudaFree(ctx->d_fetch_planes)); ctx->d_fetch_planes = nullptr; }
  if (!h.empty()) {
    CK(cudaMalloc(&ctx->d_fetch_planes, sizeof(PlaneDesc) * h.size()));
    CK(cudaMemcpy(ctx->d_fetch_planes, h.data(), sizeof(PlaneDesc) * h.size(), cudaMemcpyHostToDevice));
  }
  ctx->n_fetch_planes = (int)h.size();
  ctx->fetch = list;
  return 0;
}

// RESIDENT mode: INOUT fields (comma separated member names, "" = none) whose HOST content every noahmplsm call
// takes again before the step.  The HRLDAS driver overwrites LAI (= XLAIXY) from the forcing file before every call
// (module_hrldas_noahmp_driver.F90:335, driver/module_hrldas_netcdf_io.F90:1365/1402), so a resident drop-in must
// list "xlaixy" here to follow the reference when that array is rewritten between calls.
int noahmp_b200_set_push(noahmp_b200_ctx* ctx, const char* fields) {
  if (!ctx || !fields) return NOAHMP_ERR_ARG;
  std::vector<int> list;
  int rc = parse_fields(fields, list);
  if (rc) return rc;
  for (int f : list)
    if (kFields[f].kind != NMP_K_INOUT) { set_error(std::string("set_push: not an INOUT array: ") + kFields[f].name); return NOAHMP_ERR_ARG; }
  ctx->push = list;
  return 0;
}

// Forcing planes that need no upload in the following RESIDENT-mode calls (NOAHMP_HINT_* bits, see the header).
int noahmp_b200_set_forcing_hints(noahmp_b200_ctx* ctx, unsigned hints) {
  if (!ctx || (hints & ~(unsigned)(NOAHMP_HINT_DZ8W_CONSTANT | NOAHMP_HINT_VEGFRA_UNCHANGED | NOAHMP_HINT_P8W_LEVELS_EQUAL)))
    return NOAHMP_ERR_ARG;
  // a hint that is withdrawn makes the next call upload the plane again
  if (!(hints & NOAHMP_HINT_DZ8W_CONSTANT)) ctx->forc_valid[FC_DZ8W] = false;
  if (!(hints & NOAHMP_HINT_VEGFRA_UNCHANGED)) ctx->forc_valid[FC_VEGFRA] = false;
  ctx->hints = hints;
  return 0;
}

// Release the page-lock the library holds on a caller array (see noahmp_b200_create); a no-op for unknown pointers.
int noahmp_b200_unpin(noahmp_b200_ctx* ctx, const void* host_ptr_) {
  if (!ctx || !host_ptr_) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->registered.count(host_ptr_)) {
    CK(cudaDeviceSynchronize());
    unpin_one(ctx, host_ptr_);
  }
  return 0;
}

// ---- output / restart staging (row f3) ---------------------------------------------------------------------
// out = (mask && IVGTYP == ISWATER) ? -1.E33 : grid, for one field in the Fortran (i,k,j) layout
__global__ void stage_output_kernel(const float* __restrict__ grid, float* __restrict__ out, const float* __restrict__ ivgtyp_bits,
                                    int iswater, int mask, int layers, int ni, long long ncell) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const bool water = mask && __float_as_int(ivgtyp_bits[c]) == iswater;
  const long long i = c % ni, j = c / ni;
  for (int k = 0; k < layers; ++k) {
    const long long q = i + (long long)k * ni + j * ni * layers;
    out[q] = water ? -1.E33f : grid[q];
  }
}

int noahmp_b200_output_wait(noahmp_b200_ctx* ctx) {
  if (!ctx) return NOAHMP_ERR_ARG;
  if (!ctx->out_pending) return 0;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev_outdone));
  ctx->out_pending = false;
  return 0;
}

int noahmp_b200_output_begin(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, const char* fields, int mask_water) {
  if (!ctx || !a || !fields || !ctx->uploaded) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = check_bounds(ctx, a);
  if (rc) return rc;
  std::vector<int> list;
  if (!strcmp(fields, "*")) {
    for (int f = 0; f < NFIELDS; ++f) list.push_back(f);
  } else if ((rc = parse_fields(fields, list))) {
    return rc;
  }
  if ((rc = noahmp_b200_output_wait(ctx))) return rc;  // one snapshot in flight at a time
  if (list.empty()) return 0;
  if (!ctx->s_out) {
    CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  }
  if (!ctx->ev_outdone) {
    CK(cudaEventCreateWithFlags(&ctx->ev_outdone, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_outready, cudaEventDisableTiming));
  }
  size_t words = 0;
  for (int f : list) words += (size_t)ctx->ncell * kFields[f].layers;
  if (words > ctx->outstage_words) {
    if (ctx->d_outstage) CK(cudaFree(ctx->d_outstage));
    ctx->d_outstage = nullptr;
    ctx->outstage_words = 0;
    CK(cudaMalloc(&ctx->d_outstage, words * sizeof(float)));
    ctx->outstage_words = words;
  }
  cudaStream_t sk = ctx->stream;
  if (ctx->last_step_stream && ctx->last_step_stream != sk) {  // order the snapshot after steps on a caller stream
    CK(cudaEventRecord(ctx->ev_outready, ctx->last_step_stream));
    CK(cudaStreamWaitEvent(sk, ctx->ev_outready, 0));
  }
  const int T = 256;
  size_t off = 0;
  for (int f : list) {
    if ((rc = ensure_grid(ctx, f))) return rc;
    if (ctx->np > 0) {
      dim3 grid((unsigned)((ctx->np + T - 1) / T), kFields[f].layers);
      scatter_kernel<<<grid, T, 0, sk>>>(ctx->d_planes + kSlots.slot[f], ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
    }
    stage_output_kernel<<<(unsigned)((ctx->ncell + T - 1) / T), T, 0, sk>>>(
        ctx->d_grid[f], ctx->d_outstage + off, ctx->d_stat[ST_IVGTYP], ctx->iswater,
        // put_var_int writes integer fields unmasked (netcdf_io.F90:1985-2010)
        (mask_water && strcmp(kFields[f].name, "isnowxy")) ? 1 : 0, kFields[f].layers,
        ctx->ni, ctx->ncell);
    ctx->launches += 2;
    off += (size_t)ctx->ncell * kFields[f].layers;
  }
  CK(cudaGetLastError());
  // the snapshot is complete in stream order; later steps on this stream cannot touch it, and the copies below
  // run beside them on the download stream
  CK(cudaEventRecord(ctx->ev_outready, sk));
  CK(cudaStreamWaitEvent(ctx->s_out, ctx->ev_outready, 0));
  off = 0;
  for (int f : list) {
    const size_t n = (size_t)ctx->ncell * kFields[f].layers;
    pin(ctx, host_ptr(a, f), host_bytes(ctx, kFields[f].layers));
    CK(copy_field_rows(ctx, ctx->d_outstage + off, host_ptr(a, f), kFields[f].layers, 0, ctx->nj, false, ctx->s_out));
    off += n;
  }
  CK(cudaEventRecord(ctx->ev_outdone, ctx->s_out));
  ctx->out_pending = true;
  return 0;
}

// Re-bin the land columns every `interval` RESIDENT-mode steps (0 = never).  Default 20.
int noahmp_b200_set_rebin(noahmp_b200_ctx* ctx, int interval) {
  if (!ctx || interval < 0) return NOAHMP_ERR_ARG;
  ctx->rebin_interval = interval;
  return 0;
}
int noahmp_b200_rebin_count(const noahmp_b200_ctx* ctx) { return ctx ? ctx->rebins : 0; }

// Number of row chunks of the RESIDENT-mode pipeline (0 = automatic: 1 below 2^20 cells, else 9, the first and the
// last half as tall as the others).
int noahmp_b200_set_chunks(noahmp_b200_ctx* ctx, int nchunks) {
  if (!ctx || nchunks < 0 || nchunks > 64) return NOAHMP_ERR_ARG;
  ctx->nchunks = nchunks;
  return 0;
}

int noahmp_b200_noahmplsm(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, noahmp_status* status) {
  if (!ctx || !a) return NOAHMP_ERR_ARG;
  ++ctx->pin_clock;
  CK(cudaSetDevice(ctx->device));
  int rc;
  if (ctx->sync_mode == NOAHMP_SYNC_FULL || !ctx->uploaded) {
    if ((rc = noahmp_b200_upload(ctx, a))) { if (status) { status->code = rc; status->count = 0; } return rc; }
  }
  if (ctx->sync_mode == NOAHMP_SYNC_RESIDENT) {
    if ((rc = check_bounds(ctx, a))) return rc;
    fill_scalars(ctx, a);
    if ((rc = check_options(ctx))) return rc;
    if ((rc = step_resident_pipelined(ctx, a, auto_chunks(ctx)))) return rc;
    noahmp_status st;
    decode_status(ctx, &st);
    if (status) *status = st;
    return st.code;
  }
  if ((rc = step_device_impl(ctx, a->itimestep, a->yr, a->julian, a->dt, nullptr, /*clear_latch=*/true))) return rc;
  if (ctx->sync_mode == NOAHMP_SYNC_FULL) {
    if ((rc = scatter_fields(ctx))) return rc;
    if ((rc = download_state(ctx, a))) return rc;
  }
  CK(cudaMemcpyAsync(ctx->h_errkey, ctx->d_errkey, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_errcount, ctx->d_errcount, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  noahmp_status st;
  decode_status(ctx, &st);
  if (status) *status = st;
  return st.code;
}

// Refresh ONE host array from HBM (e.g. TSK / HFX / LH after a RESIDENT-mode step; the HRLDAS driver prints
// TSLB(1,1,1) and LAI(1,1) every step, module_hrldas_noahmp_driver.F90:567-572).
int noahmp_b200_fetch(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, const char* field) {
  if (!ctx || !a || !field || !ctx->uploaded) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int f = 0;
  for (; f < NFIELDS; ++f)
    if (!strcmp(field, kFields[f].name)) break;
  if (f == NFIELDS) { set_error(std::string("unknown field ") + field); return NOAHMP_ERR_ARG; }
  if (ctx->np > 0) {
    const int T = 256;
    dim3 grid((unsigned)((ctx->np + T - 1) / T), kFields[f].layers);
    scatter_kernel<<<grid, T, 0, ctx->stream>>>(ctx->d_planes + kSlots.slot[f], ctx->d_cell, ctx->d_state, ctx->np,
                                                 ctx->ni);
    ctx->launches++;
  }
  float* dst = host_ptr(a, f);
  int rc = check_bounds(ctx, a);
  if (rc) return rc;
  pin(ctx, dst, host_bytes(ctx, kFields[f].layers));
  CK(copy_field_rows(ctx, ctx->d_grid[f], dst, kFields[f].layers, 0, ctx->nj, false, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Point the step at caller-owned device forcing planes (same order and layout as device_forcing()); NULL
// restores the context's own staging planes.  Lets a driver keep several forcing hours resident in HBM.
int noahmp_b200_bind_forcing(noahmp_b200_ctx* ctx, float* const* dev_ptrs) {
  if (!ctx) return NOAHMP_ERR_ARG;
  for (int f = 0; f < NFORC; ++f) ctx->base.forc[f] = dev_ptrs ? dev_ptrs[f] : ctx->d_forc[f];
  return 0;
}

// ---- opt_run = 5 groundwater step ---------------------------------------------------------------------------
enum WtPlane { WT_FDEPTH = 0, WT_AREA, WT_TOPO, WT_RIVERCOND, WT_RIVERBED, WT_EQWTD, WT_PEXP, WT_QRF, WT_QSPRING,
               WT_QSLAT, WT_QRFS, WT_QSPRINGS, NWT };

// non-land bookkeeping of WTABLE_mmf_noahmp's last loops (:112-129, :186-195): QRF = 0 (QSPRING, INTENT(OUT) and
// never assigned there in the reference, is defined as 0), RECH += DEEPRECH*1e3, DEEPRECH = 0
__global__ void wt_nonland_cells_kernel(const unsigned char* __restrict__ cls, float* qrf, float* qspring,
                                        float* rech_grid, float* deeprech_grid, long long ncell) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const unsigned char k = cls[c];
  if (k == CL_LAND) return;
  qrf[c] = 0.f;
  qspring[c] = 0.f;
  if (k == CL_WATER) {  // cells without a compact column: their accumulators live in the grid-order arrays
    rech_grid[c] = rech_grid[c] + deeprech_grid[c] * 1.E3f;
    deeprech_grid[c] = 0.f;
  }
}
__global__ void wt_nonland_cols_kernel(float* state, long long np, int first, int count) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  long long n = (long long)first + t;
  float* rech = state + (long long)NMP_SLOT(rechxy) * np;
  float* deep = state + (long long)NMP_SLOT(deeprechxy) * np;
  rech[n] = rech[n] + deep[n] * 1.E3f;
  deep[n] = 0.f;
}

struct WtField { int f; size_t off; };
static const WtField kWtInout[] = {{F_smois, offsetof(noahmp_wtable_args, smois)},
                                   {F_sh2o, offsetof(noahmp_wtable_args, sh2oxy)},
                                   {F_smcwtdxy, offsetof(noahmp_wtable_args, smcwtd)},
                                   {F_zwtxy, offsetof(noahmp_wtable_args, wtd)},
                                   {F_deeprechxy, offsetof(noahmp_wtable_args, deeprech)},
                                   {F_rechxy, offsetof(noahmp_wtable_args, rech)}};
static inline float* wt_ptr(const noahmp_wtable_args* a, size_t off) {
  return *reinterpret_cast<float* const*>(reinterpret_cast<const char*>(a) + off);
}

static int wt_check(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  if (!ctx || !a) return NOAHMP_ERR_ARG;
  if (!ctx->uploaded) { set_error("wtable before the state was uploaded (call noahmplsm or upload first)"); return NOAHMP_ERR_ARG; }
  if (a->ims > a->its || a->ime < a->ite || a->jms > a->jts || a->jme < a->jte ||
      a->ite - a->its + 1 != ctx->ni || a->jte - a->jts + 1 != ctx->nj || a->nsoil != NOAHMP_NSOIL) {
    set_error("wtable: memory bounds must contain the tile bounds and the tile must match the context");
    return NOAHMP_ERR_ARG;
  }
  ctx->h_mi = a->ime - a->ims + 1;
  ctx->h_mj = a->jme - a->jms + 1;
  ctx->h_x0 = a->its - a->ims;
  ctx->h_y0 = a->jts - a->jms;
  ctx->h_its = a->its;
  ctx->h_jts = a->jts;
  return 0;
}

static int gather_field(noahmp_b200_ctx* ctx, int f) {
  if (ctx->np == 0) return 0;
  const int T = 256;
  dim3 grid((unsigned)((ctx->np + T - 1) / T), kFields[f].layers);
  gather_kernel<<<grid, T, 0, ctx->stream>>>(ctx->d_planes + kSlots.slot[f], ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
  ctx->launches++;
  return 0;
}
static int scatter_field(noahmp_b200_ctx* ctx, int f) {
  if (ctx->np == 0) return 0;
  const int T = 256;
  dim3 grid((unsigned)((ctx->np + T - 1) / T), kFields[f].layers);
  scatter_kernel<<<grid, T, 0, ctx->stream>>>(ctx->d_planes + kSlots.slot[f], ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
  ctx->launches++;
  return 0;
}

static void wt_params(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a, WtParams& w) {
  w.fdepth = ctx->d_wt[WT_FDEPTH]; w.area = ctx->d_wt[WT_AREA]; w.topo = ctx->d_wt[WT_TOPO];
  w.rivercond = ctx->d_wt[WT_RIVERCOND]; w.riverbed = ctx->d_wt[WT_RIVERBED]; w.eqwtd = ctx->d_wt[WT_EQWTD];
  w.pexp = ctx->d_wt[WT_PEXP];
  w.isltyp = ctx->d_stat[ST_ISLTYP]; w.ivgtyp = ctx->d_stat[ST_IVGTYP];
  w.wtd_grid = ctx->d_grid[F_zwtxy];
  w.qrf = ctx->d_wt[WT_QRF]; w.qspring = ctx->d_wt[WT_QSPRING]; w.qslat = ctx->d_wt[WT_QSLAT];
  w.qrfs = ctx->d_wt[WT_QRFS]; w.qsprings = ctx->d_wt[WT_QSPRINGS];
  w.kcell = ctx->d_kcell; w.head = ctx->d_head;
  w.state = ctx->d_state; w.cell = ctx->d_cell; w.tables = ctx->d_tables; w.np = ctx->np;
  w.nland = ctx->nclass[CL_LAND]; w.ni = ctx->ni; w.nj = ctx->nj;
  w.ids = a->ids; w.ide = a->ide; w.jds = a->jds; w.jde = a->jde;
  w.its = a->its; w.ite = a->ite; w.jts = a->jts; w.jte = a->jte;
  w.isurban = a->isurban;
  w.deltat = a->wtddt * 60.f;
  for (int k = 0; k < NOAHMP_NSOIL; ++k) w.dzs[k] = a->dzs[k];
}

// allocation and the uploads that happen once (static inputs, accumulators) or on every call in SYNC_FULL mode
static int wt_prepare(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc;
  const size_t plane = sizeof(float) * ctx->ncell, halo = sizeof(float) * (size_t)(ctx->ni + 2) * (ctx->nj + 2);
  if (!ctx->d_kcell) {
    for (int k = 0; k < NWT; ++k) {
      CK(cudaMalloc(&ctx->d_wt[k], plane));
      CK(cudaMemsetAsync(ctx->d_wt[k], 0, plane, ctx->stream));
    }
    CK(cudaMalloc(&ctx->d_kcell, halo));
    CK(cudaMalloc(&ctx->d_head, halo));
  }
  const bool full = ctx->sync_mode == NOAHMP_SYNC_FULL;
  if (full || !ctx->wt_init) {
    const float* src[NWT] = {a->fdepth, a->area, a->topo, a->rivercond, a->riverbed, a->eqwtd, a->pexp, nullptr, nullptr,
                             a->qslat, a->qrfs, a->qsprings};
    for (int k = 0; k < NWT; ++k)
      if (src[k]) {
        pin(ctx, src[k], host_bytes(ctx, 1));
        CK(copy_field_rows(ctx, ctx->d_wt[k], const_cast<float*>(src[k]), 1, 0, ctx->nj, true, ctx->stream));
      }
    // SMOISEQ is static input of the scheme
    pin(ctx, a->smoiseq, host_bytes(ctx, NOAHMP_NSOIL));
    CK(copy_field_rows(ctx, ctx->d_grid[F_smoiseq], const_cast<float*>(a->smoiseq), NOAHMP_NSOIL, 0, ctx->nj, true, ctx->stream));
    if ((rc = gather_field(ctx, F_smoiseq))) return rc;
    ctx->wt_init = true;
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// LATERALFLOW pass 1 on stream s: WTD of every cell in grid order, then KCELL / HEAD with an empty ring
static int wt_enqueue_head(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a, cudaStream_t s) {
  const size_t halo = sizeof(float) * (size_t)(ctx->ni + 2) * (ctx->nj + 2);
  if (ctx->np > 0) {
    const int T = 256;
    dim3 grid((unsigned)((ctx->np + T - 1) / T), 1);
    scatter_kernel<<<grid, T, 0, s>>>(ctx->d_planes + kSlots.slot[F_zwtxy], ctx->d_cell, ctx->d_state, ctx->np, ctx->ni);
    ctx->launches++;
  }
  CK(cudaMemsetAsync(ctx->d_kcell, 0, halo, s));
  CK(cudaMemsetAsync(ctx->d_head, 0, halo, s));
  WtParams w;
  wt_params(ctx, a, w);
  if (ctx->math_mode == NOAHMP_MATH_PARITY) nmp_launch_wtable_parity(w, s, &ctx->launches, 0);
  else nmp_launch_wtable_fast(w, s, &ctx->launches, 0);
  CK(cudaGetLastError());
  return 0;
}
// pass 2 + river flux + UPDATEWTD per land column, bookkeeping of the other cells, on stream s
static int wt_enqueue_columns(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a, cudaStream_t s) {
  WtParams w;
  wt_params(ctx, a, w);
  if (ctx->math_mode == NOAHMP_MATH_PARITY) nmp_launch_wtable_parity(w, s, &ctx->launches, 1);
  else nmp_launch_wtable_fast(w, s, &ctx->launches, 1);
  const int T = 256;
  wt_nonland_cells_kernel<<<(unsigned)((ctx->ncell + T - 1) / T), T, 0, s>>>(
      ctx->d_class, ctx->d_wt[WT_QRF], ctx->d_wt[WT_QSPRING], ctx->d_grid[F_rechxy], ctx->d_grid[F_deeprechxy], ctx->ncell);
  ctx->launches++;
  const int nother = (int)ctx->np - ctx->nclass[CL_LAND];
  if (nother > 0) {
    wt_nonland_cols_kernel<<<(nother + T - 1) / T, T, 0, s>>>(ctx->d_state, ctx->np, ctx->nclass[CL_LAND], nother);
    ctx->launches++;
  }
  CK(cudaGetLastError());
  return 0;
}

int noahmp_b200_wtable_begin(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc = wt_check(ctx, a);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  ++ctx->pin_clock;
  if ((rc = wt_prepare(ctx, a))) return rc;
  if (ctx->sync_mode == NOAHMP_SYNC_FULL) {
    for (const WtField& wf : kWtInout) {
      pin(ctx, wt_ptr(a, wf.off), host_bytes(ctx, kFields[wf.f].layers));
      CK(copy_field_rows(ctx, ctx->d_grid[wf.f], wt_ptr(a, wf.off), kFields[wf.f].layers, 0, ctx->nj, true, ctx->stream));
      if ((rc = gather_field(ctx, wf.f))) return rc;
    }
  } else {
    CK(cudaDeviceSynchronize());  // resident steps may have run on a caller stream
  }
  if ((rc = wt_enqueue_head(ctx, a, ctx->stream))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));  // the halo exchange that may follow runs on the caller's stream
  return 0;
}

int noahmp_b200_wtable_halo(noahmp_b200_ctx* ctx, float** kcell, float** head) {
  if (!ctx || !ctx->d_kcell || !kcell || !head) return NOAHMP_ERR_ARG;
  *kcell = ctx->d_kcell;
  *head = ctx->d_head;
  return 0;
}

static int wt_download(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc;
  for (const WtField& wf : kWtInout) {
    if ((rc = scatter_field(ctx, wf.f))) return rc;
    pin(ctx, wt_ptr(a, wf.off), host_bytes(ctx, kFields[wf.f].layers));
    CK(copy_field_rows(ctx, ctx->d_grid[wf.f], wt_ptr(a, wf.off), kFields[wf.f].layers, 0, ctx->nj, false, ctx->stream));
  }
  float* dst[5] = {a->qrf, a->qspring, a->qslat, a->qrfs, a->qsprings};
  const int src[5] = {WT_QRF, WT_QSPRING, WT_QSLAT, WT_QRFS, WT_QSPRINGS};
  for (int k = 0; k < 5; ++k) {
    pin(ctx, dst[k], host_bytes(ctx, 1));
    CK(copy_field_rows(ctx, ctx->d_wt[src[k]], dst[k], 1, 0, ctx->nj, false, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int noahmp_b200_wtable_end(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc = wt_check(ctx, a);
  if (rc) return rc;
  if (!ctx->d_kcell) { set_error("wtable_end before wtable_begin"); return NOAHMP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());  // halo writes of the caller are complete
  if ((rc = wt_enqueue_columns(ctx, a, ctx->stream))) return rc;
  if (ctx->sync_mode == NOAHMP_SYNC_FULL) return wt_download(ctx, a);
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Halo exchange of KCELL / HEAD with the neighbouring tiles over the context's communicator (between _begin and _end)
int noahmp_b200_wtable_exchange(noahmp_b200_ctx* ctx, void* stream) {
  if (!ctx || !ctx->d_kcell) { set_error("wtable_exchange before wtable_begin"); return NOAHMP_ERR_ARG; }
  if (!ctx->comm.comm || ctx->comm.nranks == 1) return 0;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  return halo_exchange(ctx->comm, ctx->d_kcell, ctx->d_head, ctx->ni, ctx->nj, s, &ctx->launches);
}

// CALL WTABLE_mmf_noahmp(...): with a communicator (noahmp_b200_comm_init) the halo is exchanged inside the call
int noahmp_b200_wtable(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc = noahmp_b200_wtable_begin(ctx, a);
  if (rc) return rc;
  if ((rc = noahmp_b200_wtable_exchange(ctx, nullptr))) return rc;
  return noahmp_b200_wtable_end(ctx, a);
}

// The same call for a device-side stepping loop (RESIDENT mode): everything — pass 1, the NCCL halo, pass 2 and the
// column update — is enqueued on `stream` behind the step that precedes it; the host does not wait.
int noahmp_b200_wtable_device(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a, void* stream) {
  int rc = wt_check(ctx, a);
  if (rc) return rc;
  if (ctx->sync_mode != NOAHMP_SYNC_RESIDENT) { set_error("wtable_device needs RESIDENT mode"); return NOAHMP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  if (!ctx->wt_init && (rc = wt_prepare(ctx, a))) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  if (ctx->last_step_stream && ctx->last_step_stream != s) {
    if (!ctx->ev_rebin) CK(cudaEventCreateWithFlags(&ctx->ev_rebin, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ev_rebin, ctx->last_step_stream));
    CK(cudaStreamWaitEvent(s, ctx->ev_rebin, 0));
  }
  ctx->last_step_stream = s;
  if ((rc = wt_enqueue_head(ctx, a, s))) return rc;
  if (ctx->comm.comm && ctx->comm.nranks > 1 &&
      (rc = halo_exchange(ctx->comm, ctx->d_kcell, ctx->d_head, ctx->ni, ctx->nj, s, &ctx->launches))) return rc;
  return wt_enqueue_columns(ctx, a, s);
}

// ---- communicator ---------------------------------------------------------------------------------------------------
// rank 0 calls comm_unique_id and the host program hands the 128 bytes to every rank (MPI_Bcast in the Fortran driver,
// torch.distributed.broadcast under torchrun); then every rank calls comm_init.  Ranks are laid out as
// mpp_land_partition does: rank = iprocx + iprocy * nprocx (mpp/module_mpp_land.F90:83-84).
int noahmp_b200_comm_unique_id(void* id128) {
  if (!id128) return NOAHMP_ERR_ARG;
  NcclApi* N = nccl_api();
  if (!N) { set_error("libnccl.so.2 not found (set NOAHMP_B200_NCCL to its path)"); return NOAHMP_ERR_CUDA; }
  ncclUniqueId id;
  NK(N->GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int noahmp_b200_comm_init(noahmp_b200_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return NOAHMP_ERR_ARG;
  NcclApi* N = nccl_api();
  if (!N) { set_error("libnccl.so.2 not found (set NOAHMP_B200_NCCL to its path)"); return NOAHMP_ERR_CUDA; }
  CK(cudaSetDevice(ctx->device));
  if (ctx->comm.comm) { N->CommDestroy(ctx->comm.comm); ctx->comm.comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NK(N->CommInitRank(&ctx->comm.comm, nranks, id, rank));
  NmpComm& C = ctx->comm;
  C.rank = rank; C.nranks = nranks;
  int nb[4];
  noahmp_b200_tile_neighbours(nranks, rank, nb);
  C.left = nb[0]; C.right = nb[1]; C.down = nb[2]; C.up = nb[3];
  if (!C.d_send) {
    CK(cudaMalloc(&C.d_send, sizeof(float) * 4 * (size_t)ctx->nj));
    CK(cudaMalloc(&C.d_recv, sizeof(float) * 4 * (size_t)ctx->nj));
    CK(cudaMalloc(&C.d_budget_sum, sizeof(double) * NBUDGET));
  }
  return 0;
}

// ranks of the tiles left, right, below (smaller j) and above `rank` in the mpp_land process grid; -1 at the edge
void noahmp_b200_tile_neighbours(int nranks, int rank, int nb[4]) {
  int npx, npy;
  noahmp_b200_proc_grid(nranks, &npx, &npy);
  const int ipx = rank % npx, ipy = rank / npx;
  nb[0] = ipx > 0 ? rank - 1 : -1;
  nb[1] = ipx < npx - 1 ? rank + 1 : -1;
  nb[2] = ipy > 0 ? rank - npx : -1;
  nb[3] = ipy < npy - 1 ? rank + npx : -1;
}

int noahmp_b200_comm_neighbours(const noahmp_b200_ctx* ctx, int nb[4]) {
  if (!ctx || !nb) return NOAHMP_ERR_ARG;
  nb[0] = ctx->comm.left; nb[1] = ctx->comm.right; nb[2] = ctx->comm.down; nb[3] = ctx->comm.up;
  return 0;
}

// ---- global water / energy budget (nmp_comm.cuh: budget_kernel) -----------------------------------------------------
int noahmp_b200_budget_enable(noahmp_b200_ctx* ctx, int enable) {
  if (!ctx) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (enable && !ctx->d_budget) CK(cudaMalloc(&ctx->d_budget, sizeof(double) * NBUDGET));
  if (enable) {
    CK(cudaMemset(ctx->d_budget, 0, sizeof(double) * NBUDGET));
    ctx->budget_steps = 0;
  }
  ctx->budget_on = enable != 0;
  return 0;
}

// called behind every step when the budget is enabled
static int budget_accumulate(noahmp_b200_ctx* ctx, float dt, cudaStream_t s) {
  const int ncol = ctx->nclass[CL_LAND] + ctx->nclass[CL_GLACIER];
  ctx->budget_steps++;
  if (ncol == 0) return 0;
  // the instantaneous sums restart with every step
  CK(cudaMemsetAsync(ctx->d_budget, 0, sizeof(double), s));
  CK(cudaMemsetAsync(ctx->d_budget + 5, 0, sizeof(double), s));
  const StepParams& b = ctx->base;
  const float dz0 = -b.zsoil[0], dz1 = b.zsoil[0] - b.zsoil[1], dz2 = b.zsoil[1] - b.zsoil[2], dz3 = b.zsoil[2] - b.zsoil[3];
  const int T = 256;
  budget_kernel<<<(ncol + T - 1) / T, T, 0, s>>>(ctx->d_state, b.forc[FC_RAINBL], ctx->d_cell, ctx->np, ctx->nclass[CL_LAND],
                                                 ncol, dt, dz0, dz1, dz2, dz3, ctx->d_budget);
  ctx->launches++;
  CK(cudaGetLastError());
  return 0;
}

// out[8] as listed in nmp_comm.cuh; global != 0 sums over the ranks of the communicator (one ncclAllReduce of 8 fp64).
// reset != 0 starts a new accumulation interval.
int noahmp_b200_budget_read(noahmp_b200_ctx* ctx, double* out, int global, int reset) {
  if (!ctx || !out || !ctx->d_budget) { set_error("budget_read before budget_enable"); return NOAHMP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->last_step_stream ? ctx->last_step_stream : ctx->stream;
  const double tail[2] = {(double)(ctx->nclass[CL_LAND] + ctx->nclass[CL_GLACIER]), (double)ctx->budget_steps};
  CK(cudaMemcpyAsync(ctx->d_budget + 6, tail, sizeof(tail), cudaMemcpyHostToDevice, s));
  const double* src = ctx->d_budget;
  if (global && ctx->comm.comm && ctx->comm.nranks > 1) {
    NcclApi* N = nccl_api();
    NK(N->AllReduce(ctx->d_budget, ctx->comm.d_budget_sum, NBUDGET, ncclFloat64, ncclSum, ctx->comm.comm, s));
    src = ctx->comm.d_budget_sum;
  }
  CK(cudaMemcpyAsync(out, src, sizeof(double) * NBUDGET, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (global && ctx->comm.nranks > 1) out[7] /= ctx->comm.nranks;  // every rank counted the same steps
  if (reset) {
    CK(cudaMemsetAsync(ctx->d_budget, 0, sizeof(double) * NBUDGET, s));
    ctx->budget_steps = 0;
  }
  return 0;
}

int noahmp_b200_wtable_sync_host(noahmp_b200_ctx* ctx, const noahmp_wtable_args* a) {
  int rc = wt_check(ctx, a);
  if (rc) return rc;
  if (!ctx->d_kcell) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  return wt_download(ctx, a);
}

// ---- on-device forcing pipeline (row f2) ------------------------------------------------------------------------
static int fb_alloc(noahmp_b200_ctx* ctx) {
  if (ctx->d_lat) return 0;
  const size_t plane = sizeof(float) * ctx->ncell;
  for (auto& b : ctx->d_fb)
    for (auto& p : b) { CK(cudaMalloc(&p, plane)); CK(cudaMemsetAsync(p, 0, plane, ctx->stream)); }
  CK(cudaMalloc(&ctx->d_lat, plane));
  CK(cudaMalloc(&ctx->d_lon, plane));
  for (auto& e : ctx->ev_fb) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if (!ctx->s_in) {
    CK(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int noahmp_b200_forcing_static(noahmp_b200_ctx* ctx, const float* lat2d, const float* lon2d, float zlvl) {
  if (!ctx || !lat2d || !lon2d) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = fb_alloc(ctx);
  if (rc) return rc;
  const size_t plane = sizeof(float) * ctx->ncell;
  CK(cudaMemcpyAsync(ctx->d_lat, lat2d, plane, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_lon, lon2d, plane, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->zlvl = zlvl;
  ctx->fb_static = true;
  return 0;
}

int noahmp_b200_forcing_upload(noahmp_b200_ctx* ctx, int slot, const noahmp_forcing_fields* f) {
  if (!ctx || !f || slot < 0 || slot > 1) return NOAHMP_ERR_ARG;
  ++ctx->pin_clock;
  CK(cudaSetDevice(ctx->device));
  int rc = fb_alloc(ctx);
  if (rc) return rc;
  const float* src[9] = {f->t, f->q, f->u, f->v, f->p, f->lw, f->sw, f->pcp, f->fpar};
  const size_t plane = sizeof(float) * ctx->ncell;
  // the physics may still be reading planes derived from this slot: order the copy after the compute stream
  CK(cudaEventRecord(ctx->ev_fb[slot], ctx->stream));
  CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_fb[slot], 0));
  for (int k = 0; k < 9; ++k) {
    if (!src[k]) { set_error("forcing_upload: null field"); return NOAHMP_ERR_ARG; }
    pin(ctx, src[k], plane);
    CK(cudaMemcpyAsync(ctx->d_fb[slot][k], src[k], plane, cudaMemcpyHostToDevice, ctx->s_in));
  }
  CK(cudaEventRecord(ctx->ev_fb[slot], ctx->s_in));
  ctx->fb_loaded[slot] = true;
  return 0;
}

int noahmp_b200_forcing_swap(noahmp_b200_ctx* ctx) {
  if (!ctx) return NOAHMP_ERR_ARG;
  for (int k = 0; k < 9; ++k) std::swap(ctx->d_fb[0][k], ctx->d_fb[1][k]);
  std::swap(ctx->ev_fb[0], ctx->ev_fb[1]);
  std::swap(ctx->fb_loaded[0], ctx->fb_loaded[1]);
  return 0;
}

int noahmp_b200_forcing_apply(noahmp_b200_ctx* ctx, float fraction, int iday, int ihour, int iminute, int isecond,
                              float model_timestep, float* julian_out) {
  if (!ctx || !ctx->fb_static || !ctx->fb_loaded[0]) { set_error("forcing_apply before forcing_static / forcing_upload"); return NOAHMP_ERR_ARG; }
  if (fraction != 1.0f && !ctx->fb_loaded[1]) { set_error("forcing_apply: bracket B not loaded"); return NOAHMP_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  const bool parity = ctx->math_mode == NOAHMP_MATH_PARITY;
  auto fsin = [&](float x) { return parity ? nmpm::sinf_(x) : sinf(x); };
  auto fcos = [&](float x) { return parity ? nmpm::cosf_(x) : cosf(x); };
  auto fasin = [&](float x) { return parity ? nmpm::asinf_(x) : asinf(x); };
  // CALC_DECLIN, the part that does not depend on the cell (:833-850)
  const float DEGRAD = 3.14159265f / 180.f, DPD = 360.f / 365.f;
  const float JULIAN = (float)iday + (float)ihour / 24.f;
  const float OBECL = 23.5f * DEGRAD;
  const float SINOB = fsin(OBECL);
  float SXLONG = 0.f;
  if (JULIAN >= 80.f) SXLONG = DPD * (JULIAN - 80.f) * DEGRAD;
  if (JULIAN < 80.f) SXLONG = DPD * (JULIAN + 285.f) * DEGRAD;
  const float ARG = SINOB * fsin(SXLONG);
  const float DECLIN = fasin(ARG);
  ForcingParams f;
  // exactly at file A (hrldas_input_copy): bracket B is not needed, so a B upload still in flight is not waited for
  // (A*1 + A*0 gives the same bits as A*1 + B*0 for finite data)
  const bool needB = fraction != 1.0f;
  for (int k = 0; k < 9; ++k) { f.A[k] = ctx->d_fb[0][k]; f.B[k] = needB ? ctx->d_fb[1][k] : ctx->d_fb[0][k]; }
  f.lat = ctx->d_lat; f.lon = ctx->d_lon;
  for (int k = 0; k < NFORC; ++k) f.out[k] = ctx->d_forc[k];
  f.ncell = ctx->ncell;
  f.fraction = fraction;
  f.sin_declin = fsin(DECLIN);
  f.cos_declin = fcos(DECLIN);
  f.hour_frac = (float)ihour + (float)iminute / 60.0f + (float)isecond / 3600.0f;
  f.dt = model_timestep;
  f.zlvl = ctx->zlvl;
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fb[0], 0));
  if (needB) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_fb[1], 0));
  if (parity) nmp_launch_forcing_parity(f, ctx->stream, &ctx->launches);
  else nmp_launch_forcing_fast(f, ctx->stream, &ctx->launches);
  CK(cudaGetLastError());
  if (julian_out) *julian_out = JULIAN;
  for (int k = 0; k < NFORC; ++k) ctx->base.forc[k] = ctx->d_forc[k];
  return 0;
}

int noahmp_b200_noahmplsm_device_forcing(noahmp_b200_ctx* ctx, const noahmp_lsm_args* a, noahmp_status* status) {
  if (!ctx || !a) return NOAHMP_ERR_ARG;
  ++ctx->pin_clock;
  if (ctx->sync_mode != NOAHMP_SYNC_RESIDENT || !ctx->uploaded) {
    set_error("noahmplsm_device_forcing needs RESIDENT mode and an uploaded state");
    return NOAHMP_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = check_bounds(ctx, a))) return rc;
  fill_scalars(ctx, a);
  if ((rc = check_options(ctx))) return rc;
  int nch = auto_chunks(ctx);
  if (ctx->fetch.empty() && ctx->push.empty()) nch = 1;  // nothing to overlap the physics with
  if ((rc = step_resident_pipelined(ctx, a, nch, /*upload=*/false))) return rc;
  noahmp_status st;
  decode_status(ctx, &st);
  if (status) *status = st;
  return st.code;
}

long long noahmp_b200_launch_count(const noahmp_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- cold start (row f1): NOAHMP_INIT on the device --------------------------------------------------------
struct InitFieldInfo {
  const char* name;
  size_t arg_offset;
  int layers, io;
};
static const InitFieldInfo kInitFields[NINITF] = {
#define X(nm, nl, io) {#nm, offsetof(noahmp_init_args, nm), nl, io},
    NMP_INIT_FIELDS(X) NMP_INIT_GW_FIELDS(X)
#undef X
};
unsigned long long noahmp_b200_sizeof_init_args(void) { return sizeof(noahmp_init_args); }

int noahmp_b200_init(noahmp_b200_ctx* ctx, const noahmp_init_args* a) {
  if (!ctx || !a) return NOAHMP_ERR_ARG;
  if (a->nsoil != NOAHMP_NSOIL) { set_error("init: nsoil must be 4"); return NOAHMP_ERR_ARG; }
  if (a->restart) return 0;  // IF (.NOT. restart): a restart run takes every field from the restart file
  CK(cudaSetDevice(ctx->device));
  const int ni = a->ime - a->ims + 1, nj = a->jme - a->jms + 1;
  if (ni != ctx->ni || nj != ctx->nj || a->ims != a->its || a->ime != a->ite || a->jms != a->jts || a->jme != a->jte) {
    set_error("init: the memory extent must be the tile the context was created for (ims=its ... jme=jte)");
    return NOAHMP_ERR_ARG;
  }
  const bool gw = a->iopt_run == 5;
  auto hp = [&](int f) { return *reinterpret_cast<float* const*>(reinterpret_cast<const char*>(a) + kInitFields[f].arg_offset); };
  for (int f = 0; f < NINITF; ++f)
    if ((f < IF_GW0 || gw) && !hp(f)) {
      set_error(f < IF_GW0 ? std::string("init: missing array ") + kInitFields[f].name
                           : std::string("Not enough fields to use groundwater option in Noah-MP: ") + kInitFields[f].name);
      return NOAHMP_ERR_ARG;
    }
  if (!a->dzs || (gw && !a->stepwtd)) { set_error("init: dzs / stepwtd missing"); return NOAHMP_ERR_ARG; }
  if (gw && (a->ids < a->its || a->ide > a->ite + 1 || a->jds < a->jts || a->jde > a->jte + 1)) {
    // LATERALFLOW would read WTD one cell outside the memory the caller passed
    set_error("init: groundwater initialisation needs ids>=its, ide<=ite+1, jds>=jts, jde<=jte+1");
    return NOAHMP_ERR_ARG;
  }
  InitParams P{};
  const size_t plane = (size_t)ni * nj;
  size_t words = 0;
  for (int f = 0; f < NINITF; ++f)
    if (f < IF_GW0 || gw) words += plane * kInitFields[f].layers;
  if (gw) words += 2 * plane;
  float* buf = nullptr;
  int* d_err = nullptr;
  CK(cudaMalloc(&buf, words * sizeof(float)));
  auto fail = [&](int rc) { cudaFree(buf); if (d_err) cudaFree(d_err); return rc; };
  if (cudaMalloc(&d_err, sizeof(int)) != cudaSuccess) { set_error("init: cudaMalloc"); return fail(NOAHMP_ERR_CUDA); }
  cudaStream_t s = ctx->stream;
  size_t off = 0;
  for (int f = 0; f < NINITF; ++f) {
    if (!(f < IF_GW0 || gw)) continue;
    P.f[f] = buf + off;
    const size_t n = plane * kInitFields[f].layers;
    off += n;
    // no page-locking here: the call runs once, and its arrays need not outlive it
    if (cudaMemcpyAsync(P.f[f], hp(f), n * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) {
      set_error(std::string("init: upload of ") + kInitFields[f].name + " failed");
      return fail(NOAHMP_ERR_CUDA);
    }
  }
  if (gw) { P.kcell = buf + off; P.head = buf + off + plane; }
  P.err = d_err;
  P.tables = ctx->d_tables;
  P.ni = ni; P.nj = nj;
  P.itf_n = std::min(a->ite, a->ide - 1) - a->its + 1;
  P.jtf_n = std::min(a->jte, a->jde - 1) - a->jts + 1;
  P.ids = a->ids; P.ide = a->ide; P.jds = a->jds; P.jde = a->jde;
  P.its = a->its; P.ite = a->ite; P.jts = a->jts; P.jte = a->jte;
  P.isurban = a->isurban; P.isice = a->isice; P.iswater = a->iswater; P.iopt_run = a->iopt_run;
  P.fndsnowh = a->fndsnowh;
  P.dx = a->dx; P.dy = a->dy; P.deltat = a->wtddt * 60.f;
  for (int k = 0; k < NOAHMP_NSOIL; ++k) P.dzs[k] = a->dzs[k];
  int herr = 0;
  if (cudaMemsetAsync(d_err, 0, sizeof(int), s) != cudaSuccess) return fail(NOAHMP_ERR_CUDA);
  nmp_launch_init(P, s, &ctx->launches);
  if (cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
      cudaStreamSynchronize(s) != cudaSuccess) {
    set_error(std::string("init: ") + cudaGetErrorString(cudaGetLastError()));
    return fail(NOAHMP_ERR_CUDA);
  }
  if (herr == 1) {
    set_error("module_sf_noahlsm.F: lsminit: out of range value of ISLTYP. Is this field in the input?");
    return fail(NOAHMP_ERR_ISLTYP);
  }
  if (herr == 2) { set_error("Problem with the logic assigning snow layers."); return fail(NOAHMP_ERR_ARG); }
  for (int f = 0; f < NINITF; ++f) {
    if (!(f < IF_GW0 || gw) || !(kInitFields[f].io & 2)) continue;
    const size_t n = plane * kInitFields[f].layers;
    if (cudaMemcpyAsync(hp(f), P.f[f], n * sizeof(float), cudaMemcpyDeviceToHost, s) != cudaSuccess) {
      set_error(std::string("init: download of ") + kInitFields[f].name + " failed");
      return fail(NOAHMP_ERR_CUDA);
    }
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) return fail(NOAHMP_ERR_CUDA);
  if (gw) {
    // STEPWTD = max(nint(WTDDT*60./DT), 1)   (:1159-1160)
    const float x = a->wtddt * 60.f / a->dt;
    const int n = (int)(x >= 0.f ? floorf(x + 0.5f) : -floorf(-x + 0.5f));
    *a->stepwtd = std::max(n, 1);
  }
  // the host arrays changed: a context that had state uploaded must take them again
  ctx->uploaded = false;
  return fail(0);
}

int noahmp_b200_census(const noahmp_b200_ctx* ctx, int64_t counts[4]) {
  if (!ctx || !ctx->classified) return NOAHMP_ERR_ARG;
  counts[0] = ctx->nclass[CL_LAND];
  counts[1] = ctx->nclass[CL_GLACIER];
  counts[2] = ctx->nclass[CL_SEAICE];
  counts[3] = ctx->nclass[CL_WATER];
  return 0;
}

// compact column -> 0-based tile-local cell index map (land | glacier | sea-ice), for tests and drivers
int noahmp_b200_column_map(noahmp_b200_ctx* ctx, int32_t* cells, long long capacity) {
  if (!ctx || !ctx->classified || !cells || capacity < ctx->np) return NOAHMP_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(cells, ctx->d_cell, sizeof(int) * ctx->np, cudaMemcpyDeviceToHost));
  return 0;
}

// Device pointer of plane `layer` of a state field in the compact SoA (np words); NULL if unknown.
float* noahmp_b200_device_state(noahmp_b200_ctx* ctx, const char* field, int layer, long long* np) {
  if (!ctx || !ctx->uploaded || !field) return nullptr;
  for (int f = 0; f < NFIELDS; ++f) {
    if (!strcmp(field, kFields[f].name)) {
      if (layer < 0 || layer >= kFields[f].layers) return nullptr;
      if (np) *np = ctx->np;
      return ctx->d_state + (size_t)(kSlots.slot[f] + layer) * (size_t)ctx->np;
    }
  }
  return nullptr;
}

// ---- domain decomposition (mpp/module_mpp_land.F90:124-141, :245-288) -----------------------------------
void noahmp_b200_proc_grid(int nproc, int* nprocx, int* nprocy) {
  // mpp_land_get_nprocsxy: over j = 1..nproc with nproc % j == 0, take (nx = nproc/j, ny = j) whenever it
  // strictly lowers |nx - ny|; the running minimum starts at nproc.
  int best = nproc, bx = nproc, by = 1;
  for (int j = 1; j <= nproc; ++j) {
    if (nproc % j == 0) {
      int i = nproc / j;
      int d = i > j ? i - j : j - i;
      if (d < best) { best = d; bx = i; by = j; }
    }
  }
  *nprocx = bx;
  *nprocy = by;
}

void noahmp_b200_tile(int global_nx, int global_ny, int nproc, int rank, int* xstart, int* xend, int* ystart,
                      int* yend) {
  int npx, npy;
  noahmp_b200_proc_grid(nproc, &npx, &npy);
  const int ipx = rank % npx, ipy = rank / npx;
  auto split = [](int n, int np, int ip, int* s, int* e) {
    const int base = n / np, rem = n % np;
    const int len = base + (ip < rem ? 1 : 0);
    int start = 1;
    for (int k = 0; k < ip; ++k) start += base + (k < rem ? 1 : 0);
    *s = start;
    *e = start + len - 1;
  };
  split(global_nx, npx, ipx, xstart, xend);
  split(global_ny, npy, ipy, ystart, yend);
}

// ---- one host process, several GPUs (row f4) --------------------------------------------------------------------------
// Replaces the reference's IO-rank data path — decompose_data_real / _int scatter the global arrays read by rank 0 to
// the MPI ranks and write_io_real / _int gather them back for output (mpp/module_mpp_land.F90:645-857) — by direct tile
// slicing: the process keeps the WHOLE-domain arrays, each GPU owns the tile mpp_land_partition_calc would give it, and
// every context copies its own rows straight out of / into the global arrays (pitched copies, memory extent = domain,
// tile bounds = the GPU's tile).  One worker thread per GPU drives its context, so the tiles' pipelines run concurrently.
struct noahmp_b200_domain {
  int gni = 0, gnj = 0, ntiles = 0;
  std::vector<noahmp_b200_ctx*> tile;
  std::vector<int> xs, xe, ys, ye;
};

noahmp_b200_domain* noahmp_b200_domain_create(const noahmp_tables* tables, int global_ni, int global_nj, int ntiles,
                                              const int* devices) {
  if (!tables || global_ni <= 0 || global_nj <= 0 || ntiles <= 0 || !devices) {
    set_error("bad arguments to noahmp_b200_domain_create");
    return nullptr;
  }
  auto* d = new noahmp_b200_domain();
  d->gni = global_ni; d->gnj = global_nj; d->ntiles = ntiles;
  d->xs.resize(ntiles); d->xe.resize(ntiles); d->ys.resize(ntiles); d->ye.resize(ntiles);
  for (int r = 0; r < ntiles; ++r) {
    noahmp_b200_tile(global_ni, global_nj, ntiles, r, &d->xs[r], &d->xe[r], &d->ys[r], &d->ye[r]);
    noahmp_b200_ctx* c = noahmp_b200_create(devices[r], tables, d->xe[r] - d->xs[r] + 1, d->ye[r] - d->ys[r] + 1);
    if (!c) { noahmp_b200_domain_destroy(d); return nullptr; }
    d->tile.push_back(c);
  }
  return d;
}

void noahmp_b200_domain_destroy(noahmp_b200_domain* d) {
  if (!d) return;
  for (auto* c : d->tile) noahmp_b200_destroy(c);
  delete d;
}

int noahmp_b200_domain_ntiles(const noahmp_b200_domain* d) { return d ? d->ntiles : 0; }
noahmp_b200_ctx* noahmp_b200_domain_tile(noahmp_b200_domain* d, int r) {
  return (d && r >= 0 && r < d->ntiles) ? d->tile[r] : nullptr;
}
int noahmp_b200_domain_tile_bounds(const noahmp_b200_domain* d, int r, int* xs, int* xe, int* ys, int* ye) {
  if (!d || r < 0 || r >= d->ntiles) return NOAHMP_ERR_ARG;
  *xs = d->xs[r]; *xe = d->xe[r]; *ys = d->ys[r]; *ye = d->ye[r];
  return 0;
}

// args of tile r: the caller's global arrays and memory bounds, tile bounds of the GPU's tile
static noahmp_lsm_args domain_tile_args(const noahmp_b200_domain* d, const noahmp_lsm_args* g, int r) {
  noahmp_lsm_args a = *g;
  a.its = g->ims + d->xs[r] - 1; a.ite = g->ims + d->xe[r] - 1;
  a.jts = g->jms + d->ys[r] - 1; a.jte = g->jms + d->ye[r] - 1;
  return a;
}
static int domain_check(const noahmp_b200_domain* d, const noahmp_lsm_args* g) {
  if (!d || !g) return NOAHMP_ERR_ARG;
  if (g->ime - g->ims + 1 != d->gni || g->jme - g->jms + 1 != d->gnj) {
    set_error("domain call: the memory bounds must span the whole domain the domain object was created for");
    return NOAHMP_ERR_ARG;
  }
  return 0;
}

}  // extern "C"
template <class F>
static int domain_for_tiles(noahmp_b200_domain* d, F&& fn) {
  std::vector<int> rc(d->ntiles, 0);
  std::vector<std::string> err(d->ntiles);
  std::vector<std::thread> th;
  for (int r = 0; r < d->ntiles; ++r)
    th.emplace_back([&, r] {
      rc[r] = fn(r);
      if (rc[r] >= NOAHMP_ERR_CUDA) err[r] = g_last_error;  // thread-local message of the worker
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < d->ntiles; ++r)
    if (rc[r] >= NOAHMP_ERR_CUDA) { set_error("tile " + std::to_string(r) + ": " + err[r]); return rc[r]; }
  return 0;
}
extern "C" {

// CALL noahmplsm(...) on the whole domain: `g` describes the global arrays (ims:ime x jms:jme = the domain).
// status: the first failing column in the reference's loop order (j outer, i inner) and the total count.
int noahmp_b200_domain_noahmplsm(noahmp_b200_domain* d, const noahmp_lsm_args* g, noahmp_status* status) {
  int rc = domain_check(d, g);
  if (rc) return rc;
  std::vector<noahmp_status> st(d->ntiles);
  rc = domain_for_tiles(d, [&](int r) {
    noahmp_lsm_args a = domain_tile_args(d, g, r);
    return noahmp_b200_noahmplsm(d->tile[r], &a, &st[r]);
  });
  if (rc) return rc;
  noahmp_status out{0, 0, 0, 0, 0.f};
  for (int r = 0; r < d->ntiles; ++r) {
    if (!st[r].code) continue;
    out.count += st[r].count;
    if (!out.code || st[r].j < out.j || (st[r].j == out.j && st[r].i < out.i)) {
      const int n = out.count;
      out = st[r];
      out.count = n;
    }
  }
  if (status) *status = out;
  return out.code;
}

int noahmp_b200_domain_sync_host(noahmp_b200_domain* d, const noahmp_lsm_args* g) {
  int rc = domain_check(d, g);
  if (rc) return rc;
  return domain_for_tiles(d, [&](int r) {
    noahmp_lsm_args a = domain_tile_args(d, g, r);
    return noahmp_b200_sync_host(d->tile[r], &a);
  });
}

// settings applied to every tile
int noahmp_b200_domain_configure(noahmp_b200_domain* d, int sync_mode, int math_mode, const char* fetch, const char* push,
                                 unsigned hints) {
  if (!d) return NOAHMP_ERR_ARG;
  for (auto* c : d->tile) {
    int rc;
    if ((rc = noahmp_b200_set_mode(c, sync_mode))) return rc;
    if ((rc = noahmp_b200_set_math(c, math_mode))) return rc;
    if (fetch && (rc = noahmp_b200_set_fetch(c, fetch))) return rc;
    if (push && (rc = noahmp_b200_set_push(c, push))) return rc;
    if ((rc = noahmp_b200_set_forcing_hints(c, hints))) return rc;
  }
  return 0;
}

}  // extern "C"
