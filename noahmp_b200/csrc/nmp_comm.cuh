// nmp_comm.cuh — the two exchanges of the path, inside the C library (included by nmp_lib.cu):
//   * the one-cell halo of the LATERALFLOW pass-1 planes KCELL / HEAD between neighbouring tiles (opt_run = 5,
//     phys/module_sf_noahmp_groundwater.F90:201-295), as grouped ncclSend / ncclRecv over NVLink;
//   * the global water / energy budget: eight fp64 sums per tile, one ncclAllReduce.
// One process per GPU; the host program (Fortran + MPI, or torchrun) only carries the 128-byte ncclUniqueId from
// rank 0 to the others.  NCCL is bound at run time with dlopen("libnccl.so.2"): a single-GPU host needs no NCCL.
//
// The reference's MPI build exchanges nothing here (its mpp_land_com* halo routines are never called for WTD and
// every rank passes ids = its), so its answer depends on the rank count; these tiles reproduce the
// sequential single-domain result for any process grid.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("NOAHMP_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      bool ok = true;
      auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) ok = false; return p; };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    }
  }
  return api.handle ? &api : nullptr;
}

#define NK(call)                                                                                     \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess) {                                                                         \
      set_error(std::string(#call) + ": " + nccl_api()->GetErrorString(r_));                         \
      return NOAHMP_ERR_CUDA;                                                                        \
    }                                                                                                \
  } while (0)

struct NmpComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  int left = -1, right = -1, down = -1, up = -1;  // ranks of the neighbouring tiles in the mpp_land process grid
  float *d_send = nullptr, *d_recv = nullptr;     // [side 0 = left, 1 = right][plane 0 = KCELL, 1 = HEAD][nj]
  double* d_budget_sum = nullptr;                 // result of the all-reduce
};

// columns 1 / ni of the interior rows -> contiguous send buffers; received columns -> ring columns 0 / ni+1
__global__ void halo_pack_kernel(const float* __restrict__ kcell, const float* __restrict__ head, float* __restrict__ send,
                                 int ni, int nj) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  const long long row = (long long)(j + 1) * (ni + 2);
  send[0 * nj + j] = kcell[row + 1];
  send[1 * nj + j] = head[row + 1];
  send[2 * nj + j] = kcell[row + ni];
  send[3 * nj + j] = head[row + ni];
}
__global__ void halo_unpack_kernel(float* __restrict__ kcell, float* __restrict__ head, const float* __restrict__ recv,
                                   int ni, int nj, int have_left, int have_right) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nj) return;
  const long long row = (long long)(j + 1) * (ni + 2);
  if (have_left) { kcell[row] = recv[0 * nj + j]; head[row] = recv[1 * nj + j]; }
  if (have_right) { kcell[row + ni + 1] = recv[2 * nj + j]; head[row + ni + 1] = recv[3 * nj + j]; }
}

// Two phases on stream `s`: left/right columns of the interior rows (packed), then the first / last interior ROW
// including the ring columns just received, which carries the four corners without diagonal messages.  Rows are
// contiguous in the (nj+2) x (ni+2) planes and go out in place.
static int halo_exchange(NmpComm& C, float* kcell, float* head, int ni, int nj, cudaStream_t s, long long* launches) {
  NcclApi* N = nccl_api();
  if (!N || !C.comm) { set_error("halo exchange without an initialised communicator"); return NOAHMP_ERR_ARG; }
  const int T = 256;
  if (C.left >= 0 || C.right >= 0) {
    halo_pack_kernel<<<(nj + T - 1) / T, T, 0, s>>>(kcell, head, C.d_send, ni, nj);
    NK(N->GroupStart());
    if (C.left >= 0) {
      NK(N->Send(C.d_send, 2 * (size_t)nj, ncclFloat32, C.left, C.comm, s));
      NK(N->Recv(C.d_recv, 2 * (size_t)nj, ncclFloat32, C.left, C.comm, s));
    }
    if (C.right >= 0) {
      NK(N->Send(C.d_send + 2 * (size_t)nj, 2 * (size_t)nj, ncclFloat32, C.right, C.comm, s));
      NK(N->Recv(C.d_recv + 2 * (size_t)nj, 2 * (size_t)nj, ncclFloat32, C.right, C.comm, s));
    }
    NK(N->GroupEnd());
    halo_unpack_kernel<<<(nj + T - 1) / T, T, 0, s>>>(kcell, head, C.d_recv, ni, nj, C.left >= 0, C.right >= 0);
    *launches += 2;
  }
  if (C.down >= 0 || C.up >= 0) {
    const size_t P = (size_t)ni + 2;
    NK(N->GroupStart());
    for (float* pl : {kcell, head}) {
      if (C.down >= 0) {
        NK(N->Send(pl + P, P, ncclFloat32, C.down, C.comm, s));
        NK(N->Recv(pl, P, ncclFloat32, C.down, C.comm, s));
      }
      if (C.up >= 0) {
        NK(N->Send(pl + P * nj, P, ncclFloat32, C.up, C.comm, s));
        NK(N->Recv(pl + P * (nj + 1), P, ncclFloat32, C.up, C.comm, s));
      }
    }
    NK(N->GroupEnd());
  }
  CK(cudaGetLastError());
  return 0;
}

// ---- global water / energy budget (north_star: "NCCL ... for global water/energy-balance diagnostics") --------------
// Per tile, in fp64, from the planes the step leaves in HBM (land + glacier columns):
//   [0] storage   canopy water + SWE + aquifer WA + sum_k SMOIS_k * DZS_k * 1000   [mm]   (instantaneous)
//   [1] precip    RAINBL                                                           [mm]   accumulated over the steps
//   [2] et        (ECAN + EDIR + ETRAN) * DT                                        [mm]   accumulated
//   [3] runoff    (RUNSF + RUNSB) * DT                                              [mm]   accumulated
//   [4] erreng    SAV + SAG - (FIRA + HFX + LH + GRDFLX): the residual ERROR tests against 0.01 W/m2
//                 (noahmplsm.F90:1188-1199)                                         [W/m2] accumulated
//   [5] swe       SNOW                                                              [mm]   (instantaneous)
//   [6] columns   number of land + glacier columns
//   [7] steps     number of steps accumulated
// Over a reporting interval the global water residual is  d(storage) - (precip - et - runoff)  = sum of the
// per-column ERRWAT (noahmplsm.F90:1204-1226), glacier columns included through SWE only (their soil is ice).
constexpr int NBUDGET = 8;
__global__ void budget_kernel(const float* __restrict__ state, const float* __restrict__ rainbl, const int* __restrict__ cell,
                              long long np, int nland, int ncol, float dt, float dz0, float dz1, float dz2, float dz3,
                              double* __restrict__ acc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double v[6] = {0., 0., 0., 0., 0., 0.};
  if (n < ncol) {
    auto S = [&](int slot) { return (double)state[(long long)slot * np + n]; };
    const bool land = n < nland;
    const double swe = S(NMP_SLOT(snow));
    double sto = swe;
    if (land) {
      sto += S(NMP_SLOT(canwat)) + S(NMP_SLOT(waxy));
      const float dz[4] = {dz0, dz1, dz2, dz3};
#pragma unroll
      for (int k = 0; k < NOAHMP_NSOIL; ++k) sto += S(NMP_SLOT(smois) + k) * (double)dz[k] * 1000.0;
    }
    v[0] = sto;
    v[1] = (double)rainbl[cell[n]];
    v[2] = (S(NMP_SLOT(ecanxy)) + S(NMP_SLOT(edirxy)) + S(NMP_SLOT(etranxy))) * (double)dt;
    v[3] = (S(NMP_SLOT(runsfxy)) + S(NMP_SLOT(runsbxy))) * (double)dt;
    const double sav = land ? S(NMP_SLOT(savxy)) : 0.0;
    v[4] = sav + S(NMP_SLOT(sagxy)) - (S(NMP_SLOT(firaxy)) + S(NMP_SLOT(hfx)) + S(NMP_SLOT(lh)) + S(NMP_SLOT(grdflx)));
    v[5] = swe;
  }
  __shared__ double red[6][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double x = v[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) red[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double x = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += red[threadIdx.x][w];
    // [0] and [5] are instantaneous: the caller zeroes them before every launch
    atomicAdd(acc + threadIdx.x, x);
  }
}
