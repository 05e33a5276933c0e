// Host->device rates for the shapes noahmplsm moves: contiguous planes, one level of an (i,k,j) array (strided rows),
// via the copy engine (1 or several streams) and via a zero-copy gather kernel reading the pinned host array.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void gather_level(const float4* __restrict__ src, float4* __restrict__ dst, int ni4, int nk, int lev, int nj) {
  // src (nj, nk, ni4) on the host, dst (nj, ni4) on the device
  const long long total = (long long)nj * ni4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t / ni4), i = (int)(t - (long long)j * ni4);
    dst[t] = __ldcs(src + ((long long)j * nk + lev) * ni4 + i);
  }
}
int main() {
  const int ni = 4608, nj = 3840, nk = 2, NA = 6;
  const size_t plane = (size_t)ni * nj * sizeof(float);
  std::vector<float*> h(NA), d(NA);
  for (int a = 0; a < NA; ++a) { CK(cudaHostAlloc(&h[a], plane * nk, cudaHostAllocMapped)); CK(cudaMalloc(&d[a], plane)); for (size_t k = 0; k < plane * nk / 4; k += 1024) h[a][k] = 1.f; }
  cudaStream_t s[NA]; for (int a = 0; a < NA; ++a) CK(cudaStreamCreateWithFlags(&s[a], cudaStreamNonBlocking));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, int mode, int nstreams, int chunks, int blocks) -> int {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0, s[0]));
      for (int a = 1; a < nstreams; ++a) CK(cudaStreamWaitEvent(s[a], e0, 0));
      for (int c = 0; c < chunks; ++c) {
        const int j0 = nj * c / chunks, j1 = nj * (c + 1) / chunks;
        for (int a = 0; a < NA; ++a) {
          cudaStream_t st = s[a % nstreams];
          if (mode == 0) CK(cudaMemcpyAsync(d[a] + (size_t)j0 * ni, h[a] + (size_t)j0 * ni, (size_t)(j1 - j0) * ni * 4, cudaMemcpyHostToDevice, st));
          if (mode == 1) CK(cudaMemcpy2DAsync(d[a] + (size_t)j0 * ni, ni * 4, h[a] + (size_t)j0 * ni * nk, (size_t)ni * nk * 4, ni * 4, j1 - j0, cudaMemcpyHostToDevice, st));
          if (mode == 2) gather_level<<<blocks, 256, 0, st>>>((const float4*)h[a] + (size_t)j0 * nk * (ni / 4), (float4*)d[a] + (size_t)j0 * (ni / 4), ni / 4, nk, 0, j1 - j0);
        }
      }
      cudaEvent_t ee[NA];
      for (int a = 1; a < nstreams; ++a) { CK(cudaEventCreateWithFlags(&ee[a], cudaEventDisableTiming)); CK(cudaEventRecord(ee[a], s[a])); CK(cudaStreamWaitEvent(s[0], ee[a], 0)); }
      CK(cudaEventRecord(e1, s[0]));
      CK(cudaDeviceSynchronize());
      for (int a = 1; a < nstreams; ++a) cudaEventDestroy(ee[a]);
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("%-44s streams %d chunks %d blocks %4d : %7.3f ms  %6.1f GB/s\n", name, nstreams, chunks, blocks, best, NA * plane / best / 1e6);
    return 0;
  };
  run("contiguous plane", 0, 1, 1, 0); run("contiguous plane", 0, 1, 8, 0);
  run("level 1 of (i,2,j): cudaMemcpy2DAsync", 1, 1, 1, 0); run("level 1 of (i,2,j): cudaMemcpy2DAsync", 1, 1, 8, 0);
  run("level 1 of (i,2,j): cudaMemcpy2DAsync", 1, 2, 8, 0); run("level 1 of (i,2,j): cudaMemcpy2DAsync", 1, 3, 8, 0);
  run("level 1 of (i,2,j): cudaMemcpy2DAsync", 1, 6, 8, 0);
  for (int b : {16, 32, 64, 148, 296, 592}) run("level 1 of (i,2,j): zero-copy gather kernel", 2, 1, 8, b);
  run("level 1 of (i,2,j): zero-copy gather kernel", 2, 3, 8, 32);
  return 0;
}
