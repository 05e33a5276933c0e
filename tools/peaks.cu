// peaks.cu — measured FP32-FMA, MUFU and warp-instruction issue peaks of the GPU (SURVEY.md §8d asks for them:
// MEASURED_PEAKS.json has only HBM and bf16 GEMM).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a peaks.cu -o peaks
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void spin(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // 8 independent FFMA chains
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    } else if (MODE == 1) {  // 8 independent MUFU.EX2
      x0 = exp2f(x0); x1 = exp2f(x1); x2 = exp2f(x2); x3 = exp2f(x3);
      x4 = exp2f(x4); x5 = exp2f(x5); x6 = exp2f(x6); x7 = exp2f(x7);
    } else {  // the physics' mix: 6 FP32 : 1 MUFU
      x0 = fmaf(x0, a, b); x1 = x1 * a; x2 = x2 + b; x3 = fmaf(x3, a, b); x4 = fmaxf(x4 * a, b); x5 = x5 * a + x0;
      x6 = exp2f(x6 * 0.001f);
      x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <int MODE>
double run(const char* name, double ops_per_iter_per_thread) {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int threads = 1024, blocks = p.multiProcessorCount * 2, iters = 1 << 16;
  float* out;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  spin<MODE><<<blocks, threads>>>(out, 1024, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  double best = 0;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    spin<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = ops_per_iter_per_thread * iters * (double)threads * blocks / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  printf("\"%s\": %.4e,\n", name, best);
  cudaFree(out);
  return best;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d,\n", p.name, p.multiProcessorCount);
  run<0>("ffma_thread_instr_per_s", 8);
  run<1>("mufu_ex2_thread_instr_per_s", 8);
  run<2>("mixed_thread_instr_per_s", 8);
  printf("\"note\": \"thread-instructions/s, best of 5, 2x1024 threads per SM, default clocks\"}\n");
  return 0;
}
