// nmp_kernels_fast.cu — production build of the column-physics kernels: libdevice fp32 math, FMA contraction.
#define NMP_PARITY 0
#include "nmp_kernels.cuh"

const char* nmp_launch_step_fast(const nmpf::StepParams& base, int nland, int nglac, cudaStream_t stream,
                                 long long* launches) {
  return launch_step(base, nland, nglac, stream, launches);
}
