// nmp_kernels_fast.cu — production build of the column-physics kernels: libdevice fp32 math, FMA contraction.
#define NMP_PARITY 0
#include "nmp_kernels.cuh"
#include "nmp_groundwater.cuh"

const char* nmp_launch_step_fast(const nmpf::StepParams& base, const nmpf::StepRange& r, cudaStream_t stream,
                                     long long* launches) {
  return launch_step(base, r, stream, launches);
}

void nmp_launch_wtable_fast(const nmpf::WtParams& w, cudaStream_t stream, long long* launches, int phase) {
  launch_wtable(w, stream, launches, phase);
}
