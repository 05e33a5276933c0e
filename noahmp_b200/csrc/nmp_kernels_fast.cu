// nmp_kernels_fast.cu — production build of the column-physics kernels: libdevice fp32 math, FMA contraction.
#define NMP_PARITY 0
#include "nmp_kernels.cuh"
#include "nmp_groundwater.cuh"
#include "nmp_forcing.cuh"

const char* nmp_launch_step_fast(const nmpf::StepParams& base, const nmpf::StepRange& r, cudaStream_t stream,
                                     long long* launches) {
  return launch_step(base, r, stream, launches);
}

void nmp_launch_wtable_fast(const nmpf::WtParams& w, cudaStream_t stream, long long* launches, int phase) {
  launch_wtable(w, stream, launches, phase);
}

void nmp_launch_forcing_fast(const nmpf::ForcingParams& f, cudaStream_t stream, long long* launches) {
  launch_forcing(f, stream, launches);
}
