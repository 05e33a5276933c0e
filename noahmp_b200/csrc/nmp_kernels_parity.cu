// nmp_kernels_parity.cu — parity build of the column-physics kernels: portable nmp_math.h transcendentals,
// compiled with -fmad=false so that every fp32 operation rounds exactly as in the CPU oracle (g++
// -ffp-contract=off).  Only the run-time-option instantiation is built here.
#define NMP_PARITY 1
#include "nmp_kernels.cuh"
#include "nmp_groundwater.cuh"
#include "nmp_forcing.cuh"
#include "nmp_init.cuh"

const char* nmp_launch_step_parity(const nmpf::StepParams& base, const nmpf::StepRange& r, cudaStream_t stream,
                                       long long* launches) {
  return launch_step(base, r, stream, launches);
}

void nmp_launch_wtable_parity(const nmpf::WtParams& w, cudaStream_t stream, long long* launches, int phase) {
  launch_wtable(w, stream, launches, phase);
}

void nmp_launch_forcing_parity(const nmpf::ForcingParams& f, cudaStream_t stream, long long* launches) {
  launch_forcing(f, stream, launches);
}

// cold start (always this build: runs once, rounds as the oracle does)
void nmp_launch_init(const nmpf::InitParams& p, cudaStream_t stream, long long* launches) { launch_init(p, stream, launches); }
