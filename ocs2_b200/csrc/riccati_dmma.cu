// riccati_dmma.cu — shape-specialised ILQR sweep (+ fused rollout) kernels. Placeholder until the DMMA/TMA kernel lands.
#include "o2c_common.cuh"

namespace o2c {

bool fast_ilqr_supported(const Layout&, const SolverSettings&, const DeviceBuffers&) { return false; }

cudaError_t launch_ilqr_fast(const Layout&, const SolverSettings&, const DeviceBuffers&, bool, double, int, int, int, cudaStream_t, int*) {
  return cudaErrorNotSupported;
}

}  // namespace o2c
