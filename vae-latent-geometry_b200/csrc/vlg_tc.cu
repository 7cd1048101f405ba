// tcgen05 (TF32) geodesic step kernel -- placeholder until the tensor-core path lands.
#include "vlg_common.cuh"
#include "vlg_kernels.h"

namespace vlg {
size_t tc_workspace_bytes(int, int, int, int) { return 0; }
cudaError_t launch_tc(const StepParams&, bool, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace vlg
