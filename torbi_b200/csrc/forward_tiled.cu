// placeholder until the batch-tiled kernel lands
#include "common.cuh"
#include "kernels.cuh"
namespace torbi_b200 {
bool forward_tiled_supported(const ForwardArgs &) { return false; }
cudaError_t launch_forward_tiled(const ForwardArgs &, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace torbi_b200
