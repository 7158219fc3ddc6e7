// K3 (tensor-core variant) -- placeholder until the bf16 kernel lands in this file.
#include "mlp.cuh"
namespace nrc {
int32_t density_mlp_fwd_bf16(cudaStream_t, const nrc_density_mlp_t*, const float*, int64_t, float*, float*,
                             float*) {
  return NRC_E_UNSUPPORTED;
}
}  // namespace nrc
