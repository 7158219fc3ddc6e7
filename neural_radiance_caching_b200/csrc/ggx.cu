// K5 -- placeholder until the GGX integration kernel lands in this file.
#include "nrc_common.cuh"
extern "C" int32_t nrc_ggx_integrate_fwd(void*, const float*, const float*, const float*, const float*,
                                         const float*, const float*, const float*, const float*,
                                         const float*, int64_t, int32_t, int32_t, float, float*, float*) {
  return NRC_E_UNSUPPORTED;
}
