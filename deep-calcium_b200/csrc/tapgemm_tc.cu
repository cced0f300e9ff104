// bf16 tcgen05/TMEM/TMA tap-GEMM kernels (placeholder until the first kernel lands).
#include "common.cuh"
#include "tapgeom.h"

namespace dcb {

int run_tc_fwd(const TapGeom&, const void*, int, const void*, int, const void*, int, void*, const float*, const float*,
               int, cudaStream_t) {
  return fail(DCB_ERR_UNSUPPORTED, "bf16 tcgen05 forward tap-GEMM not built yet");
}
int run_tc_wgrad(const TapGeom&, const void*, int, const void*, int, const void*, int, float*, void*, size_t,
                 cudaStream_t) {
  return fail(DCB_ERR_UNSUPPORTED, "bf16 tcgen05 wgrad tap-GEMM not built yet");
}
size_t tc_wgrad_workspace(const TapGeom&, int, int) { return 0; }

}  // namespace dcb
