// Tensor-core (tcgen05) sparse convolution family - placeholder until the UMMA kernels
// land: reports "unsupported" for every shape so conv_api.cu uses the SIMT family.
#include "common.cuh"

namespace b2s {
bool tc_gather_gemm_supported(int, int) { return false; }
size_t tc_gather_gemm_workspace(int, int, int) { return 0; }
int launch_gather_gemm_tc(const void*, const void*, int, int, int, int, int, const int32_t*, int64_t,
                          const void*, void*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 conv family not built");
  return B2S_ERR_UNSUPPORTED;
}
bool tc_wgrad_supported(int, int) { return false; }
int launch_wgrad_tc(const void*, const void*, const int32_t*, const int32_t*, int64_t, int64_t, int,
                    int, int, int, float*, cudaStream_t) {
  set_error("tcgen05 wgrad family not built");
  return B2S_ERR_UNSUPPORTED;
}
}  // namespace b2s
