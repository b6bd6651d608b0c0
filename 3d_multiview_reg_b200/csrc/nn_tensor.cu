// Stage 1 (tensor-core path) -- placeholder translation unit until the tcgen05 kernel lands.
#include "common.cuh"
namespace lmpcr {
size_t nn_tensor_workspace_bytes(int, int, int, int, int, int) { return 0; }
int launch_nn_tensor(const float*, int, int, const float*, int, int, int, const int32_t*, int, int32_t*, float*, void*, size_t,
                     cudaStream_t) {
  set_error("lmpcr_nn_argmin: LMPCR_NN_TENSOR is not built yet");
  return LMPCR_ERR_UNSUPPORTED;
}
}  // namespace lmpcr
