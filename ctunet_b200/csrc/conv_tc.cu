// tcgen05 / TMA implicit-GEMM convolution kernels (bf16).  Placeholder until the kernels land.
#include "common.cuh"

namespace ctu {
int conv3d_fprop_tc(const void* const*, const int*, int, const float*, const float*, void*, int, int, int, int, int,
                    int, cudaStream_t) {
    set_error("tensor path not built");
    return CTU_ERR_UNSUPPORTED;
}
int conv3d_wgrad_tc(const void* const*, const int*, int, const void*, float*, float*, int, int, int, int, int, int,
                    cudaStream_t) {
    set_error("tensor path not built");
    return CTU_ERR_UNSUPPORTED;
}
}  // namespace ctu

extern "C" int ctu_has_tensor_path(void) { return 0; }
