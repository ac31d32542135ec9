// wgrad_tc.cu -- weight gradient of a stride-1 3x3 / 1x1 convolution on the tensor cores (placeholder: the CUDA-core
// reduction k_conv_wgrad_g in train_ops.cu is used until this returns ESR_OK).
#include "tc_common.cuh"
#include "net.cuh"

namespace esr {
struct Bump;
int wgrad_tc(const float *, const __nv_bfloat16 *, int, int, int, int, int, int, float *, Bump &, cudaStream_t) { return ESR_EINVAL; }
} // namespace esr
