// rfk_favor_tc.cu — tcgen05 FAVOR+ kernel (placeholder until the GEMM path is validated on HW).
#include "rfk_common.cuh"
namespace rfk {
int favor_tc_launch(const rfk_favor_desc*, cudaStream_t) { return RFK_ERR_UNSUPPORTED; }
}  // namespace rfk
