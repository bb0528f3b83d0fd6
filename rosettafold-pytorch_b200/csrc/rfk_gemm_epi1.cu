// rfk_gemm_epi1.cu — tcgen05 GEMM instances with epilogue flavour 1 (see rfk_gemm_device.cuh).
#include "rfk_gemm_device.cuh"
namespace rfk {
int launch_tc_epi1(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles,
                   cudaStream_t s) {
  return launch_tc_bn<1>(bn, ta, tb, p, tiles, s);
}
}  // namespace rfk
