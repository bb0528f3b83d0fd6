// rfk_gemm_epi3.cu — tcgen05 GEMM instances with the TMA-store epilogue flavour 3.
#include "rfk_gemm_device.cuh"
namespace rfk {
int launch_tc_epi3(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles,
                   cudaStream_t s, const EpiMaps* em) {
  return launch_tc_bn<3>(bn, ta, tb, p, tiles, s, em);
}
}  // namespace rfk
