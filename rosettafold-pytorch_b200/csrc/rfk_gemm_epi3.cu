// rfk_gemm_epi3.cu — tcgen05 GEMM instances with the TMA-store epilogue flavour 3.
#include "rfk_gemm_device.cuh"
namespace rfk {
int launch_tc_epi3(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles,
                   cudaStream_t s, const EpiMaps* em) {
  return launch_tc_bn<3>(bn, ta, tb, p, tiles, s, em);
}
// CTA-pair (cta_group::2) instances: bf16 TMA-store epilogue, wide column blocks
int launch_tc_pair_epi3(int bn, const CUtensorMap& ta, const CUtensorMap& tb_half, const GemmDev& p, int64_t tiles,
                        cudaStream_t s, const EpiMaps* em) {
  switch (bn) {
    case 256: return launch_tc_pair<256, 3>(ta, tb_half, p, tiles, s, em);
    case 192: return launch_tc_pair<192, 3>(ta, tb_half, p, tiles, s, em);
    case 128: return launch_tc_pair<128, 3>(ta, tb_half, p, tiles, s, em);
    default: return RFK_ERR_UNSUPPORTED;
  }
}
// B-stationary instances: short-K projections with bf16 output (K <= 320: five k-blocks, K <= 384: six)
int launch_tc_bstat_epi3(int bn, int kb, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles,
                         cudaStream_t s, const EpiMaps* em) {
  if (bn == 192 && kb == 5) return launch_tc_bstat<192, 5>(ta, tb, p, tiles, s, em);
  if (bn == 128 && kb == 5) return launch_tc_bstat<128, 5>(ta, tb, p, tiles, s, em);
  if (bn == 192 && kb == 6) return launch_tc_bstat<192, 6>(ta, tb, p, tiles, s, em);
  if (bn == 128 && kb == 6) return launch_tc_bstat<128, 6>(ta, tb, p, tiles, s, em);
  return RFK_ERR_UNSUPPORTED;
}
}  // namespace rfk
