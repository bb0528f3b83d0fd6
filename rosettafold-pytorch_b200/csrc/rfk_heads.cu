// rfk_heads.cu — elementwise pieces of the prediction heads that have no twin in the trunk (reference
// rosettafold_pytorch.py:1130-1172): the symmetrisation of the projected pair map in front of the distance / omega heads,
//   y[b, i, j, :] = 0.5 * (x[b, i, j, :] + x[b, j, i, :])          (:1166)
// on a channels-last map. HBM-bound: every element is read twice (once transposed, in whole channel rows) and written once.
#include "rfk_common.cuh"

namespace rfk {
namespace {

// one thread = one 16-byte chunk (4 fp32 / 8 16-bit channels) of one (b, i, j)
template <int DT>
__global__ void __launch_bounds__(256) pair_symmetrize_kernel(const void* __restrict__ x, void* __restrict__ y, int L, int chunks,
                                                              int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % chunks);
  const int64_t pos = idx / chunks;  // (b, i, j)
  const int j = (int)(pos % L);
  const int64_t bi = pos / L;
  const int i = (int)(bi % L);
  const int64_t b = bi / L;
  const int64_t tpos = (b * L + j) * L + i;
  const uint4 a = reinterpret_cast<const uint4*>(x)[pos * chunks + c];
  const uint4 t = reinterpret_cast<const uint4*>(x)[tpos * chunks + c];
  uint4 o;
  if (DT == RFK_F32) {
    o.x = __float_as_uint(0.5f * (__uint_as_float(a.x) + __uint_as_float(t.x)));
    o.y = __float_as_uint(0.5f * (__uint_as_float(a.y) + __uint_as_float(t.y)));
    o.z = __float_as_uint(0.5f * (__uint_as_float(a.z) + __uint_as_float(t.z)));
    o.w = __float_as_uint(0.5f * (__uint_as_float(a.w) + __uint_as_float(t.w)));
  } else {
    auto avg2 = [](uint32_t p, uint32_t q) {
      const float lo = 0.5f * (h16_to_float((uint16_t)(p & 0xffffu), DT) + h16_to_float((uint16_t)(q & 0xffffu), DT));
      const float hi = 0.5f * (h16_to_float((uint16_t)(p >> 16), DT) + h16_to_float((uint16_t)(q >> 16), DT));
      return pack_h16x2(lo, hi, DT);
    };
    o.x = avg2(a.x, t.x); o.y = avg2(a.y, t.y); o.z = avg2(a.z, t.z); o.w = avg2(a.w, t.w);
  }
  reinterpret_cast<uint4*>(y)[pos * chunks + c] = o;
}

}  // namespace
}  // namespace rfk

using namespace rfk;

extern "C" int rfk_pair_symmetrize(const void* x, void* y, int dtype, int B, int L, int C, rfk_stream_t stream_) {
  if (!x || !y) return RFK_ERR_NULL_POINTER;
  if (x == y) return RFK_ERR_UNSUPPORTED;  // the transposed read would race with the write
  if (B <= 0 || L <= 0 || C <= 0) return RFK_ERR_BAD_DIMS;
  if (dtype != RFK_F32 && !is_h16(dtype)) return RFK_ERR_BAD_DTYPE;
  const int per = dtype == RFK_F32 ? 4 : 8;
  if (C % per) return RFK_ERR_BAD_DIMS;
  if (!aligned16(x) || !aligned16(y)) return RFK_ERR_MISALIGNED;
  const int chunks = C / per;
  const int64_t total = (int64_t)B * L * L * chunks;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (dtype == RFK_F32) pair_symmetrize_kernel<RFK_F32><<<blocks, 256, 0, st>>>(x, y, L, chunks, total);
  else if (dtype == RFK_BF16) pair_symmetrize_kernel<RFK_BF16><<<blocks, 256, 0, st>>>(x, y, L, chunks, total);
  else pair_symmetrize_kernel<RFK_F16><<<blocks, 256, 0, st>>>(x, y, L, chunks, total);
  return post_launch();
}
