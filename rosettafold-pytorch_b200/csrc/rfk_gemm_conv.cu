// rfk_gemm_conv.cu — 3x3 'same' convolution on a channels-last map as an implicit GEMM on the
// tcgen05 GEMM kernel (CONV instances): nn.Conv2d(d_pair, d_pair, 3, padding="same", bias=False) of
// PairUpdateWithMsa (rosettafold_pytorch.py:451-457). The im2col matrix is never built: each of the
// 9 taps is a TMA box shifted by (di, dj) over the NHWC tensor, with the border zero-filled by TMA.
#include <cstdlib>
#include "rfk_gemm_device.cuh"

namespace rfk {
int launch_tc_conv(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                   int64_t tiles, cudaStream_t s) {
  if (epi == 0) return launch_tc<32, 0, true>(ta, tb, p, tiles, s);
  if (epi == 1) {
    switch (bn) {
      case 256: return launch_tc<256, 1, true>(ta, tb, p, tiles, s);
      case 128: return launch_tc<128, 1, true>(ta, tb, p, tiles, s);
      case 96: return launch_tc<96, 1, true>(ta, tb, p, tiles, s);
      case 64: return launch_tc<64, 1, true>(ta, tb, p, tiles, s);
      default: return launch_tc<32, 1, true>(ta, tb, p, tiles, s);
    }
  }
  switch (bn) {
    case 256: return launch_tc<256, 2, true>(ta, tb, p, tiles, s);
    case 128: return launch_tc<128, 2, true>(ta, tb, p, tiles, s);
    case 96: return launch_tc<96, 2, true>(ta, tb, p, tiles, s);
    case 64: return launch_tc<64, 2, true>(ta, tb, p, tiles, s);
    default: return launch_tc<32, 2, true>(ta, tb, p, tiles, s);
  }
}
}  // namespace rfk

using namespace rfk;

extern "C" int rfk_conv3x3_nhwc_dil(const void* x, int x_dtype, const void* w_packed, void* y, int y_dtype, int B, int H,
                                    int L, int C, int Cout, int dilation, rfk_stream_t stream_) {
  if (!x || !w_packed || !y) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || H <= 0 || L <= 0 || C <= 0 || Cout <= 0 || dilation < 1 || dilation > 64) return RFK_ERR_BAD_DIMS;
  if (!is_h16(x_dtype)) return RFK_ERR_BAD_DTYPE;
  if (C % 8) return RFK_ERR_BAD_DIMS;  // TMA stride rule (16-byte rows)
  if (!is_h16(y_dtype) && y_dtype != RFK_F32) return RFK_ERR_BAD_DTYPE;
  if (!aligned16(x) || !aligned16(w_packed) || !aligned16(y)) return RFK_ERR_MISALIGNED;
  int rc = check_arch();
  if (rc != RFK_OK) return rc;
  const int cblocks = (C + 63) / 64, cpad = cblocks * 64;
  const int Lp = (L + kBlockM - 1) / kBlockM * kBlockM;
  int bn = 32;
  for (int cand : {256, 128, 96, 64, 32})
    if (Cout % cand == 0) { bn = cand; break; }
  static const char* force_bn = getenv("RFK_CONV_BN");  // A/B debugging aid: 256 / 128 / 96 / 64 / 32
  if (force_bn && atoi(force_bn) >= 32 && Cout % 32 == 0) bn = atoi(force_bn);
  const bool lean = Cout % 32 == 0 && (is_h16(y_dtype) ? Cout % 8 == 0 : Cout % 4 == 0);

  GemmDev p{};
  p.M = (int64_t)B * H * Lp;  // padded so that every tile lies inside one image row
  p.N = Cout;
  p.K = (int64_t)9 * cpad;
  p.Z0 = p.Z1 = p.Z2 = 1;
  p.MR = Lp; p.NR = Cout;
  p.alpha = 1.f; p.act = RFK_ACT_NONE; p.epi = RFK_EPI_STD;
  p.c = y; p.c_dtype = y_dtype;
  p.c_addr.ms[0] = Cout;                 // j
  p.c_addr.ms[1] = (int64_t)L * Cout;    // (b, i)
  p.c_addr.ns[0] = 1; p.c_addr.ns[1] = 0;
  p.conv_L = L; p.conv_H = H; p.conv_Lp = Lp; p.conv_cblocks = cblocks; p.conv_cpad = cpad;
  p.conv_dil = dilation;
  p.ab_f16 = x_dtype == RFK_F16;  // image and packed weights are IEEE half instead of bf16
  p.conv_last_k16 = (C - (cblocks - 1) * 64 + 15) / 16;

  CUtensorMap ta, tb;
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)L, (uint64_t)H, (uint64_t)B, 1};
    const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)L * C * 2, (uint64_t)H * L * C * 2,
                                 (uint64_t)B * H * L * C * 2};
    const uint32_t box[5] = {64, (uint32_t)kBlockM, 1, 1, 1};
    if ((rc = make_tmap_bf16_raw(&ta, x, 5, dims, strides, box)) != RFK_OK) return rc;
  }
  {
    const uint64_t dims[5] = {(uint64_t)9 * cpad, (uint64_t)Cout, 1, 1, 1};
    const uint64_t strides[4] = {(uint64_t)9 * cpad * 2, (uint64_t)9 * cpad * 2 * Cout,
                                 (uint64_t)9 * cpad * 2 * Cout, (uint64_t)9 * cpad * 2 * Cout};
    const uint32_t box[5] = {64, (uint32_t)bn, 1, 1, 1};
    if ((rc = make_tmap_bf16_raw(&tb, w_packed, 5, dims, strides, box)) != RFK_OK) return rc;
  }
  const int64_t tiles = (p.M / kBlockM) * ((Cout + bn - 1) / bn);
  return launch_tc_conv(bn, !lean ? 0 : (is_h16(y_dtype) ? 1 : 2), ta, tb, p, tiles,
                        reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int rfk_conv3x3_nhwc_hw(const void* x, const void* w_packed, void* y, int y_dtype, int B, int H,
                                   int L, int C, int Cout, rfk_stream_t stream_) {
  return rfk_conv3x3_nhwc_dil(x, RFK_BF16, w_packed, y, y_dtype, B, H, L, C, Cout, 1, stream_);
}

extern "C" int rfk_conv3x3_nhwc(const void* x, const void* w_packed, void* y, int y_dtype, int B, int L,
                                int C, int Cout, rfk_stream_t stream_) {
  return rfk_conv3x3_nhwc_hw(x, w_packed, y, y_dtype, B, L, L, C, Cout, stream_);
}

// ----------------------------------------------------------------------------------------------
// fp32 validation mode: direct 3x3 convolution on CUDA cores (reference :452, :456 in fp32).
// Block = 32 consecutive positions of one image row x 64 output channels; the input window
// (3 rows x 34 columns x 32 channels) and one tap's weights (32 x 64) are staged in shared memory;
// thread (pg, cg) accumulates positions {2 pg, 2 pg + 1} x channels [4 cg, 4 cg + 4). Fixed summation
// order (channel chunks, taps, channels): run-to-run reproducible.
// ----------------------------------------------------------------------------------------------
namespace rfk {
namespace {
constexpr int kCvW = 32, kCvCo = 64, kCvCk = 32, kCvMaxDil = 8;

__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          float* __restrict__ y, int B, int H, int W, int Cin, int Cout,
                                                          int co_blocks, int dil) {
  __shared__ float xs[3][kCvW + 2 * kCvMaxDil][kCvCk];  // window columns j0 - dil .. j0 + 31 + dil
  __shared__ __align__(16) float ws[kCvCk][kCvCo];
  const int j0 = blockIdx.x * kCvW, i = blockIdx.y;
  const int b = blockIdx.z / co_blocks, co0 = (blockIdx.z % co_blocks) * kCvCo;
  const int tid = threadIdx.x, pg = tid >> 4, cg = tid & 15;
  float acc[2][4] = {};
  for (int c0 = 0; c0 < Cin; c0 += kCvCk) {
    const int wcols = kCvW + 2 * dil;
    for (int e = tid; e < 3 * wcols * kCvCk; e += 256) {
      const int ck = e % kCvCk, col = (e / kCvCk) % wcols, r = e / (kCvCk * wcols);
      const int ii = i + (r - 1) * dil, jj = j0 + col - dil, c = c0 + ck;
      float v = 0.f;
      if (ii >= 0 && ii < H && jj >= 0 && jj < W && c < Cin) v = x[(((int64_t)b * H + ii) * W + jj) * Cin + c];
      xs[r][col][ck] = v;
    }
    for (int tap = 0; tap < 9; ++tap) {
      __syncthreads();  // xs complete (first tap) / previous tap's ws consumed
      for (int e = tid; e < kCvCk * kCvCo; e += 256) {
        const int co = e % kCvCo, ck = e / kCvCo;
        const int c = c0 + ck, o = co0 + co;
        ws[ck][co] = (c < Cin && o < Cout) ? w[((int64_t)tap * Cin + c) * Cout + o] : 0.f;
      }
      __syncthreads();
      const int di = tap / 3, dj = tap % 3;
#pragma unroll 8
      for (int ck = 0; ck < kCvCk; ++ck) {
        const float x0 = xs[di][2 * pg + dj * dil][ck], x1 = xs[di][2 * pg + 1 + dj * dil][ck];
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[ck][4 * cg]);
        acc[0][0] = fmaf(x0, w4.x, acc[0][0]); acc[0][1] = fmaf(x0, w4.y, acc[0][1]);
        acc[0][2] = fmaf(x0, w4.z, acc[0][2]); acc[0][3] = fmaf(x0, w4.w, acc[0][3]);
        acc[1][0] = fmaf(x1, w4.x, acc[1][0]); acc[1][1] = fmaf(x1, w4.y, acc[1][1]);
        acc[1][2] = fmaf(x1, w4.z, acc[1][2]); acc[1][3] = fmaf(x1, w4.w, acc[1][3]);
      }
    }
    __syncthreads();  // before the next channel chunk overwrites xs
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int j = j0 + 2 * pg + q;
    if (j >= W) continue;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int o = co0 + 4 * cg + r;
      if (o < Cout) y[(((int64_t)b * H + i) * W + j) * Cout + o] = acc[q][r];
    }
  }
}
}  // namespace
}  // namespace rfk

extern "C" int rfk_conv3x3_nhwc_f32_dil(const float* x, const float* w_packed, float* y, int B, int H, int L, int C,
                                        int Cout, int dilation, rfk_stream_t stream_) {
  if (!x || !w_packed || !y) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || H <= 0 || L <= 0 || C <= 0 || Cout <= 0 || H > 65535) return RFK_ERR_BAD_DIMS;
  if (dilation < 1 || dilation > rfk::kCvMaxDil) return RFK_ERR_BAD_DIMS;
  const int co_blocks = (Cout + rfk::kCvCo - 1) / rfk::kCvCo;
  if ((int64_t)B * co_blocks > 65535) return RFK_ERR_BAD_DIMS;
  if (rfk::check_arch() != RFK_OK) return RFK_ERR_UNSUPPORTED_ARCH;
  dim3 grid((unsigned)((L + rfk::kCvW - 1) / rfk::kCvW), (unsigned)H, (unsigned)(B * co_blocks));
  rfk::conv3x3_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, w_packed, y, B, H, L, C, Cout, co_blocks,
                                                                                     dilation);
  return rfk::post_launch();
}

extern "C" int rfk_conv3x3_nhwc_f32(const float* x, const float* w_packed, float* y, int B, int H, int L, int C, int Cout,
                                    rfk_stream_t stream_) {
  return rfk_conv3x3_nhwc_f32_dil(x, w_packed, y, B, H, L, C, Cout, 1, stream_);
}
