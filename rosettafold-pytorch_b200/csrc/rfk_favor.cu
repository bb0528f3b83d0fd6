// rfk_favor.cu — Performer FAVOR+ attention (performer_pytorch.FastAttention, non-causal), fused.
//   * SIMT fp32 kernel (validation mode, also accepts bf16 I/O): one CTA per (group, head);
//     projection matrix, context (m x 64) and key-sum stay in shared memory / registers.
//   * tcgen05 kernel (bf16): rfk_favor_tc.cu.
// Arithmetic (performer-pytorch 1.1.4):
//   dn = 64^-1/4, u = (dn*x) . proj^T
//   softmax kernel: diag = |x|^2/2 * dn^2, ratio = m^-1/2
//       q' = ratio*(exp(u - diag - max_m u) + 1e-4)     k' = ratio*(exp(u - diag - max_{n,m} u) + 1e-4)
//   relu kernel:    q' = relu(u) + 1e-3, k' likewise
//   ksum = sum_n k';  ctx = k'^T v;  out = (q' ctx) / (q' . ksum)
#include <cstdlib>

#include "rfk_common.cuh"

namespace rfk {

constexpr int kFavorDh = 64;
constexpr int kFavorMaxM = 272;
constexpr int kTokTile = 32;
constexpr int kFavorRegs = kFavorMaxM / 4;  // 68 context rows per thread

struct FavorDev {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  const float* proj;
  int dt, kind, m, heads;
  int64_t tokens, G0, G1;
  int64_t gs0, gs1, ts, ogs0, ogs1, ots;
};

__global__ void __launch_bounds__(256, 1) favor_simt_kernel(const FavorDev p) {
  extern __shared__ float sm[];
  float* proj = sm;                                   // [MaxM][65]
  float* ctx = proj + kFavorMaxM * 65;                // [MaxM][64]
  float* ksum = ctx + kFavorMaxM * 64;                // [MaxM]
  float* xs = ksum + kFavorMaxM;                      // [32][64]
  float* vs = xs + kTokTile * 64;                     // [32][64]
  float* feat = vs + kTokTile * 64;                   // [32][MaxM]
  float* red = feat + kTokTile * kFavorMaxM;          // [256]
  const int tid = threadIdx.x;
  const int m = p.m;
  const int h = blockIdx.x % p.heads;
  const int64_t g = blockIdx.x / p.heads;
  const int64_t g0 = g % p.G0, g1 = g / p.G0;
  const int64_t in_base = g1 * p.gs1 + g0 * p.gs0 + (int64_t)h * kFavorDh;
  const int64_t out_base = g1 * p.ogs1 + g0 * p.ogs0 + (int64_t)h * kFavorDh;
  const float dn = rsqrtf(sqrtf((float)kFavorDh));  // 64^-1/4
  const float ratio = rsqrtf((float)m);
  const bool softmax_kind = p.kind == 0;

  for (int i = tid; i < kFavorMaxM * kFavorDh; i += 256) {
    const int r = i / kFavorDh, c = i % kFavorDh;
    proj[r * 65 + c] = r < m ? p.proj[(int64_t)r * kFavorDh + c] : 0.f;
  }
  __syncthreads();

  const int t_loc = tid >> 3;  // token within tile (feature phase)
  const int m_lo = tid & 7;

  auto load_tile = [&](const void* src, float* dst, int64_t t0) {
    for (int i = tid; i < kTokTile * kFavorDh; i += 256) {
      const int t = i / kFavorDh, c = i % kFavorDh;
      dst[i] = (t0 + t < p.tokens) ? load_as_float(src, p.dt, in_base + (t0 + t) * p.ts + c) : 0.f;
    }
  };
  auto dot_u = [&](int mm) {
    float acc = 0.f;
#pragma unroll 16
    for (int d = 0; d < kFavorDh; ++d) acc = fmaf(xs[t_loc * 64 + d], proj[mm * 65 + d], acc);
    return acc * dn;
  };

  // ---- phase 1: global key max (softmax kernel only) ----
  float gmax = 0.f;
  if (softmax_kind) {
    float mx = -INFINITY;
    for (int64_t t0 = 0; t0 < p.tokens; t0 += kTokTile) {
      __syncthreads();
      load_tile(p.k, xs, t0);
      __syncthreads();
      if (t0 + t_loc < p.tokens)
        for (int mm = m_lo; mm < m; mm += 8) mx = fmaxf(mx, dot_u(mm));
    }
    mx = warp_max(mx);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    gmax = red[0];
    for (int i = 1; i < 8; ++i) gmax = fmaxf(gmax, red[i]);
  }

  // ---- phase 2: context and key-sum ----
  float cacc[kFavorRegs];
#pragma unroll
  for (int r = 0; r < kFavorRegs; ++r) cacc[r] = 0.f;
  float ks0 = 0.f, ks1 = 0.f;
  const int cd = tid & 63, cm = tid >> 6;
  for (int64_t t0 = 0; t0 < p.tokens; t0 += kTokTile) {
    __syncthreads();
    load_tile(p.k, xs, t0);
    load_tile(p.v, vs, t0);
    __syncthreads();
    {
      const bool valid = t0 + t_loc < p.tokens;
      float diag = 0.f;
      if (softmax_kind) {
        for (int d = 0; d < kFavorDh; ++d) diag = fmaf(xs[t_loc * 64 + d], xs[t_loc * 64 + d], diag);
        diag *= 0.5f * dn * dn;
      }
      for (int mm = m_lo; mm < kFavorMaxM; mm += 8) {
        float f = 0.f;
        if (valid && mm < m) {
          const float u = dot_u(mm);
          f = softmax_kind ? ratio * (expf(u - diag - gmax) + 1e-4f) : fmaxf(u, 0.f) + 1e-3f;
        }
        feat[t_loc * kFavorMaxM + mm] = f;
      }
    }
    __syncthreads();
    for (int t = 0; t < kTokTile; ++t) {
      const float vv = vs[t * 64 + cd];
#pragma unroll
      for (int r = 0; r < kFavorRegs; ++r) cacc[r] = fmaf(feat[t * kFavorMaxM + cm + 4 * r], vv, cacc[r]);
    }
    for (int t = 0; t < kTokTile; ++t) {
      ks0 += feat[t * kFavorMaxM + tid];
      if (tid + 256 < kFavorMaxM) ks1 += feat[t * kFavorMaxM + tid + 256];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kFavorRegs; ++r) ctx[(cm + 4 * r) * 64 + cd] = cacc[r];
  ksum[tid] = ks0;
  if (tid + 256 < kFavorMaxM) ksum[tid + 256] = ks1;
  __syncthreads();

  // ---- phase 3: queries ----
  for (int64_t t0 = 0; t0 < p.tokens; t0 += kTokTile) {
    __syncthreads();
    load_tile(p.q, xs, t0);
    __syncthreads();
    {
      float diag = 0.f, mx = -INFINITY;
      if (softmax_kind) {
        for (int d = 0; d < kFavorDh; ++d) diag = fmaf(xs[t_loc * 64 + d], xs[t_loc * 64 + d], diag);
        diag *= 0.5f * dn * dn;
      }
      for (int mm = m_lo; mm < kFavorMaxM; mm += 8) {
        float u = 0.f;
        if (mm < m) {
          u = dot_u(mm);
          mx = fmaxf(mx, u);
        }
        feat[t_loc * kFavorMaxM + mm] = u;
      }
      // the 8 threads of one token are 8 consecutive lanes
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      for (int mm = m_lo; mm < kFavorMaxM; mm += 8) {
        float f = 0.f;
        if (mm < m) {
          const float u = feat[t_loc * kFavorMaxM + mm];
          f = softmax_kind ? ratio * (expf(u - diag - mx) + 1e-4f) : fmaxf(u, 0.f) + 1e-3f;
        }
        feat[t_loc * kFavorMaxM + mm] = f;
      }
    }
    __syncthreads();
    {
      const int d0 = (tid & 7) * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      float den = 0.f;
      for (int mm = 0; mm < m; ++mm) {
        const float f = feat[t_loc * kFavorMaxM + mm];
        den = fmaf(f, ksum[mm], den);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(f, ctx[mm * 64 + d0 + j], o[j]);
      }
      if (t0 + t_loc < p.tokens) {
        const float inv = 1.f / den;
        const int64_t ob = out_base + (t0 + t_loc) * p.ots + d0;
#pragma unroll
        for (int j = 0; j < 8; ++j) store_from_float(p.out, p.dt, ob + j, o[j] * inv);
      }
    }
  }
}

int favor_tc_launch(const rfk_favor_desc* d, cudaStream_t stream);  // rfk_favor_tc.cu (round-1 kernel: features via shared memory)
int favor_tm_launch(const rfk_favor_desc* d, cudaStream_t stream);  // rfk_favor_tm.cu (features kept in tensor memory)
int favor_col_launch(const rfk_favor_desc* d, cudaStream_t stream);  // rfk_favor_col.cu (softmax kernel, tokens <= 128)

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_favor_attention(const rfk_favor_desc* d, rfk_stream_t stream_) {
  if (!d) return RFK_ERR_NULL_POINTER;
  if (!d->q || !d->k || !d->v || !d->out || !d->proj) return RFK_ERR_NULL_POINTER;
  if (d->tokens <= 0 || d->G[0] <= 0 || d->G[1] <= 0 || d->heads <= 0) return RFK_ERR_BAD_DIMS;
  if (d->m_features <= 0 || d->m_features > kFavorMaxM) return RFK_ERR_BAD_DIMS;
  if (d->kind != 0 && d->kind != 1) return RFK_ERR_UNSUPPORTED;
  if (d->io_dtype != RFK_F32 && !is_h16(d->io_dtype)) return RFK_ERR_BAD_DTYPE;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static const bool force_simt = getenv("RFK_FAVOR_FORCE_SIMT") != nullptr;  // A/B debugging aid
  if (is_h16(d->io_dtype) && !force_simt) {
    static const bool old_kernel = getenv("RFK_FAVOR_SMEM_FEATURES") != nullptr;  // A/B: the round-1 kernel
    static const bool no_col = getenv("RFK_FAVOR_NO_COL") != nullptr;  // A/B: short token axes on the general kernel
    int rc = RFK_ERR_UNSUPPORTED;
    if (old_kernel && d->io_dtype == RFK_BF16) {
      rc = favor_tc_launch(d, stream);
    } else {
      if (!no_col) rc = favor_col_launch(d, stream);  // one-tile items of the softmax kernel (MSA columns)
      if (rc == RFK_ERR_UNSUPPORTED) rc = favor_tm_launch(d, stream);
    }
    if (rc != RFK_ERR_UNSUPPORTED) return rc;
    // shapes the tensor-core kernel does not cover run on the SIMT kernel (same arithmetic)
  }
  FavorDev p{};
  p.q = d->q; p.k = d->k; p.v = d->v; p.out = d->out; p.proj = d->proj;
  p.dt = d->io_dtype; p.kind = d->kind; p.m = d->m_features; p.heads = d->heads;
  p.tokens = d->tokens; p.G0 = d->G[0]; p.G1 = d->G[1];
  p.gs0 = d->gs[0]; p.gs1 = d->gs[1]; p.ts = d->ts;
  p.ogs0 = d->out_gs[0]; p.ogs1 = d->out_gs[1]; p.ots = d->out_ts;
  const size_t smem = sizeof(float) * (size_t)(kFavorMaxM * 65 + kFavorMaxM * 64 + kFavorMaxM +
                                               2 * kTokTile * 64 + kTokTile * kFavorMaxM + 256);
  static PerDeviceOnce once;
  const int cfg_rc = per_device_once(once, [smem]() {
    return cuda_status(cudaFuncSetAttribute(favor_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  const int64_t blocks = d->G[0] * d->G[1] * d->heads;
  if (blocks > 0x7fffffffLL) return RFK_ERR_BAD_DIMS;
  favor_simt_kernel<<<(unsigned)blocks, 256, smem, stream>>>(p);
  return post_launch();
}
