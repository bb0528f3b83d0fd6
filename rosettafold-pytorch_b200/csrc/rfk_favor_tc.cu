// rfk_favor_tc.cu — fused Performer FAVOR+ attention on tcgen05 / TMEM / TMA (bf16 operands,
// fp32 accumulation). One persistent CTA per SM loops over (group, head) items; for every item
// nothing of size tokens x m ever leaves the SM:
//
//   keys   : TMA K,V tile -> U = K.Omega'^T (tcgen05, TMEM) -> feature map k' (CUDA cores, from
//            TMEM) -> bf16 k' tile in shared memory -> [ctx^T ; ksum] += [V | 1]^T k' (tcgen05,
//            both operands MN-major so neither V nor k' is ever transposed)
//   queries: TMA Q tile -> U = Q.Omega'^T -> q' -> out|den = q'.[ctx^T ; ksum]^T (tcgen05) ->
//            out/den -> global
//
// Shared memory (all tiles 1024-byte aligned, 128-byte swizzle):
//   omega  [272 m][64 d]      bf16, K-major   (dn * projection matrix, rows >= m zero)
//   cslab  [128 tok][64]      bf16, column 0 = 1: second MN-chunk of the "A = [V | 1]" operand
//   kbuf   [128 tok][64 d]    K or Q tile (TMA);   vbuf [128 tok][64 d] V tile (TMA)
//   feat   5 x [128 tok][64 m] k'/q' features: MN-major B of the context MMA *and* K-major A of
//                              the output MMA (same bytes)
//   ctxt   5 x [80][64 m]     rows 0..63 ctx^T, row 64 ksum, K-major B of the output MMA
// TMEM (512 columns): D2 = [ctx^T;ksum] cols [0,272) | U halves cols [272,416) | D3 cols [416,496);
// the full-width U of the key-max pre-pass and of the query phase reuses cols [0,272).
#include "rfk_common.cuh"

namespace rfk {

namespace {

constexpr int kFeatWarps = 16;
constexpr int kThreads = 32 * (1 + kFeatWarps);  // warp 0: control (TMA + MMA issue), the rest: feature/epilogue
constexpr int kMP = 272;       // padded feature count (17 * 16)
constexpr int kTile = 128;     // tokens per tile
constexpr uint32_t kOmegaBytes = kMP * 128;        // 34816
constexpr uint32_t kSlabBytes = kTile * 128;       // 16384
constexpr uint32_t kCtxSlabBytes = 80 * 128;       // 10240
constexpr uint32_t kOffOmega = 0;
constexpr uint32_t kOffK = kOffOmega + kOmegaBytes;
constexpr uint32_t kOffV = kOffK + kSlabBytes;
constexpr uint32_t kOffCslab = kOffV + kSlabBytes;  // directly after V: LBO of the [V | 1] operand
constexpr uint32_t kOffFeat = kOffCslab + kSlabBytes;
constexpr uint32_t kOffCtx = kOffFeat + 5 * kSlabBytes;
constexpr uint32_t kOffBar = kOffCtx + 5 * kCtxSlabBytes;  // barriers + scratch
constexpr uint32_t kOffScratch = kOffBar + 64;
constexpr uint32_t kSmemBytes = kOffScratch + (4 * 128 + 16) * 4 + 1024;
static_assert(kOffCslab % 1024 == 0 && kOffK % 1024 == 0 && kOffFeat % 1024 == 0 && kOffCtx % 1024 == 0, "align");

constexpr uint32_t kColD2 = 0, kColU = 272, kColD3 = 416;

struct FavorTcParams {
  const float* proj;
  void* out;
  int kind, m, heads;
  int tokens;
  int64_t G0, G1, items;
  int64_t ogs0, ogs1, ots;
};

// MN-major SW128 descriptor: rows are K indices (128 B each, 8-row groups SBO=1024 apart),
// 64-element MN chunks are `lbo_bytes` apart.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16_major(int M, int N, int a_mn, int b_mn) {
  return umma_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// write 16 consecutive features (m0 % 16 == 0) of token row `row` into the feature slabs
__device__ __forceinline__ void write_feat16(uint32_t feat_base, int row, int m0, const float (&f)[16]) {
  const uint32_t slab = feat_base + (uint32_t)(m0 >> 6) * kSlabBytes;
  const int c = m0 & 63;
  uint4 a, b;
  a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
  a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
  b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
  b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
  st_shared_v4(slab + sw128_offset(row, c), a);
  st_shared_v4(slab + sw128_offset(row, c + 8), b);
}

// 0.5 * dn^2 * |x|^2 of row `row` of a [128][64] bf16 swizzled tile
__device__ __forceinline__ float row_half_sqnorm(uint32_t tile, int row) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 v = ld_shared_v4(tile + sw128_offset(row, j * 8));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = bf16lo(w[i]), b = bf16hi(w[i]);
      s = fmaf(a, a, s);
      s = fmaf(b, b, s);
    }
  }
  return s * (0.5f * 0.125f);  // dn^2 = 64^-1/2 = 1/8
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads, 1)
favor_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FavorTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_omega = base + kOffOmega, s_cslab = base + kOffCslab, s_feat = base + kOffFeat,
                 s_ctx = base + kOffCtx;
  const uint32_t s_buf[2] = {base + kOffK, base + kOffV};       // [0] = "k" buffer, [1] = "v" buffer
  const uint32_t bar_ld[2] = {base + kOffBar, base + kOffBar + 8};
  const uint32_t bar_mma = base + kOffBar + 16, bar_feat = base + kOffBar + 24;
  const uint32_t tmem_slot = base + kOffBar + 32;
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kOffScratch);
  float* rowmax = scratch;            // [4][128]
  float* red = scratch + 4 * 128;     // [16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool softmax_kind = p.kind == 0;
  const int nt = (p.tokens + kTile - 1) / kTile;

  // ---- one-time setup ----
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_ld[0], 1);
      mbar_init(bar_ld[1], 1);
      mbar_init(bar_mma, 1);
      mbar_init(bar_feat, kFeatWarps);
      fence_barrier_init();
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      tma_prefetch_desc(&tm_v);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  {
    // omega' = dn * proj (rows >= m zero), K-major swizzled; constant slab: column 0 = 1
    const float dn = 0.35355339059327373f;  // 64^-1/4
    for (int i = threadIdx.x; i < kMP * 8; i += kThreads) {
      const int r = i >> 3, c = (i & 7) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = r < p.m ? __ldg(p.proj + (int64_t)r * 64 + c + j) * dn : 0.f;
      uint4 v;
      v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
      v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
      st_shared_v4(s_omega + sw128_offset(r, c), v);
    }
    for (int i = threadIdx.x; i < kTile * 8; i += kThreads) {
      const int r = i >> 3, c = (i & 7) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c == 0) v.x = 0x00003F80u;  // bf16(1.0) in element 0
      st_shared_v4(s_cslab + sw128_offset(r, c), v);
    }
    // rows 65..79 of ctxt are never written by the context read-out: zero the whole buffer once
    for (int i = threadIdx.x; i < (int)(5 * kCtxSlabBytes / 16); i += kThreads)
      st_shared_v4(s_ctx + i * 16, make_uint4(0, 0, 0, 0));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  // Loads completed so far on each TMA barrier; every thread advances the same deterministic
  // schedule (whether or not it actually waits), parity of the next wait = count & 1.
  uint32_t nld[2] = {0, 0};
  uint32_t ph_mma = 0, ph_feat = 0;
  const float ratio = rsqrtf((float)p.m);
  constexpr float kLog2e = 1.4426950408889634f;

  // feature-warp geometry: 4 warps share a TMEM lane group and split the column chunks
  const int fw = warp - 1;                  // 0..15 (valid for warp >= 1)
  const int lg = warp & 3;                  // TMEM lane group this warp may touch
  const int quarter = fw >> 2;
  const int row = lg * 32 + lane;           // token row in the tile / TMEM lane
  const uint32_t t_lane = ((uint32_t)(lg * 32) << 16);
  auto split = [&](int n, int& c0, int& c1) {
    c0 = (n * quarter) >> 2;
    c1 = (n * (quarter + 1)) >> 2;
  };

  bool pre_issued = false;  // control thread: first loads of this item already in flight
  for (int64_t item = blockIdx.x; item < p.items; item += gridDim.x) {
    const int h = (int)(item % p.heads);
    const int64_t g = item / p.heads;
    const int g0 = (int)(g % p.G0), g1 = (int)(g / p.G0);

    if (warp == 0) {
      if (lane == 0) {
        // =================== control thread ===================
        auto issue = [&](const CUtensorMap* tm, int b, int t, int hh, int gg0, int gg1) {
          mbar_arrive_expect_tx(bar_ld[b], kSlabBytes);
          tma_load_4d(tm, bar_ld[b], s_buf[b], hh * 64, t * kTile, gg0, gg1);
        };
        auto wait_ld = [&](int b) {
          mbar_wait(bar_ld[b], nld[b] & 1u);
          ++nld[b];
          tc_fence_after();
        };
        auto mma_u_full = [&](uint32_t tile) {  // U[128 x 272] = tile . omega'^T into cols [0,272)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + kColD2, umma_desc_sw128(tile) + 2 * k, umma_desc_sw128(s_omega) + 2 * k,
                      umma_idesc_bf16(128, 144), k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + kColD2 + 144, umma_desc_sw128(tile) + 2 * k,
                      umma_desc_sw128(s_omega + 144 * 128) + 2 * k, umma_idesc_bf16(128, 128), k > 0);
        };
        auto step = [&]() {  // hand the MMA results to the feature warps and wait for them
          umma_commit(bar_mma);
          mbar_wait(bar_feat, ph_feat);
          ph_feat ^= 1u;
          tc_fence_after();
        };
        // ---- S0: global key max (softmax kernel); tiles alternate between the two buffers ----
        if (softmax_kind) {
          if (!pre_issued) issue(&tm_k, 0, 0, h, g0, g1);
          for (int t = 0; t < nt; ++t) {
            if (t + 1 < nt) issue(&tm_k, (t + 1) & 1, t + 1, h, g0, g1);
            wait_ld(t & 1);
            mma_u_full(s_buf[t & 1]);
            step();
          }
          issue(&tm_k, 0, 0, h, g0, g1);
          issue(&tm_v, 1, 0, h, g0, g1);
        } else if (!pre_issued) {
          issue(&tm_k, 0, 0, h, g0, g1);
          issue(&tm_v, 1, 0, h, g0, g1);
        }
        // ---- S1: context ----
        for (int t = 0; t < nt; ++t) {
          wait_ld(0);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // half A: m 0..143
            umma_bf16(tmem + kColU, umma_desc_sw128(s_buf[0]) + 2 * k, umma_desc_sw128(s_omega) + 2 * k,
                      umma_idesc_bf16(128, 144), k > 0);
          step();
#pragma unroll
          for (int k = 0; k < 4; ++k)  // half B: m 144..271
            umma_bf16(tmem + kColU, umma_desc_sw128(s_buf[0]) + 2 * k,
                      umma_desc_sw128(s_omega + 144 * 128) + 2 * k, umma_idesc_bf16(128, 128), k > 0);
          step();
          // K buffer is free: prefetch the next K tile, or the first Q tile
          if (t + 1 < nt) issue(&tm_k, 0, t + 1, h, g0, g1);
          else issue(&tm_q, 0, 0, h, g0, g1);
          wait_ld(1);
          // [ctx^T ; ksum][128 x 272] += [V | 1]^T (K = tokens) . k'   — both operands MN-major
          const uint32_t lbo_a = s_cslab - s_buf[1];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = desc_mn_sw128(s_buf[1] + k * 2048, lbo_a);
            umma_bf16(tmem + kColD2, ad, desc_mn_sw128(s_feat + k * 2048, kSlabBytes),
                      idesc_bf16_major(128, 128, 1, 1), (t > 0 || k > 0));
            umma_bf16(tmem + kColD2 + 128, ad, desc_mn_sw128(s_feat + 2 * kSlabBytes + k * 2048, kSlabBytes),
                      idesc_bf16_major(128, 144, 1, 1), (t > 0 || k > 0));
          }
          step();  // feature warps: no-op, or the ctx^T read-out after the last tile
          if (t + 1 < nt) issue(&tm_v, 1, t + 1, h, g0, g1);
          else if (nt > 1) issue(&tm_q, 1, 1, h, g0, g1);
        }
        // ---- S3: queries (tile t lives in buffer t & 1) ----
        pre_issued = false;
        for (int t = 0; t < nt; ++t) {
          const int b = t & 1;
          wait_ld(b);
          mma_u_full(s_buf[b]);
          step();
          if (t + 2 < nt) issue(&tm_q, b, t + 2, h, g0, g1);
          // out|den [128 tok x 80] = q'[128 x 272] . [ctx^T;ksum]^T
#pragma unroll
          for (int ks = 0; ks < 17; ++ks) {
            const int kb = ks >> 2, kk = ks & 3;
            umma_bf16(tmem + kColD3, umma_desc_sw128(s_feat + kb * kSlabBytes) + 2 * kk,
                      umma_desc_sw128(s_ctx + kb * kCtxSlabBytes) + 2 * kk, umma_idesc_bf16(128, 80), ks > 0);
          }
          if (t == nt - 1) {
            // both tile buffers are free: start the next item's first loads behind this epilogue
            const int64_t nitem = item + gridDim.x;
            if (nitem < p.items) {
              const int nh = (int)(nitem % p.heads);
              const int64_t ng = nitem / p.heads;
              const int ng0 = (int)(ng % p.G0), ng1 = (int)(ng / p.G0);
              issue(&tm_k, 0, 0, nh, ng0, ng1);
              if (!softmax_kind) issue(&tm_v, 1, 0, nh, ng0, ng1);
              pre_issued = true;
            }
          }
          step();
        }
      }
      // lanes 1..31 of the control warp idle until the next item / teardown
    } else {
      // =================== feature / epilogue warps ===================
      auto wait_mma = [&]() {
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
      };
      auto arrive = [&](bool wrote_smem) {
        if (wrote_smem) fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_feat);
      };
      // consume one completed load on buffer b; only the softmax kernel reads the tile (|x|^2)
      auto consume_ld = [&](int b, bool need) {
        if (need) mbar_wait(bar_ld[b], nld[b] & 1u);
        ++nld[b];
      };
      // feature map of 16 accumulator columns -> bf16 features in shared memory
      auto feat_chunk = [&](uint32_t tcol, int m0, bool valid, bool full, float sub) {
        uint32_t r[16];
        tmem_ld_32x16(tmem + t_lane + tcol, r);
        tmem_ld_wait();
        float f[16];
        if (softmax_kind) {
          const float bias = ratio * 1e-4f;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            f[i] = fmaf(ratio, ex2_approx(fmaf(__uint_as_float(r[i]), kLog2e, -sub)), bias);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(__uint_as_float(r[i]), 0.f) + 1e-3f;
        }
        if (!full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = (valid && m0 + i < p.m) ? f[i] : 0.f;
        }
        write_feat16(s_feat, row, m0, f);
      };

      float gmax = 0.f;
      if (softmax_kind) {
        // ---- S0 ----
        float mx = -INFINITY;
        for (int t = 0; t < nt; ++t) {
          consume_ld(t & 1, false);
          wait_mma();
          const bool valid = t * kTile + row < p.tokens;
          int c0, c1;
          split(17, c0, c1);
          for (int c = c0; c < c1; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(tmem + t_lane + kColD2 + c * 16, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c * 16 + i < p.m) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          }
          arrive(false);
        }
        mx = warp_max(mx);
        if (lane == 0) red[fw] = mx;
        named_bar_sync(1, kFeatWarps * 32);
        gmax = red[0];
#pragma unroll
        for (int i = 1; i < kFeatWarps; ++i) gmax = fmaxf(gmax, red[i]);
      }

      // ---- S1 ----
      for (int t = 0; t < nt; ++t) {
        consume_ld(0, softmax_kind);
        const bool valid = t * kTile + row < p.tokens;
        const bool tile_full = (t + 1) * kTile <= p.tokens;
        float sub = 0.f;
        for (int hf = 0; hf < 2; ++hf) {
          wait_mma();
          if (hf == 0 && softmax_kind) sub = (row_half_sqnorm(s_buf[0], row) + gmax) * kLog2e;
          const int nch = hf == 0 ? 9 : 8, moff = hf == 0 ? 0 : 144;
          int c0, c1;
          split(nch, c0, c1);
          for (int c = c0; c < c1; ++c) {
            const int m0 = moff + c * 16;
            feat_chunk(kColU + c * 16, m0, valid, tile_full && m0 + 16 <= p.m, sub);
          }
          arrive(true);
        }
        wait_mma();  // context MMAs of this tile done (V, feat buffers free again)
        consume_ld(1, false);
        bool wrote = false;
        if (t == nt - 1 && lg != 3) {
          // read out [ctx^T ; ksum] rows 0..79 -> ctxt (bf16, K-major over m). Lane groups 0,1:
          // ctx^T rows; group 2: rows 64..79 (row 64 = ksum, the rest are exact zeros).
          const bool owner = lg < 2 || lane < 16;  // every lane executes the aligned TMEM loads
          int c0, c1;
          split(17, c0, c1);
          for (int c = c0; c < c1; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(tmem + t_lane + kColD2 + c * 16, r);
            tmem_ld_wait();
            if (owner) {
              const int m0 = c * 16;
              const uint32_t slab = s_ctx + (uint32_t)(m0 >> 6) * kCtxSlabBytes;
              uint4 a, b;
              a.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
              a.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
              a.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
              a.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
              b.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));
              b.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
              b.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13]));
              b.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
              st_shared_v4(slab + sw128_offset(row, m0 & 63), a);
              st_shared_v4(slab + sw128_offset(row, (m0 & 63) + 8), b);
            }
          }
          wrote = true;
        }
        arrive(wrote);
      }

      // ---- S3 ----
      for (int t = 0; t < nt; ++t) {
        const int b = t & 1;
        consume_ld(b, softmax_kind);
        wait_mma();
        const bool valid = t * kTile + row < p.tokens;
        const bool tile_full = (t + 1) * kTile <= p.tokens;
        int c0, c1;
        split(17, c0, c1);
        float sub = 0.f;
        if (softmax_kind) {
          const float diag = row_half_sqnorm(s_buf[b], row);
          float mx = -INFINITY;
          for (int c = c0; c < c1; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(tmem + t_lane + kColD2 + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c * 16 + i < p.m) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
          rowmax[quarter * 128 + row] = mx;
          named_bar_sync(1, kFeatWarps * 32);
          mx = fmaxf(fmaxf(rowmax[row], rowmax[128 + row]), fmaxf(rowmax[256 + row], rowmax[384 + row]));
          sub = (diag + mx) * kLog2e;
        }
        for (int c = c0; c < c1; ++c) {
          const int m0 = c * 16;
          feat_chunk(kColD2 + c * 16, m0, valid, tile_full && m0 + 16 <= p.m, sub);
        }
        arrive(true);
        // ---- output tile: each quarter stores 16 of the 64 head channels ----
        wait_mma();
        {
          uint32_t rd[16], r0[16];
          tmem_ld_32x16(tmem + t_lane + kColD3 + 64, rd);  // column 64 = normaliser
          tmem_ld_32x16(tmem + t_lane + kColD3 + quarter * 16, r0);
          tmem_ld_wait();
          if (valid) {
            const float inv = 1.f / __uint_as_float(rd[0]);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)g1 * p.ogs1 +
                                (int64_t)g0 * p.ogs0 + (int64_t)(t * kTile + row) * p.ots + h * 64 + quarter * 16;
            uint4 w0, w1;
            w0.x = pack_bf16x2(__uint_as_float(r0[0]) * inv, __uint_as_float(r0[1]) * inv);
            w0.y = pack_bf16x2(__uint_as_float(r0[2]) * inv, __uint_as_float(r0[3]) * inv);
            w0.z = pack_bf16x2(__uint_as_float(r0[4]) * inv, __uint_as_float(r0[5]) * inv);
            w0.w = pack_bf16x2(__uint_as_float(r0[6]) * inv, __uint_as_float(r0[7]) * inv);
            w1.x = pack_bf16x2(__uint_as_float(r0[8]) * inv, __uint_as_float(r0[9]) * inv);
            w1.y = pack_bf16x2(__uint_as_float(r0[10]) * inv, __uint_as_float(r0[11]) * inv);
            w1.z = pack_bf16x2(__uint_as_float(r0[12]) * inv, __uint_as_float(r0[13]) * inv);
            w1.w = pack_bf16x2(__uint_as_float(r0[14]) * inv, __uint_as_float(r0[15]) * inv);
            reinterpret_cast<uint4*>(op)[0] = w0;
            reinterpret_cast<uint4*>(op)[1] = w1;
          }
        }
        arrive(false);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_head_tmap(CUtensorMap* map, const void* ptr, const rfk_favor_desc* d) {
  static EncodeTiledFn enc = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  if (!enc) return RFK_ERR_TMA_ENCODE;
  // dims: (column within the heads*64 slice, token, g0, g1)
  cuuint64_t dims[4] = {(cuuint64_t)d->heads * 64, (cuuint64_t)d->tokens, (cuuint64_t)d->G[0], (cuuint64_t)d->G[1]};
  cuuint64_t strides[3] = {(cuuint64_t)d->ts * 2, (cuuint64_t)d->gs[0] * 2, (cuuint64_t)d->gs[1] * 2};
  if (d->G[0] == 1) strides[1] = strides[0] * (cuuint64_t)d->tokens;
  if (d->G[1] == 1) strides[2] = strides[1] * (cuuint64_t)d->G[0];
  cuuint32_t box[4] = {64, (cuuint32_t)kTile, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RFK_OK : RFK_ERR_TMA_ENCODE;
}

}  // namespace

int favor_tc_launch(const rfk_favor_desc* d, cudaStream_t stream) {
  // shapes the tensor-core kernel covers; anything else runs on the SIMT kernel
  if (d->m_features > kMP || d->m_features < 16) return RFK_ERR_UNSUPPORTED;
  if (d->tokens > (1 << 24)) return RFK_ERR_UNSUPPORTED;
  if (!aligned16(d->q) || !aligned16(d->k) || !aligned16(d->v) || !aligned16(d->out)) return RFK_ERR_UNSUPPORTED;
  if (d->ts % 8 || d->gs[0] % 8 || d->gs[1] % 8 || d->out_ts % 8 || d->out_gs[0] % 8 || d->out_gs[1] % 8)
    return RFK_ERR_UNSUPPORTED;
  int rc = check_arch();
  if (rc != RFK_OK) return rc;
  CUtensorMap tq, tk, tv;
  if ((rc = make_head_tmap(&tq, d->q, d)) != RFK_OK) return rc;
  if ((rc = make_head_tmap(&tk, d->k, d)) != RFK_OK) return rc;
  if ((rc = make_head_tmap(&tv, d->v, d)) != RFK_OK) return rc;
  FavorTcParams p{};
  p.proj = d->proj; p.out = d->out; p.kind = d->kind; p.m = d->m_features; p.heads = d->heads;
  p.tokens = (int)d->tokens; p.G0 = d->G[0]; p.G1 = d->G[1];
  p.items = d->G[0] * d->G[1] * d->heads;
  p.ogs0 = d->out_gs[0]; p.ogs1 = d->out_gs[1]; p.ots = d->out_ts;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(favor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return cuda_status(e);
    configured = true;
  }
  int grid = num_sms();
  if (p.items < grid) grid = (int)p.items;
  favor_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, p);
  return post_launch();
}

}  // namespace rfk
