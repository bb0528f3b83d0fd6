// rfk_favor_tc.cu — fused Performer FAVOR+ attention on tcgen05 / TMEM / TMA (bf16 operands,
// fp32 accumulation). One persistent CTA per SM loops over (group, head) items; for every item
// nothing of size tokens x m ever leaves the SM. Per item the work is a sequence of tile STEPS
//   S0 (softmax kernel only): key tiles  -> U = K.Omega'^T -> global max
//   S1: key tiles   -> U -> k' features (smem) -> ctx[m, d | 1] += k'^T [V | 1]      (TMEM)
//   S3: query tiles -> U -> q' features (smem) -> out = q'.ctx ; den = q'.ksum       -> out/den
// software-pipelined across steps: one control thread issues TMA loads and tcgen05.mma half a
// step ahead, 16 feature warps turn accumulator halves into features while the tensor pipe
// already runs the next U and the previous contraction (ctx / out).
//
// Feature halves: A = m in [0,128) (feature slabs 0,1), B = m in [128,272) (slabs 2,3,4).
// Shared memory (1024-byte aligned tiles, 128-byte swizzle):
//   omega [272 m][64 d] K-major | X0, X1 [128 tok][64] K/Q tiles | V [128 tok][64] | cslab (col 0 = 1)
//   feat 5 x [128 tok][64 m]: MN-major A of the ctx MMAs *and* K-major A of the out MMAs
//   ct [272 m][64 d]: MN-major B of the out MMAs | ksum[272] f32 | den partials | row max
// TMEM (512 cols): S1: ctx blocks 3 x 80 cols [0,240) | U_A [240,368) | U_B [368,512)
//                  S3: out tiles 2 x 64 cols [0,128)  | U_A, U_B as above
#include "rfk_common.cuh"

namespace rfk {

namespace {

constexpr int kFeatWarps = 16;
constexpr int kThreads = 32 * (1 + kFeatWarps);  // warp 0: control (TMA + MMA issue), the rest: features
constexpr int kMP = 272;       // padded feature count (17 * 16)
constexpr int kTile = 128;     // tokens per tile
constexpr uint32_t kOmegaBytes = kMP * 128;        // 34816
constexpr uint32_t kSlabBytes = kTile * 128;       // 16384
constexpr uint32_t kCtBytes = kMP * 128;           // 34816
constexpr uint32_t kOffOmega = 0;
constexpr uint32_t kOffX0 = kOffOmega + kOmegaBytes;
constexpr uint32_t kOffX1 = kOffX0 + kSlabBytes;
constexpr uint32_t kOffV = kOffX1 + kSlabBytes;
constexpr uint32_t kOffCslab = kOffV + kSlabBytes;  // directly after V: LBO of the [V | 1] operand
constexpr uint32_t kOffFeat = kOffCslab + kSlabBytes;
constexpr uint32_t kOffCt = kOffFeat + 5 * kSlabBytes;
constexpr uint32_t kOffBar = kOffCt + kCtBytes;     // 13 mbarriers + tmem slot
constexpr uint32_t kOffScratch = kOffBar + 192;
// scratch floats: ksum[272] | den[3][4][128] | rowmax[2][4][128] | red[16]
constexpr uint32_t kScratchFloats = 272 + 3 * 4 * 128 + 2 * 4 * 128 + 16;
constexpr uint32_t kSmemBytes = kOffScratch + kScratchFloats * 4 + 1024;
static_assert(kOffX0 % 1024 == 0 && kOffV % 1024 == 0 && kOffFeat % 1024 == 0 && kOffCt % 1024 == 0, "align");
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr uint32_t kColCtx = 0, kColOut = 0, kColUA = 240, kColUB = 368;

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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 16 consecutive bf16 values (c0 % 16 == 0) of row `row` into a [rows][64] swizzled slab set
__device__ __forceinline__ void write_row16(uint32_t base, uint32_t slab_bytes, int row, int c0, const float (&f)[16]) {
  const uint32_t slab = base + (uint32_t)(c0 >> 6) * slab_bytes;
  const int c = c0 & 63;
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

enum { kS0 = 0, kS1 = 1, kS3 = 2 };

__global__ void __launch_bounds__(kThreads, 1)
favor_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FavorTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_omega = base + kOffOmega, s_v = base + kOffV, s_cslab = base + kOffCslab,
                 s_feat = base + kOffFeat, s_ct = base + kOffCt;
  const uint32_t s_x[2] = {base + kOffX0, base + kOffX1};
  const uint32_t bb = base + kOffBar;
  const uint32_t bar_x[2] = {bb, bb + 8};
  const uint32_t bar_v = bb + 16, bar_uA = bb + 24, bar_uB = bb + 32, bar_fA = bb + 40, bar_fB = bb + 48,
                 bar_ctx = bb + 56, bar_ro = bb + 64;
  const uint32_t bar_out[2] = {bb + 72, bb + 80}, bar_epi[2] = {bb + 88, bb + 96};
  const uint32_t tmem_slot = bb + 128;
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kOffScratch);
  float* ksum = scratch;                 // [272]
  float* den_sm = scratch + 272;         // [3][4][128]
  float* rowmax = den_sm + 3 * 4 * 128;  // [2][4][128]
  float* red = rowmax + 2 * 4 * 128;     // [16]
  (void)s_cslab;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool softmax_kind = p.kind == 0;
  const int nt = (p.tokens + kTile - 1) / kTile;
  const int S = (softmax_kind ? nt : 0) + 2 * nt;  // tile steps per item
  const int s1_begin = softmax_kind ? nt : 0, s3_begin = s1_begin + nt;

  // ---- one-time setup ----
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_x[0], 1); mbar_init(bar_x[1], 1); mbar_init(bar_v, 1);
      mbar_init(bar_uA, 1); mbar_init(bar_uB, 1);
      mbar_init(bar_fA, kFeatWarps); mbar_init(bar_fB, kFeatWarps);
      mbar_init(bar_ctx, 1); mbar_init(bar_ro, kFeatWarps);
      mbar_init(bar_out[0], 1); mbar_init(bar_out[1], 1);
      mbar_init(bar_epi[0], kFeatWarps); mbar_init(bar_epi[1], kFeatWarps);
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
    // omega' = dn * proj (rows >= m zero), K-major swizzled; constant slab: column 0 = 1;
    // feature slab 4 (only 16 of its 64 columns are ever written) zeroed once
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
      st_shared_v4(base + kOffCslab + sw128_offset(r, c), v);
      st_shared_v4(s_feat + 4 * kSlabBytes + sw128_offset(r, c), make_uint4(0, 0, 0, 0));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  // number of items this CTA processes; global step g = k * S + s over its item sequence
  const int64_t my_items = p.items > blockIdx.x ? (p.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total_steps = my_items * S;
  auto item_coords = [&](int64_t k, int& h, int& g0, int& g1) {
    const int64_t item = blockIdx.x + k * gridDim.x;
    h = (int)(item % p.heads);
    const int64_t g = item / p.heads;
    g0 = (int)(g % p.G0);
    g1 = (int)(g / p.G0);
  };
  auto step_kind = [&](int s) { return s < s1_begin ? kS0 : (s < s3_begin ? kS1 : kS3); };
  auto step_tile = [&](int s) { return s < s1_begin ? s : (s < s3_begin ? s - s1_begin : s - s3_begin); };

  if (warp == 0) {
    if (lane == 0 && total_steps > 0) {
      // =================== control thread ===================
      auto issue_x = [&](int64_t g) {  // X tile (K or Q) of global step g -> X[g & 1]
        const int64_t k = g / S;
        const int s = (int)(g % S);
        int h, g0, g1;
        item_coords(k, h, g0, g1);
        const int b = (int)(g & 1);
        mbar_arrive_expect_tx(bar_x[b], kSlabBytes);
        tma_load_4d(step_kind(s) == kS3 ? &tm_q : &tm_k, bar_x[b], s_x[b], h * 64, step_tile(s) * kTile, g0, g1);
      };
      auto issue_v = [&](int64_t vtile) {  // V tile of global key-tile index vtile = k * nt + t
        const int64_t k = vtile / nt;
        int h, g0, g1;
        item_coords(k, h, g0, g1);
        mbar_arrive_expect_tx(bar_v, kSlabBytes);
        tma_load_4d(&tm_v, bar_v, s_v, h * 64, (int)(vtile % nt) * kTile, g0, g1);
      };
      auto issue_ua = [&](int64_t g) {
        const uint32_t x = s_x[g & 1];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + kColUA, umma_desc_sw128(x) + 2 * k, umma_desc_sw128(s_omega) + 2 * k,
                    umma_idesc_bf16(128, 128), k > 0);
      };
      auto issue_ub = [&](int64_t g) {
        const uint32_t x = s_x[g & 1];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + kColUB, umma_desc_sw128(x) + 2 * k, umma_desc_sw128(s_omega + 128 * 128) + 2 * k,
                    umma_idesc_bf16(128, 144), k > 0);
      };
      auto wait_x = [&](int64_t g) {
        mbar_wait(bar_x[g & 1], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
      };
      // prologue
      issue_x(0);
      if (total_steps > 1) issue_x(1);
      issue_v(0);
      wait_x(0);
      issue_ua(0);
      umma_commit(bar_uA);
      issue_ub(0);
      umma_commit(bar_uB);

      for (int64_t g = 0; g < total_steps; ++g) {
        const int64_t k = g / S;
        const int s = (int)(g % S);
        const int kind = step_kind(s), t = step_tile(s);
        const bool has_next = g + 1 < total_steps;
        const int64_t vt = k * nt + t;  // global key-tile / query-tile index
        const int pbuf = (int)(vt & 1);
        // ---------------- half A ----------------
        mbar_wait(bar_fA, (uint32_t)(g & 1));
        tc_fence_after();
        if (kind == kS1) {
          mbar_wait(bar_v, (uint32_t)(vt & 1));
          tc_fence_after();
          if (t == 0 && k > 0) {
            // the ctx accumulators overlap the previous item's out tiles: their epilogues must be done
            const int64_t w1 = k * nt - 1;
            if (nt >= 2) mbar_wait(bar_epi[(w1 - 1) & 1], (uint32_t)(((w1 - 1) >> 1) & 1));
            mbar_wait(bar_epi[w1 & 1], (uint32_t)((w1 >> 1) & 1));
            tc_fence_after();
          }
          // ctx block 0 (m 0..127): A = feat slabs 0,1 (MN-major), B = [V | 1] (MN-major), K = tokens
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem + kColCtx, desc_mn_sw128(s_feat + kk * 2048, kSlabBytes),
                      desc_mn_sw128(s_v + kk * 2048, kSlabBytes), idesc_bf16_major(128, 80, 1, 1),
                      (t > 0 || kk > 0) ? 1u : 0u);
        } else if (kind == kS3) {
          if (t == 0) {  // ctx read-out (ct / ksum) of this item finished
            mbar_wait(bar_ro, (uint32_t)(k & 1));
            tc_fence_after();
          }
          if (t >= 2) {  // out tile buffer drained by the epilogue two tiles ago
            mbar_wait(bar_epi[pbuf], (uint32_t)(((vt - 2) >> 1) & 1));
            tc_fence_after();
          }
          // out part A: k16 steps over feat slabs 0,1 (K-major A) x ct rows 0..127 (MN-major B)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem + kColOut + pbuf * 64, umma_desc_sw128(s_feat + (ks >> 2) * kSlabBytes) + 2 * (ks & 3),
                      desc_mn_sw128(s_ct + ks * 2048, 0), idesc_bf16_major(128, 64, 0, 1), ks > 0 ? 1u : 0u);
        }
        if (has_next) {
          wait_x(g + 1);
          issue_ua(g + 1);
          umma_commit(bar_uA);
        }
        // ---------------- half B ----------------
        mbar_wait(bar_fB, (uint32_t)(g & 1));
        tc_fence_after();
        if (kind == kS1) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t bd = desc_mn_sw128(s_v + kk * 2048, kSlabBytes);
            umma_bf16(tmem + kColCtx + 80, desc_mn_sw128(s_feat + 2 * kSlabBytes + kk * 2048, kSlabBytes), bd,
                      idesc_bf16_major(128, 80, 1, 1), (t > 0 || kk > 0) ? 1u : 0u);
            // block 2 = m 256..383: only slab 4 exists, LBO 0 mirrors it into the unused upper half
            umma_bf16(tmem + kColCtx + 160, desc_mn_sw128(s_feat + 4 * kSlabBytes + kk * 2048, 0), bd,
                      idesc_bf16_major(128, 80, 1, 1), (t > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(bar_ctx);
        } else if (kind == kS3) {
#pragma unroll
          for (int ks = 8; ks < 17; ++ks)
            umma_bf16(tmem + kColOut + pbuf * 64, umma_desc_sw128(s_feat + (ks >> 2) * kSlabBytes) + 2 * (ks & 3),
                      desc_mn_sw128(s_ct + ks * 2048, 0), idesc_bf16_major(128, 64, 0, 1), 1u);
          umma_commit(bar_out[pbuf]);
        }
        if (has_next) {
          issue_ub(g + 1);
          umma_commit(bar_uB);
        }
        // X[g & 1] is free again (its U MMAs completed before the features could be produced)
        if (g + 2 < total_steps) issue_x(g + 2);
        if (kind == kS1) {
          // V is free once the ctx MMAs of this tile are done: prefetch the next key tile's V
          mbar_wait(bar_ctx, (uint32_t)(vt & 1));
          if (vt + 1 < my_items * nt) issue_v(vt + 1);
        }
      }
      // the last out tiles of the last item are drained by the feature warps; consume their barriers
      {
        const int64_t w1 = my_items * nt - 1;
        if (nt >= 2) mbar_wait(bar_epi[(w1 - 1) & 1], (uint32_t)(((w1 - 1) >> 1) & 1));
        mbar_wait(bar_epi[w1 & 1], (uint32_t)((w1 >> 1) & 1));
      }
    }
  } else {
    // =================== feature / epilogue warps ===================
    const int fw = warp - 1;        // 0..15
    const int lg = warp & 3;        // TMEM lane group this warp may touch
    const int quarter = fw >> 2;    // 4 warps share a lane group and split the column chunks
    const int row = lg * 32 + lane; // token row in the tile / TMEM lane
    const uint32_t t_lane = ((uint32_t)(lg * 32) << 16);
    const float ratio = rsqrtf((float)p.m);
    constexpr float kLog2e = 1.4426950408889634f;
    auto split = [&](int n, int& c0, int& c1) {
      c0 = (n * quarter) >> 2;
      c1 = (n * (quarter + 1)) >> 2;
    };
    auto arrive = [&](uint32_t bar, bool wrote_smem) {
      if (wrote_smem) fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // feature map of 16 accumulator columns (feature index m0..m0+15); returns sum f * ksum if `query`
    auto feat_chunk = [&](uint32_t tcol, int m0, bool valid, bool full, float sub, bool query) -> float {
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
      write_row16(s_feat, kSlabBytes, row, m0, f);
      float d = 0.f;
      if (query) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 ks = *reinterpret_cast<const float4*>(ksum + m0 + 4 * j);
          d = fmaf(f[4 * j], ks.x, d); d = fmaf(f[4 * j + 1], ks.y, d);
          d = fmaf(f[4 * j + 2], ks.z, d); d = fmaf(f[4 * j + 3], ks.w, d);
        }
      }
      return d;
    };
    auto scan_max = [&](uint32_t tcol, int m_base, int nch, float mx) -> float {
      int c0, c1;
      split(nch, c0, c1);
      for (int c = c0; c < c1; ++c) {
        uint32_t r[16];
        tmem_ld_32x16(tmem + t_lane + tcol + c * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (m_base + c * 16 + i < p.m) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      return mx;
    };
    // epilogue of out tile w (global query-tile index), tile t of item (h, g0, g1)
    auto epilogue = [&](int64_t w, int t, int h, int g0, int g1) {
      const int pb = (int)(w & 1);
      mbar_wait(bar_out[pb], (uint32_t)((w >> 1) & 1));
      tc_fence_after();
      uint32_t r0[16];
      tmem_ld_32x16(tmem + t_lane + kColOut + pb * 64 + quarter * 16, r0);
      tmem_ld_wait();
      const float* dp = den_sm + (int)(w % 3) * 512 + row;
      const float den = (dp[0] + dp[128]) + (dp[256] + dp[384]);
      if (t * kTile + row < p.tokens) {
        const float inv = 1.f / den;
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)g1 * p.ogs1 + (int64_t)g0 * p.ogs0 +
                            (int64_t)(t * kTile + row) * p.ots + h * 64 + quarter * 16;
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
      arrive(bar_epi[pb], false);
    };

    float gmax = 0.f, mx_run = -INFINITY;
    // deferred epilogue state (out tile of the previous S3 step)
    bool pend = false;
    int64_t pend_w = 0;
    int pend_t = 0, pend_h = 0, pend_g0 = 0, pend_g1 = 0;

    for (int64_t g = 0; g < total_steps; ++g) {
      const int64_t k = g / S;
      const int s = (int)(g % S);
      const int kind = step_kind(s), t = step_tile(s);
      const int64_t vt = k * nt + t;
      const uint32_t xbuf = s_x[g & 1];
      const bool valid = t * kTile + row < p.tokens;
      const bool tile_full = (t + 1) * kTile <= p.tokens;
      const uint32_t par = (uint32_t)(g & 1);
      if (s == 0) mx_run = -INFINITY;

      if (kind == kS0) {
        mbar_wait(bar_uA, par);
        tc_fence_after();
        float mx = scan_max(kColUA, 0, 8, -INFINITY);
        arrive(bar_fA, false);
        mbar_wait(bar_uB, par);
        tc_fence_after();
        mx = scan_max(kColUB, 128, 9, mx);
        if (valid) mx_run = fmaxf(mx_run, mx);
        arrive(bar_fB, false);
        if (s == s1_begin - 1) {  // last S0 step: block-wide max
          const float wm = warp_max(mx_run);
          if (lane == 0) red[fw] = wm;
          named_bar_sync(1, kFeatWarps * 32);
          gmax = red[0];
#pragma unroll
          for (int i = 1; i < kFeatWarps; ++i) gmax = fmaxf(gmax, red[i]);
          named_bar_sync(1, kFeatWarps * 32);  // red[] may be rewritten by the next item's S0
        }
        continue;
      }

      if (kind == kS1) {
        float sub = 0.f;
        if (softmax_kind) {
          mbar_wait(bar_x[g & 1], (uint32_t)((g >> 1) & 1));
          sub = (row_half_sqnorm(xbuf, row) + gmax) * kLog2e;
        }
        mbar_wait(bar_uA, par);
        tc_fence_after();
        int c0, c1;
        split(8, c0, c1);
        for (int c = c0; c < c1; ++c) feat_chunk(kColUA + c * 16, c * 16, valid, tile_full, sub, false);
        arrive(bar_fA, true);
        mbar_wait(bar_uB, par);
        tc_fence_after();
        split(9, c0, c1);
        for (int c = c0; c < c1; ++c) {
          const int m0 = 128 + c * 16;
          feat_chunk(kColUB + c * 16, m0, valid, tile_full && m0 + 16 <= p.m, sub, false);
        }
        arrive(bar_fB, true);
        if (t == nt - 1) {
          // ---- ctx read-out: TMEM lanes are feature rows m; ct[m][0..63] (bf16) and ksum[m] ----
          mbar_wait(bar_ctx, (uint32_t)(vt & 1));
          tc_fence_after();
          // quarter j of each lane group reads ctx block j (rows m = 128 j + row); quarter 3 idles
          for (int j = quarter; j < 3; j += 4) {
            const int mrow = j * 128 + row;
#pragma unroll 1
            for (int c = 0; c < 5; ++c) {
              uint32_t r[16];
              tmem_ld_32x16(tmem + t_lane + kColCtx + j * 80 + c * 16, r);
              tmem_ld_wait();
              if (mrow < kMP) {
                if (c < 4) {
                  float f[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
                  write_row16(s_ct, kCtBytes, mrow, c * 16, f);
                } else {
                  ksum[mrow] = __uint_as_float(r[0]);
                }
              }
            }
          }
          arrive(bar_ro, true);
          named_bar_sync(1, kFeatWarps * 32);  // ksum visible to every feature warp before S3
        }
        continue;
      }

      // ---------------- kS3: query tile ----------------
      int h, g0, g1;
      item_coords(k, h, g0, g1);
      float sub = 0.f;
      mbar_wait(bar_uA, par);
      tc_fence_after();
      if (softmax_kind) {
        mbar_wait(bar_x[g & 1], (uint32_t)((g >> 1) & 1));
        const float diag = row_half_sqnorm(xbuf, row);
        mbar_wait(bar_uB, par);
        tc_fence_after();
        float mx = scan_max(kColUA, 0, 8, -INFINITY);
        mx = scan_max(kColUB, 128, 9, mx);
        float* rm = rowmax + (int)(g & 1) * 512;
        rm[quarter * 128 + row] = mx;
        named_bar_sync(1, kFeatWarps * 32);
        mx = fmaxf(fmaxf(rm[row], rm[128 + row]), fmaxf(rm[256 + row], rm[384 + row]));
        sub = (diag + mx) * kLog2e;
      }
      float den = 0.f;
      int c0, c1;
      split(8, c0, c1);
      for (int c = c0; c < c1; ++c) den += feat_chunk(kColUA + c * 16, c * 16, valid, tile_full, sub, true);
      arrive(bar_fA, true);
      if (!softmax_kind) {
        mbar_wait(bar_uB, par);
        tc_fence_after();
      }
      split(9, c0, c1);
      for (int c = c0; c < c1; ++c) {
        const int m0 = 128 + c * 16;
        den += feat_chunk(kColUB + c * 16, m0, valid, tile_full && m0 + 16 <= p.m, sub, true);
      }
      den_sm[(int)(vt % 3) * 512 + quarter * 128 + row] = den;
      arrive(bar_fB, true);
      // deferred epilogue of the previous query tile (its out MMAs ran behind this tile's features)
      if (pend) epilogue(pend_w, pend_t, pend_h, pend_g0, pend_g1);
      pend = true; pend_w = vt; pend_t = t; pend_h = h; pend_g0 = g0; pend_g1 = g1;
      if (t == nt - 1) {  // last tile of the item: drain now (the ctx accumulators get reused next)
        epilogue(pend_w, pend_t, pend_h, pend_g0, pend_g1);
        pend = false;
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
