// rfk_favor_col.cu — FAVOR+ softmax-kernel attention for SHORT token axes (tokens <= 128: ONE tile per item), i.e.
// the MSA-column attention of the trunk (rosettafold_pytorch.py:313-318: tokens = sequences of the MSA, one item per
// (residue, head)). Same arithmetic as the general kernel (rfk_favor_tm.cu, whose header explains the orientation of
// the four products); what differs is the schedule.
//
// The softmax feature map needs a stabiliser before it can exponentiate: the maximum of the WHOLE key projection of
// the item, and per query row the maximum over all 266 features. The general kernel walks 128 x 128 jobs through two
// accumulator slots, so it computes every projection twice (a max pass, then the feature pass): 12 jobs per item,
// each a dependent round trip between the MMA issuer and the feature warps. With one tile per item the whole
// projection of an item (3 feature chunks x 128 columns) fits tensor memory at once, so here
//   * every projection is computed ONCE and stays resident while the maxima are taken;
//   * the three context blocks ctx^T_c accumulate into three different regions (the dedicated 80-column block C and
//     the already-consumed U regions of chunks 0 and 1: tcgen05.mma of one thread execute in issue order), so their
//     read-outs do not serialise;
//   * the query projections go into regions that are free the moment the context MMAs have been ISSUED, and the key
//     projections of the next item follow the output MMAs in the same issue stream: no commit / mbarrier hop guards
//     an accumulator against the issuer itself.
//
// TMEM (512 columns)        key phase                      query phase
//   U0 [  0,128)   U^T chunk 0 -> k'^T_0                   ctx^T_1 (80 cols), then U chunk 1 -> q'_1
//   U1 [128,256)   U^T chunk 1 -> k'^T_1                   ctx^T_2 (80 cols, lanes 0..15)
//   U2 [256,384)   U^T chunk 2 -> k'^T_2                   U chunk 0 -> q'_0
//   C  [384,464)   ctx^T_0                                 out | den
//   S  [464,480)                                           U chunk 2 (16 features) -> q'_2
// Key chunk 1 of the NEXT item is projected into U1 as soon as ctx^T_2 has left it (before the output MMAs), so the
// feature warps take its maxima, and the |k|^2 terms, while those MMAs run. A quarter of the exponentials are
// evaluated on the FMA pipe (ex2_poly) instead of the SFU.
// Measured (B200, G = 512, T = 128, H = 12): 381 us against 457 us for the general kernel; an in-kernel clock64
// timeline (-DRFK_COL_TIMELINE, tools/favor_col_timeline.py) puts ~6.3 K of the ~14.9 K cycles per item into the two
// exponential passes (issue- / SFU-bound) and the rest into the six dependent issuer <-> feature-warp hand-offs, which
// cannot overlap with a second item because one item fills tensor memory (profiles/r02_favor_notes.md).
// Roles: warp 0 TMA producer (K, V, Q tile of an item; 6-slot ring = two items), warp 1 MMA issuer, warps 2..17
// feature warps (warp (lane group lg, column quarter cq) owns 32 lanes x 32 columns of every region).
// Every mbarrier completes exactly once per item and every role waits on the item's chain, so parities are the item
// count and no waiter can fall a phase behind.
#include "rfk_favor_device.cuh"

namespace rfk {

namespace {

constexpr int kFeatWarps = 16;
constexpr int kFeatThreads = 32 * kFeatWarps;
constexpr int kThreads = 32 * (2 + kFeatWarps);  // TMA producer, MMA issuer, feature warps
constexpr int kMP = 272;     // padded feature count
constexpr int kMRows = 384;  // omega rows in shared memory (3 chunks x 128 lanes)
constexpr int kRing = 6;     // K, V, Q of two items
constexpr uint32_t kSlabBytes = kTile * 128;    // 16384
constexpr uint32_t kOmegaBytes = kMRows * 128;  // 49152
constexpr uint32_t kCtxSlabBytes = 80 * 128;    // 10240
constexpr uint32_t kOffOmega = 0;
constexpr uint32_t kOffRing = kOffOmega + kOmegaBytes;
constexpr uint32_t kOffCslab = kOffRing + kRing * kSlabBytes;  // after the ring: LBO of [V | 1] > 0
constexpr uint32_t kOffCtx = kOffCslab + kSlabBytes;
constexpr uint32_t kOffBar = kOffCtx + 5 * kCtxSlabBytes;
constexpr uint32_t kOffScratch = kOffBar + 256;
constexpr uint32_t kScratchFloats = 4 * 128 + 4 * 128 + 128 + 32 + 4 * 128;  // key diag partials, row maxima, per-token diag, warp maxima, query diag partials
constexpr uint32_t kSmemBytes = kOffScratch + kScratchFloats * 4 + 1024;
static_assert(kOffRing % 1024 == 0 && kOffCslab % 1024 == 0 && kOffCtx % 1024 == 0, "align");
static_assert(kSmemBytes <= 232448, "shared memory budget");

// developer timeline (build with -DRFK_COL_TIMELINE): clock64 stamps of CTA 0's issuer and first feature warp for
// items 10 and 11, printed at kernel end
#ifdef RFK_COL_TIMELINE
#define RFK_TL(arr, k) do { if (blockIdx.x == 0 && lane == 0 && (n == 10 || n == 11)) arr[(n - 10) * 12 + (k)] = clock64(); } while (0)
#else
#define RFK_TL(arr, k) do { } while (0)
#endif

constexpr uint32_t kColU0 = 0, kColU1 = 128, kColU2 = 256, kColC = 384, kColS = 464;

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
favor_col_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const FavorTmParams p) {
  constexpr int kDt = F16 ? RFK_F16 : RFK_BF16;
  constexpr uint32_t kFmt = F16 ? kIdescBf16Bits : 0u;  // instruction-descriptor A / B format bits
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_omega = base + kOffOmega, s_ring = base + kOffRing, s_cslab = base + kOffCslab, s_ctx = base + kOffCtx;
  const uint32_t bars = base + kOffBar;
  auto bar_tfull = [&](uint32_t s) { return bars + 8u * s; };            // [kRing] tile landed
  auto bar_tempty = [&](uint32_t s) { return bars + 48u + 8u * s; };     // [kRing] tile consumed
  auto bar_ufull = [&](uint32_t c) { return bars + 96u + 8u * c; };      // key projection chunk c complete
  auto bar_fready = [&](uint32_t c) { return bars + 120u + 8u * c; };    // k'^T chunk c stored
  auto bar_ctxfull = [&](uint32_t c) { return bars + 144u + 8u * c; };   // ctx^T block c complete
  const uint32_t bar_ctxread1 = bars + 168u;                            // ctx^T block 1 read out of U0
  const uint32_t bar_ctxready = bars + 176u;                            // all blocks in shared memory, C / U1 free
  auto bar_uqfull = [&](uint32_t c) { return bars + 184u + 8u * c; };    // query projection chunk c complete
  auto bar_qready = [&](uint32_t c) { return bars + 208u + 8u * c; };    // q' chunk c stored
  const uint32_t bar_outfull = bars + 232u;
  const uint32_t tmem_slot = bars + 248u;
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kOffScratch);
  float* partk = scratch;         // [4][128] partial |k|^2 of the four channel quarters
  float* rmaxs = scratch + 512;   // [4][128] partial row maxima
  float* dgl = scratch + 1024;    // [128] log2(e) * 0.5 dn^2 |k|^2 per token of the key tile
  float* red = scratch + 1152;    // [16] per-warp maxima
  float* partq = scratch + 1184;  // [4][128] partial |q|^2

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t istride = gridDim.x, items = (uint32_t)p.items;  // (the launcher checks that items fit 32 bits)

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < kRing; ++s) {
      mbar_init(bar_tfull(s), 1);
      mbar_init(bar_tempty(s), 1);
    }
    for (uint32_t c = 0; c < 3; ++c) {
      mbar_init(bar_ufull(c), 1);
      mbar_init(bar_fready(c), kFeatWarps);
      mbar_init(bar_ctxfull(c), 1);
      mbar_init(bar_uqfull(c), 1);
      mbar_init(bar_qready(c), kFeatWarps);
    }
    mbar_init(bar_ctxread1, kFeatWarps);
    mbar_init(bar_ctxready, kFeatWarps);
    mbar_init(bar_outfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  {
    // omega' = dn * proj (rows >= m zero), K-major swizzled; constant slab: column 0 = 1
    const float dn = 0.35355339059327373f;  // 64^-1/4
    for (int i = threadIdx.x; i < kMRows * 8; i += kThreads) {
      const int r = i >> 3, c = (i & 7) * 8;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = r < p.m ? __ldg(p.proj + (int64_t)r * 64 + c + j) * dn : 0.f;
      uint4 v;
      v.x = pack_h16x2(f[0], f[1], kDt); v.y = pack_h16x2(f[2], f[3], kDt);
      v.z = pack_h16x2(f[4], f[5], kDt); v.w = pack_h16x2(f[6], f[7], kDt);
      st_shared_v4(s_omega + sw128_offset(r, c), v);
    }
    for (int i = threadIdx.x; i < kTile * 8; i += kThreads) {
      const int r = i >> 3, c = (i & 7) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (c == 0) v.x = F16 ? 0x00003C00u : 0x00003F80u;  // 1.0 in element 0
      st_shared_v4(s_cslab + sw128_offset(r, c), v);
    }
    // ctx rows 65..79 (never written by the read-out) feed never-read accumulator columns: zero once
    for (int i = threadIdx.x; i < (int)(5 * kCtxSlabBytes / 16); i += kThreads)
      st_shared_v4(s_ctx + i * 16, make_uint4(0, 0, 0, 0));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  // item n of this CTA uses ring slots 3 (n & 1) + {0: K, 1: V, 2: Q}; each slot is filled once per two items
  auto slot_of = [](uint32_t n, uint32_t i) { return 3u * (n & 1u) + i; };

  if (warp == 0) {
    if (lane == 0) {
      // =================== TMA producer ===================
      auto coords = [&](uint32_t it32, int& h, int& g0, int& g1) {
        const uint32_t g = it32 / (uint32_t)p.heads;
        h = (int)(it32 - g * (uint32_t)p.heads);
        g1 = (int)(g / (uint32_t)p.G0);
        g0 = (int)(g - (uint32_t)g1 * (uint32_t)p.G0);
      };
      // L2 prefetch two items ahead of the loads
      auto prefetch_item = [&](uint32_t item) {
        if (item >= items) return;
        int h, g0, g1;
        coords(item, h, g0, g1);
        tma_prefetch_4d(&tm_k, h * 64, 0, g0, g1);
        tma_prefetch_4d(&tm_v, h * 64, 0, g0, g1);
        tma_prefetch_4d(&tm_q, h * 64, 0, g0, g1);
      };
      prefetch_item(blockIdx.x);
      prefetch_item(blockIdx.x + istride);
      uint32_t n = 0;
      for (uint32_t item = blockIdx.x; item < items; item += istride, ++n) {
        prefetch_item(item + 2 * istride);
        int h, g0, g1;
        coords(item, h, g0, g1);
        const uint32_t par = (n >> 1) & 1u;
        const CUtensorMap* tms[3] = {&tm_k, &tm_v, &tm_q};
#pragma unroll
        for (uint32_t i = 0; i < 3; ++i) {
          const uint32_t slot = slot_of(n, i);
          mbar_wait(bar_tempty(slot), par ^ 1u);
          mbar_arrive_expect_tx(bar_tfull(slot), kSlabBytes);
          tma_load_4d(tms[i], bar_tfull(slot), s_ring + slot * kSlabBytes, h * 64, 0, g0, g1);
        }
      }
    }
  } else if (warp == 1) {
    // =================== MMA issuer ===================
    // all 32 lanes run the warp-uniform control flow (mbarrier waits), one elected lane issues
    const uint64_t d_omega = umma_desc_sw128(s_omega);  // + c * 1024 + 2 k
    const uint64_t d_ring = umma_desc_sw128(s_ring);    // + slot * 1024 + 2 k
    const uint64_t d_ctx = umma_desc_sw128(s_ctx);      // + slab * 640 + 2 k
    constexpr uint32_t kIdU = umma_idesc_bf16(128, 128) ^ kFmt;
    constexpr uint32_t kIdU16 = umma_idesc_bf16(128, 16) ^ kFmt;
    constexpr uint32_t kIdCtx = idesc_bf16_major(128, 80, 0, 1) ^ kFmt;
    constexpr uint32_t kIdOut = umma_idesc_bf16(128, 80) ^ kFmt;
    // key projections of item n: U^T_c[m, tok] = Omega_c . K^T (features on the lanes) into U0 | U1 | U2
    // (chunk 1 goes into U1, which the query phase never uses: it is issued EARLY, before the previous item's output
    // MMAs, so that the feature warps have a projection to take maxima of while those MMAs run)
    auto proj_keys = [&](uint32_t n, bool early) {
      mbar_wait(bar_tfull(slot_of(n, 0)), (n >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dx = d_ring + (uint64_t)(slot_of(n, 0) * 1024u);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if ((c == 1) != early) continue;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + 128u * c, d_omega + (uint64_t)(c * 1024) + 2 * k, dx + 2 * k, kIdU, k > 0);
          umma_commit(bar_ufull(c));
        }
      }
      __syncwarp();
    };
    uint32_t n = 0;
#ifdef RFK_COL_TIMELINE
    long long tli[24] = {};
#endif
    if (blockIdx.x < items) {
      proj_keys(0, true);
      proj_keys(0, false);
    }
    for (uint32_t item = blockIdx.x; item < items; item += istride, ++n) {
      const uint32_t par = n & 1u, tpar = (n >> 1) & 1u;
      const uint32_t ks = slot_of(n, 0), vs = slot_of(n, 1), qs = slot_of(n, 2);
      // ---- context: ctx^T_c[128 m x 80] = k'^T_c (A, TMEM, K = tokens) . [V | 1] (B, MN-major; second chunk = cslab)
      RFK_TL(tli, 0);
      mbar_wait(bar_tfull(vs), tpar);
      // (block C: every feature warp has read the previous item's out | den out of it before it publishes the
      // first feature chunk of this item, which the context MMAs below wait for)
      RFK_TL(tli, 1);
      const uint64_t dbv = desc_mn_sw128(s_ring + vs * kSlabBytes, s_cslab - s_ring - vs * kSlabBytes);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t dcol = c == 0 ? kColC : c == 1 ? kColU0 : kColU1;
        mbar_wait(bar_fready(c), par);
        if (c == 0) RFK_TL(tli, 2);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ts(tmem + dcol, tmem + 128u * c + 32u * (k >> 1) + 8u * (k & 1), dbv + 128 * k, kIdCtx, k > 0);
          umma_commit(bar_ctxfull(c));
          if (c == 2) {
            umma_commit(bar_tempty(ks));
            umma_commit(bar_tempty(vs));
          }
        }
        __syncwarp();
      }
      // ---- query projections: U_c[tok, m] = Q . Omega_c^T; chunk 0 -> U2, chunk 2 (16 features) -> S at once (both
      //      regions were last touched by MMAs issued above), chunk 1 -> U0 once ctx^T_1 has been read out of it
      RFK_TL(tli, 3);
      mbar_wait(bar_tfull(qs), tpar);
      tc_fence_after();
      const uint64_t dq = d_ring + (uint64_t)(qs * 1024u);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + kColU2, dq + 2 * k, d_omega + 2 * k, kIdU, k > 0);
        umma_commit(bar_uqfull(0));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + kColS, dq + 2 * k, d_omega + 2048 + 2 * k, kIdU16, k > 0);
        umma_commit(bar_uqfull(2));
      }
      __syncwarp();
      RFK_TL(tli, 4);
      mbar_wait(bar_ctxread1, par);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + kColU0, dq + 2 * k, d_omega + 1024 + 2 * k, kIdU, k > 0);
        umma_commit(bar_uqfull(1));
      }
      __syncwarp();
      // ---- output: out|den [128 tok x 80] = sum_c q'_c (A, TMEM, K = features) . ctx_c (B, K-major over m)
      RFK_TL(tli, 5);
      mbar_wait(bar_ctxready, par);
      RFK_TL(tli, 6);
      const bool has_next = item + istride < items;
      if (has_next) proj_keys(n + 1, true);  // U1 is free: ctx^T_2 has been read out of it
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = cc == 0 ? 1 : cc == 1 ? 0 : 2;  // the order the feature warps publish in
        const uint32_t acol = c == 0 ? kColU2 : c == 1 ? kColU0 : kColS;
        mbar_wait(bar_qready(c), par);
        if (cc == 0) RFK_TL(tli, 7);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t db = d_ctx + (uint64_t)(2 * c * 640);
          const int nk = c == 2 ? 1 : 8;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < nk)
              umma_bf16_ts(tmem + kColC, tmem + acol + 32u * (k >> 1) + 8u * (k & 1), db + (k >> 2) * 640 + 2 * (k & 3), kIdOut,
                           (cc > 0 || k > 0));
          if (c == 2) {
            umma_commit(bar_outfull);
            umma_commit(bar_tempty(qs));
          }
        }
        __syncwarp();
      }
      // ---- the next item's key projections follow the output MMAs in the same issue stream
      RFK_TL(tli, 8);
      if (has_next) proj_keys(n + 1, false);
      RFK_TL(tli, 9);
    }
#ifdef RFK_COL_TIMELINE
    if (blockIdx.x == 0 && lane == 0)
      for (int i = 0; i < 24; ++i) printf("issuer item %d stamp %d: %lld\n", 10 + i / 12, i % 12, tli[i] - tli[0]);
    if (blockIdx.x == 0 && lane == 0) printf("issuer base %lld\n", tli[0]);
#endif
  } else {
    // =================== feature / epilogue warps ===================
    const int fw = warp - 2;         // 0..15
    const int lg = warp & 3;         // TMEM lane group this warp may touch
    const int cq = fw >> 2;          // column quarter owned by this warp
    const int row = lg * 32 + lane;  // TMEM lane: token row (query side) / feature row of the chunk (key side)
    const uint32_t t_lane = ((uint32_t)(lg * 32) << 16);
    const uint32_t tl = tmem + t_lane;
    constexpr float kLog2e = 1.4426950408889634f;
    const uint32_t eps2 = pack_h16x2(1e-4f, 1e-4f, kDt);
    const int ntok = p.tokens;

    // two accumulator values -> one packed 16-bit feature pair exp(x - s) + eps. The exponentials are what bounds
    // this kernel once the schedule is tight (16 MUFU results per clock and SM: 69.6 K per item): POLY pairs take
    // theirs on the FMA pipe instead (ex2_poly, relative error 2.7e-6, far below the 16-bit rounding that follows)
    auto feat2 = [&](uint32_t r0, uint32_t r1, float s0, float s1, bool poly) -> uint32_t {
      const float t0 = fmaf(__uint_as_float(r0), kLog2e, -s0), t1 = fmaf(__uint_as_float(r1), kLog2e, -s1);
      const float e0 = poly ? ex2_poly(t0) : ex2_approx(t0), e1 = poly ? ex2_poly(t1) : ex2_approx(t1);
      return add_h16x2<F16>(pack_h16x2(e0, e1, kDt), eps2);
    };
    auto publish = [&](uint32_t bar, bool stored) {
      if (stored) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // 0.5 * dn^2 * |x|^2 partial of token `row` over channels [16 cq, 16 cq + 16) of a ring tile
    auto diag_partial = [&](float* part, uint32_t slot, uint32_t tpar) {
      mbar_wait(bar_tfull(slot), tpar);  // TMA bytes visible to this thread
      const uint32_t tile = s_ring + slot * kSlabBytes;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint4 v = ld_shared_v4(tile + sw128_offset(row, cq * 16 + j * 8));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = h16_to_float((uint16_t)(w[i] & 0xffffu), kDt), b = h16_to_float((uint16_t)(w[i] >> 16), kDt);
          s = fmaf(a, a, s);
          s = fmaf(b, b, s);
        }
      }
      part[cq * 128 + row] = s;
    };
    auto diag_total = [&](const float* part) { return (part[row] + part[128 + row] + part[256 + row] + part[384 + row]) * (0.5f * 0.125f); };
    auto put = [&](int m, uint32_t nn, float v) {  // ctx[m][nn] into the K-major (over m) B operand of the output MMA
      const uint32_t mc = (uint32_t)m & 63u;
      st_shared_h16<F16>(s_ctx + (uint32_t)(m >> 6) * kCtxSlabBytes + (nn >> 3) * 1024u + (nn & 7u) * 128u +
                             ((((mc >> 3) ^ nn) & 7u) << 4) + (mc & 7u) * 2u,
                         v);
    };
    // out/den epilogue: warp (lg, cq) stores channels [16 cq, 16 cq + 16) of its 32 tokens
    auto item_out_ptr = [&](uint32_t it32) {  // once per item
      const uint32_t g = it32 / (uint32_t)p.heads, h = it32 - g * (uint32_t)p.heads;
      const uint32_t g1 = g / (uint32_t)p.G0, g0 = g - g1 * (uint32_t)p.G0;
      return reinterpret_cast<uint16_t*>(p.out) + (int64_t)g1 * p.ogs1 + (int64_t)g0 * p.ogs0 + (int64_t)row * p.ots + h * 64 + 16 * cq;
    };
    auto epilogue = [&](uint16_t* out_ptr, uint32_t n) {
      mbar_wait(bar_outfull, n & 1u);
      tc_fence_after();
      uint32_t rd[1], r0[16];
      tmem_ld_32x1p(tl + kColC + 64, rd);  // column 64 = normaliser
      tmem_ld_32x16p(tl + kColC + 16u * (uint32_t)cq, r0);
      tmem_ld_wait();
      tc_fence_before();
      if (row < ntok) {
        const float inv = __fdividef(1.f, __uint_as_float(rd[0]));
        uint4* op = reinterpret_cast<uint4*>(out_ptr);
        uint4 w;
        w.x = pack_h16x2(__uint_as_float(r0[0]) * inv, __uint_as_float(r0[1]) * inv, kDt);
        w.y = pack_h16x2(__uint_as_float(r0[2]) * inv, __uint_as_float(r0[3]) * inv, kDt);
        w.z = pack_h16x2(__uint_as_float(r0[4]) * inv, __uint_as_float(r0[5]) * inv, kDt);
        w.w = pack_h16x2(__uint_as_float(r0[6]) * inv, __uint_as_float(r0[7]) * inv, kDt);
        op[0] = w;
        w.x = pack_h16x2(__uint_as_float(r0[8]) * inv, __uint_as_float(r0[9]) * inv, kDt);
        w.y = pack_h16x2(__uint_as_float(r0[10]) * inv, __uint_as_float(r0[11]) * inv, kDt);
        w.z = pack_h16x2(__uint_as_float(r0[12]) * inv, __uint_as_float(r0[13]) * inv, kDt);
        w.w = pack_h16x2(__uint_as_float(r0[14]) * inv, __uint_as_float(r0[15]) * inv, kDt);
        op[1] = w;
      }
    };

    uint32_t n = 0;
#ifdef RFK_COL_TIMELINE
    long long tlf[24] = {};
#endif
    uint16_t* prev_out = nullptr;  // output rows of the item whose out | den is still in block C
    uint32_t raw[32];
    for (uint32_t item = blockIdx.x; item < items; item += istride, ++n) {
      const uint32_t par = n & 1u, tpar = (n >> 1) & 1u;
      RFK_TL(tlf, 0);
      uint16_t* const cur_out = item_out_ptr(item);
      // ---- per-token |k|^2 term of the exponent: needs the K tile only, so it runs while the previous item's
      //      output MMAs are still in flight ----
      diag_partial(partk, slot_of(n, 0), tpar);
      named_bar_sync(1, kFeatThreads);
      if (cq == 0) dgl[row] = diag_total(partk) * kLog2e;
      // ---- key stabiliser: global max of Omega' . K^T over the valid features / tokens; this thread holds feature
      //      row (128 c + row) x tokens [32 cq, 32 cq + 32) of every chunk. Chunk 1 was projected early (before the
      //      previous item's output MMAs), chunk 0 comes last and stays in registers for the feature pass ----
      const int klim = ntok - 32 * cq;  // token columns >= klim are padding
      float kmx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c = cc == 2 ? 0 : cc + 1;
        mbar_wait(bar_ufull(c), par);
        tc_fence_after();
        tmem_ld_32x32p(tl + 128u * c + 32u * (uint32_t)cq, raw);
        tmem_ld_wait();
        if (128 * c + row < p.m) {
          if (klim >= 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) kmx = fmaxf(kmx, __uint_as_float(raw[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < klim) kmx = fmaxf(kmx, __uint_as_float(raw[i]));
          }
        }
        if (cc == 0) {
          if (prev_out) epilogue(prev_out, n - 1u);
          prev_out = cur_out;
          RFK_TL(tlf, 1);
        }
      }
      RFK_TL(tlf, 2);
      kmx = warp_max(kmx);
      if (lane == 0) red[fw] = kmx;
      named_bar_sync(1, kFeatThreads);  // (also publishes dgl)
      float gl = red[0];
#pragma unroll
      for (int i = 1; i < kFeatWarps; ++i) gl = fmaxf(gl, red[i]);
      gl *= kLog2e;
      RFK_TL(tlf, 3);
      // ---- keys: k'^T chunks, in place, feed the context MMAs (chunk 0 is still in registers) ----
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c > 0) {
          tmem_ld_32x32p(tl + 128u * c + 32u * (uint32_t)cq, raw);
          tmem_ld_wait();
        }
        uint32_t pk[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(dgl + 32 * cq + 4 * q);
          pk[2 * q] = feat2(raw[4 * q], raw[4 * q + 1], s4.x + gl, s4.y + gl, false);
          pk[2 * q + 1] = feat2(raw[4 * q + 2], raw[4 * q + 3], s4.z + gl, s4.w + gl, (q & 1) == 1);
        }
        if (128 * c + row >= p.m) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        } else if (klim < 32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (2 * i >= klim) pk[i] = 0u;
            else if (2 * i + 1 >= klim) pk[i] &= 0x0000ffffu;
          }
        }
        tmem_st_32x16(tl + 128u * c + 32u * (uint32_t)cq, pk);
        publish(bar_fready(c), true);
      }
      // ---- context read-out: ctx^T blocks (C, U0, U1) -> 16-bit K-major shared memory. Warp (lg, cq) converts
      //      columns [16 cq, 16 cq + 16) of blocks 0 and 1 for its 32 feature rows; the normaliser column 64 and the
      //      16-row block 2 are spread over the column quarters ----
      {
        uint32_t r[16], r2[1];
        RFK_TL(tlf, 4);
        mbar_wait(bar_ctxfull(0), par);
        RFK_TL(tlf, 5);
        tc_fence_after();
        tmem_ld_32x16p(tl + kColC + 16u * (uint32_t)cq, r);
        if (cq == 0) tmem_ld_32x1p(tl + kColC + 64u, r2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) put(row, 16u * cq + i, __uint_as_float(r[i]));
        if (cq == 0) put(row, 64u, __uint_as_float(r2[0]));
        mbar_wait(bar_ctxfull(1), par);
        tc_fence_after();
        tmem_ld_32x16p(tl + kColU0 + 16u * (uint32_t)cq, r);
        if (cq == 1) tmem_ld_32x1p(tl + kColU0 + 64u, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ctxread1);
#pragma unroll
        for (int i = 0; i < 16; ++i) put(128 + row, 16u * cq + i, __uint_as_float(r[i]));
        if (cq == 1) put(128 + row, 64u, __uint_as_float(r2[0]));
        diag_partial(partq, slot_of(n, 2), tpar);  // (needs the Q tile only: fills the wait for block 2)
        mbar_wait(bar_ctxfull(2), par);
        tc_fence_after();
        if (lg == 0) {  // block 2: features 256..271 live in lanes 0..15
          tmem_ld_32x16p(tmem + kColU1 + 16u * (uint32_t)cq, r);
          tmem_ld_32x1p(tmem + kColU1 + 64u, r2);
          tmem_ld_wait();
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) put(256 + lane, 16u * cq + i, __uint_as_float(r[i]));
            if (cq == 3) put(256 + lane, 64u, __uint_as_float(r2[0]));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ctxready);
      }
      RFK_TL(tlf, 6);
      // ---- query stabiliser: per token row the max over the valid features; this thread holds token `row` x
      //      features [128 c + 32 cq, + 32) (chunk 2: 16 features, cq == 0 warps). Chunks arrive as 0, 2, 1 ----
      float rmx = -INFINITY;
      {
        mbar_wait(bar_uqfull(0), par);
        tc_fence_after();
        tmem_ld_32x32p(tl + kColU2 + 32u * (uint32_t)cq, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) rmx = fmaxf(rmx, __uint_as_float(raw[i]));  // features < 128 <= m
        mbar_wait(bar_uqfull(2), par);
        tc_fence_after();
        if (cq == 0) {
          uint32_t r[16];
          tmem_ld_32x16p(tl + kColS, r);
          tmem_ld_wait();
          const int lim = p.m - 256;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < lim) rmx = fmaxf(rmx, __uint_as_float(r[i]));
        }
        mbar_wait(bar_uqfull(1), par);
        tc_fence_after();
        tmem_ld_32x32p(tl + kColU0 + 32u * (uint32_t)cq, raw);
        tmem_ld_wait();
        const int lim = p.m - (128 + 32 * cq);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < lim) rmx = fmaxf(rmx, __uint_as_float(raw[i]));
      }
      RFK_TL(tlf, 7);
      rmaxs[cq * 128 + row] = rmx;
      named_bar_sync(1, kFeatThreads);
      const float sub =
          (diag_total(partq) + fmaxf(fmaxf(rmaxs[row], rmaxs[128 + row]), fmaxf(rmaxs[256 + row], rmaxs[384 + row]))) * kLog2e;
      RFK_TL(tlf, 8);
      // ---- queries: q' chunks, in place, feed the output MMAs; chunk 1 is still in registers: order 1, 0, 2 ----
      {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = feat2(raw[2 * i], raw[2 * i + 1], sub, sub, (i & 3) == 3);
        tmem_st_32x16(tl + kColU0 + 32u * (uint32_t)cq, pk);
        publish(bar_qready(1), true);
        tmem_ld_32x32p(tl + kColU2 + 32u * (uint32_t)cq, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = feat2(raw[2 * i], raw[2 * i + 1], sub, sub, (i & 3) == 3);
        tmem_st_32x16(tl + kColU2 + 32u * (uint32_t)cq, pk);
        publish(bar_qready(0), true);
        if (cq == 0) {
          uint32_t r[16];
          tmem_ld_32x16p(tl + kColS, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = feat2(r[2 * i], r[2 * i + 1], sub, sub, false);
          tmem_st_32x8(tl + kColS, pk);
          publish(bar_qready(2), true);
        } else {
          publish(bar_qready(2), false);
        }
      }
      RFK_TL(tlf, 9);
    }
#ifdef RFK_COL_TIMELINE
    if (blockIdx.x == 0 && warp == 2 && lane == 0) {
      for (int i = 0; i < 24; ++i) printf("feat item %d stamp %d: %lld\n", 10 + i / 12, i % 12, tlf[i] - tlf[0]);
      printf("feat base %lld\n", tlf[0]);
    }
#endif
    if (prev_out) epilogue(prev_out, n - 1u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool F16>
int launch_col(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FavorTmParams& p,
               cudaStream_t stream) {
  static PerDeviceOnce once;
  const int cfg_rc = per_device_once(once, []() {
    return cuda_status(cudaFuncSetAttribute(favor_col_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  int grid = num_sms();
  if (p.items < grid) grid = (int)p.items;
  favor_col_kernel<F16><<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, p);
  return post_launch();
}

}  // namespace

// softmax kernel, tokens <= 128; anything else -> RFK_ERR_UNSUPPORTED (the caller falls through to favor_tm_launch)
int favor_col_launch(const rfk_favor_desc* d, cudaStream_t stream) {
  if (d->kind != 0 || d->tokens > kTile || d->tokens < 1) return RFK_ERR_UNSUPPORTED;
  if (d->m_features > kMP || d->m_features <= 256) return RFK_ERR_UNSUPPORTED;  // three chunks: 128 | 128 | 1..16
  if (!aligned16(d->q) || !aligned16(d->k) || !aligned16(d->v) || !aligned16(d->out)) return RFK_ERR_UNSUPPORTED;
  if (d->ts % 8 || d->gs[0] % 8 || d->gs[1] % 8 || d->out_ts % 8 || d->out_gs[0] % 8 || d->out_gs[1] % 8)
    return RFK_ERR_UNSUPPORTED;
  int rc = check_arch();
  if (rc != RFK_OK) return rc;
  CUtensorMap tq, tk, tv;
  if ((rc = make_head_tmap(&tq, d->q, d)) != RFK_OK) return rc;
  if ((rc = make_head_tmap(&tk, d->k, d)) != RFK_OK) return rc;
  if ((rc = make_head_tmap(&tv, d->v, d)) != RFK_OK) return rc;
  FavorTmParams p{};
  p.proj = d->proj; p.out = d->out; p.m = d->m_features; p.heads = d->heads;
  p.tokens = (int)d->tokens; p.G0 = d->G[0]; p.G1 = d->G[1];
  p.items = d->G[0] * d->G[1] * d->heads;
  if (p.items > 0x7fffffffLL || d->G[0] > 0x7fffffffLL) return RFK_ERR_UNSUPPORTED;  // 32-bit item decode in the kernel
  p.ogs0 = d->out_gs[0]; p.ogs1 = d->out_gs[1]; p.ots = d->out_ts;
  return d->io_dtype == RFK_F16 ? launch_col<true>(tq, tk, tv, p, stream) : launch_col<false>(tq, tk, tv, p, stream);
}

}  // namespace rfk
