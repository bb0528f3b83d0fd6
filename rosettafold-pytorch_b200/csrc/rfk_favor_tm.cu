// rfk_favor_tm.cu — fused Performer FAVOR+ attention with the random FEATURES KEPT IN TENSOR MEMORY
// (tcgen05.st + A-operand-from-TMEM MMAs), bf16 operands, fp32 accumulation. Replaces
// performer_pytorch FastAttention (softmax_kernel / generalized_kernel + linear_attention) as called at
// rosettafold_pytorch.py:313-318 (MSA columns, softmax kernel) and :505-518 (pair axes, ReLU kernel).
//
// One persistent CTA per SM walks a stream of 128 x 128 JOBS; nothing of size tokens x m ever leaves the
// SM and — unlike the round-1 kernel (rfk_favor_tc.cu) — the features never touch shared memory either:
// the feature warps read the projection accumulator U out of TMEM, apply the feature map in registers and
// write the packed bf16 features back IN PLACE over the accumulator (tcgen05.st); the consuming MMA takes
// them as its A operand straight from TMEM. Per job that removes 32 KB of shared-memory stores, 32 KB of
// shared-memory operand reads and a fence.proxy.async from the round-1 formulation, whose measured limit
// was shared-memory bandwidth (profiles/r01_favor_pipeline_notes.md).
//
// A-from-TMEM fixes the orientation of every product (A rows sit on TMEM lanes, A is K-major):
//   key side    U^T[m, tok] = Omega_c (A, smem) . K_tile^T (B, smem)    features ON THE LANES
//               k'^T = phi(U^T) -> TMEM (in place)
//               ctx^T_c[m, d|1] (+)= k'^T (A, TMEM, K = tokens) . [V | 1] (B, smem, MN-major)
//   query side  U[tok, m]   = Q_tile (A, smem) . Omega_c^T (B, smem)    tokens on the lanes
//               q' = phi(U) -> TMEM (in place)
//               out|den[tok, d|1] (+)= q' (A, TMEM, K = features) . ctx_c (B, smem, K-major)
// The 272 (padded) features are split into chunks c of 128 | 128 | 16. Job types per item (group, head):
//   relu kernel  K(t,c)*, read-out of ctx, Q(t,c)*;
//   softmax      KMAX(t,c)*, K(t,c)*, read-out, then per tile QMAX(t,c)*, Q(t,c)*   (stabiliser passes).
//
// Roles (18 warps):
//   warp 0 lane 0  TMA producer: K/V/Q tiles into a 5-slot ring + an L2 prefetch cursor 8 tiles ahead
//   warp 1         MMA issuer: per job one burst = the context / output MMAs of job j (A from TMEM) followed by
//                  the projection of job j + 2 into the same U slot (same-thread issue order hands the slot on)
//   warps 2..17    16 feature warps, all on every job: warp (lane group lg = warp % 4, column quarter cq) owns
//                  32 lanes x 32 accumulator columns: tcgen05.ld x32 -> feature map -> 16 packed words ->
//                  tcgen05.st x16 over the first half of its own columns (no warp ever writes columns another
//                  warp still has to read). MMA k-step s (16 features / tokens) therefore reads A at column
//                  32 (s / 2) + 8 (s % 2) of the slot.
// TMEM (512 columns): ctx^T blocks b = 0..2 (lanes = m - 128 b; 80 columns = d | 1) at [0,240);
//   U slots 2 x 128 columns at [256,512); out|den accumulators D3[s] alias ctx blocks 0 / 1 (dead in the
//   query phase; reuse ordered through the d3free barriers).
// Shared memory (1024-byte aligned tiles, 128-byte swizzle):
//   omega' [384][64] bf16 K-major (rows >= m zero) | tile ring 5 x [128 tok][64 d] | cslab [128 tok][64],
//   column 0 = 1 (second MN chunk of "[V | 1]") | ctx 5 x [80][64 m] K-major B of the output MMA.
#include "rfk_favor_device.cuh"

namespace rfk {

namespace {

constexpr int kFeatWarps = 16;
constexpr int kFeatThreads = 32 * kFeatWarps;
constexpr int kThreads = 32 * (2 + kFeatWarps);  // TMA producer, MMA issuer, feature warps
constexpr int kMP = 272;     // padded feature count
constexpr int kMRows = 384;  // omega rows in shared memory (3 chunks x 128 lanes)
constexpr int kRing = 5;
constexpr uint32_t kSlabBytes = kTile * 128;    // 16384
constexpr uint32_t kOmegaBytes = kMRows * 128;  // 49152
constexpr uint32_t kCtxSlabBytes = 80 * 128;    // 10240
constexpr uint32_t kOffOmega = 0;
constexpr uint32_t kOffRing = kOffOmega + kOmegaBytes;
constexpr uint32_t kOffCslab = kOffRing + kRing * kSlabBytes;  // after the ring: LBO of [V | 1] > 0
constexpr uint32_t kOffCtx = kOffCslab + kSlabBytes;
constexpr uint32_t kOffBar = kOffCtx + 5 * kCtxSlabBytes;
constexpr uint32_t kOffScratch = kOffBar + 256;
constexpr uint32_t kScratchFloats = 4 * 128 + 4 * 128 + 128 + 32;  // diag partials, row maxima, per-token sub, warp maxima
constexpr uint32_t kSmemBytes = kOffScratch + kScratchFloats * 4 + 1024;
static_assert(kOffRing % 1024 == 0 && kOffCslab % 1024 == 0 && kOffCtx % 1024 == 0, "align");
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr uint32_t kColCtx = 0, kColU = 256;

// Walks the PASSES of one CTA in issue order: item -> pass (kind, tile t); every pass has three jobs (feature
// chunks c = 0, 1, 2 of 128 | 128 | 16 features). The iterator also mirrors the producer's ring allocation (the
// producer loads tiles in exactly this order), so the pass knows the ring slots of its tiles.
constexpr int kPassM = 0, kPassK = 1, kPassX = 2, kPassQ = 3;  // key max, keys, query max, queries
struct PassInfo {
  int kind, t;
  bool valid;
  uint32_t a_slot, a_par;  // the pass's K (key side) or Q (query side) tile
  uint32_t v_slot, v_par;  // the pass's V tile (kPassK)
  uint32_t j0;             // running index of the pass's first job (U slot = j % 2)
  __device__ bool key_side() const { return kind == kPassM || kind == kPassK; }
};
template <int KIND>
struct PassIter {
  uint32_t item, items, istride;  // (the launcher checks that items fit 32 bits)
  int nt, P, ps;
  uint32_t r_slot, r_par, j;
  PassInfo last;
  __device__ void take(uint32_t& slot, uint32_t& par) {
    slot = r_slot; par = r_par;
    if (++r_slot == kRing) { r_slot = 0; r_par ^= 1u; }
  }
  __device__ void init(uint32_t first, uint32_t n_items, uint32_t stride, int n_tiles) {
    item = first; items = n_items; istride = stride; nt = n_tiles;
    P = KIND == 0 ? 4 * nt : 2 * nt;
    ps = 0; r_slot = 0; r_par = 0; j = 0;
    last.a_slot = last.a_par = last.v_slot = last.v_par = 0;
  }
  __device__ PassInfo next() {
    PassInfo x = last;
    x.valid = item < items;
    if (!x.valid) return x;
    if (KIND == 1) {
      if (ps < nt) { x.t = ps; x.kind = kPassK; } else { x.t = ps - nt; x.kind = kPassQ; }
    } else if (ps < nt) { x.t = ps; x.kind = kPassM; }
    else if (ps < 2 * nt) { x.t = ps - nt; x.kind = kPassK; }
    else { const int r = ps - 2 * nt; x.t = r >> 1; x.kind = (r & 1) ? kPassQ : kPassX; }
    if (!(KIND == 0 && x.kind == kPassQ)) take(x.a_slot, x.a_par);  // a softmax Q pass reuses its QMAX tile
    if (x.kind == kPassK) take(x.v_slot, x.v_par);
    x.j0 = j;
    j += 3u;
    if (++ps == P) { ps = 0; item += istride; }
    last = x;
    return x;
  }
};

// KIND 0: softmax kernel (exp features, stabilisers), 1: generalized ReLU kernel. F16: q, k, v, out, the projection
// matrix, the features and the context are IEEE half instead of bf16 (same tensor-core rate, 11 instead of 8 significand
// bits; meant for the softmax kernel, whose features are bounded by 1 + eps: the ReLU features and their context sums
// over up to thousands of tokens are unbounded and stay in bf16)
template <int KIND, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
favor_tm_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FavorTmParams p) {
  constexpr int kDt = F16 ? RFK_F16 : RFK_BF16;
  constexpr uint32_t kFmt = F16 ? kIdescBf16Bits : 0u;  // instruction-descriptor A / B format bits
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_omega = base + kOffOmega, s_ring = base + kOffRing, s_cslab = base + kOffCslab, s_ctx = base + kOffCtx;
  const uint32_t bars = base + kOffBar;
  auto bar_tfull = [&](uint32_t s) { return bars + 8u * s; };            // [kRing]
  auto bar_tempty = [&](uint32_t s) { return bars + 40u + 8u * s; };     // [kRing]
  auto bar_ufull = [&](uint32_t s) { return bars + 80u + 8u * s; };      // U accumulator of the slot complete
  auto bar_fready = [&](uint32_t s) { return bars + 112u + 8u * s; };    // features of the slot stored (or max taken)
  auto bar_d3full = [&](uint32_t s) { return bars + 128u + 8u * s; };
  auto bar_d3free = [&](uint32_t s) { return bars + 144u + 8u * s; };
  const uint32_t bar_ctxfull = bars + 160u, bar_ctxready = bars + 168u;
  const uint32_t tmem_slot = bars + 176u;
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kOffScratch);
  float* part = scratch;          // [4][128] partial |x|^2 of the four channel quarters
  float* rmaxs = scratch + 512;   // [4][128] partial row maxima
  float* ssub = scratch + 1024;   // [128] per-token exponent offset of the key tile
  float* red = scratch + 1152;    // [16] per-warp maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.tokens + kTile - 1) / kTile;
  const uint32_t istride = gridDim.x, items = (uint32_t)p.items;  // (the launcher checks that items fit 32 bits)

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < kRing; ++s) {
      mbar_init(bar_tfull(s), 1);
      mbar_init(bar_tempty(s), 1);
    }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(bar_ufull(s), 1);
      mbar_init(bar_fready(s), kFeatWarps);
      mbar_init(bar_d3full(s), 1);
      mbar_init(bar_d3free(s), kFeatWarps);
    }
    mbar_init(bar_ctxfull, 1);
    mbar_init(bar_ctxready, kFeatWarps);
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

  using C0 = std::integral_constant<int, 0>;
  using C1 = std::integral_constant<int, 1>;
  using C2 = std::integral_constant<int, 2>;

  if (warp == 0) {
    if (lane == 0) {
      // =================== TMA producer ===================
      struct Cursor {
        uint32_t item;
        int ph, i;  // phase: 0 = key-max pass (softmax kernel), 1 = K/V pairs, 2 = Q
      };
      auto cur_init = [&](Cursor& c) { c.item = blockIdx.x; c.ph = KIND == 0 ? 0 : 1; c.i = 0; };
      auto cur_get = [&](const Cursor& c, const CUtensorMap*& tm, int& t) {
        if (c.ph == 1) { tm = (c.i & 1) ? &tm_v : &tm_k; t = c.i >> 1; }
        else { tm = c.ph == 0 ? &tm_k : &tm_q; t = c.i; }
      };
      auto cur_next = [&](Cursor& c) {
        const int n = c.ph == 1 ? 2 * nt : nt;
        if (++c.i < n) return;
        c.i = 0;
        if (++c.ph == 3) { c.ph = KIND == 0 ? 0 : 1; c.item += istride; }
      };
      auto coords = [&](uint32_t it32, int& h, int& g0, int& g1) {
        const uint32_t g = it32 / (uint32_t)p.heads;
        h = (int)(it32 - g * (uint32_t)p.heads);
        g1 = (int)(g / (uint32_t)p.G0);
        g0 = (int)(g - (uint32_t)g1 * (uint32_t)p.G0);
      };
      constexpr int kPrefetch = 8;
      Cursor pf, ld;
      cur_init(pf);
      cur_init(ld);
      auto prefetch_one = [&]() {
        if (pf.item >= items) return;
        const CUtensorMap* tm; int t, h, g0, g1;
        cur_get(pf, tm, t);
        coords(pf.item, h, g0, g1);
        tma_prefetch_4d(tm, h * 64, t * kTile, g0, g1);
        cur_next(pf);
      };
      for (int i = 0; i < kPrefetch; ++i) prefetch_one();
      uint32_t slot = 0, par = 0;
      while (ld.item < items) {
        prefetch_one();
        const CUtensorMap* tm; int t, h, g0, g1;
        cur_get(ld, tm, t);
        coords(ld.item, h, g0, g1);
        mbar_wait(bar_tempty(slot), par ^ 1u);
        mbar_arrive_expect_tx(bar_tfull(slot), kSlabBytes);
        tma_load_4d(tm, bar_tfull(slot), s_ring + slot * kSlabBytes, h * 64, t * kTile, g0, g1);
        if (++slot == kRing) { slot = 0; par ^= 1u; }
        cur_next(ld);
      }
    }
  } else if (warp == 1) {
    // =================== MMA issuer ===================
    // ONE warp issues every MMA, in BURSTS of up to 12: the consumer MMAs of job j (A = the features in U slot
    // j % 2) followed by the projection of job j + 2 into the same slot. tcgen05.mma instructions of one thread
    // execute in issue order, so the projection cannot overwrite the slot before the consumer MMAs issued ahead of
    // it have read their A operand out of it: the slot is handed on without a commit / mbarrier round trip (across
    // issuing threads that order is NOT guaranteed: versions with separate consumer / projection issuer warps were
    // faster but produced NaNs once the bursts got longer; profiles/r02_favor_notes.md). Measured issue cost
    // (tools/micro/umma_bench.cu): ~300 cycles per burst (mbarrier try_wait, elect / reconvergence) + 50-70 cycles
    // per MMA in real code (operand descriptors into uniform registers), against 30-70 cycles of execution per
    // 128 x (64..128) x 16 MMA: long bursts keep the fixed part small.
    // The job nest is unrolled per pass (three jobs); all 32 lanes run the warp-uniform control flow, one elected
    // lane issues.
    const uint64_t d_omega = umma_desc_sw128(s_omega);  // + c * 1024 + 2 k
    const uint64_t d_ring = umma_desc_sw128(s_ring);    // + slot * 1024 + 2 k
    const uint64_t d_ctx = umma_desc_sw128(s_ctx);      // + slab * 640 + 2 k
    uint32_t nD3 = 0, d3u0 = 0, d3u1 = 0;               // output tiles started / fills per D3 slot
    uint32_t nItems = 0;
    auto wait_d3_region = [&](uint32_t s) {
      const uint32_t uses = s ? d3u1 : d3u0;
      if (uses > 0) mbar_wait(bar_d3free(s), (uses - 1u) & 1u);
    };
    // projection of job CU of pass `u` (elected lane): key side U^T[m, tok] = Omega_c . X^T (features on the lanes);
    // query side U[tok, m] = X . Omega_c^T (16 columns for chunk 2)
    auto issue_u = [&](const PassInfo& u, auto cu_c) {
      constexpr int CU = decltype(cu_c)::value;
      const uint32_t us = (u.j0 + CU) & 1u;
      const bool key = u.key_side();
      const uint64_t dx = d_ring + (uint64_t)(u.a_slot * 1024u);
      const uint64_t dw = d_omega + (uint64_t)(CU * 1024);
      const uint64_t da = key ? dw : dx;
      const uint64_t db = key ? dx : dw;
      const uint32_t idesc = ((!key && CU == 2) ? umma_idesc_bf16(128, 16) : umma_idesc_bf16(128, 128)) ^ kFmt;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + kColU + us * 128u, da + 2 * k, db + 2 * k, idesc, k > 0);
      umma_commit(bar_ufull(us));
      // the tile's last projection: hand the ring slot back (a softmax QMAX pass keeps its tile for the Q pass)
      if (CU == 2 && u.kind != kPassX) umma_commit(bar_tempty(u.a_slot));
    };
    // burst C of pass `c`: consumer MMAs of its job C, then the projection of the job two ahead (pass `u`, job CU)
    auto burst = [&](const PassInfo& c, const PassInfo& u, auto c_c, auto cu_c) {
      constexpr int C = decltype(c_c)::value;
      constexpr int CU = decltype(cu_c)::value;
      const uint32_t jc = c.j0 + C;
      const uint32_t fs = jc & 1u;
      const uint32_t par = (jc >> 1) & 1u;
      const uint32_t ds = nD3 & 1u;
      // A operand of k-step k inside the U slot: the feature warp of column quarter k / 2 wrote its 16 packed
      // columns at the start of its own 32-column range
      const uint32_t a0 = tmem + kColU + fs * 128u;
      if (c.kind == kPassK) {
        if (c.t == 0 && C < 2) wait_d3_region((uint32_t)C);  // ctx block c aliases D3[c]
        if (C == 0) mbar_wait(bar_tfull(c.v_slot), c.v_par);
      } else if (c.kind == kPassQ && C == 0) {
        if (c.t == 0) mbar_wait(bar_ctxready, nItems & 1u);
        wait_d3_region(ds);
      }
      if (u.valid && CU == 0 && !(KIND == 0 && u.kind == kPassQ)) mbar_wait(bar_tfull(u.a_slot), u.a_par);
      mbar_wait(bar_fready(fs), par);
      tc_fence_after();
      if (elect_one()) {
        if (c.kind == kPassK) {
          // ctx^T_c[128 m x 80] (+)= k'^T (A, TMEM, K = tokens) . [V | 1] (B, MN-major; second MN chunk = cslab)
          const uint64_t dbv = desc_mn_sw128(s_ring + c.v_slot * kSlabBytes, s_cslab - s_ring - c.v_slot * kSlabBytes);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ts(tmem + kColCtx + 80u * C, a0 + 32u * (k >> 1) + 8u * (k & 1), dbv + 128 * k, idesc_bf16_major(128, 80, 0, 1) ^ kFmt,
                         (c.t > 0 || k > 0));
          if (C == 2) {
            umma_commit(bar_tempty(c.v_slot));
            if (c.t == nt - 1) umma_commit(bar_ctxfull);
          }
        } else if (c.kind == kPassQ) {
          // out|den [128 tok x 80] (+)= q'_c (A, TMEM, K = features) . ctx_c (B, K-major over m)
          const uint64_t db = d_ctx + (uint64_t)(2 * C * 640);
          constexpr int NK = C == 2 ? 1 : 8;
#pragma unroll
          for (int k = 0; k < NK; ++k)
            umma_bf16_ts(tmem + kColCtx + 80u * ds, a0 + 32u * (k >> 1) + 8u * (k & 1), db + (k >> 2) * 640 + 2 * (k & 3),
                         umma_idesc_bf16(128, 80) ^ kFmt, (C > 0 || k > 0));
          if (C == 2) umma_commit(bar_d3full(ds));
        }
        // (stabiliser passes: the feature warps have taken their maxima, nothing to consume)
        if (u.valid) issue_u(u, cu_c);
      }
      __syncwarp();
    };
    PassIter<KIND> it;
    it.init(blockIdx.x, items, istride, nt);
    PassInfo cur = it.next(), nxt = it.next();
    if (cur.valid) {  // prologue: the first two projections
      mbar_wait(bar_tfull(cur.a_slot), cur.a_par);
      tc_fence_after();
      if (elect_one()) {
        issue_u(cur, C0{});
        issue_u(cur, C1{});
      }
      __syncwarp();
    }
    while (cur.valid) {
      burst(cur, cur, C0{}, C2{});
      burst(cur, nxt, C1{}, C0{});
      burst(cur, nxt, C2{}, C1{});
      if (cur.kind == kPassQ) {
        if (cur.t == 0) ++nItems;
        if (nD3 & 1u) ++d3u1; else ++d3u0;
        ++nD3;
      }
      cur = nxt;
      nxt = it.next();
    }
  } else {
    // =================== feature / epilogue warps ===================
    const int fw = warp - 2;         // 0..15
    const int lg = warp & 3;         // TMEM lane group this warp may touch
    const int cq = fw >> 2;          // column quarter of the U slot owned by this warp
    const int row = lg * 32 + lane;  // TMEM lane: token row (query side) / feature row of the chunk (key side)
    const uint32_t t_lane = ((uint32_t)(lg * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kEps = KIND == 0 ? 1e-4f : 1e-3f;
    const uint32_t eps2 = pack_h16x2(kEps, kEps, kDt);
    const uint32_t ubase = tmem + t_lane + kColU + 32u * (uint32_t)cq;  // + us * 128

    uint32_t nJ = 0;  // jobs processed (job parity = U slot)
    uint32_t nItems = 0, nD3 = 0, tile_seq = 0;
    float gmax = 0.f, sub = 0.f;
    const uint32_t last_item = blockIdx.x + ((items - 1u - blockIdx.x) / istride) * istride;

    uint32_t raw[32];
    // wait for the accumulator of job nJ and issue its TMEM loads. Q2: the job is a query-side chunk 2
    // (16 feature columns, read by the cq == 0 warps only)
    auto prefetch = [&](bool q2) {
      const uint32_t us = nJ & 1u;
      mbar_wait(bar_ufull(us), (nJ >> 1) & 1u);
      tc_fence_after();
      if (!q2) {
        tmem_ld_32x32p(ubase + us * 128u, raw);
      } else if (cq == 0) {
        tmem_ld_32x16p(ubase + us * 128u, raw);
      }
    };
    // publish: features of job nJ are in TMEM (or its maxima taken) -> the consumer issuer may proceed
    auto publish = [&](bool stored) {
      if (stored) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fready(nJ & 1u));
      ++nJ;
    };
    // two accumulator values -> one packed bf16x2 feature pair; s0 / s1: exponent offsets (softmax kernel)
    auto feat2 = [&](uint32_t r0, uint32_t r1, float s0, float s1) -> uint32_t {
      const float x0 = __uint_as_float(r0), x1 = __uint_as_float(r1);
      if (KIND == 0)
        return add_h16x2<F16>(pack_h16x2(ex2_approx(fmaf(x0, kLog2e, -s0)), ex2_approx(fmaf(x1, kLog2e, -s1)), kDt), eps2);
      return add_h16x2<F16>(cvt_relu_h16x2<F16>(x0, x1), eps2);
    };

    // ---- key-side job: this thread holds feature row (128 C + row) x tokens [32 cq, 32 cq + 32) of the tile ----
    // ntok: valid tokens of the tile; next_q2: shape of the following job
    auto key_feat_job = [&](int C, int ntok, bool has_next, bool next_q2) {
      const uint32_t us = nJ & 1u;
      tmem_ld_wait();
      const bool row_ok = 128 * C + row < p.m;
      const int lim = ntok - 32 * cq;  // token columns >= lim are padding (only in a ragged last tile)
      uint32_t pk[16];
      if (KIND == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(ssub + 32 * cq + 4 * q);
          pk[2 * q] = feat2(raw[4 * q], raw[4 * q + 1], s4.x, s4.y);
          pk[2 * q + 1] = feat2(raw[4 * q + 2], raw[4 * q + 3], s4.z, s4.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = feat2(raw[2 * i], raw[2 * i + 1], 0.f, 0.f);
      }
      if (!row_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      } else if (lim < 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (2 * i >= lim) pk[i] = 0u;
          else if (2 * i + 1 >= lim) pk[i] &= 0x0000ffffu;
        }
      }
      tmem_st_32x16(ubase + us * 128u, pk);
      publish(true);
      if (has_next) prefetch(next_q2);
    };
    // ---- key-side stabiliser job (softmax kernel): max of the raw projections over valid rows / tokens ----
    auto key_max_job = [&](int C, int ntok, float& acc) {
      tmem_ld_wait();
      const bool row_ok = 128 * C + row < p.m;
      const int lim = ntok - 32 * cq;
      float mx = -INFINITY;
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < lim) mx = fmaxf(mx, __uint_as_float(raw[i]));
      }
      acc = fmaxf(acc, mx);
      publish(false);
      prefetch(false);  // a key-side job always follows
    };
    // ---- query-side job: this thread holds token row `row` x features [128 C + 32 cq, + 32) ----
    auto query_feat_job = [&](auto cc, bool has_next, bool next_q2) {
      constexpr int C = decltype(cc)::value;
      const uint32_t us = nJ & 1u;
      if (C < 2) {
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = feat2(raw[2 * i], raw[2 * i + 1], sub, sub);
        tmem_st_32x16(ubase + us * 128u, pk);
        publish(true);
      } else if (cq == 0) {
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = feat2(raw[2 * i], raw[2 * i + 1], sub, sub);
        tmem_st_32x8(ubase + us * 128u, pk);
        publish(true);
      } else {
        publish(false);
      }
      if (has_next) prefetch(next_q2);
    };
    // ---- query-side stabiliser job (softmax kernel): per-row max over the valid feature columns ----
    auto query_max_job = [&](auto cc, float& acc) {
      constexpr int C = decltype(cc)::value;
      constexpr int NC = C < 2 ? 32 : 16;
      if (C < 2 || cq == 0) {
        tmem_ld_wait();
        const int lim = p.m - (128 * C + 32 * cq);
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i < lim) mx = fmaxf(mx, __uint_as_float(raw[i]));
        acc = fmaxf(acc, mx);
      }
      publish(false);
      prefetch(C == 1);  // QMAX chunk 2 follows chunk 1; the query chunk 0 follows QMAX chunk 2
    };

    // 0.5 * dn^2 * |x|^2 of token `row` of ring tile a_seq (softmax kernel): the four warps of a lane group sum
    // 16 channels each, the partials meet in shared memory
    auto row_diag = [&](uint32_t a_seq) {
      mbar_wait(bar_tfull(a_seq % kRing), (a_seq / kRing) & 1u);  // TMA bytes visible to this thread
      const uint32_t tile = s_ring + (a_seq % kRing) * kSlabBytes;
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
      named_bar_sync(1, kFeatThreads);
      return (part[row] + part[128 + row] + part[256 + row] + part[384 + row]) * (0.5f * 0.125f);
    };

    // out/den epilogue of one query tile: warp (lg, cq) stores channels [16 cq, 16 cq + 16) of its 32 tokens
    // (the item's output row pointer is resolved once per item, in 32-bit arithmetic: four 64-bit divisions per tile
    // in front of the stores sat on the feature warps' critical path)
    const uint16_t* out_item = nullptr;
    auto item_out_ptr = [&](uint32_t it32) {
      const uint32_t g = it32 / (uint32_t)p.heads, h = it32 - g * (uint32_t)p.heads;
      const uint32_t g1 = g / (uint32_t)p.G0, g0 = g - g1 * (uint32_t)p.G0;
      return reinterpret_cast<const uint16_t*>(p.out) + (int64_t)g1 * p.ogs1 + (int64_t)g0 * p.ogs0 + (int64_t)row * p.ots + h * 64 + 16 * cq;
    };
    auto epilogue = [&](int t) {
      const uint32_t ds = nD3 & 1u;
      mbar_wait(bar_d3full(ds), (nD3 >> 1) & 1u);
      tc_fence_after();
      uint32_t rd[1], r0[16];
      tmem_ld_32x1p(tmem + t_lane + kColCtx + 80u * ds + 64, rd);  // column 64 = normaliser
      tmem_ld_32x16p(tmem + t_lane + kColCtx + 80u * ds + 16u * (uint32_t)cq, r0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_d3free(ds));
      ++nD3;
      if (t * kTile + row < p.tokens) {
        const float inv = __fdividef(1.f, __uint_as_float(rd[0]));
        uint4* op = reinterpret_cast<uint4*>(const_cast<uint16_t*>(out_item) + (int64_t)(t * kTile) * p.ots);
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

    if (blockIdx.x < items) prefetch(false);
    for (uint32_t item = blockIdx.x; item < items; item += istride) {
      out_item = item_out_ptr(item);
      if (KIND == 0) {
        // ---- key stabiliser: global max of Omega'.K^T over the valid features / tokens ----
        float kmx = -INFINITY;
        for (int t = 0; t < nt; ++t) {
          ++tile_seq;
          const int ntok = min(kTile, p.tokens - t * kTile);
          key_max_job(0, ntok, kmx);
          key_max_job(1, ntok, kmx);
          key_max_job(2, ntok, kmx);
        }
        kmx = warp_max(kmx);
        if (lane == 0) red[fw] = kmx;
        named_bar_sync(1, kFeatThreads);
        gmax = red[0];
#pragma unroll
        for (int i = 1; i < kFeatWarps; ++i) gmax = fmaxf(gmax, red[i]);
      }
      // ---- keys: k'^T chunks feed the context MMAs ----
      for (int t = 0; t < nt; ++t) {
        if (KIND == 0) {
          const float dg = row_diag(tile_seq);
          if (cq == 0) ssub[row] = (dg + gmax) * kLog2e;
          named_bar_sync(1, kFeatThreads);
        }
        tile_seq += 2;
        const int ntok = min(kTile, p.tokens - t * kTile);
        key_feat_job(0, ntok, true, false);
        key_feat_job(1, ntok, true, false);
        key_feat_job(2, ntok, true, false);
      }
      // ---- context read-out: TMEM ctx^T blocks -> bf16 K-major smem. Warp (lg, cq) converts columns
      //      [16 cq, 16 cq + 16) of blocks 0 and 1 for its 32 feature rows; the normaliser column 64 and the
      //      16-row block 2 are spread over the column quarters ----
      {
        mbar_wait(bar_ctxfull, nItems & 1u);
        ++nItems;
        tc_fence_after();
        tmem_ld_wait();  // the prefetch of the first query job is in flight: one wait covers all loads
        auto put = [&](int m, uint32_t n, float v) {
          const uint32_t mc = (uint32_t)m & 63u;
          st_shared_h16<F16>(s_ctx + (uint32_t)(m >> 6) * kCtxSlabBytes + (n >> 3) * 1024u + (n & 7u) * 128u +
                            ((((mc >> 3) ^ n) & 7u) << 4) + (mc & 7u) * 2u,
                        v);
        };
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          uint32_t r[16];
          tmem_ld_32x16p(tmem + t_lane + kColCtx + 80u * b + 16u * (uint32_t)cq, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) put(128 * b + row, 16u * cq + i, __uint_as_float(r[i]));
        }
        if (cq < 2) {  // normaliser column of block cq
          uint32_t r[1];
          tmem_ld_32x1p(tmem + t_lane + kColCtx + 80u * (uint32_t)cq + 64u, r);
          tmem_ld_wait();
          put(128 * cq + row, 64u, __uint_as_float(r[0]));
        }
        if (lg == 0) {  // block 2: features 256..271 live in lanes 0..15
          uint32_t r[16], r2[1];
          tmem_ld_32x16p(tmem + kColCtx + 160u + 16u * (uint32_t)cq, r);
          tmem_ld_32x1p(tmem + kColCtx + 160u + 64u, r2);
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
      // ---- queries: q' chunks feed the output MMAs; the out/den epilogue of tile t runs behind
      //      the first job of tile t+1 ----
      bool pending = false;
      for (int t = 0; t < nt; ++t) {
        const bool more = t + 1 < nt || item != last_item;  // another job follows this tile
        if (KIND == 0) {
          const float diag = row_diag(tile_seq);
          float rmx = -INFINITY;
          query_max_job(C0{}, rmx);
          if (pending) { epilogue(t - 1); pending = false; }
          query_max_job(C1{}, rmx);
          query_max_job(C2{}, rmx);
          rmaxs[cq * 128 + row] = rmx;
          named_bar_sync(1, kFeatThreads);
          sub = (diag + fmaxf(fmaxf(rmaxs[row], rmaxs[128 + row]), fmaxf(rmaxs[256 + row], rmaxs[384 + row]))) * kLog2e;
        }
        ++tile_seq;
        query_feat_job(C0{}, true, false);
        if (pending) { epilogue(t - 1); pending = false; }
        query_feat_job(C1{}, true, true);
        query_feat_job(C2{}, more, false);
        pending = true;
      }
      epilogue(nt - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int KIND, bool F16>
int launch_kind(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FavorTmParams& p,
                cudaStream_t stream) {
  static PerDeviceOnce once;
  const int cfg_rc = per_device_once(once, []() {
    return cuda_status(cudaFuncSetAttribute(favor_tm_kernel<KIND, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  int grid = num_sms();
  if (p.items < grid) grid = (int)p.items;
  favor_tm_kernel<KIND, F16><<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, p);
  return post_launch();
}

}  // namespace

int favor_tm_launch(const rfk_favor_desc* d, cudaStream_t stream) {
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
  FavorTmParams p{};
  p.proj = d->proj; p.out = d->out; p.m = d->m_features; p.heads = d->heads;
  p.tokens = (int)d->tokens; p.G0 = d->G[0]; p.G1 = d->G[1];
  p.items = d->G[0] * d->G[1] * d->heads;
  if (p.items > 0x7fffffffLL || d->G[0] > 0x7fffffffLL) return RFK_ERR_UNSUPPORTED;  // 32-bit item decode in the kernel
  p.ogs0 = d->out_gs[0]; p.ogs1 = d->out_gs[1]; p.ots = d->out_ts;
  if (d->io_dtype == RFK_F16)
    return d->kind == 0 ? launch_kind<0, true>(tq, tk, tv, p, stream) : launch_kind<1, true>(tq, tk, tv, p, stream);
  return d->kind == 0 ? launch_kind<0, false>(tq, tk, tv, p, stream) : launch_kind<1, false>(tq, tk, tv, p, stream);
}

}  // namespace rfk
