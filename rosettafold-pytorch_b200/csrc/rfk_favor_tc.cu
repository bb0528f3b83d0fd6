// rfk_favor_tc.cu — fused Performer FAVOR+ attention on tcgen05 / TMEM / TMA (bf16 operands,
// fp32 accumulation), software-pipelined. One persistent CTA per SM walks a stream of
// chunk JOBS; nothing of size tokens x m ever leaves the SM.
//
// A job is (item = (group, head), type, 128-token tile t, feature chunk c); the 272 (padded)
// features are split into chunks of 128 | 128 | 16 columns. Job types:
//   KMAX  U = K.Omega_c'^T                       -> running max            (softmax kernel only)
//   K     U = K.Omega_c'^T -> k'_c (CUDA cores)  -> ctx_c[m, d|1] += k'_c^T . [V | 1]
//   QMAX  U = Q.Omega_c'^T                       -> per-row max            (softmax kernel only)
//   Q     U = Q.Omega_c'^T -> q'_c               -> out|den += q'_c . ctx_c
// Per item: relu kernel  K(t,c)*, read-out of ctx, Q(t,c)*;
//           softmax      KMAX(t,c)*, K(t,c)*, read-out, then per tile QMAX(t,c)*, Q(t,c)*.
//
// Four roles run decoupled and only meet at mbarriers:
//   warp 0 lane 0  TMA producer: K/V/Q tiles into a 3-slot ring; a second cursor runs 6 tiles ahead
//                  and pulls them into L2 (cp.async.bulk.prefetch.tensor), since two ring slots are
//                  held by the K and V tile in use and one slot cannot cover the HBM latency
//   warp 1         U issuer:  U = X.Omega_c'^T of every job, as far ahead as the two U slots allow
//   warp 2         consumer issuer: context / output MMAs of the K and Q jobs
//                  (both issuer warps run warp-uniform control flow and issue under elect.sync: a
//                  lane-0-only branch makes the compiler wrap every UTCHMMA in a divergence waterfall,
//                  and a single warp retires one dependent instruction per ~6-8 cycles, which is why
//                  the per-job bookkeeping is split over two warps and uses precomputed descriptors)
//   warps 3..14    12 feature/epilogue warps, all of them on every job: three warps share a TMEM lane
//                  group and take the column ranges [0,48) | [40,88) | [80,128) of the chunk (uniform
//                  x32+x16 TMEM loads; the 8-column overlaps are written twice with equal values).
//                  Software-pipelined: while the packed features of job j are stored to shared memory,
//                  fenced (fence.proxy.async) and published, the accumulator of job j+1 is already on
//                  its way TMEM -> registers; the U slot is handed back as soon as the load lands.
//                  15 warps leave 128 registers per thread (no spills, no per-job address recomputation).
//                  ctx read-out; out/den epilogue deferred behind the next tile's first job.
// Measured bounds (ncu + in-kernel timeline, profiles/): per 128x128 job ~130 KB of shared-memory
// traffic (MMA operands from smem + feature stores + TMA) and the F2FP/HADD2 conversion pipes, not
// the tensor pipe; the next step is to keep the features in TMEM (tcgen05.st, A-from-TMEM MMAs).
//
// TMEM (512 columns): ctx^T blocks b=0..2 (lanes = m - 128 b, 80 columns = d | 1) at [0,240);
//   U ring 2 x 128 columns at [256,512); out|den accumulators D3[s] alias ctx blocks 0/1 (dead in
//   the query phase; the issuer orders the reuse through the d3free barriers).
// Shared memory (1024-byte aligned tiles, 128-byte swizzle):
//   omega [272][64] bf16 K-major | tile ring 3 x [128 tok][64 d] | cslab [128 tok][64], col 0 = 1
//   (second MN chunk of the "[V | 1]" operand) | feat ring 2 x 2 slabs [128 tok][64 m] (MN-major A of
//   the context MMA and K-major A of the output MMA: same bytes) | ctx 5 x [80][64 m] K-major B.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "rfk_common.cuh"

namespace rfk {

namespace {

constexpr int kFeatWarps = 12;
constexpr int kThreads = 32 * (3 + kFeatWarps);  // TMA producer, two MMA issuers, feature warps
constexpr int kMP = 272;    // padded feature count
constexpr int kTile = 128;  // tokens per tile
constexpr int kRing = 3;
constexpr uint32_t kSlabBytes = kTile * 128;   // 16384
constexpr uint32_t kOmegaBytes = kMP * 128;    // 34816
constexpr uint32_t kCtxSlabBytes = 80 * 128;   // 10240
constexpr uint32_t kOffOmega = 0;
constexpr uint32_t kOffRing = kOffOmega + kOmegaBytes;
constexpr uint32_t kOffCslab = kOffRing + kRing * kSlabBytes;  // after the ring: LBO of [V | 1] > 0
constexpr uint32_t kOffFeat = kOffCslab + kSlabBytes;
constexpr uint32_t kOffCtx = kOffFeat + 4 * kSlabBytes;
constexpr uint32_t kOffBar = kOffCtx + 5 * kCtxSlabBytes;
constexpr uint32_t kOffScratch = kOffBar + 256;
constexpr uint32_t kScratchFloats = 4 * 128 + 4 * 128 + 16;  // diag partials, row maxima, block max
constexpr uint32_t kSmemBytes = kOffScratch + kScratchFloats * 4 + 1024;
static_assert(kOffRing % 1024 == 0 && kOffCslab % 1024 == 0 && kOffFeat % 1024 == 0 && kOffCtx % 1024 == 0, "align");
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr uint32_t kColCtx = 0, kColU = 256;

struct FavorTcParams {
  const float* proj;
  void* out;
  int m, heads;
  int tokens;
  int64_t G0, G1, items;
  int64_t ogs0, ogs1, ots;
  int dbg;  // timing experiments only (RFK_FAVOR_DBG): skip parts of the pipeline
  long long* trace;  // developer timeline (RFK_FAVOR_TRACE): [4 roles][kTraceMax][2] (event, clock) of CTA 0
};
constexpr int kTraceMax = 1024;
#ifdef RFK_FAVOR_TRACE_BUILD  // developer builds only: the hooks cost ~10% of the kernel time
constexpr bool kTraceBuild = true;
#else
constexpr bool kTraceBuild = false;
#endif

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
__device__ __forceinline__ void st_shared_b16(uint32_t addr, float f) {
  const unsigned short h = __bfloat16_as_ushort(__float2bfloat16_rn(f));
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// one lane of the (fully active) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// pointer flavours of the TMEM loads (sub-ranges of one register array)
__device__ __forceinline__ void tmem_ld_32x32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
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
// {bf16(max(lo,0)), bf16(max(hi,0))} in one instruction
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

template <int KIND>  // 0: softmax kernel (exp features, stabilisers), 1: generalized ReLU kernel
__global__ void __launch_bounds__(kThreads, 1)
favor_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const FavorTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_omega = base + kOffOmega, s_ring = base + kOffRing, s_cslab = base + kOffCslab,
                 s_feat = base + kOffFeat, s_ctx = base + kOffCtx;
  const uint32_t bars = base + kOffBar;
  auto bar_tfull = [&](uint32_t s) { return bars + 8u * s; };
  auto bar_tempty = [&](uint32_t s) { return bars + 24u + 8u * s; };
  auto bar_ufull = [&](uint32_t s) { return bars + 48u + 8u * s; };
  auto bar_ufree = [&](uint32_t s) { return bars + 64u + 8u * s; };
  auto bar_fready = [&](uint32_t s) { return bars + 80u + 8u * s; };
  auto bar_ffree = [&](uint32_t s) { return bars + 96u + 8u * s; };
  auto bar_d3full = [&](uint32_t s) { return bars + 112u + 8u * s; };
  auto bar_d3free = [&](uint32_t s) { return bars + 128u + 8u * s; };
  const uint32_t bar_ctxfull = bars + 144u, bar_ctxready = bars + 152u;
  const uint32_t tmem_slot = bars + 160u;
  float* scratch = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + kOffScratch);
  float* part = scratch;             // [4][128] partial |x|^2 of the four column quarters
  float* rmaxs = scratch + 512;      // [4][128] partial row maxima
  float* red = scratch + 1024;       // [16] per-warp maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (p.tokens + kTile - 1) / kTile;
  const int64_t istride = gridDim.x;
  // developer timeline: role 0 = U issuer, 1 = consumer issuer, 2 / 3 = first warp of feature group 0 / 1
  auto dbg_bits = [&]() { return kTraceBuild ? p.dbg : 0; };  // experiments exist in developer builds only
  int tr_n = 0;
  const int tr_role = warp == 1 ? 0 : warp == 2 ? 1 : warp == 3 ? 2 : warp == 7 ? 3 : -1;  // warps 3 / 7: lane group 3 of group 0 / 1
  auto TR = [&](int ev) {
    if (kTraceBuild && p.trace && blockIdx.x == 0 && lane == 0 && tr_role >= 0 && tr_n < kTraceMax) {
      p.trace[((int64_t)tr_role * kTraceMax + tr_n) * 2] = ev;
      p.trace[((int64_t)tr_role * kTraceMax + tr_n) * 2 + 1] = clock64();
      ++tr_n;
    }
  };

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    for (uint32_t s = 0; s < kRing; ++s) {
      mbar_init(bar_tfull(s), 1);
      mbar_init(bar_tempty(s), 1);
    }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(bar_ufull(s), 1);
      mbar_init(bar_ufree(s), kFeatWarps);
      mbar_init(bar_fready(s), kFeatWarps);
      mbar_init(bar_ffree(s), 1);
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
    // ctx rows 65..79 and feature slabs (stale columns feed never-read accumulator rows): zero once
    for (int i = threadIdx.x; i < (int)(5 * kCtxSlabBytes / 16); i += kThreads)
      st_shared_v4(s_ctx + i * 16, make_uint4(0, 0, 0, 0));
    for (int i = threadIdx.x; i < (int)(4 * kSlabBytes / 16); i += kThreads)
      st_shared_v4(s_feat + i * 16, make_uint4(0, 0, 0, 0));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // =================== TMA producer ===================
      // The 3-slot ring only covers ~1 tile of look-ahead (two slots are held by the K and V tile in
      // use), far less than the HBM latency; a second cursor therefore runs kPrefetch tiles ahead
      // and pulls them into L2 (cp.async.bulk.prefetch.tensor), so the ring loads hit L2.
      struct Cursor {
        int64_t item;
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
      auto coords = [&](int64_t item, int& h, int& g0, int& g1) {
        h = (int)(item % p.heads);
        const int64_t g = item / p.heads;
        g0 = (int)(g % p.G0);
        g1 = (int)(g / p.G0);
      };
      constexpr int kPrefetch = 6;
      Cursor pf, ld;
      cur_init(pf);
      cur_init(ld);
      auto prefetch_one = [&]() {
        if (pf.item >= p.items) return;
        const CUtensorMap* tm; int t, h, g0, g1;
        cur_get(pf, tm, t);
        coords(pf.item, h, g0, g1);
        tma_prefetch_4d(tm, h * 64, t * kTile, g0, g1);
        cur_next(pf);
      };
      for (int i = 0; i < kPrefetch; ++i) prefetch_one();
      uint32_t slot = 0, par = 0;
      while (ld.item < p.items) {
        if (!(dbg_bits() & 16)) prefetch_one();
        const CUtensorMap* tm; int t, h, g0, g1;
        cur_get(ld, tm, t);
        coords(ld.item, h, g0, g1);
        mbar_wait(bar_tempty(slot), par ^ 1u);
        if (dbg_bits() & 16) {
          mbar_arrive(bar_tfull(slot));
        } else {
          mbar_arrive_expect_tx(bar_tfull(slot), kSlabBytes);
          tma_load_4d(tm, bar_tfull(slot), s_ring + slot * kSlabBytes, h * 64, t * kTile, g0, g1);
        }
        if (++slot == kRing) { slot = 0; par ^= 1u; }
        cur_next(ld);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // =================== MMA issuers ===================
    // Everything an issuer does per job is on the critical path of the whole CTA, and a single warp
    // retires a dependent instruction only every ~6-8 cycles. Two warps therefore share the work:
    //   warp 1: U = X.Omega_c'^T for every job, as far ahead as the two U slots allow
    //   warp 2: the consumer MMAs (context / output accumulation) of the K and Q jobs
    // All 32 lanes run the warp-uniform control flow and one elected lane issues the tcgen05
    // instructions (a lane-0-only branch makes the compiler wrap every UTCHMMA in a divergence
    // waterfall). Both walk the same nest (item, pass = (kind, tile), chunk).
    constexpr int kPassM = 0, kPassK = 1, kPassX = 2, kPassQ = 3;  // key max, keys, query max, queries
    const int P = KIND == 0 ? 4 * nt : 2 * nt;
    auto decode = [&](int ps, int& t) -> int {
      if (KIND == 1) {
        if (ps < nt) { t = ps; return kPassK; }
        t = ps - nt;
        return kPassQ;
      }
      if (ps < nt) { t = ps; return kPassM; }
      if (ps < 2 * nt) { t = ps - nt; return kPassK; }
      const int r = ps - 2 * nt;
      t = r >> 1;
      return (r & 1) ? kPassQ : kPassX;
    };
    struct Tile { uint32_t slot, par; };
    uint32_t r_slot = 0, r_par = 0;  // ring position of the next tile to allocate
    auto alloc = [&]() {
      Tile x{r_slot, r_par};
      if (++r_slot == kRing) { r_slot = 0; r_par ^= 1u; }
      return x;
    };
    auto commit_dbg = [&](uint32_t bar) { umma_commit(bar); };
    using C0 = std::integral_constant<int, 0>;
    using C1 = std::integral_constant<int, 1>;
    using C2 = std::integral_constant<int, 2>;
    if (warp == 1) {
      const uint64_t d_omega = umma_desc_sw128(s_omega);  // + c * 1024 + 2 k
      const uint64_t d_ring = umma_desc_sw128(s_ring);    // + slot * 1024 + 2 k
      uint32_t nJ = 0;  // U chunks issued (job parity = U slot = feature buffer = warp group)
      auto issue_u = [&](const Tile& a, auto cc, bool release) {
        constexpr int C = decltype(cc)::value;
        const uint32_t us = nJ & 1u;
        mbar_wait(bar_ufree(us), ((nJ >> 1) & 1u) ^ 1u);
        TR(10 + C);
        if (C == 0) mbar_wait(bar_tfull(a.slot), a.par);
        TR(13);
        tc_fence_after();
        const uint64_t da = d_ring + (uint64_t)(a.slot * 1024u);
        const uint64_t db = d_omega + (uint64_t)(C * 1024);
        constexpr uint32_t idesc = C == 2 ? umma_idesc_bf16(128, 16) : umma_idesc_bf16(128, 128);
        if (elect_one()) {
          if (!(dbg_bits() & 4)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(tmem + kColU + us * 128u, da + 2 * k, db + 2 * k, idesc, k > 0);
          }
          commit_dbg(bar_ufull(us));
          if (C == 2 && release) commit_dbg(bar_tempty(a.slot));
        }
        __syncwarp();
        TR(14);
        ++nJ;
      };
      Tile a{0, 0};
      for (int64_t item = blockIdx.x; item < p.items; item += istride) {
        for (int ps = 0; ps < P; ++ps) {
          int t;
          const int kind = decode(ps, t);
          if (!(KIND == 0 && kind == kPassQ)) a = alloc();  // a softmax Q pass reuses its QMAX tile
          if (kind == kPassK) (void)alloc();                 // the V tile
          issue_u(a, C0{}, false);
          issue_u(a, C1{}, false);
          issue_u(a, C2{}, kind != kPassX);
        }
      }
    } else {
      const uint64_t d_feat_k = umma_desc_sw128(s_feat);              // K-major A of the output MMA
      const uint64_t d_ctx = umma_desc_sw128(s_ctx);                  // + slab * 640 + 2 k
      const uint64_t d_feat_mn = desc_mn_sw128(s_feat, kSlabBytes);   // MN-major A of the context MMA
      const uint64_t d_v0 = desc_mn_sw128(s_ring, s_cslab - s_ring);  // [V | 1] in ring slot 0 / 1 / 2
      const uint64_t d_v1 = desc_mn_sw128(s_ring + kSlabBytes, s_cslab - s_ring - kSlabBytes);
      const uint64_t d_v2 = desc_mn_sw128(s_ring + 2 * kSlabBytes, s_cslab - s_ring - 2 * kSlabBytes);
      uint32_t nC = 0;               // jobs consumed
      uint32_t nF0 = 0, nF1 = 0;     // feature jobs consumed per feature buffer
      uint32_t nD3 = 0, d3u0 = 0, d3u1 = 0;  // output tiles started / fills per D3 slot
      uint32_t nItems = 0;
      auto wait_d3_region = [&](uint32_t s) {
        const uint32_t uses = s ? d3u1 : d3u0;
        if (uses > 0) mbar_wait(bar_d3free(s), (uses - 1u) & 1u);
      };
      // ctx_c[128 m x 80] (+)= k'_c^T (A, MN-major, K = tokens) . [V | 1] (B, MN-major)
      auto consume_k = [&](const Tile& v, int t, auto cc) {
        constexpr int C = decltype(cc)::value;
        const uint32_t fs = nC & 1u;
        ++nC;
        const uint32_t nF = fs ? nF1++ : nF0++;
        if (t == 0 && C < 2) wait_d3_region((uint32_t)C);  // ctx block c aliases D3[c]
        if (C == 0) mbar_wait(bar_tfull(v.slot), v.par);
        TR(20 + C);
        mbar_wait(bar_fready(fs), nF & 1u);
        TR(23);
        tc_fence_after();
        const uint64_t da = d_feat_mn + (uint64_t)(fs * 2048u);
        const uint64_t db = v.slot == 0 ? d_v0 : (v.slot == 1 ? d_v1 : d_v2);
        if (elect_one()) {
          if (!(dbg_bits() & 1)) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem + kColCtx + 80u * C, da + 128 * k, db + 128 * k, idesc_bf16_major(128, 80, 1, 1), (t > 0 || k > 0));
          }
          commit_dbg(bar_ffree(fs));
          if (C == 2) {
            commit_dbg(bar_tempty(v.slot));
            if (t == nt - 1) commit_dbg(bar_ctxfull);
          }
        }
        __syncwarp();
        TR(24);
      };
      // out|den [128 tok x 80] (+)= q'_c (A, K-major) . ctx_c (B, K-major over m)
      auto consume_q = [&](int t, auto cc) {
        constexpr int C = decltype(cc)::value;
        const uint32_t fs = nC & 1u;
        ++nC;
        const uint32_t nF = fs ? nF1++ : nF0++;
        const uint32_t ds = nD3 & 1u;
        if (C == 0) {
          if (t == 0) {
            mbar_wait(bar_ctxready, nItems & 1u);
            ++nItems;
          }
          wait_d3_region(ds);
        }
        TR(30 + C);
        mbar_wait(bar_fready(fs), nF & 1u);
        TR(33);
        tc_fence_after();
        const uint64_t da = d_feat_k + (uint64_t)(fs * 2048u);
        const uint64_t db = d_ctx + (uint64_t)(2 * C * 640);
        if (elect_one()) {
          if (!(dbg_bits() & 2)) {
            constexpr int NK = C == 2 ? 1 : 8;
#pragma unroll
            for (int k = 0; k < NK; ++k)
              umma_bf16(tmem + kColCtx + 80u * ds, da + (k >> 2) * 1024 + 2 * (k & 3), db + (k >> 2) * 640 + 2 * (k & 3),
                        umma_idesc_bf16(128, 80), (C > 0 || k > 0));
          }
          commit_dbg(bar_ffree(fs));
          if (C == 2) commit_dbg(bar_d3full(ds));
        }
        __syncwarp();
        TR(34);
        if (C == 2) {
          if (ds) ++d3u1; else ++d3u0;
          ++nD3;
        }
      };
      Tile v{0, 0};
      for (int64_t item = blockIdx.x; item < p.items; item += istride) {
        for (int ps = 0; ps < P; ++ps) {
          int t;
          const int kind = decode(ps, t);
          if (!(KIND == 0 && kind == kPassQ)) (void)alloc();
          if (kind == kPassK) {
            v = alloc();
            consume_k(v, t, C0{});
            consume_k(v, t, C1{});
            consume_k(v, t, C2{});
          } else if (kind == kPassQ) {
            consume_q(t, C0{});
            consume_q(t, C1{});
            consume_q(t, C2{});
          } else {
            nC += 3;
          }
        }
      }
    }
  } else {
    // =================== feature / epilogue warps ===================
    // 12 warps, all of them on every job: three warps share a TMEM lane group and split the 128
    // columns of a chunk 48 | 40 | 40 (a thread owns one token row). 15 warps per CTA leave 128
    // registers per thread, so every address below stays in a register instead of being
    // recomputed per job, and the U slot is handed back as soon as the row is in registers.
    const int fw = warp - 3;           // 0..11
    const int lg = warp & 3;           // TMEM lane group this warp may touch
    const int third = fw >> 2;         // which column range of the chunk
    const int row = lg * 32 + lane;    // token row of the tile / TMEM lane
    const uint32_t t_lane = ((uint32_t)(lg * 32) << 16);
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kEps = KIND == 0 ? 1e-4f : 1e-3f;
    const uint32_t eps2 = pack_bf16x2(kEps, kEps);
    // column ranges [0,48) | [40,88) | [80,128): uniform x32 + x16 loads for every warp; the 8-column
    // overlaps are converted and stored twice with identical values
    const int col0 = third * 40;
    constexpr int ncol = 48;
    const uint32_t ubase = tmem + t_lane + kColU + (uint32_t)col0;
    // this thread's feature row: 16-byte chunk ch (0..15 over the two slabs of a buffer)
    const uint32_t frow = s_feat + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    const uint32_t r7 = (uint32_t)(row & 7);
    auto fchunk = [&](int ch) { return (uint32_t)(ch >> 3) * kSlabBytes + ((((uint32_t)(ch & 7)) ^ r7) << 4); };

    uint32_t nJ = 0;               // jobs processed (job parity = U slot = feature buffer)
    uint32_t nF0 = 0, nF1 = 0;     // feature jobs processed per feature buffer
    uint32_t nItems = 0, nD3 = 0, tile_seq = 0;
    float gmax = 0.f, sub = 0.f, diag = 0.f;
    const int64_t last_item = blockIdx.x + ((p.items - 1 - blockIdx.x) / istride) * istride;

    // out/den epilogue: the three warps of a lane group store channels [0,24) | [24,48) | [48,64)
    auto epilogue = [&](int64_t item, int t) {
      const uint32_t ds = nD3 & 1u;
      TR(50);
      mbar_wait(bar_d3full(ds), (nD3 >> 1) & 1u);
      TR(51);
      tc_fence_after();
      const int ch0 = third * 24;
      uint32_t rd[16], r0[16], r1[16];
      tmem_ld_32x16(tmem + t_lane + kColCtx + 80u * ds + 64, rd);  // column 64 = normaliser
      tmem_ld_32x16(tmem + t_lane + kColCtx + 80u * ds + ch0, r0);
      tmem_ld_32x16(tmem + t_lane + kColCtx + 80u * ds + ch0 + 8, r1);  // columns ch0+8 .. ch0+23
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_d3free(ds));
      ++nD3;
      if (t * kTile + row < p.tokens && !(dbg_bits() & 32)) {
        const int h = (int)(item % p.heads);
        const int64_t g = item / p.heads;
        const int64_t g0 = g % p.G0, g1 = g / p.G0;
        const float inv = 1.f / __uint_as_float(rd[0]);
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + g1 * p.ogs1 + g0 * p.ogs0 +
                                             (int64_t)(t * kTile + row) * p.ots + h * 64 + ch0);
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(r0[0]) * inv, __uint_as_float(r0[1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(r0[2]) * inv, __uint_as_float(r0[3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(r0[4]) * inv, __uint_as_float(r0[5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(r0[6]) * inv, __uint_as_float(r0[7]) * inv);
        op[0] = w;
        w.x = pack_bf16x2(__uint_as_float(r1[0]) * inv, __uint_as_float(r1[1]) * inv);
        w.y = pack_bf16x2(__uint_as_float(r1[2]) * inv, __uint_as_float(r1[3]) * inv);
        w.z = pack_bf16x2(__uint_as_float(r1[4]) * inv, __uint_as_float(r1[5]) * inv);
        w.w = pack_bf16x2(__uint_as_float(r1[6]) * inv, __uint_as_float(r1[7]) * inv);
        op[1] = w;
        if (third < 2) {
          w.x = pack_bf16x2(__uint_as_float(r1[8]) * inv, __uint_as_float(r1[9]) * inv);
          w.y = pack_bf16x2(__uint_as_float(r1[10]) * inv, __uint_as_float(r1[11]) * inv);
          w.z = pack_bf16x2(__uint_as_float(r1[12]) * inv, __uint_as_float(r1[13]) * inv);
          w.w = pack_bf16x2(__uint_as_float(r1[14]) * inv, __uint_as_float(r1[15]) * inv);
          op[2] = w;
        }
      }
    };

    // 0.5 * dn^2 * |x|^2 of this thread's row of the K/Q tile (softmax kernel): the three warps of a
    // lane group sum channels [0,24) | [24,48) | [48,64), the partials meet in shared memory
    auto row_diag = [&](uint32_t a_seq) {
      mbar_wait(bar_tfull(a_seq % kRing), (a_seq / kRing) & 1u);  // TMA bytes visible to this thread
      const uint32_t tile = s_ring + (a_seq % kRing) * kSlabBytes;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < 2 || third < 2) {
          const uint4 v = ld_shared_v4(tile + sw128_offset(row, third * 24 + j * 8));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = bf16lo(w[i]), b = bf16hi(w[i]);
            s = fmaf(a, a, s);
            s = fmaf(b, b, s);
          }
        }
      }
      part[third * 128 + row] = s;
      named_bar_sync(1, kFeatWarps * 32);
      return (part[row] + part[128 + row] + part[256 + row]) * (0.5f * 0.125f);
    };

    // ---- software pipeline over jobs: while the packed features of job j are stored, fenced and
    // handed to the MMA issuer, the accumulator columns of job j+1 are already on their way from
    // TMEM into `raw`. Every job: (a) finish my load, release the U slot; (b) arithmetic into packed
    // registers; (c) prefetch the next job; (d) store / publish.
    uint32_t raw[48];
    // issue the TMEM loads of job nJ (chunk shape C): 48 columns, or the 16 columns of chunk 2
    auto prefetch = [&](auto cc) {
      constexpr int C = decltype(cc)::value;
      const uint32_t us = nJ & 1u;
      mbar_wait(bar_ufull(us), (nJ >> 1) & 1u);
      tc_fence_after();
      if constexpr (C < 2) {
        tmem_ld_32x32p(ubase + us * 128u, raw);
        tmem_ld_32x16p(ubase + us * 128u + 32u, raw + 32);
      } else {
        tmem_ld_32x16p(tmem + t_lane + kColU + us * 128u, raw);  // every warp reads columns 0..15
      }
    };
    // the loads of job nJ have landed: the U slot can be overwritten
    auto release_u = [&]() {
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ufree(nJ & 1u));
    };
    // 8 accumulator columns -> 4 packed bf16x2 features (columns >= lim are padding)
    auto feat8 = [&](const uint32_t* r, int c0, int lim, bool masked, uint32_t* pk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x0 = __uint_as_float(r[2 * i]), x1 = __uint_as_float(r[2 * i + 1]);
        if (KIND == 0)
          pk[i] = add_bf16x2(pack_bf16x2(ex2_approx(fmaf(x0, kLog2e, -sub)), ex2_approx(fmaf(x1, kLog2e, -sub))), eps2);
        else
          pk[i] = add_bf16x2(cvt_relu_bf16x2(x0, x1), eps2);
      }
      if (masked) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + 2 * i >= lim) pk[i] = 0u;
          else if (c0 + 2 * i + 1 >= lim) pk[i] &= 0x0000ffffu;
        }
      }
    };

    // stabiliser pass over one chunk (softmax kernel): max of the raw projections.
    // `next`: chunk shape of the following job (std::integral_constant), has_next: it exists
    auto max_job = [&](auto cc, auto next, bool has_next, float& acc) {
      constexpr int C = decltype(cc)::value;
      release_u();
      const int lim = C < 2 ? p.m - (C * 128 + col0) : (third == 0 ? p.m - 256 : 0);
      constexpr int NC = C < 2 ? 48 : 16;
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < NC; ++i)
        if (i < lim) mx = fmaxf(mx, __uint_as_float(raw[i]));
      acc = fmaxf(acc, mx);
      ++nJ;
      if (has_next) prefetch(next);
    };

    // feature map of one chunk: packed features -> the job's smem buffer -> fready
    auto feat_job = [&](auto cc, auto next, bool has_next, bool zero_row) {
      constexpr int C = decltype(cc)::value;
      constexpr int NC = C < 2 ? 48 : 16;
      const uint32_t us = nJ & 1u;  // U slot = feature buffer = job parity
      const uint32_t nF = us ? nF1++ : nF0++;
      TR(40 + C);
      release_u();
      TR(44);
      // padding: feature columns >= m, and key rows beyond the sequence, must contribute nothing
      const int lim = zero_row ? 0 : p.m - (C * 128 + (C < 2 ? col0 : 0));
      const bool masked = lim < NC;
      uint32_t pk[NC / 2];
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) feat8(raw + 8 * q, 8 * q, lim, masked, pk + 4 * q);
      TR(45);
      ++nJ;
      if (has_next) prefetch(next);
      TR(46);
      mbar_wait(bar_ffree(us), (nF & 1u) ^ 1u);  // the MMAs that last read this buffer are done
      const uint32_t fb = frow + us * 2u * kSlabBytes;
      if (C < 2 || third == 0) {
        const int ch0 = C < 2 ? (col0 >> 3) : 0;
#pragma unroll
        for (int q = 0; q < NC / 8; ++q)
          st_shared_v4(fb + fchunk(ch0 + q), make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]));
        TR(47);
        fence_proxy_async_smem();
      }
      TR(48);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fready(us));
      TR(49);
    };
    using C0 = std::integral_constant<int, 0>;
    using C1 = std::integral_constant<int, 1>;
    using C2 = std::integral_constant<int, 2>;

    if ((int64_t)blockIdx.x < p.items) prefetch(C0{});
    for (int64_t item = blockIdx.x; item < p.items; item += istride) {
      if (KIND == 0) {
        // ---- key stabiliser: global max of K.Omega'^T over the valid tokens ----
        float kmx = -INFINITY;
        for (int t = 0; t < nt; ++t) {
          ++tile_seq;
          float mx = -INFINITY;
          max_job(C0{}, C1{}, true, mx);
          max_job(C1{}, C2{}, true, mx);
          max_job(C2{}, C0{}, true, mx);
          if (t * kTile + row < p.tokens) kmx = fmaxf(kmx, mx);
        }
        kmx = warp_max(kmx);
        if (lane == 0) red[fw] = kmx;
        named_bar_sync(1, kFeatWarps * 32);
        gmax = red[0];
#pragma unroll
        for (int i = 1; i < kFeatWarps; ++i) gmax = fmaxf(gmax, red[i]);
      }
      // ---- keys: k' chunks feed the context MMAs ----
      for (int t = 0; t < nt; ++t) {
        if (KIND == 0) sub = (row_diag(tile_seq) + gmax) * kLog2e;
        tile_seq += 2;
        const bool zero_row = t * kTile + row >= p.tokens;
        feat_job(C0{}, C1{}, true, zero_row);
        feat_job(C1{}, C2{}, true, zero_row);
        feat_job(C2{}, C0{}, true, zero_row);
      }
      // ---- context read-out: TMEM ctx^T blocks -> bf16 K-major smem; warps of third b take block b ----
      {
        TR(60);
        mbar_wait(bar_ctxfull, nItems & 1u);
        TR(61);
        ++nItems;
        tc_fence_after();
        tmem_ld_wait();  // the prefetch of the first query job is in flight: one wait covers all loads
        if (third < 2 && !(dbg_bits() & 64)) {
          const int m = 128 * third + row;
          const uint32_t mc = (uint32_t)m & 63u;
          const uint32_t slab = s_ctx + (uint32_t)(m >> 6) * kCtxSlabBytes + (mc & 7u) * 2u;
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + t_lane + kColCtx + 80u * third + 32u * hlf, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const uint32_t n = 32u * hlf + i;
              st_shared_b16(slab + (n >> 3) * 1024u + (n & 7u) * 128u + ((((mc >> 3) ^ n) & 7u) << 4), __uint_as_float(r[i]));
            }
          }
          uint32_t r2[16];
          tmem_ld_32x16(tmem + t_lane + kColCtx + 80u * third + 64u, r2);
          tmem_ld_wait();
          st_shared_b16(slab + 8u * 1024u + (((mc >> 3) & 7u) << 4), __uint_as_float(r2[0]));  // n = 64
        }
        if (third == 2 && lg == 0 && !(dbg_bits() & 64)) {
          // block 2: features 256..271 live in lanes 0..15
          const uint32_t mc = (uint32_t)lane;
          const uint32_t slab = s_ctx + 4u * kCtxSlabBytes + (mc & 7u) * 2u;
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + kColCtx + 160u + 32u * hlf, r);
            tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const uint32_t n = 32u * hlf + i;
                st_shared_b16(slab + (n >> 3) * 1024u + (n & 7u) * 128u + ((((mc >> 3) ^ n) & 7u) << 4), __uint_as_float(r[i]));
              }
            }
          }
          uint32_t r2[16];
          tmem_ld_32x16(tmem + kColCtx + 160u + 64u, r2);
          tmem_ld_wait();
          if (lane < 16) st_shared_b16(slab + 8u * 1024u + (((mc >> 3) & 7u) << 4), __uint_as_float(r2[0]));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ctxready);
        TR(62);
      }
      // ---- queries: q' chunks feed the output MMAs; the out/den epilogue of tile t runs behind
      //      the first job of tile t+1 ----
      bool pending = false;
      for (int t = 0; t < nt; ++t) {
        const bool more = t + 1 < nt || item != last_item;  // another job follows this tile
        if (KIND == 0) {
          diag = row_diag(tile_seq);
          float rmx = -INFINITY;
          max_job(C0{}, C1{}, true, rmx);
          if (pending) { epilogue(item, t - 1); pending = false; }
          max_job(C1{}, C2{}, true, rmx);
          max_job(C2{}, C0{}, true, rmx);
          rmaxs[third * 128 + row] = rmx;
          named_bar_sync(1, kFeatWarps * 32);
          sub = (diag + fmaxf(fmaxf(rmaxs[row], rmaxs[128 + row]), rmaxs[256 + row])) * kLog2e;
        }
        ++tile_seq;
        feat_job(C0{}, C1{}, true, false);
        if (pending) { epilogue(item, t - 1); pending = false; }
        feat_job(C1{}, C2{}, true, false);
        feat_job(C2{}, C0{}, more, false);
        pending = true;
      }
      epilogue(item, nt - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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

template <int KIND>
int launch_kind(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FavorTcParams& p,
                cudaStream_t stream) {
  static PerDeviceOnce once;
  const int cfg_rc = per_device_once(once, []() {
    return cuda_status(cudaFuncSetAttribute(favor_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  int grid = num_sms();
  if (p.items < grid) grid = (int)p.items;
  favor_tc_kernel<KIND><<<grid, kThreads, kSmemBytes, stream>>>(tq, tk, tv, p);
  return post_launch();
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
  p.proj = d->proj; p.out = d->out; p.m = d->m_features; p.heads = d->heads;
  p.tokens = (int)d->tokens; p.G0 = d->G[0]; p.G1 = d->G[1];
  p.items = d->G[0] * d->G[1] * d->heads;
  p.ogs0 = d->out_gs[0]; p.ogs1 = d->out_gs[1]; p.ots = d->out_ts;
  {
    static const int dbg = []() { const char* e = getenv("RFK_FAVOR_DBG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg;
  }
  static const char* trace_path = getenv("RFK_FAVOR_TRACE");
  if (trace_path) {
    // developer tool: timeline of CTA 0 (clock64 per event), dumped after a synchronous launch
    const size_t bytes = sizeof(long long) * 4 * kTraceMax * 2;
    cudaMalloc(&p.trace, bytes);
    cudaMemsetAsync(p.trace, 0, bytes, stream);
    rc = d->kind == 0 ? launch_kind<0>(tq, tk, tv, p, stream) : launch_kind<1>(tq, tk, tv, p, stream);
    cudaStreamSynchronize(stream);
    long long* h = (long long*)malloc(bytes);
    cudaMemcpy(h, p.trace, bytes, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int r = 0; r < 4; ++r)
        for (int i = 0; i < kTraceMax && h[((size_t)r * kTraceMax + i) * 2 + 1]; ++i)
          fprintf(f, "%d %lld %lld\n", r, h[((size_t)r * kTraceMax + i) * 2], h[((size_t)r * kTraceMax + i) * 2 + 1]);
      fclose(f);
    }
    free(h);
    cudaFree(p.trace);
    return rc;
  }
  return d->kind == 0 ? launch_kind<0>(tq, tk, tv, p, stream) : launch_kind<1>(tq, tk, tv, p, stream);
}

}  // namespace rfk
