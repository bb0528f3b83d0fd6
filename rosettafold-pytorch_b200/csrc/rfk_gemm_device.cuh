// rfk_gemm_device.cuh — device code of the batched TN GEMM (tcgen05 kernel + epilogues).
#pragma once
#include "rfk_common.cuh"

namespace rfk {

struct GemmDev {
  int64_t M, N, K;
  int64_t Z0, Z1, Z2;
  int64_t MR, NR;
  float alpha;
  int act;
  int epi;
  float ln_eps;
  const float* bias;
  int64_t bias_zs[3];
  void* c;
  const void* r0;
  const void* r1;
  int c_dtype, r0_dtype, r1_dtype;
  rfk_addr c_addr, r0_addr, r1_addr;
  const float* ln_gamma;
  const float* ln_beta;
  // tcgen05 path only
  int bmask[3];  // 0 -> broadcast B over that z level
  int ab_f16;    // A / B are IEEE half instead of bf16 (instruction-descriptor format bits)
  // implicit-GEMM 3x3 convolution mode (CONV kernels): A is the NHWC image [B][L][L][C]; a tile is
  // 128 consecutive columns j of one image row; K runs over 9 taps x conv_cblocks 64-channel blocks
  int conv_L, conv_H, conv_Lp, conv_cblocks, conv_last_k16, conv_cpad;  // conv_L = image width, conv_H = rows
  int conv_dil;  // dilation: tap (di, dj) reads the image at (i + di * dil, j + dj * dil)
  // TMA-store epilogues (EPI 3/4): tensor-map dimension d takes logical coordinate cmap[d] of
  // {0: n % NR, 1: n / NR, 2: m % MR, 3: m / MR, 4: z0, 5: z1, 6: z2}; -1 -> 0
  int cmap[5];
  int has_rmap;  // residual r0 is fetched with TMA (map tma_r)
  // SIMT path only
  const float* a32;
  const float* b32;
  int64_t lda, ldb;
  int64_t a_zs[3], b_zs[3];
};

// m, n, MR, NR all fit 32 bits (checked on the host): 32-bit div/mod is several times cheaper
__device__ __forceinline__ int64_t addr_zm(const rfk_addr& a, int64_t z0, int64_t z1, int64_t z2,
                                           int64_t m, int64_t MR) {
  const uint32_t mu = (uint32_t)m, mr = (uint32_t)MR;
  const uint32_t q = mu / mr, r = mu - q * mr;
  return z0 * a.zs[0] + z1 * a.zs[1] + z2 * a.zs[2] + (int64_t)r * a.ms[0] + (int64_t)q * a.ms[1];
}
__device__ __forceinline__ int64_t addr_n(const rfk_addr& a, int64_t n, int64_t NR) {
  const uint32_t nu = (uint32_t)n, nr = (uint32_t)NR;
  const uint32_t q = nu / nr, r = nu - q * nr;
  return (int64_t)r * a.ns[0] + (int64_t)q * a.ns[1];
}

// Epilogue for CH consecutive columns [n0, n0+CH) of one row m (one thread).
template <int CH>
__device__ __forceinline__ void epilogue_row_chunk(const GemmDev& p, int64_t z0, int64_t z1,
                                                   int64_t z2, int64_t m, int64_t n0,
                                                   float (&v)[CH]) {
  if (m >= p.M || n0 >= p.N) return;
  const bool full = (n0 + CH <= p.N);
  const bool same_block = (((uint32_t)n0 % (uint32_t)p.NR) + CH <= (uint32_t)p.NR);
  const float* bias = p.bias ? p.bias + z0 * p.bias_zs[0] + z1 * p.bias_zs[1] + z2 * p.bias_zs[2]
                             : nullptr;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float x = v[i] * p.alpha;
    if (bias && (full || n0 + i < p.N)) x += __ldg(bias + n0 + i);
    v[i] = apply_act(x, p.act);
  }
  // residual addends
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const void* r = which == 0 ? p.r0 : p.r1;
    if (!r) continue;
    const rfk_addr& ra = which == 0 ? p.r0_addr : p.r1_addr;
    const int rdt = which == 0 ? p.r0_dtype : p.r1_dtype;
    const int64_t base = addr_zm(ra, z0, z1, z2, m, p.MR);
    if (full && same_block && ra.ns[0] == 1) {
      const int64_t off = base + addr_n(ra, n0, p.NR);
      if (rdt == RFK_F32) {
        const float* rp = reinterpret_cast<const float*>(r) + off;
        if ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < CH; i += 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(rp + i));
            v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < CH; ++i) v[i] += __ldg(rp + i);
        }
      } else {
        const uint16_t* rp = reinterpret_cast<const uint16_t*>(r) + off;
#pragma unroll
        for (int i = 0; i < CH; ++i) v[i] += h16_to_float(rp[i], rdt);
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (full || n0 + i < p.N) v[i] += load_as_float(r, rdt, base + addr_n(ra, n0 + i, p.NR));
    }
  }
  // store
  const int64_t cbase = addr_zm(p.c_addr, z0, z1, z2, m, p.MR);
  if (full && same_block && p.c_addr.ns[0] == 1) {
    const int64_t off = cbase + addr_n(p.c_addr, n0, p.NR);
    if (p.c_dtype == RFK_F32) {
      float* cp = reinterpret_cast<float*>(p.c) + off;
      if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < CH; i += 4)
          *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < CH; ++i) cp[i] = v[i];
      }
    } else {
      uint16_t* cp = reinterpret_cast<uint16_t*>(p.c) + off;
      if (CH % 8 == 0 && (reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
#pragma unroll
        for (int i = 0; i + 7 < CH; i += 8) {
          uint4 t;
          t.x = pack_h16x2(v[i], v[i + 1], p.c_dtype);
          t.y = pack_h16x2(v[i + 2], v[i + 3], p.c_dtype);
          t.z = pack_h16x2(v[i + 4], v[i + 5], p.c_dtype);
          t.w = pack_h16x2(v[i + 6], v[i + 7], p.c_dtype);
          *reinterpret_cast<uint4*>(cp + i) = t;
        }
      } else {
#pragma unroll
        for (int i = 0; i < CH; ++i) cp[i] = cvt_h16(v[i], p.c_dtype);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      if (full || n0 + i < p.N)
        store_from_float(p.c, p.c_dtype, cbase + addr_n(p.c_addr, n0 + i, p.NR), v[i]);
  }
}

// LayerNorm over an aligned 32x32 block held by one warp: lane = m%32, v[i] = column n0+i.
// Every lane of the warp must call this (rows/columns outside the problem hold zeros).
__device__ __forceinline__ void blockln32(const GemmDev& p, int lane, float (&v)[32]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / 1024.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 1024.f) + p.ln_eps);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float x = (v[i] - mean) * rstd;
    if (p.ln_gamma) x = x * __ldg(p.ln_gamma + lane * 32 + i) + __ldg(p.ln_beta + lane * 32 + i);
    v[i] = x;
  }
}


// ----------------------------------------------------------------------------------------------
// Warp-cooperative epilogue for a 32-row x CW-column chunk (tcgen05 kernel). After tcgen05.ld a
// thread owns one ROW of the chunk; writing it out directly makes every store instruction touch 32
// different rows. The chunk is therefore staged through shared memory ([32][CW+1] floats per
// warp, conflict-free both ways) and re-read so that a store instruction covers whole 64/128-byte
// row segments ("mode R": columns contiguous in memory) or whole 64/128-byte column segments
// ("mode C": rows contiguous in memory — the K^T/V^T relayouts). Residual loads are coalesced the
// same way. Anything else falls back to the per-thread path.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4_any(const void* r, int rdt, int64_t off, float (&o)[4]) {
  if (rdt == RFK_F32) {
    const float* rp = reinterpret_cast<const float*>(r) + off;
    if ((reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(rp));
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = __ldg(rp + j);
    }
  } else {
    const uint16_t* rp = reinterpret_cast<const uint16_t*>(r) + off;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = h16_to_float(rp[j], rdt);
  }
}

template <int CW>
__device__ __forceinline__ void epilogue_warp_chunk(const GemmDev& p, float* st, int lane, int64_t z0,
                                                    int64_t z1, int64_t z2, int64_t m_warp0, int64_t n0,
                                                    float (&v)[CW], int64_t c_row, int64_t r0_row,
                                                    int64_t r1_row, const float* bias) {
  if (n0 >= p.N || m_warp0 >= p.M) return;  // warp-uniform
  const bool full = (n0 + CW <= p.N);
  const bool same_block = (((uint32_t)n0 % (uint32_t)p.NR) + CW <= (uint32_t)p.NR);
  const bool res_rowmajor = (!p.r0 || p.r0_addr.ns[0] == 1) && (!p.r1 || p.r1_addr.ns[0] == 1);
  const bool modeR = full && same_block && p.c_addr.ns[0] == 1 && res_rowmajor;
  const bool rows_ok = (m_warp0 + 32 <= p.M) && (((uint32_t)m_warp0 % (uint32_t)p.MR) + 32 <= (uint32_t)p.MR);
  const bool modeC = !modeR && full && rows_ok && p.c_addr.ms[0] == 1 && !p.r0 && !p.r1;
  if (!modeR && !modeC) {
    epilogue_row_chunk<CW>(p, z0, z1, z2, m_warp0 + lane, n0, v);
    return;
  }
  constexpr int LD = CW + 1;
  if (modeR) {
    // row-major staging with the 16-byte chunk index XOR-swizzled by (row & 7): 128-bit stores by
    // row owners and 128-bit loads by (row, chunk) owners are both bank-conflict free
    constexpr int CPRW = CW / 4;  // 16-byte chunks per row
    float4* st4 = reinterpret_cast<float4*>(st);
#pragma unroll
    for (int c = 0; c < CPRW; ++c)
      st4[lane * CPRW + (c ^ (lane & (CPRW - 1) & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    auto ld4 = [&](int row, int col, float (&o)[4]) {
      const float4 t = st4[row * CPRW + ((col >> 2) ^ (row & (CPRW - 1) & 7))];
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    };
    const int64_t ccol = addr_n(p.c_addr, n0, p.NR);
    const int64_t r0col = p.r0 ? addr_n(p.r0_addr, n0, p.NR) : 0;
    const int64_t r1col = p.r1 ? addr_n(p.r1_addr, n0, p.NR) : 0;
    if (p.c_dtype == RFK_F32) {
      constexpr int LPR = CW / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
      const int rsub = lane / LPR, col = (lane % LPR) * 4;
      float ra[ITERS][4], rb[ITERS][4];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int row = it * RPI + rsub;
        const int64_t o0 = __shfl_sync(0xffffffffu, r0_row, row);
        const int64_t o1 = __shfl_sync(0xffffffffu, r1_row, row);
        const bool ok = m_warp0 + row < p.M;
#pragma unroll
        for (int j = 0; j < 4; ++j) ra[it][j] = rb[it][j] = 0.f;
        if (ok && p.r0) load4_any(p.r0, p.r0_dtype, o0 + r0col + col, ra[it]);
        if (ok && p.r1) load4_any(p.r1, p.r1_dtype, o1 + r1col + col, rb[it]);
      }
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = __ldg(bias + n0 + col + j);
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int row = it * RPI + rsub;
        const int64_t oc = __shfl_sync(0xffffffffu, c_row, row);
        if (m_warp0 + row < p.M) {
          float x[4];
          ld4(row, col, x);
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = apply_act(x[j] * p.alpha + bv[j], p.act) + ra[it][j] + rb[it][j];
          float* cp = reinterpret_cast<float*>(p.c) + oc + ccol + col;
          if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
            *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) cp[j] = x[j];
          }
        }
      }
    } else {
      constexpr int LPR = CW / 8, RPI = 32 / LPR, ITERS = 32 / RPI;
      const int rsub = lane / LPR, col = (lane % LPR) * 8;
      float bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = bias ? __ldg(bias + n0 + col + j) : 0.f;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int row = it * RPI + rsub;
        const int64_t oc = __shfl_sync(0xffffffffu, c_row, row);
        const int64_t o0 = __shfl_sync(0xffffffffu, r0_row, row);
        const int64_t o1 = __shfl_sync(0xffffffffu, r1_row, row);
        if (m_warp0 + row < p.M) {
          float x[8];
          {
            float a4[4], b4[4];
            ld4(row, col, a4);
            ld4(row, col + 4, b4);
#pragma unroll
            for (int j = 0; j < 4; ++j) { x[j] = a4[j]; x[4 + j] = b4[j]; }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j] * p.alpha + bv[j], p.act);
          if (p.r0) {
            float t[4];
            load4_any(p.r0, p.r0_dtype, o0 + r0col + col, t);
            x[0] += t[0]; x[1] += t[1]; x[2] += t[2]; x[3] += t[3];
            load4_any(p.r0, p.r0_dtype, o0 + r0col + col + 4, t);
            x[4] += t[0]; x[5] += t[1]; x[6] += t[2]; x[7] += t[3];
          }
          if (p.r1) {
            float t[4];
            load4_any(p.r1, p.r1_dtype, o1 + r1col + col, t);
            x[0] += t[0]; x[1] += t[1]; x[2] += t[2]; x[3] += t[3];
            load4_any(p.r1, p.r1_dtype, o1 + r1col + col + 4, t);
            x[4] += t[0]; x[5] += t[1]; x[6] += t[2]; x[7] += t[3];
          }
          uint16_t* cp = reinterpret_cast<uint16_t*>(p.c) + oc + ccol + col;
          if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
            uint4 t;
            t.x = pack_h16x2(x[0], x[1], p.c_dtype); t.y = pack_h16x2(x[2], x[3], p.c_dtype);
            t.z = pack_h16x2(x[4], x[5], p.c_dtype); t.w = pack_h16x2(x[6], x[7], p.c_dtype);
            *reinterpret_cast<uint4*>(cp) = t;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) cp[j] = cvt_h16(x[j], p.c_dtype);
          }
        }
      }
    }
  } else {
    // mode C: rows are contiguous in memory; lane i (< CW) owns the offset/bias of column n0 + i
#pragma unroll
    for (int i = 0; i < CW; ++i) st[lane * LD + i] = v[i];
    __syncwarp();
    const int64_t coff_mine = lane < CW ? addr_n(p.c_addr, n0 + lane, p.NR) : 0;
    const float bias_mine = (bias && lane < CW) ? __ldg(bias + n0 + lane) : 0.f;
    const int64_t rowbase = __shfl_sync(0xffffffffu, c_row, 0);
    if (p.c_dtype != RFK_F32) {
      const int csub = lane >> 2, r8 = (lane & 3) * 8;
#pragma unroll
      for (int it = 0; it < CW / 8; ++it) {
        const int col = it * 8 + csub;
        const int64_t coff = __shfl_sync(0xffffffffu, coff_mine, col);
        const float b = __shfl_sync(0xffffffffu, bias_mine, col);
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = apply_act(st[(r8 + j) * LD + col] * p.alpha + b, p.act);
        uint16_t* cp = reinterpret_cast<uint16_t*>(p.c) + rowbase + r8 + coff;
        if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
          uint4 t;
          t.x = pack_h16x2(x[0], x[1], p.c_dtype); t.y = pack_h16x2(x[2], x[3], p.c_dtype);
          t.z = pack_h16x2(x[4], x[5], p.c_dtype); t.w = pack_h16x2(x[6], x[7], p.c_dtype);
          *reinterpret_cast<uint4*>(cp) = t;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) cp[j] = cvt_h16(x[j], p.c_dtype);
        }
      }
    } else {
      const int csub = lane >> 3, r4 = (lane & 7) * 4;
#pragma unroll
      for (int it = 0; it < CW / 4; ++it) {
        const int col = it * 4 + csub;
        const int64_t coff = __shfl_sync(0xffffffffu, coff_mine, col);
        const float b = __shfl_sync(0xffffffffu, bias_mine, col);
        float x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = apply_act(st[(r4 + j) * LD + col] * p.alpha + b, p.act);
        float* cp = reinterpret_cast<float*>(p.c) + rowbase + r4 + coff;
        if ((reinterpret_cast<uintptr_t>(cp) & 15) == 0) {
          *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) cp[j] = x[j];
        }
      }
    }
  }
  __syncwarp();
}


// ----------------------------------------------------------------------------------------------
// Lean epilogues (EPI 1 / 2). The host only selects them when: alpha == 1, N % 32 == 0, output (and
// residual) columns contiguous (ns[0] == 1) with every 32-column chunk inside one NR block, the 32
// rows of a warp inside one MR block (MR % 32 == 0 or MR >= M) so row offsets are affine, and all
// base pointers / strides 16-byte aligned. That removes every per-element predicate, alignment
// test and shuffle from the hot loop.
//   EPI 1: bf16 output, no residual.     EPI 2: f32 output, up to two f32 residual addends.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int EPI>
__device__ __forceinline__ void epilogue_fast_chunk(const GemmDev& p, float4* st4, int lane,
                                                    int rows_valid, int64_t n0, float (&v)[32],
                                                    int64_t c_base, int64_t r0_base, int64_t r1_base,
                                                    const float* bias) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    st4[lane * 8 + (c ^ (lane & 7))] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  __syncwarp();
  const int64_t ccol = addr_n(p.c_addr, n0, p.NR);
  if constexpr (EPI == 1) {
    const int rsub = lane >> 2, col = (lane & 3) * 8;
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    if (bias) {
      b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col));
      b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + col + 4));
    }
    const bool relu = p.act == RFK_ACT_RELU;
    uint16_t* cbase = reinterpret_cast<uint16_t*>(p.c) + c_base + ccol + col;
    const int64_t ms0 = p.c_addr.ms[0];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int row = it * 8 + rsub;
      const int sw = row & 7;
      float4 a = st4[row * 8 + ((col >> 2) ^ sw)];
      float4 b = st4[row * 8 + (((col >> 2) + 1) ^ sw)];
      a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      b.x += b1.x; b.y += b1.y; b.z += b1.z; b.w += b1.w;
      if (relu) {
        a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
      }
      uint4 t;
      t.x = pack_h16x2(a.x, a.y, p.c_dtype); t.y = pack_h16x2(a.z, a.w, p.c_dtype);
      t.z = pack_h16x2(b.x, b.y, p.c_dtype); t.w = pack_h16x2(b.z, b.w, p.c_dtype);
      if (row < rows_valid) *reinterpret_cast<uint4*>(cbase + row * ms0) = t;
    }
  } else {
    const int rsub = lane >> 3, col = (lane & 7) * 4;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + col));
    const bool relu = p.act == RFK_ACT_RELU;
    float4 ra[8], rb[8];
    if (p.r0) {
      const float* rp = reinterpret_cast<const float*>(p.r0) + r0_base + addr_n(p.r0_addr, n0, p.NR) + col;
      const int64_t rs = p.r0_addr.ms[0];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rsub;
        ra[it] = row < rows_valid ? __ldg(reinterpret_cast<const float4*>(rp + row * rs)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (p.r1) {
      const float* rp = reinterpret_cast<const float*>(p.r1) + r1_base + addr_n(p.r1_addr, n0, p.NR) + col;
      const int64_t rs = p.r1_addr.ms[0];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + rsub;
        rb[it] = row < rows_valid ? __ldg(reinterpret_cast<const float4*>(rp + row * rs)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float* cbase = reinterpret_cast<float*>(p.c) + c_base + ccol + col;
    const int64_t ms0 = p.c_addr.ms[0];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = it * 4 + rsub;
      float4 a = st4[row * 8 + ((col >> 2) ^ (row & 7))];
      a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
      if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
      if (p.r0) { a.x += ra[it].x; a.y += ra[it].y; a.z += ra[it].z; a.w += ra[it].w; }
      if (p.r1) { a.x += rb[it].x; a.y += rb[it].y; a.z += rb[it].z; a.w += rb[it].w; }
      if (row < rows_valid) *reinterpret_cast<float4*>(cbase + row * ms0) = a;
    }
  }
  __syncwarp();
}

// ----------------------------------------------------------------------------------------------
// tcgen05 kernel
// ----------------------------------------------------------------------------------------------
constexpr int kBlockM = 128;
#ifndef RFK_EPI3_RING
#define RFK_EPI3_RING 4
#endif
constexpr int kBlockK = 64;
#ifndef RFK_EPI3_WARPS
#define RFK_EPI3_WARPS 8
#endif
// epilogue warps: a multiple of 4 (a warp reaches the TMEM lane quarter warp % 4 only); the warps of one quarter
// take the 32-column chunks of a tile round-robin
__host__ __device__ constexpr int epi_warps(int EPI) { return EPI == 3 ? RFK_EPI3_WARPS : 8; }
__host__ __device__ constexpr int gemm_threads(int EPI) { return 32 * (2 + epi_warps(EPI)); }  // + TMA warp, MMA warp

template <int BN, int EPI = 0>
struct GemmCfg {
  static constexpr int kEpiWarps = epi_warps(EPI);
  static constexpr int kChunkStride = kEpiWarps / 4;  // warps per TMEM lane quarter
  static constexpr int kStageBytes = kBlockM * 128 + BN * 128;
  // epilogue staging: EPI 0-2 one [32][33] f32 transpose buffer per warp; EPI 3: two 2 KB bf16
  // tiles per warp; EPI 4: three 4 KB f32 tiles per warp (residual-in / result-out ring)
  static constexpr int kRing3 = RFK_EPI3_RING;
  static constexpr int kStagingBytes = EPI == 3 ? kEpiWarps * kRing3 * 2048
                                       : EPI == 4 ? kEpiWarps * 3 * 4096 : kEpiWarps * 32 * 33 * 4;
  static constexpr int kBarBytes = 512;
  static constexpr int kBudget = 232448 - 2048 - kBarBytes - kStagingBytes;
  static constexpr int kStages = kBudget / kStageBytes > 8 ? 8 : kBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128
                                   : 2 * BN <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2048 /*align*/ + kBarBytes + kStagingBytes;
};

// one lane of the (fully active) warp
__device__ __forceinline__ bool gemm_elect_one() {
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

// tile index -> (n block, m block, z0, z1, z2), n fastest
struct TileDecoder {
  uint32_t n_blocks, m_blocks, Z0, Z1, Zn;
  __device__ __forceinline__ void operator()(uint32_t t, uint32_t& nb, uint32_t& mb, uint32_t& z0, uint32_t& z1,
                                             uint32_t& z2) const {
    const uint32_t q = t / n_blocks;
    nb = t - q * n_blocks;
    if (Zn == 1) {
      mb = q; z0 = z1 = z2 = 0;
    } else {
      const uint32_t z = q / m_blocks;
      mb = q - z * m_blocks;
      const uint32_t zq = z / Z0;
      z0 = z - zq * Z0;
      z2 = zq / Z1;
      z1 = zq - z2 * Z1;
    }
  }
};

template <int BN, int EPI, bool CONV = false>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r,
               const GemmDev p) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::kStageBytes;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  // EPI 4: one residual-arrival barrier per (epilogue warp, ring slot)
  auto res_bar = [&](int w, int slot) { return bar_base + 256u + 8u * (w * 3 + slot); };
  // staging area, 1024-byte aligned (TMA-store swizzle patterns are address based)
  const uint32_t stg_base = (bar_base + Cfg::kBarBytes + 1023u) & ~1023u;
  auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + kBlockM * 128; };

  // warp index through a shuffle: the compiler then keeps it (and everything derived from it: TMEM lane
  // group, staging slots, TMA-store coordinates) in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), Cfg::kEpiWarps);
    }
    if constexpr (EPI == 4)
      for (int w = 0; w < Cfg::kEpiWarps; ++w)
        for (int sl = 0; sl < 3; ++sl) mbar_init(res_bar(w, sl), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // tile bookkeeping in 32 bits (the host checks that M, N and the tile count fit): a 64-bit division is a
  // ~100-instruction subroutine, and every role decodes its tile index once per tile
  const uint32_t m_blocks = (uint32_t)((p.M + kBlockM - 1) / kBlockM);
  const uint32_t n_blocks = (uint32_t)((p.N + BN - 1) / BN);
  const int k_blocks = CONV ? 9 * p.conv_cblocks : (int)((p.K + kBlockK - 1) / kBlockK);
  const uint32_t Zn = (uint32_t)(p.Z0 * p.Z1 * p.Z2);
  const uint32_t tiles = Zn * m_blocks * n_blocks;
  const TileDecoder tdec{n_blocks, m_blocks, (uint32_t)p.Z0, (uint32_t)p.Z1, Zn};

  if (warp == 0) {
    // ===== TMA producer =====
    // (this role and the MMA issuer run warp-uniform control flow with one elected lane issuing:
    // under a lane-0-only branch the compiler treats every operand as divergent and wraps each
    // UTMALDG / UTCHMMA in per-instruction election code)
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      uint32_t nb, mb, uz0, uz1, uz2;
      tdec(t, nb, mb, uz0, uz1, uz2);
      const int z0 = (int)uz0, z1 = (int)uz1, z2 = (int)uz2;
      if constexpr (CONV) {
        // tile -> (image b, row i, first column j0); taps shift the TMA box, out-of-image rows and
        // columns are zero-filled by the TMA unit ('same' padding for free)
        const uint32_t m0 = mb * kBlockM;
        const int img_row = (int)(m0 / (uint32_t)p.conv_Lp), j0 = (int)(m0 % (uint32_t)p.conv_Lp);
        const int bi = img_row / p.conv_H, ii = img_row % p.conv_H;
        for (int tap = 0; tap < 9; ++tap) {
          const int di = tap / 3 - 1, dj = tap % 3 - 1;
          for (int cb = 0; cb < p.conv_cblocks; ++cb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (gemm_elect_one()) {
              mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
              tma_load_5d(&tma_a, full_bar(stage), smem_a(stage), cb * kBlockK, j0 + dj * p.conv_dil, ii + di * p.conv_dil, bi, 0);
              tma_load_5d(&tma_b, full_bar(stage), smem_b(stage), tap * p.conv_cpad + cb * kBlockK,
                          (int)(nb * BN), 0, 0, 0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      } else {
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (gemm_elect_one()) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_5d(&tma_a, full_bar(stage), smem_a(stage), kb * kBlockK,
                        (int)(mb * kBlockM), z0, z1, z2);
            tma_load_5d(&tma_b, full_bar(stage), smem_b(stage), kb * kBlockK, (int)(nb * BN),
                        z0 & p.bmask[0], z1 & p.bmask[1], z2 & p.bmask[2]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t fmt = p.ab_f16 ? kIdescBf16Bits : 0u;  // operands are f16 instead of bf16
    const uint32_t idesc = umma_idesc_bf16(kBlockM, BN) ^ fmt;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // the last column block may be narrower than BN: its MMAs only cover the real columns (rounded up to the
    // instruction granularity of 16), so a wide BN costs no tensor time on the ragged edge
    const uint32_t n_tail = (uint32_t)p.N - (n_blocks - 1) * (uint32_t)BN;
    const uint32_t idesc_tail = umma_idesc_bf16(kBlockM, (int)((n_tail + 15u) & ~15u)) ^ fmt;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      const uint32_t idesc_t = (t % n_blocks == n_blocks - 1) ? idesc_tail : idesc;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_a(stage));
        const uint64_t bdesc = umma_desc_sw128(smem_b(stage));
        // conv: the last channel block of a tap may hold fewer than 64 real channels
        // (the K tail of a plain GEMM is zero-filled by TMA: skip the all-zero k16 steps too)
        const int nk16 = CONV ? ((kb % p.conv_cblocks == p.conv_cblocks - 1) ? p.conv_last_k16 : kBlockK / 16)
                              : ((kb == k_blocks - 1) ? ((int)p.K - kb * kBlockK + 15) / 16 : kBlockK / 16);
        if (gemm_elect_one()) {
          if (nk16 == kBlockK / 16) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_t, (kb > 0 || k > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              if (k < nk16)
                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_t, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (kb == k_blocks - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 2) {
    // ===== epilogue warps (TMEM lane group = warp % 4) =====
    const int lg = warp & 3;
    float* stage = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw))) + (warp - 2) * 32 * 33;
    const int chalf = (warp - 2) >> 2;  // the warps of a TMEM lane group take its chunks round-robin
    uint32_t epi_count = 0;  // chunks stored so far by this warp (ring slot / residual barrier parity)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      uint32_t unb, umb, uz0, uz1, uz2;
      tdec(t, unb, umb, uz0, uz1, uz2);
      const int64_t nb = unb, mb = umb, z0 = uz0, z1 = uz1, z2 = uz2;
      const int64_t m_warp0 = mb * kBlockM + lg * 32;
      const float* bias = p.bias ? p.bias + z0 * p.bias_zs[0] + z1 * p.bias_zs[1] + z2 * p.bias_zs[2] : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (EPI == 0 && CONV) {
        // generic conv epilogue (channel counts that are not multiples of 32): per-thread rows,
        // rows in the padding of an image row are skipped
        const int64_t m = m_warp0 + lane;
        const bool row_ok = m < p.M && (int)((uint32_t)m % (uint32_t)p.conv_Lp) < p.conv_L;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_row_chunk<32>(p, z0, z1, z2, row_ok ? m : p.M, nb * BN + c * 32, v);
        }
      } else if constexpr (EPI == 0) {
        const int64_t m = m_warp0 + lane;
        // this lane's row offsets (shared across the warp by shuffle in the staged epilogue)
        const int64_t mm = m < p.M ? m : 0;
        const int64_t c_row = addr_zm(p.c_addr, z0, z1, z2, mm, p.MR);
        const int64_t r0_row = p.r0 ? addr_zm(p.r0_addr, z0, z1, z2, mm, p.MR) : 0;
        const int64_t r1_row = p.r1 ? addr_zm(p.r1_addr, z0, z1, z2, mm, p.MR) : 0;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.epi == RFK_EPI_BLOCKLN32) blockln32(p, lane, v);
          epilogue_warp_chunk<32>(p, stage, lane, z0, z1, z2, m_warp0, nb * BN + c * 32, v, c_row, r0_row,
                                  r1_row, bias);
        }
        if (BN % 32 != 0 && chalf == ((BN / 32) & 1)) {
          uint32_t r[16];
          tmem_ld_32x16(taddr + (BN / 32) * 32, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_warp_chunk<16>(p, stage, lane, z0, z1, z2, m_warp0, nb * BN + (BN / 32) * 32, v, c_row,
                                  r0_row, r1_row, bias);
        }
      } else if constexpr (EPI == 3 || EPI == 4) {
        // ---- TMA-store epilogues: registers -> swizzled smem tile -> cp.async.bulk.tensor store.
        // No per-row address arithmetic, no transposing reads; partial tiles are clipped by TMA.
        static_assert(BN % 32 == 0, "TMA epilogues need 32-column chunks");
        constexpr uint32_t kBuf = EPI == 3 ? 2048u : 4096u;
        constexpr int kRing = EPI == 3 ? Cfg::kRing3 : 3;
        constexpr int kCS = Cfg::kChunkStride;
        const int ew = warp - 2;
        const uint32_t ring = stg_base + (uint32_t)ew * kRing * kBuf;
        // Everything the TMA instructions take is warp-uniform and lives in registers: the per-dimension
        // fixed coordinates are resolved once per tile (select chains, no indexed local array), the column
        // pair (n % NR, n / NR) is advanced incrementally (no division per chunk), and the instructions are
        // issued by an elected lane under warp-uniform control flow. (The same code under `if (lane == 0)`
        // with a division and a cmap-indexed array per chunk kept the epilogue warps busy for ~1800 cycles per
        // 32x32 chunk - more than the main loop needs per tile: profiles/r01_gemm_epilogue_notes.md.)
        const uint32_t nr = (uint32_t)p.NR;
        int fx[5];
        {
          const uint32_t mu = (uint32_t)m_warp0, mr = (uint32_t)p.MR;
          const int l3 = (int)(mu / mr), l2 = (int)(mu - (uint32_t)l3 * mr);
#pragma unroll
          for (int d = 0; d < 5; ++d) {
            const int cm = p.cmap[d];
            fx[d] = cm == 2 ? l2 : cm == 3 ? l3 : cm == 4 ? (int)uz0 : cm == 5 ? (int)uz1 : cm == 6 ? (int)uz2 : 0;
          }
        }
#define RFK_COORDS(lo, hi)                                                                                   \
  (p.cmap[0] == 0 ? (int)(lo) : p.cmap[0] == 1 ? (int)(hi) : fx[0]),                                         \
      (p.cmap[1] == 0 ? (int)(lo) : p.cmap[1] == 1 ? (int)(hi) : fx[1]),                                     \
      (p.cmap[2] == 0 ? (int)(lo) : p.cmap[2] == 1 ? (int)(hi) : fx[2]),                                     \
      (p.cmap[3] == 0 ? (int)(lo) : p.cmap[3] == 1 ? (int)(hi) : fx[3]),                                     \
      (p.cmap[4] == 0 ? (int)(lo) : p.cmap[4] == 1 ? (int)(hi) : fx[4])
        auto advance = [&](uint32_t& lo, uint32_t& hi) {  // next chunk of this warp: 64 columns on
          lo += 32u * kCS;
          while (lo >= nr) { lo -= nr; ++hi; }
        };
        const uint32_t n_tile0 = unb * (uint32_t)BN;
        const uint32_t n_first = n_tile0 + (uint32_t)chalf * 32u;
        uint32_t nhi = n_first / nr, nlo = n_first - nhi * nr;   // columns of the chunk being stored
        uint32_t phi = nhi, plo = nlo;                           // columns of the next residual tile to fetch
        // chunks of this tile owned by this warp: c = chalf + kCS i
        const int ct = min(BN / 32, (int)(((uint32_t)p.N - n_tile0 + 31u) >> 5));
        const int nch = (m_warp0 < p.M && ct > chalf) ? (ct - chalf + kCS - 1) / kCS : 0;
        const bool use_res = EPI == 4 && p.has_rmap != 0;
        if (use_res && nch > 0) {
          // every earlier store of this warp has been read out of the ring: prefetch two residual tiles
          // (bulk async-groups belong to the issuing lane; elect.sync picks the same lane every time, the
          // other lanes' waits return at once)
          bulk_wait_read<0>();
          for (int i = 0; i < 2 && i < nch; ++i) {
            const int s = (int)((epi_count + i) % 3);
            if (gemm_elect_one()) {
              mbar_arrive_expect_tx(res_bar(ew, s), kBuf);
              tma_load_5d(&tma_r, res_bar(ew, s), ring + (uint32_t)s * kBuf, RFK_COORDS(plo, phi));
            }
            advance(plo, phi);
          }
          __syncwarp();
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t myrow = (uint32_t)lane;
        // one chunk: accumulator registers -> bias / LN / activation (/ residual) -> swizzled tile -> TMA store
        auto process = [&](uint32_t (&r)[32], int i) {
          const uint32_t n0 = n_tile0 + (uint32_t)(chalf + kCS * i) * 32u;
          const int slot = (int)(epi_count % kRing);
          const uint32_t buf = ring + (uint32_t)slot * kBuf;
          float v[32];
          if (bias) {  // same addresses in every lane: broadcast loads
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + n0) + j);
              v[4 * j] = __uint_as_float(r[4 * j]) + bq.x;
              v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bq.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bq.z;
              v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (EPI == 3 && p.epi == RFK_EPI_BLOCKLN32) blockln32(p, lane, v);
          if (p.act == RFK_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if constexpr (EPI == 3) {
#ifndef RFK_GEMM_COMMIT_PER_TILE
            // the store issued four chunks ago (same slot) must have finished reading the tile
            bulk_wait_read<Cfg::kRing3 - 1>();
            __syncwarp();
#endif
            const uint32_t rowb = buf + myrow * 64u, sw = (myrow >> 1) & 3u;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4u(rowb + ((((uint32_t)q) ^ sw) << 4), pack_h16x2(v[8 * q], v[8 * q + 1], p.c_dtype),
                            pack_h16x2(v[8 * q + 2], v[8 * q + 3], p.c_dtype), pack_h16x2(v[8 * q + 4], v[8 * q + 5], p.c_dtype),
                            pack_h16x2(v[8 * q + 6], v[8 * q + 7], p.c_dtype));
          } else {
            const uint32_t rowb = buf + myrow * 128u, sw = myrow & 7u;
            if (use_res) {
              mbar_wait(res_bar(ew, slot), (uint32_t)((epi_count / 3) & 1));
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t4 = ld_shared_f4(rowb + ((((uint32_t)q) ^ sw) << 4));
                v[4 * q] += t4.x; v[4 * q + 1] += t4.y; v[4 * q + 2] += t4.z; v[4 * q + 3] += t4.w;
              }
            } else {
              bulk_wait_read<2>();
              __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st_shared_v4u(rowb + ((((uint32_t)q) ^ sw) << 4), __float_as_uint(v[4 * q]),
                            __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                            __float_as_uint(v[4 * q + 3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (gemm_elect_one()) {
            tma_store_5d(&tma_c, buf, RFK_COORDS(nlo, nhi));
#ifdef RFK_GEMM_COMMIT_PER_TILE
            if (EPI != 3) bulk_commit();  // EPI 3: one commit per tile, after the chunk loop
#else
            bulk_commit();
#endif
          }
          if (use_res && i + 2 < nch) {
            // slot of chunk i+2 was last used by chunk i-1: its store may still be in flight
            bulk_wait_read<1>();
            const int s2 = (int)((epi_count + 2) % 3);
            if (gemm_elect_one()) {
              mbar_arrive_expect_tx(res_bar(ew, s2), kBuf);
              tma_load_5d(&tma_r, res_bar(ew, s2), ring + (uint32_t)s2 * kBuf, RFK_COORDS(plo, phi));
            }
            advance(plo, phi);
          }
          __syncwarp();
          advance(nlo, nhi);
          ++epi_count;
        };
#ifdef RFK_GEMM_COMMIT_PER_TILE
        // experiment (profiles/r01_gemm_epilogue_notes.md section 4): the <= 4 stores of a tile use the 4 ring slots
        // without per-chunk waits or commits; one wait before the tile, one commit after it
        static_assert(EPI != 3 || (BN / 32 + kCS - 1) / kCS <= Cfg::kRing3, "a tile's chunks must fit the store ring");
        if (EPI == 3 && nch > 0) {
          bulk_wait_read<0>();
          __syncwarp();
        }
#endif
        if (nch > 0) {
          // the accumulator chunk after the one being processed is already on its way out of TMEM
          // (a prefetch past the last owned chunk re-reads the first one: always inside the accumulator)
          uint32_t ra[32], rb[32];
          auto chunk_addr = [&](int i) {
            const int c = chalf + kCS * i;
            return taddr + (uint32_t)((c < BN / 32 ? c : chalf) * 32);
          };
          tmem_ld_32x32(chunk_addr(0), ra);
#pragma unroll 1
          for (int i = 0; i < nch; i += 2) {
            tmem_ld_wait();
            tmem_ld_32x32(chunk_addr(i + 1), rb);
            process(ra, i);
            if (i + 1 < nch) {
              tmem_ld_wait();
              tmem_ld_32x32(chunk_addr(i + 2), ra);
              process(rb, i + 1);
            }
          }
          tmem_ld_wait();  // the trailing prefetch must land before the accumulator is handed back
#ifdef RFK_GEMM_COMMIT_PER_TILE
          if (EPI == 3 && gemm_elect_one()) bulk_commit();
          __syncwarp();
#endif
        }
#undef RFK_COORDS
      } else {
        static_assert(EPI == 0 || BN % 32 == 0, "lean epilogues need 32-column chunks");
        // rows of this warp are affine in memory: offset(row) = base + row * ms[0]
        const int64_t mw = m_warp0 < p.M ? m_warp0 : 0;
        const int64_t c_base = addr_zm(p.c_addr, z0, z1, z2, mw, p.MR);
        const int64_t r0_base = (EPI == 2 && p.r0) ? addr_zm(p.r0_addr, z0, z1, z2, mw, p.MR) : 0;
        const int64_t r1_base = (EPI == 2 && p.r1) ? addr_zm(p.r1_addr, z0, z1, z2, mw, p.MR) : 0;
        const int64_t left = CONV ? (int64_t)p.conv_L - (int64_t)((uint32_t)m_warp0 % (uint32_t)p.conv_Lp)
                                  : p.M - m_warp0;
        const int rows_valid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          const int64_t n0 = nb * BN + c * 32;
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (n0 >= p.N) continue;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (EPI == 1 && p.epi == RFK_EPI_BLOCKLN32) blockln32(p, lane, v);
          epilogue_fast_chunk<EPI>(p, reinterpret_cast<float4*>(stage), lane, rows_valid, n0, v, c_base,
                                   r0_base, r1_base, bias);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  if ((EPI == 3 || EPI == 4) && warp >= 2) bulk_wait_all();  // smem tiles must outlive their bulk stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


struct EpiMaps {
  CUtensorMap c, r;
};

template <int BN, int EPI, bool CONV = false>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles,
                     cudaStream_t stream, const EpiMaps* em = nullptr) {
  using Cfg = GemmCfg<BN, EPI>;
  static const EpiMaps kNoMaps{};
  if (!em) em = &kNoMaps;
  // the kernel decodes tiles and splits m / n in 32-bit arithmetic
  if (tiles > 0x7fffffffLL || p.M > 0x7fffffffLL || p.N > 0x7fffffffLL || p.K > 0x7fffffffLL) return RFK_ERR_BAD_DIMS;
  static PerDeviceOnce once;  // one per template instance
  const int cfg_rc = per_device_once(once, []() {
    return cuda_status(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::kSmemBytes));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  int grid = num_sms();
  if (tiles < grid) grid = (int)tiles;
  gemm_tc_kernel<BN, EPI, CONV><<<grid, gemm_threads(EPI), Cfg::kSmemBytes, stream>>>(ta, tb, em->c, em->r, p);
  return post_launch();
}

template <int EPI>
static int launch_tc_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p,
                        int64_t tiles, cudaStream_t stream, const EpiMaps* em = nullptr) {
  switch (bn) {
    case 256: return launch_tc<256, EPI>(ta, tb, p, tiles, stream, em);
    case 192: return launch_tc<192, EPI>(ta, tb, p, tiles, stream, em);
    case 128: return launch_tc<128, EPI>(ta, tb, p, tiles, stream, em);
    case 96: return launch_tc<96, EPI>(ta, tb, p, tiles, stream, em);
    case 64: return launch_tc<64, EPI>(ta, tb, p, tiles, stream, em);
    case 32: return launch_tc<32, EPI>(ta, tb, p, tiles, stream, em);
    default: return RFK_ERR_UNSUPPORTED;
  }
}

// one translation unit per epilogue flavour (parallel compilation, smaller kernels)
int launch_tc_epi0(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s);
int launch_tc_epi1(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s);
int launch_tc_epi2(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s);
int launch_tc_epi3(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s, const EpiMaps* em);
int launch_tc_epi4(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s, const EpiMaps* em);
int launch_tc_conv(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int64_t tiles, cudaStream_t s);
int make_tmap_bf16_raw(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace rfk
