// rfk_gemm.cu — batched TN GEMM with a generalised scatter epilogue.
//   mode 0 (bf16 operands): persistent, warp-specialised tcgen05 kernel. One warp drives TMA
//     (128B-swizzled K-major tiles, 5-D tensor maps so batch levels need no pointer math), one
//     thread issues tcgen05.mma into a double-buffered TMEM accumulator, four warps drain TMEM
//     with tcgen05.ld and run the epilogue while the next tile's MMAs are already in flight.
//   mode 1 (fp32 operands): plain SIMT fp32 kernel (validation mode, same epilogue code).
// See include/rfk.h for the contract and the reference lines this replaces.
#include <cstdlib>

#include "rfk_gemm_device.cuh"

namespace rfk {

// ----------------------------------------------------------------------------------------------
// SIMT fp32 kernel (validation mode): 64x64 tile, 16x16 threads, 4x4 micro-tile, BK = 16.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmDev p) {
  __shared__ float sa[16][65];
  __shared__ float sb[16][65];
  const int64_t m_blocks = (p.M + 63) / 64;
  const int64_t n_blocks = (p.N + 63) / 64;
  int64_t t = blockIdx.x;
  const int64_t nb = t % n_blocks;
  const int64_t mb = (t / n_blocks) % m_blocks;
  const int64_t z = t / (n_blocks * m_blocks);
  const int64_t z0 = z % p.Z0, z1 = (z / p.Z0) % p.Z1, z2 = z / (p.Z0 * p.Z1);
  const float* A = p.a32 + z0 * p.a_zs[0] + z1 * p.a_zs[1] + z2 * p.a_zs[2];
  const float* B = p.b32 + z0 * p.b_zs[0] + z1 * p.b_zs[1] + z2 * p.b_zs[2];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = 0; k0 < p.K; k0 += 16) {
    // each thread loads 4 elements of A and of B: row = tid/4 (+0), kk = (tid%4)*4..+3
    {
      const int r = threadIdx.x >> 2;          // 0..63
      const int kk = (threadIdx.x & 3) * 4;    // 0,4,8,12
      const int64_t gm = mb * 64 + r, gn = nb * 64 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t gk = k0 + kk + i;
        sa[kk + i][r] = (gm < p.M && gk < p.K) ? A[gm * p.lda + gk] : 0.f;
        sb[kk + i][r] = (gn < p.N && gk < p.K) ? B[gn * p.ldb + gk] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
    epilogue_row_chunk<4>(p, z0, z1, z2, mb * 64 + ty * 4 + i, nb * 64 + tx * 4, v);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 5-D bf16 tensor map over [z2][z1][z0][rows][K] with a {64, box_rows, 1, 1, 1} box, 128B swizzle.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, int64_t K, int64_t rows, int64_t ld,
                   const int64_t Z[3], const int64_t zs[3], int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return RFK_ERR_TMA_ENCODE;
  if (!aligned16(ptr) || (ld % 8) != 0) return RFK_ERR_MISALIGNED;
  cuuint64_t dims[5] = {(cuuint64_t)K, (cuuint64_t)rows, 1, 1, 1};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows, 0, 0};
  for (int i = 0; i < 3; ++i) {
    if (zs[i] != 0) {
      if (zs[i] % 8 != 0) return RFK_ERR_MISALIGNED;
      dims[2 + i] = (cuuint64_t)Z[i];
      strides[1 + i] = (cuuint64_t)zs[i] * 2;
    } else {
      dims[2 + i] = 1;
      strides[1 + i] = (i == 0) ? strides[0] * (cuuint64_t)rows : strides[i];
    }
    if (strides[1 + i] == 0) strides[1 + i] = 16;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RFK_OK : RFK_ERR_TMA_ENCODE;
}

static int make_tmap_raw(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapDataType dt,
                        CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return RFK_ERR_TMA_ENCODE;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16) return RFK_ERR_MISALIGNED;
    st[i] = strides_bytes[i];
  }
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RFK_OK : RFK_ERR_TMA_ENCODE;
}

// Tensor map of an epilogue operand described by an rfk_addr (output or residual): logical dims
// {n%NR, n/NR, m%MR, m/MR, z0, z1, z2} with size-1 dims dropped; box = 32 x 32 over (n%NR, m%MR).
// Returns RFK_ERR_UNSUPPORTED when the view needs more than 5 dims or breaks a TMA rule.
int make_epi_tmap(CUtensorMap* map, int cmap[5], const void* ptr, int dtype, const rfk_addr& a,
                  const int64_t ext[7]) {
  const int es = dtype == RFK_F32 ? 4 : 2;
  if (!aligned16(ptr) || a.ns[0] != 1) return RFK_ERR_UNSUPPORTED;
  const int64_t strides[7] = {a.ns[0], a.ns[1], a.ms[0], a.ms[1], a.zs[0], a.zs[1], a.zs[2]};
  uint64_t dims[5] = {1, 1, 1, 1, 1}, sb[4] = {16, 16, 16, 16};
  uint32_t box[5] = {1, 1, 1, 1, 1};
  int nd = 0;
  for (int l = 0; l < 7; ++l) {
    const bool keep = l == 0 || l == 2 || ext[l] > 1;
    if (!keep) continue;
    if (nd == 5) return RFK_ERR_UNSUPPORTED;
    if (l != 0) {
      const int64_t st = ext[l] > 1 ? strides[l] : (int64_t)(16 / es);
      if (st <= 0 || (st * es) % 16) return RFK_ERR_UNSUPPORTED;  // no broadcast / misaligned dims
      sb[nd - 1] = (uint64_t)st * es;
    }
    dims[nd] = (uint64_t)ext[l];
    box[nd] = (l == 0 || l == 2) ? 32 : 1;
    cmap[nd] = l;
    ++nd;
  }
  for (int d = nd; d < 5; ++d) cmap[d] = -1;
  return make_tmap_raw(map, ptr, 5, dims, sb, box,
                       dtype == RFK_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,  // (f16: same 2-byte moves)
                       dtype == RFK_F32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

int make_tmap_bf16_raw(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return RFK_ERR_TMA_ENCODE;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    if (strides_bytes[i] % 16) return RFK_ERR_MISALIGNED;
    st[i] = strides_bytes[i];
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, bx,
                   es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RFK_OK : RFK_ERR_TMA_ENCODE;
}

static int pick_bn(int64_t N) {
  // smallest tile-count first, then the least padding; N up to a few thousand
  const int cands[] = {256, 192, 128, 96, 64, 32};
  int best = 32;
  double best_cost = 1e30;
  for (int bn : cands) {
    const int64_t nb = (N + bn - 1) / bn;
    // MMA time ~ nb * max(bn, 96): small tiles pay the A-operand re-read
    double cost = (double)nb * (bn < 128 ? (bn * 0.35 + 83.0) : (double)bn);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace rfk

using namespace rfk;

extern "C" int rfk_gemm(const rfk_gemm_desc* d, rfk_stream_t stream_) {
  if (!d) return RFK_ERR_NULL_POINTER;
  if (!d->a || !d->b || !d->c) return RFK_ERR_NULL_POINTER;
  if (d->M <= 0 || d->N <= 0 || d->K <= 0) return RFK_ERR_BAD_DIMS;
  if (d->M > 0x7fffffffLL || d->N > 0x7fffffffLL || d->K > 0x7fffffffLL) return RFK_ERR_BAD_DIMS;
  for (int i = 0; i < 3; ++i)
    if (d->Z[i] <= 0) return RFK_ERR_BAD_DIMS;
  if (d->MR <= 0 || d->NR <= 0) return RFK_ERR_BAD_DIMS;
  if (d->epi == RFK_EPI_BLOCKLN32 && (d->MR != 32 || d->NR != 32 || d->M % 32 || d->N % 32))
    return RFK_ERR_BAD_DIMS;
  if (d->epi == RFK_EPI_BLOCKLN32 && ((d->ln_gamma == nullptr) != (d->ln_beta == nullptr)))
    return RFK_ERR_NULL_POINTER;
  if (d->epi == RFK_EPI_BLOCKLN32 &&
      (d->alpha != 1.f || d->bias || d->r0 || d->r1 || d->act != RFK_ACT_NONE))
    return RFK_ERR_UNSUPPORTED;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);

  GemmDev p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.Z0 = d->Z[0]; p.Z1 = d->Z[1]; p.Z2 = d->Z[2];
  p.MR = d->MR; p.NR = d->NR;
  p.alpha = d->alpha; p.act = d->act; p.epi = d->epi; p.ln_eps = d->ln_eps;
  p.bias = d->bias;
  for (int i = 0; i < 3; ++i) p.bias_zs[i] = d->bias_zs[i];
  p.c = d->c; p.r0 = d->r0; p.r1 = d->r1;
  p.c_dtype = d->c_dtype; p.r0_dtype = d->r0_dtype; p.r1_dtype = d->r1_dtype;
  p.c_addr = d->c_addr; p.r0_addr = d->r0_addr; p.r1_addr = d->r1_addr;
  p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta;
  const int64_t Z = d->Z[0] * d->Z[1] * d->Z[2];

  if (d->ab_dtype == RFK_F32) {
    if (d->epi != RFK_EPI_STD) return RFK_ERR_UNSUPPORTED;
    p.a32 = reinterpret_cast<const float*>(d->a);
    p.b32 = reinterpret_cast<const float*>(d->b);
    p.lda = d->lda; p.ldb = d->ldb;
    for (int i = 0; i < 3; ++i) { p.a_zs[i] = d->a_zs[i]; p.b_zs[i] = d->b_zs[i]; }
    const int64_t tiles = Z * ((d->M + 63) / 64) * ((d->N + 63) / 64);
    if (tiles > 0x7fffffffLL) return RFK_ERR_BAD_DIMS;
    gemm_f32_kernel<<<(unsigned)tiles, 256, 0, stream>>>(p);
    return post_launch();
  }
  if (!is_h16(d->ab_dtype)) return RFK_ERR_BAD_DTYPE;
  p.ab_f16 = d->ab_dtype == RFK_F16;
  int rc = check_arch();
  if (rc != RFK_OK) return rc;

  int bn = pick_bn(d->N);
  static const char* force_bn = getenv("RFK_GEMM_BN");  // A/B debugging aid: one of 256/192/128/96/64/32
  if (force_bn && atoi(force_bn) >= 32 && atoi(force_bn) % 32 == 0 && atoi(force_bn) <= 256) bn = atoi(force_bn);
  if (d->epi == RFK_EPI_BLOCKLN32) bn = 128;

  CUtensorMap ta, tb;
  rc = make_tmap_bf16(&ta, d->a, d->K, d->M, d->lda, d->Z, d->a_zs, kBlockM);
  if (rc != RFK_OK) return rc;
  rc = make_tmap_bf16(&tb, d->b, d->K, d->N, d->ldb, d->Z, d->b_zs, bn);
  if (rc != RFK_OK) return rc;
  for (int i = 0; i < 3; ++i) p.bmask[i] = d->b_zs[i] != 0 ? -1 : 0;
  // A broadcast over z is expressed the same way: a size-1 dim would need a mask too; require
  // a_zs != 0 whenever Z[i] > 1.
  for (int i = 0; i < 3; ++i)
    if (d->Z[i] > 1 && d->a_zs[i] == 0) return RFK_ERR_UNSUPPORTED;
  const int64_t tiles = Z * ((d->M + kBlockM - 1) / kBlockM) * ((d->N + bn - 1) / bn);
  // ---- pick the epilogue flavour (rfk_gemm_device.cuh) ----
  auto addr_aligned = [](const void* ptr, const rfk_addr& a, int es) {
    const int64_t q = 16 / es;  // elements per 16 bytes
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return false;
    for (int i = 0; i < 3; ++i)
      if (a.zs[i] % q) return false;
    return a.ms[0] % q == 0 && a.ms[1] % q == 0 && a.ns[1] % q == 0 && a.ns[0] == 1;
  };
  const bool rows_affine = (d->MR % 32 == 0) || d->MR >= d->M;
  const bool cols_chunked = (d->NR % 32 == 0) || d->NR >= d->N;
  bool lean = d->alpha == 1.f && d->N % 32 == 0 && rows_affine && cols_chunked &&
              (d->act == RFK_ACT_NONE || d->act == RFK_ACT_RELU) &&
              (!d->bias || aligned16(d->bias)) && d->ln_gamma == nullptr;
  for (int i = 0; i < 3; ++i) lean = lean && (d->bias_zs[i] % 4 == 0);
  int epi = 0;
  if (lean && is_h16(d->c_dtype) && !d->r0 && !d->r1 && addr_aligned(d->c, d->c_addr, 2)) epi = 1;
  if (lean && d->c_dtype == RFK_F32 && d->epi == RFK_EPI_STD && addr_aligned(d->c, d->c_addr, 4) &&
      (!d->r0 || (d->r0_dtype == RFK_F32 && addr_aligned(d->r0, d->r0_addr, 4))) &&
      (!d->r1 || (d->r1_dtype == RFK_F32 && addr_aligned(d->r1, d->r1_addr, 4))))
    epi = 2;
  // TMA-store flavours (3: bf16, 4: f32 + optional TMA-fetched residual) when the views fit a tensor map
  static const bool no_tma_epi = getenv("RFK_GEMM_NO_TMA_EPILOGUE") != nullptr;  // A/B debugging aid
  // flavour 4 pays for its 96 KB residual/result ring with pipeline stages: only worth it for the
  // short-K GEMMs that are bound by the fp32 residual stream, not by the tensor pipe
  static const bool epi4_any_k = getenv("RFK_GEMM_EPI4_ANYK") != nullptr;              // A/B debugging aid
  if ((epi == 1 || (epi == 2 && d->r0 && !d->r1 && (d->K <= 768 || epi4_any_k))) && !no_tma_epi) {
    const bool n_split = d->NR < d->N, m_split = d->MR < d->M;
    if ((!n_split || d->N % d->NR == 0) && (!m_split || d->M % d->MR == 0)) {
      const int64_t ext[7] = {n_split ? d->NR : d->N, n_split ? d->N / d->NR : 1,
                              m_split ? d->MR : d->M, m_split ? d->M / d->MR : 1,
                              d->Z[0], d->Z[1], d->Z[2]};
      EpiMaps em{};
      int cmap_c[5], cmap_r[5];
      bool ok = make_epi_tmap(&em.c, cmap_c, d->c, d->c_dtype, d->c_addr, ext) == RFK_OK;
      bool res_tma = false;
      if (ok && epi == 2 && d->r0) {
        res_tma = make_epi_tmap(&em.r, cmap_r, d->r0, RFK_F32, d->r0_addr, ext) == RFK_OK;
        for (int i = 0; i < 5 && res_tma; ++i) res_tma = cmap_r[i] == cmap_c[i];
        ok = res_tma;  // a residual that cannot be fetched by TMA (broadcast addends) keeps flavour 2
      }
      if (ok) {
        for (int i = 0; i < 5; ++i) p.cmap[i] = cmap_c[i];
        p.has_rmap = res_tma ? 1 : 0;
        return epi == 1 ? launch_tc_epi3(bn, ta, tb, p, tiles, stream, &em)
                        : launch_tc_epi4(bn, ta, tb, p, tiles, stream, &em);
      }
    }
  }
  if (epi == 1) return launch_tc_epi1(bn, ta, tb, p, tiles, stream);
  if (epi == 2) return launch_tc_epi2(bn, ta, tb, p, tiles, stream);
  return launch_tc_epi0(bn, ta, tb, p, tiles, stream);
}
