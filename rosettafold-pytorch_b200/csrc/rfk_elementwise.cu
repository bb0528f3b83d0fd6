// rfk_elementwise.cu — the HBM-bound kernels of the trunk: LayerNorm, row softmax, tied-attention
// symmetrisation, position-wise weight factor, OPM operand preparation, pair2att logits,
// InstanceNorm statistics/apply and dtype conversion. All are coalesced, 128-bit vectorised where
// the layout allows and reduce with warp shuffles. See include/rfk.h for the contracts.
#include "rfk_common.cuh"

namespace rfk {

// ----------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (VPL 4-element vectors per lane).
// ----------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&o)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<__half>(__half* p, const float (&v)[4]) {
  uint2 t;
  t.x = pack_h16x2(v[0], v[1], RFK_F16);
  t.y = pack_h16x2(v[2], v[3], RFK_F16);
  *reinterpret_cast<uint2*>(p) = t;
}
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  t.x = pack_bf16x2(v[0], v[1]);
  t.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

template <typename TI, typename TO, int VPL>
__global__ void __launch_bounds__(256)
layernorm_vec_kernel(const TI* __restrict__ x, int64_t xs, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, TO* __restrict__ y, int64_t ys,
                     int64_t rows, int D, const float* __restrict__ res = nullptr, int64_t rs = 0) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* xr = x + row * xs;
  float v[VPL][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      load4<TI>(xr + c, v[i]);
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    } else {
      v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = v[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  TO* yr = y + row * ys;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < D) {
      float o[4];
      if (gamma) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
        o[0] = (v[i][0] - mean) * rstd * g.x + b.x;
        o[1] = (v[i][1] - mean) * rstd * g.y + b.y;
        o[2] = (v[i][2] - mean) * rstd * g.z + b.z;
        o[3] = (v[i][3] - mean) * rstd * g.w + b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd;
      }
      if (res) {  // y = LN(x) + res (f32 addend, rfk_layernorm_residual)
        const float4 t = __ldg(reinterpret_cast<const float4*>(res + row * rs + c));
        o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
      }
      store4<TO>(yr + c, o);
    }
  }
}

// generic fallback (any D / alignment): one warp per row, three passes over the row
__global__ void __launch_bounds__(256)
layernorm_generic_kernel(const void* __restrict__ x, int xdt, int64_t xs,
                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         void* __restrict__ y, int ydt, int64_t ys, int64_t rows, int D,
                         const float* __restrict__ res = nullptr, int64_t rs = 0) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += load_as_float(x, xdt, row * xs + c);
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = load_as_float(x, xdt, row * xs + c) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int c = lane; c < D; c += 32) {
    float o = (load_as_float(x, xdt, row * xs + c) - mean) * rstd;
    if (gamma) o = o * gamma[c] + beta[c];
    if (res) o += res[row * rs + c];
    store_from_float(y, ydt, row * ys + c, o);
  }
}

template <typename TI, typename TO>
static int launch_ln_vec(const void* x, int64_t xs, const float* g, const float* b, float eps,
                         void* y, int64_t ys, int64_t rows, int D, cudaStream_t st,
                         const float* res = nullptr, int64_t rs = 0) {
  const int vecs = (D / 4 + 31) / 32;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  const TI* xi = reinterpret_cast<const TI*>(x);
  TO* yo = reinterpret_cast<TO*>(y);
  switch (vecs) {
    case 1: layernorm_vec_kernel<TI, TO, 1><<<blocks, 256, 0, st>>>(xi, xs, g, b, eps, yo, ys, rows, D, res, rs); break;
    case 2: layernorm_vec_kernel<TI, TO, 2><<<blocks, 256, 0, st>>>(xi, xs, g, b, eps, yo, ys, rows, D, res, rs); break;
    case 3: layernorm_vec_kernel<TI, TO, 3><<<blocks, 256, 0, st>>>(xi, xs, g, b, eps, yo, ys, rows, D, res, rs); break;
    case 4: layernorm_vec_kernel<TI, TO, 4><<<blocks, 256, 0, st>>>(xi, xs, g, b, eps, yo, ys, rows, D, res, rs); break;
    case 5: case 6: case 7: case 8:
      layernorm_vec_kernel<TI, TO, 8><<<blocks, 256, 0, st>>>(xi, xs, g, b, eps, yo, ys, rows, D, res, rs); break;
    default: return RFK_ERR_UNSUPPORTED;
  }
  return post_launch();
}

// ----------------------------------------------------------------------------------------------
// row softmax (fp32 in): one warp per row
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, int64_t xs, void* __restrict__ y, int ydt,
                    int64_t ys, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * xs;
  float mx = -INFINITY;
  for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __expf(xr[c] - mx);
  const float inv = 1.f / warp_sum(s);
  for (int c = lane; c < cols; c += 32)
    store_from_float(y, ydt, row * ys + c, __expf(xr[c] - mx) * inv);
}

// ----------------------------------------------------------------------------------------------
// att[b,i,j,h] = 0.5 * (A[b,h,i,j] + A[b,h,j,i])
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tied_att_sym_kernel(const void* __restrict__ A, int adt, int64_t lda, float* __restrict__ att,
                    void* __restrict__ att16, int64_t att16_stride, int B, int H, int L) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * L * L;
  if (idx >= total) return;
  const int j = (int)(idx % L);
  const int i = (int)((idx / L) % L);
  const int b = (int)(idx / ((int64_t)L * L));
  for (int h = 0; h < H; ++h) {
    const int64_t base = ((int64_t)b * H + h) * L;
    const float a = load_as_float(A, adt, (base + i) * lda + j);
    const float t = load_as_float(A, adt, (base + j) * lda + i);
    const float v = 0.5f * (a + t);
    att[idx * H + h] = v;
    if (att16) reinterpret_cast<__nv_bfloat16*>(att16)[idx * att16_stride + h] = __float2bfloat16_rn(v);
  }
}

// ----------------------------------------------------------------------------------------------
// PositionWiseWeightFactor: one warp per (b, l, h); logits over n staged in shared memory.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
poswise_kernel(const void* __restrict__ pq, int64_t pqs, const void* __restrict__ pk, int64_t pks,
               int dt, float scale, float* __restrict__ w_out, const void* __restrict__ q,
               int64_t qs, float q_scale, void* __restrict__ qt, int qtdt, int B, int N, int L,
               int H, int dh, float* __restrict__ stats) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 4 + warp;
  if (item >= (int64_t)B * L * H) return;
  const int h = (int)(item % H);
  const int l = (int)((item / H) % L);
  const int b = (int)(item / ((int64_t)H * L));
  float* lg = sm + (size_t)warp * N;
  const int64_t pq_off = ((int64_t)b * L + l) * pqs + (int64_t)h * dh;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) {
    const int64_t pk_off = (((int64_t)b * N + n) * L + l) * pks + (int64_t)h * dh;
    float acc = 0.f;
    for (int d = 0; d < dh; ++d)
      acc = fmaf(load_as_float(pq, dt, pq_off + d), load_as_float(pk, dt, pk_off + d), acc);
    acc *= scale;
    lg[n] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float s = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float e = __expf(lg[n] - mx);
    lg[n] = e;
    s += e;
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
  if (stats && lane == 0) {  // (max, sum of exp) of this sequence shard: lets the caller merge softmaxes across shards
    stats[(((int64_t)b * L + l) * H + h) * 2 + 0] = mx;
    stats[(((int64_t)b * L + l) * H + h) * 2 + 1] = s;
  }
  __syncwarp();
  if (w_out)
    for (int n = lane; n < N; n += 32)
      w_out[(((int64_t)b * N + n) * L + l) * H + h] = lg[n] * inv;
  if (qt) {
    const int64_t out_base = (((int64_t)b * H + h) * L + l) * ((int64_t)N * dh);
    for (int n = 0; n < N; ++n) {
      const float wn = lg[n] * inv * q_scale;
      const int64_t q_off = (((int64_t)b * N + n) * L + l) * qs + (int64_t)h * dh;
      for (int d = lane; d < dh; d += 32)
        store_from_float(qt, qtdt, out_base + (int64_t)n * dh + d,
                         load_as_float(q, dt, q_off + d) * wn);
    }
  }
}

// Vectorised PositionWiseWeightFactor (bf16 operands, dh % 8 == 0): one block per (b, l). A thread
// owns one 16-byte chunk (8 channels) of a row; `rows` MSA rows are processed per pass, so a warp
// reads whole 128-byte lines of pk / q and writes whole lines of qt. Phase 1: logits[n][h] into
// shared memory (the dh/8 threads of a head meet by shuffle); phase 2: softmax over n per head;
// phase 3: qt[b,h,l,n,:] = q[b,n,l,h,:] * w[n,h] * q_scale.
// F16: the 16-bit operands are IEEE half instead of bf16
template <bool F16>
__device__ __forceinline__ void ld8_h16(const uint16_t* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (F16) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    } else {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
template <bool F16>
__global__ void __launch_bounds__(256)
poswise_vec_kernel(const uint16_t* __restrict__ pq, int64_t pqs, const uint16_t* __restrict__ pk,
                   int64_t pks, float scale, float* __restrict__ w_out, const uint16_t* __restrict__ q,
                   int64_t qs, float q_scale, uint16_t* __restrict__ qt, int N, int L, int H, int dh,
                   int rows, float* __restrict__ stats) {
  constexpr int kDt = F16 ? RFK_F16 : RFK_BF16;
  extern __shared__ float lg[];  // [N][H]
  const int CH = (H * dh) >> 3, gsz = dh >> 3;
  const int c = threadIdx.x % CH, r = threadIdx.x / CH;
  const int h = c / gsz;
  const int l = blockIdx.x % L, b = blockIdx.x / L;
  float pqv[8];
  ld8_h16<F16>(pq + ((int64_t)b * L + l) * pqs + c * 8, pqv);
  for (int n0 = 0; n0 < N; n0 += rows) {
    const int n = n0 + r;
    float acc = 0.f;
    if (n < N) {
      float kv[8];
      ld8_h16<F16>(pk + (((int64_t)b * N + n) * L + l) * pks + c * 8, kv);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(pqv[i], kv[i], acc);
    }
    for (int o = gsz >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (n < N && (c % gsz) == 0) lg[n * H + h] = acc * scale;
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int hh = warp; hh < H; hh += nw) {
      float mx = -INFINITY;
      for (int n = lane; n < N; n += 32) mx = fmaxf(mx, lg[n * H + hh]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int n = lane; n < N; n += 32) {
        const float e = __expf(lg[n * H + hh] - mx);
        lg[n * H + hh] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      if (stats && lane == 0) {
        stats[(((int64_t)b * L + l) * H + hh) * 2 + 0] = mx;
        stats[(((int64_t)b * L + l) * H + hh) * 2 + 1] = sum;
      }
      for (int n = lane; n < N; n += 32) {
        const float w = lg[n * H + hh] * inv;
        lg[n * H + hh] = w;
        if (w_out) w_out[(((int64_t)b * N + n) * L + l) * H + hh] = w;
      }
    }
  }
  if (!qt) return;
  __syncthreads();
  const int64_t out_base = (((int64_t)b * H + h) * L + l) * ((int64_t)N * dh) + (c % gsz) * 8;
  for (int n = r; n < N; n += rows) {
    float qv[8];
    ld8_h16<F16>(q + (((int64_t)b * N + n) * L + l) * qs + c * 8, qv);
    const float wn = lg[n * H + h] * q_scale;
    uint4 u;
    u.x = pack_h16x2(qv[0] * wn, qv[1] * wn, kDt); u.y = pack_h16x2(qv[2] * wn, qv[3] * wn, kDt);
    u.z = pack_h16x2(qv[4] * wn, qv[5] * wn, kDt); u.w = pack_h16x2(qv[6] * wn, qv[7] * wn, kDt);
    *reinterpret_cast<uint4*>(qt + out_base + (int64_t)n * dh) = u;
  }
}

// ----------------------------------------------------------------------------------------------
// OPM operand preparation: block (32 x 8) per (b, l); transposes [n][u] -> [u][n] through smem.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
opm_prep_kernel(const float* __restrict__ m, const float* __restrict__ w, void* __restrict__ xt,
                void* __restrict__ yt, int tdt, int64_t ldt, float* __restrict__ msa1d, int B,
                int N, int L, int P) {
  __shared__ float tile[32][33];
  __shared__ float wts[32];
  __shared__ float part[8][33];
  const int l = blockIdx.x % L, b = blockIdx.x / L;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int u0 = 0; u0 < P; u0 += 32) {
    float colsum = 0.f;
    for (int n0 = 0; n0 < N; n0 += 32) {
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, u = u0 + tx;
        float val = 0.f;
        if (n < N && u < P) val = m[(((int64_t)b * N + n) * L + l) * P + u];
        tile[r][tx] = val;
        colsum += val;
      }
      if (ty == 0) wts[tx] = (n0 + tx < N) ? w[((int64_t)b * N + n0 + tx) * L + l] : 0.f;
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int u = u0 + r, n = n0 + tx;
        if (u < P && n < N) {
          const int64_t o = ((int64_t)b * L * P + (int64_t)l * P + u) * ldt + n;
          const float val = tile[tx][r];
          store_from_float(xt, tdt, o, val);
          store_from_float(yt, tdt, o, val * wts[tx]);
        }
      }
    }
    __syncthreads();
    part[ty][tx] = colsum;
    __syncthreads();
    if (ty == 0 && u0 + tx < P) {
      float s = 0.f;
      for (int r = 0; r < 8; ++r) s += part[r][tx];
      const int64_t o = ((int64_t)b * L + l) * (2 * P);
      msa1d[o + u0 + tx] = s;
      msa1d[o + P + u0 + tx] = m[(((int64_t)b * N + 0) * L + l) * P + u0 + tx];
    }
  }
}

// ----------------------------------------------------------------------------------------------
// pair2att logits: one warp per (b, i<=j) pair of positions; weights in shared memory.
// ----------------------------------------------------------------------------------------------
constexpr int kP2AMaxC = 32;
__global__ void __launch_bounds__(256)
pair2att_kernel(const float* __restrict__ pair, const float* __restrict__ Wf,
                const float* __restrict__ bf, float eps, float* __restrict__ logits, int64_t ldl,
                int B, int L, int D, int C) {
  extern __shared__ float sw[];  // [C][D]
  for (int i = threadIdx.x; i < C * D; i += blockDim.x) sw[i] = Wf[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (int64_t)B * L * L) return;
  const int j = (int)(item % L);
  const int i = (int)((item / L) % L);
  const int b = (int)(item / ((int64_t)L * L));
  if (i > j) return;
  const float* pij = pair + (((int64_t)b * L + i) * L + j) * D;
  const float* pji = pair + (((int64_t)b * L + j) * L + i) * D;
  // D <= 32 * 16 supported
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int d = t * 32 + lane;
    v[t] = d < D ? 0.5f * (pij[d] + pji[d]) : 0.f;
    s += v[t];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int d = t * 32 + lane;
    const float dd = d < D ? v[t] - mean : 0.f;
    v[t] = dd;
    q += dd * dd;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int d = t * 32 + lane;
      if (d < D) acc = fmaf(v[t], sw[c * D + d], acc);
    }
    acc = warp_sum(acc) * rstd + bf[c];
    if (lane == 0) {
      const int64_t base = ((int64_t)b * C + c) * L;
      logits[(base + i) * ldl + j] = acc;
      logits[(base + j) * ldl + i] = acc;
    }
  }
}

// D = 32 NT: eight lanes per (i <= j) pair of positions, 16-byte loads, all 2 NT loads of a lane in flight
// before the first use; a warp covers four consecutive j, a block 32; blocks below the diagonal exit.
template <int NT>
__global__ void __launch_bounds__(256)
pair2att_vec_kernel(const float* __restrict__ pair, const float* __restrict__ Wf,
                    const float* __restrict__ bf, float eps, float* __restrict__ logits, int64_t ldl,
                    int L, int C) {
  constexpr int D = NT * 32;
  extern __shared__ float sw[];  // [C][D]
  const int i = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * 32;
  if (j0 + 31 < i) return;
  for (int t = threadIdx.x; t < C * (D / 4); t += blockDim.x)
    reinterpret_cast<float4*>(sw)[t] = __ldg(reinterpret_cast<const float4*>(Wf) + t);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane & 7;
  const int j = j0 + warp * 4 + (lane >> 3);
  const bool active = j < L && j >= i;
  const int jj = active ? j : i;  // idle groups shadow the diagonal element (they take part in the shuffles)
  const float4* pij = reinterpret_cast<const float4*>(pair + (((int64_t)b * L + i) * L + jj) * D) + l8;
  const float4* pji = reinterpret_cast<const float4*>(pair + (((int64_t)b * L + jj) * L + i) * D) + l8;
  float4 u[NT], w[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) { u[t] = __ldg(pij + 8 * t); w[t] = __ldg(pji + 8 * t); }
  auto sum8 = [](float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x;
  };
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    u[t].x = 0.5f * (u[t].x + w[t].x); u[t].y = 0.5f * (u[t].y + w[t].y);
    u[t].z = 0.5f * (u[t].z + w[t].z); u[t].w = 0.5f * (u[t].w + w[t].w);
    s += (u[t].x + u[t].y) + (u[t].z + u[t].w);
  }
  const float mean = sum8(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    u[t].x -= mean; u[t].y -= mean; u[t].z -= mean; u[t].w -= mean;
    q = fmaf(u[t].x, u[t].x, q); q = fmaf(u[t].y, u[t].y, q);
    q = fmaf(u[t].z, u[t].z, q); q = fmaf(u[t].w, u[t].w, q);
  }
  const float rstd = rsqrtf(sum8(q) / (float)D + eps);
  for (int c = 0; c < C; ++c) {
    const float4* wc = reinterpret_cast<const float4*>(sw + c * D) + l8;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const float4 ww = wc[8 * t];
      acc = fmaf(u[t].x, ww.x, acc); acc = fmaf(u[t].y, ww.y, acc);
      acc = fmaf(u[t].z, ww.z, acc); acc = fmaf(u[t].w, ww.w, acc);
    }
    acc = sum8(acc) * rstd + bf[c];
    if (active && l8 == 0) {
      const int64_t base = ((int64_t)b * C + c) * L;
      logits[(base + i) * ldl + j] = acc;
      logits[(base + j) * ldl + i] = acc;
    }
  }
}

// Row-sharded pair map (one long protein over several GPUs): this device holds rows [i0, i0 + Li) of the pair
// map, `rows[b, il, j, :] = pair[b, i0 + il, j, :]`, and the transposed shard it received by all-to-all,
// `cols_t[b, j, il, :] = pair[b, j, i0 + il, :]`; it produces logits[b, c, il, j] for its rows and every j.
// Generic D (<= 512): one warp per (il, j).
__global__ void __launch_bounds__(256)
pair2att_rows_kernel(const float* __restrict__ rows, const float* __restrict__ cols_t,
                     const float* __restrict__ Wf, const float* __restrict__ bf, float eps,
                     float* __restrict__ logits, int64_t ldl, int B, int Li, int L, int D, int C) {
  extern __shared__ float sw[];  // [C][D]
  for (int i = threadIdx.x; i < C * D; i += blockDim.x) sw[i] = Wf[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= (int64_t)B * Li * L) return;
  const int j = (int)(item % L);
  const int il = (int)((item / L) % Li);
  const int b = (int)(item / ((int64_t)L * Li));
  const float* pij = rows + (((int64_t)b * Li + il) * L + j) * D;
  const float* pji = cols_t + (((int64_t)b * L + j) * Li + il) * D;
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int d = t * 32 + lane;
    v[t] = d < D ? 0.5f * (pij[d] + pji[d]) : 0.f;
    s += v[t];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int d = t * 32 + lane;
    const float dd = d < D ? v[t] - mean : 0.f;
    v[t] = dd;
    q += dd * dd;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int d = t * 32 + lane;
      if (d < D) acc = fmaf(v[t], sw[c * D + d], acc);
    }
    acc = warp_sum(acc) * rstd + bf[c];
    if (lane == 0) logits[(((int64_t)b * C + c) * Li + il) * ldl + j] = acc;
  }
}

// D = 32 NT: eight lanes per (il, j), 16-byte loads (as pair2att_vec_kernel, without the triangle)
template <int NT>
__global__ void __launch_bounds__(256)
pair2att_rows_vec_kernel(const float* __restrict__ rows, const float* __restrict__ cols_t,
                         const float* __restrict__ Wf, const float* __restrict__ bf, float eps,
                         float* __restrict__ logits, int64_t ldl, int Li, int L, int C) {
  constexpr int D = NT * 32;
  extern __shared__ float sw[];  // [C][D]
  const int il = blockIdx.y, b = blockIdx.z;
  const int j0 = blockIdx.x * 32;
  for (int t = threadIdx.x; t < C * (D / 4); t += blockDim.x)
    reinterpret_cast<float4*>(sw)[t] = __ldg(reinterpret_cast<const float4*>(Wf) + t);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane & 7;
  const int j = j0 + warp * 4 + (lane >> 3);
  const bool active = j < L;
  const int jj = active ? j : L - 1;
  const float4* pij = reinterpret_cast<const float4*>(rows + (((int64_t)b * Li + il) * L + jj) * D) + l8;
  const float4* pji = reinterpret_cast<const float4*>(cols_t + (((int64_t)b * L + jj) * Li + il) * D) + l8;
  float4 u[NT], w[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) { u[t] = __ldg(pij + 8 * t); w[t] = __ldg(pji + 8 * t); }
  auto sum8 = [](float x) {
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    return x;
  };
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    u[t].x = 0.5f * (u[t].x + w[t].x); u[t].y = 0.5f * (u[t].y + w[t].y);
    u[t].z = 0.5f * (u[t].z + w[t].z); u[t].w = 0.5f * (u[t].w + w[t].w);
    s += (u[t].x + u[t].y) + (u[t].z + u[t].w);
  }
  const float mean = sum8(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    u[t].x -= mean; u[t].y -= mean; u[t].z -= mean; u[t].w -= mean;
    q = fmaf(u[t].x, u[t].x, q); q = fmaf(u[t].y, u[t].y, q);
    q = fmaf(u[t].z, u[t].z, q); q = fmaf(u[t].w, u[t].w, q);
  }
  const float rstd = rsqrtf(sum8(q) / (float)D + eps);
  for (int c = 0; c < C; ++c) {
    const float4* wc = reinterpret_cast<const float4*>(sw + c * D) + l8;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const float4 ww = wc[8 * t];
      acc = fmaf(u[t].x, ww.x, acc); acc = fmaf(u[t].y, ww.y, acc);
      acc = fmaf(u[t].z, ww.z, acc); acc = fmaf(u[t].w, ww.w, acc);
    }
    acc = sum8(acc) * rstd + bf[c];
    if (active && l8 == 0) logits[(((int64_t)b * C + c) * Li + il) * ldl + j] = acc;
  }
}

// ----------------------------------------------------------------------------------------------
// InstanceNorm statistics and apply (channels-last)
// ----------------------------------------------------------------------------------------------
// Partial sums over `chunk` positions are formed in a fixed order in fp32; the cross-block
// accumulation uses fp64 atomics, whose order-dependent rounding (1e-16) vanishes when the
// statistics are narrowed to fp32: the result is run-to-run reproducible in practice.
__global__ void channel_stats_kernel(const void* __restrict__ x, int xdt, double* __restrict__ stats,
                                     int64_t positions, int C, int chunk) {
  const int b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = p0 + chunk < positions ? p0 + chunk : positions;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      const float v = load_as_float(x, xdt, ((int64_t)b * positions + p) * C + c);
      s += v;
      q = fmaf(v, v, q);
    }
    atomicAdd(&stats[((int64_t)b * 2 + 0) * C + c], (double)s);
    atomicAdd(&stats[((int64_t)b * 2 + 1) * C + c], (double)q);
  }
}

__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const void* __restrict__ x, int xdt, const double* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                      const void* __restrict__ res, int rdt, int elu, void* __restrict__ y, int ydt,
                      int64_t positions, int C, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int64_t b = idx / ((int64_t)C * positions);
  const double inv_n = 1.0 / (double)positions;
  const double mean_d = stats[(b * 2 + 0) * C + c] * inv_n;
  const float mean = (float)mean_d;
  const float var = fmaxf((float)(stats[(b * 2 + 1) * C + c] * inv_n - mean_d * mean_d), 0.f);
  float v = (load_as_float(x, xdt, idx) - mean) * rsqrtf(var + eps) * gamma[c] + beta[c];
  if (res) v += load_as_float(res, rdt, idx);
  if (elu) v = v > 0.f ? v : expm1f(v);
  store_from_float(y, ydt, idx, v);
}

// 128-bit vectorised InstanceNorm apply (C % 8 == 0): a thread owns 8 consecutive channels, derives
// their (scale, shift) once and walks the positions of its block's chunk; loads/stores are 16-byte
// (bf16) or 2 x 16-byte (f32) and a warp covers whole 128-byte lines.
template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void ld8<__half>(const __half* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void st8<__half>(__half* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_h16x2(v[0], v[1], RFK_F16); u.y = pack_h16x2(v[2], v[3], RFK_F16);
  u.z = pack_h16x2(v[4], v[5], RFK_F16); u.w = pack_h16x2(v[6], v[7], RFK_F16);
  *reinterpret_cast<uint4*>(p) = u;
}
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TX, typename TR, typename TY>
__global__ void __launch_bounds__(256)
instnorm_apply_vec_kernel(const TX* __restrict__ x, const double* __restrict__ stats,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                          const TR* __restrict__ res, int elu, TY* __restrict__ y, int64_t positions,
                          int C, int chunk) {
  const int groups = C >> 3;                       // 8-channel groups per position
  const int rows_per_iter = blockDim.x / groups;   // positions covered by one pass of the block
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  if (r >= rows_per_iter) return;
  const int b = blockIdx.y;
  const int c0 = g * 8;
  float sc[8], sh[8];
  const double inv_n = 1.0 / (double)positions;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double mean_d = stats[((int64_t)b * 2 + 0) * C + c0 + i] * inv_n;
    const float mean = (float)mean_d;
    const float var = fmaxf((float)(stats[((int64_t)b * 2 + 1) * C + c0 + i] * inv_n - mean_d * mean_d), 0.f);
    sc[i] = rsqrtf(var + eps) * gamma[c0 + i];
    sh[i] = beta[c0 + i] - mean * sc[i];
  }
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = p0 + chunk < positions ? p0 + chunk : positions;
  auto finish = [&](float (&v)[8], const float (&rr)[8], int64_t off) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]) + rr[i];
    if (elu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : __expf(v[i]) - 1.f;  // abs error ~1e-7
    }
    st8<TY>(y + off, v);
  };
  const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t row0 = (int64_t)b * positions;
  int64_t pos = p0 + r;
  // two positions in flight per thread (all loads issued before the first use)
  for (; pos + rows_per_iter < p1; pos += 2 * rows_per_iter) {
    const int64_t off0 = (row0 + pos) * C + c0, off1 = (row0 + pos + rows_per_iter) * C + c0;
    float v0[8], v1[8];
    ld8<TX>(x + off0, v0);
    ld8<TX>(x + off1, v1);
    if (res) {
      float r0[8], r1[8];
      ld8<TR>(res + off0, r0);
      ld8<TR>(res + off1, r1);
      finish(v0, r0, off0);
      finish(v1, r1, off1);
    } else {
      finish(v0, zero8, off0);
      finish(v1, zero8, off1);
    }
  }
  if (pos < p1) {
    const int64_t off0 = (row0 + pos) * C + c0;
    float v0[8];
    ld8<TX>(x + off0, v0);
    if (res) {
      float r0[8];
      ld8<TR>(res + off0, r0);
      finish(v0, r0, off0);
    } else {
      finish(v0, zero8, off0);
    }
  }
}

__global__ void __launch_bounds__(256)
convert_rows_kernel(const void* __restrict__ x, int xdt, int64_t xs, void* __restrict__ y, int ydt,
                    int64_t ys, int64_t rows, int cols) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols;
  const int c = (int)(idx % cols);
  store_from_float(y, ydt, r * ys + c, load_as_float(x, xdt, r * xs + c));
}

// 8 elements per thread (two 16-byte accesses on the f32 side, one on the bf16 side); cols % 8 == 0
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
convert_rows_vec_kernel(const TX* __restrict__ x, int64_t xs, TY* __restrict__ y, int64_t ys, int64_t rows,
                        int groups) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * groups) return;
  const int64_t r = idx / groups;
  const int c = (int)(idx - r * groups) * 8;
  float v[8];
  ld8<TX>(x + r * xs + c, v);
  st8<TY>(y + r * ys + c, v);
}

// InstanceNorm statistics, 8 channels per thread: a block covers `chunk` positions, its threads form
// (rows_per_iter x C/8) and each keeps 16 fp32 partial sums; the rows are folded in shared memory (fixed
// order) and one thread per channel group issues the fp64 atomics.
template <typename TX>
__global__ void __launch_bounds__(256)
channel_stats_vec_kernel(const TX* __restrict__ x, double* __restrict__ stats, int64_t positions, int C,
                         int chunk) {
  extern __shared__ float red[];  // [rows_per_iter][C][2]
  const int groups = C >> 3;
  const int rows_per_iter = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups;
  const int b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * chunk;
  const int64_t p1 = p0 + chunk < positions ? p0 + chunk : positions;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < rows_per_iter) {
    const TX* base = x + (int64_t)b * positions * C + g * 8;
    int64_t pos = p0 + r;
    // two rows in flight per thread
    for (; pos + rows_per_iter < p1; pos += 2 * rows_per_iter) {
      float v0[8], v1[8];
      ld8<TX>(base + pos * C, v0);
      ld8<TX>(base + (pos + rows_per_iter) * C, v1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += v0[i]; q[i] = fmaf(v0[i], v0[i], q[i]);
        s[i] += v1[i]; q[i] = fmaf(v1[i], v1[i], q[i]);
      }
    }
    if (pos < p1) {
      float v0[8];
      ld8<TX>(base + pos * C, v0);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += v0[i]; q[i] = fmaf(v0[i], v0[i], q[i]); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[((int64_t)r * C + g * 8 + i) * 2 + 0] = s[i];
      red[((int64_t)r * C + g * 8 + i) * 2 + 1] = q[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int rr = 0; rr < rows_per_iter; ++rr) {
      ss += red[((int64_t)rr * C + c) * 2 + 0];
      qq += red[((int64_t)rr * C + c) * 2 + 1];
    }
    atomicAdd(&stats[((int64_t)b * 2 + 0) * C + c], (double)ss);
    atomicAdd(&stats[((int64_t)b * 2 + 1) * C + c], (double)qq);
  }
}

}  // namespace rfk

using namespace rfk;

static inline bool dtype_ok(int d) { return d == RFK_F32 || d == RFK_BF16 || d == RFK_F16; }

static int layernorm_impl(const void* x, int xdt, int64_t xs, const float* gamma, const float* beta,
                          float eps, const float* res, int64_t rs, void* y, int ydt, int64_t ys,
                          int64_t rows, int D, rfk_stream_t stream) {
  if (!x || !y) return RFK_ERR_NULL_POINTER;
  if ((gamma == nullptr) != (beta == nullptr)) return RFK_ERR_NULL_POINTER;
  if (rows < 0 || D <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(xdt) || !dtype_ok(ydt)) return RFK_ERR_BAD_DTYPE;
  if (rows == 0) return RFK_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int xa = xdt == RFK_F32 ? 16 : 8, ya = ydt == RFK_F32 ? 16 : 8;
  const int xe = xdt == RFK_F32 ? 4 : 2, ye = ydt == RFK_F32 ? 4 : 2;
  const bool vec_ok = D % 4 == 0 && D <= 1024 && (reinterpret_cast<uintptr_t>(x) % xa) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) % ya) == 0 && (xs * xe) % xa == 0 &&
                      (ys * ye) % ya == 0 &&
                      (!gamma || (aligned16(gamma) && aligned16(beta))) &&
                      (!res || (aligned16(res) && rs % 4 == 0));
  if (vec_ok && !(xdt == RFK_F16 || (ydt == RFK_F16 && xdt != RFK_F32))) {  // (f16 inputs take the generic kernel)
    if (xdt == RFK_F32 && ydt == RFK_F32)
      return launch_ln_vec<float, float>(x, xs, gamma, beta, eps, y, ys, rows, D, st, res, rs);
    if (xdt == RFK_F32 && ydt == RFK_BF16)
      return launch_ln_vec<float, __nv_bfloat16>(x, xs, gamma, beta, eps, y, ys, rows, D, st, res, rs);
    if (xdt == RFK_F32 && ydt == RFK_F16)
      return launch_ln_vec<float, __half>(x, xs, gamma, beta, eps, y, ys, rows, D, st, res, rs);
    if (xdt == RFK_BF16 && ydt == RFK_BF16)
      return launch_ln_vec<__nv_bfloat16, __nv_bfloat16>(x, xs, gamma, beta, eps, y, ys, rows, D, st, res, rs);
    return launch_ln_vec<__nv_bfloat16, float>(x, xs, gamma, beta, eps, y, ys, rows, D, st, res, rs);
  }
  layernorm_generic_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, xdt, xs, gamma, beta, eps,
                                                                      y, ydt, ys, rows, D, res, rs);
  return post_launch();
}

extern "C" int rfk_layernorm(const void* x, int xdt, int64_t xs, const float* gamma,
                             const float* beta, float eps, void* y, int ydt, int64_t ys,
                             int64_t rows, int D, rfk_stream_t stream) {
  return layernorm_impl(x, xdt, xs, gamma, beta, eps, nullptr, 0, y, ydt, ys, rows, D, stream);
}

extern "C" int rfk_layernorm_residual(const void* x, int xdt, int64_t xs, const float* gamma,
                                      const float* beta, float eps, const float* res, int64_t rs,
                                      void* y, int ydt, int64_t ys, int64_t rows, int D,
                                      rfk_stream_t stream) {
  if (!res) return RFK_ERR_NULL_POINTER;
  return layernorm_impl(x, xdt, xs, gamma, beta, eps, res, rs, y, ydt, ys, rows, D, stream);
}

// logits[b,h,i,j] += -1e9 where |ca_i - ca_j| >= bins[h] (MsaUpdateWithPairAndCoord, :899-913)
__global__ void __launch_bounds__(256)
dist_mask_logits_kernel(const float* __restrict__ ca, int64_t ca_stride, const float* __restrict__ bins,
                        float* __restrict__ logits, int64_t ldl, int B, int H, int L) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * L * L) return;
  const int j = (int)(idx % L);
  const int i = (int)((idx / L) % L);
  const int b = (int)(idx / ((int64_t)L * L));
  const float* pi = ca + ((int64_t)b * L + i) * ca_stride;
  const float* pj = ca + ((int64_t)b * L + j) * ca_stride;
  const float dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
  const float d = sqrtf(dx * dx + dy * dy + dz * dz);
  for (int h = 0; h < H; ++h)
    if (!(d < bins[h])) logits[(((int64_t)b * H + h) * L + i) * ldl + j] += -1e9f;
}

extern "C" int rfk_dist_mask_logits(const float* ca, int64_t ca_stride, const float* bins, int H,
                                    float* logits, int64_t ldl, int B, int L, rfk_stream_t stream) {
  if (!ca || !bins || !logits) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || L <= 0 || H <= 0 || ldl < L || ca_stride < 3) return RFK_ERR_BAD_DIMS;
  const int64_t total = (int64_t)B * L * L;
  dist_mask_logits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ca, ca_stride, bins, logits, ldl, B, H, L);
  return post_launch();
}

extern "C" int rfk_softmax_rows(const float* x, int64_t xs, void* y, int ydt, int64_t ys,
                                int64_t rows, int cols, rfk_stream_t stream) {
  if (!x || !y) return RFK_ERR_NULL_POINTER;
  if (rows < 0 || cols <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(ydt)) return RFK_ERR_BAD_DTYPE;
  if (rows == 0) return RFK_OK;
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, xs, y, ydt, ys, rows, cols);
  return post_launch();
}

extern "C" int rfk_tied_att_symmetrize(const void* A, int adt, int64_t lda, float* att, void* att16,
                                       int64_t att16_stride, int B, int H, int L,
                                       rfk_stream_t stream) {
  if (!A || !att) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || H <= 0 || L <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(adt)) return RFK_ERR_BAD_DTYPE;
  const int64_t total = (int64_t)B * L * L;
  tied_att_sym_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                        reinterpret_cast<cudaStream_t>(stream)>>>(A, adt, lda, att, att16,
                                                                  att16_stride, B, H, L);
  return post_launch();
}

extern "C" int rfk_poswise_weight(const void* pq, int64_t pqs, const void* pk, int64_t pks, int dt,
                                  float scale, float* w_out, const void* q, int64_t qs,
                                  float q_scale, void* qt, int qtdt, int B, int N, int L, int H,
                                  int dh, rfk_stream_t stream) {
  return rfk_poswise_weight_stats(pq, pqs, pk, pks, dt, scale, w_out, q, qs, q_scale, qt, qtdt, nullptr, B, N, L, H,
                                  dh, stream);
}

extern "C" int rfk_poswise_weight_stats(const void* pq, int64_t pqs, const void* pk, int64_t pks, int dt,
                                        float scale, float* w_out, const void* q, int64_t qs,
                                        float q_scale, void* qt, int qtdt, float* stats, int B, int N,
                                        int L, int H, int dh, rfk_stream_t stream) {
  if (!pq || !pk) return RFK_ERR_NULL_POINTER;
  if (qt && !q) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || N <= 0 || L <= 0 || H <= 0 || dh <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(dt) || (qt && !dtype_ok(qtdt))) return RFK_ERR_BAD_DTYPE;
  {
    // vectorised path: bf16 operands, heads of 8k channels, everything 16-byte aligned
    const int D = H * dh, gsz = dh / 8, CH = D / 8;
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    const bool pow2 = gsz > 0 && (gsz & (gsz - 1)) == 0 && gsz <= 32;
    if (is_h16(dt) && (!qt || qtdt == dt) && dh % 8 == 0 && pow2 && CH <= 256 && pqs % 8 == 0 &&
        pks % 8 == 0 && (!qt || qs % 8 == 0) && al16(pq) && al16(pk) && (!qt || (al16(q) && al16(qt))) &&
        (size_t)N * H * sizeof(float) <= 48 * 1024) {
      int rows = 256 / CH;
      while (rows > 1 && (rows * CH) % 32 != 0) --rows;
      if ((rows * CH) % 32 == 0) {
        auto* k = dt == RFK_F16 ? poswise_vec_kernel<true> : poswise_vec_kernel<false>;
        k<<<(unsigned)(B * L), rows * CH, (size_t)N * H * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const uint16_t*>(pq), pqs, reinterpret_cast<const uint16_t*>(pk), pks, scale, w_out,
            reinterpret_cast<const uint16_t*>(q), qs, q_scale, reinterpret_cast<uint16_t*>(qt), N, L, H, dh, rows, stats);
        return post_launch();
      }
    }
  }
  const size_t smem = (size_t)4 * N * sizeof(float);
  if (smem > 48 * 1024) return RFK_ERR_BAD_DIMS;
  const int64_t items = (int64_t)B * L * H;
  poswise_kernel<<<(unsigned)((items + 3) / 4), 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      pq, pqs, pk, pks, dt, scale, w_out, q, qs, q_scale, qt, qtdt, B, N, L, H, dh, stats);
  return post_launch();
}

extern "C" int rfk_opm_prep(const float* m, const float* w, void* xt, void* yt, int tdt,
                            int64_t ldt, float* msa1d, int B, int N, int L, int P,
                            rfk_stream_t stream) {
  if (!m || !w || !xt || !yt || !msa1d) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || N <= 0 || L <= 0 || P <= 0 || ldt < N) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(tdt)) return RFK_ERR_BAD_DTYPE;
  opm_prep_kernel<<<(unsigned)(B * L), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      m, w, xt, yt, tdt, ldt, msa1d, B, N, L, P);
  return post_launch();
}

extern "C" int rfk_pair2att_logits(const float* pair, const float* Wf, const float* bf, float eps,
                                   float* logits, int64_t ldl, int B, int L, int D, int C,
                                   rfk_stream_t stream) {
  if (!pair || !Wf || !bf || !logits) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || L <= 0 || D <= 0 || D > 512 || C <= 0 || C > kP2AMaxC || ldl < L)
    return RFK_ERR_BAD_DIMS;
  const size_t smem = (size_t)C * D * sizeof(float);
  static rfk::PerDeviceOnce once;
  const int cfg_rc = rfk::per_device_once(once, []() {
    return rfk::cuda_status(cudaFuncSetAttribute(pair2att_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  if (D == 288 && (reinterpret_cast<uintptr_t>(pair) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wf) & 15) == 0 &&
      L <= 65535 && B <= 65535) {
    dim3 grid((unsigned)((L + 31) / 32), (unsigned)L, (unsigned)B);
    pair2att_vec_kernel<9><<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(pair, Wf, bf, eps, logits,
                                                                                        ldl, L, C);
    return post_launch();
  }
  const int64_t items = (int64_t)B * L * L;
  pair2att_kernel<<<(unsigned)((items + 7) / 8), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      pair, Wf, bf, eps, logits, ldl, B, L, D, C);
  return post_launch();
}

extern "C" int rfk_pair2att_logits_rows(const float* rows, const float* cols_t, const float* Wf, const float* bf,
                                        float eps, float* logits, int64_t ldl, int B, int Li, int L, int D,
                                        int C, rfk_stream_t stream) {
  if (!rows || !cols_t || !Wf || !bf || !logits) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || Li <= 0 || L <= 0 || Li > L || D <= 0 || D > 512 || C <= 0 || C > kP2AMaxC || ldl < L)
    return RFK_ERR_BAD_DIMS;
  const size_t smem = (size_t)C * D * sizeof(float);
  static rfk::PerDeviceOnce once;
  const int cfg_rc = rfk::per_device_once(once, []() {
    return rfk::cuda_status(cudaFuncSetAttribute(pair2att_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  });
  if (cfg_rc != RFK_OK) return cfg_rc;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (D == 288 && al16(rows) && al16(cols_t) && al16(Wf) && Li <= 65535 && B <= 65535) {
    dim3 grid((unsigned)((L + 31) / 32), (unsigned)Li, (unsigned)B);
    pair2att_rows_vec_kernel<9><<<grid, 256, smem, st>>>(rows, cols_t, Wf, bf, eps, logits, ldl, Li, L, C);
    return post_launch();
  }
  const int64_t items = (int64_t)B * Li * L;
  pair2att_rows_kernel<<<(unsigned)((items + 7) / 8), 256, smem, st>>>(rows, cols_t, Wf, bf, eps, logits, ldl, B, Li,
                                                                       L, D, C);
  return post_launch();
}

extern "C" int rfk_channel_stats(const void* x, int xdt, double* stats, int B, int64_t positions,
                                 int C, rfk_stream_t stream) {
  if (!x || !stats) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || positions <= 0 || C <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(xdt)) return RFK_ERR_BAD_DTYPE;
  if (xdt != RFK_F16 && C % 8 == 0 && C <= 2048 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int groups = C / 8;
    const int threads = groups >= 256 ? 256 : (256 / groups) * groups;
    const int rows_per_iter = threads / groups;
    const size_t smem = (size_t)rows_per_iter * C * 2 * sizeof(float);
    if (threads >= groups && smem <= 48 * 1024) {
      const int vchunk = 256;
      dim3 vgrid((unsigned)((positions + vchunk - 1) / vchunk), (unsigned)B);
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      if (xdt == RFK_BF16)
        channel_stats_vec_kernel<__nv_bfloat16><<<vgrid, threads, smem, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), stats, positions, C, vchunk);
      else
        channel_stats_vec_kernel<float><<<vgrid, threads, smem, st>>>(reinterpret_cast<const float*>(x), stats,
                                                                      positions, C, vchunk);
      return post_launch();
    }
  }
  const int chunk = 128;
  dim3 grid((unsigned)((positions + chunk - 1) / chunk), (unsigned)B);
  const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  channel_stats_kernel<<<grid, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, xdt, stats, positions, C, chunk);
  return post_launch();
}

extern "C" int rfk_instnorm_apply(const void* x, int xdt, const double* stats, const float* gamma,
                                  const float* beta, float eps, const void* res, int rdt, int elu,
                                  void* y, int ydt, int B, int64_t positions, int C,
                                  rfk_stream_t stream) {
  if (!x || !stats || !gamma || !beta || !y) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || positions <= 0 || C <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(xdt) || !dtype_ok(ydt) || (res && !dtype_ok(rdt))) return RFK_ERR_BAD_DTYPE;
  const int64_t total = (int64_t)B * positions * C;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool f16_out = xdt == RFK_F32 && ydt == RFK_F16;  // (the heads: fp32 convolution output -> IEEE-half operand)
  if (xdt != RFK_F16 && (ydt != RFK_F16 || f16_out) && C % 8 == 0 && C <= 2048 && al16(x) && al16(y) && (!res || al16(res)) &&
      (!res || rdt == RFK_F32)) {
    const int chunk = 256;
    const int groups = C / 8;
    const int threads = groups >= 256 ? 256 : (256 / groups) * groups;
    if (threads >= groups) {
      dim3 grid((unsigned)((positions + chunk - 1) / chunk), (unsigned)B);
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      const float* rf = reinterpret_cast<const float*>(res);
#define RFK_IN_LAUNCH(TX, TY)                                                                       \
  instnorm_apply_vec_kernel<TX, float, TY><<<grid, threads, 0, st>>>(                                \
      reinterpret_cast<const TX*>(x), stats, gamma, beta, eps, rf, elu, reinterpret_cast<TY*>(y), \
      positions, C, chunk)
      if (f16_out) RFK_IN_LAUNCH(float, __half);
      else if (xdt == RFK_BF16 && ydt == RFK_BF16) RFK_IN_LAUNCH(__nv_bfloat16, __nv_bfloat16);
      else if (xdt == RFK_BF16) RFK_IN_LAUNCH(__nv_bfloat16, float);
      else if (ydt == RFK_BF16) RFK_IN_LAUNCH(float, __nv_bfloat16);
      else RFK_IN_LAUNCH(float, float);
#undef RFK_IN_LAUNCH
      return post_launch();
    }
  }
  instnorm_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(
      x, xdt, stats, gamma, beta, eps, res, rdt, elu, y, ydt, positions, C, total);
  return post_launch();
}

extern "C" int rfk_convert_rows(const void* x, int xdt, int64_t xs, void* y, int ydt, int64_t ys,
                                int64_t rows, int cols, rfk_stream_t stream) {
  if (!x || !y) return RFK_ERR_NULL_POINTER;
  if (rows < 0 || cols <= 0) return RFK_ERR_BAD_DIMS;
  if (!dtype_ok(xdt) || !dtype_ok(ydt)) return RFK_ERR_BAD_DTYPE;
  if (rows == 0) return RFK_OK;
  const int64_t total = rows * cols;
  {
    const int xe = xdt == RFK_F32 ? 4 : 2, ye = ydt == RFK_F32 ? 4 : 2;
    auto ok = [](const void* q, int64_t stride, int es) {
      return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (stride * es) % 16 == 0;
    };
    if (cols % 8 == 0 && ok(x, xs, xe) && ok(y, ys, ye) && xdt != ydt && (xdt == RFK_F32 || ydt == RFK_F32)) {
      const int groups = cols / 8;
      const int64_t work = rows * groups;
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      if (xdt == RFK_F32 && ydt == RFK_F16)
        convert_rows_vec_kernel<float, __half><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float*>(x), xs, reinterpret_cast<__half*>(y), ys, rows, groups);
      else if (xdt == RFK_F16)
        convert_rows_vec_kernel<__half, float><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const __half*>(x), xs, reinterpret_cast<float*>(y), ys, rows, groups);
      else if (xdt == RFK_F32)
        convert_rows_vec_kernel<float, __nv_bfloat16><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float*>(x), xs, reinterpret_cast<__nv_bfloat16*>(y), ys, rows, groups);
      else
        convert_rows_vec_kernel<__nv_bfloat16, float><<<(unsigned)((work + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), xs, reinterpret_cast<float*>(y), ys, rows, groups);
      return post_launch();
    }
  }
  convert_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                        reinterpret_cast<cudaStream_t>(stream)>>>(x, xdt, xs, y, ydt, ys, rows, cols);
  return post_launch();
}
