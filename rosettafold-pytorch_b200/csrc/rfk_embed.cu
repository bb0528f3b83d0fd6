// rfk_embed.cu — the embeddings that feed the trunk (reference rosettafold_pytorch.py:57-181), as fused
// gather kernels: integer token / residue-index inputs, fp32 outputs in the trunk's layouts. Both kernels are
// HBM-write-bound (every output element is written once, all operands are small tables that live in L2 / L1):
// one thread per 16-byte output chunk, consecutive threads on consecutive chunks (whole 128-byte lines per warp).
//
// MsaEmbedding (:106-120):  out[b,n,l,:] = E[tok[b,n,l]] + PE[aa[b,l]] + Q[n == 0 ? 0 : 1]
// PairEmbedding (:123-181): the Linear over the concatenation [emb[seq_j] | emb[seq_i] | log(|aa_i - aa_j| + 1)] is
//   linear in its three parts, so with the per-vocabulary tables TL = emb W_left^T, TR = emb W_right^T (host, at
//   weight-packing time) out[b,i,j,:] = TL[seq[b,j]] + TR[seq[b,i]] + wsep * log(|aa_i - aa_j| + 1) + bias
//                                       + [PE2[aa[b,i]] | PE2[aa[b,j]]]        (2-D sinusoidal encoding, :79-103)
//   — the (B, L, L, 289) concatenation of :171 is never built.
#include "rfk_common.cuh"

namespace rfk {
namespace {

__global__ void __launch_bounds__(256)
msa_embed_kernel(const int64_t* __restrict__ tok, const int64_t* __restrict__ aa, const float* __restrict__ E,
                 const float* __restrict__ PE, const float* __restrict__ Q, float* __restrict__ out, int64_t rows,
                 int N, int L, int D4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * D4) return;
  const int c = (int)(idx % D4);
  const int64_t r = idx / D4;  // (b, n, l)
  const int l = (int)(r % L);
  const int n = (int)((r / L) % N);
  const int64_t b = r / ((int64_t)L * N);
  const int64_t t = __ldg(tok + r), a = __ldg(aa + b * L + l);
  const float4 e = __ldg(reinterpret_cast<const float4*>(E) + t * D4 + c);
  const float4 p = __ldg(reinterpret_cast<const float4*>(PE) + a * D4 + c);
  const float4 q = __ldg(reinterpret_cast<const float4*>(Q) + (n == 0 ? 0 : D4) + c);
  // the reference adds in this order: (E + PE) + Q (:119)
  reinterpret_cast<float4*>(out)[idx] = make_float4((e.x + p.x) + q.x, (e.y + p.y) + q.y, (e.z + p.z) + q.z, (e.w + p.w) + q.w);
}

__global__ void __launch_bounds__(256)
pair_embed_kernel(const int64_t* __restrict__ seq, const int64_t* __restrict__ aa, const float* __restrict__ TL,
                  const float* __restrict__ TR, const float* __restrict__ wsep, const float* __restrict__ bias,
                  const float* __restrict__ PE2, float* __restrict__ out, int64_t positions, int L, int D4, int H4) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= positions * D4) return;
  const int c = (int)(idx % D4);
  const int64_t r = idx / D4;  // (b, i, j)
  const int j = (int)(r % L);
  const int i = (int)((r / L) % L);
  const int64_t b = r / ((int64_t)L * L);
  const int64_t si = __ldg(seq + b * L + i), sj = __ldg(seq + b * L + j);
  const int64_t ai = __ldg(aa + b * L + i), aj = __ldg(aa + b * L + j);
  const int64_t d = ai - aj;
  const float sep = logf((float)(d < 0 ? -d : d) + 1.0f);
  const float4 l4 = __ldg(reinterpret_cast<const float4*>(TL) + sj * D4 + c);
  const float4 r4 = __ldg(reinterpret_cast<const float4*>(TR) + si * D4 + c);
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(wsep) + c);
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + c);
  // channels [0, d/2): encoding of residue i, [d/2, d): of residue j (:99-102)
  const float4 p4 = c < H4 ? __ldg(reinterpret_cast<const float4*>(PE2) + ai * H4 + c)
                           : __ldg(reinterpret_cast<const float4*>(PE2) + aj * H4 + (c - H4));
  float4 o;
  o.x = (((l4.x + r4.x) + w4.x * sep) + b4.x) + p4.x;
  o.y = (((l4.y + r4.y) + w4.y * sep) + b4.y) + p4.y;
  o.z = (((l4.z + r4.z) + w4.z * sep) + b4.z) + p4.z;
  o.w = (((l4.w + r4.w) + w4.w * sep) + b4.w) + p4.w;
  reinterpret_cast<float4*>(out)[idx] = o;
}

}  // namespace
}  // namespace rfk

using namespace rfk;

extern "C" int rfk_msa_embed(const int64_t* tokens, const int64_t* aa_idx, const float* emb, const float* pos_enc,
                             const float* query_enc, float* out, int B, int N, int L, int D, rfk_stream_t stream) {
  if (!tokens || !aa_idx || !emb || !pos_enc || !query_enc || !out) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || N <= 0 || L <= 0 || D <= 0 || D % 4) return RFK_ERR_BAD_DIMS;
  if (!aligned16(emb) || !aligned16(pos_enc) || !aligned16(query_enc) || !aligned16(out)) return RFK_ERR_MISALIGNED;
  if (check_arch() != RFK_OK) return RFK_ERR_UNSUPPORTED_ARCH;
  const int64_t rows = (int64_t)B * N * L, work = rows * (D / 4);
  msa_embed_kernel<<<(unsigned)((work + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tokens, aa_idx, emb, pos_enc, query_enc, out, rows, N, L, D / 4);
  return post_launch();
}

extern "C" int rfk_pair_embed(const int64_t* seq, const int64_t* aa_idx, const float* table_left, const float* table_right,
                              const float* w_sep, const float* bias, const float* pos_enc_half, float* out, int B, int L,
                              int D, rfk_stream_t stream) {
  if (!seq || !aa_idx || !table_left || !table_right || !w_sep || !bias || !pos_enc_half || !out) return RFK_ERR_NULL_POINTER;
  if (B <= 0 || L <= 0 || D <= 0 || D % 8) return RFK_ERR_BAD_DIMS;
  if (!aligned16(table_left) || !aligned16(table_right) || !aligned16(w_sep) || !aligned16(bias) || !aligned16(pos_enc_half) ||
      !aligned16(out))
    return RFK_ERR_MISALIGNED;
  if (check_arch() != RFK_OK) return RFK_ERR_UNSUPPORTED_ARCH;
  const int64_t positions = (int64_t)B * L * L, work = positions * (D / 4);
  if ((work + 255) / 256 > 0x7fffffffLL) return RFK_ERR_BAD_DIMS;
  pair_embed_kernel<<<(unsigned)((work + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      seq, aa_idx, table_left, table_right, w_sep, bias, pos_enc_half, out, positions, L, D / 4, D / 8);
  return post_launch();
}
