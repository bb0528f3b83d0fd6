/*
 * rfk.h — C ABI of librfk.so: the B200 (sm_100a) kernels behind the three-track trunk of
 * rosettafold-pytorch (MSA row/column attention, MSA->pair outer product, pair axial attention,
 * pair->MSA update and their LayerNorm / softmax / residual epilogues).
 *
 * Conventions:
 *  - every entry point is `extern "C"`, takes plain pointers + sizes, enqueues work on `stream`
 *    (a cudaStream_t passed as void*), never synchronises, never allocates persistent memory and
 *    keeps no pointer after it returns; the caller owns all buffers including workspaces;
 *  - returns 0 (RFK_OK) or an error code; never throws, never exits; rfk_strerror() names a code;
 *  - dtypes are RFK_F32 / RFK_BF16 / RFK_F16 (IEEE half: same tensor-core rate as bf16, 11 instead of 8 significand
 *    bits; used for the range-bounded operands of the MSA track); "mode 0" callers hand 16-bit operands to the tcgen05 kernels,
 *    "mode 1" (fp32 validation) callers hand fp32 operands to the SIMT fp32 kernels;
 *  - all index arithmetic is in ELEMENTS (not bytes).
 *
 * Each function cites the reference lines (rosettafold_pytorch/rosettafold_pytorch.py unless
 * stated otherwise) whose arithmetic it replaces.
 */
#ifndef RFK_H
#define RFK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rfk_stream_t; /* cudaStream_t */

enum { RFK_F32 = 0, RFK_BF16 = 1, RFK_F16 = 2 };

enum {
  RFK_OK = 0,
  RFK_ERR_BAD_DIMS = 1,
  RFK_ERR_MISALIGNED = 2,
  RFK_ERR_UNSUPPORTED_ARCH = 3,
  RFK_ERR_BAD_DTYPE = 4,
  RFK_ERR_NULL_POINTER = 5,
  RFK_ERR_TMA_ENCODE = 6,
  RFK_ERR_WORKSPACE = 7,
  RFK_ERR_UNSUPPORTED = 8,
  RFK_ERR_CUDA_BASE = 1000 /* RFK_ERR_CUDA_BASE + cudaError_t */
};

const char* rfk_strerror(int code);
int rfk_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rfk_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Generalised element addressing used by the GEMM epilogue.
 * offset(z0,z1,z2,m,n) = z0*zs[0]+z1*zs[1]+z2*zs[2] + (m%MR)*ms[0]+(m/MR)*ms[1]
 *                        + (n%NR)*ns[0]+(n/NR)*ns[1]
 * It lets one GEMM write the permuted layouts the reference produces with einops
 * rearrange copies (:38,:51,:248,:258,:403,:425,:593) without a separate transpose pass.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t zs[3];
  int64_t ms[2];
  int64_t ns[2];
} rfk_addr;

enum { RFK_ACT_NONE = 0, RFK_ACT_RELU = 1, RFK_ACT_ELU = 2 };
enum { RFK_EPI_STD = 0, RFK_EPI_BLOCKLN32 = 1 };

/*
 * Batched "TN" GEMM:  C[z][m][n] = epi( alpha * sum_k A[z][m][k] * B[z][n][k] ).
 * A is [Z][M][K] and B is [Z][N][K], both with K contiguous (nn.Linear weight layout).
 * z = (z2*Z[1] + z1)*Z[0] + z0; a_zs/b_zs/bias_zs give the per-level strides (0 = broadcast).
 * Epilogue:  x = alpha*acc + bias[n];  x = act(x);  x += r0[...] + r1[...];  c[...] = x.
 * RFK_EPI_BLOCKLN32 (needs MR == NR == 32): every aligned 32x32 block (m/32, n/32) of the
 * product is LayerNorm-ed over its 1024 entries, index (m%32)*32 + n%32, with ln_gamma/ln_beta,
 * before bias/act/residuals are skipped and the value is stored: this is the outer product
 * "b n i u, b n j v -> b i j (u v)" followed by LayerNorm(1024) (:424-425, :416).
 *
 * Replaces: nn.Linear (:195-202, :235-238, :274-277, :436, :448, :566, :574, performer
 * to_q/to_k/to_v/to_out), the einsums at :212, :254, :257, :424, :592 and the Residual adds
 * at :26-28, :346, :595.
 * ab_dtype RFK_BF16 / RFK_F16 -> tcgen05/TMEM/TMA kernel (fp32 accumulate); RFK_F32 -> SIMT fp32 kernel.
 * bf16 operands need: 16-byte aligned base pointers, lda/ldb/z-strides multiples of 8 elements.
 */
typedef struct {
  const void* a;
  const void* b;
  int32_t ab_dtype;
  int32_t act;
  int64_t M, N, K;
  int64_t Z[3];
  int64_t lda, ldb;
  int64_t a_zs[3];
  int64_t b_zs[3];
  const float* bias;
  int64_t bias_zs[3];
  float alpha;
  int32_t epi;
  int64_t MR, NR;
  void* c;
  const void* r0;
  const void* r1;
  int32_t c_dtype, r0_dtype, r1_dtype;
  float ln_eps;
  rfk_addr c_addr, r0_addr, r1_addr;
  const float* ln_gamma;
  const float* ln_beta;
} rfk_gemm_desc;

int rfk_gemm(const rfk_gemm_desc* d, rfk_stream_t stream);

/*
 * Row LayerNorm (nn.LayerNorm, eps inside the sqrt; :323,:328,:416,:435,:437,:442-443,
 * :522-524,:565,:573,:580):  y[r, :] = (x[r, :] - mean) * rsqrt(var + eps) * gamma + beta.
 * gamma/beta may be NULL (pure normalisation). Row strides let the output land in a column
 * slice of a wider buffer (the 716-channel concat of :487-496 is never materialised in fp32).
 */
int rfk_layernorm(const void* x, int x_dtype, int64_t x_row_stride, const float* gamma,
                  const float* beta, float eps, void* y, int y_dtype, int64_t y_row_stride,
                  int64_t rows, int D, rfk_stream_t stream);

/* y[r, :] = LayerNorm(x[r, :]) + res[r, :]  (res f32 with its own row stride): the
 * `msa + ln_out(out)` of MsaUpdateWithPairAndCoord (:916) in one pass. */
int rfk_layernorm_residual(const void* x, int x_dtype, int64_t x_row_stride, const float* gamma,
                           const float* beta, float eps, const float* res, int64_t res_row_stride,
                           void* y, int y_dtype, int64_t y_row_stride, int64_t rows, int D,
                           rfk_stream_t stream);

/*
 * Distance mask of MsaUpdateWithPairAndCoord (:899-913): logits[b,h,i,j] += -1e9 wherever the
 * C-alpha distance |ca[b,i] - ca[b,j]| is not below bins[h]. ca: f32, residue (b,i) at
 * ca + (b*L+i)*ca_stride (3 coordinates); logits f32 [B,H,L,ld_logits], updated in place.
 */
int rfk_dist_mask_logits(const float* ca, int64_t ca_stride, const float* bins, int H, float* logits,
                         int64_t ld_logits, int B, int L, rfk_stream_t stream);

/* Row softmax over the last dimension (:255, :569): y[r,:] = softmax(x[r,:]); x fp32. */
int rfk_softmax_rows(const float* x, int64_t x_row_stride, void* y, int y_dtype,
                     int64_t y_row_stride, int64_t rows, int cols, rfk_stream_t stream);

/*
 * Symmetrised tied-attention map (:263-264): att[b,i,j,h] = 0.5*(A[b,h,i,j] + A[b,h,j,i]).
 * A: [B,H,L,lda] (bf16 or f32), att: f32 [B,L,L,H] (+ optional bf16 copy written with row
 * stride att16_stride into a column slice of the PairUpdateWithMsa feature buffer, :493).
 */
int rfk_tied_att_symmetrize(const void* A, int a_dtype, int64_t lda, float* att, void* att16,
                            int64_t att16_stride, int B, int H, int L, rfk_stream_t stream);

/*
 * PositionWiseWeightFactor (:205-217) fused with the query scaling of :252 and the
 * "b n l (h d) -> b h l (n d)" relayout feeding the tied-logit contraction (:254).
 *   logit[b,l,h,n] = scale * sum_d pq[b,l,h*dh+d] * pk[b,n,l,h*dh+d];  w = softmax_n(logit)
 *   w_out[b,n,l,h] = w                                  (optional, f32)
 *   qt[b,h,l,n*dh+d] = q[b,n,l,h*dh+d] * w * q_scale    (optional)
 * pq: [B,L,H*dh] row stride pq_stride; pk, q: [B,N,L,H*dh] row strides pk_stride, q_stride.
 */
int rfk_poswise_weight(const void* pq, int64_t pq_stride, const void* pk, int64_t pk_stride,
                       int in_dtype, float scale, float* w_out, const void* q, int64_t q_stride,
                       float q_scale, void* qt, int qt_dtype, int B, int N, int L, int H, int dh,
                       rfk_stream_t stream);

/*
 * Same, for an MSA whose sequences are sharded over devices (the softmax of :213 runs over ALL sequences):
 * additionally writes stats[b,l,h,0..1] = (max_n logit, sum_n exp(logit - max)) of the N sequences given, so
 * that the caller can merge the shards: with M = max over shards of max_r and S = sum_r sum_r exp(max_r - M),
 * the globally normalised weights are w_r * (sum_r exp(max_r - M) / S). stats may be NULL.
 */
int rfk_poswise_weight_stats(const void* pq, int64_t pq_stride, const void* pk, int64_t pk_stride,
                             int in_dtype, float scale, float* w_out, const void* q, int64_t q_stride,
                             float q_scale, void* qt, int qt_dtype, float* stats, int B, int N, int L,
                             int H, int dh, rfk_stream_t stream);

/*
 * Operand preparation for the outer-product sum (:469-482): from m = proj_msa(msa) [B,N,L,P]
 * (f32) and w [B,N,L] (f32) write
 *   xt[b, l*P+u, n] = m[b,n,l,u]            yt[b, l*P+v, n] = m[b,n,l,v] * w[b,n,l]
 * (K-major operands of the OPM GEMM, leading dimension ldt >= N) and
 *   msa1d[b,l,0:P] = sum_n m[b,n,l,:]        msa1d[b,l,P:2P] = m[b,0,l,:]     (f32)
 */
int rfk_opm_prep(const float* m, const float* w, void* xt, void* yt, int t_dtype, int64_t ldt,
                 float* msa1d, int B, int N, int L, int P, rfk_stream_t stream);

/*
 * pair2att logits for all encoder layers at once (:563-566; the pair input is the same for
 * every layer, :607-610):  s = 0.5*(pair[b,i,j,:] + pair[b,j,i,:]);  xhat = (s-mean)*rstd;
 *   logits[b, c, i, j] = sum_d Wf[c,d]*xhat[d] + bf[c],  c in [0, C)   (C = layers*heads)
 * with Wf = W*gamma and bf = W@beta + b folded on the host. pair f32 [B,L,L,D].
 */
int rfk_pair2att_logits(const float* pair, const float* Wf, const float* bf, float eps,
                        float* logits, int64_t ld_logits, int B, int L, int D, int C,
                        rfk_stream_t stream);

/*
 * Same for a ROW-SHARDED pair map (one long protein over several devices): this device holds rows
 * [i0, i0+Li) as rows[b,il,j,:] = pair[b,i0+il,j,:] and the transposed shard cols_t[b,j,il,:] = pair[b,j,i0+il,:]
 * (what an all-to-all of the row shards delivers); writes logits[b, c, il, j] for its rows and every j
 * (f32 [B,C,Li,ld_logits]). The symmetrisation (:555-556) is the only place the pair rows meet their columns.
 */
int rfk_pair2att_logits_rows(const float* rows, const float* cols_t, const float* Wf, const float* bf,
                             float eps, float* logits, int64_t ld_logits, int B, int Li, int L, int D,
                             int C, rfk_stream_t stream);

/*
 * Per-(batch, channel) statistics over the L*L positions of a channels-last map
 * (nn.InstanceNorm2d, :453,:457): stats[b,0,c] = sum x, stats[b,1,c] = sum x^2 (f64, caller
 * zeroes `stats` first; fixed-order f32 partial sums, f64 atomics across blocks, so the values
 * narrowed to f32 are run-to-run reproducible).
 */
int rfk_channel_stats(const void* x, int x_dtype, double* stats, int B, int64_t positions, int C,
                      rfk_stream_t stream);

/*
 * Apply InstanceNorm2d(affine, eps) + optional residual + ELU on a channels-last map (:453-462):
 *   y = (x - mean_c) * rsqrt(var_c + eps) * gamma_c + beta_c;  if (res) y += res;  if (elu) y = ELU(y)
 */
int rfk_instnorm_apply(const void* x, int x_dtype, const double* stats, const float* gamma,
                       const float* beta, float eps, const void* res, int res_dtype, int elu,
                       void* y, int y_dtype, int B, int64_t positions, int C, rfk_stream_t stream);

/*
 * Performer FAVOR+ attention (performer_pytorch.FastAttention as called at :313-318 and
 * :505-518; spec in SURVEY.md section 8c), fused: feature map (softmax kernel or ReLU kernel),
 * key-sum, context, normaliser and output in one kernel, nothing of size tokens x m leaves the SM.
 *   q,k,v: element (g1, g0, t, h, d) at  base + g1*gs[1] + g0*gs[0] + t*ts + h*64 + d
 *   out  : same indexing with out_gs/out_ts.
 * `proj` is the (m x 64) f32 projection matrix buffer. kind: 0 = softmax kernel (eps 1e-4),
 * 1 = generalised ReLU kernel (eps 1e-3). io_dtype RFK_BF16 -> tcgen05 kernel, RFK_F32 -> SIMT.
 */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  const float* proj;
  int32_t io_dtype;
  int32_t kind;
  int32_t m_features;
  int32_t heads;
  int64_t tokens;
  int64_t G[2];
  int64_t gs[2];
  int64_t ts;
  int64_t out_gs[2];
  int64_t out_ts;
} rfk_favor_desc;

int rfk_favor_attention(const rfk_favor_desc* d, rfk_stream_t stream);

/*
 * 3x3 "same" convolution, no bias, on a channels-last map (nn.Conv2d(d_pair, d_pair, 3,
 * padding="same", bias=False), :452 and :456) as an implicit GEMM on the tcgen05 kernel.
 *   x: bf16 [B][L][L][C] (C % 8 == 0);  y: bf16 or f32 [B][L][L][Cout] (Cout % 32 == 0)
 *   w_packed: bf16 [Cout][9][Cpad], Cpad = ceil(C/64)*64, w_packed[o][3*di+dj][c] = W[o][c][di][dj],
 *   zero for c >= C.
 */
int rfk_conv3x3_nhwc(const void* x, const void* w_packed, void* y, int y_dtype, int B, int L, int C,
                     int Cout, rfk_stream_t stream);
/* Same on H x L images (x: [B][H][L][C]): a row shard of the pair map plus its halo rows
 * (long-protein path). */
int rfk_conv3x3_nhwc_hw(const void* x, const void* w_packed, void* y, int y_dtype, int B, int H, int L,
                        int C, int Cout, rfk_stream_t stream);

/* fp32 validation-mode form of the same convolution (SIMT, fp32 FMA, fixed summation order):
 *   x: f32 [B][H][L][C];  w_packed: f32 [9][C][Cout], w_packed[3*di+dj][c][o] = W[o][c][di][dj];  y: f32 [B][H][L][Cout]. */
int rfk_conv3x3_nhwc_f32(const float* x, const float* w_packed, float* y, int B, int H, int L, int C, int Cout,
                         rfk_stream_t stream);

/* Dilated forms (ResBlock2D of the prediction heads, resnet.py:14-44: nn.Conv2d(C, C, 3, dilation=d, padding="same",
 * bias=False)): tap (di, dj) reads the image at (i + (di - 1) d, j + (dj - 1) d), zeros outside. dilation 1 = the calls above.
 * tcgen05 form: 1 <= dilation <= 64, image and packed weights of `x_dtype` = RFK_BF16 or RFK_F16 (the heads' convolution
 * inputs follow an InstanceNorm + ELU: range-bounded, so IEEE half), y of RFK_BF16 / RFK_F16 / RFK_F32;
 * fp32 SIMT form: 1 <= dilation <= 8. */
int rfk_conv3x3_nhwc_dil(const void* x, int x_dtype, const void* w_packed, void* y, int y_dtype, int B, int H, int L,
                         int C, int Cout, int dilation, rfk_stream_t stream);
int rfk_conv3x3_nhwc_f32_dil(const float* x, const float* w_packed, float* y, int B, int H, int L, int C, int Cout,
                             int dilation, rfk_stream_t stream);

/*
 * Prediction heads (reference :1130-1172): symmetrisation of the projected pair map in
 * front of the distance / omega heads (:1166), channels-last:  y[b][i][j][:] = 0.5 (x[b][i][j][:] + x[b][j][i][:]).
 *   x, y: [B][L][L][C] of `dtype` (RFK_F32: C % 4 == 0; 16-bit: C % 8 == 0), contiguous, 16-byte aligned, x != y.
 */
int rfk_pair_symmetrize(const void* x, void* y, int dtype, int B, int L, int C, rfk_stream_t stream);

/*
 * Embeddings that feed the trunk (reference :57-181), fused gathers on the device
 * (the reference gathers CPU-resident tables in Python loops over the batch, :73, :98, :115-116).
 *   rfk_msa_embed:  MsaEmbedding.forward :114-120.  tokens i64 [B][N][L] in [0, V), aa_idx i64 [B][L] in [0, max_len);
 *     emb f32 [V][D], pos_enc f32 [max_len][D] (SinusoidalPositionalEncoding table :63-68), query_enc f32 [2][D];
 *     out f32 [B][N][L][D] = emb[tok] + pos_enc[aa] + query_enc[n == 0 ? 0 : 1].   D % 4 == 0.
 *   rfk_pair_embed: PairEmbedding.forward :147-175 without template. seq, aa_idx i64 [B][L];
 *     table_left / table_right f32 [V][D] = embed_seq.weight @ proj.weight[:, :D/2]^T / [:, D/2:D]^T (the Linear of :173
 *     applied to the two gathered halves of the concatenation), w_sep f32 [D] = proj.weight[:, D], bias f32 [D],
 *     pos_enc_half f32 [max_len][D/2] (SinusoidalPositionalEncoding2D table :86-91);
 *     out f32 [B][L][L][D] = table_left[seq_j] + table_right[seq_i] + w_sep log(|aa_i - aa_j| + 1) + bias
 *                            + [pos_enc_half[aa_i] | pos_enc_half[aa_j]].   D % 8 == 0.
 * Index ranges are the caller's contract (the Python modules check them like nn.Embedding does).
 */
int rfk_msa_embed(const int64_t* tokens, const int64_t* aa_idx, const float* emb, const float* pos_enc,
                  const float* query_enc, float* out, int B, int N, int L, int D, rfk_stream_t stream);
int rfk_pair_embed(const int64_t* seq, const int64_t* aa_idx, const float* table_left, const float* table_right,
                   const float* w_sep, const float* bias, const float* pos_enc_half, float* out, int B, int L, int D,
                   rfk_stream_t stream);

/* Cast / copy rows between dtypes with row strides (host-side plumbing for column slices). */
int rfk_convert_rows(const void* x, int x_dtype, int64_t x_row_stride, void* y, int y_dtype,
                     int64_t y_row_stride, int64_t rows, int cols, rfk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RFK_H */
