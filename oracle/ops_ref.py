"""ORACLE / TEST INFRASTRUCTURE — not part of the product.

Plain-PyTorch (CPU-capable, fp64-accumulating) restatement of every librfk op at the C-ABI
boundary (include/rfk.h). Two uses, both in tests/ only:
  * the per-op checker for the `-m gpu` parity tests (CUDA kernel vs this, same inputs);
  * `RefBackend`, swapped in through ops._set_backend_for_tests() so the host logic of the
    nn.Modules (layout juggling, weight packing, residual wiring) can be checked against the
    reference forward on a machine without a GPU.
Each function cites the reference lines (rosettafold_pytorch/rosettafold_pytorch.py) or the
performer-pytorch 1.1.4 spec it follows.
"""
from __future__ import annotations

import torch

ACT_NONE, ACT_RELU, ACT_ELU = 0, 1, 2
EPI_STD, EPI_BLOCKLN32 = 0, 1


def _act(x, act):
    if act == ACT_RELU:
        return torch.relu(x)
    if act == ACT_ELU:
        return torch.nn.functional.elu(x)
    return x


class RefBackend:
    name = "oracle-ref"

    def __init__(self, acc_dtype=torch.float64):
        self.acc = acc_dtype

    # nn.Linear / einsum contractions (:195-202,:212,:235-238,:254,:257,:274-277,:424,:592)
    def gemm(self, a, b, c_view, bias, act, alpha, r0, r1, epi, ln_gamma, ln_beta, ln_eps):
        acc = torch.matmul(a.to(self.acc), b.to(self.acc).transpose(-1, -2)) * alpha  # [Z2,Z1,Z0,M,N]
        Z2, Z1, Z0, M1, MR, N1, NR = c_view.shape
        acc = acc.reshape(Z2, Z1, Z0, M1, MR, N1, NR)
        if epi == EPI_BLOCKLN32:
            # outer product rearranged to (u v) then LayerNorm(1024) (:424-425, :416)
            blk = acc.permute(0, 1, 2, 3, 5, 4, 6).reshape(Z2, Z1, Z0, M1, N1, MR * NR)
            mean = blk.mean(-1, keepdim=True)
            var = blk.var(-1, unbiased=False, keepdim=True)
            blk = (blk - mean) / torch.sqrt(var + ln_eps)
            if ln_gamma is not None:
                blk = blk * ln_gamma.to(self.acc) + ln_beta.to(self.acc)
            acc = blk.reshape(Z2, Z1, Z0, M1, N1, MR, NR).permute(0, 1, 2, 3, 5, 4, 6)
        else:
            if bias is not None:
                acc = acc + bias.to(self.acc).reshape(1, 1, 1, 1, 1, N1, NR)
            acc = _act(acc, act)
            if r0 is not None:
                acc = acc + r0.to(self.acc)
            if r1 is not None:
                acc = acc + r1.to(self.acc)
        c_view.copy_(acc.to(c_view.dtype))

    # nn.LayerNorm (:323 etc.)
    def layernorm(self, x, gamma, beta, eps, out, res=None):
        xf = x.to(self.acc)
        mean = xf.mean(-1, keepdim=True)
        var = xf.var(-1, unbiased=False, keepdim=True)
        y = (xf - mean) / torch.sqrt(var + eps)
        if gamma is not None:
            y = y * gamma.to(self.acc) + beta.to(self.acc)
        if res is not None:
            y = y + res.to(self.acc)
        out.copy_(y.to(out.dtype))

    # distance mask of MsaUpdateWithPairAndCoord (:899-913)
    def dist_mask_logits(self, ca, bins, logits):
        L = logits.shape[2]
        pdist = torch.cdist(ca.to(self.acc), ca.to(self.acc))
        for h in range(bins.numel()):
            logits[:, h, :, :L] += ((pdist < float(bins[h])).to(logits.dtype) - 1.0) * 1e9

    # softmax (:255, :569)
    def softmax_rows(self, x, out):
        out.copy_(torch.softmax(x.to(self.acc), dim=-1).to(out.dtype))

    # (att + att^T)/2, b h i j -> b i j h (:263-264)
    def tied_att_symmetrize(self, A, att, att16):
        a = A.to(self.acc)
        s = 0.5 * (a + a.transpose(-1, -2))
        s = s.permute(0, 2, 3, 1)
        att.copy_(s.to(att.dtype))
        if att16 is not None:
            att16.copy_(s.to(att16.dtype))

    # PositionWiseWeightFactor (:205-217) + q scaling (:252) + relayout for :254
    def poswise_weight(self, pq, pk, scale, w_out, q, q_scale, qt, H, dh, stats=None):
        B, N, L, D = pk.shape
        pqh = pq.to(self.acc).reshape(B, L, H, dh)
        pkh = pk.to(self.acc).reshape(B, N, L, H, dh)
        logits = torch.einsum("blhd,bnlhd->blhn", pqh, pkh) * scale
        w = torch.softmax(logits, dim=-1)  # [B,L,H,N]
        if stats is not None:  # (max, sum of exp) over the sequences given: merges sequence shards
            mx = logits.max(dim=-1).values
            stats.copy_(torch.stack([mx, torch.exp(logits - mx.unsqueeze(-1)).sum(-1)], dim=-1).to(stats.dtype))
        if w_out is not None:
            w_out.copy_(w.permute(0, 3, 1, 2).to(w_out.dtype))
        if qt is not None:
            qh = q.to(self.acc).reshape(B, N, L, H, dh)
            qs = qh * w.permute(0, 3, 1, 2).unsqueeze(-1) * q_scale  # [B,N,L,H,dh]
            qt.copy_(qs.permute(0, 3, 2, 1, 4).reshape(B, H, L, N * dh).to(qt.dtype))

    # operands of the outer-product sum + msa_1d (:469-482)
    def opm_prep(self, m, w, xt, yt, msa1d):
        B, N, L, P = m.shape
        mf = m.to(self.acc)
        x = mf.permute(0, 2, 3, 1).reshape(B, L * P, N)
        y = (mf * w.to(self.acc).reshape(B, N, L, 1)).permute(0, 2, 3, 1).reshape(B, L * P, N)
        xt.copy_(x.to(xt.dtype))
        yt.copy_(y.to(yt.dtype))
        msa1d.copy_(torch.cat([mf.sum(1), mf[:, 0]], dim=-1).to(msa1d.dtype))

    # Symmetrization + LayerNorm + Linear of pair2att, affine folded (:554-566)
    def pair2att_logits(self, pair, Wf, bf, eps, logits):
        p = pair.to(self.acc)
        s = 0.5 * (p + p.transpose(1, 2))
        mean = s.mean(-1, keepdim=True)
        var = s.var(-1, unbiased=False, keepdim=True)
        xh = (s - mean) / torch.sqrt(var + eps)
        lg = torch.einsum("bijd,cd->bcij", xh, Wf.to(self.acc)) + bf.to(self.acc).reshape(1, -1, 1, 1)
        logits.copy_(lg.to(logits.dtype))

    def pair2att_logits_rows(self, rows, cols_t, Wf, bf, eps, logits):
        s = 0.5 * (rows.to(self.acc) + cols_t.to(self.acc).transpose(1, 2))
        mean = s.mean(-1, keepdim=True)
        var = s.var(-1, unbiased=False, keepdim=True)
        xh = (s - mean) / torch.sqrt(var + eps)
        lg = torch.einsum("bijd,cd->bcij", xh, Wf.to(self.acc)) + bf.to(self.acc).reshape(1, -1, 1, 1)
        logits.copy_(lg.to(logits.dtype))

    # InstanceNorm2d statistics / apply (:453, :457) and the ELUs (:454, :462)
    def channel_stats(self, x, stats):
        xf = x.to(self.acc)
        stats[:, 0] += xf.sum(1).to(stats.dtype)
        stats[:, 1] += (xf * xf).sum(1).to(stats.dtype)

    def instnorm_apply(self, x, stats, gamma, beta, eps, res, elu, out):
        P = x.shape[1]
        st = stats.to(self.acc)
        mean = st[:, 0:1] / P
        var = (st[:, 1:2] / P - mean * mean).clamp_min(0)
        y = (x.to(self.acc) - mean) / torch.sqrt(var + eps) * gamma.to(self.acc) + beta.to(self.acc)
        if res is not None:
            y = y + res.to(self.acc)
        if elu:
            y = torch.nn.functional.elu(y)
        out.copy_(y.to(out.dtype))

    # performer_pytorch FastAttention (softmax_kernel / generalized_kernel + linear_attention)
    def favor_attention(self, q, k, v, out, proj, kind, heads):
        G1, G0, T, _ = q.shape
        m = proj.shape[0]

        def split(t):
            return t.to(self.acc).reshape(G1, G0, T, heads, 64).permute(0, 1, 3, 2, 4)  # g1 g0 h t d

        qh, kh, vh = split(q), split(k), split(v)
        P = proj.to(self.acc)
        dn = 64 ** -0.25
        uq = torch.einsum("...td,md->...tm", qh * dn, P)
        uk = torch.einsum("...td,md->...tm", kh * dn, P)
        if kind == 0:
            ratio = m ** -0.5
            dq = (qh ** 2).sum(-1, keepdim=True) / 2.0 * dn * dn
            dk = (kh ** 2).sum(-1, keepdim=True) / 2.0 * dn * dn
            qf = ratio * (torch.exp(uq - dq - uq.amax(dim=-1, keepdim=True)) + 1e-4)
            kf = ratio * (torch.exp(uk - dk - uk.amax(dim=(-1, -2), keepdim=True)) + 1e-4)
        else:
            qf = torch.relu(uq) + 1e-3
            kf = torch.relu(uk) + 1e-3
        ksum = kf.sum(dim=-2)
        dinv = 1.0 / torch.einsum("...tm,...m->...t", qf, ksum)
        ctx = torch.einsum("...tm,...te->...me", kf, vh)
        o = torch.einsum("...me,...tm,...t->...te", ctx, qf, dinv)  # g1 g0 h t d
        out.copy_(o.permute(0, 1, 3, 2, 4).reshape(G1, G0, T, heads * 64).to(out.dtype))

    # nn.Conv2d(C, C, 3, padding="same", bias=False) on b l1 l2 d (:451-457)
    def conv3x3(self, x, w_packed, out, dilation=1):
        Cout, _, cpad = w_packed.shape
        Cin = x.shape[3]
        w = w_packed.to(self.acc)[:, :, :Cin].reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
        y = torch.nn.functional.conv2d(x.to(self.acc).permute(0, 3, 1, 2), w, padding=dilation, dilation=dilation)
        out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))

    def conv3x3_f32(self, x, w_packed, out, dilation=1):
        _, Cin, Cout = w_packed.shape
        w = w_packed.to(self.acc).reshape(3, 3, Cin, Cout).permute(3, 2, 0, 1)
        y = torch.nn.functional.conv2d(x.to(self.acc).permute(0, 3, 1, 2), w, padding=dilation, dilation=dilation)
        out.copy_(y.permute(0, 2, 3, 1).to(out.dtype))

    # PredictionHead :1166 on a channels-last map
    def pair_symmetrize(self, x, out):
        xa = x.to(self.acc)
        out.copy_((0.5 * (xa + xa.transpose(1, 2))).to(out.dtype))

    # MsaEmbedding.forward (:114-120) / PairEmbedding.forward (:147-175) at the C-ABI boundary
    def msa_embed(self, tokens, aa_idx, emb, pos_enc, query_enc, out):
        B, N, L = tokens.shape
        q = torch.ones(N, dtype=torch.long)
        q[0] = 0
        out.copy_((emb[tokens] + pos_enc[aa_idx][:, None]) + query_enc[q][None, :, None, :])

    def pair_embed(self, seq, aa_idx, table_left, table_right, w_sep, bias, pos_enc_half, out):
        sep = torch.log((aa_idx[:, :, None] - aa_idx[:, None, :]).abs().float() + 1.0)[..., None]
        x = table_left[seq][:, None, :, :] + table_right[seq][:, :, None, :] + w_sep * sep + bias
        pe = pos_enc_half[aa_idx]
        L = seq.shape[1]
        out.copy_(x + torch.cat([pe[:, :, None, :].expand(-1, -1, L, -1), pe[:, None, :, :].expand(-1, L, -1, -1)], -1))

    def convert_rows(self, x, out):
        out.copy_(x.to(out.dtype))
