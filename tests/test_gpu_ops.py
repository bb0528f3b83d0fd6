"""GPU parity tests, op level: every librfk kernel through the C ABI vs the oracle restatement
(oracle/ops_ref.py) on the same seeded inputs.

Tolerances: fp32 (validation-mode) kernels rel-L2 <= 1e-5; bf16 tensor-core kernels are compared
against the oracle evaluated on the SAME bf16-rounded inputs, so only accumulation order and the
final output rounding differ: rel-L2 <= 4e-3 for bf16 outputs, <= 1e-5 for f32 outputs.
"""
import math

import pytest
import torch

import rosettafold_pytorch_b200 as rf
from oracle.ops_ref import RefBackend
from rosettafold_pytorch_b200 import ops

pytestmark = pytest.mark.gpu
REF = RefBackend()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def tol(dtype):
    """Output rounding of the op: bf16 8, IEEE half 11 significand bits."""
    return 4e-3 if dtype == torch.bfloat16 else (5e-4 if dtype == torch.float16 else 2e-5)


OPERAND_DTYPES = [torch.bfloat16, torch.float16, torch.float32]  # tcgen05 bf16, tcgen05 f16, SIMT fp32


def _rand(shape, dtype, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(dev)


def _gemm_case(dev, dtype, Z, M, N, K, *, out_dtype, bias, act, res, lda_pad=0, alpha=1.0):
    a = _rand((*Z, M, K + lda_pad), dtype, dev, 1)[..., :K]
    b = _rand((*Z, N, K + lda_pad), dtype, dev, 2)[..., :K]
    c = torch.empty((*Z, M, N), dtype=out_dtype, device=dev)
    cv = c.reshape(*([1] * (3 - len(Z))), *Z, 1, M, 1, N)
    bias_t = _rand((N,), torch.float32, dev, 3) if bias else None
    r0 = _rand((*Z, M, N), torch.float32, dev, 4) if res else None
    r0v = None if r0 is None else r0.reshape(cv.shape)
    ops.gemm(a, b, cv, bias=bias_t, act=act, alpha=alpha, r0=r0v)
    c_ref = torch.empty_like(c)
    al, bl = a, b
    while al.dim() < 5:
        al, bl = al.unsqueeze(0), bl.unsqueeze(0)
    REF.gemm(al, bl, c_ref.reshape(cv.shape), bias_t, act, alpha, r0v, None, 0, None, None, 1e-5)
    torch.cuda.synchronize()
    return rel_l2(c, c_ref)


@pytest.mark.parametrize("dtype", OPERAND_DTYPES)
@pytest.mark.parametrize("shape", [
    ((), 128, 128, 64), ((), 256, 256, 128), ((), 384, 288, 384), ((), 1000, 384, 200),
    ((), 130, 1536, 384), ((), 512, 1152, 288), ((), 64, 32, 384), ((3,), 200, 96, 72),
    ((2, 3), 128, 512, 136), ((), 4096, 768, 384), ((), 77, 50, 33),
])
def test_gemm_plain(cuda_device, dtype, shape):
    Z, M, N, K = shape
    pad = (-K) % 8
    e = _gemm_case(cuda_device, dtype, Z, M, N, K, out_dtype=torch.float32, bias=False,
                   act=ops.ACT_NONE, res=False, lda_pad=pad)
    assert e < 2e-5, f"rel-l2 {e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_long_k_tied_logits(cuda_device, dtype):
    """The tied-logit contraction at the metric shape (1, 128, 512): per head 512 x 512 with K = N*32 = 4096
    (64 k-blocks through the TMA ring), and the A.V shape 512 x 4096 x 512, f32 and bf16 outputs."""
    scale = 4096 ** -0.25
    for Z, M, N, K, out_dtype in (((12,), 512, 512, 4096, torch.float32), ((12,), 512, 4096, 512, torch.bfloat16),
                                  ((3,), 200, 136, 4096 + 64, torch.float32)):
        if dtype == torch.float32 and out_dtype == torch.bfloat16:
            continue
        dev = cuda_device
        a = _rand((*Z, M, K), dtype, dev, 21, scale)
        b = _rand((*Z, N, K), dtype, dev, 22, scale)
        c = torch.empty((*Z, M, N), dtype=out_dtype, device=dev)
        ops.gemm(a, b, c.reshape(1, 1, *Z, 1, M, 1, N))
        ref = torch.einsum("zmk,znk->zmn", a.double(), b.double())
        torch.cuda.synchronize()
        e = rel_l2(c, ref)
        assert e < (tol(torch.bfloat16) if out_dtype == torch.bfloat16 else 2e-5), f"{(Z, M, N, K)}: rel-l2 {e}"


@pytest.mark.parametrize("shape", [(40000, 768, 384), (33000, 1536, 288), (2 * 17000, 384, 384)])
def test_gemm_large_m_bf16_out(cuda_device, shape):
    """Large-M, short-K projections with bf16 output take the TMA-store epilogue path
    (many waves of tiles per CTA, ragged last m-block)."""
    M, N, K = shape
    for act, bias in ((ops.ACT_NONE, False), (ops.ACT_RELU, True)):
        e = _gemm_case(cuda_device, torch.bfloat16, (), M, N, K, out_dtype=torch.bfloat16, bias=bias,
                       act=act, res=False)
        assert e < tol(torch.bfloat16), f"rel-l2 {e}"


@pytest.mark.parametrize("dtype", OPERAND_DTYPES)
def test_gemm_epilogues(cuda_device, dtype):
    for out_dtype in (torch.float32, torch.bfloat16, torch.float16):
        e = _gemm_case(cuda_device, dtype, (2,), 300, 384, 384, out_dtype=out_dtype, bias=True,
                       act=ops.ACT_RELU, res=True, alpha=0.5)
        assert e < tol(out_dtype), f"{out_dtype}: rel-l2 {e}"


@pytest.mark.parametrize("dtype", OPERAND_DTYPES)
def test_gemm_scatter_views(cuda_device, dtype):
    """The K^T / V^T relayouts of the tied row attention: rows (b, n, l), columns (h, d)."""
    dev = cuda_device
    B, N, L, H, dh = 2, 5, 24, 3, 32
    D = H * dh
    x = _rand((B, N * L, D), dtype, dev, 5)
    w = _rand((D, D), dtype, dev, 6, 0.1)
    bias = _rand((D,), torch.float32, dev, 7)
    kt = torch.zeros((B, H, L, N * dh), dtype=dtype, device=dev)
    vt = torch.zeros((B, H, N * dh, L), dtype=dtype, device=dev)
    kt_view = kt.view(B, H, L, N, dh).permute(0, 3, 2, 1, 4)[None, None]  # [1,1,B,N,L,H,dh]
    vt_view = vt.view(B, H, N, dh, L).permute(0, 2, 4, 1, 3)[None, None]
    ops.gemm(x, w[None], kt_view, bias=bias)
    ops.gemm(x, w[None], vt_view, bias=bias)
    ref = (x.double() @ w.double().T + bias.double()).view(B, N, L, H, dh)
    torch.cuda.synchronize()
    assert rel_l2(kt.view(B, H, L, N, dh), ref.permute(0, 3, 2, 1, 4)) < tol(dtype)
    assert rel_l2(vt.view(B, H, N, dh, L), ref.permute(0, 3, 1, 4, 2)) < tol(dtype)


def test_gemm_broadcast_residuals(cuda_device):
    """Linear(716->288) with the rank-1 row/column terms as broadcast addends (:484-496)."""
    dev = cuda_device
    B, L, P, K = 2, 20, 288, 592
    f = _rand((B, L * L, K), torch.bfloat16, dev, 8)[..., :588]
    w = _rand((P, K), torch.bfloat16, dev, 9, 0.05)[..., :588]
    R = _rand((B, L, P), torch.float32, dev, 10)
    Cc = _rand((B, L, P), torch.float32, dev, 11)
    bias = _rand((P,), torch.float32, dev, 12)
    h = torch.empty((B, L, L, P), dtype=torch.float32, device=dev)
    hv = h.view(1, 1, B, L, L, 1, P)
    r0 = R.view(1, 1, B, L, 1, 1, P).expand(1, 1, B, L, L, 1, P)
    r1 = Cc.view(1, 1, B, 1, L, 1, P).expand(1, 1, B, L, L, 1, P)
    ops.gemm(f, w[None], hv, bias=bias, r0=r0, r1=r1)
    ref = (f.double() @ w.double().T + bias.double()).view(B, L, L, P) + R.double()[:, :, None] + Cc.double()[:, None]
    torch.cuda.synchronize()
    assert rel_l2(h, ref) < 2e-5


def test_gemm_blockln32(cuda_device):
    """Outer-product sum + LayerNorm(1024) fused in the GEMM epilogue (:424-425, :416)."""
    dev = cuda_device
    B, L, N, P = 2, 12, 40, 32
    xt = _rand((B, L * P, N), torch.bfloat16, dev, 13)
    yt = _rand((B, L * P, N), torch.bfloat16, dev, 14, 0.1)
    g = _rand((1024,), torch.float32, dev, 15) + 1.0
    bt = _rand((1024,), torch.float32, dev, 16)
    o = torch.empty((B, L, L, P * P), dtype=torch.bfloat16, device=dev)
    ov = o.view(B, L, L, P, P).permute(0, 1, 3, 2, 4)[None, None]
    ops.gemm(xt, yt, ov, epi=ops.EPI_BLOCKLN32, ln_gamma=g, ln_beta=bt, ln_eps=1e-5)
    x = xt.double().view(B, L, P, N).permute(0, 3, 1, 2)  # b n i u
    y = yt.double().view(B, L, P, N).permute(0, 3, 1, 2)
    op = torch.einsum("bniu,bnjv->bijuv", x, y).reshape(B, L, L, P * P)
    ref = torch.nn.functional.layer_norm(op, (1024,), g.double(), bt.double(), 1e-5)
    torch.cuda.synchronize()
    assert rel_l2(o, ref) < tol(torch.bfloat16)


@pytest.mark.parametrize("D", [32, 288, 384, 1024, 100])
@pytest.mark.parametrize("dtypes", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32),
                                    (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float16),
                                    (torch.float16, torch.float16)])
def test_layernorm(cuda_device, D, dtypes):
    dev = cuda_device
    xi, yo = dtypes
    rows = 777
    x = _rand((rows, D), xi, dev, 20, 3.0) + 0.5
    g = _rand((D,), torch.float32, dev, 21)
    b = _rand((D,), torch.float32, dev, 22)
    wide = torch.zeros((rows, D + 72), dtype=yo, device=dev)
    out = wide[:, 8:8 + D]
    ops.layernorm(x, g, b, 1e-5, out)
    ref = torch.empty((rows, D), dtype=yo, device=dev)
    REF.layernorm(x, g, b, 1e-5, ref)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < tol(yo)
    assert float(wide[:, :8].abs().max()) == 0 and float(wide[:, 8 + D:].abs().max()) == 0


def test_softmax_and_symmetrize(cuda_device):
    dev = cuda_device
    B, H, L = 2, 12, 70
    Lp = 72
    logits = _rand((B * H * L, L), torch.float32, dev, 30, 4.0)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        A = torch.zeros((B, H, L, Lp), dtype=dt, device=dev)
        ops.softmax_rows(logits, A.view(B * H * L, Lp)[:, :L])
        ref = torch.softmax(logits.double(), -1).view(B, H, L, L)
        torch.cuda.synchronize()
        assert rel_l2(A[..., :L], ref) < tol(dt)
        att = torch.empty((B, L, L, H), dtype=torch.float32, device=dev)
        ops.tied_att_symmetrize(A[..., :L], att)
        a = A[..., :L].double()
        ref_att = (0.5 * (a + a.transpose(-1, -2))).permute(0, 2, 3, 1)
        torch.cuda.synchronize()
        assert rel_l2(att, ref_att) < 1e-6


@pytest.mark.parametrize("dtype", OPERAND_DTYPES)
@pytest.mark.parametrize("cfg", [(2, 7, 20, 12, 32), (1, 40, 16, 1, 32), (1, 130, 33, 12, 32), (2, 64, 9, 3, 16),
                                 (1, 5, 6, 2, 12)])
def test_poswise_weight(cuda_device, dtype, cfg):
    dev = cuda_device
    B, N, L, H, dh = cfg
    D = H * dh
    pq = _rand((B, L, D), dtype, dev, 40)
    buf = _rand((B, N, L, 2 * D), dtype, dev, 41)
    q, pk = buf[..., :D], buf[..., D:]
    w = torch.empty((B, N, L, H), dtype=torch.float32, device=dev)
    qt = torch.empty((B, H, L, N * dh), dtype=dtype, device=dev)
    ops.poswise_weight(pq, pk, dh ** -0.5, w_out=w, q=q, q_scale=dh ** -0.5, qt=qt, heads=H, d_head=dh)
    w_ref, qt_ref = torch.empty_like(w), torch.empty_like(qt)
    REF.poswise_weight(pq, pk, dh ** -0.5, w_ref, q, dh ** -0.5, qt_ref, H, dh)
    torch.cuda.synchronize()
    assert rel_l2(w, w_ref) < 1e-5
    assert abs(float(w.sum(1).mean()) - 1.0) < 1e-5  # reference test_module.py:180-200
    assert rel_l2(qt, qt_ref) < tol(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_poswise_weight_sequence_shards_merge(cuda_device, dtype):
    """The (max, sum) statistics of rfk_poswise_weight_stats merge two sequence shards into the softmax over
    all sequences (the identity the sequence-sharded tied row layers rely on)."""
    dev = cuda_device
    B, N, L, H, dh = 2, 12, 9, 12, 32
    D = H * dh
    pq = _rand((B, L, D), dtype, dev, 45, 2.0)
    pk = _rand((B, N, L, D), dtype, dev, 46, 2.0)
    w_full = torch.empty((B, N, L, H), dtype=torch.float32, device=dev)
    st_full = torch.empty((B, L, H, 2), dtype=torch.float32, device=dev)
    ops.poswise_weight(pq, pk, dh ** -0.5, w_out=w_full, heads=H, d_head=dh, stats=st_full)
    st_ref, w_ref = torch.empty_like(st_full), torch.empty_like(w_full)
    REF.poswise_weight(pq, pk, dh ** -0.5, w_ref, None, 1.0, None, H, dh, st_ref)
    parts, stats = [], []
    for lo, hi in ((0, 5), (5, 12)):  # ragged split
        pk_s = pk[:, lo:hi].contiguous()
        w = torch.empty((B, hi - lo, L, H), dtype=torch.float32, device=dev)
        st = torch.empty((B, L, H, 2), dtype=torch.float32, device=dev)
        ops.poswise_weight(pq, pk_s, dh ** -0.5, w_out=w, heads=H, d_head=dh, stats=st)
        parts.append(w)
        stats.append(st)
    torch.cuda.synchronize()
    assert rel_l2(st_full, st_ref) < 1e-5
    allst = torch.stack(stats)
    gmax = allst[..., 0].max(0).values
    gsum = (allst[..., 1] * torch.exp(allst[..., 0] - gmax)).sum(0)
    merged = torch.cat([w * (st[..., 1] * torch.exp(st[..., 0] - gmax) / gsum).unsqueeze(1)
                        for w, st in zip(parts, stats)], dim=1)
    assert rel_l2(merged, w_full) < 1e-5


@pytest.mark.parametrize("dtype", OPERAND_DTYPES)
def test_opm_prep(cuda_device, dtype):
    dev = cuda_device
    B, N, L, P = 2, 37, 9, 32
    m = _rand((B, N, L, P), torch.float32, dev, 50)
    w = torch.softmax(_rand((B, N, L), torch.float32, dev, 51), dim=1).contiguous()
    Np = 40
    xt = torch.zeros((B, L * P, Np), dtype=dtype, device=dev)
    yt = torch.zeros((B, L * P, Np), dtype=dtype, device=dev)
    msa1d = torch.empty((B, L, 2 * P), dtype=torch.float32, device=dev)
    ops.opm_prep(m, w, xt[..., :N], yt[..., :N], msa1d)
    xr, yr, mr = torch.empty_like(xt[..., :N]), torch.empty_like(yt[..., :N]), torch.empty_like(msa1d)
    REF.opm_prep(m, w, xr, yr, mr)
    torch.cuda.synchronize()
    assert rel_l2(xt[..., :N], xr) < 1e-6 and rel_l2(yt[..., :N], yr) < tol(dtype)
    assert rel_l2(msa1d, mr) < 1e-6


def test_pair2att_logits(cuda_device):
    dev = cuda_device
    B, L, D, Cn = 2, 33, 288, 16
    pair = _rand((B, L, L, D), torch.float32, dev, 60, 2.0)
    Wf = _rand((Cn, D), torch.float32, dev, 61, 0.1)
    bf = _rand((Cn,), torch.float32, dev, 62)
    Lp = 40
    lg = torch.zeros((B, Cn, L, Lp), dtype=torch.float32, device=dev)
    ops.pair2att_logits(pair, Wf, bf, 1e-5, lg[..., :L])
    ref = torch.empty((B, Cn, L, L), dtype=torch.float32, device=dev)
    REF.pair2att_logits(pair, Wf, bf, 1e-5, ref)
    torch.cuda.synchronize()
    assert rel_l2(lg[..., :L], ref) < 1e-5


@pytest.mark.parametrize("D", [288, 72])
def test_pair2att_logits_rows(cuda_device, D):
    """Row-sharded pair2att (long-protein path): rows + transposed shard of two ragged row ranges reproduce the
    whole-map kernel (vectorised D = 288 path and the generic one)."""
    dev = cuda_device
    B, L, Cn = 2, 33, 16
    pair = _rand((B, L, L, D), torch.float32, dev, 63, 2.0)
    Wf = _rand((Cn, D), torch.float32, dev, 64, 0.1)
    bf = _rand((Cn,), torch.float32, dev, 65)
    full = torch.empty((B, Cn, L, L), dtype=torch.float32, device=dev)
    ops.pair2att_logits(pair, Wf, bf, 1e-5, full)
    for lo, hi in ((0, 13), (13, 33)):
        rows = pair[:, lo:hi].contiguous()
        cols_t = pair[:, :, lo:hi].contiguous()
        Lp = 40
        part = torch.zeros((B, Cn, hi - lo, Lp), dtype=torch.float32, device=dev)
        ops.pair2att_logits_rows(rows, cols_t, Wf, bf, 1e-5, part[..., :L])
        ref = torch.empty((B, Cn, hi - lo, L), dtype=torch.float32, device=dev)
        REF.pair2att_logits_rows(rows, cols_t, Wf, bf, 1e-5, ref)
        torch.cuda.synchronize()
        assert rel_l2(part[..., :L], ref) < 1e-5
        assert rel_l2(part[..., :L], full[:, :, lo:hi]) < 1e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(2, 900, 288), (1, 77, 72), (3, 333, 20)])
def test_instnorm(cuda_device, dtype, shape):
    """InstanceNorm2d + residual + ELU (:453-462): vectorised path (C % 8 == 0), ragged chunks, and
    the scalar path (C = 20); both output dtypes."""
    dev = cuda_device
    B, P, Cn = shape
    x = _rand((B, P, Cn), dtype, dev, 70, 2.0) + 0.3
    res = _rand((B, P, Cn), torch.float32, dev, 71)
    g = _rand((Cn,), torch.float32, dev, 72)
    b = _rand((Cn,), torch.float32, dev, 73)
    stats = torch.zeros((B, 2, Cn), dtype=torch.float64, device=dev)
    ops.channel_stats(x, stats)
    out = torch.empty((B, P, Cn), dtype=torch.float32, device=dev)
    ops.instnorm_apply(x, stats, g, b, 1e-6, out, res=res, elu=True)
    xn = torch.nn.functional.instance_norm(x.double().permute(0, 2, 1), weight=g.double(), bias=b.double(), eps=1e-6)
    ref = torch.nn.functional.elu(xn.permute(0, 2, 1) + res.double())
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-4
    out16 = torch.empty((B, P, Cn), dtype=torch.bfloat16, device=dev)
    ops.instnorm_apply(x, stats, g, b, 1e-6, out16, elu=True)
    ref16 = torch.nn.functional.elu(xn.permute(0, 2, 1))
    torch.cuda.synchronize()
    assert rel_l2(out16, ref16) < tol(torch.bfloat16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("cfg", [(1, 3, 40, 2), (2, 2, 128, 3), (1, 2, 200, 2), (1, 40, 512, 8), (2, 32, 128, 12),
                                 (1, 70, 300, 3), (1, 200, 97, 4), (1, 300, 7, 2)])
def test_favor_attention(cuda_device, dtype, kind, cfg):
    """FAVOR+ linear attention vs the performer restatement; strided token axis as in RowWise."""
    dev = cuda_device
    G1, G0, T, H = cfg
    inner = H * 64
    g = torch.Generator().manual_seed(80)
    proj = torch.randn(266, 64, generator=g).to(dev)
    # token axis strided: buffer [G1, T, G0, 3*inner] viewed as [G1, G0, T, .]
    buf = _rand((G1, T, G0, 3 * inner), dtype, dev, 81, 0.7)
    view = buf.permute(0, 2, 1, 3)
    q, k, v = view[..., :inner], view[..., inner:2 * inner], view[..., 2 * inner:]
    out = torch.zeros((G1, T, G0, inner), dtype=dtype, device=dev).permute(0, 2, 1, 3)
    ops.favor_attention(q, k, v, out, proj, kind=kind, heads=H)
    ref = torch.empty_like(out)
    REF.favor_attention(q, k, v, ref, proj, kind, H)
    torch.cuda.synchronize()
    e = rel_l2(out, ref)
    # the 16-bit kernels round the features and the context once more: bf16 <= 1e-2, IEEE half <= 1.5e-3
    assert e < {torch.float32: 1e-4, torch.bfloat16: 1e-2, torch.float16: 1.5e-3}[dtype], f"rel-l2 {e}"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("D", [96, 384, 30])
def test_layernorm_residual(cuda_device, dtype, D):
    """y = LayerNorm(x) + res (:916): vector path (D % 4 == 0) and generic path (D = 30)."""
    dev = cuda_device
    x = _rand((77, D), torch.float32, dev, 110, 3.0) + 0.5
    res = _rand((77, D), torch.float32, dev, 111)
    g, b = _rand((D,), torch.float32, dev, 112) + 1.0, _rand((D,), torch.float32, dev, 113)
    out = torch.empty((77, D), dtype=dtype, device=dev)
    ops.layernorm(x, g, b, 1e-5, out, res=res)
    ref = torch.nn.functional.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-5) + res.double()
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < tol(dtype)


def test_dist_mask_logits(cuda_device):
    """-1e9 wherever the C-alpha distance is not below the head's threshold (:899-913)."""
    dev = cuda_device
    B, H, L, Lp = 2, 4, 37, 40
    g = torch.Generator().manual_seed(120)
    xyz = (torch.randn((B, L, 3, 3), generator=g) * 8).to(dev)
    bins = torch.tensor([8.0, 12.0, 16.0, 20.0], device=dev)
    logits = _rand((B, H, L, Lp), torch.float32, dev, 121)
    ref = logits.clone()
    ops.dist_mask_logits(xyz[:, :, 1], bins, logits[..., :L])
    pd = torch.cdist(xyz[:, :, 1], xyz[:, :, 1])
    for h in range(H):
        ref[:, h, :, :L] += (1.0 - (pd < bins[h]).float()) * -1e9
    torch.cuda.synchronize()
    assert torch.equal(logits[..., L:], ref[..., L:])  # padding untouched
    assert rel_l2(logits[..., :L], ref[..., :L]) < 1e-6
    assert int((logits[..., :L] < -1e8).sum()) == int((ref[..., :L] < -1e8).sum()) > 0


def test_conv3x3_rectangular_row_shard(cuda_device):
    """H x W images (a row shard of the pair map plus halo rows): rows 1..H-2 of the convolution of the
    haloed shard equal the same rows of the convolution of the whole map."""
    dev = cuda_device
    B, L, Cin, Cout = 1, 40, 72, 96
    x = _rand((B, L, L, Cin), torch.bfloat16, dev, 95)
    w = _rand((Cout, Cin, 3, 3), torch.float32, dev, 96, (9 * Cin) ** -0.5)
    wp = ops.pack_conv3x3_weight(w)
    full = torch.empty((B, L, L, Cout), dtype=torch.float32, device=dev)
    ops.conv3x3(x, wp, full)
    lo, hi = 10, 23
    shard = x[:, lo - 1:hi + 1].contiguous()          # 13 rows + one halo row on either side
    part = torch.empty((B, hi - lo + 2, L, Cout), dtype=torch.float32, device=dev)
    ops.conv3x3(shard, wp, part)
    torch.cuda.synchronize()
    assert rel_l2(part[:, 1:-1], full[:, lo:hi]) < 1e-6


def test_launch_counter_and_errors(cuda_device):
    n0 = rf._lib.launch_count()
    x = torch.randn(8, 32, device=cuda_device)
    ops.layernorm(x, None, None, 1e-5, torch.empty_like(x))
    assert rf._lib.launch_count() == n0 + 1
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.randn(8, 32), None, None, 1e-5, torch.empty(8, 32))  # CPU tensor: no fallback
    assert math.isfinite(float(x.sum()))


@pytest.mark.parametrize("cfg", [(1, 24, 288, 288), (2, 130, 72, 96), (1, 256, 288, 288), (2, 20, 72, 72)])
def test_conv3x3_implicit_gemm(cuda_device, cfg):
    """3x3 'same' conv as implicit GEMM vs F.conv2d on the same bf16-rounded inputs
    (reference :452, :456); L=130 exercises the padded-row tiles, C=72 the short last K block."""
    dev = cuda_device
    B, L, Cin, Cout = cfg
    x = _rand((B, L, L, Cin), torch.bfloat16, dev, 90)
    w = _rand((Cout, Cin, 3, 3), torch.float32, dev, 91, (9 * Cin) ** -0.5)
    wp = ops.pack_conv3x3_weight(w)
    for odt in (torch.bfloat16, torch.float32):
        out = torch.empty((B, L, L, Cout), dtype=odt, device=dev)
        ops.conv3x3(x, wp, out)
        ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), padding=1)
        torch.cuda.synchronize()
        e = rel_l2(out, ref.permute(0, 2, 3, 1))
        assert e < tol(odt), f"{odt}: rel-l2 {e}"


@pytest.mark.parametrize("h16", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("cols,ys", [(288, 584), (288, 288), (12, 584), (100, 100)])
def test_convert_rows(cuda_device, cols, ys, h16):
    """Row-wise dtype conversion into a strided destination: 8-wide vector path and the scalar path."""
    dev = cuda_device
    rows = 1000
    x = _rand((rows, cols), torch.float32, dev, 90)
    y = torch.zeros((rows, ys), dtype=h16, device=dev)
    ops.convert_rows(x, y[:, ys - cols:] if (ys - cols) % 8 == 0 else y[:, :cols])
    torch.cuda.synchronize()
    got = y[:, ys - cols:] if (ys - cols) % 8 == 0 else y[:, :cols]
    assert torch.equal(got, x.to(h16))
    back = torch.empty((rows, cols), dtype=torch.float32, device=dev)
    ops.convert_rows(got, back)
    torch.cuda.synchronize()
    assert torch.equal(back, got.float())


@pytest.mark.parametrize("dilation", [2, 4, 8])
@pytest.mark.parametrize("cfg", [(1, 28, 288, 288), (2, 130, 72, 96)])
def test_conv3x3_dilated(cuda_device, cfg, dilation):
    """Dilated 3x3 'same' convolution (ResBlock2D of the prediction heads, resnet.py:19-37): the tcgen05 implicit GEMM
    shifts its nine TMA boxes by the dilation (zero fill outside the image), the fp32 SIMT kernel widens its window;
    both against F.conv2d. L = 130 crosses the 128-column tile, dilation 8 reaches past it."""
    dev = cuda_device
    B, L, Cin, Cout = cfg
    x = _rand((B, L, L, Cin), torch.bfloat16, dev, 190)
    w = _rand((Cout, Cin, 3, 3), torch.float32, dev, 191, (9 * Cin) ** -0.5)
    out = torch.empty((B, L, L, Cout), dtype=torch.float32, device=dev)
    ops.conv3x3(x, ops.pack_conv3x3_weight(w), out, dilation)
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), padding=dilation,
                                     dilation=dilation).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 1e-5
    x32 = x.float()
    out32 = torch.empty_like(out)
    ops.conv3x3_f32(x32, ops.pack_conv3x3_weight_f32(w), out32, dilation)
    ref32 = torch.nn.functional.conv2d(x32.double().permute(0, 3, 1, 2), w.double(), padding=dilation,
                                       dilation=dilation).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert rel_l2(out32, ref32) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_pair_symmetrize(cuda_device, dtype):
    """0.5 (x + x^T) over the two residue axes of a channels-last map (PredictionHead :1166)."""
    dev = cuda_device
    x = _rand((2, 37, 37, 72), dtype, dev, 195)
    out = ops.pair_symmetrize(x, torch.empty_like(x))
    torch.cuda.synchronize()
    ref = 0.5 * (x.double() + x.double().transpose(1, 2))
    assert rel_l2(out, ref) < tol(dtype)
    assert torch.equal(out, out.transpose(1, 2))
    with pytest.raises(ValueError):
        ops.pair_symmetrize(x, x)
