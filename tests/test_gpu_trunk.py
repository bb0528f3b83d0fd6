"""GPU parity tests, module level: the b200 TwoTrackBlock (librfk kernels through the C ABI)
against the golden vectors of the unmodified reference and against the oracle restatement.

Tolerances are the north star's: relative L2 <= 1e-4 in the fp32 validation mode and <= 1e-2 in
the bf16 tensor-core mode, per trunk stage with each stage fed the reference's inputs AND for the free-running
chain of the whole block. The BASELINE.json shapes are in tests/test_gpu_baseline_configs.py."""
import json
import os

import pytest
import torch

import rosettafold_pytorch_b200 as rf
from oracle import trunk_ref
from tests.helpers import STAGES, build_block, load_golden, rel_l2, run_stages

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_mode():
    yield
    rf.set_mode("bf16")


@pytest.mark.parametrize("name", ["two_track_small", "two_track_default"])
def test_fp32_mode_matches_golden(cuda_device, name):
    gold = load_golden(name)
    blk, _, msa, pair = build_block(gold["config"], cuda_device)
    rf.set_mode("fp32")
    out = run_stages(blk, msa, pair)
    torch.cuda.synchronize()
    errs = {k: rel_l2(out[k], gold[k]) for k in STAGES}
    print("fp32 chain", name, errs)
    for k in STAGES:
        assert errs[k] < 1e-4, errs


@pytest.mark.parametrize("name", ["two_track_small", "two_track_default"])
def test_bf16_mode_matches_golden(cuda_device, name):
    gold = load_golden(name)
    blk, _, msa, pair = build_block(gold["config"], cuda_device)
    rf.set_mode("bf16")
    forced = run_stages(blk, msa, pair, teacher=gold)
    chain = run_stages(blk, msa, pair)
    torch.cuda.synchronize()
    e_forced = {k: rel_l2(forced[k], gold[k]) for k in STAGES}
    e_chain = {k: rel_l2(chain[k], gold[k]) for k in STAGES}
    print("bf16 teacher-forced", name, e_forced)
    print("bf16 chain", name, e_chain)
    for k in STAGES:
        assert e_forced[k] < 1e-2, e_forced
        assert e_chain[k] < 1e-2, e_chain


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_block_vs_restatement_mid_size(cuda_device, mode, tol):
    """Default widths, tile-crossing sizes (L=136 > one 128-row tile, N=40), 2 layers: whole block
    vs the CPU restatement."""
    cfg = dict(d_msa=384, d_pair=288, n_layers=2, B=1, N=40, L=136, seed=21)
    blk, sd, msa, pair = build_block(cfg, cuda_device)
    rf.set_mode(mode)
    m, p = blk(msa, pair)
    torch.cuda.synchronize()
    with torch.no_grad():
        m_ref, p_ref = trunk_ref.two_track_block(msa.cpu(), pair.cpu(), sd, cfg["n_layers"])
    e = (rel_l2(m, m_ref), rel_l2(p, p_ref))
    print(mode, "mid-size block msa/pair rel-l2", e)
    assert e[0] < tol and e[1] < tol, e


def test_block_is_deterministic_and_linear_in_batch(cuda_device):
    """Size-independent properties: samples are independent (batching == per-sample runs) and two
    runs agree bit-for-bit except for the atomics in the InstanceNorm statistics."""
    cfg = dict(d_msa=96, d_pair=72, n_layers=1, B=3, N=7, L=18, seed=5)
    blk, _, msa, pair = build_block(cfg, cuda_device)
    m, p = blk(msa, pair)
    m1, p1 = blk(msa[1:2].contiguous(), pair[1:2].contiguous())
    torch.cuda.synchronize()
    assert rel_l2(m[1:2], m1) < 1e-4 and rel_l2(p[1:2], p1) < 1e-4
    m2, p2 = blk(msa, pair)
    assert rel_l2(m2, m) < 1e-5 and rel_l2(p2, p) < 1e-5


def test_cuda_graph_replay_matches_eager(cuda_device):
    """GraphedModule: a captured TwoTrackBlock replays to the eager result, for new input values and
    after a second shape has been captured (one graph per input shape)."""
    cfg = dict(d_msa=96, d_pair=72, n_layers=1, B=1, N=9, L=40, seed=11)
    blk, _, msa, pair = build_block(cfg, cuda_device)
    gblk = rf.GraphedModule(blk)
    for scale in (1.0, 0.5):
        m_ref, p_ref = blk(msa * scale, pair * scale)
        m, p = gblk(msa * scale, pair * scale)
        torch.cuda.synchronize()
        assert rel_l2(m, m_ref) < 1e-6 and rel_l2(p, p_ref) < 1e-6
    msa2, pair2 = msa[:, :5, :24].contiguous(), pair[:, :24, :24].contiguous()
    m2_ref, p2_ref = blk(msa2, pair2)
    m2, p2 = gblk(msa2, pair2)
    torch.cuda.synchronize()
    assert rel_l2(m2, m2_ref) < 1e-6 and rel_l2(p2, p2_ref) < 1e-6
    m_ref, p_ref = blk(msa, pair)
    m, p = gblk(msa, pair)
    torch.cuda.synchronize()
    assert rel_l2(m, m_ref) < 1e-6 and rel_l2(p, p_ref) < 1e-6
    with pytest.raises(RuntimeError):
        gblk(msa.cpu(), pair.cpu())


def test_cuda_graph_follows_mode_and_weight_changes(cuda_device):
    """A captured graph has the numerics mode and the packed weights baked in: GraphedModule must re-capture when
    either changes instead of replaying the stale graph."""
    cfg = dict(d_msa=96, d_pair=72, n_layers=1, B=1, N=9, L=40, seed=12)
    blk, _, msa, pair = build_block(cfg, cuda_device)
    gblk = rf.GraphedModule(blk)
    m16, p16 = (t.clone() for t in gblk(msa, pair))
    rf.set_mode("fp32")
    m_ref, p_ref = blk(msa, pair)
    m32, p32 = gblk(msa, pair)
    torch.cuda.synchronize()
    assert rel_l2(m32, m_ref) < 1e-6 and rel_l2(p32, p_ref) < 1e-6
    assert rel_l2(m32, m16) > 1e-6  # (the two modes do differ: the check above is not vacuous)
    rf.set_mode("bf16")
    with torch.no_grad():
        blk.msa_update_with_pair.encoder_layers[0].ff.fn[1].net[3].bias.add_(0.5)
    m_ref, p_ref = blk(msa, pair)
    m, p = gblk(msa, pair)
    torch.cuda.synchronize()
    assert rel_l2(m, m_ref) < 1e-6 and rel_l2(m, m16) > 1e-3


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_msa_update_with_pair_and_coord_vs_golden(cuda_device, mode, tol):
    """MsaUpdateWithPairAndCoord (:865-920) through librfk vs the unmodified reference's output."""
    from tests.helpers import build_coord_module

    gold = load_golden("msa_pair_coord")
    mod, _, xyz, state, msa = build_coord_module(gold, cuda_device)
    rf.set_mode(mode)
    out = mod(xyz, state, msa)
    torch.cuda.synchronize()
    e = rel_l2(out, gold["msa_out"])
    print(mode, "MsaUpdateWithPairAndCoord rel-l2", e)
    assert e < tol, e


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("name", ["two_track_blocks.0", "three_track_blocks.0", "final_block"])
def test_blocks_in_situ_match_model_trace(cuda_device, name, mode, tol):
    """The trunk blocks in situ: inputs, coordinates and outputs recorded by hooks during a forward of the whole
    unmodified reference model (README widths; tests/golden/model_trace.pt), replayed through librfk."""
    from tests.helpers import build_trace_block, run_trace_block

    blk, coord, rec = build_trace_block(load_golden("model_trace"), name, cuda_device)
    rf.set_mode(mode)
    errs = run_trace_block(blk, coord, rec)
    torch.cuda.synchronize()
    print(mode, name, errs)
    assert max(errs.values()) < tol, errs


@pytest.mark.parametrize("name", ["small_template", "default"])
def test_embeddings_vs_golden(cuda_device, name):
    """MsaEmbedding / PairEmbedding (:106-181) on the device - where the reference cannot run them - against the
    unmodified reference's CPU outputs: the integer gathers exactly, the fp32 sums to 1e-6; template path through the
    LayerNorm / GEMM kernels in both modes."""
    from tests.helpers import build_embeddings

    gold = load_golden("embeddings")[name]
    m, p, _, _, (tokens, seq, aa_idx, template) = build_embeddings(gold["config"], cuda_device)
    rf.set_mode("fp32")
    msa = m(tokens.to(cuda_device), aa_idx.to(cuda_device))
    pair = p(seq, aa_idx, template) if template is not None else p(seq, aa_idx)  # CPU inputs are moved by the module
    torch.cuda.synchronize()
    assert msa.is_cuda and pair.is_cuda
    assert torch.equal(msa.cpu(), gold["msa"])
    assert rel_l2(pair, gold["pair"]) < 1e-6
    if template is not None:
        rf.set_mode("bf16")
        pair16 = p(seq, aa_idx, template)
        torch.cuda.synchronize()
        assert rel_l2(pair16, gold["pair"]) < 1e-2
    with pytest.raises(IndexError):
        m(tokens + 100, aa_idx)


def test_embeddings_large_shape_vs_oracle(cuda_device):
    """Embeddings at the metric shape (1, 128, 512), default widths, residue indices up to max_len - 1, against the CPU
    restatement (100 MB + 302 MB outputs: checked here, not stored)."""
    from oracle import embed_ref
    from oracle.weights import synth_state_dict

    B, N, L, max_len = 1, 128, 512, 5000
    m = rf.MsaEmbedding(21, 384, max_len).eval()
    p = rf.PairEmbedding(21, 288, max_len).eval()
    sd_m, sd_p = synth_state_dict(m.state_dict(), seed=80), synth_state_dict(p.state_dict(), seed=81)
    m.load_state_dict(sd_m)
    p.load_state_dict(sd_p)
    m, p = m.to(cuda_device), p.to(cuda_device)
    g = torch.Generator().manual_seed(82)
    tokens, seq = torch.randint(0, 21, (B, N, L), generator=g), torch.randint(0, 21, (B, L), generator=g)
    aa_idx = torch.sort(torch.randperm(max_len, generator=g)[:L]).values.repeat(B, 1)
    msa, pair = m(tokens, aa_idx), p(seq, aa_idx)
    torch.cuda.synchronize()
    assert torch.equal(msa.cpu(), embed_ref.msa_embedding(tokens, aa_idx, sd_m, max_len))
    assert rel_l2(pair, embed_ref.pair_embedding(seq, aa_idx, sd_p, max_len)) < 1e-6


@pytest.mark.parametrize("name", ["small", "default"])
def test_prediction_head_vs_golden(cuda_device, name):
    """PredictionHead (:1130-1172; four ResNets, dilations 1 / 2 / 4 / 8) on librfk against the unmodified reference's
    outputs: fp32 validation mode <= 1e-4, tensor-core mode <= 1e-2."""
    from tests.helpers import build_heads

    gold = load_golden("prediction_head")[name]
    head, _, pair = build_heads(gold["config"], cuda_device)
    pair = pair.to(cuda_device)
    for mode, tol_ in (("fp32", 1e-4), ("bf16", 1e-2)):
        rf.set_mode(mode)
        out = head(pair)
        torch.cuda.synchronize()
        for k in ("theta", "phi", "dist", "omega"):
            e = rel_l2(out[k], gold[k])
            assert e < tol_, f"{mode} {k}: rel-l2 {e}"
    rf.set_mode("bf16")


def test_prediction_head_tile_crossing_vs_oracle(cuda_device):
    """Default width at L = 136 (every convolution tile boundary crossed, dilation 8 reaching across it) against the CPU
    restatement (pinned to the reference by tests/test_oracle.py), tensor-core mode."""
    from oracle import heads_ref
    from oracle.weights import synth_state_dict

    C, n_blocks, L = 288, 4, 136
    head = rf.PredictionHead(C, n_blocks, 0.1).eval()
    sd = synth_state_dict(head.state_dict(), seed=95)
    head.load_state_dict(sd)
    head = head.to(cuda_device)
    pair = torch.randn((1, L, L, C), generator=torch.Generator().manual_seed(96))
    rf.set_mode("bf16")
    out = head(pair.to(cuda_device))
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = heads_ref.prediction_head(pair, sd, n_blocks)
    for k in ("theta", "phi", "dist", "omega"):
        e = rel_l2(out[k], ref[k])
        assert e < 1e-2, f"{k}: rel-l2 {e}"


def test_tokens_to_logits_pipeline_vs_oracle(cuda_device):
    """Everything this package replaces, composed on the device: MsaEmbedding / PairEmbedding -> two-track trunk blocks ->
    PredictionHead, from integer tokens to distance / orientation logits, against the same composition of the CPU
    restatements (each pinned to the unmodified reference separately). This is the part of RoseTTAFold.forward
    (:1273-1289) that does not need the SE(3) structure track; it measures how the trunk's 16-bit rounding arrives in
    the heads' logits (row g of the verdict's table, on hardware)."""
    from oracle import embed_ref, heads_ref, trunk_ref
    from oracle.weights import synth_state_dict

    B, N, L, max_len, n_blocks, n_layers, n_res = 1, 20, 72, 400, 2, 1, 4
    emb_m, emb_p = rf.MsaEmbedding(21, 384, max_len).eval(), rf.PairEmbedding(21, 288, max_len).eval()
    trunk = rf.TrunkBlocks(384, 288, n_blocks=n_blocks, n_encoder_layers=n_layers).eval()
    head = rf.PredictionHead(288, n_res, 0.1).eval()
    sds = []
    for i, m in enumerate((emb_m, emb_p, trunk, head)):
        sd = synth_state_dict(m.state_dict(), seed=300 + i)
        m.load_state_dict(sd)
        m.to(cuda_device)
        sds.append(sd)
    g = torch.Generator().manual_seed(310)
    tokens, seq = torch.randint(0, 21, (B, N, L), generator=g), torch.randint(0, 21, (B, L), generator=g)
    aa_idx = torch.arange(L).repeat(B, 1)
    aa_idx[:, L // 2:] += 30
    with torch.no_grad():
        msa_r = embed_ref.msa_embedding(tokens, aa_idx, sds[0], max_len)
        pair_r = embed_ref.pair_embedding(seq, aa_idx, sds[1], max_len)
        for b in range(n_blocks):
            msa_r, pair_r = trunk_ref.two_track_block(msa_r, pair_r, sds[2], n_layers, prefix=f"blocks.{b}.")
        ref = heads_ref.prediction_head(pair_r, sds[3], n_res)
    errs = {}
    for mode, tol_ in (("fp32", 1e-4), ("bf16", 1e-2)):
        rf.set_mode(mode)
        msa, pair = trunk(emb_m(tokens, aa_idx), emb_p(seq, aa_idx))
        out = head(pair)
        torch.cuda.synchronize()
        errs[mode] = dict(msa=rel_l2(msa, msa_r), pair=rel_l2(pair, pair_r), **{k: rel_l2(out[k], ref[k]) for k in ref})
        assert max(errs[mode].values()) < tol_, (mode, errs[mode])
    rf.set_mode("bf16")
    print("tokens -> logits:", errs)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_r02.jsonl", "a") as f:
        f.write(json.dumps(dict(test="tokens_to_logits_pipeline", shape=[B, N, L], blocks=n_blocks, errs=errs)) + "\n")


@pytest.mark.parametrize("name", ["default", "narrow"])
def test_graph_transformer_vs_golden(cuda_device, name):
    """GraphTransformer / GraphTransformerBlock (:613-677) on librfk against the unmodified reference, with and without an
    edge mask: fp32 validation mode <= 1e-4, tensor-core mode <= 1e-2."""
    from tests.helpers import build_graph_block

    gold = load_golden("graph_transformer")[name]
    blk, _, (node, edge, mask) = build_graph_block(gold["config"], cuda_device)
    node, edge, mask = node.to(cuda_device), edge.to(cuda_device), mask.to(cuda_device)
    for mode, tol_ in (("fp32", 1e-4), ("bf16", 1e-2)):
        rf.set_mode(mode)
        outs = dict(attn=blk.attn(node, edge), attn_masked=blk.attn(node, edge, mask), block=blk(node, edge, None),
                    block_masked=blk(node, edge, mask))
        torch.cuda.synchronize()
        for k, v in outs.items():
            e = rel_l2(v, gold[k])
            assert e < tol_, f"{mode} {k}: rel-l2 {e}"
    rf.set_mode("bf16")


def test_graph_transformer_large_vs_oracle(cuda_device):
    """L = 200 (tile-crossing, not a multiple of 8), model widths, against the CPU restatement, tensor-core mode."""
    from oracle import graph_ref
    from oracle.weights import synth_state_dict

    B, L, Dn, d, De, H = 1, 200, 64, 64, 64, 4
    blk = rf.GraphTransformerBlock(Dn, d, De, H).eval()
    sd = synth_state_dict(blk.state_dict(), seed=120)
    blk.load_state_dict(sd)
    blk = blk.to(cuda_device)
    g = torch.Generator().manual_seed(121)
    node, edge = torch.randn((B, L, Dn), generator=g), torch.randn((B, L, L, De), generator=g)
    rf.set_mode("bf16")
    out = blk(node.to(cuda_device), edge.to(cuda_device), None)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = graph_ref.graph_transformer_block(node, edge, None, sd, H)
    e = rel_l2(out, ref)
    assert e < 1e-2, e
