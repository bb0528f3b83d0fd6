"""CPU tests of the oracle itself: the functional restatement (oracle/trunk_ref.py) against the
golden vectors produced by the unmodified reference, and — in the build container, where
/root/reference exists — against the live reference at a second shape."""
import pytest
import torch

from oracle import reference_loader as rl
from oracle import trunk_ref
from oracle.weights import checksum, push_to_reference, synth_inputs, synth_state_dict
from tests.helpers import STAGES, build_block, load_golden, rel_l2, subset_stages


@pytest.mark.parametrize("name", ["two_track_small", "two_track_default", "two_track_tile_crossing"])
def test_restatement_matches_golden(name):
    gold = load_golden(name)
    cfg = gold["config"]
    _, sd, msa, pair = build_block(cfg)
    assert abs(checksum(sd) - gold["weight_checksum"]) < 1e-6 * gold["weight_checksum"], \
        "synthetic weights differ from the ones the fixture was generated with"
    stages = {}
    with torch.no_grad():
        trunk_ref.two_track_block(msa, pair, sd, cfg["n_layers"], stages=stages)
    stages = subset_stages(stages, cfg)
    for k in STAGES:
        assert rel_l2(stages[k], gold[k]) < 2e-5, k


@pytest.mark.skipif(not rl.available(), reason="reference source only exists in the build container")
def test_restatement_matches_live_reference():
    ref = rl.load()
    import rosettafold_pytorch_b200 as rf

    d_msa, d_pair, nl = 48, 40, 1
    mine = rf.TwoTrackBlock(d_msa, d_pair, n_encoder_layers=nl)
    sd = synth_state_dict(mine.state_dict(), seed=11)
    rblk = rl.fix_eval(push_to_reference(ref.TwoTrackBlock(d_msa, d_pair, n_encoder_layers=nl), sd))
    msa, pair = synth_inputs(1, 9, 13, d_msa, d_pair, seed=12)
    with torch.no_grad():
        m_ref, p_ref = rblk(msa, pair)
        m, p = trunk_ref.two_track_block(msa, pair, sd, nl)
    assert rel_l2(m, m_ref) < 2e-5 and rel_l2(p, p_ref) < 2e-5


def test_performer_restatement_properties():
    """FAVOR+ sanity: positive features, rows of the implied attention sum to one, and the
    softmax-kernel estimate approaches exact softmax attention as features grow."""
    from oracle.performer_ref import (gaussian_orthogonal_random_matrix, linear_attention,
                                      softmax_features)

    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 1, 16, 64, generator=g, dtype=torch.float64) * 0.5
    k = torch.randn(1, 1, 16, 64, generator=g, dtype=torch.float64) * 0.5
    v = torch.randn(1, 1, 16, 64, generator=g, dtype=torch.float64)
    proj = gaussian_orthogonal_random_matrix(4096, 64, generator=g).double()
    qf, kf = softmax_features(q, proj, True, eps=0.0), softmax_features(k, proj, False, eps=0.0)
    assert (qf > 0).all() and (kf > 0).all()
    ones = torch.ones_like(v[..., :1])
    assert torch.allclose(linear_attention(qf, kf, ones), ones)
    exact = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
    assert rel_l2(linear_attention(qf, kf, v), exact) < 0.15
    blocks = gaussian_orthogonal_random_matrix(128, 64, generator=g)
    unit = blocks / blocks.norm(dim=1, keepdim=True)
    assert torch.allclose(unit[:64] @ unit[:64].T, torch.eye(64), atol=1e-5)


def test_coord_restatement_matches_golden():
    """MsaUpdateWithPairAndCoord (:865-920): restatement vs the unmodified reference's output."""
    from tests.helpers import build_coord_module

    gold = load_golden("msa_pair_coord")
    _, sd, xyz, state, msa = build_coord_module(gold)
    with torch.no_grad():
        out = trunk_ref.msa_update_with_pair_and_coord(xyz, state, msa, trunk_ref.W(sd))
    assert rel_l2(out, gold["msa_out"]) < 2e-5


def test_performer_restatement_is_an_unbiased_softmax_attention_estimator():
    """`performer_pytorch` itself is absent (parity unpinned), so the restatement is anchored on what the published
    algorithm IS: a Monte-Carlo estimator of exp(q.k / sqrt(d)) attention. With the restated constants (inputs
    scaled by d^-1/4, the ||x||^2 / 2 term, Gaussian-orthogonal features) the result must converge to exact
    softmax attention like 1/sqrt(m); a wrong scale, sign or normaliser converges to something else (plateau)."""
    from oracle import performer_ref as P

    g = torch.Generator().manual_seed(1)
    B, H, T, d = 2, 3, 40, 64
    q = torch.randn(B, H, T, d, generator=g, dtype=torch.float64) * 0.3
    k = torch.randn(B, H, T, d, generator=g, dtype=torch.float64) * 0.3
    v = torch.randn(B, H, T, d, generator=g, dtype=torch.float64)
    exact = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v
    errs = []
    for m in (64, 256, 1024, 4096):
        e = []
        for s in range(6):
            proj = P.gaussian_orthogonal_random_matrix(m, d, generator=torch.Generator().manual_seed(10 + s)).double()
            e.append(float((P.favor_attention(q, k, v, proj, False) - exact).norm() / exact.norm()))
        errs.append(sum(e) / len(e))
    assert errs[-1] < 0.03, errs
    for a, b in zip(errs[:-1], errs[1:]):       # 4x the features -> about half the error
        assert 1.5 < a / b < 2.7, errs
    # the ReLU ("generalized") kernel is NOT a softmax estimator: it must not be confused with it
    proj = P.gaussian_orthogonal_random_matrix(4096, d, generator=torch.Generator().manual_seed(3)).double()
    gen = P.favor_attention(q, k, v, proj, True)
    assert float((gen - exact).norm() / exact.norm()) > 0.05


@pytest.mark.parametrize("generalized", [False, True])
def test_performer_restatement_matches_the_package_when_importable(generalized):
    """The pin the oracle lacks: wherever the real `performer_pytorch` can be imported (not in this image: the test
    skips, and the oracle header keeps saying "parity unpinned"), its SelfAttention with the restatement's weights and
    projection matrix must give the restatement's output. The state_dict keys are the package's own (the reference's
    checkpoints carry them), so loading is strict."""
    pp = pytest.importorskip("performer_pytorch")
    from oracle import performer_ref as P

    if "oracle" in (getattr(pp, "__file__", None) or "oracle") or pp.SelfAttention is P.SelfAttention:
        pytest.skip("`performer_pytorch` resolves to the oracle's own stand-in (oracle/shims), not to the real package")

    torch.manual_seed(5)
    dim, heads = 96, 3
    mine = P.SelfAttention(dim, heads=heads, generalized_attention=generalized).eval()
    kw = dict(generalized_attention=True, kernel_fn=torch.nn.ReLU()) if generalized else {}
    theirs = pp.SelfAttention(dim=dim, heads=heads, dropout=0.0, **kw).eval()
    assert theirs.fast_attention.nb_features == mine.fast_attention.nb_features == 266
    theirs.load_state_dict(mine.state_dict(), strict=True)
    x = torch.randn(2, 50, dim)
    with torch.no_grad():
        assert rel_l2(mine(x), theirs(x)) < 1e-5


@pytest.mark.parametrize("name", ["small_template", "default"])
def test_embedding_restatement_matches_golden(name):
    """oracle/embed_ref.py against the outputs of the unmodified reference's MsaEmbedding / PairEmbedding."""
    from oracle import embed_ref
    from tests.helpers import build_embeddings

    gold = load_golden("embeddings")[name]
    c = gold["config"]
    _, _, sd_m, sd_p, (tokens, seq, aa_idx, template) = build_embeddings(c)
    assert abs(checksum(sd_m) - gold["weight_checksums"][0]) < 1e-6 * gold["weight_checksums"][0]
    assert abs(checksum(sd_p) - gold["weight_checksums"][1]) < 1e-6 * gold["weight_checksums"][1]
    msa = embed_ref.msa_embedding(tokens, aa_idx, sd_m, c["max_len"])
    pair = embed_ref.pair_embedding(seq, aa_idx, sd_p, c["max_len"], template=template)
    assert torch.equal(msa, gold["msa"])          # integer gathers and two fp32 adds in the reference's order: exact
    assert rel_l2(pair, gold["pair"]) < 1e-6


@pytest.mark.parametrize("name", ["small", "default"])
def test_prediction_head_restatement_matches_golden(name):
    """oracle/heads_ref.py against the outputs of the unmodified reference's PredictionHead (four ResNets with
    dilations 1, 2, 4, 8; resnet.py, rosettafold_pytorch.py:1130-1172)."""
    from oracle import heads_ref
    from tests.helpers import build_heads

    gold = load_golden("prediction_head")[name]
    c = gold["config"]
    _, sd, pair = build_heads(c)
    with torch.no_grad():
        out = heads_ref.prediction_head(pair, sd, c["n_res_blocks"])
    for k in ("theta", "phi", "dist", "omega"):
        assert out[k].shape == gold[k].shape
        assert rel_l2(out[k], gold[k]) < 2e-5, k


@pytest.mark.parametrize("name", ["default", "narrow"])
def test_graph_transformer_restatement_matches_golden(name):
    """oracle/graph_ref.py against the unmodified reference's GraphTransformer / GraphTransformerBlock (:613-677)."""
    from oracle import graph_ref
    from tests.helpers import build_graph_block

    gold = load_golden("graph_transformer")[name]
    c = gold["config"]
    _, sd, (node, edge, mask) = build_graph_block(c)
    attn_sd = {k[len("attn."):]: v for k, v in sd.items() if k.startswith("attn.")}
    with torch.no_grad():
        assert rel_l2(graph_ref.graph_transformer(node, edge, None, attn_sd, c["n_heads"]), gold["attn"]) < 2e-5
        assert rel_l2(graph_ref.graph_transformer(node, edge, mask, attn_sd, c["n_heads"]), gold["attn_masked"]) < 2e-5
        assert rel_l2(graph_ref.graph_transformer_block(node, edge, None, sd, c["n_heads"]), gold["block"]) < 2e-5
        assert rel_l2(graph_ref.graph_transformer_block(node, edge, mask, sd, c["n_heads"]), gold["block_masked"]) < 2e-5
