"""CPU tests of the host logic: the b200 nn.Modules (layout views, weight packing, residual
wiring, API surface) run with the oracle's per-op emulation swapped in for librfk and are
compared with the golden vectors of the unmodified reference. No CUDA kernel runs here; the
`-m gpu` tests repeat the comparison through the real library."""
import ctypes

import pytest
import torch

import rosettafold_pytorch_b200 as rf
from oracle.ops_ref import RefBackend
from rosettafold_pytorch_b200 import _lib, ops
from tests.helpers import STAGES, build_block, load_golden, rel_l2, run_stages


@pytest.fixture()
def emulated_ops():
    prev = ops._set_backend_for_tests(RefBackend())
    yield
    ops._set_backend_for_tests(prev)
    rf.set_mode("bf16")


@pytest.mark.parametrize("name", ["two_track_small", "two_track_default"])
def test_block_fp32_mode_matches_golden(emulated_ops, name):
    gold = load_golden(name)
    blk, _, msa, pair = build_block(gold["config"])
    rf.set_mode("fp32")
    out = run_stages(blk, msa, pair)
    for k in STAGES:
        assert rel_l2(out[k], gold[k]) < 1e-4, k  # north-star fp32-validation tolerance


def test_block_bf16_storage_within_budget(emulated_ops):
    """bf16 operand storage with exact accumulation: the rounding budget of the bf16 mode."""
    gold = load_golden("two_track_small")
    blk, _, msa, pair = build_block(gold["config"])
    rf.set_mode("bf16")
    out = run_stages(blk, msa, pair, teacher=gold)
    for k in STAGES:
        assert rel_l2(out[k], gold[k]) < 1e-2, k  # north-star bf16 tolerance


def test_state_dict_keys_cover_reference_layout():
    blk = rf.TwoTrackBlock(96, 72, n_encoder_layers=1)
    keys = set(blk.state_dict())
    for k in ["msa_update_using_self_att.residue_wise_encoder_layers.0.attn.poswise_weight.to_q.0.weight",
              "msa_update_using_self_att.sequence_wise_encoder_layers.0.attn.fast_attention.projection_matrix",
              "pair_update_with_msa.resnet.1.fn.5.weight",
              "pair_update_with_axial_attention.layers.0.layer.1.fn.1.fn.to_out.bias",
              "pair_update_with_axial_attention.layers.0.row_attn.to_q.weight",
              "msa_update_with_pair.encoder_layers.0.pair2att.2.weight"]:
        assert k in keys, k
    assert blk.msa_update_using_self_att.sequence_wise_encoder_layers[0].attn.fast_attention.projection_matrix.shape == (266, 64)


def test_error_conventions():
    with pytest.raises(AssertionError):
        rf.PositionWiseWeightFactor(d_msa=100, n_heads=12)  # reference :188-190
    with pytest.raises(AssertionError):
        rf.SoftTiedAttentionOverResidues(d_msa=100, n_heads=12)  # :223-225
    with pytest.raises(NotImplementedError):
        rf.EncoderLayer(tied=False, performer=False)  # :319-320
    with pytest.raises(NotImplementedError):
        rf.EncoderLayer(performer=True, return_att=True)  # :309-312


def test_public_module_shapes(emulated_ops):
    """Shape contract of reference tests/test_module.py (:180-200, :216-231, :293-309, :322-339,
    :390-402, :425-438, :645-661)."""
    rf.set_mode("fp32")
    B, N, L = 2, 4, 10
    x = torch.randn(B, N, L, 96)
    w = rf.PositionWiseWeightFactor(96, 12)(x)
    assert w.shape == (B, N, 12, L, 1) and torch.allclose(w.sum(1), torch.ones(B, 12, L, 1), atol=1e-5)
    out, att = rf.SoftTiedAttentionOverResidues(96, 12, return_att=True)(x)
    assert out.shape == x.shape and att.shape == (B, L, L, 12)
    assert torch.allclose(att, att.transpose(1, 2), atol=1e-6)
    assert rf.EncoderLayer(96, 384, 12, performer=True)(x).shape == x.shape
    y, att = rf.MsaUpdateUsingSelfAttention(96, 384, 12, n_encoder_layers=1)(x)
    assert y.shape == x.shape and att.shape == (B, L, L, 12)
    p = rf.OuterProductMean(32, 72)(torch.randn(B, N, L, 32))
    assert p.shape == (B, L, L, 72)
    pair = torch.randn(B, L, L, 72)
    assert rf.PairUpdateWithMsa(96, 32, 72, 12)(x, pair, att).shape == pair.shape
    assert rf.PairUpdateWithAxialAttention(72, 288, 8, 0.1, 1)(pair).shape == pair.shape
    assert rf.MsaUpdateWithPair(96, 72, 4, n_encoder_layers=2)(x, pair).shape == x.shape
    s = rf.Symmetrization()(pair)
    assert torch.equal(s, s.transpose(1, 2))
    m, pr = rf.TwoTrackBlock(96, 72, 1)(x, pair)
    assert m.shape == x.shape and pr.shape == pair.shape


def test_c_abi_exports_every_declared_symbol():
    """librfk.so loads and exports every function include/rfk.h declares (no compute calls)."""
    import os
    import re

    hdr = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include", "rfk.h")).read()
    declared = set(re.findall(r"\b(rfk_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().rfk_strerror(0) == b"ok" and _lib.load().rfk_version() >= 1


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.randn(4, 32), None, None, 1e-5, torch.empty(4, 32))


@pytest.mark.skipif(not __import__("oracle.reference_loader", fromlist=["x"]).available(),
                    reason="reference source only exists in the build container")
def test_accelerate_swaps_trunk_of_reference_block(emulated_ops):
    """the reference TwoTrackBlock.forward runs unchanged on the swapped-in trunk."""
    from oracle import reference_loader as rl
    from oracle.weights import push_to_reference, synth_inputs, synth_state_dict

    ref = rl.load()
    mine = rf.TwoTrackBlock(48, 40, n_encoder_layers=1)
    sd = synth_state_dict(mine.state_dict(), seed=31)
    rblk = rl.fix_eval(push_to_reference(ref.TwoTrackBlock(48, 40, n_encoder_layers=1), sd))
    msa, pair = synth_inputs(1, 5, 12, 48, 40, seed=32)
    with torch.no_grad():
        m_ref, p_ref = rblk(msa, pair)
    rf.set_mode("fp32")
    rf.accelerate(rblk)
    assert type(rblk.msa_update_with_pair).__module__.startswith("rosettafold_pytorch_b200")
    with torch.no_grad():
        m, p = rblk(msa, pair)  # the reference's own forward (:962-968) on the b200 modules
    assert rel_l2(m, m_ref) < 1e-4 and rel_l2(p, p_ref) < 1e-4


def test_msa_update_with_pair_and_coord_fp32_matches_golden(emulated_ops):
    """Host logic of the three-track MSA update (:865-920) against the reference's golden output."""
    from tests.helpers import build_coord_module

    gold = load_golden("msa_pair_coord")
    mod, _, xyz, state, msa = build_coord_module(gold)
    rf.set_mode("fp32")
    out = mod(xyz, state, msa)
    assert out.shape == gold["msa_out"].shape
    assert rel_l2(out, gold["msa_out"]) < 1e-4
    assert set(mod.state_dict()) == {f"{m}.{p}" for m in ("ln_msa", "ln_state", "to_q", "to_k", "to_v", "ln_out",
                                                          "to_out.fn.0", "to_out.fn.1.net.0", "to_out.fn.1.net.3")
                                     for p in ("weight", "bias")}


@pytest.mark.skipif(not __import__("oracle.reference_loader", fromlist=["x"]).available(),
                    reason="reference source only exists in the build container")
def test_accelerate_swaps_coord_update_of_three_track_block(emulated_ops):
    """The MSA update of a reference ThreeTrackBlock (:1028-1035, called at :1044) is swapped too and
    reproduces the reference module on the same inputs."""
    from oracle import reference_loader as rl
    from oracle.make_golden import synth_coords
    from oracle.weights import synth_inputs, synth_state_dict

    ref = rl.load()
    rmod = ref.MsaUpdateWithPairAndCoord(d_msa=48, d_state=16, d_trfm_inner=32, d_ff=96).eval()
    rmod.load_state_dict(synth_state_dict(rmod.state_dict(), seed=41))
    mine = rf.TwoTrackBlock(48, 40, n_encoder_layers=1)
    rblk = rl.fix_eval(ref.TwoTrackBlock(48, 40, n_encoder_layers=1))
    rblk.msa_update_with_pair_and_coord = rmod  # the attribute a ThreeTrackBlock carries
    msa, _ = synth_inputs(1, 5, 12, 48, 8, seed=42)
    xyz, state = synth_coords(1, 12, 16, 43)
    with torch.no_grad():
        want = rmod(xyz, state, msa)
    rf.set_mode("fp32")
    rf.accelerate_block(rblk)
    got = rblk.msa_update_with_pair_and_coord
    assert type(got).__module__.startswith("rosettafold_pytorch_b200")
    assert rel_l2(got(xyz, state, msa), want) < 1e-4


@pytest.mark.skipif(not __import__("oracle.reference_loader", fromlist=["x"]).available(),
                    reason="reference source only exists in the build container")
def test_accelerate_whole_reference_model(emulated_ops, tmp_path, monkeypatch):
    """Drop-in at the top: the unmodified reference `RoseTTAFold` (:1160-1271; two-track blocks, three-track
    blocks with their SE(3) structure track on the oracle's dgl / lie_learn shims, output heads) gives the same
    logits, coordinates and pLDDT before and after `rf.accelerate(model)` swaps the trunk under it."""
    from oracle import reference_loader as rl

    monkeypatch.chdir(tmp_path)  # the reference caches its SE(3) bases under ./cache
    ref = rl.load()
    torch.manual_seed(0)
    model = rl.fix_eval(ref.RoseTTAFold(d_input=21, d_msa=48, d_pair=40, d_node=16, d_edge=16, d_state=16,
                                        n_two_track_blocks=1, n_three_track_blocks=2, n_encoder_layers=1,
                                        n_neighbors=[8, 8], p_dropout=0.1, max_len=64))
    g = torch.Generator().manual_seed(1234)
    B, N, L = 1, 4, 12
    msa, seq = torch.randint(0, 21, (B, N, L), generator=g), torch.randint(0, 21, (B, L), generator=g)
    aa_idx = torch.arange(L).repeat(B, 1)
    with torch.no_grad():
        logits, xyz, plddt = model(msa, seq, aa_idx)
    rf.set_mode("fp32")
    rf.accelerate(model)
    swapped = [n for n, m in model.named_modules() if type(m).__module__.startswith("rosettafold_pytorch_b200")]
    for stage in ("msa_update_using_self_att", "pair_update_with_msa", "pair_update_with_axial_attention",
                  "msa_update_with_pair", "msa_update_with_pair_and_coord"):
        assert any(n.endswith(stage) for n in swapped), stage
    with torch.no_grad():
        logits2, xyz2, plddt2 = model(msa, seq, aa_idx)
    for k in logits:
        assert rel_l2(logits2[k], logits[k]) < 1e-4, k
    assert rel_l2(xyz2, xyz) < 1e-4 and rel_l2(plddt2, plddt) < 1e-4


@pytest.mark.skipif(not __import__("oracle.reference_loader", fromlist=["x"]).available(),
                    reason="reference source only exists in the build container")
def test_whole_model_drift_with_16bit_operands(emulated_ops, tmp_path, monkeypatch, capsys):
    """Row (g) of the scope table, bounded on the CPU: the whole unmodified reference model (README widths: d_msa 384,
    d_pair 288; 2 two-track + 2 three-track blocks + final block, kNN graphs with n_neighbors < L so that a trunk
    perturbation CAN flip a neighbour) with its trunk swapped for the b200 modules in the TENSOR-CORE mode, the ops
    emulated by oracle/ops_ref.py on tensors stored in the kernels' operand formats (bf16 / IEEE half).

    Asserted: what the trunk hands to the final block stays inside the 1e-2 budget. Measured and printed (DESIGN.md
    section 4 quotes them): the four logits, xyz and pLDDT of the whole model. Those sit DOWNSTREAM of reference code
    that amplifies any perturbation with random-init weights (InstanceNorm ResNet heads: x2.5; the SE(3) track behind
    a top-k kNN graph: x10 and more, SURVEY.md section 7 "hard parts"), so for them only a sanity bound is asserted;
    the fp32 validation mode reproduces all of them to 1e-4 (test_accelerate_whole_reference_model)."""
    from oracle import reference_loader as rl

    monkeypatch.chdir(tmp_path)
    ref = rl.load()
    torch.manual_seed(0)
    model = rl.fix_eval(ref.RoseTTAFold(d_input=21, d_msa=384, d_pair=288, d_node=32, d_edge=32, d_state=32,
                                        n_two_track_blocks=2, n_three_track_blocks=3, n_encoder_layers=2,
                                        n_neighbors=[12, 12], p_dropout=0.1, max_len=64))
    g = torch.Generator().manual_seed(1234)
    B, N, L = 1, 6, 24
    msa, seq = torch.randint(0, 21, (B, N, L), generator=g), torch.randint(0, 21, (B, L), generator=g)
    aa_idx = torch.arange(L).repeat(B, 1)
    seen = []
    model.final_block.register_forward_pre_hook(lambda m, a: seen.append((a[0].clone(), a[1].clone())))
    with torch.no_grad():
        logits, xyz, plddt = model(msa, seq, aa_idx)
    rf.set_mode("bf16")
    rf.accelerate(model)
    with torch.no_grad():
        logits2, xyz2, plddt2 = model(msa, seq, aa_idx)
    trunk = dict(msa=rel_l2(seen[1][0], seen[0][0]), pair=rel_l2(seen[1][1], seen[0][1]))
    errs = {k: rel_l2(logits2[k], logits[k]) for k in logits}
    errs.update(xyz=rel_l2(xyz2, xyz), plddt=rel_l2(plddt2, plddt))
    with capsys.disabled():
        print("\nwhole model, 16-bit operand emulation vs reference: trunk into the final block",
              {k: f"{v:.2e}" for k, v in trunk.items()}, "model outputs", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(trunk.values()) < 1e-2, trunk
    assert max(errs[k] for k in logits) < 5e-2 and errs["xyz"] < 0.5 and errs["plddt"] < 0.5, errs


@pytest.mark.parametrize("name", ["two_track_blocks.0", "three_track_blocks.0", "final_block"])
def test_blocks_in_situ_match_model_trace(emulated_ops, name):
    """Host logic on the activations / coordinates of a whole reference-model forward (tests/golden/model_trace.pt)."""
    from tests.helpers import build_trace_block, run_trace_block

    blk, coord, rec = build_trace_block(load_golden("model_trace"), name)
    rf.set_mode("fp32")
    errs = run_trace_block(blk, coord, rec)
    assert max(errs.values()) < 1e-4, errs


def test_single_residue_is_refused_like_the_reference(emulated_ops):
    """L = 1: the reference's InstanceNorm2d (:453) raises ValueError on a 1 x 1 map; so does the replacement."""
    from oracle import trunk_ref

    cfg = dict(d_msa=48, d_pair=40, n_layers=1, B=2, N=3, L=1, seed=23)
    blk, sd, msa, pair = build_block(cfg)
    rf.set_mode("fp32")
    with pytest.raises(ValueError, match="more than 1 spatial element"):
        trunk_ref.two_track_block(msa, pair, sd, 1)
    with pytest.raises(ValueError, match="more than 1 spatial element"):
        blk(msa, pair)


@pytest.mark.parametrize("B,N,L", [(1, 1, 9), (2, 3, 2), (1, 2, 8), (3, 9, 17)])
def test_block_degenerate_shapes_match_restatement(emulated_ops, B, N, L):
    """Host logic at the edges the reference accepts: a single sequence (the query row of the position-wise
    weights is the only row), two residues (3x3 convolution on a 2 x 2 map), batches > 1, sizes below / across the
    8-element padding of the operand buffers."""
    from oracle import trunk_ref

    cfg = dict(d_msa=48, d_pair=40, n_layers=1, B=B, N=N, L=L, seed=23)
    blk, sd, msa, pair = build_block(cfg)
    rf.set_mode("fp32")
    m, p = blk(msa, pair)
    with torch.no_grad():
        m_ref, p_ref = trunk_ref.two_track_block(msa, pair, sd, cfg["n_layers"])
    assert m.shape == msa.shape and p.shape == pair.shape
    assert rel_l2(m, m_ref) < 1e-4 and rel_l2(p, p_ref) < 1e-4


@pytest.mark.skipif(not __import__("oracle.reference_loader", fromlist=["x"]).available(),
                    reason="reference source only exists in the build container")
def test_accelerate_with_device_hops_keeps_structure_and_results(emulated_ops, tmp_path, monkeypatch):
    """`rf.accelerate(model, device, hop=True)` (for the CPU-only reference model with the trunk on a GPU): the
    forward pre-hooks sit on the swapped modules and on the reference modules that consume their outputs, the
    module tree and `state_dict` keys are untouched, results are unchanged. (Here both devices are the CPU: this
    checks the wiring, not a transfer.)"""
    from oracle import reference_loader as rl

    monkeypatch.chdir(tmp_path)  # the reference caches its SE(3) bases under ./cache
    ref = rl.load()
    torch.manual_seed(0)
    model = rl.fix_eval(ref.RoseTTAFold(d_input=21, d_msa=48, d_pair=40, d_node=16, d_edge=16, d_state=16,
                                        n_two_track_blocks=1, n_three_track_blocks=2, n_encoder_layers=1,
                                        n_neighbors=[8, 8], p_dropout=0.1, max_len=64))
    g = torch.Generator().manual_seed(1234)
    msa, seq = torch.randint(0, 21, (1, 4, 12), generator=g), torch.randint(0, 21, (1, 12), generator=g)
    aa_idx = torch.arange(12).repeat(1, 1)
    with torch.no_grad():
        logits, xyz, plddt = model(msa, seq, aa_idx)
    rf.set_mode("fp32")
    rf.accelerate(model, device="cpu")
    keys = set(model.state_dict())
    rf.accelerate_block(model.final_block, device="cpu", hop=True)   # idempotent per module
    rf.accelerate(model, device="cpu", hop=True)
    assert set(model.state_dict()) == keys
    hopped = {n for n, m in model.named_modules() if getattr(m, "_rfk_hop", False)}
    assert {"msa_emb", "pair_emb", "initial_coord_generation_with_msa_and_pair", "prediction_head",
            "final_block.plddt_head", "final_block.coord_update_with_msa_and_pair",
            "three_track_blocks.0.coord_update_with_msa_and_pair", "three_track_blocks.0.msa_update_with_pair_and_coord",
            "two_track_blocks.0.msa_update_using_self_att", "two_track_blocks.0.msa_update_with_pair"} <= hopped
    assert not any(n in hopped for n in ("two_track_blocks", "three_track_blocks", "final_block", ""))
    with torch.no_grad():
        logits2, xyz2, plddt2 = model(msa, seq, aa_idx)
    for k in logits:
        assert rel_l2(logits2[k], logits[k]) < 1e-4
    assert rel_l2(xyz2, xyz) < 1e-4 and rel_l2(plddt2, plddt) < 1e-4


@pytest.mark.parametrize("name", ["small_template", "default"])
def test_embeddings_match_golden(emulated_ops, name):
    """Host logic of the embedding modules (table packing of the split Linear, template path, index handling)
    against the unmodified reference's outputs; the kernels themselves are checked by `-m gpu`."""
    from tests.helpers import build_embeddings

    gold = load_golden("embeddings")[name]
    m, p, _, _, (tokens, seq, aa_idx, template) = build_embeddings(gold["config"])
    rf.set_mode("fp32")
    assert rel_l2(m(tokens, aa_idx), gold["msa"]) < 1e-6
    out = p(seq, aa_idx, template) if template is not None else p(seq, aa_idx)
    assert rel_l2(out, gold["pair"]) < 1e-6
    with pytest.raises(IndexError):
        m(tokens + 100, aa_idx)
    if template is None:
        with pytest.raises(ValueError):
            p(seq, aa_idx, torch.zeros(1))


@pytest.mark.parametrize("name", ["small", "default"])
def test_prediction_head_matches_golden(emulated_ops, name):
    """Host logic of PredictionHead / ResNet / ResBlock2D (channels-last bookkeeping, weight packing of the 1 x 1 and
    dilated 3 x 3 convolutions, residual / InstanceNorm wiring, symmetrised input of the dist / omega heads) against
    the unmodified reference's outputs; the kernels themselves are checked by `-m gpu`."""
    from tests.helpers import build_heads

    gold = load_golden("prediction_head")[name]
    head, _, pair = build_heads(gold["config"])
    rf.set_mode("fp32")
    out = head(pair)
    for k in ("theta", "phi", "dist", "omega"):
        assert rel_l2(out[k], gold[k]) < 1e-4, k
    # the NCHW entry points of the building blocks (API parity with resnet.py)
    rn = head.theta_head[0]
    x = torch.randn(1, gold["config"]["in_channels"], 9, 9)
    assert rn(x).shape == (1, 37, 9, 9)
    assert rn.layer[3](x).shape == x.shape
    with pytest.raises(ValueError):
        head(pair[:, :1, :1])


@pytest.mark.parametrize("name", ["default", "narrow"])
def test_graph_transformer_matches_golden(emulated_ops, name):
    """Host logic of GraphTransformer(Block): the strided batched GEMM views that replace the reference's einsums over the
    materialised edge embedding, the folded scale, the additive mask, ELU + residual epilogue."""
    from tests.helpers import build_graph_block

    gold = load_golden("graph_transformer")[name]
    blk, _, (node, edge, mask) = build_graph_block(gold["config"])
    rf.set_mode("fp32")
    assert rel_l2(blk.attn(node, edge), gold["attn"]) < 1e-5
    assert rel_l2(blk.attn(node, edge, mask), gold["attn_masked"]) < 1e-5
    assert rel_l2(blk(node, edge, None), gold["block"]) < 1e-5
    assert rel_l2(blk(node, edge, mask), gold["block_masked"]) < 1e-5
    with pytest.raises(ValueError):
        blk.attn(node, edge[:, :, :-1])
