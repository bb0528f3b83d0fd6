"""Shared test helpers: golden fixtures, deterministic weights, error metrics."""
import os

import torch

from oracle.weights import checksum, synth_inputs, synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGES = ("msa_a", "att", "pair_b", "pair_c", "msa_d")


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def subset_stages(stages, cfg):
    """Restrict full stage outputs to the rows a subset fixture holds (oracle/make_golden.py SUBSET_CONFIGS);
    a config without row lists keeps the whole tensors."""
    if "msa_rows" not in cfg:
        return stages
    out = {}
    for k, v in stages.items():
        idx = torch.tensor(cfg["msa_rows"] if k.startswith("msa") else cfg["pair_rows"], device=v.device)
        out[k] = v.index_select(1, idx)
    return out


def build_block(cfg, device="cpu"):
    """A b200 TwoTrackBlock carrying the fixture's synthetic weights, and the fixture inputs."""
    import rosettafold_pytorch_b200 as rf

    blk = rf.TwoTrackBlock(cfg["d_msa"], cfg["d_pair"], n_encoder_layers=cfg["n_layers"]).eval()
    sd = synth_state_dict(blk.state_dict(), seed=cfg["seed"])
    blk.load_state_dict(sd, strict=True)
    msa, pair = synth_inputs(cfg["B"], cfg["N"], cfg["L"], cfg["d_msa"], cfg["d_pair"], seed=cfg["seed"] + 100)
    return blk.to(device), sd, msa.to(device), pair.to(device)


def run_stages(blk, msa, pair, teacher=None):
    """Run the four trunk calls; with `teacher` (a golden dict) each stage gets the reference's
    inputs (teacher forcing), otherwise the block's own chain."""
    out = {}
    m, att = blk.msa_update_using_self_att(msa)
    out["msa_a"], out["att"] = m, att
    if teacher is not None:
        m, att = teacher["msa_a"].to(msa.device), teacher["att"].to(msa.device)
    p1 = blk.pair_update_with_msa(m, pair, att)
    out["pair_b"] = p1
    if teacher is not None:
        p1 = teacher["pair_b"].to(msa.device)
    p2 = blk.pair_update_with_axial_attention(p1)
    out["pair_c"] = p2
    if teacher is not None:
        p2 = teacher["pair_c"].to(msa.device)
    out["msa_d"] = blk.msa_update_with_pair(m, p2)
    return out


def build_coord_module(gold, device="cpu"):
    """The b200 MsaUpdateWithPairAndCoord with the fixture's synthetic weights, and its inputs."""
    import rosettafold_pytorch_b200 as rf

    c = gold["config"]
    mod = rf.MsaUpdateWithPairAndCoord(c["d_msa"], c["d_state"], c["d_inner"], c["d_ff"]).eval()
    sd = synth_state_dict(mod.state_dict(), seed=c["seed"])
    assert abs(checksum(sd) - gold["weight_checksum"]) < 1e-6 * gold["weight_checksum"]
    mod.load_state_dict(sd, strict=True)
    msa, _ = synth_inputs(c["B"], c["N"], c["L"], c["d_msa"], 8, seed=c["seed"] + 100)
    return mod.to(device), sd, gold["xyz"].to(device), gold["state"].to(device), msa.to(device)


TRACE_BLOCKS = ("two_track_blocks.0", "three_track_blocks.0", "final_block")


def build_trace_block(gold, name, device="cpu"):
    """The b200 TwoTrackBlock (+ MsaUpdateWithPairAndCoord for a three-track block) carrying the synthetic
    weights that block `name` of the traced reference model held (oracle/make_golden.py --trace-only)."""
    import rosettafold_pytorch_b200 as rf

    c = gold["config"]
    k = TRACE_BLOCKS.index(name)
    blk = rf.TwoTrackBlock(c["d_msa"], c["d_pair"], n_encoder_layers=c["n_encoder_layers"]).eval()
    sd = synth_state_dict(blk.state_dict(), seed=c["seed"] + k)
    assert abs(checksum(sd) - gold["weight_checksums"][name]) < 1e-6 * gold["weight_checksums"][name]
    blk.load_state_dict(sd, strict=True)
    coord = None
    if "xyz" in gold["blocks"][name]:
        coord = rf.MsaUpdateWithPairAndCoord(c["d_msa"], c["d_state"], 32, 4 * c["d_msa"]).eval()
        coord.load_state_dict(synth_state_dict(coord.state_dict(), seed=c["seed"] + 10 + k), strict=True)
        coord = coord.to(device)
    return blk.to(device), coord, {n: v.to(device) for n, v in gold["blocks"][name].items()}


def run_trace_block(blk, coord, rec):
    """-> errors of the block's trunk outputs (and the coordinate-conditioned MSA update) against the trace."""
    msa, pair = blk(rec["msa_in"], rec["pair_in"])
    errs = dict(msa=rel_l2(msa, rec["msa_trunk_out"]), pair=rel_l2(pair, rec["pair_out"]))
    if coord is not None:
        errs["msa_coord"] = rel_l2(coord(rec["xyz"], rec["state"], rec["msa_coord_in"]), rec["msa_coord_out"])
        errs["msa_coord_chain"] = rel_l2(coord(rec["xyz"], rec["state"], msa), rec["msa_coord_out"])
    return errs


def build_embeddings(c, device="cpu"):
    """The b200 MsaEmbedding / PairEmbedding with a fixture's synthetic weights (oracle/make_golden.py EMBED_CONFIGS),
    and the fixture's inputs."""
    import rosettafold_pytorch_b200 as rf
    from oracle.make_golden import synth_embed_inputs

    m = rf.MsaEmbedding(c["d_input"], c["d_msa"], c["max_len"]).eval()
    p = rf.PairEmbedding(c["d_input"], c["d_pair"], c["max_len"], use_template=c["use_template"], d_template=c["d_template"]).eval()
    sd_m = synth_state_dict(m.state_dict(), seed=c["seed"])
    sd_p = synth_state_dict(p.state_dict(), seed=c["seed"] + 1)
    m.load_state_dict(sd_m, strict=True)
    p.load_state_dict(sd_p, strict=True)
    return m.to(device), p.to(device), sd_m, sd_p, synth_embed_inputs(c)


def build_heads(c, device="cpu"):
    """The b200 PredictionHead with a fixture's synthetic weights (oracle/make_golden.py HEAD_CONFIGS) and its input."""
    import rosettafold_pytorch_b200 as rf
    from oracle.make_golden import synth_head_input

    h = rf.PredictionHead(c["in_channels"], c["n_res_blocks"], 0.1).eval()
    sd = synth_state_dict(h.state_dict(), seed=c["seed"])
    h.load_state_dict(sd, strict=True)
    return h.to(device), sd, synth_head_input(c)


def build_graph_block(c, device="cpu"):
    """The b200 GraphTransformerBlock with a fixture's synthetic weights (oracle/make_golden.py GRAPH_CONFIGS) and inputs."""
    import rosettafold_pytorch_b200 as rf
    from oracle.make_golden import synth_graph_inputs

    blk = rf.GraphTransformerBlock(c["d_node"], c["d_out"], c["d_edge"], c["n_heads"]).eval()
    sd = synth_state_dict(blk.state_dict(), seed=c["seed"])
    blk.load_state_dict(sd, strict=True)
    return blk.to(device), sd, synth_graph_inputs(c)
