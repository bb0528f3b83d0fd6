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
