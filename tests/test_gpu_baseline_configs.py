"""GPU parity at the shapes BASELINE.json names — the shapes bench.py reports numbers for.

  * metric shape (1, 128, 512): one TwoTrackBlock x 4 encoder layers, bf16 mode, every stage teacher-forced
    AND the free-running chain, against the CPU oracle computed on this box (tied logits contract over
    K = N*32 = 4096, FAVOR walks 4 token tiles x 512 groups, the 3x3 convolution runs 2048 tiles);
  * config 2 (1, 64, 256): the whole 13-block trunk (README depth), bf16 mode, per-block drift;
  * config 1 (4, 8, 128): the whole 13-block trunk in the fp32 validation mode (README dummy config);
  * a tile-crossing fixture written by the UNMODIFIED reference (tests/golden/two_track_tile_crossing.pt,
    L = 136, default widths), both modes.

Tolerances are the north star's and nothing looser: relative L2 <= 1e-2 (bf16 tensor-core mode), <= 1e-4 (fp32
validation mode). The oracle (oracle/trunk_ref.py, pinned to the reference by tests/test_oracle.py) runs on the
host cores at test time; a fixture of these sizes would be hundreds of MB. Measured errors are appended to
gpurun_out/parity_r02.jsonl when that directory exists (profiles/r02_parity.md is built from it)."""
import json
import os
import time

import pytest
import torch

import rosettafold_pytorch_b200 as rf
from oracle import trunk_ref
from oracle.weights import synth_inputs, synth_state_dict
from tests.helpers import STAGES, build_block, load_golden, rel_l2, run_stages, subset_stages

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(name, payload):
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_r02.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **payload)) + "\n")
    print(name, payload)


@pytest.fixture(autouse=True)
def _reset_mode():
    yield
    rf.set_mode("bf16")


def test_metric_shape_block_bf16(cuda_device):
    """(1, 128, 512), 4 encoder layers per stage: what one of bench.py's 13 blocks computes."""
    cfg = dict(d_msa=384, d_pair=288, n_layers=4, B=1, N=128, L=512, seed=31)
    blk, sd, msa, pair = build_block(cfg, cuda_device)
    rf.set_mode("bf16")
    chain = run_stages(blk, msa, pair)
    torch.cuda.synchronize()
    chain = {k: v.cpu() for k, v in chain.items()}
    t0 = time.time()
    gold = {}
    with torch.no_grad():
        trunk_ref.two_track_block(msa.cpu(), pair.cpu(), sd, cfg["n_layers"], stages=gold)
    oracle_s = time.time() - t0
    forced = run_stages(blk, msa, pair, teacher=gold)
    torch.cuda.synchronize()
    e_forced = {k: rel_l2(forced[k], gold[k]) for k in STAGES}
    e_chain = {k: rel_l2(chain[k], gold[k]) for k in STAGES}
    _record("metric_shape_block_bf16", dict(shape=[1, 128, 512], layers=4, teacher_forced=e_forced, chain=e_chain,
                                            oracle_cpu_s=round(oracle_s, 1)))
    for k in STAGES:
        assert e_forced[k] < 1e-2, ("teacher-forced", e_forced)
        assert e_chain[k] < 1e-2, ("chain", e_chain)


def _trunk_drift(cfg, n_blocks, mode, device):
    """Run `n_blocks` TwoTrackBlocks (distinct weights per block) on the GPU and on the CPU oracle, each side
    feeding its own outputs forward; returns per-block (msa, pair) rel-L2."""
    rf.set_mode(mode)
    msa, pair = synth_inputs(cfg["B"], cfg["N"], cfg["L"], cfg["d_msa"], cfg["d_pair"], seed=cfg["seed"] + 100)
    m_g, p_g = msa.to(device), pair.to(device)
    m_c, p_c = msa, pair
    blk = rf.TwoTrackBlock(cfg["d_msa"], cfg["d_pair"], n_encoder_layers=cfg["n_layers"]).eval()
    # the template must be the CPU module's own state_dict: synth_state_dict recognises the aliased entries of
    # the axial layers (row_attn / col_attn / ff are registered twice, reference :505-525) by their storage
    template = blk.state_dict()
    blk_dev = rf.TwoTrackBlock(cfg["d_msa"], cfg["d_pair"], n_encoder_layers=cfg["n_layers"]).eval().to(device)
    errs = []
    for b in range(n_blocks):
        sd = synth_state_dict(template, seed=cfg["seed"] + b)
        blk_dev.load_state_dict(sd, strict=True)
        blk = blk_dev
        m_g, p_g = blk(m_g, p_g)
        with torch.no_grad():
            m_c, p_c = trunk_ref.two_track_block(m_c, p_c, sd, cfg["n_layers"])
        errs.append((rel_l2(m_g, m_c), rel_l2(p_g, p_c)))
        assert torch.isfinite(m_c).all() and torch.isfinite(p_c).all()
    return errs


def test_config2_full_trunk_bf16_drift(cuda_device):
    """BASELINE config 2: single protein L=256, Nseq=64, full default trunk (13 blocks x 4 layers), bf16 kernels
    against the fp32 oracle; the free-running outputs must still be within 1e-2 after the last block."""
    cfg = dict(d_msa=384, d_pair=288, n_layers=4, B=1, N=64, L=256, seed=41)
    t0 = time.time()
    errs = _trunk_drift(cfg, 13, "bf16", cuda_device)
    _record("config2_full_trunk_bf16", dict(shape=[1, 64, 256], blocks=13, layers=4,
                                            per_block_msa=[round(e[0], 6) for e in errs],
                                            per_block_pair=[round(e[1], 6) for e in errs], wall_s=round(time.time() - t0, 1)))
    assert errs[-1][0] < 1e-2 and errs[-1][1] < 1e-2, errs
    assert max(max(e) for e in errs) < 1e-2, errs


def test_config1_readme_trunk_fp32(cuda_device):
    """BASELINE config 1 (README dummy config: bsz 4, n_seq 8, max_len 128, 4 encoder layers) in the fp32
    validation mode: <= 1e-4 after every block. The CPU oracle needs ~10-18 s per block at this shape, so the
    default run chains 4 blocks (fp32 drift is at the 1e-6 level per block); RFK_FULL_DEPTH=1 runs all 8 + 5
    (profiles/r02_parity.md holds that run)."""
    cfg = dict(d_msa=384, d_pair=288, n_layers=4, B=4, N=8, L=128, seed=51)
    t0 = time.time()
    n_blocks = 13 if os.environ.get("RFK_FULL_DEPTH") else 4
    errs = _trunk_drift(cfg, n_blocks, "fp32", cuda_device)
    _record("config1_readme_trunk_fp32", dict(shape=[4, 8, 128], blocks=n_blocks, layers=4,
                                              per_block_msa=[float(f"{e[0]:.3e}") for e in errs],
                                              per_block_pair=[float(f"{e[1]:.3e}") for e in errs],
                                              wall_s=round(time.time() - t0, 1)))
    assert max(max(e) for e in errs) < 1e-4, errs


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_tile_crossing_golden_from_reference(cuda_device, mode, tol):
    """L = 136 (> one 128-row tile), N = 12, default widths: outputs of the UNMODIFIED reference block on the
    fixture's row subset (free-running chain; the fixture holds no full intermediates to teacher-force with)."""
    gold = load_golden("two_track_tile_crossing")
    cfg = gold["config"]
    blk, _, msa, pair = build_block(cfg, cuda_device)
    rf.set_mode(mode)
    out = subset_stages(run_stages(blk, msa, pair), cfg)
    torch.cuda.synchronize()
    errs = {k: rel_l2(out[k], gold[k]) for k in STAGES}
    _record(f"tile_crossing_golden_{mode}", dict(shape=[1, 12, 136], errs=errs))
    for k in STAGES:
        assert errs[k] < tol, errs
