"""Developer tool: where does the bf16 error of a deep trunk come from? Runs the config-2 chain (1,64,256) on the
CPU oracle and on the GPU up to block `nb`, then block nb+1 (a) teacher-forced per stage from the oracle's inputs,
(b) from the GPU's own drifted inputs, printing per-stage rel-L2."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rosettafold_pytorch_b200 as rf
from oracle import trunk_ref
from oracle.weights import synth_inputs, synth_state_dict
from tests.helpers import STAGES, rel_l2, run_stages

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device("cuda:0")
cfg = dict(d_msa=384, d_pair=288, n_layers=4, B=1, N=64, L=256, seed=41)
msa, pair = synth_inputs(1, 64, 256, 384, 288, seed=141)
cpu_blk = rf.TwoTrackBlock(384, 288, n_encoder_layers=4).eval()
tmpl = cpu_blk.state_dict()
blk = rf.TwoTrackBlock(384, 288, n_encoder_layers=4).eval().to(dev)
m_c, p_c, m_g, p_g = msa, pair, msa.to(dev), pair.to(dev)
for b in range(nb):
    sd = synth_state_dict(tmpl, seed=41 + b)
    blk.load_state_dict(sd)
    m_g, p_g = blk(m_g, p_g)
    with torch.no_grad():
        m_c, p_c = trunk_ref.two_track_block(m_c, p_c, sd, 4)
    print(b, "%.2e %.2e" % (rel_l2(m_g, m_c), rel_l2(p_g, p_c)), flush=True)
sd = synth_state_dict(tmpl, seed=41 + nb)
blk.load_state_dict(sd)
gold = {}
with torch.no_grad():
    trunk_ref.two_track_block(m_c, p_c, sd, 4, stages=gold)
forced = run_stages(blk, m_c.to(dev), p_c.to(dev), teacher=gold)
print("block", nb, "exact inputs, teacher-forced:", {k: "%.2e" % rel_l2(forced[k], gold[k]) for k in STAGES})
chain = run_stages(blk, m_c.to(dev), p_c.to(dev))
print("block", nb, "exact inputs, chain:         ", {k: "%.2e" % rel_l2(chain[k], gold[k]) for k in STAGES})
drift = run_stages(blk, m_g, p_g)
print("block", nb, "drifted inputs, chain:       ", {k: "%.2e" % rel_l2(drift[k], gold[k]) for k in STAGES})
# which drifted input matters: msa only / pair only
d1 = run_stages(blk, m_g, p_c.to(dev))
print("block", nb, "drifted msa, exact pair:     ", {k: "%.2e" % rel_l2(d1[k], gold[k]) for k in STAGES})
d2 = run_stages(blk, m_c.to(dev), p_g)
print("block", nb, "exact msa, drifted pair:     ", {k: "%.2e" % rel_l2(d2[k], gold[k]) for k in STAGES})
