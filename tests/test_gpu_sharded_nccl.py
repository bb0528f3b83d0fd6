"""Multi-GPU parity of the long-protein path (BASELINE.json config 4; reference :31-54 RowWise / ColWise are what the
all-to-all replaces): two ranks over NCCL, sharded trunk vs the same trunk on one GPU. Needs >= 2 visible GPUs
(skipped otherwise; the host logic of the sharding is covered on the CPU by tests/test_sharded_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
sys.path.insert(0, os.environ["RFK_ROOT"])
import torch, torch.distributed as dist
import rosettafold_pytorch_b200 as rf
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
trunk = rf.TrunkBlocks(384, 288, n_blocks=2, n_encoder_layers=2).eval().to(dev)
g = torch.Generator().manual_seed(5)
N, L = 24, 136   # ragged against the 128-row tiles, divisible by the 2 ranks
msa = torch.randn((1, N, L, 384), generator=g).to(dev)
pair = torch.randn((1, L, L, 288), generator=g).to(dev)
out = {}
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for mode in ("fp32", "bf16"):
    rf.set_mode(mode)
    sharded = rf.ShardedTrunkBlocks(trunk)
    m_s, p_s = sharded(msa, pair)
    m_1, p_1 = trunk(msa, pair)
    torch.cuda.synchronize()
    out[mode] = [rel(m_s, m_1), rel(p_s, p_1)]
    # the same through CUDA graphs of the compute segments (collectives eager in between): record, then replay twice
    seg = rf.sharded.SegmentedGraph(sharded)
    for scale in (1.0, 0.5):
        m_g, p_g = seg(msa * scale, pair * scale)
        m_e, p_e = sharded(msa * scale, pair * scale)
        torch.cuda.synchronize()
        out[mode + "_graph_vs_eager_%g" % scale] = [rel(m_g, m_e), rel(p_g, p_e)]
    out[mode + "_segments"] = seg.segments()
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_sharded_trunk_matches_single_gpu_over_nccl(cuda_device, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, RFK_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    print(res)
    assert max(res["fp32"]) < 1e-4, res   # same arithmetic, different summation order of the sharded reductions
    assert max(res["bf16"]) < 1e-2, res
    for k, v in res.items():
        if "graph_vs_eager" in k:
            assert max(v) < 1e-5, (k, res)  # same kernels, same order (atomics of the InstanceNorm statistics aside)
