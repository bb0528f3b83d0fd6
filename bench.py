#!/usr/bin/env python
"""bench.py — trunk forward throughput (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps 3 --warmup 3            # b200 arm (librfk kernels)
  python bench.py --impl reference --steps 1 --warmup 0     # CPU reference arm (oracle port)
  torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling

Workload: the three-track trunk (13 block executions = README depth 8 two-track + 5 three-track,
4 encoder layers, d_msa 384, d_pair 288) on one protein of L=512 residues, Nseq=128 per GPU,
synthetic MSA/pair embeddings, random-init weights, eval mode. A step = one trunk forward of one
sample per GPU; value = samples/s over all GPUs. Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D_MSA, D_PAIR, N_LAYERS = 384, 288, 4
METRIC = "trunk fwd samples/s (L=512, Nseq=128)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--L", type=int, default=512)
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=13)
    ap.add_argument("--batch", type=int, default=1, help="samples per GPU per step")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-shapes", default=None, help="write the eager pass's per-(op, algorithmic work) device times here")
    ap.add_argument("--no-long-protein", action="store_true", help="skip the config-4 record of multi-GPU runs")
    ap.add_argument("--long-L", type=int, default=1024)
    ap.add_argument("--long-N", type=int, default=256)
    ap.add_argument("--long-blocks", type=int, default=2)
    ap.add_argument("--long-timeout", type=float, default=300.0, help="seconds before the config-4 record is abandoned")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU reference leg (oracle port of the reference's trunk, timed on the host cores)
# ------------------------------------------------------------------------------------------------
def workload_desc(args):
    """config.workload: the same string in both arms."""
    return (f"trunk forward, {args.blocks} blocks x {N_LAYERS} encoder layers, B={args.batch}/GPU, Nseq={args.N}, "
            f"L={args.L}, d_msa {D_MSA}, d_pair {D_PAIR}")


class CpuBlock:
    """One whole TwoTrackBlock of the workload (all four stages, all N_LAYERS encoder layers of each, full
    (B, N, L) shape, no slicing) on the CPU oracle port: what a timed CPU step runs. The trunk is `blocks`
    such blocks of identical cost in sequence, so the only scale factor is x blocks."""

    def __init__(self, args, torch, shape=None):
        from oracle import trunk_ref
        from oracle.weights import synth_inputs, synth_state_dict
        import rosettafold_pytorch_b200 as rf

        self.torch, self.trunk_ref = torch, trunk_ref
        B, N, L = shape or (args.batch, args.N, args.L)
        blk = rf.TwoTrackBlock(D_MSA, D_PAIR, n_encoder_layers=N_LAYERS)
        self.sd = synth_state_dict(blk.state_dict(), seed=1)
        self.msa, self.pair = synth_inputs(B, N, L, D_MSA, D_PAIR, seed=2)

    def step(self):
        t0 = time.perf_counter()
        with self.torch.no_grad():
            m, p = self.trunk_ref.two_track_block(self.msa, self.pair, self.sd, N_LAYERS)
        assert bool(self.torch.isfinite(m).all()) and bool(self.torch.isfinite(p).all())
        return time.perf_counter() - t0


def sample_desc(args):
    return (f"oracle port (oracle/trunk_ref.py), fp32, torch CPU kernels: one WHOLE TwoTrackBlock (all four stages, "
            f"all {N_LAYERS} encoder layers of each) at the full ({args.batch},{args.N},{args.L}) shape per timed step, nothing "
            f"sliced; the trunk is {args.blocks} such blocks of identical cost in sequence, so seconds per trunk = "
            f"{args.blocks} x the mean timed block")


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    CpuBlock(args, torch, shape=(1, 16, 64)).step()  # first-call overheads (thread pools, oneDNN primitives)
    blk = CpuBlock(args, torch)
    for _ in range(args.warmup):
        blk.step()
    times = [blk.step() for _ in range(max(1, args.steps))]
    block_s = sum(times) / len(times)
    trunk_s = args.blocks * block_s
    value = args.batch / trunk_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": block_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args),
                   "timed_step": f"one TwoTrackBlock = 1/{args.blocks} of the workload (ms_per_step is the mean of the "
                                 f"{max(1, args.steps)} timed blocks; value = batch / ({args.blocks} x that))",
                   "ms_per_trunk": trunk_s * 1e3, "block_s_min": min(times), "block_s_max": max(times)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": sample_desc(args)},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# b200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import rosettafold_pytorch_b200 as rf
    from rosettafold_pytorch_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rf.set_mode(args.mode)

    torch.manual_seed(1234 + rank)
    trunk = rf.TrunkBlocks(D_MSA, D_PAIR, n_blocks=args.blocks, n_encoder_layers=N_LAYERS).eval().to(dev)
    B, N, L = args.batch, args.N, args.L
    g = torch.Generator().manual_seed(99 + rank)
    msa_h = torch.randn((B, N, L, D_MSA), generator=g).pin_memory()
    pair_h = torch.randn((B, L, L, D_PAIR), generator=g).pin_memory()
    msa_out_h = torch.empty_like(msa_h).pin_memory()
    pair_out_h = torch.empty_like(pair_h).pin_memory()
    msa_d, pair_d = msa_h.to(dev), pair_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The product path replays the trunk as ONE CUDA graph (rf.GraphedModule); the eager path is
    # kept for the per-kernel-family device timing, whose events cannot live inside a graph.
    gtrunk = rf.GraphedModule(trunk)
    msa_s, pair_s = gtrunk.static_inputs(msa_d, pair_d)  # graph-owned input buffers
    msa_s.copy_(msa_d)
    pair_s.copy_(pair_d)

    def step_eager():
        return trunk(msa_d, pair_d)

    def step_resident():
        return gtrunk(msa_s, pair_s)  # inputs already resident in HBM: no copy, one graph launch

    # End-to-end step through the public API with HOST buffers: H2D of this step's inputs (pinned
    # memory), graph replay, D2H of its outputs. Like a serving loop, the copies run on their own
    # streams: while step i computes, the inputs of step i+1 are already crossing PCIe and the outputs
    # of step i-1 are on their way back. Every step's copies are inside the timed region.
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    stage_in = [(torch.empty_like(msa_d), torch.empty_like(pair_d)) for _ in range(2)]
    stage_out = [(torch.empty_like(msa_d), torch.empty_like(pair_d)) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free_in = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_free_out = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0}

    def upload(slot):
        with torch.cuda.stream(copy_in):
            copy_in.wait_event(ev_free_in[slot])            # the compute stream has consumed this slot
            stage_in[slot][0].copy_(msa_h, non_blocking=True)
            stage_in[slot][1].copy_(pair_h, non_blocking=True)
            ev_in[slot].record(copy_in)

    def step_e2e():
        i = e2e_state["i"]
        slot = i & 1
        cur = torch.cuda.current_stream()
        if i == 0:
            for sl in range(2):
                ev_free_in[sl].record(cur)
                ev_free_out[sl].record(cur)
            upload(0)
        upload((i + 1) & 1)                                  # next step's inputs, behind this step's compute
        cur.wait_event(ev_in[slot])
        msa_s.copy_(stage_in[slot][0], non_blocking=True)    # device-to-device into the graph's inputs
        pair_s.copy_(stage_in[slot][1], non_blocking=True)
        ev_free_in[slot].record(cur)
        mo, po = gtrunk(msa_s, pair_s)
        cur.wait_event(ev_free_out[slot])
        stage_out[slot][0].copy_(mo, non_blocking=True)
        stage_out[slot][1].copy_(po, non_blocking=True)
        ev_out[slot].record(cur)
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(ev_out[slot])
            msa_out_h.copy_(stage_out[slot][0], non_blocking=True)
            pair_out_h.copy_(stage_out[slot][1], non_blocking=True)
            ev_free_out[slot].record(copy_out)
        e2e_state["i"] = i + 1

    def drain_e2e():
        torch.cuda.current_stream().wait_stream(copy_out)    # the last outputs must have landed
        torch.cuda.current_stream().wait_stream(copy_in)
        e2e_state["i"] = 0

    def timed(fn, steps, after=None):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(3, args.warmup)):
        step_eager()
    for _ in range(max(3, args.warmup)):
        step_resident()
    clocks = Clocks(local)
    clocks.start()
    # per-kernel-family device time: CUDA events on the launching stream around every librfk call
    # of an eager pass over the same step (same kernels, same shapes as the graph replays)
    n0 = rf._lib.launch_count()
    ops.start_timing(["gemm_bf16", "gemm_f32", "favor_attention", "conv3x3", "layernorm"])
    eager_ms = timed(step_eager, args.steps)
    fam = ops.stop_timing()
    launches = (rf._lib.launch_count() - n0) // max(1, args.steps) * args.steps
    if args.dump_shapes and rank == 0:
        rows = []
        for k, v in fam.items():
            for w, e in v["shapes"].items():
                if e["calls"]:
                    rows.append(dict(op=k, work=w, calls_per_step=e["calls"] / args.steps, avg_us=1e3 * e["ms"] / e["calls"],
                                     ms_per_step=e["ms"] / args.steps, rate=w / (e["ms"] / e["calls"] * 1e-3) / 1e12))
        rows.sort(key=lambda r: -r["ms_per_step"])
        with open(args.dump_shapes, "w") as f:
            json.dump(rows, f, indent=1)
    total_ms = timed(step_resident, args.steps)
    step_e2e()  # warm the pinned-copy path
    drain_e2e()
    torch.cuda.synchronize()
    e2e_ms = timed(step_e2e, args.steps, after=drain_e2e)
    clk = clocks.stop()

    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)

    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "MEASURED_PEAKS.json (sustained bf16)" if peaks else "fallback"

    # Dominant KERNEL = the (op, shape) bucket with the largest device time in the step: launches of one
    # op with equal algorithmic work are one kernel at one shape (e.g. the 8 pair axial FAVOR calls
    # of a block). achieved = algorithmic FLOPs per launch / average launch duration (CUDA events).
    tensor_ops = ("gemm_bf16", "favor_attention", "gemm_f32", "conv3x3")
    buckets = [(k, w, e) for k, v in fam.items() if k in tensor_ops for w, e in v["shapes"].items() if e["calls"]]
    roofline = None
    if buckets:
        dom, work, e = max(buckets, key=lambda t: t[2]["ms"])
        avg_ms = e["ms"] / e["calls"]
        ach = work / (avg_ms * 1e-3) / 1e12
        kernel = {"gemm_bf16": "rfk::gemm_tc_kernel (tcgen05)", "favor_attention": "rfk::favor_tm_kernel (tcgen05 FAVOR+, features in TMEM)",
                  "gemm_f32": "rfk::gemm_f32_kernel", "conv3x3": "rfk::gemm_tc_kernel<CONV> (implicit GEMM)"}[dom]
        # DRAM bytes per launch of that kernel from the round's `ncu --set full` capture (profiles/)
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tr_path):
            for ent in json.load(open(tr_path)):
                if ent["op"] == dom and abs(ent["flops_per_launch"] - work) <= 1e-6 * work:
                    traffic = ent["dram_bytes_per_launch"]
        roofline = {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": ach / tf_peak, "traffic": traffic, "peak_source": peak_src,
                    "flops_per_launch": work, "launches_per_step": e["calls"] / args.steps, "avg_launch_ms": avg_ms,
                    "share_of_step": e["ms"] / eager_ms}
    families = {k: {"ms_per_step": v["ms"] / args.steps, "calls_per_step": v["calls"] / args.steps,
                    ("GB/s" if k == "layernorm" else "TFLOP/s"): (v["work"] / (v["ms"] * 1e-3) / (1e9 if k == "layernorm" else 1e12)) if v["ms"] > 0 else 0.0}
                for k, v in fam.items() if v["calls"]}
    if "layernorm" in families:
        families["layernorm"]["frac_of_hbm_peak"] = families["layernorm"]["GB/s"] / hbm_peak

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args),
                   "parallelism": f"replicas x{world} (one sample per GPU, no data-path collective)",
                   "l2": "working set (msa 101 MB + pair 302 MB fp32 + intermediates) exceeds the 126 MB L2; no explicit flush",
                   "ms_per_block": ms_per_step / args.blocks,
                   "execution": "timed region = CUDA-graph replays of the whole trunk (rf.GraphedModule, one launch per step); "
                                "kernel-family durations and gpu_launches come from an eager pass of the same step",
                   "eager_ms_per_step": eager_ms / args.steps,
                   "e2e": "every step: H2D of its inputs from pinned host memory, graph replay, D2H of its outputs; the copies run on "
                          "their own streams so step i+1's upload and step i-1's download overlap step i's compute"},
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": int(msa_h.numel() * 4 + pair_h.numel() * 4),
                "d2h_bytes_per_step": int(msa_out_h.numel() * 4 + pair_out_h.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "kernel_families": families,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        CpuBlock(args, torch, shape=(1, 16, 64)).step()  # first-call overheads (thread pools, oneDNN primitives)
        block_s = CpuBlock(args, torch).step()
        line["cpu_baseline"] = {"value": B / (args.blocks * block_s), "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": sample_desc(args) + " (one timed block here; `--impl reference` averages K)",
                                "block_s": round(block_s, 2)}
    if world > 1 and not args.no_long_protein:
        # outside the timed regions above; appended to the same JSON line so that the driver's scaling run carries it.
        # The headline (replicas) number must survive any failure of this extra record: exceptions are caught, and a
        # watchdog on every rank prints the line without the record and leaves if the section does not finish in time
        # (a collective that never completes cannot be caught any other way).
        def bail():
            line["long_protein"] = {"error": f"did not finish within {args.long_timeout} s"}
            if rank == 0:
                print(json.dumps(line), flush=True)
            os._exit(0)

        dog = threading.Timer(args.long_timeout, bail)
        dog.daemon = True
        dog.start()
        try:
            line["long_protein"] = long_protein_record(args, torch, dist, rf, dev, rank, world)
        except Exception as e:
            line["long_protein"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        dog.cancel()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def long_protein_record(args, torch, dist, rf, dev, rank, world):
    """BASELINE.json config 4 under the driver's multi-GPU runs: ONE long protein (1, 256, 1024) through a trunk of
    `long_blocks` blocks sharded over the ranks (rosettafold_pytorch_b200/sharded.py: MSA sequence- / residue-sharded,
    pair map row-sharded, NCCL all-to-all / all-reduce / all-gather between the stages), against the same trunk on one
    GPU of the same box: ms per block (CUDA events, max over ranks), speed-up, and rel-L2 of the sharded result."""
    L, N, nb = args.long_L, args.long_N, args.long_blocks
    if N % world or L % world:
        return {"skipped": f"Nseq {N} / L {L} not divisible by {world} ranks"}
    torch.manual_seed(4321)  # identical weights (and FAVOR projections) on every rank
    trunk = rf.TrunkBlocks(D_MSA, D_PAIR, n_blocks=nb, n_encoder_layers=N_LAYERS).eval().to(dev)
    g = torch.Generator().manual_seed(77)
    msa = torch.randn((1, N, L, D_MSA), generator=g).to(dev)
    pair = torch.randn((1, L, L, D_PAIR), generator=g).to(dev)
    eager = rf.ShardedTrunkBlocks(trunk)
    sharded = rf.sharded.SegmentedGraph(eager)  # CUDA graphs of the compute segments, collectives eager in between

    def timed(fn, steps=3):
        for _ in range(2):
            out = fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / nb, out

    ms_sharded, (m_s, p_s) = timed(lambda: sharded(msa, pair))
    ms_eager, _ = timed(lambda: eager(msa, pair))
    ms_single, (m_1, p_1) = timed(lambda: trunk(msa, pair))  # every rank runs the whole protein on its own GPU

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm())

    rec = {"config": f"one protein (1,{N},{L}), {nb} blocks x {N_LAYERS} encoder layers, sharded over {world} GPUs "
                     "(MSA by sequence / residue, pair map by row; NCCL all-to-all, all-reduce, all-gather)",
           "execution": "rf.sharded.SegmentedGraph: one CUDA graph per compute segment, NCCL collectives issued eagerly in between",
           "segments": sharded.segments(),
           "n_gpus": world, "ms_per_block_sharded": ms_sharded, "ms_per_block_sharded_eager_launches": ms_eager,
           "ms_per_block_one_gpu": ms_single,
           "speedup": ms_single / ms_sharded, "scaling": "strong",
           "rel_l2_vs_one_gpu": {"msa": rel(m_s, m_1), "pair": rel(p_s, p_1)}}
    del trunk, sharded, eager, msa, pair, m_s, p_s, m_1, p_1
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
