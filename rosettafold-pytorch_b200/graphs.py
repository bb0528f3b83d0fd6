"""CUDA-graph execution of trunk modules.

One trunk block issues ~300 kernels through ctypes; at small protein sizes (L <= 256) the host cannot
enqueue them as fast as the GPU retires them (tools/launch_overhead.py: ~30 us of host time per
launch), and even at the metric size every kernel boundary costs launch latency. `GraphedModule`
captures a module's forward once per input shape into a CUDA graph (all intermediates live in the
graph's private memory pool, the TMA descriptors baked into the kernel parameters stay valid because
those addresses never change) and replays it with a single launch.

The librfk launchers are capture-safe: they only enqueue on `torch.cuda.current_stream()`, allocate
nothing and never synchronise. One-time `cudaFuncSetAttribute` calls and the weight-packing caches are
exercised by the warm-up forwards that run before the capture.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _key(tensors):
    return tuple((tuple(t.shape), t.dtype, t.device) for t in tensors)


def _state_signature(module: nn.Module, track_weights: bool):
    """What a captured graph has baked in besides the input shapes: the numerics mode, the operand format and the
    packed weights (the packing caches of modules.py are keyed by the parameters' (data_ptr, _version))."""
    from . import modules as M

    sig = (M.get_mode(), M._BOUNDED_DT)
    if track_weights:
        sig += tuple((p.data_ptr(), p._version) for p in module.parameters())
        sig += tuple((b.data_ptr(), b._version) for b in module.buffers())
    return sig


class GraphedModule(nn.Module):
    """Wraps an inference module whose forward takes and returns tensors (or a tuple of tensors).

    The outputs of `forward` are views of graph-owned buffers: they are overwritten by the next call
    with the same input shapes, so copy them if they must survive it.
    """

    def __init__(self, module: nn.Module, warmup: int = 2, track_weights: bool = True):
        """track_weights: re-capture when a parameter / buffer was modified in place or replaced (costs one pass over
        the module's tensors per call on the host, ~1 ms for a 13-block trunk; the replay itself is asynchronous).
        With track_weights=False call `invalidate()` after changing weights."""
        super().__init__()
        self.module = module
        self.warmup = warmup
        self.track_weights = track_weights
        self._graphs = {}
        self._sig = None

    def invalidate(self):
        """Drop every captured graph (after set_mode / weight updates when track_weights is off)."""
        self._graphs.clear()

    def _check_state(self):
        sig = _state_signature(self.module, self.track_weights)
        if sig != self._sig:
            self._graphs.clear()  # stale graphs would replay the old mode / packed weights silently
            self._sig = sig

    def _capture(self, inputs):
        static_in = [torch.empty_like(t) for t in inputs]
        for s, t in zip(static_in, inputs):
            s.copy_(t)
        side = torch.cuda.Stream(device=inputs[0].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):
                self.module(*static_in)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            out = self.module(*static_in)
        return graph, static_in, out

    @torch.no_grad()
    def forward(self, *inputs):
        if not inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in inputs):
            raise RuntimeError("GraphedModule: inputs must be CUDA tensors (there is no CPU path)")
        self._check_state()
        key = _key(inputs)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._graphs[key] = self._capture(inputs)
        graph, static_in, out = entry
        for s, t in zip(static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        graph.replay()
        return out

    def static_inputs(self, *example_inputs):
        """The graph-owned input buffers for this shape (capture if needed): callers that fill them
        in place (e.g. straight from pinned host memory) save the device-to-device copy."""
        self._check_state()
        key = _key(example_inputs)
        if key not in self._graphs:
            self._graphs[key] = self._capture(example_inputs)
        return self._graphs[key][1]
