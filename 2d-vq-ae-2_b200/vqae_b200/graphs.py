"""CUDA-graph replay of a whole encode / decode step.

One step of the hot path is 13 (encode, 256-model) to ~35 (512-model round trip) kernel launches
issued from Python through ctypes: 1.4 - 3.9 ms of host time per step, and ~10 us of device idle time
between consecutive launches.  Both vanish when the step is captured once and replayed
(profiles/host_overhead.py: encode at batch 256 3.87 -> 3.74 ms, and no host work left to hide).

The step functions of this package only launch kernels of ``libvqae_b200.so`` on torch's current
stream and allocate through torch's caching allocator, so ``torch.cuda.graph`` captures them as they
are.  Everything a step reads besides its inputs (packed weights, codebooks) is packed during the
eager warm-up calls and stays alive in the plans' caches.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import engine as E

class CapturedStep:
    """``fn(*tensors)`` as a CUDA graph per distinct set of input buffers.

    The first ``warmup`` calls with a new combination of input shapes run eagerly (they pack weights,
    set kernel attributes and size the allocator's pools -- none of which may happen during a
    capture); from then on the first call with a given set of buffers is captured (and replayed
    once), later ones are replays.  Replays return the SAME
    output tensors every time (owned by the graph): consume or copy them before the next call with
    the same buffers.  All graphs of one ``CapturedStep`` share one memory pool, so they must not
    run concurrently (they never do on one stream)."""

    def __init__(self, fn: Callable, warmup: int = 1, key_extra: Callable = None, max_graphs: int = 512):
        """``key_extra()``: extra hashable state a captured graph depends on (e.g. the precision the
        step would pick); a change selects / captures another graph.  ``max_graphs``: graphs are keyed on
        the input BUFFERS, so a caller that passes freshly allocated tensors every time would capture
        without end -- beyond this many the oldest graph is dropped (use persistent staging buffers)."""
        self.fn, self.warmup, self.key_extra, self.max_graphs = fn, warmup, key_extra, max_graphs
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, object, int]] = {}
        self._seen: Dict[Tuple, int] = {}
        self._eager_only = set()
        self._pool = None

    def _key(self, tensors) -> Tuple:
        extra = self.key_extra() if self.key_extra is not None else None
        return (extra,) + tuple((t.data_ptr(), tuple(t.shape), t.dtype) for t in tensors)

    def reset(self) -> None:
        """Drop every captured graph (REQUIRED after the model's weights change: a graph reads the
        packed weight buffers that existed when it was captured)."""
        self._graphs.clear()
        self._seen.clear()
        self._eager_only.clear()

    @torch.no_grad()
    def __call__(self, *tensors: torch.Tensor):
        key = self._key(tensors)
        hit = self._graphs.get(key)
        if hit is not None:
            graph, out, n_launch = hit
            graph.replay()
            E._graph_launch_adjust += n_launch
            return out
        shape_key = (key[0],) + tuple(k[1:] for k in key[1:])      # without the buffer addresses
        seen = self._seen.get(shape_key, 0)
        if seen < self.warmup:
            self._seen[shape_key] = seen + 1
            return self.fn(*tensors)
        if key in self._eager_only:
            return self.fn(*tensors)
        graph = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()
        l0 = E.raw_launch_count()
        try:
            with torch.cuda.graph(graph, pool=self._pool, capture_error_mode="thread_local"):
                out = self.fn(*tensors)
        except RuntimeError as exc:
            # a step that cannot be captured (it allocates or synchronises outside torch's allocator)
            # keeps running as plain launches of the same kernels; say so once
            E._graph_launch_adjust -= E.raw_launch_count() - l0
            self._eager_only.add(key)
            import warnings
            warnings.warn(f"vqae_b200.graphs: capture failed, this step runs eagerly ({exc}); note that "
                          "torch leaves its CUDA random generator in capture mode after a failed capture")
            torch.cuda.synchronize()
            return self.fn(*tensors)
        n_launch = E.raw_launch_count() - l0    # recorded during the capture, not executed
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = (graph, out, n_launch)
        graph.replay()                          # adjust: - n_launch (recorded) + n_launch (this replay)
        return out
