"""Opt-in CUDA-graph cache for the latency-bound calls (BASELINE configs 1-2).

At N = 1 a render is a ~6 us kernel behind ~20 us of Python + ctypes + allocator work (bench ``variants``).  With
``splat_features(..., cuda_graph=True)`` the whole call — canonicalisation copies, every launch, the output buffers — is
captured once per (argument addresses, shapes, dtypes, options) and later calls with the same key are one
``cudaGraphLaunch``.  The C ABI never allocates or synchronises, which is what makes the capture legal.

Semantics differ from the eager call in ONE way, which is why it is opt-in: the returned tensors are the graph's
static buffers, overwritten (in stream order) by the next call with the same key.  The graph reads the argument tensors
at replay time, so updating them in place between calls (``xs.copy_(...)``) re-renders with the new values.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Tuple

import torch

_MAX_ENTRIES = 16
_cache: "OrderedDict[Tuple, Tuple[torch.cuda.CUDAGraph, Any, tuple]]" = OrderedDict()


def tensor_key(tensors) -> Tuple:
    """Identity + address of every tensor argument (None allowed): ~1 us for six tensors, against ~6 us for a key built
    from shapes, strides and dtypes.  The cache keeps the argument tensors alive, so an id is never recycled while its
    entry exists; the address catches a tensor whose storage was swapped (``set_`` / ``resize_``)."""
    return tuple(0 if t is None else id(t) for t in tensors) + tuple(0 if t is None else t.data_ptr() for t in tensors)


def capturable(*tensors) -> bool:
    """Every tensor argument must already live on the device: a pageable host -> device copy cannot be captured."""
    return all(t is None or (torch.is_tensor(t) and t.is_cuda) for t in tensors)


def call(key: Tuple, fn: Callable[[], Any], keepalive: tuple = ()):
    """Replay the graph cached under ``key`` or capture ``fn`` (which must launch on the current stream only).
    ``keepalive`` pins the argument tensors whose addresses are baked into the graph."""
    hit = _cache.get(key)
    if hit is not None:
        _cache.move_to_end(key)
        hit[0].replay()
        return hit[1]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()                                   # warm-up outside capture: module loads, cudaFuncSetAttribute, allocator
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn()
    graph.replay()
    _cache[key] = (graph, out, keepalive)
    while len(_cache) > _MAX_ENTRIES:
        _cache.popitem(last=False)
    return out


def clear() -> None:
    _cache.clear()
