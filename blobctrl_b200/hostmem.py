"""Pinned host buffers for the host-input path (streaming.HostRenderer).

``pinned_like`` returns page-locked host tensors the caller fills and hands to HostRenderer.  With
``write_combined=True`` the pages are allocated by ``cudaHostAlloc(cudaHostAllocWriteCombined)``: the CPU writes them
through its write-combining buffers and never caches them, so the GPU's DMA reads are not snooped against the CPU caches.
Such memory is for buffers the host only WRITES (sequential fills); reading it back on the CPU is slow.
The reference keeps its blob parameters and features in ordinary pageable tensors (scripts/blobctrl_inference.py:101-109)
and ships the rendered maps; here the inputs are what crosses the link.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

_cudaHostAllocDefault = 0x00
_cudaHostAllocWriteCombined = 0x04
_live = {}                                   # data_ptr -> cudart handle: freed by free_pinned


def _cudart():
    """The CUDA runtime torch already loaded (by soname), else the wheel's or the toolkit's copy.  Any instance will do:
    they share the device's primary context, and pinned allocations are visible to every runtime in the process."""
    import glob
    import os
    names = ["libcudart.so.12", "libcudart.so"]
    names += glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    names += glob.glob("/usr/local/cuda/lib64/libcudart.so*")
    for name in names:
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("libcudart not found (tried the loaded runtime, torch's wheel and /usr/local/cuda)")


def pinned_empty(shape, dtype=torch.float32, write_combined: bool = False) -> torch.Tensor:
    """A page-locked host tensor; write_combined=True -> cudaHostAllocWriteCombined (freed by free_pinned)."""
    if not write_combined:
        return torch.empty(shape, dtype=dtype).pin_memory()
    torch.cuda.init()
    rt = _cudart()
    n = 1
    for s in shape:
        n *= int(s)
    nbytes = max(n * torch.empty((), dtype=dtype).element_size(), 1)
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(_cudaHostAllocWriteCombined))
    if rc != 0 or not p.value:
        raise RuntimeError(f"cudaHostAlloc(write-combined, {nbytes} B) failed: {rc}")
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    t = torch.frombuffer(buf, dtype=dtype, count=n).view(*shape)
    _live[t.data_ptr()] = (rt, p)
    return t


def pinned_like(tensors: Dict[str, torch.Tensor], write_combined: bool = False) -> Dict[str, torch.Tensor]:
    """Pinned copies of host tensors (filled)."""
    out = {}
    for k, v in tensors.items():
        t = pinned_empty(tuple(v.shape), v.dtype, write_combined)
        t.copy_(v)
        out[k] = t
    return out


def free_pinned(t: torch.Tensor) -> None:
    """Release a write-combined buffer from pinned_empty (ordinary pinned tensors are freed by torch)."""
    h = _live.pop(t.data_ptr(), None)
    if h is not None:
        h[0].cudaFreeHost(h[1])
