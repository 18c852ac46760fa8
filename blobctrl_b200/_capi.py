"""ctypes binding of libblobsplat.so — the thin C-ABI layer (include/blobsplat.h).

PyTorch is used only for device memory and streams; every tensor crosses this boundary as a raw
device pointer plus sizes.  There is no CPU path: if the shared library is missing the import of
any renderer entry point fails loudly (``BlobSplatLibraryError``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

F32, F64, BF16, F16 = 0, 1, 2, 3
SELECT_ALL, SELECT_FG, SELECT_BG = 0, 1, 2
COMPOSITE_AUTO, COMPOSITE_LANE_PIXEL, COMPOSITE_WARP_SCAN = 0, 1, 2
ENGINE_AUTO, ENGINE_FMA, ENGINE_TENSOR, ENGINE_TMA = 0, 1, 2, 3
ABI_VERSION = 4

_DTYPES = {torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16, torch.float16: F16}

LIB_PATH = os.environ.get("BLOBSPLAT_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libblobsplat.so")   # BLOBSPLAT_LIB: A/B builds (scripts/ab_variants.py)


class BlobSplatLibraryError(RuntimeError):
    """libblobsplat.so is missing or does not match include/blobsplat.h."""


class BlobSplatError(RuntimeError):
    """A C-ABI call returned a negative status (base class of the two below)."""


class BlobSplatUnsupported(BlobSplatError):
    """Status -2: the shape / dtype is outside the envelope of the requested kernel.  Nothing was launched, the CUDA
    context is untouched — the only status a caller may answer by choosing another engine."""


class BlobSplatCudaError(BlobSplatError):
    """Status -3: a CUDA call failed (launch error, trap in a kernel, sticky context error).  Never retried, never
    answered with another engine: the context may be poisoned and the caller has to see it."""


class Caps(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int), ("sm_arch", ctypes.c_int), ("max_blobs", ctypes.c_int),
                ("tensor_max_k", ctypes.c_int), ("tensor_c_multiple", ctypes.c_int), ("tensor_max_c", ctypes.c_int)]


_I, _P, _L = ctypes.c_int, ctypes.c_void_p, ctypes.c_int64
# name -> argtypes; mirrors include/blobsplat.h one to one (tests/test_capi_symbols.py checks the header)
SIGNATURES = {
    "blobsplat_abi_version": [],
    "blobsplat_get_caps": [ctypes.POINTER(Caps)],
    "blobsplat_last_error": [ctypes.c_char_p, ctypes.c_size_t],
    "blobsplat_scores": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _I, _P],
    "blobsplat_scores_ellipse": [_P, _P, ctypes.c_float, ctypes.c_float, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P],
    "blobsplat_preview": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "blobsplat_preview_u8": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "blobsplat_composite": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_resize_bilinear": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_pyramid": [_P, ctypes.POINTER(_P), _I, _I, _I, _I, _I, _P],
    "blobsplat_feature_splat": [_P, _L, _L, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_feature_splat_levels": [_I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _P],
    "blobsplat_conditioning_fill": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_residual_inject": [_P, _P, _P, ctypes.c_float, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_conv_in_weights": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_conv_in_hoisted": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "blobsplat_render": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    "blobsplat_render_small": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "blobsplat_render_multiscale": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P],
}

_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BlobSplatLibraryError(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C blobctrl_b200/csrc`). blobctrl_b200 has no CPU or PyTorch fallback.")
    try:
        handle = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise BlobSplatLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise BlobSplatLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    if handle.blobsplat_abi_version() != ABI_VERSION:
        raise BlobSplatLibraryError(f"ABI version mismatch: library {handle.blobsplat_abi_version()}, "
                                    f"binding {ABI_VERSION}; rebuild the library")
    _lib = handle
    return handle


def caps() -> Caps:
    c = Caps()
    check(lib().blobsplat_get_caps(ctypes.byref(c)))
    return c


def last_error() -> str:
    buf = ctypes.create_string_buffer(512)
    lib().blobsplat_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(status: int) -> None:
    if status == 0:
        return
    msg = last_error()
    if status == -1:
        raise ValueError(f"blobsplat: invalid argument: {msg}")
    if status == -2:
        raise BlobSplatUnsupported(f"blobsplat: unsupported: {msg}")
    raise BlobSplatCudaError(f"blobsplat: CUDA failure (status {status}): {msg}")


def f32_exact() -> bool:
    """BLOBSPLAT_F32_EXACT=1 (read per call, like the library does): float32 stage 3 stays on the FMA engine."""
    e = os.environ.get("BLOBSPLAT_F32_EXACT")
    return bool(e) and e[0] == "1"


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"blobsplat supports float32/float64/bfloat16/float16 tensors, not {dt}") from None


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: blobctrl_b200 renders on CUDA (sm_100a) only and has no CPU "
                           f"fallback; move the blob tensors to the GPU first")


def ptr(t: Optional[torch.Tensor]):
    """Device address as a plain int (the declared argtypes convert it): no ctypes object per argument on the latency path."""
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_of(t: torch.Tensor):
    """The current CUDA stream of the tensor's device as a raw handle.  torch.cuda.current_stream() builds a Stream
    object (~8 us per call — a third of a small launch); the raw accessor is the same lookup without it."""
    if _raw_stream is not None:
        return _raw_stream(dev_of(t)) or None
    return torch.cuda.current_stream(t.device).cuda_stream or None


def dev_of(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


_POISON = os.environ.get("BLOBSPLAT_POISON_OUTPUTS") == "1"


def new_output(shape, dtype, device):
    """Output buffer for a kernel that writes every element: uninitialised, or — BLOBSPLAT_POISON_OUTPUTS=1, set by the
    GPU test-suite — NaN-filled, so that an element the kernel failed to write cannot pass a comparison by holding
    stale data of an earlier, identical call from the caching allocator."""
    import torch
    if _POISON and dtype.is_floating_point:
        return torch.full(tuple(shape), float("nan"), dtype=dtype, device=device)
    return torch.empty(tuple(shape), dtype=dtype, device=device)
