"""CPU: the C-ABI library loads and exports exactly what include/blobsplat.h declares, and the ctypes
binding agrees with the header's parameter lists.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "blobsplat.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"BLOBSPLAT_API\s+int\s+(blobsplat_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_declares_the_expected_entry_points():
    d = _declared()
    assert set(d) == {"blobsplat_abi_version", "blobsplat_get_caps", "blobsplat_last_error", "blobsplat_scores",
                      "blobsplat_scores_ellipse", "blobsplat_composite", "blobsplat_resize_bilinear", "blobsplat_pyramid",
                      "blobsplat_feature_splat", "blobsplat_feature_splat_levels", "blobsplat_conditioning_fill", "blobsplat_residual_inject", "blobsplat_render", "blobsplat_render_small", "blobsplat_render_multiscale",
                      "blobsplat_preview", "blobsplat_preview_u8", "blobsplat_conv_in_weights", "blobsplat_conv_in_hoisted"}


def test_library_exports_every_declared_symbol():
    from blobctrl_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared():
        assert hasattr(handle, name), f"{name} not exported"


def test_binding_matches_header_arity_and_abi_version():
    from blobctrl_b200 import _capi
    d = _declared()
    assert set(_capi.SIGNATURES) == set(d)
    for name, n in d.items():
        assert len(_capi.SIGNATURES[name]) == n, f"{name}: header has {n} params, binding {len(_capi.SIGNATURES[name])}"
    assert _capi.lib().blobsplat_abi_version() == _capi.ABI_VERSION
    hdr = open(HEADER).read()
    assert f"#define BLOBSPLAT_ABI_VERSION {_capi.ABI_VERSION}" in hdr
    for name, val in (("F32", 0), ("F64", 1), ("BF16", 2), ("F16", 3)):
        assert re.search(rf"BLOBSPLAT_{name}\s*=\s*{val}\b", hdr) and getattr(_capi, name) == val


def test_caps_and_argument_validation_without_a_gpu():
    from blobctrl_b200 import _capi
    c = _capi.caps()
    assert c.abi_version == _capi.ABI_VERSION == 4 and c.sm_arch == 100 and c.max_blobs >= 65536
    L = _capi.lib()
    # invalid arguments are rejected before any CUDA call
    assert L.blobsplat_scores(None, None, None, None, 0, 1, 1, 0, 8, 0, None, 0, None, 0, 0, -1, None) == -1
    assert "bad shape" in _capi.last_error()
    assert L.blobsplat_scores(None, None, None, None, 0, 1, 1, 8, 8, 0, None, 0, None, 0, 0, -1, None) == -1
    assert "NULL" in _capi.last_error()
    assert L.blobsplat_scores(None, None, None, None, 0, 0, 1, 8, 8, 0, None, 0, None, 0, 0, -1, None) == 0   # N == 0
    assert L.blobsplat_pyramid(None, None, 3, 1, 20, 0, -1, None) == -1
    assert "divisible" in _capi.last_error()
    with pytest.raises(ValueError):
        _capi.check(-1)
    with pytest.raises(TypeError):
        import torch
        _capi.dtype_code(torch.int32)
    # N == 0 is a no-op success
    assert L.blobsplat_composite(None, None, 0, 3, 4, 4, 0, -1, None) == 0
    assert L.blobsplat_preview(None, None, None, None, 0, None, 0, 0, 1, 8, 8, None, None, -1, None) == 0
    assert L.blobsplat_preview(None, None, None, None, 2, None, 0, 1, 1, 8, 8, None, None, -1, None) == -1
    assert L.blobsplat_preview_u8(None, None, None, None, 0, None, 0, 0, 1, 8, 8, None, -1, None) == 0
    assert L.blobsplat_preview_u8(None, None, None, None, 0, None, 0, 1, 1, 8, 8, None, -1, None) == -1     # NULL image
    assert L.blobsplat_conv_in_weights(None, None, None, 1, 8, 30, 4, 20, 1, 0, -1, None) == -1      # Cin != lc + 1 + C
    assert "input planes" in _capi.last_error()
    assert L.blobsplat_conv_in_hoisted(None, None, None, None, None, None, 0, 8, 25, 4, 2, 4, 4, 2, 0, -1, None) == 0
