"""Import the unmodified reference (blobctrl + its diffusers 0.30.0 fork) from baseline/_ref.

baseline/_ref is produced by scripts/install_reference.sh (pip install --target, git-ignored, shipped to the GPU box
with the snapshot); /root/reference itself does not exist there.  Three shims, none of which touches reference code
(SURVEY.md Appendix B): the fork wants ``transformers.utils.FLAX_WEIGHTS_NAME`` (gone in transformers 5), it must be
imported before a spec-less ``matplotlib`` stub exists, and ``blobctrl/utils/utils.py:11`` imports the (unused)
``matplotlib.cm``.
"""
from __future__ import annotations

import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


class ReferenceUnavailable(RuntimeError):
    pass


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "blobctrl")) and os.path.isdir(os.path.join(REF_DIR, "diffusers"))


_ns = None


def load(need_pipeline: bool = True):
    """Returns a namespace with the reference's public objects.  need_pipeline=False imports only blobctrl.utils.utils
    (the renderer), which needs nothing but torch + einops + cv2."""
    global _ns
    if not available():
        raise ReferenceUnavailable(f"{REF_DIR} is missing: run scripts/install_reference.sh where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    ns = _ns or types.SimpleNamespace(utils=None, pipeline=None)
    if "diffusers" not in sys.modules:
        # always, even when only the renderer is wanted: the fork must be imported BEFORE the matplotlib stub exists (its
        # find_spec("matplotlib") probe rejects a spec-less module) or a later load(need_pipeline=True) would fail
        import transformers.utils as TU
        if not hasattr(TU, "FLAX_WEIGHTS_NAME"):
            TU.FLAX_WEIGHTS_NAME = "flax_model.msgpack"
        import diffusers
        assert os.path.abspath(diffusers.__file__).startswith(REF_DIR), f"diffusers resolved to {diffusers.__file__}"
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            m = types.ModuleType("matplotlib"); m.cm = types.ModuleType("matplotlib.cm")
            sys.modules["matplotlib"] = m; sys.modules["matplotlib.cm"] = m.cm
    if ns.utils is None:
        import blobctrl.utils.utils as U
        assert os.path.abspath(U.__file__).startswith(REF_DIR), f"blobctrl resolved to {U.__file__}"
        ns.utils = U
    if need_pipeline and ns.pipeline is None:
        from blobctrl.models.blobnet import BlobNetModel
        from blobctrl.pipelines.pipeline_blobnet import StableDiffusionBlobNetPipeline
        from diffusers import AutoencoderKL, UNet2DConditionModel, UniPCMultistepScheduler
        ns.pipeline = StableDiffusionBlobNetPipeline
        ns.BlobNetModel, ns.UNet2DConditionModel = BlobNetModel, UNet2DConditionModel
        ns.UniPCMultistepScheduler, ns.AutoencoderKL = UniPCMultistepScheduler, AutoencoderKL
    _ns = ns
    return ns
