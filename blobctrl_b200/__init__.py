"""blobctrl_b200 — B200-native (sm_100a) blob splatting behind BlobCtrl's own renderer API.

Scope: the one data-parallel hot path of TencentARC/BlobCtrl — ``blobctrl/utils/utils.py``'s
``splat_features`` / ``splat_features_from_scores`` / ``pyramid_resize`` and the pipeline's
conditioning prologue — as hand-written CUDA behind the C ABI in ``include/blobsplat.h``.
Nothing else of BlobCtrl (BlobNet, UNet, diffusers, the Gradio app) is reimplemented here.
"""
from . import _capi
from .utils.utils import (BLOB_VIS_COLORS, pyramid_resize, splat_features, splat_features_from_scores,
                          splat_ellipses, splat_features_multiscale, visualize_features, viz_score_fn)

__version__ = "0.1.0"
__all__ = ["splat_features", "splat_features_from_scores", "pyramid_resize", "visualize_features",
           "splat_features_multiscale", "splat_ellipses", "viz_score_fn", "BLOB_VIS_COLORS", "_capi"]
