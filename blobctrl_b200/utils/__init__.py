"""Same import surface as the reference's ``blobctrl/utils/__init__.py`` (plus the stage functions)."""
from .utils import (BLOB_VIS_COLORS, pyramid_resize, splat_features, splat_features_from_scores,
                    splat_ellipses, splat_features_multiscale, visualize_features, viz_score_fn)

__all__ = ["splat_features", "viz_score_fn", "BLOB_VIS_COLORS", "splat_features_from_scores", "pyramid_resize",
           "visualize_features", "splat_features_multiscale", "splat_ellipses"]
