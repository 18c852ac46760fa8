"""The reference's ``blobctrl/utils/__init__.py`` import surface (splat_features, viz_score_fn, BLOB_VIS_COLORS,
vis_gt_ellipse_from_ellipse — blobctrl/utils/__init__.py:1-2), plus the stage functions and the extensions."""
from .utils import (BLOB_VIS_COLORS, pyramid_resize, splat_ellipses, splat_features, splat_features_from_scores,
                    splat_features_multiscale, vis_gt_ellipse_from_ellipse, vis_gt_ellipse_from_norm_ellipse,
                    vis_gt_ellipse_from_norm_gs, vis_scores, visualize_features, viz_score_fn)

__all__ = ["splat_features", "viz_score_fn", "BLOB_VIS_COLORS", "vis_gt_ellipse_from_ellipse", "splat_features_from_scores",
           "pyramid_resize", "visualize_features", "splat_features_multiscale", "splat_ellipses", "vis_scores",
           "vis_gt_ellipse_from_norm_gs", "vis_gt_ellipse_from_norm_ellipse"]
