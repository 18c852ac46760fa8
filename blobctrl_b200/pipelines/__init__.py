from .conditioning import (BlobConditioning, BlobConditioningMixin, BlobNetInputBuffers, construct_blobnet_input,
                           inject_residual, prepare_blob_conditioning, splat_features_from_scores)

__all__ = ["BlobConditioning", "BlobConditioningMixin", "BlobNetInputBuffers", "construct_blobnet_input", "inject_residual", "prepare_blob_conditioning",
           "splat_features_from_scores"]
from .conv_in_hoist import HoistedConvIn  # noqa: E402

__all__.append("HoistedConvIn")
