"""Public API of the product package (re-exported by the ``sfvos_b200`` alias)."""
from . import _lib, ops  # noqa: F401
from .slowfast import SlowFastLayers  # noqa: F401
from .roi_heads import (FastRCNNPredictor, MaskRCNNHeads, MaskRCNNPredictor, MultiScaleRoIAlign, RoIHeads,  # noqa: F401
                        TwoMLPHead, fastrcnn_loss, install, maskrcnn_inference, maskrcnn_loss, paste_masks_in_image,
                        pool_pair, postprocess, project_masks_on_boxes)

from .backbone import AnchorGenerator, FeaturePyramidNetwork, RPNHead, install_backbone  # noqa: F401,E402

__all__ = ["FeaturePyramidNetwork", "RPNHead", "AnchorGenerator", "install_backbone", "SlowFastLayers", "MultiScaleRoIAlign", "MaskRCNNHeads", "MaskRCNNPredictor", "RoIHeads", "TwoMLPHead", "FastRCNNPredictor", "fastrcnn_loss", "paste_masks_in_image", "postprocess", "install",
           "maskrcnn_loss", "maskrcnn_inference", "project_masks_on_boxes", "pool_pair", "ops", "_lib"]
