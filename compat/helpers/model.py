"""Shim with the reference's module path: put ``<repo>/compat`` (and ``<repo>``) ahead of the reference's ``code/``
directory on sys.path and ``from helpers.model import SegmentationModel, SlowFastLayers`` -- as train.py:9,
prediction.py:3 and osvos/osvos_model.py do -- resolves to the libsfvos-backed classes."""
from sfvos_b200.model import SegmentationModel, get_model_instance_segmentation  # noqa: F401
from sfvos_b200.slowfast import SlowFastLayers  # noqa: F401
