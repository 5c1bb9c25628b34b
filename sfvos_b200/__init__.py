"""Importable alias of the product package.

The contract names the package directory ``applying-slowfast-networks-to-video-object-segmentation_b200`` (hyphens
are not importable), so this alias package points its ``__path__`` at that directory: ``import sfvos_b200`` and
``from sfvos_b200.slowfast import SlowFastLayers`` resolve to the files there.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "applying-slowfast-networks-to-video-object-segmentation_b200")
__path__.insert(0, _REAL)

from ._api import *  # noqa: E402,F401,F403
