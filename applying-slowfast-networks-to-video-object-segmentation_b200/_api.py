"""Public API of the product package (re-exported by the ``sfvos_b200`` alias)."""
from .slowfast import SlowFastLayers  # noqa: F401
from . import ops, _lib  # noqa: F401

__all__ = ["SlowFastLayers", "ops", "_lib"]
