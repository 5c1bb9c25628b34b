"""Run an UNCHANGED script of the reference (train.py, prediction.py, helpers/davis_evaluate.py, osvos/*.py ...) with the
hot path swapped in:

    cd /path/to/reference/code
    PYTHONPATH=/path/to/this/repo python -m sfvos_b200.run_reference train.py [script args]

Why a launcher and not just PYTHONPATH: ``python train.py`` puts the script's own directory at sys.path[0], ahead of
every PYTHONPATH entry, so ``from helpers.model import SegmentationModel`` (code/train.py:49, code/prediction.py:5,
code/helpers/davis_evaluate.py:17, code/osvos/osvos_model.py) always finds the reference's ``helpers/model.py`` first.
``patch_reference()`` therefore installs an import hook for exactly ONE module name, ``helpers.model``; every other
module of the reference (helpers.constants / dataset / evaluation / utils, davis2017_evaluation, ...) is imported from the
reference tree untouched.  The script then runs under ``runpy`` as ``__main__`` with sys.path[0] = its directory, as
``python script.py`` would.

``patch_reference()`` also papers over API drift between the reference's pins (torch>=1.5, numpy<1.24, Pillow<10:
code/requirements.txt) and current releases, all OUTSIDE the hot path: numpy's removed ``np.float`` / ``np.int``
/ ``np.bool`` aliases (helpers/evaluation.py:44,58, helpers/davis_evaluate.py:38), Pillow's ``Image.ANTIALIAS``
(helpers/utils.py:11), and ``torch.as_tensor(list of bool ndarrays, dtype=uint8)`` (helpers/dataset.py:124), which current
torch / numpy 2 reject element-wise -- the list is stacked first, which is what old releases did implicitly.  Missing checkpoints / hub weights are handled by SegmentationModel itself (warning + initial
weights)."""
import importlib.abc
import importlib.util
import os
import runpy
import sys

_HOOKED = "helpers.model"


class _HelpersModelFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``helpers.model`` -> the libsfvos-backed classes; nothing else is intercepted."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname == _HOOKED:
            return importlib.util.spec_from_loader(fullname, self, origin="sfvos_b200.model")
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        from .model import SegmentationModel, get_model_instance_segmentation
        from .slowfast import SlowFastLayers
        module.SegmentationModel = SegmentationModel
        module.SlowFastLayers = SlowFastLayers
        module.get_model_instance_segmentation = get_model_instance_segmentation
        module.__file__ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model.py")


def patch_reference():
    """Idempotent: install the ``helpers.model`` hook and the numpy / Pillow compatibility aliases."""
    if not any(isinstance(f, _HelpersModelFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _HelpersModelFinder())
    sys.modules.pop(_HOOKED, None)                       # a previously imported reference module would win otherwise
    import numpy as np
    for name, typ in (("float", float), ("int", int), ("bool", bool), ("object", object)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    import torch
    if not getattr(torch.as_tensor, "_sfvos_compat", False):
        _as_tensor = torch.as_tensor

        def as_tensor(data, *args, **kwargs):
            if isinstance(data, (list, tuple)) and data and all(isinstance(d, np.ndarray) for d in data):
                data = np.stack(data)
            return _as_tensor(data, *args, **kwargs)
        as_tensor._sfvos_compat = True
        torch.as_tensor = as_tensor
    try:
        from PIL import Image
        if not hasattr(Image, "ANTIALIAS"):
            Image.ANTIALIAS = Image.LANCZOS
    except ImportError:
        pass


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m sfvos_b200.run_reference <reference script.py> [args...]", file=sys.stderr)
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        print(f"run_reference: {script} not found", file=sys.stderr)
        return 2
    patch_reference()
    sys.argv = [script] + argv[1:]
    sys.path.insert(0, os.path.dirname(script))          # what ``python script.py`` does
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
