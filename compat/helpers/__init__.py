"""Shim with the reference's package name for code that imports it from OUTSIDE the reference tree (a notebook, a test):
with ``<repo>/compat`` ahead of the reference's ``code/`` directory on sys.path, ``helpers.model`` resolves to the
libsfvos-backed classes (``compat/helpers/model.py``) and every other submodule -- ``helpers.constants``, ``helpers.dataset``,
``helpers.evaluation``, ``helpers.utils``, ``helpers.davis_evaluate`` -- to the reference's own file: this package appends the
reference's ``helpers/`` directory (the first other ``helpers`` package found on sys.path) to its ``__path__``.

For the reference's SCRIPTS use ``python -m sfvos_b200.run_reference train.py`` instead: ``python train.py`` puts the script's
directory ahead of PYTHONPATH, so this shim would never be found."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
for _entry in list(_sys.path):
    _cand = _os.path.abspath(_os.path.join(_entry or _os.getcwd(), "helpers"))
    if _cand != _here and _os.path.isfile(_os.path.join(_cand, "__init__.py")):
        __path__.append(_cand)
        break
