"""CPU test of the windowing / slicing / zero-padding host logic of the SegmentationModel mirror
against tests/golden/window.npz, which the unmodified reference produced (tests/golden/make_window_golden.py)."""
import importlib.util
import json
import os

import numpy as np

from conftest import GOLDEN


def _load_generator():
    spec = importlib.util.spec_from_file_location("make_window_golden", os.path.join(GOLDEN, "make_window_golden.py"))
    src = open(spec.origin).read()
    # reuse the generator's scenario driver, but never its import of the reference
    src = src.replace('sys.path.insert(0, "/root/reference/code")', "").replace(
        "from helpers.model import SegmentationModel as Ref  # noqa: E402  (the reference itself; never instantiated)", "")
    mod = {"__file__": spec.origin, "__name__": "make_window_golden"}
    exec(compile(src, spec.origin, "exec"), mod)
    return mod


def test_window_slice_and_padding_match_reference_golden():
    from sfvos_b200.model import SegmentationModel
    gen = _load_generator()
    golden = json.loads(str(np.load(os.path.join(GOLDEN, "window.npz"))["golden"]))
    assert len(golden) == len(gen["SCENARIOS"]) >= 30
    for n, fp, sp in gen["SCENARIOS"]:
        got = gen["run"](SegmentationModel, n, fp, sp)
        assert got == golden[f"{n},{fp},{sp}"], (n, fp, sp)


def test_zero_padded_frames_are_exact_zeros_and_slow_window_is_centred():
    from sfvos_b200.model import SegmentationModel
    gen = _load_generator()
    out = gen["run"](SegmentationModel, 10, 8, 1)
    assert out[0]["fast"] == [0, 0, 0, 0, 1, 2, 3, 4] and out[0]["slow"] == [1]
    assert out[9]["fast"] == [6, 7, 8, 9, 10, 0, 0, 0] and out[9]["slow"] == [10]
    out = gen["run"](SegmentationModel, 10, 7, 3)
    assert out[5]["fast"] == [3, 4, 5, 6, 7, 8, 9] and out[5]["slow"] == [5, 6, 7] and out[5]["centre"] == [6]
