"""Golden fixture for the windowing / slicing / zero-padding host logic, produced by the UNMODIFIED
reference: SegmentationModel.compute_maskrcnn_features / _slice_features / _index_features (code/helpers/model.py:190-266)
are called unbound on a stub whose "backbone" tags every frame with its index.

Run in the build container only (needs /root/reference):  python tests/golden/make_window_golden.py
Writes tests/golden/window.npz: for every (n_frames, fp, sp) scenario and every centre frame, the frame ids of the fast
window and of the slow slice (0 = zero-padded frame) and the feature-cache keys left behind."""
import json
import os
import sys
import types
from collections import OrderedDict
from math import ceil, floor

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/code")
from helpers.model import SegmentationModel as Ref  # noqa: E402  (the reference itself; never instantiated)

SCENARIOS = [(n, fp, sp) for n in (1, 3, 10) for fp in (1, 7, 8, 16) for sp in (1, 3, 4) if sp <= fp]


def make_stub(cls):
    stub = types.SimpleNamespace()
    stub.device = torch.device("cpu")
    stub.features_cache = {}
    stub.use_caching = True
    stub.maskrcnn_model = types.SimpleNamespace(
        backbone=lambda img: OrderedDict([("0", img.expand(1, 2, 3, 4).clone()), ("pool", img.expand(1, 2, 2, 2).clone())]))
    for name in ("compute_maskrcnn_features", "_slice_features", "_index_features", "_detach_features"):
        setattr(stub, name, types.MethodType(getattr(cls, name), stub))
    return stub


def run(cls, n, fp, sp):
    stub = make_stub(cls)
    images = torch.arange(1, n + 1, dtype=torch.float32).view(n, 1, 1, 1)       # frame i carries the value i + 1
    out = []
    for idx in range(n):
        indices = range(idx - floor(fp / 2), idx + ceil(fp / 2))
        window = stub.compute_maskrcnn_features(images, indices)
        slow = stub._slice_features(window, fp // 2, sp)
        centre = stub._index_features(window, fp // 2, fp // 2 + 1)
        assert list(window.keys()) == ["0", "pool"]
        out.append({"fast": [int(v) for v in window["0"][:, 0, 0, 0]], "fast_pool_shape": list(window["pool"].shape),
                    "slow": [int(v) for v in slow["0"][:, 0, 0, 0]], "centre": [int(v) for v in centre["pool"][:, 0, 0, 0]],
                    "cache": sorted(int(k) for k in stub.features_cache)})
    return out


if __name__ == "__main__":
    golden = {f"{n},{fp},{sp}": run(Ref, n, fp, sp) for n, fp, sp in SCENARIOS}
    np.savez_compressed(os.path.join(HERE, "window.npz"), golden=np.array(json.dumps(golden)))
    print("wrote window.npz:", len(golden), "scenarios")
