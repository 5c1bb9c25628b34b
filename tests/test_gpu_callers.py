"""GPU test of the callers of the hot path: the call sequences of the reference's scripts replayed on a synthetic DAVIS-shaped
dataset ON DISK (tests/davis_like.py) through the reference's import path (``helpers.model`` via sfvos_b200.run_reference's
hook).  The GPU box has no reference tree, so the loops below restate what the scripts do around the model -- the scripts
themselves run unchanged in tests/test_reference_scripts.py where the tree exists:

  * code/prediction.py:8-14 + code/helpers/evaluation.py:16-81: build SegmentationModel(device, sp, fp) -> .to(device) ->
    load_state_dict(torch.load(best_model_path)) -> DataLoader(dataset, batch_size=None) -> model.eval() -> under no_grad
    ``_, detections = model(imgs, deepcopy(targets))`` -> OR of the thresholded masks per frame -> IoU with the ground truth;
  * code/train.py:61-121: SGD(model.parameters(), lr, momentum=0.9, weight_decay) -> model.train() ->
    ``batch_loss, _ = model(imgs, targets, optimizer=opt)`` per sequence (backward and optimizer steps happen inside) ->
    torch.save(model.state_dict()) + {'epoch', 'optimizer_state_dict'} -> resume with load_state_dict on both."""
import os
from copy import deepcopy

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from davis_like import DavisLikeDataset, write_davis_tree

pytestmark = pytest.mark.gpu


def _model(sp, fp):
    from sfvos_b200.run_reference import patch_reference
    patch_reference()
    import importlib
    import types
    import sys
    if "helpers" not in sys.modules:                      # no reference tree here: an empty package named like the reference's
        pkg = types.ModuleType("helpers")
        pkg.__path__ = []
        sys.modules["helpers"] = pkg
    SegmentationModel = importlib.import_module("helpers.model").SegmentationModel      # the reference's import path
    dev = torch.device("cuda")
    torch.manual_seed(63)
    m = SegmentationModel(device=dev, slow_pathway_size=sp, fast_pathway_size=fp, maskrcnn_weights=None, pretrained=False)
    m.to(dev)
    return m


def test_prediction_and_training_call_sequences(tmp_path):
    root = str(tmp_path / "DAVIS")
    write_davis_tree(root, sequences=(("seq_a", 4), ("seq_b", 3)), hw=(120, 160), empty_frame=("seq_a", 2))
    sp, fp = 1, 4
    model = _model(sp, fp)

    # ---- train.py: one epoch over the sequences, checkpoint, resume ----------------------------------------------------
    loader = DataLoader(DavisLikeDataset(root, "train"), batch_size=None)
    opt = torch.optim.SGD(model.parameters(), lr=0.001, momentum=0.9, weight_decay=0.0001)     # all parameters, like train.py:80
    total_loss = 0.0
    for imgs, targets, _ in loader:
        model.train()
        batch_loss, _ = model(imgs, targets, optimizer=opt)
        assert isinstance(batch_loss, float) and np.isfinite(batch_loss) and batch_loss > 0
        total_loss += batch_loss
    model_path, ckpt_path = str(tmp_path / "model.pth"), str(tmp_path / "ckpt.pth")
    torch.save(model.state_dict(), model_path)
    torch.save({"epoch": 0, "optimizer_state_dict": opt.state_dict()}, ckpt_path)
    resumed = _model(sp, fp)
    opt2 = torch.optim.SGD(resumed.parameters(), lr=0.001, momentum=0.9, weight_decay=0.0001)
    opt2.load_state_dict(torch.load(ckpt_path)["optimizer_state_dict"])       # index-keyed: needs the reference's parameter order
    resumed.load_state_dict(torch.load(model_path))
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), resumed.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    # the momentum buffers line up with the parameters they belong to
    for p, p2 in zip(model.parameters(), resumed.parameters()):
        s1, s2 = opt.state.get(p, {}), opt2.state.get(p2, {})
        assert ("momentum_buffer" in s1) == ("momentum_buffer" in s2)
        if "momentum_buffer" in s1:
            assert s1["momentum_buffer"].shape == p.shape and torch.equal(s1["momentum_buffer"], s2["momentum_buffer"])
    assert any("momentum_buffer" in opt.state.get(p, {}) for p in model.slow_fast.parameters())

    # ---- prediction.py / evaluation.evaluate -------------------------------------------------------------------------
    loader = DataLoader(DavisLikeDataset(root, "val"), batch_size=None)
    resumed.eval()
    ious = []
    for imgs, targets, seq_name in loader:
        with torch.no_grad():
            _, detections = resumed(imgs, deepcopy(targets))
        assert len(detections) == len(imgs)
        for i, target in enumerate(targets):
            if len(target) == 0:
                assert detections[i] == {}                      # frames without objects are skipped (model.py:289-296,377-387)
                continue
            gt = np.zeros(imgs[i].shape[-2:], dtype=bool)
            for m in target["masks"]:
                gt |= m.cpu().numpy() >= 0.5
            pred = np.zeros_like(gt)
            assert detections[i]["masks"].device.type == "cpu" and len(detections[i]["masks"]) <= 10
            for m in detections[i]["masks"]:
                pred |= (m.cpu().numpy() >= 0.5)[0]
            union = np.logical_or(gt, pred).sum()
            ious.append(np.logical_and(gt, pred).sum() / union if union else 0.0)
    assert len(ious) == 6 and all(0.0 <= v <= 1.0 for v in ious)
