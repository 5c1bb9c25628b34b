"""GPU smoke/contract test of the caller of the hot path: the SegmentationModel mirror (code/helpers/model.py:168-389)
run exactly the way train.py / evaluation.py call it -- ``model(imgs, targets, optimizer=opt)`` in train mode,
``model(imgs, targets)`` in eval mode -- on a short synthetic DAVIS-like sequence (moving rectangle), random-init
Mask R-CNN (no network for the pretrained weights), SlowFast module / ROIAlign / mask branch on libsfvos.so."""
import sys
import os

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _sequence(n=3, h=120, w=160):
    g = torch.Generator().manual_seed(0)
    imgs, targets = [], []
    for i in range(n):
        imgs.append(torch.rand(3, h, w, generator=g))
        x1, y1 = 20 + 10 * i, 30 + 5 * i
        x2, y2 = x1 + 60, y1 + 50
        m = torch.zeros(1, h, w, dtype=torch.uint8)
        m[0, y1:y2, x1:x2] = 1
        targets.append({"boxes": torch.tensor([[x1, y1, x2, y2]], dtype=torch.float32), "labels": torch.ones(1, dtype=torch.int64),
                        "masks": m})
    return imgs, targets


def test_segmentation_model_train_and_eval_contract():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from helpers.model import SegmentationModel, SlowFastLayers          # the reference's import path, via the shim
    from sfvos_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(63)
    model = SegmentationModel(device=dev, slow_pathway_size=1, fast_pathway_size=4, maskrcnn_weights=None, pretrained=False)
    model.to(dev)
    assert isinstance(model.slow_fast, SlowFastLayers)
    keys = list(model.state_dict().keys())
    assert keys[0].startswith("maskrcnn_model.") and any(k.startswith("slow_fast.fast_conv1") for k in keys)
    assert keys.index("slow_fast.fast_conv1.weight") > keys.index("maskrcnn_model.roi_heads.mask_predictor.mask_fcn_logits.bias")
    trainable = [p for p in model.parameters() if p.requires_grad]
    assert all(not p.requires_grad for p in model.maskrcnn_model.backbone.parameters())
    opt = torch.optim.SGD(trainable, lr=1e-4, momentum=0.9)
    imgs, targets = _sequence()

    model.train()
    before = ops.launches()
    w0 = model.slow_fast.slow_conv3.weight.detach().clone()
    loss, dets = model(imgs, targets, optimizer=opt)
    torch.cuda.synchronize()
    assert ops.launches() - before > 100                                   # the libsfvos kernels did the work
    assert isinstance(loss, float) and loss == loss and loss > 0
    assert dets == []                                                      # torchvision's RoIHeads returns no detections in train mode
    assert not torch.equal(w0, model.slow_fast.slow_conv3.weight.detach())  # optimizer stepped after the 2nd frame
    g = model.slow_fast.fast_conv1.weight.grad                              # the 3rd frame's gradient is still there
    assert g is not None and torch.isfinite(g).all() and g.abs().sum() > 0
    assert model.slow_fast.bn_s1.num_batches_tracked.item() == 3 * 5        # 3 frames x 5 pyramid levels

    model.eval()
    with torch.no_grad():
        loss, dets = model(imgs, targets)
    assert loss == 0. and len(dets) == 3
    for d in dets:
        assert set(d.keys()) == {"boxes", "labels", "scores", "masks"}
        assert d["masks"].device.type == "cpu" and d["masks"].shape[1:] == (1, 120, 160) and d["boxes"].shape[0] <= 10
