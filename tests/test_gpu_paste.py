"""GPU parity of the mask paste-back kernel against torchvision's paste_masks_in_image /
GeneralizedRCNNTransform.postprocess (the step after the hot path, code/helpers/model.py:347) and the CPU oracle."""
import pytest
import torch
from torchvision.models.detection.roi_heads import paste_masks_in_image as tv_paste
from torchvision.models.detection.transform import GeneralizedRCNNTransform

from oracle import roi_oracle as ro
from test_oracle import _paste_cases

pytestmark = pytest.mark.gpu


def test_paste_masks_matches_torchvision_and_oracle():
    from sfvos_b200 import paste_masks_in_image
    masks, boxes, shape = _paste_cases()
    got = paste_masks_in_image(masks.cuda(), boxes.cuda(), shape)
    ref_gpu = tv_paste(masks.cuda(), boxes.cuda(), shape)
    ref_cpu = ro.paste_masks_in_image(masks, boxes, shape)
    assert got.shape == ref_gpu.shape and got.dtype == ref_gpu.dtype
    assert (got - ref_gpu).abs().max().item() <= 2e-6
    assert (got.cpu() - ref_cpu).abs().max().item() <= 2e-6
    assert ((got > 1e-5) == (ref_gpu > 1e-5)).all()           # identical integer boxes and clipping
    assert paste_masks_in_image(masks[:0].cuda(), boxes[:0].cuda(), shape).shape == (0, 1, 120, 160)
    outside = paste_masks_in_image(masks[:1].cuda(), torch.tensor([[-50.0, -40.0, -10.0, -5.0]]).cuda(), shape)
    assert outside.abs().sum().item() == 0.0                  # torchvision raises here; detections never are outside


def test_paste_masks_davis_size_many_detections():
    from sfvos_b200 import paste_masks_in_image
    g = torch.Generator().manual_seed(2)
    K, shape = 40, (480, 854)
    masks = torch.rand(K, 1, 28, 28, generator=g).cuda()
    x1 = torch.rand(K, generator=g) * 700
    y1 = torch.rand(K, generator=g) * 400
    boxes = torch.stack([x1, y1, (x1 + 8 + torch.rand(K, generator=g) * 500).clamp(max=854.0),
                         (y1 + 8 + torch.rand(K, generator=g) * 300).clamp(max=480.0)], 1).cuda()
    got = paste_masks_in_image(masks, boxes, shape)
    ref = tv_paste(masks, boxes, shape)
    assert (got - ref).abs().max().item() <= 2e-6
    # same integer boxes and clipping (a sample whose interpolation weight is ~1e-7 in one arithmetic and 0 in the other is not a footprint difference)
    assert ((got > 1e-5) == (ref > 1e-5)).all()


def test_postprocess_matches_transform_postprocess():
    from sfvos_b200 import postprocess
    g = torch.Generator().manual_seed(3)
    tr = GeneralizedRCNNTransform(800, 1333, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]).eval()
    image_shapes, original = [(749, 1333)] * 3, [(480, 854)] * 3

    def dets():
        gg = torch.Generator().manual_seed(4)
        out = []
        for n in (3, 0, 5):
            x1, y1 = torch.rand(n, generator=gg) * 900, torch.rand(n, generator=gg) * 500
            b = torch.stack([x1, y1, x1 + 20 + torch.rand(n, generator=gg) * 300, y1 + 20 + torch.rand(n, generator=gg) * 200], 1)
            out.append({"boxes": b.cuda(), "labels": torch.ones(n, dtype=torch.int64).cuda(), "scores": torch.rand(n, generator=gg).cuda(),
                        "masks": torch.rand(n, 1, 28, 28, generator=gg).cuda()})
        return out
    ref = tr.postprocess(dets(), image_shapes, original)
    got = postprocess(dets(), image_shapes, original)
    for a, b in zip(got, ref):
        assert torch.equal(a["boxes"], b["boxes"]) and a["masks"].shape == b["masks"].shape
        if a["masks"].numel():
            assert (a["masks"] - b["masks"]).abs().max().item() <= 2e-6
