"""GPU parity of ROIAlign / MultiScaleRoIAlign / mask head / mask predictor / mask loss (through the C ABI) against
the live torchvision modules the reference calls (code/helpers/model.py:346), the CPU oracle, and the golden
fixture tests/golden/roi_mask.npz.  ROI indexing (levels, output order) must be identical."""
import copy
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torchvision
from torchvision.models.detection.mask_rcnn import MaskRCNNHeads as TVHeads, MaskRCNNPredictor as TVPredictor
from torchvision.models.detection.roi_heads import maskrcnn_loss as tv_maskrcnn_loss, project_masks_on_boxes as tv_project
from torchvision.ops import MultiScaleRoIAlign as TVPool

from conftest import GOLDEN, report
from oracle import roi_oracle as ro

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _nerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def _check_grad(name, got, ref, precision):
    """Gradients that pass through ReLUs are compared in relative L2: a ReLU whose pre-activation sits within rounding
    distance of zero flips between two implementations and moves individual gradient entries by percents of the max while
    leaving the L2 error small.  Unlike the SlowFast fixtures (tests/golden/make_golden.py) the mask branch cannot be given
    an input with a ReLU margin: 4 conv layers + the ConvTranspose on even 9 ROIs are 3.6 M ReLU inputs with a standard
    deviation of ~0.1, i.e. an expected minimum |pre-activation| of ~7e-8 (a scan of 400 input seeds found none above
    4e-6), so a handful of undetermined masks is a property of the test, not of the kernels.  Measured (round 2, B200,
    gpurun_out/parity_report.jsonl -> profiles/parity_r2.jsonl): fp32 validation mode <= 3.2e-3 (the first head conv, which
    sees the flips of all layers above it; 5e-5..2e-4 elsewhere; 3e-7 downstream of every ReLU), bf16 <= 0.11 (first
    conv) / <= 2e-2 (other layers) - the sqrt(flip fraction) effect described in test_gpu_slowfast._check_grad."""
    rel = (got.double() - ref.double()).norm().item() / (ref.double().norm().item() + 1e-30)
    report("roi_mask_grad", name=name, precision=precision, rel_l2=rel, max_norm=_nerr(got, ref))
    assert rel < (1e-2 if precision == "fp32" else 0.25), (name, rel)
    if name.startswith(("mask_predictor.mask_fcn_logits", "mask_fcn_logits")):     # downstream of every ReLU: no masks involved
        # measured 3.9e-7 (fp32 mode) / 2.6e-3 (bf16)
        assert _nerr(got, ref) < (1e-5 if precision == "fp32" else 1e-2), (name, _nerr(got, ref))


def _feats(n=2, c=256, seed=0, shapes=((48, 84), (24, 42), (12, 21), (6, 11))):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((str(i), torch.randn(n, c, h, w, generator=g).cuda()) for i, (h, w) in enumerate(shapes))


def _edge_boxes():
    return torch.tensor([[10.0, 10.0, 10.4, 10.3], [0.0, 0.0, 333.0, 187.0], [5.0, 5.0, 61.0, 61.0],
                         [300.0, 150.0, 333.0, 187.0], [-20.0, -10.0, 40.0, 30.0], [320.0, 180.0, 400.0, 260.0]])


def test_roi_levels_identical_to_torchvision():
    from sfvos_b200 import ops
    from torchvision.ops.poolers import LevelMapper
    sides = [10, 111, 112, 223, 224, 447, 448, 896]
    boxes = torch.tensor([[0.0, 0.0, float(s), float(s)] for s in sides])
    rnd = torch.cat(ro.synthetic_rois(8, 4000, seed=99) + [boxes])
    rois = torch.cat([torch.zeros(len(rnd), 1), rnd], 1).cuda()
    got = ops.roi_levels(rois, 2, 5).cpu().long()
    assert torch.equal(got, LevelMapper(2, 5)([rnd]))
    assert got[-8:].tolist() == [0, 0, 1, 1, 2, 2, 3, 3]


@pytest.mark.parametrize("P", [7, 14])
def test_multiscale_roi_align_matches_torchvision_fwd_bwd(P):
    from sfvos_b200 import MultiScaleRoIAlign
    feats = _feats()
    image_shapes = [(187, 333)] * 2
    boxes = [b.cuda() for b in ro.synthetic_rois(2, 64, image_hw=image_shapes[0], seed=4321, lo=4.0, hi=175.0)]
    boxes[0] = torch.cat([boxes[0], _edge_boxes().cuda()])
    fa = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in feats.items())
    fb = OrderedDict((k, v.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)) for k, v in feats.items())
    ref_pool = TVPool(["0", "1", "2", "3"], P, 2)
    ref = ref_pool(fa, boxes, image_shapes)
    ours = MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)(fb, boxes, image_shapes)
    assert ours.shape == ref.shape and ours.is_contiguous()
    assert _nerr(ours, ref) < 1e-5
    g = torch.randn_like(ref)
    ref.backward(g)
    ours.backward(g)
    for k in fa:
        assert _nerr(fb[k].grad, fa[k].grad) < 1e-5, k
    # channels-last bf16 output feeding the mask head
    nhwc = MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2, out_layout="nhwc", precision="bf16")(feats, boxes, image_shapes)
    assert nhwc.shape == ref.shape and nhwc.dtype == torch.bfloat16 and nhwc.permute(0, 2, 3, 1).is_contiguous()
    assert _nerr(nhwc, ref) < 1e-2
    # empty box list for one image keeps the output order of the rest
    empty = [boxes[0][:0], boxes[1]]
    o2 = MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)(feats, empty, image_shapes)
    r2 = TVPool(["0", "1", "2", "3"], P, 2)(feats, empty, image_shapes)
    assert o2.shape == r2.shape and _nerr(o2, r2) < 1e-5


def test_pool_pair_equals_two_separate_poolings():
    """The shared-backward pair (box + mask branch of a training step in ONE autograd node) must give the outputs of
    the two separate poolings and the SUM of their feature gradients."""
    from sfvos_b200 import MultiScaleRoIAlign, pool_pair
    boxes = [b.cuda() for b in ro.synthetic_rois(2, 60, image_hw=(187, 333), seed=11, lo=4.0, hi=300.0)]
    boxes[0] = torch.cat([boxes[0], _edge_boxes().cuda()])
    sub = [b[:17] for b in boxes]
    shapes = [(187, 333)] * 2
    pa = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, out_layout="nchw", precision="fp32")
    pb = MultiScaleRoIAlign(["0", "1", "2", "3"], 14, 2, out_layout="nhwc", precision="fp32")
    f1 = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in _feats(seed=4).items())
    f2 = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in _feats(seed=4).items())
    oa, ob = pool_pair(pa, pb, f1, boxes, sub, shapes)
    ra, rb = pa(f2, boxes, shapes), pb(f2, sub, shapes)
    assert torch.equal(oa, ra) and torch.equal(ob, rb)
    g = torch.Generator().manual_seed(1)
    wa, wb = torch.randn(oa.shape, generator=g).cuda(), torch.randn(ob.shape, generator=g).cuda()
    ((oa * wa).sum() + (ob * wb).sum()).backward()
    ((ra * wa).sum() + (rb * wb).sum()).backward()
    for k in f1:
        assert _nerr(f1[k].grad, f2[k].grad) < 1e-5, k
    # only one branch used: the other's gradient is simply absent
    f3 = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in _feats(seed=4).items())
    oa3, _ = pool_pair(pa, pb, f3, boxes, sub, shapes)
    (oa3 * wa).sum().backward()
    f4 = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in _feats(seed=4).items())
    (pa(f4, boxes, shapes) * wa).sum().backward()
    for k in f3:
        assert _nerr(f3[k].grad, f4[k].grad) < 1e-5, k


def test_roi_and_mask_golden_fixture():
    from sfvos_b200 import MultiScaleRoIAlign, MaskRCNNHeads, MaskRCNNPredictor, maskrcnn_loss
    gold = np.load(os.path.join(GOLDEN, "roi_mask.npz"))
    feats = OrderedDict((str(i), torch.from_numpy(gold["feat" + str(i)]).cuda()) for i in range(4))
    boxes = [torch.from_numpy(gold["boxes0"]).cuda(), torch.from_numpy(gold["boxes1"]).cuda()]
    for P in (7, 14):
        out = MultiScaleRoIAlign(["0", "1", "2", "3"], P, 2)(feats, boxes, [(187, 333)] * 2)
        assert _nerr(out, torch.from_numpy(gold[f"pool{P}"])) < 1e-5
    from sfvos_b200 import ops
    rois = torch.cat([torch.zeros(sum(len(b) for b in boxes), 1).cuda(), torch.cat(boxes)], 1)
    assert torch.equal(ops.roi_levels(rois, 2, 5).cpu().long(), torch.from_numpy(gold["levels"]))
    torch.manual_seed(11)
    head = MaskRCNNHeads(256, (256, 256, 256, 256), 1).cuda()
    pred = MaskRCNNPredictor(256, 256, 2).cuda()
    for prec, tol in (("fp32", 1e-4), ("bf16", 1e-2)):      # north star: <= 1e-4 / <= 1e-2; measured 5.7e-7 / 6.4e-3
        head.precision = pred.precision = prec
        x = torch.from_numpy(gold["mh_x"]).cuda()
        logits = pred(head(x))
        report("mask_logits_golden", precision=prec, max_norm=_nerr(logits, torch.from_numpy(gold["mh_logits"])))
        assert _nerr(logits, torch.from_numpy(gold["mh_logits"])) < tol, prec
        gt = torch.zeros(2, 187, 333, dtype=torch.uint8)
        gt[0, 40:120, 60:200] = 1
        gt[1, 100:180, 150:300] = 1
        props = [torch.tensor([[50.0, 30.0, 210.0, 130.0], [140.0, 90.0, 310.0, 186.0], [0.0, 0.0, 20.5, 17.25]]).cuda()]
        loss = maskrcnn_loss(logits, props, [gt.cuda()], [torch.tensor([1, 1]).cuda()], [torch.tensor([0, 1, 0]).cuda()])
        assert abs(loss.item() - float(gold["mh_loss"])) < tol * abs(float(gold["mh_loss"])), prec


def test_mask_targets_match_torchvision():
    from sfvos_b200 import project_masks_on_boxes
    g = torch.Generator().manual_seed(5)
    masks = (torch.rand(3, 120, 200, generator=g) > 0.5).to(torch.uint8).cuda()
    boxes = torch.cat(ro.synthetic_rois(1, 40, image_hw=(120, 200), seed=7, lo=3.0, hi=190.0)).cuda()
    boxes = torch.cat([boxes, torch.tensor([[0.0, 0.0, 200.0, 120.0], [10.0, 10.0, 10.5, 10.2], [-5.0, -5.0, 30.0, 20.0]]).cuda()])
    idx = torch.randint(0, 3, (len(boxes),), generator=g).cuda()
    ref = tv_project(masks, boxes, idx, 28)
    out = project_masks_on_boxes(masks, boxes, idx, 28)
    assert out.shape == ref.shape and _nerr(out, ref) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mask_branch_forward_backward_matches_torchvision(precision):
    from sfvos_b200 import MaskRCNNHeads, MaskRCNNPredictor, maskrcnn_loss
    torch.manual_seed(3)
    head_ref, pred_ref = TVHeads(256, (256, 256, 256, 256), 1).cuda(), TVPredictor(256, 256, 2).cuda()
    head, pred = MaskRCNNHeads(256, (256, 256, 256, 256), 1).cuda(), MaskRCNNPredictor(256, 256, 2).cuda()
    head.load_state_dict(head_ref.state_dict())
    pred.load_state_dict(pred_ref.state_dict())
    assert list(head.state_dict().keys()) == list(head_ref.state_dict().keys())
    head.precision = pred.precision = precision
    K = 9
    g = torch.Generator().manual_seed(1)
    x = torch.randn(K, 256, 14, 14, generator=g).cuda()
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    gt = (torch.rand(2, 96, 160, generator=g) > 0.5).to(torch.uint8).cuda()
    props = [torch.cat(ro.synthetic_rois(1, K, image_hw=(96, 160), seed=3, lo=8.0, hi=150.0)).cuda()]
    labels, matched = [torch.tensor([1, 1]).cuda()], [torch.randint(0, 2, (K,), generator=g).cuda()]
    if precision == "fp32":              # the exact reference for the validation mode: torchvision's modules in fp64
        head_ref, pred_ref, xr = head_ref.double(), pred_ref.double(), x.double().clone().requires_grad_(True)
    lr = pred_ref(head_ref(xr))
    loss_r = tv_maskrcnn_loss(lr, props, [gt], labels, matched)
    loss_r.backward()
    lo = pred(head(xo))
    loss_o = maskrcnn_loss(lo, props, [gt], labels, matched)
    loss_o.backward()
    ftol = 1e-4 if precision == "fp32" else 1e-2         # north star bounds; measured 7.2e-6 / 6.9e-3
    report("mask_branch_logits", precision=precision, max_norm=_nerr(lo, lr), loss_rel=abs(loss_o.item() - loss_r.item()) / abs(loss_r.item()))
    assert lo.shape == lr.shape and _nerr(lo, lr) < ftol
    assert abs(loss_o.item() - loss_r.item()) < ftol * abs(loss_r.item())
    for (n1, p1), (n2, p2) in zip(list(head.named_parameters()) + list(pred.named_parameters()),
                                  list(head_ref.named_parameters()) + list(pred_ref.named_parameters())):
        assert n1 == n2
        _check_grad(n1, p1.grad, p2.grad, precision)
    _check_grad("x", xo.grad, xr.grad, precision)


def _roi_heads_pair(precision):
    from sfvos_b200 import install
    torch.manual_seed(5)
    model = torchvision.models.detection.maskrcnn_resnet50_fpn(weights=None, weights_backbone=None, num_classes=2)
    ref = model.roi_heads.cuda()
    ours = install(copy.deepcopy(ref), precision=precision)
    return ref, ours


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_roi_heads_train_and_eval_match_torchvision(precision):
    ref, ours = _roi_heads_pair(precision)
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    feats = _feats(n=1, seed=2, shapes=((48, 84), (24, 42), (12, 21), (6, 11)))
    feats["pool"] = torch.randn(1, 256, 3, 6).cuda()
    image_shapes = [(187, 333)]
    props = [torch.cat(ro.synthetic_rois(1, 300, image_hw=image_shapes[0], seed=8, lo=6.0, hi=170.0)).cuda()]
    gt_boxes = torch.tensor([[60.0, 40.0, 200.0, 120.0], [150.0, 100.0, 300.0, 180.0]]).cuda()
    masks = torch.zeros(2, 187, 333, dtype=torch.uint8).cuda()
    masks[0, 40:120, 60:200] = 1
    masks[1, 100:180, 150:300] = 1
    targets = [{"boxes": gt_boxes, "labels": torch.tensor([1, 1]).cuda(), "masks": masks}]
    tol = 1e-4 if precision == "fp32" else 2e-2
    ref.train(); ours.train()
    torch.manual_seed(0)
    _, l_ref = ref(feats, [p.clone() for p in props], image_shapes, copy.deepcopy(targets))
    torch.manual_seed(0)
    _, l_our = ours(feats, [p.clone() for p in props], image_shapes, copy.deepcopy(targets))
    assert set(l_ref) == set(l_our)
    for k in l_ref:
        assert abs(l_our[k].item() - l_ref[k].item()) <= tol * max(1.0, abs(l_ref[k].item())), (k, l_our[k].item(), l_ref[k].item())
    sum(l_ref.values()).backward()
    sum(l_our.values()).backward()
    gr = dict(ref.named_parameters())
    for n, p in ours.named_parameters():
        _check_grad(n, p.grad, gr[n].grad, precision)
    ref.eval(); ours.eval()
    ref.score_thresh = ours.score_thresh = 0.0          # random-init scores are ~0.5: keep detections
    with torch.no_grad():
        d_ref, _ = ref(feats, props, image_shapes)
        d_our, _ = ours(feats, props, image_shapes)
    assert len(d_ref) == len(d_our) == 1
    assert d_ref[0]["boxes"].shape == d_our[0]["boxes"].shape
    if precision == "fp32":
        assert torch.equal(d_ref[0]["labels"], d_our[0]["labels"])
        assert _nerr(d_our[0]["boxes"], d_ref[0]["boxes"]) < 1e-4
        assert _nerr(d_our[0]["masks"], d_ref[0]["masks"]) < 1e-3

@pytest.mark.parametrize("n_cls,pixel_order", [(2, 1), (2, 0), (1, 1)])
def test_mask_logits_relu_bwd_fast_path_equals_the_general_kernel(n_cls, pixel_order, monkeypatch):
    """bf16 / C = 256 / <= 2 classes runs a kernel with four rows in flight per warp (sfvos_mask_logits_relu_bwd); it keeps the
    general kernel's pixel assignment and accumulation order, so dx must agree bit for bit (the reduced quantities up to the order of the CTAs' atomics) - and with a torch restatement of
    d/dx [logits = W relu-output + b] followed by the ReLU backward (TV mask_rcnn.py:342-344)."""
    from sfvos_b200 import ops
    from sfvos_b200._lib import call
    K, S, C = 5, 28, 256
    g = torch.Generator().manual_seed(11)
    x = torch.randn(K, S * S, C, generator=g).relu().bfloat16().to(DEV)            # rows in kernel order
    w = (torch.randn(n_cls, C, generator=g) / 16).to(DEV)
    gl = torch.randn(K, n_cls, S, S, generator=g).to(DEV)
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("SFVOS_MASK_LOGITS_GENERIC", "1")
        dx = torch.empty_like(x)
        dw, db, dbx = torch.zeros(n_cls, C, device=DEV), torch.zeros(n_cls, device=DEV), torch.zeros(C, device=DEV)
        call("sfvos_mask_logits_relu_bwd", ops._p(x), ops.BF16, ops._p(w), ops._p(gl), ops._p(dx), ops.BF16, ops._p(dw), ops._p(db),
             ops._p(dbx), K, S, C, n_cls, pixel_order, ops.stream())
        torch.cuda.synchronize()
        outs.append((dx, dw, db, dbx))
    assert torch.equal(outs[0][0], outs[1][0])                    # dx: same arithmetic per element
    for a, b in zip(outs[0][1:], outs[1][1:]):                    # per-CTA partial sums are equal; the CTAs' atomics land in any order
        assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item() + 1e-7
    # torch reference (rows -> spatial order for pixel_order 1: row (h*(S/2)+w)*4 + 2i + j = pixel (2h+i, 2w+j))
    if pixel_order:
        r = torch.arange(S * S, device=DEV)
        tap, lw = r & 3, r >> 2
        h, ww = lw // (S // 2), lw % (S // 2)
        sp = (2 * h + (tap >> 1)) * S + 2 * ww + (tap & 1)
    else:
        sp = torch.arange(S * S, device=DEV)
    g_rows = gl.reshape(K, n_cls, S * S)[:, :, sp]                                   # [K, cls, row]
    xf = x.float()
    dx_ref = torch.einsum("kcr,cd->krd", g_rows, w) * (xf > 0)
    dw_ref = torch.einsum("kcr,krd->cd", g_rows, xf)
    dx, dw, db, dbx = outs[0]
    assert (dx.float() - dx_ref).abs().max().item() <= 4e-3 * dx_ref.abs().max().item()          # one bf16 rounding
    assert (dw - dw_ref).abs().max().item() <= 1e-4 * dw_ref.abs().max().item()
    assert (db - g_rows.sum((0, 2))).abs().max().item() <= 1e-4 * g_rows.sum((0, 2)).abs().max().item() + 1e-5
    assert (dbx - dx_ref.sum((0, 1))).abs().max().item() <= 1e-3 * dx_ref.sum((0, 1)).abs().max().item()
