"""GPU parity of the box branch through the C ABI: the 256-column N tiling of the tcgen05 conv / wgrad
kernels that turns them into the fc6 / fc7 GEMMs, the tiled fc weight transposes, fastrcnn_loss forward / backward, and the
TwoMLPHead / FastRCNNPredictor drop-ins against the live torchvision modules the reference calls (code/helpers/model.py:346),
the CPU oracle and the golden fixture tests/golden/box_head.npz."""
import os

import numpy as np
import pytest
import torch
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor as TVPredictor, TwoMLPHead as TVHead
from torchvision.models.detection.roi_heads import fastrcnn_loss as tv_fastrcnn_loss

from conftest import GOLDEN
from oracle import roi_oracle as ro

pytestmark = pytest.mark.gpu


def _nerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).norm().item() / (b.norm().item() + 1e-20)


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("M,K,N", [(300, 1024, 1024), (77, 12544, 512), (256, 64, 1024), (130, 1024, 64)])
def test_fc_as_conv_n_tiles_match_matmul(M, K, N):
    """y = relu(x W^T + b) through sfvos_conv_umma with N tiled in 256-column chunks (ragged M: TMA zero-fill + row masks)."""
    from sfvos_b200 import ops
    from sfvos_b200._lib import BF16
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    xa = ops.Act(x.to(torch.bfloat16).reshape(-1), 1, 1, 1, M, K)
    wp = ops.pack_weights(w.view(N, K, 1, 1, 1), 0, BF16, K)
    for dtype in (torch.float32, torch.bfloat16):
        y = ops.Act.empty(1, 1, 1, M, N, dtype, x.device)
        ops.conv(xa, wp, K, N, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=b)
        ref = torch.relu(_bf16(x).double() @ _bf16(w).double().t() + b.double()).float()
        got = y.buf.view(M, N).float()
        assert _nerr(got, ref) < (2e-5 if dtype == torch.float32 else 5e-3), (dtype, _nerr(got, ref))


@pytest.mark.parametrize("M,K,N", [(300, 1024, 1024), (200, 12544, 512), (4096, 1024, 64)])
def test_fc_wgrad_and_dgrad_n_tiles_match_matmul(M, K, N):
    from sfvos_b200 import ops
    from sfvos_b200._lib import BF16
    g = torch.Generator().manual_seed(M + K + N + 1)
    x = torch.randn(M, K, generator=g).cuda()
    dy = torch.randn(M, N, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / N ** 0.5).cuda()
    xa = ops.Act(x.to(torch.bfloat16).reshape(-1), 1, 1, 1, M, K)
    da = ops.Act(dy.to(torch.bfloat16).reshape(-1), 1, 1, 1, M, N)
    dwp = torch.zeros(K * N, device=x.device)
    ops.wgrad(xa, da, (1, 1, 1), (0, 0, 0), dwp, umma=True)
    gw = torch.zeros(N, K, device=x.device)
    ops.unpack_wgrad(dwp, gw.view(N, K, 1, 1, 1), 0)
    ref = (_bf16(dy).double().t() @ _bf16(x).double()).float()
    assert _nerr(gw, ref) < 2e-5
    ops.unpack_wgrad(dwp, gw.view(N, K, 1, 1, 1), 0)              # accumulates
    assert _nerr(gw, 2 * ref) < 2e-5
    # dgrad: dx = dy W, operand packed by the tiled transpose
    wd = ops.pack_weights(w.view(N, K, 1, 1, 1), 1, BF16, (N + 63) // 64 * 64)
    assert torch.equal(wd.view(K, -1)[:, :N], w.t().to(torch.bfloat16))
    if K % 256 == 0 or K <= 256:
        dx = ops.Act.empty(1, 1, 1, M, K, torch.float32, x.device)
        ops.conv(da, wd, (N + 63) // 64 * 64, K, (1, 1, 1), (0, 0, 0), 1, dx, umma=True)
        refx = (_bf16(dy).double() @ _bf16(w).double()).float()
        assert _nerr(dx.buf.view(M, K), refx) < 2e-5


@pytest.mark.parametrize("n_cls", [2, 5])
def test_fastrcnn_loss_matches_torchvision_and_oracle(n_cls):
    from sfvos_b200 import fastrcnn_loss
    g = torch.Generator().manual_seed(n_cls)
    M = 1500
    fused = torch.randn(M, 64, generator=g).cuda()               # column slices of a wider buffer, like the predictor output
    fused[:, n_cls:] *= 0.2
    fused.requires_grad_(True)
    labels = [torch.randint(0, n_cls, (700,), generator=g).cuda(), torch.randint(0, n_cls, (800,), generator=g).cuda()]
    targets = [(0.2 * torch.randn(700, 4, generator=g)).cuda(), (0.2 * torch.randn(800, 4, generator=g)).cuda()]
    z, r = fused[:, :n_cls], fused[:, n_cls:5 * n_cls]
    lc, lb = fastrcnn_loss(z, r, labels, targets)
    (2.0 * lc + 0.5 * lb).backward()
    f2 = fused.detach().clone().requires_grad_(True)
    rc, rb = tv_fastrcnn_loss(f2[:, :n_cls], f2[:, n_cls:5 * n_cls], labels, targets)
    (2.0 * rc + 0.5 * rb).backward()
    assert abs(lc.item() - rc.item()) < 2e-6 * max(1.0, abs(rc.item())) and abs(lb.item() - rb.item()) < 2e-6 * max(1.0, abs(rb.item()))
    assert _nerr(fused.grad, f2.grad) < 1e-5
    oc, ob = ro.fastrcnn_loss(fused.detach().cpu()[:, :n_cls], fused.detach().cpu()[:, n_cls:5 * n_cls],
                              [t.cpu() for t in labels], [t.cpu() for t in targets])
    assert abs(lc.item() - oc.item()) < 2e-6 * max(1.0, abs(oc.item())) and abs(lb.item() - ob.item()) < 2e-6 * max(1.0, abs(ob.item()))


def _pair(precision, seed=21):
    from sfvos_b200 import FastRCNNPredictor, TwoMLPHead
    torch.manual_seed(seed)
    head_ref, pred_ref = TVHead(256 * 7 * 7, 1024).cuda(), TVPredictor(1024, 2).cuda()
    head, pred = TwoMLPHead(256 * 7 * 7, 1024).cuda(), FastRCNNPredictor(1024, 2).cuda()
    head.load_state_dict(head_ref.state_dict())
    pred.load_state_dict(pred_ref.state_dict())
    assert list(head.state_dict().keys()) == list(head_ref.state_dict().keys()) == ["fc6.weight", "fc6.bias", "fc7.weight", "fc7.bias"]
    assert list(pred.state_dict().keys()) == list(pred_ref.state_dict().keys())
    head.precision = pred.precision = precision
    return head_ref, pred_ref, head, pred


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_box_branch_forward_backward_matches_torchvision(precision):
    from sfvos_b200 import fastrcnn_loss
    head_ref, pred_ref, head, pred = _pair(precision)
    g = torch.Generator().manual_seed(4)
    M = 333                                                       # ragged: not a multiple of the 128-row tile
    x = torch.randn(M, 256, 7, 7, generator=g).cuda()
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    labels = [(torch.rand(M, generator=g) < 0.25).long().cuda()]
    targets = [(0.3 * torch.randn(M, 4, generator=g)).cuda()]
    zr, rr = pred_ref(head_ref(xr))
    sum(tv_fastrcnn_loss(zr, rr, labels, targets)).backward()
    zo, ro_ = pred(head(xo))
    lc, lb = fastrcnn_loss(zo, ro_, labels, targets)
    (lc + lb).backward()
    ftol = 1e-4 if precision == "fp32" else 1e-2
    assert zo.shape == zr.shape and ro_.shape == rr.shape and zo.dtype == torch.float32
    assert _nerr(zo, zr) < ftol and _nerr(ro_, rr) < ftol
    rc, rb = tv_fastrcnn_loss(zr, rr, labels, targets)
    assert abs(lc.item() - rc.item()) < ftol * max(1.0, rc.item()) and abs(lb.item() - rb.item()) < ftol * max(1.0, rb.item())
    gtol = 1e-3 if precision == "fp32" else 0.1                   # relative L2 (single ReLU flips move individual entries)
    for (n1, p1), (n2, p2) in zip(list(head.named_parameters()) + list(pred.named_parameters()),
                                  list(head_ref.named_parameters()) + list(pred_ref.named_parameters())):
        assert n1 == n2 and p1.grad.shape == p2.grad.shape
        assert _rel(p1.grad, p2.grad) < gtol, (n1, _rel(p1.grad, p2.grad))
        if n1.startswith(("cls_score", "bbox_pred")):             # downstream of every ReLU: max-normalised too
            assert _nerr(p1.grad, p2.grad) < (1e-4 if precision == "fp32" else 2e-2), (n1, _nerr(p1.grad, p2.grad))
    assert _rel(xo.grad, xr.grad) < gtol


def test_box_branch_matches_golden_fixture():
    """fp32 validation mode against tests/golden/box_head.npz (live torchvision CPU outputs, make_box_golden.py)."""
    from sfvos_b200 import fastrcnn_loss
    gold = np.load(os.path.join(GOLDEN, "box_head.npz"))
    _, _, head, pred = _pair("fp32", seed=21)                    # same seed and construction order as the generator
    x = torch.from_numpy(gold["x"]).cuda().requires_grad_(True)
    labels = [torch.from_numpy(gold["labels0"]).cuda(), torch.from_numpy(gold["labels1"]).cuda()]
    targets = [torch.from_numpy(gold["targets0"]).cuda(), torch.from_numpy(gold["targets1"]).cuda()]
    feat = head(x)
    z, r = pred(feat)
    lc, lb = fastrcnn_loss(z, r, labels, targets)
    (lc + lb).backward()
    assert _nerr(feat, torch.from_numpy(gold["feat"])) < 1e-4
    assert _nerr(z, torch.from_numpy(gold["scores"])) < 1e-4 and _nerr(r, torch.from_numpy(gold["deltas"])) < 1e-4
    assert abs(lc.item() - float(gold["loss_cls"])) < 1e-5 and abs(lb.item() - float(gold["loss_box"])) < 1e-5
    assert _nerr(x.grad, torch.from_numpy(gold["gx"])) < 1e-4
    for mod, pre in ((head, "box_head."), (pred, "box_predictor.")):
        for n, p in mod.named_parameters():
            ref64 = torch.from_numpy(gold["g64_" + pre + n])
            got64 = p.grad.reshape(-1)[:: max(1, p.numel() // 64)][:64].cpu()
            assert (got64 - ref64).abs().max().item() <= 1e-4 * max(ref64.abs().max().item(), 1e-6), n
            assert abs(p.grad.double().sum().item() - float(gold["gsum_" + pre + n])) <= 1e-4 * float(gold["gabs_" + pre + n]) + 1e-9, n


def test_box_branch_empty_batch():
    _, _, head, pred = _pair("bf16")
    x = torch.zeros(0, 256, 7, 7).cuda().requires_grad_(True)
    z, r = pred(head(x))
    assert z.shape == (0, 2) and r.shape == (0, 8)
    (z.sum() + r.sum()).backward()
    assert head.fc6.weight.grad.abs().sum().item() == 0.0
