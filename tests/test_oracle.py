"""CPU tests: the oracle against the committed golden fixtures (outputs of the unmodified reference,
tests/golden/make_golden.py), against the reference's published parameter counts, and against the live
torchvision ops it restates."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import slowfast_oracle as so
from oracle import roi_oracle as ro

CONFIGS = [(1, 8), (3, 7), (2, 16), (4, 32), (1, 1)]
LEVELS = OrderedDict([("0", (8, 12)), ("pool", (4, 6))])


def _inputs(sp, fp, seed0=1234):
    fast, slow = [], []
    for clip in range(2):
        f = so.synthetic_clip(LEVELS, fp, seed=seed0 + 100 * clip, zero_left=(fp // 2 if clip == 1 else 0))
        fast.append(f)
        slow.append(so.slice_window(f, fp // 2, sp))
    return slow, fast


def _sample_idx(numel, k=64):
    g = torch.Generator().manual_seed(numel % 9973 + 17)
    return torch.randint(0, numel, (k,), generator=g)


# published: final_report/chapters/Experiments.tex:20-24 totals minus torchvision Mask R-CNN (43,922,395)
@pytest.mark.parametrize("sp,fp,total", [(1, 1, 45421851), (3, 3, 46398747), (7, 7, 48407835),
                                         (1, 7, 45618459), (3, 7, 46570779)])
def test_param_counts_match_report(sp, fp, total):
    assert so.param_count(sp, fp) == total - 43922395


@pytest.mark.parametrize("sp,fp,ks,kf,kl", [
    (1, 8, (1, 1, 1), (3, 3, 4), (6, 4)), (4, 32, (2, 2, 2), (11, 11, 12), (20, 11)),
    (2, 16, (1, 1, 2), (6, 6, 6), (10, 5)), (3, 7, (1, 2, 2), (3, 3, 3), (3, 2)), (8, 8, (3, 3, 4), (3, 3, 4), (1, 1))])
def test_kernel_size_rules(sp, fp, ks, kf, kl):
    assert so.calc_kernel_sizes(sp) == ks and so.calc_kernel_sizes(fp) == kf
    spec = so.layer_specs(sp, fp)
    assert (spec["conv_f2s1"][2], spec["conv_f2s2"][2]) == kl
    # every pathway ends at temporal extent 1
    assert sp - sum(ks) + 3 == 1 and fp - sum(kf) + 3 == 1


@pytest.mark.parametrize("sp,fp", CONFIGS)
def test_slowfast_oracle_matches_reference_golden(sp, fp):
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    sd = so.init_state_dict(sp, fp, seed=63)
    assert so.param_count(sp, fp) == int(gold["n_params"])
    slow, fast = _inputs(sp, fp, int(gold["input_seed"]))
    # the fixture's inputs were chosen for their ReLU margin (make_golden.py): no pre-activation within 1e-5 of zero
    margin, n_relu = so.relu_margin(sd, slow, fast)
    assert margin >= 1e-5 and abs(margin - float(gold["relu_margin"])) <= 1e-9 and n_relu == int(gold["n_relu_inputs"])
    merged, loss, grads, buffers = so.grads_of(sd, slow, fast)
    for k, v in merged.items():
        ref = torch.from_numpy(gold["train_out_" + k])
        assert v.shape == ref.shape
        assert (v.detach() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert abs(loss.item() - float(gold["loss"])) < 1e-5
    for name, g in grads.items():
        flat = g.flatten()
        smp = torch.from_numpy(gold["grad_smp_" + name])
        scale = float(gold["grad_abs_" + name]) / flat.numel() + 1e-12
        assert (flat[_sample_idx(flat.numel())] - smp).abs().max().item() <= 5e-3 * scale + 1e-9, name
        assert abs(flat.double().abs().sum().item() - float(gold["grad_abs_" + name])) <= 1e-3 * float(gold["grad_abs_" + name]) + 1e-7, name
    for name, b in buffers.items():
        ref = torch.from_numpy(gold["buf_" + name])
        assert torch.allclose(b.to(ref.dtype), ref, rtol=1e-5, atol=1e-6), name
    # with that margin the fp32 run has the exact masks, so fp32 and fp64 gradients agree to rounding
    # (measured <= 3e-6 max-normalised; with a flipped mask it is ~1e-3)
    masks32 = so.relu_preacts(sd, slow, fast, dtype=torch.float32)
    masks64 = so.relu_preacts(sd, slow, fast, dtype=torch.float64)
    assert all(torch.equal(a > 0, b > 0) for (_, a), (_, b) in zip(masks32, masks64))
    _, _, grads64, _ = so.grads_of(sd, slow, fast, dtype=torch.float64)
    for name, g in grads.items():
        if name.endswith(("conv1.bias", "conv2.bias", "conv3.bias")):
            continue                      # exactly zero through train-mode BN (rounding residue only)
        err = (g.double() - grads64[name]).abs().max().item() / grads64[name].abs().max().item()
        assert err <= 2e-5, (name, err)
    # eval forward with the updated running stats
    for k, v in buffers.items():
        sd[k] = v
    out_eval = so.temporally_enhance_features(sd, slow, fast, False)
    for k, v in out_eval.items():
        ref = torch.from_numpy(gold["eval_out_" + k])
        assert (v - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_conv3d_c_restatement_matches_functional():
    import ctypes
    from oracle import lib
    torch.manual_seed(0)
    B, Cin, T, H, W, Cout, kt = 2, 5, 4, 6, 7, 3, 2
    x = torch.randn(B, Cin, T, H, W)
    w = torch.randn(Cout, Cin, kt, 3, 3)
    b = torch.randn(Cout)
    y = np.zeros((B, Cout, T - kt + 1, H, W), dtype=np.float32)
    P = ctypes.POINTER(ctypes.c_float)
    i64 = ctypes.c_int64
    lib().conv3d_ncdhw_f32(x.numpy().ctypes.data_as(P), w.numpy().ctypes.data_as(P), b.numpy().ctypes.data_as(P),
                           y.ctypes.data_as(P), i64(B), i64(Cin), i64(T), i64(H), i64(W), i64(Cout), i64(kt), i64(3), i64(3), i64(1))
    ref = torch.nn.functional.conv3d(x, w, b, padding=(0, 1, 1))
    assert np.abs(y - ref.numpy()).max() < 1e-4
    # batch-norm (train) + relu
    g, be = torch.rand(Cout) + 0.5, torch.randn(Cout)
    rm, rv = torch.zeros(Cout), torch.ones(Cout)
    rm_c, rv_c = rm.numpy().copy(), rv.numpy().copy()
    y2 = y.copy()
    lib().batchnorm3d_ncdhw_f32(y2.ctypes.data_as(P), g.numpy().ctypes.data_as(P), be.numpy().ctypes.data_as(P),
                                rm_c.ctypes.data_as(P), rv_c.ctypes.data_as(P), i64(B), i64(Cout),
                                i64((T - kt + 1) * H * W), ctypes.c_int(1), ctypes.c_int(1),
                                ctypes.c_double(0.1), ctypes.c_double(1e-5))
    ref2 = torch.relu(torch.nn.functional.batch_norm(ref, rm, rv, g, be, True, 0.1, 1e-5))
    assert np.abs(y2 - ref2.numpy()).max() < 1e-4
    assert np.abs(rm_c - rm.numpy()).max() < 1e-5 and np.abs(rv_c - rv.numpy()).max() < 1e-5


def _edge_rois():
    return torch.tensor([[0, -5.0, -3.0, 20.0, 18.0], [0, 30.0, 20.0, 80.0, 60.0], [1, 10.0, 10.0, 10.2, 10.1],
                         [1, 0.0, 0.0, 84.0, 48.0], [0, 200.0, 200.0, 260.0, 240.0], [1, 60.0, 30.0, 100.0, 70.0]])


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("p,sr,scale", [(7, 2, 0.25), (14, 2, 0.5), (28, -1, 1.0)])
def test_roi_align_c_matches_torchvision(dtype, tol, p, sr, scale):
    import torchvision
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 48, 84, generator=g, dtype=dtype)
    rois = _edge_rois().to(dtype)
    if scale == 1.0:
        rois[:, 1:] *= 0.5
    ref = torchvision.ops.roi_align(x, rois, (p, p), scale, sr, False)
    out = ro.roi_align(x, rois, p, scale, sr)
    assert (out - ref).abs().max().item() <= tol
    gout = torch.randn(ref.shape, generator=g, dtype=dtype)
    ref_b = torch.ops.torchvision._roi_align_backward(gout, rois, scale, p, p, 2, 3, 48, 84, sr, False)
    out_b = ro.roi_align_backward(gout, rois, (2, 3, 48, 84), scale, sr)
    assert (out_b - ref_b).abs().max().item() <= tol * 10


def test_level_mapper_matches_torchvision_and_known_answers():
    from torchvision.ops.poolers import LevelMapper
    sides = [10, 111, 112, 223, 224, 447, 448, 896]
    boxes = torch.tensor([[0.0, 0.0, float(s), float(s)] for s in sides])
    assert ro.level_mapper(boxes).tolist() == [0, 0, 1, 1, 2, 2, 3, 3]       # SURVEY 8(c) spot values
    rnd = torch.cat(ro.synthetic_rois(4, 4000, seed=99))
    assert torch.equal(ro.level_mapper(rnd), LevelMapper(2, 5)([rnd]))


def test_multiscale_and_mask_head_match_golden():
    gold = np.load(os.path.join(GOLDEN, "roi_mask.npz"))
    feats = [torch.from_numpy(gold["feat" + str(i)]) for i in range(4)]
    boxes = [torch.from_numpy(gold["boxes0"]), torch.from_numpy(gold["boxes1"])]
    for p in (7, 14):
        out, levels = ro.multiscale_roi_align(feats, boxes, [(187, 333)] * 2, p, 2)
        assert torch.equal(levels, torch.from_numpy(gold["levels"]))
        assert (out - torch.from_numpy(gold[f"pool{p}"])).abs().max().item() < 2e-5
    from torchvision.models.detection.mask_rcnn import MaskRCNNHeads, MaskRCNNPredictor
    torch.manual_seed(11)
    head = MaskRCNNHeads(256, (256, 256, 256, 256), 1)
    pred = MaskRCNNPredictor(256, 256, 2)
    sd = {"mask_head." + k: v for k, v in head.state_dict().items()}
    sd.update({"mask_predictor." + k: v for k, v in pred.state_dict().items()})
    x = torch.from_numpy(gold["mh_x"])
    logits = ro.mask_predictor_forward(sd, ro.mask_head_forward(sd, x))
    ref = torch.from_numpy(gold["mh_logits"])
    assert (logits - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    gt = torch.zeros(2, 187, 333, dtype=torch.uint8)
    gt[0, 40:120, 60:200] = 1
    gt[1, 100:180, 150:300] = 1
    props = [torch.tensor([[50.0, 30.0, 210.0, 130.0], [140.0, 90.0, 310.0, 186.0], [0.0, 0.0, 20.5, 17.25]])]
    loss = ro.maskrcnn_loss(logits, props, [gt], [torch.tensor([1, 1])], [torch.tensor([0, 1, 0])])
    assert abs(loss.item() - float(gold["mh_loss"])) <= 1e-4 * abs(float(gold["mh_loss"]))


def _box_golden_modules():
    """The torchvision modules of tests/golden/make_box_golden.py, rebuilt from the seed (the 13 M weights are not stored)."""
    from torchvision.models.detection.faster_rcnn import FastRCNNPredictor, TwoMLPHead
    torch.manual_seed(21)
    return TwoMLPHead(256 * 7 * 7, 1024), FastRCNNPredictor(1024, 2)


def test_box_head_and_fastrcnn_loss_match_golden_and_torchvision():
    gold = np.load(os.path.join(GOLDEN, "box_head.npz"))
    head, pred = _box_golden_modules()
    sd = {"box_head." + k: v for k, v in head.state_dict().items()}
    sd.update({"box_predictor." + k: v for k, v in pred.state_dict().items()})
    x = torch.from_numpy(gold["x"]).requires_grad_(True)
    labels = [torch.from_numpy(gold["labels0"]), torch.from_numpy(gold["labels1"])]
    targets = [torch.from_numpy(gold["targets0"]), torch.from_numpy(gold["targets1"])]
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    feat = ro.box_head_forward(leaves, x)
    scores, deltas = ro.box_predictor_forward(leaves, feat)
    l_cls, l_box = ro.fastrcnn_loss(scores, deltas, labels, targets)
    for got, key in ((feat, "feat"), (scores, "scores"), (deltas, "deltas")):
        ref = torch.from_numpy(gold[key])
        assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), key
    assert abs(l_cls.item() - float(gold["loss_cls"])) <= 1e-5 and abs(l_box.item() - float(gold["loss_box"])) <= 1e-5
    (l_cls + l_box).backward()
    gx = torch.from_numpy(gold["gx"])
    assert (x.grad - gx).abs().max().item() <= 1e-4 * gx.abs().max().item()
    for k, v in leaves.items():
        assert abs(v.grad.double().sum().item() - float(gold["gsum_" + k])) <= 1e-4 * float(gold["gabs_" + k]) + 1e-9, k
        got64 = v.grad.reshape(-1)[:: max(1, v.numel() // 64)][:64]
        ref64 = torch.from_numpy(gold["g64_" + k])
        assert (got64 - ref64).abs().max().item() <= 1e-4 * max(ref64.abs().max().item(), 1e-6), k
    # the written-out loss against the live torchvision function on a larger random case (incl. |d| on both sides of beta)
    from torchvision.models.detection.roi_heads import fastrcnn_loss as tv_loss
    g = torch.Generator().manual_seed(5)
    z, r = torch.randn(300, 3, generator=g) * 2, torch.randn(300, 12, generator=g) * 0.2
    lab = [torch.randint(0, 3, (300,), generator=g)]
    tgt = [torch.randn(300, 4, generator=g) * 0.2]
    a, b = ro.fastrcnn_loss(z, r, lab, tgt)
    ra, rb = tv_loss(z, r, lab, tgt)
    assert abs(a.item() - ra.item()) <= 1e-6 and abs(b.item() - rb.item()) <= 1e-6


@pytest.mark.parametrize("sp,fp", [(1, 8), (3, 7), (2, 16), (1, 1)])
def test_eval_mode_is_a_shift_invariant_temporal_filter(sp, fp):
    """The identity behind SlowFastLayers.temporally_enhance_sequence, on the reference restatement itself: in eval mode
    the zero-padded sequence fed as ONE clip yields at temporal index t the output of the reference's window around t."""
    from math import ceil, floor
    levels = OrderedDict([("0", (6, 8))])
    sd = so.init_state_dict(sp, fp, seed=63)
    g = torch.Generator().manual_seed(2)
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    n = 6
    frames = torch.randn(n, 256, 6, 8, generator=g)
    lo, hi = fp // 2, fp - fp // 2 - 1
    padded = torch.cat([torch.zeros(lo, 256, 6, 8), frames, torch.zeros(hi, 256, 6, 8)])
    s_off = fp // 2 - sp // 2
    fast = padded.unsqueeze(0).transpose(1, 2)                                   # [1,256,n+fp-1,H,W]
    slow = padded[s_off:s_off + n + sp - 1].unsqueeze(0).transpose(1, 2)
    s, f = so.forward(sd, slow, fast, False)
    assert s.shape[2] == n and f.shape[2] == n
    sweep = torch.cat([s, f], dim=1)[0].transpose(0, 1)                          # [n,256,H,W]
    for t in range(n):
        win = torch.stack([frames[i] if 0 <= i < n else torch.zeros(256, 6, 8) for i in range(t - floor(fp / 2), t + ceil(fp / 2))])
        p = fp // 2
        ref = so.temporally_enhance_features(sd, [OrderedDict([("0", win[p - floor(sp / 2):p + ceil(sp / 2)])])],
                                             [OrderedDict([("0", win)])], False)["0"][0]
        assert (sweep[t] - ref).abs().max().item() <= 2e-5 * ref.abs().max().item(), t


def _paste_cases():
    g = torch.Generator().manual_seed(9)
    masks = torch.rand(9, 1, 28, 28, generator=g)
    boxes = torch.tensor([[10.3, 12.7, 90.2, 70.9], [-15.0, -8.0, 30.5, 25.0], [100.0, 60.0, 170.0, 125.0],
                          [40.0, 30.0, 40.5, 30.2], [0.0, 0.0, 160.0, 120.0], [150.2, 100.1, 200.0, 140.0],
                          [155.0, 110.0, 190.0, 150.0], [20.0, 20.0, 48.0, 48.0], [33.3, 44.4, 77.7, 99.9]])
    # (a box entirely outside the image makes torchvision's slice assignment raise; detections are clipped to the image,
    # so the reference never meets one -- the kernel writes zeros there)
    return masks, boxes, (120, 160)


def test_paste_masks_restatement_matches_torchvision():
    from torchvision.models.detection.roi_heads import paste_masks_in_image as tv_paste
    masks, boxes, shape = _paste_cases()
    ref = tv_paste(masks, boxes, shape)
    got = ro.paste_masks_in_image(masks, boxes, shape)
    assert got.shape == ref.shape == (9, 1, 120, 160)
    assert (got - ref).abs().max().item() <= 2e-6
    assert ((got != 0) == (ref != 0)).all()                   # identical footprint: same integer boxes, same clipping
    assert ro.paste_masks_in_image(masks[:0], boxes[:0], shape).shape == (0, 1, 120, 160)
