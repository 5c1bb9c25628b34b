"""Eval-mode sequence sweep: SlowFastLayers.temporally_enhance_sequence must return, for every frame t
of a sequence, what temporally_enhance_features returns for the reference's window around t (code/helpers/model.py:215-248,
322-340: fast = [t - fp//2, t + ceil(fp/2)), slow = its centre sp frames, out-of-sequence frames all-zero) -- and the CPU
oracle's window output.  Also the SegmentationModel eval path built on it against the reference-faithful per-frame loop."""
import os
import sys
from collections import OrderedDict
from math import ceil, floor

import pytest
import torch

from conftest import ROOT
from oracle import slowfast_oracle as so

pytestmark = pytest.mark.gpu

LEVELS = OrderedDict([("0", (24, 40)), ("pool", (6, 11))])


def _nerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def _module(sp, fp, precision):
    from sfvos_b200 import SlowFastLayers
    torch.manual_seed(63)
    mod = SlowFastLayers(256, torch.device("cuda"), sp, fp).cuda()
    mod.precision = precision
    g = torch.Generator().manual_seed(5)
    for name, buf in mod.named_buffers():            # non-trivial running statistics (as after training)
        if name.endswith("running_mean"):
            buf.copy_(0.2 * torch.randn(buf.shape, generator=g).cuda())
        elif name.endswith("running_var"):
            buf.copy_((0.5 + torch.rand(buf.shape, generator=g)).cuda())
    return mod.eval()


def _frames(n, seed=11):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((k, torch.randn(n, 256, h, w, generator=g).cuda()) for k, (h, w) in LEVELS.items())


def _window(frames, t, sp, fp):
    """The reference's window around frame t: zero frames outside the sequence, slow = centre rows (model.py:215-248)."""
    n = next(iter(frames.values())).shape[0]
    idx = range(t - floor(fp / 2), t + ceil(fp / 2))
    fast = OrderedDict()
    for k, v in frames.items():
        fast[k] = torch.stack([v[i] if 0 <= i < n else torch.zeros_like(v[0]) for i in idx])
    p = fp // 2
    slow = OrderedDict((k, v[p - floor(sp / 2):p + ceil(sp / 2)]) for k, v in fast.items())
    return slow, fast


@pytest.mark.parametrize("sp,fp", [(1, 8), (3, 7), (2, 16), (4, 32), (1, 1)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sequence_sweep_equals_per_window_eval(sp, fp, precision):
    mod = _module(sp, fp, precision)
    n = 11
    frames = _frames(n)
    seq = mod.temporally_enhance_sequence(frames)
    chunked = mod.temporally_enhance_sequence(frames, max_frames=4)
    tol = 2e-5 if precision == "fp32" else 1e-2
    for k in LEVELS:
        assert seq[k].shape == (n, 256) + LEVELS[k] and seq[k].dtype == torch.float32
        assert _nerr(chunked[k], seq[k]) <= tol, k
    sd = {k: v.detach().cpu() for k, v in mod.state_dict().items()}
    for t in (0, 1, n // 2, n - 2, n - 1):
        slow, fast = _window(frames, t, sp, fp)
        with torch.no_grad():
            win = mod.temporally_enhance_features([slow], [fast])
        ref = so.temporally_enhance_features(sd, [OrderedDict((k, v.cpu()) for k, v in slow.items())],
                                             [OrderedDict((k, v.cpu()) for k, v in fast.items())], False)
        for k in LEVELS:
            assert _nerr(seq[k][t:t + 1], win[k]) <= tol, (k, t, _nerr(seq[k][t:t + 1], win[k]))
            assert _nerr(seq[k][t:t + 1], ref[k]) <= (1e-4 if precision == "fp32" else 1e-2), (k, t)


def test_sequence_sweep_halo_and_train_mode_guard():
    sp, fp = 1, 8
    mod = _module(sp, fp, "bf16")
    frames = _frames(14, seed=3)
    full = mod.temporally_enhance_sequence(frames)
    lo, hi = fp // 2, fp - fp // 2 - 1
    c0, c1 = 5, 9                                    # stream frames 5..8 with their real neighbours as halo
    part = OrderedDict((k, v[c0 - lo:c1 + hi]) for k, v in frames.items())
    piece = mod.temporally_enhance_sequence(part, halo=(lo, hi))
    for k in LEVELS:
        assert piece[k].shape[0] == c1 - c0
        assert _nerr(piece[k], full[k][c0:c1]) <= 1e-2
    mod.train()
    with pytest.raises(RuntimeError):
        mod.temporally_enhance_sequence(frames)


def test_segmentation_model_sequence_mode_equals_per_frame_loop():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    from helpers.model import SegmentationModel
    from test_gpu_model import _sequence
    dev = torch.device("cuda")
    torch.manual_seed(63)
    model = SegmentationModel(device=dev, slow_pathway_size=1, fast_pathway_size=4, maskrcnn_weights=None, pretrained=False)
    model.to(dev).eval()
    model.slow_fast.precision = "fp32"
    for m in (model.maskrcnn_model.roi_heads.box_roi_pool, model.maskrcnn_model.roi_heads.mask_roi_pool,
              model.maskrcnn_model.roi_heads.mask_head, model.maskrcnn_model.roi_heads.mask_predictor,
              model.maskrcnn_model.roi_heads.box_head, model.maskrcnn_model.roi_heads.box_predictor):
        m.precision = "fp32"
    # the libsfvos FPN / RPN head in validation mode too, and one backbone call per frame in BOTH modes: the torchvision body's
    # cuDNN kernels may differ in the last bits between batch sizes, which random-init detection scores (all ~equal) turn
    # into a different top-k / NMS selection
    model.maskrcnn_model.backbone.fpn.precision = model.maskrcnn_model.rpn.head.precision = "fp32"
    model.backbone_batch = 1
    model.maskrcnn_model.roi_heads.score_thresh = 0.0
    imgs, targets = _sequence(n=5)
    targets[2] = {}                                  # a frame without objects is skipped (model.py:289-296) but still feeds its neighbours' windows
    model.sequence_mode, model.sequence_chunk = True, 2
    with torch.no_grad():
        _, seq = model(imgs, targets)
    model.sequence_mode = False
    with torch.no_grad():
        _, loop = model(imgs, targets)
    assert len(seq) == len(loop) == 5 and seq[2] == {} and loop[2] == {}
    for a, b in zip(seq, loop):
        assert set(a.keys()) == set(b.keys())
        if not a:
            continue
        assert a["boxes"].shape == b["boxes"].shape and torch.equal(a["labels"], b["labels"])
        assert _nerr(a["boxes"], b["boxes"]) < 1e-3 and _nerr(a["scores"], b["scores"]) < 1e-3
        assert (a["masks"] - b["masks"]).abs().mean().item() < 1e-3
