"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Writes tests/golden/slowfast_sp{sp}_fp{fp}.npz and tests/golden/roi_mask.npz.

SlowFast fixtures: reference ``SlowFastLayers`` (code/helpers/model.py:30-165) constructed under
torch.manual_seed(63) (code/helpers/constants.py:11), fed seeded synthetic FPN features, CPU fp32.
Stored: train-mode outputs, BN buffers after that step, eval-mode outputs (with those buffers), the scalar
module loss, and per-parameter gradient summaries (sum, abs-sum, 64 sampled entries).

ReLU margin.  A weight gradient upstream of a ReLU is a discontinuous function of the forward pass: an element whose
pre-activation sits within arithmetic noise of 0 has an undetermined mask, and one flipped mask moves single gradient
entries by ~1e-3 of the maximum (round 1's red GPU test: (sp,fp)=(3,7) had a pre-activation at 2.7e-7).  So the input
seed of every fixture is CHOSEN: the first seed 1234 + 1000*k for which min |pre-activation| over all ReLU layers (fp64
oracle) is >= RELU_MARGIN = 1e-5 -- ~10x the fp32 summation noise measured on this graph (<= 6e-6 max, ~5e-7 typical) and
~50x the noise of the GPU validation mode (fp64 accumulation, f32 storage: ~2e-7).  The script also asserts that the
REFERENCE's own fp32 run has exactly the fp64 masks, and stores the seed and the margin in the fixture; the tests
re-check the margin on the oracle before they compare gradients.
ROI/mask fixtures: outputs of the live torchvision ops/modules the reference calls at model.py:346.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/code")

from helpers.model import SlowFastLayers  # noqa: E402  (the reference itself)
from oracle import slowfast_oracle as so  # noqa: E402
from oracle import roi_oracle as ro  # noqa: E402

LEVELS = OrderedDict([("0", (8, 12)), ("pool", (4, 6))])
CONFIGS = [(1, 8), (3, 7), (2, 16), (4, 32), (1, 1)]
N_CLIPS = 2
SAMPLES = 64
RELU_MARGIN = 1e-5


def sample_idx(numel, k=SAMPLES):
    g = torch.Generator().manual_seed(numel % 9973 + 17)
    return torch.randint(0, numel, (k,), generator=g)


def make_inputs(sp, fp, seed0=1234):
    fast, slow = [], []
    for clip in range(N_CLIPS):
        f = so.synthetic_clip(LEVELS, fp, seed=seed0 + 100 * clip, zero_left=(fp // 2 if clip == 1 else 0))
        fast.append(f)
        slow.append(so.slice_window(f, fp // 2, sp))
    return slow, fast


def pick_seed(sp, fp):
    """First input seed whose ReLU margin (fp64 oracle) is >= RELU_MARGIN."""
    sd = so.init_state_dict(sp, fp, seed=63)
    for k in range(500):
        seed0 = 1234 + 1000 * k
        margin, n = so.relu_margin(sd, *make_inputs(sp, fp, seed0))
        if margin >= RELU_MARGIN:
            return seed0, margin, n
    raise RuntimeError("no seed with the required ReLU margin")


def reference_relu_masks(ref, slow, fast):
    """ReLU masks of the UNMODIFIED reference module's own fp32 forward, via forward hooks on its BatchNorm3d modules (the
    shared nn.ReLU is in-place, so the pre-activation is captured before it).  Layer-3 BNs feed no ReLU."""
    masks, hooks = [], []
    for name in ("bn_s1", "bn_f1", "bn_f2s1", "bn_s2", "bn_f2", "bn_f2s2"):
        hooks.append(getattr(ref, name).register_forward_hook(lambda m, i, o, name=name: masks.append((name, (o.detach() > 0).clone()))))
    out = ref.temporally_enhance_features(slow, fast)
    for h in hooks:
        h.remove()
    return out, masks


def slowfast_fixture(sp, fp):
    torch.manual_seed(63)
    ref = SlowFastLayers(256, torch.device("cpu"), sp, fp)
    # the oracle's init must reproduce the reference's parameters exactly
    sd0 = so.init_state_dict(sp, fp, seed=63)
    ref_sd = ref.state_dict()
    assert list(ref_sd.keys()) == list(sd0.keys()), (list(ref_sd.keys()), list(sd0.keys()))
    for k in ref_sd:
        assert torch.equal(ref_sd[k], sd0[k]), k
    n_params = sum(p.numel() for p in ref.parameters())

    seed0, margin, n_relu = pick_seed(sp, fp)
    slow, fast = make_inputs(sp, fp, seed0)
    ref.train()
    out_train, masks = reference_relu_masks(ref, slow, fast)
    # the reference's fp32 masks must be the exact (fp64) ones, layer by layer in call order
    bn_of = dict(so.LAYER_ORDER)
    exact = so.relu_preacts(sd0, slow, fast)
    assert len(exact) == len(masks)
    for (conv, y64), (bn, m32) in zip(exact, masks):
        assert bn_of[conv] == bn and torch.equal(y64 > 0, m32), f"reference fp32 ReLU mask differs from fp64 at {conv}"
    loss = so.module_loss(out_train)
    loss.backward()
    rec = {"n_params": np.int64(n_params), "loss": np.float64(loss.item()), "input_seed": np.int64(seed0),
           "relu_margin": np.float64(margin), "n_relu_inputs": np.int64(n_relu)}
    for k, v in out_train.items():
        rec["train_out_" + k] = v.detach().numpy()
    for name, p in ref.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        flat = g.flatten()
        rec["grad_sum_" + name] = np.float64(flat.double().sum().item())
        rec["grad_abs_" + name] = np.float64(flat.double().abs().sum().item())
        rec["grad_smp_" + name] = flat[sample_idx(flat.numel())].numpy()
    for name, b in ref.named_buffers():
        rec["buf_" + name] = b.detach().numpy()
    ref.eval()
    with torch.no_grad():
        out_eval = ref.temporally_enhance_features(slow, fast)
    for k, v in out_eval.items():
        rec["eval_out_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"slowfast_sp{sp}_fp{fp}.npz"), **rec)
    print(f"sp={sp} fp={fp}: params={n_params} loss={loss.item():.6f} input_seed={seed0} relu_margin={margin:.3e} over {n_relu} ReLU inputs")


def roi_mask_fixture():
    import torchvision
    from torchvision.ops import MultiScaleRoIAlign
    from torchvision.models.detection.mask_rcnn import MaskRCNNHeads, MaskRCNNPredictor
    from torchvision.models.detection.roi_heads import maskrcnn_loss

    g = torch.Generator().manual_seed(7)
    n = 2
    image_shapes = [(187, 333)] * n               # quarter-size DAVIS frame after the transform
    shapes = [(48, 84), (24, 42), (12, 21), (6, 11)]
    feats = OrderedDict((str(i), torch.randn(n, 16, h, w, generator=g)) for i, (h, w) in enumerate(shapes))
    boxes = ro.synthetic_rois(n, 24, image_hw=image_shapes[0], seed=4321, lo=4.0, hi=175.0)
    # edge cases the reference path can meet: sub-cell box, full image, box on the level boundary (56 px = 224/4)
    boxes[0] = torch.cat([boxes[0], torch.tensor([[10.0, 10.0, 10.4, 10.3], [0.0, 0.0, 333.0, 187.0],
                                                  [5.0, 5.0, 61.0, 61.0], [300.0, 150.0, 333.0, 187.0]])])
    rec = {}
    for p in (7, 14):
        pool = MultiScaleRoIAlign(["0", "1", "2", "3"], p, 2)
        out = pool(feats, boxes, image_shapes)
        rec[f"pool{p}"] = out.numpy()
        if p == 7:
            rec["levels"] = pool.map_levels(boxes).numpy()
            rec["scales"] = np.array(pool.scales)
    for k, v in feats.items():
        rec["feat" + k] = v.numpy()
    rec["boxes0"], rec["boxes1"] = boxes[0].numpy(), boxes[1].numpy()

    torch.manual_seed(11)
    head = MaskRCNNHeads(256, (256, 256, 256, 256), 1)
    pred = MaskRCNNPredictor(256, 256, 2)
    x = torch.randn(3, 256, 14, 14, generator=g)
    x.requires_grad_(True)
    logits = pred(head(x))
    gt_masks = [torch.zeros(2, 187, 333, dtype=torch.uint8)]
    gt_masks[0][0, 40:120, 60:200] = 1
    gt_masks[0][1, 100:180, 150:300] = 1
    props = [torch.tensor([[50.0, 30.0, 210.0, 130.0], [140.0, 90.0, 310.0, 186.0], [0.0, 0.0, 20.5, 17.25]])]
    gt_labels = [torch.tensor([1, 1])]
    matched = [torch.tensor([0, 1, 0])]
    loss = maskrcnn_loss(logits, props, gt_masks, gt_labels, matched)
    loss.backward()
    rec["mh_x"] = x.detach().numpy()
    rec["mh_logits"] = logits.detach().numpy()
    rec["mh_loss"] = np.float64(loss.item())
    rec["mh_dx_smp"] = x.grad.flatten()[sample_idx(x.grad.numel())].numpy()
    for name, p in list(head.named_parameters()) + list(pred.named_parameters()):
        rec["mh_gsum_" + name] = np.float64(p.grad.double().sum().item())
        rec["mh_gabs_" + name] = np.float64(p.grad.double().abs().sum().item())
    rec["tv_version"] = np.array(torchvision.__version__)
    np.savez_compressed(os.path.join(HERE, "roi_mask.npz"), **rec)
    print("roi/mask fixture: loss", loss.item())


if __name__ == "__main__":
    for sp, fp in CONFIGS:
        slowfast_fixture(sp, fp)
    if "--slowfast-only" not in sys.argv:
        roi_mask_fixture()
