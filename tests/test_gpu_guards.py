"""Out-of-bounds checks of our own (compute-sanitizer is closed on this GPU pool: profiles/sanitizer_r2.txt).

Every activation buffer the engine allocates through ``ops.Act.empty`` is placed in the middle of a larger allocation whose
head and tail GUARD regions are filled with NaN.  After a full forward + backward
  * every guard must still be all-NaN                      -> no kernel wrote outside its buffer (stores, TMA stores, atomics);
  * the results must equal those of the unguarded run      -> no kernel read outside its buffer (a NaN would have spread:
    the neighbours of a buffer are NaN here, zeros or stale data otherwise; TMA out-of-bounds fills come from the tensor-map
    extents, which this pins too).
Level sizes are odd on purpose (13x21, 7x11, 5x3): partial tiles on both spatial axes in every kernel."""
from collections import OrderedDict

import pytest
import torch

from oracle import slowfast_oracle as so

pytestmark = pytest.mark.gpu
GUARD = 4096          # elements on each side (16 KB of f32): a multiple of every vector / TMA alignment in use


class _Guarded:
    def __init__(self, monkeypatch):
        from sfvos_b200 import ops
        self.ops, self.allocs = ops, []

        def empty(B, T, H, W, C, dtype, device, cstride=None):
            cs = cstride if cstride is not None else C
            n = B * T * H * W * cs
            big = torch.full((n + 2 * GUARD,), float("nan"), dtype=dtype, device=device)
            self.allocs.append((big, n))
            return ops.Act(big[GUARD:GUARD + n], B, T, H, W, C, cs, 0)
        monkeypatch.setattr(ops.Act, "empty", staticmethod(empty))

    def check(self):
        assert self.allocs, "no guarded allocation happened"
        torch.cuda.synchronize()
        for big, n in self.allocs:
            assert torch.isnan(big[:GUARD]).all() and torch.isnan(big[GUARD + n:]).all(), (tuple(big.shape), n)
        return len(self.allocs)


def _run_slowfast(precision, sp, fp, levels):
    from sfvos_b200 import SlowFastLayers
    torch.manual_seed(63)
    m = SlowFastLayers(256, torch.device("cuda"), sp, fp).cuda().train()
    m.precision = precision
    fast = [so.synthetic_clip(levels, fp, seed=1234 + 100 * c) for c in range(2)]
    fast_c = [OrderedDict((k, v.cuda()) for k, v in f.items()) for f in fast]
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    out = m.temporally_enhance_features(slow_c, fast_c)
    so.module_loss(out).backward()
    torch.cuda.synchronize()
    return [v.detach().clone() for v in out.values()], [p.grad.detach().clone() for p in m.parameters()]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sp,fp", [(1, 8), (3, 7)])
def test_slowfast_kernels_stay_inside_their_buffers(precision, sp, fp, monkeypatch):
    levels = OrderedDict([("0", (13, 21)), ("1", (7, 11)), ("pool", (5, 3))])
    plain_out, plain_grads = _run_slowfast(precision, sp, fp, levels)
    guard = _Guarded(monkeypatch)
    out, grads = _run_slowfast(precision, sp, fp, levels)
    assert guard.check() >= 30
    for a, b in zip(out + grads, plain_out + plain_grads):
        assert torch.isfinite(a).all()
        if precision == "fp32":
            assert torch.equal(a, b)                                   # the validation mode is bit-reproducible
        else:                                                          # bf16: atomically ordered statistics (last-bit noise)
            assert (a - b).norm().item() <= 0.1 * b.norm().item() + 1e-9


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_roi_and_mask_kernels_stay_inside_their_buffers(precision, monkeypatch):
    from oracle import roi_oracle as ro
    from sfvos_b200 import MaskRCNNHeads, MaskRCNNPredictor, MultiScaleRoIAlign, TwoMLPHead, FastRCNNPredictor

    def run():
        torch.manual_seed(5)
        g = torch.Generator().manual_seed(2)
        shapes = ((45, 83), (23, 41), (12, 21), (6, 11))
        feats = OrderedDict((str(i), torch.randn(2, 256, h, w, generator=g).cuda().requires_grad_(True)) for i, (h, w) in enumerate(shapes))
        boxes = [b.cuda() for b in ro.synthetic_rois(2, 37, image_hw=(180, 330), seed=4321, lo=4.0, hi=175.0)]
        pool_m = MultiScaleRoIAlign(["0", "1", "2", "3"], 14, 2, out_layout="nhwc", precision=precision)
        pool_b = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2, out_layout="nchw", precision=precision, out_dtype="act")
        head, pred = MaskRCNNHeads(256, (256, 256, 256, 256), 1).cuda(), MaskRCNNPredictor(256, 256, 2).cuda()
        bh, bp = TwoMLPHead(256 * 49, 1024).cuda(), FastRCNNPredictor(1024, 2).cuda()
        for mod in (head, pred, bh, bp):
            mod.precision = precision
        logits = pred(head(pool_m(feats, [b[:11] for b in boxes], [(180, 330)] * 2)))
        scores, deltas = bp(bh(pool_b(feats, boxes, [(180, 330)] * 2)))
        (logits.float().square().mean() + scores.float().square().mean() + deltas.float().square().mean()).backward()
        torch.cuda.synchronize()
        return [logits.detach().clone(), scores.detach().clone()] + [f.grad.detach().clone() for f in feats.values()]
    plain = run()
    guard = _Guarded(monkeypatch)
    got = run()
    assert guard.check() >= 10
    for a, b in zip(got, plain):
        assert torch.isfinite(a).all()
        assert (a.float() - b.float()).norm().item() <= (1e-5 if precision == "fp32" else 0.1) * b.float().norm().item() + 1e-9
