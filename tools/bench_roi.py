"""Time the multi-level ROIAlign launches of the bench workload (box 7x7 f32 NCHW, mask 14x14 bf16 NHWC), fwd and bwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops, workload as wl
from sfvos_b200.roi_heads import MultiScaleRoIAlign

dev = "cuda"
B = 8
feats = {k: torch.randn(B, h, w, 256, device=dev).permute(0, 3, 1, 2).requires_grad_(True) for k, (h, w) in wl.LEVELS.items()}
box = [b.to(dev) for b in wl.synthetic_rois(B, 512)]
mask = [b[:128] for b in box]
shapes = [wl.IMAGE_HW] * B
pools = {"box p7 f32 nchw": (MultiScaleRoIAlign(wl.POOL_LEVELS, 7, 2, out_layout="nchw"), box, (7, 4, 4)),
         "mask p14 bf16 nhwc": (MultiScaleRoIAlign(wl.POOL_LEVELS, 14, 2, out_layout="nhwc"), mask, (14, 4, 2))}
for name, (pool, rois, (P, ei, eo)) in pools.items():
    ops.TIMING = []
    for it in range(6):
        out = pool(feats, rois, shapes)
        out.backward(torch.ones_like(out))
        for f in feats.values(): f.grad = None
    torch.cuda.synchronize()
    t = ops.TIMING[-8:]            # last 4 iterations
    ops.TIMING = None
    for tag, bwd in (("fwd", False), ("bwd", True)):
        ms = sorted(a.elapsed_time(b) for n, _, a, b in t if tag in n)[1]
        by = wl.roi_align_bytes(rois, P, ei, eo, backward=bwd)
        print(f"{name:20s} {tag}: {ms*1e3:8.1f} us  {by/1e9:6.2f} GB algorithmic -> {by/ms/1e6:7.1f} GB/s ({by/ms/1e6/6542.1*100:.0f}% of 6542)", flush=True)
