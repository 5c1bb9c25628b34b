"""ncu --set full capture of the kernels on the round-2 fix list, one measured launch each after a warm-up
(`ncu --profile-from-start off`: only the launches between cudaProfilerStart/Stop are profiled):
  conv_tstack<32,9> fast_conv2 fprop | wgrad_c32 fast_conv2 | conv_pair slow_conv1 fprop |
  roi_align fwd p7 (4096 ROIs, bf16 rows) | roi_align fwd p14 (1024 ROIs) | roi_align bwd p7 | roi_align bwd p14"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops, workload as wl
from sfvos_b200.roi_heads import MultiScaleRoIAlign

dev = "cuda"
B, H, W = 8, 192, 336


def act(b, t, h, w, c):
    return ops.Act(torch.randn(b * t * h * w * c, device=dev).bfloat16(), b, t, h, w, c)


def conv_case(T, cin, cout, kt):
    To = T - kt + 1
    wt = torch.randn(cout, cin, kt, 3, 3, device=dev) / math.sqrt(cin * kt * 9)
    x = act(B, T, H, W, cin)
    cp = 32 if cin <= 32 else (cin + 63) // 64 * 64
    wp = ops.pack_weights(wt, 0, ops.BF16, cp)
    y = ops.Act.empty(B, To, H, W, cout, torch.float32, dev)
    stats = torch.zeros(2 * cout, device=dev)
    return lambda: ops.conv(x, wp, cp, cout, (kt, 3, 3), (0, 1, 1), To, y, umma=True, stats=stats)


def wgrad_case(T, cin, cout, kt):
    To = T - kt + 1
    x, dy = act(B, T, H, W, cin), act(B, To, H, W, cout)
    dwp = torch.zeros(kt * 9 * cin * cout, device=dev)
    return lambda: ops.wgrad(x, dy, (kt, 3, 3), (0, 1, 1), dwp, umma=True)


cases = [conv_case(6, 32, 32, 3), wgrad_case(6, 32, 32, 3), conv_case(1, 256, 192, 1)]
feats = {k: torch.randn(B, h, w, 256, device=dev).permute(0, 3, 1, 2).requires_grad_(True) for k, (h, w) in wl.LEVELS.items()}
box = [b.to(dev) for b in wl.synthetic_rois(B, 512)]
mask = [b[:128] for b in box]
shapes = [wl.IMAGE_HW] * B
pools = [(MultiScaleRoIAlign(wl.POOL_LEVELS, 7, 2, out_layout="nchw", out_dtype=torch.bfloat16), box),
         (MultiScaleRoIAlign(wl.POOL_LEVELS, 14, 2, out_layout="nhwc"), mask)]


def roi_pass():
    for pool, rois in pools:
        out = pool(feats, rois, shapes)
        out.backward(torch.ones_like(out))


for fn in cases:
    fn()
roi_pass()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for fn in cases:
    fn()
roi_pass()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
