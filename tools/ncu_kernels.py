"""One launch (after one warm-up launch) of every hot kernel at its full bench size, for `ncu --set full`.
Order of the captured launches (each twice: warm-up, measured):
  conv_umma slow_conv1 fprop | conv_umma slow_conv2 dgrad | conv_umma mask-head 3x3 (1024 ROIs) | conv_tstack fast_conv1 fprop |
  conv_tstack fast_conv2 fprop | wgrad_umma slow_conv1 | wgrad_halo fast_conv1 | wgrad_c32 fast_conv2 |
  roi_align fwd p7 | roi_align bwd p7 | roi_align fwd p14 | roi_align bwd p14 |
  box branch (M = 4096 ROIs): conv_umma fc6 fprop (4 N chunks) | conv_umma fc6 dgrad (49 N chunks) | wgrad_umma fc6 (98 x 4 blocks) |
  conv_umma ConvTranspose tap (resident weights, 1024 ROIs x 14x14) | roi_align fwd p7 with bf16 [K, C*P*P] rows"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops, workload as wl
from sfvos_b200.roi_heads import MultiScaleRoIAlign

dev = "cuda"
B, H, W = 8, 192, 336


def act(b, t, h, w, c):
    return ops.Act(torch.randn(b * t * h * w * c, device=dev).bfloat16(), b, t, h, w, c)


def conv_case(T, cin, cout, kt, khw, dgrad=False, b=B, h=H, w=W):
    pad = 1 if khw == 3 else 0
    To = T - kt + 1
    wt = torch.randn(cout, cin, kt, khw, khw, device=dev) / math.sqrt(cin * kt * khw * khw)
    if not dgrad:
        x = act(b, T, h, w, cin)
        cp = 32 if cin <= 32 else (cin + 63) // 64 * 64
        wp = ops.pack_weights(wt, 0, ops.BF16, cp)
        y = ops.Act.empty(b, To, h, w, cout, torch.float32, dev)
        stats = torch.zeros(2 * cout, device=dev)
        return lambda: ops.conv(x, wp, cp, cout, (kt, khw, khw), (0, pad, pad), To, y, umma=True, stats=stats)
    dy = act(b, To, h, w, cout)
    cp = 32 if cout <= 32 else (cout + 63) // 64 * 64
    wd = ops.pack_weights(wt, 1, ops.BF16, cp)
    dx = ops.Act.empty(b, T, h, w, cin, torch.float32, dev)
    return lambda: ops.conv(dy, wd, cp, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, dx, umma=True)


def wgrad_case(T, cin, cout, kt, khw):
    pad = 1 if khw == 3 else 0
    To = T - kt + 1
    x, dy = act(B, T, H, W, cin), act(B, To, H, W, cout)
    dwp = torch.zeros(kt * khw * khw * cin * cout, device=dev)
    return lambda: ops.wgrad(x, dy, (kt, khw, khw), (0, pad, pad), dwp, umma=True)


cases = [conv_case(1, 256, 192, 1, 3), conv_case(1, 256, 192, 1, 3, dgrad=True), conv_case(1, 256, 256, 1, 3, b=1024, h=14, w=14),
         conv_case(8, 256, 32, 3, 3), conv_case(6, 32, 32, 3, 3),
         wgrad_case(1, 256, 192, 1, 3), wgrad_case(8, 256, 32, 3, 3), wgrad_case(6, 32, 32, 3, 3)]
for fn in cases:
    fn(); fn()
torch.cuda.synchronize()

feats = {k: torch.randn(B, h, w, 256, device=dev).permute(0, 3, 1, 2).requires_grad_(True) for k, (h, w) in wl.LEVELS.items()}
box = [b.to(dev) for b in wl.synthetic_rois(B, 512)]
mask = [b[:128] for b in box]
shapes = [wl.IMAGE_HW] * B
for pool, rois in ((MultiScaleRoIAlign(wl.POOL_LEVELS, 7, 2, out_layout="nchw"), box), (MultiScaleRoIAlign(wl.POOL_LEVELS, 14, 2, out_layout="nhwc"), mask)):
    for _ in range(2):
        out = pool(feats, rois, shapes)
        out.backward(torch.ones_like(out))
torch.cuda.synchronize()

# ---- box branch GEMMs + the resident-weight ConvTranspose tap + bf16 box pooling ----
M = 4096


def rows(m, c):
    return ops.Act(torch.randn(m * c, device=dev).bfloat16(), 1, 1, 1, m, c)


def fc_case(K, N):
    x = rows(M, K)
    wp = ops.pack_weights((torch.randn(N, K, device=dev) / math.sqrt(K)).view(N, K, 1, 1, 1), 0, ops.BF16, K)
    y = ops.Act.empty(1, 1, 1, M, N, torch.bfloat16, dev)
    bias = torch.zeros(N, device=dev)
    return lambda: ops.conv(x, wp, K, N, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=bias)


def fc_wgrad_case(K, N):
    x, dy = rows(M, K), rows(M, N)
    dwp = torch.zeros(K * N, device=dev)
    return lambda: ops.wgrad(x, dy, (1, 1, 1), (0, 0, 0), dwp, umma=True)


def convt_tap_case():
    x = act(1024, 1, 14, 14, 256)
    wt = torch.randn(256, 256, 2, 2, device=dev) / 16
    wp = ops.pack_weights(wt, 2, ops.BF16, 256, (0, 0))
    up = ops.Act.empty(1024, 1, 28, 28, 256, torch.bfloat16, dev)
    bias = torch.zeros(256, device=dev)
    return lambda: ops.conv(x, wp, 256, 256, (1, 1, 1), (0, 0, 0), 1, up, umma=True, relu=True, shift=bias, scatter=(28, 28, 2, 0, 2, 0))


for fn in (fc_case(12544, 1024), fc_case(1024, 12544), fc_wgrad_case(12544, 1024), convt_tap_case()):
    fn(); fn()
pool = MultiScaleRoIAlign(wl.POOL_LEVELS, 7, 2, out_layout="nchw", out_dtype=torch.bfloat16)
for _ in range(2):
    pool(feats, box, shapes)
torch.cuda.synchronize()
print("ok")
