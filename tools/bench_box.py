"""Per-launch timing of the box-branch GEMMs (fc6 / fc7 / predictor as N-tiled 1x1 convolutions) at the bench size M = 4096 ROIs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops
from sfvos_b200._lib import BF16

dev = "cuda"
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def act(m, c, dtype=torch.bfloat16):
    return ops.Act(torch.randn(m * c, device=dev).to(dtype), 1, 1, 1, m, c)


def timed(name, fn, flops, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{name:34s} {us:9.1f} us  {flops / us / 1e6:8.1f} TFLOP/s" if flops else f"{name:34s} {us:9.1f} us")


def fwd(K, N, out_dtype=torch.bfloat16):
    x, w = act(M, K), torch.randn(N, K, device=dev) / K ** 0.5
    wp = ops.pack_weights(w.view(N, K, 1, 1, 1), 0, BF16, (K + 63) // 64 * 64)
    b = torch.zeros(N, device=dev)
    y = ops.Act.empty(1, 1, 1, M, N, out_dtype, dev)
    return lambda: ops.conv(x, wp, (K + 63) // 64 * 64, N, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=b)


def wg(K, N):
    x, dy = act(M, K), act(M, N)
    dwp = torch.zeros(K * N, device=dev)
    return lambda: ops.wgrad(x, dy, (1, 1, 1), (0, 0, 0), dwp, umma=True)


for name, K, N in (("fc6 fprop  12544 -> 1024", 12544, 1024), ("fc7 fprop   1024 -> 1024", 1024, 1024), ("pred fprop  1024 -> 64", 1024, 64),
                   ("pred dgrad    64 -> 1024", 64, 1024), ("fc7 dgrad   1024 -> 1024", 1024, 1024), ("fc6 dgrad   1024 -> 12544", 1024, 12544)):
    timed(name, fwd(K, N), 2.0 * M * K * N)
for name, K, N in (("fc6 wgrad  12544 x 1024", 12544, 1024), ("fc7 wgrad   1024 x 1024", 1024, 1024), ("pred wgrad  1024 x 64", 1024, 64)):
    timed(name, wg(K, N), 2.0 * M * K * N)
w6 = torch.randn(1024, 12544, device=dev)
timed("pack fc6 fprop (f32->bf16)", lambda: ops.pack_weights(w6.view(1024, 12544, 1, 1, 1), 0, BF16, 12544), 0)
timed("pack fc6 dgrad (transpose)", lambda: ops.pack_weights(w6.view(1024, 12544, 1, 1, 1), 1, BF16, 1024), 0)
dwp = torch.zeros(12544 * 1024, device=dev); gw = torch.zeros(1024, 12544, device=dev)
timed("unpack fc6 wgrad (transpose +=)", lambda: ops.unpack_wgrad(dwp, gw.view(1024, 12544, 1, 1, 1), 0), 0)
timed("zero fc6 dw (51 MB)", lambda: dwp.zero_(), 0)
