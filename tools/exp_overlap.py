"""Experiment: how much does running pyramid level 0 and levels 1..4 of the SlowFast fwd+bwd on two CUDA streams overlap?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
from sfvos_b200 import SlowFastLayers, workload as wl

dev = torch.device("cuda")
B, sp, fp = 8, 1, 8
torch.manual_seed(63)
m = SlowFastLayers(256, dev, sp, fp).cuda().train()
g = torch.Generator(device="cuda").manual_seed(1)
fast = [OrderedDict((k, torch.randn(fp, 256, h, w, device=dev, generator=g)) for k, (h, w) in wl.LEVELS.items()) for _ in range(B)]
lo = fp // 2 - sp // 2


def sub(keys):
    f = [OrderedDict((k, c[k]) for k in keys) for c in fast]
    s = [OrderedDict((k, v[lo:lo + sp]) for k, v in c.items()) for c in f]
    proj = {k: torch.randn(B, 256, *wl.LEVELS[k], device=dev, generator=g).contiguous(memory_format=torch.channels_last) for k in keys}
    return s, f, proj


def step(part):
    s, f, proj = part
    out = m.temporally_enhance_features(s, f)
    loss = sum((out[k] * proj[k]).sum() for k in out)
    loss.backward()


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


p0, p1, pall = sub(["0"]), sub(["1", "2", "3", "pool"]), sub(list(wl.LEVELS))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa):
        step(p0)
    with torch.cuda.stream(sb):
        step(p1)
    cur.wait_stream(sa); cur.wait_stream(sb)


print(f"level 0 alone        {timed(lambda: step(p0)):7.2f} ms")
print(f"levels 1..pool alone {timed(lambda: step(p1)):7.2f} ms")
print(f"all levels, 1 stream {timed(lambda: step(pall)):7.2f} ms")
print(f"two streams          {timed(both):7.2f} ms")
