"""Eval-mode SlowFast throughput at full DAVIS-shaped sizes: the reference's per-frame windows (B windows per call) vs the
single temporal sweep of SlowFastLayers.temporally_enhance_sequence.  Prints one JSON line per mode.

    python tools/bench_inference.py --sp 1 --fp 8 --frames 32 [--steps 3]
"""
import argparse
import json
import os
import sys
from collections import OrderedDict
from math import ceil, floor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sfvos_b200 import SlowFastLayers, ops, workload as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sp", type=int, default=1)
    ap.add_argument("--fp", type=int, default=8)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--window-batch", type=int, default=8)
    ap.add_argument("--chunk", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda")
    ops.device_check()
    torch.manual_seed(63)
    mod = SlowFastLayers(256, dev, a.sp, a.fp).cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = OrderedDict((k, torch.randn(a.frames, 256, h, w, device=dev, generator=g)) for k, (h, w) in wl.LEVELS.items())
    n, sp, fp = a.frames, a.sp, a.fp
    zero = {k: torch.zeros_like(v[0]) for k, v in frames.items()}

    def window(t):
        idx = range(t - floor(fp / 2), t + ceil(fp / 2))
        fast = OrderedDict((k, torch.stack([v[i] if 0 <= i < n else zero[k] for i in idx])) for k, v in frames.items())
        p = fp // 2
        slow = OrderedDict((k, v[p - floor(sp / 2):p + ceil(sp / 2)]) for k, v in fast.items())
        return slow, fast

    def per_window():
        outs = []
        for t0 in range(0, n, a.window_batch):
            ws = [window(t) for t in range(t0, min(n, t0 + a.window_batch))]
            with torch.no_grad():
                outs.append(mod.temporally_enhance_features([w[0] for w in ws], [w[1] for w in ws]))
        return outs

    def sweep():
        return mod.temporally_enhance_sequence(frames, max_frames=a.chunk or None)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    ref = per_window()
    seq = sweep()
    err = 0.0
    for i, o in enumerate(ref):
        for k in o:
            r = o[k]
            s = seq[k][i * a.window_batch:i * a.window_batch + r.shape[0]]
            err = max(err, (s - r).abs().max().item() / (r.abs().max().item() + 1e-12))
    fwd_window = wl.conv_flops(sp, fp, fwd_only=True)
    for name, fn in (("per_window", per_window), ("sequence_sweep", sweep)):
        ms = timed(fn)
        print(json.dumps({"mode": name, "sp": sp, "fp": fp, "frames": n, "ms": round(ms, 2), "frames_per_s": round(n / ms * 1e3, 1),
                          "window_equiv_tflops": round(fwd_window * n / (ms * 1e-3) / 1e12, 1),
                          "max_norm_diff_vs_per_window": float(f"{err:.3e}"),
                          "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)


if __name__ == "__main__":
    main()
