"""Time the SlowFast module fwd+bwd at full DAVIS-shaped sizes and print a per-kernel breakdown (torch.profiler)."""
import sys, os, time, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
from sfvos_b200 import SlowFastLayers, ops

LEVELS = OrderedDict([("0", (192, 336)), ("1", (96, 168)), ("2", (48, 84)), ("3", (24, 42)), ("pool", (12, 21))])

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sp", type=int, default=1); ap.add_argument("--fp", type=int, default=8)
    ap.add_argument("--B", type=int, default=8); ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--levels", default="0,1,2,3,pool"); ap.add_argument("--no-prof", action="store_true")
    ap.add_argument("--fwd-only", action="store_true")
    a = ap.parse_args()
    levels = OrderedDict((k, LEVELS[k]) for k in a.levels.split(","))
    torch.manual_seed(63)
    m = SlowFastLayers(256, torch.device("cuda"), a.sp, a.fp).cuda().train()
    g = torch.Generator(device="cuda").manual_seed(1)
    fast = [OrderedDict((k, torch.randn(a.fp, 256, h, w, device="cuda", generator=g)) for k, (h, w) in levels.items()) for _ in range(a.B)]
    lo = a.fp // 2 - a.sp // 2
    slow = [OrderedDict((k, v[lo:lo + a.sp]) for k, v in f.items()) for f in fast]
    proj = {k: torch.randn(a.B, 256, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last) / (a.B * 256 * h * w) for k, (h, w) in levels.items()}

    def step():
        out = m.temporally_enhance_features(slow, fast)
        if a.fwd_only:
            return
        loss = sum((out[k] * proj[k]).sum() for k in out)
        loss.backward()
        for p in m.parameters():
            p.grad = None

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"sp={a.sp} fp={a.fp} B={a.B} levels={a.levels}: {ms:.2f} ms/step -> {a.B * a.fp / ms * 1e3:.1f} clip-frames/s; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    if a.no_prof:
        return
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("umma" in e.name or "wgrad_stack" in e.name)]
    evs.sort(key=lambda e: e.time_range.start)
    print("per-launch timeline of the tensor-core kernels (us):")
    for i, e in enumerate(evs):
        kind = 'wstck' if 'stack' in e.name else 'wgrad' if 'wgrad' in e.name else 'conv '
        print(f"  {i:3d} {kind} {e.time_range.end - e.time_range.start:9.1f}")

if __name__ == "__main__":
    main()
