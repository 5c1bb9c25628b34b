"""torch.profiler breakdown of one full bench step (SlowFast module + ROIAlign + mask branch, fwd+bwd) by kernel."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops, workload as wl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sp", type=int, default=1); ap.add_argument("--fp", type=int, default=8)
    ap.add_argument("--B", type=int, default=8); ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=45)
    a = ap.parse_args()
    step = wl.HotPathStep(a.sp, a.fp, a.B, 512, 128, device="cuda")
    feats = wl.synthetic_features(a.B, a.fp, device="cuda")
    for _ in range(3):
        step.step(feats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        step.step(feats)
    e1.record()
    t_launch = (time.perf_counter() - t0) / a.steps
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"sp={a.sp} fp={a.fp} B={a.B}: {ms:.2f} ms/step GPU, {t_launch*1e3:.2f} ms/step host launch time -> "
          f"{a.B * a.fp / ms * 1e3:.1f} clip-frames/s; peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.step(feats); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=70))


if __name__ == "__main__":
    main()
