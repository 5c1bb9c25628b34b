"""Warm per-kernel durations of the bench step (C2 defaults), from torch.profiler (CUPTI) around eager steps on ONE stream:
writes every kernel launch of one step in order (name, us) and the per-kernel totals.
usage: python tools/profile_timeline.py [--out gpurun_out/timeline.csv] [--streams]   (--streams: keep the level streams)"""
import argparse, os, re, sys
from collections import OrderedDict, defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    ap.add_argument("--sp", type=int, default=1); ap.add_argument("--fp", type=int, default=8); ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--streams", action="store_true")
    a = ap.parse_args()
    if not a.streams:
        os.environ["SFVOS_LEVEL_STREAMS"] = "0"
    import torch
    from sfvos_b200 import dp, ops, workload as wl
    dev = torch.device("cuda", 0)
    step = wl.HotPathStep(a.sp, a.fp, a.B, 512, 128, device=dev, precision="bf16")
    params = step.parameters()
    seq = wl.synthetic_sequence(a.B + a.fp - 1, seed=1234, device=dev, dtype=torch.bfloat16)
    clips = wl.sequence_windows(seq, a.fp, 0, a.B)
    arena = dp.GradArena(step.groups(), dev)
    ops.GRAD_ARENA = arena
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4, foreach=True)

    def eager_step():
        arena.zero()
        for p in params:
            p.grad = None
        loss, merged = step.forward(clips)
        step.backward_split(loss, merged, None)
        opt.step()

    for _ in range(4):
        eager_step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eager_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
    evs = sorted(evs, key=lambda e: e.time_range.start)
    short = lambda n: re.sub(r"^void ", "", re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "")))
    tot = defaultdict(lambda: [0, 0.0])
    with open(a.out, "w") as f:
        f.write("idx,start_us,dur_us,kernel\n")
        t0 = evs[0].time_range.start if evs else 0
        for i, e in enumerate(evs):
            d = e.time_range.end - e.time_range.start
            n = short(e.name)
            f.write(f"{i},{e.time_range.start - t0:.1f},{d:.1f},{n.replace(',', ';')}\n")
            tot[n][0] += 1; tot[n][1] += d
    total = sum(v[1] for v in tot.values())
    span = (evs[-1].time_range.end - evs[0].time_range.start) if evs else 0
    print(f"kernel time {total/1e3:.2f} ms over {len(evs)} launches; span {span/1e3:.2f} ms")
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t:9.1f} us {c:4d}x  {n[:110]}")


if __name__ == "__main__":
    main()
