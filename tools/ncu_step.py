"""One eager bench step (C2, one stream) between cudaProfilerStart/Stop after warm-ups, for
`ncu --profile-from-start off -k regex:<kernels> --set full ... python tools/ncu_step.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SFVOS_LEVEL_STREAMS"] = "0"
import torch
from sfvos_b200 import dp, ops, workload as wl

dev = torch.device("cuda", 0)
B, FP = 8, 8
step = wl.HotPathStep(1, FP, B, 512, 128, device=dev, precision="bf16")
params = step.parameters()
seq = wl.synthetic_sequence(B + FP - 1, seed=1234, device=dev, dtype=torch.bfloat16)
clips = wl.sequence_windows(seq, FP, 0, B)
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


for _ in range(3):
    eager_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eager_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
