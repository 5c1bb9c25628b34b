"""Print per-parameter gradient errors (max-normalised and relative L2) of the GPU module vs the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
from oracle import slowfast_oracle as so
from sfvos_b200 import SlowFastLayers

def run(sp, fp, precision, levels, n_clips=2, emulate=False):
    fast = [so.synthetic_clip(levels, fp, seed=1234 + 100 * c, zero_left=(fp // 2 if c == 1 else 0)) for c in range(n_clips)]
    slow = [so.slice_window(f, fp // 2, sp) for f in fast]
    sd = so.init_state_dict(sp, fp, seed=63)
    ref_out, ref_loss, grads, _ = so.grads_of(sd, slow, fast, emulate_bf16=emulate)
    torch.manual_seed(63)
    m = SlowFastLayers(256, torch.device("cuda"), sp, fp).cuda().train()
    m.precision = precision
    fc = [OrderedDict((k, v.cuda()) for k, v in f.items()) for f in fast]
    sc = [so.slice_window(f, fp // 2, sp) for f in fc]
    out = m.temporally_enhance_features(sc, fc)
    so.module_loss(out).backward()
    print(f"--- sp={sp} fp={fp} {precision} levels={dict(levels)} vs {'bf16-emulated' if emulate else 'fp32'} oracle")
    for k in out:
        d = out[k].detach().cpu() - ref_out[k].detach()
        print(f"  out[{k}] maxnorm {d.abs().max()/ref_out[k].abs().max():.3e}  relL2 {d.norm()/ref_out[k].norm():.3e}")
    for name, p in m.named_parameters():
        ref = grads[name]; d = p.grad.cpu() - ref
        print(f"  {name:22s} maxnorm {d.abs().max().item()/(ref.abs().max().item()+1e-20):.3e}  relL2 {d.norm().item()/(ref.norm().item()+1e-20):.3e}  |ref|max {ref.abs().max().item():.2e}")

if __name__ == "__main__":
    run(1, 8, "bf16", OrderedDict([("0", (8, 12)), ("pool", (4, 6))]))
    run(1, 8, "bf16", OrderedDict([("0", (48, 84))]))
    run(1, 8, "fp32", OrderedDict([("0", (8, 12)), ("pool", (4, 6))]))
