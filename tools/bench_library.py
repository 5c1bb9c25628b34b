"""Same-box LIBRARY baselines, op by op ("the kernel to beat"): what the reference's stock GPU path would
dispatch to for every hot op of the C2 step, timed on this box with CUDA events next to the libsfvos kernel for the same
shape.

  * Conv3d layers of SlowFastLayers at pyramid level 0 (192x336), B clips: cuDNN through torch.nn.functional.conv3d /
    aten::convolution_backward -- bf16 channels_last_3d (the fastest layout cuDNN offers here) and the reference's own fp32 NCDHW
    (TF32 as PyTorch defaults to) -- fprop, dgrad, wgrad separately (code/helpers/model.py:72-76,83-90)
  * mask-head Conv2d 256->256 3x3 on [K,256,14,14]: cuDNN bf16 channels_last (TV/models/detection/mask_rcnn.py:284-296)
  * torchvision.ops.roi_align CUDA fwd/bwd through torchvision's MultiScaleRoIAlign on the bench's ROI set
    (TV/ops/roi_align.py:258, TV/ops/poolers.py:289-321)

bench.py calls ``run()`` OUTSIDE its timed region and puts the table under ``roofline.vs_library``; every entry is
{ours_us, lib_us, speedup = lib_us / ours_us}.  CLI: python tools/bench_library.py [--B 8]"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def _time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3          # median, us


def _layers(sp, fp):
    """name -> (T_in, Cin, Cout, kt, khw, needs_dgrad) for (sp, fp), from the module's own construction rules."""
    from sfvos_b200 import SlowFastLayers
    m = SlowFastLayers(256, torch.device("cpu"), sp, fp)
    s = m._specs
    t1s, t1f = sp - s["slow_conv1"].kt + 1, fp - s["fast_conv1"].kt + 1
    t2s, t2f = t1s - s["slow_conv2"].kt + 1, t1f - s["fast_conv2"].kt + 1
    tin = {"fast_conv1": fp, "slow_conv1": sp, "conv_f2s1": t1f, "fast_conv2": t1f, "slow_conv2": t1s, "conv_f2s2": t2f,
           "fast_conv3": t2f, "slow_conv3": t2s}
    return {n: (tin[n], s[n].cin, s[n].cout, s[n].kt, s[n].khw, n not in ("fast_conv1", "slow_conv1")) for n in tin}


def conv3d_table(sp=1, fp=8, B=8, H=192, W=336, dev="cuda", fp32_too=True):
    from sfvos_b200 import ops
    out = {}
    for name, (T, cin, cout, kt, khw, need_dx) in _layers(sp, fp).items():
        pad = 1 if khw == 3 else 0
        To = T - kt + 1
        flops = 2.0 * B * To * H * W * cout * cin * kt * khw * khw
        g = torch.Generator(device=dev).manual_seed(1)
        x5 = torch.randn(B, cin, T, H, W, device=dev, generator=g, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        w5 = (torch.randn(cout, cin, kt, khw, khw, device=dev, generator=g) / math.sqrt(cin * kt * khw * khw))
        wb = w5.bfloat16().contiguous(memory_format=torch.channels_last_3d)
        dy5 = torch.randn(B, cout, To, H, W, device=dev, generator=g, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        padding = (0, pad, pad)
        rec = {"gflop": round(flops / 1e9, 1)}

        def bwd(mask, x=x5, w=wb, dy=dy5):
            return torch.ops.aten.convolution_backward(dy, x, w, None, (1, 1, 1), padding, (1, 1, 1), False, (0, 0, 0), 1, mask)
        lib = {"fprop": _time(lambda: F.conv3d(x5, wb, None, padding=padding)),
               "wgrad": _time(lambda: bwd([False, True, False]))}
        if need_dx:
            lib["dgrad"] = _time(lambda: bwd([True, False, False]))
        # libsfvos on the same shape (channels-last bf16 activations; fprop with the fused BN-statistics epilogue)
        xa = ops.Act(x5.permute(0, 2, 3, 4, 1).reshape(-1), B, T, H, W, cin)
        dya = ops.Act(dy5.permute(0, 2, 3, 4, 1).reshape(-1), B, To, H, W, cout)
        cp = 32 if cin <= 32 else (cin + 63) // 64 * 64
        wp = ops.pack_weights(w5, 0, ops.BF16, cp)
        y = ops.Act.empty(B, To, H, W, cout, torch.float32, dev)
        stats = torch.zeros(2 * cout, device=dev)
        ours = {"fprop": _time(lambda: ops.conv(xa, wp, cp, cout, (kt, khw, khw), padding, To, y, umma=True, stats=stats))}
        dwp = torch.zeros(kt * khw * khw * cin * cout, device=dev)
        ours["wgrad"] = _time(lambda: ops.wgrad(xa, dya, (kt, khw, khw), padding, dwp, umma=True))
        if need_dx:
            cpd = 32 if cout <= 32 else (cout + 63) // 64 * 64
            wd = ops.pack_weights(w5, 1, ops.BF16, cpd)
            dx = ops.Act.empty(B, T, H, W, cin, torch.float32, dev)
            ours["dgrad"] = _time(lambda: ops.conv(dya, wd, cpd, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, dx, umma=True))
        for op in lib:
            rec[op] = {"ours_us": round(ours[op], 1), "cudnn_bf16_ndhwc_us": round(lib[op], 1), "speedup": round(lib[op] / ours[op], 2),
                       "ours_tflops": round(flops / ours[op] / 1e6, 1), "cudnn_tflops": round(flops / lib[op] / 1e6, 1)}
        if fp32_too:                     # the reference's own dtype / layout on a GPU: fp32 NCDHW, TF32 as PyTorch defaults to
            del x5, dy5, xa, dya, y
            xf = torch.randn(B, cin, T, H, W, device=dev, generator=g)
            wf = w5
            rec["fprop"]["cudnn_fp32_ncdhw_us"] = round(_time(lambda: F.conv3d(xf, wf, None, padding=padding), reps=3, warm=1), 1)
            del xf
        out[name] = rec
        torch.cuda.empty_cache()
    return out


def mask_head_table(K=1024, dev="cuda"):
    from sfvos_b200 import ops
    g = torch.Generator(device=dev).manual_seed(2)
    x = torch.randn(K, 256, 14, 14, device=dev, generator=g, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(256, 256, 3, 3, device=dev, generator=g) / 48.0
    wb = w.bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(K, 256, 14, 14, device=dev, generator=g, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    flops = 2.0 * K * 196 * 256 * 2304

    def bwd(mask):
        return torch.ops.aten.convolution_backward(dy, x, wb, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, mask)
    lib = {"fprop": _time(lambda: F.conv2d(x, wb, None, padding=1)), "dgrad": _time(lambda: bwd([True, False, False])),
           "wgrad": _time(lambda: bwd([False, True, False]))}
    xa = ops.Act(x.permute(0, 2, 3, 1).reshape(-1), K, 1, 14, 14, 256)
    dya = ops.Act(dy.permute(0, 2, 3, 1).reshape(-1), K, 1, 14, 14, 256)
    wp = ops.pack_weights(w, 0, ops.BF16, 256)
    wd = ops.pack_weights(w, 1, ops.BF16, 256)
    y = ops.Act.empty(K, 1, 14, 14, 256, torch.bfloat16, dev)
    dx = ops.Act.empty(K, 1, 14, 14, 256, torch.bfloat16, dev)
    bias = torch.zeros(256, device=dev)
    dwp = torch.zeros(9 * 256 * 256, device=dev)
    ours = {"fprop": _time(lambda: ops.conv(xa, wp, 256, 256, (1, 3, 3), (0, 1, 1), 1, y, umma=True, relu=True, shift=bias)),
            "dgrad": _time(lambda: ops.conv(dya, wd, 256, 256, (1, 3, 3), (0, 1, 1), 1, dx, umma=True)),
            "wgrad": _time(lambda: ops.wgrad(xa, dya, (1, 3, 3), (0, 1, 1), dwp, umma=True))}
    return {op: {"ours_us": round(ours[op], 1), "cudnn_bf16_nhwc_us": round(lib[op], 1), "speedup": round(lib[op] / ours[op], 2),
                 "ours_tflops": round(flops / ours[op] / 1e6, 1), "cudnn_tflops": round(flops / lib[op] / 1e6, 1)} for op in lib}


def roi_align_table(B=8, k_box=512, k_mask=128, dev="cuda"):
    """torchvision's MultiScaleRoIAlign (fp32 NCHW, per-level torchvision.ops.roi_align launches + index glue) vs the one-launch
    multi-level kernel, on the bench's ROI set; both produce the gradient w.r.t. all four feature maps in the backward."""
    from torchvision.ops import MultiScaleRoIAlign as TVPool
    from sfvos_b200 import workload as wl
    from sfvos_b200.roi_heads import MultiScaleRoIAlign
    g = torch.Generator(device=dev).manual_seed(3)
    base = {k: torch.randn(B, 256, h, w, device=dev, generator=g) for k, (h, w) in wl.LEVELS.items() if k in wl.POOL_LEVELS}
    box = [b.to(dev) for b in wl.synthetic_rois(B, k_box)]
    sets = {"box_p7": (7, box, "nchw"), "mask_p14": (14, [b[:k_mask] for b in box], "nhwc")}
    shapes = [wl.IMAGE_HW] * B
    out = {}
    for tag, (P, rois, layout) in sets.items():
        rec = {}
        for who in ("lib", "ours"):
            if who == "lib":
                feats = {k: v.clone().requires_grad_(True) for k, v in base.items()}
                pool = TVPool(wl.POOL_LEVELS, P, 2)
            else:
                feats = {k: v.contiguous(memory_format=torch.channels_last).clone().requires_grad_(True) for k, v in base.items()}
                pool = MultiScaleRoIAlign(wl.POOL_LEVELS, P, 2, out_layout=layout, precision="bf16")
            y = pool(feats, rois, shapes)
            gy = torch.ones_like(y)
            rec[who + "_fwd"] = _time(lambda: pool(feats, rois, shapes))

            def fb():
                yy = pool(feats, rois, shapes)
                yy.backward(gy)
                for f in feats.values():
                    f.grad = None
            rec[who + "_fwd_bwd"] = _time(fb)
        out[tag] = {"fwd": {"ours_us": round(rec["ours_fwd"], 1), "torchvision_us": round(rec["lib_fwd"], 1),
                            "speedup": round(rec["lib_fwd"] / rec["ours_fwd"], 2)},
                    "fwd_bwd": {"ours_us": round(rec["ours_fwd_bwd"], 1), "torchvision_us": round(rec["lib_fwd_bwd"], 1),
                                "speedup": round(rec["lib_fwd_bwd"] / rec["ours_fwd_bwd"], 2)}}
    return out


def run(sp=1, fp=8, B=8, k_box=512, k_mask=128, dev="cuda", fp32_too=True):
    torch.backends.cudnn.benchmark = True         # let cuDNN pick its best algorithm per shape: the baseline at its best
    res = {"conv3d_level0": conv3d_table(sp, fp, B, dev=dev, fp32_too=fp32_too), "mask_head_conv": mask_head_table(B * k_mask, dev),
           "roi_align": roi_align_table(B, k_box, k_mask, dev),
           "note": "median of 5 launches after 2 warm-ups, CUDA events, one op at a time (burst clocks); cuDNN with benchmark=True"}
    torch.backends.cudnn.benchmark = False
    return res


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8); ap.add_argument("--sp", type=int, default=1); ap.add_argument("--fp", type=int, default=8)
    a = ap.parse_args()
    print(json.dumps(run(a.sp, a.fp, a.B), indent=1))
