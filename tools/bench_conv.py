"""Time single conv / wgrad launches at full level-0 sizes (CUDA events), for kernel tuning.
usage: python tools/bench_conv.py [--B 8] [--reps 5] case [case ...]   (cases: fast1 fast2 fast3 slow1 slow2 slow3 f2s1 f2s2, +d = dgrad, +w = wgrad)"""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sfvos_b200 import ops

# name: (T_in, Cin, Cout, kt, khw) at (sp,fp) = (1,8)
CASES = {"fast1": (8, 256, 32, 3, 3), "fast2": (6, 32, 32, 3, 3), "fast3": (4, 32, 32, 4, 3), "slow1": (1, 256, 192, 1, 3),
         "slow2": (1, 256, 192, 1, 3), "slow3": (1, 256, 224, 1, 3), "f2s1": (6, 32, 64, 6, 1), "f2s2": (4, 32, 64, 4, 1),
         "fast1_16": (16, 256, 32, 6, 3), "fast1_32": (32, 256, 32, 11, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8); ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--H", type=int, default=192); ap.add_argument("--W", type=int, default=336)
    ap.add_argument("cases", nargs="+")
    a = ap.parse_args()
    dev = "cuda"
    for case in a.cases:
        mode = "f"
        name = case
        if case.endswith("+d"): mode, name = "d", case[:-2]
        if case.endswith("+w"): mode, name = "w", case[:-2]
        if name == "convt":            # one tap of the mask predictor's ConvTranspose2d(256,256,2,2): 1x1 GEMM + scatter to 28x28
            K, Hm = 1024, 14
            x = ops.Act(torch.randn(K * Hm * Hm * 256, device=dev).bfloat16(), K, 1, Hm, Hm, 256)
            wt = torch.randn(256, 256, 2, 2, device=dev) / 16
            wp = ops.pack_weights(wt, 2, ops.BF16, 256, (0, 0))
            up = ops.Act.empty(K, 1, 2 * Hm, 2 * Hm, 256, torch.bfloat16, dev)
            bias = torch.zeros(256, device=dev)
            fn = lambda: ops.conv(x, wp, 256, 256, (1, 1, 1), (0, 0, 0), 1, up, umma=True, relu=True, shift=bias, scatter=(2 * Hm, 2 * Hm, 2, 0, 2, 0))
            flops = 2.0 * K * Hm * Hm * 256 * 256
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            print(f"{case:12s} 1024 ROIs 14x14: {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)
            continue
        if name in ("convt4", "maskconv", "fc6"):
            # mask predictor ConvTranspose2d as ONE 1x1 GEMM 256 -> 1024 (bf16 out, bias + ReLU) / mask-head 3x3 conv fprop (bias + ReLU,
            # bf16 out) or dgrad with the fused ReLU backward / box head fc6 (12544 -> 1024) as a 1x1 conv over 4096 "pixels"
            K, Hm = 1024, 14
            if name == "convt4" and mode == "d":      # data gradient of the ConvTranspose: 1x1 GEMM 1024 -> 256
                x = ops.Act(torch.randn(K * Hm * Hm * 1024, device=dev).bfloat16(), K, 1, Hm, Hm, 1024)
                wp = (torch.randn(256, 1024, device=dev) / 32).bfloat16().contiguous()
                y = ops.Act.empty(K, 1, Hm, Hm, 256, torch.bfloat16, dev)
                fn = lambda: ops.conv(x, wp, 1024, 256, (1, 1, 1), (0, 0, 0), 1, y, umma=True)
                flops = 2.0 * K * Hm * Hm * 256 * 1024
            elif name == "convt4":
                x = ops.Act(torch.randn(K * Hm * Hm * 256, device=dev).bfloat16(), K, 1, Hm, Hm, 256)
                wp = (torch.randn(1024, 256, device=dev) / 16).bfloat16().contiguous()
                y = ops.Act.empty(K, 1, Hm, Hm, 1024, torch.bfloat16, dev)
                bias = torch.zeros(1024, device=dev)
                fn = lambda: ops.conv(x, wp, 256, 1024, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=bias)
                flops = 2.0 * K * Hm * Hm * 256 * 1024
            elif name == "fc6":
                M = 4096
                x = ops.Act(torch.randn(M * 12544, device=dev).bfloat16(), 1, 1, 1, M, 12544)
                wp = (torch.randn(1024, 12544, device=dev) / 112).bfloat16().contiguous()
                y = ops.Act.empty(1, 1, 1, M, 1024, torch.bfloat16, dev)
                bias = torch.zeros(1024, device=dev)
                fn = lambda: ops.conv(x, wp, 12544, 1024, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=bias)
                flops = 2.0 * M * 12544 * 1024
            else:
                x = ops.Act(torch.randn(K * Hm * Hm * 256, device=dev).bfloat16(), K, 1, Hm, Hm, 256)
                wt = torch.randn(256, 256, 1, 3, 3, device=dev) / 48
                wp = ops.pack_weights(wt, 1 if mode == "d" else 0, ops.BF16, 256)
                y = ops.Act.empty(K, 1, Hm, Hm, 256, torch.bfloat16, dev)
                bias = torch.zeros(256, device=dev)
                if mode == "d":
                    act = ops.Act(torch.randn(K * Hm * Hm * 256, device=dev).relu().bfloat16(), K, 1, Hm, Hm, 256)
                    fn = lambda: ops.conv(x, wp, 256, 256, (1, 3, 3), (0, 1, 1), 1, y, umma=True, relu_mask=act, dbias=bias)
                else:
                    fn = lambda: ops.conv(x, wp, 256, 256, (1, 3, 3), (0, 1, 1), 1, y, umma=True, relu=True, shift=bias)
                flops = 2.0 * K * Hm * Hm * 256 * 256 * 9
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            envs = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SFVOS_"))
            print(f"{case:12s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s   [{envs}]", flush=True)
            continue
        T, cin, cout, kt, khw = CASES[name]
        pad = 1 if khw == 3 else 0
        To = T - kt + 1
        B, H, W = a.B, a.H, a.W
        w = torch.randn(cout, cin, kt, khw, khw, device=dev) / math.sqrt(cin * kt * khw * khw)
        flops = 2.0 * B * To * H * W * cout * cin * kt * khw * khw
        if mode == "f":
            x = ops.Act(torch.randn(B * T * H * W * cin, device=dev).bfloat16(), B, T, H, W, cin)
            cp = 32 if cin <= 32 else (cin + 63) // 64 * 64
            wp = ops.pack_weights(w, 0, ops.BF16, cp)
            y = ops.Act.empty(B, To, H, W, cout, torch.float32, dev)
            stats = torch.zeros(2 * cout, device=dev)
            fn = lambda: ops.conv(x, wp, cp, cout, (kt, khw, khw), (0, pad, pad), To, y, umma=True, stats=stats)
        elif mode == "d":
            dy = ops.Act(torch.randn(B * To * H * W * cout, device=dev).bfloat16(), B, To, H, W, cout)
            cp = 32 if cout <= 32 else (cout + 63) // 64 * 64
            wd = ops.pack_weights(w, 1, ops.BF16, cp)
            dx = ops.Act.empty(B, T, H, W, cin, torch.bfloat16 if os.environ.get("BENCH_DX_BF16") else torch.float32, dev)
            addend = None
            if os.environ.get("BENCH_ADDEND"):          # f32 | bf16: y = dgrad + addend (the lateral on top of the fast conv's partial gradient)
                adt = torch.bfloat16 if os.environ["BENCH_ADDEND"] == "bf16" else torch.float32
                addend = ops.Act(torch.randn(B * T * H * W * cin, device=dev).to(adt), B, T, H, W, cin)
            fn = lambda: ops.conv(dy, wd, cp, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, dx, umma=True, addend=addend)
        else:
            x = ops.Act(torch.randn(B * T * H * W * cin, device=dev).bfloat16(), B, T, H, W, cin)
            dy = ops.Act(torch.randn(B * To * H * W * cout, device=dev).bfloat16(), B, To, H, W, cout)
            dwp = torch.zeros(kt * khw * khw * cin * cout, device=dev)
            fn = lambda: ops.wgrad(x, dy, (kt, khw, khw), (0, pad, pad), dwp, umma=True)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        envs = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("SFVOS_"))
        print(f"{case:12s} B={B} {H}x{W}: {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s   [{envs}]", flush=True)


if __name__ == "__main__":
    main()
