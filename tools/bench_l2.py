"""L2 -> SM stream bandwidth of this B200, the denominator for kernels whose bytes are served by L2 rather than HBM (ROIAlign's
gather).  A read-only reduction (torch.sum, a plain grid-stride streaming kernel) over a buffer that FITS in the 126 MB L2
(32 / 64 MB, warmed), against the same reduction over a 4 GB buffer (HBM).  CUDA events, median of 20."""
import json

import torch


def bw(nbytes, reps=20):
    x = torch.empty(nbytes // 4, dtype=torch.float32, device="cuda").normal_()
    for _ in range(5):
        x.sum()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x.sum(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    return nbytes / ms / 1e6


def probe(kind, nbytes, iters, reps=9):
    """libsfvos probe kernel: kind 0 = 16-byte loads, 1 = red.global.add.v4.f32; returns GB/s of vectors touched."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from sfvos_b200 import ops
    from sfvos_b200._lib import call
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device="cuda")
    sink = torch.zeros(1, device="cuda")
    for _ in range(2):
        call("sfvos_probe_l2", kind, ops._p(buf), nbytes, iters, ops._p(sink), ops.stream())
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call("sfvos_probe_l2", kind, ops._p(buf), nbytes, iters, ops._p(sink), ops.stream()); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    return nbytes * iters / ms / 1e6


if __name__ == "__main__":
    out = {"l2_resident_32MB_gbs": round(bw(32 << 20), 1), "l2_resident_64MB_gbs": round(bw(64 << 20), 1),
           "hbm_4GB_gbs": round(bw(4 << 30), 1), "how": "torch.sum over f32, CUDA events, median of 20 after 5 warm-ups",
           "probe_ld16_l2_resident_48MB_gbs": round(probe(0, 48 << 20, 20), 1), "probe_ld16_hbm_2GB_gbs": round(probe(0, 2 << 30, 1), 1),
           "probe_red_v4_l2_resident_48MB_gbs": round(probe(1, 48 << 20, 20), 1), "probe_red_v4_hbm_2GB_gbs": round(probe(1, 2 << 30, 1), 1),
           "probe_how": "sfvos_probe_l2 (grid-stride kernel, 16 CTAs x 256 threads per SM), GB/s = bytes of 16-byte vectors touched / time; "
                        "red_v4 = red.global.add.v4.f32, one 16-byte vector per thread per step; median of 9"}
    print(json.dumps(out))
