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


if __name__ == "__main__":
    out = {"l2_resident_32MB_gbs": round(bw(32 << 20), 1), "l2_resident_64MB_gbs": round(bw(64 << 20), 1),
           "hbm_4GB_gbs": round(bw(4 << 30), 1), "how": "torch.sum over f32, CUDA events, median of 20 after 5 warm-ups"}
    print(json.dumps(out))
