"""BASELINE config C4: the full inference pipeline (torchvision ResNet-50-FPN backbone + RPN -> SlowFast module -> roi_heads
-> paste-back) on synthetic DAVIS-shaped sequences, sharded BY SEQUENCE across the ranks (no data-path collective), eval mode.

    python tools/bench_pipeline.py [--sp 1 --fp 8 --sequences 4 --frames 24]              # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_pipeline.py ...

Prints one JSON line per mode (sequence sweep on / off): frames/s over all ranks (max-over-ranks time).  Backbone (ResNet-50
body + FPN) and RPN head run on libsfvos (SFVOS_NATIVE_BACKBONE=0 = torchvision fp32); anchors, box decoding
and NMS are torchvision.  Random-init weights, synthetic frames (no network for DAVIS)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sfvos_b200 import dp, ops  # noqa: E402
from sfvos_b200.model import SegmentationModel  # noqa: E402


def synthetic_sequence(n, seed, h=480, w=854):
    g = torch.Generator().manual_seed(seed)
    imgs, targets = [], []
    for i in range(n):
        imgs.append(torch.rand(3, h, w, generator=g))
        x1, y1 = 100 + 6 * i, 120 + 3 * i
        m = torch.zeros(1, h, w, dtype=torch.uint8)
        m[0, y1:y1 + 200, x1:x1 + 300] = 1
        targets.append({"boxes": torch.tensor([[x1, y1, x1 + 300, y1 + 200]], dtype=torch.float32),
                        "labels": torch.ones(1, dtype=torch.int64), "masks": m})
    return imgs, targets


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sp", type=int, default=1); ap.add_argument("--fp", type=int, default=8)
    ap.add_argument("--sequences", type=int, default=4); ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--sweep-only", action="store_true", help="skip the reference-faithful per-frame loop (4-5x slower)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ops.device_check()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(63)
    import warnings
    warnings.simplefilter("ignore")
    model = SegmentationModel(device=dev, slow_pathway_size=a.sp, fast_pathway_size=a.fp, maskrcnn_weights=None, pretrained=False)
    model.to(dev).eval()
    model.maskrcnn_model.roi_heads.score_thresh = 0.0          # random-init scores: keep the 10 detections per frame
    model.sequence_chunk = a.chunk
    lengths = [a.frames + 4 * (i % 3) for i in range(a.sequences)]            # unequal lengths, like DAVIS
    mine = dp.shard_sequences(lengths, world)[rank]
    seqs = {i: synthetic_sequence(lengths[i], seed=i) for i in mine}

    def run(mode):
        model.sequence_mode = mode
        n = 0
        for i in mine:
            with torch.no_grad():
                _, dets = model(*seqs[i])
            n += len(dets)
        return n

    for mode in ((True,) if a.sweep_only else (True, False)):
        run(mode)                                              # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        model.phase_times.clear()
        t0 = time.perf_counter()
        run(mode)
        torch.cuda.synchronize()
        sec = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"config": "C4 full inference pipeline, sharded by sequence", "sequence_sweep": mode, "sp": a.sp, "fp": a.fp,
                              "n_gpus": world, "sequences": lengths, "frames_total": sum(lengths), "seconds": round(sec.item(), 3),
                              "frames_per_s": round(sum(lengths) / sec.item(), 2),
                              "phase_seconds_rank0": {k: round(v, 3) for k, v in model.phase_times.items()},
                              "backbone": "libsfvos ResNet-50 body + FPN + RPN head (bf16 NHWC)" if os.environ.get("SFVOS_NATIVE_BACKBONE", "1") != "0" else "torchvision fp32",
                              "note": "wall clock incl. backbone, RPN (torchvision anchors / NMS), SlowFast sweep, roi_heads, paste-back and the D2H of the pasted masks (as the reference does)"}),
                  flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
