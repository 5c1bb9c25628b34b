#!/usr/bin/env python
"""bench.py -- headline benchmark of the SlowFast-VOS hot path (BASELINE.json metric: clip-frames/s fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: SlowFast temporal module (sp=1, fp=8) on B=8 clips x 5 FPN levels of a
480x854 DAVIS frame + multi-level ROIAlign (512 box / 128 mask ROIs per clip) + box head (fc6/fc7/predictor) +
fastrcnn_loss + mask head + mask loss, forward AND backward, bf16 tensor-core path.  One rank per GPU; every rank runs its own B clips (weak scaling)
and the trainable gradients are summed with one NCCL all-reduce per step when N > 1.

Prints ONE JSON line (rank 0): value = device-timed clip-frames/s with inputs resident in HBM; e2e = the same step
through the public API fed from pinned HOST buffers (H2D of the features and D2H of the loss inside the timed
region); roofline = achieved TFLOP/s of the tensor-core conv kernels from CUDA events around every launch of the
timed region; cpu_baseline = the reference's CPU path (oracle port, torch/torchvision CPU fp32) on a bounded
sample, timed on this box's host cores.  `--impl reference` times only that CPU path.
"""
import argparse
import json
from collections import OrderedDict
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "clip_frames_per_sec_fwd_bwd"
UNIT = "clip-frames/s"
SP, FP, B_PER_GPU, K_BOX, K_MASK = 1, 8, 8, 512, 128
# BASELINE.json configs by (sp, fp, clips per GPU): "alpha = 8" reads as sp = fp / 8 (notation)
CONFIGS = {"c2": (1, 8, 8, "C2"), "c3": (4, 32, 1, "C3 (long context, alpha=8)"), "c5": (2, 16, 8, "C5 (DP training, 16-frame clips, global batch 64)")}
CONFIG_NAME = "C2"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"bf16_sustained": float(p["bf16_tflops_sustained"]), "bf16_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's CPU path)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(state):
    """One fwd+bwd of the hot path for ONE clip on the host CPU: oracle restatement of SlowFastLayers (torch CPU fp32,
    what the reference's nn modules dispatch to) + torchvision's own CPU ROIAlign / box head / fastrcnn_loss / mask head /
    mask loss."""
    import torch
    from oracle import slowfast_oracle as so
    sd, feats, slow, pools, head, pred, props_box, props_mask, gt, lab, matched, shapes, box_head, box_pred, box_lab, box_tgt = state
    from torchvision.models.detection.roi_heads import fastrcnn_loss, maskrcnn_loss
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    work = {k: (leaves[k] if k in leaves else v.clone()) for k, v in sd.items()}
    merged = so.temporally_enhance_features(work, slow, feats, True)
    box = pools[0](merged, props_box, shapes)
    mask = pools[1](merged, props_mask, shapes)
    logits = pred(head(mask))
    loss_cls, loss_reg = fastrcnn_loss(*box_pred(box_head(box)), box_lab, box_tgt)
    loss = maskrcnn_loss(logits, props_mask, gt, lab, matched) + loss_cls + loss_reg
    loss.backward()
    for m in (head, pred, box_head, box_pred):
        m.zero_grad()
    return float(loss.detach())


def build_cpu_state():
    import torch
    from torchvision.models.detection.faster_rcnn import FastRCNNPredictor, TwoMLPHead
    from torchvision.models.detection.mask_rcnn import MaskRCNNHeads, MaskRCNNPredictor
    from torchvision.ops import MultiScaleRoIAlign
    from oracle import slowfast_oracle as so
    from sfvos_b200 import workload as wl
    torch.set_num_threads(os.cpu_count() or 1)
    sd = so.init_state_dict(SP, FP, seed=63)
    feats = wl.synthetic_features(1, FP)
    lo = FP // 2 - SP // 2
    slow = [so.slice_window(feats[0], FP // 2, SP)]
    torch.manual_seed(63)
    pools = (MultiScaleRoIAlign(wl.POOL_LEVELS, 7, 2), MultiScaleRoIAlign(wl.POOL_LEVELS, 14, 2))
    head, pred = MaskRCNNHeads(256, (256, 256, 256, 256), 1), MaskRCNNPredictor(256, 256, 2)
    box_head, box_pred = TwoMLPHead(256 * 7 * 7, 1024), FastRCNNPredictor(1024, 2)
    box_lab, box_tgt = wl.synthetic_box_targets(1, K_BOX, K_MASK)
    box = wl.synthetic_rois(1, K_BOX)
    gt = torch.zeros(1, wl.IMAGE_HW[0], wl.IMAGE_HW[1], dtype=torch.uint8)
    gt[0, 200:500, 400:600] = 1
    return (sd, feats, slow, pools, head, pred, box, [box[0][:K_MASK]], [gt], [torch.ones(1, dtype=torch.int64)],
            [torch.zeros(K_MASK, dtype=torch.int64)], [wl.IMAGE_HW], box_head, box_pred, box_lab, box_tgt)


def time_cpu_reference(steps, warmup, threads=None):
    """-> (clip-frames/s, seconds per step): MEDIAN of ``steps`` timed repetitions after ``warmup`` (SURVEY 8(d) protocol)."""
    import torch
    state = build_cpu_state()
    if threads:
        torch.set_num_threads(threads)
    for _ in range(warmup):
        cpu_reference_step(state)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(state)
        times.append(time.perf_counter() - t0)
    sec = sorted(times)[len(times) // 2]
    return FP / sec, sec


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    steps, warmup = max(3, min(steps, 5)), max(1, min(warmup, 2))     # each step is seconds of CPU work; >= 3 timed, median
    value, sec = time_cpu_reference(steps, warmup)
    cores = os.cpu_count() or 1
    sample = f"1 clip (B=1, sp={SP}, fp={FP}, 5 levels, {K_BOX} box + {K_MASK} mask ROIs) fwd+bwd per step; median of {steps} timed steps after {warmup} warm-up"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{CONFIG_NAME}: SlowFast(sp={SP},fp={FP}) + ROIAlign + box head/losses + mask head/loss fwd+bwd (CPU sample: 1 clip/step)"},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from sfvos_b200 import dp, ops, workload as wl

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ops.device_check()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = _peaks()
    micro = max(1, min(args.micro, B_PER_GPU))
    n_micro = (B_PER_GPU + micro - 1) // micro
    assert n_micro * micro == B_PER_GPU, "clips per GPU must be a multiple of --micro"
    feat_dtype = torch.bfloat16 if args.feat_dtype == "bf16" else torch.float32
    step = wl.HotPathStep(SP, FP, micro, K_BOX, K_MASK, device=dev, precision="bf16")
    params = step.parameters()
    # The rank's clips are the B consecutive windows of ONE synthetic sequence of B + fp - 1 frames -- how the reference forms
    # its clips (code/helpers/model.py:318-337: window w = frames [w, w+fp) of the cached backbone features) -- kept as ONE
    # [frames,256,H,W] tensor per level; the clips are views of it.
    n_frames = B_PER_GPU + FP - 1
    seq = wl.synthetic_sequence(n_frames, seed=1234 + 1000 * rank, device=dev, dtype=feat_dtype)
    chunks = lambda sq: [wl.sequence_windows(sq, FP, j * micro, micro) for j in range(n_micro)]
    # gradients live in one flat arena: [roi_heads | slow_fast]; each range is all-reduced as soon as it is complete
    arena = dp.GradArena(step.groups(), dev)
    ops.GRAD_ARENA = arena
    # code/train.py:80.  foreach, not fused: the fused implementation updated the parameters without the version bump that the
    # packed-weight caches of the eager path key on (measured: eager loss 1.1 % off the graph's after a fused step)
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4, foreach=True)
    opt_impl = "foreach"

    def reduce_and_step(roi_work):
        """Tail of the step: the (small) SlowFast range joins the roi_heads range already in flight, 1/world, SGD."""
        if world > 1:
            if roi_work is None:        # unsplit step: the whole arena in one collective
                dist.all_reduce(arena.flat, op=dist.ReduceOp.AVG)
            else:
                sf_work = dist.all_reduce(arena.range("slow_fast"), op=dist.ReduceOp.AVG, async_op=True)
                roi_work.wait(); sf_work.wait()
        opt.step()

    def launch_roi_allreduce():          # ncclAvg: the 1/world scale happens inside the collective, no extra pass over the arena
        return dist.all_reduce(arena.range("roi_heads"), op=dist.ReduceOp.AVG, async_op=True) if world > 1 else None

    def eager_step(sq):
        arena.zero()
        work = [None]
        for j, clips in enumerate(chunks(sq)):
            for p in params:
                p.grad = None
            loss, merged = step.forward(clips)
            last = j == n_micro - 1
            step.backward_split(loss, merged, (lambda: work.__setitem__(0, launch_roi_allreduce())) if last else None)
        reduce_and_step(work[0])
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(3, args.warmup)):
        eager_step(seq)
    assert arena.adopted(), "gradient arena slices were not adopted as .grad"
    # ---- eager pass: K steps with CUDA events around every tensor-core / ROIAlign launch (per-kernel rooflines) ----
    # (single stream for this pass: with the pyramid levels on concurrent streams an event pair around a small level's launch
    # brackets the time it waits for SMs behind level 0's persistent kernels, not the kernel.  The timed region below runs the
    # levels concurrently.)
    prev_streams = os.environ.get("SFVOS_LEVEL_STREAMS")
    os.environ["SFVOS_LEVEL_STREAMS"] = "0"
    eager_step(seq)
    ops.TIMING = []
    l0 = ops.launches()
    ms_eager = timed(lambda: eager_step(seq), args.steps)
    launches = ops.launches() - l0
    timing, ops.TIMING = ops.TIMING, None
    if prev_streams is None:
        os.environ.pop("SFVOS_LEVEL_STREAMS")
    else:
        os.environ["SFVOS_LEVEL_STREAMS"] = prev_streams
    step_peak = torch.cuda.max_memory_allocated(dev)           # features + one eager step
    torch.cuda.empty_cache()

    # ---- timed region: K steps, inputs resident in HBM.  Every micro-batch is replayed from TWO CUDA graphs (forward +
    # roi_heads backward | SlowFast backward) so that the all-reduce of the roi_heads gradient range can be launched in
    # between and run under the second graph; SFVOS_GRAPH=0 or a failed capture falls back to the eager launches above ----
    # Two input slots (device copies of the sequence) so that the e2e leg can fill one while the other is consumed.
    h2d = sum(v.numel() * v.element_size() for v in seq.values())
    slots = [seq]
    free_now = torch.cuda.mem_get_info(dev)[0]
    if free_now > (step_peak - h2d) + 2 * h2d + (6 << 30):
        slots.append(OrderedDict((k, torch.empty_like(v)) for k, v in seq.items()))
    graphs, mode = None, "eager"
    if os.environ.get("SFVOS_GRAPH", "1") != "0":
        try:
            pool = torch.cuda.graph_pool_handle()
            split = world > 1 and os.environ.get("SFVOS_DP_SPLIT", "1") != "0"
            if split:           # two graphs per micro-batch: the roi_heads all-reduce is launched between them
                graphs = [[step.capture_split(clips, zero_arena=(j == 0), pool=pool) for j, clips in enumerate(chunks(sq))] for sq in slots]
            else:               # nothing to overlap (or SFVOS_DP_SPLIT=0): one graph per micro-batch, no join between the backward phases
                graphs = []
                for sq in slots:
                    per = []
                    for j, clips in enumerate(chunks(sq)):
                        g, loss_t = step.capture(clips, warmup=1, pool=pool, zero_arena=(j == 0))
                        per.append((g, None, loss_t))
                    graphs.append(per)
            mode = "cuda_graph"
        except Exception as exc:                        # keep the run valid: report the eager number
            print(f"bench: CUDA-graph capture failed ({exc.__class__.__name__}: {exc}); timing eager launches", file=sys.stderr)
            graphs = None
            for p in params:
                p.grad = None

    def graph_step(slot=0):
        work = None
        for j, (g1, g2, _) in enumerate(graphs[slot]):
            g1.replay()
            if g2 is not None:
                if j == n_micro - 1:
                    work = launch_roi_allreduce()
                g2.replay()
        reduce_and_step(work)
        return graphs[slot][-1][2]

    run_step = (lambda slot=0: graph_step(slot)) if graphs is not None else (lambda slot=0: eager_step(slots[slot]))
    if graphs is not None:
        # the replayed step must be the step: same loss as an eager forward on the same inputs and parameters
        graphs[0][-1][0].replay()
        g_loss = float(graphs[0][-1][2].detach())
        ref_loss = float(step.forward(chunks(seq)[-1])[0].detach())
        assert abs(g_loss - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss)), (g_loss, ref_loss)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(2):
        run_step()
    ms = timed(run_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- rooflines from the per-launch CUDA events of the eager pass ----
    # tensor-core GEMM kernels: work = algorithmic FLOPs of the launch (no padding / halo / zero-tap FLOPs);
    # ROIAlign: work = algorithmic bytes of the launch (workload.roi_align_bytes).
    roi_bytes = {"roi_align_fwd_p7": wl.roi_align_bytes(step.box_props, 7, 4, 2),
                 "roi_align_fwd_p14": wl.roi_align_bytes(step.mask_props, 14, 4, 2),
                 "roi_align_bwd_p7": wl.roi_align_bytes(step.box_props, 7, 4, 2, backward=True),
                 "roi_align_bwd_p14": wl.roi_align_bytes(step.mask_props, 14, 4, 2, backward=True)}
    fam = {}
    for name, flops, a, b in timing:
        d = fam.setdefault(name, [0.0, 0.0, 0])
        d[0] += roi_bytes.get(name, flops); d[1] += a.elapsed_time(b); d[2] += 1
    tensor = {k: v for k, v in fam.items() if k not in roi_bytes}
    kernels = {k: {"tflops": round(v[0] / (v[1] * 1e-3) / 1e12, 1) if v[1] else None, "ms_per_step": round(v[1] / args.steps, 3),
                   "launches_per_step": v[2] // args.steps, "share_of_step": round(v[1] / args.steps / ms_eager, 3)}
               for k, v in sorted(tensor.items(), key=lambda kv: -kv[1][1])}
    tot_f = sum(v[0] for v in tensor.values()); tot_ms = sum(v[1] for v in tensor.values())
    achieved_all = tot_f / (tot_ms * 1e-3) / 1e12 if tot_ms else 0.0
    dom = max(tensor, key=lambda k: tensor[k][1]) if tensor else None          # the kernel with the largest share of the step
    dom_f, dom_ms, dom_n = tensor[dom] if dom else (0.0, 0.0, 0)
    achieved = dom_f / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0
    traffic, traffic_src, dram, tj = None, None, {}, {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"{dom}_dram_bytes_per_launch")
            traffic_src = tj.get("source")
            dram = tj.get("roi_align_dram", {})
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": f"{dom}_kernel ({dom_n // max(1, args.steps)} launches/step: the dominant kernel by time)",
                "achieved": round(achieved, 1), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["bf16_sustained"], 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["source"] + ", sustained bf16 (cuBLAS 8192^3 back to back for 4 s: the kernel is timed inside the step)",
                "frac_of_burst_peak": round(achieved / peaks["bf16_burst"], 4), "burst_peak": peaks["bf16_burst"],
                "flops_per_launch": dom_f / max(1, dom_n), "us_per_launch": round(dom_ms * 1e3 / max(1, dom_n), 1),
                "all_tensor_kernels": {"achieved": round(achieved_all, 1), "frac": round(achieved_all / peaks["bf16_sustained"], 4),
                                       "flops_per_step": tot_f / args.steps, "kernel_ms_per_step": round(tot_ms / args.steps, 3)},
                "per_kernel": kernels}
    # ROIAlign, "achieved HBM GB/s from ncu" (north star): bytes = what ncu's DRAM counters saw for one launch of that kind on
    # this ROI set (dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json <- profiles/ncu_kernels_r2.md), time = this
    # run's CUDA events.  The SURVEY 8(d) byte model (unique footprint per ROI, no credit for cache hits) is kept beside it: it
    # over-credits by 2-3x because the footprints of neighbouring ROIs and bins overlap in L1 / L2.
    # The backward kernels are bound by the L2's vector-reduction throughput, not by HBM: their work is red.global.add.v4.f32
    # into f32 maps, and ncu shows 83-89 % of what the L2 delivers on a streaming reduction probe.  So next to the DRAM figure
    # every launch kind carries an "l2" record: bytes = ncu's L2 sector counts for that launch (reads for the forward gather,
    # reductions for the backward scatter; profiles/traffic.json), peak = this GPU's L2 measured LIVE with sfvos_probe_l2 on an
    # L2-resident 48 MB buffer (16-byte loads / red.v4.f32), time = this run's CUDA events.
    l2_peak, l2_bytes = {}, {}
    try:
        from sfvos_b200._lib import call as _call
        pb = torch.zeros((48 << 20) // 4, dtype=torch.float32, device=dev)
        sink = torch.zeros(1, device=dev)
        for kind, key in ((0, "read"), (1, "red")):
            for _ in range(2):
                _call("sfvos_probe_l2", kind, ops._p(pb), 48 << 20, 20, ops._p(sink), ops.stream())
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); _call("sfvos_probe_l2", kind, ops._p(pb), 48 << 20, 20, ops._p(sink), ops.stream()); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            l2_peak[key] = (48 << 20) * 20 / (sorted(ts)[3] * 1e-3) / 1e9
        del pb
        l2_bytes = tj.get("roi_align_l2", {})
    except Exception as exc:
        print(f"bench: L2 probe skipped ({exc.__class__.__name__}: {exc})", file=sys.stderr)
    roi = {}
    for k, v in fam.items():
        if k.startswith("roi_align_") and v[1]:
            us = v[1] * 1e3 / v[2]
            rec = {"us_per_launch": round(us, 1), "algorithmic_model_gbs": round(v[0] / v[2] / us / 1e3, 1)}
            if k in dram:
                rec.update({"bound": "hbm", "achieved": round(dram[k] / us / 1e3, 1), "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": round(dram[k] / us / 1e3 / peaks["hbm"], 4), "ncu_dram_bytes_per_launch": dram[k]})
            if k in l2_bytes and l2_bytes[k]["kind"] in l2_peak:
                kind = l2_bytes[k]["kind"]
                gbs = l2_bytes[k]["bytes"] / us / 1e3
                rec["l2"] = {"bound": "l2 " + ("vector reductions (red.global.add.v4.f32)" if kind == "red" else "-> SM reads"),
                             "achieved": round(gbs, 1), "peak": round(l2_peak[kind], 1), "unit": "GB/s", "frac": round(gbs / l2_peak[kind], 4),
                             "ncu_l2_bytes_per_launch": l2_bytes[k]["bytes"], "peak_source": "sfvos_probe_l2, 48 MB L2-resident buffer, this run"}
            roi[k] = rec
    for tag in ("fwd", "bwd", ""):
        ks = [k for k in fam if k.startswith("roi_align_" + tag) and fam[k][1]]
        t = sum(fam[k][1] for k in ks)
        if not t:
            continue
        model_b = sum(fam[k][0] for k in ks)
        rec = {"ms_per_step": round(t / args.steps, 3), "algorithmic_model_gbs": round(model_b / (t * 1e-3) / 1e9, 1),
               "algorithmic_model_frac": round(model_b / (t * 1e-3) / 1e9 / peaks["hbm"], 4)}
        if all(k in dram for k in ks):
            dram_b = sum(dram[k] * fam[k][2] for k in ks)
            rec.update({"bound": "hbm", "achieved": round(dram_b / (t * 1e-3) / 1e9, 1), "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": round(dram_b / (t * 1e-3) / 1e9 / peaks["hbm"], 4), "basis": "ncu DRAM bytes per launch x launches / live launch time"})
        roi["roi_align_" + (tag or "fwd_bwd")] = rec
    roofline["roi_align"] = roi

    # ---- end to end through the public API from pinned host buffers ----
    # Every step's inputs come from pinned HOST memory and every step's loss goes back to the host, all inside the timed
    # region.  What crosses PCIe is the sequence's features, every frame ONCE (the windows are views; the reference's
    # features_cache does the same on its side, model.py:191-227), in the feature dtype.  The copies are software-pipelined
    # like a data loader: step i+1's sequence goes into the other device slot on a copy stream while step i computes.
    e2e_steps = max(2, min(args.steps, 20))      # the same K steps as the resident-input region (the first step's whole-sequence copy is not overlapped)
    host = wl.synthetic_sequence(n_frames, seed=1234 + 1000 * rank, device="cpu", dtype=feat_dtype, pin=True)
    ns = len(slots)
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(ns)]
    freed = [torch.cuda.Event() for _ in range(ns)]
    loss_host = torch.zeros(e2e_steps + 1, dtype=torch.float32).pin_memory()

    # Streaming: consecutive steps work on consecutive stretches of ONE long sequence (step i: windows [iB, (i+1)B) = frames
    # [iB, (i+1)B + fp - 1)), so fp - 1 of a step's frames are already on the device from the previous step -- the reference's
    # features_cache keeps them there too (model.py:191-227) -- and only the B NEW frames cross PCIe: every frame is copied from
    # the host exactly once.  The carried-over frames move slot to slot with a device copy.
    n_old, n_new = FP - 1, B_PER_GPU
    h2d_stream = sum(v[n_old:].numel() * v.element_size() for v in host.values())

    def issue_copy(slot, prev, streaming):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for k, v in host.items():
                if streaming:
                    if prev != slot:
                        copy_stream.wait_event(ready[prev])
                    carried = slots[prev][k][n_new:n_new + n_old]
                    if prev == slot and n_old > n_new:
                        carried = carried.clone()                                                            # ranges overlap in place
                    slots[slot][k][:n_old].copy_(carried, non_blocking=True)                                  # carried over (device)
                    slots[slot][k][n_old:].copy_(v[n_old:], non_blocking=True)                                # new frames (PCIe)
                else:
                    slots[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_run(n, streaming=True):
        cur = torch.cuda.current_stream()
        for ev in freed:
            ev.record(cur)
        issue_copy(0, 0, False)                                  # the first stretch of the sequence: all of its frames
        for i in range(n):
            if ns > 1 and i + 1 < n:
                issue_copy((i + 1) % ns, i % ns, streaming)
            cur.wait_event(ready[i % ns])
            loss = run_step(i % ns)
            freed[i % ns].record(cur)
            if ns == 1 and i + 1 < n:
                issue_copy(0, 0, streaming)
            loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)     # D2H of the step's result

    def e2e_time(streaming):
        e2e_run(2, streaming)
        sync_all()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_run(e2e_steps, streaming)
        t1.record()
        sync_all()
        assert all(math.isfinite(float(x)) for x in loss_host[:e2e_steps])
        t = torch.tensor([t0.elapsed_time(t1) / e2e_steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_ms = e2e_time(True)
    e2e_full_ms = e2e_time(False)
    e2e = {"value": round(world * B_PER_GPU * FP / (e2e_ms * 1e-3), 2), "unit": UNIT, "h2d_bytes_per_step": h2d_stream * world,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
           "h2d_gbs_per_gpu": round(h2d_stream / (e2e_ms * 1e-3) / 1e9, 1),
           "note": (f"streaming through one long sequence: every step copies its {n_new} NEW frames of {args.feat_dtype} FPN features from pinned "
                    f"host memory (each frame crosses PCIe once; the {n_old} frames it shares with the previous step stay on the device, as the "
                    f"reference's features_cache keeps them), " +
                    ("double-buffered on a copy stream so step i+1's H2D overlaps step i's compute" if ns > 1 else
                     "single device buffer (a second one does not fit next to this configuration's step), so H2D and compute alternate")),
           "every_step_a_new_sequence": {"value": round(world * B_PER_GPU * FP / (e2e_full_ms * 1e-3), 2), "ms_per_step": round(e2e_full_ms, 3),
                                         "h2d_bytes_per_step": h2d * world, "h2d_gbs_per_gpu": round(h2d / (e2e_full_ms * 1e-3) / 1e9, 1),
                                         "note": f"all {n_frames} frames of the step copied every step (no frame shared with the previous step)"}}

    vs_library = None
    if rank == 0 and world == 1 and not args.no_lib:
        graphs = None                                           # release the graphs' private pool first
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_library
            vs_library = bench_library.run(SP, FP, min(B_PER_GPU, 8), K_BOX, K_MASK, dev="cuda")
        except Exception as exc:
            vs_library = {"error": f"{exc.__class__.__name__}: {exc}"}
    roofline["vs_library"] = vs_library

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            v, sec = time_cpu_reference(3, 1)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"1 clip of the same workload (B=1) fwd+bwd on the host CPU, 1 warm-up + 3 timed steps, median ({sec:.1f} s/step)"}
            if not args.no_cpu_1t:
                v1, sec1 = time_cpu_reference(1, 0, threads=1)
                cpu["one_thread"] = {"value": round(v1, 4), "unit": UNIT, "cores": 1, "sample": f"the same clip on ONE thread, 1 timed step ({sec1:.1f} s)"}
        conv_f, mask_f = step.flops_per_step()
        conv_f, mask_f = conv_f * n_micro, mask_f * n_micro
        line = {"metric": METRIC, "value": round(world * B_PER_GPU * FP / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{CONFIG_NAME}: SlowFast temporal module (sp={SP}, fp={FP}) + multi-level ROIAlign + box head (fc6/fc7/predictor/fastrcnn_loss) + mask head/predictor/loss, fwd+bwd, gradient all-reduce (N>1), SGD step",
                           "clips_per_gpu": B_PER_GPU, "frames_per_clip": FP, "levels": "192x336,96x168,48x84,24x42,12x21 x256ch",
                           "clips": f"the {B_PER_GPU} windows of one synthetic {n_frames}-frame sequence per GPU (views; code/helpers/model.py:318-337), features {args.feat_dtype} [frames,256,H,W]",
                           "micro_batches": n_micro, "clips_per_micro_batch": micro,
                           "rois_per_clip": {"box": K_BOX, "mask": K_MASK},
                           "parallelism": f"dp{world} by clip; gradients in one flat arena, roi_heads range all-reduced under the SlowFast backward, SlowFast range at the end",
                           "optimizer": f"torch.optim.SGD(lr 1e-3, momentum 0.9, weight_decay 1e-4, {opt_impl}) inside the timed region",
                           "l2": f"inputs ({h2d / 1e9:.2f} GB of features) and activations (> 10 GB per step) far exceed the 126 MB L2; no explicit flush",
                           "launch": mode, "eager_ms_per_step": round(ms_eager, 3),
                           "streams": "timed region: pyramid levels 1..4 on side streams inside the graphs; per-kernel event pass: eager, one stream"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "model_tflops": round((conv_f + mask_f) * world / (ms * 1e-3) / 1e12, 1)}
        print(json.dumps(line), flush=True)
    ops.GRAD_ARENA = None
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cpu-1t", action="store_true", help="skip the single-thread figure of the cpu_baseline leg")
    ap.add_argument("--no-lib", action="store_true", help="skip the same-box library baselines (cuDNN / torchvision, roofline.vs_library)")
    ap.add_argument("--feat-dtype", default="bf16", choices=["bf16", "fp32"], help="dtype of the FPN features handed to the module (host and device)")
    ap.add_argument("--micro", type=int, default=8, help="clips per micro-batch (forward+backward call); a rank's clips run as ceil(clips/micro) calls whose gradients accumulate")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: c2 (default, the configuration the metric is quoted on), c3 long context, c5 DP training")
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: the config's)")
    args = ap.parse_args()
    global SP, FP, B_PER_GPU, CONFIG_NAME
    SP, FP, B_PER_GPU, CONFIG_NAME = CONFIGS[args.config]
    if args.config == "c5":
        B_PER_GPU = max(8, 64 // max(1, args.gpus))      # global batch 64 over the ranks, 8 clips per micro-batch (one GPU alone: 8)
    if args.clips:
        B_PER_GPU = args.clips
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--config", args.config, "--micro", str(args.micro), "--feat-dtype", args.feat_dtype] + (
                   ["--no-cpu"] if args.no_cpu else []) + (["--clips", str(args.clips)] if args.clips else [])
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
