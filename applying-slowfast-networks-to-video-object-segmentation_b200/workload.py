"""Synthetic DAVIS-shaped hot-path step: SlowFast temporal module on B clips (all 5 FPN levels) ->
multi-level ROIAlign for the box (7x7) and mask (14x14) branches -> box head (fc6/fc7) -> box predictor ->
fastrcnn_loss, and mask head -> mask predictor -> mask loss, forward + backward: what ``roi_heads`` computes in a
training step of the reference (code/helpers/model.py:340-346) once the proposals are sampled.  Used by bench.py (the
measured step) and the data-parallel trainer; everything numeric runs in libsfvos.so."""
import math
from collections import OrderedDict

import torch

from .roi_heads import (FastRCNNPredictor, MaskRCNNHeads, MaskRCNNPredictor, MultiScaleRoIAlign, TwoMLPHead,
                        fastrcnn_loss, maskrcnn_loss, pool_pair)
from .slowfast import SlowFastLayers

# 480x854 DAVIS frame -> GeneralizedRCNNTransform -> 749x1333 -> padded 768x1344 -> FPN strides 4..64
IMAGE_HW = (749, 1333)
LEVELS = OrderedDict([("0", (192, 336)), ("1", (96, 168)), ("2", (48, 84)), ("3", (24, 42)), ("pool", (12, 21))])
POOL_LEVELS = ["0", "1", "2", "3"]


def synthetic_rois(n_clips, k_per_clip, image_hw=IMAGE_HW, seed=4321, lo=16.0, hi=700.0):
    """Log-uniform side 16..700 px, aspect U(0.5,2), uniform centre, clipped; all 4 levels are hit."""
    g = torch.Generator().manual_seed(seed)
    h_img, w_img = image_hw
    out = []
    for _ in range(n_clips):
        s = torch.exp(torch.rand(k_per_clip, generator=g) * (math.log(hi) - math.log(lo)) + math.log(lo))
        a = torch.rand(k_per_clip, generator=g) * 1.5 + 0.5
        bw, bh = s * torch.sqrt(a), s / torch.sqrt(a)
        cx = torch.rand(k_per_clip, generator=g) * w_img
        cy = torch.rand(k_per_clip, generator=g) * h_img
        x1 = (cx - bw / 2).clamp(0, w_img - 1)
        y1 = (cy - bh / 2).clamp(0, h_img - 1)
        x2 = torch.maximum((cx + bw / 2).clamp(0, w_img), x1 + 1)
        y2 = torch.maximum((cy + bh / 2).clamp(0, h_img), y1 + 1)
        out.append(torch.stack([x1, y1, x2, y2], dim=1))
    return out


def synthetic_features(n_clips, fp, levels=LEVELS, seed=1234, device="cpu", pin=False):
    """Per clip an OrderedDict {level: f32 [fp,256,H,W]} of seeded unit-variance features."""
    clips = []
    for b in range(n_clips):
        d = OrderedDict()
        for i, (k, (h, w)) in enumerate(levels.items()):
            if device == "cpu":
                g = torch.Generator().manual_seed(seed + 100 * b + i)
                t = torch.randn(fp, 256, h, w, generator=g)
                if pin:
                    t = t.pin_memory()
            else:
                g = torch.Generator(device=device).manual_seed(seed + 100 * b + i)
                t = torch.randn(fp, 256, h, w, generator=g, device=device)
            d[k] = t
        clips.append(d)
    return clips


def conv_flops(sp, fp, levels=LEVELS, fwd_only=False):
    """Algorithmic FLOPs per clip of the 8 Conv3d layers: fwd, and fwd + wgrad + dgrad (no dgrad for
    the two first-layer convs, whose inputs carry no gradient)."""
    def ks(p):
        d, r = divmod(p, 3)
        return (d, d + 1, d + 1) if r == 0 else (d + 1, d + 1, d + 1) if r == 1 else (d + 1, d + 1, d + 2)
    s, f = ks(sp), ks(fp)
    ts = [sp - s[0] + 1]; ts.append(ts[0] - s[1] + 1); ts.append(ts[1] - s[2] + 1)
    tf = [fp - f[0] + 1]; tf.append(tf[0] - f[1] + 1); tf.append(tf[1] - f[2] + 1)
    kl1, kl2 = tf[0] - ts[0] + 1, tf[1] - ts[1] + 1
    hw = sum(h * w for h, w in levels.values())
    layers = [  # (T_out, Cout, Cin, taps, first)
        (ts[0], 192, 256, s[0] * 9, True), (tf[0], 32, 256, f[0] * 9, True), (ts[0], 64, 32, kl1, False),
        (ts[1], 192, 256, s[1] * 9, False), (tf[1], 32, 32, f[1] * 9, False), (ts[1], 64, 32, kl2, False),
        (ts[2], 224, 256, s[2] * 9, False), (tf[2], 32, 32, f[2] * 9, False)]
    fwd = sum(2.0 * t * hw * co * ci * taps for t, co, ci, taps, _ in layers)
    first = sum(2.0 * t * hw * co * ci * taps for t, co, ci, taps, fl in layers if fl)
    return fwd if fwd_only else 3.0 * fwd - first


def roi_align_bytes(boxes, P, e_in, e_out, backward=False, C=256, image_hw=IMAGE_HW, levels=LEVELS):
    """Algorithmic bytes of one multi-level ROIAlign launch: the output (or its gradient), the UNIQUE
    feature footprint of every ROI at its level -- C * min(H_l, ceil(h)+1) * min(W_l, ceil(w)+1) elements, no credit
    for re-reads -- and 20 bytes of ROI record.  Backward reads the output gradient and read-modify-writes the f32
    footprint (2x); the zero-fill of the gradient maps is a separate kernel and not counted.  boxes: list of [K_i,4]
    tensors in image pixels (any device)."""
    names = [k for k in levels if k in POOL_LEVELS]
    shapes = [levels[k] for k in names]
    scales = [2.0 ** float(round(math.log2(h / float(image_hw[0])))) for h, _ in shapes]
    k_min, k_max = int(-math.log2(scales[0])), int(-math.log2(scales[-1]))
    b = torch.cat([x.detach().float().cpu() for x in boxes])
    K = b.shape[0]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lvl = torch.floor(4.0 + torch.log2(torch.sqrt(area) / 224.0) + 1e-6).clamp(k_min, k_max).long() - k_min
    sc = torch.tensor(scales)[lvl]
    Hl = torch.tensor([float(s[0]) for s in shapes])[lvl]
    Wl = torch.tensor([float(s[1]) for s in shapes])[lvl]
    h = torch.clamp((b[:, 3] - b[:, 1]) * sc, min=1.0)
    w = torch.clamp((b[:, 2] - b[:, 0]) * sc, min=1.0)
    cells = (torch.minimum(Hl, torch.ceil(h) + 1) * torch.minimum(Wl, torch.ceil(w) + 1)).sum().item()
    out_bytes = K * C * P * P * e_out
    if backward:
        return out_bytes + 2.0 * cells * C * 4 + K * 20
    return out_bytes + cells * C * e_in + K * 20


MASK_HEAD_FLOPS_PER_ROI = 2.0 * 196 * 256 * 2304 * 4 + 2.0 * 196 * 256 * 1024 + 2.0 * 784 * 2 * 256   # fwd
BOX_HEAD_FLOPS_PER_ROI = 2.0 * (256 * 49 * 1024 + 1024 * 1024 + 1024 * 10)                            # fwd: fc6, fc7, cls+bbox


def synthetic_box_targets(n_clips, k_box, k_pos, seed=977):
    """Per clip: labels int64 [k_box] (the first k_pos ROIs are foreground = the mask branch's ROIs, the rest background,
    1:3 like torchvision's sampler) and regression targets f32 [k_box,4] ~ 0.5 * N(0,1) (both sides of beta = 1/9)."""
    g = torch.Generator().manual_seed(seed)
    labels, targets = [], []
    for _ in range(n_clips):
        lab = torch.zeros(k_box, dtype=torch.int64)
        lab[:k_pos] = 1
        labels.append(lab)
        targets.append(0.5 * torch.randn(k_box, 4, generator=g))
    return labels, targets


class HotPathStep:
    """One forward+backward of the hot path on a batch of clips.  ``features``: list (len B) of {level: [fp,256,H,W]}
    CUDA f32 tensors (e.g. the frozen backbone's output windows)."""

    def __init__(self, sp=1, fp=8, n_clips=8, k_box=512, k_mask=128, levels=LEVELS, device="cuda", precision="bf16", seed=63):
        self.sp, self.fp, self.B, self.levels = sp, fp, n_clips, levels
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.slow_fast = SlowFastLayers(256, self.device, sp, fp).to(self.device).train()
        self.slow_fast.precision = precision
        self.mask_head = MaskRCNNHeads(256, (256, 256, 256, 256), 1).to(self.device)
        self.mask_predictor = MaskRCNNPredictor(256, 256, 2).to(self.device)
        self.mask_head.precision = self.mask_predictor.precision = precision
        names = [k for k in levels if k in POOL_LEVELS]
        self.box_head = TwoMLPHead(256 * 7 * 7, 1024).to(self.device)
        self.box_predictor = FastRCNNPredictor(1024, 2).to(self.device)
        self.box_head.precision = self.box_predictor.precision = precision
        self.box_roi_pool = MultiScaleRoIAlign(names, 7, 2, out_layout="nchw", precision=precision, out_dtype="act")
        self.mask_roi_pool = MultiScaleRoIAlign(names, 14, 2, out_layout="nhwc", precision=precision)
        self.image_shapes = [IMAGE_HW] * n_clips
        box = synthetic_rois(n_clips, k_box, seed=4321)
        self.box_props = [b.to(self.device) for b in box]
        self.mask_props = [b[:k_mask].to(self.device) for b in box]
        gt = torch.zeros(1, IMAGE_HW[0], IMAGE_HW[1], dtype=torch.uint8)
        gt[0, 200:500, 400:600] = 1                                     # one 300x200 object per clip
        self.gt_masks = [gt.to(self.device) for _ in range(n_clips)]
        self.gt_labels = [torch.ones(1, dtype=torch.int64, device=self.device) for _ in range(n_clips)]
        self.matched = [torch.zeros(k_mask, dtype=torch.int64, device=self.device) for _ in range(n_clips)]
        lab, tgt = synthetic_box_targets(n_clips, k_box, k_mask)
        self.box_labels = [t.to(self.device) for t in lab]
        self.box_targets = [t.to(self.device) for t in tgt]
        self.k_box, self.k_mask = k_box, k_mask

    def modules(self):
        """(name, module) in the registration order of the reference's roi_heads, after slow_fast."""
        return [("slow_fast", self.slow_fast), ("box_head", self.box_head), ("box_predictor", self.box_predictor),
                ("mask_head", self.mask_head), ("mask_predictor", self.mask_predictor)]

    def parameters(self):
        return [p for _, m in self.modules() for p in m.parameters()]

    def state_dict(self):
        sd = OrderedDict(("slow_fast." + k, v) for k, v in self.slow_fast.state_dict().items())
        sd.update(("box_head." + k, v) for k, v in self.box_head.state_dict().items())
        sd.update(("box_predictor." + k, v) for k, v in self.box_predictor.state_dict().items())
        sd.update(("mask_head." + k, v) for k, v in self.mask_head.state_dict().items())
        sd.update(("mask_predictor." + k, v) for k, v in self.mask_predictor.state_dict().items())
        return sd

    def forward(self, features):
        lo = self.fp // 2 - self.sp // 2
        slow = [OrderedDict((k, v[lo:lo + self.sp]) for k, v in f.items()) for f in features]
        merged = self.slow_fast.temporally_enhance_features(slow, features)
        # both poolings in one autograd node, as RoIHeads.forward does in training (one set of gradient maps)
        box_feats, mask_feats = pool_pair(self.box_roi_pool, self.mask_roi_pool, merged, self.box_props, self.mask_props,
                                          self.image_shapes)
        class_logits, box_regression = self.box_predictor(self.box_head(box_feats))
        loss_cls, loss_reg = fastrcnn_loss(class_logits, box_regression, self.box_labels, self.box_targets)
        logits = self.mask_predictor(self.mask_head(mask_feats))
        loss_mask = maskrcnn_loss(logits, self.mask_props, self.gt_masks, self.gt_labels, self.matched)
        return loss_cls + loss_reg + loss_mask, merged

    def step(self, features, zero_grad=True):
        loss, _ = self.forward(features)
        loss.backward()
        if zero_grad:
            for p in self.parameters():
                p.grad = None
        return loss

    # ---- data-parallel step pieces ----------------------------------------------------------------------------------------
    def groups(self):
        """Parameter groups in the order the backward pass COMPLETES them: every roi_heads gradient (87 % of the bytes) is final
        once the ROI pooling's backward has run, before the SlowFast module's backward starts."""
        roi = [p for name, m in self.modules() if name != "slow_fast" for p in m.parameters()]
        return [("roi_heads", roi), ("slow_fast", list(self.slow_fast.parameters()))]

    def backward_split(self, loss, merged, on_roi_done=None):
        """``loss.backward()`` in two phases with a callback in between: (1) the box / mask branches and the ROI pooling --
        afterwards every roi_heads gradient is complete (``on_roi_done()`` may launch their all-reduce) -- (2) the SlowFast
        module.  Gradients ACCUMULATE into p.grad like backward() does; with a gradient arena (ops.GRAD_ARENA) set p.grad = None
        first so the arena slices are adopted instead of added to themselves."""
        roi = [p for p in self.groups()[0][1] if p.requires_grad]
        outs = [v for v in merged.values() if v.requires_grad]
        grads = torch.autograd.grad(loss, roi + outs, allow_unused=True)
        for p, g in zip(roi, grads):
            if g is not None:
                p.grad = g if p.grad is None else p.grad.add_(g)
        if on_roi_done is not None:
            on_roi_done()
        live = [(o, g) for o, g in zip(outs, grads[len(roi):]) if g is not None]
        torch.autograd.backward([o for o, _ in live], [g for _, g in live])

    def capture(self, features, warmup=2, pool=None, zero_arena=False):
        """Capture one forward+backward on ``features`` (static device buffers) into a CUDA graph: ~430 kernel launches
        become one graph launch, which removes the launch gaps between the (many short) kernels of the small pyramid
        levels.  Returns (graph, loss): ``graph.replay()`` recomputes ``loss`` and every ``p.grad`` in place from the
        current contents of ``features`` and the current parameter values (weight packing is part of the graph).
        With a gradient arena (ops.GRAD_ARENA) the gradients are its slices; ``zero_arena`` puts its clear at the head."""
        from . import ops
        params = self.parameters()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(features)
        torch.cuda.current_stream().wait_stream(side)
        for p in params:
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        try:    # the parameters' AccumulateGrad nodes were created on another stream than the capture stream: expected here
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        with torch.cuda.graph(graph, pool=pool):
            if zero_arena and ops.GRAD_ARENA is not None:
                ops.GRAD_ARENA.zero()
            loss, _ = self.forward(features)
            loss.backward()
        return graph, loss

    def capture_split(self, features, zero_arena=False, warmup=1, pool=None):
        """The same step as TWO graphs split where ``backward_split`` calls back: (forward + roi_heads backward, SlowFast
        backward).  A data-parallel trainer replays the first, launches the all-reduce of the roi_heads gradient range, and
        replays the second underneath it.  Needs ``ops.GRAD_ARENA`` (the gradients are static arena slices that successive
        micro-batch graphs accumulate into; ``zero_arena`` puts the arena's clear at the head of the first graph).
        Returns (graph1, graph2, loss).  Drop every reference to earlier losses / outputs of this step first: a live autograd
        graph keeps the parameters' AccumulateGrad nodes of the stream it ran on, and a capture cannot wait on the legacy stream."""
        from . import ops
        arena = ops.GRAD_ARENA
        assert arena is not None, "capture_split needs a gradient arena"
        params = self.parameters()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in params:
                    p.grad = None
                loss, merged = self.forward(features)
                self.backward_split(loss, merged)
        torch.cuda.current_stream().wait_stream(side)
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        for p in params:
            p.grad = None
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        pool = pool if pool is not None else torch.cuda.graph_pool_handle()
        box = {}
        with torch.cuda.graph(g1, pool=pool):
            if zero_arena:
                arena.zero()
            loss, merged = self.forward(features)
            roi = [p for p in self.groups()[0][1] if p.requires_grad]
            outs = [v for v in merged.values() if v.requires_grad]
            grads = torch.autograd.grad(loss, roi + outs, allow_unused=True)
            for p, g in zip(roi, grads):
                if g is not None:
                    p.grad = g
            box["live"] = [(o, g) for o, g in zip(outs, grads[len(roi):]) if g is not None]
        with torch.cuda.graph(g2, pool=pool):
            torch.autograd.backward([o for o, _ in box["live"]], [g for _, g in box["live"]])
        assert arena.adopted(), "autograd did not adopt the arena slices as .grad"
        return g1, g2, loss

    def flops_per_step(self):
        conv = conv_flops(self.sp, self.fp, self.levels) * self.B
        mask = 3.0 * MASK_HEAD_FLOPS_PER_ROI * self.k_mask * self.B
        box = 3.0 * BOX_HEAD_FLOPS_PER_ROI * self.k_box * self.B
        return conv, mask + box


def flat_grads(params):
    """One contiguous f32 bucket holding every parameter gradient (a COPY; the data-parallel step uses dp.GradArena, where the
    gradients are written into the flat buffer in the first place)."""
    gs = [p.grad for p in params if p.grad is not None]
    return torch.cat([g.reshape(-1) for g in gs]) if gs else None


def synthetic_sequence(n_frames, levels=LEVELS, seed=1234, device="cpu", dtype=torch.float32, pin=False):
    """OrderedDict {level: [n_frames,256,H,W]} of seeded unit-variance features: the backbone output of ONE sequence."""
    d = OrderedDict()
    for i, (k, (h, w)) in enumerate(levels.items()):
        if device == "cpu":
            g = torch.Generator().manual_seed(seed + i)
            t = torch.randn(n_frames, 256, h, w, generator=g).to(dtype)
            if pin:
                t = t.pin_memory()
        else:
            g = torch.Generator(device=device).manual_seed(seed + i)
            t = torch.randn(n_frames, 256, h, w, generator=g, device=device).to(dtype)
        d[k] = t
    return d


def sequence_windows(seq, fp, first=0, count=None):
    """The reference's clips (code/helpers/model.py:318-337): window w = frames [w, w+fp) of the sequence's features, as VIEWS
    (consecutive windows share fp-1 frames; nothing is copied)."""
    n = next(iter(seq.values())).shape[0] - fp + 1
    count = n - first if count is None else count
    return [OrderedDict((k, v[w:w + fp]) for k, v in seq.items()) for w in range(first, first + count)]
