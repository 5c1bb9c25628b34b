"""SlowFastLayers -- drop-in for the reference module (code/helpers/model.py:30-165) running on libsfvos.so.

Same constructor arguments, attribute names, parameter registration order and state_dict keys (the nn.Conv3d /
nn.BatchNorm3d submodules are kept as *parameter containers* only -- their forward is never called; all
arithmetic goes through the C ABI in include/sfvos.h).  Differences that are deliberate:
  * activations are channels-last (NDHWC) bf16 internally (fp32 in validation mode, ``precision="fp32"`` or
    SFVOS_PRECISION=fp32); conv outputs are written straight into channel slices of the concat buffers, so the
    torch.cat / stack / transpose copies of model.py:115,157-158,162 do not exist
  * returned feature maps have the reference's shapes and dtype (fp32 [B,256,H,W]) but channels_last strides
  * there is no CPU path: inputs are moved to ``self.device`` (model.py:157-158 does the same) and the call fails
    loudly if that is not an sm_100 GPU.
"""
import os
from collections import OrderedDict

import torch
from torch import nn

from . import _lib, ops
from ._lib import BF16, F32
from .ops import Act

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class _Spec:
    __slots__ = ("conv", "bn", "cin", "cout", "kt", "khw", "pad", "relu")

    def __init__(self, conv, bn, cin, cout, kt, khw, relu):
        self.conv, self.bn, self.cin, self.cout, self.kt, self.khw, self.relu = conv, bn, cin, cout, kt, khw, relu
        self.pad = 1 if khw == 3 else 0

    @property
    def k(self):
        return (self.kt, self.khw, self.khw)


def _round_up(x, m):
    return (x + m - 1) // m * m


class SlowFastLayers(nn.Module):
    def __init__(self, input_size, device, slow_pathway_size, fast_pathway_size):
        super().__init__()
        self.device = device
        self.slow_pathway_size = slow_pathway_size
        self.fast_pathway_size = fast_pathway_size

        ks1, ks2, ks3 = self._calc_kernel_sizes(slow_pathway_size)
        kf1, kf2, kf3 = self._calc_kernel_sizes(fast_pathway_size)
        kl1, slow_out1, fast_out1 = self._calc_fuse_kernel_size(slow_pathway_size, ks1, fast_pathway_size, kf1)
        kl2, _, _ = self._calc_fuse_kernel_size(slow_out1, ks2, fast_out1, kf2)

        # registration order == reference (model.py:47-67): it fixes state_dict order and SGD state indices
        self.fast_conv1, self.bn_f1 = self._init_conv_and_bn(kf1, input_size, 32)
        self.slow_conv1, self.bn_s1 = self._init_conv_and_bn(ks1, input_size, 192)
        self.fast_conv2, self.bn_f2 = self._init_conv_and_bn(kf2, 32, 32)
        self.slow_conv2, self.bn_s2 = self._init_conv_and_bn(ks2, 256, 192)
        self.fast_conv3, self.bn_f3 = self._init_conv_and_bn(kf3, 32, 32)
        self.slow_conv3, self.bn_s3 = self._init_conv_and_bn(ks3, 256, 224)
        self.conv_f2s1, self.bn_f2s1 = self._init_fuse_and_bn(kl1)
        self.conv_f2s2, self.bn_f2s2 = self._init_fuse_and_bn(kl2)
        self.relu = nn.ReLU(inplace=True)

        self.input_size = input_size
        self.precision = os.environ.get("SFVOS_PRECISION", "bf16")
        self._specs = OrderedDict((s.conv, s) for s in [
            _Spec("fast_conv1", "bn_f1", input_size, 32, kf1, 3, True),
            _Spec("slow_conv1", "bn_s1", input_size, 192, ks1, 3, True),
            _Spec("fast_conv2", "bn_f2", 32, 32, kf2, 3, True),
            _Spec("slow_conv2", "bn_s2", 256, 192, ks2, 3, True),
            _Spec("fast_conv3", "bn_f3", 32, 32, kf3, 3, False),
            _Spec("slow_conv3", "bn_s3", 256, 224, ks3, 3, False),
            _Spec("conv_f2s1", "bn_f2s1", 32, 64, kl1, 1, True),
            _Spec("conv_f2s2", "bn_f2s2", 32, 64, kl2, 1, True),
        ])
        self._pack_cache = {}

    # ---- construction helpers (same names as the reference) ------------------------------------------------------
    def _init_conv_and_bn(self, temporal_kernelsize, in_channels, out_channels):
        conv = nn.Conv3d(in_channels, out_channels, kernel_size=(temporal_kernelsize, 3, 3), padding=(0, 1, 1))
        return conv, nn.BatchNorm3d(out_channels)

    def _init_fuse_and_bn(self, temporal_kernelsize):
        conv = nn.Conv3d(32, 64, kernel_size=[temporal_kernelsize, 1, 1], stride=[1, 1, 1], padding=[0, 0, 0], bias=False)
        return conv, nn.BatchNorm3d(64)

    def _calc_kernel_sizes(self, pathway_size):
        div, rem = divmod(pathway_size, 3)
        if rem == 0:
            return (div, div + 1, div + 1)
        if rem == 1:
            return (div + 1, div + 1, div + 1)
        return (div + 1, div + 1, div + 2)

    def _calc_fuse_kernel_size(self, slow_in, slow_kernel, fast_in, fast_kernel):
        out_slow = slow_in - slow_kernel + 1
        out_fast = fast_in - fast_kernel + 1
        return out_fast - out_slow + 1, out_slow, out_fast

    # ---- helpers ----------------------------------------------------------------------------------------------------
    @property
    def _umma(self):
        return self.precision != "fp32"

    @property
    def _act_dtype(self):
        return torch.bfloat16 if self._umma else torch.float32

    def _packed(self, name, mode):
        """Packed GEMM operand of a conv weight, cached until the parameter changes."""
        w = getattr(self, name).weight
        spec = self._specs[name]
        kc = spec.cin if mode == 0 else spec.cout
        cp = (32 if kc <= 32 else _round_up(kc, 64)) if self._umma else kc     # 32-wide K steps for Cin = 32 layers
        key = (name, mode, self._umma)
        if w.is_cuda and torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture the packing kernel must be part of the graph (a replay has to see the
            # parameter values of ITS step), so the cache is bypassed
            return ops.pack_weights(w, mode, BF16 if self._umma else F32, cp), cp
        tag = (w.data_ptr(), w._version, str(w.device))
        hit = self._pack_cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_weights(w, mode, BF16 if self._umma else F32, cp), cp)
            self._pack_cache[key] = hit
        return hit[1], hit[2]

    def _param_list(self):
        ps = []
        for s in self._specs.values():
            conv, bn = getattr(self, s.conv), getattr(self, s.bn)
            ps.append(conv.weight)
            if conv.bias is not None:
                ps.append(conv.bias)
            ps.extend([bn.weight, bn.bias])
        return ps

    # ---- reference API ------------------------------------------------------------------------------------------------
    def fuse(self, slow, fast, conv, bn):
        """model.py:111-116: ``(cat([slow, relu(bn(conv(fast)))], 1), fast)`` for one of the two lateral connections
        (``conv`` / ``bn`` = ``self.conv_f2s{1,2}`` / ``self.bn_f2s{1,2}``).  The fused pipeline (forward /
        temporally_enhance_features) never calls it - there the lateral convolution writes straight into channels 192..255
        of the slow buffer - but it is part of the reference's interface, so it is a differentiable op of its own here."""
        name = next((n for n in ("conv_f2s1", "conv_f2s2") if getattr(self, n) is conv), None)
        if name is None or getattr(self, self._specs[name].bn) is not bn:
            raise ValueError("fuse: conv / bn must be one of this module's lateral pairs (conv_f2s1, bn_f2s1) / (conv_f2s2, bn_f2s2)")
        lateral = _FuseFn.apply(self, name, torch.is_grad_enabled(), fast, conv.weight, bn.weight, bn.bias)
        return torch.cat([slow, lateral.to(slow.dtype)], 1), fast

    def forward(self, slow, fast):
        """slow [B,256,sp,H,W], fast [B,256,fp,H,W] (any strides) -> (slow [B,224,1,H,W], fast [B,32,1,H,W])."""
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        slow, fast = slow.to(dev), fast.to(dev)
        merged = _SlowFastLevelFn.apply(self, torch.is_grad_enabled(), None, None, slow, fast, *self._param_list())
        b, _, h, w = merged.shape
        return merged[:, :224].unsqueeze(2), merged[:, 224:].unsqueeze(2)

    def temporally_enhance_features(self, slow_features, fast_features):
        """list (len B) of {level: [T,256,H,W]} -> OrderedDict{level: [B,256,H,W]} (model.py:151-165)."""
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        params = self._param_list()
        keys = list(slow_features[0].keys())
        slow_lists = [[d[key].to(dev) for d in slow_features] for key in keys]
        fast_lists = [[d[key].to(dev) for d in fast_features] for key in keys]
        merged = OrderedDict()
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for lst in slow_lists + fast_lists for t in lst)
        if not needs_grad:
            outs = _SlowFastPyramidFn.apply(self, torch.is_grad_enabled(), slow_lists, fast_lists, *params)
            for key, out in zip(keys, outs):
                merged[key] = out
            return merged
        for key, slow_list, fast_list in zip(keys, slow_lists, fast_lists):     # inputs carry gradient: per-level nodes
            s = torch.stack(slow_list).transpose(1, 2)
            f = torch.stack(fast_list).transpose(1, 2)
            merged[key] = _SlowFastLevelFn.apply(self, True, None, None, s, f, *params)
        return merged


    @torch.no_grad()
    def temporally_enhance_sequence(self, frame_features, max_frames=None, halo=(0, 0)):
        """Eval-mode features of EVERY frame of one sequence in a single temporal sweep.

        ``frame_features``: OrderedDict{level: [F,256,H,W]} -- the backbone features of the F frames of a sequence, in
        order.  Returns OrderedDict{level: [F,256,H,W]} whose row t equals what ``temporally_enhance_features`` returns for
        the window centred on frame t (fast = frames [t - fp//2, t + ceil(fp/2)), slow = its centre sp frames, frames
        outside the sequence all-zero: code/helpers/model.py:215-248,322-340).

        Why it is the same arithmetic: every Conv3d has "valid" temporal padding and eval-mode BatchNorm is a fixed
        per-channel affine map, so the whole stack is a shift-invariant temporal filter; consecutive windows share
        fp - 1 frames and every intermediate activation at an absolute time is identical in all of them.  Feeding the
        zero-padded sequence ([F + fp - 1] frames, the slow pathway a frame range of the same buffer) as ONE clip makes each
        layer compute every such activation once: per frame fp/k-fold less convolution work (497.7 -> ~293 GFLOP at
        (1,8), 4261 -> ~668 GFLOP at (4,32)) and one layout conversion per frame instead of fp.  Not valid in train mode
        (batch statistics are per window), where it raises.  ``max_frames`` bounds the frames per sweep (memory).
        ``halo=(hl, hr)``: the first hl / last hr given frames are context only (a caller streaming a long sequence in
        pieces passes up to fp//2 / ceil(fp/2)-1 neighbouring frames); outputs are returned for the frames in between and
        only the part of a window that the given frames do not cover is zero-filled."""
        if self.training:
            raise RuntimeError("temporally_enhance_sequence needs eval mode: train-mode BatchNorm statistics are per window")
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):                    # libsfvos launches on the current device
                return self.temporally_enhance_sequence(frame_features, max_frames, halo)
        ops.device_check()
        sp, fp = self.slow_pathway_size, self.fast_pathway_size
        lo, hi = fp // 2, fp - fp // 2 - 1                  # zero frames before / after the sequence
        s_off = fp // 2 - sp // 2                           # first slow frame inside a fast window (_slice_features)
        keys = list(frame_features.keys())
        hl, hr = halo
        n_in = frame_features[keys[0]].shape[0]
        n = n_in - hl - hr                                  # frames that get an output
        assert n >= 1 and 0 <= hl <= lo and 0 <= hr <= hi, "halo must leave >= 1 frame and fit inside one window"
        chunk = n if not max_frames else max(1, int(max_frames))
        outs = OrderedDict((k, []) for k in keys)
        main = torch.cuda.current_stream(dev)
        streams = _level_streams(dev, [tuple(frame_features[k].shape[-2:]) for k in keys])   # smaller levels on side streams
        order = sorted(range(len(keys)), key=lambda j: (streams[j] is None, j))
        if any(st is not None for st in streams):
            _prepack(self, (0,))
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            chunk_outs = {}
            for ki in order:
                key = keys[ki]
                with _on_stream(streams[ki], main):
                    x = frame_features[key]
                    _, c, h, w = x.shape
                    t_in = (c1 - c0) + fp - 1                    # padded frames [c0 - lo, c1 + hi) of the output numbering
                    fast_in = Act.empty(1, t_in, h, w, c, self._act_dtype, dev)
                    f0, f1 = max(-hl, c0 - lo), min(n + hr, c1 + hi)   # given frames inside the padded range
                    per = h * w * c
                    left, right = f0 - (c0 - lo), (c1 + hi) - f1
                    if left:
                        fast_in.buf[:left * per].zero_()
                    if right:
                        fast_in.buf[(t_in - right) * per:].zero_()
                    _fill_frames(fast_in, left, x[f0 + hl:f1 + hl].to(dev))
                    slow_in = fast_in.frames(s_off, s_off + (c1 - c0) + sp - 1)
                    chunk_outs[key] = _level_forward(self, slow_in, fast_in, False, None).as_nchw()
            for st in streams:
                if st is not None:
                    main.wait_stream(st)
            for key in keys:
                outs[key].append(chunk_outs[key])
        return OrderedDict((k, v[0] if len(v) == 1 else torch.cat(v)) for k, v in outs.items())


# ----------------------------------------------------------------------------------------------------------------------
# engine
# ----------------------------------------------------------------------------------------------------------------------
def _frame_layout(t):
    """"nchw" / "nhwc": how the frames of a [F,C,H,W] tensor are laid out when each frame is dense and frames are
    consecutive (a frame range of such a tensor is again one); None otherwise.  "nhwc" = torch.channels_last, what the
    libsfvos FPN (backbone.py) emits."""
    if t.dim() != 4:
        return None
    if t.is_contiguous():
        return "nchw"
    if t.permute(0, 2, 3, 1).is_contiguous():
        return "nhwc"
    return None


def _fill_frames(dst, frame_off, src):
    """Frames [frame_off, frame_off+F) of the dense channels-last Act ``dst`` <- src [F,C,H,W] of any layout / float dtype:
    one layout-conversion launch, or a plain copy when the source already is channels-last in the activation dtype."""
    f, c, h, w = src.shape
    if (src.dtype == dst.buf.dtype and dst.cstride == c and _frame_layout(src) == "nhwc" and src.device == dst.buf.device):
        per = h * w * c
        o = dst.ch_off + frame_off * per
        dst.buf[o:o + f * per].view(f, h, w, c).copy_(src.permute(0, 2, 3, 1))
    else:
        ops.nchw_to_nhwc(_as_source(src), dst, frame_off=frame_off)


def _as_source(x):
    """Contiguous f32 or bf16 tensor (the layout kernel reads both; anything else is converted to f32 first)."""
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return x if x.is_contiguous() else x.contiguous()


def _to_act(x5, dtype):
    """[B,C,T,H,W] tensor (any strides) -> dense channels-last Act."""
    b, c, t, h, w = x5.shape
    frames = _as_source(x5.permute(0, 2, 1, 3, 4)).view(b * t, c, h, w)
    act = Act.empty(b, t, h, w, c, dtype, x5.device)
    ops.nchw_to_nhwc(frames, act)
    return act


def _clips_to_act(clips, dtype):
    """list of per-clip [T,C,H,W] fp32 tensors -> one channels-last Act [B,T,H,W,C] (no torch.stack copy)."""
    t, c, h, w = clips[0].shape
    act = Act.empty(len(clips), t, h, w, c, dtype, clips[0].device)
    for b, clip in enumerate(clips):
        _fill_frames(act, b * t, clip)
    return act


def _lists_to_acts(slow_list, fast_list, dt_act):
    """Per-clip [T,256,H,W] tensors -> (fast Act, slow Act); the slow window aliases the fast buffer when it is a frame
    range of it (what the reference's _slice_features yields).  When the fast clips themselves are consecutive windows of one
    sequence tensor (clip b = frames [b, b+T): how the reference forms its clips, model.py:318-337) every frame is converted
    ONCE into a channels-last sequence buffer and the clips are overlapping views of it (batch stride = one frame)."""
    off = _alias_offset(slow_list, fast_list)
    seq = _window_sequence(fast_list)
    if seq is not None:
        b, (t, c, h, w) = len(fast_list), fast_list[0].shape
        # tcgen05 path: the kernels address clips through a 5-D tensor map whose batch stride becomes one frame -- ask the
        # driver once whether it encodes such a map; the CUDA-core validation mode uses plain pointer arithmetic
        if dt_act == torch.float32 or _lib.load().sfvos_tma_overlap_supported(ops._p(seq)) == 1:
            if _frame_layout(seq) == "nhwc" and seq.dtype == dt_act and seq.data_ptr() % 16 == 0:
                flat = seq.permute(0, 2, 3, 1).reshape(-1)            # already channels-last in the activation dtype: no copy
            else:
                buf = Act.empty(1, b + t - 1, h, w, c, dt_act, seq.device)
                _fill_frames(buf, 0, seq)
                flat = buf.buf
            fast_in = Act(flat, b, t, h, w, c, c, 0, bstride=h * w * c)
            slow_in = fast_in.frames(off, off + slow_list[0].shape[0]) if off is not None else _clips_to_act(slow_list, dt_act)
            return fast_in, slow_in
    fast_in = _clips_to_act(fast_list, dt_act)
    if off is not None:
        return fast_in, fast_in.frames(off, off + slow_list[0].shape[0])
    return fast_in, _clips_to_act(slow_list, dt_act)


def _window_sequence(fast_list):
    """If clip b is frames [b, b+T) of ONE contiguous [F,C,H,W] tensor (same storage, consecutive start frames) return that
    tensor's frames [0, B+T-1) as a view, else None.  SFVOS_WINDOW_DEDUP=0 disables the detection."""
    if len(fast_list) < 2 or os.environ.get("SFVOS_WINDOW_DEDUP", "1") == "0":
        return None
    f0 = fast_list[0]
    layout = _frame_layout(f0)
    if f0.dtype not in (torch.float32, torch.bfloat16) or layout is None:
        return None
    frame = f0[0].numel()
    base = f0.untyped_storage().data_ptr()
    for b, f in enumerate(fast_list):
        if (f.shape != f0.shape or f.dtype != f0.dtype or _frame_layout(f) != layout or f.untyped_storage().data_ptr() != base
                or f.storage_offset() != f0.storage_offset() + b * frame):
            return None
    n = len(fast_list) + f0.shape[0] - 1
    if (f0.storage_offset() + n * frame) * f0.element_size() > f0.untyped_storage().nbytes():
        return None
    return torch.as_strided(f0, (n,) + tuple(f0.shape[1:]), f0.stride(), f0.storage_offset())


def _alias_offset(slow_list, fast_list):
    """If every slow clip is a contiguous frame range of its fast clip (the reference's _slice_features,
    model.py:242-248, yields exactly such views) return the common first-frame offset, else None."""
    off = None
    for s, f in zip(slow_list, fast_list):
        if not (_frame_layout(f) is not None and _frame_layout(s) == _frame_layout(f) and s.dtype == f.dtype
                and f.dtype in (torch.float32, torch.bfloat16)):
            return None
        if s.untyped_storage().data_ptr() != f.untyped_storage().data_ptr() or s.shape[1:] != f.shape[1:]:
            return None
        frame_bytes = f[0].numel() * f.element_size()
        delta = s.data_ptr() - f.data_ptr()
        if delta < 0 or delta % frame_bytes:
            return None
        o = delta // frame_bytes
        if o + s.shape[0] > f.shape[0] or (off is not None and o != off):
            return None
        off = o
    return off


class _Scratch:
    """Bump allocator over ONE zero-filled f32 tensor: every small statistics / gradient buffer of a step is a view
    of it, so a step issues one fill instead of hundreds (each torch.zeros is a kernel launch plus host overhead)."""

    def __init__(self, n, device):
        self.buf = torch.zeros(n, dtype=torch.float32, device=device)
        self.off = 0
        self.deferred = None         # see _conv_bn_forward

    def take(self, n):
        v = self.buf[self.off:self.off + n]
        assert v.numel() == n, "scratch exhausted"
        self.off += (n + 3) // 4 * 4            # keep every view 16-byte aligned (vector atomics)
        return v


def _apply_deferred_running_stats(mod, deferred):
    """deferred: {layer: [(stats, count) per level, in LEVEL order]} -> one EMA kernel per layer (level order preserved)."""
    for name, calls in deferred.items():
        spec = mod._specs[name]
        conv, bn = getattr(mod, spec.conv), getattr(mod, spec.bn)
        ops.bn_running_update(calls, conv.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                              BN_MOMENTUM if bn.momentum is None else bn.momentum)


def _fwd_scratch_size(mod, n_levels):
    return n_levels * sum(8 * s.cout + 8 for s in mod._specs.values())       # (sum, sumsq) [f64 in validation mode] + bn4


def _conv_bn_forward(mod, spec, x, out, training, saved, scratch=None):
    """conv -> BN (-> ReLU) of one layer, writing into ``out`` (an Act, possibly a channel slice).
    If ``scratch.deferred`` is a dict, the running-statistics update is NOT done here: (stats, count) is appended to
    ``scratch.deferred[layer]`` and the caller applies the updates of all pyramid levels in level order after they joined
    (``_apply_deferred_running_stats``) - the levels run concurrently, the EMA is order-dependent."""
    conv, bn = getattr(mod, spec.conv), getattr(mod, spec.bn)
    umma = mod._umma
    wp, cp = mod._packed(spec.conv, 0)
    to = x.T - spec.kt + 1
    dev = x.buf.device
    pad = (0, spec.pad, spec.pad)
    if training:
        raw = Act.empty(x.B, to, x.H, x.W, spec.cout, torch.float32, dev)
        if umma:
            stats = scratch.take(2 * spec.cout) if scratch is not None else torch.zeros(2 * spec.cout, dtype=torch.float32, device=dev)
            ops.conv(x, wp, cp, spec.cout, spec.k, pad, to, raw, umma=True, stats=stats)
        else:
            # validation mode: fp64 statistics reduced in a fixed order (bit-reproducible, no cancellation)
            stats = (scratch.take(4 * spec.cout).view(torch.float64) if scratch is not None
                     else torch.empty(2 * spec.cout, dtype=torch.float64, device=dev))
            ops.conv(x, wp, cp, spec.cout, spec.k, pad, to, raw, umma=False)
            ops.channel_stats(raw, stats)
        bn4 = scratch.take(4 * spec.cout) if scratch is not None else torch.empty(4 * spec.cout, dtype=torch.float32, device=dev)
        track = bn.track_running_stats and bn.running_mean is not None
        deferred = getattr(scratch, "deferred", None) if scratch is not None else None
        if track and deferred is not None:
            deferred.setdefault(spec.conv, []).append((stats, raw.npix))
            track = False
        ops.bn_finalize(stats, raw.npix, conv.bias, bn.weight, bn.bias, bn.running_mean if track else None,
                        bn.running_var if track else None, bn.num_batches_tracked if track else None,
                        BN_MOMENTUM if bn.momentum is None else bn.momentum, bn.eps, bn4)
        ops.affine_act(raw, out, bn4[:spec.cout], bn4[spec.cout:2 * spec.cout], spec.relu)
        if saved is not None:
            saved[spec.conv] = (raw, bn4, False)
    elif saved is not None:
        # eval mode WITH autograd (fine-tuning against frozen BatchNorm statistics; the reference module backpropagates
        # through eval-mode BN like any nn.Module): keep the raw conv output like the train path does; BatchNorm is the fixed
        # affine map of the running statistics and its backward has no batch-statistics terms (fixed_stats)
        raw = Act.empty(x.B, to, x.H, x.W, spec.cout, torch.float32, dev)
        ops.conv(x, wp, cp, spec.cout, spec.k, pad, to, raw, umma=umma)
        bn4 = torch.empty(4 * spec.cout, dtype=torch.float32, device=dev)
        ops.bn_fold_eval(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn4)
        ops.affine_act(raw, out, bn4[:spec.cout], bn4[spec.cout:2 * spec.cout], spec.relu)
        saved[spec.conv] = (raw, bn4, True)
    else:
        fold = torch.empty(2 * spec.cout, dtype=torch.float32, device=dev)
        ops.bn_fold_eval(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, fold)
        ops.conv(x, wp, cp, spec.cout, spec.k, pad, to, out, umma=umma, relu=spec.relu, scale=fold[:spec.cout],
                 shift=fold[spec.cout:])


def _level_forward(mod, slow_in, fast_in, training, saved, scratch=None, fast_stream=None):
    """One pyramid level (model.py:118-149 + the concat of :162).  Returns the merged f32 Act [B,T3,H,W,256]: T3 = 1 for
    the reference's windows (sp / fp frames in); in sequence mode (temporally_enhance_sequence) the inputs are whole
    zero-padded sequences and T3 = the number of frames.

    ``fast_stream``: run the fast pathway (and the lateral convolutions it feeds into the slow buffers) on that CUDA stream,
    concurrently with the slow pathway on the current one.  The two only meet at the fuse points - slow_conv{2,3} read the
    lateral channels - so the current stream waits for the fast one there; the fast stream never waits.  The HBM-bound
    BatchNorm passes of one pathway then run under the tensor-bound convolutions of the other."""
    sp = mod._specs
    dt_act = mod._act_dtype
    dev = fast_in.buf.device
    B, H, W = fast_in.B, fast_in.H, fast_in.W
    t1s, t1f = slow_in.T - sp["slow_conv1"].kt + 1, fast_in.T - sp["fast_conv1"].kt + 1
    t2s, t2f = t1s - sp["slow_conv2"].kt + 1, t1f - sp["fast_conv2"].kt + 1
    t3 = t2s - sp["slow_conv3"].kt + 1
    assert t3 >= 1 and t3 == t2f - sp["fast_conv3"].kt + 1, "slow / fast temporal extents do not meet after layer 3"
    # every buffer both pathways write is allocated BEFORE the fork (the caching allocator orders reuse per stream)
    s1 = Act.empty(B, t1s, H, W, 256, dt_act, dev)
    s2 = Act.empty(B, t2s, H, W, 256, dt_act, dev)
    out = Act.empty(B, t3, H, W, 256, torch.float32, dev)     # = cat([slow, fast], 1).squeeze(2) when T3 = 1
    main = torch.cuda.current_stream(dev) if fast_stream is not None else None

    def fast(fn):
        if fast_stream is None:
            return fn()
        with torch.cuda.stream(fast_stream):
            return fn()

    def join():
        if fast_stream is not None:
            main.wait_stream(fast_stream)

    if fast_stream is not None:
        fast_stream.wait_stream(main)                          # inputs (layout conversion) are ready
    f1 = fast(lambda: Act.empty(B, t1f, H, W, 32, dt_act, dev))
    f2 = fast(lambda: Act.empty(B, t2f, H, W, 32, dt_act, dev))
    # layer 1 (+ lateral 1 straight into channels 192..255 of the slow buffer)
    fast(lambda: (_conv_bn_forward(mod, sp["fast_conv1"], fast_in, f1, training, saved, scratch),
                  _conv_bn_forward(mod, sp["conv_f2s1"], f1, s1.slice(192, 64), training, saved, scratch)))
    _conv_bn_forward(mod, sp["slow_conv1"], slow_in, s1.slice(0, 192), training, saved, scratch)
    join()
    # layer 2
    fast(lambda: (_conv_bn_forward(mod, sp["fast_conv2"], f1, f2, training, saved, scratch),
                  _conv_bn_forward(mod, sp["conv_f2s2"], f2, s2.slice(192, 64), training, saved, scratch)))
    _conv_bn_forward(mod, sp["slow_conv2"], s1, s2.slice(0, 192), training, saved, scratch)
    join()
    # layer 3: both pathways land in one f32 [B,T3,H,W,256] buffer
    fast(lambda: _conv_bn_forward(mod, sp["fast_conv3"], f2, out.slice(224, 32), training, saved, scratch))
    _conv_bn_forward(mod, sp["slow_conv3"], s2, out.slice(0, 224), training, saved, scratch)
    join()
    if saved is not None:
        saved["_acts"] = dict(slow_in=slow_in, fast_in=fast_in, f1=f1, s1=s1, f2=f2, s2=s2)
    return out


class _GradSlot:
    """One set of gradient accumulators: parameter-gradient views + packed weight-gradient accumulators."""
    __slots__ = ("grads", "dwp", "deterministic", "swapped")

    def __init__(self, deterministic):
        self.grads, self.dwp, self.deterministic = {}, {}, deterministic
        self.swapped = set()          # convs whose dwp accumulator holds the operand-swapped layout (see _layer_backward)


class _GradBank:
    """Every parameter gradient of the module as a view of ONE zero-filled f32 buffer, plus the packed weight-gradient
    accumulators and the BN-backward sums.

    Product path (bf16): all pyramid levels accumulate into the same views inside the kernels (atomics), so the 5 levels
    cost no per-parameter torch.add and one fill.  Validation mode (``deterministic``): every level gets its own slot of
    accumulators - the fp32 kernels add without atomics, and the levels may run concurrently - and ``finish`` sums the
    slots in LEVEL order, so the result is bit-identical whatever the scheduling."""

    def __init__(self, mod, n_levels, device, deterministic=False):
        specs = list(mod._specs.values())
        per_slot = 0
        for s in specs:
            conv = getattr(mod, s.conv)
            per_slot += 2 * (conv.weight.numel() + 4) + 3 * (s.cout + 4)
        per_slot = (per_slot + 3) // 4 * 4
        n_slots = n_levels if deterministic else 1
        sums = sum(n_levels * (2 * s.cout + 4) for s in specs)
        self.scratch = _Scratch(n_slots * per_slot + sums, device)
        self.per_slot, self.deterministic = per_slot, deterministic
        self.slots = []
        # product path with a gradient arena (ops.GRAD_ARENA, data-parallel step): the parameter gradients live in the arena
        # (accumulated into, cleared once per optimizer step by its owner); only the packed accumulators stay in the scratch
        arena = None if deterministic else ops.GRAD_ARENA

        def grad_view(param):
            v = arena.view(param) if arena is not None else None
            return v if v is not None else self.scratch.take(param.numel()).view(param.shape)
        for i in range(n_slots):
            assert self.scratch.off == i * per_slot
            slot = _GradSlot(deterministic)
            for s in specs:
                conv, bn = getattr(mod, s.conv), getattr(mod, s.bn)
                if conv.weight.requires_grad:
                    slot.grads[s.conv + ".weight"] = grad_view(conv.weight)
                    slot.dwp[s.conv] = self.scratch.take(s.kt * s.khw * s.khw * s.cin * s.cout)
                if conv.bias is not None:
                    # a per-channel constant added before train-mode BN has exactly zero gradient
                    slot.grads[s.conv + ".bias"] = grad_view(conv.bias)
                slot.grads[s.bn + ".weight"] = grad_view(bn.weight)
                slot.grads[s.bn + ".bias"] = grad_view(bn.bias)
            self.scratch.off = (i + 1) * per_slot
            self.slots.append(slot)

    def slot(self, i):
        return self.slots[i if self.deterministic else 0]

    def finish(self, mod):
        first = self.scratch.buf[:self.per_slot]
        for i in range(1, len(self.slots)):                    # fixed order: level 0 + level 1 + ...
            ops.axpby(self.scratch.buf[i * self.per_slot:(i + 1) * self.per_slot], first, 1.0, 1.0)
        grads = self.slots[0].grads
        for name, dwp in self.slots[0].dwp.items():
            if name in self.slots[0].swapped:
                # operand-swapped lateral weight gradient: dwp = [kt][Cout][Cin] with the taps reversed (_layer_backward)
                g = grads[name + ".weight"]
                cout, cin, kt = g.shape[0], g.shape[1], g.shape[2]
                g.add_(dwp.view(kt, cout, cin).flip(0).permute(1, 2, 0).reshape(g.shape))
            else:
                ops.unpack_wgrad(dwp, grads[name + ".weight"], 0)
        return grads


def _layer_backward(mod, spec, dy, x_in, saved, bank, dx=None, dx_accumulate=False, need_dx=True, dx_dtype=torch.float32,
                    dx_addend=None):
    """BN(+ReLU) backward -> weight gradient -> (optionally) data gradient of one layer.
    dy: Act gradient wrt the layer's post-activation output; returns dx Act (``dx_dtype``) or None.
    ``dx_addend``: an Act like dx; the returned gradient is dgrad + dx_addend, stored once in ``dx_dtype``.
    ``bank``: a (_GradBank, _GradSlot) pair - the level's accumulators."""
    bank, slot = bank
    conv, bn = getattr(mod, spec.conv), getattr(mod, spec.bn)
    umma = mod._umma
    raw, bn4, fixed_stats = saved[spec.conv]
    dev = raw.buf.device
    dconv = Act.empty(raw.B, raw.T, raw.H, raw.W, spec.cout, mod._act_dtype, dev)
    ops.bn_bwd(dy, raw, bn4, bn.weight, spec.relu, dconv, slot.grads[spec.bn + ".weight"], slot.grads[spec.bn + ".bias"],
               sums=bank.scratch.take(2 * spec.cout), deterministic=slot.deterministic, fixed_stats=fixed_stats,
               dbias=slot.grads.get(spec.conv + ".bias") if fixed_stats else None)
    pad = (0, spec.pad, spec.pad)
    if conv.weight.requires_grad:
        if (umma and not slot.deterministic and spec.khw == 1 and spec.cin == 32 and spec.cout % 64 == 0 and spec.cout <= 128
                and os.environ.get("SFVOS_LATERAL_WGRAD_SWAP", "1") != "0"):
            # Lateral connection (32 -> 64 channels, k_t x 1 x 1): as dw[ta][c][n] = sum_t x[t+ta][c] dy[t][n] it is k_t GEMMs
            # of 32 x 64 that each re-read dy and leave 3/4 of the MMA rows empty.  With the operands SWAPPED - "x" := dy,
            # "dy" := x, temporal padding k_t - 1 - the same sum is dw'[k_t-1-ta][n][c] = sum_t' dy[t'-ta][n] x[t'][c], a problem
            # with 32 output channels: the stacked weight-gradient kernel (wgrad_stack) fetches the dy tile once and
            # multiplies it by all k_t input frames stacked along N (one 128 x 32 k_t x 16 MMA per K step; every byte of x and
            # dy read once).  The accumulator keeps the swapped layout; _GradBank.finish un-swaps it.
            ops.wgrad(dconv, x_in, spec.k, (spec.kt - 1, 0, 0), slot.dwp[spec.conv], umma=True,
                      flops=2.0 * dconv.npix * spec.cin * spec.cout * spec.kt)
            slot.swapped.add(spec.conv)
        else:
            slot.swapped.discard(spec.conv)
            ops.wgrad(x_in, dconv, spec.k, pad, slot.dwp[spec.conv], umma=umma)
    if not need_dx:
        return None
    wd, cpd = mod._packed(spec.conv, 1)
    if dx is None:
        dx = Act.empty(x_in.B, x_in.T, x_in.H, x_in.W, spec.cin, dx_dtype, dev)
    dpad = (spec.kt - 1, spec.khw - 1 - spec.pad, spec.khw - 1 - spec.pad)
    ops.conv(dconv, wd, cpd, spec.cin, spec.k, dpad, x_in.T, dx, umma=umma, accumulate=dx_accumulate, addend=dx_addend)
    return dx


def _level_backward(mod, saved, g_out, need_input_grad, bank, fast_stream=None):
    """g_out: f32 Act [B,1,H,W,256] = gradient of the merged output; parameter gradients accumulate into ``bank`` =
    (_GradBank, this level's _GradSlot).
    Returns (d_slow, d_fast).  ``fast_stream``: the fast pathway's (and the laterals') backward on that stream, mirroring
    _level_forward: it waits for the slow pathway only where a lateral needs the slow gradient (d_s2 / d_s1)."""
    sp = mod._specs
    a = saved["_acts"]
    dev = g_out.buf.device
    main = torch.cuda.current_stream(dev) if fast_stream is not None else None

    def fast(fn):
        if fast_stream is None:
            return fn()
        with torch.cuda.stream(fast_stream):
            return fn()

    def lateral_sum(spec, dy, x_in, d_part):
        """Gradient wrt a fast-pathway activation = its fast convolution's dgrad (d_part, f32) + the lateral's dgrad.
        Tensor-core path: the lateral's epilogue reads the f32 partial sum and stores the total ONCE in the activation
        dtype for the BatchNorm backward that consumes it (instead of a read-modify-write of the f32 buffer followed by two
        f32 reads): 6 of 20 bytes per element less.  Validation mode: f32 accumulate, as before."""
        if use_addend:
            return _layer_backward(mod, spec, dy, x_in, saved, bank, dx_dtype=gdt, dx_addend=d_part)
        _layer_backward(mod, spec, dy, x_in, saved, bank, dx=d_part, dx_accumulate=True)
        return d_part

    # the slow pathway's data gradients are consumed once, by the BN-backward passes of the layer below, which round
    # to the activation dtype anyway: store them in it (bf16 on the product path) -- half the bytes of three passes
    gdt = mod._act_dtype if os.environ.get("SFVOS_BF16_DGRAD", "1") != "0" else torch.float32
    if fast_stream is not None:
        fast_stream.wait_stream(main)                          # g_out is ready
    # the fast convolutions' dgrads are PARTIAL sums that the lateral's epilogue reads once (lateral_sum): store them in the
    # activation dtype as well (what bf16 autocast training does with every activation gradient); SFVOS_FAST_PARTIAL_BF16=0: f32
    use_addend = mod._umma and gdt != torch.float32 and os.environ.get("SFVOS_LATERAL_ADDEND", "1") != "0"
    pdt = gdt if use_addend and os.environ.get("SFVOS_FAST_PARTIAL_BF16", "1") != "0" else torch.float32
    d_f2 = fast(lambda: _layer_backward(mod, sp["fast_conv3"], g_out.slice(224, 32), a["f2"], saved, bank, dx_dtype=pdt))
    d_s2 = _layer_backward(mod, sp["slow_conv3"], g_out.slice(0, 224), a["s2"], saved, bank, dx_dtype=gdt)
    if fast_stream is not None:
        fast_stream.wait_stream(main)                          # lateral 2 reads d_s2[192:]
    d_f1 = fast(lambda: _layer_backward(mod, sp["fast_conv2"], lateral_sum(sp["conv_f2s2"], d_s2.slice(192, 64), a["f2"], d_f2),
                                        a["f1"], saved, bank, dx_dtype=pdt))
    d_s1 = _layer_backward(mod, sp["slow_conv2"], d_s2.slice(0, 192), a["s1"], saved, bank, dx_dtype=gdt)
    if fast_stream is not None:
        fast_stream.wait_stream(main)                          # lateral 1 reads d_s1[192:]
    d_fast = fast(lambda: _layer_backward(mod, sp["fast_conv1"], lateral_sum(sp["conv_f2s1"], d_s1.slice(192, 64), a["f1"], d_f1),
                                          a["fast_in"], saved, bank, need_dx=need_input_grad))
    d_slow = _layer_backward(mod, sp["slow_conv1"], d_s1.slice(0, 192), a["slow_in"], saved, bank, need_dx=need_input_grad)
    if fast_stream is not None:
        main.wait_stream(fast_stream)
    return d_slow, d_fast


def _grad_to_act(g):
    b, c, h, w = g.shape
    gl = g.permute(0, 2, 3, 1)
    if gl.is_contiguous() and g.dtype == torch.float32:
        return Act(gl.reshape(-1), b, 1, h, w, c, c, 0)
    g_act = Act.empty(b, 1, h, w, c, torch.float32, g.device)
    ops.nchw_to_nhwc(g.float().contiguous(), g_act)
    return g_act


def _param_names(mod):
    names = []
    for s in mod._specs.values():
        names.append(s.conv + ".weight")
        if getattr(mod, s.conv).bias is not None:
            names.append(s.conv + ".bias")
        names.extend([s.bn + ".weight", s.bn + ".bias"])
    return names


def _act_to_ncdhw(act):
    out = torch.empty(act.B * act.T, act.C, act.H, act.W, dtype=torch.float32, device=act.buf.device)
    ops.nhwc_to_nchw(act, out)
    return out.view(act.B, act.T, act.C, act.H, act.W).permute(0, 2, 1, 3, 4)


class _FuseFn(torch.autograd.Function):
    """One lateral connection (Conv3d k x 1 x 1, no bias -> BatchNorm3d -> ReLU) as a differentiable op: fast [B,32,T,H,W]
    -> [B,64,T-k+1,H,W] (f32, channels-last memory)."""

    @staticmethod
    @ops.device_guard
    def forward(ctx, mod, name, grad_enabled, fast, weight, gamma, beta):
        ops.device_check()
        spec = mod._specs[name]
        f = _to_act(fast, mod._act_dtype)
        out = Act.empty(f.B, f.T - spec.kt + 1, f.H, f.W, spec.cout, torch.float32, fast.device)
        want_grad = grad_enabled and (fast.requires_grad or weight.requires_grad or gamma.requires_grad or beta.requires_grad)
        saved = {} if want_grad else None
        _conv_bn_forward(mod, spec, f, out, mod.training, saved)
        ctx.mod, ctx.spec, ctx.saved, ctx.x_in = mod, spec, saved, f
        lateral = out.buf.view(f.B, out.T, f.H, f.W, spec.cout).permute(0, 4, 1, 2, 3)
        if saved is None:
            ctx.mark_non_differentiable(lateral)
        return lateral

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        mod, spec, saved, f = ctx.mod, ctx.spec, ctx.saved, ctx.x_in
        if saved is None:
            raise RuntimeError("SlowFastLayers.fuse: backward needs a forward with grad enabled")
        b, c, t, h, w = g.shape
        gl = g.permute(0, 2, 3, 4, 1)
        if not (gl.is_contiguous() and g.dtype == torch.float32):
            gl = gl.float().contiguous()
        bank = _GradBank(mod, 1, g.device, deterministic=not mod._umma)
        dx = _layer_backward(mod, spec, Act(gl.reshape(-1), b, t, h, w, c), f, saved, (bank, bank.slot(0)),
                             need_dx=ctx.needs_input_grad[3])
        grads = bank.finish(mod)
        ctx.saved = None
        return (None, None, None, _act_to_ncdhw(dx) if dx is not None else None, grads[spec.conv + ".weight"],
                grads[spec.bn + ".weight"], grads[spec.bn + ".bias"])


class _SlowFastLevelFn(torch.autograd.Function):
    """forward(slow, fast) of one pyramid level with a hand-written backward over the saved raw conv outputs."""

    @staticmethod
    @ops.device_guard
    def forward(ctx, mod, grad_enabled, slow_list, fast_list, slow5, fast5, *params):
        ops.device_check()
        dt_act = mod._act_dtype
        if fast5 is not None:
            fast_in = _to_act(fast5, dt_act)
            slow_in = _to_act(slow5, dt_act)
        else:
            fast_in, slow_in = _lists_to_acts(slow_list, fast_list, dt_act)
        training = mod.training
        want_grad = grad_enabled and (
            any(p.requires_grad for p in params) or (fast5 is not None and (fast5.requires_grad or slow5.requires_grad)))
        saved = {} if want_grad else None
        scratch = _Scratch(_fwd_scratch_size(mod, 1), fast_in.buf.device) if training else None
        out = _level_forward(mod, slow_in, fast_in, training, saved, scratch)
        ctx.mod, ctx.saved_acts = mod, saved
        ctx.need_input_grad = fast5 is not None and (fast5.requires_grad or slow5.requires_grad)
        merged = out.as_nchw()                      # [B,256,H,W] f32 view, channels_last strides
        if saved is None:
            ctx.mark_non_differentiable(merged)
        return merged

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        mod, saved = ctx.mod, ctx.saved_acts
        if saved is None:
            raise RuntimeError("SlowFastLayers: backward needs a forward with grad enabled")
        bank = _GradBank(mod, 1, g.device, deterministic=not mod._umma)
        d_slow, d_fast = _level_backward(mod, saved, _grad_to_act(g), ctx.need_input_grad, (bank, bank.slot(0)))
        grads = bank.finish(mod)
        ctx.saved_acts = None
        g_slow = _act_to_ncdhw(d_slow) if d_slow is not None else None
        g_fast = _act_to_ncdhw(d_fast) if d_fast is not None else None
        return (None, None, None, None, g_slow, g_fast) + tuple(grads.get(n) for n in _param_names(mod))


_SIDE_STREAMS = {}


def _level_streams(dev, shapes):
    """One side stream per pyramid level except the largest (None = the current stream).  The launches of the smaller levels
    fill a fraction of the 148 SMs and are latency chains, not throughput; run concurrently they soak into each other's idle
    SMs and into the tails of the largest level's persistent kernels (measured on the bench step, same box: 23.16 -> 22.13 ms
    with levels 2..4 on side streams, 21.54 -> 21.2 ms with level 1 as well).  Every level still runs its own kernels in order
    on ONE stream, and the parameter-gradient accumulators are atomics, so nothing else changes.  SFVOS_LEVEL_STREAMS=0 puts
    everything back on the current stream; SFVOS_LEVEL_STREAMS_MAXPIX=<H*W> restricts the side streams to levels up to that size."""
    if os.environ.get("SFVOS_LEVEL_STREAMS", "1") == "0" or len(shapes) < 3:
        return [None] * len(shapes)
    sizes = [h * w for h, w in shapes]
    max_pix = int(os.environ.get("SFVOS_LEVEL_STREAMS_MAXPIX", 0)) or (max(sizes) - 1)
    small = [i for i, n in enumerate(sizes) if n <= max_pix and n < max(sizes)]
    if not small:
        return [None] * len(shapes)
    pool = _SIDE_STREAMS.setdefault(str(dev), [])
    while len(pool) < len(small):
        pool.append(torch.cuda.Stream(device=dev))
    out = [None] * len(shapes)
    for n, i in enumerate(small):
        out[i] = pool[n]
    return out


def _prepack(mod, modes):
    """Fill the packed-weight cache on the CURRENT stream before the level streams fork: a cached operand packed on one
    side stream would otherwise be read by the other streams without any ordering.  (During a CUDA-graph capture the cache is
    bypassed and every use packs on its own stream.)"""
    if torch.cuda.is_current_stream_capturing():
        return
    for name, spec in mod._specs.items():
        for mode in modes:
            if mode == 1 and name in ("slow_conv1", "fast_conv1"):
                continue                                        # first-layer inputs carry no gradient on this path
            mod._packed(name, mode)


def _pathway_stream(dev):
    """Side stream for the fast pathway of the level that runs on the current stream.  Opt-in (SFVOS_PATH_STREAMS=1): measured
    on the bench step it does not pay (21.3-21.4 ms without, 21.6-21.7 ms with, same box) - both pathways' kernels are
    persistent, whole-GPU launches, and the chip is at its power cap - unlike the level streams above."""
    if os.environ.get("SFVOS_PATH_STREAMS", "0") != "1":
        return None
    key = "path:" + str(dev)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


class _on_stream:
    """``with _on_stream(side, main)``: fork ``side`` from ``main`` and make it current (no-op for side = None)."""

    def __init__(self, side, main):
        self.side, self.main, self.cm = side, main, None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(self.main)
            self.cm = torch.cuda.stream(self.side)
            self.cm.__enter__()

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)


class _SlowFastPyramidFn(torch.autograd.Function):
    """All pyramid levels of temporally_enhance_features (model.py:151-165) in ONE autograd node: the levels share the
    statistics scratch and accumulate their parameter gradients in-kernel into one buffer (no per-level torch.add)."""

    @staticmethod
    @ops.device_guard
    def forward(ctx, mod, grad_enabled, slow_lists, fast_lists, *params):
        ops.device_check()
        ctx.set_materialize_grads(False)
        dt_act = mod._act_dtype
        training = mod.training
        want_grad = grad_enabled and any(p.requires_grad for p in params)
        dev = fast_lists[0][0].device
        scratch = _Scratch(_fwd_scratch_size(mod, len(fast_lists)), dev) if training else None
        outs, saved_all = [], []
        main = torch.cuda.current_stream(dev)
        streams = _level_streams(dev, [tuple(fl[0].shape[-2:]) for fl in fast_lists])
        n_lv = len(fast_lists)
        outs, saved_all = [None] * n_lv, [None] * n_lv
        # small levels are enqueued FIRST (their streams fork from here), the large ones follow on the current stream
        concurrent = any(st is not None for st in streams)
        if concurrent:
            _prepack(mod, (0,))
        per_level = [dict() for _ in range(n_lv)]
        for i in sorted(range(n_lv), key=lambda j: (streams[j] is None, j)):
            with _on_stream(streams[i], main):
                if scratch is not None:
                    scratch.deferred = per_level[i] if concurrent else None
                fast_in, slow_in = _lists_to_acts(slow_lists[i], fast_lists[i], dt_act)
                saved = {} if want_grad else None
                outs[i] = _level_forward(mod, slow_in, fast_in, training, saved, scratch,
                                         fast_stream=_pathway_stream(dev) if streams[i] is None and any(st is not None for st in streams) else None).as_nchw()
                saved_all[i] = saved
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        if scratch is not None and concurrent:
            scratch.deferred = None
            merged = OrderedDict()
            for name in mod._specs:                          # level order inside every layer = the reference's call order
                calls = [c for lv in per_level for c in lv.get(name, [])]
                if calls:
                    merged[name] = calls
            _apply_deferred_running_stats(mod, merged)
        ctx.mod, ctx.saved_all, ctx.streams = mod, (saved_all if want_grad else None), streams
        if not want_grad:
            ctx.mark_non_differentiable(*outs)
        return tuple(outs)

    @staticmethod
    @ops.device_guard
    def backward(ctx, *gs):
        mod, saved_all = ctx.mod, ctx.saved_all
        if saved_all is None:
            raise RuntimeError("SlowFastLayers: backward needs a forward with grad enabled")
        live = [i for i, g in enumerate(gs) if g is not None]
        dev = gs[live[0]].device
        bank = _GradBank(mod, max(1, len(live)), dev, deterministic=not mod._umma)
        main = torch.cuda.current_stream(dev)
        if any(ctx.streams[i] is not None for i in live):
            _prepack(mod, (1,))
        for i in reversed(live):
            with _on_stream(ctx.streams[i], main):
                _level_backward(mod, saved_all[i], _grad_to_act(gs[i]), False, (bank, bank.slot(live.index(i))),
                                fast_stream=_pathway_stream(dev) if ctx.streams[i] is None and any(st is not None for st in ctx.streams) else None)
            saved_all[i] = None
        for i in live:
            if ctx.streams[i] is not None:
                main.wait_stream(ctx.streams[i])
        grads = bank.finish(mod)
        ctx.saved_all = None
        return (None, None, None, None) + tuple(grads.get(n) for n in _param_names(mod))
