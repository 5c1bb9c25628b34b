"""ROIAlign + mask branch on libsfvos.so, behind torchvision's own module interfaces.

The reference reaches this arithmetic at code/helpers/model.py:346 (``self.maskrcnn_model.roi_heads(...)``); the
modules below keep torchvision's constructor arguments, call signatures and state_dict keys so they can be swapped
into ``maskrcnn_model.roi_heads`` (see ``install`` and INTEGRATION.md):

  MultiScaleRoIAlign   <- torchvision.ops.MultiScaleRoIAlign            (TV/ops/poolers.py:230-321)
  MaskRCNNHeads        <- torchvision...mask_rcnn.MaskRCNNHeads          (TV/models/detection/mask_rcnn.py:271-303)
  MaskRCNNPredictor    <- torchvision...mask_rcnn.MaskRCNNPredictor      (TV/models/detection/mask_rcnn.py:337-353)
  maskrcnn_loss / maskrcnn_inference <- TV/models/detection/roi_heads.py:56-129
  TwoMLPHead           <- torchvision...faster_rcnn.TwoMLPHead           (TV/models/detection/faster_rcnn.py:286-307)
  FastRCNNPredictor    <- torchvision...faster_rcnn.FastRCNNPredictor    (TV/models/detection/faster_rcnn.py:347-370)
  fastrcnn_loss        <- TV/models/detection/roi_heads.py:12-53
  RoIHeads             <- torchvision...roi_heads.RoIHeads.forward       (TV/models/detection/roi_heads.py:739-887)
  paste_masks_in_image / postprocess <- TV/models/detection/roi_heads.py:415-501, transform.py:257-279 (8(f) rank 4)

Tensors exchanged between these modules keep torchvision's logical shapes ([K,C,P,P]) but are channels_last in
memory and bf16 on the product path (fp32 when ``precision == "fp32"``), which is what the tensor-core kernels
consume directly.  The box head (fc6/fc7), the box predictor and fastrcnn_loss run on the same tensor-core GEMM
kernels: a Linear layer is a 1x1x1 convolution over the M ROIs.  Box sampling, box decoding and NMS
stay torchvision/torch.
"""
import math
import os
import weakref
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn, Tensor
from torchvision.models.detection import faster_rcnn as tv_faster_rcnn
from torchvision.models.detection import mask_rcnn as tv_mask_rcnn
from torchvision.models.detection import roi_heads as tv_roi_heads

from . import ops
from ._lib import BF16, F32, call
from .ops import Act, _p, stream


def _default_precision():
    return os.environ.get("SFVOS_PRECISION", "bf16")


def _act_dtype(precision):
    return torch.float32 if precision == "fp32" else torch.bfloat16


def _is_channels_last(x):
    return x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()


def _nchw_view(buf, K, H, W, C):
    """[K,C,H,W]-shaped view (channels_last strides) of a flat channels-last buffer."""
    return buf.view(K, H, W, C).permute(0, 3, 1, 2)


def _to_cl_act(x, dtype):
    """[K,C,H,W] tensor of any layout/dtype -> dense channels-last Act [K,1,H,W,C] in ``dtype`` (zero-copy if possible)."""
    K, C, H, W = x.shape
    if _is_channels_last(x) and x.dtype == dtype:
        return Act(x.permute(0, 2, 3, 1).reshape(-1), K, 1, H, W, C)
    act = Act.empty(K, 1, H, W, C, dtype, x.device)
    if K:
        ops.nchw_to_nhwc(x.float().contiguous(), act)
    return act


# ----------------------------------------------------------------------------------------------------------------------
# MultiScaleRoIAlign
# ----------------------------------------------------------------------------------------------------------------------
class _RoiAlignFn(torch.autograd.Function):
    @staticmethod
    @ops.device_guard
    def forward(ctx, rois, levels, scales, P, sr, out_nchw, out_dtype, *feats):
        ops.device_check()
        N, C = feats[0].shape[:2]
        shapes = [tuple(f.shape[-2:]) for f in feats]
        cl = []
        for f in feats:                                    # channels-last f32 (the SlowFast output already is)
            if _is_channels_last(f) and f.dtype == torch.float32:
                cl.append(f.permute(0, 2, 3, 1))
            else:
                a = Act.empty(N, 1, f.shape[2], f.shape[3], C, torch.float32, f.device)
                ops.nchw_to_nhwc(f.float().contiguous(), a)
                cl.append(a.buf)
        K = rois.shape[0]
        if out_nchw:
            out = torch.empty(K, C, P, P, dtype=out_dtype, device=rois.device)
        else:
            out = torch.empty(K, P, P, C, dtype=out_dtype, device=rois.device)
        if K:
            ops.roi_align_fwd(cl, shapes, scales, N, C, rois, levels, P, sr, out, out_nchw)
        ctx.save_for_backward(rois, levels)
        ctx.meta = (shapes, scales, N, C, P, sr, out_nchw, [f.dtype for f in feats])
        return out if out_nchw else out.permute(0, 3, 1, 2)

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        rois, levels = ctx.saved_tensors
        shapes, scales, N, C, P, sr, out_nchw, dtypes = ctx.meta
        dfeats = [torch.zeros(N, h, w, C, dtype=torch.float32, device=g.device) for (h, w) in shapes]
        if rois.shape[0]:
            if out_nchw:
                gg = g if g.dtype in (torch.float32, torch.bfloat16) else g.float()
                gg = gg.contiguous()
            else:
                gl = g.permute(0, 2, 3, 1)
                gg = gl if gl.is_contiguous() else gl.contiguous()
            ops.roi_align_bwd(dfeats, shapes, scales, N, C, rois, levels, P, sr, gg, out_nchw)
        grads = tuple(d.permute(0, 3, 1, 2).to(dt) for d, dt in zip(dfeats, dtypes))
        return (None,) * 7 + grads


class MultiScaleRoIAlign(nn.Module):
    """Same arguments and call signature as torchvision.ops.MultiScaleRoIAlign (aligned=False legacy ROIAlign).
    ``out_layout="nchw"`` returns a contiguous f32 [K,C,P,P] (what the torch box head flattens);
    ``out_layout="nhwc"`` returns the same logical shape with channels_last strides in the activation dtype."""

    def __init__(self, featmap_names: List[str], output_size, sampling_ratio: int, *, canonical_scale: int = 224,
                 canonical_level: int = 4, out_layout: str = "nchw", precision: Optional[str] = None,
                 out_dtype: Optional[torch.dtype] = None):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert output_size[0] == output_size[1], "square pooling only"
        self.featmap_names = featmap_names
        self.sampling_ratio = sampling_ratio
        self.output_size = tuple(output_size)
        self.scales = None
        self.map_levels = None
        self.canonical_scale = canonical_scale
        self.canonical_level = canonical_level
        self.out_layout = out_layout
        self.precision = precision or _default_precision()
        # None: f32 for "nchw" (a torch consumer), the activation dtype for "nhwc"; "act": always the activation dtype of
        # ``self.precision`` (the native box head's [K, C*P*P] rows); or an explicit torch dtype
        self.out_dtype = out_dtype

    def _setup_scales(self, feats, image_shapes):
        max_h = max(s[0] for s in image_shapes)
        self.scales = [2.0 ** float(round(math.log2(float(f.shape[-2]) / float(max_h)))) for f in feats]
        self.k_min = int(-math.log2(self.scales[0]))
        self.k_max = int(-math.log2(self.scales[-1]))

    @ops.device_guard
    def _prepare(self, x, boxes, image_shapes):
        """-> (feature maps, rois [K,5] = (image idx, x1, y1, x2, y2), level ids)."""
        feats = [v for k, v in x.items() if k in self.featmap_names]
        if self.scales is None:
            self._setup_scales(feats, image_shapes)
        dev = feats[0].device
        ids = torch.cat([torch.full((b.shape[0], 1), float(i), dtype=torch.float32, device=dev) for i, b in enumerate(boxes)])
        rois = torch.cat([ids, torch.cat(boxes).to(device=dev, dtype=torch.float32)], dim=1).contiguous()
        levels = ops.roi_levels(rois, self.k_min, self.k_max) if len(feats) > 1 else None
        return feats, rois, levels

    def _out_spec(self):
        nchw = self.out_layout == "nchw"
        if self.out_dtype == "act":
            return nchw, _act_dtype(self.precision)
        if self.out_dtype is not None:
            return nchw, self.out_dtype
        return nchw, (torch.float32 if nchw else _act_dtype(self.precision))

    def forward(self, x: Dict[str, Tensor], boxes: List[Tensor], image_shapes: List[Tuple[int, int]]) -> Tensor:
        feats, rois, levels = self._prepare(x, boxes, image_shapes)
        nchw, out_dtype = self._out_spec()
        return _RoiAlignFn.apply(rois, levels, self.scales, self.output_size[0], self.sampling_ratio, nchw, out_dtype, *feats)


class _RoiAlignPairFn(torch.autograd.Function):
    """Two ROI poolings of the SAME feature maps (the box and the mask branch of a training step) as one autograd node:
    their backward passes scatter into ONE zero-filled set of gradient maps, instead of two sets that autograd then has
    to add (a fill plus a read-read-write pass over every pyramid level saved)."""

    @staticmethod
    @ops.device_guard
    def forward(ctx, specs, *feats):
        ops.device_check()
        ctx.set_materialize_grads(False)
        N, C = feats[0].shape[:2]
        shapes = [tuple(f.shape[-2:]) for f in feats]
        cl = []
        for f in feats:
            if _is_channels_last(f) and f.dtype == torch.float32:
                cl.append(f.permute(0, 2, 3, 1))
            else:
                a = Act.empty(N, 1, f.shape[2], f.shape[3], C, torch.float32, f.device)
                ops.nchw_to_nhwc(f.float().contiguous(), a)
                cl.append(a.buf)
        outs = []
        for rois, levels, scales, P, sr, nchw, out_dtype in specs:
            K = rois.shape[0]
            out = torch.empty((K, C, P, P) if nchw else (K, P, P, C), dtype=out_dtype, device=rois.device)
            if K:
                ops.roi_align_fwd(cl, shapes, scales, N, C, rois, levels, P, sr, out, nchw)
            outs.append(out if nchw else out.permute(0, 3, 1, 2))
        ctx.specs = specs
        ctx.meta = (shapes, N, C, [f.dtype for f in feats])
        return tuple(outs)

    @staticmethod
    @ops.device_guard
    def backward(ctx, *gs):
        shapes, N, C, dtypes = ctx.meta
        dev = next(g for g in gs if g is not None).device
        dfeats = [torch.zeros(N, h, w, C, dtype=torch.float32, device=dev) for (h, w) in shapes]
        for g, (rois, levels, scales, P, sr, nchw, _) in zip(gs, ctx.specs):
            if g is None or rois.shape[0] == 0:
                continue
            if nchw:
                gg = (g if g.dtype in (torch.float32, torch.bfloat16) else g.float()).contiguous()
            else:
                gl = g.permute(0, 2, 3, 1)
                gg = gl if gl.is_contiguous() else gl.contiguous()
            ops.roi_align_bwd(dfeats, shapes, scales, N, C, rois, levels, P, sr, gg, nchw)
        return (None,) + tuple(d.permute(0, 3, 1, 2).to(dt) for d, dt in zip(dfeats, dtypes))


def pool_pair(pool_a: "MultiScaleRoIAlign", pool_b: "MultiScaleRoIAlign", x, boxes_a, boxes_b, image_shapes):
    """``(pool_a(x, boxes_a, image_shapes), pool_b(x, boxes_b, image_shapes))`` with a shared backward (see _RoiAlignPairFn)."""
    feats, rois_a, lv_a = pool_a._prepare(x, boxes_a, image_shapes)
    feats_b, rois_b, lv_b = pool_b._prepare(x, boxes_b, image_shapes)
    if len(feats) != len(feats_b) or any(fa is not fb for fa, fb in zip(feats, feats_b)):
        return pool_a(x, boxes_a, image_shapes), pool_b(x, boxes_b, image_shapes)
    specs = []
    for pool, rois, lv in ((pool_a, rois_a, lv_a), (pool_b, rois_b, lv_b)):
        nchw, out_dtype = pool._out_spec()
        specs.append((rois, lv, pool.scales, pool.output_size[0], pool.sampling_ratio, nchw, out_dtype))
    return _RoiAlignPairFn.apply(tuple(specs), *feats)


# ----------------------------------------------------------------------------------------------------------------------
# mask head: 4 x (conv3x3 + bias + ReLU) on [K,256,14,14]
# ----------------------------------------------------------------------------------------------------------------------
_PACKED = {}      # id(parameter) -> (weakref to it, {key: (version, data_ptr, value)}); entries die with the parameter


def _cached(owner, key, make):
    """``make()`` cached per PARAMETER OBJECT until it is written to (``_version``) or re-allocated.  Keyed by object identity
    with a weak reference - not by address, which a new model's parameter can inherit from a freed one - and bypassed inside a
    CUDA-graph capture, where the packing kernels must be part of the graph (a replay has to see ITS step's weights)."""
    if not isinstance(owner, nn.Parameter) or (owner.is_cuda and torch.cuda.is_current_stream_capturing()):
        return make()
    ent = _PACKED.get(id(owner))
    if ent is None or ent[0]() is not owner:
        k = id(owner)
        ent = (weakref.ref(owner, lambda _r, k=k: _PACKED.pop(k, None)), {})
        _PACKED[k] = ent
    hit = ent[1].get(key)
    if hit is None or hit[0] != owner._version or hit[1] != owner.data_ptr():
        hit = (owner._version, owner.data_ptr(), make())
        ent[1][key] = hit
    return hit[2]


def _pack(w, mode, umma, kc, tap=(0, 0), owner=None):
    """Packed GEMM operand of a weight.  ``owner``: the nn.Parameter ``w`` is (a view of) - inference and evaluation call the
    heads with unchanged weights thousands of times, so the operand is cached on it."""
    cp = (kc + 63) // 64 * 64 if umma else kc
    owner = w if owner is None else owner
    return _cached(owner, ("pack", mode, umma, cp, tap), lambda: ops.pack_weights(w, mode, BF16 if umma else F32, cp, tap)), cp


class _MaskHeadFn(torch.autograd.Function):
    @staticmethod
    @ops.device_guard
    def forward(ctx, x, precision, *wb):
        ops.device_check()
        umma = precision != "fp32"
        dt_act = _act_dtype(precision)
        K, C, H, W = x.shape
        cur = _to_cl_act(x, dt_act)
        acts = [cur]
        for i in range(len(wb) // 2):
            w, b = wb[2 * i], wb[2 * i + 1]
            wp, cp = _pack(w, 0, umma, w.shape[1])
            y = Act.empty(K, 1, H, W, w.shape[0], dt_act, x.device)
            if K:
                ops.conv(cur, wp, cp, w.shape[0], (1, 3, 3), (0, 1, 1), 1, y, umma=umma, relu=True, shift=b.detach().float())
            acts.append(y)
            cur = y
        ctx.acts, ctx.umma, ctx.dt_act, ctx.x_dtype = acts, umma, dt_act, x.dtype
        ctx.save_for_backward(*wb)
        return _nchw_view(cur.buf, K, H, W, cur.C)

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        wb = ctx.saved_tensors
        acts, umma, dt_act = ctx.acts, ctx.umma, ctx.dt_act
        K, C, H, W = g.shape
        dy = _to_cl_act(g, g.dtype if g.dtype in (torch.float32, torch.bfloat16) else torch.float32)
        if dt_act == torch.float32 and dy.dtype != torch.float32:
            dy = _to_cl_act(g.float(), torch.float32)
        grads = [None] * len(wb)
        n = len(wb) // 2
        dconv = db = None          # gradient wrt the pre-activation output of layer i (and its bias gradient), if already known
        for i in range(n - 1, -1, -1):
            w, b = wb[2 * i], wb[2 * i + 1]
            x_in, y = acts[i], acts[i + 1]
            gw = ops.new_grad(w)
            if dconv is None:
                dconv = Act.empty(K, 1, H, W, y.C, dt_act, g.device)
                db = ops.new_grad(b)
                if K:
                    ops.relu_bwd(dy, y, dconv, db)
            if K:
                dwp = torch.zeros(9 * x_in.C * y.C, dtype=torch.float32, device=g.device)
                ops.wgrad(x_in, dconv, (1, 3, 3), (0, 1, 1), dwp, umma=umma)
                ops.unpack_wgrad(dwp, gw, 0)
            grads[2 * i], grads[2 * i + 1] = gw.to(w.dtype), db.to(b.dtype)
            if i > 0 and umma and K:
                # data gradient with the ReLU backward of the layer below fused into the epilogue: the result IS that layer's
                # pre-activation gradient (bf16) and its bias gradient - no f32 round trip, no separate relu_bwd pass
                wd, cpd = _pack(w, 1, umma, w.shape[0])
                nxt = Act.empty(K, 1, H, W, x_in.C, dt_act, g.device)
                db = ops.new_grad(wb[2 * i - 1])
                ops.conv(dconv, wd, cpd, x_in.C, (1, 3, 3), (0, 1, 1), 1, nxt, umma=True, relu_mask=x_in, dbias=db)
                dconv = nxt
            elif i > 0 or ctx.needs_input_grad[0]:
                wd, cpd = _pack(w, 1, umma, w.shape[0])
                last = i == 0
                dx = Act.empty(K, 1, H, W, x_in.C, (dt_act if last else torch.float32), g.device)
                if K:
                    ops.conv(dconv, wd, cpd, x_in.C, (1, 3, 3), (0, 1, 1), 1, dx, umma=umma)
                dy = dx
                dconv = db = None
        gx = None
        if ctx.needs_input_grad[0]:
            gx = _nchw_view(dy.buf, K, H, W, dy.C).to(ctx.x_dtype)
        ctx.acts = None
        return (gx, None) + tuple(grads)


class MaskRCNNHeads(tv_mask_rcnn.MaskRCNNHeads):
    """torchvision's container (same ctor, same state_dict keys ``{i}.0.weight|bias``); forward on libsfvos."""

    precision = None

    def forward(self, x):
        wb = []
        for blk in self:
            conv = blk[0]
            assert conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.dilation == (1, 1), "3x3 pad-1 convs only"
            assert len(blk) == 2 and isinstance(blk[1], nn.ReLU), "norm layers in the mask head are not supported"
            wb.extend([conv.weight, conv.bias])
        return _MaskHeadFn.apply(x, self.precision or _default_precision(), *wb)


# ----------------------------------------------------------------------------------------------------------------------
# mask predictor: ConvTranspose2d(k2,s2)+ReLU -> 1x1 conv to n_cls logits
# ----------------------------------------------------------------------------------------------------------------------
_TAPS = [(0, 0), (0, 1), (1, 0), (1, 1)]


class _MaskPredictorFn(torch.autograd.Function):
    """ConvTranspose2d(k2, s2) + ReLU -> 1x1 conv to n_cls logits.  The four taps of the transposed convolution are ONE 1x1
    convolution C -> 4*co whose output stays in 2x2 space-to-depth order ([K,H,W,(tap, co)]: every low-res pixel holds its
    2x2 block of high-res pixels); the logits kernels read it in that order (``pixel_order=1``) and write the reference's raster
    order.  Backward: the weight gradient of all taps is one GEMM (N = 4*co), the data gradient one GEMM with K = 4*co -
    no per-tap launches, no read-modify-write of the input gradient."""

    @staticmethod
    def _fprop_weights(wt, umma, C):
        """ConvTranspose2d weight [C, co, 2, 2] -> operand of the 1x1 convolution C -> 4*co (output channel = tap*co + n)."""
        def make():
            parts = [_pack(wt, 2, umma, C, tap)[0] for tap in _TAPS]     # umma: bf16 [co][C] each; simt: f32 [C][co]
            return torch.cat(parts, dim=0) if umma else torch.cat(parts, dim=1).contiguous()
        return _cached(wt, ("convt_fprop", umma, C), make)

    @staticmethod
    def _dgrad_weights(wt, umma, co):
        """-> operand of the 1x1 convolution 4*co -> C of the data gradient (input channel = tap*co + n)."""
        def make():
            parts = [_pack(wt, 3, umma, co, tap)[0] for tap in _TAPS]    # umma: bf16 [C][co] each; simt: f32 [co][C]
            return torch.cat(parts, dim=1).contiguous() if umma else torch.cat(parts, dim=0)
        return _cached(wt, ("convt_dgrad", umma, co), make)

    @staticmethod
    @ops.device_guard
    def forward(ctx, x, precision, wt, bt, wl, bl):
        ops.device_check()
        umma = precision != "fp32"
        dt_act = _act_dtype(precision)
        K, C, H, W = x.shape
        assert H == W, f"MaskRCNNPredictor: square ROI features only (the logits kernels take one side length), got {H}x{W}"
        co = wt.shape[1]
        n_cls = wl.shape[0]
        assert not umma or (C % 64 == 0 and co % 64 == 0), "conv5_mask channels must be multiples of 64 on the tensor-core path"
        xin = _to_cl_act(x, dt_act)
        up = Act.empty(K, 1, H, W, 4 * co, dt_act, x.device)             # space-to-depth: channel = tap * co + c
        logits = torch.empty(K, n_cls, 2 * H, 2 * W, dtype=torch.float32, device=x.device)
        if K:
            wf = _MaskPredictorFn._fprop_weights(wt, umma, C)
            ops.conv(xin, wf, C, 4 * co, (1, 1, 1), (0, 0, 0), 1, up, umma=umma, relu=True, shift=bt.detach().float().repeat(4))
            call("sfvos_mask_logits_fwd", up.ptr(), ops.dt(up.buf), _p(wl.detach().float().contiguous()),
                 _p(bl.detach().float().contiguous()), _p(logits), K, 2 * H, co, n_cls, 1, stream())
        ctx.acts = (xin, up)
        ctx.meta = (umma, dt_act, x.dtype, K, C, H, W, co, n_cls)
        ctx.save_for_backward(wt, bt, wl, bl)
        return logits

    @staticmethod
    @ops.device_guard
    def backward(ctx, glogits):
        wt, bt, wl, bl = ctx.saved_tensors
        xin, up = ctx.acts
        umma, dt_act, x_dtype, K, C, H, W, co, n_cls = ctx.meta
        dev = glogits.device
        gwl = ops.new_grad(wl, (n_cls, co))
        gbl = ops.new_grad(bl)
        gwt = ops.new_grad(wt)
        gbt = ops.new_grad(bt)
        gx = None
        if K:
            gl = glogits.float().contiguous()
            # logits-layer backward and the ReLU backward of the ConvTranspose output in ONE pass over ``up``
            dconv = Act.empty(K, 1, H, W, 4 * co, dt_act, dev)
            call("sfvos_mask_logits_relu_bwd", up.ptr(), ops.dt(up.buf), _p(wl.detach().float().contiguous()), _p(gl),
                 dconv.ptr(), ops.dt(dconv.buf), _p(gwl), _p(gbl), _p(gbt), K, 2 * H, co, n_cls, 1, stream())
            # weight gradient of the four taps: dw[c][tap*co + n]
            dwp = torch.zeros(C * 4 * co, dtype=torch.float32, device=dev)
            ops.wgrad(xin, dconv, (1, 1, 1), (0, 0, 0), dwp, umma=umma)
            dw4 = dwp.view(C, 4, co)
            for n, (i, j) in enumerate(_TAPS):
                ops.unpack_wgrad(dw4[:, n].contiguous(), gwt, 2, (i, j))
            if ctx.needs_input_grad[0]:
                wb = _MaskPredictorFn._dgrad_weights(wt, umma, co)
                dx = Act.empty(K, 1, H, W, C, dt_act, dev)
                ops.conv(dconv, wb, 4 * co, C, (1, 1, 1), (0, 0, 0), 1, dx, umma=umma)
                gx = _nchw_view(dx.buf, K, H, W, C).to(x_dtype)
        elif ctx.needs_input_grad[0]:
            gx = torch.zeros(K, C, H, W, dtype=x_dtype, device=dev)
        ctx.acts = None
        return gx, None, gwt.to(wt.dtype), gbt.to(bt.dtype), gwl.view_as(wl).to(wl.dtype), gbl.to(bl.dtype)


class MaskRCNNPredictor(tv_mask_rcnn.MaskRCNNPredictor):
    """Same ctor / state_dict keys (conv5_mask.*, mask_fcn_logits.*) as torchvision; forward on libsfvos."""

    precision = None

    def forward(self, x):
        return _MaskPredictorFn.apply(x, self.precision or _default_precision(), self.conv5_mask.weight, self.conv5_mask.bias,
                                      self.mask_fcn_logits.weight, self.mask_fcn_logits.bias)


# ----------------------------------------------------------------------------------------------------------------------
# loss / inference
# ----------------------------------------------------------------------------------------------------------------------
class _MaskBceFn(torch.autograd.Function):
    @staticmethod
    @ops.device_guard
    def forward(ctx, logits, labels, targets):
        K, n_cls, S, _ = logits.shape
        logits = logits.float().contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        call("sfvos_mask_bce_fwd", _p(logits), _p(labels), _p(targets), _p(loss), K, S, n_cls, stream())
        ctx.save_for_backward(logits, labels, targets)
        return loss[0]

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        logits, labels, targets = ctx.saved_tensors
        K, n_cls, S, _ = logits.shape
        gl = torch.empty_like(logits)
        gg = g.reshape(1).float().contiguous()
        call("sfvos_mask_bce_bwd", _p(logits), _p(labels), _p(targets), _p(gg), _p(gl), K, S, n_cls, stream())
        return gl, None, None


@ops.device_guard
def project_masks_on_boxes(gt_masks, boxes, matched_idxs, M):
    """TV roi_heads.py:85-97 on the GPU: adaptive-sampling ROIAlign of the u8 GT masks to M x M."""
    rois = torch.cat([matched_idxs[:, None].to(boxes), boxes], dim=1).float().contiguous()
    masks = gt_masks.to(torch.uint8).contiguous()
    return ops.mask_targets(masks, rois, M)


def maskrcnn_loss(mask_logits, proposals, gt_masks, gt_labels, mask_matched_idxs):
    """Same signature and value as torchvision's maskrcnn_loss (TV roi_heads.py:100-129)."""
    M = mask_logits.shape[-1]
    labels = torch.cat([gl[idxs] for gl, idxs in zip(gt_labels, mask_matched_idxs)], dim=0)
    if len(gt_masks) > 1 and all(m.shape[1:] == gt_masks[0].shape[1:] for m in gt_masks):
        # same-sized images: one launch for the whole batch (object indices offset into the concatenated masks)
        offs, n = [], 0
        for m in gt_masks:
            offs.append(n)
            n += m.shape[0]
        idx = torch.cat([i + o for i, o in zip(mask_matched_idxs, offs)])
        targets = project_masks_on_boxes(torch.cat(gt_masks), torch.cat(proposals), idx, M)
    else:
        targets = torch.cat([project_masks_on_boxes(m, p, i, M) for m, p, i in zip(gt_masks, proposals, mask_matched_idxs)], dim=0)
    if targets.numel() == 0:
        return mask_logits.sum() * 0
    return _MaskBceFn.apply(mask_logits, labels.to(torch.int64).contiguous(), targets.contiguous())


@ops.device_guard
def maskrcnn_inference(x, labels):
    """TV roi_heads.py:56-82: sigmoid of the predicted-class channel, split per image."""
    per_img = [lab.shape[0] for lab in labels]
    lab = torch.cat(labels).to(torch.int64).contiguous()
    K, n_cls, S, _ = x.shape
    prob = torch.empty(K, 1, S, S, dtype=torch.float32, device=x.device)
    if K:
        call("sfvos_mask_probs", _p(x.float().contiguous()), _p(lab), _p(prob), K, S, n_cls, stream())
    return prob.split(per_img, dim=0)



# ----------------------------------------------------------------------------------------------------------------------
# box head: fc6 (C*P*P -> 1024) + ReLU, fc7 (1024 -> 1024) + ReLU; predictor: cls_score / bbox_pred; fastrcnn_loss
# A Linear layer over M ROIs is the 1x1x1 convolution of a [1,1,1,M] "image" with in_features channels: the same
# tcgen05 kernels (N tiled in 256-column chunks), the same weight-gradient GEMM, the same ReLU-backward pass.
# ----------------------------------------------------------------------------------------------------------------------
def _rows_act(t2d):
    """[M, C] contiguous tensor -> Act over M "pixels" (B=T=H=1, W=M)."""
    M, C = t2d.shape
    return Act(t2d.reshape(-1), 1, 1, 1, M, C)


def _linear_fwd(x, w, b, umma, dt_act, relu, out_dtype=None):
    """x: Act [M,K]; w [N,K] f32 parameter (N a multiple of 32 <= 256 or of 256); -> Act [M,N]."""
    N, K = w.shape
    wp, cp = _pack(w.view(N, K, 1, 1, 1), 0, umma, K, owner=w)
    y = Act.empty(1, 1, 1, x.W, N, out_dtype or dt_act, x.buf.device)
    if x.W:
        ops.conv(x, wp, cp, N, (1, 1, 1), (0, 0, 0), 1, y, umma=umma, relu=relu, shift=b)
    return y


def _linear_bwd(x, dy, w, umma, dx_dtype, need_dx=True):
    """dy = gradient of the PRE-activation output [M,N] (Act); returns (gw [N,K] f32, dx Act [M,K] or None)."""
    N, K = w.shape
    dev = dy.buf.device
    gw = ops.new_grad(w)
    dx = None
    if x.W:
        dwp = torch.zeros(K * N, dtype=torch.float32, device=dev)
        ops.wgrad(x, dy, (1, 1, 1), (0, 0, 0), dwp, umma=umma)
        ops.unpack_wgrad(dwp, gw.view(N, K, 1, 1, 1), 0)
    if need_dx:
        dx = Act.empty(1, 1, 1, x.W, K, dx_dtype, dev)
        if x.W:
            wd, cpd = _pack(w.view(N, K, 1, 1, 1), 1, umma, N, owner=w)
            ops.conv(dy, wd, cpd, K, (1, 1, 1), (0, 0, 0), 1, dx, umma=umma)
    return gw, dx


def _as_rows(x, dt_act):
    """Any [M, ...] tensor -> dense Act [M, features] in the activation dtype."""
    x2 = x.flatten(start_dim=1)
    if x2.dtype != dt_act or not x2.is_contiguous():
        x2 = x2.to(dt_act).contiguous()
    return _rows_act(x2)


class _BoxHeadFn(torch.autograd.Function):
    @staticmethod
    @ops.device_guard
    def forward(ctx, x, precision, w6, b6, w7, b7):
        ops.device_check()
        umma = precision != "fp32"
        dt_act = _act_dtype(precision)
        xin = _as_rows(x, dt_act)
        y6 = _linear_fwd(xin, w6, b6.detach().float(), umma, dt_act, True)
        y7 = _linear_fwd(y6, w7, b7.detach().float(), umma, dt_act, True)
        ctx.acts = (xin, y6, y7)
        ctx.meta = (umma, dt_act, x.dtype, tuple(x.shape))
        ctx.save_for_backward(w6, b6, w7, b7)
        return y7.buf.view(xin.W, w7.shape[0])

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        w6, b6, w7, b7 = ctx.saved_tensors
        xin, y6, y7 = ctx.acts
        umma, dt_act, x_dtype, x_shape = ctx.meta
        M, dev = xin.W, g.device
        gb6, gb7 = ops.new_grad(b6), ops.new_grad(b7)
        dy7 = _as_rows(g, dt_act)
        dc7 = Act.empty(1, 1, 1, M, y7.C, dt_act, dev)
        if M:
            ops.relu_bwd(dy7, y7, dc7, gb7)
        gw7, dy6 = _linear_bwd(y6, dc7, w7, umma, dt_act)
        dc6 = Act.empty(1, 1, 1, M, y6.C, dt_act, dev)
        if M:
            ops.relu_bwd(dy6, y6, dc6, gb6)
        gw6, dx = _linear_bwd(xin, dc6, w6, umma, dt_act, need_dx=ctx.needs_input_grad[0])
        gx = dx.buf.view(x_shape).to(x_dtype) if dx is not None else None
        ctx.acts = None
        return gx, None, gw6.to(w6.dtype), gb6.to(b6.dtype), gw7.to(w7.dtype), gb7.to(b7.dtype)


class TwoMLPHead(tv_faster_rcnn.TwoMLPHead):
    """Same ctor / state_dict keys (fc6.*, fc7.*) as torchvision; flatten -> fc6 -> ReLU -> fc7 -> ReLU on libsfvos."""

    precision = None

    def forward(self, x):
        return _BoxHeadFn.apply(x, self.precision or _default_precision(), self.fc6.weight, self.fc6.bias,
                                self.fc7.weight, self.fc7.bias)


def _pred_pad(n):
    return 64 if n <= 64 else (n + 255) // 256 * 256 if n > 256 else (n + 31) // 32 * 32


class _BoxPredictorFn(torch.autograd.Function):
    """cls_score and bbox_pred as ONE GEMM: rows [0,n_cls) of the stacked weight are the class logits, rows
    [n_cls, 5 n_cls) the box deltas; zero rows pad the output to a tensor-core tile.  Returns f32 [M, Npad]."""

    @staticmethod
    @ops.device_guard
    def forward(ctx, x, precision, wc, bc, wb, bb):
        ops.device_check()
        umma = precision != "fp32"
        dt_act = _act_dtype(precision)
        n_out = wc.shape[0] + wb.shape[0]
        n_pad = _pred_pad(n_out)
        K = wc.shape[1]

        def stack():
            w = torch.zeros(n_pad, K, dtype=torch.float32, device=x.device)
            w[:wc.shape[0]] = wc.detach()
            w[wc.shape[0]:n_out] = wb.detach()
            b = torch.zeros(n_pad, dtype=torch.float32, device=x.device)
            b[:wc.shape[0]] = bc.detach()
            b[wc.shape[0]:n_out] = bb.detach()
            return w, b
        w, b = _cached(wc, ("stack", wb.data_ptr(), wb._version, bc.data_ptr(), bc._version, bb.data_ptr(), bb._version, str(x.device)), stack)
        xin = _as_rows(x, dt_act)
        y = _linear_fwd(xin, w, b, umma, dt_act, False, out_dtype=torch.float32)
        ctx.acts = (xin, w)
        ctx.meta = (umma, dt_act, x.dtype, tuple(x.shape), n_pad)
        ctx.save_for_backward(wc, bc, wb, bb)
        return y.buf.view(xin.W, n_pad)

    @staticmethod
    @ops.device_guard
    def backward(ctx, g):
        wc, bc, wb, bb = ctx.saved_tensors
        xin, w = ctx.acts
        umma, dt_act, x_dtype, x_shape, n_pad = ctx.meta
        M, dev = xin.W, g.device
        nc, nb = wc.shape[0], wb.shape[0]
        g32 = _rows_act(g.float().contiguous())
        stats = torch.zeros(2, n_pad, dtype=torch.float64, device=dev)     # row 0 = column sums = the bias gradients
        if M:
            ops.channel_stats(g32, stats.view(-1))
        if umma:
            dy = Act.empty(1, 1, 1, M, n_pad, dt_act, dev)
            if M:
                ops.affine_act(g32, dy, torch.ones(n_pad, device=dev), torch.zeros(n_pad, device=dev), False)
        else:
            dy = g32
        gw, dx = _linear_bwd(xin, dy, w, umma, dt_act, need_dx=ctx.needs_input_grad[0])
        gx = dx.buf.view(x_shape).to(x_dtype) if dx is not None else None
        ctx.acts = None
        # the four parameter gradients are row ranges of the stacked GEMM's: added into their own accumulators
        outs = []
        for prm, val in ((wc, gw[:nc]), (bc, stats[0, :nc]), (wb, gw[nc:nc + nb]), (bb, stats[0, nc:nc + nb])):
            acc = ops.new_grad(prm)
            acc.add_(val.to(torch.float32))
            outs.append(acc.to(prm.dtype))
        return (gx, None) + tuple(outs)


class FastRCNNPredictor(tv_faster_rcnn.FastRCNNPredictor):
    """Same ctor / state_dict keys (cls_score.*, bbox_pred.*) as torchvision; returns (scores [M,n_cls], bbox_deltas
    [M,4 n_cls]) as column slices of one fused f32 GEMM output."""

    precision = None

    def forward(self, x):
        if x.dim() == 4:
            assert list(x.shape[2:]) == [1, 1], f"x has the wrong shape, expecting the last two dimensions to be [1,1] instead of {list(x.shape[2:])}"
        fused = _BoxPredictorFn.apply(x, self.precision or _default_precision(), self.cls_score.weight, self.cls_score.bias,
                                      self.bbox_pred.weight, self.bbox_pred.bias)
        nc, nb = self.cls_score.out_features, self.bbox_pred.out_features
        return fused[:, :nc], fused[:, nc:nc + nb]


class _FastRCNNLossFn(torch.autograd.Function):
    @staticmethod
    @ops.device_guard
    def forward(ctx, class_logits, box_regression, labels, regression_targets, beta):
        ops.device_check()
        M, n_cls = class_logits.shape

        def rows(t):   # f32 rows with unit column stride (column slices of the fused predictor output pass as they are)
            return t if (t.dtype == torch.float32 and t.stride(1) == 1) else t.float().contiguous()
        cls, box = rows(class_logits), rows(box_regression)
        losses = torch.empty(2, dtype=torch.float32, device=cls.device)
        call("sfvos_fastrcnn_loss_fwd", _p(cls), cls.stride(0), _p(box), box.stride(0), _p(labels), _p(regression_targets), M,
             n_cls, float(beta), _p(losses), stream())
        ctx.save_for_backward(cls, box, labels, regression_targets)
        ctx.beta = float(beta)
        ctx.dtypes = (class_logits.dtype, box_regression.dtype)
        return losses[0].clone(), losses[1].clone()

    @staticmethod
    @ops.device_guard
    def backward(ctx, g_cls, g_box):
        cls, box, labels, tgt = ctx.saved_tensors
        M, n_cls = cls.shape
        gl = torch.zeros(2, dtype=torch.float32, device=cls.device)
        if g_cls is not None:
            gl[0] = g_cls
        if g_box is not None:
            gl[1] = g_box
        dcls = torch.empty(M, n_cls, dtype=torch.float32, device=cls.device)
        dbox = torch.empty(M, 4 * n_cls, dtype=torch.float32, device=cls.device)
        call("sfvos_fastrcnn_loss_bwd", _p(cls), cls.stride(0), _p(box), box.stride(0), _p(labels), _p(tgt), _p(gl), M, n_cls,
             ctx.beta, _p(dcls), dcls.stride(0), _p(dbox), dbox.stride(0), stream())
        return dcls.to(ctx.dtypes[0]), dbox.to(ctx.dtypes[1]), None, None, None


def fastrcnn_loss(class_logits, box_regression, labels, regression_targets):
    """Same signature and values as torchvision's fastrcnn_loss (TV roi_heads.py:12-53): (cross-entropy over the class
    logits, smooth-L1(beta=1/9, sum) / M over the matched class's deltas of the positive ROIs)."""
    labels = torch.cat(labels, dim=0).to(torch.int64).contiguous()
    regression_targets = torch.cat(regression_targets, dim=0).float().contiguous()
    if labels.numel() == 0:
        return tv_roi_heads.fastrcnn_loss(class_logits, box_regression, [labels], [regression_targets])
    return _FastRCNNLossFn.apply(class_logits, box_regression, labels, regression_targets, 1.0 / 9)


# ----------------------------------------------------------------------------------------------------------------------
# RoIHeads
# ----------------------------------------------------------------------------------------------------------------------
class RoIHeads(tv_roi_heads.RoIHeads):
    """torchvision RoIHeads with both branches on libsfvos kernels: box pool -> fc6/fc7 -> predictor -> fastrcnn_loss and
    mask pool -> head -> predictor -> loss/inference.  Box sampling, box decoding and NMS are inherited unchanged."""

    def forward(self, features, proposals, image_shapes, targets=None):
        if self.training:
            proposals, matched_idxs, labels, regression_targets = self.select_training_samples(proposals, targets)
        else:
            labels = regression_targets = matched_idxs = None
        mask_features = mask_proposals = pos_matched_idxs = None
        if (self.training and self.has_mask() and isinstance(self.box_roi_pool, MultiScaleRoIAlign)
                and isinstance(self.mask_roi_pool, MultiScaleRoIAlign)):
            # training: the mask branch's proposals (the positives) are known up front, so both poolings share one
            # autograd node and one set of gradient maps
            mask_proposals, pos_matched_idxs = [], []
            for img_id in range(len(proposals)):
                pos = torch.where(labels[img_id] > 0)[0]
                mask_proposals.append(proposals[img_id][pos])
                pos_matched_idxs.append(matched_idxs[img_id][pos])
            box_features, mask_features = pool_pair(self.box_roi_pool, self.mask_roi_pool, features, proposals, mask_proposals,
                                                    image_shapes)
        else:
            box_features = self.box_roi_pool(features, proposals, image_shapes)
        box_features = self.box_head(box_features)
        class_logits, box_regression = self.box_predictor(box_features)

        result: List[Dict[str, Tensor]] = []
        losses = {}
        if self.training:
            loss_classifier, loss_box_reg = fastrcnn_loss(class_logits, box_regression, labels, regression_targets)
            losses = {"loss_classifier": loss_classifier, "loss_box_reg": loss_box_reg}
        else:
            boxes, scores, labels = self.postprocess_detections(class_logits, box_regression, proposals, image_shapes)
            for i in range(len(boxes)):
                result.append({"boxes": boxes[i], "labels": labels[i], "scores": scores[i]})

        if self.has_mask():
            if mask_features is None:
                mask_proposals = [p["boxes"] for p in result]
                if self.training:
                    mask_proposals, pos_matched_idxs = [], []
                    for img_id in range(len(proposals)):
                        pos = torch.where(labels[img_id] > 0)[0]
                        mask_proposals.append(proposals[img_id][pos])
                        pos_matched_idxs.append(matched_idxs[img_id][pos])
                mask_features = self.mask_roi_pool(features, mask_proposals, image_shapes)
            mask_features = self.mask_head(mask_features)
            mask_logits = self.mask_predictor(mask_features)
            if self.training:
                gt_masks = [t["masks"] for t in targets]
                gt_labels = [t["labels"] for t in targets]
                losses["loss_mask"] = maskrcnn_loss(mask_logits, mask_proposals, gt_masks, gt_labels, pos_matched_idxs)
            else:
                labels = [r["labels"] for r in result]
                for mask_prob, r in zip(maskrcnn_inference(mask_logits, labels), result):
                    r["masks"] = mask_prob
        return result, losses


# ----------------------------------------------------------------------------------------------------------------------
# mask paste-back (the step after the path: code/helpers/model.py:347 -> transform.postprocess)
# ----------------------------------------------------------------------------------------------------------------------
@ops.device_guard
def paste_masks_in_image(masks, boxes, img_shape, padding=1):
    """Same signature and values as torchvision's paste_masks_in_image (TV roi_heads.py:474-501): masks [K,1,M,M], boxes
    [K,4] in output-image pixels -> [K,1,im_h,im_w].  One kernel for all K masks instead of a Python loop per mask."""
    ops.device_check()
    im_h, im_w = int(img_shape[0]), int(img_shape[1])
    K, M = masks.shape[0], masks.shape[-1]
    out = torch.empty(K, 1, im_h, im_w, dtype=torch.float32, device=masks.device)
    if K:
        call("sfvos_paste_masks", _p(masks.float().contiguous()), _p(boxes.float().contiguous()), K, M, int(padding), im_h, im_w,
             _p(out), stream())
    return out.to(masks.dtype)


def postprocess(result, image_shapes, original_image_sizes, to_cpu=False):
    """GeneralizedRCNNTransform.postprocess in eval mode (TV transform.py:257-279): boxes rescaled to the original image,
    masks pasted back.  Images that share one original size (all frames of a sequence, model.py:342) are pasted in ONE launch.
    ``to_cpu=True`` also performs the reference's move of the detections to the host (model.py:348) -- the pasted masks of a
    group cross PCIe as ONE asynchronous copy into pinned memory instead of one pageable copy per tensor."""
    from torchvision.models.detection.transform import resize_boxes
    for pred, im_s, o_im_s in zip(result, image_shapes, original_image_sizes):
        pred["boxes"] = resize_boxes(pred["boxes"], im_s, o_im_s)
    groups = {}
    for i, (pred, o_im_s) in enumerate(zip(result, original_image_sizes)):
        if "masks" in pred:
            groups.setdefault((int(o_im_s[0]), int(o_im_s[1])), []).append(i)
    pending = []
    for size, idxs in groups.items():
        counts = [result[i]["masks"].shape[0] for i in idxs]
        pasted = paste_masks_in_image(torch.cat([result[i]["masks"] for i in idxs]), torch.cat([result[i]["boxes"] for i in idxs]), size)
        if to_cpu and pasted.is_cuda:
            host = torch.empty(pasted.shape, dtype=pasted.dtype, pin_memory=True)     # caching host allocator: reused across chunks
            host.copy_(pasted, non_blocking=True)
            pending.append(pasted)                                                    # keep the source alive until the sync below
            pasted = host
        for i, part in zip(idxs, pasted.split(counts)):
            result[i]["masks"] = part
    if to_cpu:
        for pred in result:
            for k in list(pred.keys()):
                if torch.is_tensor(pred[k]) and pred[k].is_cuda:
                    pred[k] = pred[k].cpu()
        if pending:
            torch.cuda.current_stream().synchronize()
    return result


def install(roi_heads: tv_roi_heads.RoIHeads, precision: Optional[str] = None):
    """Swap the libsfvos modules into an existing torchvision ``roi_heads`` IN PLACE, keeping every parameter
    (same tensors, same state_dict keys, same registration order).  Returns the same object, now a ``RoIHeads``."""
    precision = precision or _default_precision()
    native_box = type(roi_heads.box_head) in (tv_faster_rcnn.TwoMLPHead, TwoMLPHead)
    for name, layout in (("box_roi_pool", "nchw"), ("mask_roi_pool", "nhwc")):
        old = getattr(roi_heads, name)
        if old is None:
            continue
        # the native box head consumes the [K, C*P*P] rows (torchvision's flatten order) in the activation dtype
        od = "act" if (name == "box_roi_pool" and native_box) else None
        new = MultiScaleRoIAlign(list(old.featmap_names), old.output_size, old.sampling_ratio,
                                 canonical_scale=old.canonical_scale, canonical_level=old.canonical_level,
                                 out_layout=layout, precision=precision, out_dtype=od)
        setattr(roi_heads, name, new)
    if native_box:
        roi_heads.box_head.__class__ = TwoMLPHead
        roi_heads.box_head.precision = precision
    if type(roi_heads.box_predictor) in (tv_faster_rcnn.FastRCNNPredictor, FastRCNNPredictor):
        roi_heads.box_predictor.__class__ = FastRCNNPredictor
        roi_heads.box_predictor.precision = precision
    if roi_heads.mask_head is not None:
        roi_heads.mask_head.__class__ = MaskRCNNHeads
        roi_heads.mask_head.precision = precision
    if roi_heads.mask_predictor is not None:
        roi_heads.mask_predictor.__class__ = MaskRCNNPredictor
        roi_heads.mask_predictor.precision = precision
    roi_heads.__class__ = RoIHeads
    return roi_heads
