"""The step BEFORE the hot path: the feature-pyramid half of the frozen backbone and the RPN head on
libsfvos.so, behind torchvision's own module interfaces, producing the channels-last bf16 features the temporal module
consumes without a layout pass.

  FeaturePyramidNetwork <- torchvision.ops.FeaturePyramidNetwork        (TV/ops/feature_pyramid_network.py)
                           lateral 1x1 convs C_l -> 256, top-down nearest-upsample + add, 3x3 output convs, LastLevelMaxPool
  RPNHead               <- torchvision...rpn.RPNHead                    (TV/models/detection/rpn.py)
                           3x3 conv + ReLU, then cls_logits (A) and bbox_pred (4A) as ONE 1x1 GEMM
  AnchorGenerator       <- torchvision...anchor_utils.AnchorGenerator   (anchors stay f32 whatever the feature dtype)
  install_backbone(maskrcnn_model)   swaps them in IN PLACE (same parameters / state_dict keys), like roi_heads.install

Reference call sites: code/helpers/model.py:204 (``self.maskrcnn_model.backbone(batch_imgs)``) and :236-240
(``self.maskrcnn_model.rpn(...)``); both are frozen (model.py:176-179), so there is no backward here and the packed weights
are cached until a parameter changes.

Per 480x854 frame this is 219 of the ~390 GFLOP of backbone + RPN (FPN output convs 101, RPN conv 101, laterals 16, heads
0.7); the ResNet-50 body (stride-2 convolutions, 7x7 stem, max-pool) stays torchvision.  Anchor generation, box decoding
and NMS stay torchvision."""
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor
from torchvision.models.detection import anchor_utils as tv_anchor_utils
from torchvision.models.detection import rpn as tv_rpn
from torchvision.models import _utils as tv_utils
from torchvision.ops import feature_pyramid_network as tv_fpn

from . import ops
from ._lib import BF16, F32, call
from .ops import Act, _p, stream
from .roi_heads import _act_dtype, _default_precision, _nchw_view, _to_cl_act

def _packed(owner, name, w, umma):
    """[Cout,Cin,kh,kw] parameter -> fprop GEMM operand, cached ON THE OWNING MODULE until the parameter is written to or
    replaced (backbone and RPN are frozen, code/helpers/model.py:176-179).  The cache lives and dies with the module: a
    process-wide cache keyed by address would hand a new model the operands of a freed one."""
    cin = w.shape[1]
    cp = (cin + 63) // 64 * 64 if umma else cin
    capturing = w.is_cuda and torch.cuda.is_current_stream_capturing()
    cache = owner.__dict__.setdefault("_sfvos_packed", {})
    tag = (w.data_ptr(), w._version, str(w.device), tuple(w.shape), umma)
    hit = cache.get(name)
    if hit is None or hit[0] != tag or capturing:
        hit = (tag, ops.pack_weights(w, 0, BF16 if umma else F32, cp), cp)
        if not capturing:
            cache[name] = hit
    return hit[1], hit[2]


def _conv(owner, name, x, w, b, umma, out_dtype, relu=False):
    """x: channels-last Act [N,1,H,W,Cin]; w [Cout,Cin,k,k] (k = 1 or 3, stride 1, 'same' padding), b [Cout] -> Act."""
    cout, _, kh, kw = w.shape
    assert kh == kw and kh in (1, 3), "1x1 and 3x3 convolutions only"
    wp, cp = _packed(owner, name, w, umma)
    y = Act.empty(x.B, 1, x.H, x.W, cout, out_dtype, x.buf.device)
    if x.npix:
        ops.conv(x, wp, cp, cout, (1, kh, kw), (0, kh // 2, kw // 2), 1, y, umma=umma, relu=relu,
                 shift=None if b is None else b.detach().float())
    return y


# ----------------------------------------------------------------------------------------------------------------------
# ResNet-50 body: stem + 16 bottleneck blocks, frozen BatchNorm folded into the convolution epilogues
# ----------------------------------------------------------------------------------------------------------------------
def _fold_bn(owner, name, bn):
    """Frozen / eval-mode BatchNorm2d -> (scale, shift) f32 [C]: y = x * scale + shift.  Cached on ``owner``."""
    cache = owner.__dict__.setdefault("_sfvos_folded", {})
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    tag = tuple((t.data_ptr(), t._version) for t in ts)
    hit = cache.get(name)
    if hit is None or hit[0] != tag:
        eps = float(getattr(bn, "eps", 1e-5))
        scale = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + eps))
        shift = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
        hit = (tag, scale.float().contiguous(), shift.float().contiguous())
        cache[name] = hit
    return hit[1], hit[2]


def _derived(owner, name, w, fn):
    """A re-arranged copy of a frozen weight (stem patch order, stride-2 phase sub-kernels), cached on ``owner``."""
    cache = owner.__dict__.setdefault("_sfvos_derived", {})
    tag = (w.data_ptr(), w._version)
    hit = cache.get(name)
    if hit is None or hit[0] != tag:
        hit = (tag, fn(w.detach().float()).contiguous())
        cache[name] = hit
    return hit[1]


def _conv_bn(owner, name, x, w, scale, shift, umma, out_dtype, relu, stride=1):
    """Conv2d (1x1 or 3x3, 'same' padding, stride 1 or 2, no bias) + folded BatchNorm (+ ReLU) on a channels-last Act."""
    cout, cin, kh, kw = w.shape
    dev = x.buf.device
    if stride == 1:
        wp, cp = _packed(owner, name, w, umma)
        y = Act.empty(x.B, 1, x.H, x.W, cout, out_dtype, dev)
        ops.conv(x, wp, cp, cout, (1, kh, kw), (0, kh // 2, kw // 2), 1, y, umma=umma, relu=relu, scale=scale, shift=shift)
        return y
    assert stride == 2 and x.H % 2 == 0 and x.W % 2 == 0 and x.cstride == x.C and x.bstride == 0
    ho, wo = x.H // 2, x.W // 2
    frame = x.H * x.W * x.C

    def phase(pi, pj):      # x[:, pi::2, pj::2, :] as a strided Act: pixel stride 2C, row stride 2WC
        v = Act(x.buf, x.B, 1, ho, wo, x.C, 2 * x.C, x.ch_off + (pi * x.W + pj) * x.C)
        return v, (2 * x.W * x.C, frame, frame)
    if kh == 1:             # 1x1 stride 2 (the downsample branch): every second pixel of every second row
        v, strides = phase(0, 0)
        wp, cp = _packed(owner, name, w, umma)
        y = Act.empty(x.B, 1, ho, wo, cout, out_dtype, dev)
        ops.conv(v, wp, cp, cout, (1, 1, 1), (0, 0, 0), 1, y, umma=umma, relu=relu, scale=scale, shift=shift, x_strides=strides)
        return y
    # 3x3 stride 2, padding 1: out[y,x] = sum_ij in[2y+i-1, 2x+j-1] w[i,j].  Rows 2y-1 / 2y+1 are rows y-1 / y of the odd
    # row phase (taps i = 0, 2), row 2y is row y of the even phase (tap i = 1); the same for columns: four stride-1
    # convolutions (1x1, 1x2, 2x1, 2x2 taps) over the four phase images, accumulated in f32, then scale / shift / ReLU
    acc = Act.empty(x.B, 1, ho, wo, cout, torch.float32, dev)
    first = True
    for pi, rows in ((0, [1]), (1, [0, 2])):
        for pj, cols in ((0, [1]), (1, [0, 2])):
            sub = _derived(owner, f"{name}.p{pi}{pj}", w, lambda t, r=rows, c=cols: t[:, :, r][:, :, :, c])
            wp, cp = _packed(owner, f"{name}.p{pi}{pj}", sub, umma)
            v, strides = phase(pi, pj)
            ops.conv(v, wp, cp, cout, (1, len(rows), len(cols)), (0, len(rows) - 1, len(cols) - 1), 1, acc, umma=umma,
                     accumulate=not first, x_strides=strides)
            first = False
    y = Act.empty(x.B, 1, ho, wo, cout, out_dtype, dev)
    ops.affine_act(acc, y, scale, shift, relu)
    return y


class ResNetBody(tv_utils.IntermediateLayerGetter):
    """torchvision's ``backbone.body`` (IntermediateLayerGetter over ResNet-50: conv1 / bn1 / relu / maxpool / layer1..4) with
    the forward on libsfvos: same modules, same parameters, same state_dict keys.  The 7x7 stride-2 stem is a patch GEMM
    (sfvos_im2col + a 1x1 convolution), every other convolution runs on sfvos_conv_umma with the frozen BatchNorm folded into
    its scale / shift epilogue, stride-2 3x3 convolutions as four phase convolutions; the residual stream stays f32.
    Returns ``{name: [N,C,H,W]}`` channels-last in the activation dtype -- the FPN below consumes them in place.
    Anything it cannot express (trainable or train-mode BatchNorm, autograd, basic blocks, odd sizes) goes to torchvision's
    forward."""

    precision = None

    def _native_ok(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[-2] % 32 or x.shape[-1] % 32 or not x.is_cuda:
            return False
        mods = dict(self.items())
        if not all(k in mods for k in ("conv1", "bn1", "maxpool", "layer1", "layer2", "layer3", "layer4")):
            return False
        for m in self.modules():
            if isinstance(m, torch.nn.BatchNorm2d) and m.training:
                return False
        c1 = mods["conv1"]
        if c1.kernel_size != (7, 7) or c1.stride != (2, 2) or c1.padding != (3, 3) or c1.bias is not None:
            return False
        return all(type(b).__name__ == "Bottleneck" and b.conv2.groups == 1 and b.conv2.dilation == (1, 1)
                   for k in ("layer1", "layer2", "layer3", "layer4") for b in mods[k])

    @ops.device_guard
    def forward(self, x):
        if not self._native_ok(x):
            return super().forward(x)
        ops.device_check()
        precision = self.precision or _default_precision()
        umma, dt_act = precision != "fp32", _act_dtype(precision)
        mods = dict(self.items())
        n, _, h, w = x.shape
        dev = x.device
        # ---- stem: 7x7 stride-2 conv on 3 channels as a [pixels, 192] x [192, 64] GEMM, then 3x3 stride-2 max-pool
        conv1 = mods["conv1"]
        kp = 192
        ho, wo = h // 2, w // 2
        rows = Act.empty(n, 1, ho, wo, kp, dt_act, dev)
        call("sfvos_im2col", _p(x.float().contiguous()), rows.ptr(), ops.dt(rows.buf), n, 3, h, w, 7, 7, 2, 3, kp, stream())
        w_stem = _derived(self, "stem", conv1.weight, lambda t: torch.nn.functional.pad(
            t.permute(0, 2, 3, 1).reshape(t.shape[0], -1), (0, kp - 147)).reshape(t.shape[0], kp, 1, 1))
        sc, sh = _fold_bn(self, "bn1", mods["bn1"])
        stem = _conv_bn(self, "stem", rows, w_stem, sc, sh, umma, dt_act, relu=True)
        cur = Act.empty(n, 1, ho // 2, wo // 2, 64, dt_act, dev)
        call("sfvos_maxpool3x3s2", stem.ptr(), cur.ptr(), ops.dt(cur.buf), n, ho, wo, 64, stream())
        cur32 = None                                           # f32 residual stream (None before the first block: it downsamples)
        out = OrderedDict()
        for lname in ("layer1", "layer2", "layer3", "layer4"):
            for bi, blk in enumerate(mods[lname]):
                tag = f"{lname}.{bi}"
                s = blk.conv2.stride[0]
                o = _conv_bn(self, tag + ".conv1", cur, blk.conv1.weight, *_fold_bn(self, tag + ".bn1", blk.bn1), umma, dt_act, True,
                             stride=blk.conv1.stride[0])
                o = _conv_bn(self, tag + ".conv2", o, blk.conv2.weight, *_fold_bn(self, tag + ".bn2", blk.bn2), umma, dt_act, True, stride=s)
                o = _conv_bn(self, tag + ".conv3", o, blk.conv3.weight, *_fold_bn(self, tag + ".bn3", blk.bn3), umma, torch.float32, False)
                if blk.downsample is not None:
                    ds_conv, ds_bn = blk.downsample[0], blk.downsample[1]
                    idt = _conv_bn(self, tag + ".down", cur, ds_conv.weight, *_fold_bn(self, tag + ".dbn", ds_bn), umma, torch.float32,
                                   False, stride=ds_conv.stride[0])
                else:
                    idt = cur32
                nxt32 = Act.empty(o.B, 1, o.H, o.W, o.C, torch.float32, dev)
                nxt = nxt32 if not umma else Act.empty(o.B, 1, o.H, o.W, o.C, torch.bfloat16, dev)
                call("sfvos_add_relu", o.ptr(), idt.ptr(), nxt32.ptr(), nxt.ptr() if umma else None, o.npix * o.C, stream())
                cur, cur32 = nxt, nxt32
            if lname in self.return_layers:
                out[self.return_layers[lname]] = _nchw_view(cur.buf, cur.B, cur.H, cur.W, cur.C)
        return out


class FeaturePyramidNetwork(tv_fpn.FeaturePyramidNetwork):
    """Same ctor / parameters / state_dict keys (``inner_blocks.{i}.0.*``, ``layer_blocks.{i}.0.*``) as torchvision's; forward
    on libsfvos.  Returns ``{name: [N,256,H,W]}`` with channels_last strides, bf16 on the product path (f32 with
    ``precision == "fp32"``): exactly what SlowFastLayers lays out as its input, so no conversion pass follows."""

    precision = None

    @ops.device_guard
    def forward(self, x: Dict[str, Tensor]) -> Dict[str, Tensor]:
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(v.requires_grad for v in x.values())):
            # the reference freezes backbone and RPN (code/helpers/model.py:176-179) and calls them under no_grad (:324-333);
            # someone fine-tuning the FPN gets torchvision's differentiable forward -- this module has no backward
            return super().forward(x)
        ops.device_check()
        precision = self.precision or _default_precision()
        umma, dt_act = precision != "fp32", _act_dtype(precision)
        names, feats = list(x.keys()), list(x.values())
        n = len(feats)
        for blk in list(self.inner_blocks) + list(self.layer_blocks):
            assert len(blk) == 1, "norm / activation layers inside the FPN blocks are not supported"
        # lateral 1x1 convolutions, f32 results (the top-down chain adds up to n of them)
        inner = [_conv(self, f"inner{i}", _to_cl_act(f, dt_act), self.inner_blocks[i][0].weight, self.inner_blocks[i][0].bias, umma,
                       torch.float32) for i, f in enumerate(feats)]
        results = [None] * n
        for i in range(n - 1, -1, -1):
            a = inner[i]
            top = inner[i + 1] if i + 1 < n else None
            merged = a if not umma else Act.empty(a.B, 1, a.H, a.W, a.C, torch.bfloat16, a.buf.device)
            if a.npix and (top is not None or umma):
                call("sfvos_upsample_add", top.ptr() if top is not None else None, top.H if top is not None else 0,
                     top.W if top is not None else 0, a.ptr(), merged.ptr() if umma else None, a.B, a.H, a.W, a.C, stream())
            out = _conv(self, f"layer{i}", merged, self.layer_blocks[i][0].weight, self.layer_blocks[i][0].bias, umma, dt_act)
            results[i] = _nchw_view(out.buf, out.B, out.H, out.W, out.C)
        if self.extra_blocks is not None:
            if isinstance(self.extra_blocks, tv_fpn.LastLevelMaxPool):
                # max_pool2d(kernel 1, stride 2) = every second row / column of the coarsest map
                results.append(results[-1][:, :, ::2, ::2].contiguous(memory_format=torch.channels_last))
                names = names + ["pool"]
            else:
                results, names = self.extra_blocks(results, feats, names)
        return OrderedDict(zip(names, results))


class RPNHead(tv_rpn.RPNHead):
    """Same ctor / parameters / state_dict keys (``conv.{i}.0.*``, ``cls_logits.*``, ``bbox_pred.*``) as torchvision's.
    Returns (objectness [N,A,H,W], box deltas [N,4A,H,W]) per level, f32."""

    precision = None

    @ops.device_guard
    def forward(self, x: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(v.requires_grad for v in x)):
            return super().forward([v.float() for v in x])         # trainable RPN head: torchvision's differentiable forward
        ops.device_check()
        precision = self.precision or _default_precision()
        umma, dt_act = precision != "fp32", _act_dtype(precision)
        a_cls, a_box = self.cls_logits.out_channels, self.bbox_pred.out_channels
        n_pad = (a_cls + a_box + 31) // 32 * 32
        dev = self.cls_logits.weight.device
        # cls_logits and bbox_pred as one GEMM over the stacked (zero-padded) weight
        capturing = dev.type == "cuda" and torch.cuda.is_current_stream_capturing()
        prm = (self.cls_logits.weight, self.bbox_pred.weight, self.cls_logits.bias, self.bbox_pred.bias)
        key = tuple((t.data_ptr(), t._version) for t in prm) + (str(dev),)
        hit = self.__dict__.get("_sfvos_pred")
        if hit is None or hit[0] != key or capturing:
            w = torch.zeros(n_pad, self.cls_logits.in_channels, 1, 1, dtype=torch.float32, device=dev)
            b = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            w[:a_cls] = self.cls_logits.weight.detach(); w[a_cls:a_cls + a_box] = self.bbox_pred.weight.detach()
            b[:a_cls] = self.cls_logits.bias.detach(); b[a_cls:a_cls + a_box] = self.bbox_pred.bias.detach()
            hit = (key, w, b)
            if not capturing:
                self.__dict__["_sfvos_pred"] = hit
        _, w_pred, b_pred = hit
        logits, bbox_reg = [], []
        for feature in x:
            t = _to_cl_act(feature, dt_act)
            for j, blk in enumerate(self.conv):
                t = _conv(self, f"conv{j}", t, blk[0].weight, blk[0].bias, umma, dt_act, relu=True)
            pred = _conv(self, "pred", t, w_pred, b_pred, umma, torch.float32)
            out = torch.empty(t.B, n_pad, t.H, t.W, dtype=torch.float32, device=dev)
            if t.npix:
                ops.nhwc_to_nchw(pred, out)
            logits.append(out[:, :a_cls].contiguous())
            bbox_reg.append(out[:, a_cls:a_cls + a_box].contiguous())
        return logits, bbox_reg


class AnchorGenerator(tv_anchor_utils.AnchorGenerator):
    """torchvision's generator takes the anchors' dtype from the feature maps; with bf16 features that would quantise box
    coordinates to 8 px.  Only the maps' spatial sizes are needed, so hand it f32 stand-ins of the same shape."""

    def forward(self, image_list, feature_maps: List[Tensor]) -> List[Tensor]:
        proxies = [torch.empty((1, 1) + tuple(f.shape[-2:]), dtype=torch.float32, device=f.device) for f in feature_maps]
        return super().forward(image_list, proxies)


def install_backbone(maskrcnn_model, precision: Optional[str] = None):
    """Swap the libsfvos FPN / RPN head into a torchvision Mask R-CNN IN PLACE (same parameters, same state_dict keys).
    Returns the model."""
    precision = precision or _default_precision()
    body = getattr(maskrcnn_model.backbone, "body", None)
    if body is not None and type(body) in (tv_utils.IntermediateLayerGetter, ResNetBody):
        body.__class__ = ResNetBody
        body.precision = precision
    fpn = getattr(maskrcnn_model.backbone, "fpn", None)
    if fpn is not None and type(fpn) in (tv_fpn.FeaturePyramidNetwork, FeaturePyramidNetwork) and \
            all(len(b) == 1 for b in list(fpn.inner_blocks) + list(fpn.layer_blocks)):
        fpn.__class__ = FeaturePyramidNetwork
        fpn.precision = precision
    head = maskrcnn_model.rpn.head
    if type(head) in (tv_rpn.RPNHead, RPNHead) and all(len(b) == 2 for b in head.conv):
        head.__class__ = RPNHead
        head.precision = precision
    gen = maskrcnn_model.rpn.anchor_generator
    if type(gen) in (tv_anchor_utils.AnchorGenerator, AnchorGenerator):
        gen.__class__ = AnchorGenerator
    return maskrcnn_model
