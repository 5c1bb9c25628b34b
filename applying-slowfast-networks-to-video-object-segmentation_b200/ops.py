"""Torch-facing wrappers over the C ABI (include/sfvos.h).  torch is used only for device memory and streams;
every arithmetic kernel is in libsfvos.so.  All wrappers require CUDA tensors and raise otherwise (no fallback)."""
import ctypes
import functools

import torch

from . import _lib
from ._lib import BF16, F32, F64, ConvParams, RoiParams, WgradParams, call


TIMING = None   # bench.py sets this to a list: every conv / wgrad launch then appends (kernel, flops, start, end)

# Optional gradient arena (dp.GradArena): when set, the backward passes write each parameter gradient straight into its slice
# of ONE flat f32 buffer (and return that slice, which autograd adopts as p.grad without a copy), so the data-parallel step
# all-reduces contiguous ranges of the arena -- no torch.cat / pack pass over the 73 MB of gradients.
GRAD_ARENA = None


def new_grad(param, shape=None):
    """Accumulator for ``param``'s gradient: its arena slice (NOT re-zeroed: kernels add to it, the arena is cleared once per
    optimizer step) or a fresh zero-filled f32 tensor.  ``shape``: view the slice differently (same number of elements)."""
    if GRAD_ARENA is not None:
        v = GRAD_ARENA.view(param, shape)
        if v is not None:
            return v
    return torch.zeros(tuple(shape) if shape is not None else param.shape, dtype=torch.float32, device=param.device)



def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _find_cuda(obj, depth=0):
    if torch.is_tensor(obj):
        return obj if obj.is_cuda else None
    if depth < 3:
        if isinstance(obj, dict):
            obj = tuple(obj.values())
        if isinstance(obj, (list, tuple)):
            for o in obj:
                t = _find_cuda(o, depth + 1)
                if t is not None:
                    return t
    return None


def device_guard(fn):
    """Run ``fn`` with the CUDA device of its first CUDA tensor argument current.  libsfvos launches on the CURRENT device
    (``stream()`` above, sfvos_device_check, the TMA descriptors and the SM count all follow it), while the reference API
    lets a module live on any device (``SlowFastLayers(256, device='cuda:1', ...)``): every entry that launches kernels -
    the autograd Function forward/backward methods and the free functions of roi_heads - is wrapped with this."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _find_cuda(args)
        if t is None and kwargs:
            t = _find_cuda(tuple(kwargs.values()))
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def _timed_call(kernel, flops, name, params):
    if TIMING is None:
        call(name, ctypes.byref(params), stream())
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, ctypes.byref(params), stream())
    e1.record()
    TIMING.append((_lib.last_kernel() or kernel, flops, e0, e1))


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def torch_dtype(code):
    return torch.float32 if code == F32 else torch.bfloat16


def _p(t, byte_off=0):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libsfvos ops need CUDA tensors (there is no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr() + byte_off)


class Act:
    """A channels-last activation [B,T,H,W,C] living in (a channel slice of) a dense [B*T*H*W, cstride] buffer."""

    __slots__ = ("buf", "B", "T", "H", "W", "C", "cstride", "ch_off", "bstride")

    def __init__(self, buf, B, T, H, W, C, cstride=None, ch_off=0, bstride=0):
        self.buf, self.B, self.T, self.H, self.W, self.C = buf, B, T, H, W, C
        self.cstride = cstride if cstride is not None else C
        self.ch_off = ch_off          # element offset of the first channel / first frame inside buf
        self.bstride = bstride        # elements between clips; 0 = dense (T*H*W*cstride)

    @staticmethod
    def empty(B, T, H, W, C, dtype, device, cstride=None):
        cs = cstride if cstride is not None else C
        return Act(torch.empty(B * T * H * W * cs, dtype=dtype, device=device), B, T, H, W, C, cs, 0)

    @property
    def npix(self):
        return self.B * self.T * self.H * self.W

    @property
    def dtype(self):
        return self.buf.dtype

    def ptr(self):
        return _p(self.buf, self.ch_off * self.buf.element_size())

    def slice(self, ch_off, C):
        return Act(self.buf, self.B, self.T, self.H, self.W, C, self.cstride, self.ch_off + ch_off, self.bstride)

    def frames(self, t0, t1):
        """Temporal sub-range [t0,t1) of every clip (same storage; clips keep the parent's batch stride)."""
        per = self.H * self.W * self.cstride
        bs = self.bstride if self.bstride else self.T * per
        return Act(self.buf, self.B, t1 - t0, self.H, self.W, self.C, self.cstride, self.ch_off + t0 * per, bs)

    def as_nchw(self):
        """[B*T, C, H, W]-shaped strided view (channels_last memory) of a dense activation -- no copy."""
        assert self.bstride == 0 and self.ch_off < self.cstride
        v = self.buf.view(self.B * self.T, self.H, self.W, self.cstride)[..., self.ch_off:self.ch_off + self.C]
        return v.permute(0, 3, 1, 2)


def device_check():
    call("sfvos_device_check")


def pack_weights(w, mode, out_dtype, Cp, tap=(0, 0)):
    """mode 0/1: w [Cout,Cin,kt,kh,kw]; mode 2/3: ConvTranspose2d w [Cin,Cout,kh,kw], single tap."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    if mode <= 1:
        if w.dim() == 4:
            w = w.unsqueeze(2)
        Cout, Cin, kt, kh, kw = w.shape
        taps = kt * kh * kw
    else:
        Cin, Cout, kh, kw = w.shape
        kt, taps = 1, 1
    N = Cout if mode in (0, 2) else Cin
    if out_dtype == BF16:
        out = torch.empty(N, taps * Cp, dtype=torch.bfloat16, device=w.device)
    else:
        out = torch.empty(taps * Cp, N, dtype=torch.float32, device=w.device)
    call("sfvos_pack_weights", _p(w), _p(out), out_dtype, mode, Cout, Cin, kt, kh, kw, Cp, tap[0], tap[1], stream())
    return out


def unpack_wgrad(dw, grad, mode, tap=(0, 0)):
    if mode == 0:
        g5 = grad if grad.dim() == 5 else grad.unsqueeze(2)
        Cout, Cin, kt, kh, kw = g5.shape
    else:
        Cin, Cout, kh, kw = grad.shape
        kt = 1
    assert grad.is_contiguous() and grad.dtype == torch.float32
    call("sfvos_unpack_wgrad", _p(dw), _p(grad), mode, Cout, Cin, kt, kh, kw, tap[0], tap[1], stream())


def conv(x, w_packed, Cp, N, k, pad, To, y, *, umma, relu=False, scale=None, shift=None, stats=None, accumulate=False,
         x_strides=None, scatter=None, relu_mask=None, dbias=None, addend=None):
    """x: Act (input), y: Act (output, B*To*OH*OW pixels).  k=(kt,kh,kw), pad=(pt,ph,pw).
    stats: optional f32 [2,N] tensor receiving (sum, sumsq) of the raw accumulators (umma only)."""
    p = ConvParams()
    p.x = x.ptr(); p.B, p.T, p.H, p.W, p.C, p.x_cstride = x.B, x.T, x.H, x.W, x.C, x.cstride
    p.x_bstride = x.bstride
    if x_strides is not None:
        p.x_hstride, p.x_tstride, p.x_bstride = x_strides
    p.w = _p(w_packed); p.Cp = Cp; p.N = N
    p.kt, p.kh, p.kw = k
    p.pad_t, p.pad_h, p.pad_w = pad
    p.To = To
    p.y = y.ptr(); p.y_dtype = dt(y.buf); p.relu = int(relu); p.y_cstride = y.cstride
    p.scale = _p(scale); p.shift = _p(shift)
    if stats is not None:
        p.sum = _p(stats); p.sumsq = _p(stats, N * 4)
    p.accumulate = int(accumulate)
    if scatter is not None:
        p.OH, p.OW, p.oy_mul, p.oy_off, p.ox_mul, p.ox_off = scatter
    if relu_mask is not None:       # fused ReLU backward: relu_mask = Act of the post-ReLU activation, dbias f32 [N] accumulates
        assert umma and stats is None and relu_mask.dtype == torch.bfloat16
        p.relu_mask = relu_mask.ptr(); p.relu_mask_cstride = relu_mask.cstride
        if dbias is not None:
            p.sum = _p(dbias)
    if addend is not None:          # y = act(...) + addend: an Act indexed like y (f32 or bf16); tensor-core path only
        assert umma and not accumulate and addend.npix == y.npix
        p.addend = addend.ptr(); p.addend_dtype = dt(addend.buf); p.addend_cstride = addend.cstride
    # algorithmic FLOPs: no zero-padded taps.  fprop: To = T - kt + 1 output frames; a data-gradient launch
    # (full temporal padding, To = the forward INPUT frames) does the forward layer's work, one product per forward output
    # frame and tap, so it counts x.T (= the forward output frames) -- min() covers both
    flops = 2.0 * x.B * min(To, x.T) * x.H * x.W * N * x.C * k[0] * k[1] * k[2]
    _timed_call("conv_umma" if umma else "conv_simt", flops, "sfvos_conv_umma" if umma else "sfvos_conv_simt", p)


def wgrad(x, dy, k, pad, dw, *, umma, dy_strides=None, flops=None):
    """dw f32 [taps, x.C, dy.C] += x^T dy over all pixels (x shifted per tap).  ``flops``: algorithmic FLOPs of the launch when
    they are not 2 * pixels(dy) * x.C * dy.C * taps (the operand-swapped lateral weight gradient)."""
    p = WgradParams()
    p.x = x.ptr(); p.B, p.T, p.H, p.W, p.C, p.x_cstride = x.B, x.T, x.H, x.W, x.C, x.cstride
    p.x_bstride = x.bstride
    p.dy = dy.ptr(); p.To = dy.T; p.N = dy.C; p.dy_cstride = dy.cstride
    if dy_strides is not None:
        p.dy_hstride, p.dy_tstride, p.dy_bstride = dy_strides
    p.kt, p.kh, p.kw = k
    p.pad_t, p.pad_h, p.pad_w = pad
    p.dw = _p(dw)
    if not umma:
        # validation mode: fp64 split-K partials reduced in a fixed order need scratch (the caller owns all memory)
        nbytes = _lib.load().sfvos_wgrad_simt_workspace_bytes(ctypes.byref(p))
        if nbytes:
            ws = torch.empty(nbytes // 8, dtype=torch.float64, device=dw.device)
            p.workspace, p.workspace_bytes = _p(ws), nbytes
    if flops is None:
        flops = 2.0 * x.B * dy.T * x.H * x.W * dy.C * x.C * k[0] * k[1] * k[2]
    _timed_call("wgrad_umma" if umma else "wgrad_simt", flops, "sfvos_wgrad_umma" if umma else "sfvos_wgrad_simt", p)


def _reduce_workspace(npix, C, device):
    nbytes = _lib.load().sfvos_reduce_workspace_bytes(npix, C)
    return torch.empty(max(1, nbytes // 8), dtype=torch.float64, device=device), nbytes


def channel_stats(x, stats):
    """stats: f64 [2C] <- per-channel (sums, sums of squares) of the f32 activation x; fixed reduction order."""
    assert x.dtype == torch.float32 and stats.dtype == torch.float64 and stats.numel() >= 2 * x.C
    if x.npix == 0:
        stats.zero_()
        return
    ws, nbytes = _reduce_workspace(x.npix, x.C, x.buf.device)
    call("sfvos_channel_stats", x.ptr(), x.npix, x.C, x.cstride, _p(stats), _p(ws), nbytes, stream())


def _stats_dtype(stats):
    return (F64, 8) if stats.dtype == torch.float64 else (F32, 4)


def bn_finalize(stats, count, conv_bias, gamma, beta, running_mean, running_var, nbt, momentum, eps, out4):
    """stats: [2C] (sum, sumsq), f32 (umma epilogues) or f64 (channel_stats); out4: f32 [4,C] = (scale, shift, mean, rstd)."""
    C = gamma.numel()
    code, esz = _stats_dtype(stats)
    call("sfvos_bn_finalize", _p(stats), _p(stats, C * esz), code, float(count), _p(conv_bias), _p(gamma), _p(beta),
         _p(running_mean), _p(running_var), _p(nbt), float(momentum), float(eps),
         _p(out4), _p(out4, C * 4), _p(out4, 2 * C * 4), _p(out4, 3 * C * 4), C, stream())


def bn_running_update(calls, conv_bias, running_mean, running_var, nbt, momentum):
    """calls: list of (stats [2C] = (sum, sumsq), f32 or f64, count) in forward-call order (see sfvos_bn_running_update)."""
    C = running_mean.numel()
    for i in range(0, len(calls), _lib.BN_MAX_CALLS):
        part = calls[i:i + _lib.BN_MAX_CALLS]
        p = _lib.BnRunningParams()
        code, esz = _stats_dtype(part[0][0])
        for j, (stats, count) in enumerate(part):
            assert _stats_dtype(stats)[0] == code
            p.sum[j] = stats.data_ptr()
            p.sumsq[j] = stats.data_ptr() + C * esz
            p.count[j] = float(count)
        p.n_calls = len(part); p.stats_dtype = code
        p.conv_bias = _p(conv_bias); p.running_mean = _p(running_mean); p.running_var = _p(running_var)
        p.num_batches_tracked = _p(nbt); p.momentum = float(momentum); p.C = C
        call("sfvos_bn_running_update", ctypes.byref(p), stream())


def bn_fold_eval(conv_bias, gamma, beta, running_mean, running_var, eps, out):
    """out: f32 [2C] = (scale, shift), or [4C] = (scale, shift, mean, rstd) like bn_finalize's (for the eval-mode backward)."""
    C = gamma.numel()
    full = out.numel() >= 4 * C
    call("sfvos_bn_fold_eval", _p(conv_bias), _p(gamma), _p(beta), _p(running_mean), _p(running_var), float(eps),
         _p(out), _p(out, C * 4), _p(out, 2 * C * 4) if full else None, _p(out, 3 * C * 4) if full else None, C, stream())


def affine_act(x, y, scale, shift, relu):
    call("sfvos_affine_act", x.ptr(), dt(x.buf), x.cstride, y.ptr(), dt(y.buf), y.cstride, _p(scale), _p(shift),
         int(relu), x.npix, x.C, stream())


def bn_bwd(dy, raw, bn4, gamma, relu, dx, dgamma, dbeta, sums=None, deterministic=False, fixed_stats=False, dbias=None):
    """dy, raw (f32), dx: Acts over the same pixels; bn4 = (scale, shift, mean, rstd) [4,C]; ``sums``: optional
    zero-filled f32 [2C] scratch.  ``deterministic`` (validation mode, f32 dy): the two per-channel sums are formed in
    fp64 in a fixed order instead of with float atomics.  ``fixed_stats``: eval-mode BatchNorm (bn4 from bn_fold_eval): the
    statistics are constants, dx = gamma*rstd*dy_m, and ``dbias`` receives the conv-bias gradient."""
    C = raw.C
    if sums is None:
        sums = torch.zeros(2 * C, dtype=torch.float32, device=raw.buf.device)
    sc, sh, mu, rs = _p(bn4), _p(bn4, C * 4), _p(bn4, 2 * C * 4), _p(bn4, 3 * C * 4)
    ws, nbytes = _reduce_workspace(raw.npix, C, raw.buf.device) if deterministic and raw.npix else (None, 0)
    call("sfvos_bn_bwd_reduce", dy.ptr(), dt(dy.buf), dy.cstride, raw.ptr(), raw.cstride, sc, sh, mu, rs, int(relu),
         raw.npix, C, _p(sums), _p(ws), nbytes, stream())
    call("sfvos_bn_bwd_apply", dy.ptr(), dt(dy.buf), dy.cstride, raw.ptr(), raw.cstride, sc, sh, mu, rs, _p(gamma),
         int(relu), raw.npix, C, _p(sums), dx.ptr(), dt(dx.buf), dx.cstride, _p(dgamma), _p(dbeta), int(fixed_stats),
         _p(dbias), stream())


def relu_bwd(dy, y, dx, dbias):
    call("sfvos_relu_bwd", dy.ptr(), dt(dy.buf), dy.cstride, y.ptr(), dt(y.buf), y.cstride, dx.ptr(), dt(dx.buf),
         dx.cstride, _p(dbias), y.npix, y.C, stream())


def nchw_to_nhwc(src, dst_act, frame_off=0):
    """src (f32 | bf16) [F,C,H,W] contiguous -> frames [frame_off, frame_off+F) of dst_act."""
    assert src.dtype in (torch.float32, torch.bfloat16) and src.is_contiguous()
    F, C, H, W = src.shape
    esz = dst_act.buf.element_size()
    off = (frame_off * H * W * dst_act.cstride + dst_act.ch_off) * esz
    call("sfvos_nchw_to_nhwc", _p(src), dt(src), C * H * W, _p(dst_act.buf, off), dt(dst_act.buf), dst_act.cstride, F, C, H * W,
         stream())


def nhwc_to_nchw(src_act, dst):
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    call("sfvos_nhwc_to_nchw", src_act.ptr(), dt(src_act.buf), src_act.cstride, _p(dst), src_act.B * src_act.T,
         src_act.C, src_act.H * src_act.W, stream())


def roi_levels(rois, k_min, k_max):
    K = rois.shape[0]
    levels = torch.empty(K, dtype=torch.int32, device=rois.device)
    call("sfvos_roi_levels", _p(rois), K, k_min, k_max, _p(levels), stream())
    return levels


def _roi_params(feats, dfeats, shapes, scales, N, C, cstride, feat_dtype, rois, levels, P, sr, out, out_nchw):
    p = RoiParams()
    for i in range(len(shapes)):
        if feats is not None:
            p.feat[i] = feats[i].data_ptr()
        if dfeats is not None:
            p.dfeat[i] = dfeats[i].data_ptr()
        p.H[i], p.W[i] = shapes[i]
        p.scale[i] = scales[i]
    p.n_levels = len(shapes); p.feat_dtype = feat_dtype
    p.N, p.C, p.cstride = N, C, cstride
    p.rois = _p(rois); p.levels = _p(levels); p.K = rois.shape[0]; p.P = P; p.sampling_ratio = sr
    p.out = _p(out); p.out_dtype = dt(out); p.out_nchw = int(out_nchw)
    return p


def _timed_plain(kernel, name, params):
    """bench.py timing hook for the non-GEMM kernels (work = 0: bench.py knows their algorithmic bytes)."""
    if TIMING is None:
        call(name, ctypes.byref(params), stream())
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, ctypes.byref(params), stream())
    e1.record()
    TIMING.append((kernel, 0.0, e0, e1))


def roi_align_fwd(feats, shapes, scales, N, C, rois, levels, P, sr, out, out_nchw):
    p = _roi_params(feats, None, shapes, scales, N, C, C, dt(feats[0]), rois, levels, P, sr, out, out_nchw)
    _timed_plain(f"roi_align_fwd_p{P}", "sfvos_roi_align_fwd", p)


def roi_align_bwd(dfeats, shapes, scales, N, C, rois, levels, P, sr, gout, out_nchw):
    p = _roi_params(None, dfeats, shapes, scales, N, C, C, F32, rois, levels, P, sr, gout, out_nchw)
    _timed_plain(f"roi_align_bwd_p{P}", "sfvos_roi_align_bwd", p)


def mask_targets(masks_u8, rois, M):
    n_obj, H, W = masks_u8.shape
    K = rois.shape[0]
    out = torch.empty(K, M, M, dtype=torch.float32, device=rois.device)
    call("sfvos_mask_targets", _p(masks_u8), n_obj, H, W, _p(rois), K, M, _p(out), stream())
    return out


def axpby(x, y, a, b):
    call("sfvos_axpby", _p(x), _p(y), float(a), float(b), x.numel(), stream())


def launches():
    return _lib.LAUNCHES
