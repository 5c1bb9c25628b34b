"""ORACLE (test infrastructure only) -- CPU fp32 restatement of the reference SlowFast temporal module.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg
may import this file.  The product path (``sfvos_b200``) never does.

Parity pin: ``tests/golden/slowfast_*.npz`` were produced by importing the *unmodified* reference
(``/root/reference/code/helpers/model.py``) in the build container with
``tests/golden/make_golden.py``; ``tests/test_oracle.py`` checks this restatement against them
bit-for-bit on CPU (same torch build), so the oracle is pinned to outputs of the reference itself.

What is restated (reference file:line):
  * temporal kernel-size rules            code/helpers/model.py:96-109
  * layer construction order / shapes     code/helpers/model.py:37-69, 71-94
  * forward (two pathways + 2 laterals)   code/helpers/model.py:111-149
  * per-level driver / concat             code/helpers/model.py:151-165
The arithmetic itself (conv3d, batch-norm) is delegated to torch's CPU fp32 functional ops, which is what
the reference's nn.Conv3d / nn.BatchNorm3d dispatch to.  ``oracle/conv3d_ref.c`` is an independent plain-C
direct convolution + batch-norm used to cross-check those functional calls on small cases.
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# (name, kind) in the reference's registration order, code/helpers/model.py:47-67
LAYER_ORDER = [
    ("fast_conv1", "bn_f1"), ("slow_conv1", "bn_s1"),
    ("fast_conv2", "bn_f2"), ("slow_conv2", "bn_s2"),
    ("fast_conv3", "bn_f3"), ("slow_conv3", "bn_s3"),
    ("conv_f2s1", "bn_f2s1"), ("conv_f2s2", "bn_f2s2"),
]


def calc_kernel_sizes(pathway_size):
    """code/helpers/model.py:96-103"""
    div = pathway_size // 3
    rem = pathway_size % 3
    if rem == 0:
        return (div, div + 1, div + 1)
    if rem == 1:
        return (div + 1, div + 1, div + 1)
    return (div + 1, div + 1, div + 2)


def calc_fuse_kernel_size(slow_in, slow_kernel, fast_in, fast_kernel):
    """code/helpers/model.py:105-109"""
    out_slow = slow_in - slow_kernel + 1
    out_fast = fast_in - fast_kernel + 1
    return out_fast - out_slow + 1, out_slow, out_fast


def layer_specs(sp, fp, input_size=256):
    """Shapes of every conv in registration order: name -> (cout, cin, k_t, k_hw, has_bias)."""
    ks = calc_kernel_sizes(sp)
    kf = calc_kernel_sizes(fp)
    kl1, so1, fo1 = calc_fuse_kernel_size(sp, ks[0], fp, kf[0])
    kl2, _, _ = calc_fuse_kernel_size(so1, ks[1], fo1, kf[1])
    return OrderedDict([
        ("fast_conv1", (32, input_size, kf[0], 3, True)),
        ("slow_conv1", (192, input_size, ks[0], 3, True)),
        ("fast_conv2", (32, 32, kf[1], 3, True)),
        ("slow_conv2", (192, 256, ks[1], 3, True)),
        ("fast_conv3", (32, 32, kf[2], 3, True)),
        ("slow_conv3", (224, 256, ks[2], 3, True)),
        ("conv_f2s1", (64, 32, kl1, 1, False)),
        ("conv_f2s2", (64, 32, kl2, 1, False)),
    ])


def init_state_dict(sp, fp, seed=63, input_size=256):
    """Default-PyTorch init in the reference's construction order (model.py:47-67) under ``seed``
    (code/helpers/constants.py:11 uses 63).  Conv: kaiming_uniform(a=sqrt(5)) on weight, U(+-1/sqrt(fan_in))
    on bias; BN: gamma=1, beta=0, running_mean=0, running_var=1, num_batches_tracked=0.
    Draw order (weight then bias per conv, in construction order) reproduces nn.Conv3d.reset_parameters."""
    torch.manual_seed(seed)
    specs = layer_specs(sp, fp, input_size)
    bn_of = dict(LAYER_ORDER)
    sd = OrderedDict()
    for name, (cout, cin, kt, khw, has_bias) in specs.items():
        w = torch.empty(cout, cin, kt, khw, khw)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        sd[name + ".weight"] = w
        if has_bias:
            fan_in = cin * kt * khw * khw
            bound = 1.0 / math.sqrt(fan_in)
            b = torch.empty(cout)
            torch.nn.init.uniform_(b, -bound, bound)
            sd[name + ".bias"] = b
        bn = bn_of[name]
        sd[bn + ".weight"] = torch.ones(cout)
        sd[bn + ".bias"] = torch.zeros(cout)
        sd[bn + ".running_mean"] = torch.zeros(cout)
        sd[bn + ".running_var"] = torch.ones(cout)
        sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    # the reference registers the two laterals last but constructs them last as well -> order already right
    return sd


def param_count(sp, fp):
    return sum(v.numel() for k, v in init_state_dict(sp, fp).items()
               if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 and back, gradient passed straight through (used by the bf16 EMULATION mode below)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


# bf16 emulation: the SAME reference graph with the two operands of every convolution (activations and weights) rounded
# to bf16 and everything else -- accumulation, BatchNorm, gradients -- in fp32.  This is what "bf16 with fp32
# accumulation" means for this module in plain PyTorch; the GPU bf16 path is compared against it with a tight tolerance,
# because the ~5e-3 forward rounding flips ~0.3 % of the ReLU masks and that alone moves weight gradients by 5-8 %
# (relative L2) away from the fp32 reference in ANY bf16 implementation.
EMULATE_BF16 = False

# Optional recorder: when a list, every ReLU layer appends (conv name, pre-activation tensor) during forward().  Used to
# measure the ReLU MARGIN of a test input: min |pre-activation| over all ReLU layers.  A gradient comparison through
# ReLU layers is only meaningful when that margin is well above the arithmetic noise of both sides - an element whose
# pre-activation sits at 1e-7 has an undetermined mask in fp32, and one flipped mask moves single weight-gradient entries by
# ~1e-3 of the maximum (tests/golden/make_golden.py picks input seeds by this margin, the GPU tests re-check it).
PREACT_SINK = None


def _conv_bn(sd, conv, bn, x, training, relu, spatial_pad):
    w = sd[conv + ".weight"]
    b = sd.get(conv + ".bias")
    if EMULATE_BF16:
        w, x = _RoundBF16.apply(w), _RoundBF16.apply(x)
    y = F.conv3d(x, w, b, padding=(0, spatial_pad, spatial_pad))
    rm, rv = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    if training:
        sd[bn + ".num_batches_tracked"] += 1
    y = F.batch_norm(y, rm, rv, sd[bn + ".weight"], sd[bn + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if relu and PREACT_SINK is not None:
        PREACT_SINK.append((conv, y.detach()))
    return F.relu(y) if relu else y


def forward(sd, slow, fast, training):
    """code/helpers/model.py:118-149.  ``sd`` is mutated in train mode (running stats), like the module."""
    slow = _conv_bn(sd, "slow_conv1", "bn_s1", slow, training, True, 1)
    fast = _conv_bn(sd, "fast_conv1", "bn_f1", fast, training, True, 1)
    fuse = _conv_bn(sd, "conv_f2s1", "bn_f2s1", fast, training, True, 0)      # model.py:111-116
    slow = torch.cat([slow, fuse], 1)
    slow = _conv_bn(sd, "slow_conv2", "bn_s2", slow, training, True, 1)
    fast = _conv_bn(sd, "fast_conv2", "bn_f2", fast, training, True, 1)
    fuse = _conv_bn(sd, "conv_f2s2", "bn_f2s2", fast, training, True, 0)
    slow = torch.cat([slow, fuse], 1)
    slow = _conv_bn(sd, "slow_conv3", "bn_s3", slow, training, False, 1)      # no ReLU, model.py:144-148
    fast = _conv_bn(sd, "fast_conv3", "bn_f3", fast, training, False, 1)
    return slow, fast


def temporally_enhance_features(sd, slow_features, fast_features, training):
    """code/helpers/model.py:151-165: list of per-clip dicts {level: [T,256,H,W]} -> {level: [B,256,H,W]}."""
    slow_by_level = {k: [d[k] for d in slow_features] for k in slow_features[0]}
    fast_by_level = {k: [d[k] for d in fast_features] for k in fast_features[0]}
    merged = OrderedDict()
    for key in slow_by_level:
        s = torch.stack(slow_by_level[key]).transpose(1, 2)
        f = torch.stack(fast_by_level[key]).transpose(1, 2)
        s, f = forward(sd, s, f, training)
        merged[key] = torch.cat([s, f], dim=1).squeeze(dim=2)
    return merged


def slice_window(features, center_idx, pathway_size):
    """code/helpers/model.py:242-248."""
    lo = center_idx - math.floor(pathway_size / 2)
    hi = center_idx + math.ceil(pathway_size / 2)
    return OrderedDict((k, v[lo:hi]) for k, v in features.items())


def synthetic_clip(level_shapes, fp, seed=1234, zero_left=0, scale=1.0):
    """SURVEY 8(d) synthetic FPN features: per level randn(fp,256,H,W) seeded 1234+level_index; the first
    ``zero_left`` frames are all-zero (sequence-start padding, model.py:215-225)."""
    feats = OrderedDict()
    for i, (key, (h, w)) in enumerate(level_shapes.items()):
        g = torch.Generator().manual_seed(seed + i)
        x = torch.randn(fp, 256, h, w, generator=g) * scale
        if zero_left:
            x[:zero_left] = 0
        feats[key] = x
    return feats


def module_loss(merged):  # noqa: D401
    """Module-only scalar used to seed gradients in tests/benches: sum_l mean(out_l * r_l) with a fixed seeded
    random projection r_l.  (A plain sum_l mean(out_l^2), as SURVEY 8(d) suggests for benches, is invariant under
    the final BatchNorm -- its true gradient is zero -- so it cannot pin a backward pass.)"""
    total = 0
    for i, v in enumerate(merged.values()):
        g = torch.Generator().manual_seed(777 + i)
        r = torch.randn(v.shape, generator=g).to(device=v.device, dtype=v.dtype)
        total = total + (v * r).mean()
    return total


def cast_inputs(sd, slow_features, fast_features, dtype):
    """The same state dict / clips in another floating-point dtype (fp64 = the exact-arithmetic version of the oracle)."""
    sd2 = OrderedDict((k, (v.to(dtype) if v.is_floating_point() else v.clone())) for k, v in sd.items())
    cast = lambda feats: [OrderedDict((k, v.to(dtype)) for k, v in d.items()) for d in feats]
    return sd2, cast(slow_features), cast(fast_features)


def grads_of(sd, slow_features, fast_features, loss_fn=module_loss, emulate_bf16=False, dtype=None, training=True):
    """Forward + backward through the functional graph (train mode unless ``training=False``: eval-mode BatchNorm, the
    running statistics as constants); returns (merged, loss, {param: grad}, buffers).
    ``emulate_bf16``: see EMULATE_BF16 above.  ``dtype=torch.float64`` runs the same graph in double precision."""
    global EMULATE_BF16
    prev, EMULATE_BF16 = EMULATE_BF16, bool(emulate_bf16)
    if dtype is not None:
        sd, slow_features, fast_features = cast_inputs(sd, slow_features, fast_features, dtype)
    try:
        return _grads_of(sd, slow_features, fast_features, loss_fn, training)
    finally:
        EMULATE_BF16 = prev


def relu_preacts(sd, slow_features, fast_features, dtype=torch.float64, training=True):
    """[(layer, pre-activation tensor)] of every ReLU layer, level by level, in ``dtype``."""
    global PREACT_SINK
    sd2, slow, fast = cast_inputs(sd, slow_features, fast_features, dtype)
    prev, PREACT_SINK = PREACT_SINK, []
    try:
        with torch.no_grad():
            temporally_enhance_features(sd2, slow, fast, training)
        return PREACT_SINK
    finally:
        PREACT_SINK = prev


def relu_margin(sd, slow_features, fast_features):
    """(min |pre-activation| over all ReLU layers in fp64, number of ReLU inputs)."""
    pre = relu_preacts(sd, slow_features, fast_features)
    return min(float(y.abs().min()) for _, y in pre), sum(y.numel() for _, y in pre)


def _grads_of(sd, slow_features, fast_features, loss_fn, training=True):
    leaves = {}
    work = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            leaves[k] = v.clone().requires_grad_(True)
            work[k] = leaves[k]
        else:
            work[k] = v.clone()
    merged = temporally_enhance_features(work, slow_features, fast_features, training)
    loss = loss_fn(merged)
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    grads = {n: (g if g is not None else torch.zeros_like(leaves[n])) for n, g in zip(names, gs)}
    buffers = {k: v for k, v in work.items() if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    return merged, loss, grads, buffers
