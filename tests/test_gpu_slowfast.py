"""GPU parity of the SlowFastLayers drop-in (through the C ABI) against the CPU oracle and the committed golden
fixtures produced by the unmodified reference.  Metric: max|d| / max|ref| per tensor;
<= 1e-4 in the fp32 validation mode, <= 1e-2 in bf16.

Gradients and ReLU masks.  A weight gradient upstream of a ReLU is a DIScontinuous function of the forward pass: an
element whose pre-activation sits within arithmetic noise of zero has an undetermined mask, and one flipped mask moves
single gradient entries by ~1e-3 of the maximum (round 1's red test: (3,7) had a pre-activation at 2.7e-7 and the GPU's
atomically ordered fp32 sums flipped it on some runs).  What makes the fp32 comparison below sound:
  * the inputs of every golden fixture are chosen for their ReLU margin (min |pre-activation| >= 1e-5 in the fp64 oracle,
    tests/golden/make_golden.py) and the margin is re-checked here;
  * the validation mode accumulates in fp64 and rounds once (pre-activation noise ~2e-7, 50x below the margin) and uses no
    float atomics, so its masks equal the exact ones and two runs are bit-identical (asserted below);
  * gradients are compared with the fp64 oracle (the exact values for those masks), bound 1e-5, measured ~1e-7.
Every measured value is written to gpurun_out/parity_report.jsonl (conftest.report) and summarised in DESIGN.md section 5."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import GOLDEN, report
from oracle import slowfast_oracle as so

pytestmark = pytest.mark.gpu
LEVELS = OrderedDict([("0", (8, 12)), ("pool", (4, 6))])
TOL = {"fp32": 1e-4, "bf16": 1e-2}
BIAS_GRADS = ("conv1.bias", "conv2.bias", "conv3.bias")      # exactly zero through train-mode BN


def _nerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


def _rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).norm().item() / (b.norm().item() + 1e-30)


def _check_grad(name, got, ref, precision):
    """fp32 validation mode: 1e-5 max-normalised against the fp64 oracle (measured <= 3e-7: plain rounding, the masks are exact).
    bf16: parameters downstream of the last ReLU (layer 3) see only bf16 rounding (<= 3e-2 max-normalised, measured 5-7e-3);
    parameters upstream of a ReLU additionally see ReLU-mask flips caused by the ~5e-3 forward error -- a fraction p of
    flipped gradient terms gives a relative error ~sqrt(p) in ANY bf16 implementation (the bf16-emulated oracle measures
    the same, see test_bf16_path_matches_bf16_emulated_oracle) -- so they are held to a relative-L2 bound (measured 5-8e-2)."""
    if precision == "fp32":
        e = _nerr(got, ref)
        assert e <= 1e-5, (name, e)
        return e
    if name.startswith(("fast_conv3", "slow_conv3", "bn_f3", "bn_s3")):
        e = _nerr(got, ref)
        assert e <= 3e-2, (name, e)
        return e
    e = _rel_l2(got, ref)
    assert e <= 0.2, (name, e)
    return e


def _inputs(sp, fp, levels=LEVELS, n_clips=2, seed0=1234):
    fast, slow = [], []
    for clip in range(n_clips):
        f = so.synthetic_clip(levels, fp, seed=seed0 + 100 * clip, zero_left=(fp // 2 if clip == 1 else 0))
        fast.append(f)
        slow.append(so.slice_window(f, fp // 2, sp))
    return slow, fast


def _to_cuda(list_of_dicts):
    return [OrderedDict((k, v.cuda()) for k, v in d.items()) for d in list_of_dicts]


def _module(sp, fp, precision):
    from sfvos_b200 import SlowFastLayers
    torch.manual_seed(63)
    m = SlowFastLayers(256, torch.device("cuda"), sp, fp).cuda()
    m.precision = precision
    return m


def _train_step(m, slow_c, fast_c):
    for p in m.parameters():
        p.grad = None
    out = m.temporally_enhance_features(slow_c, fast_c)
    so.module_loss(out).backward()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sp,fp", [(1, 8), (3, 7), (2, 16), (4, 32), (1, 1)])
def test_train_forward_backward_matches_reference_golden(sp, fp, precision):
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    m = _module(sp, fp, precision).train()
    assert sum(p.numel() for p in m.parameters()) == int(gold["n_params"])
    slow, fast = _inputs(sp, fp, seed0=int(gold["input_seed"]))
    sd = so.init_state_dict(sp, fp, seed=63)
    margin, _ = so.relu_margin(sd, slow, fast)
    assert margin >= 1e-5, margin                     # the fixture's inputs keep every ReLU input away from zero
    # views into the fast clip, exactly like the reference's _slice_features (exercises the aliasing path)
    fast_c = _to_cuda(fast)
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    out = m.temporally_enhance_features(slow_c, fast_c)
    tol = TOL[precision]
    assert list(out.keys()) == list(LEVELS.keys())
    measured = {}
    for k, v in out.items():
        ref = torch.from_numpy(gold["train_out_" + k])
        assert v.shape == ref.shape and v.dtype == torch.float32
        measured["out_" + k] = _nerr(v, ref)
        assert measured["out_" + k] <= tol, (k, measured["out_" + k])
    loss = so.module_loss(out)
    assert abs(loss.item() - float(gold["loss"])) <= tol
    loss.backward()
    # oracle gradients: fp64 (exact for these masks) in the validation mode, the reference's own fp32 arithmetic for bf16
    _, _, grads, buffers = so.grads_of(sd, slow, fast, dtype=torch.float64 if precision == "fp32" else None)
    worst = 0.0
    for name, p in m.named_parameters():
        ref = grads[name]
        if name.endswith(BIAS_GRADS):
            assert p.grad.abs().max().item() <= 1e-6 + 1e-3 * float(ref.abs().max())   # exactly zero through train BN
            continue
        worst = max(worst, _check_grad(name, p.grad, ref, precision))
        # and the REFERENCE's own gradient samples (fp32 CPU run of the unmodified module, make_golden.py)
        if precision == "fp32":
            flat = p.grad.detach().flatten().cpu()
            g = torch.Generator().manual_seed(flat.numel() % 9973 + 17)
            idx = torch.randint(0, flat.numel(), (64,), generator=g)
            smp = torch.from_numpy(gold["grad_smp_" + name])
            assert (flat[idx] - smp).abs().max().item() <= 1e-4 * float(ref.abs().max()), name
    measured["grad_worst"] = worst
    btol = 1e-4 if precision == "fp32" else 5e-3
    for name, b in m.named_buffers():
        ref = torch.from_numpy(gold["buf_" + name])
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(ref)
        else:
            measured["buf_worst"] = max(measured.get("buf_worst", 0.0), _nerr(b, ref.float()))
            assert _nerr(b, ref.float()) <= btol, (name, _nerr(b, ref.float()))
    report("train_golden", sp=sp, fp=fp, precision=precision, relu_margin=margin, **measured)


@pytest.mark.parametrize("sp,fp", [(3, 7), (2, 16)])
def test_fp32_validation_mode_is_bit_reproducible(sp, fp):
    """No float atomics in the validation mode: two training steps on the same inputs give bit-identical outputs, parameter
    gradients and BatchNorm buffers (round 1: atomically ordered sums differed in the last bits and could flip a ReLU)."""
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    slow, fast = _inputs(sp, fp, seed0=int(gold["input_seed"]))
    fast_c = _to_cuda(fast)
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    runs = []
    for _ in range(3):
        m = _module(sp, fp, "fp32").train()
        out = _train_step(m, slow_c, fast_c)
        runs.append(([v.detach().clone() for v in out.values()], [p.grad.detach().clone() for p in m.parameters()],
                     [b.detach().clone() for b in m.buffers()]))
    for other in runs[1:]:
        for a_list, b_list in zip(runs[0], other):
            for a, b in zip(a_list, b_list):
                assert torch.equal(a, b)


@pytest.mark.parametrize("sp,fp,levels", [(1, 8, OrderedDict([("0", (24, 42)), ("pool", (6, 11))])),
                                          (2, 16, LEVELS), (3, 7, LEVELS)])
def test_bf16_path_matches_bf16_emulated_oracle(sp, fp, levels):
    """The bf16 product path against the oracle run in bf16-EMULATION mode (the reference graph with both operands of
    every convolution rounded to bf16, fp32 accumulation / BatchNorm / gradients).  Against the fp32 oracle the weight
    gradients of the layers upstream of a ReLU sit 5-8 % (relative L2) away in ANY bf16 implementation (ReLU-mask flips:
    the emulation itself measures exactly that); against the emulation the kernels must be within
    plain rounding distance, which is what pins the backward pass of the product path."""
    m = _module(sp, fp, "bf16").train()
    slow, fast = _inputs(sp, fp, levels)
    fast_c = _to_cuda(fast)
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    out = m.temporally_enhance_features(slow_c, fast_c)
    so.module_loss(out).backward()
    sd = so.init_state_dict(sp, fp, seed=63)
    ref_out, _, grads, _ = so.grads_of(sd, slow, fast, emulate_bf16=True)
    worst_out = 0.0
    for k, v in out.items():
        worst_out = max(worst_out, _nerr(v, ref_out[k]))
        assert _nerr(v, ref_out[k]) <= 4e-3, (k, _nerr(v, ref_out[k]))      # measured 1-2e-3
    worst = 0.0
    for name, p in m.named_parameters():
        if name.endswith(BIAS_GRADS):
            continue
        rel = _rel_l2(p.grad, grads[name])
        worst = max(worst, rel)
        # measured (round 1, run37): <= 2.7e-2 at 24x42, <= 7.2e-2 at the 8x12 / 4x6 toy levels -- a few hundred pixels per
        # channel, so the handful of masks that differ between two bf16 roundings of the same value weigh percents
        assert rel <= 0.1, (name, rel)
    assert worst > 0           # the comparison did run on non-trivial gradients
    report("bf16_vs_emulated", sp=sp, fp=fp, level0=str(list(levels.values())[0]), out_worst=worst_out, grad_rel_l2_worst=worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sp,fp", [(1, 8), (4, 32)])
def test_eval_forward_matches_reference_golden(sp, fp, precision):
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    m = _module(sp, fp, precision)
    sd = m.state_dict()
    for k in sd:
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            sd[k] = torch.from_numpy(gold["buf_" + k]).to(sd[k].dtype)
    m.load_state_dict(sd)
    m.eval()
    slow, fast = _inputs(sp, fp, seed0=int(gold["input_seed"]))
    with torch.no_grad():
        out = m.temporally_enhance_features(_to_cuda(slow), _to_cuda(fast))
    worst = 0.0
    for k, v in out.items():
        worst = max(worst, _nerr(v, torch.from_numpy(gold["eval_out_" + k])))
        assert _nerr(v, torch.from_numpy(gold["eval_out_" + k])) <= TOL[precision]
    report("eval_golden", sp=sp, fp=fp, precision=precision, out_worst=worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_eval_mode_backward_matches_oracle(precision):
    """eval() with autograd enabled (fine-tuning against frozen BatchNorm statistics): the reference nn.Module
    backpropagates through eval-mode BN, where the statistics are constants -- and the conv biases DO get a gradient."""
    sp, fp = 3, 7
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    slow, fast = _inputs(sp, fp, seed0=int(gold["input_seed"]))
    sd = so.init_state_dict(sp, fp, seed=63)
    for k in sd:                                               # non-trivial running statistics: the golden train step's
        if k.endswith(("running_mean", "running_var")):
            sd[k] = torch.from_numpy(gold["buf_" + k]).clone()
    margin = min(float(y.abs().min()) for _, y in so.relu_preacts(sd, slow, fast, training=False))
    m = _module(sp, fp, precision)
    m.load_state_dict(sd)
    m.eval()
    out = m.temporally_enhance_features(_to_cuda(slow), _to_cuda(fast))
    assert all(v.requires_grad for v in out.values())
    so.module_loss(out).backward()
    ref_out, _, grads, _ = so.grads_of(sd, slow, fast, training=False, dtype=torch.float64 if precision == "fp32" else None)
    for k, v in out.items():
        assert _nerr(v, ref_out[k]) <= TOL[precision], (k, _nerr(v, ref_out[k]))
    worst = 0.0
    for name, p in m.named_parameters():
        ref = grads[name]
        if precision == "fp32" and margin >= 2e-6:
            e = _nerr(p.grad, ref)
            assert e <= 1e-5, (name, e)
        else:                                  # bf16 (or an fp32 input without margin): mask flips -> relative L2
            e = _rel_l2(p.grad, ref)
            assert e <= (0.2 if precision == "bf16" else 1e-2), (name, e)
        worst = max(worst, e)
    report("eval_backward", precision=precision, relu_margin=margin, grad_worst=worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_api_and_input_grads(precision):
    """forward(slow, fast) on [B,C,T,H,W] tensors requiring grad (the OSVOS caller trains the backbone)."""
    sp, fp = 3, 7
    m = _module(sp, fp, precision).train()
    g = torch.Generator().manual_seed(9)
    fast = torch.randn(2, fp, 256, 6, 10, generator=g)
    slow = fast[:, 2:5].clone()
    sd = so.init_state_dict(sp, fp, seed=63)
    dt = torch.float64 if precision == "fp32" else torch.float32
    sd_ref = OrderedDict((k, (v.to(dt) if v.is_floating_point() else v.clone())) for k, v in sd.items())
    fr = fast.transpose(1, 2).to(dt).clone().requires_grad_(True)
    sr = slow.transpose(1, 2).to(dt).clone().requires_grad_(True)
    s_ref, f_ref = so.forward(sd_ref, sr, fr, True)
    rs, rf = torch.randn(s_ref.shape, generator=g), torch.randn(f_ref.shape, generator=g)
    ((s_ref * rs.to(dt)).mean() + (f_ref * rf.to(dt)).mean()).backward()
    fc, sc = fast.cuda().transpose(1, 2).requires_grad_(True), slow.cuda().transpose(1, 2).requires_grad_(True)
    s_out, f_out = m(sc, fc)
    assert s_out.shape == s_ref.shape and f_out.shape == f_ref.shape
    tol = TOL[precision]
    assert _nerr(s_out, s_ref) <= tol and _nerr(f_out, f_ref) <= tol
    ((s_out * rs.cuda()).mean() + (f_out * rf.cuda()).mean()).backward()
    # (random inputs without a margin guarantee: input gradients are compared in relative L2 - a flipped mask changes a few
    # entries, not the norm; measured 2e-7 in fp32, 6e-2 in bf16)
    for name, got, ref in (("fast_in", fc.grad, fr.grad), ("slow_in", sc.grad, sr.grad)):
        e = _rel_l2(got, ref)
        assert e <= (1e-3 if precision == "fp32" else 0.2), (name, e)
        report("input_grads", precision=precision, tensor=name, rel_l2=e)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fuse_is_the_reference_lateral_connection_and_differentiable(precision):
    """SlowFastLayers.fuse(slow, fast, conv, bn) (model.py:111-116) as a public op: cat([slow, relu(bn(conv(fast)))], 1)."""
    import torch.nn.functional as F
    sp, fp = 1, 8
    m = _module(sp, fp, precision).train()
    g = torch.Generator().manual_seed(21)
    fast = torch.randn(2, 32, 6, 6, 10, generator=g)
    slow = torch.randn(2, 192, 1, 6, 10, generator=g)
    w = m.conv_f2s1.weight.detach().cpu().double().requires_grad_(True)
    gam = torch.ones(64, dtype=torch.float64, requires_grad=True)
    bet = torch.zeros(64, dtype=torch.float64, requires_grad=True)
    fr = fast.double().requires_grad_(True)
    lat = F.relu(F.batch_norm(F.conv3d(fr, w), None, None, gam, bet, True, 0.1, 1e-5))
    ref = torch.cat([slow.double(), lat], 1)
    r = torch.randn(ref.shape, generator=g)
    (ref * r.double()).mean().backward()
    fc = fast.cuda().requires_grad_(True)
    cat, fast_back = m.fuse(slow.cuda(), fc, m.conv_f2s1, m.bn_f2s1)
    assert fast_back is fc and cat.shape == ref.shape
    assert _nerr(cat, ref) <= TOL[precision]
    (cat * r.cuda()).mean().backward()
    for name, got, want in (("weight", m.conv_f2s1.weight.grad, w.grad), ("gamma", m.bn_f2s1.weight.grad, gam.grad),
                            ("beta", m.bn_f2s1.bias.grad, bet.grad), ("fast", fc.grad, fr.grad)):
        e = _rel_l2(got, want)
        assert e <= (1e-3 if precision == "fp32" else 0.1), (name, e)
        report("fuse", precision=precision, tensor=name, rel_l2=e)
    with pytest.raises(ValueError):
        m.fuse(slow.cuda(), fc, m.fast_conv2, m.bn_f2)


def test_state_dict_roundtrip_and_no_cpu_fallback():
    from sfvos_b200 import SlowFastLayers
    m = _module(1, 8, "bf16")
    ref_keys = list(so.init_state_dict(1, 8).keys())
    assert list(m.state_dict().keys()) == ref_keys
    m2 = SlowFastLayers(256, torch.device("cuda"), 1, 8).cuda()
    m2.load_state_dict(m.state_dict())
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    cpu = SlowFastLayers(256, torch.device("cpu"), 1, 8)
    x = torch.randn(1, 256, 8, 4, 6)
    with pytest.raises(RuntimeError):
        cpu(x[:, :, 4:5], x)


# 5 pyramid levels small enough that an input with a ReLU margin exists (seed found by scanning 1234 + 1000 k on the fp64
# oracle: 708,864 ReLU inputs, margin 5.4e-6); the fp32 half of the test below asserts the margin before relying on it
CONC_LEVELS = {"fp32": (OrderedDict([("0", (12, 20)), ("1", (8, 12)), ("2", (6, 10)), ("3", (4, 6)), ("pool", (2, 3))]), 5234),
               "bf16": (OrderedDict([("0", (24, 40)), ("1", (12, 20)), ("2", (8, 12)), ("3", (6, 10)), ("pool", (4, 6))]), 1234)}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_concurrent_levels_and_pathways_match_single_stream_and_oracle(precision, monkeypatch):
    """With >= 3 pyramid levels the smaller levels run on side streams and the fast pathway of the largest one on its own
    stream (slowfast._level_streams / _pathway_stream).  Two consecutive training steps must leave the same outputs, parameter
    gradients and BatchNorm running statistics as the single-stream order (the running-stat EMA is order-dependent: it is
    applied in level order after the join) and match the CPU oracle.  fp32 validation mode: BIT-identical (no atomics; every
    level accumulates into its own gradient slot and the slots are summed in level order).  bf16: to reduction-order noise."""
    levels, seed0 = CONC_LEVELS[precision]
    sp, fp = 1, 8
    slow, fast = _inputs(sp, fp, levels, seed0=seed0)
    slow_c, fast_c = _to_cuda(slow), _to_cuda(fast)
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SFVOS_LEVEL_STREAMS", mode)
        monkeypatch.setenv("SFVOS_PATH_STREAMS", mode)
        m = _module(sp, fp, precision).train()
        for _ in range(2):                                   # two steps: the second EMA step sees the first one's buffers
            out = _train_step(m, slow_c, fast_c)
        results[mode] = (OrderedDict((k, v.detach().clone()) for k, v in out.items()),
                         OrderedDict((n, p.grad.detach().clone()) for n, p in m.named_parameters()),
                         OrderedDict((n, b.detach().clone()) for n, b in m.named_buffers()))
    (o0, g0, b0), (o1, g1, b1) = results["0"], results["1"]
    if precision == "fp32":
        for d0, d1 in ((o0, o1), (g0, g1), (b0, b1)):
            for k in d0:
                assert torch.equal(d0[k], d1[k]), k
    else:
        # the statistics are reduced with atomics on the product path: a last-bit difference of a statistic moves individual
        # activations by a bf16 ulp and can flip single masks -> outputs to 1e-2, gradients in relative L2
        # (measured: outputs 4e-3, buffers 2e-4, gradients <= 4e-2)
        for k in o0:
            assert _nerr(o1[k], o0[k]) <= 1e-2, (k, _nerr(o1[k], o0[k]))
        for n in b0:
            # an EMA applied in the wrong level order (or a lost update) would show at ~momentum * |mean_i - mean_j| ~ 1e-2
            assert (b1[n].double() - b0[n].double()).abs().max().item() <= 2e-3 * (1.0 + b0[n].double().abs().max().item()), n
        for n in g0:
            if n.endswith(".weight"):
                assert _rel_l2(g1[n], g0[n]) <= 0.1, (n, _rel_l2(g1[n], g0[n]))
    # and against the oracle (one step, fresh state)
    monkeypatch.setenv("SFVOS_LEVEL_STREAMS", "1")
    monkeypatch.setenv("SFVOS_PATH_STREAMS", "1")
    sd = so.init_state_dict(sp, fp, seed=63)
    if precision == "fp32":
        margin, n_relu = so.relu_margin(sd, slow, fast)
        assert margin >= 5e-6, (margin, n_relu)
    ref_out, ref_loss, ref_grads, ref_sd = so.grads_of(sd, slow, fast, dtype=torch.float64 if precision == "fp32" else None)
    m = _module(sp, fp, precision).train()
    out = _train_step(m, slow_c, fast_c)
    for k in out:
        assert _nerr(out[k], ref_out[k]) <= TOL[precision], k
    worst = 0.0
    for n, p in m.named_parameters():
        if n.endswith(BIAS_GRADS):
            continue
        if precision == "fp32":
            e = _nerr(p.grad, ref_grads[n])
            assert e <= 1e-5, (n, e)
        else:
            e = _rel_l2(p.grad, ref_grads[n])
            assert e <= 0.2, (n, e)
        worst = max(worst, e)
    for n, b in m.named_buffers():
        if "running" in n:
            assert _nerr(b, ref_sd[n]) <= (1e-4 if precision == "fp32" else 1e-2), n
        elif n.endswith("num_batches_tracked"):
            assert int(b) == 5, (n, int(b))                  # one increment per pyramid level, like the reference
    report("concurrent_levels", precision=precision, grad_worst=worst)


def test_bn_statistics_survive_a_large_mean():
    """|mean| / std = 1e3 per channel: sumsq/n - mean^2 cancels 6 digits.  The validation mode sums in fp64 (fixed order), so
    scale / shift / running_var still match torch's batch_norm; with fp32 sums the variance would be off by >10 %."""
    import torch.nn.functional as F
    from sfvos_b200 import ops
    B, T, H, W, C = 2, 3, 24, 42, 64
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, T, H, W, C, generator=g) * 1e-2 + 10.0 + 0.1 * torch.arange(C).float()).cuda()
    xa = ops.Act(x.reshape(-1).contiguous(), B, T, H, W, C)
    stats = torch.empty(2 * C, dtype=torch.float64, device="cuda")
    ops.channel_stats(xa, stats)
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    bn4 = torch.empty(4 * C, device="cuda")
    ops.bn_finalize(stats, xa.npix, None, gamma, beta, rm, rv, nbt, 0.1, 1e-5, bn4)
    xd = x.double().reshape(-1, C)
    var = xd.var(0, unbiased=False)
    rstd_ref = 1.0 / torch.sqrt(var + 1e-5)
    e_rstd = ((bn4[3 * C:].double() - rstd_ref).abs() / rstd_ref).max().item()
    e_rv = ((rv.double() - (0.9 + 0.1 * xd.var(0, unbiased=True))).abs()).max().item()
    assert e_rstd <= 1e-5 and e_rv <= 2e-7, (e_rstd, e_rv)
    report("bn_large_mean", rstd_rel_err=e_rstd, running_var_abs_err=e_rv)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_a_non_current_device():
    """The reference API lets the module live on any device; libsfvos launches on the CURRENT one, so every entry makes the
    tensors' device current (ops.device_guard).  Module on cuda:1 while cuda:0 is current == the same module on cuda:0."""
    from sfvos_b200 import SlowFastLayers
    sp, fp = 1, 8
    slow, fast = _inputs(sp, fp)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        torch.manual_seed(63)
        m = SlowFastLayers(256, torch.device(dev), sp, fp).to(dev).train()
        m.precision = "fp32"
        fast_d = [OrderedDict((k, v.to(dev)) for k, v in d.items()) for d in fast]
        slow_d = [so.slice_window(f, fp // 2, sp) for f in fast_d]
        assert torch.cuda.current_device() == 0
        out = m.temporally_enhance_features(slow_d, fast_d)
        so.module_loss(out).backward()
        torch.cuda.synchronize(dev)
        assert all(v.device == torch.device(dev) for v in out.values())
        outs.append(([v.detach().cpu() for v in out.values()], [p.grad.detach().cpu() for p in m.parameters()]))
        m.eval()
        with torch.no_grad():
            seq = m.temporally_enhance_sequence(OrderedDict((k, v.to(dev)) for k, v in fast[0].items()))
        assert all(v.device == torch.device(dev) for v in seq.values())
    for a_list, b_list in zip(outs[0], outs[1]):
        for a, b in zip(a_list, b_list):
            assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sp,fp,dtype", [(1, 8, torch.float32), (3, 7, torch.bfloat16)])
def test_consecutive_windows_of_a_sequence_are_converted_once(precision, sp, fp, dtype, monkeypatch):
    """Clips that are consecutive windows of ONE sequence tensor (clip b = frames [b, b+fp): the reference's own clips,
    model.py:318-337) take the de-duplicated input path: every frame is laid out channels-last once and the clips are
    overlapping views (batch stride = one frame).  Same results as converting every window separately (SFVOS_WINDOW_DEDUP=0),
    bit for bit in the validation mode; f32 and bf16 feature tensors."""
    from sfvos_b200 import ops
    n_win = 5
    levels = OrderedDict([("0", (13, 21)), ("1", (8, 12)), ("pool", (4, 6))])
    g = torch.Generator().manual_seed(31)
    seq = OrderedDict((k, torch.randn(n_win + fp - 1, 256, h, w, generator=g).to(dtype).cuda()) for k, (h, w) in levels.items())
    fast_c = [OrderedDict((k, v[b:b + fp]) for k, v in seq.items()) for b in range(n_win)]
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SFVOS_WINDOW_DEDUP", flag)
        m = _module(sp, fp, precision).train()
        before = ops.launches()
        out = _train_step(m, slow_c, fast_c)
        res[flag] = ([v.detach().clone() for v in out.values()], [p.grad.detach().clone() for p in m.parameters()], ops.launches() - before)
    assert res["1"][2] < res["0"][2]                          # fewer layout launches: one per level instead of one per clip
    for a_list, b_list in zip(res["1"][:2], res["0"][:2]):
        for a, b in zip(a_list, b_list):
            if precision == "fp32":
                assert torch.equal(a, b)
            else:
                assert _rel_l2(a, b) <= 5e-2 or float(b.abs().max()) == 0
    # and against the oracle on the same windows
    slow = [OrderedDict((k, v.float().cpu()) for k, v in d.items()) for d in slow_c]
    fast = [OrderedDict((k, v.float().cpu()) for k, v in d.items()) for d in fast_c]
    ref_out, _, _, _ = so.grads_of(so.init_state_dict(sp, fp, seed=63), slow, fast)
    for v, r in zip(res["1"][0], ref_out.values()):
        assert _nerr(v, r) <= TOL[precision]
