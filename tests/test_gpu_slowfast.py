"""GPU parity of the SlowFastLayers drop-in (through the C ABI) against the CPU oracle and the committed golden
fixtures produced by the unmodified reference.  Metric: max|d| / max|ref| per tensor;
<= 1e-4 in the fp32 validation mode, <= 1e-2 in bf16."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import slowfast_oracle as so

pytestmark = pytest.mark.gpu
LEVELS = OrderedDict([("0", (8, 12)), ("pool", (4, 6))])
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def _nerr(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item() / (b.detach().abs().max().item() + 1e-12)


def _check_grad(name, got, ref, precision):
    """fp32 validation mode: tight.  bf16: parameters downstream of the last ReLU (layer 3) see only bf16 rounding
    (<= 3e-2 max-normalised); parameters upstream of a ReLU additionally see ReLU-mask flips caused by the ~5e-3
    forward error -- a fraction p of flipped gradient terms gives a relative error ~sqrt(p) ~ 5-8 % in ANY bf16
    implementation -- so they are held to a relative-L2 bound instead (measured 5-8e-2)."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if precision == "fp32":
        assert _nerr(got, ref) <= 1e-4, (name, _nerr(got, ref))
    elif name.startswith(("fast_conv3", "slow_conv3", "bn_f3", "bn_s3")):
        assert _nerr(got, ref) <= 3e-2, (name, _nerr(got, ref))
    else:
        rel = (got - ref).norm().item() / (ref.norm().item() + 1e-20)
        assert rel <= 0.2, (name, rel)


def _inputs(sp, fp, levels=LEVELS, n_clips=2):
    fast, slow = [], []
    for clip in range(n_clips):
        f = so.synthetic_clip(levels, fp, seed=1234 + 100 * clip, zero_left=(fp // 2 if clip == 1 else 0))
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


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sp,fp", [(1, 8), (3, 7), (2, 16), (4, 32), (1, 1)])
def test_train_forward_backward_matches_reference_golden(sp, fp, precision):
    gold = np.load(os.path.join(GOLDEN, f"slowfast_sp{sp}_fp{fp}.npz"))
    m = _module(sp, fp, precision).train()
    assert sum(p.numel() for p in m.parameters()) == int(gold["n_params"])
    slow, fast = _inputs(sp, fp)
    # views into the fast clip, exactly like the reference's _slice_features (exercises the aliasing path)
    fast_c = _to_cuda(fast)
    slow_c = [so.slice_window(f, fp // 2, sp) for f in fast_c]
    out = m.temporally_enhance_features(slow_c, fast_c)
    tol = TOL[precision]
    assert list(out.keys()) == list(LEVELS.keys())
    for k, v in out.items():
        ref = torch.from_numpy(gold["train_out_" + k])
        assert v.shape == ref.shape and v.dtype == torch.float32
        assert _nerr(v, ref) <= tol, (k, _nerr(v, ref))
    loss = so.module_loss(out)
    assert abs(loss.item() - float(gold["loss"])) <= tol
    loss.backward()
    # oracle gradients (full tensors) on CPU
    sd = so.init_state_dict(sp, fp, seed=63)
    _, _, grads, buffers = so.grads_of(sd, slow, fast)
    for name, p in m.named_parameters():
        ref = grads[name]
        if name.endswith(("conv1.bias", "conv2.bias", "conv3.bias")):
            assert p.grad.abs().max().item() <= 1e-6 + 1e-3 * ref.abs().max().item()   # exactly zero through train BN
            continue
        _check_grad(name, p.grad, ref, precision)
    btol = 1e-4 if precision == "fp32" else 5e-3
    for name, b in m.named_buffers():
        ref = torch.from_numpy(gold["buf_" + name])
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(ref)
        else:
            assert _nerr(b, ref.float()) <= btol, (name, _nerr(b, ref.float()))


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
    for k, v in out.items():
        assert _nerr(v, ref_out[k]) <= 4e-3, (k, _nerr(v, ref_out[k]))
    worst = 0.0
    for name, p in m.named_parameters():
        if name.endswith(("conv1.bias", "conv2.bias", "conv3.bias")):
            continue
        got, ref = p.grad.detach().float().cpu(), grads[name]
        rel = (got - ref).norm().item() / (ref.norm().item() + 1e-20)
        worst = max(worst, rel)
        assert rel <= 0.1, (name, rel)          # measured: <= 2.7e-2 at 24x42, <= 7.2e-2 at the 8x12 / 4x6 toy levels
    assert worst > 0           # the comparison did run on non-trivial gradients


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
    slow, fast = _inputs(sp, fp)
    with torch.no_grad():
        out = m.temporally_enhance_features(_to_cuda(slow), _to_cuda(fast))
    for k, v in out.items():
        assert _nerr(v, torch.from_numpy(gold["eval_out_" + k])) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_api_and_input_grads(precision):
    """forward(slow, fast) on [B,C,T,H,W] tensors requiring grad (the OSVOS caller trains the backbone)."""
    sp, fp = 3, 7
    m = _module(sp, fp, precision).train()
    g = torch.Generator().manual_seed(9)
    fast = torch.randn(2, fp, 256, 6, 10, generator=g)
    slow = fast[:, 2:5].clone()
    sd = so.init_state_dict(sp, fp, seed=63)
    fr, sr = fast.transpose(1, 2).clone().requires_grad_(True), slow.transpose(1, 2).clone().requires_grad_(True)
    s_ref, f_ref = so.forward({k: v.clone() for k, v in sd.items()}, sr, fr, True)
    rs, rf = torch.randn(s_ref.shape, generator=g), torch.randn(f_ref.shape, generator=g)
    ((s_ref * rs).mean() + (f_ref * rf).mean()).backward()
    fc, sc = fast.cuda().transpose(1, 2).requires_grad_(True), slow.cuda().transpose(1, 2).requires_grad_(True)
    s_out, f_out = m(sc, fc)
    assert s_out.shape == s_ref.shape and f_out.shape == f_ref.shape
    tol = TOL[precision]
    assert _nerr(s_out, s_ref) <= tol and _nerr(f_out, f_ref) <= tol
    ((s_out * rs.cuda()).mean() + (f_out * rf.cuda()).mean()).backward()
    _check_grad("fast_in", fc.grad, fr.grad, precision)
    _check_grad("slow_in", sc.grad, sr.grad, precision)


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


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_concurrent_levels_and_pathways_match_single_stream_and_oracle(precision, monkeypatch):
    """With >= 3 pyramid levels the smaller levels run on side streams and the fast pathway of the largest one on its own
    stream (slowfast._level_streams / _pathway_stream).  Two consecutive training steps must leave the same outputs, parameter
    gradients and BatchNorm running statistics (to reduction-order noise) as the single-stream order (the running-stat EMA is
    order-dependent: it is applied in level order after the join), and match the CPU oracle."""
    levels = OrderedDict([("0", (24, 40)), ("1", (12, 20)), ("2", (8, 12)), ("3", (6, 10)), ("pool", (4, 6))])
    sp, fp = 1, 8
    slow, fast = _inputs(sp, fp, levels)
    slow_c, fast_c = _to_cuda(slow), _to_cuda(fast)
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SFVOS_LEVEL_STREAMS", mode)
        monkeypatch.setenv("SFVOS_PATH_STREAMS", mode)
        m = _module(sp, fp, precision).train()
        for _ in range(2):                                   # two steps: the second EMA step sees the first one's buffers
            for p in m.parameters():
                p.grad = None
            out = m.temporally_enhance_features(slow_c, fast_c)
            so.module_loss(out).backward()
        torch.cuda.synchronize()
        results[mode] = (OrderedDict((k, v.detach().clone()) for k, v in out.items()),
                         OrderedDict((n, p.grad.detach().clone()) for n, p in m.named_parameters()),
                         OrderedDict((n, b.detach().clone()) for n, b in m.named_buffers()))
    (o0, g0, b0), (o1, g1, b1) = results["0"], results["1"]
    # (not bit-equal even on one stream: the BatchNorm statistics are reduced with atomics)
    # fp32 pins the ordering tightly; in bf16 a last-bit difference of a statistic moves individual activations by a bf16 ulp.
    # Gradients are compared in relative L2 in both precisions: the statistics are reduced with atomics, and a last-bit
    # difference can flip the mask of a ReLU whose pre-activation sits at ~0, which moves single gradient entries by
    # percents of the max (measured: 3.5e-3 max-normalised on slow_conv2.weight between two single-stream fp32 runs).
    otol, btol, gtol = (1e-5, 1e-6, 1e-2) if precision == "fp32" else (1e-2, 2e-3, 0.1)
    for k in o0:
        assert _nerr(o1[k], o0[k]) <= otol, (k, _nerr(o1[k], o0[k]))
    for n in b0:
        # an EMA applied in the wrong level order (or a lost update) would show at ~momentum * |mean_i - mean_j| ~ 1e-2
        assert (b1[n].double() - b0[n].double()).abs().max().item() <= btol * (1.0 + b0[n].double().abs().max().item()), n
    for n in g0:
        if n.endswith(".weight"):
            ref = g0[n].float()
            assert (g1[n].float() - ref).norm().item() <= gtol * ref.norm().item() + 1e-9, n
    # and against the oracle (one step, fresh state)
    monkeypatch.setenv("SFVOS_LEVEL_STREAMS", "1")
    monkeypatch.setenv("SFVOS_PATH_STREAMS", "1")
    sd = so.init_state_dict(sp, fp, seed=63)
    ref_out, ref_loss, ref_grads, ref_sd = so.grads_of(sd, slow, fast)
    m = _module(sp, fp, precision).train()
    out = m.temporally_enhance_features(slow_c, fast_c)
    so.module_loss(out).backward()
    for k in out:
        assert _nerr(out[k], ref_out[k]) <= TOL[precision], k
    for n, p in m.named_parameters():
        if n.endswith(".weight"):
            ref = ref_grads[n].float()
            rel = (p.grad.detach().float().cpu() - ref).norm().item() / (ref.norm().item() + 1e-20)
            assert rel <= (1e-2 if precision == "fp32" else 0.2), (n, rel)
    for n, b in m.named_buffers():
        if "running" in n:
            assert _nerr(b, ref_sd[n]) <= (1e-4 if precision == "fp32" else 1e-2), n
        elif n.endswith("num_batches_tracked"):
            assert int(b) == 5, (n, int(b))                  # one increment per pyramid level, like the reference
