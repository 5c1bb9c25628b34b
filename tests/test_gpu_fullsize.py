"""Parity at BASELINE.json's FULL sizes (level 0 = 192x336 of a 480x854 DAVIS frame), where the CPU oracle is too slow
to run: size-independent properties of the domain, cross-checks between the two independent kernel families
(tcgen05 bf16 vs fp32 CUDA-core validation mode) and torchvision's own CUDA ROIAlign on the bench's ROI set."""
import math
from collections import OrderedDict

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
H0, W0 = 192, 336


def _nerr(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def _act(ops, B, T, H, W, C, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return ops.Act(torch.randn(B * T * H * W * C, device=DEV, generator=g).bfloat16(), B, T, H, W, C)


@pytest.mark.parametrize("name,T,cin,cout,kt", [("fast_conv1 -> conv_tstack", 8, 256, 32, 3), ("slow_conv2 -> conv_pair", 1, 256, 192, 1),
                                                ("fast_conv2 -> conv_tstack<32>", 6, 32, 32, 3)])
def test_conv_scaling_is_exact_at_full_size(name, T, cin, cout, kt):
    """conv(2x) == 2 conv(x) BIT FOR BIT (a power-of-two scale is exact in bf16 and in the fp32 accumulators), and the
    fused BN statistics scale by 2 and 4: exercises every tile, tap and temporal group of a full 192x336 launch."""
    from sfvos_b200 import ops
    B = 2
    x = _act(ops, B, T, H0, W0, cin, 1)
    x2 = ops.Act((x.buf.float() * 2).bfloat16(), B, T, H0, W0, cin)
    g = torch.Generator(device=DEV).manual_seed(2)
    w = torch.randn(cout, cin, kt, 3, 3, device=DEV, generator=g) / math.sqrt(cin * kt * 9)
    cp = 32 if cin <= 32 else cin
    wp = ops.pack_weights(w, 0, ops.BF16, cp)
    To = T - kt + 1
    ys, stats = [], []
    for inp in (x, x2):
        y = ops.Act.empty(B, To, H0, W0, cout, torch.float32, DEV)
        st = torch.zeros(2 * cout, device=DEV)
        ops.conv(inp, wp, cp, cout, (kt, 3, 3), (0, 1, 1), To, y, umma=True, stats=st)
        ys.append(y.buf); stats.append(st)
    assert torch.equal(ys[1], ys[0] * 2)
    assert ys[0].abs().max().item() > 0.1
    assert _nerr(stats[1][:cout], 2 * stats[0][:cout]) < 1e-5 and _nerr(stats[1][cout:], 4 * stats[0][cout:]) < 1e-5


@pytest.mark.parametrize("name,T,cin,cout,kt", [("fast_conv1 -> wgrad_halo", 8, 256, 32, 3), ("slow_conv1 -> wgrad_pair", 1, 256, 192, 1),
                                                ("fast_conv2 -> wgrad_c32", 6, 32, 32, 3)])
def test_wgrad_is_linear_in_dy_at_full_size(name, T, cin, cout, kt):
    """dw(x, 2 dy) == 2 dw(x, dy) up to the order of the split-K atomics, and dw accumulates (+=) across calls."""
    from sfvos_b200 import ops
    B = 2
    To = T - kt + 1
    x = _act(ops, B, T, H0, W0, cin, 3)
    dy = _act(ops, B, To, H0, W0, cout, 4)
    dy2 = ops.Act((dy.buf.float() * 2).bfloat16(), B, To, H0, W0, cout)
    n = kt * 9 * cin * cout
    a, b = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ops.wgrad(x, dy, (kt, 3, 3), (0, 1, 1), a, umma=True)
    ops.wgrad(x, dy2, (kt, 3, 3), (0, 1, 1), b, umma=True)
    assert a.abs().max().item() > 1.0
    assert _nerr(b, 2 * a) < 1e-5
    ops.wgrad(x, dy, (kt, 3, 3), (0, 1, 1), b, umma=True)
    assert _nerr(b, 3 * a) < 1e-5


def test_module_full_size_bf16_vs_fp32_validation_mode_and_bn_invariants():
    """One full-size level (192x336, B = 2 clips of 8 frames): the tcgen05 bf16 path against the independent fp32
    CUDA-core path (itself pinned to the oracle at small sizes), plus the invariants train-mode BatchNorm gives the
    output of layer 3 at ANY size: per-channel mean 0 and variance 1 (gamma = 1, beta = 0 at init)."""
    from sfvos_b200 import SlowFastLayers
    sp, fp, B = 1, 8, 2
    g = torch.Generator(device=DEV).manual_seed(5)
    fast = [OrderedDict([("0", torch.randn(fp, 256, H0, W0, device=DEV, generator=g))]) for _ in range(B)]
    fast[1]["0"][:4] = 0                                           # sequence start: zero-padded frames
    slow = [OrderedDict((k, v[fp // 2:fp // 2 + sp]) for k, v in f.items()) for f in fast]
    outs, grads = {}, {}
    for prec in ("bf16", "fp32"):
        torch.manual_seed(63)
        m = SlowFastLayers(256, torch.device(DEV), sp, fp).to(DEV).train()
        m.precision = prec
        out = m.temporally_enhance_features(slow, fast)["0"]
        proj = torch.randn(out.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(7))
        (out * proj).mean().backward()
        outs[prec] = out.detach()
        grads[prec] = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        del m
    assert outs["bf16"].shape == (B, 256, H0, W0)
    assert _nerr(outs["bf16"], outs["fp32"]) <= 1e-2
    for prec in outs:
        flat = outs[prec].permute(1, 0, 2, 3).reshape(256, -1).double()
        assert flat.mean(1).abs().max().item() < 1e-4, prec
        assert (flat.var(1, unbiased=False) - 1).abs().max().item() < 1e-3, prec
    for n in grads["fp32"]:
        a, b = grads["bf16"][n].float(), grads["fp32"][n].float()
        if n.endswith(("conv1.bias", "conv2.bias", "conv3.bias")):
            assert a.abs().max().item() == 0 and b.abs().max().item() == 0        # exactly zero through train-mode BN
        elif n.startswith(("fast_conv3", "slow_conv3", "bn_f3", "bn_s3")):
            assert _nerr(a, b) <= 3e-2, (n, _nerr(a, b))
        else:
            rel = (a - b).norm().item() / (b.norm().item() + 1e-20)
            assert rel <= 0.2, (n, rel)


@pytest.mark.parametrize("P,k_per_clip", [(7, 512), (14, 128)])
def test_roi_align_bench_rois_match_torchvision_cuda_op(P, k_per_clip):
    """The bench's ROI set (8 clips, all 4 pyramid levels at full size) against torchvision's own multi-level pooler
    running its CUDA roi_align kernel: forward values and feature gradients, ROI order identical."""
    from torchvision.ops import MultiScaleRoIAlign as TVPool
    from sfvos_b200 import MultiScaleRoIAlign, workload as wl
    B = 8
    g = torch.Generator(device=DEV).manual_seed(9)
    names = wl.POOL_LEVELS
    feats = OrderedDict((k, torch.randn(B, h, w, 256, device=DEV, generator=g).permute(0, 3, 1, 2)) for k, (h, w) in wl.LEVELS.items() if k in names)
    boxes = [b[:k_per_clip].to(DEV) for b in wl.synthetic_rois(B, 512)]
    shapes = [wl.IMAGE_HW] * B
    f_a = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in feats.items())
    f_b = OrderedDict((k, v.contiguous().clone().requires_grad_(True)) for k, v in feats.items())
    ours = MultiScaleRoIAlign(names, P, 2, out_layout="nchw", precision="fp32")(f_a, boxes, shapes)
    ref = TVPool(names, P, 2)(f_b, boxes, shapes)
    assert ours.shape == ref.shape == (B * k_per_clip, 256, P, P)
    assert _nerr(ours, ref) < 2e-6
    wgt = torch.randn(ref.shape, device=DEV, generator=g)
    (ours * wgt).sum().backward()
    (ref * wgt).sum().backward()
    for k in names:
        assert _nerr(f_a[k].grad, f_b[k].grad) < 1e-5, k


def test_box_branch_bench_size_properties_and_cublas_crosscheck():
    """The box branch at the bench size (M = 4096 ROIs, fc6 12544 -> 1024): exact power-of-two scaling of the N-tiled GEMM
    (every pixel tile x N chunk), linear + accumulating weight gradient, and the whole fc6 output against a cuBLAS bf16 GEMM
    with fp32 accumulation of the same bf16 operands (an independent implementation of the same arithmetic)."""
    from sfvos_b200 import ops
    M, K, N = 4096, 12544, 1024
    g = torch.Generator(device=DEV).manual_seed(7)
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K))
    b = torch.randn(N, device=DEV, generator=g)
    wp = ops.pack_weights(w.view(N, K, 1, 1, 1), 0, ops.BF16, K)
    assert torch.equal(wp.view(N, K), w.bfloat16())                       # the vectorised conversion is RNE like torch
    outs = []
    for scale in (1.0, 2.0):
        xa = ops.Act((x.float() * scale).bfloat16().reshape(-1), 1, 1, 1, M, K)
        y = ops.Act.empty(1, 1, 1, M, N, torch.float32, DEV)
        ops.conv(xa, wp, K, N, (1, 1, 1), (0, 0, 0), 1, y, umma=True)
        outs.append(y.buf.view(M, N).clone())
    assert torch.equal(outs[1], 2 * outs[0])
    ref = torch.matmul(x, w.bfloat16().t()).float()                        # cuBLAS, bf16 output of fp32 accumulators
    assert _nerr(outs[0], ref) < 4e-3                                      # bf16 output rounding of the cuBLAS side
    ref32 = x.float() @ w.bfloat16().float().t()
    assert _nerr(outs[0], ref32) < 2e-5
    # bias + ReLU epilogue on all N chunks
    y = ops.Act.empty(1, 1, 1, M, N, torch.float32, DEV)
    ops.conv(ops.Act(x.reshape(-1), 1, 1, 1, M, K), wp, K, N, (1, 1, 1), (0, 0, 0), 1, y, umma=True, relu=True, shift=b)
    assert _nerr(y.buf.view(M, N), torch.relu(ref32 + b)) < 2e-5
    # weight gradient: linear in dy, accumulating
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    dwp = torch.zeros(K * N, device=DEV)
    xa = ops.Act(x.reshape(-1), 1, 1, 1, M, K)
    ops.wgrad(xa, ops.Act(dy.reshape(-1), 1, 1, 1, M, N), (1, 1, 1), (0, 0, 0), dwp, umma=True)
    one = dwp.clone()
    ops.wgrad(xa, ops.Act((dy.float() * 2).bfloat16().reshape(-1), 1, 1, 1, M, N), (1, 1, 1), (0, 0, 0), dwp, umma=True)
    assert _nerr(dwp, 3 * one) < 1e-6
    refw = (x.float().t() @ dy.float())                                    # [K, N] = the packed gradient layout
    assert _nerr(one.view(K, N), refw) < 2e-5


def test_sequence_sweep_full_size_level_is_bit_identical_to_windows():
    """Eval-mode sequence sweep at a full-size pyramid level (96x168): identical bits to the per-window path."""
    from sfvos_b200 import SlowFastLayers
    torch.manual_seed(63)
    sp, fp, n = 1, 8, 10
    mod = SlowFastLayers(256, torch.device(DEV), sp, fp).cuda().eval()
    g = torch.Generator(device=DEV).manual_seed(3)
    frames = OrderedDict([("1", torch.randn(n, 256, 96, 168, device=DEV, generator=g))])
    seq = mod.temporally_enhance_sequence(frames)["1"]
    zero = torch.zeros_like(frames["1"][0])
    for t in (0, 4, n - 1):
        idx = range(t - fp // 2, t + (fp + 1) // 2)
        fast = OrderedDict([("1", torch.stack([frames["1"][i] if 0 <= i < n else zero for i in idx]))])
        slow = OrderedDict([("1", fast["1"][fp // 2 - sp // 2:fp // 2 - sp // 2 + sp])])
        with torch.no_grad():
            win = mod.temporally_enhance_features([slow], [fast])["1"]
        assert torch.equal(seq[t:t + 1], win), t


# ---- the ORACLE itself at BASELINE sizes: its torch graph runs on the GPU in IEEE fp32 (conftest.force_ieee_fp32) -------------
@pytest.mark.parametrize("cfg,sp,fp,hw,B", [("C2 level 0", 1, 8, (192, 336), 2), ("C5 level 1", 2, 16, (96, 168), 2),
                                            ("C3 level 1", 4, 32, (96, 168), 1)])
def test_module_matches_the_oracle_at_baseline_sizes(cfg, sp, fp, hw, B):
    """tcgen05 path and validation mode against oracle.slowfast_oracle.grads_of evaluated on CUDA tensors (the same functional
    graph the CPU tests pin to the reference, cuDNN IEEE fp32 instead of oneDNN), at sizes where every kernel runs many waves
    of full and partial tiles: outputs <= 1e-2 (bf16) / <= 1e-4 (validation mode) max-normalised; bf16 gradients against the
    bf16-EMULATED oracle (the reference graph with conv operands rounded to bf16) in relative L2.  At these sizes (10^7..10^8
    ReLU inputs) no input has a ReLU margin, so the validation mode's gradients upstream of a ReLU are held in relative L2 as
    well: a handful of elements with |pre-activation| < 1e-6 have masks that cuDNN's fp32 summation order decides."""
    from conftest import report
    from oracle import slowfast_oracle as so
    from sfvos_b200 import SlowFastLayers
    g = torch.Generator(device=DEV).manual_seed(11)
    fast = [OrderedDict([("0", torch.randn(fp, 256, hw[0], hw[1], device=DEV, generator=g))]) for _ in range(B)]
    if B > 1:
        fast[1]["0"][:fp // 2] = 0                                  # sequence start: zero-padded frames
    slow = [so.slice_window(f, fp // 2, sp) for f in fast]
    sd = OrderedDict((k, v.to(DEV)) for k, v in so.init_state_dict(sp, fp, seed=63).items())
    ref_out, _, ref_grads, _ = so.grads_of(sd, slow, fast)
    ref_out = {k: v.detach() for k, v in ref_out.items()}
    emu_out, _, emu_grads, _ = so.grads_of(sd, slow, fast, emulate_bf16=True)
    emu_out = {k: v.detach() for k, v in emu_out.items()}
    torch.cuda.empty_cache()
    skip = ("conv1.bias", "conv2.bias", "conv3.bias")
    for prec in ("bf16", "fp32"):
        torch.manual_seed(63)
        m = SlowFastLayers(256, torch.device(DEV), sp, fp).to(DEV).train()
        m.precision = prec
        out = m.temporally_enhance_features(slow, fast)
        so.module_loss(out).backward()
        torch.cuda.synchronize()
        e_out = _nerr(out["0"], ref_out["0"])
        assert e_out <= (1e-2 if prec == "bf16" else 1e-4), (cfg, prec, e_out)
        rec = {"out_vs_fp32_oracle": e_out}
        if prec == "bf16":
            rec["out_vs_bf16_emulated"] = _nerr(out["0"], emu_out["0"])
            assert rec["out_vs_bf16_emulated"] <= 4e-3, (cfg, rec)
        worst3, worst = 0.0, 0.0
        for n, p in m.named_parameters():
            if n.endswith(skip):
                continue
            ref = (emu_grads if prec == "bf16" else ref_grads)[n]
            rel = (p.grad.double() - ref.double()).norm().item() / (ref.double().norm().item() + 1e-30)
            if n.startswith(("fast_conv3", "slow_conv3", "bn_f3", "bn_s3")):
                worst3 = max(worst3, _nerr(p.grad, ref))            # downstream of every ReLU: no masks involved
            else:
                worst = max(worst, rel)
        rec["layer3_grad_max_norm"], rec["other_grads_rel_l2"] = worst3, worst
        report("fullsize_oracle", cfg=cfg, precision=prec, **rec)
        # measured (round 2, B200, profiles/parity_r2.jsonl): bf16 outputs 5.4-6.7e-3 vs the fp32 oracle and 1.9-2.0e-3 vs the
        # emulation, layer-3 gradients 2.4-3.2e-3, other gradients 2.1-3.1e-2 (relative L2 vs the emulation); validation mode
        # outputs 2.7-4.3e-6, layer-3 gradients 3.1-5.1e-6, other gradients 2.2-3.5e-3 (mask flips against cuDNN's fp32 sums).
        # Bounds = 2-3x the measured values.
        assert worst3 <= (1e-2 if prec == "bf16" else 2e-5), (cfg, prec, rec)
        assert worst <= (6e-2 if prec == "bf16" else 1e-2), (cfg, prec, rec)
        del m, out
        torch.cuda.empty_cache()
