"""GPU parity tests of the individual kernels, called through the C ABI (ctypes), against plain PyTorch fp32
references of the same op on the same (bf16-rounded) inputs.  Tolerances are max-normalised: max|d|/max|ref|."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from sfvos_b200 import ops
    return ops


def _err(a, b):
    return (a.float() - b.float()).abs().max().item() / (b.float().abs().max().item() + 1e-12)


def _mk_act(ops, t_ndhwc, dtype, cstride=None, ch_off=0):
    """torch [B,T,H,W,C] -> Act holding it (optionally inside a wider buffer at channel offset ch_off)."""
    B, T, H, W, C = t_ndhwc.shape
    cs = cstride or C
    buf = torch.zeros(B * T * H * W, cs, dtype=dtype, device=DEV)
    buf[:, ch_off:ch_off + C] = t_ndhwc.reshape(-1, C).to(dtype)
    return ops.Act(buf.reshape(-1), B, T, H, W, C, cs, ch_off)


def _read_act(act):
    return act.buf.view(act.B, act.T, act.H, act.W, act.cstride)[..., act.ch_off:act.ch_off + act.C]


CONV_CASES = [
    # B, T, H, W, C, N, kt, khw, name
    (1, 1, 16, 16, 256, 192, 1, 3, "slow1-like exact tiles"),
    (2, 8, 12, 21, 256, 32, 3, 3, "fast1-like ragged 12x21"),
    (1, 6, 24, 42, 32, 32, 3, 3, "fast2-like Cin=32"),
    (2, 6, 12, 21, 32, 64, 6, 1, "lateral k=(6,1,1)"),
    (1, 2, 48, 84, 256, 224, 2, 3, "slow3-like N=224 kt=2"),
    (3, 1, 14, 14, 256, 256, 1, 3, "mask-head 14x14 N=256"),
    (1, 3, 7, 5, 64, 96, 2, 3, "tiny odd"),
    (1, 16, 12, 21, 256, 32, 6, 3, "fast1 (2,16): kt=6, 11 output frames in 2 groups"),
    (1, 32, 8, 9, 64, 32, 11, 3, "fast1 (4,32): kt=11, 22 output frames, 4 tap groups"),
    (2, 12, 12, 21, 32, 32, 12, 3, "fast3 (4,32): kt=12 -> 1 frame, Cin=32"),
    (2, 4, 40, 23, 32, 32, 4, 3, "fast3 (1,8): kt=4 -> 1 frame, ragged"),
]


@pytest.mark.parametrize("umma", [True, False], ids=["umma", "simt"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[-1] for c in CONV_CASES])
def test_conv_fprop(case, umma):
    ops = _ops()
    B, T, H, W, C, N, kt, khw, _ = case
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, T, H, W, C, generator=g).to(DEV)
    w = (torch.randn(N, C, kt, khw, khw, generator=g) / math.sqrt(C * kt * khw * khw)).to(DEV)
    dtype = torch.bfloat16 if umma else torch.float32
    if umma:
        x = x.bfloat16().float()
        w = w.bfloat16().float()
    pad = 1 if khw == 3 else 0
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=(0, pad, pad)).permute(0, 2, 3, 4, 1)
    xa = _mk_act(ops, x, dtype)
    cp = (32 if C <= 32 else (C + 63) // 64 * 64) if umma else C
    wp = ops.pack_weights(w, 0, ops.BF16 if umma else ops.F32, cp)
    To = T - kt + 1
    y = ops.Act.empty(B, To, H, W, N, torch.float32, DEV)
    stats = torch.zeros(2 * N, device=DEV) if umma else None
    ops.conv(xa, wp, cp, N, (kt, khw, khw), (0, pad, pad), To, y, umma=umma, stats=stats)
    torch.cuda.synchronize()
    out = _read_act(y)
    assert _err(out, ref) < (2e-5 if umma else 1e-5)
    if umma:
        n = ref.numel() // N
        s_ref = ref.reshape(-1, N).double().sum(0)
        q_ref = (ref.reshape(-1, N).double() ** 2).sum(0)
        assert ((stats[:N].double() - s_ref).abs() / (n * q_ref).sqrt()).max().item() < 1e-5
        assert _err(stats[N:], q_ref.float()) < 1e-4


@pytest.mark.parametrize("umma", [True, False], ids=["umma", "simt"])
def test_conv_epilogue_options_and_channel_slices(umma):
    """scale/shift/ReLU epilogue, bf16 output into a channel slice of a 256-wide buffer, input read from a slice,
    accumulate into f32."""
    ops = _ops()
    B, T, H, W, C, N, kt = 2, 3, 12, 21, 64, 64, 2
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, T, H, W, C, generator=g).to(DEV)
    w = (torch.randn(N, C, kt, 3, 3, generator=g) / math.sqrt(C * kt * 9)).to(DEV)
    scale = (torch.rand(N, generator=g) + 0.5).to(DEV)
    shift = torch.randn(N, generator=g).to(DEV)
    dtype = torch.bfloat16 if umma else torch.float32
    if umma:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=(0, 1, 1)).permute(0, 2, 3, 4, 1)
    ref_act = torch.relu(ref * scale + shift)
    xa = _mk_act(ops, x, dtype, cstride=256, ch_off=128)
    cp = 64
    wp = ops.pack_weights(w, 0, ops.BF16 if umma else ops.F32, cp)
    To = T - kt + 1
    ybuf = ops.Act.empty(B, To, H, W, 256, dtype, DEV)
    ybuf.buf.zero_()
    ops.conv(xa, wp, cp, N, (kt, 3, 3), (0, 1, 1), To, ybuf.slice(192, 64), umma=umma, relu=True, scale=scale, shift=shift)
    out = ybuf.buf.view(-1, 256)
    assert _err(out[:, 192:], ref_act.reshape(-1, N)) < (1e-2 if umma else 1e-5)
    assert out[:, :192].abs().max().item() == 0
    # accumulate
    acc = ops.Act.empty(B, To, H, W, N, torch.float32, DEV)
    acc.buf.fill_(1.5)
    ops.conv(xa, wp, cp, N, (kt, 3, 3), (0, 1, 1), To, acc, umma=umma, accumulate=True)
    assert _err(_read_act(acc), ref + 1.5) < 2e-5


@pytest.mark.parametrize("umma", [True, False], ids=["umma", "simt"])
@pytest.mark.parametrize("cin,cout,kt,khw,T", [(256, 192, 2, 3, 4), (32, 64, 3, 1, 5), (256, 224, 1, 3, 1), (32, 32, 3, 3, 6),
                                                (32, 32, 4, 3, 4), (32, 32, 11, 3, 22), (32, 32, 6, 3, 11),
                                                (32, 64, 6, 1, 6), (32, 64, 20, 1, 22)])
def test_conv_dgrad(cin, cout, kt, khw, T, umma):
    """Data gradient = the same kernel on dy with flipped/transposed packed weights and 'full' temporal padding."""
    ops = _ops()
    B, H, W = 2, 12, 21
    pad = 1 if khw == 3 else 0
    g = torch.Generator().manual_seed(3)
    w = (torch.randn(cout, cin, kt, khw, khw, generator=g) / math.sqrt(cin * kt * khw * khw)).to(DEV)
    To = T - kt + 1
    dy = torch.randn(B, To, H, W, cout, generator=g).to(DEV)
    dtype = torch.bfloat16 if umma else torch.float32
    if umma:
        w, dy = w.bfloat16().float(), dy.bfloat16().float()
    x = torch.zeros(B, cin, T, H, W, device=DEV, requires_grad=True)
    yref = F.conv3d(x, w, padding=(0, pad, pad))
    (dx_ref,) = torch.autograd.grad(yref, x, dy.permute(0, 4, 1, 2, 3))
    dx_ref = dx_ref.permute(0, 2, 3, 4, 1)
    cp = (32 if cout <= 32 else (cout + 63) // 64 * 64) if umma else cout
    wd = ops.pack_weights(w, 1, ops.BF16 if umma else ops.F32, cp)
    dya = _mk_act(ops, dy, dtype)
    dx = ops.Act.empty(B, T, H, W, cin, torch.float32, DEV)
    ops.conv(dya, wd, cp, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, dx, umma=umma)
    assert _err(_read_act(dx), dx_ref) < 2e-5
    # accumulate into an existing f32 gradient (how the lateral path merges into the fast pathway's gradient)
    dx.buf.fill_(0.25)
    ops.conv(dya, wd, cp, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, dx, umma=umma, accumulate=True)
    assert _err(_read_act(dx), dx_ref + 0.25) < 2e-5
    if umma:
        # addend: the total of two gradient paths stored once, in bf16 or f32, from an f32 or bf16 partial sum that is left
        # untouched (sfvos_conv_params.addend: the lateral dgrad on top of the fast convolution's dgrad)
        part = torch.randn(B, T, H, W, cin, generator=g).to(DEV)
        for part_dtype in (torch.float32, torch.bfloat16):
            pa = _mk_act(ops, part, part_dtype)
            before = pa.buf.clone()
            for out_dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 4e-3)):     # bf16: one rounding of the total (2^-9)
                tot = ops.Act.empty(B, T, H, W, cin, out_dtype, DEV)
                ops.conv(dya, wd, cp, cin, (kt, khw, khw), (kt - 1, khw - 1 - pad, khw - 1 - pad), T, tot, umma=True, addend=pa)
                assert _err(_read_act(tot), dx_ref + _read_act(pa).float()) < tol
            assert torch.equal(pa.buf, before)


@pytest.mark.parametrize("umma", [True, False], ids=["umma", "simt"])
@pytest.mark.parametrize("cin,cout,kt,khw,T,H,W", [(256, 192, 1, 3, 1, 16, 16), (256, 32, 3, 3, 5, 12, 21), (32, 64, 4, 1, 6, 12, 21),
                                                   (32, 32, 3, 3, 5, 24, 42), (256, 224, 2, 3, 3, 12, 21), (256, 256, 1, 3, 1, 14, 14),
                                                   (32, 32, 4, 3, 4, 40, 23), (32, 32, 6, 3, 11, 12, 21), (32, 32, 12, 3, 12, 9, 10),
                                                   (256, 32, 6, 3, 16, 9, 21), (64, 32, 2, 3, 3, 7, 5)])
def test_wgrad(cin, cout, kt, khw, T, H, W, umma):
    ops = _ops()
    B = 2
    pad = 1 if khw == 3 else 0
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, H, W, cin, generator=g).to(DEV)
    To = T - kt + 1
    dy = torch.randn(B, To, H, W, cout, generator=g).to(DEV)
    dtype = torch.bfloat16 if umma else torch.float32
    if umma:
        x, dy = x.bfloat16().float(), dy.bfloat16().float()
    w = torch.zeros(cout, cin, kt, khw, khw, device=DEV, requires_grad=True)
    yref = F.conv3d(x.permute(0, 4, 1, 2, 3), w, padding=(0, pad, pad))
    (dw_ref,) = torch.autograd.grad(yref, w, dy.permute(0, 4, 1, 2, 3))
    xa, dya = _mk_act(ops, x, dtype), _mk_act(ops, dy, dtype)
    taps = kt * khw * khw
    dwp = torch.zeros(taps * cin * cout, device=DEV)
    ops.wgrad(xa, dya, (kt, khw, khw), (0, pad, pad), dwp, umma=umma)
    gw = torch.zeros(dw_ref.shape, device=DEV)
    ops.unpack_wgrad(dwp, gw, 0)
    assert _err(gw, dw_ref) < 5e-5


def test_bn_pieces_match_torch_batchnorm():
    ops = _ops()
    B, T, H, W, C = 2, 3, 12, 21, 64
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, T, H, W, C, generator=g) * 2 + 0.7).to(DEV)
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    xa = _mk_act(ops, x, torch.float32)
    stats = torch.zeros(2 * C, device=DEV, dtype=torch.float64)
    ops.channel_stats(xa, stats)
    bn4 = torch.empty(4 * C, device=DEV)
    ops.bn_finalize(stats, xa.npix, bias, gamma, beta, rm, rv, nbt, 0.1, 1e-5, bn4)
    y = ops.Act.empty(B, T, H, W, C, torch.float32, DEV)
    ops.affine_act(xa, y, bn4[:C], bn4[C:2 * C], True)
    # torch reference: BN of (x + bias) in train mode
    xr = (x + bias).permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    yr = torch.relu(F.batch_norm(xr, rm_r, rv_r, gr, br, True, 0.1, 1e-5))
    assert _err(_read_act(y), yr.permute(0, 2, 3, 4, 1)) < 1e-5
    assert _err(rm, rm_r) < 1e-5 and _err(rv, rv_r) < 1e-5 and int(nbt) == 1
    # backward
    dy = torch.randn(B, T, H, W, C, generator=g).to(DEV)
    yr.backward(dy.permute(0, 4, 1, 2, 3))
    dya = _mk_act(ops, dy, torch.float32)
    dx = ops.Act.empty(B, T, H, W, C, torch.float32, DEV)
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ops.bn_bwd(dya, xa, bn4, gamma, True, dx, dgamma, dbeta)
    assert _err(_read_act(dx), xr.grad.permute(0, 2, 3, 4, 1)) < 2e-5
    assert _err(dgamma, gr.grad) < 2e-5 and _err(dbeta, br.grad) < 2e-5
    # eval fold
    fold = torch.empty(2 * C, device=DEV)
    ops.bn_fold_eval(bias, gamma, beta, rm, rv, 1e-5, fold)
    ops.affine_act(xa, y, fold[:C], fold[C:], False)
    ye = F.batch_norm((x + bias).permute(0, 4, 1, 2, 3), rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert _err(_read_act(y), ye.permute(0, 2, 3, 4, 1)) < 1e-5


@pytest.mark.parametrize("C,H,W", [(256, 12, 22), (64, 8, 16), (128, 40, 23), (256, 12, 21), (72, 5, 8)])
def test_layout_bf16_source(C, H, W):
    """bf16 [frames,C,H,W] -> bf16 channels-last (the product path's layout pass): a pure transposition, bit-exact.  H*W % 8 == 0
    and C % 64 == 0 take the byte-permute kernel (full, partial and single tiles); the others the general kernel."""
    ops = _ops()
    F_ = 3
    x = torch.randn(F_, C, H, W, device=DEV).bfloat16()
    act = ops.Act.empty(1, F_, H, W, C, torch.bfloat16, DEV)
    ops.nchw_to_nhwc(x, act)
    assert torch.equal(_read_act(act)[0], x.permute(0, 2, 3, 1))
    # into a channel slice of a wider channels-last buffer (cstride > C), frames at an offset
    cs = C + 64
    buf = torch.zeros((F_ + 1) * H * W * cs, dtype=torch.bfloat16, device=DEV)
    act2 = ops.Act(buf, 1, F_ + 1, H, W, C, cs, 32)
    ops.nchw_to_nhwc(x, act2, frame_off=1)
    got = buf.view(F_ + 1, H, W, cs)
    assert torch.equal(got[1:, :, :, 32:32 + C], x.permute(0, 2, 3, 1))
    assert int(got[0].abs().sum()) == 0 and int(got[1:, :, :, :32].abs().sum()) == 0 and int(got[1:, :, :, 32 + C:].abs().sum()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_roundtrip(dtype):
    ops = _ops()
    F_, C, H, W = 3, 256, 12, 21
    x = torch.randn(F_, C, H, W, device=DEV)
    act = ops.Act.empty(1, F_, H, W, C, dtype, DEV)
    ops.nchw_to_nhwc(x, act)
    ref = x.permute(0, 2, 3, 1).to(dtype)
    assert torch.equal(_read_act(act)[0], ref)
    back = torch.empty(F_, C, H, W, device=DEV)
    ops.nhwc_to_nchw(act, back)
    assert torch.equal(back, ref.float().permute(0, 3, 1, 2))


def test_relu_bwd_and_bias_grad():
    ops = _ops()
    npix, C = 1000, 256
    y = torch.relu(torch.randn(1, 1, npix, 1, C, device=DEV)).bfloat16()
    dy = torch.randn(1, 1, npix, 1, C, device=DEV)
    ya, dya = _mk_act(ops, y, torch.bfloat16), _mk_act(ops, dy, torch.float32)
    dx = ops.Act.empty(1, 1, npix, 1, C, torch.bfloat16, DEV)
    db = torch.zeros(C, device=DEV)
    ops.relu_bwd(dya, ya, dx, db)
    ref = dy * (y.float() > 0)
    assert _err(_read_act(dx), ref) < 5e-3
    assert _err(db, ref.reshape(-1, C).sum(0)) < 1e-5

@pytest.mark.parametrize("kt,T", [(6, 6), (4, 4), (3, 5), (12, 12)])
def test_lateral_wgrad_with_swapped_operands(kt, T):
    """The lateral connection's weight gradient (32 -> 64 channels, k_t x 1 x 1) is computed with the operands swapped so that it
    runs on the stacked 32-output-channel kernel: wgrad("x" := dy, "dy" := x, pad_t = k_t - 1) = dw'[k_t-1-ta][n][c]
    (slowfast._layer_backward / _GradBank.finish).  Against torch's conv3d weight gradient and the direct launch."""
    ops = _ops()
    B, H, W, cin, cout = 2, 12, 21, 32, 64
    To = T - kt + 1
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, H, W, cin, generator=g).to(DEV).bfloat16().float()
    dy = torch.randn(B, To, H, W, cout, generator=g).to(DEV).bfloat16().float()
    w = torch.zeros(cout, cin, kt, 1, 1, device=DEV, requires_grad=True)
    (dw_ref,) = torch.autograd.grad(F.conv3d(x.permute(0, 4, 1, 2, 3), w), w, dy.permute(0, 4, 1, 2, 3))
    xa, dya = _mk_act(ops, x, torch.bfloat16), _mk_act(ops, dy, torch.bfloat16)
    direct = torch.zeros(kt * cin * cout, device=DEV)
    ops.wgrad(xa, dya, (kt, 1, 1), (0, 0, 0), direct, umma=True)
    g_direct = torch.zeros(dw_ref.shape, device=DEV)
    ops.unpack_wgrad(direct, g_direct, 0)
    swapped = torch.zeros(kt * cin * cout, device=DEV)
    ops.wgrad(dya, xa, (kt, 1, 1), (kt - 1, 0, 0), swapped, umma=True)
    g_swapped = swapped.view(kt, cout, cin).flip(0).permute(1, 2, 0).reshape(dw_ref.shape)
    assert _err(g_direct, dw_ref) < 5e-5
    assert _err(g_swapped, dw_ref) < 5e-5
