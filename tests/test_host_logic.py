"""CPU tests of the host-side logic added in round 2 (no kernel launches): clip-layout detection for the overlapping-window
input path, the per-parameter packed-weight cache, the frozen-BatchNorm conversion of the offline-built backbone (parameter
count = the reference's report), the gradient arena's slicing, and the bench's algorithmic-FLOP accounting."""
import gc

import pytest
import torch


def test_frame_layout_and_window_sequence_detection():
    from sfvos_b200.slowfast import _alias_offset, _frame_layout, _window_sequence
    for fmt in (torch.contiguous_format, torch.channels_last):
        seq = torch.randn(11, 8, 3, 5).contiguous(memory_format=fmt)
        assert _frame_layout(seq) == ("nchw" if fmt == torch.contiguous_format else "nhwc")
        assert _frame_layout(seq[2:9]) == _frame_layout(seq)            # a frame range keeps the layout
        assert _frame_layout(seq[:, :4]) is None                        # a channel slice does not
        clips = [seq[i:i + 8] for i in range(4)]                        # the reference's clips: consecutive windows (model.py:318-337)
        union = _window_sequence(clips)
        assert union is not None and union.shape[0] == 11 and union.data_ptr() == seq.data_ptr() and torch.equal(union, seq)
        assert _alias_offset([c[4:5] for c in clips], clips) == 4       # slow window = frame range of the fast one, either layout
        sub = seq[1:]
        union = _window_sequence([sub[i:i + 8] for i in range(3)])      # windows that do not start at the tensor's first frame
        assert torch.equal(union, seq[1:11])
        assert _window_sequence([seq[0:8], seq[2:10]]) is None          # stride of two frames: not consecutive
        assert _window_sequence([seq[i:i + 8].clone() for i in range(4)]) is None      # copies: different storages
        assert _window_sequence([seq[0:8]]) is None                     # a single clip has nothing to share
    assert _window_sequence([torch.randn(8, 8, 3, 5).half()[i:i + 4] for i in range(3)]) is None    # unsupported dtype


def test_packed_weight_cache_follows_the_parameter_object():
    from sfvos_b200 import roi_heads as rh
    w = torch.nn.Parameter(torch.randn(4, 4))
    made = []
    make = lambda: made.append(1) or len(made)
    assert rh._cached(w, "k", make) == 1 and rh._cached(w, "k", make) == 1        # cached while the parameter is unchanged
    with torch.no_grad():
        w.add_(1.0)                                                                # an optimizer step bumps the version
    assert rh._cached(w, "k", make) == 2
    assert rh._cached(w, "other", make) == 3
    assert rh._cached(torch.randn(2), "k", make) == 4 and rh._cached(torch.randn(2), "k", make) == 5   # plain tensors: never cached
    n = len(rh._PACKED)
    del w
    gc.collect()
    assert len(rh._PACKED) == n - 1                                                # the entry dies with its parameter


def test_offline_backbone_is_frozen_batchnorm_and_counts_match_the_report():
    """code/helpers/model.py:13 loads the pretrained Mask R-CNN, whose ResNet has FrozenBatchNorm2d; built offline the model must
    have the same structure: total parameters = final_report/chapters/Experiments.tex:20 (slow-fast 1-1: 45,421,851)."""
    import warnings
    from torchvision.ops.misc import FrozenBatchNorm2d
    from sfvos_b200.model import SegmentationModel
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = SegmentationModel(torch.device("cpu"), 1, 1, maskrcnn_weights=None, pretrained=False)
    body = m.maskrcnn_model.backbone.body
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in body.modules())
    assert sum(isinstance(x, FrozenBatchNorm2d) for x in body.modules()) == 53
    assert sum(p.numel() for p in m.parameters()) == 45421851
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 1499456 + 16529164      # slow_fast (1,1) + roi_heads
    assert type(body).__name__ == "ResNetBody" and type(m.maskrcnn_model.backbone.fpn).__name__ == "FeaturePyramidNetwork"
    keys = list(m.state_dict().keys())
    assert keys[0] == "maskrcnn_model.backbone.body.conv1.weight" and "maskrcnn_model.rpn.head.conv.0.0.weight" in keys


def test_gradient_arena_slices_and_ranges():
    from sfvos_b200 import dp
    a, b, c = (torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2, 2)))
    frozen = torch.nn.Parameter(torch.zeros(4), requires_grad=False)
    arena = dp.GradArena([("head", [a, frozen, b]), ("trunk", [c])])
    assert arena.flat.numel() == 16 + 8 + 8 and arena.view(frozen) is None      # 16-byte aligned slices, frozen parameters skipped
    arena.view(a).fill_(1.0); arena.view(b).fill_(2.0); arena.view(c).fill_(3.0)
    assert arena.range("head").tolist() == [1.0] * 15 + [0.0] + [2.0] * 7 + [0.0] and arena.range("trunk").tolist() == [3.0] * 8
    assert arena.view(a, (5, 3)).shape == (5, 3) and arena.view(a) is not arena.view(a)          # a NEW tensor object per call
    assert not arena.adopted()
    for p in (a, b, c):
        p.grad = arena.view(p)
    assert arena.adopted()
    arena.zero()
    assert float(a.grad.abs().sum()) == 0.0


def test_dgrad_flops_exclude_zero_padded_temporal_taps():
    """ops.conv counts the algorithmic FLOPs of a launch: a data-gradient launch does the forward layer's work -
    one product per forward OUTPUT frame and tap - not T_in x k_t frame-taps."""
    from sfvos_b200 import ops
    recorded = []

    class _X:       # stands in for an Act: only the attributes the accounting reads
        def __init__(self, B, T, H, W, C):
            self.B, self.T, self.H, self.W, self.C = B, T, H, W, C
    flops = lambda x, To, N, k: 2.0 * x.B * min(To, x.T) * x.H * x.W * N * x.C * k[0] * k[1] * k[2]
    fwd = flops(_X(8, 8, 192, 336, 256), 6, 32, (3, 3, 3))             # fast_conv1 fprop: 8 frames in, 6 out
    bwd = flops(_X(8, 4, 192, 336, 32), 6, 32, (3, 3, 3))              # fast_conv2 dgrad: dy has 4 frames, dx 6
    assert abs(fwd / 1e9 - 1369.8) < 1.0 and abs(bwd / 1e9 - 114.2) < 0.2     # = the layer's forward GFLOP
    import inspect
    assert "min(To, x.T)" in inspect.getsource(ops.conv)


@pytest.mark.parametrize("kt,T", [(6, 6), (4, 4), (3, 5)])
def test_swapped_lateral_weight_gradient_identity(kt, T):
    """Host-side algebra of slowfast._layer_backward / _GradBank.finish for the lateral connections (code/helpers/model.py:83-90,
    invoked :128-131,140-143): the weight gradient of the valid temporal convolution y = conv3d(x, W), W [Cout,Cin,kt,1,1], equals
    the weight-gradient problem with the operands swapped ("x" := dy, "dy" := x, temporal padding kt - 1), un-swapped with
    flip(taps) + transpose(channels).  Here the swapped problem is evaluated with plain torch on the CPU, in the layout the kernel
    accumulates: dw'[ta'][c' = Cout][n' = Cin] = sum_{b,t',px} dy[b, t' + ta' - (kt-1), px, c'] * x[b, t', px, n']."""
    import torch.nn.functional as F
    B, H, W, cin, cout = 2, 3, 4, 8, 16
    To = T - kt + 1
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, cin, T, H, W, generator=g, dtype=torch.float64)
    dy = torch.randn(B, cout, To, H, W, generator=g, dtype=torch.float64)
    w = torch.zeros(cout, cin, kt, 1, 1, dtype=torch.float64, requires_grad=True)
    (dw_ref,) = torch.autograd.grad(F.conv3d(x, w), w, dy)
    dwp = torch.zeros(kt, cout, cin, dtype=torch.float64)
    for tap in range(kt):
        for t in range(T):                       # t' runs over the frames of the swapped problem's "dy" (= x)
            src = t + tap - (kt - 1)             # frame of the swapped problem's "x" (= dy); out of range = zero padding
            if 0 <= src < To:
                dwp[tap] += torch.einsum("bchw,bnhw->cn", dy[:, :, src], x[:, :, t])
    unswapped = dwp.flip(0).permute(1, 2, 0).reshape(dw_ref.shape)      # exactly _GradBank.finish's expression
    assert torch.allclose(unswapped, dw_ref, rtol=1e-12, atol=1e-12)
