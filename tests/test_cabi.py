"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/sfvos.h declares, the
ctypes structures match the C structs byte for byte, and the product path refuses to run without its CUDA
extension / device instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "sfvos.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sfvos_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sfvos_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sfvos.h but not exported by libsfvos.so"
    assert set(_lib.EXPORTED) == set(declared), set(_lib.EXPORTED) ^ set(declared)
    assert lib.sfvos_version() == 100


def test_ctypes_structs_match_c_layout(tmp_path):
    from sfvos_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sfvos.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(sfvos_conv_params), sizeof(sfvos_wgrad_params), sizeof(sfvos_roi_params),'
                    'offsetof(sfvos_conv_params, OH), offsetof(sfvos_wgrad_params, dw), offsetof(sfvos_roi_params, out),'
                    'sizeof(sfvos_bn_running_params), offsetof(sfvos_bn_running_params, momentum),'
                    'offsetof(sfvos_conv_params, addend), offsetof(sfvos_conv_params, addend_cstride));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.ConvParams), ctypes.sizeof(_lib.WgradParams), ctypes.sizeof(_lib.RoiParams),
            _lib.ConvParams.OH.offset, _lib.WgradParams.dw.offset, _lib.RoiParams.out.offset,
            ctypes.sizeof(_lib.BnRunningParams), _lib.BnRunningParams.momentum.offset,
            _lib.ConvParams.addend.offset, _lib.ConvParams.addend_cstride.offset]
    assert got == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_no_cpu_fallback_without_gpu():
    from sfvos_b200 import SlowFastLayers, _lib, ops
    lib = _lib.load()
    assert lib.sfvos_device_check() != 0                      # SFVOS_ERR_UNSUPPORTED / CUDA error, with a message
    assert len(lib.sfvos_last_error()) > 0
    with pytest.raises(RuntimeError):
        ops.device_check()
    m = SlowFastLayers(256, torch.device("cpu"), 1, 8)
    x = torch.randn(1, 256, 8, 4, 6)
    with pytest.raises(RuntimeError):
        m(x[:, :, 4:5], x)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "applying-slowfast-networks-to-video-object-segmentation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "/root/reference" not in text, f


def test_module_interface_matches_reference_contract():
    """ctor signature, attributes, state_dict keys / shapes / order -- against the oracle's restatement of
    the reference's registration order, which tests/test_oracle.py pins to the reference itself."""
    import inspect
    from oracle import slowfast_oracle as so
    from sfvos_b200 import SlowFastLayers
    assert list(inspect.signature(SlowFastLayers.__init__).parameters) == ["self", "input_size", "device", "slow_pathway_size", "fast_pathway_size"]
    for sp, fp in [(1, 8), (4, 32), (2, 16), (3, 7), (1, 1), (8, 8)]:
        torch.manual_seed(63)
        m = SlowFastLayers(256, torch.device("cpu"), sp, fp)
        ref = so.init_state_dict(sp, fp, seed=63)
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in ref:
            assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k
        assert [n for n, _ in m.named_parameters()] == [k for k in ref if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        for attr in ("device", "slow_pathway_size", "fast_pathway_size", "relu", "fuse", "forward", "temporally_enhance_features",
                     "_calc_kernel_sizes", "_calc_fuse_kernel_size"):
            assert hasattr(m, attr)


def test_slow_window_alias_detection():
    from sfvos_b200.slowfast import _alias_offset
    fast = [torch.randn(8, 4, 3, 5) for _ in range(2)]
    assert _alias_offset([f[4:5] for f in fast], fast) == 4
    assert _alias_offset([f[3:6] for f in fast], fast) == 3
    assert _alias_offset([fast[0][4:5], fast[1][3:4]], fast) is None          # different offsets per clip
    assert _alias_offset([f[4:5].clone() for f in fast], fast) is None        # not a view
    assert _alias_offset([f[:, :2][4:5] for f in fast], fast) is None         # channel slice: not contiguous


def test_workload_flop_formulas_match_survey():
    from sfvos_b200 import workload as wl
    assert abs(wl.conv_flops(1, 8, fwd_only=True) / 1e9 - 497.7) < 0.1
    assert abs(wl.conv_flops(1, 8) / 1e9 - 1189.0) < 0.1
    assert abs(wl.conv_flops(4, 32) / 1e9 - 9260.4) < 0.5
    assert abs(wl.conv_flops(2, 16) / 1e9 - 3196.4) < 0.5
    assert abs(wl.MASK_HEAD_FLOPS_PER_ROI / 1e9 - 1.028) < 1e-3
    assert abs(wl.BOX_HEAD_FLOPS_PER_ROI / 1e9 - 0.0278) < 1e-4          # box head 0.0278 GFLOP/ROI fwd
    lab, tgt = wl.synthetic_box_targets(2, 512, 128)
    assert [int(l.sum()) for l in lab] == [128, 128] and tgt[0].shape == (512, 4) and not torch.equal(tgt[0], tgt[1])


def test_bench_configs_map_to_baseline():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.CONFIGS["c2"][:3] == (1, 8, 8) and bench.CONFIGS["c3"][:2] == (4, 32) and bench.CONFIGS["c5"][:2] == (2, 16)
    assert bench.METRIC == "clip_frames_per_sec_fwd_bwd" and bench.UNIT == "clip-frames/s"
