"""CPU tests of the script-level drop-in (north star: "train.py and prediction.py use it unchanged"), run against the REAL
reference tree where it exists (the build container; the GPU box has none -- there tests/test_gpu_callers.py replays the same
call sequences on an independent restatement of the dataset contract).

``python -m sfvos_b200.run_reference <script>`` runs an unchanged reference script with ``helpers.model`` -- and only that
module -- replaced by the libsfvos-backed classes.  Without a GPU the run must get all the way through the reference's own
imports, constants, dataset, checkpoint loading and the torchvision transform / backbone, and then FAIL LOUDLY at the first
libsfvos launch (there is no CPU fallback); with a GPU the same command completes."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from davis_like import write_davis_tree

REF_CODE = "/root/reference/code"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_CODE), reason="needs the reference tree (build container only)")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests", "stubs")])     # stubs: matplotlib is not in this image
    env["MPLBACKEND"] = "Agg"
    return env


def _run(workdir, script, *args, timeout=600):
    return subprocess.run([sys.executable, "-m", "sfvos_b200.run_reference", script, *args], cwd=workdir, env=_env(),
                          capture_output=True, text=True, timeout=timeout)


def test_only_helpers_model_is_replaced(tmp_path):
    probe = tmp_path / "probe.py"
    probe.write_text(
        "import sys\nsys.path.insert(0, %r)\n"
        "import helpers.model, helpers.constants, helpers.dataset, helpers.utils\n"
        "from helpers.model import SegmentationModel, SlowFastLayers\n"
        "import sfvos_b200\n"
        "assert SegmentationModel is sfvos_b200.model.SegmentationModel and SlowFastLayers is sfvos_b200.SlowFastLayers\n"
        "for m in (helpers, helpers.constants, helpers.dataset, helpers.utils):\n"
        "    assert m.__file__.startswith(%r), m.__file__\n"
        "print('PROBE_OK', helpers.constants.model_name)\n" % (REF_CODE, REF_CODE))
    r = _run(str(tmp_path), str(probe))
    assert r.returncode == 0 and "PROBE_OK model_maskrcnn_slowfast_sp_1fp_1" in r.stdout, r.stderr[-2000:]


def test_compat_package_extends_to_the_reference_helpers(tmp_path):
    """Importing from outside the reference tree: compat/helpers shadows only helpers.model."""
    code = ("import sys\nsys.path[:0] = [%r, %r, %r, %r]\n"
            "import helpers.model, helpers.constants, helpers.dataset\n"
            "assert helpers.model.__file__.startswith(%r) and helpers.constants.__file__.startswith(%r)\nprint('OK')\n"
            % (os.path.join(ROOT, "compat"), ROOT, os.path.join(ROOT, "tests", "stubs"), REF_CODE, os.path.join(ROOT, "compat"), REF_CODE))
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_prediction_py_runs_unchanged_up_to_the_first_kernel(tmp_path):
    """code/prediction.py: constants -> SegmentationModel(...) -> .to(device) -> load_state_dict(best_model_path) ->
    evaluation.evaluate -> DAVISDataset -> model.eval() -> model(imgs, targets)."""
    work = str(tmp_path)
    write_davis_tree(os.path.join(work, "data", "DAVIS"), sequences=(("seq_a", 3),), hw=(96, 128))
    # a checkpoint in the reference's format at the path prediction.py loads (helpers/constants.py: models/<model_name>_best.pth)
    from sfvos_b200.model import SegmentationModel
    torch.manual_seed(63)
    m = SegmentationModel(device=torch.device("cpu"), slow_pathway_size=1, fast_pathway_size=1, maskrcnn_weights=None, pretrained=False)
    os.makedirs(os.path.join(work, "models"), exist_ok=True)
    torch.save(m.state_dict(), os.path.join(work, "models", "model_maskrcnn_slowfast_sp_1fp_1_best.pth"))
    r = _run(work, os.path.join(REF_CODE, "prediction.py"))
    out = r.stdout + r.stderr
    assert "Environment is local" in out                               # the reference's own helpers/constants.py ran
    assert "Evaluating with Sequence" in out                           # evaluate() reached the model call on the dataset
    if torch.cuda.is_available():
        assert r.returncode == 0 and "Mean_IoU" in out, out[-3000:]
    else:
        assert r.returncode != 0 and "no CPU fallback" in out, out[-3000:]
