timeout 600 python -m pytest tests/test_gpu_slowfast.py -m gpu -q --no-header -p no:cacheprovider -k "emulated" 2>&1 | grep -E "assert|Error|passed|failed" | head
python - <<'PY'
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
from collections import OrderedDict
from grad_err_report import run
run(2, 16, "bf16", OrderedDict([("0", (8, 12)), ("pool", (4, 6))]), emulate=True)
PY
