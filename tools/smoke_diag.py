"""Per-parameter gradient errors for the smoke() configuration, under the kernel-selection toggles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from grad_err_report import run
if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    run(1, 8, prec, OrderedDict([("0", (24, 42)), ("pool", (6, 11))]), emulate=(len(sys.argv) > 2))
