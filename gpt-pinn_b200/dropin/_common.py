"""Shared plumbing of the drop-in modules: locate the gptpinn package that sits next to this folder."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402

# one rank per GPU under torchrun: the reference scripts say torch.device("cuda") (= the current device), so the
# current device is set to LOCAL_RANK here, at import time of the first drop-in module, before the script allocates
if "LOCAL_RANK" in os.environ and torch.cuda.is_available():
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())

import gptpinn  # noqa: E402,F401
from gptpinn import offline, precomp, sweep, pinn  # noqa: E402,F401
