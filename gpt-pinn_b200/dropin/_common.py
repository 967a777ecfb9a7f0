"""Shared plumbing of the drop-in modules: locate the gptpinn package that sits next to this folder."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import gptpinn  # noqa: E402,F401
from gptpinn import offline, precomp, sweep, pinn  # noqa: E402,F401
