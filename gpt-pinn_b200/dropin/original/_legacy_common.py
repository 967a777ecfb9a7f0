"""Plumbing of the legacy (`original/`) adapter: puts the drop-in folder and the gptpinn package on sys.path."""
import os
import sys

_DROPIN = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _DROPIN not in sys.path:
    sys.path.insert(1, _DROPIN)
from _common import precomp, sweep, torch  # noqa: E402,F401
