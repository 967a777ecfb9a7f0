"""gptpinn -- host side of the B200-native GPT-PINN hot path (ctypes over libgptpinn_b200.so)."""
from . import _lib  # noqa: F401
from ._lib import GptPinnError  # noqa: F401
