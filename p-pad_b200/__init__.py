"""ppad_b200 -- a B200-native batched Puzzle & Dragons environment.

Keeps the Python env API of nuwapi/P-PAD (``PAD``: reset / step / calculate_reward / backtrack,
actions up/down/left/right/pass, ``(board, finger)`` state, combo-based float64 reward) on top of
hand-written sm_100a kernels reached through the C ABI of ``include/ppad_b200.h``.

    import ppad_b200
    env = ppad_b200.PAD()                       # drop-in for ppad.PAD (ppad/pad/game.py:28)
    venv = ppad_b200.BatchedPAD(1 << 20)        # a million boards on one GPU

There is no CPU path: compute calls raise if the CUDA library or a CUDA device is missing.
"""
from . import _cabi, layout  # noqa: F401
from .config import make_config  # noqa: F401

__all__ = ["PAD", "BatchedPAD", "make_config", "layout", "discount", "board_finger_to_state", "install_as_ppad"]

_LAZY = {
    "BatchedPAD": ("batched", "BatchedPAD"),
    "PAD": ("pad.game", "PAD"),
    "discount": ("agent.utils", "discount"),
    "board_finger_to_state": ("agent.encode", "board_finger_to_state"),
    "install_as_ppad": ("compat", "install_as_ppad"),
    "smart_data": ("data.smart_data", "smart_data"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        value = getattr(importlib.import_module("." + mod, __name__), attr)
        globals()[name] = value
        return value
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
