"""discount -- ``ppad.discount`` (ppad/agent/utils.py:22-51) on the GPU (ppad_discount)."""
import numpy as np

_venv = None


def _slab():
    global _venv
    if _venv is None:
        from ..batched import BatchedPAD
        _venv = BatchedPAD(1)
    return _venv


def discount(rewards, gamma, norm=False, log10=True):
    """Discounted returns of one trajectory (list / array of raw rewards) -> float64 numpy array."""
    r = np.array(rewards, dtype=np.float64)
    if r.size == 0:
        return r
    return _slab().discount(r[None], gamma, norm=norm, log10=log10)[0].cpu().numpy()
