"""board_finger_to_state -- the one-hot encoding of ``Agent01`` (ppad/agent/agent01.py:254-273) on the GPU."""
import numpy as np

_slabs = {}


def board_finger_to_state(boards, fingers, st_shape=None, as_numpy=True):
    """boards [B, R, C] ints, fingers [B, 2] -> [B, R, C, 7] float64 (numpy, like the reference) or, with
    as_numpy=False, the float32 CUDA tensor the kernel wrote.  Channels 0..5 = (board == j), channel 6 = finger."""
    import torch
    from ..batched import BatchedPAD
    is_tensor = torch.is_tensor(boards)
    boards = boards if is_tensor else np.asarray(boards)
    b, rows, cols = boards.shape
    if st_shape is not None and (tuple(st_shape)[:2] != (rows, cols) or tuple(st_shape)[2] != 7):
        raise ValueError("st_shape %r does not match boards of %dx%d with 7 channels" % (st_shape, rows, cols))
    key = (b, rows, cols)
    if key not in _slabs:
        if len(_slabs) > 8:
            _slabs.clear()
        _slabs[key] = BatchedPAD(b, rows, cols)
    venv = _slabs[key]
    if is_tensor:
        masked = torch.where((boards > 6) | (boards < -1), torch.full_like(boards, -1), boards)
    else:
        masked = np.where((boards > 6) | (boards < -1), -1, boards)   # types 7..9 have no channel: all-zero either way
    venv.set_state(masked, fingers)
    obs = venv.encode_obs('f32')
    return obs.cpu().numpy().astype(np.float64) if as_numpy else obs
