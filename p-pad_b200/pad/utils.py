"""Free functions of the reference env (ppad/pad/utils.py:26-111) on top of the CUDA kernels."""
import numpy as np

from . import parameters

_slabs = {}


def _slab(rows, cols, orb_bits=3):
    key = (rows, cols, orb_bits)
    if key not in _slabs:
        from ..batched import BatchedPAD
        _slabs[key] = BatchedPAD(1, rows, cols, orb_bits=orb_bits)
    return _slabs[key]


def cancel(board):
    """Cancel all 3+ connected orbs of `board` IN PLACE and return the combos, as ppad.pad.utils.cancel
    (utils.py:26-70): a list of {'color', 'N', 'enhanced', 'shape'} in the reference's order."""
    rows, cols = board.shape
    venv = _slab(rows, cols, 4 if (np.asarray(board) > 6).any() else 3)   # types 7..9: the 4-bit kernels
    flags = venv.set_state(np.asarray(board)[None], np.zeros((1, 2), dtype=np.int64))
    if int(flags[0]):
        raise NotImplementedError('orb types are -1 (empty) and 0..9 (game.py:236-246)')
    n_combos, log = venv.cancel(log_cap=(rows * cols) // 3)
    n = int(n_combos[0])
    log = log[0, :n].cpu().numpy().view(np.uint16)
    board[...] = venv.get_state().boards[0].cpu().numpy()
    return [{'color': parameters.int2english[int(v) & 0xFF], 'N': np.float64(int(v) >> 8),
             'enhanced': None, 'shape': None} for v in log]


def illegal_actions(finger, dim_row=5, dim_col=6):
    """Set of off-board directions (utils.py:73-83), read from bits 4..7 of ppad_legal_mask."""
    venv = _slab(dim_row, dim_col)
    venv.set_state(np.zeros((1, dim_row, dim_col), dtype=np.int64), np.asarray(finger).reshape(1, 2))
    mask = int(venv.legal_mask()[0]) >> 4
    names = ['up', 'down', 'left', 'right']
    out = {names[a] for a in range(4) if (mask >> a) & 1}
    # the reference pairs the checks with `elif` (utils.py:75-82): on a one-row / one-column board it reports 'up' /
    # 'left' only
    if dim_row == 1:
        out.discard('down')
    if dim_col == 1:
        out.discard('right')
    return out


def detect_island(board, island, x, y, orb_type):
    """Mark, IN PLACE in `island`, the 4-connected island of cells of `orb_type` around (x, y) = (row, col), as
    ppad.pad.utils.detect_island (utils.py:86-111): cells already non-zero in `island` are not entered.  One launch of
    ppad_island (the closure of the same-colour edge masks the cancel kernel uses)."""
    board = np.asarray(board)
    rows, cols = board.shape
    venv = _slab(rows, cols, 4 if (board > 6).any() else 3)
    flags = venv.set_state(board[None], np.zeros((1, 2), dtype=np.int64))
    if int(flags[0]):
        raise NotImplementedError('orb types are -1 (empty) and 0..9 (game.py:236-246)')
    if not (0 <= x < rows and 0 <= y < cols):
        raise IndexError('index ({}, {}) is out of bounds for a {} x {} board'.format(x, y, rows, cols))
    cells = np.arange(rows * cols, dtype=np.uint64)
    pre = int((np.uint64(1) << cells)[np.asarray(island).reshape(-1) != 0].sum()) if np.any(island) else 0
    if pre >= 1 << 63:
        pre -= 1 << 64
    t = int(orb_type) if -1 <= int(orb_type) <= 9 else -2        # -2: no cell holds it
    mask = int(venv.islands(np.array([[x, y]]), np.array([t], np.int8), np.array([pre], np.int64))[0].item()) & (2 ** 64 - 1)
    new = np.array([(mask >> c) & 1 for c in range(rows * cols)], bool).reshape(rows, cols)
    island[new] = 1


def episode2gif(*args, **kwargs):
    raise NotImplementedError('rendering (utils.py:114-268) is out of scope of ppad_b200 (SURVEY.md section 2, #16)')
