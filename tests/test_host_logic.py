"""Host-side logic that needs no GPU: the documented byte layouts, config translation, import hygiene."""
import os
import re

import numpy as np
import pytest

import oracle as orc
import ppad_b200
from ppad_b200 import layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("rows,cols", [(5, 6), (6, 7)])
def test_state_layout_round_trip(rows, cols):
    rng = np.random.default_rng(rows)
    n = 500
    boards = rng.integers(-1, 7, size=(n, rows, cols))
    fingers = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1)
    prev = rng.integers(0, 5, n)
    moves = rng.integers(0, 65536, n)
    st = layout.pack_state(boards, fingers, prev, moves)
    assert st.dtype == np.uint32 and st.shape == (n, layout.state_words(rows, cols))
    b, f, p, m = layout.unpack_state(st, rows, cols)
    assert np.array_equal(b, boards) and np.array_equal(f, fingers) and np.array_equal(p, prev) and np.array_equal(m, moves)
    # the documented bit positions: plane j bit i = bit j of the code of cell i; 7 = empty
    i = int(rng.integers(n))
    code = np.where(boards[i] < 0, 7, boards[i]).reshape(-1)
    planes = [int(st[i, j]) for j in range(3)] if rows * cols <= 32 else \
        [int(st[i, 2 * j]) | (int(st[i, 2 * j + 1]) << 32) for j in range(3)]
    for cell in range(rows * cols):
        assert sum(((planes[j] >> cell) & 1) << j for j in range(3)) == code[cell]
    assert all(p >> (rows * cols) == 0 for p in planes)
    with pytest.raises(ValueError):
        layout.pack_state(np.full((1, rows, cols), 7), np.zeros((1, 2), int))


def test_slab_layout():
    rng = np.random.default_rng(0)
    orbs = rng.integers(0, 7, size=(9, 37)).astype(np.uint8)
    slab = layout.pack_slab(orbs)
    assert slab.shape == (9, 5) and slab.dtype == np.uint32
    for d in range(37):
        assert np.array_equal((slab[:, d >> 3] >> (4 * (d & 7))) & 15, orbs[:, d])


def test_make_config_translates_team_and_skyfall():
    team = [dict(color_1='red', color_2=None, atk=12.5), dict(color_1='dark', color_2='light', atk=3)]
    cfg = ppad_b200.make_config(6, 7, skyfall=[0.5, 0.5, 0, 0, 0, 0, 0, 0, 0, 0], skyfall_damage=False, team=team, seed=(1 << 63) + 5)
    assert (cfg.rows, cfg.cols, cfg.skyfall_damage, cfg.team_size) == (6, 7, 0, 2)
    assert list(cfg.team_color_1)[:2] == [0, 4] and list(cfg.team_color_2)[:2] == [-1, 3]
    assert list(cfg.team_atk)[:2] == [12.5, 3.0] and cfg.seed == (1 << 63) + 5
    from ppad_b200.pad import parameters
    assert parameters.int2english[3] == 'light' and parameters.match_n == 3
    assert [u['color_1'] for u in parameters.default_team] == ['dark', 'light', 'blue', 'green', 'red', 'blue']
    o = orc.make_cfg()
    d = ppad_b200._cabi.default_config()
    assert [parameters.english2int[u['color_1']] for u in parameters.default_team] == list(d.team_color_1)[:6] == list(o.team_c1)[:6]


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may import oracle/ (the checker is never the product)"""
    pkg = os.path.join(ROOT, "p-pad_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "libppad_oracle" not in text and "ppad_oracle.h" not in text and "hostcore" not in text.replace(
                    "tests/hostcore", ""), f
    for f in ("__graft_entry__.py", "bench.py"):
        text = open(os.path.join(ROOT, f)).read()
        assert "import oracle" in text          # allowed there, as the checker / the CPU baseline
    shim = open(os.path.join(ROOT, "ppad_b200.py")).read()
    assert "oracle" not in shim


def test_solved_board_generator_is_fully_matched():
    from ppad_b200.data import solved_boards
    assert len(solved_boards.boards) == 16 and solved_boards.boards[0].shape == (5, 6)
    for rows, cols in ((5, 6), (6, 7)):
        for b in solved_boards.generate_solved_boards(rows, cols, 20, seed=3):
            bb = np.ascontiguousarray(b, np.int8)
            assert orc.cancel(bb) and (bb == -1).all()     # a pass cancels every orb
