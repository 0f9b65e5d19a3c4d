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


def test_slab_layouts_planes3_and_chunks10():
    """the two 3-bit slab layouts as include/ppad_b200.h states them"""
    rng = np.random.default_rng(1)
    orbs = rng.integers(0, 7, size=(11, 47)).astype(np.uint8)
    p3 = layout.pack_slab_planes3(orbs)
    assert p3.shape == (11, 6) and p3.dtype == np.uint32
    c10 = layout.pack_slab_chunks10(orbs, stride=16)
    assert c10.shape == (5, 16) and c10.dtype == np.uint32 and not c10[:, 11:].any() and not (c10 >> 30).any()
    for d in range(47):
        g, k = d >> 5, d & 31
        code = sum(((p3[:, 3 * g + j] >> k) & 1) << j for j in range(3))
        assert np.array_equal(code, orbs[:, d])
        c, k = divmod(d, 10)
        code = sum(((c10[c, :11] >> (10 * j + k)) & 1) << j for j in range(3))
        assert np.array_equal(code, orbs[:, d])


def _encode_compact(n, words, action, reward, info, state, dictionary=None, seed=0):
    """Test-side writer of the compact result stream exactly as include/ppad_b200.h describes it (tiles of 256 boards,
    record blocks claimed in arbitrary order)."""
    T = layout.TILE
    tiles = (n + T - 1) // T
    nib = np.zeros(tiles * T, np.uint8)
    has = info != 0
    nib[:n] = np.where(action > 5, 7, action) | (has.astype(np.uint8) << 3)
    nibbles = (nib[0::2] | (nib[1::2] << 4)).view(np.uint32)
    index = {}
    if dictionary is not None:
        for j, (r, i) in enumerate(zip(dictionary[0].view(np.uint64), dictionary[1])):
            index.setdefault((int(r), int(i)), j)
    blocks, tile_line = {}, np.zeros(tiles, np.uint32)
    for t in range(tiles):
        sel = np.flatnonzero(has[t * T:(t + 1) * T]) + t * T
        k = len(sel)
        done = sel[(info[sel] >> 31) != 0]
        if dictionary is None:
            body = reward[sel].tobytes() + info[sel].tobytes()
        else:
            codes = np.array([index.get((int(reward[i:i + 1].view(np.uint64)[0]), int(info[i])), 0xFFFF) for i in sel], np.uint16)
            esc = sel[codes == 0xFFFF]
            body = codes.tobytes()
            body += b"\0" * (-len(body) % 8) + reward[esc].tobytes() + info[esc].tobytes()
        body += b"\0" * (-len(body) % 16) + state[done].tobytes()
        body += b"\0" * (-len(body) % 128)
        blocks[t] = body
    order = np.random.default_rng(seed).permutation(tiles)
    records, line = b"", 0
    for t in order:
        tile_line[t] = line
        records += blocks[t]
        line += len(blocks[t]) // 128
    summary = np.array([line, tiles, 0, 0], np.uint32)
    return nibbles, tile_line, np.frombuffer(records + b"\0" * 128, np.uint8), summary


@pytest.mark.parametrize("n", [1, 255, 256, 1000, 5000])
@pytest.mark.parametrize("with_dict", [False, True])
def test_decode_compact_round_trip(n, with_dict):
    """layout.decode_compact against a writer of the documented format: plain 12-byte records and dictionary-coded ones
    (16-bit codes, escapes in full), fresh-episode state records, ragged last tile."""
    rng = np.random.default_rng(n)
    words = 4
    action = rng.integers(0, 5, n).astype(np.uint8)
    action[rng.random(n) < 0.05] = 255
    pool_r = np.round(rng.random(40) * 1e4, 2)
    pool_i = rng.integers(1, 2 ** 31, 40).astype(np.uint32)
    pick = rng.integers(0, 40, n)
    reward, info = pool_r[pick].copy(), pool_i[pick].copy()
    rare = rng.random(n) < 0.1                                   # pairs no dictionary holds
    reward[rare] = rng.random(int(rare.sum())) * 1e5
    info[rng.random(n) < 0.1] |= np.uint32(1 << 31)             # finished episodes: a state record follows
    none = rng.random(n) < 0.4                                   # no record at all
    reward[none], info[none] = 0.0, 0
    state = rng.integers(0, 2 ** 32, size=(n, words), dtype=np.uint64).astype(np.uint32)
    dictionary = (pool_r[:30].copy(), pool_i[:30].copy()) if with_dict else None     # 10 pool pairs stay outside
    stream = _encode_compact(n, words, action, reward, info, state, dictionary, seed=n)
    got = layout.decode_compact(n, words, *stream, dictionary=dictionary)
    assert np.array_equal(got["action"], action)
    assert np.array_equal(got["reward"].view(np.uint64), reward.view(np.uint64)) and np.array_equal(got["info"], info)
    done = np.flatnonzero((info >> 31) != 0)
    assert np.array_equal(got["reset_index"], done) and np.array_equal(got["reset_state"], state[done])
    if with_dict and n >= 1000:
        assert 0 < got["escapes"] < int((info != 0).sum())


def test_chunk_bytes_on_demand_counts_sectors():
    """the e2e leg's host-to-device accounting for chunk-major slabs: a chunk beyond the eager ones is fetched for a
    32-byte sector of neighbouring boards when any of them drew far enough"""
    info = np.zeros(64, np.uint32)
    assert layout.chunk_bytes_on_demand(info, 6, 2, 3, 9) == 0
    info[5] = 19 << 16                      # 19 orbs: reaches chunk 3 (orbs 18..23) of a uint16 slab with 3 eager chunks
    assert layout.chunk_bytes_on_demand(info, 6, 2, 3, 9) == 32
    info[6] = 25 << 16                      # same sector of 16 boards: chunk 3 once, chunk 4 once
    assert layout.chunk_bytes_on_demand(info, 6, 2, 3, 9) == 64
    info[40] = 18 << 16                     # exactly the eager orbs: nothing more
    assert layout.chunk_bytes_on_demand(info, 6, 2, 3, 9) == 64
    info[41] = 255 << 16                    # saturated: every chunk the slab has (3..8) in that other sector
    assert layout.chunk_bytes_on_demand(info, 6, 2, 3, 9) == 64 + 6 * 32
    assert layout.chunk_bytes_on_demand(info, 10, 4, 2, 5) == 32 * (1 + 3)   # uint32 chunks of 10 orbs, sectors of 8 boards


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
