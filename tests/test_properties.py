"""Property tests (hypothesis) of the per-board device code, host build (tests/hostcore): invariants the
game logic must satisfy for ANY board, independent of the oracle.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle as orc
import hostcore_lib as hc

GEOMS = [(4, 5), (5, 6), (6, 7)]


@st.composite
def boards(draw, allow_empty=False):
    """a board plus the kernel family that runs it: 3-bit orb codes (types 0..6) or 4-bit (types 0..9)"""
    rows, cols = draw(st.sampled_from(GEOMS))
    orb_bits = draw(st.sampled_from([3, 4]))
    ncol = draw(st.integers(2, 7 if orb_bits == 3 else 10))
    lo = -1 if allow_empty else 0
    cells = draw(st.lists(st.integers(lo, ncol - 1), min_size=rows * cols, max_size=rows * cols))
    if orb_bits == 4 and draw(st.booleans()):
        cells = [c if c < 0 else 9 - c for c in cells]          # make the high types (poison .. bomb) the common ones
    return np.array(cells, np.int8).reshape(1, rows, cols), orb_bits


@settings(max_examples=300, deadline=None)
@given(boards(allow_empty=True))
def test_cancel_is_idempotent_and_erases_exactly_the_counted_cells(bo):
    b, orb_bits = bo
    before = b.copy()
    n1, log1 = hc.cancel(b, log_cap=32, orb_bits=orb_bits)
    erased = int(((before != -1) & (b == -1)).sum())
    assert erased == sum(int(v) >> 8 for v in log1[0, :n1[0]])          # sum of N = erased cells
    assert all((int(v) >> 8) >= 3 for v in log1[0, :n1[0]])
    assert ((b == before) | (b == -1)).all()
    after = b.copy()
    n2, _ = hc.cancel(b, log_cap=32, orb_bits=orb_bits)
    assert n2[0] == 0 and np.array_equal(b, after)                       # nothing left to cancel


@settings(max_examples=300, deadline=None)
@given(boards(allow_empty=True))
def test_drop_is_a_stable_column_compaction(bo):
    b, orb_bits = bo
    before = b.copy()
    hc.drop(b, orb_bits=orb_bits)
    for c in range(b.shape[2]):
        col_before = [v for v in before[0, :, c] if v != -1]
        col_after = list(b[0, :, c])
        k = len(col_after) - len(col_before)
        assert col_after[:k] == [-1] * k and col_after[k:] == col_before
    again = b.copy()
    hc.drop(again, orb_bits=orb_bits)
    assert np.array_equal(again, b)


@settings(max_examples=200, deadline=None)
@given(boards(), st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 40), st.booleans())
def test_resolve_invariants(bo, seed, env, skyfall_damage):
    b, orb_bits = bo
    rows, cols = b.shape[1:]
    sky = [0.1] * 10 if orb_bits == 4 else None
    cfg = hc.cfg_from_oracle(orc.make_cfg(rows, cols, skyfall_damage=skyfall_damage, seed=seed, cascade_cap=0, skyfall=sky),
                             orb_bits=orb_bits)
    fingers = np.zeros((1, 2), np.int32)
    keep = b.copy()
    _, reward, info, final, log = hc.step_batch(cfg, b, fingers, np.full(1, 4, np.uint8), None, actions=np.full(1, 4, np.uint8),
                                                env0=env, step=3, want_final=True, log_cap=64)
    combos, depth, consumed, flags = int(info[0]) & 0xFF, (int(info[0]) >> 8) & 0xFF, (int(info[0]) >> 16) & 0xFF, int(info[0]) >> 24
    assert np.array_equal(b, keep)                                       # a pass evaluates a copy (game.py:366)
    assert flags == 0 and reward[0] >= 0 and combos >= depth and (combos > 0) == (depth > 0)
    if combos < 64 and consumed < 255:
        sizes = [int(v) >> 8 for v in log[0, :combos]]
        if skyfall_damage:
            assert consumed == sum(sizes)                                # every erased orb is refilled exactly once
            assert final.min() >= 0 and orc.cancel(np.ascontiguousarray(final[0])) == []   # stable, full board
        else:
            assert consumed == 0 and depth <= 1 and int((final == -1).sum()) == sum(sizes)
    # the same board under another env id / batch position gives the same no-skyfall evaluation
    if not skyfall_damage:
        _, r2, i2, _, _ = hc.step_batch(cfg, keep.copy(), fingers, np.full(1, 4, np.uint8), None,
                                        actions=np.full(1, 4, np.uint8), env0=env + 12345, step=9)
        assert r2[0] == reward[0] and i2[0] == info[0]


@settings(max_examples=200, deadline=None)
@given(st.sampled_from(GEOMS), st.integers(0, 41), st.integers(0, 4), st.integers(0, 2 ** 32 - 1))
def test_sampled_moves_are_legal_and_moves_are_reversible(geom, cell, prev, seed):
    rows, cols = geom
    cell %= rows * cols
    cfg = hc.cfg_from_oracle(orc.make_cfg(rows, cols, seed=seed))
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 6, size=(1, rows, cols)).astype(np.int8)
    f = np.array([[cell // cols, cell % cols]], np.int32)
    keep_b, keep_f = b.copy(), f.copy()
    p = np.array([prev], np.uint8)
    a, _, info, _, _ = hc.step_batch(cfg, b, f, p, None, actions=None, eval_moves=False, step=1)
    assert a[0] < 4 and (orc.legal_mask(rows, cols, keep_f[0], prev) >> a[0]) & 1 and int(info[0]) == 0
    assert sorted(b.reshape(-1)) == sorted(keep_b.reshape(-1))           # a move permutes the orbs
    back = np.array([a[0] ^ 1], np.uint8)                                # the opposite move undoes it
    hc.step_batch(cfg, b, f, p, None, actions=back, eval_moves=False, step=2)
    assert np.array_equal(b, keep_b) and np.array_equal(f, keep_f)
