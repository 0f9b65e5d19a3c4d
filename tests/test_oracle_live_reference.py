"""Live differential: the C oracle against the UNMODIFIED reference imported in-process, on fresh random
cases (not the committed fixtures).  Runs only where /root/reference exists (the build container); skipped on
the GPU box.  CPU only."""
import os
import random
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_harness as rh  # noqa: E402
import oracle as orc  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference sources not present")


def bits(x):
    return np.asarray(x, np.float64).view(np.uint64)


@pytest.mark.parametrize("rows,cols", [(4, 5), (5, 6), (6, 7)])
def test_resolve_live(rows, cols):
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    env = rh.make_env(rows, cols, skyfall_damage=True)
    cfgs = {sd: orc.make_cfg(rows, cols, skyfall_damage=sd, cascade_cap=0) for sd in (0, 1)}
    for k in range(160):
        ncol = int(rng.integers(2, 7))
        # every fourth case draws its colours from all ten orb types (poison, mortal poison, bomb: game.py:243-246)
        pal = np.arange(ncol) if k % 4 else rng.permutation(10)[:ncol]
        board = pal[rng.integers(0, ncol, size=(rows, cols))].astype(np.int8)
        orbs = pal[rng.integers(0, ncol, size=400)].astype(np.uint8)
        sd = bool(rng.integers(2))
        try:
            want = rh.resolve(env, board, sd, orbs)
        except IndexError:
            continue
        got = orc.resolve(cfgs[int(sd)], board, inj=orbs)
        assert bits(got["reward"]) == bits(want["reward"])
        assert got["combos"] == [c for it in want["combos"] for c in it]
        assert (got["depth"], got["consumed"]) == (want["depth"], want["consumed"])
        assert np.array_equal(got["final_board"], want["final_board"])


def test_reset_sample_and_onehot_live():
    ref = rh.load()
    seed = int.from_bytes(os.urandom(2), "little")
    rh.set_clock(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = ref.game.PAD()
    # the reference's reset leaves a combo-free board; its sampled moves are always in the oracle's legal set
    assert orc.cancel(np.ascontiguousarray(env.board, np.int8)) == []
    prev = 4
    for _ in range(60):
        a = env.action_space.sample()
        k = ["up", "down", "left", "right"].index(a)
        assert (orc.legal_mask(5, 6, env.finger, prev) >> k) & 1
        board, finger = np.ascontiguousarray(env.board, np.int8), env.finger.astype(np.int32).copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env.step(a)
        assert orc.apply_action(board, finger, k) == 1
        assert np.array_equal(board, env.board) and np.array_equal(finger, env.finger)
        prev = k
    import types
    me = types.SimpleNamespace(st_shape=(5, 6, 7))
    states = ref.agent01.Agent01.board_finger_to_state(me, env.board[None], env.finger[None])
    assert np.array_equal(orc.onehot(env.board[None].astype(np.int8), env.finger[None].astype(np.int32)), states.astype(np.float32))
    rewards = [0.0] * 7 + [float(random.Random(seed).uniform(1, 5e4))]
    assert np.allclose(orc.discount(rewards, 0.93, log10=True), ref.agent_utils.discount(rewards, 0.93, log10=True), rtol=4e-16)
