"""Harness that runs the UNMODIFIED reference (nuwapi/P-PAD) in-process as the parity oracle.

TEST INFRASTRUCTURE ONLY.  It is importable only where ``/root/reference`` exists (the build
container); the GPU box never has it, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py``
imports this module.  It is used by ``gen_golden.py`` to produce the committed fixtures under
``tests/golden/`` and by the (skippable) live differential tests.

Three in-process patches, none of which touch the game logic:
  1. ``tensorflow`` is stubbed (``ppad/__init__.py:42`` imports the TF1 agent).
  2. ``ppad.pad.game.datetime.now()`` returns an int (``game.py:72,302`` seed with a datetime,
     a TypeError on Python >= 3.11).
  3. ``PAD.random_orb`` (``game.py:458-470``) can be replaced by a callable that pops from an
     injected orb sequence; draws are consumed row-major per empty cell (``game.py:253-255``).
"""
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PPAD_REFERENCE_ROOT", "/root/reference")

ENGLISH2INT = {"red": 0, "blue": 1, "green": 2, "light": 3, "dark": 4, "heal": 5,
               "jammer": 6, "poison": 7, "mortal poison": 8, "bomb": 9}

_ref = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ppad", "pad"))


class _FrozenClock:
    """Stands in for ``datetime`` in game.py: ``random.seed(datetime.now())`` gets a settable int."""
    value = 20181019

    @classmethod
    def now(cls):
        return cls.value


def set_clock(value):
    _FrozenClock.value = int(value)


def load():
    """Import the reference with patches 1-2 and return a namespace of its modules."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import ppad  # noqa: F401
        import ppad.pad.game as game
        import ppad.pad.utils as pad_utils
        import ppad.pad.player as player
        import ppad.pad.parameters as parameters
        import ppad.agent.agent01 as agent01
        import ppad.agent.utils as agent_utils
        import ppad.data.dataset01.smart_data as smart_data
        import ppad.data.dataset01.solved_boards as solved_boards
    game.datetime = _FrozenClock
    _ref = types.SimpleNamespace(ppad=ppad, game=game, pad_utils=pad_utils, player=player,
                                 parameters=parameters, agent01=agent01, agent_utils=agent_utils,
                                 smart_data=smart_data, solved_boards=solved_boards)
    return _ref


class InjectedSkyfall:
    """Patch 3: a stand-in for ``PAD.random_orb`` that pops from a fixed orb sequence."""

    def __init__(self, orbs):
        self.orbs = [int(o) for o in orbs]
        self.cursor = 0

    def __call__(self, prob):
        if self.cursor >= len(self.orbs):
            raise IndexError("injected skyfall exhausted")
        o = self.orbs[self.cursor]
        self.cursor += 1
        return o


def make_env(rows=5, cols=6, skyfall_damage=True, board=None, finger=None, team=None, skyfall=None):
    ref = load()
    kw = dict(dim_row=rows, dim_col=cols, skyfall_damage=skyfall_damage)
    if team is not None:
        kw["team"] = team
    if skyfall is not None:
        kw["skyfall"] = skyfall
    if board is None:
        # constructor would rejection-sample with the global MT; give it a dummy and let callers reset
        board = -np.ones((rows, cols), dtype=int)
    if finger is None:
        finger = np.zeros(2, dtype=int)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = ref.game.PAD(board=np.array(board, dtype=int), finger=np.array(finger, dtype=int), **kw)
    return env


def resolve(env, board, skyfall_damage, orbs):
    """Run ``PAD.calculate_reward`` (game.py:364-405) with an injected skyfall sequence.

    Returns dict(reward, combos=[[(colour, N), ...] per cascade iteration], final_board, consumed, depth).
    The final board is captured from the last ``cancel`` call's (mutated) argument.
    """
    ref = load()
    inj = InjectedSkyfall(orbs)
    seen = []
    real_cancel = ref.pad_utils.cancel

    def spy_cancel(b):
        combos = real_cancel(b)
        seen.append((b, combos))
        return combos

    old_orb = env.random_orb
    env.random_orb = inj
    ref.pad_utils.cancel = spy_cancel
    ref.game.pad_utils.cancel = spy_cancel
    try:
        reward = env.calculate_reward(board=np.array(board, dtype=int), skyfall_damage=skyfall_damage)
    finally:
        ref.pad_utils.cancel = real_cancel
        ref.game.pad_utils.cancel = real_cancel
        env.random_orb = old_orb
    per_iter = [[(ENGLISH2INT[c["color"]], int(c["N"])) for c in combos] for _, combos in seen if len(combos)]
    return dict(reward=float(reward), combos=per_iter, final_board=np.array(seen[-1][0], dtype=np.int8),
                consumed=inj.cursor, depth=len(per_iter))
