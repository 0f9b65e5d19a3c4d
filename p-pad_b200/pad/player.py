"""PlayerAction -- the action space object of the reference env (ppad/pad/player.py:22-66)."""
import numpy as np


class PlayerAction:
    # player.py:24-30
    player_action = ['up', 'down', 'left', 'right', 'pass']
    opposite_actions = {'up': 'down', 'down': 'up', 'left': 'right', 'right': 'left', 'pass': 'pass'}

    def __init__(self, finger, dim_row, dim_col, env=None, seed=0):
        self.previous_action = None
        self.finger = finger
        self.dim_row = dim_row
        self.dim_col = dim_col
        self._env = env          # the PAD facade that owns this action space (shares its GPU slab)
        self._seed = seed
        self._venv = None
        self._draws = 0

    def _slab(self):
        if self._env is not None:
            return self._env._venv
        if self._venv is None:
            from ..batched import BatchedPAD
            self._venv = BatchedPAD(1, self.dim_row, self.dim_col, seed=self._seed)
        return self._venv

    def sample(self, type='random', include_pass=False):
        """A uniformly random legal move that is not the exact reverse of previous_action (player.py:38-66),
        drawn on the GPU (ppad_sample_actions): same support and distribution as the reference's rejection loop,
        Philox instead of the process-global Mersenne Twister."""
        if type != 'random':
            raise Exception('ERROR: Unknown player action type!')
        env = self._env
        if env is not None and getattr(env, 'rng', 'philox') == 'reference':
            # the reference's own rejection loop over the global Mersenne Twister (player.py:46-62): in this mode the
            # host draws the numbers so that a seeded run replays the reference draw for draw
            import random
            num_action = len(self.player_action) - 1 + (1 if include_pass else 0)
            while True:
                current = self.player_action[random.randint(0, num_action - 1)]
                if self.opposite_actions[current] == self.previous_action:
                    continue
                if (self.finger[0] >= self.dim_row - 1 and current == 'down') or (self.finger[0] <= 0 and current == 'up'):
                    continue
                if (self.finger[1] >= self.dim_col - 1 and current == 'right') or (self.finger[1] <= 0 and current == 'left'):
                    continue
                return current
        self._draws += 1
        if env is not None:
            env._want_presample = True     # the caller samples: let the next step bring the following move along
        if (env is not None and not include_pass and env._presample[0] == self._draws
                and np.array_equal(np.asarray(self.finger), env.finger) and env._device_is_current()):
            # fastest path: the last step already ran the sample kernel for exactly this draw on the state it left
            a = env._presample[1]
        elif env is not None and np.array_equal(np.asarray(self.finger), env.finger):
            # fast path: the env's device record already holds (finger, previous_action)
            env._sync_to_device()
            a = int(env._venv.sample_actions(include_pass=include_pass, step=0x40000000 + self._draws)[0])
        else:
            venv = self._slab()
            prev = self.previous_action if self.previous_action in self.player_action else 'pass'
            board = np.zeros((1, self.dim_row, self.dim_col), dtype=np.int64)
            venv.set_state(board, np.asarray(self.finger).reshape(1, 2),
                           prev=np.array([self.player_action.index(prev)], np.uint8))
            a = int(venv.sample_actions(include_pass=include_pass, step=0x40000000 + self._draws)[0])
            if env is not None:
                env._dev_shadow = None     # the env's device record was used as scratch: re-pack before its next step
        if a == 255:
            raise Exception('ERROR: no legal move from finger {0}'.format(self.finger))
        return self.player_action[a]
