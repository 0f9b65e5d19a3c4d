"""PAD -- single-env drop-in for ``ppad.PAD`` (ppad/pad/game.py:28-470) backed by the CUDA kernels.

Same constructor keywords, attributes, methods, action strings and error behaviour as the
reference, so ``projects/simulation.py``, ``ppad/data/dataset01/smart_data.py`` and ``Agent01.act``
can use it unchanged.  Every method is a 1-board launch of the batched kernels (ppad_step,
ppad_reset, ...): this facade exists for compatibility, the throughput path is ``BatchedPAD``.

Differences that are deliberate ("quirks"):
  * randomness is Philox4x32-10 keyed by ``seed`` (new keyword; default: time based, like the
    reference's ``random.seed(datetime.now())``), not the process-global Mersenne Twister;
  * an unknown action string raises ValueError (the reference silently records it as accepted);
  * ``action is 'pass'`` identity tests become equality tests;
  * orb types 7..9 and boards other than 5x6 / 6x7 raise (no kernel instantiation).
"""
import time

import numpy as np

from . import parameters
from . import player
from . import utils as pad_utils
from ..batched import ACTION_IDS, BatchedPAD
from .. import _cabi


class PAD:
    """
    Board orientation (game.py:30-38): board[row, col], row 0 is the top, gravity towards larger rows,
    finger = (row, col).
    """

    def __init__(self,
                 dim_row=5,
                 dim_col=6,
                 skyfall=parameters.default_skyfall,
                 skyfall_locked=parameters.default_skyfall_locked,
                 skyfall_enhanced=parameters.default_skyfall_enhanced,
                 skyfall_damage=True,
                 buff=None,
                 team=parameters.default_team,
                 enemy=parameters.default_enemy,
                 board=None,
                 finger=None,
                 seed=None,
                 device=None):
        # the argument checks of game.py:54-68, with the reference's messages
        def require(ok, message):
            if not ok:
                raise Exception('ERROR: ' + message)
        require(abs(sum(skyfall) - 1) <= 1e-4, 'Skyfall rates should add up to 1!')
        require(all(p >= 0 for p in skyfall), 'Skyfall rates should be a positive number!')
        require(all(0 <= p <= 1 for p in skyfall_locked), 'Locked orb probability is between 0 and 1!')
        require(all(0 <= p <= 1 for p in skyfall_enhanced), 'Enhanced orb probability is between 0 and 1!')
        require(dim_col >= 0 and dim_row >= 0, 'Board dimensions should be positive integers!')
        require(dim_col * dim_row >= 13, 'The board should allow at least 13 orbs. Please increase board size!')

        self.dim_row = dim_row
        self.dim_col = dim_col
        self.skyfall = skyfall
        self.skyfall_locked = skyfall_locked
        self.skyfall_enhanced = skyfall_enhanced
        self.skyfall_damage = skyfall_damage
        self.buff = buff
        self.team = team
        self.enemy = enemy
        self.observation_space = None
        self.render_dict = parameters.default_render_dict
        self.seed = int(time.time_ns()) & 0xFFFFFFFFFFFFFFFF if seed is None else int(seed)
        try:
            self._venv = BatchedPAD(1, dim_row, dim_col, skyfall=list(skyfall), skyfall_damage=skyfall_damage,
                                    team=team, seed=self.seed, device=device, cascade_cap=0, reset_cap=0)
        except _cabi.PPADError as e:
            if e.status == _cabi.ERR_UNSUPPORTED:
                raise NotImplementedError(e.detail)
            raise Exception(e.detail)
        self._tick = 0            # Philox step counter: one per randomised call

        self.board = -1 * np.ones((self.dim_row, self.dim_col), dtype=int)
        self.locked = np.zeros((self.dim_row, self.dim_col), dtype=int)
        self.enhanced = np.zeros((self.dim_row, self.dim_col), dtype=int)
        self.finger = np.zeros(2, dtype=int)
        # episode record (game.py:105-115)
        self.boards = []
        self.fingers = []
        self.observations = []
        self.actions = []
        self.rewards = []
        self.action_space = player.PlayerAction(finger=self.finger, dim_row=self.dim_row, dim_col=self.dim_col,
                                                env=self, seed=self.seed)
        self.reset(board=board, finger=finger)

    # ------------------------------------------------------------------ device round trips
    def _next_tick(self):
        self._tick += 1
        return self._tick & 0x3FFFFFFF

    def _push(self, board=None):
        prev = self.action_space.previous_action
        prev_id = ACTION_IDS.get(prev, 4)
        flags = self._venv.set_state(np.asarray(self.board if board is None else board)[None],
                                     np.asarray(self.finger).reshape(1, 2), prev=np.array([prev_id], np.uint8))
        if int(flags[0]) & _cabi.FLAG_UNSUPPORTED_ORB:
            raise NotImplementedError('orb types 7..9 do not fit the 3-bit board state of ppad_b200')

    def _pull(self):
        st = self._venv.get_state()
        self.board = st.boards[0].cpu().numpy().astype(int)
        self.finger = st.fingers[0].cpu().numpy().astype(int)

    # ------------------------------------------------------------------ game.py:122-155
    def apply_action(self, action):
        """Swap the finger orb with its neighbour in direction `action`; 'rejected' when that leaves the board."""
        if action not in ('up', 'down', 'left', 'right'):
            return 'accepted'      # the reference falls through its if/elif chain without doing anything
        self._push()
        out = self._venv.step(np.array([ACTION_IDS[action]], np.uint8), step=self._next_tick())
        flags = (int(out.info[0]) >> 24) & 0xFF
        if flags & _cabi.FLAG_REJECTED:
            return 'rejected'
        self._pull()
        return 'accepted'

    # ------------------------------------------------------------------ game.py:157-199
    def damage(self, combos):
        """Total team damage of a list of combos (dicts as produced by cancel())."""
        if len(combos) == 0:
            return 0.0
        log = np.zeros((1, len(combos)), np.uint16)
        for i, combo in enumerate(combos):
            n = float(combo['N'])
            if n != int(n) or not 0 <= int(n) <= 255:
                raise NotImplementedError('combo sizes must be integers in 0..255')
            log[0, i] = parameters.english2int[combo['color']] | (int(n) << 8)
        if len(combos) > 255:
            raise NotImplementedError('at most 255 combos per damage() call')
        reward = self._venv.damage(log.view(np.int16), np.array([len(combos)], np.uint8))
        return np.float64(reward[0].item())

    # ------------------------------------------------------------------ game.py:201-221
    def drop(self, board):
        """Move all orbs down into empty cells; `board` is updated in place and returned."""
        keep = self._venv.state.clone()
        self._push(board)
        self._venv.drop()
        board[...] = self._venv.get_state().boards[0].cpu().numpy()
        self._venv.state.copy_(keep)
        return board

    def visualize(self, filename, shrink=3, animate=True):
        raise NotImplementedError('GIF rendering (game.py:223-230) is out of scope of ppad_b200')

    # ------------------------------------------------------------------ game.py:232-256
    def fill_board(self, board=None, reset=False):
        """Fill empty cells (all cells when reset) row-major with skyfall orbs; in place."""
        if board is None:
            board = self.board
        keep = self._venv.state.clone()
        self._push(board)
        self._venv.fill_board(reset=reset, step=self._next_tick())
        board[...] = self._venv.get_state().boards[0].cpu().numpy()
        self._venv.state.copy_(keep)
        return board

    def render(self, board=None):
        """Print a coloured board to the terminal (game.py:260-280)."""
        if board is None:
            board = self.board
        print('+-----------+')
        for x in range(self.dim_row):
            if x > 0:
                print('|-----------|')
            line = '|'
            for y in range(self.dim_col):
                letter, colour = self.render_dict[int(board[x, y])]
                if x == self.finger[0] and y == self.finger[1]:
                    line += letter + '|'
                else:
                    line += colour + letter + '\033[0m' + '|'
            print(line)
        print('+-----------+')

    # ------------------------------------------------------------------ game.py:282-327
    def reset(self, board=None, finger=None):
        """Start a new episode; returns the state (board, finger)."""
        self.observations = []
        self.boards = []
        self.fingers = []
        self.actions = []
        self.rewards = []
        self.action_space.previous_action = 'pass'

        tick = self._next_tick()
        if board is None or finger is None:
            # combo-free random board by rejection and/or a random finger, on the GPU
            fin = None if finger is None else np.asarray(finger).reshape(1, 2)
            self._venv.reset(fingers=fin, step=tick)
            st = self._venv.get_state()
            if board is None:
                self.board = st.boards[0].cpu().numpy().astype(int)
            if finger is None:
                self.finger = st.fingers[0].cpu().numpy().astype(int)
        if board is not None:
            self.board = np.copy(board)
        if finger is not None:
            self.finger = np.copy(finger)
        self.action_space.finger = self.finger

        self._record_observation()
        return self.state()

    def _record_observation(self):
        """append copies of (board, finger); boards[-1] / fingers[-1] alias the tuple's members like the reference's"""
        snapshot = (self.board.copy(), self.finger.copy())
        self.observations.append(snapshot)
        self.boards.append(snapshot[0])
        self.fingers.append(snapshot[1])

    def state(self):
        return self.boards[-1], self.fingers[-1]

    # ------------------------------------------------------------------ game.py:329-362
    def step(self, action, verbose=False):
        """One env step; returns None like the reference.  'pass' evaluates the board (reward = damage of all
        combos, cascades included when skyfall_damage) without modifying it; a move swaps orbs, reward 0."""
        reward = 0
        self.action_space.finger = self.finger
        if action == 'pass':
            reward = self.calculate_reward(board=self.board, skyfall_damage=self.skyfall_damage, verbose=verbose)
        else:
            if action not in ('up', 'down', 'left', 'right'):
                raise ValueError('Unknown action {!r}: expected up, down, left, right or pass'.format(action))
            comment = self.apply_action(action)
            if comment == 'accepted':
                if verbose is True:
                    print('Action: {0}. Board after orb move:'.format(action))
                    self.render()
                self.action_space.previous_action = action
                self.action_space.finger = self.finger
            else:
                raise ValueError('Invalid move, you cannot move off the board!\n'
                                 'Finger = {}, action = {}'.format(self.finger, action))
        if action != 'pass':
            self._record_observation()     # a pass records action and reward only (game.py:357-362)
        self.actions.append(action)
        self.rewards.append(reward)

    # ------------------------------------------------------------------ game.py:364-405
    def calculate_reward(self, board=None, skyfall_damage=True, verbose=False):
        """Damage of all combos of `board` (a copy is evaluated; cancel -> drop -> skyfall until stable when
        skyfall_damage).  One fused kernel launch."""
        keep = self._venv.state.clone()
        self._push(self.board if board is None else board)
        out = self._venv.calculate_reward(step=self._next_tick(), skyfall_damage=skyfall_damage,
                                          want_final=verbose, log_cap=64 if verbose else 0)
        reward = np.float64(out.reward[0].item())
        if verbose is True:
            n = int(out.info[0]) & 0xFF
            print('Board after the cascade:')
            self.render(self._venv.get_state(out.final_state).boards[0].cpu().numpy())
            log = out.combo_log[0, :min(n, 64)].cpu().numpy().view(np.uint16)
            print([{'color': parameters.int2english[int(v) & 0xFF], 'N': float(int(v) >> 8)} for v in log])
            print('The total damange is:', reward)
        self._venv.state.copy_(keep)
        return reward

    # ------------------------------------------------------------------ game.py:407-427
    def backtrack(self, n=1):
        """Undo the last n recorded steps."""
        if n >= len(self.boards):
            print('ERROR: Cannot backtrack more than the length of the trajectory.')
            exit()
        for record in (self.observations, self.boards, self.fingers, self.actions, self.rewards):
            del record[len(record) - n:]
        self.board, self.finger = self.boards[-1].copy(), self.fingers[-1].copy()
        self.action_space.previous_action = self.actions[-1] if self.actions else 'pass'
        self.action_space.finger = self.finger

    # ------------------------------------------------------------------ game.py:429-456
    def swap(self, x1, y1, x2, y2, move_finger):
        """Exchange two orbs of self.board; optionally move the finger onto the second one.  Host-side
        bookkeeping only: the batched path does the swap inside ppad_step."""
        self.board[x1, y1], self.board[x2, y2] = self.board[x2, y2], self.board[x1, y1]
        if move_finger:
            self.finger[0] = x2
            self.finger[1] = y2

    # ------------------------------------------------------------------ game.py:458-470
    def random_orb(self, prob=None):
        """One skyfall draw following `prob` (defaults to self.skyfall), on the GPU's Philox stream."""
        if prob is not None and list(prob) != list(self.skyfall):
            env = PAD(self.dim_row, self.dim_col, skyfall=list(prob), seed=self.seed + self._next_tick())
            return env.random_orb()
        one = -1 * np.ones((self.dim_row, self.dim_col), dtype=int)
        one[1:, :] = 0
        one[0, 1:] = 0
        return int(self.fill_board(board=one)[0, 0])
