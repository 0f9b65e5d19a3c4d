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
  * orb types 7..9 (poison, mortal poison, bomb) switch the env to the 4-bit board kernels (orb_bits = 4), on
    construction when the skyfall produces them, or as soon as a board holding one is handed in; boards other
    than 4x5 / 5x6 / 6x7 raise (no kernel instantiation).
"""
import random
import time

import numpy as np

from . import parameters
from . import player
from . import utils as pad_utils  # noqa: F401  (the reference's game module exposes it under this name)
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
                 device=None,
                 rng='philox',
                 clock=None):
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
        # rng='philox' (default): every draw comes from the GPU's counter-based Philox streams keyed by `seed`.
        # rng='reference': every draw comes from Python's process-global `random` in exactly the order and number
        # the reference makes them (game.py:72,253-255,302-315,464; player.py:49) and is handed to the kernels as an
        # injected orb sequence -- after the same random.seed(...) the env reproduces the reference's trajectories
        # bit for bit.  `clock` stands in for datetime.now() in the reference's random.seed(datetime.now()) calls
        # (a callable returning the seed; None = do not reseed).
        if rng not in ('philox', 'reference'):
            raise Exception("ERROR: rng must be 'philox' or 'reference'")
        self.rng = rng
        self.clock = clock
        self._reseed()
        self._device = device
        self._make_venv(None)
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

    def _make_venv(self, orb_bits):
        """(re)create the one-board device slab; orb_bits None = 4 only if the skyfall produces orb types 7..9"""
        try:
            self._venv = BatchedPAD(1, self.dim_row, self.dim_col, skyfall=list(self.skyfall),
                                    skyfall_damage=self.skyfall_damage, team=self.team, seed=self.seed,
                                    device=self._device, cascade_cap=0, reset_cap=0, orb_bits=orb_bits)
        except _cabi.PPADError as e:
            if e.status == _cabi.ERR_UNSUPPORTED:
                raise NotImplementedError(e.detail)
            raise Exception(e.detail)
        self._io_setup()

    def _fit_codes(self, board):
        """A board with poison / mortal poison / bomb orbs (types 7..9, game.py:243-246) needs the 4-bit kernels."""
        b = np.asarray(board)
        if ((b < -1) | (b > 9)).any():
            raise NotImplementedError('orb types are -1 (empty) and 0..9 (game.py:236-246)')
        if self._venv.orb_bits == 3 and (b > 6).any():
            self._make_venv(4)

    # ------------------------------------------------------------------ rng='reference': the host only draws numbers
    def _reseed(self):
        if self.rng == 'reference' and self.clock is not None:
            random.seed(self.clock())          # game.py:72,302

    def _draw_orb(self):
        """one random_orb(self.skyfall) draw (game.py:458-470) from the global Mersenne Twister"""
        u = random.random()
        acc = 0
        for i, p in enumerate(self.skyfall):
            acc += p
            if acc >= u and p > 0:
                return i
        return len(self.skyfall) - 1

    def _speculate(self, run, unit, first=2):
        """Run `run(orbs)` -> number of orbs the kernel consumed, on a pre-drawn orb sequence, then leave the
        Mersenne Twister exactly where the reference would have left it (rewind, burn `consumed` draws)."""
        k = first * unit
        while True:
            state0 = random.getstate()
            orbs = np.array([self._draw_orb() for _ in range(k)], np.uint8)
            consumed = run(orbs)
            if consumed is not None and consumed <= k:
                random.setstate(state0)
                for _ in range(consumed):
                    random.random()
                return consumed
            random.setstate(state0)
            k *= 4
            if k > 1 << 14:
                raise RuntimeError('cascade / reset did not terminate within %d skyfall orbs' % k)

    # ------------------------------------------------------------------ device round trips
    # The env's board lives on the GPU as one packed record; `self.board` / `self.finger` are the host mirror the
    # reference's callers read.  A step is ONE kernel launch plus ONE 80-byte read-back (state, reward, info, action);
    # the host only converts formats (layout.pack_state / unpack_state).  If a caller assigned or edited the mirror,
    # the record is re-packed before the next launch.
    def _next_tick(self):
        self._tick += 1
        return self._tick & 0x3FFFFFFF

    def _io_setup(self):
        import ctypes
        import torch
        v = self._venv
        self._torch, self._ctypes = torch, ctypes
        # two 128-byte slots: [0] the env's record, [1] scratch for calculate_reward(board=...).
        # slot layout: state 0..64 (16..48 bytes used) | reward 64..72 | info 72..76 | action 76 | pre-sampled move 77 |
        # skyfall orbs consumed (uint16, not saturated at 255 like the info byte) 78..80
        self._io = torch.zeros(256, dtype=torch.uint8, device=v.device)
        self._io_host = torch.zeros(256, dtype=torch.uint8).pin_memory()
        self._io_np = self._io_host.numpy()
        v.state = self._io[:4 * v.words].view(torch.int32).view(1, v.words)   # the env slab IS slot 0
        self._action_bytes = torch.arange(5, dtype=torch.uint8, device=v.device)
        self._dev_shadow = None           # (board, finger, prev id) currently packed in slot 0
        # action_space.sample() right after a step is the common loop (visualize_episode.py:29-31): when the caller
        # samples, the next step's launch is followed by the sample kernel for the state it leaves behind and the
        # move comes back with the same read-back -- one device round trip per env.step(sample()) instead of two
        self._want_presample = False
        self._presample = (-1, 255)       # (draw index it was made for, action id)

    def _device_is_current(self):
        sh = self._dev_shadow
        return (sh is not None and sh[2] == self._prev_id() and np.array_equal(sh[0], self.board)
                and np.array_equal(sh[1], self.finger))

    def _prev_id(self):
        return ACTION_IDS.get(self.action_space.previous_action, 4)

    def _write_slot(self, slot, board, finger, prev_id):
        from .. import layout
        b = np.asarray(board)
        self._fit_codes(b)
        rec = layout.pack_state(b[None], np.asarray(finger).reshape(1, 2), [prev_id],
                                orb_bits=self._venv.orb_bits).view(np.uint8).reshape(-1)
        off = 128 * slot
        self._io_np[off:off + rec.size] = rec
        self._io[off:off + 64].copy_(self._io_host[off:off + 64], non_blocking=True)

    def _sync_to_device(self):
        cur = (self.board, self.finger, self._prev_id())
        sh = self._dev_shadow
        if sh is None or sh[2] != cur[2] or not np.array_equal(sh[0], cur[0]) or not np.array_equal(sh[1], cur[1]):
            self._write_slot(0, *cur)
            self._dev_shadow = (np.array(cur[0]), np.array(cur[1]), cur[2])

    def _launch(self, slot, action_id, skyfall_damage=None, final=False, log_cap=0, injected=None):
        """ppad_step for the one board in `slot`; returns (reward, info, action) after one read-back"""
        v, torch = self._venv, self._torch
        slab = None
        if injected is not None:
            from .. import layout
            slab = torch.from_numpy(layout.pack_slab(np.asarray(injected, np.uint8)[None]).view(np.int32)).to(v.device)
        base = self._io.data_ptr() + 128 * slot
        a = _cabi.StepArgs()
        a.n, a.env0, a.step = 1, v.env0, self._next_tick()
        a.skyfall_damage = -1 if skyfall_damage is None else int(bool(skyfall_damage))
        a.state, a.reward, a.info, a.action_out = base, base + 64, base + 72, base + 76
        a.consumed_out = base + 78
        a.actions = self._action_bytes.data_ptr() + action_id
        if slab is not None:
            a.slab, a.slab_words, a.n_inj = slab.data_ptr(), slab.shape[1], len(injected)
        extra = None
        if final or log_cap:
            extra = (torch.empty((1, v.words), dtype=torch.int32, device=v.device),
                     torch.zeros((1, max(1, log_cap)), dtype=torch.int16, device=v.device))
            a.final_state, a.combo_log, a.log_cap = extra[0].data_ptr(), extra[1].data_ptr(), max(1, log_cap)
        presample = slot == 0 and self._want_presample and self.rng == 'philox'
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_step(v._ctx, self._ctypes.byref(a), v._stream()))
            if presample:
                idx = self.action_space._draws + 1
                _cabi.check(v.lib.ppad_sample_actions(v._ctx, 1, v.env0, (0x40000000 + idx) & 0xFFFFFFFF, base, 0,
                                                      base + 77, v._stream()))
        self._io_host[128 * slot:128 * slot + 80].copy_(self._io[128 * slot:128 * slot + 80])   # blocking read-back
        raw = self._io_np[128 * slot:128 * slot + 80]
        if slot == 0:
            self._presample = (idx, int(raw[77])) if presample else (-1, 255)
        reward = np.float64(raw[64:72].view(np.float64)[0])
        info = int(raw[72:76].view(np.uint32)[0])
        self._last_consumed = int(raw[78:80].view(np.uint16)[0])
        return reward, info, int(raw[76]), extra

    def _pull(self):
        """slot 0 of the read-back -> self.board / self.finger (format conversion of ONE record: the bit planes are
        spread with np.unpackbits; layout.unpack_state is the batched equivalent and the tests compare the two)"""
        v = self._venv
        cells = self.dim_row * self.dim_col
        from .. import layout
        pb = 4 * layout.plane_words(self.dim_row, self.dim_col)   # bytes per plane
        bits = np.unpackbits(self._io_np[:pb * v.orb_bits].reshape(v.orb_bits, pb), axis=1, bitorder='little')[:, :cells]
        codes = bits[0].astype(int)
        for j in range(1, v.orb_bits):
            codes += bits[j].astype(int) << j
        codes[codes == (1 << v.orb_bits) - 1] = -1
        self.board = codes.reshape(self.dim_row, self.dim_col)
        f = int(self._io_np[pb * v.orb_bits]) & 63            # meta word: finger cell in bits 0..5
        self.finger = np.array([f // self.dim_col, f % self.dim_col])
        self._dev_shadow = (self.board.copy(), self.finger.copy(), self._dev_shadow[2] if self._dev_shadow else 4)

    def _push(self, board=None):
        """general path for the per-stage methods: (board or self.board, self.finger) -> slot 0 via ppad_pack"""
        self._fit_codes(self.board if board is None else board)
        flags = self._venv.set_state(np.asarray(self.board if board is None else board)[None],
                                     np.asarray(self.finger).reshape(1, 2), prev=np.array([self._prev_id()], np.uint8))
        self._dev_shadow = None
        if int(flags[0]) & _cabi.FLAG_UNSUPPORTED_ORB:
            raise NotImplementedError('orb types are -1 (empty) and 0..9 (game.py:236-246)')

    # ------------------------------------------------------------------ game.py:122-155
    def apply_action(self, action):
        """Swap the finger orb with its neighbour in direction `action`; 'rejected' when that leaves the board."""
        if action not in ('up', 'down', 'left', 'right'):
            return 'accepted'      # the reference falls through its if/elif chain without doing anything
        self._sync_to_device()
        _, info, _, _ = self._launch(0, ACTION_IDS[action])
        if (info >> 24) & _cabi.FLAG_REJECTED:
            return 'rejected'
        self._pull()
        self._dev_shadow = (self._dev_shadow[0], self._dev_shadow[1], ACTION_IDS[action])   # the kernel set prev = action
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
        self._push(board)
        self._venv.drop()
        board[...] = self._venv.get_state().boards[0].cpu().numpy()
        return board

    def visualize(self, filename, shrink=3, animate=True):
        raise NotImplementedError('GIF rendering (game.py:223-230) is out of scope of ppad_b200')

    # ------------------------------------------------------------------ game.py:232-256
    def fill_board(self, board=None, reset=False):
        """Fill empty cells (all cells when reset) row-major with skyfall orbs; in place."""
        if board is None:
            board = self.board
        self._push(board)
        if self.rng == 'reference':
            need = board.size if reset else int((np.asarray(board) == -1).sum())
            orbs = np.array([self._draw_orb() for _ in range(need)], np.uint8).reshape(1, -1)
            if need:
                self._venv.fill_board(reset=reset, injected=orbs, step=self._next_tick())
        else:
            self._venv.fill_board(reset=reset, step=self._next_tick())
        board[...] = self._venv.get_state().boards[0].cpu().numpy()
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
        if board is not None:
            self._fit_codes(board)      # poison / mortal poison / bomb orbs: the 4-bit kernels from here on

        self._reseed()
        if self.rng == 'reference':
            cells = self.dim_row * self.dim_col
            if board is None:
                def attempt(orbs):
                    # the kernel rejection-samples over the pre-drawn attempts; which one it kept tells the draw count
                    flags = self._venv.reset(fingers=np.zeros((1, 2), int), injected=orbs[None], step=self._next_tick())
                    if int(flags[0]) & _cabi.FLAG_SKYFALL_EXHAUSTED:
                        return None
                    kept = self._venv.get_state().boards[0].cpu().numpy().reshape(-1)
                    for a in range(len(orbs) // cells):
                        if np.array_equal(orbs[a * cells:(a + 1) * cells], kept):
                            self.board = kept.reshape(self.dim_row, self.dim_col).astype(int)
                            return (a + 1) * cells
                    return None
                self._speculate(attempt, cells, first=6)
            else:
                self.board = np.copy(board)
            if finger is None:
                self.finger = np.array([random.randint(0, self.dim_row - 1), random.randint(0, self.dim_col - 1)])
            else:
                self.finger = np.copy(finger)
            self.action_space.finger = self.finger
            self._dev_shadow = None
            self._record_observation()
            return self.state()
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
        self._dev_shadow = None

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
        skyfall_damage).  One fused kernel launch on the scratch record; the env's own record is not touched."""
        if verbose is True:
            return self._calculate_reward_staged(board, skyfall_damage)
        self._write_slot(1, self.board if board is None else board, self.finger, self._prev_id())
        if self.rng == 'reference' and skyfall_damage:
            result = []

            def cascade(orbs):
                self._write_slot(1, self.board if board is None else board, self.finger, self._prev_id())
                r = self._launch(1, _cabi.PASS, skyfall_damage=True, final=verbose, log_cap=64 if verbose else 0,
                                 injected=orbs)
                if (r[1] >> 24) & _cabi.FLAG_SKYFALL_EXHAUSTED or self._last_consumed == 0xFFFF:
                    return None
                result[:] = r
                return self._last_consumed       # the unsaturated count (the info byte stops at 255)
            self._speculate(cascade, 32, first=2)
            reward, info, _, extra = result
        else:
            reward, info, _, extra = self._launch(1, _cabi.PASS, skyfall_damage=skyfall_damage, final=verbose,
                                                  log_cap=64 if verbose else 0)
        return reward

    def _calculate_reward_staged(self, board, skyfall_damage):
        """calculate_reward(verbose=True): the reference prints the board at every stage of every cascade iteration
        (game.py:372-403), so this path runs the stages one kernel at a time -- cancel, drop, fill_board -- instead of
        the fused step.  The skyfall continues at the draw index the previous iteration stopped at, at ONE step index:
        the result is the one the fused kernel gives (verbose does not change the outcome in the reference either)."""
        from . import utils as pad_utils
        board = np.copy(self.board if board is None else board)
        tick = self._next_tick()
        drawn = 0
        all_combos = []
        while True:
            print('Board before combo canceling:')
            self.render(board)
            combos = pad_utils.cancel(board)
            print('Board after combo canceling:')
            self.render(board)
            if len(combos) < 1:
                break
            all_combos += combos
            if skyfall_damage is False:
                break
            board = self.drop(board=board)
            print('Board after orb drop:')
            self.render(board)
            need = int((np.asarray(board) == -1).sum())
            if self.rng == 'reference':
                board = self.fill_board(board=board)
            else:
                self._push(board)
                self._venv.fill_board(reset=False, draw0=drawn, step=tick)
                board[...] = self._venv.get_state().boards[0].cpu().numpy()
            drawn += need
            print('Board after fill board:')
            self.render(board)
        reward = self.damage(all_combos)
        print(all_combos)
        print('The total damange is:', reward)
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
