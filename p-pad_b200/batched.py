"""BatchedPAD -- n independent Puzzle & Dragons boards resident on one GPU.

The tensor-level API over the C ABI (include/ppad_b200.h).  PyTorch is plumbing here: it owns the
device memory and the stream; every operation is one hand-written sm_100a kernel launched through
ctypes.  The semantics per board are those of the reference's ``PAD`` object
(ppad/pad/game.py:28-470); the single-env drop-in facade lives in ``ppad_b200.pad.game``.
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _cabi
from .config import make_config
from .layout import state_words, decode_compact, chunk_bytes_on_demand, TILE

ACTION_NAMES = ['up', 'down', 'left', 'right', 'pass']        # player.py:24
ACTION_IDS = {name: i for i, name in enumerate(ACTION_NAMES)}  # simulation.py:236


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BatchedPAD:
    """n boards of dim_row x dim_col on ``device``; keyword arguments as ``PAD.__init__`` (game.py:39-50).

    ``env0`` is the global id of board 0 (sharding: rank * n); Philox streams are keyed by
    (seed, env0 + i, step), so a rollout does not depend on how boards are split across GPUs.
    """

    def __init__(self, n, dim_row=5, dim_col=6, skyfall=None, skyfall_damage=True, team=None, seed=0,
                 device=None, env0=0, cascade_cap=255, reset_cap=64, orb_bits=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ppad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _cabi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = int(n)
        self.dim_row, self.dim_col = int(dim_row), int(dim_col)
        self.cells = self.dim_row * self.dim_col
        self.skyfall_damage = bool(skyfall_damage)
        self.env0 = int(env0)
        self.step_index = 0
        self.cfg = make_config(dim_row, dim_col, skyfall, skyfall_damage, team, seed, cascade_cap, reset_cap, orb_bits)
        self.orb_bits = int(self.cfg.orb_bits)   # 3: orb types 0..6; 4: all ten orb types (game.py:236-246)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):      # the library selects its device; the guard restores torch's
            _cabi.check(self.lib.ppad_create(C.byref(self.cfg), self.device.index, C.byref(handle)))
        self._ctx = handle
        self.words = state_words(self.dim_row, self.dim_col, self.orb_bits)
        assert self.lib.ppad_state_bytes_ex(self.dim_row, self.dim_col, self.orb_bits) == 4 * self.words
        # packed state records (int32 view of the uint32 words of include/ppad_b200.h)
        self.state = torch.zeros((self.n, self.words), dtype=torch.int32, device=self.device)

    def like(self, n, env0=None):
        """A fresh slab of n boards with the same game configuration (used for search expansions)."""
        other = BatchedPAD.__new__(BatchedPAD)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("state", "_ctx")})
        other.n = int(n)
        other.env0 = self.env0 if env0 is None else int(env0)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_create(C.byref(self.cfg), self.device.index, C.byref(handle)))
        other._ctx = handle
        other.state = torch.zeros((other.n, self.words), dtype=torch.int32, device=self.device)
        return other

    def __del__(self):
        ctx, self._ctx = getattr(self, "_ctx", None), None
        if ctx:
            self.lib.ppad_destroy(ctx)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, x, dtype):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(self.device)

    def _slab(self, injected, fmt='nibbles'):
        """injected: None, or a uint8 [n, S] orb sequence (tensor/array), or (slab_int32 [n, W], n_inj) already packed.
        fmt: 'nibbles' (PPAD_SLAB_NIBBLES, 16 B per 32 orbs), 'planes3' (PPAD_SLAB_PLANES3, 12 B per 32 orbs) or
        'chunks10' (PPAD_SLAB_CHUNKS10, chunk-major [W, n]: 4 B per 10 orbs, read on demand; step() only)."""
        if injected is None:
            return None, 0, 0
        if isinstance(injected, tuple):
            slab, n_inj = injected[0], injected[1]
            slab = self._dev(slab, torch.int16 if fmt in ('chunks6', 'chunks6t') else torch.int32)
            return slab, slab.shape[1 if fmt == 'chunks6t' else (0 if fmt in _cabi.CHUNK_MAJOR else 1)], int(n_inj)
        orbs = self._dev(injected, torch.uint8)
        n, s = orbs.shape
        if fmt in ('chunks6', 'chunks6t'):
            chunks = max(1, (s + 5) // 6)
            pad = torch.zeros((n, chunks * 6), dtype=torch.int64, device=self.device)
            pad[:, :s] = orbs
            word = (pad.view(n, chunks, 6) * (6 ** torch.arange(6, device=self.device, dtype=torch.int64))).sum(dim=2)
            word = torch.where(word >= 2 ** 15, word - 2 ** 16, word).to(torch.int16)   # uint16 bit pattern in an int16 tensor
            if fmt == 'chunks6t':                          # [tiles, chunks, 256]
                tiles = max(1, (n + 255) // 256)
                pad2 = torch.zeros((tiles * 256, chunks), dtype=torch.int16, device=self.device)
                pad2[:n] = word
                return pad2.view(tiles, 256, chunks).permute(0, 2, 1).contiguous(), chunks, s
            return word.t().contiguous(), chunks, s
        if fmt == 'chunks10':
            chunks = max(1, (s + 9) // 10)
            pad = torch.zeros((n, chunks * 10), dtype=torch.int64, device=self.device)
            pad[:, :s] = orbs & 7
            pad = pad.view(n, chunks, 10)
            k = torch.arange(10, device=self.device, dtype=torch.int64)
            word = sum(((((pad >> j) & 1) << (k + 10 * j)).sum(dim=2)) for j in range(3))   # [n, chunks], < 2^30
            return word.to(torch.int32).t().contiguous(), chunks, s
        if fmt == 'planes3':
            groups = max(1, (s + 31) // 32)
            pad = torch.zeros((n, groups * 32), dtype=torch.int64, device=self.device)
            pad[:, :s] = orbs & 7
            pad = pad.view(n, groups, 32)
            shifts = torch.arange(32, device=self.device, dtype=torch.int64)
            planes = torch.stack([(((pad >> j) & 1) << shifts).sum(dim=2) for j in range(3)], dim=2)   # [n, groups, 3]
            slab = torch.where(planes >= 2 ** 31, planes - 2 ** 32, planes).to(torch.int32).reshape(n, groups * 3)
            return slab.contiguous(), groups * 3, s
        words = (s + 7) // 8
        pad = torch.zeros((n, words * 8), dtype=torch.int64, device=self.device)
        pad[:, :s] = orbs & 15
        shifts = 4 * torch.arange(8, device=self.device, dtype=torch.int64)
        packed = (pad.view(n, words, 8) << shifts).sum(dim=2)
        slab = (packed & 0xFFFFFFFF).to(torch.int64)
        slab = torch.where(slab >= 2 ** 31, slab - 2 ** 32, slab).to(torch.int32).contiguous()
        return slab, words, s

    # ------------------------------------------------------------------ state in / out
    def set_state(self, boards, fingers, prev=None, moves=None):
        """PAD.board / PAD.finger / previous_action (game.py:88,94,117) -> packed state. Returns flags [n] uint8."""
        boards = self._dev(boards, torch.int64).reshape(self.n, self.cells)
        fingers = self._dev(fingers, torch.int64).reshape(self.n, 2)
        prev = self._dev(prev, torch.uint8)
        moves = None if moves is None else self._dev(moves, torch.int16)
        flags = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_pack(self._ctx, self.n, _ptr(boards), _ptr(fingers), _cabi.I64, _ptr(prev),
                                           _ptr(moves), _ptr(self.state), _ptr(flags), self._stream()))
        return flags

    def get_state(self, state=None):
        """-> SimpleNamespace(boards int64 [n, rows, cols], fingers int64 [n, 2], prev uint8 [n], moves int16 [n])"""
        state = self.state if state is None else state
        n = state.shape[0]
        boards = torch.empty((n, self.dim_row, self.dim_col), dtype=torch.int64, device=self.device)
        fingers = torch.empty((n, 2), dtype=torch.int64, device=self.device)
        prev = torch.empty(n, dtype=torch.uint8, device=self.device)
        moves = torch.empty(n, dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_unpack(self._ctx, n, _ptr(state), _ptr(boards), _ptr(fingers), _cabi.I64,
                                             _ptr(prev), _ptr(moves), self._stream()))
        return SimpleNamespace(boards=boards, fingers=fingers, prev=prev, moves=moves)

    def snapshot(self):
        """Checkpoint of the env slab (the batched analogue of the record PAD.backtrack pops, game.py:407-427)."""
        return self.state.clone(), self.step_index

    def restore(self, snap):
        self.state.copy_(snap[0])
        self.step_index = snap[1]

    # ------------------------------------------------------------------ reset / step
    def reset(self, boards=None, fingers=None, injected=None, step=None):
        """PAD.reset (game.py:282-323) for every board; returns flags [n] uint8.

        boards given: adopt them (fingers default to Philox draws unless given).
        boards None : rejection-sample combo-free boards (injected orbs first, then Philox)."""
        step = self._next_step(step)
        if boards is not None:
            if fingers is None:
                # random fingers from the Philox FINGER stream: do a plain reset, keep only its fingers
                self.step_index = step
                self.reset(step=step)
                fingers = self.get_state().fingers
            return self.set_state(boards, fingers)
        slab, words, n_inj = self._slab(injected)
        fin = None if fingers is None else self._dev(fingers, torch.int32).reshape(self.n, 2)
        flags = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self.last_reset_consumed = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_reset(self._ctx, self.n, self.env0, step, _ptr(slab), words, n_inj, _ptr(fin),
                                            _ptr(self.state), _ptr(flags), _ptr(self.last_reset_consumed),
                                            self._stream()))
        return flags

    def _next_step(self, step):
        if step is None:
            step = self.step_index
        self.step_index = int(step) + 1
        return int(step) & 0xFFFFFFFF

    def step(self, actions=None, eval_moves=False, auto_reset=False, max_moves=0, injected=None, obs=None,
             want_final=False, log_cap=0, step=None, out=None, skyfall_damage=None, paths=None, path_lens=None,
             slab_format='nibbles', want_consumed=False):
        """One env step of all boards (ppad_step).  actions: None (random legal moves / forced pass at
        max_moves), or uint8 [n] ids 0..4 (255 = sample).  obs: None, 'f32' or 'u8', or a preallocated tensor.
        injected: see _slab (``slab_format`` names the layout of an already packed slab / the layout to pack into).
        Returns SimpleNamespace(action, reward, info, obs, final_state, combo_log, consumed); ``out`` reuses buffers."""
        step = self._next_step(step)
        n = self.n
        out = out if out is not None else SimpleNamespace()
        if getattr(out, "reward", None) is None:
            out.action = torch.empty(n, dtype=torch.uint8, device=self.device)
            out.reward = torch.empty(n, dtype=torch.float64, device=self.device)
            out.info = torch.empty(n, dtype=torch.int32, device=self.device)
        obs_dtype = _cabi.OBS_F32
        if isinstance(obs, torch.Tensor):
            out.obs = obs
            obs_dtype = _cabi.OBS_F32 if obs.dtype == torch.float32 else _cabi.OBS_U8
        elif obs is not None:
            obs_dtype = {'f32': _cabi.OBS_F32, 'u8': _cabi.OBS_U8}[obs]
            want = torch.float32 if obs == 'f32' else torch.uint8
            if getattr(out, "obs", None) is None or out.obs.dtype != want:
                out.obs = torch.empty((n, self.dim_row, self.dim_col, 7), dtype=want, device=self.device)
        else:
            out.obs = None
        out.final_state = torch.empty_like(self.state) if want_final else None
        out.combo_log = torch.zeros((n, log_cap), dtype=torch.int16, device=self.device) if log_cap else None
        out.consumed = torch.empty(n, dtype=torch.int16, device=self.device) if want_consumed else None
        acts = self._dev(actions, torch.uint8)
        slab, words, n_inj = self._slab(injected, slab_format)
        a = _cabi.StepArgs()
        a.slab_format = _cabi.SLAB_FORMATS[slab_format]
        a.slab_stride = slab.shape[-1] if slab is not None and slab_format in _cabi.CHUNK_MAJOR else 0
        a.consumed_out = out.consumed.data_ptr() if want_consumed else None
        a.n, a.env0, a.step = n, self.env0, step
        a.eval_moves, a.auto_reset, a.max_moves = int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves)
        a.state = self.state.data_ptr()
        a.actions = acts.data_ptr() if acts is not None else None
        a.slab = slab.data_ptr() if slab is not None else None
        a.slab_words, a.n_inj = words, n_inj
        a.action_out, a.reward, a.info = out.action.data_ptr(), out.reward.data_ptr(), out.info.data_ptr()
        a.final_state = out.final_state.data_ptr() if want_final else None
        a.combo_log = out.combo_log.data_ptr() if log_cap else None
        a.log_cap = int(log_cap)
        a.obs = out.obs.data_ptr() if out.obs is not None else None
        a.obs_dtype = obs_dtype
        a.skyfall_damage = -1 if skyfall_damage is None else int(bool(skyfall_damage))
        if paths is not None:
            # whole finger paths: (packed int32 [n, W], L) or a uint8 [n, L] array of move ids 0..3
            if isinstance(paths, tuple):
                packed, plen = self._dev(paths[0], torch.int32), int(paths[1])
            else:
                packed, plen = self.pack_paths(paths)
            lens = None if path_lens is None else self._dev(path_lens, torch.int16)
            a.paths, a.path_words, a.path_len = packed.data_ptr(), packed.shape[1], plen
            a.path_lens = lens.data_ptr() if lens is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_step(self._ctx, C.byref(a), self._stream()))
        return out

    def pack_paths(self, moves):
        """uint8 [n, L] move ids 0..3 -> (int32 [n, ceil(L / 16)] with 2 bits per move, L)"""
        m = self._dev(moves, torch.uint8)
        n, length = m.shape
        words = max(1, (length + 15) // 16)
        pad = torch.zeros((n, words * 16), dtype=torch.int64, device=self.device)
        pad[:, :length] = m & 3
        packed = (pad.view(n, words, 16) << (2 * torch.arange(16, device=self.device, dtype=torch.int64))).sum(dim=2)
        packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).to(torch.int32).contiguous()
        return packed, length

    def rollout(self, steps, eval_moves=False, auto_reset=True, max_moves=50, step=None, record=False):
        """`steps` env steps of the in-kernel random policy in ONE launch (ppad_rollout): equals that many
        step(None, ...) calls.  Returns SimpleNamespace(reward_sum float64 [n], episodes int32 [n],
        combos_sum int32 [n], flags_or uint8 [n]); step_index advances by `steps`.  record=True adds the trajectory:
        states int32 [steps, n, words] (the state after every step) and actions uint8 [steps, n], time-major."""
        step0 = self.step_index if step is None else int(step)
        self.step_index = step0 + int(steps)
        out = SimpleNamespace(reward_sum=torch.empty(self.n, dtype=torch.float64, device=self.device),
                              episodes=torch.empty(self.n, dtype=torch.int32, device=self.device),
                              combos_sum=torch.empty(self.n, dtype=torch.int32, device=self.device),
                              flags_or=torch.empty(self.n, dtype=torch.uint8, device=self.device),
                              states=None, actions=None)
        if record:
            out.states = torch.empty((int(steps), self.n, self.words), dtype=torch.int32, device=self.device)
            out.actions = torch.empty((int(steps), self.n), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_rollout_record(self._ctx, self.n, self.env0, step0 & 0xFFFFFFFF, int(steps),
                                                     int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves),
                                                     _ptr(self.state), _ptr(out.reward_sum), _ptr(out.episodes),
                                                     _ptr(out.combos_sum), _ptr(out.flags_or), _ptr(out.states),
                                                     _ptr(out.actions), self._stream()))
        return out

    def calculate_reward(self, injected=None, want_final=False, log_cap=0, step=None, skyfall_damage=None):
        """PAD.calculate_reward (game.py:364-405) of every board, on a copy: the state is not modified."""
        acts = torch.full((self.n,), _cabi.PASS, dtype=torch.uint8, device=self.device)
        return self.step(acts, injected=injected, want_final=want_final, log_cap=log_cap, step=step,
                         skyfall_damage=skyfall_damage)

    def sample_actions(self, include_pass=False, step=None):
        """PlayerAction.sample (player.py:38-66) for every board -> uint8 [n] (255 = no legal move).
        Uses the draw the in-kernel policy of step(actions=None) would use at the same step index and does
        not advance step_index."""
        step = self.step_index if step is None else int(step)
        actions = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_sample_actions(self._ctx, self.n, self.env0, step & 0xFFFFFFFF, _ptr(self.state),
                                                     int(bool(include_pass)), _ptr(actions), self._stream()))
        return actions

    def drop(self):
        """PAD.drop (game.py:201-221) in place"""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_drop(self._ctx, self.n, _ptr(self.state), self._stream()))

    def fill_board(self, reset=False, injected=None, draw0=0, step=None):
        """PAD.fill_board (game.py:232-256) in place -> (flags uint8 [n], consumed uint8 [n])"""
        step = self.step_index if step is None else int(step)
        slab, words, n_inj = self._slab(injected)
        flags = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        consumed = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_fill(self._ctx, self.n, self.env0, step & 0xFFFFFFFF, int(draw0), int(bool(reset)),
                                           _ptr(slab), words, n_inj, _ptr(self.state), _ptr(flags), _ptr(consumed),
                                           self._stream()))
        return flags, consumed

    def damage(self, combo_log, n_combos):
        """PAD.damage (game.py:157-199): combo_log int16/uint16 [n, cap] (colour | N << 8), n_combos uint8 [n]"""
        combo_log = self._dev(combo_log, torch.int16)
        n_combos = self._dev(n_combos, torch.uint8)
        n = combo_log.shape[0]
        reward = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_damage(self._ctx, n, _ptr(combo_log), combo_log.shape[1], _ptr(n_combos),
                                             _ptr(reward), self._stream()))
        return reward

    # ------------------------------------------------------------------ single-purpose kernels
    def encode_obs(self, dtype='f32', state=None, out=None):
        """Agent01.board_finger_to_state (agent01.py:254-273): [n, rows, cols, 7]"""
        state = self.state if state is None else state
        n = state.shape[0]
        tdt = torch.float32 if dtype == 'f32' else torch.uint8
        if out is None:
            out = torch.empty((n, self.dim_row, self.dim_col, 7), dtype=tdt, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_encode_obs(self._ctx, n, _ptr(state), _ptr(out),
                                                 _cabi.OBS_F32 if dtype == 'f32' else _cabi.OBS_U8, self._stream()))
        return out

    def cancel(self, log_cap=16):
        """pad_utils.cancel (utils.py:26-70) in place on every board -> (n_combos uint8 [n], combo_log int16 [n, log_cap])"""
        n_combos = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        log = torch.zeros((self.n, log_cap), dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_cancel(self._ctx, self.n, _ptr(self.state), _ptr(n_combos), _ptr(log), log_cap,
                                             self._stream()))
        return n_combos, log

    def islands(self, seeds, orb_types, island_in=None):
        """pad_utils.detect_island (utils.py:86-111) for every board: seeds int [n, 2] (row, col), orb_types int [n]
        (-1 = empty cells), island_in int64 [n] bit masks of cells already marked or None -> int64 [n] bit masks (bit = cell)."""
        seeds = self._dev(seeds, torch.int64).reshape(self.n, 2)
        cell = (seeds[:, 0] * self.dim_col + seeds[:, 1]).clamp(0, 255).to(torch.uint8).contiguous()
        off = (seeds[:, 0] < 0) | (seeds[:, 0] >= self.dim_row) | (seeds[:, 1] < 0) | (seeds[:, 1] >= self.dim_col)
        cell = torch.where(off, torch.full_like(cell, 255), cell)
        types = self._dev(orb_types, torch.int8).reshape(self.n).contiguous()
        pre = None if island_in is None else self._dev(island_in, torch.int64).reshape(self.n).contiguous()
        out = torch.empty(self.n, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_island(self._ctx, self.n, _ptr(self.state), _ptr(cell), _ptr(types), _ptr(pre),
                                             _ptr(out), self._stream()))
        return out

    def legal_mask(self):
        """uint8 [n]: bits 0..3 = legal non-reversing moves (player.py:52-62), bits 4..7 = off-board moves (utils.py:73-83)"""
        mask = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_legal_mask(self._ctx, self.n, _ptr(self.state), _ptr(mask), self._stream()))
        return mask

    def permute_colours(self, maps):
        """permutation_mapping (smart_data.py:94-108) in place: maps uint8 [n, <= 8], orb type t -> maps[i][t]"""
        m = self._dev(maps, torch.uint8)
        full = torch.arange(8, dtype=torch.uint8, device=self.device).repeat(self.n, 1)
        full[:, :m.shape[1]] = m
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_permute(self._ctx, self.n, _ptr(self.state), _ptr(full.contiguous()), self._stream()))

    # ------------------------------------------------------------------ callers either side of the step
    def select_actions(self, q, method='max', beta=None, epsilon=0.0, max_moves=0, step=None):
        """Agent01.act after the network (agent01.py:173-218): q float32 [n, 5] -> uint8 [n] action ids.
        Uses the Philox POLICY stream of the step index the NEXT step() call will use; does not advance it."""
        method = str(method).lower()
        if method not in ('max', 'boltzmann'):
            raise Exception('ERROR: Unknown action method!')
        if method == 'boltzmann' and beta is None:
            raise Exception('ERROR: Provide a value for beta!')
        step = self.step_index if step is None else int(step)
        q = self._dev(q, torch.float32).reshape(self.n, 5)
        actions = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_select_actions(
                self._ctx, self.n, self.env0, step & 0xFFFFFFFF, _ptr(self.state), _ptr(q),
                _cabi.SELECT_MAX if method == 'max' else _cabi.SELECT_BOLTZMANN, float(beta or 0.0), float(epsilon),
                int(max_moves), _ptr(actions), self._stream()))
        return actions

    def policy_step(self, q, method='boltzmann', beta=None, epsilon=0.0, max_moves=0, eval_moves=False, auto_reset=True,
                    injected=None, slab_format='nibbles', obs=None, ring=None, stats=None, step=None, out=None,
                    ctas_per_sm=0):
        """The loop body of model_simulation (simulation.py:34-68) for every board in ONE kernel (ppad_policy_step):
        action selection from q float32 [n, 5] (agent01.py:173-218) with the episode length cap, the env step with
        auto-reset, the one-hot observation of the new state, the push of (state before the step, action, reward /
        discounted return) into ``ring`` (a replay.RolloutRing) and the episode statistics into ``stats`` (a
        replay.RolloutStats).  Returns SimpleNamespace(action, reward, info, obs); ``out`` reuses the buffers."""
        method = str(method).lower()
        if method not in ('max', 'boltzmann'):
            raise Exception('ERROR: Unknown action method!')
        if method == 'boltzmann' and beta is None:
            raise Exception('ERROR: Provide a value for beta!')
        step = self._next_step(step)
        n = self.n
        out = out if out is not None else SimpleNamespace()
        if getattr(out, "reward", None) is None:
            out.action = torch.empty(n, dtype=torch.uint8, device=self.device)
            out.reward = torch.empty(n, dtype=torch.float64, device=self.device)
            out.info = torch.empty(n, dtype=torch.int32, device=self.device)
        obs_dtype = _cabi.OBS_F32
        if isinstance(obs, torch.Tensor):
            out.obs = obs
            obs_dtype = _cabi.OBS_F32 if obs.dtype == torch.float32 else _cabi.OBS_U8
        elif obs is not None:
            obs_dtype = {'f32': _cabi.OBS_F32, 'u8': _cabi.OBS_U8}[obs]
            want = torch.float32 if obs == 'f32' else torch.uint8
            if getattr(out, "obs", None) is None or out.obs.dtype != want:
                out.obs = torch.empty((n, self.dim_row, self.dim_col, 7), dtype=want, device=self.device)
        else:
            out.obs = None
        q = self._dev(q, torch.float32).reshape(n, 5)
        slab, words, n_inj = self._slab(injected, slab_format)
        a = _cabi.PolicyArgs()
        a.n, a.env0, a.step = n, self.env0, step
        a.eval_moves, a.auto_reset, a.max_moves = int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves)
        a.state, a.q = self.state.data_ptr(), q.data_ptr()
        a.method = _cabi.SELECT_MAX if method == 'max' else _cabi.SELECT_BOLTZMANN
        a.beta, a.epsilon = float(beta or 0.0), float(epsilon)
        a.slab = slab.data_ptr() if slab is not None else None
        a.slab_words, a.n_inj = words, n_inj
        a.slab_format = {'nibbles': _cabi.SLAB_NIBBLES, 'planes3': _cabi.SLAB_PLANES3}[slab_format]   # no chunks10 here
        a.obs_dtype = obs_dtype
        a.action_out, a.reward, a.info = out.action.data_ptr(), out.reward.data_ptr(), out.info.data_ptr()
        a.obs = out.obs.data_ptr() if out.obs is not None else None
        if ring is not None:
            a.ring_state, a.ring_action, a.ring_reward = ring.state.data_ptr(), ring.action.data_ptr(), ring.reward.data_ptr()
            a.ring_steps, a.ring_step = ring.steps, ring.pushed
            a.gamma, a.log10_reward = float(ring.gamma), int(bool(ring.log10))
        a.ctas_per_sm = int(ctas_per_sm)
        a.stats_rows = stats.rows.data_ptr() if stats is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_policy_step(self._ctx, C.byref(a), self._stream()))
        if ring is not None:
            ring.pushed += 1
        return out

    def discount(self, rewards, gamma, lengths=None, norm=False, log10=True):
        """ppad.discount (agent/utils.py:22-51) for a batch of trajectories: rewards float64 [m, T] (a copy is
        discounted and returned), lengths int32 [m] or None."""
        r = self._dev(rewards, torch.float64).clone()
        if r.dim() == 1:
            r = r[None]
        lengths = None if lengths is None else self._dev(lengths, torch.int32)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.ppad_discount(self._ctx, r.shape[0], r.shape[1], _ptr(r), _ptr(lengths), float(gamma),
                                               int(bool(log10)), int(bool(norm)), self._stream()))
        return r

    def host_pipe(self, slab_orbs=32, chunks=2, zero_copy=True, depth=1, compact=False, slab_format='nibbles',
                  slab_eager=0):
        """A HostPipe for this slab: env steps whose inputs and results are pinned HOST tensors (ppad_pipe_*)."""
        return HostPipe(self, slab_orbs, chunks, zero_copy, depth, compact, slab_format, slab_eager)

    @staticmethod
    def split_info(info):
        """info int32 [n] -> (n_combos, depth, consumed, flags) int32 tensors"""
        u = info.to(torch.int64) & 0xFFFFFFFF
        return u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF, (u >> 24) & 0xFF


class HostPipe:
    """Env steps for callers whose inputs and results live in host memory (the reference's situation:
    PAD.step takes a Python value and appends to Python lists, game.py:329-362).

    zero_copy=True (default): the kernel reads the skyfall slab from and writes its results to the pinned host buffers
    directly (unified addressing), so the PCIe traffic overlaps the kernel store by store.  zero_copy='hybrid': only
    the results are written in place; the slab and the actions travel by DMA, one of `chunks` at a time, so that no
    PCIe read sits inside a kernel (hosts with a long read latency: 0.75 instead of 1.1 ms per 2^20-board step on
    one box of the pool; 2-3 % slower than full zero copy elsewhere, with 4 or more chunks).
    zero_copy=False: the batch is cut into `chunks`; each chunk copies its inputs host -> device, runs the fused
    step kernel and copies its results device -> host on its own stream.
    The env state and the one-hot observation stay in HBM.  Buffers are pinned and owned by the pipe; with
    depth = 2 there are two sets (``buffers[0]``, ``buffers[1]``) so that step t + 1 can be enqueued before the
    results of step t are read (steps of one pipe always execute in order):
      in : slab int32 [n, slab_orbs / 8] (nibble-packed skyfall, layout.pack_slab), actions uint8 [n]
      out: action uint8 [n], reward float64 [n], info int32 [n], state int32 [n, words]
    ``pipe.slab`` etc. are the buffers of set 0.
    """

    def __init__(self, venv, slab_orbs=32, chunks=2, zero_copy=True, depth=1, compact=False, slab_format='nibbles',
                 slab_eager=0):
        self.venv = venv
        self.slab_eager = int(slab_eager)   # chunks10: chunks per board the kernel fetches up front (0 = default, 2)
        self._dict = None                   # (rewards float64 [D], infos uint32 [D]) of the dictionary-coded stream
        self.slab_orbs = int(slab_orbs)
        self.slab_format = slab_format
        self.slab_words = {'planes3': 3 * ((self.slab_orbs + 31) // 32), 'chunks10': (self.slab_orbs + 9) // 10,
                           'chunks6': (self.slab_orbs + 5) // 6, 'chunks6t': (self.slab_orbs + 5) // 6,
                           'nibbles': (self.slab_orbs + 7) // 8}[slab_format]
        self.compact = bool(compact)
        n = venv.n
        self.tiles = (n + TILE - 1) // TILE
        pin = dict(pin_memory=True)
        cap = int(venv.lib.ppad_compact_capacity(venv._ctx, n)) if self.compact else 0
        self.buffers = [SimpleNamespace(
            # chunks10 / chunks6 are chunk-major ([W, n], layout.pack_slab_chunks10 / _chunks6); the row formats are [n, W]
            slab=torch.zeros(((n + 255) // 256, max(1, self.slab_words), 256) if slab_format == 'chunks6t' else
                             ((max(1, self.slab_words), n) if slab_format in _cabi.CHUNK_MAJOR else (n, max(1, self.slab_words))),
                             dtype=torch.int16 if slab_format in ('chunks6', 'chunks6t') else torch.int32, **pin),
            actions=torch.zeros(n, dtype=torch.uint8, **pin),
            action=torch.zeros(n, dtype=torch.uint8, **pin),
            reward=torch.zeros(n, dtype=torch.float64, **pin),
            info=torch.zeros(n, dtype=torch.int32, **pin),
            state=torch.zeros((n, venv.words), dtype=torch.int32, **pin),
            # compact result stream (include/ppad_b200.h: ppad_host_step_args)
            nibbles=torch.zeros(self.tiles * 32 if self.compact else 1, dtype=torch.int32, **pin),
            tile_line=torch.zeros(self.tiles if self.compact else 1, dtype=torch.int32, **pin),
            records=torch.zeros(max(cap, 128), dtype=torch.uint8, **pin),
            summary=torch.zeros(4, dtype=torch.int32, **pin)) for _ in range(2 if depth > 1 else 1)]
        for name, value in vars(self.buffers[0]).items():
            setattr(self, name, value)
        handle = C.c_void_p()
        with torch.cuda.device(venv.device):
            _cabi.check(venv.lib.ppad_pipe_create(venv._ctx, n, self.slab_words, int(chunks), C.byref(handle)))
        self._pipe = handle
        self.mode = {True: 'zero_copy', False: 'staged'}.get(zero_copy, zero_copy)
        _cabi.check(venv.lib.ppad_pipe_set_zero_copy(handle, 2 if zero_copy == 'hybrid' else int(bool(zero_copy))))

    def __del__(self):
        pipe, self._pipe = getattr(self, "_pipe", None), None
        if pipe:
            self.venv.lib.ppad_pipe_destroy(pipe)

    def step(self, use_actions=False, use_slab=True, eval_moves=False, auto_reset=False, max_moves=0, obs=None,
             want_state=None, skyfall_damage=None, step=None, wait=True, buf=0, join=True):
        """Enqueue one step (inputs are read from buffers[buf].slab / .actions).  Dense pipes: the results land in
        .reward, .info, .action and, with want_state (default), .state.  Compact pipes: they land in .nibbles /
        .tile_line / .records / .summary -- decode them with results(buf); the state is only copied when want_state is
        set (the host can follow the board from the actions; fresh boards after a reset come with the records).
        wait=False returns right after enqueueing; call wait(buf) before reading the results or refilling that set.
        join=True (default) orders work queued later on torch's current stream behind this step; join=False leaves that
        to the caller (join() / wait()): one cross-stream dependency less per step when steps are enqueued back to back."""
        v = self.venv
        b = self.buffers[buf]
        if want_state is None:
            want_state = not self.compact
        a = _cabi.HostStepArgs()
        a.n, a.env0, a.step = v.n, v.env0, v._next_step(step)
        a.eval_moves, a.auto_reset, a.max_moves = int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves)
        a.skyfall_damage = -1 if skyfall_damage is None else int(bool(skyfall_damage))
        a.state = v.state.data_ptr()
        a.actions_host = b.actions.data_ptr() if use_actions else None
        a.slab_host = b.slab.data_ptr() if use_slab and self.slab_words else None
        a.slab_words, a.n_inj = self.slab_words, self.slab_orbs
        a.slab_format = _cabi.SLAB_FORMATS[self.slab_format]
        a.slab_stride = b.slab.shape[-1] if self.slab_format in _cabi.CHUNK_MAJOR else 0
        a.slab_eager = self.slab_eager
        if self.compact:
            a.results_compact = 2 if self._dict is not None else 1
            b.dict = self._dict                     # the stream of this step decodes with the dictionary it was coded with
            a.nibbles_host, a.tile_line_host = b.nibbles.data_ptr(), b.tile_line.data_ptr()
            a.records_host, a.records_capacity = b.records.data_ptr(), b.records.numel()
            a.summary_host = b.summary.data_ptr()
        else:
            a.action_out_host, a.reward_host, a.info_host = b.action.data_ptr(), b.reward.data_ptr(), b.info.data_ptr()
        a.state_out_host = b.state.data_ptr() if want_state else None
        a.no_join = 0 if join else 1
        if obs is not None:
            a.obs = obs.data_ptr()
            a.obs_dtype = _cabi.OBS_F32 if obs.dtype == torch.float32 else _cabi.OBS_U8
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_pipe_step(self._pipe, C.byref(a), v._stream(), int(buf)))
            if wait:
                _cabi.check(v.lib.ppad_pipe_wait(self._pipe, int(buf)))

    def results(self, buf=0):
        """Dense view of the last completed step of buffer set `buf`: dict(action uint8 [n], reward float64 [n],
        info uint32 [n]) as numpy arrays; compact pipes add reset_index / reset_state (the fresh boards of the episodes
        that ended) and `lines`.  Call after wait(buf)."""
        b = self.buffers[buf]
        if not self.compact:
            return dict(action=b.action.numpy(), reward=b.reward.numpy(), info=b.info.numpy().view(np.uint32))
        return decode_compact(self.venv.n, self.venv.words, b.nibbles.numpy().view(np.uint32),
                              b.tile_line.numpy().view(np.uint32), b.records.numpy(), b.summary.numpy().view(np.uint32),
                              dictionary=getattr(b, "dict", None))

    def set_dictionary(self, rewards=None, infos=None):
        """Dictionary of the compact stream (ppad_pipe_set_dictionary): records whose (reward, info) pair is in it travel
        as a 16-bit index.  Waits for the steps in flight; decode their results before calling.  None removes it."""
        if not self.compact:
            raise ValueError("only compact pipes take a dictionary")
        if rewards is None or len(rewards) == 0:
            self._dict = None
            r = i = None
            count = 0
        else:
            r = np.ascontiguousarray(rewards, np.float64)
            i = np.ascontiguousarray(infos, np.uint32)
            assert r.shape == i.shape and r.ndim == 1
            count = len(r)
            self._dict = (r, i)
        with torch.cuda.device(self.venv.device):
            _cabi.check(self.venv.lib.ppad_pipe_set_dictionary(
                self._pipe, r.ctypes.data if count else None, i.ctypes.data if count else None, count))

    def learn_dictionary(self, buf=0, max_entries=4096):
        """Build the dictionary from the results of the last completed step of `buf`: its most frequent (reward, info)
        pairs.  A host-side frequency count over values the device computed -- no game logic.  Returns the number of
        entries and the share of that step's records they cover."""
        got = self.results(buf)
        has = got["info"] != 0
        if not has.any():
            self.set_dictionary(None)
            return 0, 0.0
        key = np.empty(int(has.sum()), dtype=[("r", np.uint64), ("i", np.uint32)])
        key["r"] = got["reward"][has].view(np.uint64)
        key["i"] = got["info"][has]
        uniq, counts = np.unique(key, return_counts=True)
        order = np.argsort(-counts, kind="stable")[:int(max_entries)]
        self.set_dictionary(uniq["r"][order].view(np.float64), uniq["i"][order])
        return len(order), float(counts[order].sum()) / float(counts.sum())

    def d2h_bytes(self, buf=0, want_state=None):
        """Bytes the last completed step of `buf` wrote to host memory."""
        b = self.buffers[buf]
        if want_state is None:
            want_state = not self.compact
        state = b.state.numel() * 4 if want_state else 0
        if not self.compact:
            return b.reward.numel() * 8 + b.info.numel() * 4 + b.action.numel() + state
        return b.nibbles.numel() * 4 + b.tile_line.numel() * 4 + 16 + int(b.summary[0]) * 128 + state

    def h2d_bytes(self, use_actions=False):
        """Host-to-device bytes of one step.  Row slabs are read whole; of a chunks10 slab the kernel reads the eager
        chunks of every board (the further chunks a cascade reaches are not counted here: see h2d_bytes_on_demand)."""
        slab = self.slab.numel() * 4
        if self.slab_format in _cabi.CHUNK_MAJOR:
            per, size, default = _cabi.CHUNK_MAJOR[self.slab_format]
            boards = self.slab.shape[0] * 256 if self.slab_format == 'chunks6t' else self.slab.shape[1]
            slab = min(self.slab_eager or default, self.slab_words) * boards * size
        return slab + (self.actions.numel() if use_actions else 0)

    def h2d_bytes_on_demand(self, info):
        """Chunk-major slabs: bytes of the chunks beyond the eager ones that the cascades of a step reached
        (layout.chunk_bytes_on_demand: from the consumed counts in its info words, in 32-byte sectors)."""
        if self.slab_format not in _cabi.CHUNK_MAJOR:
            return 0
        per, size, default = _cabi.CHUNK_MAJOR[self.slab_format]
        return chunk_bytes_on_demand(info, per, size, min(self.slab_eager or default, self.slab_words), self.slab_words)

    def wait(self, buf=0):
        with torch.cuda.device(self.venv.device):
            _cabi.check(self.venv.lib.ppad_pipe_wait(self._pipe, int(buf)))

    def join(self, buf=0):
        """make torch's current stream wait for the step last enqueued with `buf` (e.g. before reading obs on the device)"""
        with torch.cuda.device(self.venv.device):
            _cabi.check(self.venv.lib.ppad_pipe_join(self._pipe, self.venv._stream(), int(buf)))
