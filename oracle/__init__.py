"""ctypes binding of the CPU oracle (oracle/ppad_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
legs.  Nothing under the product package imports this module (tests/test_host_logic.py::test_product_never_touches_the_oracle
greps for it).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libppad_oracle.so")
MAX_TEAM = 8

FLAG_REJECTED, FLAG_SKYFALL_EXHAUSTED, FLAG_CASCADE_CAP, FLAG_NO_LEGAL_MOVE = 0x01, 0x02, 0x04, 0x08
FLAG_BAD_ACTION, FLAG_RESET_CAP, FLAG_UNSUPPORTED_ORB, FLAG_EPISODE_DONE = 0x10, 0x20, 0x40, 0x80
STREAM_SKYFALL, STREAM_ACTION, STREAM_RESET, STREAM_FINGER, STREAM_POLICY = 0, 1, 2, 3, 4


class Cfg(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("skyfall_damage", C.c_int32),
                ("cascade_cap", C.c_int32), ("reset_cap", C.c_int32), ("team_n", C.c_int32),
                ("team_c1", C.c_int32 * MAX_TEAM), ("team_c2", C.c_int32 * MAX_TEAM),
                ("team_atk", C.c_double * MAX_TEAM), ("skyfall", C.c_double * 10), ("seed", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("reward", C.c_double), ("n_combos", C.c_int32), ("depth", C.c_int32),
                ("consumed", C.c_int32), ("flags", C.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "ppad_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "ppad_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ppad_oracle_draw.restype = C.c_uint32
        L.ppad_oracle_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ppad_oracle_damage.restype = C.c_double
        L.ppad_oracle_orb_from_u32.argtypes = [C.POINTER(C.c_double), C.c_uint32]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


_ENGLISH2INT = {"red": 0, "blue": 1, "green": 2, "light": 3, "dark": 4, "heal": 5,
                "jammer": 6, "poison": 7, "mortal poison": 8, "bomb": 9, None: -1}


def make_cfg(rows=5, cols=6, skyfall_damage=True, team=None, skyfall=None, seed=0, cascade_cap=255, reset_cap=64):
    cfg = Cfg()
    lib().ppad_oracle_default_cfg(C.byref(cfg))
    cfg.rows, cfg.cols, cfg.skyfall_damage = rows, cols, int(bool(skyfall_damage))
    cfg.cascade_cap, cfg.reset_cap, cfg.seed = cascade_cap, reset_cap, seed
    if team is not None:
        assert len(team) <= MAX_TEAM
        cfg.team_n = len(team)
        for i, u in enumerate(team):
            cfg.team_c1[i] = _ENGLISH2INT[u["color_1"]]
            cfg.team_c2[i] = _ENGLISH2INT[u["color_2"]]
            cfg.team_atk[i] = float(u["atk"])
    if skyfall is not None:
        for i in range(10):
            cfg.skyfall[i] = float(skyfall[i])
    return cfg


def philox4x32(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().ppad_oracle_philox4x32(c, k, out)
    return list(out)


def draw(seed, env, step, stream, d):
    return lib().ppad_oracle_draw(seed, env, step, stream, d)


def orb_from_u32(prob, x):
    p = (C.c_double * 10)(*[float(v) for v in prob])
    return lib().ppad_oracle_orb_from_u32(p, x)


def cancel(board):
    """utils.py:26-70 on an int8 [rows, cols] array (mutated). Returns [(colour, N), ...]."""
    assert board.dtype == np.int8 and board.flags.c_contiguous
    rows, cols = board.shape
    colors = np.zeros(64, np.int32)
    counts = np.zeros(64, np.int32)
    n = lib().ppad_oracle_cancel(_p(board, C.c_int8), rows, cols, _p(colors, C.c_int32), _p(counts, C.c_int32))
    return [(int(colors[i]), int(counts[i])) for i in range(n)]


def drop(board):
    assert board.dtype == np.int8 and board.flags.c_contiguous
    lib().ppad_oracle_drop(_p(board, C.c_int8), board.shape[0], board.shape[1])


def damage(cfg, combos):
    colors = np.array([c for c, _ in combos], np.int32)
    counts = np.array([n for _, n in combos], np.int32)
    return lib().ppad_oracle_damage(C.byref(cfg), _p(colors, C.c_int32), _p(counts, C.c_int32), len(combos))


def resolve(cfg, board, inj=None, env=0, step=0, log_cap=256):
    """calculate_reward (game.py:364-405). inj: uint8 orb sequence or None for pure Philox."""
    board = np.ascontiguousarray(board, np.int8)
    final = np.zeros_like(board)
    log = np.zeros(log_cap, np.uint16)
    res = Result()
    if inj is not None:
        inj = np.ascontiguousarray(inj, np.uint8)
    lib().ppad_oracle_resolve(C.byref(cfg), _p(board, C.c_int8), _p(inj, C.c_uint8),
                              len(inj) if inj is not None else -1, C.c_uint64(env), C.c_uint32(step),
                              C.byref(res), _p(final, C.c_int8), _p(log, C.c_uint16), log_cap)
    n = min(res.n_combos, log_cap)
    return dict(reward=res.reward, n_combos=res.n_combos, depth=res.depth, consumed=res.consumed,
                flags=res.flags, final_board=final, combos=[(int(v) & 0xFF, int(v) >> 8) for v in log[:n]])


def apply_action(board, finger, action):
    assert board.dtype == np.int8 and finger.dtype == np.int32
    return lib().ppad_oracle_apply_action(_p(board, C.c_int8), board.shape[0], board.shape[1],
                                          _p(finger, C.c_int32), int(action))


def legal_mask(rows, cols, finger, prev):
    f = np.ascontiguousarray(finger, np.int32)
    return lib().ppad_oracle_legal_mask(rows, cols, _p(f, C.c_int32), int(prev))


def sample_action(seed, env, step, mask):
    return lib().ppad_oracle_sample_action(C.c_uint64(seed), C.c_uint64(env), C.c_uint32(step), int(mask))


def reset(cfg, inj=None, env=0, step=0, finger=None):
    board = np.zeros((cfg.rows, cfg.cols), np.int8)
    fout = np.zeros(2, np.int32)
    consumed = C.c_int32(0)
    if inj is not None:
        inj = np.ascontiguousarray(inj, np.uint8)
    fin = np.ascontiguousarray(finger, np.int32) if finger is not None else None
    flags = lib().ppad_oracle_reset(C.byref(cfg), _p(inj, C.c_uint8), len(inj) if inj is not None else -1,
                                    C.c_uint64(env), C.c_uint32(step), _p(fin, C.c_int32),
                                    _p(board, C.c_int8), _p(fout, C.c_int32), C.byref(consumed))
    return board, fout, consumed.value, flags


def onehot(boards, fingers, out=None):
    boards = np.ascontiguousarray(boards, np.int8)
    fingers = np.ascontiguousarray(fingers, np.int32)
    n, rows, cols = boards.shape
    if out is None:
        out = np.zeros((n, rows, cols, 7), np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n, rows, cols, 7)
    lib().ppad_oracle_onehot(_p(boards, C.c_int8), _p(fingers, C.c_int32), C.c_int64(n), rows, cols,
                             _p(out, C.c_float))
    return out


def discount(rewards, gamma, log10=True):
    r = np.array(rewards, np.float64)
    lib().ppad_oracle_discount(_p(r, C.c_double), len(r), C.c_double(gamma), int(bool(log10)))
    return r


def step_batch(cfg, boards, fingers, prev, moves=None, actions=None, eval_moves=True, auto_reset=False,
               max_moves=0, inj=None, env0=0, step=0, want_final=False, threads=1):
    """One batched env step; boards/fingers/prev/moves are updated in place (kernel contract).
    threads > 1 splits the batch into contiguous slices run on a thread pool (ctypes drops the GIL)."""
    n = boards.shape[0]
    if threads > 1 and n >= 2 * threads:
        from concurrent.futures import ThreadPoolExecutor
        edges = [n * k // threads for k in range(threads + 1)]

        def run(k):
            sl = slice(edges[k], edges[k + 1])
            return step_batch(cfg, boards[sl], fingers[sl], prev[sl], None if moves is None else moves[sl],
                              None if actions is None else np.ascontiguousarray(actions[sl]), eval_moves, auto_reset,
                              max_moves, None if inj is None else inj[sl], env0 + edges[k], step, want_final, 1)
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(run, range(threads)))
        return tuple(None if parts[0][j] is None else np.concatenate([p[j] for p in parts]) for j in range(4))
    assert boards.dtype == np.int8 and fingers.dtype == np.int32 and prev.dtype == np.uint8
    assert boards.flags.c_contiguous and fingers.flags.c_contiguous
    n = boards.shape[0]
    action_out = np.zeros(n, np.uint8)
    reward = np.zeros(n, np.float64)
    info = np.zeros(n, np.uint32)
    final = np.zeros_like(boards) if want_final else None
    if inj is not None:
        assert inj.dtype == np.uint8 and inj.ndim == 2 and inj.flags.c_contiguous
    if moves is not None:
        assert moves.dtype == np.uint16
    if actions is not None:
        actions = np.ascontiguousarray(actions, np.uint8)
    lib().ppad_oracle_step_batch(
        C.byref(cfg), C.c_int64(n), C.c_uint64(env0), C.c_uint32(step), _p(boards, C.c_int8),
        _p(fingers, C.c_int32), _p(prev, C.c_uint8), _p(moves, C.c_uint16), _p(actions, C.c_uint8),
        int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves), _p(inj, C.c_uint8),
        inj.shape[1] if inj is not None else 0, inj.shape[1] if inj is not None else 0,
        _p(action_out, C.c_uint8), _p(reward, C.c_double), _p(info, C.c_uint32), _p(final, C.c_int8))
    return action_out, reward, info, final


def path_batch(cfg, boards, fingers, prev, moves, paths, lens=None, eval_moves=True, inj=None, env0=0, step=0,
               want_final=False, threads=1):
    """Whole finger paths: paths uint8 [n, L] move ids; boards/fingers/prev/moves updated in place."""
    n = boards.shape[0]
    paths = np.ascontiguousarray(paths, np.uint8)
    lens = np.ascontiguousarray(np.full(n, paths.shape[1]) if lens is None else lens, np.uint16)
    if threads > 1 and n >= 2 * threads:
        from concurrent.futures import ThreadPoolExecutor
        edges = [n * k // threads for k in range(threads + 1)]

        def run(k):
            sl = slice(edges[k], edges[k + 1])
            return path_batch(cfg, boards[sl], fingers[sl], prev[sl], None if moves is None else moves[sl], paths[sl],
                              lens[sl], eval_moves, None if inj is None else inj[sl], env0 + edges[k], step, want_final, 1)
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(run, range(threads)))
        return tuple(None if parts[0][j] is None else np.concatenate([p[j] for p in parts]) for j in range(4))
    action_out = np.zeros(n, np.uint8)
    reward = np.zeros(n, np.float64)
    info = np.zeros(n, np.uint32)
    final = np.zeros_like(boards) if want_final else None
    lib().ppad_oracle_path_batch(
        C.byref(cfg), C.c_int64(n), C.c_uint64(env0), C.c_uint32(step), _p(boards, C.c_int8), _p(fingers, C.c_int32),
        _p(prev, C.c_uint8), _p(moves, C.c_uint16), _p(paths, C.c_uint8), paths.shape[1], _p(lens, C.c_uint16),
        int(bool(eval_moves)), _p(inj, C.c_uint8), inj.shape[1] if inj is not None else 0,
        inj.shape[1] if inj is not None else 0, _p(action_out, C.c_uint8), _p(reward, C.c_double), _p(info, C.c_uint32),
        _p(final, C.c_int8))
    return action_out, reward, info, final
