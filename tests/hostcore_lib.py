"""ctypes wrapper of tests/hostcore/libhostcore.so (host build of the device core). TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

import ppad_b200
from ppad_b200 import _cabi

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcore")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "-s"], stderr=subprocess.DEVNULL)
        _lib = C.CDLL(os.path.join(_DIR, "libhostcore.so"))
        _lib.hostcore_last_error.restype = C.c_char_p
    return _lib


def cfg_from_oracle(o, orb_bits=None):
    """oracle.Cfg -> ppad_config (same fields, different struct)"""
    cfg = _cabi.Config()
    cfg.rows, cfg.cols, cfg.skyfall_damage = o.rows, o.cols, o.skyfall_damage
    cfg.cascade_cap, cfg.reset_cap, cfg.team_size, cfg.seed = o.cascade_cap, o.reset_cap, o.team_n, o.seed
    for i in range(8):
        cfg.team_color_1[i], cfg.team_color_2[i], cfg.team_atk[i] = o.team_c1[i], o.team_c2[i], o.team_atk[i]
    for i in range(10):
        cfg.skyfall[i] = o.skyfall[i]
    cfg.orb_bits = 4 if (orb_bits == 4 or any(o.skyfall[i] > 0 for i in range(7, 10))) else 3
    return cfg


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def step_batch(cfg, boards, fingers, prev, moves=None, actions=None, eval_moves=True, auto_reset=False,
               max_moves=0, inj=None, env0=0, step=0, want_final=False, log_cap=0, paths=None, path_lens=None,
               slab_format=0):
    n = boards.shape[0]
    action_out = np.zeros(n, np.uint8)
    reward = np.zeros(n, np.float64)
    info = np.zeros(n, np.uint32)
    final = np.zeros_like(boards) if want_final else None
    log = np.zeros((n, log_cap), np.uint16) if log_cap else None
    slab = None
    if inj is not None:
        slab = {0: ppad_b200.layout.pack_slab, 1: ppad_b200.layout.pack_slab_planes3,
                2: ppad_b200.layout.pack_slab_chunks10, 3: ppad_b200.layout.pack_slab_chunks6,
                4: ppad_b200.layout.pack_slab_chunks6t}[slab_format](inj)
    if actions is not None:
        actions = np.ascontiguousarray(actions, np.uint8)
    packed = None
    if paths is not None:
        paths = np.asarray(paths, np.uint8)
        words = max(1, (paths.shape[1] + 15) // 16)
        pad = np.zeros((n, words * 16), np.uint64)
        pad[:, :paths.shape[1]] = paths & 3
        packed = np.ascontiguousarray((pad.reshape(n, words, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(2).astype(np.uint32))
        path_lens = np.ascontiguousarray(path_lens if path_lens is not None else np.full(n, paths.shape[1]), np.uint16)
    r = lib().hostcore_step_batch(
        C.byref(cfg), C.c_int64(n), C.c_uint64(env0), C.c_uint32(step), _p(boards), _p(fingers), _p(prev), _p(moves),
        _p(actions), int(bool(eval_moves)), int(bool(auto_reset)), int(max_moves), _p(slab),
        (slab.shape[0] if slab_format in (2, 3) else slab.shape[1]) if slab is not None else 0,   # chunks / words per board
        inj.shape[1] if inj is not None else 0, _p(action_out), _p(reward),
        _p(info), _p(final), _p(log), log_cap, _p(packed), packed.shape[1] if packed is not None else 0, _p(path_lens),
        int(slab_format))
    if r:
        raise RuntimeError(lib().hostcore_last_error().decode())
    return action_out, reward, info, final, log


def reset(cfg, n, inj=None, env0=0, step=0, fingers_in=None):
    boards = np.zeros((n, cfg.rows, cfg.cols), np.int8)
    fingers = np.zeros((n, 2), np.int32)
    flags = np.zeros(n, np.uint8)
    consumed = np.zeros(n, np.int32)
    slab = ppad_b200.layout.pack_slab(inj) if inj is not None else None
    fin = np.ascontiguousarray(fingers_in, np.int32) if fingers_in is not None else None
    r = lib().hostcore_reset(C.byref(cfg), C.c_int64(n), C.c_uint64(env0), C.c_uint32(step), _p(slab),
                             slab.shape[1] if slab is not None else 0, inj.shape[1] if inj is not None else 0,
                             _p(fin), _p(boards), _p(fingers), _p(flags), _p(consumed))
    if r:
        raise RuntimeError(lib().hostcore_last_error().decode())
    return boards, fingers, flags, consumed


def cancel(boards, log_cap=32, orb_bits=3):
    n, rows, cols = boards.shape
    n_combos = np.zeros(n, np.uint8)
    log = np.zeros((n, log_cap), np.uint16)
    assert lib().hostcore_cancel(rows, cols, orb_bits, C.c_int64(n), _p(boards), _p(n_combos), _p(log), log_cap) == 0
    return n_combos, log


def drop(boards, orb_bits=3):
    n, rows, cols = boards.shape
    assert lib().hostcore_drop(rows, cols, orb_bits, C.c_int64(n), _p(boards)) == 0


def obs(boards, fingers, orb_bits=3):
    n, rows, cols = boards.shape
    out = np.zeros((n, rows, cols, 7), np.uint8)
    assert lib().hostcore_obs(rows, cols, orb_bits, C.c_int64(n), _p(boards), _p(np.ascontiguousarray(fingers, np.int32)),
                              _p(out)) == 0
    return out


def sky_table(cfg):
    thr = (C.c_uint32 * 10)()
    valid = C.c_uint32(0)
    r = lib().hostcore_sky_table(C.byref(cfg), thr, C.byref(valid))
    if r:
        raise RuntimeError(lib().hostcore_last_error().decode())
    return list(thr), valid.value


def uniform_k(cfg):
    return lib().hostcore_uniform_k(C.byref(cfg))


def orb_from_u32(cfg, x):
    return lib().hostcore_orb_from_u32(C.byref(cfg), C.c_uint32(x))


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().hostcore_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)
