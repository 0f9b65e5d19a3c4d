"""ctypes binding of libppad_b200.so (include/ppad_b200.h).  There is no CPU path: if the library
is missing or a call fails, this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPAD_B200_LIB: an alternative build of the same library (profiles/variants.py A/B experiments)
LIB_PATH = os.environ.get("PPAD_B200_LIB") or os.path.join(_HERE, "libppad_b200.so")
MAX_TEAM_UNITS = 8

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
UP, DOWN, LEFT, RIGHT, PASS, NOOP, SAMPLE = 0, 1, 2, 3, 4, 5, 255
FLAG_REJECTED, FLAG_SKYFALL_EXHAUSTED, FLAG_CASCADE_CAP, FLAG_NO_LEGAL_MOVE = 0x01, 0x02, 0x04, 0x08
FLAG_BAD_ACTION, FLAG_RESET_CAP, FLAG_UNSUPPORTED_ORB, FLAG_EPISODE_DONE = 0x10, 0x20, 0x40, 0x80
I8, I32, I64 = 1, 4, 8
OBS_F32, OBS_U8 = 0, 1
SELECT_MAX, SELECT_BOLTZMANN = 0, 1
SLAB_NIBBLES, SLAB_PLANES3, SLAB_CHUNKS10, SLAB_CHUNKS6, SLAB_CHUNKS6T = 0, 1, 2, 3, 4
SLAB_FORMATS = {'nibbles': SLAB_NIBBLES, 'planes3': SLAB_PLANES3, 'chunks10': SLAB_CHUNKS10, 'chunks6': SLAB_CHUNKS6,
                'chunks6t': SLAB_CHUNKS6T}
CHUNK_MAJOR = {'chunks10': (10, 4, 2), 'chunks6': (6, 2, 3), 'chunks6t': (6, 2, 3)}   # orbs per chunk, bytes per chunk, default eager chunks
ABI_VERSION = 4


class Config(C.Structure):
    """struct ppad_config"""
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("skyfall_damage", C.c_int32),
                ("cascade_cap", C.c_int32), ("reset_cap", C.c_int32), ("team_size", C.c_int32),
                ("team_color_1", C.c_int32 * MAX_TEAM_UNITS), ("team_color_2", C.c_int32 * MAX_TEAM_UNITS),
                ("team_atk", C.c_double * MAX_TEAM_UNITS), ("skyfall", C.c_double * 10), ("seed", C.c_uint64),
                ("orb_bits", C.c_int32), ("reserved", C.c_int32)]


class StepArgs(C.Structure):
    """struct ppad_step_args"""
    _fields_ = [("n", C.c_int64), ("env0", C.c_uint64), ("step", C.c_uint32), ("eval_moves", C.c_int32),
                ("auto_reset", C.c_int32), ("max_moves", C.c_int32), ("state", C.c_void_p),
                ("actions", C.c_void_p), ("slab", C.c_void_p), ("slab_words", C.c_int32), ("n_inj", C.c_int32),
                ("action_out", C.c_void_p), ("reward", C.c_void_p), ("info", C.c_void_p),
                ("final_state", C.c_void_p), ("state_copy", C.c_void_p), ("combo_log", C.c_void_p),
                ("log_cap", C.c_int32),
                ("obs", C.c_void_p), ("obs_dtype", C.c_int32), ("skyfall_damage", C.c_int32),
                ("paths", C.c_void_p), ("path_lens", C.c_void_p), ("path_words", C.c_int32), ("path_len", C.c_int32),
                ("slab_format", C.c_int32), ("slab_stride", C.c_int32), ("consumed_out", C.c_void_p),
                ("slab_eager", C.c_int32), ("reserved", C.c_int32)]


class HostStepArgs(C.Structure):
    """struct ppad_host_step_args"""
    _fields_ = [("n", C.c_int64), ("env0", C.c_uint64), ("step", C.c_uint32), ("eval_moves", C.c_int32),
                ("auto_reset", C.c_int32), ("max_moves", C.c_int32), ("skyfall_damage", C.c_int32),
                ("state", C.c_void_p), ("actions_host", C.c_void_p), ("slab_host", C.c_void_p),
                ("slab_words", C.c_int32), ("n_inj", C.c_int32), ("action_out_host", C.c_void_p),
                ("reward_host", C.c_void_p), ("info_host", C.c_void_p), ("state_out_host", C.c_void_p),
                ("obs", C.c_void_p), ("obs_dtype", C.c_int32), ("slab_format", C.c_int32),
                ("results_compact", C.c_int32), ("nibbles_host", C.c_void_p), ("tile_line_host", C.c_void_p),
                ("records_host", C.c_void_p), ("records_capacity", C.c_int64), ("summary_host", C.c_void_p),
                ("no_join", C.c_int32), ("slab_stride", C.c_int32), ("slab_eager", C.c_int32), ("reserved2", C.c_int32)]


class PolicyArgs(C.Structure):
    """struct ppad_policy_args"""
    _fields_ = [("n", C.c_int64), ("env0", C.c_uint64), ("step", C.c_uint32), ("eval_moves", C.c_int32),
                ("auto_reset", C.c_int32), ("max_moves", C.c_int32), ("state", C.c_void_p), ("q", C.c_void_p),
                ("method", C.c_int32), ("beta", C.c_float), ("epsilon", C.c_float), ("slab", C.c_void_p),
                ("slab_words", C.c_int32), ("n_inj", C.c_int32), ("slab_format", C.c_int32), ("obs_dtype", C.c_int32),
                ("action_out", C.c_void_p), ("reward", C.c_void_p), ("info", C.c_void_p), ("obs", C.c_void_p),
                ("ring_state", C.c_void_p), ("ring_action", C.c_void_p), ("ring_reward", C.c_void_p),
                ("ring_steps", C.c_int64), ("ring_step", C.c_int64), ("gamma", C.c_double),
                ("log10_reward", C.c_int32), ("ctas_per_sm", C.c_int32), ("stats_rows", C.c_void_p)]


# every symbol include/ppad_b200.h declares (tests/test_cabi_symbols.py checks the header against this)
SYMBOLS = ["ppad_default_config", "ppad_create", "ppad_destroy", "ppad_abi_version", "ppad_strerror",
           "ppad_last_error", "ppad_state_bytes", "ppad_state_bytes_ex", "ppad_launch_count", "ppad_pack", "ppad_unpack", "ppad_reset",
           "ppad_step", "ppad_encode_obs", "ppad_cancel", "ppad_legal_mask", "ppad_sample_actions", "ppad_drop",
           "ppad_fill", "ppad_damage", "ppad_select_actions", "ppad_discount", "ppad_replay_push",
           "ppad_replay_sample", "ppad_permute", "ppad_pipe_create", "ppad_pipe_destroy", "ppad_pipe_step",
           "ppad_pipe_wait", "ppad_pipe_join", "ppad_rollout", "ppad_pipe_set_zero_copy", "ppad_rollout_record",
           "ppad_compact_capacity", "ppad_pipe_set_dictionary", "ppad_policy_step", "ppad_policy_stats", "ppad_policy_stats_rows", "ppad_island"]


class PPADError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("ppad_b200 status %d: %s" % (status, message))
        self.status = status
        self.detail = message


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
                          "ppad_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    L.ppad_default_config.argtypes = [C.POINTER(Config)]
    L.ppad_default_config.restype = None
    L.ppad_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(vp)]
    L.ppad_destroy.argtypes = [vp]
    L.ppad_destroy.restype = None
    L.ppad_strerror.argtypes = [C.c_int]
    L.ppad_strerror.restype = C.c_char_p
    L.ppad_last_error.restype = C.c_char_p
    L.ppad_state_bytes.argtypes = [C.c_int, C.c_int]
    L.ppad_state_bytes.restype = i64
    L.ppad_state_bytes_ex.argtypes = [C.c_int, C.c_int, C.c_int]
    L.ppad_state_bytes_ex.restype = i64
    L.ppad_launch_count.restype = u64
    L.ppad_pack.argtypes = [vp, i64, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    L.ppad_unpack.argtypes = [vp, i64, vp, vp, vp, C.c_int, vp, vp, vp]
    L.ppad_reset.argtypes = [vp, i64, u64, u32, vp, i32, i32, vp, vp, vp, vp, vp]
    L.ppad_step.argtypes = [vp, C.POINTER(StepArgs), vp]
    L.ppad_encode_obs.argtypes = [vp, i64, vp, vp, C.c_int, vp]
    L.ppad_cancel.argtypes = [vp, i64, vp, vp, vp, i32, vp]
    L.ppad_legal_mask.argtypes = [vp, i64, vp, vp, vp]
    L.ppad_island.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    L.ppad_sample_actions.argtypes = [vp, i64, u64, u32, vp, i32, vp, vp]
    L.ppad_drop.argtypes = [vp, i64, vp, vp]
    L.ppad_fill.argtypes = [vp, i64, u64, u32, i32, i32, vp, i32, i32, vp, vp, vp, vp]
    L.ppad_damage.argtypes = [vp, i64, vp, i32, vp, vp, vp]
    L.ppad_permute.argtypes = [vp, i64, vp, vp, vp]
    L.ppad_rollout.argtypes = [vp, i64, u64, u32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ppad_rollout_record.argtypes = [vp, i64, u64, u32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ppad_pipe_create.argtypes = [vp, i64, i32, i32, C.POINTER(vp)]
    L.ppad_pipe_set_dictionary.argtypes = [vp, vp, vp, i32]
    L.ppad_pipe_destroy.argtypes = [vp]
    L.ppad_pipe_destroy.restype = None
    L.ppad_pipe_step.argtypes = [vp, C.POINTER(HostStepArgs), vp, i32]
    L.ppad_pipe_wait.argtypes = [vp, i32]
    L.ppad_pipe_set_zero_copy.argtypes = [vp, i32]
    L.ppad_pipe_join.argtypes = [vp, vp, i32]
    L.ppad_select_actions.argtypes = [vp, i64, u64, u32, vp, vp, i32, C.c_float, C.c_float, i32, vp, vp]
    L.ppad_discount.argtypes = [vp, i64, i32, vp, vp, C.c_double, i32, i32, vp]
    L.ppad_replay_push.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ppad_policy_step.argtypes = [vp, C.POINTER(PolicyArgs), vp]
    L.ppad_policy_stats.argtypes = [vp, i64, vp, vp, vp]
    L.ppad_policy_stats_rows.argtypes = [i64]
    L.ppad_policy_stats_rows.restype = i64
    L.ppad_compact_capacity.argtypes = [vp, i64]
    L.ppad_compact_capacity.restype = i64
    L.ppad_replay_sample.argtypes = [vp, i64, i64, u32, vp, vp, vp, vp, vp, vp, vp, vp]
    for name in SYMBOLS:
        getattr(L, name)
    if L.ppad_abi_version() != ABI_VERSION:
        raise ImportError("libppad_b200.so ABI version mismatch")
    _lib = L
    return L


def check(status):
    if status != OK:
        L = lib()
        raise PPADError(status, "%s: %s" % (L.ppad_strerror(status).decode(), L.ppad_last_error().decode()))


def default_config():
    cfg = Config()
    lib().ppad_default_config(C.byref(cfg))
    return cfg
