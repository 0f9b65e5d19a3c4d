"""The C-ABI library loads here (no GPU) and exports every symbol include/ppad_b200.h declares;
constants agree between the header, the ctypes binding and the oracle.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "ppad_b200.h")).read()


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__
    __graft_entry__.build()
    from ppad_b200 import _cabi
    return _cabi


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(ppad_[a-z0-9_]+)\s*\(", body)))


def test_every_declared_symbol_is_exported_and_bound(cabi):
    names = declared_functions()
    assert len(names) >= 20
    L = cabi.lib()
    for name in names:
        assert hasattr(L, name), name
    assert sorted(cabi.SYMBOLS) == names


def test_header_cites_the_reference_for_every_entry_point():
    # each prototype's comment names the reference file it replaces
    for name in declared_functions():
        if name in ("ppad_destroy", "ppad_abi_version", "ppad_strerror", "ppad_last_error", "ppad_state_bytes",
                    "ppad_launch_count", "ppad_replay_sample", "ppad_unpack", "ppad_compact_capacity") or name.startswith("ppad_pipe_"):
            continue      # helpers of an entry point that carries the citation (the pipe: its struct comment)
        idx = HEADER.index(name + "(")
        assert re.search(r"\.py:\d+", HEADER[max(0, idx - 2500):idx]), name


def test_constants_agree(cabi):
    def henum(name):
        return int(re.search(name + r"\s*=\s*(0x[0-9A-Fa-f]+|\d+)", HEADER).group(1), 0)
    for flag in ("REJECTED", "SKYFALL_EXHAUSTED", "CASCADE_CAP", "NO_LEGAL_MOVE", "BAD_ACTION", "RESET_CAP",
                 "UNSUPPORTED_ORB", "EPISODE_DONE"):
        assert henum("PPAD_FLAG_" + flag) == getattr(cabi, "FLAG_" + flag) == getattr(orc, "FLAG_" + flag)
    assert [henum("PPAD_" + a) for a in ("UP", "DOWN", "LEFT", "RIGHT", "PASS")] == [0, 1, 2, 3, 4]
    assert C.sizeof(cabi.Config) == 6 * 4 + 2 * 8 * 4 + 8 * 8 + 10 * 8 + 8 + 2 * 4   # ... seed, orb_bits, reserved
    assert cabi.lib().ppad_abi_version() == int(re.search(r"#define PPAD_ABI_VERSION\s+(\d+)", HEADER).group(1)) == cabi.ABI_VERSION == 4
    for name in ("SLAB_NIBBLES", "SLAB_PLANES3"):
        assert henum("PPAD_" + name) == getattr(cabi, name)
    # struct sizes the ctypes mirrors must agree on (a mismatch would shift every later field)
    assert C.sizeof(cabi.StepArgs) == 8 + 8 + 4 * 4 + 3 * 8 + 2 * 4 + 6 * 8 + 4 + 4 + 8 + 2 * 4 + 2 * 8 + 2 * 4 + 2 * 4 + 8 + 2 * 4
    assert C.sizeof(cabi.HostStepArgs) % 8 == 0
    assert cabi.lib().ppad_state_bytes(5, 6) == 16 and cabi.lib().ppad_state_bytes(6, 7) == 32
    assert cabi.lib().ppad_state_bytes(9, 9) == -1 and cabi.lib().ppad_state_bytes(4, 5) == 16
    # every other size up to 64 cells runs on the run-time-dimension kernels: 64-bit planes
    assert cabi.lib().ppad_state_bytes(4, 4) == 32 and cabi.lib().ppad_state_bytes_ex(8, 8, 4) == 48


def test_default_config_is_the_reference_default(cabi):
    cfg = cabi.default_config()
    o = orc.make_cfg()
    assert (cfg.rows, cfg.cols, cfg.skyfall_damage, cfg.team_size) == (5, 6, 1, 6) == (o.rows, o.cols, o.skyfall_damage, o.team_n)
    assert list(cfg.team_color_1)[:6] == list(o.team_c1)[:6] == [4, 3, 1, 2, 0, 1]
    assert list(cfg.team_color_2)[:6] == list(o.team_c2)[:6] == [4, 3, 2, 0, 1, -1]
    assert list(cfg.skyfall) == list(o.skyfall) == [1 / 6] * 6 + [0] * 4


def test_no_cpu_fallback_without_a_device(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = C.c_void_p()
    cfg = cabi.default_config()
    status = cabi.lib().ppad_create(C.byref(cfg), 0, C.byref(handle))
    assert status == cabi.ERR_CUDA and not handle.value
    assert b"no CPU path" in cabi.lib().ppad_last_error()
    import ppad_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ppad_b200.BatchedPAD(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ppad_b200.PAD()


def test_config_validation_messages_match_the_reference(cabi):
    # validation happens before the device check, so it is testable without a GPU (game.py:54-68)
    handle = C.c_void_p()
    for mutate, msg, status in (
            (lambda c: setattr(c, "rows", 2), b"at least 13 orbs", cabi.ERR_INVALID),
            (lambda c: c.skyfall.__setitem__(0, 0.5), b"add up to 1", cabi.ERR_INVALID),
            (lambda c: (c.skyfall.__setitem__(0, -0.1), c.skyfall.__setitem__(1, 1 / 6 + 0.1 + 1 / 6)), b"positive number", cabi.ERR_INVALID),
            (lambda c: (setattr(c, "rows", 9), setattr(c, "cols", 9)), b"more than 64 cells", cabi.ERR_UNSUPPORTED),
            (lambda c: (c.skyfall.__setitem__(9, 1 / 6), c.skyfall.__setitem__(0, 0.0)), b"3-bit", cabi.ERR_UNSUPPORTED),
            (lambda c: c.team_color_1.__setitem__(0, 5), b"damage_tracker", cabi.ERR_UNSUPPORTED)):
        cfg = cabi.default_config()
        mutate(cfg)
        assert cabi.lib().ppad_create(C.byref(cfg), 0, C.byref(handle)) == status
        assert msg in cabi.lib().ppad_last_error(), cabi.lib().ppad_last_error()


def test_ctypes_struct_layouts_match_the_c_header(cabi, tmp_path):
    """Compile include/ppad_b200.h with gcc (the header is plain C) and compare sizeof / offsetof of every struct with
    the ctypes mirrors: a drifted field would silently shift every later argument."""
    import subprocess
    structs = {"ppad_config": cabi.Config, "ppad_step_args": cabi.StepArgs, "ppad_host_step_args": cabi.HostStepArgs,
               "ppad_policy_args": cabi.PolicyArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ppad_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
