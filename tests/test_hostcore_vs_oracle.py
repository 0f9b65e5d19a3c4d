"""The per-board device code (p-pad_b200/csrc/ppad_core.cuh), compiled for the host as test
infrastructure, against the CPU oracle and the reference-generated fixtures.  CPU only: this is how
the bit-sliced kernels are debugged in the GPU-less container; tests/test_gpu_parity.py repeats the
comparison through the C ABI on the B200."""
import json
import os

import numpy as np
import pytest

import oracle as orc
import hostcore_lib as hc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(x):
    return np.asarray(x, np.float64).view(np.uint64)


def random_boards(rng, n, rows, cols, kind):
    if kind == "uniform6":
        b = rng.integers(0, 6, size=(n, rows, cols))
    elif kind == "few":
        b = rng.integers(0, rng.integers(2, 5), size=(n, rows, cols))
    elif kind == "jammer_empty":
        b = rng.integers(-1, 7, size=(n, rows, cols))
    elif kind == "all10":                                   # every orb type of game.py:236-246 and empties
        b = rng.integers(-1, 10, size=(n, rows, cols))
    elif kind == "poison_few":                              # few colours, among them types 7..9: big islands
        pal = np.array([0, 7, 8, 9, 5])[: rng.integers(2, 6)]
        b = pal[rng.integers(0, len(pal), size=(n, rows, cols))]
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(b, np.int8)


@pytest.mark.parametrize("rows,cols", [(4, 5), (5, 6), (6, 7)])
@pytest.mark.parametrize("kind,orb_bits", [("uniform6", 3), ("few", 3), ("jammer_empty", 3), ("uniform6", 4),
                                           ("jammer_empty", 4), ("all10", 4), ("poison_few", 4)])
def test_cancel_and_drop_fuzz(rows, cols, kind, orb_bits):
    rng = np.random.default_rng(sum(map(ord, kind)) * 7 + rows + orb_bits)
    n = 4000
    boards = random_boards(rng, n, rows, cols, kind)
    mine = boards.copy()
    n_combos, log = hc.cancel(mine, log_cap=32, orb_bits=orb_bits)
    for i in range(n):
        b = boards[i].copy()
        combos = orc.cancel(b)
        assert n_combos[i] == len(combos), i
        assert [int(v) for v in log[i, :len(combos)]] == [c | (k << 8) for c, k in combos], i
        assert np.array_equal(b, mine[i]), i
        orc.drop(b)
        boards[i] = b
    hc.drop(mine, orb_bits=orb_bits)
    assert np.array_equal(mine, boards)


@pytest.mark.parametrize("rows,cols", [(4, 5), (5, 6), (6, 7)])
@pytest.mark.parametrize("mode", ["injected", "philox", "injected_short", "no_skyfall", "wide_injected", "wide_philox",
                                  "wide3_injected", "planes3_injected", "planes3_short", "chunks10_injected",
                                  "chunks10_short", "chunks6_injected", "chunks6_short", "chunks6t_injected"])
def test_step_batch_fuzz(rows, cols, mode):
    """wide_*: the 4-bit kernels with all ten orb types (skyfall rates and injected orbs for types 7..9);
    wide3_injected: the 4-bit kernels on a workload the 3-bit kernels also run;
    planes3_*: the injected skyfall as a bit-sliced 3-bit slab (PPAD_SLAB_PLANES3) of 96 / 40 orbs;
    chunks10_*: as a chunk-major slab of ten orbs per word (PPAD_SLAB_CHUNKS10) of 95 / 23 orbs;
    chunks6_*: as a chunk-major slab of six base-6 orbs per uint16 (PPAD_SLAB_CHUNKS6) of 95 / 23 orbs"""
    rng = np.random.default_rng(1000 + rows + len(mode))
    n, steps = 3000, 12
    wide = mode.startswith("wide")
    sky = [0.05, 0.15, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1] if mode in ("wide_injected", "wide_philox") else None
    ocfg = orc.make_cfg(rows, cols, skyfall_damage=(mode != "no_skyfall"), seed=0x1234567890ABCDEF, skyfall=sky)
    cfg = hc.cfg_from_oracle(ocfg, orb_bits=4 if wide else None)
    kind = ["uniform6", "few", "jammer_empty"][rows % 3] if not mode.endswith("philox") else "uniform6"
    if sky:
        kind = ["all10", "poison_few", "all10"][rows % 3]
    boards = random_boards(rng, n, rows, cols, kind)
    fingers = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1).astype(np.int32)
    prev = np.full(n, 4, np.uint8)
    moves = np.zeros(n, np.uint16)
    mine = [boards.copy(), fingers.copy(), prev.copy(), moves.copy()]
    S = {"injected": 256, "philox": 0, "injected_short": 8, "no_skyfall": 16, "wide_injected": 256, "wide_philox": 0,
         "wide3_injected": 256, "planes3_injected": 96, "planes3_short": 40,
         "chunks10_injected": 95, "chunks10_short": 23, "chunks6_injected": 95, "chunks6_short": 23, "chunks6t_injected": 95}[mode]
    fmt = {"planes3": 1, "chunks10": 2, "chunks6": 3, "chunks6t": 4}.get(mode.split("_")[0], 0)
    top = 10 if sky else 6
    seen_flags = 0
    for t in range(steps):
        inj = rng.integers(0, top, size=(n, S)).astype(np.uint8) if S else None
        if t % 3 == 2:
            actions = None                                   # in-kernel random policy
        else:
            actions = rng.integers(0, 6, size=n).astype(np.uint8)   # includes off-board moves, passes and no-ops (5)
            if t == 4:
                actions[:50] = 9                             # bad action ids
        kw = dict(actions=actions, eval_moves=(t % 4 != 3), auto_reset=(t % 2 == 0), max_moves=5,
                  inj=inj, env0=77777, step=t, want_final=True)
        a0, r0, i0, f0 = orc.step_batch(ocfg, boards, fingers, prev, moves, **kw)
        a1, r1, i1, f1, _ = hc.step_batch(cfg, *mine, slab_format=fmt, **kw)
        assert np.array_equal(a0, a1), t
        assert np.array_equal(boards, mine[0]) and np.array_equal(fingers, mine[1]), t
        assert np.array_equal(prev, mine[2]) and np.array_equal(moves, mine[3]), t
        assert np.array_equal(i0, i1), (t, np.nonzero(i0 != i1)[0][:5])
        assert np.array_equal(bits(r0), bits(r1)), t
        assert np.array_equal(f0, f1), t
        seen_flags |= int(np.bitwise_or.reduce(i0 >> 24))
    assert seen_flags & orc.FLAG_REJECTED and seen_flags & orc.FLAG_BAD_ACTION and seen_flags & orc.FLAG_EPISODE_DONE
    if mode in ("injected_short", "chunks10_short", "chunks6_short"):
        assert seen_flags & orc.FLAG_SKYFALL_EXHAUSTED


ODD_SIZES = [(3, 5), (7, 9), (8, 8), (2, 7), (1, 13), (13, 1), (6, 5), (4, 8), (3, 21), (5, 5)]


@pytest.mark.parametrize("rows,cols", ODD_SIZES)
def test_any_board_size_up_to_64_cells(rows, cols):
    """PAD.__init__ takes every dim_row x dim_col with at least 13 orbs (game.py:54-68): the run-time-dimension
    geometry (GeomDyn) against the oracle -- cancel incl. combo order, drop, full steps with cascades, reset."""
    rng = np.random.default_rng(rows * 100 + cols)
    n = 1500
    boards = random_boards(rng, n, rows, cols, "few")
    mine = boards.copy()
    n_combos, log = hc.cancel(mine, log_cap=24)
    for i in range(0, n, 3):
        b = boards[i].copy()
        combos = orc.cancel(b)
        assert n_combos[i] == len(combos) and [int(v) for v in log[i, :len(combos)]] == [c | (k << 8) for c, k in combos], i
        assert np.array_equal(b, mine[i]), i
    for wide in (False, True):
        sky = [0.05, 0.15, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1] if wide else None
        ocfg = orc.make_cfg(rows, cols, True, seed=99, skyfall=sky)
        cfg = hc.cfg_from_oracle(ocfg)
        b0 = random_boards(rng, n, rows, cols, "all10" if wide else "uniform6")
        fingers = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1).astype(np.int32)
        a = [b0.copy(), fingers.copy(), np.full(n, 4, np.uint8), np.zeros(n, np.uint16)]
        b = [b0.copy(), fingers.copy(), np.full(n, 4, np.uint8), np.zeros(n, np.uint16)]
        for t in range(6):
            inj = rng.integers(0, 10 if wide else 6, size=(n, 40)).astype(np.uint8) if t % 2 else None
            actions = None if t % 3 == 2 else rng.integers(0, 6, size=n).astype(np.uint8)
            kw = dict(actions=actions, eval_moves=(t != 3), auto_reset=(t % 2 == 0), max_moves=4, inj=inj, env0=5, step=t,
                      want_final=True)
            a0, r0, i0, f0 = orc.step_batch(ocfg, *a, **kw)
            a1, r1, i1, f1, _ = hc.step_batch(cfg, *b, **kw)
            assert np.array_equal(a0, a1) and np.array_equal(i0, i1) and np.array_equal(bits(r0), bits(r1)) and np.array_equal(f0, f1), t
            for u, v in zip(a, b):
                assert np.array_equal(u, v), t
        rb, rf, rfl, rcons = hc.reset(cfg, 64, env0=3, step=9)
        for i in range(64):
            ob, of, ocons, _ = orc.reset(ocfg, env=3 + i, step=9)
            assert np.array_equal(ob, rb[i]) and np.array_equal(of, rf[i]) and ocons == rcons[i], i


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_injected_codes_that_are_no_orb_type_are_flagged(fmt):
    """a nibble of 7 (3-bit context) writes an empty cell, 8..15 alias: flagged, never silent."""
    rng = np.random.default_rng(5)
    n = 512
    ocfg = orc.make_cfg(5, 6, True, seed=3)
    cfg = hc.cfg_from_oracle(ocfg)
    boards = np.zeros((n, 5, 6), np.int8)
    boards[:, ::2, :] = 1                                  # rows of one colour: every board cascades and draws orbs
    args = lambda: (boards.copy(), np.zeros((n, 2), np.int32), np.full(n, 4, np.uint8), np.zeros(n, np.uint16))
    good = rng.integers(0, 6, size=(n, 64)).astype(np.uint8)
    bad = good.copy()
    bad[::2, 3] = 7                                        # the 4th orb drawn: every board draws at least 18
    passes = np.full(n, 4, np.uint8)
    _, _, i_good, _, _ = hc.step_batch(cfg, *args(), actions=passes, inj=good, slab_format=fmt)
    _, _, i_bad, _, _ = hc.step_batch(cfg, *args(), actions=passes, inj=bad, slab_format=fmt)
    assert not ((i_good >> 24) & orc.FLAG_UNSUPPORTED_ORB).any()
    assert ((i_bad[::2] >> 24) & orc.FLAG_UNSUPPORTED_ORB).all()
    assert not ((i_bad[1::2] >> 24) & orc.FLAG_UNSUPPORTED_ORB).any()


@pytest.mark.parametrize("rows,cols", [(4, 5), (5, 6), (6, 7)])
def test_finger_paths(rows, cols):
    rng = np.random.default_rng(300 + rows)
    n, L = 2000, 200
    ocfg = orc.make_cfg(rows, cols, True, seed=5)
    cfg = hc.cfg_from_oracle(ocfg)
    boards = random_boards(rng, n, rows, cols, "uniform6")
    fingers = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1).astype(np.int32)
    paths = rng.integers(0, 4, size=(n, L)).astype(np.uint8)
    lens = rng.integers(0, L + 1, size=n).astype(np.uint16)
    lens[:10] = 0
    inj = rng.integers(0, 6, size=(n, 64)).astype(np.uint8)
    a = [boards.copy(), fingers.copy(), np.full(n, 4, np.uint8), np.zeros(n, np.uint16)]
    b = [boards.copy(), fingers.copy(), np.full(n, 4, np.uint8), np.zeros(n, np.uint16)]
    a0, r0, i0, f0 = orc.path_batch(ocfg, *a, paths, lens, inj=inj, step=2, want_final=True)
    a1, r1, i1, f1, _ = hc.step_batch(cfg, *b, inj=inj, step=2, want_final=True, paths=paths, path_lens=lens)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    assert np.array_equal(a0, a1) and np.array_equal(i0, i1) and np.array_equal(bits(r0), bits(r1)) and np.array_equal(f0, f1)
    assert ((i0 >> 24) & orc.FLAG_REJECTED).any() and (a0[:10] == 255).all()


def test_cascade_cap_flag():
    # a one-colour skyfall never stabilises: both sides stop at the cap and say so
    ocfg = orc.make_cfg(5, 6, True, skyfall=[1, 0, 0, 0, 0, 0, 0, 0, 0, 0], cascade_cap=7)
    cfg = hc.cfg_from_oracle(ocfg)
    boards = np.zeros((4, 5, 6), np.int8)
    args = lambda: (boards.copy(), np.zeros((4, 2), np.int32), np.full(4, 4, np.uint8), np.zeros(4, np.uint16))
    _, r0, i0, _ = orc.step_batch(ocfg, *args(), actions=np.full(4, 4, np.uint8))
    _, r1, i1, _, _ = hc.step_batch(cfg, *args(), actions=np.full(4, 4, np.uint8))
    assert np.array_equal(i0, i1) and np.array_equal(bits(r0), bits(r1))
    assert ((i0 >> 24) & orc.FLAG_CASCADE_CAP).all() and (((i0 >> 8) & 0xFF) == 7).all()


@pytest.mark.parametrize("tag,rows,cols,team,orb_bits", [
    ("5x6", 5, 6, None, 3), ("6x7", 6, 7, None, 3), ("4x5", 4, 5, None, 3), ("5x6_custom_team", 5, 6, "custom", 3),
    ("5x6", 5, 6, None, 4), ("6x7", 6, 7, None, 4), ("5x6_all10", 5, 6, None, 4), ("6x7_all10", 6, 7, None, 4)])
def test_resolve_golden_vectors(tag, rows, cols, team, orb_bits):
    """straight against what the reference produced; with the combo log (islands in the reference's order) and
    without it (the colour-major tally the production kernels run)"""
    z = np.load(os.path.join(G, "resolve_%s.npz" % tag))
    t = json.load(open(os.path.join(G, "custom_team.json"))) if team else None
    n = len(z["boards"])
    for sd in (0, 1):
        sel = np.nonzero(z["skyfall_damage"] == sd)[0]
        cfg = hc.cfg_from_oracle(orc.make_cfg(rows, cols, skyfall_damage=sd, team=t, cascade_cap=0), orb_bits=orb_bits)
        k = len(sel)
        for log_cap in (64, 0):
            boards = np.ascontiguousarray(z["boards"][sel])
            _, reward, info, final, log = hc.step_batch(
                cfg, boards, np.zeros((k, 2), np.int32), np.full(k, 4, np.uint8), None, actions=np.full(k, 4, np.uint8),
                inj=np.ascontiguousarray(z["orbs"][sel]), want_final=True, log_cap=log_cap)
            assert np.array_equal(bits(reward), bits(z["reward"][sel]))
            assert np.array_equal(info & 0xFF, np.minimum(z["n_combos"][sel], 255))
            assert np.array_equal((info >> 8) & 0xFF, z["depth"][sel])
            assert np.array_equal((info >> 16) & 0xFF, np.minimum(z["consumed"][sel], 255))
            assert not (info >> 24).any()
            assert np.array_equal(final, z["final_boards"][sel])
            if log_cap:
                assert np.array_equal(log, z["combo_log"][sel])
    assert n >= 360


def test_kat_golden():
    cases = json.load(open(os.path.join(G, "kat.json")))["cases"]
    cfg = hc.cfg_from_oracle(orc.make_cfg(5, 6, skyfall_damage=False))
    for c in cases:
        if "actions" in c:
            continue
        b = np.array([c["board"]], np.int8)
        _, reward, info, final, log = hc.step_batch(cfg, b, np.zeros((1, 2), np.int32), np.full(1, 4, np.uint8), None,
                                                    actions=np.full(1, 4, np.uint8), want_final=True, log_cap=32)
        assert float(reward[0]).hex() == c["reward_hex"], c["name"]
        assert [[int(v) & 0xFF, int(v) >> 8] for v in log[0, :info[0] & 0xFF]] == c["combos"], c["name"]
        assert final[0].tolist() == c["final_board"], c["name"]


def test_reset_golden_and_philox():
    z = np.load(os.path.join(G, "reset_injected.npz"))
    for rows, cols in ((5, 6), (6, 7)):
        s = "%dx%d" % (rows, cols)
        ocfg = orc.make_cfg(rows, cols, reset_cap=0, seed=99)
        cfg = hc.cfg_from_oracle(ocfg)
        k = len(z["orbs_" + s])
        boards, fingers, flags, consumed = hc.reset(cfg, k, inj=z["orbs_" + s], fingers_in=np.tile([1, 2], (k, 1)))
        assert np.array_equal(boards, z["boards_" + s]) and np.array_equal(consumed, z["consumed_" + s])
        assert not flags.any() and (fingers == [1, 2]).all()
        # pure Philox reset: hostcore vs oracle, env ids straddling 2^32
        boards, fingers, flags, consumed = hc.reset(cfg, 300, env0=(1 << 32) - 100, step=5)
        for i in range(300):
            b, f, c, fl = orc.reset(ocfg, env=(1 << 32) - 100 + i, step=5)
            assert np.array_equal(b, boards[i]) and np.array_equal(f, fingers[i]) and c == consumed[i] and fl == flags[i]
            assert orc.cancel(b.copy()) == []


@pytest.mark.parametrize("rows,cols", [(4, 5), (5, 6), (6, 7)])
def test_obs_mask_matches_onehot(rows, cols):
    rng = np.random.default_rng(77 + rows)
    n = 5000
    boards = random_boards(rng, n, rows, cols, "jammer_empty")
    fingers = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1).astype(np.int32)
    assert np.array_equal(hc.obs(boards, fingers), orc.onehot(boards, fingers).astype(np.uint8))
    if rows == 4:
        return
    z = np.load(os.path.join(G, "onehot.npz"))
    s = "%dx%d" % (rows, cols)
    b = np.where(z["boards_" + s] > 6, -1, z["boards_" + s]).astype(np.int8)
    assert np.array_equal(hc.obs(b, z["fingers_" + s]), z["states_" + s])
    # the 4-bit kernels take the reference's boards as they are (orb types up to 9)
    raw = np.ascontiguousarray(z["boards_" + s], np.int8)
    assert raw.max() > 6
    assert np.array_equal(hc.obs(raw, z["fingers_" + s], orb_bits=4), z["states_" + s])
    wide = random_boards(rng, n, rows, cols, "all10")
    assert np.array_equal(hc.obs(wide, fingers, orb_bits=4), orc.onehot(wide, fingers).astype(np.uint8))


def test_sky_thresholds_match_float_rule():
    z = np.load(os.path.join(G, "random_orb.npz"))
    for i, p in enumerate(z["probs"]):
        cfg = hc.cfg_from_oracle(orc.make_cfg(skyfall=p))
        thr, valid = hc.sky_table(cfg)
        assert valid == sum(1 << j for j in range(7) if p[j] > 0)
        assert (z["xs"] == 0).any()                          # x = 0 with leading zero rates is the special case
        for j in range(len(z["xs"])):                        # random draws, 0, 2^32 - 1 and the threshold neighbours
            assert hc.orb_from_u32(cfg, int(z["xs"][j])) == z["orbs"][i, j], (i, j)


def test_uniform_skyfall_fast_path_is_the_float_rule():
    """orb_from_u32 replaces the threshold compares by one multiply when K leading types are equally likely; it must
    return exactly what the reference's rule returns (game.py:464-470: first i with cumsum[i] >= u and prob[i] > 0,
    u = x / 2^32), including at every threshold and its neighbours."""
    rng = np.random.default_rng(9)
    for K in range(2, 11):
        p = [1.0 / K] * K + [0.0] * (10 - K)
        cfg = hc.cfg_from_oracle(orc.make_cfg(skyfall=p), orb_bits=4 if K > 7 else None)
        assert hc.uniform_k(cfg) == K
        xs = {0, 1, 2, 2 ** 32 - 1, 2 ** 32 - 2, 2 ** 31, 2 ** 31 - 1, 2 ** 31 + 1}
        for k in range(1, K):
            t = (k << 32) // K
            xs.update({t - 2, t - 1, t, t + 1, t + 2})
        xs.update(int(v) for v in rng.integers(0, 2 ** 32, 2000))
        for x in sorted(xs):
            u, acc, want = x / 2.0 ** 32, 0.0, 9
            for i, pi in enumerate(p):
                acc += pi
                if acc >= u and pi > 0:
                    want = i
                    break
            assert hc.orb_from_u32(cfg, x) == want, (K, x)
    # not uniform, or not the leading types: the compare loop stays
    assert hc.uniform_k(hc.cfg_from_oracle(orc.make_cfg(skyfall=[0.5, 0.25, 0.25, 0, 0, 0, 0, 0, 0, 0]))) == 0
    assert hc.uniform_k(hc.cfg_from_oracle(orc.make_cfg(skyfall=[0, 0.5, 0.5, 0, 0, 0, 0, 0, 0, 0]))) == 0
    assert hc.uniform_k(hc.cfg_from_oracle(orc.make_cfg())) == 6            # the reference's default skyfall


def test_philox_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(50):
        ctr = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        key = [int(v) for v in rng.integers(0, 2 ** 32, 2)]
        assert hc.philox(ctr, key) == orc.philox4x32(ctr, key)
    assert hc.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
