"""The C oracle (oracle/ppad_oracle.c) against the fixtures the reference itself produced
(tests/golden/gen_golden.py).  CPU only.  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

import oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(x):
    return np.asarray(x, np.float64).view(np.uint64)


def test_kat_solved_boards_and_specials():
    cases = json.load(open(os.path.join(G, "kat.json")))["cases"]
    assert len(cases) >= 20
    cfg = orc.make_cfg(5, 6, skyfall_damage=False)
    for c in cases:
        if "actions" in c:
            continue
        r = orc.resolve(cfg, np.array(c["board"], np.int8), inj=np.zeros(0, np.uint8))
        assert r["combos"] == [tuple(x) for x in c["combos"]], c["name"]
        assert float(r["reward"]).hex() == c["reward_hex"], c["name"]
        assert r["final_board"].tolist() == c["final_board"], c["name"]
    # spot values quoted in SURVEY.md section 8c
    byname = {c["name"]: c for c in cases}
    assert byname["solved_00"]["reward"] == 5731.25
    assert byname["solved_15"]["reward"] == 23962.5
    assert byname["bridged_red_runs"]["combos"] == [[0, 6]]
    assert byname["row_col_row"]["combos"] == [[0, 7]]
    assert byname["two_disjoint_red_runs"]["reward"] == 3250.0


def test_kat_moves_then_skyfall():
    c = [c for c in json.load(open(os.path.join(G, "kat.json")))["cases"] if "actions" in c][0]
    board = np.array(c["board"], np.int8)
    finger = np.zeros(2, np.int32)
    for a in c["actions"]:
        assert orc.apply_action(board, finger, a) == 1
    assert board.tolist() == c["board_after_moves"] and finger.tolist() == c["finger_after_moves"]
    r = orc.resolve(orc.make_cfg(5, 6, True), board, inj=np.array(c["orbs"], np.uint8))
    assert r["consumed"] == c["consumed"] == 3 and r["depth"] == c["depth"]
    assert float(r["reward"]).hex() == c["reward_hex"] and r["reward"] == 1300.0
    assert r["final_board"].tolist() == c["final_board"]


@pytest.mark.parametrize("tag,rows,cols,team", [("5x6", 5, 6, None), ("6x7", 6, 7, None), ("4x5", 4, 5, None),
                                                 ("5x6_custom_team", 5, 6, "custom"), ("5x6_all10", 5, 6, None),
                                                 ("6x7_all10", 6, 7, None)])
def test_resolve_vectors(tag, rows, cols, team):
    z = np.load(os.path.join(G, "resolve_%s.npz" % tag))
    t = json.load(open(os.path.join(G, "custom_team.json"))) if team else None
    cfgs = {sd: orc.make_cfg(rows, cols, skyfall_damage=sd, team=t, cascade_cap=0) for sd in (0, 1)}
    n = len(z["boards"])
    assert n >= 360
    if tag.endswith("all10"):
        assert z["boards"].max() == 9 and z["orbs"].max() == 9      # poison, mortal poison and bomb orbs take part
    for i in range(n):
        r = orc.resolve(cfgs[int(z["skyfall_damage"][i])], z["boards"][i], inj=z["orbs"][i])
        assert bits(r["reward"]) == bits(z["reward"][i]), i
        assert (r["n_combos"], r["depth"], r["consumed"], r["flags"]) == (
            z["n_combos"][i], z["depth"][i], z["consumed"][i], 0), i
        assert np.array_equal(r["final_board"], z["final_boards"][i]), i
        log = np.zeros(64, np.uint16)
        for j, (c, k) in enumerate(r["combos"]):
            log[j] = c | (k << 8)
        assert np.array_equal(log, z["combo_log"][i]), i


@pytest.mark.parametrize("tag,rows,cols", [("5x6", 5, 6), ("6x7", 6, 7)])
def test_episode_vectors(tag, rows, cols):
    z = np.load(os.path.join(G, "episodes_%s.npz" % tag))
    cfg = orc.make_cfg(rows, cols, True, cascade_cap=0)
    E, L = z["actions"].shape
    for e in range(E):
        board = z["init_boards"][e].copy()
        finger = z["init_fingers"][e].copy()
        prev = 4
        for t in range(L):
            a = int(z["actions"][e, t])
            assert (orc.legal_mask(rows, cols, finger, prev) >> a) & 1   # reference sample() only returns legal moves
            assert orc.apply_action(board, finger, a) == 1
            prev = a
            assert np.array_equal(board, z["boards"][e, t]) and np.array_equal(finger, z["fingers"][e, t])
            r = orc.resolve(cfg, board, inj=z["orbs"][e, t])
            assert bits(r["reward"]) == bits(z["reward"][e, t])
            assert (r["n_combos"], r["depth"], r["consumed"]) == (z["n_combos"][e, t], z["depth"][e, t], z["consumed"][e, t])
        r = orc.resolve(cfg, board, inj=z["pass_orbs"][e])
        assert bits(r["reward"]) == bits(z["pass_reward"][e])
        # step(move) records reward 0, step('pass') records the damage (game.py:337,343,362)
        assert np.array_equal(z["recorded_rewards"][e][:-1], np.zeros(L)) and z["recorded_rewards"][e][-1] == z["pass_reward"][e]


def test_step_batch_matches_episode_vectors():
    z = np.load(os.path.join(G, "episodes_5x6.npz"))
    cfg = orc.make_cfg(5, 6, True, cascade_cap=0)
    E, L = z["actions"].shape
    boards = z["init_boards"].copy()
    fingers = z["init_fingers"].copy()
    prev = np.full(E, 4, np.uint8)
    moves = np.zeros(E, np.uint16)
    for t in range(L):
        a_out, reward, info, final = orc.step_batch(cfg, boards, fingers, prev, moves, actions=z["actions"][:, t],
                                                    inj=np.ascontiguousarray(z["orbs"][:, t]), want_final=True)
        assert np.array_equal(boards, z["boards"][:, t]) and np.array_equal(fingers, z["fingers"][:, t])
        assert np.array_equal(bits(reward), bits(z["reward"][:, t]))
        assert np.array_equal(info & 0xFF, z["n_combos"][:, t]) and np.array_equal((info >> 8) & 0xFF, z["depth"][:, t])
        assert np.array_equal((info >> 16) & 0xFF, z["consumed"][:, t]) and not (info >> 24).any()
    assert (moves == L).all() and np.array_equal(prev, z["actions"][:, -1])
    _, reward, info, _ = orc.step_batch(cfg, boards, fingers, prev, moves, actions=np.full(E, 4, np.uint8),
                                        inj=np.ascontiguousarray(z["pass_orbs"]))
    assert np.array_equal(bits(reward), bits(z["pass_reward"]))
    assert np.array_equal(boards, z["boards"][:, -1])   # a pass leaves env.board untouched (game.py:366)


def test_damage_vectors():
    z = np.load(os.path.join(G, "damage.npz"))
    team = json.load(open(os.path.join(G, "custom_team.json")))
    for tag, t in (("default", None), ("custom", team)):
        cfg = orc.make_cfg(team=t)
        for i in range(len(z["lens_" + tag])):
            n = int(z["lens_" + tag][i])
            combos = list(zip(z["colors_" + tag][i, :n].tolist(), z["counts_" + tag][i, :n].tolist()))
            assert bits(orc.damage(cfg, combos)) == bits(z["rewards_" + tag][i]), (tag, i)
    # per-combo constants quoted in SURVEY.md (a9)
    cfg = orc.make_cfg()
    assert orc.damage(cfg, [(0, 3)]) == 1300.0 and orc.damage(cfg, [(1, 3)]) == 2300.0
    assert orc.damage(cfg, [(3, 5)]) == 1650.0000000000002 and orc.damage(cfg, [(5, 4)]) == 0.0


def test_onehot_vectors():
    z = np.load(os.path.join(G, "onehot.npz"))
    for s in ("5x6", "6x7"):
        out = orc.onehot(z["boards_" + s], z["fingers_" + s])
        assert np.array_equal(out, z["states_" + s].astype(np.float32))


def test_reset_injected_vectors():
    z = np.load(os.path.join(G, "reset_injected.npz"))
    for rows, cols in ((5, 6), (6, 7)):
        s = "%dx%d" % (rows, cols)
        cfg = orc.make_cfg(rows, cols, reset_cap=0)
        for i in range(len(z["orbs_" + s])):
            board, finger, consumed, flags = orc.reset(cfg, inj=z["orbs_" + s][i], finger=[1, 2])
            assert np.array_equal(board, z["boards_" + s][i]) and consumed == z["consumed_" + s][i] and flags == 0
            assert finger.tolist() == [1, 2]
            assert orc.cancel(board.copy()) == []


def test_discount_vectors():
    z = np.load(os.path.join(G, "discount.npz"))
    for i in range(len(z["lens"])):
        n = int(z["lens"][i])
        for key, flag in (("out_log", True), ("out_lin", False)):
            got = orc.discount(z["rewards"][i, :n], z["gammas"][i], log10=flag)
            if flag:   # log10 comes from libm on both sides but numpy may use its own SIMD loop: 2 ulp
                assert np.allclose(got, z[key][i, :n], rtol=4e-16, atol=0)
            else:
                assert np.array_equal(bits(got), bits(z[key][i, :n]))


def test_legal_sets():
    z = np.load(os.path.join(G, "legal.npz"))
    for rows, cols in ((5, 6), (6, 7)):
        s = "%dx%d" % (rows, cols)
        for r in range(rows):
            for c in range(cols):
                for prev in range(5):
                    assert orc.legal_mask(rows, cols, [r, c], prev) == z["sample_" + s][r, c, prev]
                # illegal_actions (utils.py:73-83) = off-board moves, with its elif quirk on 1-wide boards only
                off = (r == 0) * 1 | (r == rows - 1) * 2 | (c == 0) * 4 | (c == cols - 1) * 8
                assert z["illegal_" + s][r, c] == off


def test_random_orb_thresholds():
    z = np.load(os.path.join(G, "random_orb.npz"))
    for i, p in enumerate(z["probs"]):
        for j, x in enumerate(z["xs"]):
            assert orc.orb_from_u32(p, int(x)) == z["orbs"][i, j]


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert orc.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
