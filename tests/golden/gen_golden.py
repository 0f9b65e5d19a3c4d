#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/gen_golden.py          # needs /root/reference; writes tests/golden/*.npz|json

The reference has no tests or vectors of its own, so these fixtures -- outputs
of the reference's own Python game logic on seeded inputs -- are what pins the C oracle
(oracle/ppad_oracle.c), which in turn checks the CUDA kernels on the GPU box where the reference
does not exist.  Inputs are produced with numpy's PCG64 (seeded); skyfall is injected (patch 3 of
ref_harness.py), so nothing depends on the reference's wall-clock-seeded Mersenne Twister except
the episode fixtures, which seed ``random`` explicitly.
"""
import json
import os
import random
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

INT2ENGLISH = {v: k for k, v in rh.ENGLISH2INT.items()}
ACTIONS = ["up", "down", "left", "right", "pass"]
LOGCAP = 64


def random_walk_board(rng, rows, cols, ncolors, max_moves):
    """A full random board, then a random finger walk (swaps), like a played position."""
    board = rng.integers(0, ncolors, size=(rows, cols)).astype(np.int8)
    f = [int(rng.integers(rows)), int(rng.integers(cols))]
    for _ in range(int(rng.integers(0, max_moves + 1))):
        d = int(rng.integers(4))
        t = [f[0] + (d == 1) - (d == 0), f[1] + (d == 3) - (d == 2)]
        if 0 <= t[0] < rows and 0 <= t[1] < cols:
            board[f[0], f[1]], board[t[0], t[1]] = board[t[0], t[1]], board[f[0], f[1]]
            f = t
    return board, f


def pack_log(per_iter):
    flat = [c | (n << 8) for it in per_iter for (c, n) in it]
    out = np.zeros(LOGCAP, np.uint16)
    out[:min(len(flat), LOGCAP)] = flat[:LOGCAP]
    return out, len(flat)


def gen_kat(ref):
    env5 = rh.make_env(5, 6, skyfall_damage=False)
    cases = []
    boards = [np.array(b) for b in ref.solved_boards.boards]
    names = ["solved_%02d" % i for i in range(len(boards))]
    extra = {
        "two_disjoint_red_runs": [[0, 0, 0, 1, 2, 3], [1, 2, 1, 2, 3, 1], [2, 1, 2, 0, 0, 0], [1, 2, 3, 1, 2, 3], [2, 3, 1, 2, 3, 1]],
        "bridged_red_runs": [[0, 0, 0, 1, 2, 3], [1, 2, 0, 0, 3, 1], [2, 1, 2, 0, 0, 0], [1, 2, 3, 1, 2, 3], [2, 3, 1, 2, 3, 1]],
        "row_col_row": [[0, 0, 0, 1, 2, 3], [1, 2, 0, 2, 3, 1], [2, 1, 0, 0, 0, 2], [1, 2, 3, 1, 2, 3], [2, 3, 1, 2, 3, 1]],
        "island_min_is_unmatched": [[0, 1, 1, 1, 2, 3], [0, 2, 3, 4, 5, 4], [3, 0, 0, 0, 4, 5], [0, 4, 5, 3, 2, 1], [0, 0, 0, 2, 3, 4]],
        "all_one_colour": [[2] * 6] * 5,
        "no_match": [[0, 1, 2, 3, 4, 5], [1, 2, 3, 4, 5, 0], [2, 3, 4, 5, 0, 1], [3, 4, 5, 0, 1, 2], [4, 5, 0, 1, 2, 3]],
        "with_empties": [[-1, -1, -1, 1, 2, 3], [0, 0, 0, 2, 3, 1], [2, 1, -1, 0, 0, 0], [1, 2, -1, 1, 2, 3], [2, 3, -1, 2, 3, 1]],
        "jammers": [[6, 6, 6, 1, 2, 3], [0, 0, 0, 2, 3, 1], [2, 1, 5, 5, 5, 5], [1, 2, 3, 1, 2, 3], [2, 3, 1, 2, 3, 1]],
    }
    for k, v in extra.items():
        names.append(k)
        boards.append(np.array(v))
    for name, b in zip(names, boards):
        r = rh.resolve(env5, b, False, [])
        combos = [list(c) for it in r["combos"] for c in it]
        cases.append(dict(name=name, board=b.astype(int).tolist(), combos=combos,
                          reward_hex=float(r["reward"]).hex(), reward=r["reward"],
                          final_board=r["final_board"].astype(int).tolist()))
    # the survey's move + skyfall known answer (SURVEY.md section 8c last row)
    env = rh.make_env(5, 6, skyfall_damage=True, board=np.array(extra["two_disjoint_red_runs"]), finger=[0, 0])
    rr = random.Random(42)
    orbs = [rr.randrange(6) for _ in range(64)]
    env.step("right")
    env.step("down")
    r = rh.resolve(env, env.board, True, orbs)
    cases.append(dict(name="move_right_down_then_pass_skyfall", board=np.array(extra["two_disjoint_red_runs"]).tolist(),
                      actions=[3, 1], orbs=orbs, board_after_moves=env.board.astype(int).tolist(),
                      finger_after_moves=[int(v) for v in env.finger],
                      combos=[list(c) for it in r["combos"] for c in it], consumed=r["consumed"], depth=r["depth"],
                      reward_hex=float(r["reward"]).hex(), reward=r["reward"],
                      final_board=r["final_board"].astype(int).tolist()))
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/gen_golden.py", cases=cases), f, indent=0)
    print("kat.json:", len(cases), "cases")


def gen_resolve(rows, cols, n_walk, n_adv, seed, tag, team=None, skyfall_rates=None, ncolors=6):
    """ncolors = 10: boards and injected skyfall over ALL orb types of game.py:236-246 (poison, mortal poison, bomb)"""
    rng = np.random.default_rng(seed)
    env_on = rh.make_env(rows, cols, skyfall_damage=True, team=team)
    S = 512
    recs = []
    t0 = time.time()
    i = 0
    while len(recs) < n_walk + n_adv:
        adversarial = len(recs) >= n_walk
        if adversarial:
            ncol = int(rng.integers(2, 5))
            pal = np.arange(ncol) if ncolors <= 6 else rng.permutation(np.array([0, 3, 5, 6, 7, 8, 9]))[:ncol]
            board = pal[rng.integers(0, ncol, size=(rows, cols))].astype(np.int8)
            p = rng.dirichlet(np.ones(ncol) * 0.7)
            orbs = pal[rng.choice(ncol, size=S, p=p)].astype(np.uint8)
        else:
            board, _ = random_walk_board(rng, rows, cols, ncolors, 60)
            orbs = rng.integers(0, ncolors, size=S).astype(np.uint8)
        sd = bool(rng.random() < 0.85) or adversarial
        i += 1
        try:
            r = rh.resolve(env_on, board, sd, orbs)
        except IndexError:
            continue  # cascade outlived the injected sequence: not a usable vector
        log, nlog = pack_log(r["combos"])
        if nlog > LOGCAP:
            continue
        recs.append((board, orbs[:256] if r["consumed"] <= 256 else None, sd, r, log))
        if recs[-1][1] is None:
            recs.pop()
    K = len(recs)
    np.savez_compressed(
        os.path.join(HERE, "resolve_%s.npz" % tag),
        boards=np.stack([r[0] for r in recs]).astype(np.int8),
        orbs=np.stack([r[1] for r in recs]).astype(np.uint8),
        skyfall_damage=np.array([r[2] for r in recs], np.uint8),
        reward=np.array([r[3]["reward"] for r in recs], np.float64),
        n_combos=np.array([sum(len(it) for it in r[3]["combos"]) for r in recs], np.int32),
        depth=np.array([r[3]["depth"] for r in recs], np.int32),
        consumed=np.array([r[3]["consumed"] for r in recs], np.int32),
        final_boards=np.stack([r[3]["final_board"] for r in recs]).astype(np.int8),
        combo_log=np.stack([r[4] for r in recs]))
    d = np.array([r[3]["depth"] for r in recs])
    print("resolve_%s.npz: %d cases (%d tried) in %.0fs; depth hist %s max consumed %d" % (
        tag, K, i, time.time() - t0, np.bincount(d).tolist(), max(r[3]["consumed"] for r in recs)))


def gen_episodes(rows, cols, n_ep, length, seed, tag):
    """Reference reset() + PlayerAction.sample() walks; every step is also evaluated with
    calculate_reward(board=env.board, skyfall_damage=True) (the pattern of simulation.py:117-119)."""
    rng = np.random.default_rng(seed)
    S = 64
    out = dict(init_boards=[], init_fingers=[], actions=[], boards=[], fingers=[], orbs=[], reward=[],
               n_combos=[], depth=[], consumed=[], pass_reward=[], pass_orbs=[], recorded_rewards=[])
    for e in range(n_ep):
        rh.set_clock(seed * 1000 + e)   # game.py:72,302 seed the global MT from "datetime.now()"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = rh.load().game.PAD(dim_row=rows, dim_col=cols, skyfall_damage=True)
        out["init_boards"].append(env.board.astype(np.int8).copy())
        out["init_fingers"].append(env.finger.astype(np.int32).copy())
        acts, boards, fingers, orbs_l, rew, nc, dep, con = [], [], [], [], [], [], [], []
        for t in range(length):
            env.action_space.finger = env.finger
            a = env.action_space.sample()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                env.step(a)
            orbs = rng.integers(0, 6, size=S).astype(np.uint8)
            r = rh.resolve(env, env.board, True, orbs)
            acts.append(ACTIONS.index(a))
            boards.append(env.board.astype(np.int8).copy())
            fingers.append(env.finger.astype(np.int32).copy())
            orbs_l.append(orbs)
            rew.append(r["reward"])
            nc.append(sum(len(it) for it in r["combos"]))
            dep.append(r["depth"])
            con.append(r["consumed"])
        porbs = rng.integers(0, 6, size=S).astype(np.uint8)
        inj = rh.InjectedSkyfall(porbs)
        env.random_orb = inj
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env.step("pass")
        out["pass_reward"].append(float(env.rewards[-1]))
        out["pass_orbs"].append(porbs)
        out["recorded_rewards"].append(np.array(env.rewards, np.float64))
        assert len(env.observations) == length + 1 and len(env.actions) == length + 1
        for k, v in (("actions", acts), ("boards", boards), ("fingers", fingers), ("orbs", orbs_l),
                     ("reward", rew), ("n_combos", nc), ("depth", dep), ("consumed", con)):
            out[k].append(np.array(v))
    np.savez_compressed(os.path.join(HERE, "episodes_%s.npz" % tag), **{k: np.stack(v) for k, v in out.items()})
    print("episodes_%s.npz: %d episodes x %d steps" % (tag, n_ep, length))


def gen_onehot(ref):
    rng = np.random.default_rng(7)
    res = {}
    import types
    for rows, cols in ((5, 6), (6, 7)):
        B = 64
        boards = rng.integers(-1, 8, size=(B, rows, cols)).astype(int)
        fingers = np.stack([rng.integers(0, rows, B), rng.integers(0, cols, B)], 1).astype(int)
        me = types.SimpleNamespace(st_shape=(rows, cols, 7))
        states = ref.agent01.Agent01.board_finger_to_state(me, boards, fingers)
        assert states.dtype == np.float64 and set(np.unique(states)) <= {0.0, 1.0}
        res["boards_%dx%d" % (rows, cols)] = boards.astype(np.int8)
        res["fingers_%dx%d" % (rows, cols)] = fingers.astype(np.int32)
        res["states_%dx%d" % (rows, cols)] = states.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "onehot.npz"), **res)
    print("onehot.npz")


def gen_reset():
    rng = np.random.default_rng(11)
    res = {}
    for rows, cols in ((5, 6), (6, 7)):
        K, S = 48, rows * cols * 40
        streams, boards, consumed = [], [], []
        env = rh.make_env(rows, cols)
        for _ in range(K):
            orbs = rng.integers(0, 6, size=S).astype(np.uint8)
            inj = rh.InjectedSkyfall(orbs)
            env.random_orb = inj
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                b, f = env.reset(finger=np.array([1, 2]))
            streams.append(orbs)
            boards.append(b.astype(np.int8).copy())
            consumed.append(inj.cursor)
        res["orbs_%dx%d" % (rows, cols)] = np.stack(streams)
        res["boards_%dx%d" % (rows, cols)] = np.stack(boards)
        res["consumed_%dx%d" % (rows, cols)] = np.array(consumed, np.int32)
    np.savez_compressed(os.path.join(HERE, "reset_injected.npz"), **res)
    print("reset_injected.npz: attempts hist", np.bincount(res["consumed_5x6"] // 30).tolist())


CUSTOM_TEAM = [dict(name="A", color_1="red", color_2="red", atk=1234, hp=1, rcv=1),
               dict(name="B", color_1="blue", color_2="red", atk=987.5, hp=1, rcv=1),
               dict(name="C", color_1="green", color_2=None, atk=1511, hp=1, rcv=1),
               dict(name="D", color_1="light", color_2="dark", atk=2048, hp=1, rcv=1),
               dict(name="E", color_1="dark", color_2="green", atk=777, hp=1, rcv=1)]


def gen_damage():
    rng = np.random.default_rng(13)
    res = {}
    for tag, team in (("default", None), ("custom", CUSTOM_TEAM)):
        env = rh.make_env(5, 6, team=team)
        K, M = 4000, 24
        colors = np.full((K, M), -1, np.int32)
        counts = np.zeros((K, M), np.int32)
        lens = rng.integers(0, M + 1, size=K)
        rewards = np.zeros(K, np.float64)
        for i in range(K):
            n = int(lens[i])
            colors[i, :n] = rng.integers(0, 7, size=n)
            counts[i, :n] = rng.integers(3, 31, size=n)
            combos = [dict(color=INT2ENGLISH[int(colors[i, j])], N=np.float64(counts[i, j]), enhanced=None, shape=None)
                      for j in range(n)]
            rewards[i] = float(env.damage(combos))
        res.update({"colors_" + tag: colors, "counts_" + tag: counts, "lens_" + tag: lens.astype(np.int32),
                    "rewards_" + tag: rewards})
    np.savez_compressed(os.path.join(HERE, "damage.npz"), **res)
    with open(os.path.join(HERE, "custom_team.json"), "w") as f:
        json.dump(CUSTOM_TEAM, f)
    print("damage.npz")


def gen_discount(ref):
    rng = np.random.default_rng(17)
    K, L = 200, 40
    rewards = np.zeros((K, L))
    lens = rng.integers(1, L + 1, size=K)
    out_log, out_lin = np.zeros((K, L)), np.zeros((K, L))
    gammas = rng.uniform(0.5, 0.99, size=K)
    for i in range(K):
        n = int(lens[i])
        r = np.where(rng.random(n) < 0.3, rng.uniform(0, 40000, n), 0.0)
        if i % 10 == 0:
            r[:] = 0
        rewards[i, :n] = r
        out_log[i, :n] = ref.agent_utils.discount(list(r), gammas[i], log10=True)
        out_lin[i, :n] = ref.agent_utils.discount(list(r), gammas[i], log10=False)
    np.savez_compressed(os.path.join(HERE, "discount.npz"), rewards=rewards, lens=lens.astype(np.int32),
                        gammas=gammas, out_log=out_log, out_lin=out_lin)
    print("discount.npz")


def gen_legal(ref):
    """illegal_actions (utils.py:73-83) and the support of PlayerAction.sample (player.py:38-66)."""
    res = {}
    for rows, cols in ((5, 6), (6, 7)):
        illegal = np.zeros((rows, cols), np.uint8)
        sample_mask = np.zeros((rows, cols, 5), np.uint8)
        for r in range(rows):
            for c in range(cols):
                f = np.array([r, c])
                s = ref.pad_utils.illegal_actions(f, dim_row=rows, dim_col=cols)
                illegal[r, c] = sum(1 << ACTIONS.index(a) for a in s)
                for prev in range(5):
                    pa = ref.player.PlayerAction(finger=f, dim_row=rows, dim_col=cols)
                    pa.previous_action = ACTIONS[prev]
                    random.seed(r * 100 + c * 10 + prev)
                    m = 0
                    for _ in range(300):
                        m |= 1 << ACTIONS.index(pa.sample())
                    sample_mask[r, c, prev] = m
        res["illegal_%dx%d" % (rows, cols)] = illegal
        res["sample_%dx%d" % (rows, cols)] = sample_mask
    np.savez_compressed(os.path.join(HERE, "legal.npz"), **res)
    print("legal.npz")


def gen_random_orb(ref):
    """random_orb (game.py:458-470) with random.random() pinned to x / 2^32."""
    rng = np.random.default_rng(19)
    probs = [[1 / 6] * 6 + [0] * 4,
             [0.3, 0.2, 0.1, 0.1, 0.1, 0.1, 0.1, 0, 0, 0],
             [0.5, 0, 0.25, 0, 0.25, 0, 0, 0, 0, 0],
             [0, 0, 0, 0.125, 0.125, 0.75, 0, 0, 0, 0],
             [0.05, 0.15, 0.2, 0.1, 0.3, 0.1, 0.1, 0, 0, 0]]
    xs = np.concatenate([rng.integers(0, 2 ** 32, size=3000, dtype=np.uint64),
                         np.array([0, 1, 2 ** 32 - 1, 2 ** 31, 715827882, 715827883, 1431655765, 1431655766,
                                   2147483648, 2863311530, 2863311531, 3579139413, 3579139414], np.uint64)])
    # add the exact threshold neighbours of every cumulative sum
    extra = []
    for p in probs:
        c = 0.0
        for v in p:
            c += v
            t = int(c * 2 ** 32)
            extra += [max(0, min(2 ** 32 - 1, t + d)) for d in (-1, 0, 1)]
    xs = np.concatenate([xs, np.array(extra, np.uint64)])
    out = np.zeros((len(probs), len(xs)), np.uint8)
    real = ref.game.random.random
    try:
        for i, p in enumerate(probs):
            for j, x in enumerate(xs):
                ref.game.random.random = lambda x=x: int(x) / 4294967296.0
                out[i, j] = ref.game.PAD.random_orb(p)
    finally:
        ref.game.random.random = real
    np.savez_compressed(os.path.join(HERE, "random_orb.npz"), probs=np.array(probs, np.float64),
                        xs=xs.astype(np.uint32), orbs=out)
    print("random_orb.npz")


def gen_mt_episodes():
    """Whole episodes of the reference driven by its OWN Mersenne Twister (no injection): seeded through the
    patched clock, so that PAD(rng='reference', clock=...) of ppad_b200 can be asked to replay them draw for draw."""
    out = {}
    for rows, cols in ((5, 6), (6, 7)):
        recs = dict(seed=[], init_board=[], init_finger=[], actions=[], boards=[], fingers=[], pass_reward=[],
                    extra_reward=[], board2=[], finger2=[], actions2=[], pass_reward2=[], mt_after=[])
        for s in range(12):
            seed = 7000 + 100 * rows + s
            rh.set_clock(seed)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                env = rh.load().game.PAD(dim_row=rows, dim_col=cols)
                recs["seed"].append(seed)
                recs["init_board"].append(env.board.astype(np.int8).copy())
                recs["init_finger"].append(env.finger.astype(np.int32).copy())
                acts, boards, fingers = [], [], []
                for _ in range(25):
                    a = env.action_space.sample()
                    env.step(a)
                    acts.append(ACTIONS.index(a))
                    boards.append(env.board.astype(np.int8).copy())
                    fingers.append(env.finger.astype(np.int32).copy())
                env.step("pass")
                recs["actions"].append(np.array(acts))
                recs["boards"].append(np.stack(boards))
                recs["fingers"].append(np.stack(fingers))
                recs["pass_reward"].append(float(env.rewards[-1]))
                recs["extra_reward"].append(float(env.calculate_reward(board=env.board, skyfall_damage=True)))
                rh.set_clock(seed + 50)
                env.reset()
                recs["board2"].append(env.board.astype(np.int8).copy())
                recs["finger2"].append(env.finger.astype(np.int32).copy())
                acts = []
                for _ in range(6):
                    a = env.action_space.sample(include_pass=True)
                    env.step(a)
                    acts.append(ACTIONS.index(a))
                env.step("pass")
                recs["actions2"].append(np.array(acts))
                recs["pass_reward2"].append(float(env.rewards[-1]))
                recs["mt_after"].append(random.random())       # where the Mersenne Twister stands at the end
        for k, v in recs.items():
            out["%s_%dx%d" % (k, rows, cols)] = np.stack(v) if isinstance(v[0], np.ndarray) else np.array(v)
    np.savez_compressed(os.path.join(HERE, "mt_episodes.npz"), **out)
    print("mt_episodes.npz")


def gen_agent_act(ref):
    """Agent01.act (agent01.py:157-218) called UNBOUND with a stub self (no TensorFlow): the -inf masks for every
    (finger, last action), argmax incl. ties, the support of the epsilon-greedy branch and the Boltzmann probabilities
    np.random.choice is handed (captured), for q-values the test feeds to ppad_select_actions / ppad_policy_step."""
    import types
    A = ref.agent01.Agent01
    rng = np.random.default_rng(77)
    rows, cols = 5, 6

    def make_stub(q, last):
        stub = types.SimpleNamespace(st_shape=(rows, cols, 7), last_action=int(last))
        stub.board_finger_to_state = lambda b, f: A.board_finger_to_state(stub, b, f)
        stub.predict_from_state = lambda state, model: np.array([q], dtype=np.float32)
        stub.action_verb = lambda a: A.action_verb(stub, a)
        return stub

    verbs = ['up', 'down', 'left', 'right', 'pass']
    board = np.zeros((rows, cols), int)
    # (1) masks + argmax: every finger x last action, q = descending so that the first allowed action wins, then random q
    fingers, lasts, qs, acts = [], [], [], []
    for r in range(rows):
        for c in range(cols):
            for last in range(5):
                for k in range(4):
                    q = np.array([5, 4, 3, 2, 1], np.float32) if k == 0 else \
                        (np.round(rng.standard_normal(5) * 2) * 500).astype(np.float32) if k == 1 else \
                        (rng.standard_normal(5) * 1000).astype(np.float32)
                    stub = make_stub(q.copy(), last)
                    a = verbs.index(A.act(stub, board, np.array([r, c]), 'A', 'max'))
                    assert stub.last_action == a
                    fingers.append([r, c]); lasts.append(last); qs.append(q); acts.append(a)
    # (2) epsilon-greedy support: epsilon = 1 -> uniform over the actions that are not -inf
    supp = []
    np.random.seed(5)
    for r in range(rows):
        for c in range(cols):
            for last in range(5):
                seen = 0
                for _ in range(60):
                    stub = make_stub(np.zeros(5, np.float32), last)
                    seen |= 1 << verbs.index(A.act(stub, board, np.array([r, c]), 'A', 'max', epsilon=1.0))
                supp.append(seen)
    # (3) Boltzmann: capture the probability vector handed to np.random.choice
    bf, bl, bq, bb, bp = [], [], [], [], []
    captured = {}
    real_choice = np.random.choice

    def fake_choice(n, p=None):
        captured['p'] = np.array(p, np.float64)
        return int(np.argmax(p))
    ref.agent01.np.random.choice = fake_choice
    try:
        for _ in range(600):
            r, c, last = int(rng.integers(rows)), int(rng.integers(cols)), int(rng.integers(5))
            beta = float(10 ** rng.uniform(-4, -2))
            q = (rng.standard_normal(5) * 1000).astype(np.float32)
            stub = make_stub(q.copy(), last)
            A.act(stub, board, np.array([r, c]), 'A', 'boltzmann', beta)
            bf.append([r, c]); bl.append(last); bq.append(q); bb.append(beta); bp.append(captured['p'])
    finally:
        ref.agent01.np.random.choice = real_choice
    np.savez_compressed(os.path.join(HERE, "agent_act.npz"), fingers=np.array(fingers), lasts=np.array(lasts), qs=np.array(qs),
                        acts=np.array(acts), eps_support=np.array(supp), b_fingers=np.array(bf), b_lasts=np.array(bl),
                        b_qs=np.array(bq), b_betas=np.array(bb), b_probs=np.array(bp))
    print("agent_act.npz", len(acts), len(supp), len(bp))


def gen_islands(ref):
    """pad_utils.detect_island (utils.py:86-111) run as it is: random boards (few colours, some empties), random seeds,
    orb types that do and do not match the seed cell, island arrays that start clear or partly marked."""
    rng = np.random.default_rng(31)
    out = {}
    for rows, cols, n in ((5, 6, 1500), (6, 7, 800)):
        boards = rng.integers(-1, 4, size=(n, rows, cols))
        seeds = np.stack([rng.integers(0, rows, n), rng.integers(0, cols, n)], 1)
        types = np.where(rng.random(n) < 0.8, boards[np.arange(n), seeds[:, 0], seeds[:, 1]], rng.integers(-1, 5, n))
        pre = (rng.random((n, rows, cols)) < 0.15) & (rng.random(n) < 0.4)[:, None, None]
        islands = np.zeros((n, rows, cols), np.uint8)
        for i in range(n):
            isl = pre[i].astype(float).copy()
            ref.pad_utils.detect_island(boards[i], isl, int(seeds[i, 0]), int(seeds[i, 1]), int(types[i]))
            islands[i] = isl != 0
        tag = "%dx%d" % (rows, cols)
        out.update({"boards_" + tag: boards.astype(np.int8), "seeds_" + tag: seeds.astype(np.int8),
                    "types_" + tag: types.astype(np.int8), "pre_" + tag: pre.astype(np.uint8), "islands_" + tag: islands})
    np.savez_compressed(os.path.join(HERE, "islands.npz"), **out)
    print("islands.npz", {k: v.shape for k, v in out.items() if k.startswith("islands")})


def gen_2sla(ref):
    """model_simulation_2sla (projects/simulation.py:71-183), the reference driver's data-collection policy, run
    UNMODIFIED: the two function definitions are compiled from the reference's source file (the module itself is a
    training script that needs TensorFlow) into a namespace that provides what the script's globals provide.  Patches,
    none of them game or search logic: env.reset() adopts prepared boards; random.shuffle puts the candidate moves into
    the fixed order up, down, left, right (the reference's order is random and only decides ties); the agent is
    Agent01.act UNBOUND on a stub whose network is a fixed integer-weight linear map of the one-hot observation."""
    import ast
    import types
    src = open(os.path.join(rh.REFERENCE_ROOT, "projects", "simulation.py")).read()
    tree = ast.parse(src)
    funcs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("model_simulation", "model_simulation_2sla")]
    assert len(funcs) == 2
    A = ref.agent01.Agent01
    rows, cols = 5, 6
    rng = np.random.default_rng(2024)
    W = rng.integers(-9, 10, size=(rows * cols * 7, 5)).astype(np.int64)
    W[:, 4] -= 1                                 # a slight reluctance to pass
    order = ['up', 'down', 'left', 'right']
    fake_random = types.SimpleNamespace(shuffle=lambda lst: lst.sort(key=order.index))
    env = ref.game.PAD(skyfall_damage=False)
    starts = []
    real_reset = env.reset

    def reset_from_list(board=None, finger=None):
        b, f = starts.pop(0)
        return real_reset(board=b, finger=f)
    env.reset = reset_from_list
    stub = types.SimpleNamespace(st_shape=(rows, cols, 7), last_action=4)
    stub.board_finger_to_state = lambda b, f: A.board_finger_to_state(stub, b, f)
    stub.predict_from_state = lambda state, model: (state.reshape(1, -1).astype(np.int64) @ W).astype(np.float32)
    stub.action_verb = lambda a: A.action_verb(stub, a)
    stub.act = lambda board, finger, model='A', method='max', beta=None, epsilon=0: A.act(stub, board, finger, model, method, beta, epsilon)
    ns = dict(env=env, ppad=ref.ppad, illegal_actions=ref.pad_utils.illegal_actions, random=fake_random, np=np,
              ACTION2ID={'up': 0, 'down': 1, 'left': 2, 'right': 3, 'pass': 4},
              ID2ACTION={0: 'up', 1: 'down', 2: 'left', 3: 'right', 4: 'pass'},
              NON_PASS_ACTIONS={'up', 'down', 'left', 'right'})
    exec(compile(ast.Module(body=funcs, type_ignores=[]), "simulation.py", "exec"), ns)
    out = dict(W=W, boards=[], fingers=[], n_actions=[], actions=[], final_reward=[], discounted=[], obs_boards=[], obs_fingers=[])
    E, max_ep, max_la, gamma = 160, 12, 5, 0.95
    verbs = ['up', 'down', 'left', 'right', 'pass']
    walk = np.random.default_rng(9)
    for e in range(E):
        b, f = random_walk_board(walk, rows, cols, 6 if e % 3 else 4, 30)
        b, f = b.astype(int), np.array(f)
        starts.append((b.copy(), f.copy()))
        sars, n_ep = ns["model_simulation_2sla"](stub, 1, gamma, log10_reward=False, policy='max', beta=None,
                                                  max_episode_len=max_ep, max_lookahead_len=max_la)
        assert n_ep == 1 and stub.last_action == 4
        acts = [verbs.index(a) for a in env.actions]
        pad = max_ep + max_la + 2
        out["boards"].append(b); out["fingers"].append(f); out["n_actions"].append(len(acts))
        out["actions"].append(acts + [255] * (pad - len(acts)))
        out["final_reward"].append(float(env.rewards[-1]))
        disc = [float(s[2]) for s in sars]
        out["discounted"].append(disc + [0.0] * (pad - len(disc)))
        ob = np.stack([o[0] for o in env.observations]); of = np.stack([o[1] for o in env.observations])
        out["obs_boards"].append(np.concatenate([ob, np.zeros((pad - len(ob), rows, cols), ob.dtype)]))
        out["obs_fingers"].append(np.concatenate([of, np.zeros((pad - len(of), 2), of.dtype)]))
    res = {k: np.array(v) for k, v in out.items()}
    res.update(max_episode_len=max_ep, max_lookahead_len=max_la, gamma=gamma)
    np.savez_compressed(os.path.join(HERE, "sim_2sla.npz"), **res)
    n = res["n_actions"]
    print("sim_2sla.npz episodes", E, "actions per episode min/mean/max", n.min(), n.mean(), n.max(),
          "nonzero rewards", int((res["final_reward"] > 0).sum()))


def main():
    ref = rh.load()
    if "--agent" in sys.argv:      # only the fixtures of the callers either side of the step
        gen_agent_act(ref)
        gen_2sla(ref)
        gen_islands(ref)
        return
    if "--all10" in sys.argv:      # only the fixtures of the 4-bit (orb types 0..9) kernels
        gen_resolve(5, 6, 700, 150, 205, "5x6_all10", ncolors=10)
        gen_resolve(6, 7, 300, 60, 206, "6x7_all10", ncolors=10)
        return
    gen_agent_act(ref)
    gen_2sla(ref)
    gen_islands(ref)
    gen_kat(ref)
    gen_onehot(ref)
    gen_damage()
    gen_discount(ref)
    gen_legal(ref)
    gen_random_orb(ref)
    gen_reset()
    gen_mt_episodes()
    gen_episodes(5, 6, 24, 30, 101, "5x6")
    gen_episodes(6, 7, 10, 24, 102, "6x7")
    gen_resolve(5, 6, 2500, 300, 201, "5x6")
    gen_resolve(6, 7, 1000, 150, 202, "6x7")
    gen_resolve(5, 6, 400, 0, 203, "5x6_custom_team", team=CUSTOM_TEAM)
    gen_resolve(4, 5, 500, 60, 204, "4x5")
    gen_resolve(5, 6, 700, 150, 205, "5x6_all10", ncolors=10)
    gen_resolve(6, 7, 300, 60, 206, "6x7_all10", ncolors=10)


if __name__ == "__main__":
    main()
