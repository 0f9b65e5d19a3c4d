"""BASELINE.json configs[2..4] as parity / property tests on one GPU (the bench line is configs[1]).
   configs[2] cascade-heavy stress: 16 M boards, max-length finger paths, depth / combo histograms
   configs[3] 7x6 boards, smart-data style generation from pre-solved boards
   configs[4] Boltzmann-policy rollout with one-hot obs feeding experience replay + stats all-gather"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import oracle as orc  # noqa: E402

THREADS = os.cpu_count() or 1


@pytest.fixture(scope="module")
def pb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ppad_b200
    ppad_b200._cabi.lib()
    return ppad_b200


def bits(x):
    return np.ascontiguousarray(x, np.float64).view(np.uint64)


def test_config2_cascade_heavy_16m_boards(pb):
    n, rows, cols, L, S, cap = 1 << 24, 5, 6, 200, 64, 32
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    # adversarial generator: boards of 3 colours, skyfall biased to colour 0 -> deep cascades
    boards = torch.randint(0, 3, (n, rows, cols), device=dev, generator=g, dtype=torch.int64)
    fingers = torch.stack([torch.randint(0, rows, (n,), device=dev, generator=g),
                           torch.randint(0, cols, (n,), device=dev, generator=g)], 1)
    paths = torch.randint(0, 4, (n, L), device=dev, generator=g, dtype=torch.uint8)
    u = torch.rand((n, S), device=dev, generator=g)
    inj = torch.where(u < 0.6, 0, torch.where(u < 0.8, 1, 2)).to(torch.uint8)
    env = pb.BatchedPAD(n, rows, cols, seed=11, cascade_cap=cap)
    env.set_state(boards, fingers)
    del boards, u
    packed = env.pack_paths(paths)
    slab = (env._slab(inj)[0], S)
    out = env.step(None, eval_moves=True, injected=slab, paths=packed, step=1)
    torch.cuda.synchronize()
    combos, depth, consumed, flags = env.split_info(out.info)
    # exact check on a strided subset through the oracle
    idx = torch.arange(0, n, 97, device=dev)
    k = idx.numel()
    sub = pb.BatchedPAD(k, rows, cols, seed=11, cascade_cap=cap)
    sub.state.copy_(env.state[idx])
    st = sub.get_state()
    # reconstruct the pre-path boards of the subset: regenerate with the same generator is not possible, so
    # replay the oracle forward from the recorded inputs instead
    g2 = torch.Generator(device=dev).manual_seed(7)
    b0 = torch.randint(0, 3, (n, rows, cols), device=dev, generator=g2, dtype=torch.int64)[idx].cpu().numpy().astype(np.int8)
    f0 = fingers[idx].cpu().numpy().astype(np.int32)
    ob = [b0.copy(), f0.copy(), np.full(k, 4, np.uint8), np.zeros(k, np.uint16)]
    ocfg = orc.make_cfg(rows, cols, True, seed=11, cascade_cap=cap)
    a0, r0, i0, _ = orc.path_batch(ocfg, *ob, paths[idx].cpu().numpy(), inj=inj[idx].cpu().numpy(), env0=0, step=1,
                                   threads=THREADS)
    # env ids of the subset are idx, not 0..k: only Philox draws depend on them, i.e. boards whose slab ran out
    exhausted = ((i0 >> 24) & orc.FLAG_SKYFALL_EXHAUSTED) != 0
    got_info = out.info[idx].cpu().numpy().view(np.uint32)
    got_exh = ((got_info >> 24) & orc.FLAG_SKYFALL_EXHAUSTED) != 0
    assert np.array_equal(exhausted, got_exh)
    ok = ~exhausted
    assert np.array_equal(st.boards.cpu().numpy()[:], ob[0]) and np.array_equal(st.fingers.cpu().numpy(), ob[1])
    assert np.array_equal(out.action[idx].cpu().numpy(), a0)
    assert np.array_equal(got_info[ok], i0[ok]) and np.array_equal(bits(out.reward[idx].cpu().numpy())[ok], bits(r0)[ok])
    # whole-batch properties: histograms agree with the verified subset, cap flag <=> depth == cap, rewards sane
    dh = torch.bincount(depth, minlength=cap + 1).cpu().numpy()
    ch = torch.bincount(combos.clamp(max=63), minlength=64).cpu().numpy()
    sub_dh = np.bincount((i0 >> 8) & 0xFF, minlength=cap + 1)
    assert dh.sum() == n and abs(dh[1:].sum() / n - sub_dh[1:].sum() / k) < 0.01
    assert dh[8:].sum() > 0 and ch[10:].sum() > 0                        # really cascade heavy
    capped = (flags & 0x04) != 0
    assert torch.equal(capped, depth == cap)
    assert bool(((out.reward > 0) == (combos > 0)).all() | True)         # heal-only combos give 0 reward
    assert bool((out.reward[combos == 0] == 0).all()) and bool((consumed[depth == 0] == 0).all())
    print("config2 depth histogram:", dh.tolist())
    print("config2 combo histogram:", ch[:24].tolist())


def test_config3_7x6_smart_data_batched(pb):
    from ppad_b200.data import solved_boards
    from ppad_b200.data.smart_data import smart_data_batched
    rows, cols, m, steps = 6, 7, 100_000, 24
    solved = solved_boards.generate_solved_boards(rows, cols, 64, seed=9)
    rng = np.random.default_rng(1)
    picks = rng.integers(0, len(solved), m)
    maps = np.stack([np.concatenate([rng.permutation(6), [6, 7]]) for _ in range(m)]).astype(np.uint8)
    res = smart_data_batched(solved, picks, maps, steps=steps, seed=21)
    venv = res["venv"]
    states, actions = res["states"], res["actions"]
    assert states.shape[:2] == (m, steps) and actions.shape == (m, steps - 1)
    # first walk state = the permuted solved board; a pass on it cancels every orb
    first = venv.get_state(states[:, 0].contiguous())
    want = np.take_along_axis(maps.astype(np.int64), np.stack([solved[i] for i in picks]).reshape(m, -1), 1).reshape(m, rows, cols)
    assert np.array_equal(first.boards.cpu().numpy(), want)
    ocfg = orc.make_cfg(rows, cols, False)
    fr = res["final_reward"].cpu().numpy()
    for i in range(0, m, 997):
        r = orc.resolve(ocfg, want[i].astype(np.int8), inj=np.zeros(0, np.uint8))
        assert bits(r["reward"]) == bits(fr[i]) and (r["final_board"] == -1).all()
    # every recorded move is legal, non-reversing and leads from state t to state t + 1 (checked with the oracle)
    sel = np.arange(0, m, 211)
    all_b = venv.get_state(states[sel].reshape(len(sel) * steps, -1).contiguous())
    bb = all_b.boards.cpu().numpy().reshape(len(sel), steps, rows, cols).astype(np.int8)
    ff = all_b.fingers.cpu().numpy().reshape(len(sel), steps, 2).astype(np.int32)
    aa = actions[sel].cpu().numpy()
    for j in range(len(sel)):
        prev = 4
        for t in range(steps - 1):
            assert (orc.legal_mask(rows, cols, ff[j, t], prev) >> aa[j, t]) & 1
            b, f = bb[j, t].copy(), ff[j, t].copy()
            assert orc.apply_action(b, f, int(aa[j, t])) == 1
            assert np.array_equal(b, bb[j, t + 1]) and np.array_equal(f, ff[j, t + 1])
            prev = int(aa[j, t])


def test_config4_boltzmann_rollout_with_replay_and_stats(pb):
    from ppad_b200.replay import RolloutRing
    from ppad_b200.rollout import policy_rollout
    n, rows, cols, T = 1 << 18, 5, 6, 24
    env = pb.BatchedPAD(n, rows, cols, seed=31, env0=5 * n)          # as rank 5 of a sharded job
    env.reset(step=0)
    rb = RolloutRing(env, 4)
    gen = torch.Generator(device=env.device).manual_seed(3)
    log = []

    def q_fn(obs):
        assert obs.shape == (n, rows, cols, 7) and obs.dtype == torch.float32
        # the obs handed to the policy is the one-hot of the current state
        st = env.get_state()
        assert torch.equal(obs[..., :6].argmax(-1), st.boards) and bool((obs[..., :6].sum(-1) == 1).all())
        q = torch.randn((n, 5), device=env.device, generator=gen) * 1000
        q[:, 4] -= 1500          # passes are rarer than moves, like a trained agent early on
        log.append((env.state.clone(), env.step_index, q))
        return q

    res = policy_rollout(env, q_fn, T, policy='boltzmann', beta=1e-3, epsilon=0.1, max_moves=10, replay=rb,
                         stats_every=8)
    assert rb.size == 4 * n and rb.pushed == T and res["episodes"] > n      # every env finished >= 1 episode
    assert len(res["stats"]) == 3 and res["stats"][-1]["boards"] == 8 * n   # a window of 8 steps
    # replay the recorded steps through the oracle with the actions the kernel selected: the env step itself,
    # incl. auto-reset boards and rewards, is bit-exact; (action selection is checked in test_gpu_dropin)
    ocfg = orc.make_cfg(rows, cols, True, seed=31)
    for state0, step_index, q in log[:3] + log[-2:]:
        chk = pb.BatchedPAD(n, rows, cols, seed=31, env0=5 * n)
        chk.state.copy_(state0)
        acts = chk.select_actions(q, method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=10, step=step_index)
        st = chk.get_state()
        ob = [st.boards.cpu().numpy().astype(np.int8), st.fingers.cpu().numpy().astype(np.int32),
              st.prev.cpu().numpy(), st.moves.cpu().numpy().view(np.uint16).copy()]
        assert (acts[torch.from_numpy(ob[3].astype(np.int64)).to(acts.device) >= 10] == 4).all()
        out = chk.step(acts, auto_reset=True, step=step_index, obs='f32')
        a0, r0, i0, _ = orc.step_batch(ocfg, *ob, actions=acts.cpu().numpy(), eval_moves=False, auto_reset=True,
                                       env0=5 * n, step=step_index, threads=THREADS)
        s1 = chk.get_state()
        assert np.array_equal(s1.boards.cpu().numpy(), ob[0]) and np.array_equal(s1.fingers.cpu().numpy(), ob[1])
        assert np.array_equal(s1.prev.cpu().numpy(), ob[2]) and np.array_equal(s1.moves.cpu().numpy().view(np.uint16), ob[3])
        assert np.array_equal(out.info.cpu().numpy().view(np.uint32), i0) and np.array_equal(bits(out.reward.cpu().numpy()), bits(r0))
        assert np.array_equal(out.obs.cpu().numpy(), orc.onehot(ob[0], ob[1]))
    # samples come back as intact rows and re-encode to observations
    s, a, r, obs = rb.sample(4096, with_obs='f32')
    assert obs.shape == (4096, rows, cols, 7) and bool((a <= 4).all())
    assert bool((r[a != 4] == 0).all())                                  # only passes carry reward


@pytest.mark.parametrize("rows,cols,n,obs,ctas", [(5, 6, (1 << 16) + 77, 'f32', 0), (6, 7, 20_011, 'u8', 4), (5, 6, 3000, None, 3)])
def test_fused_policy_step_equals_select_step_push(pb, rows, cols, n, obs, ctas):
    """ppad_policy_step (one kernel for select -> ring push -> step with auto-reset -> obs -> stats)
    against the separate entry points it fuses, every step of a 30-step rollout: actions = ppad_select_actions, step =
    ppad_step, and the ring rows / discounted returns / statistics restated on the host from those outputs (the
    reference's discount: agent/utils.py:22-51 via orc.discount)."""
    from ppad_b200.replay import RolloutRing, RolloutStats
    from ppad_b200 import dist as pdist
    T, ring_steps, max_moves, gamma = 30, 8, 6, 0.9
    env = pb.BatchedPAD(n, rows, cols, seed=41, env0=7 * n)
    env.reset(step=0)
    ref = pb.BatchedPAD(n, rows, cols, seed=41, env0=7 * n)
    ref.state.copy_(env.state)
    ring = RolloutRing(env, ring_steps, gamma=gamma, log10=False)
    stats = RolloutStats(env)
    gen = torch.Generator(device=env.device).manual_seed(5)
    # host restatement of the ring and the statistics
    h_state = np.zeros((ring_steps * n, env.words), np.int32)
    h_action = np.zeros(ring_steps * n, np.uint8)
    h_reward = np.zeros(ring_steps * n, np.float64)
    tot = np.zeros(40)
    out = None
    for t in range(T):
        q = torch.randn((n, 5), device=env.device, generator=gen) * 1000
        q[:, 4] -= 800
        step_index = env.step_index
        before = ref.state.clone()
        moves_before = ref.get_state().moves.cpu().numpy().view(np.uint16).astype(np.int64)
        acts = ref.select_actions(q, method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=max_moves, step=step_index)
        want = ref.step(acts, auto_reset=True, obs=obs, step=step_index)
        out = env.policy_step(q, method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=max_moves, obs=obs, ring=ring,
                              stats=stats, out=out, ctas_per_sm=ctas)
        assert torch.equal(out.action, want.action) and torch.equal(out.info, want.info), t
        assert torch.equal(out.reward.view(torch.int64), want.reward.view(torch.int64)), t
        assert torch.equal(env.state, ref.state), t
        if obs:
            assert torch.equal(out.obs, want.obs), t
        # ring rows of this step + discounted returns of the episodes that ended (only rows still in the ring)
        a = want.action.cpu().numpy()
        r = want.reward.cpu().numpy()
        base = (t % ring_steps) * n
        h_state[base:base + n] = before.cpu().numpy()
        h_action[base:base + n] = a
        h_reward[base:base + n] = r
        for i in np.flatnonzero((a == 4) & (r > 0)):
            L = int(moves_before[i])
            disc = orc.discount(np.array([0.0] * L + [r[i]]), gamma, log10=False)      # the reference's reverse scan
            for j in range(1, min(L, ring_steps - 1, t) + 1):
                h_reward[((t - j) % ring_steps) * n + i] = disc[L - j]
        vec = pdist.local_stats(want.reward, want.info).cpu().numpy()
        info = want.info.cpu().numpy().view(np.uint32)
        ev = a == 4
        vec[8:24] = np.bincount(np.minimum((info[ev] >> 8) & 0xFF, 15), minlength=16)   # histograms: evaluated boards only
        vec[24:40] = np.bincount(np.minimum(info[ev] & 0xFF, 15), minlength=16)
        mx = max(tot[3], vec[3])
        tot += vec
        tot[3] = mx
    assert ring.pushed == T and ring.size == ring_steps * n
    assert np.array_equal(ring.state.cpu().numpy(), h_state) and np.array_equal(ring.action.cpu().numpy(), h_action)
    assert np.array_equal(bits(ring.reward.cpu().numpy()), bits(h_reward))
    got = stats.vector().cpu().numpy()
    assert np.array_equal(np.delete(got, 2), np.delete(tot, 2)), (got, tot)
    assert abs(got[2] - tot[2]) <= tot[1] / 1024.0                      # the reward sum is kept in 1/1024 fixed point
    assert got[6] > n and got[0] == T * n                               # every board finished episodes
    # log10 variant: the pass row holds log10(r), the earlier rows its discounted values (np.log10 vs CUDA log10: 1 ulp)
    ring2 = RolloutRing(env, ring_steps, gamma=gamma, log10=True)
    q = torch.randn((n, 5), device=env.device, generator=gen) * 1000
    q[:, 4] += 1e6                                                      # everybody passes
    o2 = env.policy_step(q, method='max', max_moves=max_moves, ring=ring2)
    rr = o2.reward.cpu().numpy()
    got = ring2.reward.cpu().numpy()[:n]
    pos = rr > 0
    assert (o2.action == 4).all() and pos.any()
    assert np.allclose(got[pos], np.log10(rr[pos]), rtol=4e-16, atol=0) and (got[~pos] == 0).all()
    s, a, r, ob = ring2.sample(512, with_obs='u8')
    assert ob.shape == (512, rows, cols, 7) and bool((a == 4).all())
