"""Batched 2-step look-ahead (projects/simulation.py:71-183, ``model_simulation_2sla``), all on the device.

The reference driver collects its training data like this, one env at a time: play an episode with the agent until it
passes (the pass is withheld) or ``max_episode_len`` moves were made; then compare, with cascades off
(skyfall_damage=False):
  priority 1  passing right away,
  priority 2  every legal non-reversing single move followed by a pass,
  priority 3  every legal pair of moves followed by a free roam of the agent (until it passes or the look-ahead made
              ``max_lookahead_len`` moves after the first one) and a pass,
and replay the best sequence (later candidates win ties, simulation.py:161-166); the (observation, action, discounted
reward) tuples of the whole episode are the data.

Here every board of a BatchedPAD slab runs that procedure in lock step, WITHOUT a host synchronisation: the episode
phase is ``max_episode_len`` rounds of (q_fn, ppad_select_actions, ppad_step) in which boards that already passed sit
out (PPAD_NOOP); the candidates of all boards are evaluated as expanded slabs (n * 4 and n * 16 boards, illegal ones
masked); the roam is a fixed ``max_lookahead_len - 1`` rounds.  The reference shuffles the candidate order
(random.shuffle), which only decides ties; here the order is fixed: up, down, left, right.
tests/golden/sim_2sla.npz holds episodes of the UNMODIFIED reference function (with that fixed order and a deterministic
linear q-function) and tests/test_gpu_dropin.py replays them.
"""
import torch

from . import _cabi

NEG = float("-inf")


def _noop_unless(mask, actions):
    return torch.where(mask, actions, torch.full_like(actions, _cabi.NOOP))


def two_step_lookahead(venv, q_fn=None, policy='max', beta=None, max_lookahead_len=10, step0=1000):
    """venv: BatchedPAD holding the states at the end of the episodes (not modified).
    q_fn: callable(obs float32 [m, R, C, 7]) -> q float32 [m, 5] for the roam phase, or None (pass at once).
    Returns dict(best_reward float64 [n], best_slot int64 [n], rewards float64 [n, 21], first_moves uint8 [n, 21, 2],
    roam_actions uint8 [n, 16, max_lookahead_len - 1] (moves, then PASS where the agent passed, NOOP after),
    one_states int32 [n, 4, words], two_states int32 [n, 16, words] (after the pair, before the roam),
    roam_states int32 [max_lookahead_len - 1, n, 16, words] (after every roam round), final_states int32 [n, 16, words])
    -- slot 0 = pass, 1 + d = move d, 5 + 4 * d1 + d2 = pair."""
    n, dev = venv.n, venv.device
    rewards = torch.full((n, 21), NEG, dtype=torch.float64, device=dev)
    first = torch.full((n, 21, 2), 255, dtype=torch.uint8, device=dev)
    # priority 1
    rewards[:, 0] = venv.calculate_reward(skyfall_damage=False, step=0).reward
    # priority 2: the four single moves; off-board and reversing moves stay masked (simulation.py:112-114)
    one = venv.like(n * 4)
    one.state.copy_(venv.state.repeat_interleave(4, 0))
    d1 = torch.arange(4, dtype=torch.uint8, device=dev).repeat(n)
    legal1 = ((venv.legal_mask()[:, None] >> torch.arange(4, device=dev)) & 1).bool().reshape(-1)   # [n*4]
    out = one.step(_noop_unless(legal1, d1), eval_moves=True, skyfall_damage=False, step=0)
    rewards[:, 1:5] = torch.where(legal1, out.reward, torch.full_like(out.reward, NEG)).reshape(n, 4)
    first[:, 1:5, 0] = d1.reshape(n, 4)
    # priority 3: pairs (simulation.py:127-129), then the agent roams (simulation.py:137-148)
    two = venv.like(n * 16)
    two.state.copy_(one.state.repeat_interleave(4, 0))
    d2 = torch.arange(4, dtype=torch.uint8, device=dev).repeat(n * 4)
    legal2 = ((one.legal_mask()[:, None] >> torch.arange(4, device=dev)) & 1).bool().reshape(-1) & legal1.repeat_interleave(4)
    two.step(_noop_unless(legal2, d2), step=0)
    two_states = two.state.clone()
    rounds = max(0, max_lookahead_len - 1)      # steps_made starts at 1 with the second move and stops at max_lookahead_len
    roam = torch.full((n * 16, rounds), _cabi.NOOP, dtype=torch.uint8, device=dev)
    roam_states = torch.empty((rounds, n * 16, venv.words), dtype=torch.int32, device=dev)
    active = legal2.clone()
    for t in range(rounds):
        if q_fn is None:
            acts = torch.full((n * 16,), _cabi.PASS, dtype=torch.uint8, device=dev)
        else:
            acts = two.select_actions(q_fn(two.encode_obs('f32')), method=policy, beta=beta, step=step0 + t)
        acts = _noop_unless(active, acts)
        roam[:, t] = acts
        active = active & (acts != _cabi.PASS)
        two.step(_noop_unless(acts != _cabi.PASS, acts), step=step0 + t)     # the pass itself is evaluated below
        roam_states[t] = two.state
    out = two.calculate_reward(skyfall_damage=False, step=0)
    rewards[:, 5:] = torch.where(legal2, out.reward, torch.full_like(out.reward, NEG)).reshape(n, 16)
    first[:, 5:, 0] = d1.repeat_interleave(4).reshape(n, 16)
    first[:, 5:, 1] = d2.reshape(n, 16)
    # the LAST candidate holding the maximum wins (simulation.py:161-166)
    best = rewards.max(1).values
    is_best = rewards == best[:, None]
    best_slot = 20 - torch.flip(is_best, [1]).to(torch.int64).argmax(1)
    return dict(best_reward=best, best_slot=best_slot, rewards=rewards, first_moves=first,
                roam_actions=roam.reshape(n, 16, rounds), one_states=one.state.reshape(n, 4, -1),
                two_states=two_states.reshape(n, 16, -1), roam_states=roam_states.reshape(rounds, n, 16, -1),
                final_states=two.state.reshape(n, 16, -1))


def model_simulation_2sla(venv, q_fn, policy='max', beta=None, max_episode_len=50, max_lookahead_len=10, gamma=0.95,
                          log10_reward=False, reset=True, step0=0):
    """One episode with the 2-step look-ahead for EVERY board of ``venv`` (simulation.py:71-183); no host sync.
    reset=True starts from fresh boards (env.reset(), simulation.py:84); otherwise from the current states.
    Returns dict(
      actions   uint8 [n, T]   env.actions of the episode: the agent's moves, then the best look-ahead sequence incl. its
                               final pass; 255 = padding; T = max_episode_len + max_lookahead_len + 2
      n_actions int64 [n]
      states    int32 [n, T, words]  the packed state every action was taken in (env.observations)
      final_reward float64 [n]       env.rewards[-1] (cascades off, as the driver's env: simulation.py:252)
      discounted float64 [n, T]      ppad.discount(env.rewards, gamma, log10) (agent/utils.py:22-51)
      best_slot int64 [n], candidate_rewards float64 [n, 21])
    and leaves venv at the final states."""
    n, dev, words = venv.n, venv.device, venv.words
    L = int(max_lookahead_len)
    T = int(max_episode_len) + L + 2
    if reset:
        venv.reset(step=step0)
    # ---- the episode: the agent plays until it passes (the pass is withheld, simulation.py:99-100) or the cap
    ep_states = torch.empty((max_episode_len, n, words), dtype=torch.int32, device=dev)
    ep_actions = torch.full((max_episode_len, n), 255, dtype=torch.uint8, device=dev)
    playing = torch.ones(n, dtype=torch.bool, device=dev)
    for t in range(max_episode_len):
        ep_states[t] = venv.state
        acts = venv.select_actions(q_fn(venv.encode_obs('f32')), method=policy, beta=beta, step=step0 + 1 + t)
        playing = playing & (acts != _cabi.PASS)
        moves = _noop_unless(playing, acts)
        ep_actions[t] = torch.where(playing, acts, ep_actions[t])
        venv.step(moves, step=step0 + 1 + t)
    ep_len = (ep_actions != 255).sum(0)                                  # moves of the episode, [n]
    # ---- the look-ahead over the end-of-episode states
    la = two_step_lookahead(venv, q_fn=q_fn, policy=policy, beta=beta, max_lookahead_len=L, step0=step0 + 100000)
    slot = la["best_slot"]
    is_pair, is_single = slot >= 5, (slot >= 1) & (slot < 5)
    pair = (slot - 5).clamp(min=0)                                       # 4 * d1 + d2
    d1 = torch.where(is_pair, pair // 4, (slot - 1).clamp(min=0))
    ar = torch.arange(n, device=dev)
    rounds = L - 1
    roam = la["roam_actions"][ar, pair]                                  # [n, rounds]
    roam_moves = (roam < 4)
    m_r = roam_moves.sum(1)
    # best sequence as (action, state it is taken in) columns 0 .. L + 1
    seq = torch.full((n, L + 2), 255, dtype=torch.uint8, device=dev)
    seq_states = torch.zeros((n, L + 2, words), dtype=torch.int32, device=dev)
    PASS = torch.full((n,), _cabi.PASS, dtype=torch.uint8, device=dev)
    seq[:, 0] = torch.where(slot == 0, PASS, d1.to(torch.uint8))
    seq_states[:, 0] = venv.state
    one_state = la["one_states"][ar, d1]
    seq[:, 1] = torch.where(is_single, PASS, torch.where(is_pair, (pair % 4).to(torch.uint8), torch.full_like(PASS, 255)))
    seq_states[:, 1] = one_state
    if rounds > 0:
        seq[:, 2:2 + rounds] = torch.where(is_pair[:, None] & roam_moves, roam, torch.full_like(roam, 255))
        after_pair = la["two_states"][ar, pair]
        roam_states = la["roam_states"][:, ar, pair].permute(1, 0, 2)    # [n, rounds, words]: state after roam round r
        seq_states[:, 2] = after_pair
        if rounds > 1:
            seq_states[:, 3:2 + rounds] = roam_states[:, :rounds - 1]
        seq_states[:, 2 + rounds] = roam_states[:, rounds - 1]
    else:
        seq_states[:, 2] = la["two_states"][ar, pair]
    pass_col = torch.where(is_pair, 2 + m_r, torch.zeros_like(m_r))
    seq.scatter_(1, pass_col[:, None], torch.where(is_pair, PASS, seq[ar, pass_col])[:, None])
    la_len = torch.where(is_pair, 3 + m_r, torch.where(is_single, 2, 1))
    # ---- env.actions / env.observations of the whole episode
    actions = torch.full((n, T), 255, dtype=torch.uint8, device=dev)
    states = torch.zeros((n, T, words), dtype=torch.int32, device=dev)
    actions[:, :max_episode_len] = ep_actions.t()
    states[:, :max_episode_len] = ep_states.permute(1, 0, 2) * (ep_actions.t() != 255)[:, :, None]
    col = ep_len[:, None] + torch.arange(L + 2, device=dev)[None, :]     # where the look-ahead sequence goes
    valid = torch.arange(L + 2, device=dev)[None, :] < la_len[:, None]
    col = torch.where(valid, col, torch.full_like(col, T - 1))           # padding is parked in the last column ...
    actions.scatter_(1, col, torch.where(valid, seq, torch.full_like(seq, 255)))
    states.scatter_(1, col[:, :, None].expand(-1, -1, words), seq_states * valid[:, :, None])
    n_actions = ep_len + la_len
    last = (n_actions - 1)
    # ... which is only a real entry for an episode of maximal length: restore it
    full = n_actions == T
    actions[:, T - 1] = torch.where(full, PASS, torch.full_like(PASS, 255))
    states[:, T - 1] = torch.where(full[:, None], seq_states[ar, (la_len - 1)], torch.zeros_like(states[:, T - 1]))
    final_reward = la["best_reward"]
    raw = torch.zeros((n, T), dtype=torch.float64, device=dev)
    raw.scatter_(1, last[:, None], final_reward[:, None])
    discounted = venv.discount(raw, gamma, lengths=n_actions.to(torch.int32), log10=log10_reward)
    # the env ends at the final state of the best sequence (a pass does not change the board)
    venv.state.copy_(seq_states[ar, la_len - 1])
    return dict(actions=actions, n_actions=n_actions, states=states, final_reward=final_reward, discounted=discounted,
                best_slot=slot, candidate_rewards=la["rewards"], episode_len=ep_len)
