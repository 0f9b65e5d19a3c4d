"""Device-resident rollouts (BASELINE configs[4]): policy -> env step -> replay, no host round trip per step.

The loop of projects/simulation.py:34-68 (``model_simulation``) for a whole slab at once.  Q-values come from
``q_fn`` (the agent network in a real run, synthetic in the benchmark); everything after the network is ONE kernel
per step, ``ppad_policy_step``: masking + epsilon-greedy + argmax / Boltzmann selection with the episode length cap,
the env step with auto-reset on pass, the one-hot observation of the new state, the push of (state, action, reward /
discounted return) into the time-major replay ring and the episode statistics.
"""
from . import dist as pdist
from .replay import RolloutStats


def policy_rollout(venv, q_fn, steps, policy='boltzmann', beta=1e-3, epsilon=0.0, max_moves=50, replay=None,
                   obs='f32', injected=None, stats_every=0):
    """Run `steps` env steps of every board.  q_fn(obs [n, R, C, 7]) -> q float32 [n, 5]; replay: a
    replay.RolloutRing or None.  Returns dict(obs=last observation, episodes, reward_sum, stats=list of all-gathered
    stats dicts, one per `stats_every` steps, each over its own window)."""
    out = None
    cur_obs = venv.encode_obs(obs)
    window = RolloutStats(venv)
    acc = None
    stats = []

    def close_window():
        nonlocal acc
        vec = window.vector()
        if acc is None:
            acc = vec.clone()
        else:
            mx = max(acc[3], vec[3])
            acc += vec
            acc[3] = mx
        window.clear()
        return vec

    for t in range(steps):
        q = q_fn(cur_obs)
        out = venv.policy_step(q, method=policy, beta=beta, epsilon=epsilon, max_moves=max_moves, obs=cur_obs,
                               injected=None if injected is None else injected(t), ring=replay, stats=window, out=out)
        if stats_every and (t + 1) % stats_every == 0:
            stats.append(pdist.gather_stats(close_window()))
    close_window()
    return dict(obs=cur_obs, episodes=int(acc[6].item()), reward_sum=float(acc[2].item()), stats=stats, last=out)
