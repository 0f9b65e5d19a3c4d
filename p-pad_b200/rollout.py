"""Device-resident rollouts (BASELINE configs[4]): policy -> env step -> replay, no host round trip per step.

The loop of projects/simulation.py:34-68 (``model_simulation``) for a whole slab at once: Q-values come
from ``q_fn`` (the agent network in a real run, synthetic in the benchmark), actions from
``ppad_select_actions`` (masking + epsilon-greedy + argmax/Boltzmann), the env step with auto-reset on
pass from ``ppad_step`` (one-hot observation of the new state emitted by the same kernel), and the
(state, action, reward) rows go into the device replay ring.
"""
import torch

from . import dist as pdist


def policy_rollout(venv, q_fn, steps, policy='boltzmann', beta=1e-3, epsilon=0.0, max_moves=50, replay=None,
                   obs='f32', injected=None, stats_every=0):
    """Run `steps` env steps of every board.  q_fn(obs [n, R, C, 7]) -> q float32 [n, 5].
    Returns dict(obs=last observation, episodes, reward_sum, stats=list of all-gathered stats dicts)."""
    out = None
    cur_obs = venv.encode_obs(obs)
    episodes = torch.zeros((), dtype=torch.int64, device=venv.device)
    reward_sum = torch.zeros((), dtype=torch.float64, device=venv.device)
    stats = []
    for t in range(steps):
        q = q_fn(cur_obs)
        actions = venv.select_actions(q, method=policy, beta=beta, epsilon=epsilon, max_moves=max_moves)
        before = venv.state.clone() if replay is not None else None
        out = venv.step(actions, auto_reset=True, injected=None if injected is None else injected(t), obs=cur_obs, out=out)
        done = (out.info.to(torch.int64) >> 24) & 0x80
        episodes += (done != 0).sum()
        reward_sum += out.reward.sum()
        if replay is not None:
            replay.push(before, out.action, out.reward)
        if stats_every and (t + 1) % stats_every == 0:
            stats.append(pdist.gather_stats(pdist.local_stats(out.reward, out.info)))
    return dict(obs=cur_obs, episodes=int(episodes.item()), reward_sum=float(reward_sum.item()), stats=stats, last=out)
