"""Batched 2-step look-ahead (projects/simulation.py:71-183, ``model_simulation_2sla``).

After an episode's moves, the reference compares, with cascades off (skyfall_damage=False):
  priority 1  passing right away,
  priority 2  every legal non-reversing single move followed by a pass,
  priority 3  every legal pair of moves followed by a free roam of the agent (until it passes or
              max_lookahead_len moves were made) and a pass,
and replays the best sequence (later candidates win ties, simulation.py:161-166).  Here all candidates of
all envs are evaluated as one expanded slab: each candidate is a board of a BatchedPAD of n * 21 boards
(1 + 4 + 16 slots, illegal ones masked), and every evaluation is one ppad_step launch.
The reference shuffles the candidate order (random.shuffle); here the order is fixed (up, down, left,
right), which only changes which of several equally good sequences is picked.
"""
import torch

from . import _cabi

NEG = float("-inf")


def two_step_lookahead(venv, q_fn=None, policy='max', beta=None, max_lookahead_len=10):
    """venv: BatchedPAD holding the states at the end of the episodes (not modified).
    q_fn: callable(obs float32 [m, R, C, 7]) -> q float32 [m, 5] for the roam phase, or None (pass at once).
    Returns dict(best_reward float64 [n], best_slot int64 [n], rewards float64 [n, 21], first_moves uint8 [n, 21, 2],
    roam_actions uint8 [n, 16, max_lookahead_len]) -- slot 0 = pass, 1 + d = move d, 5 + 4 * d1 + d2 = pair."""
    n, dev = venv.n, venv.device
    rewards = torch.full((n, 21), NEG, dtype=torch.float64, device=dev)
    first = torch.full((n, 21, 2), 255, dtype=torch.uint8, device=dev)
    # priority 1
    rewards[:, 0] = venv.calculate_reward(skyfall_damage=False, step=0).reward
    # priority 2: the four single moves; REJECTED (off-board) and reversing moves stay masked
    one = venv.like(n * 4)
    one.state.copy_(venv.state.repeat_interleave(4, 0))
    d1 = torch.arange(4, dtype=torch.uint8, device=dev).repeat(n)
    legal1 = ((venv.legal_mask()[:, None] >> torch.arange(4, device=dev)) & 1).bool().reshape(-1)   # [n*4]
    out = one.step(torch.where(legal1, d1, torch.full_like(d1, _cabi.PASS)), eval_moves=True, skyfall_damage=False, step=0)
    rewards[:, 1:5] = torch.where(legal1, out.reward, torch.full_like(out.reward, NEG)).reshape(n, 4)
    first[:, 1:5, 0] = d1.reshape(n, 4)
    # priority 3: pairs, then the agent roams
    two = venv.like(n * 16)
    two.state.copy_(one.state.repeat_interleave(4, 0))
    d2 = torch.arange(4, dtype=torch.uint8, device=dev).repeat(n * 4)
    legal2 = ((one.legal_mask()[:, None] >> torch.arange(4, device=dev)) & 1).bool().reshape(-1) & legal1.repeat_interleave(4)
    two.step(torch.where(legal2, d2, torch.full_like(d2, _cabi.PASS)), step=0)
    roam = torch.full((n * 16, max_lookahead_len), 255, dtype=torch.uint8, device=dev)
    active = legal2.clone()
    steps_made = 2
    t = 0
    while q_fn is not None and steps_made < max_lookahead_len and bool(active.any()):
        q = q_fn(two.encode_obs('f32'))
        acts = two.select_actions(q, method=policy, beta=beta, step=1 + t)
        acts = torch.where(active, acts, torch.full_like(acts, _cabi.PASS))
        roam[:, t] = torch.where(active, acts, roam[:, t])
        active &= acts != _cabi.PASS
        two.step(acts, step=1 + t)          # passes are no-ops here; the evaluation happens below
        steps_made += 1
        t += 1
    out = two.calculate_reward(skyfall_damage=False, step=0)
    rewards[:, 5:] = torch.where(legal2, out.reward, torch.full_like(out.reward, NEG)).reshape(n, 16)
    first[:, 5:, 0] = d1.repeat_interleave(4).reshape(n, 16)
    first[:, 5:, 1] = d2.reshape(n, 16)
    # the LAST candidate holding the maximum wins (simulation.py:161-166)
    best = rewards.max(1).values
    is_best = rewards == best[:, None]
    best_slot = 20 - torch.flip(is_best, [1]).to(torch.int64).argmax(1)
    return dict(best_reward=best, best_slot=best_slot, rewards=rewards, first_moves=first,
                roam_actions=roam.reshape(n, 16, max_lookahead_len), final_states=two.state.reshape(n, 16, -1))
