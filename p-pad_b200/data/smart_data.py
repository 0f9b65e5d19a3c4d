"""smart_data -- training trajectories walked BACKWARDS from solved boards
(ppad/data/dataset01/smart_data.py:27-141), all trajectories in parallel on the GPU.

Per trajectory the reference does: reset(board), step('pass') -> final_reward, reset(board) (new random
finger), `steps - 1` random legal non-reversing moves (or, with steps = -1, moves until the board has
no combo left), step('pass'); then it reverses observations, reverses + inverts the moves and puts
final_reward on the last step.  Here every trajectory is one board of a BatchedPAD slab: the colour
permutation, the pass evaluations and the random walk are kernels (ppad_permute, ppad_step).
"""
import numpy as np
import torch

from .. import _cabi
from ..batched import ACTION_NAMES, BatchedPAD
from . import solved_boards

_OPPOSITE = np.array([1, 0, 3, 2, 4], np.uint8)


def smart_data_batched(solved, picks, maps, steps=100, max_walk=5000, seed=0, device=None):
    """solved: list of [R, C] boards; picks int [m] index of the solved board of each trajectory;
    maps uint8 [m, 6+] colour permutation of each trajectory.
    With steps = -1 the walk of a board ends at its first combo-free position (the reference loops without
    bound, smart_data.py:75-79); max_walk bounds the launch count, boards still unfinished then are cut there.
    Returns dict(states int32 [m, T, words] walk states in WALK order, actions uint8 [m, T-1] walk moves,
    lengths int32 [m] number of observations, final_reward float64 [m], venv)."""
    rows, cols = solved[0].shape
    m = len(picks)
    venv = BatchedPAD(m, rows, cols, skyfall_damage=False, seed=seed, device=device)
    start = np.stack([solved[i] for i in picks]).astype(np.int64)
    venv.reset(boards=start)                       # Philox fingers
    venv.permute_colours(maps)
    final_reward = venv.calculate_reward(skyfall_damage=False).reward.clone()     # smart_data.py:65-67
    venv.reset(boards=venv.get_state().boards)     # second reset: fresh random fingers (smart_data.py:68)
    states = [venv.state.clone()]
    actions = []
    if steps != -1:
        lengths = torch.full((m,), steps, dtype=torch.int32, device=venv.device)
        if steps > 1:
            # the whole random walk (legal non-reversing moves, no evaluation) in ONE launch, trajectory recorded
            walk = venv.rollout(steps - 1, eval_moves=False, auto_reset=False, max_moves=0, record=True)
            states = torch.cat([states[0][None], walk.states], 0).permute(1, 0, 2).contiguous()
            actions = walk.actions.t().contiguous()
            return dict(states=states, actions=actions, lengths=lengths, final_reward=final_reward, venv=venv)
    else:
        # walk until the first time no combo is left (smart_data.py:73-79); finished boards take PASS, a no-op
        done = torch.zeros(m, dtype=torch.bool, device=venv.device)
        lengths = torch.ones(m, dtype=torch.int32, device=venv.device)
        for _ in range(max_walk):
            acts = venv.sample_actions()
            acts = torch.where(done, torch.full_like(acts, _cabi.PASS), acts)
            out = venv.step(acts, eval_moves=True, skyfall_damage=False)
            actions.append(out.action.clone())
            states.append(venv.state.clone())
            lengths += (~done).to(torch.int32)
            done |= (out.info & 0xFF) == 0
            if bool(done.all()):
                break
    states = torch.stack(states, 1)
    actions = torch.stack(actions, 1) if actions else torch.zeros((m, 0), dtype=torch.uint8, device=venv.device)
    return dict(states=states, actions=actions, lengths=lengths, final_reward=final_reward, venv=venv)


def smart_data(boards=1, permutations=1, trajectories=1, steps=100, discount=True, gamma=0.9, log10=True,
               allowed_orbs=solved_boards.allowed_orbs, solved=None, seed=0, device=None, max_walk=5000):
    """Same arguments and return format as the reference's smart_data (smart_data.py:27-91):
    (observations, actions, rewards), one entry per trajectory; observations are lists of (board, finger)."""
    solved = solved_boards.boards if solved is None else solved
    if boards < 0 or boards > len(solved):
        raise Exception('Invalid input value for board = {0}.'.format(boards))
    if trajectories < 0:
        raise Exception('Invalid input value for traj = {0}.'.format(trajectories))
    if permutations < 0:
        raise Exception('Invalid input value for shuffle = {0}.'.format(permutations))
    rng = np.random.default_rng(seed)
    picks, maps = [], []
    for index in rng.choice(len(solved), size=boards, replace=False):
        current = np.arange(8, dtype=np.uint8)
        for _ in range(permutations):
            perm = rng.permutation(np.array(allowed_orbs, np.uint8))
            step_map = np.arange(8, dtype=np.uint8)
            step_map[np.array(solved_boards.allowed_orbs)] = perm     # original_orbs[i] -> mapping[i]
            current = step_map[current]                               # permutations compound (smart_data.py:60-63)
            for _ in range(trajectories):
                picks.append(int(index))
                maps.append(current.copy())
    if not picks:
        return [], [], []
    res = smart_data_batched(solved, np.array(picks), np.stack(maps), steps=steps, max_walk=max_walk, seed=seed,
                             device=device)
    venv = res['venv']
    m, t_max = res['states'].shape[:2]
    st = venv.get_state(res['states'].reshape(m * t_max, -1))
    all_boards = st.boards.cpu().numpy().reshape(m, t_max, *st.boards.shape[1:]).astype(int)
    all_fingers = st.fingers.cpu().numpy().reshape(m, t_max, 2).astype(int)
    acts = res['actions'].cpu().numpy()
    lengths = res['lengths'].cpu().numpy()
    final = res['final_reward'].cpu().numpy()
    # rewards: zeros with final_reward on the last action (smart_data.py:137-141), discounted on the GPU
    raw = np.zeros((m, t_max), np.float64)
    raw[np.arange(m), lengths - 1] = final
    rew = venv.discount(raw, gamma, lengths=lengths, log10=log10).cpu().numpy() if discount else raw
    observations_sd, actions_sd, rewards_sd = [], [], []
    for i in range(m):
        n_obs = int(lengths[i])
        observations_sd.append([(all_boards[i, t].copy(), all_fingers[i, t].copy()) for t in reversed(range(n_obs))])
        rev = [ACTION_NAMES[_OPPOSITE[a]] for a in reversed(acts[i, :n_obs - 1])]
        actions_sd.append(rev + ['pass'])
        rewards_sd.append(rew[i, :n_obs].copy() if discount else list(raw[i, :n_obs]))
    return observations_sd, actions_sd, rewards_sd
