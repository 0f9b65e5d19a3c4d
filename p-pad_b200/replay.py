"""Device-side experience replay ring (projects/simulation.py:294-306) of (state, action, reward) rows.

The reference keeps a Python list of (observation, action, discounted reward) tuples, shuffles and
truncates it to MAX_DATA_SIZE and samples minibatches from it.  Here the rows live in HBM as the
16/32-byte packed state + 1-byte action + float64 reward; observations are re-encoded from the state
on demand (ppad_encode_obs) instead of being stored (840 B vs 16 B per row).
"""
import torch

from . import _cabi
from .batched import _ptr


class ReplayBuffer:
    def __init__(self, venv, capacity):
        self.venv = venv
        self.capacity = int(capacity)
        dev = venv.device
        self.state = torch.zeros((self.capacity, venv.words), dtype=torch.int32, device=dev)
        self.action = torch.zeros(self.capacity, dtype=torch.uint8, device=dev)
        self.reward = torch.zeros(self.capacity, dtype=torch.float64, device=dev)
        self.cursor = 0      # next ring slot
        self.size = 0        # valid rows
        self.samples = 0

    def push(self, state, action, reward, keep=None):
        """Append rows (only those with keep[i] != 0 when given; kept rows stay in order and contiguous).  A push that
        keeps more rows than the ring holds stores its last `capacity` rows (simulation.py:296-299 truncates, too)."""
        v = self.venv
        n = state.shape[0]
        state = state.contiguous()
        action = v._dev(action, torch.uint8)
        reward = v._dev(reward, torch.float64)
        slot = None
        count = n
        if keep is not None:
            keep = v._dev(keep, torch.uint8)
            k64 = (keep != 0).to(torch.int64)
            slot = (torch.cumsum(k64, 0) - k64).contiguous()    # exclusive prefix sum: ring offset of each kept row
            count = int(k64.sum().item())
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_replay_push(v._ctx, n, self.capacity, self.cursor, count, _ptr(state), _ptr(action),
                                               _ptr(reward), _ptr(keep), _ptr(slot), _ptr(self.state),
                                               _ptr(self.action), _ptr(self.reward), v._stream()))
        self.cursor = (self.cursor + count) % self.capacity
        self.size = min(self.capacity, self.size + count)
        return count

    def sample(self, k, with_obs=None):
        """k uniform rows -> (state int32 [k, words], action uint8 [k], reward float64 [k], obs or None)"""
        if self.size == 0:
            raise ValueError("replay buffer is empty")
        v = self.venv
        state = torch.empty((k, v.words), dtype=torch.int32, device=v.device)
        action = torch.empty(k, dtype=torch.uint8, device=v.device)
        reward = torch.empty(k, dtype=torch.float64, device=v.device)
        self.samples += 1
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_replay_sample(v._ctx, k, self.size, self.samples & 0xFFFFFFFF, _ptr(self.state),
                                                 _ptr(self.action), _ptr(self.reward), _ptr(state), _ptr(action),
                                                 _ptr(reward), None, v._stream()))
        obs = v.encode_obs(with_obs, state=state) if with_obs else None
        return state, action, reward, obs


class RolloutRing:
    """Time-major replay ring of a fused policy rollout (BatchedPAD.policy_step): `steps` slabs of n rows, the row of
    (step t, board i) is (t % steps) * n + i.  Rows hold the packed state the action was chosen in, the action and the
    reward -- with gamma > 0 the discounted return of ppad.discount (agent/utils.py:22-51), written back by the kernel
    when the episode ends, i.e. the (observation, action, discounted reward) tuples of simulation.py:59-61."""

    def __init__(self, venv, steps, gamma=0.0, log10=False):
        self.venv = venv
        self.steps = int(steps)
        self.n = venv.n
        self.gamma, self.log10 = float(gamma), bool(log10)
        dev = venv.device
        self.state = torch.zeros((self.steps * self.n, venv.words), dtype=torch.int32, device=dev)
        self.action = torch.zeros(self.steps * self.n, dtype=torch.uint8, device=dev)
        self.reward = torch.zeros(self.steps * self.n, dtype=torch.float64, device=dev)
        self.pushed = 0        # slabs pushed so far (the kernel's ring_step)
        self.samples = 0

    @property
    def size(self):
        return min(self.pushed, self.steps) * self.n

    def sample(self, k, with_obs=None):
        """k uniform rows of the filled part -> (state, action, reward, obs or None), as ReplayBuffer.sample"""
        if self.size == 0:
            raise ValueError("replay ring is empty")
        v = self.venv
        state = torch.empty((k, v.words), dtype=torch.int32, device=v.device)
        action = torch.empty(k, dtype=torch.uint8, device=v.device)
        reward = torch.empty(k, dtype=torch.float64, device=v.device)
        self.samples += 1
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_replay_sample(v._ctx, k, self.size, self.samples & 0xFFFFFFFF, _ptr(self.state),
                                                 _ptr(self.action), _ptr(self.reward), _ptr(state), _ptr(action),
                                                 _ptr(reward), None, v._stream()))
        obs = v.encode_obs(with_obs, state=state) if with_obs else None
        return state, action, reward, obs


class RolloutStats:
    """Per-CTA episode statistics accumulated by BatchedPAD.policy_step (the report block of simulation.py:336-349):
    ``vector()`` sums them into the float64 [40] layout of dist.STAT_FIELDS + depth / combo histograms, ready for
    dist.gather_stats; ``clear()`` starts a new window."""

    REWARD_SCALE = 1024.0

    def __init__(self, venv):
        self.venv = venv
        rows = int(venv.lib.ppad_policy_stats_rows(venv.n))
        self.rows = torch.zeros((rows, 40), dtype=torch.int64, device=venv.device)
        self.out = torch.zeros(40, dtype=torch.int64, device=venv.device)

    def clear(self):
        self.rows.zero_()

    def vector(self):
        v = self.venv
        with torch.cuda.device(v.device):
            _cabi.check(v.lib.ppad_policy_stats(v._ctx, v.n, _ptr(self.rows), _ptr(self.out), v._stream()))
        vec = self.out.to(torch.float64)
        vec[2] = vec[2] / self.REWARD_SCALE
        vec[3] = self.out[3:4].view(torch.float64)[0]
        return vec
