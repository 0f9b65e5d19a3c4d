"""Multi-GPU plumbing: env sharding and the two small NCCL all-gathers the rollout needs.

The hot path has no exchange step: boards are independent, each rank owns a contiguous slab of env
ids and its Philox streams are keyed by the GLOBAL env id, so a rollout is bit-identical whether it
runs on 1, 2, 4 or 8 GPUs (tests/test_gpu_parity.py::test_full_size_one_million_boards checks the
slab-split invariance on one GPU; tests/test_dist_gloo.py runs these helpers at world_size 2).
Collectives are only used to all-gather O(100 B)/rank episode stats and sampled replay rows; both
are latency-bound and go through torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
import torch
import torch.distributed as dist

STAT_FIELDS = ["boards", "evaluated", "reward_sum", "reward_max", "combos_sum", "orbs_consumed", "episodes_done",
               "faults"]
DEPTH_BINS = 16
COMBO_BINS = 16


def shard_env0(rank, n_local):
    """global id of board 0 of `rank`'s slab"""
    return int(rank) * int(n_local)


def local_stats(reward, info):
    """reward float64 [n], info int32 [n] (ppad_step outputs) -> float64 [8 + 16 + 16] stats vector
    (STAT_FIELDS, depth histogram, combo-count histogram; last bins are '>= 15')."""
    u = info.to(torch.int64) & 0xFFFFFFFF
    combos, depth, consumed, flags = u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF, (u >> 24) & 0xFF
    head = torch.stack([
        torch.tensor(float(reward.numel()), dtype=torch.float64, device=reward.device),
        (depth > 0).sum().double(), reward.sum(), reward.max() if reward.numel() else reward.sum(),
        combos.sum().double(), consumed.sum().double(), ((flags & 0x80) != 0).sum().double(),
        ((flags & 0x7F) != 0).sum().double()])
    dh = torch.bincount(depth.clamp(max=DEPTH_BINS - 1), minlength=DEPTH_BINS).double()
    ch = torch.bincount(combos.clamp(max=COMBO_BINS - 1), minlength=COMBO_BINS).double()
    return torch.cat([head, dh, ch])


def merge_stats(per_rank):
    """[world, 40] -> dict of job-wide stats (sums, max for reward_max)"""
    s = per_rank.sum(0)
    out = {name: float(s[i]) for i, name in enumerate(STAT_FIELDS)}
    out["reward_max"] = float(per_rank[:, 3].max())
    out["depth_hist"] = [int(v) for v in s[8:8 + DEPTH_BINS]]
    out["combo_hist"] = [int(v) for v in s[8 + DEPTH_BINS:8 + DEPTH_BINS + COMBO_BINS]]
    return out


def gather_stats(local):
    """all-gather the per-rank stats vector and merge; works without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        buf = [torch.empty_like(local) for _ in range(dist.get_world_size())]
        dist.all_gather(buf, local)
        per_rank = torch.stack(buf)
    else:
        per_rank = local[None]
    return merge_stats(per_rank.cpu())


def all_gather_rows(*tensors):
    """all-gather already sampled replay rows (e.g. RolloutRing.sample): every tensor [k, ...] -> [world * k, ...]"""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return tuple(t.contiguous() for t in tensors)
    out = []
    for p in tensors:
        p = p.contiguous()
        buf = [torch.empty_like(p) for _ in range(dist.get_world_size())]
        dist.all_gather(buf, p)
        out.append(torch.cat(buf))
    return tuple(out)


def gather_rollout_window(stats_vec, state, action, reward, out=None):
    """ONE all-gather for what a rollout exchanges every K steps (BASELINE configs[4]): the float64 [40] stats vector
    of this rank's window and its k sampled replay rows (state int32 [k, words], action uint8 [k], reward float64 [k])
    travel as one packed byte row per rank.  Nothing is copied to the host: returns device tensors
    (stats [world, 40] float64, state [world * k, words], action [world * k], reward [world * k]); merge the stats with
    merge_stats(stats.cpu()) when they are wanted.  `out` reuses the receive buffer of an earlier call."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    k, words = state.shape
    parts = [stats_vec.contiguous().view(torch.uint8), reward.contiguous().view(torch.uint8),
             state.contiguous().view(torch.uint8).reshape(-1), action.contiguous().view(torch.uint8)]
    sizes = [p.numel() for p in parts]
    row = (sum(sizes) + 15) // 16 * 16
    send = torch.zeros(row, dtype=torch.uint8, device=state.device)
    torch.cat(parts, out=send[:sum(sizes)])
    if out is None or out.numel() != world * row:
        out = torch.empty(world * row, dtype=torch.uint8, device=state.device)
    if world > 1:
        dist.all_gather_into_tensor(out, send)
    else:
        out.copy_(send)
    rows = out.view(world, row)
    o0, o1, o2 = sizes[0], sizes[0] + sizes[1], sizes[0] + sizes[1] + sizes[2]
    g_stats = rows[:, :o0].contiguous().view(torch.float64).view(world, -1)
    g_reward = rows[:, o0:o1].contiguous().view(torch.float64).reshape(world * k)
    g_state = rows[:, o1:o2].contiguous().view(torch.int32).reshape(world * k, words)
    g_action = rows[:, o2:o2 + sizes[3]].contiguous().reshape(world * k)
    return g_stats, g_state, g_action, g_reward, out


def gather_replay_sample(state, action, reward, k, generator=None):
    """Uniformly sample k rows of this rank's (state, action, reward) slab and all-gather them:
    returns (state [world*k, words] int32, action [world*k] uint8, reward [world*k] float64).
    Observations are not shipped: every rank re-encodes them from the 16-byte state (ppad_encode_obs)."""
    n = state.shape[0]
    idx = torch.randint(0, n, (k,), device=state.device, generator=generator)
    parts = (state[idx].contiguous(), action[idx].contiguous(), reward[idx].contiguous())
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return parts
    out = []
    for p in parts:
        buf = [torch.empty_like(p) for _ in range(dist.get_world_size())]
        dist.all_gather(buf, p)
        out.append(torch.cat(buf))
    return tuple(out)
