"""The multi-rank plumbing at world_size 2 on CPU (gloo): env sharding, stats all-gather, replay sample gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ppad_b200 import dist as pdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = 1000
    assert pdist.shard_env0(rank, n_local) == rank * n_local
    g = torch.Generator().manual_seed(rank)
    reward = torch.rand(n_local, generator=g, dtype=torch.float64) * 1000 * (rank + 1)
    combos = torch.randint(0, 20, (n_local,), generator=g)
    depth = torch.randint(0, 18, (n_local,), generator=g)
    consumed = torch.randint(0, 40, (n_local,), generator=g)
    flags = torch.where(torch.arange(n_local) % 10 == 0, 0x80, 0) | torch.where(torch.arange(n_local) % 100 == 1, 0x02, 0)
    info = (combos | (depth << 8) | (consumed << 16) | (flags << 24))
    info = torch.where(info >= 2 ** 31, info - 2 ** 32, info).to(torch.int32)
    stats = pdist.gather_stats(pdist.local_stats(reward, info))
    state = torch.arange(n_local * 4, dtype=torch.int32).reshape(n_local, 4) + rank * 100000
    action = torch.full((n_local,), rank, dtype=torch.uint8)
    s, a, r = pdist.gather_replay_sample(state, action, reward, 32, generator=g)
    # the rollout's single packed all-gather: stats vector + sampled rows of every rank in one collective
    vec = pdist.local_stats(reward, info)
    g_stats, g_state, g_action, g_reward, buf = pdist.gather_rollout_window(vec, state[:8], action[:8], reward[:8])
    assert g_stats.shape == (world, 40) and torch.equal(g_stats[rank], vec)
    for peer in range(world):
        assert torch.equal(g_state[8 * peer:8 * peer + 8], torch.arange(32, dtype=torch.int32).reshape(8, 4) + peer * 100000)
        assert bool((g_action[8 * peer:8 * peer + 8] == peer).all())
    assert torch.equal(g_reward[8 * rank:8 * rank + 8], reward[:8])
    assert pdist.merge_stats(g_stats) == stats
    again = pdist.gather_rollout_window(vec, state[:8], action[:8], reward[:8], out=buf)
    assert again[4].data_ptr() == buf.data_ptr()
    if rank == 0:
        out.put((stats, s.numpy(), a.numpy(), r.numpy(), reward.numpy(), combos.numpy(), depth.numpy()))
    else:
        out.put((reward.numpy(), combos.numpy(), depth.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_and_replay_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [g for g in got if len(g) == 7][0]
    r1 = [g for g in got if len(g) == 3][0]
    stats, s, a, r, reward0, combos0, depth0 = r0
    reward1, combos1, depth1 = r1
    assert stats["boards"] == 2000 and stats["episodes_done"] == 200 and stats["faults"] == 20
    assert np.isclose(stats["reward_sum"], reward0.sum() + reward1.sum())
    assert stats["reward_max"] == max(reward0.max(), reward1.max())
    assert stats["combos_sum"] == combos0.sum() + combos1.sum()
    want_depth = np.bincount(np.minimum(np.concatenate([depth0, depth1]), 15), minlength=16)
    assert stats["depth_hist"] == want_depth.tolist() and sum(stats["combo_hist"]) == 2000
    # 32 rows from each rank, rank-major; rows are intact (state row <-> action <-> reward)
    assert s.shape == (64, 4) and (a[:32] == 0).all() and (a[32:] == 1).all()
    assert (s[:32] < 100000).all() and (s[32:] >= 100000).all()
    idx0 = s[:32, 0] // 4
    assert np.array_equal(r[:32], reward0[idx0])
