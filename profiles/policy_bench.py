"""Component timing of the fused policy step (configs[4]) on one GPU:  python profiles/policy_bench.py [n_log2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
from ppad_b200.replay import RolloutRing, RolloutStats

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 23)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
qs = [torch.randn((n, 5), device=dev, generator=gen) * 1000 for _ in range(4)]
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)


def run(label, ctas, use_obs=True, use_ring=True, use_stats=True, gamma=0.99, log10=True, K=30):
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=3, device=dev)
    env.reset(step=0)
    ring = RolloutRing(env, 4, gamma=gamma, log10=log10) if use_ring else None
    stats = RolloutStats(env) if use_stats else None
    out = None
    for t in range(8):
        out = env.policy_step(qs[t & 3], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs if use_obs else None,
                              ring=ring, stats=stats, out=out, ctas_per_sm=ctas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        env.policy_step(qs[t & 3], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs if use_obs else None,
                        ring=ring, stats=stats, out=out, ctas_per_sm=ctas)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%-44s ctas %d: %.4f ms per step = %.4f ms per 2^20 boards, %.0f GB/s of 930 B/board" % (
        label, ctas, ms, ms * (1 << 20) / n, 930 * n / (ms * 1e-3) / 1e9), flush=True)


for ctas in (3, 4):
    run("full (obs + ring + discount + stats)", ctas)
run("no stats", 3, use_stats=False)
run("no ring", 3, use_ring=False)
run("ring without discount/log10", 3, gamma=0.0, log10=False)
run("no obs", 3, use_obs=False)
run("no obs, no ring, no stats", 3, use_obs=False, use_ring=False, use_stats=False)
run("no obs, no ring, no stats", 4, use_obs=False, use_ring=False, use_stats=False)
