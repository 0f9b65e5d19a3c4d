"""policy_step_kernel only, for ncu:  python profiles/ncu_policy.py [n_log2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
from ppad_b200.replay import RolloutRing, RolloutStats
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 21)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev)
env.reset(step=0)
ring, stats = RolloutRing(env, 4, gamma=0.99, log10=True), RolloutStats(env)
qs = [torch.randn((n, 5), device=dev, generator=gen) * 1000 for _ in range(2)]
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)
out = None
for t in range(14):     # 12 steps to reach the steady mix of episode lengths, then 2 to profile
    out = env.policy_step(qs[t & 1], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs, ring=ring, stats=stats, out=out)
torch.cuda.synchronize()
print("done", ppad_b200._cabi.lib().ppad_launch_count())
