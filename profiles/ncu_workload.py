"""Short workload for ncu captures (round 2): a few launches of every kernel VERDICT r1 asks evidence for.
    python profiles/ncu_workload.py [n_log2]
Launch order (after the resets / setup kernels): step_obs_kernel<5x6> x2, step_obs_kernel<5x6, compact> (host pipe) x2,
policy_step_kernel<5x6> x2, step_obs_kernel<6x7> x2, reset_kernel<6x7> x1, step_pool_kernel<5x6> x2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppad_b200
from ppad_b200.replay import RolloutRing, RolloutStats

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev)
env.reset(step=0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=dev, dtype=torch.uint8, generator=gen), 'planes3')[0] for _ in range(4)]
out = env.step(None, eval_moves=True, injected=(slabs[0], 32), slab_format='planes3', obs='f32', step=1)
for t in range(2, 4):
    env.step(None, eval_moves=True, injected=(slabs[t & 3], 32), slab_format='planes3', obs=out.obs, step=t, out=out)
pipe = env.host_pipe(slab_orbs=32, compact=True, slab_format='planes3')
pipe.slab.copy_(slabs[0].cpu())
for t in range(4, 6):
    pipe.step(eval_moves=True, obs=out.obs, step=t, wait=True)
ring, stats = RolloutRing(env, 4, gamma=0.99, log10=True), RolloutStats(env)
qs = [torch.randn((n, 5), device=dev, generator=gen) * 1000 for _ in range(2)]
for t in range(6, 8):
    env.policy_step(qs[t & 1], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=out.obs, ring=ring, stats=stats, out=out)
env7 = ppad_b200.BatchedPAD(n, 6, 7, seed=2, device=dev)
env7.reset(step=0)
slab7 = env7._slab(torch.randint(0, 6, (n, 32), device=dev, dtype=torch.uint8, generator=gen), 'planes3')[0]
out7 = env7.step(None, eval_moves=True, injected=(slab7, 32), slab_format='planes3', obs='f32', step=1)
env7.step(None, eval_moves=True, injected=(slab7, 32), slab_format='planes3', obs=out7.obs, step=2, out=out7)
env7.reset(step=3)
for t in range(8, 10):
    env.step(None, eval_moves=True, injected=(slabs[t & 3], 32), slab_format='planes3', step=t, out=out)
torch.cuda.synchronize()
print("ncu workload done:", n, "boards; launches", ppad_b200._cabi.lib().ppad_launch_count())
