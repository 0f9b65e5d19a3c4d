"""Workload for the ncu captures of round 2: every kernel VERDICT r1 asks evidence for, in its STEADY state (the boards
have been stepped until the cascade statistics settle) and at the bench's sizes.  Only the section between
cudaProfilerStart / Stop is profiled:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python profiles/ncu_workload.py

Profiled launches, in order: step_obs_kernel<5x6> x2 (the headline), step_obs_kernel<5x6, compact> x2 (host pipe, zero
copy: chunk-major base-6 slab in pinned memory, 3 eager chunks, dictionary-coded result stream -- the e2e leg of
bench.py), step_wpool_kernel<5x6> x2 (state only), policy_step_kernel<5x6> x2 (configs[4], 2^21 boards), step_obs_kernel<6x7>
x2, reset_kernel<6x7> x1, reset_kernel<5x6> x1, rollout_kernel<5x6> x1 (51 fused steps), stats_reduce_kernel etc. as
they come."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
from ppad_b200.replay import RolloutRing, RolloutStats

n = 1 << 20
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev)
env.reset(step=0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=dev, dtype=torch.uint8, generator=gen), 'planes3')[0] for _ in range(4)]
# the headline step of bench.py: 50 injected orbs per board as a chunk-major slab in HBM
oslabs = [env._slab(torch.randint(0, 6, (n, 50), device=dev, dtype=torch.uint8, generator=gen), 'chunks10')[0] for _ in range(4)]
out = env.step(None, eval_moves=True, injected=(oslabs[0], 50), slab_format='chunks10', obs='f32', step=1)
for t in range(2, 40):
    env.step(None, eval_moves=True, injected=(oslabs[t & 3], 50), slab_format='chunks10', obs=out.obs, step=t, out=out)
pipe = env.host_pipe(slab_orbs=54, compact=True, slab_format='chunks6')
pipe.slab.copy_(env._slab(torch.randint(0, 6, (n, 54), device=dev, dtype=torch.uint8, generator=gen), 'chunks6')[0].cpu())
pipe.step(eval_moves=True, obs=out.obs, step=40, wait=True)
pipe.learn_dictionary(max_entries=4096)
for t in range(3):
    pipe.step(eval_moves=True, obs=out.obs, step=37 + t, wait=True)
n2 = 1 << 21
penv = ppad_b200.BatchedPAD(n2, 5, 6, seed=3, device=dev)
penv.reset(step=0)
ring, stats = RolloutRing(penv, 4, gamma=0.99, log10=True), RolloutStats(penv)
qs = [torch.randn((n2, 5), device=dev, generator=gen) * 1000 for _ in range(2)]
pobs = torch.empty((n2, 5, 6, 7), dtype=torch.float32, device=dev)
pout = None
for t in range(14):
    pout = penv.policy_step(qs[t & 1], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=pobs, ring=ring, stats=stats, out=pout)
env7 = ppad_b200.BatchedPAD(n, 6, 7, seed=2, device=dev)
env7.reset(step=0)
slab7 = env7._slab(torch.randint(0, 6, (n, 32), device=dev, dtype=torch.uint8, generator=gen), 'planes3')[0]
out7 = env7.step(None, eval_moves=True, injected=(slab7, 32), slab_format='planes3', obs='f32', step=1)
for t in range(2, 30):
    env7.step(None, eval_moves=True, injected=(slab7, 32), slab_format='planes3', obs=out7.obs, step=t, out=out7)
renv = ppad_b200.BatchedPAD(n, 5, 6, seed=4, device=dev)
renv.reset(step=0)
renv.rollout(51, max_moves=50)
torch.cuda.synchronize()

torch.cuda.profiler.start()
for t in range(40, 42):
    env.step(None, eval_moves=True, injected=(oslabs[t & 3], 50), slab_format='chunks10', obs=out.obs, step=t, out=out)
for t in range(42, 44):
    pipe.step(eval_moves=True, obs=out.obs, step=t, wait=True)
for t in range(44, 46):
    env.step(None, eval_moves=True, injected=(slabs[t & 3], 32), slab_format='planes3', step=t, out=out)
for t in range(14, 16):
    penv.policy_step(qs[t & 1], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=pobs, ring=ring, stats=stats, out=pout)
stats.vector()
for t in range(30, 32):
    env7.step(None, eval_moves=True, injected=(slab7, 32), slab_format='planes3', obs=out7.obs, step=t, out=out7)
env7.reset(step=99)
renv.reset(step=98)
renv.rollout(51, max_moves=50)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu workload done; launches", ppad_b200._cabi.lib().ppad_launch_count())
