"""One policy-rollout-like step (30 % passes, auto reset, fp32 obs) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppad_b200
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
acts = torch.where(torch.rand(n, device=env.device, generator=g) < 0.3, torch.full((n,), 4, dtype=torch.uint8, device=env.device),
                   torch.full((n,), 255, dtype=torch.uint8, device=env.device))
out = env.step(acts, auto_reset=True, obs='f32', step=1)
for t in range(2, 12):
    env.step(acts, auto_reset=True, obs=out.obs, step=t, out=out)
torch.cuda.synchronize()
print("done")
