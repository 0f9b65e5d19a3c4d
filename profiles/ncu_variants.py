"""Four launches for ncu: step+eval+obs f32, step+eval (state only), encode_obs, reset (2^20 boards)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
if os.environ.get('PPAD_AB_LIB'):   # profile a variant built by profiles/build_variant.sh
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ['PPAD_AB_LIB'])

n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
slab = env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g))[0]
out = env.step(None, eval_moves=True, injected=(slab, 32), obs='f32', step=1)
obs = out.obs
for t in range(2, 30):
    env.step(None, eval_moves=True, injected=(slab, 32), obs=obs, step=t, out=out)
torch.cuda.synchronize()
env.step(None, eval_moves=True, injected=(slab, 32), obs=obs, step=100, out=out)
env.step(None, eval_moves=True, injected=(slab, 32), step=101, out=out)
env.step(None, eval_moves=True, step=102, out=out)      # Philox skyfall
env.encode_obs('f32', out=obs)
torch.cuda.synchronize()
print("done")
