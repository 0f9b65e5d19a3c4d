"""state-only step kernel (warp pool) in its steady state, for ncu:  python profiles/ncu_pool.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
n = 1 << 20
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev)
env.reset(step=0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=dev, dtype=torch.uint8, generator=gen), 'planes3')[0] for _ in range(4)]
out = env.step(None, eval_moves=True, injected=(slabs[0], 32), slab_format='planes3', step=1)
for t in range(2, 40):
    env.step(None, eval_moves=True, injected=(slabs[t & 3], 32), slab_format='planes3', step=t, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in range(40, 42):
    env.step(None, eval_moves=True, injected=(slabs[t & 3], 32), slab_format='planes3', step=t, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
