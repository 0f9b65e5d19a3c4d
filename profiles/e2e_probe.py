"""e2e leg of bench.py alone: HostPipe.step (zero copy) per 2^20-board step, synchronous and two in flight."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
if os.environ.get('PPAD_AB_LIB'):
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ['PPAD_AB_LIB'])
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=env.device)
g = torch.Generator().manual_seed(0)
for zc, chunks in ((True, 2), ('hybrid', 2), ('hybrid', 4), ('hybrid', 8), (False, 2)):
    pipe = env.host_pipe(slab_orbs=32, chunks=chunks, zero_copy=zc)
    g.manual_seed(0)
    pipe.slab.copy_(torch.randint(0, 2**31 - 1, pipe.slab.shape, dtype=torch.int32, generator=g) & 0x55555555)
    env.reset(step=0)
    for with_obs in (True, False):
        for t in range(3):
            pipe.step(eval_moves=True, obs=obs if with_obs else None, step=t, wait=True)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for t in range(30):
                pipe.step(eval_moves=True, obs=obs if with_obs else None, step=10 + t, wait=True)
            best = min(best, (time.perf_counter() - t0) / 30)
        print("zero_copy=%-6s chunks=%d obs=%d  %.3f ms/step  %.3fe9 steps/s   reward sum %.1f" % (zc, chunks, with_obs, best * 1e3, n / best / 1e9, float(pipe.reward.sum())))
