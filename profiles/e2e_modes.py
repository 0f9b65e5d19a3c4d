"""Synchronous HostPipe.step on this box: full zero copy against hybrid (slab by DMA in chunks), bench-like slab."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppad_b200
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=env.device)
g = torch.Generator(device=env.device).manual_seed(0)
slab = env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g))[0].cpu()
os.system("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader")
for rnd in range(2):
    for zc, chunks in ((True, 2), ('hybrid', 4), ('hybrid', 8)):
        pipe = env.host_pipe(slab_orbs=32, chunks=chunks, zero_copy=zc)
        pipe.slab.copy_(slab)
        for t in range(20):
            pipe.step(eval_moves=True, obs=obs, step=t, wait=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(50):
            pipe.step(eval_moves=True, obs=obs, step=100 + t, wait=True)
        dt = (time.perf_counter() - t0) / 50
        print("  %-9s chunks=%d  %.3f ms/step  %.3fe9 steps/s" % (zc, chunks, dt * 1e3, n / dt / 1e9), flush=True)
        del pipe
