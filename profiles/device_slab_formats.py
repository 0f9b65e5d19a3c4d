"""The device-timed bench step (2^20 boards, random move + cascade + fp32 obs) with the injected skyfall slab resident in HBM
in each layout: ms per step (CUDA events, best of 3 x 300 steps, slabs rotating over 8 buffers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
n = 1 << 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for fmt, orbs in (("planes3", 32), ("chunks6", 54), ("chunks10", 50), ("chunks6t", 54), ("nibbles", 32), ("planes3", 64), ("planes3", 32)):
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev)
    env.reset(step=0)
    slabs = [env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=g), fmt)[0] for _ in range(8)]
    out = env.step(None, eval_moves=True, injected=(slabs[0], orbs), slab_format=fmt, obs='f32', step=1)
    for t in range(2, 40):
        env.step(None, eval_moves=True, injected=(slabs[t % 8], orbs), slab_format=fmt, obs=out.obs, step=t, out=out)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(300):
            env.step(None, eval_moves=True, injected=(slabs[t % 8], orbs), slab_format=fmt, obs=out.obs, step=100 + t, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 300)
    faults = int(((out.info.cpu().numpy().view('uint32') >> 24) & 0x02 != 0).sum())
    print("%-9s %2d orbs: %.4f ms/step  %.3fe9 steps/s   boards past their slab in the last step: %d" % (fmt, orbs, best, n / best / 1e6, faults), flush=True)
    del env, slabs, out
