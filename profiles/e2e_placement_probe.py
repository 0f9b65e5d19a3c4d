"""Which PCIe stream of the zero-copy host pipe costs what?  The compact-result step of 2^20 boards with the skyfall slab
and the result buffers placed in pinned host memory or in device memory, timed back to back with CUDA events (kernel
time, no host synchronisation per step) and synchronously (the e2e call).  PPAD_PIPE_ANY_MEMORY=1 lifts the pinned check."""
import sys, os, time
os.environ["PPAD_PIPE_ANY_MEMORY"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppad_b200

n = 1 << 20
dev = torch.device("cuda:0")
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(7)


def run(fmt, orbs, slab_on_host, results_on_host, with_obs=True, reps=100, dict_entries=0):
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=20260, device=dev)
    env.reset(step=0)
    pipe = env.host_pipe(slab_orbs=orbs, compact=True, slab_format=fmt)
    if dict_entries:      # learn the dictionary with everything in pinned memory, then move the buffers
        pipe.slab.copy_(env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0].cpu())
        pipe.step(eval_moves=True, obs=obs, step=0, wait=True)
        pipe.learn_dictionary(max_entries=dict_entries)
    slab = env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0]
    b = pipe.buffers[0]
    if slab_on_host:
        b.slab.copy_(slab.cpu())
    else:
        b.slab = slab
    if not results_on_host:
        for name in ("nibbles", "tile_line", "records", "summary"):
            setattr(b, name, torch.zeros_like(getattr(b, name), device=dev))
    o = obs if with_obs else None
    torch.cuda.synchronize()
    t0, w = time.time(), 0
    while w < 20 or time.time() - t0 < 0.3:
        pipe.step(eval_moves=True, obs=o, step=1 + w, wait=True); w += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for t in range(reps):
        pipe.step(eval_moves=True, obs=o, step=1000 + t, wait=True)
    e1.record(); torch.cuda.synchronize()
    sync_ms = e0.elapsed_time(e1) / reps
    torch.cuda.synchronize(); e0.record()
    for t in range(reps):
        pipe.step(eval_moves=True, obs=o, step=2000 + t, wait=False)
    e1.record(); torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / reps
    print("%-9s %2d orbs dict=%d obs=%d slab=%-6s results=%-6s: back to back %.4f ms, synchronous %.4f ms (%.3fe9 steps/s)" %
          (fmt, orbs, dict_entries, with_obs, "host" if slab_on_host else "device", "host" if results_on_host else "device", b2b_ms,
           sync_ms, n / sync_ms / 1e6), flush=True)
    del pipe, env


mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode == "ncu":     # two launches per variant for an ncu pcie__ capture
    for fmt, orbs in (("planes3", 32), ("chunks10", 50)):
        env = ppad_b200.BatchedPAD(n, 5, 6, seed=20260, device=dev)
        env.reset(step=0)
        pipe = env.host_pipe(slab_orbs=orbs, compact=True, slab_format=fmt)
        pipe.slab.copy_(env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0].cpu())
        for t in range(3):
            pipe.step(eval_moves=True, obs=obs, step=1 + t, wait=True)
    sys.exit(0)
if mode == "dict":
    for rep in range(2):
        for slab_on_host in (True, False):
            for results_on_host in (True, False):
                run("chunks6", 54, slab_on_host, results_on_host, True, dict_entries=4096)
    sys.exit(0)
for with_obs in (True, False):
    for fmt, orbs in (("planes3", 32), ("chunks10", 50)):
        for slab_on_host in (True, False):
            for results_on_host in (True, False):
                run(fmt, orbs, slab_on_host, results_on_host, with_obs)
