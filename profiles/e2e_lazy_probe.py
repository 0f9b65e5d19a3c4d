"""e2e probe: HostPipe.step (compact results, zero copy) with the skyfall slab as a 12-byte row per board (planes3, 32
orbs) against the chunk-major slab read on demand (chunks10, 40 / 50 / 70 orbs offered per board).
Prints ms per 2^20-board step, steps/s, and the host->device bytes: offered (buffer size) and fetched (chunk 0 of every
board + every later chunk a board's cascade reached, counted in 32-byte sectors of 8 boards from the consumed counts)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppad_b200

n = 1 << 20
dev = torch.device("cuda:0")
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(7)


def fetched_bytes(consumed, words, sector_boards=8):
    """chunk 0 for everybody (one line per warp) + chunk c >= 1 for every sector with a board that drew more than 10 c"""
    total = 4 * n
    pad = (-n) % sector_boards
    c = np.pad(consumed, (0, pad)).reshape(-1, sector_boards).max(axis=1)
    for w in range(1, words):
        total += int((c > 10 * w).sum()) * 4 * sector_boards
    return total


def run(fmt, orbs, with_obs=True, reps=200):
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=20260, device=dev)
    env.reset(step=0)
    pipe = env.host_pipe(slab_orbs=orbs, compact=True, slab_format=fmt)
    slab = env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0]
    pipe.slab.copy_(slab.cpu())
    torch.cuda.synchronize()
    o = obs if with_obs else None
    t0, w = time.time(), 0
    while w < 20 or time.time() - t0 < 0.3:
        pipe.step(eval_moves=True, obs=o, step=1 + w, wait=True); w += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for t in range(reps):
        pipe.step(eval_moves=True, obs=o, step=1000 + t, wait=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    got = pipe.results()
    consumed = ((got["info"] >> 16) & 0xFF).astype(np.int64)
    exhausted = int((((got["info"] >> 24) & 0xFF) & 8 != 0).sum())   # counted loosely: any flag bit 3
    d2h = pipe.d2h_bytes()
    h2d_offered = pipe.h2d_bytes()
    h2d = fetched_bytes(consumed, pipe.slab_words) if fmt == "chunks10" else h2d_offered
    h2d_128 = fetched_bytes(consumed, pipe.slab_words, 32) if fmt == "chunks10" else h2d_offered
    print("%-9s %3d orbs obs=%d: %.4f ms/step %.3fe9 steps/s | h2d offered %.2f B/board, fetched %.2f (32 B sectors) %.2f (128 B lines); "
          "d2h %.2f B/board; mean consumed %.2f, flags-or 0x%x" %
          (fmt, orbs, with_obs, ms, n / ms / 1e6, h2d_offered / n, h2d / n, h2d_128 / n, d2h / n, consumed.mean(),
           int(np.bitwise_or.reduce(got["info"] >> 24))), flush=True)
    del pipe, env


for with_obs in (True, False):
    run("planes3", 32, with_obs)
    run("chunks10", 40, with_obs)
    run("chunks10", 50, with_obs)
    run("chunks10", 70, with_obs)
    run("nibbles", 32, with_obs)
