"""Per-step device time of the zero-copy host pipe step (compact results), measured with CUDA events around single steps:
isolated (2 ms of idle between steps), synchronous loop, back to back -- for the row slab (planes3), the on-demand slab
(chunks10, 50 orbs) and the on-demand slab cut to its first chunk (10 orbs: no lazy host reads at all)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppad_b200

n = 1 << 20
dev = torch.device("cuda:0")
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(7)


def run(fmt, orbs, with_obs=True, reps=60, eager=0, dict_entries=0):
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=20260, device=dev)
    env.reset(step=0)
    pipe = env.host_pipe(slab_orbs=orbs, compact=True, slab_format=fmt, slab_eager=eager)
    pipe.slab.copy_(env._slab(torch.randint(0, 6, (n, orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0].cpu())
    o = obs if with_obs else None
    torch.cuda.synchronize()
    t0, w = time.time(), 0
    while w < 20 or time.time() - t0 < 0.3:
        pipe.step(eval_moves=True, obs=o, step=1 + w, wait=True); w += 1

    cover = 0.0
    if dict_entries:
        entries, cover = pipe.learn_dictionary(max_entries=dict_entries)
        for w in range(10):
            pipe.step(eval_moves=True, obs=o, step=500 + w, wait=True)

    def one(t, idle):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.step(eval_moves=True, obs=o, step=t, wait=False)      # join=True: the current stream waits for the step
        e1.record()
        if idle is not None:
            pipe.wait()
            torch.cuda.synchronize()
            if idle:
                time.sleep(idle)
        return e0, e1

    out = {}
    for label, idle in (("isolated", 0.002), ("synchronous", 0.0), ("back_to_back", None)):
        torch.cuda.synchronize()
        evs = [one(5000 + t, idle) for t in range(reps)]
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        out[label] = (np.median(ms), ms.min(), ms.max())
    d2h = pipe.d2h_bytes()
    got = pipe.results()
    print("%-9s %2d orbs eager=0x%x dict=%d (cover %.3f, escapes %.3f) obs=%d h2d %.2f + %.2f B/board: " % (fmt, orbs, eager, dict_entries, cover, got["escapes"] / max(1, int((got["info"] != 0).sum())), with_obs, pipe.h2d_bytes() / n, pipe.h2d_bytes_on_demand(got["info"]) / n) +
          "; ".join("%s median %.4f ms (min %.4f max %.4f)" % ((k,) + v) for k, v in out.items()) +
          "; d2h %.2f B/board" % (d2h / n), flush=True)
    del pipe, env


for rep in range(3):
    run("chunks6", 54, True, reps=100, eager=3, dict_entries=4096)
    run("chunks6t", 54, True, reps=100, eager=3, dict_entries=4096)
    run("chunks6t", 54, True, reps=100, eager=4, dict_entries=4096)
    run("chunks6t", 54, True, reps=100, eager=2, dict_entries=4096)
