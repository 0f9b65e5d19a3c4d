"""Per-board time of the obs step and of encode_obs against the batch size (ramp / tail share of a launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
if os.environ.get('PPAD_AB_LIB'):
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ['PPAD_AB_LIB'])
for logn in (18, 19, 20, 21, 22, 23):
    n = 1 << logn
    env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
    env.reset(step=0)
    g = torch.Generator(device=env.device).manual_seed(0)
    slabs = [env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g))[0] for _ in range(4)]
    out = env.step(None, eval_moves=True, injected=(slabs[0], 32), obs='f32', step=1)
    obs = out.obs
    for t in range(2, 12):
        env.step(None, eval_moves=True, injected=(slabs[t % 4], 32), obs=obs, step=t, out=out)
    res = []
    for what in ("obs", "state", "encode"):
        best = 1e9
        reps = max(20, (1 << 26) // n)
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(reps):
                if what == "obs":
                    env.step(None, eval_moves=True, injected=(slabs[t % 4], 32), obs=obs, step=100 + t, out=out)
                elif what == "state":
                    env.step(None, eval_moves=True, injected=(slabs[t % 4], 32), step=100 + t, out=out)
                else:
                    env.encode_obs('f32', out=obs)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        res.append(best)
    print("n=2^%d  obs step %.4f ms (%.1f ps/board, %.0f GB/s)  state-only %.4f ms (%.1f ps/board)  encode %.4f ms (%.0f GB/s)" % (
        logn, res[0], res[0] * 1e9 / n, 901 * n / res[0] / 1e6, res[1], res[1] * 1e9 / n, res[2], 856 * n / res[2] / 1e6))
    del env, slabs, out, obs
    torch.cuda.empty_cache()
