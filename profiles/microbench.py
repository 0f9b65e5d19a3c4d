"""Back-to-back timing of the step kernel variants (no L2 flush, 2^20 boards): python profiles/microbench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200

n = 1 << 20
ROWS, COLS = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5, 6)
env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g))[0] for _ in range(8)]
out = env.step(None, eval_moves=True, injected=(slabs[0], 32), obs='f32', step=1)
obs = out.obs
for t in range(2, 40):
    env.step(None, eval_moves=True, injected=(slabs[t % 8], 32), obs=obs, step=t, out=out)


def run(label, reps, **kw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(reps):
        env.step(None, injected=(slabs[t % 8], 32), step=100 + t, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %.4f ms/step" % (label, e0.elapsed_time(e1) / reps))


for _ in range(2):
    run("step+eval+obs f32", 200, eval_moves=True, obs=obs)
    run("step+eval (state only)", 200, eval_moves=True)
    run("step, no eval", 200, eval_moves=False)
    run("step+eval+obs u8", 200, eval_moves=True, obs='u8')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(200):
        env.encode_obs('f32', out=obs)
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %.4f ms/step" % ("encode_obs f32", e0.elapsed_time(e1) / 200))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(20):
        env.reset(step=1000 + t)
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %.4f ms/step" % ("reset (Philox)", e0.elapsed_time(e1) / 20))
    env.reset(step=0)
