"""A/B timing of kernel variants: python profiles/ab.py LIB.so [LIB2.so ...]
Each library is loaded in its own subprocess (the package binds one library per process), runs the same
2^20-board workloads back to back and prints ms per step plus a checksum of the results (variants must agree)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os
sys.path.insert(0, %(root)r)
import torch
import ppad_b200
from ppad_b200 import _cabi
_cabi.LIB_PATH = %(lib)r
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g))[0] for _ in range(8)]
out = env.step(None, eval_moves=True, injected=(slabs[0], 32), obs='f32', step=1)
obs = out.obs
for t in range(2, 30):
    env.step(None, eval_moves=True, injected=(slabs[t %% 8], 32), obs=obs, step=t, out=out)
chk = [float(out.reward.sum()), int(out.info.long().sum())]
def run(label, reps, inj=True, **kw):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(reps):
            if inj: env.step(None, injected=(slabs[t %% 8], 32), step=100 + t, out=out, **kw)
            else: env.step(None, step=100 + t, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    chk.append(float(out.reward.sum())); chk.append(int(out.info.long().sum()))
    print("  %%-30s %%.4f ms/step" %% (label, best))
run("step+eval+obs f32", 300, eval_moves=True, obs=obs)
run("step+eval (state only)", 300, eval_moves=True)
run("step+eval philox skyfall", 300, inj=False, eval_moves=True)
run("step+eval+obs u8", 300, eval_moves=True, obs='u8')
st = env.state.clone()
import time
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    env.state.copy_(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout(51, step=5000, eval_moves=False, auto_reset=True, max_moves=50)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("  %%-30s %%.4f ms (51 steps)" %% ("fused rollout episode", best))
chk.append(int(env.state.long().sum()))
print("  checksum", chk)
'''

for lib in sys.argv[1:]:
    lib = os.path.abspath(lib)
    print(os.path.basename(lib), flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": lib}], check=False)
