"""A/B timing of kernel variants: python profiles/ab.py LIB.so [LIB2.so ...]
Each library is loaded in its own subprocess (PPAD_B200_LIB; the package binds one library per process), runs the same
workloads back to back and prints ms per step plus a checksum of the results (variants must agree)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, os
sys.path.insert(0, %(root)r)
import torch
import ppad_b200
from ppad_b200.replay import RolloutRing, RolloutStats
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
slabs = [env._slab(torch.randint(0, 6, (n, 32), device=env.device, dtype=torch.uint8, generator=g), 'planes3')[0] for _ in range(8)]
out = env.step(None, eval_moves=True, injected=(slabs[0], 32), slab_format='planes3', obs='f32', step=1)
obs = out.obs
for t in range(2, 30):
    env.step(None, eval_moves=True, injected=(slabs[t %% 8], 32), slab_format='planes3', obs=obs, step=t, out=out)
chk = [float(out.reward.sum()), int(out.info.long().sum())]
def timeit(fn, reps, rounds=3):
    best = 1e9
    for rep in range(rounds):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(reps):
            fn(t)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best
def run(label, reps, inj=True, **kw):
    def fn(t):
        if inj: env.step(None, injected=(slabs[t %% 8], 32), slab_format='planes3', step=100 + t, out=out, **kw)
        else: env.step(None, step=100 + t, out=out, **kw)
    best = timeit(fn, reps)
    chk.append(float(out.reward.sum())); chk.append(int(out.info.long().sum()))
    print("  %%-34s %%.4f ms/step" %% (label, best))
run("step+eval+obs f32", 300, eval_moves=True, obs=obs)
run("step+eval (state only)", 300, eval_moves=True)
run("step+eval philox skyfall", 300, inj=False, eval_moves=True)
# the state-only kernel is shorter than one Python call: replay 16 steps from a CUDA graph for the device rate
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for t in range(3):
        env.step(None, injected=(slabs[t %% 8], 32), slab_format='planes3', step=900 + t, out=out, eval_moves=True)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for t in range(16):
        env.step(None, injected=(slabs[t %% 8], 32), slab_format='planes3', step=1000 + t, out=out, eval_moves=True)
graph.replay()
print("  %%-34s %%.4f ms/step" %% ("state only, CUDA graph of 16", timeit(lambda t: graph.replay(), 20) / 16))
if os.environ.get("PPAD_AB_ONLY") == "state":
    print("  checksum", chk)
    sys.exit(0)
st = env.state.clone()
def roll(t):
    env.state.copy_(st)
    env.rollout(51, step=5000, eval_moves=False, auto_reset=True, max_moves=50)
print("  %%-34s %%.4f ms (51 steps, incl. state copy)" %% ("fused rollout episode", timeit(roll, 1)))
chk.append(int(env.state.long().sum()))
print("  %%-34s %%.4f ms" %% ("reset 2^20", timeit(lambda t: env.reset(step=7 + t), 5)))
n2 = 1 << 22
env2 = ppad_b200.BatchedPAD(n2, 5, 6, seed=3)
env2.reset(step=0)
ring, stats = RolloutRing(env2, 4, gamma=0.99, log10=True), RolloutStats(env2)
qs = [torch.randn((n2, 5), device=env2.device, generator=g) * 1000 for _ in range(4)]
obs2 = torch.empty((n2, 5, 6, 7), dtype=torch.float32, device=env2.device)
o2 = None
for t in range(12):
    o2 = env2.policy_step(qs[t & 3], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs2, ring=ring, stats=stats, out=o2)
for ctas in (3, 4):
    ms = timeit(lambda t: env2.policy_step(qs[t & 3], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs2, ring=ring, stats=stats, out=o2, ctas_per_sm=ctas), 30)
    print("  %%-34s %%.4f ms per 2^20 boards (%%.0f GB/s)" %% ("policy step full, ctas %%d" %% ctas, ms / 4, 930 * n2 / (ms * 1e-3) / 1e9))
chk.append(int(env2.state.long().sum()))
print("  checksum", chk)
'''

for lib in sys.argv[1:]:
    lib = os.path.abspath(lib)
    print(os.path.basename(lib), flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], check=False, env=dict(os.environ, PPAD_B200_LIB=lib))
