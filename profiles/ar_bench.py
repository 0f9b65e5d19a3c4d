"""Step kernels with auto_reset on a policy-rollout-like action mix (30 % passes): python profiles/ar_bench.py LIB..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %(root)r)
import torch, ppad_b200
from ppad_b200 import _cabi
_cabi.LIB_PATH = %(lib)r
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
g = torch.Generator(device=env.device).manual_seed(0)
acts = [torch.where(torch.rand(n, device=env.device, generator=g) < 0.3, torch.full((n,), 4, dtype=torch.uint8, device=env.device),
                    torch.full((n,), 255, dtype=torch.uint8, device=env.device)) for _ in range(8)]
out = env.step(acts[0], auto_reset=True, obs='f32', step=1)
obs = out.obs
def run(label, **kw):
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(100):
            env.step(acts[t %% 8], auto_reset=True, step=100 + t, out=out, **kw)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    print("  %%-40s %%.4f ms/step   checksum %%d" %% (label, best, int(env.state.long().sum())))
run("30%% pass + auto reset + obs f32", obs=obs)
run("30%% pass + auto reset + obs u8", obs='u8')
run("30%% pass + auto reset (state only)")
'''
for lib in sys.argv[1:]:
    print(os.path.basename(lib), flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": os.path.abspath(lib)}], check=False)
