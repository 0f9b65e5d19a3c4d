"""calculate_reward(skyfall_damage=False) (one cancel per board: the look-ahead's evaluation) and moves evaluated without
skyfall: python profiles/nosky_bench.py LIB..."""
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
for t in range(20): env.step(None, step=t)
def run(label, fn):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(100): r = fn(t)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    print("  %%-52s %%.4f ms   checksum %%.1f" %% (label, best, float(r.reward.sum())))
passes = torch.full((n,), 4, dtype=torch.uint8, device=env.device)
out = env.step(passes, skyfall_damage=False)
run("calculate_reward(skyfall_damage=False), all boards", lambda t: env.step(passes, skyfall_damage=False, out=out, step=50))
run("random move + evaluation without skyfall", lambda t: env.step(None, eval_moves=True, skyfall_damage=False, out=out, step=100 + t))
run("calculate_reward(skyfall_damage=True), all boards", lambda t: env.step(passes, skyfall_damage=True, out=out, step=50))
'''
for lib in sys.argv[1:]:
    print(os.path.basename(lib), flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": os.path.abspath(lib)}], check=False)
