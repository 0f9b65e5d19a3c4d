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
out = env.step(None, max_moves=50, auto_reset=True)
for t in range(51): env.step(None, max_moves=50, auto_reset=True, out=out)
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(102): env.step(None, max_moves=50, auto_reset=True, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 102)
print("  episode mode, one launch per step: %%.4f ms/step  (%%.2fe10 env.step calls/s)  checksum %%d" %% (best, n / best / 1e7, int(env.state.long().sum())))
best = 1e9
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(100): env.step(None, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
print("  plain random move steps:           %%.4f ms/step" %% best)
'''
for lib in sys.argv[1:]:
    print(os.path.basename(lib), flush=True)
    subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "lib": os.path.abspath(lib)}], check=False)
