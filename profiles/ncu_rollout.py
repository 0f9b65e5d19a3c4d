"""One fused 51-step episode launch (ppad_rollout) and one swap-only step for ncu (2^20 boards, lock step)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
if os.environ.get('PPAD_AB_LIB'):
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ['PPAD_AB_LIB'])
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
env.rollout(51, max_moves=50)
env.rollout(51, max_moves=50)
torch.cuda.synchronize()
env.rollout(51, max_moves=50)
out = env.step(None, max_moves=50, auto_reset=True)
torch.cuda.synchronize()
print("done")
