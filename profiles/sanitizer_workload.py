import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import ppad_b200
for rows, cols in ((5, 6), (6, 7), (4, 5)):
    for n in (1, 33, 1000, 5000):
        env = ppad_b200.BatchedPAD(n, rows, cols, seed=3)
        env.reset()
        rng = np.random.default_rng(n)
        inj = rng.integers(0, 6, size=(n, 40)).astype(np.uint8)
        for t in range(3):
            env.step(None, eval_moves=True, injected=inj, obs='f32', auto_reset=True, max_moves=2, want_final=True, log_cap=8)
            env.step(None, eval_moves=True, injected=inj, obs='u8')
            env.step(None, eval_moves=True)
            env.step(rng.integers(0, 5, n).astype(np.uint8), eval_moves=True, paths=rng.integers(0, 4, (n, 37)).astype(np.uint8))
        env.encode_obs('f32'); env.encode_obs('u8'); env.cancel(); env.legal_mask(); env.sample_actions(); env.drop(); env.fill_board()
        env.select_actions(torch.randn(n, 5, device=env.device), method='boltzmann', beta=0.01, epsilon=0.1)
        pipe = env.host_pipe(32, 3); pipe.step(eval_moves=True); del pipe
torch.cuda.synchronize(); print("sanitizer workload done")
