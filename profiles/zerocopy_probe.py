"""Zero-copy experiment: the step kernel reads the skyfall slab from and writes reward/info/action/(a state copy)
to PINNED HOST memory directly (UVA), instead of staging through device buffers + cudaMemcpyAsync."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
from ppad_b200 import _cabi
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=env.device)
h_slab = torch.randint(0, 2**31 - 1, (n, 4), dtype=torch.int32).pin_memory()
d_slab = h_slab.to(env.device)
h_reward = torch.zeros(n, dtype=torch.float64).pin_memory()
h_info = torch.zeros(n, dtype=torch.int32).pin_memory()
h_action = torch.zeros(n, dtype=torch.uint8).pin_memory()
h_state = torch.zeros((n, 4), dtype=torch.int32).pin_memory()
d_reward = torch.zeros(n, dtype=torch.float64, device=env.device)
d_info = torch.zeros(n, dtype=torch.int32, device=env.device)
d_action = torch.zeros(n, dtype=torch.uint8, device=env.device)
d_final = torch.zeros((n, 4), dtype=torch.int32, device=env.device)

def step(t, slab, reward, info, action, final, with_obs=True):
    a = _cabi.StepArgs()
    a.n, a.env0, a.step, a.eval_moves = n, 0, t, 1
    a.state = env.state.data_ptr(); a.slab = slab.data_ptr(); a.slab_words = 4; a.n_inj = 32
    a.reward, a.info, a.action_out = reward.data_ptr(), info.data_ptr(), action.data_ptr()
    a.final_state = final.data_ptr() if final is not None else None
    a.obs = obs.data_ptr() if with_obs else None
    a.skyfall_damage = -1
    _cabi.check(env.lib.ppad_step(env._ctx, C.byref(a), env._stream()))

def timeit(label, fn, reps=30):
    for t in range(3): fn(t)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(reps):
        fn(100 + t); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-62s %.3f ms/step (%.2fe9 steps/s)" % (label, dt * 1e3, n / dt / 1e9))

timeit("all device buffers (no host traffic)", lambda t: step(t, d_slab, d_reward, d_info, d_action, d_final))
timeit("results -> pinned host (reward, info, action), slab device", lambda t: step(t, d_slab, h_reward, h_info, h_action, None))
timeit("results + state copy -> pinned host, slab device", lambda t: step(t, d_slab, h_reward, h_info, h_action, h_state))
timeit("results + state copy -> pinned host, slab <- pinned host", lambda t: step(t, h_slab, h_reward, h_info, h_action, h_state))
timeit("same, no obs", lambda t: step(t, h_slab, h_reward, h_info, h_action, h_state, with_obs=False))
