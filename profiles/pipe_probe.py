"""HostPipe.step time vs chunk count (2^20 boards, 32-orb slab, obs f32 on device)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
import ppad_b200
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=env.device)
for chunks in (1, 2, 3, 4, 6, 8):
    for want_state in (True, False):
        pipe = env.host_pipe(32, chunks)
        pipe.slab.random_(0, 2**31 - 1)
        for t in range(5):
            pipe.step(eval_moves=True, obs=obs, want_state=want_state)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for t in range(30):
            pipe.step(eval_moves=True, obs=obs, want_state=want_state)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 30
        print("chunks %d want_state %-5s: %.3f ms/step  (%.2fe9 steps/s)" % (chunks, want_state, dt * 1e3, n / dt / 1e9))
        del pipe
