"""Per-geometry timings, 2^20 boards per launch: 4x5 / 5x6 / 6x7 with 3-bit codes and 5x6 / 6x7 with 4-bit codes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppad_b200
if os.environ.get("PPAD_AB_LIB"):
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ["PPAD_AB_LIB"])
n = 1 << 20
def t(fn, reps=40):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best
print("%-14s %10s %10s %12s %12s %10s %12s" % ("geometry", "reset", "step+eval", "+obs f32", "+obs u8", "encode f32", "episode(51)"))
for rows, cols, ob in ((4, 5, 3), (5, 6, 3), (6, 7, 3), (5, 6, 4), (6, 7, 4)):
    sky = [0.1] * 10 if ob == 4 else None
    env = ppad_b200.BatchedPAD(n, rows, cols, seed=1, skyfall=sky, orb_bits=ob)
    env.reset(step=0)
    o32 = torch.empty((n, rows, cols, 7), dtype=torch.float32, device=env.device)
    o8 = torch.empty((n, rows, cols, 7), dtype=torch.uint8, device=env.device)
    out = env.step(None, eval_moves=True, obs=o8)
    cnt = [0]
    def rs():
        cnt[0] += 1; env.reset(step=100 + cnt[0])
    r = t(rs, 10)
    a = t(lambda: env.step(None, eval_moves=True, out=out))
    b = t(lambda: env.step(None, eval_moves=True, obs=o32, out=out))
    c = t(lambda: env.step(None, eval_moves=True, obs=o8, out=out))
    d = t(lambda: env.encode_obs('f32', out=o32))
    env.reset(step=0)
    e = t(lambda: env.rollout(51, max_moves=50), 4)
    print("%dx%d %d-bit     %8.4f ms %8.4f ms %9.4f ms %9.4f ms %8.4f ms %9.4f ms   obs f32: %.0f GB/s" % (
        rows, cols, ob, r, a, b, c, d, e, (rows * cols * 28 + 2 * 4 * env.words + 13) * n / b / 1e6))
    del env, o32, o8, out
    torch.cuda.empty_cache()
