import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ppad_b200
n = 1 << 20
def t(label, fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.4f ms" % (label, e0.elapsed_time(e1) / reps))
for rows, cols in ((4, 5), (5, 6), (6, 7)):
    env = ppad_b200.BatchedPAD(n, rows, cols, seed=1)
    env.reset(step=0)
    o8 = torch.empty((n, rows, cols, 7), dtype=torch.uint8, device=env.device)
    o32 = torch.empty((n, rows, cols, 7), dtype=torch.float32, device=env.device)
    cnt = [0]
    def rs():
        cnt[0] += 1; env.reset(step=100 + cnt[0])
    t("%dx%d reset" % (rows, cols), rs)
    print("   mean attempts:", float(env.last_reset_consumed.float().mean()) / (rows * cols))
    t("%dx%d encode u8" % (rows, cols), lambda: env.encode_obs('u8', out=o8))
    t("%dx%d encode f32" % (rows, cols), lambda: env.encode_obs('f32', out=o32))
    out = env.step(None, eval_moves=True, obs=o8)
    t("%dx%d step+eval+obs u8 (philox)" % (rows, cols), lambda: env.step(None, eval_moves=True, obs=o8, out=out))
    t("%dx%d step+eval+obs f32 (philox)" % (rows, cols), lambda: env.step(None, eval_moves=True, obs=o32, out=out))
    t("%dx%d step+eval (philox)" % (rows, cols), lambda: env.step(None, eval_moves=True, out=out))
