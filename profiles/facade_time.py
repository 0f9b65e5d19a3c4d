import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, ppad_b200
env = ppad_b200.PAD(seed=1)
for _ in range(20): env.step(env.action_space.sample())
env.reset()
t0 = time.perf_counter(); n = 0
for ep in range(20):
    env.reset()
    for _ in range(50):
        env.step(env.action_space.sample()); n += 1
    env.step('pass'); n += 1
dt = time.perf_counter() - t0
print("facade: %.1f env.step calls/s, %.2f ms/episode (50 sampled moves + pass + reset)" % (n / dt, dt / 20 * 1e3))
moves = ['right', 'left'] * 25
t0 = time.perf_counter()
for ep in range(20):
    env.reset(finger=np.array([2, 2]))
    for m in moves: env.step(m)
    env.step('pass')
dt = time.perf_counter() - t0
print("facade, given moves: %.1f calls/s, %.2f ms/episode" % (20 * 51 / dt, dt / 20 * 1e3))
