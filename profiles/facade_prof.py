import sys, os, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, ppad_b200
env = ppad_b200.PAD(seed=1)
for _ in range(20): env.step(env.action_space.sample())
moves = ['right', 'left'] * 25
def run():
    for ep in range(40):
        env.reset(finger=np.array([2, 2]))
        for m in moves: env.step(m)
        env.step('pass')
run()
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
def run2():
    for ep in range(40):
        env.reset()
        for _ in range(50): env.step(env.action_space.sample())
        env.step('pass')
pr = cProfile.Profile(); pr.enable(); run2(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
