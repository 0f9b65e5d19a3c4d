"""Times the UNMODIFIED Python reference (nuwapi/P-PAD) on one host core -- BASELINE.md section 3.

TEST / BENCH INFRASTRUCTURE: run by bench.py's cpu_baseline leg, one process per host core.  The reference's sources
are not part of this repository; __graft_entry__.build() stages them from /root/reference into the git-ignored
oracle/_ref/ where that directory exists, and this script exits with {"unavailable": ...} otherwise.  Two in-process
patches that do not touch game logic: a stub `tensorflow` module (ppad/__init__.py:42 imports the
TF1 agent) and an integer clock for random.seed(datetime.now()) (game.py:72,302: TypeError on Python >= 3.11).

    --mode resolved : env.step(env.action_space.sample()); env.calculate_reward(board=env.board, skyfall_damage=True)
                      (the look-ahead pattern of projects/simulation.py:117-119), reset() every 50 steps
    --mode episode  : reset(); 50 x env.step(env.action_space.sample()); env.step('pass')
                      (projects/visualize_episode.py:28-33 with L = 50, simulation.py:199)
"""
import argparse
import json
import os
import sys
import time
import types
import warnings


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="resolved", choices=["resolved", "episode"])
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--root", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref"))
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(args.root, "ppad", "pad")):
        print(json.dumps({"unavailable": "no staged reference under %s" % args.root}))
        return
    sys.path.insert(0, args.root)
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import ppad
        import ppad.pad.game as game

    class Clock:
        @staticmethod
        def now():
            return args.seed
    game.datetime = Clock
    env = ppad.PAD()
    steps = episodes = 0
    # 1 s warm-up, then the timed loop
    for phase, budget in (("warm", 1.0), ("timed", args.seconds)):
        steps = episodes = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget:
            env.reset()
            if args.mode == "resolved":
                for _ in range(50):
                    env.step(env.action_space.sample())
                    env.calculate_reward(board=env.board, skyfall_damage=True)
                    steps += 1
            else:
                for _ in range(50):
                    env.step(env.action_space.sample())
                env.step('pass')
                steps += 51
            episodes += 1
        elapsed = time.perf_counter() - t0
    print(json.dumps({"mode": args.mode, "steps": steps, "episodes": episodes, "seconds": elapsed}))


if __name__ == "__main__":
    main()
