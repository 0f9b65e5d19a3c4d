#!/usr/bin/env python
"""bench.py -- env steps/sec including cascades (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU port of the reference's path

Workload (BASELINE.json configs[1]): 2^20 independent 5x6 boards per GPU, one env step per board per
"step": a uniformly random legal non-reversing move (PlayerAction.sample, in-kernel Philox), the
orb swap, the full cascade resolution on a copy (cancel -> drop -> skyfall until stable) with the
skyfall taken from an injected 50-orb slab per board per step (chunk-major, ten 3-bit codes per uint32: the kernel
reads 8 bytes per board up front and a later chunk only where a cascade draws more than 20 orbs), the float64
reward, the packed state written back and the fp32 one-hot observation [n,5,6,7] of the new state --
ONE fused kernel launch.  N > 1: every rank owns its own 2^20-board slab (env ids rank * 2^20 + i); no
data-path collective (weak scaling).

Timing: W >= 3 warm-up steps, then EXACTLY K steps between barrier + synchronize, CUDA events on the launch
stream, max over ranks -> `value`.  Because K steps of 0.16 ms are a short region, a steady-state leg of
>= 0.6 s follows with one CUDA event per step (`steady`: median / p10 / p90 per step) while nvidia-smi is
sampled every 100 ms (>= 5 samples).  Every step streams 881 MB of obs (7x the 126 MB L2) and the skyfall
slabs rotate over 16 buffers (336 MB), so no step finds its inputs in L2.

Sub-records of the same JSON line (all measured in this run, on every rank, max over ranks):
  state_only    the same step without the observation stream (persistent pool kernel), L2 flushed
  e2e           the same step through HostPipe.step (ppad_pipe_step): pinned HOST buffers, the skyfall slab (54 orbs
                offered per board, chunk-major; three chunks = 6 B fetched per board, more on demand) crosses PCIe host ->
                device and the compact, dictionary-coded result stream device -> host inside the timed region
  rollout       BASELINE configs[4]: 2^23 boards per GPU, Boltzmann policy on synthetic Q-values, auto-reset, fp32 obs,
                replay ring with discounted returns, statistics; NCCL all-gathers of stats + replay samples INSIDE the
                timed region
  config2       BASELINE configs[2]: 2^24 boards, finger paths of 200 moves applied in one launch, then resolved
  config3       BASELINE configs[3]: 7x6 boards, obs step + smart-data style trajectory generation
  episode_mode  50 random moves + pass + reset, per-step launches and fused
  cpu_baseline  the C port of the reference's path on all host threads and, where the reference's sources were staged
                by build() (oracle/_ref), the unmodified Python reference itself on all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LOCAL = 1 << 20
ROWS, COLS = 5, 6
# the headline (device-timed) leg: 50 injected orbs per board-step resident in HBM as a chunk-major slab (PPAD_SLAB_CHUNKS10,
# ten 3-bit codes per uint32, 20 B per board); the kernel reads the first two chunks of every board (8 B) and a later one
# only where a cascade draws more than 20 orbs.  Faster than the 32-orb row slab it replaces (0.151 against 0.154 ms,
# profiles/r02av_device_slab_formats.txt), and one board-step in two million runs past its slab instead of one in 12 000
# (SURVEY.md section 8d asks for >= 42 orbs).
OBS_SLAB_ORBS = 50
OBS_SLAB_FORMAT = 'chunks10'
OBS_SLAB_BYTES = 8
# the state-only leg (the pooled kernels read row slabs) and configs[3]: 32 orbs, bit-sliced rows
SLAB_ORBS = 32
SLAB_FORMAT = 'planes3'          # PPAD_SLAB_PLANES3: 3 bits per orb, 12 bytes per 32 orbs
SLAB_BYTES = 12
# the e2e leg (host buffers): the host OFFERS 54 orbs per board-step as a chunk-major slab (PPAD_SLAB_CHUNKS6: six base-6
# orbs per uint16, 18 B per board in pinned memory); the kernel fetches the first three chunks of every board (6 B) and a
# later chunk only for the boards whose cascade draws more than 18 orbs (about 7 in 1000).  (The tile-major variant
# PPAD_SLAB_CHUNKS6T is 2 % faster on one GPU and 10 % slower on eight: profiles/r02ao_e2e_multi_8gpu.txt.)
E2E_SLAB_ORBS = 54
E2E_SLAB_FORMAT = 'chunks6'
E2E_DICT_ENTRIES = 4096
N_SLABS = 16
SEED = 20181019
# algorithmic bytes per board-step (restated in DESIGN.md):
#   state in 16 + state out 16 + action 1 + reward 8 + info 4 = 45; + skyfall read: 8 (two eager chunks of the 50-orb slab;
#   the row slab of the state-only leg: 12); + fp32 obs 840
BYTES_STATE_ONLY = 45 + SLAB_BYTES
BYTES_WITH_OBS = 45 + OBS_SLAB_BYTES + ROWS * COLS * 7 * 4
# configs[4] policy step: state 32, q 20, results 13, obs 840, replay ring row 16 + 1 + 8
BYTES_ROLLOUT = 32 + 20 + 13 + 840 + 25


def workload_config(n, world):
    return {"workload": "configs[1]: 2^20 5x6 boards/GPU, random legal move + full cascade resolve + "
                        "fp32 one-hot obs, injected 50-orb skyfall slab per board-step",
            "boards_per_gpu": n, "rows": ROWS, "cols": COLS, "skyfall_orbs_per_step": OBS_SLAB_ORBS,
            "skyfall_slab": "device-timed leg: 50 injected orbs per board-step resident in HBM (chunk-major, ten 3-bit codes "
                            "per uint32: 20 B per board, of which the kernel reads 8 B for every board and the rest where a "
                            "cascade draws it); about one board-step in two million draws more than 50 orbs and continues on the "
                            "Philox stream at the same draw index, flagged SKYFALL_EXHAUSTED, as the oracle does (stats_last_step.faults).  e2e leg: the host offers 54 orbs per board-step "
                            "in pinned memory (chunk-major, base 6), of which 18 (6 B) cross PCIe for every board and the "
                            "rest on demand.  (state_only sub-record: 32-orb row slab, 12 B; a board that draws more "
                            "continues on the Philox stream at the same draw index, flagged SKYFALL_EXHAUSTED, as the "
                            "oracle does)",
            "l2": "inputs > L2: 881 MB obs stream per step, skyfall slabs rotate over 16 buffers (336 MB)",
            "parallelism": "env-sharded x%d, no data-path collective" % world}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        sm = [float(r[0]) for r in rows]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(rows[0][1]), "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# CPU legs
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_rate(n_sample, steps, threads):
    """The oracle (C port of the reference's path) timed on `threads` host cores: same workload, bounded sample."""
    import oracle as orc
    rng = np.random.default_rng(1)
    cfg = orc.make_cfg(ROWS, COLS, True, seed=SEED)
    boards = np.zeros((n_sample, ROWS, COLS), np.int8)
    fingers = np.zeros((n_sample, 2), np.int32)
    edges = [n_sample * k // threads for k in range(threads + 1)]
    # combo-free start boards like PAD.reset; cheap enough to do through the scalar binding on a subsample
    base = min(n_sample, 4096)
    for i in range(base):
        b, f, _, _ = orc.reset(cfg, env=i, step=0)
        boards[i], fingers[i] = b, f
    reps = (n_sample + base - 1) // base
    boards[:] = np.tile(boards[:base], (reps, 1, 1))[:n_sample]
    fingers[:] = np.tile(fingers[:base], (reps, 1))[:n_sample]
    prev = np.full(n_sample, 4, np.uint8)
    moves = np.zeros(n_sample, np.uint16)
    inj = rng.integers(0, 6, size=(n_sample, OBS_SLAB_ORBS)).astype(np.uint8)
    obs = np.zeros((n_sample, ROWS, COLS, 7), np.float32)
    from concurrent.futures import ThreadPoolExecutor

    def work(k, step):
        sl = slice(edges[k], edges[k + 1])
        orc.step_batch(cfg, boards[sl], fingers[sl], prev[sl], moves[sl], actions=None, eval_moves=True, inj=inj[sl],
                       env0=edges[k], step=step)
        orc.onehot(boards[sl], fingers[sl], out=obs[sl])      # the fp32 one-hot obs, like the GPU step

    times = []
    with ThreadPoolExecutor(threads) as ex:
        for t in range(steps):
            t0 = time.perf_counter()
            list(ex.map(lambda k: work(k, t + 1), range(threads)))
            times.append(time.perf_counter() - t0)
    return n_sample / float(np.mean(times)), float(np.mean(times))


def python_reference_rate(seconds=8.0):
    """the UNMODIFIED Python reference (staged by build() into the git-ignored oracle/_ref/ from
    /root/reference; absent -> None) on P = os.cpu_count() processes, one env each: (a) the resolved-step loop of
    projects/simulation.py:117-119, (b) the episode loop of projects/visualize_episode.py:28-33 with L = 50."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "ppad", "pad")):
        return None
    procs = os.cpu_count() or 1
    script = os.path.join(ROOT, "oracle", "ref_bench.py")
    out = {}
    for mode in ("resolved", "episode"):
        ps = [subprocess.Popen([sys.executable, script, "--mode", mode, "--seconds", str(seconds), "--seed", str(100 + k),
                                "--root", ref_root], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
              for k in range(procs)]
        rates = []
        for p in ps:
            txt, _ = p.communicate(timeout=seconds * 6 + 60)
            try:
                rates.append(json.loads(txt.strip().splitlines()[-1]))
            except Exception:
                pass
        if not rates:
            return None
        out[mode] = rates
    cpu = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    res = sum(r["steps"] / r["seconds"] for r in out["resolved"])
    epi = sum(r["steps"] / r["seconds"] for r in out["episode"])
    return {"resolved_steps_per_s": res, "resolved_steps_per_s_per_core": res / len(out["resolved"]),
            "episode_env_step_calls_per_s": epi, "episode_env_step_calls_per_s_per_core": epi / len(out["episode"]),
            "episodes_per_s": sum(r["episodes"] / r["seconds"] for r in out["episode"]),
            "processes": len(out["resolved"]), "cpu_model": cpu, "seconds_per_loop": seconds,
            "what": "unmodified ppad/pad/game.py + utils.py + player.py (TF stub + int clock patches only), PAD() defaults "
                    "(5x6, skyfall on): resolved = env.step(env.action_space.sample()); env.calculate_reward(board=env.board, "
                    "skyfall_damage=True) (simulation.py:117-119); episode = reset() + 50 sampled moves + step('pass') "
                    "(visualize_episode.py:28-33)"}


def run_reference(args):
    """--impl reference: the reference is pure Python; this arm times the oracle port (oracle/ppad_oracle.c) of the same
    path on all host cores -- a far FASTER baseline than the reference itself, which is reported beside it
    (cpu_baseline.python_reference) wherever its sources were staged."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # warm-up doubles as calibration: size the per-step sample so that K timed steps take about two minutes at most
    rate0, _ = cpu_port_rate(min(N_LOCAL, (1 << 12) * threads), max(1, min(args.warmup, 2)), threads)
    budget_s = 100.0
    n_sample = int(min(N_LOCAL, max((1 << 10) * threads, rate0 * budget_s / max(1, args.steps))))
    n_sample -= n_sample % threads
    rate, sec = cpu_port_rate(n_sample, args.steps, threads)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec_incl_cascades", "value": rate, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 boards + f64 reward + f32 obs",
        "data": "synthetic",
        "config": workload_config(N_LOCAL, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port",
                         "sample": "%d boards per step x %d steps of the same workload (move + cascade + fp32 obs), C port "
                                   "of ppad/pad/game.py + utils.py + agent01.py:254-273 on %d threads" % (
                                       n_sample, args.steps, threads)},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_python_reference:
        pr = python_reference_rate(6.0)
        if pr:
            line["cpu_baseline"]["python_reference"] = pr
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, torch, dist, dev, world):
        self.torch, self.dist, self.dev, self.world = torch, dist, dev, world

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, k0, count, finish=None):
        """barrier + sync, `count` steps between CUDA events on the launch stream, sync; max over ranks (ms).
        finish(): joins side streams into the launch stream before the closing event."""
        torch = self.torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(k0, k0 + count):
            fn(t)
        if finish is not None:
            finish()
        e1.record()
        torch.cuda.synchronize()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def per_step(self, fn, k0, count):
        """one CUDA event per step: per-step durations in ms (this rank)"""
        torch = self.torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(count + 1)]
        ev[0].record()
        for j in range(count):
            fn(k0 + j)
            ev[j + 1].record()
        torch.cuda.synchronize()
        return np.array([ev[j].elapsed_time(ev[j + 1]) for j in range(count)])


def leg_rollout(ppad_b200, pdist, T, args, rank, world, dev, peak):
    """BASELINE configs[4] at --rollout-boards per GPU (default 2^23): per step the fused policy kernel
    (ppad_policy_step: Boltzmann selection on Q-values resident in HBM -- 4 rotating synthetic buffers q ~ N(0,1) * 1000,
    the stand-in for Agent01's output --, episode cap 50, env step, auto-reset, fp32 one-hot obs, time-major replay ring
    with discounted returns, statistics); every 10 steps the statistics vector and 32 replay rows per GPU are
    all-gathered over NCCL, inside the timed region."""
    import torch
    from ppad_b200.replay import RolloutRing, RolloutStats
    n = args.rollout_boards
    K, W = 50, 5
    env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED + 4, device=dev, env0=pdist.shard_env0(rank, n))
    env.reset(step=0)
    ring = RolloutRing(env, 4, gamma=0.99, log10=True)
    stats = RolloutStats(env)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    qs = [torch.randn((n, 5), device=dev, dtype=torch.float32, generator=gen) * 1000.0 for _ in range(4)]
    obs = torch.empty((n, ROWS, COLS, 7), dtype=torch.float32, device=dev)
    out = env.policy_step(qs[0], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs, ring=ring, stats=stats)
    last = {}

    def one(t):
        env.policy_step(qs[t & 3], method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50, obs=obs, ring=ring,
                        stats=stats, out=out)
        if (t + 1) % 10 == 0:
            # one NCCL all-gather of (stats vector, 32 sampled replay rows) per rank; results stay on the device
            vec = stats.vector()
            stats.clear()
            s, a, r, _ = ring.sample(32)
            last["gathered"] = pdist.gather_rollout_window(vec, s, a, r, out=last.get("buf"))
            last["buf"] = last["gathered"][4]

    for t in range(W):
        one(t)
    ms = T.timed(one, W, K)
    per = ms / K
    achieved = BYTES_ROLLOUT * n / (per * 1e-3) / 1e9
    g_stats, g_state = last["gathered"][0], last["gathered"][1]
    last["stats"] = pdist.merge_stats(g_stats.cpu())
    last["replay_rows"] = int(g_state.shape[0])
    del env, ring, stats, qs, obs
    torch.cuda.empty_cache()
    return {"metric": "env_steps_per_sec_policy_rollout", "value": n * world * K / (ms * 1e-3), "unit": "steps/s",
            "boards_per_gpu": n, "steps": K, "warmup": W, "ms_per_step": per,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_board": BYTES_ROLLOUT, "kernel": "policy_step_kernel<Geom<5,6>>"},
            "collectives_in_timed_region": "every 10 steps: ONE NCCL all_gather carrying the 40-double stats vector of the window "
                                           "and 32 sampled replay rows (state, action, discounted return) per GPU",
            "workload": "configs[4]: Boltzmann policy (beta 1e-3, epsilon 0.1) on synthetic Q-values resident in HBM, episode cap "
                        "50, auto-reset, fp32 obs, replay ring (4 steps deep, gamma 0.99, log10) -- one fused launch per step",
            "stats_last_window": last.get("stats"), "replay_rows_gathered": last.get("replay_rows")}


def leg_config2(ppad_b200, pdist, T, args, rank, world, dev):
    """BASELINE configs[2]: 2^24 boards per GPU, a finger path of L = 200 moves (the largest cap in the reference,
    simulation.py:35) applied in one launch and then resolved with Philox skyfall; uniform boards and an adversarial
    variant (3 colours, skyfall 70 % one colour) for the depth / combo histograms."""
    import torch
    n, L = args.config2_boards, 200
    res = {}
    gen = torch.Generator(device=dev)
    gen.manual_seed(7 + rank)
    for name, sky, colours in (("uniform", None, 6), ("adversarial", [0.7, 0.15, 0.15, 0, 0, 0, 0, 0, 0, 0], 3)):
        env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED + 2, device=dev, env0=pdist.shard_env0(rank, n), skyfall=sky,
                                   cascade_cap=32)
        boards = torch.randint(0, colours, (n, ROWS, COLS), device=dev, dtype=torch.int64, generator=gen)
        fingers = torch.stack([torch.randint(0, ROWS, (n,), device=dev, generator=gen),
                               torch.randint(0, COLS, (n,), device=dev, generator=gen)], 1)
        env.set_state(boards, fingers)
        del boards, fingers
        words = (L + 15) // 16
        packed = torch.randint(-2 ** 31, 2 ** 31, (n, words), device=dev, dtype=torch.int64, generator=gen).to(torch.int32)
        paths = (packed, L)
        snap = env.state.clone()
        out = env.step(None, eval_moves=True, paths=paths, step=1)

        def one(t):
            env.state.copy_(snap)
            env.step(None, eval_moves=True, paths=paths, step=t, out=out)
        ms_copy = T.timed(lambda t: env.state.copy_(snap), 0, 3) / 3
        ms = T.timed(one, 2, 3) / 3 - ms_copy
        st = pdist.gather_stats(pdist.local_stats(out.reward, out.info))
        res[name] = {"value": n * world / (ms * 1e-3), "unit": "path-resolves/s", "ms_per_launch": ms,
                     "moves_per_s": n * world * L / (ms * 1e-3), "depth_hist": st["depth_hist"], "combo_hist": st["combo_hist"],
                     "cascade_cap_hits": int(((out.info.to(torch.int64) >> 24) & 4).ne(0).sum().item())}
        del env, packed, snap, out
        torch.cuda.empty_cache()
    res["workload"] = "configs[2]: %d 5x6 boards/GPU, %d-move finger paths (2 bit/move) + full resolve, Philox skyfall, cascade cap 32" % (n, L)
    return res


def leg_config3(ppad_b200, pdist, T, args, rank, world, dev, peak):
    """BASELINE configs[3]: 7x6 boards -- the fused obs step at 2^20 boards and smart-data style trajectory generation
    (permuted solved boards, pass for the final reward, random walk recorded in one launch; smart_data.py:27-91)."""
    import torch
    from ppad_b200.data import solved_boards
    from ppad_b200.data.smart_data import smart_data_batched
    rows, cols, n = 6, 7, N_LOCAL
    env = ppad_b200.BatchedPAD(n, rows, cols, seed=SEED + 3, device=dev, env0=pdist.shard_env0(rank, n))
    env.reset(step=0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(55 + rank)
    slabs = [env._slab(torch.randint(0, 6, (n, SLAB_ORBS), device=dev, dtype=torch.uint8, generator=gen), SLAB_FORMAT)[0]
             for _ in range(8)]
    out = env.step(None, eval_moves=True, injected=(slabs[0], SLAB_ORBS), slab_format=SLAB_FORMAT, obs='f32', step=1)

    def one(t):
        env.step(None, eval_moves=True, injected=(slabs[t & 7], SLAB_ORBS), slab_format=SLAB_FORMAT, obs=out.obs, step=t, out=out)
    for t in range(2, 7):
        one(t)
    ms = T.timed(one, 7, 200) / 200
    bytes_per = 32 + 32 + 13 + SLAB_BYTES + rows * cols * 7 * 4
    achieved = bytes_per * n / (ms * 1e-3) / 1e9
    env.reset(step=0)
    reset_ms = T.timed(lambda t: env.reset(step=t), 1, 5) / 5
    del env, slabs, out
    torch.cuda.empty_cache()
    # smart data: m trajectories of 100 observations each
    m, steps = args.config3_trajectories, 100
    solved = solved_boards.generate_solved_boards(rows, cols, 64, seed=9)
    rng = np.random.default_rng(1 + rank)
    picks = rng.integers(0, len(solved), m)
    maps = np.stack([np.concatenate([rng.permutation(6), [6, 7]]) for _ in range(64)]).astype(np.uint8)[rng.integers(0, 64, m)]
    # warm-up at full size: the first call allocates the trajectory tensors (GBs of cudaMalloc), the timed ones reuse them
    res = smart_data_batched(solved, picks, maps, steps=steps, seed=21 + rank, device=dev)
    del res
    torch.cuda.synchronize()
    secs = []
    for rep in range(3):
        if world > 1:
            T.dist.barrier()
        t0 = time.perf_counter()
        res = smart_data_batched(solved, picks, maps, steps=steps, seed=22 + rank + rep, device=dev)
        torch.cuda.synchronize()
        secs.append(T.max_over_ranks(time.perf_counter() - t0))
        del res
    sec = float(np.median(secs))
    torch.cuda.empty_cache()
    return {"obs_step": {"value": n * world / (ms * 1e-3), "unit": "steps/s", "ms_per_step": ms,
                         "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                                      "frac": achieved / peak, "algorithmic_bytes_per_board": bytes_per,
                                      "kernel": "step_obs_kernel<Geom<6,7>>"}},
            "reset_ms_per_2p20": reset_ms,
            "smart_data": {"value": m * world / sec, "unit": "trajectories/s", "steps_per_trajectory": steps,
                           "trajectories_per_gpu": m, "seconds": sec,
                           "note": "host to host, median of 3 calls: solved-board upload, colour permutation, pass, the recorded random "
                                   "walk (one launch), trajectory tensors left in HBM"},
            "workload": "configs[3]: 7x6 boards; obs step as configs[1]; smart-data generation as smart_data.py:27-91"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boards", type=int, default=N_LOCAL, help="boards per GPU (default 2^20 = configs[1])")
    ap.add_argument("--rollout-boards", type=int, default=1 << 23, help="boards per GPU of the configs[4] sub-record")
    ap.add_argument("--config2-boards", type=int, default=1 << 24, help="boards per GPU of the configs[2] sub-record")
    ap.add_argument("--config3-trajectories", type=int, default=1 << 18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the rollout / config2 / config3 / episode sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ppad_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import ppad_b200
    from ppad_b200 import dist as pdist
    L = ppad_b200._cabi.lib()
    dev = torch.device("cuda", local_rank)
    n = args.boards
    K, W = args.steps, args.warmup
    T = Timer(torch, dist, dev, world)
    peak, peak_src = measured_peak()

    env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED, device=dev, env0=pdist.shard_env0(rank, n))
    env.reset(step=0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    host_orbs = [torch.randint(0, 6, (n, OBS_SLAB_ORBS), device=dev, dtype=torch.uint8, generator=gen) for _ in range(N_SLABS)]
    obs_slabs = [env._slab(s, OBS_SLAB_FORMAT)[0] for s in host_orbs]             # chunk-major slabs, resident in HBM
    slabs = [env._slab(s[:, :SLAB_ORBS].contiguous(), SLAB_FORMAT)[0] for s in host_orbs]   # row slabs of the same orbs
    del host_orbs
    out = env.step(None, eval_moves=True, injected=(obs_slabs[0], OBS_SLAB_ORBS), slab_format=OBS_SLAB_FORMAT, obs='f32', step=1)
    obs_buf = out.obs

    so_out = [None]

    def one_step(t, with_obs=True):
        if with_obs:
            return env.step(None, eval_moves=True, injected=(obs_slabs[t % N_SLABS], OBS_SLAB_ORBS),
                            slab_format=OBS_SLAB_FORMAT, obs=obs_buf, step=t, out=out)
        # its own result buffers: `out` keeps the last step of the headline leg for stats_last_step
        so_out[0] = env.step(None, eval_moves=True, injected=(slabs[t % N_SLABS], SLAB_ORBS), slab_format=SLAB_FORMAT,
                             obs=None, step=t, out=so_out[0])
        return so_out[0]

    for t in range(2, 2 + W):
        one_step(t)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    c0 = L.ppad_launch_count()
    w0 = time.time()
    ms = T.timed(one_step, 2 + W, K)                 # the contract's region: EXACTLY K steps
    launches = L.ppad_launch_count() - c0
    # steady state: >= 0.6 s of the same step with one event per step (a short K gives the clocks no time to show)
    k_steady = max(K, int(0.6 / max(ms / K * 1e-3, 1e-6)) + 1)
    k_steady = min(k_steady, 20000)
    per = T.per_step(one_step, 2 + W + K, k_steady)
    w1 = time.time()
    clocks = sampler.stop(w0, w1) if sampler else None
    steady = {"steps": int(k_steady), "median_ms": float(np.median(per)), "p10_ms": float(np.percentile(per, 10)),
              "p90_ms": float(np.percentile(per, 90)), "seconds": float(per.sum() * 1e-3),
              "value_at_median": n * world / (T.max_over_ranks(float(np.median(per))) * 1e-3)}
    base_t = 2 + W + K + k_steady

    # secondary: state-only step (no obs stream), L2 flushed between steps, per-step events
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    so_ms = []
    for t in range(base_t, base_t + 10):
        flush.fill_(t & 0xFF)
        so_ms.append(T.timed(lambda tt: one_step(tt, with_obs=False), t, 1))
    so_ms = float(np.median(so_ms))
    so_b2b = T.timed(lambda tt: one_step(tt, with_obs=False), base_t + 10, 50) / 50
    # the same back to back, replayed from a CUDA graph of 16 steps: a 0.065 ms kernel is shorter than one Python
    # call of env.step, so the plain loop above measures the host; the graph shows the device rate
    so_graph = None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for tt in range(3):
                one_step(base_t + 60 + tt, with_obs=False)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for tt in range(16):
                one_step(base_t + 70 + tt, with_obs=False)
        graph.replay()
        so_graph = T.timed(lambda tt: graph.replay(), 0, 20) / (20 * 16)
        del graph
    except Exception as exc:                      # a capture problem must not cost the whole bench line
        so_graph = None
        graph_error = repr(exc)
    del flush
    base_t += 100

    # end to end through the public host-buffer API (BatchedPAD.host_pipe -> ppad_pipe_step): every step the skyfall slab
    # of that step is read from pinned host memory (12 B per board) and the compact result stream (4 bits per board: action
    # + has-record; 12 B per board whose info word is not 0: reward + info) is written to pinned host memory; the call
    # returns only when the results are on the host.  The fp32 one-hot obs stays in HBM for the device-side agent.
    e2e_slabs = [env._slab(torch.randint(0, 6, (n, E2E_SLAB_ORBS), device=dev, dtype=torch.uint8, generator=gen),
                           E2E_SLAB_FORMAT)[0].cpu() for _ in range(2)]
    pipe = env.host_pipe(slab_orbs=E2E_SLAB_ORBS, compact=True, slab_format=E2E_SLAB_FORMAT)
    pipe.slab.copy_(e2e_slabs[0])
    torch.cuda.synchronize()
    # the result dictionary of the compact stream: the (reward, info) pairs one earlier step reported most often (a
    # frequency count on the host over values the device computed); records in it travel as 16-bit codes, the others
    # in full -- the decoded results are exact either way (tests/test_gpu_dropin.py)
    pipe.step(eval_moves=True, obs=obs_buf, step=base_t - 1, wait=True)
    dict_entries, dict_cover = pipe.learn_dictionary(max_entries=E2E_DICT_ENTRIES)
    e2e_dict = pipe._dict

    def e2e_step(t):
        pipe.step(eval_moves=True, obs=obs_buf, step=t, wait=True)

    # warm-up by time, not by count: the first ~50-100 ms of host traffic run slower on every box and mode (PCIe link /
    # host uncore leaving their idle states; profiles/e2e_modes.py)
    def warm_by_time(fn, t0, seconds=0.3, at_least=20):
        """every pipe has its own pinned buffers, and every fresh set of them starts slow"""
        t_warm, done = time.time(), 0
        while done < at_least or time.time() - t_warm < seconds:
            fn(t0 + done)
            done += 1
        return done

    w = warm_by_time(e2e_step, base_t)
    e2e_k = max(5, min(K, 200))
    e2e_ms = T.timed(e2e_step, base_t + w, e2e_k)
    got = pipe.results()                                  # decode once: the stream is complete and self-consistent
    assert got["lines"] > 0 and (got["info"] != 0).sum() > 0
    h2d_eager = pipe.h2d_bytes()
    h2d_lazy = pipe.h2d_bytes_on_demand(got["info"])
    pipe_slab_bytes = pipe.slab.numel() * pipe.slab.element_size()
    h2d = h2d_eager + h2d_lazy
    d2h = pipe.d2h_bytes()
    e2e_records = int((got["info"] != 0).sum())
    e2e_escapes = int(got["escapes"])
    # the decoded stream of the last timed step against the plain device step on a copy of the state before it
    if rank == 0:
        chk = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED, device=dev, env0=pdist.shard_env0(rank, n))
        chk.state.copy_(env.state)
        pipe.step(eval_moves=True, obs=obs_buf, step=base_t + w + e2e_k, wait=True)
        got2 = pipe.results()
        ref = chk.step(None, eval_moves=True, injected=(pipe.slab.to(dev), E2E_SLAB_ORBS), slab_format=E2E_SLAB_FORMAT,
                       step=base_t + w + e2e_k)
        assert np.array_equal(got2["reward"].view(np.uint64), ref.reward.cpu().numpy().view(np.uint64))
        assert np.array_equal(got2["info"], ref.info.cpu().numpy().view(np.uint32))
        assert np.array_equal(got2["action"], ref.action.cpu().numpy()) and torch.equal(env.state, chk.state)
        del chk, ref
    # the same with two steps in flight (two sets of host buffers on one pipe; the results of step t are read while
    # step t + 1 is running) -- reported separately, the headline e2e above is the synchronous call
    pipe2 = env.host_pipe(slab_orbs=E2E_SLAB_ORBS, compact=True, slab_format=E2E_SLAB_FORMAT, depth=2)
    for k, b in enumerate(pipe2.buffers):
        b.slab.copy_(e2e_slabs[k])
    pipe2.set_dictionary(*e2e_dict)

    def e2e_async_step(t):
        k = t & 1
        pipe2.wait(k)                              # results of step t - 2 (this buffer set) are on the host
        _ = int(pipe2.buffers[k].summary[0])       # "read" them
        pipe2.step(eval_moves=True, obs=obs_buf, step=t, wait=False, buf=k, join=False)

    wa = warm_by_time(e2e_async_step, base_t + 50000)
    wa += wa & 1                                   # keep the buffer parity
    e2e_async_ms = T.timed(e2e_async_step, base_t + 50000 + wa, e2e_k, finish=lambda: (pipe2.join(0), pipe2.join(1)))
    pipe2.wait(0), pipe2.wait(1)
    # for comparison: round 1's dense result arrays + state copy + nibble slab (47 B per board-step across PCIe)
    pipe_d = env.host_pipe(slab_orbs=SLAB_ORBS, chunks=2)
    pipe_d.slab.copy_(torch.from_numpy(ppad_b200.layout.pack_slab(
        np.random.default_rng(5 + rank).integers(0, 6, size=(n, SLAB_ORBS)).astype(np.uint8)).view(np.int32)))
    wd = warm_by_time(lambda t: pipe_d.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 60000)
    dense_ms = T.timed(lambda t: pipe_d.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 60000 + wd, min(e2e_k, 50)) / min(e2e_k, 50)
    dense_bytes = pipe_d.h2d_bytes() + pipe_d.d2h_bytes()
    # round 2's first compact format: 12-byte slab rows (32 orbs), 12-byte records, no dictionary
    pipe_c = env.host_pipe(slab_orbs=SLAB_ORBS, compact=True, slab_format=SLAB_FORMAT)
    pipe_c.slab.copy_(slabs[0].cpu())
    wc = warm_by_time(lambda t: pipe_c.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 70000)
    rows_ms = T.timed(lambda t: pipe_c.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 70000 + wc, min(e2e_k, 50)) / min(e2e_k, 50)
    rows_bytes = pipe_c.h2d_bytes() + pipe_c.d2h_bytes()
    # the headline's slab (chunk-major, on demand) with plain 12-byte records: what the dictionary alone is worth
    pipe_n = env.host_pipe(slab_orbs=E2E_SLAB_ORBS, compact=True, slab_format=E2E_SLAB_FORMAT)
    pipe_n.slab.copy_(e2e_slabs[0])
    wn = warm_by_time(lambda t: pipe_n.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 80000)
    nodict_ms = T.timed(lambda t: pipe_n.step(eval_moves=True, obs=obs_buf, step=t, wait=True), base_t + 80000 + wn, min(e2e_k, 50)) / min(e2e_k, 50)
    nodict_bytes = pipe_n.h2d_bytes() + pipe_n.h2d_bytes_on_demand(pipe_n.results()["info"]) + pipe_n.d2h_bytes()
    del pipe, pipe2, pipe_d, pipe_c, pipe_n

    stats = pdist.gather_stats(pdist.local_stats(out.reward, out.info))

    extras = {}
    if not args.no_extras:
        # episode mode of visualize_episode.py:28-33 / simulation.py:199 -- L = 50 random moves (swap only), then a pass
        # (full cascade) and an automatic reset, all boards in lock step: 51 launches per episode
        ep_env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED + 1, device=dev, env0=pdist.shard_env0(rank, n))
        ep_env.reset(step=0)
        ep_out = ep_env.step(None, max_moves=50, auto_reset=True)
        for t in range(51):
            ep_env.step(None, max_moves=50, auto_reset=True, out=ep_out)
        ep_ms = T.timed(lambda t: ep_env.step(None, max_moves=50, auto_reset=True, out=ep_out), 0, 102)
        ep_env.rollout(51, max_moves=50)
        ep_fused_ms = T.timed(lambda t: ep_env.rollout(51, max_moves=50), 0, 4)
        del ep_env, ep_out
        extras["episode_mode"] = (ep_ms, ep_fused_ms)
        del env, slabs, obs_slabs, out, obs_buf
        torch.cuda.empty_cache()
        extras["rollout"] = leg_rollout(ppad_b200, pdist, T, args, rank, world, dev, peak)
        extras["config2"] = leg_config2(ppad_b200, pdist, T, args, rank, world, dev)
        extras["config3"] = leg_config3(ppad_b200, pdist, T, args, rank, world, dev, peak)

    if rank == 0:
        total = n * world
        ms_per_step = ms / K
        achieved = BYTES_WITH_OBS * n / (ms_per_step * 1e-3) / 1e9          # GB/s per GPU (max-over-ranks time)
        traffic, prof = None, {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = prof["step_kernel_obs_bytes_per_launch"]
        except Exception:
            pass
        # integer-issue roofline of the state-only and rollout kernels: warp instructions per launch (ncu, static) over
        # the launch time measured here, against SMs x 4 schedulers x 1 warp instruction per clock at the sampled clock
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        issue_peak = sm_count * 4 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6

        def issue(inst, ms_launch):
            if not inst:
                return None
            rate = inst / (ms_launch * 1e-3)
            return {"warp_instructions_per_launch": inst, "achieved": rate / 1e9, "peak": issue_peak / 1e9,
                    "unit": "G warp-instr/s", "frac": rate / issue_peak,
                    "note": "instruction count from the ncu capture under profiles/ (same kernel, same workload)"}
        line = {
            "metric": "env_steps_per_sec_incl_cascades", "value": total * K / (ms * 1e-3), "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards + f64 reward + f32 obs", "data": "synthetic",
            "config": workload_config(n, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of this kernel and workload: "
                                                               "dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                         "peak_source": peak_src, "kernel": "step_obs_kernel<Geom<5,6>>",
                         "algorithmic_bytes_per_board": BYTES_WITH_OBS},
            "steady": steady,
            "state_only": {"value": total / (so_ms * 1e-3), "unit": "steps/s", "ms_per_step": so_ms,
                           "back_to_back": {"value": total / (so_b2b * 1e-3), "ms_per_step": so_b2b,
                                            "note": "one Python call per step: host bound below ~0.07 ms per step"},
                           "cuda_graph": None if not so_graph else {
                               "value": total / (so_graph * 1e-3), "ms_per_step": so_graph,
                               "issue": issue(prof.get("step_pool_warp_instructions_per_launch"), so_graph),
                               "note": "16 steps captured in a CUDA graph, replayed 20 times: the device rate of the kernel"},
                           "hbm_gbs": BYTES_STATE_ONLY * n / (so_ms * 1e-3) / 1e9,
                           "issue": issue(prof.get("step_pool_warp_instructions_per_launch"), so_ms),
                           "note": "same step without the obs stream (step_pool_kernel); L2 flushed (256 MiB fill) before every step"},
            "e2e": {"value": total * e2e_k / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_k, "steps": e2e_k,
                    "bytes_per_board_step": (h2d + d2h) / n,
                    "h2d_bytes_eager": h2d_eager, "h2d_bytes_on_demand": h2d_lazy,
                    "h2d_bytes_offered": pipe_slab_bytes,
                    "skyfall_orbs_offered_per_board": E2E_SLAB_ORBS,
                    "dictionary": {"entries": dict_entries, "share_of_records_covered_when_learned": dict_cover,
                                   "records_last_step": e2e_records, "escapes_last_step": e2e_escapes},
                    "checked": "the decoded stream of one more step equals ppad_step on device buffers bit for bit "
                               "(reward, info, action, state), in this run",
                    "note": "HostPipe.step (ppad_pipe_step), synchronous, zero copy.  In: the host offers 54 skyfall orbs "
                            "per board in pinned memory (chunk-major, six base-6 orbs per uint16); the kernel fetches the "
                            "first three chunks of every board (18 orbs, 6 B) and later chunks only where a cascade gets "
                            "that far (h2d_bytes_on_demand, counted from the consumed counts in 32-byte sectors).  Out: the compact "
                            "result stream in pinned memory: 4 bits per board (action + has-record); a board whose info "
                            "word is not 0 (the others have reward 0 and info 0 by definition) adds a 16-bit code when "
                            "its (float64 reward, uint32 info) pair is in the pipe's dictionary -- the 4096 most frequent "
                            "pairs of an earlier step, counted on the host -- and the 12 bytes themselves otherwise; the "
                            "host follows the boards from the actions, fresh boards after a reset travel with the "
                            "records (tests/test_gpu_dropin.py::test_host_pipe_compact_results_equal_device_step); the "
                            "fp32 one-hot obs stays in HBM for the device-side agent",
                    "two_steps_in_flight": {"value": total * e2e_k / (e2e_async_ms * 1e-3), "unit": "steps/s",
                                            "ms_per_step": e2e_async_ms / e2e_k},
                    "without_dictionary": {"value": total / (nodict_ms * 1e-3), "unit": "steps/s", "ms_per_step": nodict_ms,
                                           "bytes_per_board_step": nodict_bytes / n,
                                           "note": "the same slab, every record as its 12 bytes (results_compact = 1)"},
                    "row_slab_plain_records": {"value": total / (rows_ms * 1e-3), "unit": "steps/s", "ms_per_step": rows_ms,
                                               "bytes_per_board_step": rows_bytes / n,
                                               "note": "12-byte slab rows (32 orbs) + 12-byte records, no dictionary"},
                    "round1_dense_format": {"value": total / (dense_ms * 1e-3), "unit": "steps/s", "ms_per_step": dense_ms,
                                            "bytes_per_board_step": dense_bytes / n}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stats_last_step": stats,
        }
        if "episode_mode" in extras:
            ep_ms, ep_fused_ms = extras["episode_mode"]
            line["episode_mode"] = {
                "env_step_calls_per_s": total * 102 / (ep_ms * 1e-3), "episodes_per_s": total * 2 / (ep_ms * 1e-3),
                "fused_env_step_calls_per_s": total * 51 * 4 / (ep_fused_ms * 1e-3),
                "fused_episodes_per_s": total * 4 / (ep_fused_ms * 1e-3),
                "fused_issue": issue(prof.get("rollout_warp_instructions_per_51_step_launch"), ep_fused_ms / 4),
                "note": "50 random moves + pass (full cascade, Philox skyfall) + auto reset per episode, "
                        "state-only kernels, one launch per step; fused_* = the 51 steps in one ppad_rollout launch"}
        for key in ("rollout", "config2", "config3"):
            if key in extras:
                line[key] = extras[key]
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n_sample = N_LOCAL                                   # the whole workload, 3 steps: ~10-20 s of CPU time
            rate, sec = cpu_port_rate(n_sample, 3, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d boards x 3 steps of the same workload (move + cascade + fp32 obs), C port "
                                              "of the reference's Python path on %d threads, %.1f s/step" % (n_sample, threads, sec)}
            if not args.no_python_reference:
                pr = python_reference_rate(8.0)
                if pr:
                    line["cpu_baseline"]["python_reference"] = pr
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
