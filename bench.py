#!/usr/bin/env python
"""bench.py -- env steps/sec including cascades (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the CPU port of the reference's path

Workload (BASELINE.json configs[1]): 2^20 independent 5x6 boards per GPU, one env step per board per
"step": a uniformly random legal non-reversing move (PlayerAction.sample, in-kernel Philox), the
orb swap, the full cascade resolution on a copy (cancel -> drop -> skyfall until stable) with the
skyfall taken from an injected 32-orb slab per board per step, the float64 reward, the packed state
written back and the fp32 one-hot observation [n,5,6,7] of the new state -- ONE fused kernel launch.
N > 1: every rank owns its own 2^20-board slab (env ids rank * 2^20 + i); no data-path collective
(weak scaling); a stats all-gather follows the timed region.

Timing: W >= 3 warm-up steps, then K steps between barrier + synchronize, CUDA events on the launch
stream, max over ranks.  Every step streams 881 MB of obs (7x the 126 MB L2) and the skyfall slabs
rotate over 16 buffers (268 MB), so no step finds its inputs in L2.
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
SLAB_ORBS = 32
N_SLABS = 16
SEED = 20181019
# algorithmic bytes per board-step (restated in DESIGN.md):
#   state in 16 + state out 16 + action 1 + reward 8 + info 4 + skyfall slab 16 = 61; + fp32 obs 840
BYTES_STATE_ONLY = 61
BYTES_WITH_OBS = 61 + ROWS * COLS * 7 * 4


def workload_config(n, world):
    return {"workload": "configs[1]: 2^20 5x6 boards/GPU, random legal move + full cascade resolve + "
                        "fp32 one-hot obs, injected 32-orb skyfall slab per board-step",
            "boards_per_gpu": n, "rows": ROWS, "cols": COLS, "skyfall_orbs_per_step": SLAB_ORBS,
            "l2": "inputs > L2: 881 MB obs stream per step, skyfall slabs rotate over 16 buffers (268 MB)",
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
    inj = rng.integers(0, 6, size=(n_sample, SLAB_ORBS)).astype(np.uint8)
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


def run_reference(args):
    """--impl reference: the reference is pure Python and does not travel to the GPU box, so this arm times
    the oracle port (oracle/ppad_oracle.c) of the same path on all host cores."""
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
                                   "of ppad/pad/game.py + utils.py + agent01.py:254-273 on %d threads; the Python "
                                   "reference itself measured ~157 steps/s/core (BASELINE.md section 2)" % (
                                       n_sample, args.steps, threads)},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_rollout(args, rank, local_rank, world):
    """BASELINE configs[4] (scaled to --boards per GPU): per step synthetic Q-values q ~ N(0, 1) * 1000 (stand-in for
    Agent01), masking + Boltzmann selection, env step with auto-reset and fp32 one-hot obs, push of (state, action,
    reward) into the per-GPU replay ring; every 10 steps an NCCL all-gather of the stats vector and of 32 replay rows
    per GPU.  The Q-value generation (torch.randn) is inside the timed region, like the agent's forward pass would be."""
    import torch
    import torch.distributed as dist
    import ppad_b200
    from ppad_b200 import dist as pdist
    from ppad_b200.replay import ReplayBuffer
    dev = torch.device("cuda", local_rank)
    n, K, W = args.boards, args.steps, args.warmup
    env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED, device=dev, env0=pdist.shard_env0(rank, n))
    env.reset(step=0)
    rb = ReplayBuffer(env, 4 * n)
    obs = env.encode_obs('f32')
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    q = torch.empty((n, 5), device=dev, dtype=torch.float32)
    out = None
    before = torch.empty_like(env.state)
    stats = None

    def one(t):
        nonlocal out, stats
        q.normal_(0.0, 1000.0, generator=gen)
        actions = env.select_actions(q, method='boltzmann', beta=1e-3, epsilon=0.1, max_moves=50)
        before.copy_(env.state)
        out = env.step(actions, auto_reset=True, obs=obs, out=out)
        rb.push(before, out.action, out.reward)
        if (t + 1) % 10 == 0:
            stats = pdist.gather_stats(pdist.local_stats(out.reward, out.info))
            pdist.gather_replay_sample(rb.state[:rb.size], rb.action[:rb.size], rb.reward[:rb.size], 32, generator=gen)

    for t in range(W):
        one(t)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(W, W + K):
        one(t)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        print(json.dumps({
            "metric": "env_steps_per_sec_policy_rollout", "value": n * world * K / (ms * 1e-3), "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bitboards + f64 reward + f32 obs", "data": "synthetic",
            "config": {"workload": "configs[4]: Boltzmann-policy rollout, synthetic Q-values, auto-reset, fp32 obs, device "
                                   "replay ring, NCCL all-gather of stats + 32 replay rows per GPU every 10 steps",
                       "boards_per_gpu": n, "collectives_per_10_steps": 4}, "stats_last": stats}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--boards", type=int, default=N_LOCAL, help="boards per GPU (default 2^20 = configs[1])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="configs1", choices=["configs1", "rollout"],
                    help="configs1 = the headline bench (default); rollout = BASELINE configs[4]: Boltzmann-policy "
                         "rollout with fp32 obs, auto-reset, device replay ring and NCCL stats / replay-sample all-gathers")
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
    if args.workload == "rollout":
        return run_rollout(args, rank, local_rank, world)
    L = ppad_b200._cabi.lib()
    dev = torch.device("cuda", local_rank)
    n = args.boards
    K, W = args.steps, args.warmup

    env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED, device=dev, env0=pdist.shard_env0(rank, n))
    env.reset(step=0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    slabs = [torch.randint(0, 6, (n, SLAB_ORBS), device=dev, dtype=torch.uint8, generator=gen) for _ in range(N_SLABS)]
    slabs = [env._slab(s)[0] for s in slabs]                       # packed nibble slabs, resident in HBM
    out = env.step(None, eval_moves=True, injected=(slabs[0], SLAB_ORBS), obs='f32', step=1)   # allocates buffers
    obs_buf = out.obs

    def one_step(t, with_obs=True):
        return env.step(None, eval_moves=True, injected=(slabs[t % N_SLABS], SLAB_ORBS),
                        obs=obs_buf if with_obs else None, step=t, out=out)

    def timed(fn, k0, count, finish=None):
        """barrier + sync, `count` steps between CUDA events on the launch stream, sync; max over ranks (ms).
        finish(): joins side streams into the launch stream before the closing event."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(k0, k0 + count):
            fn(t)
        if finish is not None:
            finish()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for t in range(2, 2 + W):
        one_step(t)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    c0 = L.ppad_launch_count()
    w0 = time.time()
    ms = timed(one_step, 2 + W, K)
    w1 = time.time()
    launches = L.ppad_launch_count() - c0
    clocks = sampler.stop(w0, w1) if sampler else None

    # secondary: state-only step (no obs stream), L2 flushed between steps, per-step events
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    so_ms = []
    for t in range(2 + W + K, 2 + W + K + 10):
        flush.fill_(t & 0xFF)
        so_ms.append(timed(lambda tt: one_step(tt, with_obs=False), t, 1))
    so_ms = float(np.median(so_ms))

    # secondary: episode mode of visualize_episode.py:28-33 / simulation.py:199 -- L = 50 random moves (swap only),
    # then a pass (full cascade) and an automatic reset, all boards in lock step: 51 launches per episode
    ep_env = ppad_b200.BatchedPAD(n, ROWS, COLS, seed=SEED + 1, device=dev, env0=pdist.shard_env0(rank, n))
    ep_env.reset(step=0)
    ep_out = ep_env.step(None, max_moves=50, auto_reset=True)
    for t in range(51):
        ep_env.step(None, max_moves=50, auto_reset=True, out=ep_out)
    ep_ms = timed(lambda t: ep_env.step(None, max_moves=50, auto_reset=True, out=ep_out), 0, 102)
    # the same 51 steps fused into one launch (ppad_rollout: boards stay in registers between steps)
    ep_env.rollout(51, max_moves=50)
    ep_fused_ms = timed(lambda t: ep_env.rollout(51, max_moves=50), 0, 4)
    del ep_env

    # end to end through the public host-buffer API (BatchedPAD.host_pipe -> ppad_pipe_step): every step copies that
    # step's skyfall slab from pinned host memory and reads reward, info, action and the packed state back to pinned
    # host memory; the call returns only when the results are on the host.  The fp32 one-hot obs stays in HBM.
    pipe = env.host_pipe(slab_orbs=SLAB_ORBS, chunks=2)     # best of 1..8 on this box (profiles/pipe_probe.py)
    pipe.slab.copy_(slabs[0].cpu())
    torch.cuda.synchronize()

    def e2e_step(t):
        pipe.step(eval_moves=True, obs=obs_buf, step=t, wait=True)

    base = 2 + W + K + 10 + 100000          # step indices of the e2e legs: clear of everything before
    # warm-up by time, not by count: the first ~50-100 ms of host traffic run at 0.9-1.1 ms per step on every box and
    # mode (PCIe link / host uncore leaving their idle states; profiles/e2e_modes.py), then it settles at ~0.72 ms
    t_warm, w = time.time(), 0
    while w < 20 or time.time() - t_warm < 0.3:
        e2e_step(base + w)
        w += 1
    e2e_k = max(5, min(K, 50))
    e2e_ms = timed(e2e_step, base + w, e2e_k)
    h2d = pipe.slab.numel() * 4
    d2h = pipe.reward.numel() * 8 + pipe.info.numel() * 4 + pipe.action.numel() + pipe.state.numel() * 4
    # the same with two steps in flight (two sets of host buffers on one pipe; the results of step t are read while
    # step t + 1 is running) -- reported separately, the headline e2e above is the synchronous call
    pipe2 = env.host_pipe(slab_orbs=SLAB_ORBS, chunks=2, depth=2)
    for k, b in enumerate(pipe2.buffers):
        b.slab.copy_(slabs[k].cpu())

    def e2e_async_step(t):
        k = t & 1
        pipe2.wait(k)                              # results of step t - 2 (this buffer set) are on the host
        _ = pipe2.buffers[k].reward[0].item()      # "read" them
        pipe2.step(eval_moves=True, obs=obs_buf, step=t, wait=False, buf=k)

    for t in range(base + 50000, base + 50010):
        e2e_async_step(t)
    e2e_async_ms = timed(e2e_async_step, base + 50010, e2e_k, finish=lambda: (pipe2.join(0), pipe2.join(1)))
    pipe2.wait(0), pipe2.wait(1)

    # episode stats of the last step, all-gathered over NVLink (tiny; outside the timed region)
    stats = pdist.gather_stats(pdist.local_stats(out.reward, out.info))

    if rank == 0:
        peak, peak_src = measured_peak()
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
                         "traffic": traffic, "peak_source": peak_src, "kernel": "step_obs_kernel<Geom<5,6>>",
                         "algorithmic_bytes_per_board": BYTES_WITH_OBS},
            "state_only": {"value": total / (so_ms * 1e-3), "unit": "steps/s", "ms_per_step": so_ms,
                           "hbm_gbs": BYTES_STATE_ONLY * n / (so_ms * 1e-3) / 1e9,
                           "issue": issue(prof.get("step_pool_warp_instructions_per_launch"), so_ms),
                           "note": "same step without the obs stream (step_pool_kernel); L2 flushed (256 MiB fill) before every step"},
            "e2e": {"value": total * e2e_k / (e2e_ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_k,
                    "note": "HostPipe.step (ppad_pipe_step, zero copy: the kernel reads / writes the pinned host buffers), synchronous: "
                            "skyfall slab H2D; reward, info, action and packed state (the reference's (board, finger) "
                            "observation) D2H every step; the fp32 one-hot obs stays in HBM for the device-side agent",
                    "two_steps_in_flight": {"value": total * e2e_k / (e2e_async_ms * 1e-3), "unit": "steps/s",
                                            "ms_per_step": e2e_async_ms / e2e_k}},
            "episode_mode": {"env_step_calls_per_s": total * 102 / (ep_ms * 1e-3), "episodes_per_s": total * 2 / (ep_ms * 1e-3),
                             "fused_env_step_calls_per_s": total * 51 * 4 / (ep_fused_ms * 1e-3),
                             "fused_episodes_per_s": total * 4 / (ep_fused_ms * 1e-3),
                             "fused_issue": issue(prof.get("rollout_warp_instructions_per_51_step_launch"), ep_fused_ms / 4),
                             "note": "50 random moves + pass (full cascade, Philox skyfall) + auto reset per episode, "
                                     "state-only kernels, one launch per step; fused_* = the 51 steps in one ppad_rollout launch; "
                                     "the reference does ~3.8e3 env.step calls/s per core in this mode"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stats_last_step": stats,
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n_sample = N_LOCAL                                   # the whole workload, 3 steps: ~10-20 s of CPU time
            rate, sec = cpu_port_rate(n_sample, 3, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "steps/s", "cores": threads, "kind": "port",
                                    "sample": "%d boards x 3 steps of the same workload (move + cascade + fp32 obs), C port "
                                              "of the reference's Python path on %d threads, %.1f s/step" % (n_sample, threads, sec)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
