"""The kernels either side of the env step and the per-stage entry points, 2^20 5x6 boards:
ms per launch and achieved GB/s over their algorithmic bytes (read + written per board)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppad_b200
from ppad_b200.replay import ReplayBuffer
if os.environ.get('PPAD_AB_LIB'):
    ppad_b200._cabi.LIB_PATH = os.path.abspath(os.environ['PPAD_AB_LIB'])
n = 1 << 20
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1)
env.reset(step=0)
dev = env.device


def t(label, fn, bytes_per_board, reps=50):
    for _ in range(3):
        fn()
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    print("%-44s %8.4f ms  %7.0f GB/s  (%d B/board)" % (label, best, bytes_per_board * n / best / 1e6, bytes_per_board))


q = torch.randn((n, 5), device=dev) * 1000
t("select_actions max (q 20 + state 16 + act 1)", lambda: env.select_actions(q, method='max', step=3), 37)
t("select_actions boltzmann", lambda: env.select_actions(q, method='boltzmann', beta=1e-3, step=3), 37)
t("sample_actions (state 16 + act 1)", lambda: env.sample_actions(step=5), 17)
t("legal_mask (state 16 + mask 1)", lambda: env.legal_mask(), 17)
T = 51
rew = torch.rand((n, T), dtype=torch.float64, device=dev) * 1e4
lens = torch.randint(1, T + 1, (n,), dtype=torch.int32, device=dev)
import ctypes as C
from ppad_b200 import _cabi
def disc(r, log10):
    _cabi.check(env.lib.ppad_discount(env._ctx, n, T, C.c_void_p(r.data_ptr()), C.c_void_p(lens.data_ptr()), 0.9, int(log10), 0, env._stream()))
t("discount [n, 51] dense rewards, log10", lambda: disc(rew, True), 2 * 8 * T, reps=10)
sparse = torch.zeros((n, T), dtype=torch.float64, device=dev)
sparse[torch.arange(n, device=dev), (lens - 1).long()] = 5000.0          # one reward per episode, at the pass
t("discount [n, 51] episode rewards, log10", lambda: disc(sparse.clone() if False else sparse, True), 2 * 8 * T, reps=10)
t("discount [n, 51] dense rewards, linear", lambda: disc(rew, False), 2 * 8 * T, reps=10)
rb = ReplayBuffer(env, 4 * n)
act = torch.zeros(n, dtype=torch.uint8, device=dev)
r1 = torch.zeros(n, dtype=torch.float64, device=dev)
t("replay push (2 x 25)", lambda: rb.push(env.state, act, r1), 50)
t("replay sample 2^20 rows (2 x 25)", lambda: rb.sample(n), 50)
maps = torch.stack([torch.randperm(6, device=dev) for _ in range(64)]).to(torch.uint8)
maps = torch.cat([maps, torch.tensor([[6, 7]] * 64, dtype=torch.uint8, device=dev)], 1).repeat(n // 64, 1).contiguous()
t("permute_colours (2 x 16 + 8)", lambda: env.permute_colours(maps), 40)
t("cancel in place (2 x 16 + 1)", lambda: env.cancel(log_cap=0) if False else env.cancel(log_cap=1), 33 + 2)
t("drop in place (2 x 16)", lambda: env.drop(), 32)
t("fill_board Philox (2 x 16 + 2)", lambda: env.fill_board(step=9), 34)
st = env.get_state()
t("unpack -> int64 boards (16 + 30*8 + 2*8 + 3)", lambda: env.get_state(), 16 + 240 + 16 + 3)
t("pack <- int64 boards", lambda: env.set_state(st.boards, st.fingers), 16 + 240 + 16 + 1)
t("encode_obs u8 (16 + 210)", lambda: env.encode_obs('u8'), 226)

# ---- the callers built from those kernels
import time
from ppad_b200.lookahead import two_step_lookahead
from ppad_b200.data.smart_data import smart_data_batched
from ppad_b200.data import solved_boards
m = 1 << 16
small = ppad_b200.BatchedPAD(m, 5, 6, seed=2, skyfall_damage=False)
small.reset(step=0)
for _ in range(2):
    two_step_lookahead(small, q_fn=None)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    res = two_step_lookahead(small, q_fn=None)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print("two_step_lookahead, %d envs x 21 candidates, no roam: %.2f ms  (%.2fe6 envs/s, %.2fe6 candidate evaluations/s)" % (
    m, dt * 1e3, m / dt / 1e6, 21 * m / dt / 1e6))
qf = lambda obs: torch.randn((obs.shape[0], 5), device=obs.device) * 1000
torch.cuda.synchronize(); t0 = time.perf_counter()
res = two_step_lookahead(small, q_fn=qf, policy='boltzmann', beta=1e-3, max_lookahead_len=10)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("two_step_lookahead, %d envs, Boltzmann roam up to 10 moves (synthetic q): %.2f ms  (%.2fe6 envs/s)" % (m, dt * 1e3, m / dt / 1e6))
from ppad_b200.lookahead import model_simulation_2sla
W = torch.randint(-9, 10, (210, 5), device=small.device).float()
qlin = lambda obs: obs.reshape(obs.shape[0], -1) @ W
model_simulation_2sla(small, qlin, policy='boltzmann', beta=0.05, max_episode_len=50, max_lookahead_len=10)
torch.cuda.synchronize(); t0 = time.perf_counter()
res = model_simulation_2sla(small, qlin, policy='boltzmann', beta=0.05, max_episode_len=50, max_lookahead_len=10)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("model_simulation_2sla, %d envs, episode <= 50 + look-ahead <= 10 (linear q on the obs): %.1f ms  (%.2fe6 episodes/s, %.1fe6 SAR tuples/s)" % (
    m, dt * 1e3, m / dt / 1e6, float(res["n_actions"].sum()) / dt / 1e6))
rngp = torch.Generator().manual_seed(0)
picks = torch.randint(0, 16, (m,), generator=rngp).numpy()
maps = torch.stack([torch.randperm(6, generator=rngp) for _ in range(m)]).to(torch.uint8).numpy()
torch.cuda.synchronize(); t0 = time.perf_counter()
out = smart_data_batched(solved_boards.boards, picks, maps, steps=100, seed=3)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("smart_data_batched, %d trajectories x 100 steps: %.1f ms  (%.2fe6 trajectories/s, %.1fe6 walk steps/s)" % (
    m, dt * 1e3, m / dt / 1e6, 99 * m / dt / 1e6))
