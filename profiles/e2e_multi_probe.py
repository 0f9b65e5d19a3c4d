"""e2e leg only, under torchrun on N GPUs of one box: HostPipe.step (compact dictionary-coded stream) with the skyfall slab
as chunks6 (chunk-major over the batch) and chunks6t (chunk-major per tile of 256 boards), alternating, max over ranks.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/e2e_multi_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import ppad_b200

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
obs = torch.empty((n, 5, 6, 7), dtype=torch.float32, device=dev)
env = ppad_b200.BatchedPAD(n, 5, 6, seed=1, device=dev, env0=rank * n)
env.reset(step=0)
pipes = {}
for fmt, eager in (("chunks6", 3), ("chunks6t", 3), ("chunks10", 2)):
    p = env.host_pipe(slab_orbs=54 if fmt != "chunks10" else 50, compact=True, slab_format=fmt, slab_eager=eager)
    p.slab.copy_(env._slab(torch.randint(0, 6, (n, p.slab_orbs), device=dev, dtype=torch.uint8, generator=gen), fmt)[0].cpu())
    for w in range(30):
        p.step(eval_moves=True, obs=obs, step=1 + w, wait=True)
    p.learn_dictionary(max_entries=4096)
    t0 = time.time()
    while time.time() - t0 < 0.3:
        p.step(eval_moves=True, obs=obs, step=50, wait=True)
    pipes[fmt] = p


def timed(p, reps=200):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(reps):
        p.step(eval_moves=True, obs=obs, step=1000 + t, wait=True)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


for rep in range(3):
    for fmt, p in pipes.items():
        ms = timed(p)
        if rank == 0:
            print("%d GPUs %-9s %.4f ms/step  %.3fe9 steps/s" % (world, fmt, ms, world * n / ms / 1e6), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
