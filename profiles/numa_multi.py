"""8 ranks copying D2H + H2D at once (torchrun): does binding each rank to its GPU's NUMA node before it pins
host memory change the aggregate PCIe rate?   torchrun --nproc-per-node 8 profiles/numa_multi.py [bind]"""
import os, sys, time
import torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
def rd(f):
    try: return open("/sys/bus/pci/devices/%s/%s" % (bus, f)).read().strip()
    except Exception as e: return "n/a"
node, cl = rd("numa_node"), rd("local_cpulist")
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
aff0 = len(os.sched_getaffinity(0))
if bind and cl not in ("", "n/a"):
    cpus = set()
    for part in cl.split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
n = 32 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
h1.fill_(1); h2.fill_(2)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=30):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
a = run(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = run(both)
print("rank %d gpu %s numa %s cpulist %s affinity %d -> %d" % (rank, bus, node, cl, aff0, len(os.sched_getaffinity(0))), flush=True)
dist.barrier()
if rank == 0:
    print("%s: aggregate D2H alone %.1f GB/s; D2H + H2D together %.1f GB/s each way (%d ranks)" % (
        "bound to GPU-local cpus" if bind else "default affinity", world * n / a / 1e9, world * n / c / 1e9, world), flush=True)
    os.system("lscpu | grep -i -E 'numa|socket' | head -6")
dist.destroy_process_group()
