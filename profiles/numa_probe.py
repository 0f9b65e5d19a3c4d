"""Does CPU/NUMA placement change pinned-copy rates on this box?"""
import os, time, torch
p = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
base = "/sys/bus/pci/devices/" + bus
def rd(f):
    try: return open(os.path.join(base, f)).read().strip()
    except Exception as e: return "n/a (%s)" % e
print("gpu", bus, "numa_node", rd("numa_node"), "local_cpulist", rd("local_cpulist"), "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("nvidia-smi topo -m | head -8; lscpu | grep -i -E 'numa|socket|model name' | head -8")
def probe(tag):
    n = 32 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h1.fill_(1); h2.fill_(2)
    d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    a = run(lambda: d1.copy_(h1, non_blocking=True)); b = run(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    c = run(both)
    print("%-18s H2D %.1f  D2H %.1f  both %.1f GB/s each" % (tag, n / a / 1e9, n / b / 1e9, n / c / 1e9))
probe("default affinity")
cl = rd("local_cpulist")
if not cl.startswith("n/a") and cl:
    cpus = set()
    for part in cl.split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    cpus &= os.sched_getaffinity(0)
    if cpus:
        os.sched_setaffinity(0, cpus)
        probe("gpu-local cpus")
