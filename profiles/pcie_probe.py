"""PCIe copy rates on the bench box: H2D, D2H, both at once (pinned memory, 16-64 MiB)."""
import torch, time
dev = torch.device("cuda")
for mb in (4, 16, 64):
    n = mb << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    t_h2d = run(lambda: d1.copy_(h1, non_blocking=True))
    t_d2h = run(lambda: h2.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    t_both = run(both)
    print("%3d MiB: H2D %.1f GB/s  D2H %.1f GB/s  both at once: %.1f GB/s each (%.3f ms vs %.3f + %.3f)" % (
        mb, n / t_h2d / 1e9, n / t_d2h / 1e9, n / t_both / 1e9, t_both * 1e3, t_h2d * 1e3, t_d2h * 1e3))
