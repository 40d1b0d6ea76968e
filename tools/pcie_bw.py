import torch, time
dev = torch.device("cuda:0")
n = 1 << 28  # 2 GiB of float64
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device=dev)
d2 = torch.empty(n, dtype=torch.float64, device=dev); h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 8 / 1e9
print("H2D GB/s", gb / t(lambda: d.copy_(h, non_blocking=True)))
print("D2H GB/s", gb / t(lambda: h2.copy_(d2, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("bidirectional: each direction GB/s", gb / t(both))
import os; print("cpus", os.cpu_count())
