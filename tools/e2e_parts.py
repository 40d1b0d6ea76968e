"""Dev tool: where the end-to-end time of solve_from_wavenumbers goes (host sines vs the C call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
from superresolution_for_pdes_b200 import _lib
_lib.init(0)
n, B = 640, 512
rs = np.random.RandomState(3)
th = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True); th.copy_(torch.from_numpy(rs.uniform(0.5, 2.0, size=(B, n, n))))
u = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True)
k = rs.uniform(8, 12, size=(B, 2)); x = np.linspace(0, 1, n)
k1, k2 = k[:, 0].copy(), k[:, 1].copy()
sx = np.empty((B, n)); sy = np.empty((B, n))
def rng(lo, hi):
    np.sin((2 * np.pi * k1[lo:hi])[:, None] * x[None, :], out=sx[lo:hi]); np.sin((2 * np.pi * k2[lo:hi])[:, None] * x[None, :], out=sy[lo:hi])
for nt in (1, 2, 4, 8, 16):
    cuts = [(B * t) // nt for t in range(nt + 1)]
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=nt) as tp: list(tp.map(lambda t: rng(cuts[t], cuts[t + 1]), range(nt)))
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"sines with {nt:2d} threads: {min(ts):.2f} ms   (cpu_count {os.cpu_count()})")
for rep in range(3): _lib.solve_batch(n, B, theta=th.numpy(), sin_x=sx, sin_y=sy, out=u.numpy())
t0 = time.perf_counter()
for rep in range(4): _lib.solve_batch(n, B, theta=th.numpy(), sin_x=sx, sin_y=sy, out=u.numpy())
print(f"solve_batch (numpy views of pinned buffers, pageable sines): {(time.perf_counter() - t0) / 4 * 1e3:.1f} ms")
for rep in range(2): _lib.solve_from_wavenumbers(n, k1, k2, theta=th.numpy(), out=u.numpy())
t0 = time.perf_counter()
for rep in range(4): _lib.solve_from_wavenumbers(n, k1, k2, theta=th.numpy(), out=u.numpy())
print(f"solve_from_wavenumbers: {(time.perf_counter() - t0) / 4 * 1e3:.1f} ms")
