"""Dev tool: end-to-end (pinned host buffers through the C ABI) throughput for one workload."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from superresolution_for_pdes_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=640); ap.add_argument("--batch", type=int, default=512); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
_lib.init(0)
for o in a.opt:
    k, v = o.split("="); _lib.set_option(k, int(v))
n, B = a.n, a.batch
rs = np.random.RandomState(3)
th = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True); th.copy_(torch.from_numpy(rs.uniform(0.5, 2.0, size=(B, n, n))))
k = rs.uniform(8, 12, size=(B, 2)); x = np.linspace(0, 1, n)
sx = torch.from_numpy(np.sin(2*np.pi*k[:, :1]*x)).pin_memory(); sy = torch.from_numpy(np.sin(2*np.pi*k[:, 1:]*x)).pin_memory()
u = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True)
for _ in range(2): _, info = _lib.solve_batch(n, B, theta=th, sin_x=sx, sin_y=sy, out=u)
t0 = time.perf_counter()
for _ in range(a.steps): _, info = _lib.solve_batch(n, B, theta=th, sin_x=sx, sin_y=sy, out=u)
dt = (time.perf_counter() - t0) / a.steps
print(f"e2e n={n} B={B} opts={a.opt}: {dt*1e3:.1f} ms/step  {B/dt:.0f} solves/s  (H2D {th.numel()*8/1e9:.2f} GB, D2H {u.numel()*8/1e9:.2f} GB; status max {info.status.max()})")

_lib.profile_begin()
t0 = time.perf_counter(); _lib.solve_batch(n, B, theta=th, sin_x=sx, sin_y=sy, out=u); dtp = time.perf_counter() - t0
prof = _lib.profile_read(); _lib.profile_end()
print(f"   profiled call {dtp*1e3:.1f} ms; sum of kernel times {sum(r['ms'] for r in prof.values()):.1f} ms over {sum(r['launches'] for r in prof.values())} launches")
for kname, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:6]:
    print(f"      {kname:20s} {r['launches']:4d} launches {r['ms']:8.3f} ms")
