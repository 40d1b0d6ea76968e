"""Dev tool: per-kernel timing of one workload (device-resident), e.g.
   python tools/quick_bench.py --n 640 --batch 512 --steps 2"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_for_pdes_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=640); ap.add_argument("--batch", type=int, default=512); ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--sines", action="store_true", help="forcing as sine factors (the product's call form) instead of device-evaluated k1, k2")
a = ap.parse_args()
dev = torch.device("cuda:0"); _lib.init(0)
for o in a.opt:
    k, v = o.split("="); _lib.set_option(k, int(v))
g = torch.Generator(device=dev); g.manual_seed(3)
theta = torch.rand((a.batch, a.n, a.n), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.5
k = torch.rand((a.batch, 2), dtype=torch.float64, device=dev, generator=g) * 4 + 8
k1, k2 = k[:, 0].contiguous(), k[:, 1].contiguous()
u = torch.empty((a.batch, a.n, a.n), dtype=torch.float64, device=dev)
xg = torch.linspace(0, 1, a.n, dtype=torch.float64, device=dev)
sx = torch.sin((2 * torch.pi * k1)[:, None] * xg[None, :]).contiguous()
sy = torch.sin((2 * torch.pi * k2)[:, None] * xg[None, :]).contiguous()
forcing = dict(sin_x=sx, sin_y=sy) if a.sines else dict(k1=k1, k2=k2)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): _lib.solve_batch(a.n, a.batch, theta=theta, out=u, stream=st, **forcing)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): _, info = _lib.solve_batch(a.n, a.batch, theta=theta, out=u, stream=st, **forcing)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
_lib.profile_begin()
_lib.solve_batch(a.n, a.batch, theta=theta, out=u, stream=st, **forcing)
prof = _lib.profile_read(); _lib.profile_end()
info = info.to_host()
print(f"lib={os.environ.get('PSN_B200_LIB','default')} n={a.n} B={a.batch} ms/step={ms:.2f} solves/s={a.batch/ms*1e3:.0f} iters={info.iters.mean():.2f} status_max={info.status.max()}")
for kname, r in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"   {kname:20s} {r['launches']:4d} launches {r['ms']:8.3f} ms  ({r['ms']/r['launches']*1e3:8.1f} us each)")
