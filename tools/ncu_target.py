"""Small driver for ncu captures: one warm-up call and one measured call of psn_solve_batch with
device-resident inputs.  usage: python tools/ncu_target.py --n 640 --batch 32"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_for_pdes_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=640)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--calls", type=int, default=2)
ap.add_argument("--sines", action="store_true", help="pass host-style sine factors (device arrays) instead of k1,k2")
a = ap.parse_args()
dev = torch.device("cuda:0")
_lib.init(0)
g = torch.Generator(device=dev); g.manual_seed(3)
theta = torch.rand((a.batch, a.n, a.n), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.5
k = torch.rand((a.batch, 2), dtype=torch.float64, device=dev, generator=g) * 4 + 8
u = torch.empty((a.batch, a.n, a.n), dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
x = torch.linspace(0, 1, a.n, dtype=torch.float64, device=dev)
sx = torch.sin(2 * torch.pi * k[:, :1] * x[None, :]).contiguous()
sy = torch.sin(2 * torch.pi * k[:, 1:] * x[None, :]).contiguous()
for _ in range(a.calls):
    if a.sines:
        _, info = _lib.solve_batch(a.n, a.batch, theta=theta, sin_x=sx, sin_y=sy, out=u, stream=st)
    else:
        _, info = _lib.solve_batch(a.n, a.batch, theta=theta, k1=k[:, 0].contiguous(), k2=k[:, 1].contiguous(), out=u, stream=st)
torch.cuda.synchronize()
info = info.to_host()
print("iters", info.iters.min(), info.iters.max(), "status max", info.status.max(), "relres max", info.relres.max())
