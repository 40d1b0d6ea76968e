"""Dev tool (GPU box): BASELINE.json configs[4] on one GPU -- device-resident solves/s over grid size 20..640 and
batch 1..65,536 (theta~U(0.5,2) iid, forcing as sine factors, k~U(8,12) capped at the grid's Nyquist range by n/4).
A (n, B) pair is skipped when one [B, n, n] fp64 array exceeds --cap-gb (the workspace is ~10 such arrays).
   python tools/size_batch_sweep.py --out gpurun_out/sweep.json"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from superresolution_for_pdes_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--cap-gb", type=float, default=3.5)
ap.add_argument("--sizes", default="20,40,80,160,320,640")
ap.add_argument("--batches", default="1,8,64,512,4096,32768,65536")
a = ap.parse_args()
dev = torch.device("cuda:0")
_lib.init(0)
st = torch.cuda.current_stream().cuda_stream
rows = []
for n in [int(s) for s in a.sizes.split(",")]:
    for B in [int(s) for s in a.batches.split(",")]:
        if B * n * n * 8 > a.cap_gb * 1e9:
            rows.append({"n": n, "batch": B, "skipped": "array larger than cap"})
            continue
        g = torch.Generator(device=dev)
        g.manual_seed(4)
        theta = torch.rand((B, n, n), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.5
        khi = min(12.0, n / 4.0)
        k = torch.rand((B, 2), dtype=torch.float64, device=dev, generator=g) * (khi - 0.5) + 0.5
        xg = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
        sx = torch.sin((2 * torch.pi * k[:, :1]) * xg[None, :]).contiguous()
        sy = torch.sin((2 * torch.pi * k[:, 1:]) * xg[None, :]).contiguous()
        u = torch.empty((B, n, n), dtype=torch.float64, device=dev)
        for _ in range(3):
            _, info = _lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy, out=u, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy, out=u, stream=st)
        e1.record()
        torch.cuda.synchronize()
        steps = max(3, min(200, int(60.0 / max(e0.elapsed_time(e1), 1e-3))))
        e0.record()
        for _ in range(steps):
            _, info = _lib.solve_batch(n, B, theta=theta, sin_x=sx, sin_y=sy, out=u, stream=st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        info = info.to_host()
        row = {"n": n, "batch": B, "ms_per_call": round(ms, 4), "solves_per_s": round(B / ms * 1e3, 1), "steps": steps,
               "mean_iters": float(info.iters.mean()), "max_relres": float(info.relres.max()), "status_max": int(info.status.max())}
        rows.append(row)
        print(row, flush=True)
        del theta, u, sx, sy
        torch.cuda.empty_cache()
if a.out:
    json.dump({"what": "device-resident solves/s, one B200, CUDA events on the launch stream, 3 warm-ups", "rows": rows},
              open(a.out, "w"), indent=1)
