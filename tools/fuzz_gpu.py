"""Dev tool (GPU box): random grid sizes / batch sizes / option combinations through the C ABI against the oracle
(SciPy spsolve restatement of the reference).  Prints one line per case and a summary; exits 1 on any failure."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superresolution_for_pdes_b200 import _lib
from oracle import poisson_oracle as O

_lib.init(0)
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sizes = [20, 20, 40, 40, 64, 80, 80, 160, 160, 320, 320, 640, 88, 90, 96, 100, 104, 112, 120, 128, 131, 144, 160, 168, 176, 192, 200, 208, 224, 240, 256, 272, 288, 300, 320, 352, 384, 416, 448]
bad = 0
t0 = time.time()
for case in range(ncases):
    n = int(rs.choice(sizes))
    B = int(rs.choice([1, 2, 3, 5, 17, 40]))
    opts = {"vcycle_f32": int(rs.rand() < 0.8), "upd_strip": int(rs.rand() < 0.8), "p_f32": int(rs.rand() < 0.8),
            "x_defer": int(rs.rand() < 0.8), "nu": int(rs.choice([1, 2, 2, 2, 3])), "pdl": int(rs.rand() < 0.7),
            "vcycle_mode": int(rs.choice([2, 2, 2, 1, 0])), "fuse_upd": int(rs.choice([2, 2, 1, 0])),
            "copy_streams": int(rs.rand() < 0.7), "coarsest_max": int(rs.choice([0, 0, 6, 10])),
            "coarse_block": int(rs.rand() < 0.75), "res_block": int(rs.rand() < 0.85)}
    host = rs.rand() < 0.5
    theta = rs.uniform(0.5, 2.0, size=(B, n, n))
    k = rs.uniform(0.5, 12.0, size=(B, 2))
    x = np.linspace(0, 1, n)
    sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
    scale = float(rs.choice([1.0, 1.0, 1e-20, 1e12]))
    for kname, v in opts.items():
        _lib.set_option(kname, v)
    try:
        if host:
            u, info = _lib.solve_batch(n, B, theta=theta, sin_x=sx * scale, sin_y=sy)
        else:
            import torch
            d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            ud = torch.empty((B, n, n), dtype=torch.float64, device="cuda")
            _, info = _lib.solve_batch(n, B, theta=d(theta), sin_x=d(sx * scale), sin_y=d(sy), out=ud)
            torch.cuda.synchronize()
            u = ud.cpu().numpy()
            info = info.to_host()
        errs = []
        for b in sorted(set([0, B - 1])):
            ref = O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)
            errs.append(O.rel_l2(u[b] / scale, ref))
        ok = (np.asarray(info.status) <= 1).all() and max(errs) <= 1e-10 and np.isfinite(u).all()
    except Exception as e:   # noqa: BLE001
        ok, errs, info = False, [repr(e)], None
    bad += 0 if ok else 1
    print(f"{'ok  ' if ok else 'FAIL'} n={n:4d} B={B:3d} host={int(host)} scale={scale:g} opts={opts} iters={None if info is None else int(np.max(info.iters))} err={errs}")
for kname, v in {"vcycle_f32": 1, "upd_strip": 1, "p_f32": 1, "x_defer": 1, "nu": 2, "pdl": 1, "vcycle_mode": 2, "fuse_upd": 2,
                 "copy_streams": 1, "coarsest_max": 0, "coarse_block": 1, "res_block": 1}.items():
    _lib.set_option(kname, v)
print(f"{ncases - bad}/{ncases} cases ok in {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
