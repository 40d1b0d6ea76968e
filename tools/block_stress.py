"""Dev tool (GPU box): the register-blocked kernels (80x80 by default; `block_stress.py SEED 40` / `20` for the others) on awkward data -- forcing scales from 1e-30 to 1e30, theta contrasts
up to 1e6, high wave numbers, a zero right-hand side -- against the oracle (SciPy spsolve restatement of the reference).
The V-cycle of that kernel runs in float on a power-of-two-scaled residual; this checks that the scaling makes it
insensitive to the magnitude of the data.  Exits 1 on any failure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superresolution_for_pdes_b200 import _lib
from oracle import poisson_oracle as O

_lib.init(0)
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 80
x = np.linspace(0, 1, n)
bad = 0
iters_seen = []
for scale in [1e-30, 1e-12, 1.0, 1e9, 1e30]:
    for lo, hi in [(0.5, 2.0), (1e-3, 1e3), (1.0, 1.0)]:
        for B in [1, 7, 300]:
            theta = np.exp(rs.uniform(np.log(lo), np.log(hi), size=(B, n, n)))
            k = rs.uniform(0.5, 0.375 * n, size=(B, 2))
            sx, sy = np.sin(2 * np.pi * k[:, :1] * x), np.sin(2 * np.pi * k[:, 1:] * x)
            if B == 7:
                sx[3] = 0.0   # zero right-hand side: u = 0, status converged, no NaN from the 1 / ||r|| scaling
            u, info = _lib.solve_batch(n, B, theta=theta, sin_x=sx * scale, sin_y=sy)
            errs = []
            for b in sorted(set([0, B // 2 + 1 if B == 7 else B // 2, B - 1])):   # (sample 3 of B = 7 has the zero rhs)
                ref = O.solve_poisson(np.outer(sy[b], sx[b]), theta[b], n)
                errs.append(O.rel_l2(u[b] / scale, ref))
            ok = (np.asarray(info.status) == 0).all() and max(errs) <= 1e-10 and np.isfinite(u).all()
            if B == 7:
                ok = ok and not u[3].any() and int(info.iters[3]) == 0
            bad += 0 if ok else 1
            iters_seen.append(int(np.max(info.iters)))
            print(f"{'ok  ' if ok else 'FAIL'} scale={scale:g} theta=[{lo:g},{hi:g}] B={B:3d} iters={int(np.min(info.iters))}..{int(np.max(info.iters))} "
                  f"status_max={int(np.max(info.status))} relres_max={float(np.max(info.relres)):.2e} err_max={max(errs):.2e}")
print(f"{bad} failures; iterations {min(iters_seen)}..{max(iters_seen)}")
sys.exit(1 if bad else 0)
