"""Measure the other BASELINE.json configurations through the public (reference-signature) API on one GPU:
C1 standard dataset (1,000 pairs 20^2+40^2), C2 sub-domain dataset, C3 multi-resolution chain (B=1) and
batched 160^2 / 320^2 references.  Host numpy in/out, wall-clock, first call excluded (CUDA init)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superresolution_for_pdes_b200 import PoissonSolver, EnhancedPoissonSolver, solve_multi_resolution, solve_multi_resolution_batch, _lib
_lib.init(0)
def timed(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    return (time.perf_counter() - t0) / reps, out
res = {}
ps = PoissonSolver(); es = EnhancedPoissonSolver()
np.random.seed(0); t, ds = timed(lambda: ps.generate_dataset(1000, k_range=(0.5, 5.0)))
res["C1 generate_dataset(1000) [20^2+40^2 pairs]"] = {"seconds": t, "samples_per_s": 1000 / t}
np.random.seed(1); t, sd = timed(lambda: es.generate_subdomain_dataset(5000, k_range=(0.5, 12.0)))
res["C2 generate_subdomain_dataset(5000) [80^2 solve + window + downsample]"] = {"seconds": t, "samples_per_s": 5000 / t}
np.random.seed(2); t, mr = timed(lambda: solve_multi_resolution(40, [80, 160, 320, 640]))
res["C3 solve_multi_resolution chain 40..640, one example"] = {"seconds": t}
np.random.seed(2); t, mb = timed(lambda: solve_multi_resolution_batch(16, 40, [80, 160, 320, 640]), reps=2)
res["C3 solve_multi_resolution_batch(16 examples)"] = {"seconds": t, "chains_per_s": 16 / t, "max_iters": int(mb["info"].iters.max())}
for n, B in ((160, 256), (320, 256)):
    rs = np.random.RandomState(n); th = rs.uniform(0.5, 2.0, size=(B, n, n)); k = rs.uniform(8, 12, size=(B, 2))
    t, _ = timed(lambda: _lib.solve_from_wavenumbers(n, k[:, 0], k[:, 1], theta=th))
    res[f"C3 batched {n}^2 references, B={B} (pageable numpy in/out)"] = {"seconds": t, "solves_per_s": B / t}
n = 640; th = np.random.RandomState(3).uniform(0.5, 2.0, size=(1, n, n)); f = np.zeros((1, n, n)); f[0] = np.sin(np.linspace(0, 30, n))[None, :] * np.sin(np.linspace(0, 20, n))[:, None]
big = PoissonSolver(320, 640)
t, _ = timed(lambda: big.solve_poisson(f[0], th[0]), reps=5)
res["single solve_poisson 640^2 (latency, host arrays)"] = {"seconds": t}
print(json.dumps(res, indent=1))
