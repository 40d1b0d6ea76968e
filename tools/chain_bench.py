"""Dev tool (GPU box): end-to-end cost of the multi-resolution chain (BASELINE configs[2]) through
the public API, host RNG vs device RNG, and the raw throughput of the device MT19937 replay.
   python tools/chain_bench.py --examples 64 --out gpurun_out/chain.json"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from superresolution_for_pdes_b200 import _lib, solve_multi_resolution_batch, solve_multi_resolution

ap = argparse.ArgumentParser()
ap.add_argument("--examples", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="")
a = ap.parse_args()
_lib.init(0)
res = {"examples": a.examples, "chain": [40, 80, 160, 320, 640]}


def timed(fn, reps):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


for name, kw in (("host_rng", dict(device_rng=False)), ("device_rng", dict(device_rng=True)),
                 ("device_rng_theta_stays_on_device", dict(device_rng=True, theta_on_host=False))):
    np.random.seed(1)
    solve_multi_resolution_batch(2, **kw)   # warm-up (plans, polynomials, arenas)
    np.random.seed(1)
    solve_multi_resolution_batch(a.examples, **kw)
    def run():
        np.random.seed(1)
        solve_multi_resolution_batch(a.examples, **kw)
    t = timed(run, a.reps)
    res[name] = {"s_per_call": t, "examples_per_s": a.examples / t}
    print(name, res[name], flush=True)

for name, kw in (("single_host_rng", dict(device_rng=False)), ("single_device_rng", dict(device_rng=True))):
    np.random.seed(2); solve_multi_resolution(**kw)
    def run1():
        solve_multi_resolution(**kw)
    t = timed(run1, 5)
    res[name] = {"s_per_chain": t}
    print(name, res[name], flush=True)

# raw generator: 512 blocks of (2 + 640^2) doubles into device memory
B, cnt = 512, 640 * 640
dev = torch.empty((B, cnt), dtype=torch.float64, device="cuda")
rs = np.random.RandomState(3)
_lib.mt19937_uniform_blocks(B, 2, cnt, 0.5, 2.0, out=dev, random_state=rs)
_lib.profile_begin()
t = timed(lambda: _lib.mt19937_uniform_blocks(B, 2, cnt, 0.5, 2.0, out=dev, random_state=rs), 3)
prof = _lib.profile_read(); _lib.profile_end()
t = timed(lambda: _lib.mt19937_uniform_blocks(B, 2, cnt, 0.5, 2.0, out=dev, random_state=rs), 5)
t0 = time.perf_counter(); np.random.RandomState(3).uniform(0.5, 2.0, (8, 640, 640)); th = (time.perf_counter() - t0) / 8
res["raw"] = {"blocks": B, "doubles_per_block": cnt, "s_per_call": t, "GB_per_s_written": B * cnt * 8 / t / 1e9,
              "kernels_ms": prof, "numpy_host_s_per_block": th, "numpy_host_s_for_all": th * B}
print(res["raw"], flush=True)
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
