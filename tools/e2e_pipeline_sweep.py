"""Dev tool (GPU box): end-to-end (pinned host buffers through the C ABI) throughput of one workload under several
host-pipeline settings (staging slots, upload prefetch depth, dedicated upload stream), one process.
   python tools/e2e_pipeline_sweep.py --n 640 --batch 512"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from superresolution_for_pdes_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=640); ap.add_argument("--batch", type=int, default=512); ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--sets", default="3,1,0;5,1,0;5,3,0;5,3,1;6,3,1;3,1,0")
ap.add_argument("--chunks", default="0")
ap.add_argument("--copy", default="1", help="copy_streams settings to try, e.g. 0,1")
a = ap.parse_args()
_lib.init(0)
n, B = a.n, a.batch
rs = np.random.RandomState(3)
th = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True); th.uniform_(0.5, 2.0)
k = rs.uniform(8, 12, size=(B, 2)); x = np.linspace(0, 1, n)
sx = torch.from_numpy(np.sin(2*np.pi*k[:, :1]*x)).pin_memory(); sy = torch.from_numpy(np.sin(2*np.pi*k[:, 1:]*x)).pin_memory()
u = torch.empty((B, n, n), dtype=torch.float64, pin_memory=True)
ref = None
for hc, cp in [(int(c), int(d)) for c in a.chunks.split(",") for d in a.copy.split(",")]:
    _lib.set_option("host_chunks", hc)
    _lib.set_option("copy_streams", cp)
    for st in a.sets.split(";"):
        slots, pre, hs = [int(v) for v in st.split(",")]
        _lib.set_option("host_slots", slots); _lib.set_option("host_prefetch", pre); _lib.set_option("host_h2d_stream", hs)
        for _ in range(2): _, info = _lib.solve_batch(n, B, theta=th, sin_x=sx, sin_y=sy, out=u)
        ts = []
        for _ in range(a.steps):
            t0 = time.perf_counter(); _, info = _lib.solve_batch(n, B, theta=th, sin_x=sx, sin_y=sy, out=u); ts.append(time.perf_counter() - t0)
        if ref is None: ref = u.clone()
        same = bool(torch.equal(ref, u))
        dt = sum(ts) / len(ts)
        print(f"n={n} B={B} chunks={hc} copy_streams={cp} slots={slots} prefetch={pre} h2d_stream={hs}: {dt*1e3:.1f} ms/step (min {min(ts)*1e3:.1f})  {B/dt:.0f} solves/s  "
              f"status_max={int(info.status.max())} iters={info.iters.mean():.2f} bit_identical_to_first={same}", flush=True)
