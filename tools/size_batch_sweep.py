"""Dev tool (GPU box): BASELINE.json configs[4] -- device-resident solves/s over grid size 20..640 and total batch
1..65,536 (theta~U(0.5,2) iid, forcing as sine factors, k~U(0.5, min(12, n/4))), on one GPU or, under torchrun, sharded
by sample over all ranks (contiguous ranges, no collective on the solve path; time = max over ranks).  A (n, B) pair is
skipped when one rank's [B/ranks, n, n] fp64 array exceeds --cap-gb.  With --cpu the reference's own CPU solve (oracle
port: SciPy spsolve / SuperLU, one process per host core) is timed next to it on a bounded sample per size.
   python tools/size_batch_sweep.py --out gpurun_out/sweep.json
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
       tools/size_batch_sweep.py --cpu --out gpurun_out/sweep8.json"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from superresolution_for_pdes_b200 import _lib
from superresolution_for_pdes_b200.sharding import shard_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
ap.add_argument("--cap-gb", type=float, default=3.5)
ap.add_argument("--sizes", default="20,40,80,160,320,640")
ap.add_argument("--batches", default="1,8,64,512,4096,32768,65536")
ap.add_argument("--cpu", action="store_true", help="also time the CPU reference (rank 0, bounded sample per size)")
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
_lib.init(local)
st = torch.cuda.current_stream().cuda_stream


def sync_all():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_ranks(v):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rows = []
for n in [int(s) for s in a.sizes.split(",")]:
    for B in [int(s) for s in a.batches.split(",")]:
        lo, hi = shard_bounds(B, world, rank)
        Bl = hi - lo
        per_rank_max = -(-B // world)
        if per_rank_max * n * n * 8 > a.cap_gb * 1e9:
            rows.append({"n": n, "batch": B, "skipped": "array larger than cap"})
            continue
        g = torch.Generator(device=dev)
        g.manual_seed(4 + rank)
        Ba = max(Bl, 1)
        theta = torch.rand((Ba, n, n), dtype=torch.float64, device=dev, generator=g) * 1.5 + 0.5
        khi = min(12.0, n / 4.0)
        k = torch.rand((Ba, 2), dtype=torch.float64, device=dev, generator=g) * (khi - 0.5) + 0.5
        xg = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
        sx = torch.sin((2 * torch.pi * k[:, :1]) * xg[None, :]).contiguous()
        sy = torch.sin((2 * torch.pi * k[:, 1:]) * xg[None, :]).contiguous()
        u = torch.empty((Ba, n, n), dtype=torch.float64, device=dev)

        def call():
            if Bl > 0:
                return _lib.solve_batch(n, Bl, theta=theta[:Bl], sin_x=sx[:Bl], sin_y=sy[:Bl], out=u[:Bl], stream=st)[1]
            return None

        for _ in range(3):
            info = call()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        sync_all()
        steps = max(3, min(200, int(60.0 / max(max_ranks(e0.elapsed_time(e1)), 1e-3))))
        sync_all()
        e0.record()
        for _ in range(steps):
            info = call()
        e1.record()
        sync_all()
        ms = max_ranks(e0.elapsed_time(e1)) / steps
        it_mean, rr_max, st_max = 0.0, 0.0, 0
        if info is not None:
            ih = info.to_host()
            it_mean, rr_max, st_max = float(ih.iters.mean()), float(ih.relres.max()), int(ih.status.max())
        st_max = int(max_ranks(float(st_max)))
        row = {"n": n, "batch": B, "n_gpus": world, "ms_per_call": round(ms, 4), "solves_per_s": round(B / ms * 1e3, 1), "steps": steps,
               "mean_iters_rank0": it_mean, "max_relres_rank0": rr_max, "status_max": st_max}
        rows.append(row)
        if rank == 0:
            print(row, flush=True)
        del theta, u, sx, sy
        torch.cuda.empty_cache()

cpu = None
if a.cpu and rank == 0:
    # the reference's CPU path on this box's host cores: oracle port, one process per core, one thread each
    import multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B_

    cpu = {}
    for n in [int(s) for s in a.sizes.split(",")]:
        P = B_.host_workers(n)
        per = {20: 400, 40: 100, 80: 24, 160: 6, 320: 2, 640: 1}.get(n, 1)
        ctx = mp.get_context("fork")
        with ctx.Pool(P, initializer=B_._worker_init) as pool:
            pool.map(B_._worker_solve, [(n, w) for w in range(P)])          # warm-up (imports, first factorisation)
            t0 = time.perf_counter()
            pool.map(B_._worker_solve, [(n, 1000 + w) for w in range(P * per)], chunksize=per)
            dt = time.perf_counter() - t0
        cpu[str(n)] = {"solves_per_s": P * per / dt, "cores": P, "sample": f"{P * per} solves of {n}x{n}, {per} per worker process"}
        print("cpu", n, cpu[str(n)], flush=True)

if a.out and rank == 0:
    json.dump({"what": f"device-resident solves/s, {world} B200 (samples sharded contiguously over ranks, max over ranks), CUDA events on the "
                       "launch stream, 3 warm-ups", "rows": rows, "cpu_reference": cpu}, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
